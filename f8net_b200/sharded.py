"""Batch sharding across the GPUs of one box.

Images are independent, so the batch dimension is partitioned contiguously, one process per
GPU, weights replicated; the only exchange step is ONE all-gather of the float32 logits.
Each rank's classifier kernel writes straight into its slice of the gather buffer, so the
collective needs no staging copy.  (The reference's int_op_only path is single-process CPU:
``distributed: False`` in res18_fix_quant_test_int_op_only.yml:41.)
"""
from typing import Callable, Tuple


def shard_range(n: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) slice of a batch of n for ``rank``; the first n % world ranks get
    one extra image, so ragged batches are covered without padding."""
    if world <= 0 or not (0 <= rank < world) or n < 0:
        raise ValueError(f"bad shard request n={n} world={world} rank={rank}")
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


class ShardedRunner:
    """Runs ``forward_local(x_shard, out=slice)`` on this rank's shard and all-gathers the
    logits.  ``forward_local`` is Engine.run_device on GPUs; the gloo CPU tests pass a
    stand-in so the host logic can be exercised without a device."""

    def __init__(self, forward_local: Callable, num_classes: int, group=None):
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.forward_local = forward_local
        self.num_classes = num_classes
        self._buf = None

    def gather_buffer(self, n_per_rank, like):
        import torch
        need = self.world * n_per_rank * self.num_classes
        if self._buf is None or self._buf.numel() < need or self._buf.device != like.device:
            self._buf = torch.empty(need, dtype=torch.float32, device=like.device)
        return self._buf[:need].view(self.world, n_per_rank, self.num_classes)

    def __call__(self, x_shard):
        """x_shard: this rank's images (every rank must pass the same count).  Returns the
        logits of the whole batch, [world * n_local, classes], rank-major = batch order."""
        n = x_shard.shape[0]
        full = self.gather_buffer(n, x_shard)
        mine = full[self.rank]
        out = self.forward_local(x_shard, out=mine)
        if out.data_ptr() != mine.data_ptr():
            mine.copy_(out)
        if self.world > 1:
            # in-place all-gather: the input is this rank's slice of the output
            self.dist.all_gather_into_tensor(full.view(-1), mine.reshape(-1), group=self.group)
        return full.view(self.world * n, self.num_classes)


class OverlappedGather:
    """Steady-state form of the same exchange for a stream of batches: the all-gather of step i
    overlaps the forward pass of step i+1.  Two gather buffers alternate; ``slot(i)`` is this
    rank's slice of buffer i % 2 (where step i's classifier writes), ``submit(i)`` starts the
    asynchronous all-gather of that buffer behind the work already enqueued on the current
    stream, and ``ready(i)`` makes the current stream (CUDA) or the host (gloo) wait until the
    gather that last used buffer i % 2 -- step i-2's -- has finished, which is what the forward
    pass of step i needs before it overwrites the slice.  ``result(i)`` waits for step i's own
    gather and returns the [world * n, classes] logits."""

    def __init__(self, n_per_rank: int, num_classes: int, like, group=None):
        import torch
        import torch.distributed as dist
        self.dist, self.group = dist, group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.n, self.num_classes = n_per_rank, num_classes
        self.bufs = [torch.empty((self.world, n_per_rank, num_classes), dtype=torch.float32,
                                 device=like.device) for _ in range(2)]
        self.work = [None, None]

    def slot(self, i):
        return self.bufs[i % 2][self.rank]

    def ready(self, i):
        w = self.work[i % 2]
        if w is not None:
            w.wait()
            self.work[i % 2] = None

    def submit(self, i):
        full = self.bufs[i % 2]
        self.work[i % 2] = self.dist.all_gather_into_tensor(
            full.view(-1), full[self.rank].reshape(-1), group=self.group, async_op=True)

    def result(self, i):
        self.ready(i)
        return self.bufs[i % 2].view(self.world * self.n, self.num_classes)

    def drain(self):
        for j in range(2):
            self.ready(j)

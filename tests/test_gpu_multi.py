"""N > 1 on real GPUs: two NCCL ranks, one per GPU, each running its shard of a
seeded batch through the engine; the all-gathered logits (rank-major order, the classifier
writing straight into the rank's slice of the gather buffer) must equal the single-GPU logits of
the whole batch bit for bit, for the plain gather and for the overlapped double-buffered one
bench.py times.  Skipped when fewer than two GPUs are visible (bench.py performs the same check
at every N before it times anything: ``parity.gathered_logits_exact``)."""
import os
import socket
import sys

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, arch, n_local, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        import f8net_b200
        from f8net_b200 import synth
        from f8net_b200.sharded import OverlappedGather, ShardedRunner
        hs = synth.HEAD_SIGNED[arch]
        sd = synth.make_state_dict(arch, hs)
        eng = f8net_b200.compile(sd, arch=arch, head_signed=hs, device=dev, chunk=n_local)
        x = torch.from_numpy(synth.make_input(arch, world * n_local, hs, seed=31)).to(dev)
        whole = eng.run_device(x)
        mine = x[rank * n_local:(rank + 1) * n_local].contiguous()
        gathered = ShardedRunner(eng.run_device, 1000)(mine)
        ok_plain = torch.equal(gathered, whole)
        # overlapped form: three steps over rotated shards, each step's gather checked
        og = OverlappedGather(n_local, 1000, x)
        ok_over = True
        perms = [torch.roll(torch.arange(world * n_local), s).to(dev) for s in (0, 1, 2)]
        for i, p in enumerate(perms):
            og.ready(i)
            eng.run_device(x[p][rank * n_local:(rank + 1) * n_local].contiguous(), out=og.slot(i))
            og.submit(i)
            if i:
                ok_over &= torch.equal(og.result(i - 1), whole[perms[i - 1]])
        ok_over &= torch.equal(og.result(len(perms) - 1), whole[perms[-1]])
        og.drain()
        torch.cuda.synchronize()
        q.put((rank, bool(ok_plain), bool(ok_over), whole[:2].cpu().numpy() if rank == 0 else None))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("arch", ["resnet18", "mobilenet_v2"])
def test_two_gpu_all_gather_equals_single_gpu(cuda, f8lib, arch):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from f8net_b200 import synth
    from oracle import nets
    world, n_local = 2, 5
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, arch, n_local, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get() for _ in range(world)]
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    for rank, ok_plain, ok_over, first in res:
        assert ok_plain, f"rank {rank}: gathered logits differ from the single-GPU logits"
        assert ok_over, f"rank {rank}: overlapped gather returned another step's logits"
        if first is not None:
            hs = synth.HEAD_SIGNED[arch]
            want = nets.forward(arch, synth.make_state_dict(arch, hs),
                                synth.make_input(arch, world * n_local, hs, seed=31)[:2], hs)
            assert np.array_equal(first, want)

"""N > 1 host logic on CPU: world_size-2 gloo processes shard a batch, run a stand-in for the
local engine (the oracle -- test infrastructure, allowed here as the checker), all-gather
the logits and compare with the unsharded result."""
import os
import socket
import sys

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_range_partitions_every_batch():
    from f8net_b200.sharded import shard_range
    for n in (0, 1, 7, 8, 255, 256, 2048):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(8, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_total, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from f8net_b200 import synth
        from f8net_b200.sharded import ShardedRunner, shard_range
        from oracle import nets
        arch = "mobilenet_v2"
        sd = synth.make_state_dict(arch)
        x = synth.make_input(arch, n_total)
        lo, hi = shard_range(n_total, world, rank)

        def forward_local(xs, out=None):
            y = torch.from_numpy(nets.forward(arch, sd, xs.numpy()))
            out.copy_(y)
            return out

        runner = ShardedRunner(forward_local, 1000)
        full = runner(torch.from_numpy(x[lo:hi]))
        if rank == 0:
            q.put(full.numpy().copy())
    finally:
        dist.destroy_process_group()


def test_two_rank_shard_and_all_gather_matches_single_process():
    from f8net_b200 import synth
    from oracle import nets
    n_total, world = 4, 2
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_total, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    arch = "mobilenet_v2"
    want = nets.forward(arch, synth.make_state_dict(arch), synth.make_input(arch, n_total))
    assert got.shape == (n_total, 1000)
    assert np.array_equal(got, want)


def _overlap_worker(rank, world, port, steps, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from f8net_b200.sharded import OverlappedGather
        n, classes = 3, 5
        og = OverlappedGather(n, classes, torch.empty(1))
        seen = []
        for i in range(steps):
            og.ready(i)                      # buffer i % 2 is free again (step i-2 gathered)
            og.slot(i).copy_(torch.full((n, classes), float(100 * i + rank)))
            og.submit(i)
            if i >= 1:
                seen.append(og.result(i - 1).clone())     # step i-1's gather, overlapped with step i
        seen.append(og.result(steps - 1).clone())
        og.drain()
        if rank == 1:
            q.put(torch.stack(seen).numpy())
    finally:
        dist.destroy_process_group()


def test_overlapped_gather_double_buffering_keeps_step_order():
    """OverlappedGather (bench.py's N>1 steady state): the gather of step i runs while step i+1
    writes the other buffer; every step's gathered logits are rank-major and belong to that step."""
    world, steps = 2, 5
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_overlap_worker, args=(r, world, port, steps, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert got.shape == (steps, world * 3, 5)
    for i in range(steps):
        for r in range(world):
            assert (got[i, 3 * r:3 * r + 3] == 100 * i + r).all()

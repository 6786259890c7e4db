#!/usr/bin/env python
"""Per-launch device times of one batch through the engine (f8_plan_profile: CUDA events
around every launch) with the algorithmic bytes / int8 ops of each launch beside them.

    python tools/profile_ops.py --arch resnet18 --batch 256 --chunk 32 --backend 1
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

import f8net_b200  # noqa: E402
from f8net_b200 import synth  # noqa: E402
from f8net_b200.roofline import op_work  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--arch", default="resnet18")
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--chunk", type=int, default=32)
    ap.add_argument("--backend", type=int, default=-1)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--no-fuse-tail", action="store_true")
    a = ap.parse_args()
    hs = synth.HEAD_SIGNED.get(a.arch, False)
    sd = synth.make_state_dict(a.arch, hs)
    kw = {} if a.backend < 0 else {"backend": a.backend}
    if a.no_fuse_tail:
        kw["fuse_tail"] = False
    eng = f8net_b200.compile(sd, arch=a.arch, head_signed=hs, chunk=a.chunk, **kw)
    S = eng.net.image_size
    x = torch.randint(0, 128, (a.batch, S, S, 4), dtype=torch.int8 if hs else torch.uint8, device="cuda")
    x[..., 3] = 0
    eng.run_device(x)
    torch.cuda.synchronize()
    work = op_work(eng.plan)
    acc = [0.0] * len(work)
    for _ in range(a.reps):
        for i, (_, _, ms) in enumerate(eng.profile(x)):
            acc[i] += ms / a.reps
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.reps):
        eng.run_device(x)
    e1.record()
    torch.cuda.synchronize()
    print(f"# {a.arch} batch {a.batch} chunk {a.chunk} backend {eng.backend}: "
          f"{e0.elapsed_time(e1) / a.reps:.3f} ms/batch back-to-back, {sum(acc):.3f} ms sum of launches")
    print(f"{'op':34s} {'geom':>26s} {'ms':>8s} {'TOPS':>8s} {'GB/s(alg)':>10s}")
    for w, op, ms in zip(work, eng.plan.ops, acc):
        geom = f"{op.cin}->{op.cout} k{op.k}s{op.stride} {op.hin}->{op.hout}"
        tops = w["ops"] * a.batch / (ms / 1e3) / 1e12 if ms > 0 else 0
        gbs = (w["bytes_per_image"] * a.batch + w["weight_bytes"]) / (ms / 1e3) / 1e9 if ms > 0 else 0
        print(f"{op.name:34s} {geom:>26s} {ms:8.4f} {tops:8.1f} {gbs:10.1f}")


if __name__ == "__main__":
    main()

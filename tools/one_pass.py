#!/usr/bin/env python
"""Runs whole forward passes of one network (engine-native NHWC input, batch 256) -- the process ncu wraps.

    python tools/one_pass.py --arch resnet18 --count        # launches per pass
    python tools/one_pass.py --arch resnet18 --passes 2 --names names.json
The first pass warms every kernel (function attributes, tensor maps); `ncu -s <launches> -c <launches>`
captures the second.  --names writes the kernel template of every launch in launch order.
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

import f8net_b200  # noqa: E402
from f8net_b200 import _capi as C  # noqa: E402
from f8net_b200 import synth  # noqa: E402
from f8net_b200.roofline import op_work  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--arch", default="resnet18")
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--passes", type=int, default=2)
    ap.add_argument("--count", action="store_true")
    ap.add_argument("--names", default="")
    a = ap.parse_args()
    hs = synth.HEAD_SIGNED.get(a.arch, False)
    eng = f8net_b200.compile(synth.make_state_dict(a.arch, hs), arch=a.arch, head_signed=hs, chunk=a.batch)
    if a.count:
        print(eng.launches(a.batch, C.F8_IN_NHWC4_8))
        return
    S = eng.net.image_size
    g = torch.Generator(device="cuda").manual_seed(7)
    x = torch.randint(-127 if hs else 0, 128 if hs else 256, (a.batch, S, S, 4), dtype=torch.int8 if hs else torch.uint8,
                      device="cuda", generator=g)
    x[..., 3] = 0
    for _ in range(a.passes):
        eng.run_device(x)
    torch.cuda.synchronize()
    if a.names:
        eng.profile(x)
        work = op_work(eng.plan)
        rows = [dict(op=op.name, kernel=k, algorithmic_bytes=w["bytes_per_image"] * a.batch + w["weight_bytes"],
                     int8_ops=w["ops"] * a.batch)
                for op, k, w in zip(eng.plan.ops, eng.kernel_names(), work) if k]
        json.dump({"arch": a.arch, "batch": a.batch, "launches": rows}, open(a.names, "w"), indent=1)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Times ONE dense layer through f8_conv_dense (back-to-back launches, CUDA events), for kernel work:
    python tools/bench_layer.py --cin 64 --cout 384 --hw 14 --k 1 [--n 256] [--carry] [--reps 50]"""
import argparse
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from f8net_b200 import _capi as C  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cin", type=int, default=64)
    ap.add_argument("--cout", type=int, default=384)
    ap.add_argument("--hw", type=int, default=14)
    ap.add_argument("--k", type=int, default=1)
    ap.add_argument("--stride", type=int, default=1)
    ap.add_argument("--n", type=int, default=256)
    ap.add_argument("--reps", type=int, default=50)
    ap.add_argument("--carry", action="store_true")
    a = ap.parse_args()
    lib = C.lib()
    pad = a.k // 2
    cpad = lambda c: (c + 15) // 16 * 16
    cin_p, cout_p = cpad(a.cin), cpad(a.cout)
    ho = (a.hw + 2 * pad - a.k) // a.stride + 1
    rng = np.random.default_rng(1)
    w = rng.integers(-127, 128, (a.cout, a.cin, a.k, a.k)).astype(np.int32)
    nb = lib.f8_pack_weights_bytes(C.F8_OP_CONV_DENSE, a.cin, a.cout, cin_p, cout_p, a.k, a.k)
    wp = np.zeros(nb, np.uint8)
    C.check(lib.f8_pack_weights(C.F8_OP_CONV_DENSE, w.ctypes.data, a.cin, a.cout, cin_p, cout_p, a.k, a.k, wp.ctypes.data))
    dev = "cuda"
    x = torch.randint(0, 256, (a.n, a.hw, a.hw, cin_p), dtype=torch.uint8, device=dev)
    wd = torch.from_numpy(wp).to(dev)
    bd = torch.zeros(cout_p, dtype=torch.int32, device=dev)
    y = torch.empty((a.n, ho, ho, cout_p), dtype=torch.uint8, device=dev)
    args = C.f8_conv_args()
    args.n, args.cin, args.cout, args.cin_pad, args.cout_pad = a.n, a.cin, a.cout, cin_p, cout_p
    args.kh = args.kw = a.k
    args.stride, args.pad = a.stride, pad
    args.hin = args.win = a.hw
    args.hout = args.wout = ho
    args.in_, args.wpack, args.bias = x.data_ptr(), wd.data_ptr(), bd.data_ptr()
    args.relu = 1
    args.out[0] = y.data_ptr()
    args.out_shift[0] = 12
    keep = []
    if a.carry:
        pix = (a.n * ho * ho + 127) // 128 * 128
        ci = torch.zeros(pix * cout_p, dtype=torch.int32, device=dev)
        co = torch.zeros(pix * cout_p, dtype=torch.int32, device=dev)
        args.carry_in, args.carry_out = ci.data_ptr(), co.data_ptr()
        keep = [ci, co]
    st = torch.cuda.current_stream()
    for _ in range(3):
        C.check(lib.f8_conv_dense(ctypes.byref(args), 1, st.cuda_stream))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.reps):
        C.check(lib.f8_conv_dense(ctypes.byref(args), 1, st.cuda_stream))
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.reps
    ops = 2.0 * a.n * ho * ho * a.cin * a.cout * a.k * a.k
    byts = a.n * (a.hw * a.hw * a.cin + ho * ho * a.cout * (9 if a.carry else 1))
    print(f"{a.cin}->{a.cout} k{a.k}s{a.stride} {a.hw}x{a.hw} n={a.n}{' carry' if a.carry else ''}: {ms * 1e3:.1f} us  "
          f"{ops / ms / 1e9:.0f} TOPS  {byts / ms / 1e6:.0f} GB/s")


if __name__ == "__main__":
    main()

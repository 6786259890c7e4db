"""Measured int8 tensor-pipe ceiling of this library's own GEMM kernel: a large
point-wise convolution (1x1, Cin = Cout = 1024, 131072 pixels -- 275 GOP on 268 MB of
activations, far above the int8 ridge) through ``f8_conv_dense`` on the tcgen05 backend, timed
with CUDA events.  bench.py reports it as ``roofline.int8_peak_tops`` and divides the achieved
TOPS of each kernel family by it."""
import ctypes

import numpy as np

from . import _capi as C


def measure_int8_peak(device, reps=10, cin=1024, cout=1024, n=128, hw=32):
    import torch
    lib = C.lib()
    dev = torch.device(device)
    rng = np.random.default_rng(7)
    w = rng.integers(-127, 128, size=(cout, cin, 1, 1), dtype=np.int64).astype(np.int32)
    nbytes = lib.f8_pack_weights_bytes(C.F8_OP_CONV_DENSE, cin, cout, cin, cout, 1, 1)
    wp = np.zeros(nbytes, dtype=np.uint8)
    C.check(lib.f8_pack_weights(C.F8_OP_CONV_DENSE, w.ctypes.data, cin, cout, cin, cout, 1, 1,
                                wp.ctypes.data))
    with torch.cuda.device(dev):
        g = torch.Generator(device=dev).manual_seed(3)
        x = torch.randint(0, 256, (n * hw * hw, cin), dtype=torch.uint8, device=dev, generator=g)
        wd = torch.from_numpy(wp).to(dev)
        bd = torch.zeros(cout, dtype=torch.int32, device=dev)
        y = torch.empty((n * hw * hw, cout), dtype=torch.uint8, device=dev)
        a = C.f8_conv_args()
        a.n, a.cin, a.cout, a.cin_pad, a.cout_pad = n, cin, cout, cin, cout
        a.kh = a.kw = a.stride = 1
        a.pad = 0
        a.hin = a.win = a.hout = a.wout = hw
        a.in_, a.wpack, a.bias = x.data_ptr(), wd.data_ptr(), bd.data_ptr()
        a.relu = 1
        a.out[0] = y.data_ptr()
        a.out_shift[0], a.out_signed[0] = 14, 0
        st = torch.cuda.current_stream(dev)
        for _ in range(2):
            C.check(lib.f8_conv_dense(ctypes.byref(a), 1, st.cuda_stream))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(reps):
            C.check(lib.f8_conv_dense(ctypes.byref(a), 1, st.cuda_stream))
        e1.record(st)
        e1.synchronize()
        ms = e0.elapsed_time(e1) / reps
    ops = 2.0 * n * hw * hw * cin * cout
    return {"tops": ops / (ms / 1e3) / 1e12, "ms": ms,
            "shape": f"1x1 conv {cin}->{cout} on {n * hw * hw} pixels ({ops / 1e9:.0f} GOP), f8_conv_dense tcgen05"}

"""Calls the per-kernel C-ABI entry points with numpy inputs (device memory via torch) and
computes the expected result with the CPU oracle.  Used by the -m gpu parity tests."""
import ctypes

import numpy as np
import torch

from f8net_b200 import _capi as C
from oracle import oracle as O

from util import (carry_elems, carry_to_nchw, cpad, nchw_to_carry, nchw_to_nhwc8,  # noqa: F401
                  nchw_to_nhwc32, nhwc_to_nchw)

DEV = "cuda:0"


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def pack(lib, kind, w, cin_pad, cout_pad):
    cout, cin_g, kh, kw = w.shape
    cin = cout if kind == C.F8_OP_CONV_DW else cin_g
    n = lib.f8_pack_weights_bytes(kind, cin, cout, cin_pad, cout_pad, kh, kw)
    dst = np.zeros(n, dtype=np.uint8)
    w = np.ascontiguousarray(w, dtype=np.int32)
    C.check(lib.f8_pack_weights(kind, w.ctypes.data, cin, cout, cin_pad, cout_pad, kh, kw,
                                dst.ctypes.data))
    return dst


def oracle_epilogue(acc, carry, carry_shift, relu, outs):
    """The reference's chain after a conv accumulator: residual add, ReLU, requants."""
    v = acc
    if carry is not None:
        v, _ = O.residual_add(v, carry, max(carry_shift, 0), max(-carry_shift, 0))
    if relu:
        v = O.relu(v)
    q = [O.requant(v, max(0, -s), max(0, s), bool(g)) for s, g in outs]
    return v, q


def run_conv(lib, x, w, b, stride, pad, *, depthwise=False, in_signed=False, relu=False,
             carry=None, carry_shift=0, outs=((0, False),), want_carry=True, want_f32=False,
             backend=0, in_pad=None):
    """x int32 NCHW in the 8-bit range; w reference layout; returns (v_int32 NCHW or None,
    [8-bit images as int32 NCHW], float logits or None) from the GPU."""
    n, cin, hin, win = x.shape
    cout, _, kh, kw = w.shape
    cin_pad = in_pad or (4 if cin == 3 else cpad(cin))
    cout_pad = cpad(cout)
    hout = (hin + 2 * pad - kh) // stride + 1
    wout = (win + 2 * pad - kw) // stride + 1
    kind = C.F8_OP_CONV_DW if depthwise else C.F8_OP_CONV_DENSE
    xd = dev(nchw_to_nhwc8(x, cin_pad, in_signed))
    wd = dev(pack(lib, kind, w, cin_pad, cout_pad))
    bp = np.zeros(cout_pad, np.int32)
    bp[:cout] = b
    bd = dev(bp)
    a = C.f8_conv_args()
    a.n, a.cin, a.cout, a.cin_pad, a.cout_pad = n, cin, cout, cin_pad, cout_pad
    a.kh, a.kw, a.stride, a.pad = kh, kw, stride, pad
    a.hin, a.win, a.hout, a.wout = hin, win, hout, wout
    a.in_signed = int(in_signed)
    a.in_, a.wpack, a.bias = xd.data_ptr(), wd.data_ptr(), bd.data_ptr()
    keep = [xd, wd, bd]
    if not depthwise and kh == 3 and kw == 3 and pad == 1 and cin_pad % 64 == 0:
        # the stage-major copy the plan hands to the resident-patch kernel (f8_pack_weights_stage3x3)
        dense = pack(lib, kind, w, cin_pad, cout_pad)
        st = np.zeros(lib.f8_pack_weights_stage3x3_bytes(cin_pad, cout_pad), dtype=np.uint8)
        C.check(lib.f8_pack_weights_stage3x3(dense.ctypes.data, cin_pad, cout_pad, st.ctypes.data))
        sd = dev(st)
        keep.append(sd)
        a.wpack_stage = sd.data_ptr()
    if carry is not None:
        cd = dev(nchw_to_carry(carry, cout_pad))
        keep.append(cd)
        a.carry_in = cd.data_ptr()
    a.carry_shift, a.relu = carry_shift, int(relu)
    M = n * hout * wout
    co = None
    if want_carry:
        co = torch.full((carry_elems(n, hout, wout, cout_pad),), -12345, dtype=torch.int32, device=DEV)
        a.carry_out = co.data_ptr()
    q = []
    for j, (s, g) in enumerate(outs):
        t = torch.full((M, cout_pad), 0x77, dtype=torch.uint8, device=DEV)
        q.append(t)
        a.out[j] = t.data_ptr()
        a.out_shift[j], a.out_signed[j] = s, int(g)
    f = None
    if want_f32:
        f = torch.full((M, cout), float("nan"), dtype=torch.float32, device=DEV)
        a.out_f32, a.out_f32_ld = f.data_ptr(), cout
    st = torch.cuda.current_stream().cuda_stream
    if depthwise:
        C.check(lib.f8_conv_dw3x3(ctypes.byref(a), st))
    else:
        C.check(lib.f8_conv_dense(ctypes.byref(a), backend, st))
    torch.cuda.synchronize()
    del keep
    shape = (n, hout, wout, cout_pad)
    v = carry_to_nchw(co.cpu().numpy(), n, cout, hout, wout, cout_pad) if co is not None else None
    qi = []
    for t, (s, g) in zip(q, outs):
        arr = t.cpu().numpy().reshape(shape)
        if g:
            arr = arr.view(np.int8)
        # padded channels must be exact zeros only when bias/weights there are zero and the
        # requant of 0 is 0 -- which holds for every (shift, signedness)
        assert not arr[..., cout:].any(), "padded output channels are not zero"
        qi.append(nhwc_to_nchw(arr, cout))
    fo = f.cpu().numpy().reshape(n, hout, wout, cout).transpose(0, 3, 1, 2) if f is not None else None
    return v, qi, fo


def expect_conv(x, w, b, stride, pad, *, depthwise=False, relu=False, carry=None, carry_shift=0,
                outs=((0, False),)):
    groups = x.shape[1] if depthwise else 1
    acc = O.conv2d(x, w, b, stride, pad, groups)
    return oracle_epilogue(acc, carry, carry_shift, relu, outs)

"""ctypes front-end of the CPU oracle (oracle/f8_oracle.c).

TEST INFRASTRUCTURE ONLY -- imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference leg.  The product package f8net_b200 never imports it.

Every function takes / returns contiguous numpy int32 arrays in the reference's NCHW
layout.  See f8_oracle.c for the reference file:line each primitive restates.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libf8oracle.so")
_lib = None

_i32p = ctypes.POINTER(ctypes.c_int32)
_f32p = ctypes.POINTER(ctypes.c_float)


def build(force=False):
    """Compile oracle/libf8oracle.so with the committed Makefile."""
    src = os.path.join(_HERE, "f8_oracle.c")
    if (force or not os.path.exists(_LIB_PATH)
            or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src)):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B", "libf8oracle.so"])
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB_PATH)
        _lib.f8o_max_threads.restype = ctypes.c_int
    return _lib


def _p(a):
    return a.ctypes.data_as(_i32p)


def _c(a):
    a = np.ascontiguousarray(a, dtype=np.int32)
    return a


def set_threads(n):
    lib().f8o_set_threads(int(n))


def max_threads():
    return int(lib().f8o_max_threads())


def requant(x, fl, input_fl, signed):
    """int_op_only_fix_quant(x, 8, fl, input_fl, signed) -- fix_quant_ops.py:90-114."""
    assert fl >= 0 and (fl <= 7 if signed else fl <= 8)  # fix_quant_ops.py:91-96
    x = _c(x)
    y = np.empty_like(x)
    rc = lib().f8o_requant(_p(x), _p(y), ctypes.c_size_t(x.size), int(fl), int(input_fl),
                           int(bool(signed)))
    if rc:
        raise ValueError("requant: shift out of range")
    return y


def conv2d(x, w, b, stride, pad, groups=1):
    """int32 nn.Conv2d (fix_quant_ops.py:680-714)."""
    x, w = _c(x), _c(w)
    b = _c(b)
    N, C, H, W = x.shape
    O, Cg, kh, kw = w.shape
    assert Cg * groups == C
    Ho = (H + 2 * pad - kh) // stride + 1
    Wo = (W + 2 * pad - kw) // stride + 1
    y = np.empty((N, O, Ho, Wo), dtype=np.int32)
    rc = lib().f8o_conv2d(_p(x), N, C, H, W, _p(w), _p(b), O, kh, kw, int(stride), int(pad),
                          int(groups), _p(y))
    if rc:
        raise ValueError("conv2d: bad groups")
    return y


def relu(x):
    x = _c(x)
    y = np.empty_like(x)
    lib().f8o_relu(_p(x), _p(y), ctypes.c_size_t(x.size))
    return y


def maxpool_float_rt(x, k=3, stride=2, pad=1):
    """head[-1](x.float()).int() -- fix_resnet.py:358-359."""
    x = _c(x)
    N, C, H, W = x.shape
    Ho = (H + 2 * pad - k) // stride + 1
    Wo = (W + 2 * pad - k) // stride + 1
    y = np.empty((N, C, Ho, Wo), dtype=np.int32)
    lib().f8o_maxpool_float_rt(_p(x), N, C, H, W, k, stride, pad, _p(y))
    return y


def maxpool_int(x, k=3, stride=2, pad=1):
    """FXQMaxPool2d.forward -- fix_quant_ops.py:141-157 (quant_maxpool True only)."""
    x = _c(x)
    N, C, H, W = x.shape
    Ho = (H + 2 * pad - k) // stride + 1
    Wo = (W + 2 * pad - k) // stride + 1
    y = np.empty((N, C, Ho, Wo), dtype=np.int32)
    lib().f8o_maxpool_int(_p(x), N, C, H, W, k, stride, pad, _p(y))
    return y


def avgpool_sum(x):
    """FXQAvgPool2d.forward int branch -- fix_quant_ops.py:126-134. Returns [N,C]."""
    x = _c(x)
    N, C, H, W = x.shape
    y = np.empty((N, C), dtype=np.int32)
    rc = lib().f8o_avgpool_sum(_p(x), N, C, H, W, _p(y))
    if rc:
        raise AssertionError("FXQAvgPool2d: res <= 2**32-1 violated (fix_quant_ops.py:132)")
    return y


def linear(q, w, b):
    """int nn.Linear then .float() -- fix_quant_ops.py:1165-1195, fix_resnet.py:383.
    Returns (int32 logits, float32 logits)."""
    q, w, b = _c(q), _c(w), _c(b)
    N, K = q.shape
    O = w.shape[0]
    yi = np.empty((N, O), dtype=np.int32)
    yf = np.empty((N, O), dtype=np.float32)
    lib().f8o_linear(_p(q), N, K, _p(w), _p(b), O, _p(yi), yf.ctypes.data_as(_f32p))
    return yi, yf


def residual_add(res, x, res_fl, x_fl):
    """fix_resnet.py:40-76 / fix_mobilenet_v2.py:34-48. Returns (tensor, fraclen)."""
    res, x = _c(res), _c(x)
    assert res.shape == x.shape
    out = np.empty_like(res)
    fl = lib().f8o_residual_add(_p(res), _p(x), _p(out), ctypes.c_size_t(res.size), int(res_fl),
                                int(x_fl))
    if fl < 0:
        raise ValueError("residual_add: shift out of range")
    return out, fl


def input_u8(x):
    """fix_train.py:689-692: (255*x).round_().int(), fraclen 8."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    y = np.empty(x.shape, dtype=np.int32)
    lib().f8o_input_u8(x.ctypes.data_as(_f32p), _p(y), ctypes.c_size_t(x.size))
    return y


def input_s8(x, fl):
    """fix_train.py:682-687 with fix_quant signed (fix_quant_ops.py:64-87)."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    y = np.empty(x.shape, dtype=np.int32)
    lib().f8o_input_s8(x.ctypes.data_as(_f32p), _p(y), ctypes.c_size_t(x.size), int(fl))
    return y


def image_prep_u8(pix, normalize, fl=8, mean=(0.485, 0.456, 0.406), std=(0.229, 0.224, 0.225)):
    """Decoded uint8 pixels [..., H, W, 3] -> the int32 NCHW tensor forward_loss hands to the head:
    torchvision ToTensor (p / 255 in float32) and Normalize ((x - mean) / std, float32,
    /root/reference/fix_train.py:299-318; mean 0 / std 1 when normalize is False) followed by
    input_u8 / input_s8 (fix_train.py:676-692).  One float32 rounding per operation, like torch."""
    p = np.asarray(pix, dtype=np.uint8)
    x = p.astype(np.float32) / np.float32(255.0)
    if normalize:
        x = (x - np.asarray(mean, dtype=np.float32)) / np.asarray(std, dtype=np.float32)
    x = np.moveaxis(x.astype(np.float32), -1, -3)          # HWC -> CHW
    return input_s8(x, fl) if normalize else input_u8(x)

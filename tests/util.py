"""Layout helpers shared by the parity tests: the reference's int32 NCHW tensors <-> the
engine's channel-padded NHWC buffers."""
import numpy as np

CH_ALIGN = 16


def cpad(c):
    return (c + CH_ALIGN - 1) // CH_ALIGN * CH_ALIGN


def nchw_to_nhwc8(x, c_pad=None, signed=False):
    """int32 [N,C,H,W] with 8-bit-range values -> uint8/int8 [N,H,W,c_pad] (zero padded)."""
    n, c, h, w = x.shape
    c_pad = c_pad or cpad(c)
    out = np.zeros((n, h, w, c_pad), dtype=np.int8 if signed else np.uint8)
    out[..., :c] = np.transpose(x, (0, 2, 3, 1)).astype(out.dtype)
    return out


def nchw_to_nhwc32(x, c_pad=None):
    n, c, h, w = x.shape
    c_pad = c_pad or cpad(c)
    out = np.zeros((n, h, w, c_pad), dtype=np.int32)
    out[..., :c] = np.transpose(x, (0, 2, 3, 1))
    return out


def nhwc_to_nchw(y, c):
    """[N,H,W,c_pad] (any integer dtype) -> int32 [N,c,H,W]."""
    return np.ascontiguousarray(np.transpose(y[..., :c], (0, 3, 1, 2))).astype(np.int32)


def checksum(a):
    """Position-weighted checksum mod 2^64 (same as tests/golden/make_golden.py)."""
    a = np.ascontiguousarray(a).reshape(-1).astype(np.int64).view(np.uint64)
    wts = (np.arange(a.size, dtype=np.uint64) % np.uint64(65521)) + np.uint64(1)
    with np.errstate(over="ignore"):
        return np.uint64((a * wts).sum(dtype=np.uint64))

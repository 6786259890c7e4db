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


def nchw_to_carry(x, c_pad=None):
    """int32 [N,C,H,W] -> the engine's int32 carry layout (csrc/f8_common.cuh): with p the pixel
    index in image-major NHW order, element (p, c) at ((p>>7)*(C/4) + (c>>2))*512 + (p&127)*4 + (c&3).
    Returns a flat int32 array of ceil128(N*H*W) * c_pad elements."""
    n, c, h, w = x.shape
    c_pad = c_pad or cpad(c)
    P = n * h * w
    P128 = (P + 127) // 128 * 128
    flat = np.zeros((P128, c_pad), dtype=np.int32)
    flat[:P, :c] = np.transpose(x, (0, 2, 3, 1)).reshape(P, c)
    return np.ascontiguousarray(flat.reshape(P128 // 128, 128, c_pad // 4, 4).transpose(0, 2, 1, 3)).reshape(-1)


def carry_to_nchw(buf, n, c, h, w, c_pad=None):
    """Inverse of nchw_to_carry."""
    c_pad = c_pad or cpad(c)
    P = n * h * w
    P128 = (P + 127) // 128 * 128
    flat = np.asarray(buf).reshape(-1)[:P128 * c_pad].reshape(P128 // 128, c_pad // 4, 128, 4)
    flat = flat.transpose(0, 2, 1, 3).reshape(P128, c_pad)[:P, :c]
    return np.ascontiguousarray(flat.reshape(n, h, w, c).transpose(0, 3, 1, 2)).astype(np.int32)


def carry_elems(n, h, w, c_pad):
    return (n * h * w + 127) // 128 * 128 * c_pad


# ---- round-2 fixture families (tests/golden/make_variant_golden.py) --------------------------
import os as _os

_GOLD = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "golden")
TRAINED_SEED, QMP_SEED = 5150, 808


def trained_fixture(name):
    """(arch, head_signed, state_dict, x, golden) of the trained-fraclen family."""
    from f8net_b200 import synth
    arch, hs, sd = synth.make_trained_state_dict(name)
    x = synth.make_input(arch, 2, hs, seed=TRAINED_SEED)
    return arch, hs, sd, x, np.load(_os.path.join(_GOLD, f"trained_{name}_n2.npz"))


def qmaxpool_fixture(arch):
    """(head_signed, state_dict, x, golden) of the FXQMaxPool2d-vs-float-pool family."""
    from f8net_b200 import synth
    hs = synth.HEAD_SIGNED[arch]
    sd = synth.make_maxpool_state_dict(arch, hs)
    x = synth.make_input(arch, 2, hs, seed=QMP_SEED)
    return hs, sd, x, np.load(_os.path.join(_GOLD, f"qmaxpool_{arch}_n2.npz"))

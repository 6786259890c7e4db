"""ctypes binding of libf8b200.so (include/f8b200.h).

The library is the only compute path of this package.  If it is missing or does not export
a symbol the header declares, importing fails loudly -- there is no CPU / eager fallback.
"""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libf8b200.so")

F8_ABI_VERSION = 2
F8_OK, F8_ERR_ARG, F8_ERR_CUDA, F8_ERR_UNSUPPORTED, F8_ERR_NOMEM, F8_ERR_RANGE = 0, -1, -2, -3, -4, -5
F8_IN_NCHW_I32, F8_IN_NHWC4_8, F8_IN_NCHW_F32, F8_IN_NHWC3_U8 = 0, 1, 2, 3
F8_OPF_INT_MAXPOOL = 1
F8_OP_CONVERT_INPUT, F8_OP_CONV_DENSE, F8_OP_CONV_DW, F8_OP_MAXPOOL, F8_OP_POOL_REQUANT, \
    F8_OP_HEAD_POOL, F8_OP_POOL_FC = range(7)

_i32 = ctypes.c_int32
_i32p = ctypes.POINTER(ctypes.c_int32)
_vp = ctypes.c_void_p


class f8_op(ctypes.Structure):
    _fields_ = [
        ("kind", _i32), ("cin", _i32), ("cout", _i32), ("cin_pad", _i32), ("cout_pad", _i32),
        ("kh", _i32), ("kw", _i32), ("stride", _i32), ("pad", _i32),
        ("hin", _i32), ("win", _i32), ("hout", _i32), ("wout", _i32),
        ("in_signed", _i32), ("in_buf", _i32),
        ("weight", _vp), ("bias", _vp),
        ("carry_in_buf", _i32), ("carry_shift", _i32), ("relu", _i32), ("carry_out_buf", _i32),
        ("out_buf", _i32 * 2), ("out_shift", _i32 * 2), ("out_signed", _i32 * 2),
        ("out_f32", _i32), ("flags", _i32),
    ]


class f8_buffer(ctypes.Structure):
    _fields_ = [("bytes_per_image", ctypes.c_int64), ("offset_per_image", ctypes.c_int64)]


class f8_model_desc(ctypes.Structure):
    _fields_ = [
        ("abi_version", _i32), ("n_ops", _i32), ("ops", ctypes.POINTER(f8_op)),
        ("n_buffers", _i32), ("buffers", ctypes.POINTER(f8_buffer)),
        ("workspace_per_image", ctypes.c_int64),
        ("image_h", _i32), ("image_w", _i32), ("num_classes", _i32), ("head_signed", _i32),
    ]


class f8_conv_args(ctypes.Structure):
    _fields_ = [
        ("n", _i32), ("cin", _i32), ("cout", _i32), ("cin_pad", _i32), ("cout_pad", _i32),
        ("kh", _i32), ("kw", _i32), ("stride", _i32), ("pad", _i32),
        ("hin", _i32), ("win", _i32), ("hout", _i32), ("wout", _i32),
        ("in_signed", _i32),
        ("in_", _vp), ("wpack", _vp), ("bias", _vp), ("carry_in", _vp),
        ("carry_shift", _i32), ("relu", _i32),
        ("carry_out", _vp),
        ("out", _vp * 2), ("out_shift", _i32 * 2), ("out_signed", _i32 * 2),
        ("out_f32", _vp), ("out_f32_ld", _i32), ("flags", _i32), ("wpack_stage", _vp),
    ]


# every symbol include/f8b200.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "f8_plan_create": (ctypes.c_int, [ctypes.POINTER(f8_model_desc), ctypes.c_int,
                                      ctypes.POINTER(_vp)]),
    "f8_plan_destroy": (None, [_vp]),
    "f8_plan_workspace_bytes": (ctypes.c_int, [_vp, ctypes.c_int, ctypes.POINTER(ctypes.c_size_t)]),
    "f8_plan_run": (ctypes.c_int, [_vp, _vp, ctypes.c_int, ctypes.c_int, _vp, _vp,
                                   ctypes.c_size_t, ctypes.c_int, _vp]),
    "f8_plan_run_host": (ctypes.c_int, [_vp, _vp, ctypes.c_int, ctypes.c_int, _vp, _vp, _vp, _vp,
                                        ctypes.c_size_t, ctypes.c_int, ctypes.c_int, _vp]),
    "f8_plan_profile": (ctypes.c_int, [_vp, _vp, ctypes.c_int, ctypes.c_int, _vp, _vp,
                                       ctypes.c_size_t, ctypes.c_int, _vp,
                                       ctypes.POINTER(ctypes.c_float), ctypes.c_int]),
    "f8_plan_kernel_name": (ctypes.c_int, [_vp, ctypes.c_int, ctypes.c_char_p, ctypes.c_int]),
    "f8_plan_read_buffer": (ctypes.c_int, [_vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, _vp, _vp,
                                           ctypes.c_size_t, _vp]),
    "f8_host_pack_info": (ctypes.c_char_p, [ctypes.POINTER(ctypes.c_int)]),
    "f8_plan_last_raw_images": (ctypes.c_int, [_vp]),
    "f8_plan_input_range": (ctypes.c_int, [_vp, ctypes.c_int]),
    "f8_plan_launch_count": (ctypes.c_int, [_vp, ctypes.c_int, ctypes.c_int, ctypes.c_int]),
    "f8_plan_set_backend": (ctypes.c_int, [_vp, ctypes.c_int]),
    "f8_pack_weights_bytes": (ctypes.c_size_t, [ctypes.c_int] * 7),
    "f8_pack_weights": (ctypes.c_int, [ctypes.c_int, _vp] + [ctypes.c_int] * 6 + [_vp]),
    "f8_pack_weights_stage3x3_bytes": (ctypes.c_size_t, [ctypes.c_int, ctypes.c_int]),
    "f8_pack_weights_stage3x3": (ctypes.c_int, [_vp, ctypes.c_int, ctypes.c_int, _vp]),
    "f8_conv_dense": (ctypes.c_int, [ctypes.POINTER(f8_conv_args), ctypes.c_int, _vp]),
    "f8_conv_dw3x3": (ctypes.c_int, [ctypes.POINTER(f8_conv_args), _vp]),
    "f8_maxpool3x3s2": (ctypes.c_int, [ctypes.POINTER(f8_conv_args), _vp]),
    "f8_pool_requant": (ctypes.c_int, [ctypes.POINTER(f8_conv_args), _vp]),
    "f8_head_pool": (ctypes.c_int, [ctypes.POINTER(f8_conv_args), _vp]),
    "f8_pool_fc": (ctypes.c_int, [ctypes.POINTER(f8_conv_args), _vp]),
    "f8_convert_input": (ctypes.c_int, [_vp, _vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, _vp]),
    "f8_requant_i32": (ctypes.c_int, [_vp, _vp, ctypes.c_size_t, ctypes.c_int, ctypes.c_int,
                                      ctypes.c_int, _vp]),
    "f8_plan_set_input_prep": (ctypes.c_int, [_vp, ctypes.c_int, ctypes.c_int,
                                              ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_float)]),
    "f8_integerize_f32": (ctypes.c_int, [_vp, _vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                         ctypes.c_int, _vp]),
    "f8_integerize_u8": (ctypes.c_int, [_vp, _vp, _vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, _vp]),
    "f8_pack_input_host": (ctypes.c_int, [_vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, _vp, ctypes.c_int, ctypes.c_int]),
    "f8_make_input_lut": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_float),
                                         ctypes.POINTER(ctypes.c_float), _vp]),
    "f8_last_error": (ctypes.c_char_p, []),
    "f8_abi_version": (ctypes.c_int, []),
    "f8_has_umma": (ctypes.c_int, [ctypes.c_int]),
}

_lib = None


class F8Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libf8b200 error {code}: {msg}")
        self.code = code


def lib():
    """The loaded library.  Raises if the CUDA extension has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: the CUDA extension is the only compute path of "
                "f8net_b200 (no CPU fallback).  Build it with `python -m f8net_b200.build`.")
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)       # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        if L.f8_abi_version() != F8_ABI_VERSION:
            raise ImportError("libf8b200.so ABI version mismatch; rebuild it")
        _lib = L
    return _lib


def check(rc):
    if rc != F8_OK:
        msg = lib().f8_last_error()
        raise F8Error(rc, msg.decode() if msg else "?")

"""f8net_b200 -- B200 (sm_100a) engine for F8Net's integer-only (int_op_only) forward path.

Only what the hot path needs lives here: the CUDA kernels + C ABI (csrc/, libf8b200.so), the
integer-graph description of the four networks (arch), the host planner that fuses the
reference's tensor-op chain into kernel epilogues (planner) and the reference-facing call
surface (engine).  ``synth`` makes the seeded synthetic workloads of bench.py / the tests.
"""
from .arch import ARCHS, graph_for, graph_from_module  # noqa: F401


def compile(*args, **kwargs):
    from .engine import compile as _compile
    return _compile(*args, **kwargs)


def build(force=False):
    from .build import build as _build
    return _build(force=force)


def export_int_state_dict(*args, **kwargs):
    """Float-simulation checkpoint -> IntModel state_dict (Model.int_model())."""
    from .export import export_int_state_dict as _e
    return _e(*args, **kwargs)


def compile_float(*args, **kwargs):
    from .export import compile_float as _c
    return _c(*args, **kwargs)


def export_onnx(*args, **kwargs):
    """Opset-11 ONNX file of the integer graph (the reference's onnx_export for the int_op_only model)."""
    from .onnx_export import export_onnx as _e
    return _e(*args, **kwargs)


def save_engine(*args, **kwargs):
    """Self-contained int8 engine file (weights + formats + graph meta)."""
    from .engine_file import save_engine as _s
    return _s(*args, **kwargs)


def load_engine(*args, **kwargs):
    from .engine_file import load_engine as _l
    return _l(*args, **kwargs)

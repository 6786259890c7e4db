"""Builds f8net_b200/libf8b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m f8net_b200.build [--force]

The library is the product's only compute path: the Python host fails loudly when it is
missing (f8net_b200/_capi.py), there is no eager / CPU fallback.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libf8b200.so")
SOURCES = ["plan.cu", "conv_mma.cu", "conv_umma.cu", "conv3x3_umma.cu", "head_pool2_umma.cu", "head3x3_umma.cu", "dw_conv.cu", "pool_misc.cu", "host_pack.cpp"]
HEADERS = [os.path.join(CSRC, "f8_common.cuh"), os.path.join(CSRC, "umma_common.cuh"), os.path.join(CSRC, "tma_common.cuh"), os.path.join(CSRC, "host_pack.h"), os.path.join(HERE, "..", "include", "f8b200.h")]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC=/path/to/nvcc)")


def sources():
    return [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]


def stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(p) > t for p in sources() + HEADERS if os.path.exists(p))


def build(force=False, verbose=False):
    """Compile every .cu under csrc/ into one shared library.  Returns its path."""
    if not force and not stale():
        return LIB
    srcs = sources()
    cmd = [_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3",
           "-std=c++17", "-Xcompiler", "-fPIC,-fvisibility=hidden", "--threads", "4",
           "-shared", "-o", LIB]
    if any(s.endswith("conv_umma.cu") for s in srcs):
        cmd += ["-DF8_WITH_UMMA"]
    if os.environ.get("F8_DEBUG_PROBES"):      # in-kernel wait counters / timing probes (never in the shipping build)
        cmd += ["-DF8_DEBUG_PROBES"]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    cmd += srcs
    subprocess.check_call(cmd, cwd=CSRC)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))

"""The C-ABI boundary without a GPU: the library loads, exports every symbol the header
declares, the ctypes mirrors of the structs have the C compiler's layout, and the host-only
weight packer produces the documented images.  No compute call is made here."""
import ctypes
import os
import re
import subprocess
import sys
import tempfile

import numpy as np
import pytest

from f8net_b200 import _capi as C

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "f8b200.h")


def _declared_symbols():
    src = open(HEADER).read()
    return sorted(set(re.findall(r"F8_API\s+[\w\s\*]+?\b(f8_\w+)\s*\(", src)))


def test_header_symbols_are_exported_and_bound(f8lib):
    names = _declared_symbols()
    assert len(names) >= 18
    for n in names:
        assert hasattr(f8lib, n), f"{n} declared in include/f8b200.h but not exported"
        assert n in C.SYMBOLS, f"{n} has no ctypes prototype in f8net_b200/_capi.py"
    assert sorted(C.SYMBOLS) == names
    assert f8lib.f8_abi_version() == C.F8_ABI_VERSION


def test_struct_layouts_match_the_c_compiler():
    prog = r"""
#include <stdio.h>
#include <stddef.h>
#include "f8b200.h"
int main(void) {
  printf("%zu %zu %zu %zu ", sizeof(f8_op), sizeof(f8_buffer), sizeof(f8_model_desc), sizeof(f8_conv_args));
  printf("%zu %zu %zu %zu ", offsetof(f8_op, weight), offsetof(f8_op, carry_in_buf), offsetof(f8_op, out_buf), offsetof(f8_op, out_f32));
  printf("%zu %zu %zu %zu %zu\n", offsetof(f8_conv_args, in), offsetof(f8_conv_args, carry_shift), offsetof(f8_conv_args, carry_out), offsetof(f8_conv_args, out_f32), offsetof(f8_model_desc, workspace_per_image));
  printf("%zu %zu %d %zu\n", offsetof(f8_op, flags), offsetof(f8_conv_args, flags), (int)F8_OPF_INT_MAXPOOL, offsetof(f8_conv_args, wpack_stage));
  return 0; }
"""
    with tempfile.TemporaryDirectory() as td:
        c = os.path.join(td, "t.c")
        open(c, "w").write(prog)
        exe = os.path.join(td, "t")
        subprocess.check_call(["gcc", "-std=c11", "-I", os.path.join(ROOT, "include"), c, "-o", exe])
        got = [int(v) for v in subprocess.check_output([exe]).split()]
    want = [ctypes.sizeof(C.f8_op), ctypes.sizeof(C.f8_buffer), ctypes.sizeof(C.f8_model_desc),
            ctypes.sizeof(C.f8_conv_args),
            C.f8_op.weight.offset, C.f8_op.carry_in_buf.offset, C.f8_op.out_buf.offset,
            C.f8_op.out_f32.offset,
            C.f8_conv_args.in_.offset, C.f8_conv_args.carry_shift.offset,
            C.f8_conv_args.carry_out.offset, C.f8_conv_args.out_f32.offset,
            C.f8_model_desc.workspace_per_image.offset,
            C.f8_op.flags.offset, C.f8_conv_args.flags.offset, C.F8_OPF_INT_MAXPOOL, C.f8_conv_args.wpack_stage.offset]
    assert got == want


def _pack(f8lib, kind, w, cin_pad, cout_pad):
    cout, cin_g, kh, kw = w.shape
    cin = cout if kind == C.F8_OP_CONV_DW else cin_g
    n = f8lib.f8_pack_weights_bytes(kind, cin, cout, cin_pad, cout_pad, kh, kw)
    dst = np.full(n, 0x5A, dtype=np.uint8)
    w = np.ascontiguousarray(w, dtype=np.int32)
    rc = f8lib.f8_pack_weights(kind, w.ctypes.data, cin, cout, cin_pad, cout_pad, kh, kw,
                               dst.ctypes.data)
    return rc, dst


def test_pack_dense_generic_layout(f8lib):
    rng = np.random.default_rng(1)
    w = rng.integers(-127, 128, (24, 20, 3, 3)).astype(np.int32)
    rc, img = _pack(f8lib, C.F8_OP_CONV_DENSE, w, 32, 32)
    assert rc == 0
    K = 3 * 3 * 32
    kp = (K + 63) // 64 * 64
    # image is [K_pad/16][256][16]; un-chunk it to [rows][K_pad]
    img = img.view(np.int8).reshape(kp // 16, 256, 16).transpose(1, 0, 2).reshape(256, kp)
    want = np.zeros((256, kp), np.int8)
    # k = (r*kw + s)*cin_pad + c
    want[:24, :K].reshape(24, 3, 3, 32)[..., :20] = np.transpose(w, (0, 2, 3, 1))
    assert np.array_equal(img, want)


def test_pack_dense_small_c_row_window(f8lib):
    rng = np.random.default_rng(2)
    for k in (7, 3):
        w = rng.integers(-127, 128, (32, 3, k, k)).astype(np.int32)
        rc, img = _pack(f8lib, C.F8_OP_CONV_DENSE, w, 4, 32)
        assert rc == 0
        px = (k + 1 + 1) // 2 * 2            # window pixels: one extra on the left, even count
        K = k * px * 4
        kp = (K + 63) // 64 * 64
        img = img.view(np.int8).reshape(kp // 16, 256, 16).transpose(1, 0, 2).reshape(256, kp)
        want = np.zeros((256, kp), np.int8)
        v = want[:32, :K].reshape(32, k, px, 4)
        v[:, :, 1:1 + k, :3] = np.transpose(w, (0, 2, 3, 1))
        assert np.array_equal(img, want)


def test_pack_depthwise_dp4a_operands(f8lib):
    rng = np.random.default_rng(3)
    w = rng.integers(-127, 128, (24, 1, 3, 3)).astype(np.int32)
    rc, img = _pack(f8lib, C.F8_OP_CONV_DW, w, 32, 32)
    assert rc == 0
    raw = img
    img = raw[:12 * 32].view(np.uint32).reshape(12, 8)        # dp4a words of the CUDA-core kernel
    for ch in range(24):
        taps = w[ch, 0].reshape(9)
        for k in range(3):
            word = int(img[(ch % 4) * 3 + k, ch // 4])
            for byte in range(4):
                t = 4 * k + byte
                want = int(taps[t]) & 0xFF if t < 9 else 0
                assert (word >> (8 * byte)) & 0xFF == want
    assert not img[:, 6:].any()              # padded channels stay zero
    # tensor-core form behind it (256-byte aligned): one diagonal 64 x 64 image per channel group,
    # [36 chunks][64 rows][16 B], K byte k = tap * 64 + c of row o = c
    dense = raw[512:].view(np.int8).reshape(36, 64, 16)
    want = np.zeros((36, 64, 16), np.int8)
    for ch in range(24):
        for t in range(9):
            want[t * 4 + ch // 16, ch, ch % 16] = w[ch, 0].reshape(9)[t]
    assert raw.size == 512 + 36 * 64 * 16 and np.array_equal(dense, want)


def test_pack_rejects_weights_outside_8_bits(f8lib):
    w = np.zeros((16, 16, 1, 1), np.int32)
    w[3, 2, 0, 0] = 200
    rc, _ = _pack(f8lib, C.F8_OP_CONV_DENSE, w, 16, 16)
    assert rc == C.F8_ERR_UNSUPPORTED
    assert b"8-bit" in f8lib.f8_last_error()


def test_plan_create_rejects_bad_descriptors(f8lib):
    h = ctypes.c_void_p()
    assert f8lib.f8_plan_create(None, 0, ctypes.byref(h)) == C.F8_ERR_ARG
    d = C.f8_model_desc()
    d.abi_version = 99
    assert f8lib.f8_plan_create(ctypes.byref(d), 0, ctypes.byref(h)) == C.F8_ERR_ARG
    assert b"ABI" in f8lib.f8_last_error()


def test_missing_library_fails_loudly(monkeypatch):
    monkeypatch.setattr(C, "_lib", None)
    monkeypatch.setattr(C, "LIB_PATH", "/nonexistent/libf8b200.so")
    with pytest.raises(ImportError, match="no CPU fallback"):
        C.lib()


def test_input_lut_matches_reference(f8lib):
    """f8_make_input_lut (host code of the product) against the reference-generated vectors."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "input_prep.npz"))
    pix = g["u8_pix"]
    mean = (ctypes.c_float * 3)(0.485, 0.456, 0.406)
    std = (ctypes.c_float * 3)(0.229, 0.224, 0.225)
    for normalize, fl, key in ((0, 8, "u8_u_y"), (1, 3, "u8_s3_y"), (1, 5, "u8_s5_y"), (1, 7, "u8_s7_y")):
        lut = np.zeros(768, dtype=np.uint8)
        assert f8lib.f8_make_input_lut(normalize, fl, mean, std, lut.ctypes.data) == 0
        got = np.stack([lut[c * 256 + pix[..., c].astype(np.int64)] for c in range(3)])   # [3,16,16]
        want = g[key][0]
        got = got.view(np.int8).astype(np.int32) if normalize else got.astype(np.int32)
        assert np.array_equal(got, want), key


def test_pack_input_host_low_bytes_nhwc4(f8lib):
    """The host-side narrowing of f8_plan_run_host (int32 NCHW -> NHWC4 low bytes), every thread
    count, ragged shapes (row tails, fewer rows than threads), unsigned and signed ranges; repeated
    calls reuse the persistent helper threads."""
    rng = np.random.default_rng(3)
    for n, h, w in [(1, 1, 1), (2, 5, 7), (3, 64, 30), (5, 224, 224), (9, 33, 18)]:
        for lo, hi in [(0, 256), (-127, 128)]:
            x = rng.integers(lo, hi, (n, 3, h, w)).astype(np.int32)
            want = np.zeros((n, h, w, 4), np.uint8)
            want[..., :3] = (x.transpose(0, 2, 3, 1) & 0xff).astype(np.uint8)
            for threads in (0, 1, 3, 16):
                out = np.full((n, h, w, 4), 0x55, np.uint8)
                assert f8lib.f8_pack_input_host(x.ctypes.data, n, h, w, out.ctypes.data, threads, int(lo < 0)) == 0
                assert np.array_equal(out, want), (n, h, w, lo, threads)
    assert f8lib.f8_pack_input_host(None, 1, 1, 1, None, 1, 0) != 0


def test_pack_input_host_reports_out_of_range_values(f8lib):
    """The narrowing pass is also the range check (the reference's head conv consumes the full
    int32, fix_resnet.py:355): one value outside the head's 8 bits anywhere -- vector body, row tail, any
    channel, any helper thread's share -- gives F8_ERR_RANGE, with the low bytes still written; the bounds
    themselves pass.  Unsigned head: [0, 255]; signed head: [-128, 127]."""
    from f8net_b200 import _capi as C
    rng = np.random.default_rng(17)
    n, h, w = 3, 70, 37                                     # 37 = two AVX-512 vectors + a 5-wide tail
    for signed, lo, hi in [(0, 0, 255), (1, -128, 127)]:
        base = rng.integers(lo, hi + 1, (n, 3, h, w)).astype(np.int32)
        base[0, 0, 0, 0], base[n - 1, 2, h - 1, w - 1] = lo, hi            # the bounds are in range
        out = np.zeros((n, h, w, 4), np.uint8)
        for threads in (1, 4):
            assert f8lib.f8_pack_input_host(base.ctypes.data, n, h, w, out.ctypes.data, threads, signed) == 0
        spots = [(0, 0, 0, 0), (1, 1, 35, 36), (2, 2, 69, 20), (0, 1, 10, 31), (2, 0, 64, 33)]
        for bad in (lo - 1, hi + 1, 1 << 20, -(1 << 31), 0x100 + lo if lo else 0x1ff):
            for spot in spots:
                x = base.copy()
                x[spot] = bad
                for threads in (1, 4):
                    rc = f8lib.f8_pack_input_host(x.ctypes.data, n, h, w, out.ctypes.data, threads, signed)
                    assert rc == C.F8_ERR_RANGE, (signed, bad, spot, threads, rc)
                    assert np.array_equal(out[..., :3], (x.transpose(0, 2, 3, 1) & 0xff).astype(np.uint8))
        # the other head's range is out of range here
        other = rng.integers(-128, 0, (n, 3, h, w)).astype(np.int32) if not signed else \
            rng.integers(128, 256, (n, 3, h, w)).astype(np.int32)
        assert f8lib.f8_pack_input_host(other.ctypes.data, n, h, w, out.ctypes.data, 2, signed) == C.F8_ERR_RANGE


@pytest.mark.parametrize("isa", [0, 1, 2, 3])
def test_pack_input_host_every_simd_body(f8lib, isa):
    """The scalar / SSE2 / AVX2 / AVX-512 bodies of the host-side narrowing (host_pack.cpp; selected once
    per process, F8_HOST_PACK_ISA overrides the CPU detection) give the same bytes, also for widths that
    leave vector tails and outputs that are not vector aligned."""
    code = (
        "import numpy as np, ctypes\n"
        "from f8net_b200 import _capi as C\n"
        "lib = C.lib()\n"
        "t = ctypes.c_int(0)\n"
        "name = lib.f8_host_pack_info(ctypes.byref(t)).decode()\n"
        "rng = np.random.default_rng(5)\n"
        "for n, h, w in [(2, 7, 224), (3, 5, 37), (1, 3, 16), (2, 2, 9)]:\n"
        "    x = rng.integers(0, 256, (n, 3, h, w)).astype(np.int32)\n"
        "    want = np.zeros((n, h, w, 4), np.uint8)\n"
        "    want[..., :3] = (x.transpose(0, 2, 3, 1) & 0xff).astype(np.uint8)\n"
        "    for off in (0, 4, 16):\n"
        "        buf = np.full(n * h * w * 4 + 64 + off, 0x55, np.uint8)\n"
        "        base = (-buf.ctypes.data) % 64 + off\n"
        "        out = buf[base:base + n * h * w * 4]\n"
        "        assert lib.f8_pack_input_host(x.ctypes.data, n, h, w, out.ctypes.data, 3, 0) == 0\n"
        "        assert np.array_equal(out.reshape(n, h, w, 4), want), (n, h, w, off)\n"
        "        for k in range(w):\n"
        "            y = x.copy(); y[n - 1, k % 3, h - 1, k] = 256 + k\n"
        "            assert lib.f8_pack_input_host(y.ctypes.data, n, h, w, out.ctypes.data, 1, 0) == C.F8_ERR_RANGE, (w, k)\n"
        "            y[n - 1, k % 3, h - 1, k] = -1 - k\n"
        "            assert lib.f8_pack_input_host(y.ctypes.data, n, h, w, out.ctypes.data, 2, 0) == C.F8_ERR_RANGE, (w, k)\n"
        "print('ISA', name, t.value)\n")
    env = dict(os.environ, F8_HOST_PACK_ISA=str(isa))
    r = subprocess.run([sys.executable, "-c", code], env=env, cwd=ROOT, capture_output=True, text=True, timeout=120)
    if r.returncode != 0 and "Illegal instruction" in r.stderr + str(r.returncode):
        pytest.skip("this CPU lacks the instruction set")
    assert r.returncode == 0 and "ISA " + ["scalar", "sse2", "avx2", "avx512"][isa] in r.stdout, (r.stdout, r.stderr[-1500:])


def test_pack_stage3x3_is_a_permutation_of_the_dense_pack(f8lib):
    """f8_pack_weights_stage3x3: tiles of 128 (or 64) output rows; per tile, 64-channel group, filter row, filter
    column: four 16-byte K chunks of T rows.  Every weight lands where the kernel's stage copy expects it."""
    rng = np.random.default_rng(11)
    for cin, cout in [(64, 64), (128, 200), (64, 48)]:
        cin_pad, cout_pad = (cin + 15) // 16 * 16, (cout + 15) // 16 * 16
        w = rng.integers(-127, 128, (cout, cin, 3, 3)).astype(np.int32)
        rc, dense = _pack(f8lib, C.F8_OP_CONV_DENSE, w, cin_pad, cout_pad)
        assert rc == 0
        n = f8lib.f8_pack_weights_stage3x3_bytes(cin_pad, cout_pad)
        T = 128 if cout_pad > 64 else 64
        assert n == -(-cout_pad // T) * T * cin_pad * 9
        st = np.full(n, 0x5A, np.uint8)
        assert f8lib.f8_pack_weights_stage3x3(dense.ctypes.data, cin_pad, cout_pad, st.ctypes.data) == 0
        img = st.view(np.int8).reshape(-1, cin_pad // 64, 3, 3, 4, T, 16)      # [tile][group][r][s][chunk][row][byte]
        for o, c, r, s_ in [(0, 0, 0, 0), (cout - 1, cin - 1, 2, 2), (cout // 2, 17, 1, 2), (5, 63, 2, 0)]:
            assert img[o // T, c // 64, r, s_, (c % 64) // 16, o % T, c % 16] == w[o, c, r, s_]
        assert not img[-1, :, :, :, :, (cout_pad - 1) % T + 1:, :].any() or cout_pad % T == 0
    assert f8lib.f8_pack_weights_stage3x3_bytes(48, 64) == 0

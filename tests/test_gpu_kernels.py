"""Layer-level parity on the GPU: every kernel of libf8b200.so, called through the C ABI,
against the CPU oracle on the same seeded inputs.  Bit-exact (integer path): equality, no
tolerance."""
import ctypes

import numpy as np
import pytest

torch = pytest.importorskip("torch")

from f8net_b200 import _capi as C  # noqa: E402
from oracle import oracle as O  # noqa: E402

from util import carry_elems, carry_to_nchw, cpad, nchw_to_carry, nhwc_to_nchw  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def G(cuda, f8lib):
    import gpu_util
    return gpu_util


def _rand_layer(rng, cin, cout, k, signed, groups=1, big_bias=False):
    lo, hi = (-127, 128) if signed else (0, 256)
    w = rng.integers(-127, 128, (cout, cin // groups, k, k)).astype(np.int32)
    bmax = 2 ** 31 if big_bias else 2 ** 14
    b = rng.integers(-bmax, bmax, (cout,)).astype(np.int32)
    return lo, hi, w, b


def test_requant_i32_matches_reference_formula(G, f8lib):
    rng = np.random.default_rng(0)
    xs = np.concatenate([
        rng.integers(-2 ** 31, 2 ** 31, 5000), rng.integers(-70000, 70000, 5000),
        np.array([0, 1, -1, 2 ** 31 - 1, -2 ** 31, 2 ** 31 - 2, -2 ** 31 + 1, 2 ** 24 + 1,
                  -1610612736, 2 ** 29, 2 ** 30])]).astype(np.int32)
    for n in list(range(-8, 17)) + [24, 30]:
        ties = np.array([(2 * k + 1) << (n - 1) for k in range(-64, 64)] if n > 0 else [0],
                        dtype=np.int64).astype(np.int32)
        v = np.concatenate([xs, ties])
        xd = G.dev(v)
        for signed in (False, True):
            fl = max(0, -n)
            if fl > (7 if signed else 8):
                continue
            yd = torch.empty_like(xd)
            C.check(f8lib.f8_requant_i32(xd.data_ptr(), yd.data_ptr(), v.size, fl, fl + n,
                                         int(signed), torch.cuda.current_stream().cuda_stream))
            want = O.requant(v, fl, fl + n, signed)
            assert np.array_equal(yd.cpu().numpy(), want), (n, signed)


# (cin, cout, k, stride, pad, h) -- every dense shape class of SURVEY.md Appendix A, shrunk
DENSE_SHAPES = [
    (3, 64, 7, 2, 3, 32),      # ResNet head (small-C row-window gather)
    (3, 32, 3, 2, 1, 32),      # MobileNet head
    (64, 64, 3, 1, 1, 14),     # 3x3 s1
    (64, 128, 3, 2, 1, 14),    # 3x3 s2
    (64, 128, 1, 2, 0, 14),    # 1x1 s2 shortcut
    (256, 64, 1, 1, 0, 7),     # 1x1 reduce
    (32, 16, 1, 1, 0, 12),     # MBV2 project to 16
    (16, 96, 1, 1, 0, 12),     # K = 16 < one MMA K step
    (24, 144, 1, 1, 0, 9),     # 24 -> padded 32
    (144, 24, 1, 1, 0, 9),     # K = 144 (not a multiple of 64)
    (160, 960, 1, 1, 0, 7),
    (512, 512, 3, 1, 1, 7),    # K = 4608
    (64, 64, 3, 1, 1, 56),     # resident-patch kernel: 2 rows per M segment
    (128, 192, 3, 1, 1, 28),   # 2 channel groups, 3 N tiles
    (256, 256, 3, 1, 1, 14),
]


BACKENDS = [pytest.param(0, id="imma"), pytest.param(1, id="tcgen05")]


@pytest.mark.parametrize("backend", BACKENDS)
@pytest.mark.parametrize("shape", DENSE_SHAPES, ids=lambda s: "x".join(map(str, s)))
@pytest.mark.parametrize("signed", [False, True])
def test_conv_dense_plain(G, f8lib, shape, signed, backend):
    cin, cout, k, st, pd, h = shape
    rng = np.random.default_rng(hash(shape) % 2 ** 31)
    lo, hi, w, b = _rand_layer(rng, cin, cout, k, signed)
    x = rng.integers(lo, hi, (3, cin, h, h)).astype(np.int32)
    outs = ((9, False), (7, True))
    v, q, _ = G.run_conv(f8lib, x, w, b, st, pd, in_signed=signed, relu=True, outs=outs,
                         backend=backend)
    ev, eq = G.expect_conv(x, w, b, st, pd, relu=True, outs=outs)
    assert np.array_equal(v, ev)
    assert np.array_equal(q[0], eq[0]) and np.array_equal(q[1], eq[1])


@pytest.mark.parametrize("signed", [False, True])
@pytest.mark.parametrize("geom", [(1, 8, 8), (3, 32, 32), (2, 30, 44), (5, 224, 224)],
                         ids=lambda g: "x".join(map(str, g)))
def test_mobilenet_head_space_to_depth(G, f8lib, geom, signed):
    """head[0] of the MobileNets (3 -> 32, k3 s2 p1) + ReLU + one unsigned requant in the
    space-to-depth kernel (head3x3_umma.cu): ragged batches, non-square even images, every tile
    boundary of the padded linear space, exact ties; against the generic gather path as well."""
    if not f8lib.f8_has_umma(0):
        pytest.skip("tcgen05 backend not available")
    n, h, wd = geom
    rng = np.random.default_rng(n * 1000 + h + signed)
    lo, hi, w, b = _rand_layer(rng, 3, 32, 3, signed)
    w[:, :, 1, 1] = (w[:, :, 1, 1] // 2) * 2
    b[:8] = (b[:8] // 16) * 16 + 8                      # ties of the 4-bit shift below
    x = rng.integers(lo, hi, (n, 3, h, wd)).astype(np.int32)
    x[0, :, 0, :] = hi - 1                              # first / last rows and columns saturated:
    x[-1, :, -1, :] = lo                                # the zero halo must not leak in
    x[:, :, :, 0] = hi - 1
    for shift in (4, 9):
        outs = ((shift, False),)
        _, q, _ = G.run_conv(f8lib, x, w, b, 2, 1, in_signed=signed, relu=True, outs=outs, want_carry=False,
                             backend=1)
        _, eq = G.expect_conv(x, w, b, 2, 1, relu=True, outs=outs)
        assert np.array_equal(q[0], eq[0]), (geom, signed, shift)


@pytest.mark.parametrize("backend", BACKENDS)
@pytest.mark.parametrize("carry_shift", [-3, 0, 2, 29])
def test_conv_dense_residual_epilogue(G, f8lib, carry_shift, backend):
    """IntBlock shift-align / wrapping add / clamp / ReLU (fix_resnet.py:40-77) including
    accumulators at the int32 limits (huge biases) and left-shift requants."""
    rng = np.random.default_rng(10 + carry_shift)
    cin, cout, h = 64, 48, 9
    _, _, w, b = _rand_layer(rng, cin, cout, 3, False, big_bias=True)
    x = rng.integers(0, 256, (2, cin, h, h)).astype(np.int32)
    carry = rng.integers(-2 ** 31, 2 ** 31, (2, cout, h, h)).astype(np.int32)
    carry[0, 0, 0, :4] = [-2 ** 31, 2 ** 31 - 1, 0, 1]
    for relu in (True, False):
        outs = ((11, False), (-2, True))
        v, q, _ = G.run_conv(f8lib, x, w, b, 1, 1, relu=relu, carry=carry,
                             carry_shift=carry_shift, outs=outs, backend=backend)
        ev, eq = G.expect_conv(x, w, b, 1, 1, relu=relu, carry=carry, carry_shift=carry_shift,
                               outs=outs)
        assert np.array_equal(v, ev)
        assert np.array_equal(q[0], eq[0]) and np.array_equal(q[1], eq[1])


@pytest.mark.parametrize("backend", BACKENDS)
def test_conv_dense_ties_round_half_even(G, f8lib, backend):
    """Even weights + odd-half biases make every accumulator an exact tie of the shift."""
    rng = np.random.default_rng(5)
    cin, cout, h = 32, 32, 8
    w = (rng.integers(-63, 64, (cout, cin, 1, 1)) * 2).astype(np.int32)
    x = (rng.integers(0, 64, (2, cin, h, h)) * 4).astype(np.int32)      # acc multiple of 8
    b = np.full(cout, 4, np.int32)                                        # + 4 -> tie for n=3
    outs = ((3, True), (3, False))
    v, q, _ = G.run_conv(f8lib, x, w, b, 1, 0, outs=outs, backend=backend)
    ev, eq = G.expect_conv(x, w, b, 1, 0, outs=outs)
    assert (ev % 8 == 4).all()
    assert np.array_equal(v, ev) and np.array_equal(q[0], eq[0]) and np.array_equal(q[1], eq[1])


@pytest.mark.parametrize("backend", BACKENDS)
def test_linear_float_logits(G, f8lib, backend):
    """nn.Linear + .float() (fix_quant_ops.py:1165-1195, fix_resnet.py:383) as a 1x1 conv on
    a 1x1 image; ragged batch, 1000 classes (cout_pad 1008)."""
    rng = np.random.default_rng(6)
    for n, k in [(1, 512), (5, 1280), (130, 2048)]:
        w = rng.integers(-127, 128, (1000, k, 1, 1)).astype(np.int32)
        b = rng.integers(-2 ** 20, 2 ** 20, (1000,)).astype(np.int32)
        q8 = rng.integers(0, 256, (n, k, 1, 1)).astype(np.int32)
        v, _, f = G.run_conv(f8lib, q8, w, b, 1, 0, outs=(), want_f32=True, backend=backend)
        yi, yf = O.linear(q8.reshape(n, k), w.reshape(1000, k), b)
        assert np.array_equal(v.reshape(n, 1000), yi)
        assert np.array_equal(f.reshape(n, 1000), yf)
    # logits beyond 2^24 round to nearest-even float32 like torch's .float()
    w = np.full((1000, 512, 1, 1), 127, np.int32)
    q8 = np.full((2, 512, 1, 1), 255, np.int32)
    b = np.arange(1000, dtype=np.int32) * 3 + 2 ** 24
    _, _, f = G.run_conv(f8lib, q8, w, b, 1, 0, outs=(), want_f32=True, backend=backend)
    _, yf = O.linear(q8.reshape(2, 512), w.reshape(1000, 512), b)
    assert np.array_equal(f.reshape(2, 1000), yf)


# (channels, stride, input size).  Stride 1 and full-group even-size stride 2 run on the tensor core
# (diagonal weight blocks: conv3x3_umma.cu), the others on the CUDA-core kernel (dw_conv.cu).
DW_SHAPES = [(32, 1, 16), (96, 2, 16), (144, 1, 9), (144, 2, 14), (24, 1, 7), (960, 1, 7),
             (64, 2, 15), (128, 2, 28), (192, 2, 14), (64, 2, 112), (32, 1, 112), (576, 2, 14),
             (160, 1, 14)]


@pytest.mark.parametrize("shape", DW_SHAPES, ids=lambda s: "x".join(map(str, s)))
@pytest.mark.parametrize("signed", [False, True])
def test_conv_depthwise(G, f8lib, shape, signed):
    c, st, h = shape
    rng = np.random.default_rng(c * 7 + st)
    lo, hi, w, b = _rand_layer(rng, c, c, 3, signed, groups=c)
    x = rng.integers(lo, hi, (3, c, h, h)).astype(np.int32)
    outs = ((5, False), (4, True))
    v, q, _ = G.run_conv(f8lib, x, w, b, st, 1, depthwise=True, in_signed=signed, relu=True,
                         outs=outs)
    ev, eq = G.expect_conv(x, w, b, st, 1, depthwise=True, relu=True, outs=outs)
    assert np.array_equal(v, ev)
    assert np.array_equal(q[0], eq[0]) and np.array_equal(q[1], eq[1])


def _args_for_pool(n, c, hin, hout, k, stride, pad):
    a = C.f8_conv_args()
    a.n, a.cin, a.cout, a.cin_pad, a.cout_pad = n, c, c, cpad(c), cpad(c)
    a.kh, a.kw, a.stride, a.pad = k, k, stride, pad
    a.hin, a.win, a.hout, a.wout = hin, hin, hout, hout
    return a


@pytest.mark.parametrize("int_pool", [False, True], ids=["float_rt", "FXQMaxPool2d"])
def test_maxpool_float_round_trip(G, f8lib, int_pool):
    """x = head[-1](x.float()).int()  (fix_resnet.py:358-359) incl. values above 2^24 -- and the
    FLAGS.quant_maxpool variant FXQMaxPool2d (fix_quant_ops.py:141-157): integer max, no round trip."""
    rng = np.random.default_rng(8)
    n, c, h = 3, 64, 18
    x = rng.integers(0, 5_000_000, (n, c, h, h)).astype(np.int32)
    x[0, :, 3, 3] = 2 ** 24 + 1
    x[1, 1, 5, 5] = 2 ** 31 - 1
    x[2, 2] = rng.integers(2 ** 24, 2 ** 31 - 1, (h, h))
    want = O.maxpool_int(x, 3, 2, 1) if int_pool else O.maxpool_float_rt(x, 3, 2, 1)
    assert int_pool == bool((O.maxpool_int(x, 3, 2, 1) == want).all())      # the data tells them apart
    ho = want.shape[2]
    xd = G.dev(nchw_to_carry(x))
    a = _args_for_pool(n, c, h, ho, 3, 2, 1)
    a.flags = C.F8_OPF_INT_MAXPOOL if int_pool else 0
    a.in_ = xd.data_ptr()
    co = torch.empty((carry_elems(n, ho, ho, cpad(c)),), dtype=torch.int32, device="cuda:0")
    q0 = torch.empty((n, ho, ho, cpad(c)), dtype=torch.uint8, device="cuda:0")
    a.carry_out = co.data_ptr()
    a.out[0] = q0.data_ptr()
    a.out_shift[0], a.out_signed[0] = 15, 0
    C.check(f8lib.f8_maxpool3x3s2(ctypes.byref(a), torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    assert np.array_equal(carry_to_nchw(co.cpu().numpy(), n, c, ho, ho), want)
    assert np.array_equal(nhwc_to_nchw(q0.cpu().numpy(), c), O.requant(want, 0, 15, False))


@pytest.mark.parametrize("int_pool", [False, True], ids=["float_rt", "FXQMaxPool2d"])
@pytest.mark.parametrize("signed", [False, True])
def test_head_conv_pool_fused(G, f8lib, signed, int_pool):
    """x = head[:-1](x); x = head[-1](x.float()).int() in one launch (fix_resnet.py:355-362),
    incl. accumulators above 2^24 (float rounding) and at INT_MAX (x86 .int() indefinite)."""
    if not f8lib.f8_has_umma(0):
        pytest.skip("tcgen05 backend not available")
    rng = np.random.default_rng(21 + signed)
    n = 3
    lo, hi, w, b = _rand_layer(rng, 3, 64, 7, signed)
    b[5] = 2 ** 24 + 12345          # every value of the channel needs float32 rounding
    b[9] = 2 ** 31 - 1000           # float(x) == 2^31 for most pixels -> INT_MIN after .int()
    b[11] = -2 ** 31 + 5
    x = rng.integers(lo, hi, (n, 3, 224, 224)).astype(np.int32)
    acc = O.relu(O.conv2d(x, w, b, 2, 3, 1))
    pooled = O.maxpool_int(acc, 3, 2, 1) if int_pool else O.maxpool_float_rt(acc, 3, 2, 1)
    outs = ((14, False), (13, True))
    want_q = [O.requant(pooled, 0, s, g) for s, g in outs]
    xd = G.dev(G.nchw_to_nhwc8(x, 4, signed))
    wd = G.dev(G.pack(f8lib, C.F8_OP_CONV_DENSE, w, 4, 64))
    bd = G.dev(b)
    a = C.f8_conv_args()
    a.n, a.cin, a.cout, a.cin_pad, a.cout_pad = n, 3, 64, 4, 64
    a.kh, a.kw, a.stride, a.pad = 7, 7, 2, 3
    a.hin, a.win, a.hout, a.wout = 224, 224, 56, 56
    a.in_signed = int(signed)
    a.flags = C.F8_OPF_INT_MAXPOOL if int_pool else 0
    a.in_, a.wpack, a.bias = xd.data_ptr(), wd.data_ptr(), bd.data_ptr()
    co = torch.full((carry_elems(n, 56, 56, 64),), -7, dtype=torch.int32, device="cuda:0")
    q = [torch.full((n, 56, 56, 64), 0x77, dtype=torch.uint8, device="cuda:0") for _ in outs]
    a.carry_out = co.data_ptr()
    for j, (s_, g_) in enumerate(outs):
        a.out[j] = q[j].data_ptr()
        a.out_shift[j], a.out_signed[j] = s_, int(g_)
    C.check(f8lib.f8_head_pool(ctypes.byref(a), torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    assert np.array_equal(carry_to_nchw(co.cpu().numpy(), n, 64, 56, 56), pooled)
    assert np.array_equal(nhwc_to_nchw(q[0].cpu().numpy(), 64), want_q[0])
    assert np.array_equal(nhwc_to_nchw(q[1].cpu().numpy().view(np.int8), 64), want_q[1])


def test_pool_requant(G, f8lib):
    """FXQAvgPool2d int branch + classifier requant (fix_quant_ops.py:126-134)."""
    rng = np.random.default_rng(9)
    n, c = 5, 1280
    x = rng.integers(-2 ** 26, 2 ** 26, (n, c, 7, 7)).astype(np.int32)
    x[0, 0] = 2 ** 26                 # 49 * 2^26 wraps negative as int32
    want = O.avgpool_sum(np.minimum(x, 2 ** 26))
    xd = G.dev(nchw_to_carry(x))
    a = _args_for_pool(n, c, 7, 1, 7, 1, 0)
    a.in_ = xd.data_ptr()
    co = torch.empty((n, cpad(c)), dtype=torch.int32, device="cuda:0")
    q0 = torch.empty((n, cpad(c)), dtype=torch.uint8, device="cuda:0")
    a.carry_out = co.data_ptr()
    a.out[0] = q0.data_ptr()
    a.out_shift[0], a.out_signed[0] = 12, 0
    C.check(f8lib.f8_pool_requant(ctypes.byref(a), torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    assert np.array_equal(co.cpu().numpy()[:, :c], want)
    assert np.array_equal(q0.cpu().numpy()[:, :c].astype(np.int32), O.requant(want, 0, 12, False))


@pytest.mark.parametrize("shape", [(5, 512, 1000, False), (19, 1280, 1000, False), (3, 2048, 1000, True),
                                   (9, 48, 37, False)], ids=str)
def test_pool_fc_fused_tail(G, f8lib, shape):
    """FXQAvgPool2d + requant + classifier + .float() in one launch (fix_quant_ops.py:126-134,
    fix_resnet.py:367-383): ragged image counts (CTAs own 8 images), wrap of the pooled sum, both
    signednesses of the requantised vector, bias near INT_MAX (wrapping add, float rounding)."""
    n, c, classes, signed = shape
    rng = np.random.default_rng(n * c)
    x = rng.integers(-2 ** 20, 2 ** 22, (n, c, 7, 7)).astype(np.int32)
    x[0, 0] = 2 ** 26                                       # 49 * 2^26 wraps negative as int32
    w = rng.integers(-127, 128, (classes, c)).astype(np.int32)
    b = rng.integers(-2 ** 20, 2 ** 20, classes).astype(np.int32)
    b[1], b[2] = 2 ** 31 - 5, 2 ** 24 + 3
    shift = 15
    q = O.requant(O.avgpool_sum(x), 0, shift, signed)
    _, want = O.linear(q, w, b)
    a = _args_for_pool(n, c, 7, 1, 7, 1, 0)
    a.cout, a.cout_pad = classes, cpad(classes)
    xd = G.dev(nchw_to_carry(x))
    wd = G.dev(G.pack(f8lib, C.F8_OP_CONV_DENSE, w.reshape(classes, c, 1, 1), cpad(c), cpad(classes)))
    bd = G.dev(b)
    out = torch.full((n, classes), -1.0, dtype=torch.float32, device="cuda:0")
    a.in_, a.wpack, a.bias = xd.data_ptr(), wd.data_ptr(), bd.data_ptr()
    a.out_shift[0], a.out_signed[0] = shift, int(signed)
    a.out_f32, a.out_f32_ld = out.data_ptr(), classes
    C.check(f8lib.f8_pool_fc(ctypes.byref(a), torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    assert np.array_equal(out.cpu().numpy(), want)


def test_convert_input(G, f8lib):
    rng = np.random.default_rng(11)
    for signed in (False, True):
        lo, hi = (-127, 128) if signed else (0, 256)
        x = rng.integers(lo, hi, (3, 3, 20, 24)).astype(np.int32)
        xd = G.dev(x)
        out = torch.full((3, 20, 24, 4), 0x55, dtype=torch.uint8, device="cuda:0")
        C.check(f8lib.f8_convert_input(xd.data_ptr(), out.data_ptr(), 3, 20, 24,
                                       torch.cuda.current_stream().cuda_stream))
        got = out.cpu().numpy()
        if signed:
            got = got.view(np.int8)
        assert np.array_equal(got[..., :3].astype(np.int32), x.transpose(0, 2, 3, 1))
        assert not got[..., 3].any()


def test_errors_do_not_throw(G, f8lib):
    a = C.f8_conv_args()
    assert f8lib.f8_conv_dense(ctypes.byref(a), 0, None) == C.F8_ERR_ARG
    assert f8lib.f8_requant_i32(1, 1, 4, 0, 40, 0, None) == C.F8_ERR_UNSUPPORTED
    assert b"shift" in f8lib.f8_last_error()

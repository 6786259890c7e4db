"""The CPU oracle against (a) the committed golden vectors produced by the UNMODIFIED
reference (tests/golden/make_golden.py ran /root/reference through oracle/ref_harness.py)
and (b) a literal pure-Python restatement of the reference formulas on small cases."""
import os

import numpy as np
import pytest

from f8net_b200 import synth
from oracle import nets
from oracle import oracle as O

from util import checksum, qmaxpool_fixture, trained_fixture

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
ARCHS = list(synth.HEAD_SIGNED)


def _fixture(arch, family):
    hs = synth.HEAD_SIGNED[arch]
    if family == "calibrated":
        return (synth.make_state_dict(arch, hs), synth.make_input(arch, 2, hs),
                np.load(os.path.join(GOLD, f"{arch}_n2.npz")))
    return (synth.make_edge_state_dict(arch, hs), synth.make_input(arch, 2, hs, seed=777),
            np.load(os.path.join(GOLD, f"edge_{arch}_n2.npz")))


@pytest.mark.parametrize("family", ["calibrated", "edge"])
@pytest.mark.parametrize("arch", ARCHS)
def test_oracle_matches_reference_golden(arch, family):
    sd, x, gold = _fixture(arch, family)
    trace = {}
    y = nets.forward(arch, sd, x, synth.HEAD_SIGNED[arch], trace)
    # logits: float32 holding exact int32 values (fix_resnet.py:383)
    assert np.array_equal(y.astype(np.int64), gold["logits"].astype(np.int64))
    # every int layer's 8-bit input and int32 accumulator, pinned by checksum
    for name, want in zip(gold["layer_names"], gold["layer_checksums"]):
        assert checksum(trace[str(name)]) == want, f"{arch}/{family}: {name}"


@pytest.mark.parametrize("name", synth.TRAINED)
def test_oracle_matches_reference_on_trained_fraclens(name):
    """Second fixture family: the per-layer formats of the networks the reference's
    authors trained (fraclen_visual/*.out), golden logits / layer tensors from the unmodified reference
    (tests/golden/make_variant_golden.py)."""
    arch, hs, sd, x, gold = trained_fixture(name)
    fis = {int(sd[k][0]) for k in sd if k.endswith(".input_fraclen")}
    fws = {int(sd[k]) for k in sd if k.endswith(".weight_fraclen")}
    if name == "mobilenet_v2":
        assert {0, 1} <= fws and {1, 8} <= fis          # the fw in {0,1} layers SURVEY names
    trace = {}
    y = nets.forward(arch, sd, x, hs, trace)
    assert np.array_equal(y.astype(np.int64), gold["logits"].astype(np.int64))
    for lname, want in zip(gold["layer_names"], gold["layer_checksums"]):
        assert checksum(trace[str(lname)]) == want, f"{name}: {lname}"


@pytest.mark.parametrize("arch", ["resnet18", "resnet50"])
def test_oracle_matches_reference_with_both_head_pools(arch):
    """FLAGS.quant_maxpool: FXQMaxPool2d (integer max, fix_quant_ops.py:141-157) against
    nn.MaxPool2d on x.float() (fix_resnet.py:358-359); the fixture makes them disagree."""
    hs, sd, x, gold = qmaxpool_fixture(arch)
    yf = nets.forward(arch, sd, x, hs)
    ti = {}
    yi = nets.forward(arch, sd, x, hs, ti, quant_maxpool=True)
    assert np.array_equal(yf.astype(np.int64), gold["logits_float_pool"].astype(np.int64))
    assert np.array_equal(yi.astype(np.int64), gold["logits"].astype(np.int64))
    assert (yf != yi).mean() > 0.5
    for lname, want in zip(gold["layer_names"], gold["layer_checksums"]):
        assert checksum(ti[str(lname)]) == want, f"{arch}: {lname}"


# ---- literal restatement of fix_quant_ops.py:90-114 with Python ints -------------------
def _wrap32(v):
    v &= 0xFFFFFFFF
    return v - (1 << 32) if v >= (1 << 31) else v


def _ref_requant(x, fl, input_fl, signed):
    n = input_fl - fl
    if n > 0:
        res = _wrap32(x + (1 << (n - 1)))
        if x % (1 << n) == (1 << (n - 1)):          # Python floor-mod == torch.remainder
            res = _wrap32((res >> (n + 1)) << 1)
        else:
            res = res >> n
    else:
        res = _wrap32(x << (-n))
    return max(-127, min(127, res)) if signed else max(0, min(255, res))


def test_requant_against_python_restatement():
    rng = np.random.default_rng(0)
    special = [0, 1, -1, 2, 3, -2, -3, 127, 128, 255, 256, -127, -128, 2 ** 31 - 1, -2 ** 31,
               2 ** 31 - 2, -2 ** 31 + 1, 2 ** 24, 2 ** 24 + 1, -1610612736, 2 ** 29, 2 ** 30]
    xs = np.array(special + list(rng.integers(-2 ** 31, 2 ** 31, 400)) +
                  list(rng.integers(-70000, 70000, 600)), dtype=np.int64).astype(np.int32)
    for n in range(-9, 17):
        # ties: odd multiples of 2^(n-1)
        extra = np.array([(2 * k + 1) << (n - 1) for k in range(-40, 40)] if n > 0 else [0],
                         dtype=np.int64).astype(np.int32)
        v = np.concatenate([xs, extra])
        for signed in (False, True):
            fl = 7 if signed else 8
            got = O.requant(v, fl, fl + n, signed)
            want = np.array([_ref_requant(int(t), fl, fl + n, signed) for t in v], dtype=np.int32)
            assert np.array_equal(got, want), (n, signed)


def test_requant_survey_probes():
    """Facts probed on the reference (SURVEY.md 7 'Exactness traps')."""
    # ReLU-before-requant matters for left shifts: x = -1610612736, n = -1
    assert O.requant(np.array([-1610612736], np.int32), 8, 7, False)[0] == 255
    assert O.requant(np.array([0], np.int32), 8, 7, False)[0] == 0
    # 2^29 << 3 wraps to 0
    assert O.requant(np.array([2 ** 29], np.int32), 8, 5, False)[0] == 0
    # x + half wraps at INT_MAX -> negative -> clamps to -127 (signed)
    assert O.requant(np.array([2 ** 31 - 1], np.int32), 7, 11, True)[0] == -127
    # round half to even: 2.5 -> 2, 3.5 -> 4, -2.5 -> -2
    assert list(O.requant(np.array([5, 7, -5], np.int32), 7, 8, True)) == [2, 4, -2]


def test_residual_add_wrap_and_clamp():
    r = np.array([2 ** 31 - 1, -2 ** 31, 5, 2 ** 30], np.int32)
    s = np.array([1, 0, -2 ** 31, 2 ** 30], np.int32)
    out, fl = O.residual_add(r, s, 10, 10)
    # INT_MAX + 1 wraps to INT_MIN -> clamped to INT_MIN+1 ; INT_MIN stays clamped
    assert list(out) == [-2 ** 31 + 1, -2 ** 31 + 1, -2 ** 31 + 5, -2 ** 31 + 1] and fl == 10
    out, fl = O.residual_add(np.array([3], np.int32), np.array([2 ** 29], np.int32), 12, 9)
    assert out[0] == 3 and fl == 12          # (2^29 << 3) wraps to 0
    out, fl = O.residual_add(np.array([3], np.int32), np.array([5], np.int32), 9, 11)
    assert out[0] == 17 and fl == 11         # res <<= 2


def test_maxpool_float_round_trip():
    x = np.zeros((1, 1, 4, 4), np.int32)
    x[0, 0, 1, 1] = 2 ** 24 + 1              # not representable in float32 -> 2^24
    x[0, 0, 3, 3] = 7
    y = O.maxpool_float_rt(x, 3, 2, 1)
    assert y.shape == (1, 1, 2, 2)
    assert y[0, 0, 0, 0] == 2 ** 24 and y[0, 0, 1, 1] == 2 ** 24 and y[0, 0, 1, 0] == 2 ** 24
    big = np.full((1, 1, 2, 2), 2 ** 31 - 1, np.int32)   # float(INT_MAX) = 2^31 -> .int() indefinite
    assert O.maxpool_float_rt(big, 3, 2, 1)[0, 0, 0, 0] == -2 ** 31


def test_avgpool_sum_wraps_like_int64_then_int():
    x = np.full((1, 2, 7, 7), 2 ** 26, np.int32)          # 49 * 2^26 = 3288334336 < 2^32
    x[0, 1] = -5
    y = O.avgpool_sum(x)
    assert y[0, 0] == _wrap32(49 * 2 ** 26) and y[0, 1] == -245
    with pytest.raises(AssertionError):                   # fix_quant_ops.py:132
        O.avgpool_sum(np.full((1, 1, 7, 7), 2 ** 27, np.int32))


def test_conv_and_linear_against_numpy():
    rng = np.random.default_rng(3)
    for (c, o, k, st, pd, g, h) in [(3, 8, 7, 2, 3, 1, 20), (8, 8, 3, 1, 1, 8, 9), (8, 8, 3, 2, 1, 8, 10),
                                    (16, 24, 1, 2, 0, 1, 8), (5, 7, 3, 2, 1, 1, 11), (4, 6, 3, 1, 1, 1, 7)]:
        x = rng.integers(-127, 256, (2, c, h, h)).astype(np.int32)
        w = rng.integers(-127, 128, (o, c // g, k, k)).astype(np.int32)
        b = rng.integers(-2 ** 31, 2 ** 31, (o,)).astype(np.int32)
        got = O.conv2d(x, w, b, st, pd, g)
        ho = (h + 2 * pd - k) // st + 1
        xp = np.pad(x.astype(np.int64), ((0, 0), (0, 0), (pd, pd), (pd, pd)))
        want = np.zeros((2, o, ho, ho), np.int64)
        og, cg = o // g, c // g
        for oo in range(o):
            gi = oo // og
            for cc in range(cg):
                for r in range(k):
                    for s in range(k):
                        want[:, oo] += (xp[:, gi * cg + cc, r:r + st * ho:st, s:s + st * ho:st]
                                        * int(w[oo, cc, r, s]))
            want[:, oo] += int(b[oo])
        want = ((want + 2 ** 31) % 2 ** 32 - 2 ** 31).astype(np.int32)
        assert np.array_equal(got, want), (c, o, k, st, pd, g)
    q = rng.integers(0, 256, (3, 40)).astype(np.int32)
    w = rng.integers(-127, 128, (10, 40)).astype(np.int32)
    b = rng.integers(-2 ** 20, 2 ** 20, (10,)).astype(np.int32)
    yi, yf = O.linear(q, w, b)
    assert np.array_equal(yi, (q.astype(np.int64) @ w.T.astype(np.int64) + b).astype(np.int32))
    assert np.array_equal(yf, yi.astype(np.float32))


def test_input_integerisation():
    x = np.array([0.0, 0.5, 1.0, 0.0019607844, 0.00980392, 0.9999], np.float32)
    assert list(O.input_u8(x)) == [int(np.rint(np.float32(255.0) * v)) for v in x]
    xs = np.array([-3.0, -0.26, 0.0, 0.124, 0.126, 5.0], np.float32)
    got = O.input_s8(xs, 5)
    want = [int(max(-127, min(127, np.rint(v * 32)))) for v in xs]
    assert list(got) == want


# ---------------------------------------------------------------------------------------------
# forward_loss input preparation (A12) against vectors produced by the reference's own code
# (tests/golden/make_input_golden.py: fix_train.py:676-692 + models.fix_quant_ops.fix_quant)
# ---------------------------------------------------------------------------------------------
def _input_golden():
    import os
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "input_prep.npz"))


def test_input_prep_float_matches_reference():
    g = _input_golden()
    assert np.array_equal(O.input_u8(g["f32_u_x"]), g["f32_u_y"])
    for fl in (3, 5, 7):
        assert np.array_equal(O.input_s8(g[f"f32_s{fl}_x"], fl), g[f"f32_s{fl}_y"])


def test_input_prep_uint8_matches_reference():
    g = _input_golden()
    assert np.array_equal(O.image_prep_u8(g["u8_pix"][None], False), g["u8_u_y"])
    for fl in (3, 5, 7):
        assert np.array_equal(O.image_prep_u8(g["u8_pix"][None], True, fl), g[f"u8_s{fl}_y"])
    # normalize False is the identity on pixel values: (255 * (p / 255)).round() == p
    p = np.arange(256, dtype=np.uint8).reshape(1, 16, 16, 1).repeat(3, axis=3)
    assert np.array_equal(O.image_prep_u8(p, False)[0, 0].reshape(-1), np.arange(256))

"""Seeded synthetic workloads for the int_op_only path (bench.py and the tests).

No trained F8Net checkpoint is reachable offline (Google-Drive links only,
/root/reference/README.md:95), so weights are synthesised with trained-like statistics
(SURVEY.md 8(d)):  w_int = clamp(round(N(0, sigma_l)), -127, 127),  b_int ~ U{-2^14..2^14},
fw in {5,6,7} (mostly 7), and per-layer ``input_fraclen`` taken from a committed table
(f8net_b200/data/fraclens_<arch>.json) that tests/golden/make_golden.py calibrated so the
activations stay alive through the whole network.

The produced dict has exactly the reference IntModel.state_dict() layout:
``<p>.weight`` int32 [O,C/g,kh,kw] | [O,K], ``<p>.bias`` int32 [O], ``<p>.weight_fraclen``
int32 0-dim, ``<p>.input_fraclen`` int32 [1].
"""
import json
import math
import os

import numpy as np

from .arch import graph_for

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")

WEIGHT_SEED = 1234
INPUT_SEED = 1995    # the reference's random_seed (res18_fix_quant_test_int_op_only.yml:36)

# BASELINE.json configs: which nets run with a signed (normalize: True) head
HEAD_SIGNED = {"resnet18": False, "resnet50": True, "mobilenet_v1": False, "mobilenet_v2": False}


def weight_format_for(K, index, gain=1.0):
    """(sigma_int, fw) for a layer with fan-in K: He-scaled float weights sigma = gain*sqrt(2/K)
    (BN-folded trained nets keep the per-layer gain near 1; the conv that feeds a residual
    add gets a smaller gain so the carry does not blow up with depth) quantised with the
    reference's own rule fw = clamp(floor(log2(40 / sigma)), 0, 7) (metric2fraclen,
    fix_quant_ops.py:30-37), lowered by one on a few layers for variety."""
    sigma = gain * math.sqrt(2.0 / K)
    fw = int(min(7, max(0, math.floor(math.log2(40.0 / sigma)))))
    if index % 7 == 3 and fw > 0:
        fw -= 1
    return sigma * (1 << fw), fw


def load_fraclens(arch):
    path = os.path.join(_DATA, f"fraclens_{arch}.json")
    with open(path) as f:
        return json.load(f)


def bias_from(u, fw, fi):
    """b_int for a real-valued bias u/4 (u ~ U(-1,1)) at fraclen fw+fi -- what int_bias
    (fix_quant_ops.py:598-615) produces for a bias that is O(activation scale)."""
    return np.rint(u * 0.25 * float(1 << (fw + fi))).astype(np.int64).astype(np.int32)


def make_state_dict(arch, head_signed=None, seed=WEIGHT_SEED, input_fraclens=None,
                    default_fi=6, aux=None):
    """Synthetic reference-layout state dict (numpy int32 arrays, insertion-ordered).

    ``input_fraclens``: mapping prefix -> fi; defaults to the committed calibrated table,
    falling back to ``default_fi`` for prefixes the table lacks (used while calibrating).
    ``aux``: optional dict that receives prefix -> the uniform bias draws (calibration
    re-derives the bias once a layer's fi is chosen).
    """
    if head_signed is None:
        head_signed = HEAD_SIGNED.get(arch, False)
    net = graph_for(arch, head_signed)
    if input_fraclens is None:
        try:
            input_fraclens = load_fraclens(arch)
        except FileNotFoundError:
            input_fraclens = {}
    rng = np.random.default_rng(seed)
    residual_feeders = set()
    for blk in net.blocks:
        if blk.identity or blk.shortcut is not None:
            residual_feeders.add(blk.body[-1].prefix)
    sd = {}
    for idx, L in enumerate(net.convs()):
        shape = L.weight_shape()
        K = int(np.prod(shape[1:]))
        sigma_int, fw = weight_format_for(K, idx, 0.35 if L.prefix in residual_feeders else 1.0)
        w = np.rint(rng.normal(0.0, sigma_int, size=shape))
        w = np.clip(w, -127, 127).astype(np.int32)
        u = rng.uniform(-1.0, 1.0, size=(L.cout,))
        if aux is not None:
            aux[L.prefix] = u
        if L.prefix == "head.0":
            fi = (5 if head_signed else 8)          # weight_only head => 8 (fix_quant_ops.py:486-488)
        else:
            fi = int(input_fraclens.get(L.prefix, default_fi))
        b = bias_from(u, fw, fi)
        sd[L.prefix + ".weight"] = w
        sd[L.prefix + ".bias"] = b
        sd[L.prefix + ".weight_fraclen"] = np.array(fw, dtype=np.int32)
        sd[L.prefix + ".input_fraclen"] = np.array([fi], dtype=np.int32)
    return sd


def make_edge_state_dict(arch, head_signed=None):
    """Adversarial parameters (third fixture family): left-shift requants
    (fi > fa), accumulators pushed to the int32 limits by huge biases (wrap in the residual
    add, INT_MIN clamp), and even weights / odd-half biases that produce many exact ties."""
    if head_signed is None:
        head_signed = HEAD_SIGNED.get(arch, False)
    sd = make_state_dict(arch, head_signed, seed=4321)
    rng = np.random.default_rng(99)
    net = graph_for(arch, head_signed)
    convs = net.convs()
    # the conv feeding the 7x7 sum keeps sane biases: FXQAvgPool2d asserts sum <= 2^32-1
    # (fix_quant_ops.py:132) and the reference would raise instead of producing logits
    protected = {"classifier.0"}
    if net.tail is not None:
        protected.add(net.tail.prefix)
    for blk in net.blocks[-3:]:
        protected.update(c.prefix for c in blk.body)
        if blk.shortcut is not None:
            protected.add(blk.shortcut.prefix)
    for i, L in enumerate(convs):
        p = L.prefix
        b = sd[p + ".bias"]
        if i % 3 == 1 and p not in protected:
            big = rng.integers(-(1 << 31), (1 << 31), size=b.shape, dtype=np.int64)
            mask = rng.random(b.shape) < 0.25
            sd[p + ".bias"] = np.where(mask, big, b).astype(np.int32)
        if i % 5 == 2 and p != "head.0":
            sd[p + ".input_fraclen"] = np.array([7 if L.sym else 8], dtype=np.int32)
            sd[p + ".weight_fraclen"] = np.array(0, dtype=np.int32)
        if i % 4 == 3:
            sd[p + ".weight"] = (sd[p + ".weight"] // 2 * 2).astype(np.int32)
            if p not in protected:
                sd[p + ".bias"] = (sd[p + ".bias"] // 64 * 64 + 32).astype(np.int32)
    return sd


def make_input(arch, n, head_signed=None, seed=INPUT_SEED, size=224):
    """int32 NCHW [n,3,size,size] in the head's 8-bit range: 0..255 (what (255*x).round()
    yields, fix_train.py:691-692) or -127..127 for a signed head (fix_train.py:682-687)."""
    if head_signed is None:
        head_signed = HEAD_SIGNED.get(arch, False)
    rng = np.random.default_rng(seed)
    if head_signed:
        return rng.integers(-127, 128, size=(n, 3, size, size), dtype=np.int64).astype(np.int32)
    return rng.integers(0, 256, size=(n, 3, size, size), dtype=np.int64).astype(np.int32)


def to_torch_state_dict(sd):
    """numpy dict -> torch int32 tensors with the reference's exact shapes (for
    IntModel.load_state_dict)."""
    import torch
    return {k: torch.from_numpy(np.ascontiguousarray(v)).to(torch.int32).reshape(v.shape)
            for k, v in sd.items()}


def make_float_state_dict(arch, seed=77, flags=None):
    """A seeded synthetic FLOAT-simulation checkpoint (the ``best_model.pt`` layout of the
    reference's ``Model.state_dict()``: ``<p>.conv.weight``, ``<p>.bn.*``, ``<p>.alpha``,
    ``<p>.input_fraclen``; classifier ``weight`` / ``bias`` / ``alpha`` / ``input_fraclen``) for the
    export tests (f8net_b200/export.py).  Per-layer gains spread the folded weights over
    several weight fraclens; input fraclens carry fractional noise so the export's round()
    and clamp() are exercised."""
    import torch
    from .export import ExportFlags, float_layers
    flags = flags or ExportFlags()
    net = graph_for(arch, bool(flags.normalize))
    g = torch.Generator().manual_seed(seed)
    by_int = {c.prefix: c for c in net.convs()}
    sd = {}
    for i, L in enumerate(float_layers(net, flags)):
        c = by_int[L.iprefix]
        p = L.fprefix
        if L.kind == "fc":
            sd[p + ".weight"] = torch.randn(c.weight_shape(), generator=g) * (0.01 * (1 + i % 3))
            sd[p + ".bias"] = torch.randn(c.cout, generator=g) * 0.2
        else:
            shape = c.weight_shape()
            K = shape[1] * shape[2] * shape[3]
            gain = (0.3, 1.0, 3.0, 0.6, 8.0)[i % 5]
            sd[p + ".conv.weight"] = torch.randn(shape, generator=g) * (gain * math.sqrt(2.0 / K))
            sd[p + ".bn.weight"] = torch.rand(c.cout, generator=g) + 0.5
            sd[p + ".bn.bias"] = torch.randn(c.cout, generator=g) * 0.3
            sd[p + ".bn.running_mean"] = torch.randn(c.cout, generator=g) * 0.2
            sd[p + ".bn.running_var"] = torch.rand(c.cout, generator=g) * 1.5 + 0.5
        sd[p + ".alpha"] = torch.rand((), generator=g) * 8.0 + 2.0
        fl = float(torch.randint(3, 10, (1,), generator=g)) + float(torch.rand((), generator=g)) * 0.8 - 0.4
        sd[p + ".input_fraclen"] = torch.ones(1) * fl
    return sd


TRAINED = ("mobilenet_v2", "resnet50_ptcv", "resnet50_nvidia")


def load_trained_fraclens(name):
    """Per-layer (input_fraclen, weight_fraclen) of a network the reference's authors trained,
    parsed from the logs they ship (tools/parse_fraclen_logs.py): returns (arch, head_signed,
    {int prefix: (fi, fw)})."""
    with open(os.path.join(_DATA, f"trained_fraclens_{name}.json")) as f:
        doc = json.load(f)
    return doc["arch"], bool(doc["head_signed"]), {k: tuple(v) for k, v in doc["fraclens"].items()}


def make_trained_state_dict(name, seed=2468, gain=1.0):
    """Second fixture family of SURVEY.md 8(d): the TRAINED per-layer formats (fi 1..8, fw 0..7,
    including the fw in {0, 1} layers of MobileNetV2) with synthetic weights.  A layer's real-valued
    gain has to carry the activation range from its own input format to its consumer's
    (range ~ 2^(8 - fi)), so sigma_float = gain * sqrt(2 / K) * 2^(fi - fi_next) and
    w_int = clamp(round(N(0, sigma_float * 2^fw)), -127, 127); biases as in make_state_dict.
    Returns (arch, head_signed, state_dict)."""
    arch, head_signed, table = load_trained_fraclens(name)
    net = graph_for(arch, head_signed)
    rng = np.random.default_rng(seed)
    convs = net.convs()
    # consumer of every layer: the next body conv, or the first conv of the next block
    order = []
    for blk in net.blocks:
        order.append([c.prefix for c in blk.body])
    consumer = {}
    seq = [net.head.prefix] + [p for body in order for p in body]
    seq += [net.tail.prefix] if net.tail is not None else []
    seq += [net.fc.prefix]
    for a, b in zip(seq, seq[1:]):
        consumer[a] = b
    for i, blk in enumerate(net.blocks):
        if blk.shortcut is not None:
            consumer[blk.shortcut.prefix] = consumer[blk.body[-1].prefix]
    residual_feeders = {blk.body[-1].prefix for blk in net.blocks if blk.identity or blk.shortcut is not None}
    sd = {}
    for L in convs:
        fi, fw = table[L.prefix]
        shape = L.weight_shape()
        K = int(np.prod(shape[1:]))
        fi_next = table[consumer[L.prefix]][0] if L.prefix in consumer else fi
        sigma = gain * math.sqrt(2.0 / K) * 2.0 ** (fi - fi_next)
        if L.prefix in residual_feeders:
            sigma *= 0.5
        sigma_int = min(40.0, max(0.6, sigma * (1 << fw)))
        w = np.clip(np.rint(rng.normal(0.0, sigma_int, size=shape)), -127, 127).astype(np.int32)
        u = rng.uniform(-1.0, 1.0, size=(L.cout,))
        sd[L.prefix + ".weight"] = w
        sd[L.prefix + ".bias"] = bias_from(u, fw, fi)
        sd[L.prefix + ".weight_fraclen"] = np.array(fw, dtype=np.int32)
        sd[L.prefix + ".input_fraclen"] = np.array([fi], dtype=np.int32)
    return arch, head_signed, sd


def make_maxpool_state_dict(arch, head_signed=None):
    """Fixture that tells the two head max-pools apart (FLAGS.quant_maxpool, fix_resnet.py:355-359):
    the calibrated parameters with the head bias of eight channels pushed just under 2^31, so that
    many pooled values land in [2^31 - 64, 2^31) -- ``nn.MaxPool2d(x.float()).int()`` rounds those to
    2^31 and the x86 conversion returns INT_MIN (0 after the next requant), FXQMaxPool2d keeps them
    (255 after the next requant) -- and others lose low bits above 2^24 in the float round trip."""
    if head_signed is None:
        head_signed = HEAD_SIGNED.get(arch, False)
    sd = make_state_dict(arch, head_signed)
    b = sd["head.0.bias"].astype(np.int64)
    b[:8] = (1 << 31) - (25000 if head_signed else 60000) + 1500 * np.arange(8)
    b[8:12] = (1 << 24) + 12345 + 2 * np.arange(4)
    sd["head.0.bias"] = b.astype(np.int32)
    return sd

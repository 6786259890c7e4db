"""Offline export: float-simulation checkpoint -> IntModel
state_dict (f8net_b200/export.py) against the vectors the UNMODIFIED reference produced
(``Model.int_model()`` on the same seeded checkpoint; tests/golden/make_export_golden.py).

CPU tests: every exported tensor's SHA-256, every fraclen and every int bias equals the
reference's.  GPU test: a float checkpoint compiled straight into an Engine gives the oracle's
logits for the exported integers."""
import hashlib
import os
import sys

import numpy as np
import pytest

torch = pytest.importorskip("torch")

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))

from f8net_b200 import synth  # noqa: E402
from f8net_b200.arch import graph_for  # noqa: E402
from f8net_b200.export import ExportFlags, export_int_state_dict, float_layers  # noqa: E402
from make_export_golden import SEED, VARIANTS, flags_for  # noqa: E402

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
ARCHS = ["resnet18", "resnet50", "mobilenet_v1", "mobilenet_v2"]


def _sha(t):
    return hashlib.sha256(np.ascontiguousarray(t.numpy()).tobytes()).hexdigest()


@pytest.mark.parametrize("arch", ARCHS)
def test_export_equals_reference_int_model(arch):
    gold = np.load(os.path.join(GOLD, f"export_{arch}.npz"))
    assert int(gold["seed"]) == SEED
    flags = flags_for(arch)
    sd = export_int_state_dict(synth.make_float_state_dict(arch, SEED, flags), arch, flags)
    keys = [str(k) for k in gold["keys"]]
    assert list(sd.keys()) == keys                         # IntModel.state_dict() order
    for k, want in zip(keys, gold["sha256"]):
        t = sd[k]
        assert t.dtype == torch.int32, k
        if not k.endswith(".weight"):                      # small tensors are stored whole
            assert t.shape == gold[k].shape and np.array_equal(t.numpy(), gold[k]), k
        assert _sha(t) == str(want), k
    # state_dict layout of SURVEY.md 8(b)(1)
    assert sd["head.0.weight_fraclen"].dim() == 0 and tuple(sd["head.0.input_fraclen"].shape) == (1,)
    assert int(sd["head.0.weight"].abs().max()) <= 127


@pytest.mark.parametrize("arch,variant", [("resnet18", v) for v in VARIANTS] + [("mobilenet_v2", "sharing")])
def test_export_config_variants_equal_reference(arch, variant):
    """Config switches the shipped int_op_only yml files leave at their defaults -- input_fraclen_sharing
    (master layers share the buffer, fix_quant_ops.py:462-471), metric rms / mae (metric2fraclen
    coefficients, :30-37), conv weight rescaling (:306-316, 534-544), no classifier rescaling (:1046-1056) --
    against vectors the reference produced with the same FLAGS overrides."""
    gold = np.load(os.path.join(GOLD, f"export_{arch}_{variant}.npz"))
    flags = flags_for(arch, variant)
    sd = export_int_state_dict(synth.make_float_state_dict(arch, SEED, flags), arch, flags)
    keys = [str(k) for k in gold["keys"]]
    assert list(sd.keys()) == keys
    for k, want in zip(keys, gold["sha256"]):
        assert _sha(sd[k]) == str(want), (variant, k)


def test_wiring_master_and_following():
    """fix_resnet.py:148-153, 194-199, 456-467: a downsample block's body[0] and shortcut inherit
    the previous identity block's master; block-last convs and shortcuts are followed by the
    next block's body[0]; the last block by the classifier."""
    L = {l.fprefix: l for l in float_layers(graph_for("resnet18"), ExportFlags())}
    assert L["head.0"].weight_only and L["head.0"].following is L["stage_0_layer_0.body.0"]
    assert L["stage_0_layer_0.body.0"].master is None
    assert L["stage_0_layer_1.body.0"].master is L["stage_0_layer_0.body.0"]
    assert L["stage_1_layer_0.body.0"].master is L["stage_0_layer_1.body.0"]
    assert L["stage_1_layer_0.shortcut.0"].master is L["stage_0_layer_1.body.0"]
    assert L["stage_1_layer_1.body.0"].master is None
    assert L["stage_1_layer_0.shortcut.0"].following is L["stage_1_layer_1.body.0"]
    assert L["stage_3_layer_1.body.1"].following is L["classifier.0"]
    assert L["stage_3_layer_1.body.1"].avgpool_scale == 64 / 49
    assert L["stage_3_layer_1.body.0"].avgpool_scale == 1.0
    M = {l.fprefix: l for l in float_layers(graph_for("mobilenet_v2"), ExportFlags())}
    assert M["tail.0"].avgpool_scale == 64 / 49 and M["tail.0"].double_side
    assert M["stage_1_layer_1.body.0"].master is None and M["stage_1_layer_1.body.0"].double_side
    assert M["stage_2_layer_0.body.0"].master is M["stage_1_layer_1.body.0"]
    assert not M["stage_0_layer_0.body.0"].double_side
    V = {l.fprefix: l for l in float_layers(graph_for("mobilenet_v1"), ExportFlags())}
    assert all(l.master is None for l in V.values())
    assert V["stage_4_layer_1.body.1"].avgpool_scale == 64 / 49


def test_conv_constant_rescale_fails_like_the_reference():
    flags = ExportFlags(rescale_forward_conv=True, rescale_type="constant")
    with pytest.raises(NotImplementedError):
        export_int_state_dict(synth.make_float_state_dict("resnet18", SEED, flags), "resnet18", flags)


def test_rejects_an_int_state_dict():
    with pytest.raises(KeyError):
        export_int_state_dict(synth.to_torch_state_dict(synth.make_state_dict("resnet18")), "resnet18")


@pytest.mark.gpu
@pytest.mark.parametrize("arch", ["resnet18", "mobilenet_v2"])
def test_float_checkpoint_to_engine(cuda, f8lib, arch):
    from f8net_b200.export import compile_float
    from oracle import nets
    flags = flags_for(arch)
    fsd = synth.make_float_state_dict(arch, SEED, flags)
    eng = compile_float(fsd, arch, flags, chunk=2)
    x = synth.make_input(arch, 3, bool(flags.normalize), seed=5)
    sd_np = {k: v.numpy() for k, v in export_int_state_dict(fsd, arch, flags).items()}
    want = nets.forward(arch, sd_np, x, bool(flags.normalize))
    got = eng(torch.from_numpy(x).cuda()).cpu().numpy()
    assert np.array_equal(got, want)

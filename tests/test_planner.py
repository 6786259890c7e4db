"""Host logic (no GPU): the integer graph, the fused launch plan and its buffer table."""
import numpy as np
import pytest

from f8net_b200 import _capi as C
from f8net_b200 import synth
from f8net_b200.arch import graph_for
from f8net_b200.engine import infer_arch
from f8net_b200.planner import build_plan, cpad

ARCHS = list(synth.HEAD_SIGNED)
# int layers per model, state_dict tensors = 4 per layer
N_LAYERS = {"resnet18": 21, "resnet50": 54, "mobilenet_v1": 28, "mobilenet_v2": 53}
MMACS = {"resnet18": 1814.1, "resnet50": 4089.2, "mobilenet_v1": 568.7, "mobilenet_v2": 300.8}


@pytest.mark.parametrize("arch", ARCHS)
def test_graph_matches_survey_inventory(arch):
    net = graph_for(arch, synth.HEAD_SIGNED[arch])
    convs = net.convs()
    assert len(convs) == N_LAYERS[arch]
    sd = synth.make_state_dict(arch)
    assert len(sd) == 4 * N_LAYERS[arch]
    assert infer_arch(sd) == arch
    # MACs per image from the graph geometry
    P = build_plan(net, sd)
    macs = 0
    for op in P.ops:
        if op.kind == C.F8_OP_CONV_DENSE:
            macs += op.hout * op.wout * op.cout * op.cin * op.k * op.k
        elif op.kind == C.F8_OP_CONV_DW:
            macs += op.hout * op.wout * op.cout * 9
    assert abs(macs / 1e6 - MMACS[arch]) < 0.06


@pytest.mark.parametrize("arch", ARCHS)
def test_plan_structure(arch):
    net = graph_for(arch, synth.HEAD_SIGNED[arch])
    sd = synth.make_state_dict(arch)
    P = build_plan(net, sd)
    kinds = [op.kind for op in P.ops]
    assert kinds[0] == C.F8_OP_CONVERT_INPUT and kinds[-1] == C.F8_OP_CONV_DENSE
    assert kinds[-2] == C.F8_OP_POOL_REQUANT and P.ops[-1].out_f32 == 1
    assert kinds.count(C.F8_OP_MAXPOOL) == (1 if net.maxpool else 0)
    # one launch per int layer + convert + pool (+ maxpool): no standalone requant / ReLU /
    # residual kernels
    assert len(P.ops) == N_LAYERS[arch] + 2 + (1 if net.maxpool else 0)
    written = set()
    for op in P.ops:
        for b in (op.in_buf, op.carry_in_buf):
            if b >= 0:
                assert b in written, f"{op.name} reads buffer {b} before it is written"
        for b in [op.carry_out_buf] + [o[0] for o in op.outs]:
            if b >= 0:
                assert b not in written, "buffers are single-assignment"
                written.add(b)
    # every residual block's last conv reads a carry
    n_res = sum(1 for b in net.blocks if b.identity or b.shortcut is not None)
    assert sum(1 for op in P.ops if op.carry_in_buf >= 0) == n_res


@pytest.mark.parametrize("arch", ARCHS)
def test_buffers_never_overlap_while_live(arch):
    net = graph_for(arch, synth.HEAD_SIGNED[arch])
    P = build_plan(net, synth.make_state_dict(arch))
    bufs = P.bufs
    for i, a in enumerate(bufs):
        assert a.offset % 256 == 0 and a.offset + a.bytes_per_image <= P.workspace_per_image
        for b in bufs[i + 1:]:
            live = not (a.last < b.first or b.last < a.first)
            overlap = not (a.offset + a.bytes_per_image <= b.offset or
                           b.offset + b.bytes_per_image <= a.offset)
            assert not (live and overlap), (a.name, b.name)
    # reuse must actually happen: the workspace is far below the sum of all buffers
    assert P.workspace_per_image < 0.6 * sum(b.bytes_per_image for b in bufs)


def test_shifts_and_signedness_follow_the_reference_rule():
    arch = "resnet50"
    net = graph_for(arch, True)
    sd = synth.make_state_dict(arch, True)
    P = build_plan(net, sd)
    by_name = {op.name: op for op in P.ops}
    fi = lambda p: int(sd[p + ".input_fraclen"][0])
    fw = lambda p: int(sd[p + ".weight_fraclen"])
    # first block: maxpool output feeds body.0 and shortcut.0 (possibly different fi)
    mp = by_name["head.maxpool"]
    fa_head = fw("head.0") + fi("head.0")
    want = {(fa_head - fi("stage_0_layer_0.body.0"), 0), (fa_head - fi("stage_0_layer_0.shortcut.0"), 0)}
    assert {(s, g) for _, s, g in mp.outs} == want
    # residual shift of the first block: fa(body.4) - fa(shortcut)
    last = by_name["stage_0_layer_0.body.4"]
    fa_r = fw("stage_0_layer_0.body.4") + fi("stage_0_layer_0.body.4")
    fa_s = fw("stage_0_layer_0.shortcut.0") + fi("stage_0_layer_0.shortcut.0")
    assert last.carry_shift == fa_r - fa_s and last.relu == 1
    assert last.carry_in_buf == by_name["stage_0_layer_0.shortcut.0"].carry_out_buf
    # second block is an identity block: reads the int32 carry of the first block's output
    assert by_name["stage_0_layer_1.body.4"].carry_in_buf == last.carry_out_buf
    # classifier requant: avgpool adds 6 fractional bits
    pool = by_name["avgpool"]
    assert pool.outs[0][1] == pool.fa - fi("classifier.0") and pool.fa == by_name["stage_3_layer_2.body.4"].fa + 6


def test_mbv2_signed_block_entries():
    net = graph_for("mobilenet_v2")
    P = build_plan(net, synth.make_state_dict("mobilenet_v2"))
    by_name = {op.name: op for op in P.ops}
    assert by_name["stage_0_layer_0.body.0"].in_signed == 0       # stage 0 entry is unsigned
    assert by_name["stage_1_layer_0.body.0"].in_signed == 1
    assert by_name["tail.0"].in_signed == 1 and by_name["tail.0"].relu == 1
    # linear bottleneck: no ReLU on the projection, identity add without ReLU
    proj = by_name["stage_2_layer_1.body.4"]
    assert proj.relu == 0 and proj.carry_in_buf >= 0
    assert cpad(24) == 32 and cpad(144) == 144 and cpad(1000) == 1008


def test_descriptor_round_trip():
    net = graph_for("resnet18")
    P = build_plan(net, synth.make_state_dict("resnet18"))
    d, keep = P.to_desc()
    assert d.n_ops == len(P.ops) and d.n_buffers == len(P.bufs)
    assert d.ops[1].kind == C.F8_OP_CONV_DENSE and d.ops[1].cin_pad == 4 and d.ops[1].kh == 7
    assert d.ops[1].weight and d.ops[1].bias
    assert d.workspace_per_image == P.workspace_per_image and d.num_classes == 1000


def test_planner_rejects_wrong_shapes():
    sd = synth.make_state_dict("resnet18")
    sd["stage_1_layer_0.body.0.weight"] = np.zeros((128, 64, 1, 1), np.int32)
    with pytest.raises(ValueError, match="stage_1_layer_0.body.0.weight"):
        build_plan(graph_for("resnet18"), sd)


def test_fused_tail_plan_keeps_the_algorithmic_work():
    """fuse_tail folds FXQAvgPool2d + requant + classifier into one POOL_FC launch; the roofline
    accounting must not change."""
    from f8net_b200.roofline import network_work, op_work
    for arch in ("resnet18", "mobilenet_v2"):
        hs = synth.HEAD_SIGNED[arch]
        sd = synth.make_state_dict(arch, hs)
        net = graph_for(arch, hs)
        a = build_plan(net, sd, fuse_head=True)
        b = build_plan(net, sd, fuse_head=True, fuse_tail=True)
        assert len(b.ops) == len(a.ops) - 1 and b.ops[-1].kind == C.F8_OP_POOL_FC and b.ops[-1].out_f32 == 1
        assert b.ops[-1].outs[0][1:] == a.ops[-2].outs[0][1:]          # same requant of the pooled sum
        wa, wb = op_work(a), op_work(b)
        assert sum(r["bytes_per_image"] for r in wa) == sum(r["bytes_per_image"] for r in wb)
        assert sum(r["ops"] for r in wa) == sum(r["ops"] for r in wb)
        assert network_work(net)[1] == sum(r["bytes_per_image"] for r in wb)


def test_roofline_accounting_reproduces_the_survey_table():
    """the per-image algorithmic work `roofline.achieved` must be computed from -- MACs, algorithmic
    HBM bytes (8-bit activations once in / once out per layer + the int32 residual carries) and int8 weight bytes of
    the four networks, probed from the reference's own layer shapes."""
    from f8net_b200.roofline import network_work
    table = {   # arch: (MACs/img in M, algorithmic MB/img, weights MB)
        "resnet18": (1814.1, 10.69, 11.68), "resnet50": (4089.2, 65.93, 25.50),
        "mobilenet_v1": (568.7, 10.19, 4.21), "mobilenet_v2": (300.8, 15.18, 3.47)}
    for arch, (macs_m, mb, w_mb) in table.items():
        ops, nbytes, wbytes = network_work(graph_for(arch, synth.HEAD_SIGNED[arch]))
        assert abs(ops / 2 / 1e6 - macs_m) < 0.06, (arch, ops)
        assert abs(nbytes / 1e6 - mb) < 0.006, (arch, nbytes)
        assert abs(wbytes / 1e6 - w_mb) < 0.006, (arch, wbytes)

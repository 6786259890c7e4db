"""ONNX export of the integer graph (the reference's onnx_export call for the
int_op_only model, myutils/export.py:4-31 / fix_train.py:948-954): the file is decoded again and the
decoded graph re-executed node by node (tests/onnx_eval.py, Conv / Gemm through the CPU oracle); its
logits must equal the golden logits the unmodified reference produced for the same parameters."""
import os

import numpy as np
import pytest

from f8net_b200 import synth
from f8net_b200.onnx_export import FLOAT, INT32, export_onnx, read_model
from oracle import nets
from util import qmaxpool_fixture, trained_fixture

import onnx_eval

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("arch", list(synth.HEAD_SIGNED))
def test_exported_graph_reproduces_the_reference_logits(tmp_path, arch):
    hs = synth.HEAD_SIGNED[arch]
    sd, x = synth.make_state_dict(arch, hs), synth.make_input(arch, 2, hs)
    path = str(tmp_path / f"{arch}.onnx")
    nbytes = export_onnx(sd, path, arch=arch, head_signed=hs)
    assert nbytes == os.path.getsize(path)
    m = read_model(path)
    # the reference's export surface: opset 11, input / output names, dynamic batch axis, int32 in, float out
    assert m["opset"] == 11
    assert m["inputs"] == [{"name": "input", "elem_type": INT32, "dims": ["batch_size", 3, 224, 224]}]
    assert m["outputs"] == [{"name": "output", "elem_type": FLOAT, "dims": ["batch_size", 1000]}]
    # every parameter is an int32 initialiser under its state_dict key, bit for bit
    for k, v in sd.items():
        if k.endswith((".weight", ".bias")):
            assert m["initializers"][k].dtype == np.int32 and np.array_equal(m["initializers"][k], v)
    ops = [n["op"] for n in m["nodes"]]
    assert ops.count("Conv") + ops.count("Gemm") == len(sd) // 4 and ops.count("Gemm") == 1
    y = onnx_eval.run(m, x)
    gold = np.load(os.path.join(GOLD, f"{arch}_n2.npz"))["logits"]
    assert y.dtype == np.float32 and np.array_equal(y.astype(np.int64), gold.astype(np.int64))


def test_exported_graph_edge_family_wraps_like_the_reference(tmp_path):
    """Adversarial parameters: left-shift requants, INT32 wrap in the residual add, ties."""
    arch = "resnet18"
    hs = synth.HEAD_SIGNED[arch]
    sd, x = synth.make_edge_state_dict(arch, hs), synth.make_input(arch, 2, hs, seed=777)
    path = str(tmp_path / "edge.onnx")
    export_onnx(sd, path, arch=arch, head_signed=hs)
    y = onnx_eval.run(read_model(path), x)
    gold = np.load(os.path.join(GOLD, f"edge_{arch}_n2.npz"))["logits"]
    assert np.array_equal(y.astype(np.int64), gold.astype(np.int64))


def test_exported_graph_trained_fraclens_and_arch_inference(tmp_path):
    arch, hs, sd, x, gold = trained_fixture("mobilenet_v2")
    path = str(tmp_path / "t.onnx")
    export_onnx(sd, path)                         # architecture inferred from the key set
    y = onnx_eval.run(read_model(path), x)
    assert np.array_equal(y.astype(np.int64), gold["logits"].astype(np.int64))


def test_exported_graph_head_pool_variants(tmp_path):
    """quant_maxpool False: Cast(float) -> MaxPool -> Cast(int32); True (FXQMaxPool2d): integer MaxPool."""
    hs, sd, x, gold = qmaxpool_fixture("resnet18")
    pf, pi = str(tmp_path / "f.onnx"), str(tmp_path / "i.onnx")
    export_onnx(sd, pf, arch="resnet18", head_signed=hs)
    export_onnx(sd, pi, arch="resnet18", head_signed=hs, quant_maxpool=True)
    mf, mi = read_model(pf), read_model(pi)
    assert [n["op"] for n in mf["nodes"]].count("Cast") == [n["op"] for n in mi["nodes"]].count("Cast") + 2
    assert np.array_equal(onnx_eval.run(mf, x).astype(np.int64), gold["logits_float_pool"].astype(np.int64))
    assert np.array_equal(onnx_eval.run(mi, x).astype(np.int64), gold["logits"].astype(np.int64))


def test_reader_matches_protobuf_runtime_if_present(tmp_path):
    """The hand-written wire encoder against google.protobuf's generic decoder (no onnx schema needed):
    every top-level field parses and the graph field round-trips byte for byte."""
    pb = pytest.importorskip("google.protobuf.internal.decoder")
    sd = synth.make_state_dict("mobilenet_v1")
    path = str(tmp_path / "m.onnx")
    export_onnx(sd, path, arch="mobilenet_v1")
    data = open(path, "rb").read()
    pos, seen = 0, []
    while pos < len(data):
        tag, pos = pb._DecodeVarint(data, pos)
        field, wt = tag >> 3, tag & 7
        if wt == 0:
            _, pos = pb._DecodeVarint(data, pos)
        else:
            assert wt == 2
            ln, pos = pb._DecodeVarint(data, pos)
            pos += ln
        seen.append(field)
    assert pos == len(data) and seen == [1, 2, 3, 7, 8]

"""Serialised engine file: int8 weights + int32 biases + formats + graph meta in
one .npz; loading gives back the reference-layout state_dict bit for bit, and (GPU) an Engine whose
logits equal the golden logits of the unmodified reference."""
import os

import numpy as np
import pytest

import f8net_b200
from f8net_b200 import engine_file, synth

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("arch", list(synth.HEAD_SIGNED))
def test_engine_file_round_trip(tmp_path, arch):
    hs = synth.HEAD_SIGNED[arch]
    sd = synth.make_edge_state_dict(arch, hs)          # huge biases, left-shift formats
    path = str(tmp_path / f"{arch}.f8e")
    meta = f8net_b200.save_engine(path, lambda: sd, head_signed=hs)      # bound-method style, arch inferred
    assert meta["arch"] == arch and meta["head_signed"] == hs
    m2, back = engine_file.load_state_dict(path)
    assert m2 == meta and list(back.keys()) == list(sd.keys())
    for k, v in sd.items():
        assert back[k].dtype == np.int32 and back[k].shape == v.shape and np.array_equal(back[k], v), k
    # a quarter of the int32 checkpoint, roughly (weights dominate)
    raw = sum(v.nbytes for v in sd.values())
    assert os.path.getsize(path) < 0.3 * raw


def test_engine_file_rejects_non_8bit_weights_and_foreign_files(tmp_path):
    sd = synth.make_state_dict("mobilenet_v1")
    sd["head.0.weight"] = sd["head.0.weight"].copy()
    sd["head.0.weight"][0, 0, 0, 0] = 128
    with pytest.raises(ValueError, match="8-bit range"):
        f8net_b200.save_engine(str(tmp_path / "x.f8e"), sd)
    other = str(tmp_path / "other.npz")
    np.savez(other, a=np.zeros(3))
    with pytest.raises(ValueError, match="not an f8net-b200 engine file"):
        engine_file.load_state_dict(other)


@pytest.mark.gpu
@pytest.mark.parametrize("arch", ["resnet18", "mobilenet_v2"])
def test_engine_file_loads_into_an_engine(cuda, f8lib, tmp_path, arch):
    import torch
    hs = synth.HEAD_SIGNED[arch]
    path = str(tmp_path / "m.f8e")
    f8net_b200.save_engine(path, synth.make_state_dict(arch, hs), arch=arch, head_signed=hs)
    eng = f8net_b200.load_engine(path)
    y = eng(torch.from_numpy(synth.make_input(arch, 2, hs)))
    gold = np.load(os.path.join(GOLD, f"{arch}_n2.npz"))["logits"]
    assert np.array_equal(y.numpy().astype(np.int64), gold.astype(np.int64))


@pytest.mark.gpu
def test_engine_file_keeps_the_head_pool_variant(cuda, f8lib, tmp_path):
    import torch
    from util import qmaxpool_fixture
    hs, sd, x, gold = qmaxpool_fixture("resnet18")
    path = str(tmp_path / "q.f8e")
    f8net_b200.save_engine(path, sd, arch="resnet18", head_signed=hs, quant_maxpool=True)
    y = f8net_b200.load_engine(path)(torch.from_numpy(x))
    assert np.array_equal(y.numpy().astype(np.int64), gold["logits"].astype(np.int64))

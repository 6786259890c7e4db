"""Per-layer parity INSIDE the whole network on the GPU (logits and every captured
layer tensor).  The engine is compiled with keep_buffers=True (every plan buffer gets its own
workspace range) and, after one pass, every buffer is copied out through f8_plan_read_buffer:

 * the 8-bit input image of every int layer  == the tensor the UNMODIFIED reference fed that layer
   (forward-pre-hook capture, pinned by the committed checksums in tests/golden/*.npz);
 * every int32 tensor the engine keeps (residual carries, shortcut outputs, head / tail outputs, the
   average pool's input) == the CPU oracle's trace of the same run.
"""
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")

import f8net_b200  # noqa: E402
from f8net_b200 import _capi as C  # noqa: E402
from f8net_b200 import synth  # noqa: E402
from oracle import nets  # noqa: E402
from util import carry_to_nchw, checksum, qmaxpool_fixture, trained_fixture  # noqa: E402

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _cases():
    out = []
    for arch in synth.HEAD_SIGNED:
        out.append(pytest.param(("calibrated", arch), id=f"calibrated-{arch}"))
        out.append(pytest.param(("edge", arch), id=f"edge-{arch}"))
    for name in synth.TRAINED:
        out.append(pytest.param(("trained", name), id=f"trained-{name}"))
    out.append(pytest.param(("qmaxpool", "resnet18"), id="qmaxpool-resnet18"))
    return out


def _load(case):
    family, name = case
    qmp = False
    if family == "calibrated":
        arch, hs = name, synth.HEAD_SIGNED[name]
        sd, x = synth.make_state_dict(arch, hs), synth.make_input(arch, 2, hs)
        gold = np.load(os.path.join(GOLD, f"{arch}_n2.npz"))
    elif family == "edge":
        arch, hs = name, synth.HEAD_SIGNED[name]
        sd, x = synth.make_edge_state_dict(arch, hs), synth.make_input(arch, 2, hs, seed=777)
        gold = np.load(os.path.join(GOLD, f"edge_{arch}_n2.npz"))
    elif family == "trained":
        arch, hs, sd, x, gold = trained_fixture(name)
    else:
        arch = name
        hs, sd, x, gold = qmaxpool_fixture(arch)
        qmp = True
    return arch, hs, sd, x, gold, qmp


@pytest.mark.parametrize("backend", [pytest.param(0, id="imma"), pytest.param(1, id="tcgen05")])
@pytest.mark.parametrize("case", _cases())
def test_every_layer_tensor_inside_the_network(cuda, f8lib, case, backend):
    arch, hs, sd, x, gold, qmp = _load(case)
    n = x.shape[0]
    eng = f8net_b200.compile(sd, arch=arch, head_signed=hs, backend=backend, chunk=n, keep_buffers=True,
                             quant_maxpool=qmp)
    y = eng(torch.from_numpy(x).cuda()).cpu().numpy()
    assert np.array_equal(y.astype(np.int64), gold["logits"].astype(np.int64))
    trace = {}
    nets.forward(arch, sd, x, hs, trace, quant_maxpool=qmp)
    want8 = dict(zip((str(s) for s in gold["layer_names"]), gold["layer_checksums"]))
    last_of_block = {blk.body[-1].prefix: blk.name for blk in eng.net.blocks}
    checked8 = checked32 = 0
    for op in eng.plan.ops:
        # ---- the 8-bit image this layer reads: the reference's forward-pre-hook capture ----
        if op.kind in (C.F8_OP_CONV_DENSE, C.F8_OP_CONV_DW) and op.in_buf >= 0 and op.name + ":in8" in want8 \
                and op.name != "head.0":
            raw = eng.read_buffer(op.in_buf).reshape(n, op.hin, op.win, op.cin_pad)
            img = raw.view(np.int8) if op.in_signed else raw
            assert not img[..., op.cin:].any(), f"{op.name}: padded input channels are not zero"
            nchw = np.ascontiguousarray(img[..., :op.cin].transpose(0, 3, 1, 2)).astype(np.int32)
            assert checksum(nchw) == want8[op.name + ":in8"], f"{case}: 8-bit input of {op.name}"
            assert np.array_equal(nchw, trace[op.name + ":in8"].reshape(nchw.shape))
            checked8 += 1
        # ---- the int32 tensor this launch keeps ----
        if op.carry_out_buf >= 0:
            if op.name.endswith(".shortcut.0"):
                key = op.name + ":acc"
            elif op.name in ("head.0+maxpool", "head.maxpool"):
                key = "head:pool"
            elif op.name in ("head.0", "tail.0"):
                key = op.name + ":acc"                       # traced after the in-place ReLU
            else:
                key = last_of_block[op.name] + ":out"
            got = carry_to_nchw(eng.read_buffer(op.carry_out_buf).view(np.int32), n, op.cout, op.hout, op.wout,
                                op.cout_pad)
            assert np.array_equal(got, trace[key].reshape(got.shape)), f"{case}: int32 output of {op.name} ({key})"
            checked32 += 1
    n_layers = len(eng.net.convs())
    assert checked8 >= n_layers - 2, (checked8, n_layers)       # all but the head and the fused classifier
    assert checked32 >= 1

"""The "loads unchanged" surface against REAL reference modules: builds
the reference's own IntModel (Model.int_model(), fix_resnet.py:526-544, fix_mobilenet_v1.py:262-281,
fix_mobilenet_v2.py:405-423) through oracle/ref_harness.py and checks that

 * arch.graph_from_module(int_model) equals arch.graph_for(arch) -- stride / padding / groups /
   input_symmetric / block wiring read from the live module tree agree with the tables the engine
   uses when it is handed a bare state_dict;
 * the planner lowers both to the same launch list;
 * FLAGS.quant_maxpool (FXQMaxPool2d head) is detected, and an IntModel without FXQAvgPool2d
   (quant_avgpool off: float average pool) is rejected instead of being computed differently;
 * the unmodified reference, run here, still produces the committed golden logits.

Authoring tier only: skipped where /root/reference does not exist (the GPU box).  One subprocess
per case: the reference keeps one global config per process (myutils/config.py:152-178)."""
import os
import subprocess
import sys

import pytest

from oracle import ref_harness

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.skipif(not ref_harness.available(), reason="/root/reference not present")

CASE = r"""
import sys, numpy as np
sys.path.insert(0, {root!r})
from oracle import ref_harness
from f8net_b200 import synth
from f8net_b200.arch import graph_for, graph_from_module
from f8net_b200.planner import build_plan
arch, overrides, mode = {arch!r}, {overrides!r}, {mode!r}
im, FLAGS = ref_harness.build_int_model(arch, flag_overrides=overrides)
hs = synth.HEAD_SIGNED[arch]
assert bool(getattr(FLAGS, 'normalize', False)) == hs
if mode == 'reject':
    try:
        graph_from_module(im)
    except ValueError as e:
        assert 'FXQAvgPool2d' in str(e)
        print('CASE-OK')
        sys.exit(0)
    raise SystemExit('an IntModel with a float average pool was accepted')
net = graph_from_module(im)
want = graph_for(arch, hs, quant_maxpool=bool(overrides.get('quant_maxpool', False)))
assert net == want, 'graph_from_module(IntModel) != graph_for(arch)'
assert net.maxpool_int == bool(overrides.get('quant_maxpool', False))
# the state_dict layout the engine is documented to accept (SURVEY.md 8(b)(1))
ref_sd = im.state_dict()
sd = synth.make_state_dict(arch, hs)
assert list(ref_sd.keys()) == list(sd.keys())
for k, v in ref_sd.items():
    assert tuple(v.shape) == tuple(sd[k].shape) and str(v.dtype) == 'torch.int32', k
# same launch list from the module walk and from the architecture table
a = build_plan(net, sd, fuse_head=True, fuse_tail=True)
b = build_plan(want, sd, fuse_head=True, fuse_tail=True)
key = lambda op: (op.kind, op.name, op.cin, op.cout, op.k, op.stride, op.pad, op.in_signed, op.relu,
                  op.carry_shift, tuple(op.outs), op.in_buf, op.carry_in_buf, op.carry_out_buf, op.flags)
assert [key(o) for o in a.ops] == [key(o) for o in b.ops]
if mode == 'golden':
    # the unmodified reference on the committed fixture reproduces the committed logits
    import torch
    im.load_state_dict({{k: torch.from_numpy(np.ascontiguousarray(v)).to(torch.int32).reshape(ref_sd[k].shape)
                        for k, v in sd.items()}})
    x = torch.from_numpy(synth.make_input(arch, 2, hs))
    x.output_fraclen = int(im.head[0].input_fraclen.item()) if hs else 8
    with torch.no_grad():
        y = im(x).numpy()
    gold = np.load({root!r} + '/tests/golden/' + arch + '_n2.npz')['logits']
    assert np.array_equal(y.astype(np.int64), gold.astype(np.int64)), 'reference logits != committed golden'
print('CASE-OK')
"""


def _run(arch, overrides=None, mode="graph"):
    code = CASE.format(root=ROOT, arch=arch, overrides=overrides or {}, mode=mode)
    env = dict(os.environ, PYTHONDONTWRITEBYTECODE="1")
    r = subprocess.run([sys.executable, "-W", "ignore", "-c", code], cwd=ROOT, env=env,
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "CASE-OK" in r.stdout, (r.stdout[-2000:], r.stderr[-3000:])


@pytest.mark.parametrize("arch", ["resnet18", "resnet50", "mobilenet_v1", "mobilenet_v2"])
def test_real_int_model_walk_equals_architecture_table_and_golden(arch):
    _run(arch, mode="golden")


def test_real_int_model_with_fxq_maxpool_is_detected():
    _run("resnet18", {"quant_maxpool": True})


def test_real_int_model_with_float_avgpool_is_rejected():
    _run("resnet18", {"quant_avgpool": False}, mode="reject")

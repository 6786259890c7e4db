"""Generates the committed fixtures.  Run ONLY in the authoring container (needs
/root/reference):

    python tests/golden/make_golden.py [arch ...]

For each architecture it
 1. calibrates per-layer input_fraclen on the seeded synthetic weights (oracle forward with
    the CALIB hook: largest fi whose 8-bit saturation rate is <= 3 %) and writes
    f8net_b200/data/fraclens_<arch>.json;
 2. runs the UNMODIFIED reference IntModel (oracle/ref_harness.py, subprocess: the reference
    keeps one global config per process) on N=2 seeded inputs;
 3. checks the C oracle against the reference on the logits and on every layer's 8-bit
    input and int32 accumulator, and
 4. writes tests/golden/<arch>_n2.npz: reference logits + a position-weighted checksum of
    every captured tensor (the full tensors are tens of MB; the checksums pin them).

A second, adversarial fixture (tests/golden/edge_<arch>.npz, same recipe) uses fraclens that
force left shifts, huge biases (accumulators near INT32 limits so wrap / clamp paths fire)
and ties.
"""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from f8net_b200 import synth  # noqa: E402
from f8net_b200.arch import graph_for  # noqa: E402
from oracle import nets  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
DATA = os.path.join(ROOT, "f8net_b200", "data")
N_GOLD = 2


def checksum(a):
    """Position-weighted checksum mod 2^64 of an int tensor in C order."""
    a = np.ascontiguousarray(a).reshape(-1).astype(np.int64).view(np.uint64)
    wts = (np.arange(a.size, dtype=np.uint64) % np.uint64(65521)) + np.uint64(1)
    with np.errstate(over="ignore"):
        return np.uint64((a * wts).sum(dtype=np.uint64))


def calibrate(arch, head_signed, n=2):
    aux = {}
    sd = synth.make_state_dict(arch, head_signed, input_fraclens={}, aux=aux)
    x = synth.make_input(arch, n, head_signed)
    table = {}

    def calib(layer, t, fa):
        hi = 7 if layer.sym else 8
        bound = 127 if layer.sym else 255
        v = np.abs(t.astype(np.int64)) if layer.sym else np.maximum(t.astype(np.int64), 0)
        best = 0
        for fi in range(hi, -1, -1):
            nshift = fa - fi
            q = (v >> nshift) if nshift >= 0 else (v << (-nshift))
            if (q > bound).mean() <= 0.03:
                best = fi
                break
        # variety: make a shortcut disagree with its sibling body conv where possible
        if ".shortcut." in layer.prefix and best > 0 and "stage_2" in layer.prefix:
            best -= 1
        table[layer.prefix] = int(best)
        layer.b = synth.bias_from(aux[layer.prefix], layer.fw, best)   # bias scale follows fi
        return best

    nets.CALIB = calib
    try:
        nets.forward(arch, sd, x, head_signed)
    finally:
        nets.CALIB = None
    return table


def run_reference(arch, x, sd, flags=()):
    with tempfile.TemporaryDirectory() as td:
        inp, outp = os.path.join(td, "in.npz"), os.path.join(td, "out.npz")
        np.savez(inp, x=x, **sd)
        env = dict(os.environ, PYTHONDONTWRITEBYTECODE="1")
        subprocess.check_call([sys.executable, "-W", "ignore", "-m", "oracle.ref_harness", arch,
                               inp, outp, *flags], cwd=ROOT, env=env)
        o = np.load(outp)
        return {k: o[k] for k in o.files}


def compare_and_pack(arch, head_signed, x, sd, ref, quant_maxpool=False):
    tr = {}
    y = nets.forward(arch, sd, x, head_signed, tr, quant_maxpool=quant_maxpool)
    assert list(ref["keys"]) == list(sd.keys()), "state_dict key order differs from reference"
    assert np.array_equal(y, ref["logits"]), f"{arch}: oracle logits != reference logits"
    refsym = dict(zip(ref["sym_names"], ref["sym_vals"]))
    for c in graph_for(arch, head_signed).convs():
        assert bool(refsym[c.prefix]) == c.sym, f"input_symmetric mismatch at {c.prefix}"
    names, sums = [], []
    for k in sorted(ref.keys()):
        if ":" not in k or k == "head.0:in8":
            continue
        b = ref[k]
        if k in ("head.0:acc", "tail.0:acc"):
            b = np.maximum(b, 0)      # the oracle traces these after the in-place ReLU
        a = tr[k].reshape(b.shape)
        assert np.array_equal(a, b), f"{arch}: oracle != reference at {k}"
        names.append(k)
        sums.append(checksum(b))
    live = float((ref["logits"] != ref["logits"][0, 0]).mean())
    print(f"  {arch}: oracle == reference on logits and {len(names)} layer tensors; "
          f"logit range [{ref['logits'].min():.0f}, {ref['logits'].max():.0f}], live={live:.2f}")
    return dict(logits=ref["logits"].astype(np.int32), layer_names=np.array(names),
                layer_checksums=np.array(sums, dtype=np.uint64))


def main(archs):
    os.makedirs(DATA, exist_ok=True)
    for arch in archs:
        hs = synth.HEAD_SIGNED[arch]
        print(f"[{arch}] calibrating input fraclens on the oracle ...")
        table = calibrate(arch, hs)
        with open(os.path.join(DATA, f"fraclens_{arch}.json"), "w") as f:
            json.dump(table, f, indent=0, sort_keys=True)
        hist = {}
        for v in table.values():
            hist[v] = hist.get(v, 0) + 1
        print(f"  fi histogram: {dict(sorted(hist.items()))}")
        sd = synth.make_state_dict(arch, hs, input_fraclens=table)
        x = synth.make_input(arch, N_GOLD, hs)
        ref = run_reference(arch, x, sd)
        np.savez_compressed(os.path.join(GOLD, f"{arch}_n{N_GOLD}.npz"),
                            **compare_and_pack(arch, hs, x, sd, ref))
        # adversarial family
        esd = synth.make_edge_state_dict(arch, hs)
        ex = synth.make_input(arch, N_GOLD, hs, seed=777)
        try:
            eref = run_reference(arch, ex, esd)
        except subprocess.CalledProcessError:
            print(f"  {arch}: reference asserted on the edge fixture (avgpool bound); skipped")
            continue
        np.savez_compressed(os.path.join(GOLD, f"edge_{arch}_n{N_GOLD}.npz"),
                            **compare_and_pack(arch, hs, ex, esd, eref))


if __name__ == "__main__":
    main(sys.argv[1:] or list(synth.HEAD_SIGNED))

#!/usr/bin/env python
"""Pins f8net_b200/export.py against the UNMODIFIED reference.

Authoring container only (needs /root/reference).  Per architecture (one subprocess each: the
reference's FLAGS singleton hosts one config per process):

  1. synth.make_float_state_dict(arch, seed) -- a seeded float-simulation checkpoint;
  2. the reference float-sim ``Model`` loads it (fix_train.py:877-891) and ``Model.int_model()``
     converts it (fix_train.py:930-934)  -> reference IntModel.state_dict();
  3. export_int_state_dict(same checkpoint) must be identical, key for key, bit for bit;
  4. commit per-tensor SHA-256 + the fraclens + every int bias to tests/golden/export_<arch>.npz
     (tests/test_export.py re-checks the exporter against them without the reference).

    python tests/golden/make_export_golden.py [arch ...]
"""
import hashlib
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
SEED = 77


# config variants beyond the shipped *_int_op_only*.yml files: name -> (ExportFlags fields, FLAGS overrides)
VARIANTS = {
    "sharing": (dict(input_fraclen_sharing=True), dict(input_fraclen_sharing=True)),
    "rms": (dict(metric="rms"), dict(metric="rms")),
    "mae": (dict(metric="mae"), dict(metric="mae")),
    "convrescale": (dict(rescale_forward_conv=True, rescale_type="stddev"),
                    dict(rescale_forward_conv=True, rescale_type="stddev")),
    "norescale": (dict(rescale_forward=False), dict(rescale_forward=False)),
}


def flags_for(arch, variant=None):
    from f8net_b200.export import ExportFlags
    if arch == "resnet50":      # res50 tiny_finetuning config: normalize, no_clipping, grid search
        f = ExportFlags(normalize=True, no_clipping=True, format_grid_search=True)
    else:
        f = ExportFlags()
    if variant:
        for k, v in VARIANTS[variant][0].items():
            setattr(f, k, v)
    return f


def digest(t):
    import numpy as np
    a = np.ascontiguousarray(t.detach().cpu().numpy())
    return hashlib.sha256(a.tobytes()).hexdigest()


def one(arch, variant=None):
    import numpy as np
    import torch
    from f8net_b200 import synth
    from f8net_b200.export import export_int_state_dict
    from oracle import ref_harness
    torch.set_num_threads(8)
    flags = flags_for(arch, variant)
    fsd = synth.make_float_state_dict(arch, SEED, flags)
    im, FLAGS = ref_harness.build_int_model(arch, float_state_dict=fsd, keep_grid_search=True,
                                            flag_overrides=VARIANTS[variant][1] if variant else None)
    for k in ("normalize", "no_clipping", "format_grid_search"):
        assert bool(getattr(FLAGS, k, False)) == bool(getattr(flags, k)), (k, getattr(FLAGS, k, None))
    ref = im.state_dict()
    mine = export_int_state_dict(fsd, arch, flags)
    assert list(ref.keys()) == list(mine.keys()), "key order differs"
    bad = []
    for k in ref:
        r, m = ref[k], mine[k]
        if r.dtype != m.dtype or r.shape != m.shape or not torch.equal(r, m):
            bad.append((k, str(r.dtype), str(m.dtype), tuple(r.shape), tuple(m.shape),
                        int((r != m).sum()) if r.shape == m.shape else -1))
    if bad:
        for b in bad[:20]:
            print("MISMATCH", b)
        raise SystemExit(f"{arch}: {len(bad)} tensors differ from the reference")
    out = {"keys": np.array(list(ref.keys())), "sha256": np.array([digest(ref[k]) for k in ref]),
           "seed": np.array(SEED)}
    for k in ref:
        if not k.endswith(".weight"):
            out[k] = ref[k].numpy()
    tag = f"{arch}_{variant}" if variant else arch
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", f"export_{tag}.npz"), **out)
    fws = [int(ref[k]) for k in ref if k.endswith("weight_fraclen")]
    print(f"{tag}: {len(ref)} tensors identical to the reference int_model(); weight fraclens "
          f"{sorted(set(fws))}")


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--one":
        one(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else None)
    else:
        archs = sys.argv[1:] or ["resnet18", "resnet50", "mobilenet_v1", "mobilenet_v2"]
        env = dict(os.environ, PYTHONDONTWRITEBYTECODE="1")
        for a in archs:
            subprocess.check_call([sys.executable, os.path.abspath(__file__), "--one", a], env=env, cwd=ROOT)
        if not sys.argv[1:]:        # config variants on the two small networks
            for a, v in [("resnet18", v) for v in VARIANTS] + [("mobilenet_v2", "sharing")]:
                subprocess.check_call([sys.executable, os.path.abspath(__file__), "--one", a, v], env=env, cwd=ROOT)

"""Generates the round-2 fixture families with the UNMODIFIED reference (authoring container only):

    python tests/golden/make_variant_golden.py

 * trained_<name>_n2.npz -- the TRAINED per-layer formats parsed from the reference's own logs
   (fraclen_visual/*.out -> f8net_b200/data/trained_fraclens_*.json, tools/parse_fraclen_logs.py)
   with synthetic weights (synth.make_trained_state_dict): MobileNetV2 and both ResNet50s.
 * qmaxpool_<arch>_n2.npz -- ResNet18 / ResNet50 on synth.make_maxpool_state_dict, run twice through
   the reference: quant_maxpool False (nn.MaxPool2d on floats, the shipped int_op_only configs) and
   True (FXQMaxPool2d, integer max).  The two logits differ; both are pinned.

Every file: reference logits + a position-weighted checksum of every layer's 8-bit input and int32
accumulator, after checking that the C oracle reproduces all of them.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from f8net_b200 import synth  # noqa: E402
from make_golden import GOLD, compare_and_pack, run_reference  # noqa: E402

TRAINED_SEED, QMP_SEED = 5150, 808


def main():
    for name in synth.TRAINED:
        arch, hs, sd = synth.make_trained_state_dict(name)
        x = synth.make_input(arch, 2, hs, seed=TRAINED_SEED)
        ref = run_reference(arch, x, sd)
        np.savez_compressed(os.path.join(GOLD, f"trained_{name}_n2.npz"),
                            **compare_and_pack(arch, hs, x, sd, ref))
    for arch in ("resnet18", "resnet50"):
        hs = synth.HEAD_SIGNED[arch]
        sd = synth.make_maxpool_state_dict(arch, hs)
        x = synth.make_input(arch, 2, hs, seed=QMP_SEED)
        f = compare_and_pack(arch, hs, x, sd, run_reference(arch, x, sd))
        q = compare_and_pack(arch, hs, x, sd, run_reference(arch, x, sd, ["quant_maxpool=1"]),
                             quant_maxpool=True)
        assert (f["logits"] != q["logits"]).mean() > 0.5, "fixture does not tell the two pools apart"
        np.savez_compressed(os.path.join(GOLD, f"qmaxpool_{arch}_n2.npz"),
                            logits_float_pool=f["logits"], logits=q["logits"],
                            layer_names=q["layer_names"], layer_checksums=q["layer_checksums"],
                            layer_checksums_float_pool=f["layer_checksums"])
        print(f"  {arch}: FXQMaxPool2d vs float round trip differ in "
              f"{100 * (f['logits'] != q['logits']).mean():.1f} % of the logits")


if __name__ == "__main__":
    main()

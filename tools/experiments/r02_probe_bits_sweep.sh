#!/bin/bash
out=gpurun_out
F8_DEBUG_PROBES=1 python -m f8net_b200.build --force > $out/r02aa_build.log 2>&1
for p in 0 1 2 3; do
  echo "== F8_UPROBE=$p"; F8_UPROBE=$p timeout 90 python tools/profile_ops.py --arch mobilenet_v1 --batch 256 --chunk 256 2>&1 | grep "stage_3_layer_[12].body.2\|stage_4_layer_1.body.2\|stage_1_layer_1.body.2\|back-to-back"
done

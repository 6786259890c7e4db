#!/bin/bash
out=gpurun_out; tag=r02h
timeout 90 python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_smoke.log 2>&1 || { echo "SMOKE FAILED/HUNG"; exit 1; }
for m in 0 2; do
  F8_MC_GENERIC=$m timeout 90 python tools/profile_ops.py --arch resnet18 --batch 256 --chunk 256 > $out/${tag}_per_layer_resnet18_mcg$m.txt 2>&1
  head -n 1 $out/${tag}_per_layer_resnet18_mcg$m.txt
  F8_MC_GENERIC=$m timeout 90 python tools/profile_ops.py --arch resnet50 --batch 256 --chunk 256 > $out/${tag}_per_layer_resnet50_mcg$m.txt 2>&1
  head -n 1 $out/${tag}_per_layer_resnet50_mcg$m.txt
done

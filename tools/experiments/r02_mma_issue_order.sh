#!/bin/bash
# MMA issue-order probe for the 3x3 kernel (debug build on the box; the shipping .so is not touched in the repo)
out=gpurun_out
F8_DEBUG_PROBES=1 python -m f8net_b200.build --force > $out/r02d_build.log 2>&1
for p in 0 64 128; do
  F8_PROBE=$p python tools/profile_ops.py --arch resnet18 --batch 256 --chunk 256 > $out/r02d_r18_probe$p.txt 2>&1
done
F8_STATS=1 F8_PROBE=0 python tools/profile_ops.py --arch resnet18 --batch 256 --chunk 256 --reps 1 2>&1 | grep "f8 stats" | head -n 80 > $out/r02d_stats0.txt
paste <(awk '{print $1, $4}' $out/r02d_r18_probe0.txt) <(awk '{print $4}' $out/r02d_r18_probe64.txt) <(awk '{print $4}' $out/r02d_r18_probe128.txt)

#!/bin/bash
out=gpurun_out
F8_DEBUG_PROBES=1 python -m f8net_b200.build --force > $out/r02g_build.log 2>&1
rm -f $out/r02g_stats.txt
for p in 784 1808 3856 12048 ; do
  echo "== F8_PROBE=$p" >> $out/r02g_stats.txt
  F8_STATS=1 F8_PROBE=$p timeout 120 python tools/profile_ops.py --arch resnet18 --batch 256 --chunk 256 --reps 1 2>&1 | grep "f8 stats\] conv3x3" | sed -n '1p;13p' | cut -c1-330 >> $out/r02g_stats.txt
done
cat $out/r02g_stats.txt

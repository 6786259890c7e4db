#!/bin/bash
# what slows the MMA loop of the 3x3 kernel below the isolated tcgen05 rate? (debug build on the box)
out=gpurun_out
F8_DEBUG_PROBES=1 python -m f8net_b200.build --force > $out/r02f_build.log 2>&1
for p in 0 16 256 512 768 784; do
  echo "== F8_PROBE=$p" >> $out/r02f_stats.txt
  F8_STATS=1 F8_PROBE=$p python tools/profile_ops.py --arch resnet18 --batch 256 --chunk 256 --reps 1 2>&1 | grep "f8 stats\] conv3x3" | sed -n '1p;5p;9p;13p' | cut -c1-330 >> $out/r02f_stats.txt
done
cat $out/r02f_stats.txt

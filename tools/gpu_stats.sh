#!/bin/bash
# debug build on the box + in-kernel wait counters (F8_STATS) for one network
tag=${1:-stats}; arch=${2:-resnet18}; probe=${3:-0}
out=gpurun_out
F8_DEBUG_PROBES=1 python -m f8net_b200.build --force > $out/${tag}_build.log 2>&1
F8_STATS=1 F8_PROBE=$probe timeout 300 python tools/profile_ops.py --arch $arch --batch 256 --chunk 256 --reps 1 2>&1 | grep "f8 stats" | cut -c1-420 > $out/${tag}_stats_$arch.txt
head -n 40 $out/${tag}_stats_$arch.txt

#!/bin/bash
# source-level ncu capture of single launches of one forward pass (batch 256):
#   bash tools/gpu_src_profile.sh <tag> <arch> <launch index in the pass> [<launch index> ...]
# writes gpurun_out/<tag>_src_<arch>_<idx>.csv (per-SASS-instruction samples / stall reasons) and the raw page
tag=$1; arch=$2; shift 2
out=gpurun_out
mkdir -p $out
KF='regex:umma_kernel|conv1x1_res|head3x3s2|head_pool2|pool_fc|dw3x3|conv_mma|maxpool_kernel|pool_requant|convert_input|integerize|requant_i32'
n=$(timeout 120 python tools/one_pass.py --arch $arch --count | tail -n 1)
for idx in "$@"; do
  timeout 300 ncu --set full --clock-control none --import-source on -k "$KF" -s $((n + idx)) -c 1 -f -o $out/${tag}_src_${arch}_$idx \
    python tools/one_pass.py --arch $arch --passes 2 > $out/${tag}_src_${arch}_$idx.log 2>&1
  timeout 120 ncu -i $out/${tag}_src_${arch}_$idx.ncu-rep --page source --csv > $out/${tag}_src_${arch}_$idx.csv 2> /dev/null
  timeout 120 ncu -i $out/${tag}_src_${arch}_$idx.ncu-rep --page raw --csv > $out/${tag}_raw_${arch}_$idx.csv 2> /dev/null
  ls -la $out/${tag}_src_${arch}_$idx.ncu-rep
  rm -f $out/${tag}_src_${arch}_$idx.ncu-rep
done

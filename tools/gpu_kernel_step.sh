#!/bin/bash
# one kernel iteration on the GPU box: smoke (90 s cap: a deadlocked kernel must not eat the budget), kernel parity, per-layer tables, whole-net + per-layer-in-net parity
tag=${1:-r02e}
archs=${2:-"resnet18 resnet50"}
out=gpurun_out
mkdir -p $out
timeout 90 python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_smoke.log 2>&1 || { echo "SMOKE FAILED/HUNG"; tail -n 5 $out/${tag}_smoke.log; exit 1; }
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q 2>&1 | tail -n 8 > $out/${tag}_tests_a.log
cat $out/${tag}_tests_a.log
for a in $archs; do
  timeout 90 python tools/profile_ops.py --arch $a --batch 256 --chunk 256 > $out/${tag}_per_layer_${a}.txt 2>&1
  head -n 1 $out/${tag}_per_layer_${a}.txt
done
timeout 420 python -m pytest tests/test_gpu_nets.py tests/test_gpu_layers_in_net.py -m gpu -x -q 2>&1 | tail -n 6 > $out/${tag}_tests_b.log
cat $out/${tag}_tests_b.log

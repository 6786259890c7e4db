#!/bin/bash
# one build-measure iteration on the GPU box: kernel + whole-net parity tests, per-layer tables.
# Every step runs under its own short timeout: a deadlocked kernel must not eat the GPU budget.
tag=${1:-step}
archs=${2:-"resnet18 mobilenet_v2"}
out=gpurun_out
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_smoke.log 2>&1 || { echo "SMOKE FAILED/HUNG"; tail -n 5 $out/${tag}_smoke.log; exit 1; }
timeout 420 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_nets.py -m gpu -x -q 2>&1 | tail -n 6 > $out/${tag}_tests.log
for a in $archs; do
  timeout 90 python tools/profile_ops.py --arch $a --batch 256 --chunk 256 > $out/${tag}_per_layer_$a.txt 2>&1
done
cat $out/${tag}_tests.log
for a in $archs; do head -n 1 $out/${tag}_per_layer_$a.txt; done

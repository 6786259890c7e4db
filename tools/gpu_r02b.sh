#!/bin/bash
# round-2 step: input range check tests + residual-carry L2 prefetch A/B (per-layer tables with the prefetch off / on)
tag=${1:-r02b}
out=gpurun_out
mkdir -p $out
timeout 90 python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_smoke.log 2>&1 || { echo "SMOKE FAILED/HUNG"; tail -n 5 $out/${tag}_smoke.log; exit 1; }
timeout 300 python -m pytest tests/test_gpu_input.py tests/test_gpu_kernels.py -m gpu -x -q 2>&1 | tail -n 12 > $out/${tag}_tests_a.log
cat $out/${tag}_tests_a.log
for a in resnet18 resnet50 mobilenet_v2; do
  F8_CARRY_PREFETCH=0 timeout 90 python tools/profile_ops.py --arch $a --batch 256 --chunk 256 > $out/${tag}_per_layer_${a}_pf0.txt 2>&1
  timeout 90 python tools/profile_ops.py --arch $a --batch 256 --chunk 256 > $out/${tag}_per_layer_${a}_pf1.txt 2>&1
  head -n 1 $out/${tag}_per_layer_${a}_pf0.txt $out/${tag}_per_layer_${a}_pf1.txt
done
timeout 420 python -m pytest tests/test_gpu_nets.py -m gpu -x -q 2>&1 | tail -n 6 > $out/${tag}_tests_b.log
cat $out/${tag}_tests_b.log

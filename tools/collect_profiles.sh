#!/bin/bash
# Runs on the GPU box (under gpurun): tests, bench lines, per-layer tables, ncu launch list and the
# ncu --set full capture of one forward pass.  Outputs under gpurun_out/<tag>_*.
tag=${1:-r01}
out=gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > $out/${tag}_tests.log
python bench.py > $out/${tag}_bench_resnet18.json 2> $out/${tag}_bench_resnet18.err
for a in resnet50 mobilenet_v1 mobilenet_v2; do
  python bench.py --arch $a --no-cpu-baseline > $out/${tag}_bench_$a.json 2> $out/${tag}_bench_$a.err
done
python bench.py --impl reference --steps 3 --warmup 1 > $out/${tag}_bench_reference.json 2>/dev/null
for a in resnet18 resnet50 mobilenet_v1 mobilenet_v2; do
  python tools/profile_ops.py --arch $a --batch 256 --chunk 256 > $out/${tag}_per_layer_$a.txt 2>&1
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $out/${tag}_resnet18_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > $out/${tag}_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"umma_kernel|head_pool|pool_fc" -s 63 -c 21 \
  -o $out/${tag}_dense python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph > $out/${tag}_ncu_full.log 2>&1
cat $out/${tag}_tests.log
cat $out/${tag}_bench_resnet18.json

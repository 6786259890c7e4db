#!/bin/bash
# Round-2 evidence collection on the GPU box (under gpurun).  Everything lands in gpurun_out/<tag>_*; copy what
# should be judged into profiles/ afterwards (tools/make_profiles.py turns the ncu outputs into summaries).
#   bash tools/collect_profiles.sh r02 [full]
tag=${1:-r02}
full=${2:-}
out=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -n 4 > $out/${tag}_tests.log
timeout 600 python bench.py --steps 20 --warmup 5 > $out/${tag}_bench_n1.json 2> $out/${tag}_bench_n1.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $out/${tag}_bench_reference.json 2>/dev/null
for a in resnet18 resnet50 mobilenet_v1 mobilenet_v2; do
  timeout 120 python tools/profile_ops.py --arch $a --batch 256 --chunk 256 > $out/${tag}_per_layer_$a.txt 2>&1
done
# only this library's kernels count towards ncu's launch-skip / launch-count (torch's input generation also launches)
KF='regex:umma_kernel|conv1x1_res|head3x3s2|head_pool2|pool_fc|dw3x3|conv_mma|maxpool_kernel|pool_requant|convert_input|integerize|requant_i32'
if [ -n "$full" ]; then
  for a in resnet18 resnet50 mobilenet_v1 mobilenet_v2; do
    n=$(timeout 120 python tools/one_pass.py --arch $a --count | tail -n 1)
    # launch list: per-launch gpu time of the second pass (cold caches, serialised: shares, not absolutes)
    timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KF" -s $n -c $n --csv --log-file $out/${tag}_launches_$a.csv \
      python tools/one_pass.py --arch $a --passes 2 --names $out/${tag}_names_$a.json > $out/${tag}_ncu_launches_$a.log 2>&1
    # full capture of the same pass
    timeout 900 ncu --set full --clock-control none --import-source on -k "$KF" -s $n -c $n -f -o $out/${tag}_full_$a \
      python tools/one_pass.py --arch $a --passes 2 > $out/${tag}_ncu_full_$a.log 2>&1
    timeout 300 ncu -i $out/${tag}_full_$a.ncu-rep --page raw --csv > $out/${tag}_full_$a.csv 2> /dev/null
    rm -f $out/${tag}_full_$a.ncu-rep          # gpurun brings back at most 64 MiB: keep the CSV export only
  done
fi
if [ -n "$full" ]; then
  for tool in memcheck initcheck racecheck; do
    timeout 600 compute-sanitizer --tool $tool --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_sanitizer_$tool.log 2>&1
    echo "exit $?" >> $out/${tag}_sanitizer_$tool.log
  done
fi
du -sh $out
cat $out/${tag}_tests.log
cut -c1-300 $out/${tag}_bench_n1.json

#!/bin/bash
# round-2 first GPU call: regression tests, epilogue/MMA probes, sanitizer passes on smoke()
out=gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > $out/r02a_tests.log
./tools/probes/epi2_probe > $out/r02a_epi2_probe.txt 2>&1
for a in resnet18 mobilenet_v2; do
  python tools/profile_ops.py --arch $a --batch 256 --chunk 256 > $out/r02a_per_layer_$a.txt 2>&1
done
for tool in memcheck initcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > $out/r02a_sanitizer_$tool.log 2>&1
  echo "exit $?" >> $out/r02a_sanitizer_$tool.log
done
cat $out/r02a_tests.log
tail -3 $out/r02a_sanitizer_*.log

#!/bin/bash
# round-2: new bench line (configs array, parity block, per-template roofline, int8 peak) at N=1 and N=2,
# NCCL gathered-logits test on two GPUs
out=gpurun_out
python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -n 5 > $out/r02b_multi_tests.log
python bench.py --steps 20 --warmup 5 > $out/r02b_bench_n1.json 2> $out/r02b_bench_n1.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus 2 --steps 20 --warmup 5 > $out/r02b_bench_n2.json 2> $out/r02b_bench_n2.err
python bench.py --impl reference --steps 2 --warmup 1 > $out/r02b_bench_ref.json 2>/dev/null
cat $out/r02b_multi_tests.log
tail -n 5 $out/r02b_bench_n1.err $out/r02b_bench_n2.err
cut -c1-600 $out/r02b_bench_n1.json
cut -c1-600 $out/r02b_bench_n2.json

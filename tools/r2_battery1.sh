#!/bin/bash
# round 2, battery 1: whole GPU suite with the second-generation LSTM kernels, A/B against the first generation, benches
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/b1_gpu.txt 2>&1
( time timeout 1500 python -m pytest tests -m gpu -q -s -p no:cacheprovider ) > gpurun_out/b1_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/b1_tests.log
if ! grep -q " passed" gpurun_out/b1_tests.log || grep -q "failed" gpurun_out/b1_tests.log; then
  ( B200_LSTM_V1=1 B200_CLUSTER_V1=1 timeout 900 python -m pytest tests/test_forward_gpu.py tests/test_golden.py -m gpu -q -x -p no:cacheprovider ) > gpurun_out/b1_tests_v1.log 2>&1
fi
timeout 600 python bench.py > gpurun_out/b1_bench_default.json 2> gpurun_out/b1_bench_default.err
for r in 1 3 4; do
  timeout 200 python bench.py --runners $r --no-sub-models --no-cpu-baseline > gpurun_out/b1_bench_fast_r$r.json 2>> gpurun_out/b1_bench.err
done
B200_LSTM_V1=1 timeout 200 python bench.py --no-sub-models --no-cpu-baseline > gpurun_out/b1_bench_fast_v1.json 2>> gpurun_out/b1_bench.err
B200_CLUSTER_V1=1 timeout 300 python bench.py --model hac --batch 512 --steps 8 --no-cpu-baseline > gpurun_out/b1_bench_hac_v1.json 2>> gpurun_out/b1_bench.err
timeout 300 python bench.py --model hac --batch 512 --steps 8 --runners 1 --no-cpu-baseline > gpurun_out/b1_bench_hac_r1.json 2>> gpurun_out/b1_bench.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/b1_bench_reference.json 2>> gpurun_out/b1_bench.err
echo done > gpurun_out/b1_done

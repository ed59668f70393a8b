#!/bin/bash
# round 2, battery 8: folded RMSNorm, FMA-pipe reciprocal in the gate activations, GEMM epilogue policy, chunk benchmark tables
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q -s -p no:cacheprovider ) > gpurun_out/b8_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/b8_tests.log
if grep -q "failed" gpurun_out/b8_tests.log; then
  ( B200_TX_RMSNORM_PASS=1 timeout 600 python -m pytest tests/test_forward_gpu.py tests/test_golden.py -m gpu -q -k "tx or sup" -p no:cacheprovider ) > gpurun_out/b8_tests_normpass.log 2>&1
fi
for cfg in "fast 512" "hac 512"; do
  echo "== $cfg" >> gpurun_out/b8_timeline.txt
  timeout 120 python tools/lstm_timeline.py $cfg 2>> gpurun_out/b8_timeline.txt >/dev/null
done
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/b8_bench_default.json 2> gpurun_out/b8_bench_default.err
B200_TX_RMSNORM_PASS=1 timeout 300 python bench.py --model sup --batch 128 --steps 6 --no-cpu-baseline > gpurun_out/b8_bench_sup_normpass.json 2>> gpurun_out/b8_bench.err
timeout 900 python tools/gen_chunk_benchmarks.py > gpurun_out/b8_chunk_benchmarks_b200.inc 2> gpurun_out/b8_chunk_benchmarks.err
echo done > gpurun_out/b8_done

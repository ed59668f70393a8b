#!/bin/bash
# round 2, battery 2: suite with the staged GEMM epilogue + pool + lifecycle, kernel timelines, GEMM A/B
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q -s -p no:cacheprovider ) > gpurun_out/b2_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/b2_tests.log
if grep -q "failed" gpurun_out/b2_tests.log; then
  ( B200_GEMM_DIRECT=1 timeout 900 python -m pytest tests/test_gemm_gpu.py tests/test_forward_gpu.py -m gpu -q -x -p no:cacheprovider ) > gpurun_out/b2_tests_direct.log 2>&1
fi
for cfg in "fast 512" "hac 512" "hac 256"; do
  echo "== v2 $cfg" >> gpurun_out/b2_timeline.txt
  timeout 120 python tools/lstm_timeline.py $cfg 2>> gpurun_out/b2_timeline.txt >/dev/null
done
echo "== v1 fast 512" >> gpurun_out/b2_timeline.txt
B200_LSTM_V1=1 timeout 120 python tools/lstm_timeline.py fast 512 2>> gpurun_out/b2_timeline.txt >/dev/null
for sl in 3 4 5; do
  echo "== beam state_len $sl" >> gpurun_out/b2_timeline.txt
  timeout 120 python tools/beam_timeline.py $sl 128 2>> gpurun_out/b2_timeline.txt >/dev/null
done
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/b2_bench_default.json 2> gpurun_out/b2_bench_default.err
B200_GEMM_DIRECT=1 timeout 300 python bench.py --model sup --batch 128 --steps 6 --no-cpu-baseline > gpurun_out/b2_bench_sup_direct.json 2>> gpurun_out/b2_bench.err
B200_GEMM_DIRECT=1 timeout 300 python bench.py --model hac --batch 512 --steps 8 --no-cpu-baseline > gpurun_out/b2_bench_hac_direct.json 2>> gpurun_out/b2_bench.err
echo done > gpurun_out/b2_done

#!/bin/bash
# round 2, battery 7: GEMM with four epilogue warp sets, cluster mask fix, chunk benchmark tables
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests/test_gemm_gpu.py tests/test_forward_gpu.py tests/test_golden.py -m gpu -q -p no:cacheprovider ) > gpurun_out/b7_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/b7_tests.log
echo "== hac 512" >> gpurun_out/b7_timeline.txt
timeout 120 python tools/lstm_timeline.py hac 512 2>> gpurun_out/b7_timeline.txt >/dev/null
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/b7_bench_default.json 2> gpurun_out/b7_bench_default.err
B200_GEMM_DIRECT=1 timeout 300 python bench.py --model sup --batch 128 --steps 6 --no-cpu-baseline > gpurun_out/b7_bench_sup_direct.json 2>> gpurun_out/b7_bench.err
B200_GEMM_DIRECT=1 timeout 300 python bench.py --model hac --batch 512 --steps 8 --no-cpu-baseline > gpurun_out/b7_bench_hac_direct.json 2>> gpurun_out/b7_bench.err
timeout 900 python tools/gen_chunk_benchmarks.py > gpurun_out/b7_chunk_benchmarks_b200.inc 2> gpurun_out/b7_chunk_benchmarks.err
echo done > gpurun_out/b7_done

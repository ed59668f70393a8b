#!/bin/bash
# round 2, battery 22: operand-stationary GEMM schedule (hac x-projection: W_ih tiles resident; sup QKV / FC1: W slice resident) --
# GEMM and forward parity first, then hac / sup lines with the schedule on and off
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests/test_gemm_gpu.py tests/test_forward_gpu.py tests/test_full_size_gpu.py -m gpu -q -x -s -p no:cacheprovider ) > gpurun_out/b22_tests_fwd.log 2>&1
echo "tests rc=$?" >> gpurun_out/b22_tests_fwd.log
H="timeout 300 python bench.py --model hac --batch 512 --steps 8 --no-cpu-baseline"
$H > gpurun_out/b22_hac_wstat.json 2> gpurun_out/b22_bench.err
B200_GEMM_NO_WSTAT=1 $H > gpurun_out/b22_hac_default.json 2>> gpurun_out/b22_bench.err
S="timeout 300 python bench.py --model sup --batch 128 --steps 6 --no-cpu-baseline"
$S > gpurun_out/b22_sup_wstat.json 2>> gpurun_out/b22_bench.err
B200_GEMM_NO_WSTAT=1 $S > gpurun_out/b22_sup_default.json 2>> gpurun_out/b22_bench.err
echo done > gpurun_out/b22_done

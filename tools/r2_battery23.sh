#!/bin/bash
# round 2, battery 23: GEMM tile order (the shorter dimension runs fastest: the hac x-projection no longer sweeps its 655 MB of
# activations once per weight tile) -- GEMM / forward / full-size parity, then hac, fast and sup lines; old order for the A/B
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests/test_gemm_gpu.py tests/test_forward_gpu.py tests/test_full_size_gpu.py -m gpu -q -x -s -p no:cacheprovider ) > gpurun_out/b23_tests_fwd.log 2>&1
echo "tests rc=$?" >> gpurun_out/b23_tests_fwd.log
H="timeout 300 python bench.py --model hac --batch 512 --steps 8 --no-cpu-baseline"
$H > gpurun_out/b23_hac.json 2> gpurun_out/b23_bench.err
B200_GEMM_NFAST=1 $H > gpurun_out/b23_hac_old_order.json 2>> gpurun_out/b23_bench.err
$H --runners 3 > gpurun_out/b23_hac_r3.json 2>> gpurun_out/b23_bench.err
$H --runners 5 > gpurun_out/b23_hac_r5.json 2>> gpurun_out/b23_bench.err
timeout 300 python bench.py --no-cpu-baseline --no-sub-models > gpurun_out/b23_fast.json 2>> gpurun_out/b23_bench.err
timeout 300 python bench.py --model sup --batch 128 --steps 6 --no-cpu-baseline > gpurun_out/b23_sup.json 2>> gpurun_out/b23_bench.err
echo done > gpurun_out/b23_done

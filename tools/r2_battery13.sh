#!/bin/bash
# round 2, battery 13: GEMM epilogue changes (residual by TMA, RoPE table layout), gate activation, full suite; A/B switches; GEMM grid cap with 3-4 hac runners
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q -s -p no:cacheprovider ) > gpurun_out/b13_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/b13_tests.log
S="timeout 300 python bench.py --model sup --batch 128 --steps 6 --no-cpu-baseline"
$S > gpurun_out/b13_sup.json 2> gpurun_out/b13_bench.err
B200_GEMM_RES_LDG=1 $S > gpurun_out/b13_sup_res_ldg.json 2>> gpurun_out/b13_bench.err
B200_GEMM_STAGE_ALL=1 $S > gpurun_out/b13_sup_stage_all.json 2>> gpurun_out/b13_bench.err
$S --runners 3 > gpurun_out/b13_sup_r3.json 2>> gpurun_out/b13_bench.err
H="timeout 300 python bench.py --model hac --batch 512 --steps 8 --no-cpu-baseline"
$H > gpurun_out/b13_hac.json 2>> gpurun_out/b13_bench.err
for cap in 100 116; do for r in 3 4; do
  B200_GEMM_MAX_CTAS=$cap $H --runners $r > gpurun_out/b13_hac_cap${cap}_r${r}.json 2>> gpurun_out/b13_bench.err
done; done
timeout 600 python bench.py --no-cpu-baseline --no-sub-models > gpurun_out/b13_fast.json 2>> gpurun_out/b13_bench.err
echo done > gpurun_out/b13_done

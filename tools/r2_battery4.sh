#!/bin/bash
# round 2, battery 4: variable chunk sizes, C++ adapter executed, GEMM regression bisect (library with the battery-1 GEMM)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q -s -p no:cacheprovider ) > gpurun_out/b4_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/b4_tests.log
for lib in new old; do
  if [ $lib = old ]; then export B200CALL_LIB=$PWD/dorado_b200/libb200call_oldgemm.so; else unset B200CALL_LIB; fi
  timeout 300 python bench.py --model hac --batch 512 --steps 8 --no-cpu-baseline > gpurun_out/b4_bench_hac_$lib.json 2>> gpurun_out/b4_bench.err
  timeout 300 python bench.py --model sup --batch 128 --steps 6 --no-cpu-baseline > gpurun_out/b4_bench_sup_$lib.json 2>> gpurun_out/b4_bench.err
done
unset B200CALL_LIB
B200_GEMM_DIRECT=1 timeout 300 python bench.py --model sup --batch 128 --steps 6 --no-cpu-baseline > gpurun_out/b4_bench_sup_direct.json 2>> gpurun_out/b4_bench.err
timeout 300 python bench.py --no-sub-models --no-cpu-baseline > gpurun_out/b4_bench_fast.json 2>> gpurun_out/b4_bench.err
echo done > gpurun_out/b4_done

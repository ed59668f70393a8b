#!/bin/bash
# round 2, battery 5: GEMM regression bisect (library with the battery-1 GEMM), all-gather by remote vector stores
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( B200_CLUSTER_GATHER=st timeout 600 python -m pytest tests/test_forward_gpu.py -m gpu -q -k "hac or variable" -p no:cacheprovider ) > gpurun_out/b5_tests_st.log 2>&1
for mode in st dsmem; do
  echo "== hac 512 gather=$mode" >> gpurun_out/b5_timeline.txt
  B200_CLUSTER_GATHER=$mode timeout 120 python tools/lstm_timeline.py hac 512 2>> gpurun_out/b5_timeline.txt >/dev/null
done
B200_CLUSTER_GATHER=st timeout 300 python bench.py --model hac --batch 512 --steps 8 --no-cpu-baseline > gpurun_out/b5_bench_hac_st.json 2>> gpurun_out/b5_bench.err
B200_CLUSTER_GATHER=st timeout 300 python bench.py --model hac --batch 512 --steps 8 --runners 1 --no-cpu-baseline > gpurun_out/b5_bench_hac_st_r1.json 2>> gpurun_out/b5_bench.err
B200CALL_LIB=$PWD/dorado_b200/libb200call_oldgemm.so timeout 300 python bench.py --model hac --batch 512 --steps 8 --no-cpu-baseline > gpurun_out/b5_bench_hac_oldgemm.json 2>> gpurun_out/b5_bench.err
B200CALL_LIB=$PWD/dorado_b200/libb200call_oldgemm.so timeout 300 python bench.py --model sup --batch 128 --steps 6 --no-cpu-baseline > gpurun_out/b5_bench_sup_oldgemm.json 2>> gpurun_out/b5_bench.err
echo done > gpurun_out/b5_done

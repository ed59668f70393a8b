#!/bin/bash
# round 2, battery 3: GEMM ring fix (staged vs direct), cluster all-gather through L2 vs DSMEM
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests/test_gemm_gpu.py tests/test_forward_gpu.py tests/test_golden.py -m gpu -q -p no:cacheprovider ) > gpurun_out/b3_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/b3_tests.log
( B200_CLUSTER_GATHER=dsmem timeout 600 python -m pytest tests/test_forward_gpu.py -m gpu -q -k "hac" -p no:cacheprovider ) > gpurun_out/b3_tests_dsmem.log 2>&1
for mode in l2 dsmem; do
  echo "== hac 512 gather=$mode" >> gpurun_out/b3_timeline.txt
  B200_CLUSTER_GATHER=$mode timeout 120 python tools/lstm_timeline.py hac 512 2>> gpurun_out/b3_timeline.txt >/dev/null
done
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/b3_bench_default.json 2> gpurun_out/b3_bench_default.err
B200_CLUSTER_GATHER=dsmem timeout 300 python bench.py --model hac --batch 512 --steps 8 --no-cpu-baseline > gpurun_out/b3_bench_hac_dsmem.json 2>> gpurun_out/b3_bench.err
B200_GEMM_DIRECT=1 timeout 300 python bench.py --model sup --batch 128 --steps 6 --no-cpu-baseline > gpurun_out/b3_bench_sup_direct.json 2>> gpurun_out/b3_bench.err
B200_GEMM_DIRECT=1 timeout 300 python bench.py --model hac --batch 512 --steps 8 --no-cpu-baseline > gpurun_out/b3_bench_hac_direct.json 2>> gpurun_out/b3_bench.err
timeout 300 python bench.py --model hac --batch 512 --steps 8 --runners 1 --no-cpu-baseline > gpurun_out/b3_bench_hac_r1.json 2>> gpurun_out/b3_bench.err
echo done > gpurun_out/b3_done

#!/bin/bash
# round 2, battery 6: GEMM templated on the activation, cluster MMA issuer polling both groups + staggered start
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests/test_gemm_gpu.py tests/test_forward_gpu.py tests/test_golden.py -m gpu -q -p no:cacheprovider ) > gpurun_out/b6_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/b6_tests.log
for st in 1400 0 2000; do
  echo "== hac 512 stagger=$st" >> gpurun_out/b6_timeline.txt
  B200_CLUSTER_STAGGER=$st timeout 120 python tools/lstm_timeline.py hac 512 2>> gpurun_out/b6_timeline.txt >/dev/null
  B200_CLUSTER_STAGGER=$st timeout 300 python bench.py --model hac --batch 512 --steps 8 --runners 1 --no-cpu-baseline > gpurun_out/b6_bench_hac_r1_st$st.json 2>> gpurun_out/b6_bench.err
done
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/b6_bench_default.json 2> gpurun_out/b6_bench_default.err
B200_GEMM_DIRECT=1 timeout 300 python bench.py --model sup --batch 128 --steps 6 --no-cpu-baseline > gpurun_out/b6_bench_sup_direct.json 2>> gpurun_out/b6_bench.err
B200_GEMM_DIRECT=1 timeout 300 python bench.py --model hac --batch 512 --steps 8 --no-cpu-baseline > gpurun_out/b6_bench_hac_direct.json 2>> gpurun_out/b6_bench.err
echo done > gpurun_out/b6_done

#!/bin/bash
# round 2, battery 9: tcgen05 attention (guarded first: a wrong barrier protocol would hang), full suite, A/B against the mma.sync kernel
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( timeout 240 python -m pytest tests/test_forward_gpu.py -m gpu -q -x -k "tx_model_scores" -p no:cacheprovider ) > gpurun_out/b9_attn_first.log 2>&1
rc=$?
echo "attn first rc=$rc" >> gpurun_out/b9_attn_first.log
if [ $rc -ne 0 ]; then
  export B200_ATTN_LEGACY=1
  echo "tc attention failed; rest of the battery runs the legacy kernel" >> gpurun_out/b9_attn_first.log
  ( timeout 240 python -m pytest tests/test_forward_gpu.py -m gpu -q -x -k "tx_model_scores" -p no:cacheprovider ) > gpurun_out/b9_attn_legacy.log 2>&1
fi
( time timeout 1500 python -m pytest tests -m gpu -q -s -p no:cacheprovider ) > gpurun_out/b9_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/b9_tests.log
timeout 300 python bench.py --model sup --batch 128 --steps 6 --no-cpu-baseline > gpurun_out/b9_bench_sup.json 2> gpurun_out/b9_bench_sup.err
B200_ATTN_LEGACY=1 timeout 300 python bench.py --model sup --batch 128 --steps 6 --no-cpu-baseline > gpurun_out/b9_bench_sup_legacy.json 2>> gpurun_out/b9_bench_sup.err
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/b9_bench_default.json 2> gpurun_out/b9_bench_default.err
echo done > gpurun_out/b9_done

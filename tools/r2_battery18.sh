#!/bin/bash
# round 2, battery 18: decode v3 (instruction diet: sequential bisection, table merge, lean kmer probability, no record warp), conv1+conv2
# with conv2 on tcgen05 (guarded first), full suite, fast / hac lines, conv12 A/B
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_forward_gpu.py -m gpu -q -x -s -k "conv12 or (lstm_model_scores and 1200)" -p no:cacheprovider ) > gpurun_out/b18_conv12_first.log 2>&1
echo "conv12 first rc=$?" >> gpurun_out/b18_conv12_first.log
for sl in 3 4 5; do
  n=512; [ $sl = 5 ] && n=128
  B200_DEBUG_DECODE_TIMES=1 timeout 300 python tools/beam_timeline.py $sl $n > gpurun_out/b18_timeline_sl$sl.txt 2>&1
done
( time timeout 1200 python -m pytest tests -m gpu -q -s -p no:cacheprovider ) > gpurun_out/b18_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/b18_tests.log
timeout 600 python bench.py --no-cpu-baseline --no-sub-models > gpurun_out/b18_fast.json 2> gpurun_out/b18_bench.err
B200_CONV12_FMA=1 timeout 600 python bench.py --no-cpu-baseline --no-sub-models > gpurun_out/b18_fast_conv12_fma.json 2>> gpurun_out/b18_bench.err
timeout 300 python bench.py --model hac --batch 512 --steps 8 --no-cpu-baseline > gpurun_out/b18_hac.json 2>> gpurun_out/b18_bench.err
timeout 300 python bench.py --model sup --batch 128 --steps 6 --no-cpu-baseline > gpurun_out/b18_sup.json 2>> gpurun_out/b18_bench.err
echo done > gpurun_out/b18_done

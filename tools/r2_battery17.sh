#!/bin/bash
# round 2, battery 17: decode v2 (table merge, speculative bisection, record warp, branch-free exp, prefetching traceback on split planes)
# -- bit-exactness suite first, then per-kernel decode times and the in-kernel timeline, new library vs the battery-16 build
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_decode_gpu.py tests/test_golden.py -m gpu -q -x -p no:cacheprovider ) > gpurun_out/b17_tests_decode.log 2>&1
echo "decode tests rc=$?" >> gpurun_out/b17_tests_decode.log
for sl in 3 4 5; do
  n=512; [ $sl = 5 ] && n=128
  B200_DEBUG_DECODE_TIMES=1 timeout 300 python tools/beam_timeline.py $sl $n > gpurun_out/b17_timeline_new_sl$sl.txt 2>&1
  B200CALL_LIB=$PWD/dorado_b200/libb200call_base.so B200_DEBUG_DECODE_TIMES=1 timeout 300 python tools/beam_timeline.py $sl $n > gpurun_out/b17_timeline_base_sl$sl.txt 2>&1
done
( time timeout 1200 python -m pytest tests -m gpu -q -s -p no:cacheprovider ) > gpurun_out/b17_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/b17_tests.log
timeout 600 python bench.py --no-cpu-baseline --no-sub-models > gpurun_out/b17_fast.json 2> gpurun_out/b17_bench.err
timeout 300 python bench.py --model hac --batch 512 --steps 8 --no-cpu-baseline > gpurun_out/b17_hac.json 2>> gpurun_out/b17_bench.err
echo done > gpurun_out/b17_done

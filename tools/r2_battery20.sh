#!/bin/bash
# round 2, battery 20: MUFU.TANH gates / swish everywhere, trimmed exp (no F2I, no upper clamp on max-shifted arguments), min/max score
# clamp -- full suite, then fast / hac / sup lines and runner-count sweeps
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -q -s -p no:cacheprovider ) > gpurun_out/b20_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/b20_tests.log
B200_DEBUG_DECODE_TIMES=1 timeout 300 python tools/beam_timeline.py 3 512 > gpurun_out/b20_timeline_sl3.txt 2>&1
B200_DEBUG_DECODE_TIMES=1 timeout 300 python tools/beam_timeline.py 4 512 > gpurun_out/b20_timeline_sl4.txt 2>&1
F="timeout 300 python bench.py --no-cpu-baseline --no-sub-models"
$F > gpurun_out/b20_fast.json 2> gpurun_out/b20_bench.err
$F --runners 5 > gpurun_out/b20_fast_r5.json 2>> gpurun_out/b20_bench.err
$F --runners 6 > gpurun_out/b20_fast_r6.json 2>> gpurun_out/b20_bench.err
H="timeout 300 python bench.py --model hac --batch 512 --steps 8 --no-cpu-baseline"
$H > gpurun_out/b20_hac.json 2>> gpurun_out/b20_bench.err
$H --runners 3 > gpurun_out/b20_hac_r3.json 2>> gpurun_out/b20_bench.err
$H --runners 5 > gpurun_out/b20_hac_r5.json 2>> gpurun_out/b20_bench.err
timeout 300 python bench.py --model sup --batch 128 --steps 6 --no-cpu-baseline > gpurun_out/b20_sup.json 2>> gpurun_out/b20_bench.err
timeout 300 python bench.py --model sup --batch 128 --steps 6 --no-cpu-baseline --runners 3 > gpurun_out/b20_sup_r3.json 2>> gpurun_out/b20_bench.err
echo done > gpurun_out/b20_done

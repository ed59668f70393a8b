#!/bin/bash
# round 2, battery 19: LSTM gate activations on one MUFU.TANH (build EXTRA=-DB200_LSTM_TANH_APPROX -> libb200call_tanh.so) vs ex2 + rcp:
# score error statistics against the oracles, then fast / hac lines with 8 and 16 chunks per group
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=$PWD/dorado_b200/libb200call_tanh.so
timeout 300 python tools/err_stats.py fast 64 3000 > gpurun_out/b19_err_fast_default.txt 2>&1
B200CALL_LIB=$T timeout 300 python tools/err_stats.py fast 64 3000 > gpurun_out/b19_err_fast_tanh.txt 2>&1
B200CALL_LIB=$T timeout 400 python tools/err_stats.py hac 32 1998 > gpurun_out/b19_err_hac_tanh.txt 2>&1
F="timeout 300 python bench.py --no-cpu-baseline --no-sub-models"
B200CALL_LIB=$T $F > gpurun_out/b19_fast_tanh_nbr8_r4.json 2> gpurun_out/b19_bench.err
B200CALL_LIB=$T B200_LSTM_CHUNKS_PER_CTA=16 $F > gpurun_out/b19_fast_tanh_nbr16_r4.json 2>> gpurun_out/b19_bench.err
B200CALL_LIB=$T B200_LSTM_CHUNKS_PER_CTA=16 $F --runners 8 > gpurun_out/b19_fast_tanh_nbr16_r8.json 2>> gpurun_out/b19_bench.err
B200_LSTM_CHUNKS_PER_CTA=16 $F --runners 8 > gpurun_out/b19_fast_default_nbr16_r8.json 2>> gpurun_out/b19_bench.err
$F --runners 8 > gpurun_out/b19_fast_default_nbr8_r8.json 2>> gpurun_out/b19_bench.err
B200CALL_LIB=$T timeout 300 python bench.py --model hac --batch 512 --steps 8 --no-cpu-baseline > gpurun_out/b19_hac_tanh.json 2>> gpurun_out/b19_bench.err
echo done > gpurun_out/b19_done

#!/bin/bash
# round 2, battery 11: attention with chunk-level masking (parity + bench), fast LSTM groups x chunks-per-group x runners sweep
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_forward_gpu.py tests/test_golden.py tests/test_full_size_gpu.py -m gpu -q -x -k "tx or sup" -p no:cacheprovider ) > gpurun_out/b11_tests_sup.log 2>&1
echo "rc=$?" >> gpurun_out/b11_tests_sup.log
timeout 300 python bench.py --model sup --batch 128 --steps 6 --no-cpu-baseline > gpurun_out/b11_bench_sup.json 2> gpurun_out/b11_bench_sup.err
B="timeout 200 python bench.py --no-cpu-baseline --no-sub-models --steps 20"
for ng in 1 2; do for nbr in 8 16; do for r in 2 4 6 8; do
  B200_LSTM_GROUPS=$ng B200_LSTM_CHUNKS_PER_CTA=$nbr $B --runners $r > gpurun_out/b11_sweep_ng${ng}_nbr${nbr}_r${r}.json 2>> gpurun_out/b11_bench.err
done; done; done
for r in 2 3 4; do
  B200_LSTM_GROUPS=1 B200_LSTM_CHUNKS_PER_CTA=4 $B --runners $r > gpurun_out/b11_sweep_ng1_nbr4_r${r}.json 2>> gpurun_out/b11_bench.err
done
for r in 3 4; do
  timeout 300 python bench.py --model hac --batch 512 --steps 8 --no-cpu-baseline --runners $r > gpurun_out/b11_bench_hac_r${r}.json 2>> gpurun_out/b11_bench.err
done
echo done > gpurun_out/b11_done

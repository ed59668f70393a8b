#!/bin/bash
# round 2, battery 10: two-group fast LSTM kernel -- parity, timeline, A/B against one group, chunks-per-group x runners-in-flight sweep
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_forward_gpu.py tests/test_golden.py tests/test_full_size_gpu.py -m gpu -q -x -k "fast or lstm" -p no:cacheprovider ) > gpurun_out/b10_tests_fast.log 2>&1
echo "rc=$?" >> gpurun_out/b10_tests_fast.log
echo "== fast 512" > gpurun_out/b10_timeline.txt
timeout 120 python tools/lstm_timeline.py fast 512 2>> gpurun_out/b10_timeline.txt >/dev/null
B="timeout 200 python bench.py --no-cpu-baseline --no-sub-models --steps 20"
$B > gpurun_out/b10_bench_ng2.json 2> gpurun_out/b10_bench.err
B200_LSTM_GROUPS=1 $B > gpurun_out/b10_bench_ng1.json 2>> gpurun_out/b10_bench.err
for nbr in 2 4 8; do for r in 2 3 4; do
  B200_LSTM_CHUNKS_PER_CTA=$nbr $B --runners $r > gpurun_out/b10_sweep_nbr${nbr}_r${r}.json 2>> gpurun_out/b10_bench.err
done; done
for n in 1024 2048; do
  $B --batch $n > gpurun_out/b10_bench_batch$n.json 2>> gpurun_out/b10_bench.err
done
echo done > gpurun_out/b10_done

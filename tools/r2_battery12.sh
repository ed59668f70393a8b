#!/bin/bash
# round 2, battery 12: hac cluster recurrence with 64 chunks per cluster -- parity, timeline, chunks-per-cluster x runners sweep; FLSTM test
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_forward_gpu.py tests/test_golden.py tests/test_full_size_gpu.py -m gpu -q -x -k "hac or cluster or variable or flstm or lstm" -p no:cacheprovider ) > gpurun_out/b12_tests_hac.log 2>&1
echo "rc=$?" >> gpurun_out/b12_tests_hac.log
for un in 32 64; do
  echo "== hac 512 chunks/cluster $un" >> gpurun_out/b12_timeline.txt
  B200_CLUSTER_CHUNKS=$un timeout 120 python tools/lstm_timeline.py hac 512 2>> gpurun_out/b12_timeline.txt >/dev/null
done
B="timeout 300 python bench.py --model hac --batch 512 --steps 8 --no-cpu-baseline"
for un in 32 64; do for r in 2 3 4; do
  B200_CLUSTER_CHUNKS=$un $B --runners $r > gpurun_out/b12_hac_un${un}_r${r}.json 2>> gpurun_out/b12_bench.err
done; done
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/b12_bench_default.json 2> gpurun_out/b12_bench_default.err
echo done > gpurun_out/b12_done

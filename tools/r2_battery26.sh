#!/bin/bash
# round 2, battery 26: hac x-projection grid cap next to the co-running recurrences, after the tile-order change (default 116)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
H="timeout 120 python bench.py --model hac --batch 512 --steps 8 --no-cpu-baseline"
for cap in 132 148 100; do
  B200_GEMM_MAX_CTAS=$cap $H > gpurun_out/b26_hac_cap$cap.json 2>> gpurun_out/b26_bench.err
done
echo done > gpurun_out/b26_done

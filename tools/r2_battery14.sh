#!/bin/bash
# round 2, battery 14: default bench line (with auto-batch sub-results), hac batch sweep 64-2048 (BASELINE configs[4], 1 GPU), pool bench on 1 GPU
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 900 python bench.py ) > gpurun_out/b14_bench_default.json 2> gpurun_out/b14_bench_default.err
: > gpurun_out/b14_hac_sweep.jsonl
for b in 64 128 256 512 1024 2048; do
  timeout 300 python bench.py --model hac --batch $b --steps 6 --no-cpu-baseline >> gpurun_out/b14_hac_sweep.jsonl 2>> gpurun_out/b14_bench.err
done
for m in fast hac sup; do
  timeout 300 python tools/bench_pool.py --model $m --gpus 1 > gpurun_out/b14_pool_${m}_g1.json 2>> gpurun_out/b14_bench.err
done
timeout 300 python tools/bench_pool.py --model hac --gpus 1 --batch auto > gpurun_out/b14_pool_hac_auto_g1.json 2>> gpurun_out/b14_bench.err
echo done > gpurun_out/b14_done

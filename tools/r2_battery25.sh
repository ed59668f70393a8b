#!/bin/bash
# round 2, battery 25: the reference's multi-device shape -- one process, one engine per device, NUMA-pinned feeder threads on one
# shared cursor (b200_pool_*), measured end to end from host chunks.  usage: r2_battery25.sh <gpus> ["models"]
cd "$(dirname "$0")/.."
G=${1:-1}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/b25_topo_g$G.txt 2>&1
for m in ${2:-fast hac sup}; do
  timeout 300 python tools/bench_pool.py --model $m --gpus $G > gpurun_out/b25_pool_${m}_g$G.json 2>> gpurun_out/b25_g$G.err
done
echo done > gpurun_out/b25_done_g$G

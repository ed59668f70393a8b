#!/bin/bash
# ncu --set full captures of the sup hot kernels (layer 0: QKV+RoPE, attention, out-proj, FC1+SwiGLU, FC2) -- one process per capture
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
CMD="python bench.py --model sup --batch 128 --steps 1 --warmup 1 --no-cpu-baseline --runners 1"
NCU="ncu --set full --clock-control none --import-source on -f"
# GEMM launches of a step: conv2..conv5 (4), then per layer qkv, out_proj, fc1, fc2
$NCU -k regex:gemm_f16 -s 4 -c 4 -o gpurun_out/r02_gemm_sup_layer0 $CMD > gpurun_out/ncu_gemm_sup.log 2>&1
$NCU -k regex:tx_attention -s 1 -c 1 -o gpurun_out/r02_tx_attention_sup_n128 $CMD > gpurun_out/ncu_attn_sup.log 2>&1
ls -la gpurun_out/*.ncu-rep

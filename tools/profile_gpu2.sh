#!/bin/bash
# r01 profiles for hac and sup (run on the GPU box via gpurun)
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_hac.csv python bench.py --model hac --batch 512 --steps 1 --warmup 1 > gpurun_out/b_hac.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 160 --csv --log-file gpurun_out/launches_sup.csv python bench.py --model sup --batch 128 --steps 1 --warmup 1 > gpurun_out/b_sup.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_f16 -s 12 -c 1 -o gpurun_out/prof_gemm_fc1 -f python bench.py --model sup --batch 128 --steps 1 --warmup 1 > gpurun_out/ncu_gemm.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tx_attention -s 2 -c 1 -o gpurun_out/prof_attn -f python bench.py --model sup --batch 128 --steps 1 --warmup 1 > gpurun_out/ncu_attn.log 2>&1
ls -la gpurun_out | tail -8

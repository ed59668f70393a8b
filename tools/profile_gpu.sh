#!/bin/bash
# Run on the GPU box (via gpurun): launch list + full ncu capture of the two dominant kernels of the fast bench.
set -x
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_fast.csv python bench.py --steps 2 --warmup 1 > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:lstm_layer -s 5 -c 1 -o gpurun_out/prof_lstm -f python bench.py --steps 1 --warmup 1 > gpurun_out/ncu_lstm.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:crf_fwd_beam -s 1 -c 1 -o gpurun_out/prof_fwd_beam -f python bench.py --steps 1 --warmup 1 > gpurun_out/ncu_fwd.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:crf_bwd_scan -s 1 -c 1 -o gpurun_out/prof_bwd -f python bench.py --steps 1 --warmup 1 > gpurun_out/ncu_bwd.log 2>&1
ls -la gpurun_out

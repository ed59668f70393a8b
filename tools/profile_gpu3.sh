#!/bin/bash
# r01 end-of-round profiles (run on the GPU box via gpurun): launch lists with the final kernels, one full capture of
# the weights-in-TMEM cluster LSTM, and the hac / sup bench lines
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 66 --csv --log-file gpurun_out/launches_fast_final.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/b_fast.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 45 --csv --log-file gpurun_out/launches_hac_final.csv python bench.py --model hac --batch 512 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/b_hac.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:lstm_cluster -s 2 -c 1 -o gpurun_out/prof_lstm_cluster_tmem -f python bench.py --model hac --batch 512 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_cluster.log 2>&1
timeout 300 python bench.py --model hac --batch 512 > gpurun_out/bench_hac_final.json 2> gpurun_out/bench_hac_final.err
timeout 400 python bench.py --model sup --batch 128 --steps 10 > gpurun_out/bench_sup_final.json 2> gpurun_out/bench_sup_final.err
ls -la gpurun_out | tail -8
python - <<'PY'
import json
for k in ("hac", "sup"):
    try:
        d = json.loads(open(f"gpurun_out/bench_{k}_final.json").read().strip().splitlines()[-1])
        print(k, f'{d["value"]:.3e}', f'e2e {d["e2e"]["value"]:.3e}', d["ms_per_step"], d["cpu_baseline"], d["clocks"])
    except Exception as e:
        print(k, "failed", e)
PY

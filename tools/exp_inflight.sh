#!/bin/bash
# round-2 experiment: batches in flight x chunks per LSTM CTA (fast@v5, batch 512); run on the GPU box via gpurun
mkdir -p gpurun_out; : > gpurun_out/inflight.jsonl
for nbr in 4 8 16; do
  for r in 1 2 3 4; do
    B200_LSTM_CHUNKS_PER_CTA=$nbr timeout 200 python bench.py --runners $r --steps 24 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 |
      python -c "import sys,json; d=json.loads(sys.stdin.read()); d['nbr']=$nbr; print(json.dumps(d))" >> gpurun_out/inflight.jsonl
  done
done
python - <<'PY'
import json
for l in open("gpurun_out/inflight.jsonl"):
    d = json.loads(l)
    print("chunks/CTA", d["nbr"], "runners", d["config"]["runners_per_gpu"], f'value {d["value"]:.3e}', f'{d["ms_per_step"]:.2f} ms',
          f'e2e {d["e2e"]["value"]:.3e}')
PY

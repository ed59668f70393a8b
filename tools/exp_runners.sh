#!/bin/bash
# runners-in-flight experiment: parity under concurrency, then 1 vs 2 runners per GPU
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_forward_gpu.py -x -q -m gpu -k "concurrent or call_chunks or rejects" 2>&1 | tail -5
: > gpurun_out/runners.jsonl
for spec in "fast 512 1" "fast 512 2" "fast 512 3" "hac 512 1" "hac 512 2" "sup 128 1" "sup 128 2"; do
  set -- $spec
  timeout 200 python bench.py --model $1 --batch $2 --runners $3 --steps 6 --warmup 3 --no-cpu-baseline 2>gpurun_out/err_$1_$3.txt | tail -1 >> gpurun_out/runners.jsonl
done
python - <<'PY'
import json
for l in open("gpurun_out/runners.jsonl"):
    try: d = json.loads(l)
    except Exception: print("bad line", l[:200]); continue
    print(d["config"]["model"], d["config"]["batch_per_gpu"], "runners", d["config"]["runners_per_gpu"], f'value {d["value"]:.3e}', f'{d["ms_per_step"]:.2f} ms', f'e2e {d["e2e"]["value"]:.3e}', f'fwd {d["forward_ms_per_step"]:.2f} dec {d["decode_ms_per_step"]:.2f}', d["clocks"])
PY

#!/bin/bash
# batch-size sweep (BASELINE.json configs 3-5), run on the GPU box; one JSON line per (model, batch)
mkdir -p gpurun_out; : > gpurun_out/sweep.jsonl
if [ $# -eq 0 ]; then set -- "fast 512" "fast 1024" "fast 2048" "fast 4096" "hac 64" "hac 128" "hac 256" "hac 512" "hac 1024" "hac 2048" "sup 32" "sup 128" "sup 256"; fi
for spec in "$@"; do
  set -- $spec
  timeout 300 python bench.py --model $1 --batch $2 --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 >> gpurun_out/sweep.jsonl
done
python - <<'PY'
import json
for l in open("gpurun_out/sweep.jsonl"):
    try: d = json.loads(l)
    except Exception: continue
    print(d["config"]["model"], d["config"]["batch_per_gpu"], f'{d["value"]:.3e}', f'{d["ms_per_step"]:.2f} ms', f'fwd {d["forward_ms_per_step"]:.2f} dec {d["decode_ms_per_step"]:.2f}', d["roofline"]["kernel"], f'{d["roofline"]["frac"]:.3f}',
          {k: v["ms_per_launch"] * v["launches"] for k, v in d["roofline"].get("per_kernel", {}).items()})
PY

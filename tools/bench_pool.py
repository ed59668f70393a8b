#!/usr/bin/env python
"""One process, several GPUs: the reference's multi-device shape (one caller per device, runners fed from shared queues;
dorado/api/runner_creation.cpp:91-123, read_pipeline/nodes/BasecallerNode.cpp:300-352) measured end to end through
b200_pool_call_chunks -- host chunks in, calls out, H2D/D2H inside, NUMA-pinned feeder threads, no collective.

usage: python tools/bench_pool.py --model hac --gpus 2 [--batch 512 | --batch auto] [--runners 2] [--batches-per-runner 6]
Prints one JSON line: samples/s over the whole job (host wall clock of the pool call, which ends when the last result is copied out).
"""
import argparse
import json
import pathlib
import sys
import time

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="hac", choices=["fast", "hac", "sup"])
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--batch", default="512")
    ap.add_argument("--runners", type=int, default=None, help="runners per device (default: 4 fast / hac, 2 sup)")
    ap.add_argument("--chunksize", type=int, default=10000)
    ap.add_argument("--batches-per-runner", type=int, default=6)
    ap.add_argument("--repeats", type=int, default=3)
    args = ap.parse_args()
    from conftest import MODELS, model_dir
    from dorado_b200 import batching
    from dorado_b200.config import load_model_config
    from dorado_b200.runner import B200Caller, B200Pool
    from dorado_b200.weights import synthetic_weights
    cfg = load_model_config(model_dir(args.model))
    w = synthetic_weights(cfg, 42)
    T = cfg.normalise_chunk_size(args.chunksize)
    R = args.runners or {"fast": 4, "hac": 4, "sup": 2}[args.model]
    source = "fixed"
    if args.batch == "auto":
        # CudaCaller::determine_batch_dims: pre-computed B200 table + exact memory cap (80 % of the device like the reference)
        import ctypes as C
        caller = B200Caller(cfg, w, device=0, num_runners=R)
        free = 0.8 * 180e9
        dims, source = batching.determine_batch_dims(caller, MODELS[args.model], args.chunksize, int(free), num_runners=R)
        batch = dims[0][0]
        caller.close()
    else:
        batch = int(args.batch)
    pool = B200Pool(cfg, w, list(range(args.gpus)), R, batch, T)
    n_chunks = batch * R * args.gpus * args.batches_per_runner
    rng = np.random.default_rng(7)
    base = rng.standard_normal((batch, T)).astype(np.float16)
    chunks = np.tile(base, (n_chunks // batch, 1))
    pool.call_chunks(chunks[: batch * R * args.gpus], want_output=False)   # warm-up: every runner once
    best, times = None, []
    for _ in range(args.repeats):
        secs, moves, seq, qs, nb = pool.call_chunks(chunks)
        times.append(secs)
    secs = float(np.median(times))
    info = [pool.runner_info(i) for i in range(pool.num_runners())]
    line = {"metric": "basecalled samples/s", "value": n_chunks * T / secs, "unit": "samples/s", "n_gpus": args.gpus,
            "model": args.model, "batch_per_runner": batch, "batch_source": source, "runners_per_gpu": R, "chunk_samples": T,
            "chunks_per_call": n_chunks, "seconds_per_call": [round(t, 4) for t in times], "timing": "host wall clock of b200_pool_call_chunks (median)",
            "bases_called": int(nb.sum()), "runner_numa_nodes": [i["numa_node"] for i in info],
            "batches_per_runner": [i["batches"] for i in info], "h2d_d2h": "inside the timed region"}
    print(json.dumps(line))
    pool.close()


if __name__ == "__main__":
    main()

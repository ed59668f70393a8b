#!/usr/bin/env python
"""bench.py -- headline metric of BASELINE.json: basecalled samples/s on B200.

One "step" = one pass of the hot path (network forward + CRF decode) over one batch of synthetic chunks.
Workload at N=1: BASELINE.json configs[1] -- dna_r10.4.1 fast@v5 topology, batch 512, chunksize 10000
(normalised to 9996 = 1666 blocks of stride 6), synthetic N(0,1) signal, seeded synthetic weights.
Multi-GPU: one process per GPU (torchrun), each with its own engine replica and its own batch (reads shard
embarrassingly; no collective on the data path) -> weak scaling; time = max over ranks of the device time.

  value  whole-job samples/s with the batches already resident in HBM (CUDA events bracketing the runners' streams);
         `--runners` batches are in flight per GPU (default 2, dorado's num_runners per device), so one batch's decode
         overlaps the next batch's network -- every step is still a full forward + decode of one batch
  e2e    same metric through the C ABI call the adapter makes (b200_runner_call_chunks via
         B200ModelRunner.call_chunks_raw) from pinned host buffers, one host thread per runner:
         H2D of the fp16 batch and D2H of moves/sequence/qstring inside the timed region
  roofline      dominant kernel of the step, timed live per launch with CUDA events
  cpu_baseline  the reference's own CPU path (oracle/_ref, compiled from the reference sources) on a bounded
                sample of the same workload, all host cores (torch intra-op threads = 1 per runner, one runner
                per core, as dorado does)

`--impl reference` times only that CPU path and prints the same JSON line with "impl": "reference".
"""
from __future__ import annotations

import argparse
import json
import os
import pathlib
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

MODELS = {
    "fast": "dna_r10.4.1_e8.2_400bps_fast@v5.0.0",
    "hac": "dna_r10.4.1_e8.2_400bps_hac@v5.0.0",
    "sup": "dna_r10.4.1_e8.2_400bps_sup@v5.0.0",
}
# algorithmic FLOP per input sample
FLOP_PER_SAMPLE = {"fast": 0.1435e6, "hac": 2.139e6, "sup": 14.35e6}


def model_dir(kind):
    return ROOT / "tests" / "data" / "model_configs" / MODELS[kind]


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return dict(hbm_gbs=d["hbm_gbs"], tflops=d["bf16_tflops"], tflops_sustained=d.get("bf16_tflops_sustained"),
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm_gbs=6650.0, tflops=1590.0, tflops_sustained=1400.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu), "-lms", "20"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self, t_begin=None, t_end=None):
        """Summarise the samples whose nvidia-smi timestamp falls inside [t_begin, t_end] (time.time() values)."""
        import datetime
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        out = self._summarise(t_begin, t_end, 0.02)
        if t_begin is not None and out["samples"] < 2:
            # timed region shorter than nvidia-smi's real sampling period: also take the samples of the warm-up
            # (the same kernels, run back to back just before the timed region)
            out = self._summarise(t_begin, t_end, 0.4)
            out["window_widened_s"] = 0.4
        return out

    def _summarise(self, t_begin, t_end, slack):
        import datetime
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                if t_begin is not None:
                    ts = datetime.datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                    if ts < t_begin - slack or ts > t_end + slack:
                        continue
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def run_reference_cpu(kind, chunk_size, budget_chunks_per_core=2, repeats=1):
    """The reference's own CPU path (ModelRunner::call_chunks semantics) on all host cores."""
    from dorado_b200.config import load_model_config
    from dorado_b200.weights import save_b2w, synthetic_weights
    from oracle.oracle import Reference
    import concurrent.futures as cf
    import tempfile
    cfg = load_model_config(model_dir(kind))
    T = cfg.normalise_chunk_size(chunk_size)
    cores = os.cpu_count() or 1
    ref = Reference()
    ref.set_num_threads(1)  # dorado/torch_utils/torch_utils.cpp:20
    with tempfile.TemporaryDirectory() as td:
        wpath = os.path.join(td, "w.b2w")
        save_b2w(wpath, synthetic_weights(cfg, 42))
        handles = [ref.load_model(model_dir(kind), wpath) for _ in range(cores)]
    rng = np.random.default_rng(1234)
    sig = rng.standard_normal((cores, budget_chunks_per_core, T)).astype(np.float32)

    def work(i):
        n = 0
        for c in range(budget_chunks_per_core):
            scores = ref.forward(handles[i], sig[i, c:c + 1])
            ref.decode(scores, q_shift=cfg.qbias, q_scale=cfg.qscale)
            n += T
        return n

    with cf.ThreadPoolExecutor(cores) as ex:
        list(ex.map(work, range(cores)))  # warm-up (also pages libtorch in)
        best = None
        for _ in range(repeats):
            t0 = time.perf_counter()
            total = sum(ex.map(work, range(cores)))
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
    for h in handles:
        ref.free_model(h)
    return dict(value=total / best, unit="samples/s", cores=cores, kind="reference",
                sample=f"{cores} runners x {budget_chunks_per_core} chunks of {T} samples, {kind} topology, "
                       f"forward + CPUDecoder, torch threads=1 per runner", seconds=best, samples=total)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--model", default="fast", choices=list(MODELS))
    ap.add_argument("--batch", type=int, default=512)
    ap.add_argument("--chunksize", type=int, default=10000)
    ap.add_argument("--runners", type=int, default=2,
                    help="runners (batches in flight) per GPU; dorado's default is 2 per device (api/runner_creation.cpp)")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the CPU reference leg (batch sweeps)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    kind = args.model

    from dorado_b200.config import load_model_config
    cfg = load_model_config(model_dir(kind))
    T = cfg.normalise_chunk_size(args.chunksize)
    config = {"workload": f"{MODELS[kind]} topology (synthetic weights), batch {args.batch} per GPU, chunksize "
                          f"{args.chunksize} -> {T} samples/chunk, synthetic N(0,1) fp16 signal",
              "model": kind, "batch_per_gpu": args.batch, "chunk_samples": T, "parallelism": f"replica x{args.gpus}",
              "runners_per_gpu": args.runners,
              "l2": "per-step working set (conv activations + scores > 400 MB) exceeds the 126 MB L2; no explicit flush"}
    metric = "basecalled samples/s"

    if args.impl == "reference":
        if rank != 0:
            return 0
        w = max(0, args.warmup)
        res = run_reference_cpu(kind, args.chunksize, budget_chunks_per_core=1, repeats=max(1, min(args.steps, 3)))
        line = {"impl": "reference", "metric": metric, "value": res["value"], "unit": "samples/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": w, "ms_per_step": res["seconds"] * 1e3, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
                "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": res["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return 0

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from dorado_b200.runner import B200Caller, B200ModelRunner
    from dorado_b200.weights import synthetic_weights
    weights = synthetic_weights(cfg, 42)
    caller = B200Caller(cfg, weights, device=local_rank)
    R = max(1, args.runners)
    runners = [B200ModelRunner(caller, args.batch, args.chunksize) for _ in range(R)]
    runner = runners[0]
    rng = np.random.default_rng(1234 + rank)
    for r in runners:
        r.input_view()[:] = rng.standard_normal((args.batch, T)).astype(np.float16)
    N = args.batch
    samples_per_step = N * T

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident throughput -------------------------------------------------------------
    # step i runs on runner i % R (own stream, own buffers); with R = 2 one batch's decode overlaps the next one's network
    for r in runners:
        r.upload()
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.3)  # let nvidia-smi reach its sampling loop; warm-up and timed region then run back to back
    B200ModelRunner.step_device_runners(runners, N, max(3, args.warmup) * R)
    launches0 = caller.stats()["gpu_launches"]
    barrier()
    t_begin = time.time()
    tot_ms = B200ModelRunner.step_device_runners(runners, N, args.steps)
    barrier()
    t_end = time.time()
    time.sleep(0.05)
    clocks = sampler.stop(t_begin, t_end)
    launches = caller.stats()["gpu_launches"] - launches0
    _, fwd_ms, dec_ms = runner.step_device(N, args.steps)  # un-overlapped stage split, outside the timed region
    tot_ms = max_over_ranks(tot_ms)
    value = world * samples_per_step * args.steps / (tot_ms * 1e-3)

    # ---- end to end through the public API (host buffers) ----------------------------------------
    # one host thread per runner (BasecallerNode drives each runner from its own thread); args.steps calls in total
    import threading
    last = [None] * R

    def drive(i, n_calls):
        for _ in range(n_calls):
            last[i] = runners[i].call_chunks_raw(N)  # the C-ABI call the C++ adapter makes; results land in pinned host memory

    def run_calls(total):
        ths = [threading.Thread(target=drive, args=(i, total // R + (1 if i < total % R else 0))) for i in range(R)]
        for th in ths:
            th.start()
        for th in ths:
            th.join()

    run_calls(max(3, args.warmup) * R)
    barrier()
    t0 = time.perf_counter()
    run_calls(args.steps)
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    bases_last = int(next(c for c in last if c is not None)[3][:N].sum())
    e2e_value = world * samples_per_step * args.steps / e2e_s
    h2d = N * T * 2
    d2h = N * runner.out_len() * 3 + 4 * N

    if rank != 0:
        return 0

    # ---- roofline of the dominant kernel (rank 0, live CUDA events per launch) --------------------
    pk = peaks()
    prof = {}
    for _ in range(3):
        for name, ms in runner.profile(N):
            prof.setdefault(name, []).append(ms)
    agg = {k: (float(np.mean(v)), len(v) // 3) for k, v in prof.items()}  # mean ms per launch, launches per step
    step_ms = sum(m * c for m, c in agg.values())
    dom = max(agg, key=lambda k: agg[k][0] * agg[k][1])
    dom_ms, dom_cnt = agg[dom]
    C, T_out = cfg.outsize, runner.out_len()
    # algorithmic work of ONE launch of each kernel
    work = {}
    if cfg.is_tx_model:
        M = N * (T // cfg.stride_inner())          # transformer tokens in the batch
        d, ff = cfg.tx.d_model, cfg.tx.dim_feedforward
        work = {"qkv_gemm": ("tensor", 2.0 * M * d * 3 * d), "out_proj_gemm": ("tensor", 2.0 * M * d * d),
                "fc1_swiglu_gemm": ("tensor", 2.0 * M * d * 2 * ff), "fc2_gemm": ("tensor", 2.0 * M * ff * d),
                "tx_attention": ("tensor", 4.0 * M * cfg.tx.nhead * 64 * (sum(cfg.tx.attn_window) + 1)),
                "upsample_gemm": ("tensor", 2.0 * M * d * cfg.tx.upsample_scale * d),
                "crf_gemm": ("tensor", 2.0 * M * cfg.tx.upsample_scale * d * C),
                "rmsnorm": ("hbm", 2.0 * M * d * 2)}
    else:
        Cl = cfg.lstm_size
        work = {"lstm_layer": ("tensor", 16.0 * Cl * Cl * T_out * N),   # 2*(2C)*(4C) per chunk-step
                "lstm_rec": ("tensor", 8.0 * Cl * Cl * T_out * N),      # W_hh half; the W_ih half is lstm_gx_gemm
                "lstm_gx_gemm": ("tensor", 8.0 * Cl * Cl * T_out * N),
                "conv3_gemm": ("tensor", 2.0 * cfg.convs[2].winlen * 16 * Cl * T_out * N),
                "linear_gemm": ("tensor", 2.0 * Cl * C * T_out * N),
                "conv12": ("hbm", (2.0 + 32.0) * T * N)}
    for k in ("crf_bwd_scan", "crf_fwd_beam", "crf_traceback"):
        work[k] = ("hbm", (2.0 * C + 3.0) * T_out * N)  # whole-decode algorithmic bytes, charged to each kernel
    bound, amount = work.get(dom, ("tensor", FLOP_PER_SAMPLE[kind] * samples_per_step))
    if bound == "tensor":
        roof = {"kernel": dom, "bound": "tensor", "achieved": amount / (dom_ms * 1e-3) / 1e12, "peak": pk["tflops"],
                "unit": "TFLOP/s", "traffic": None}
    else:
        roof = {"kernel": dom, "bound": "hbm", "achieved": amount / (dom_ms * 1e-3) / 1e9, "peak": pk["hbm_gbs"],
                "unit": "GB/s", "traffic": None}
    roof["per_kernel"] = {k: {"bound": work[k][0], "ms_per_launch": round(agg[k][0], 4), "launches": agg[k][1],
                              "achieved": round(work[k][1] / (agg[k][0] * 1e-3) / (1e12 if work[k][0] == "tensor" else 1e9), 2),
                              "frac": round(work[k][1] / (agg[k][0] * 1e-3) /
                                            ((pk["tflops"] * 1e12) if work[k][0] == "tensor" else (pk["hbm_gbs"] * 1e9)), 4)}
                          for k in agg if k in work}
    roof["frac"] = roof["achieved"] / roof["peak"]
    # DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture of this workload
    # (profiles/*_traffic.json, written by tools/ncu_summary.py); null for workloads that were not captured
    capture = {("fast", 512, "lstm_layer"): "r01_lstm_layer_fast_n512",
               ("hac", 512, "lstm_rec"): "r01_lstm_cluster_hac_n512_tmem",
               ("sup", 128, "fc1_swiglu_gemm"): "r01_gemm_sup_n128"}.get((kind, N, dom))
    try:
        tr = json.loads((ROOT / "profiles" / "r01_traffic.json").read_text())[capture]
        roof["traffic"] = tr["dram_bytes"]
        roof["traffic_source"] = f"profiles/{capture}.ncu-rep ({tr['kernel']})"
    except (OSError, KeyError, ValueError):
        pass
    roof["peak_source"] = pk["source"]
    roof["ms_per_launch"] = dom_ms
    roof["launches_per_step"] = dom_cnt
    roof["share_of_step"] = dom_ms * dom_cnt / step_ms
    roof["kernels_ms"] = {k: round(m * c, 4) for k, (m, c) in agg.items()}
    # whole-forward tensor roofline for context
    roof["forward_tflops"] = FLOP_PER_SAMPLE[kind] * samples_per_step / (fwd_ms / args.steps * 1e-3) / 1e12
    roof["decode_gbs"] = (2.0 * C + 3.0) * T_out * N / (dec_ms / args.steps * 1e-3) / 1e9

    try:
        if args.no_cpu_baseline:
            raise RuntimeError("skipped (--no-cpu-baseline)")
        if world > 1:
            raise RuntimeError("reported at N=1 only")
        cpu = run_reference_cpu(kind, args.chunksize, budget_chunks_per_core=1, repeats=1)
        cpu_baseline = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}
    except Exception as e:  # the checker is optional for the headline number
        cpu_baseline = {"value": None, "unit": "samples/s", "cores": os.cpu_count(), "kind": "reference",
                        "sample": f"unavailable: {e}"}

    line = {"metric": metric, "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": tot_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f16 (fp32 accumulate; fp32 decode)", "data": "synthetic",
            "config": config, "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_s / args.steps * 1e3},
            "gpu_launches": int(launches), "forward_ms_per_step": fwd_ms / args.steps,
            "decode_ms_per_step": dec_ms / args.steps, "roofline": roof, "cpu_baseline": cpu_baseline,
            "bases_called_last_step": bases_last}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())

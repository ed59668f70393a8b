"""Batch-size selection: the B200 counterpart of CudaCaller::determine_batch_dims (dorado/basecall/CudaCaller.cpp:372-632).

  1. memory cap: the largest batch (multiple of the granularity) whose runners fit the memory limit -- from the exact
     arena size of the launch plan instead of the reference's bytes-per-chunk-timestep tables (:323-370);
  2. timing table: ms per chunk for every batch size up to the cap, measured at a short chunk (288 strides, :499-506);
  3. selection: first entry within (1 + penalty) of the best time per chunk, capped by memory (:560-631).
"""
from __future__ import annotations

import ctypes as C
from typing import List, Sequence, Tuple

from . import lib as L


def batch_size_granularity(cfg) -> int:
    """CudaCaller::get_batch_size_granularity (dorado/basecall/include/basecall/CudaCaller.h:60-63): 32 for transformer
    models, 64 for LSTM models -- both are multiples of what the kernels here need (16 for fast, 32 for hac, 1 for sup)."""
    return 32 if cfg.is_tx_model else 64


def select_batch_size(table: Sequence[Tuple[int, float]], max_batch_size: int, granularity: int,
                      time_penalty: float = 0.0) -> int:
    """table: (batch_size, ms per chunk), ascending batch sizes.  Pure host logic in libb200call.so."""
    lib = L.load_library()
    n = len(table)
    bs = (C.c_int32 * max(1, n))(*[int(b) for b, _ in table])
    ms = (C.c_float * max(1, n))(*[float(t) for _, t in table])
    out = C.c_int32()
    L.check(lib.b200_select_batch_size(bs, ms, n, int(max_batch_size), int(granularity), float(time_penalty), C.byref(out)))
    return int(out.value)


def max_batch_size_for_memory(caller, chunk_size: int, mem_limit_bytes: int, granularity: int, num_runners: int = 2,
                              hard_limit: int = 10240) -> int:
    """Largest multiple of `granularity` (<= hard_limit, the reference's benchmarking cap :492) whose `num_runners`
    runners fit mem_limit_bytes; `granularity` if even that does not fit (the reference warns and does the same)."""
    best = granularity
    lo, hi = 1, max(1, hard_limit // granularity)
    while lo <= hi:  # arena bytes grow monotonically with the batch size
        mid = (lo + hi) // 2
        if caller.runner_bytes(mid * granularity, chunk_size) * num_runners <= mem_limit_bytes:
            best = mid * granularity
            lo = mid + 1
        else:
            hi = mid - 1
    return best


def determine_batch_size(caller, chunk_size: int, mem_limit_bytes: int, granularity: int, time_penalty: float = 0.0,
                         num_runners: int = 2, benchmark_limit: int = 2048) -> Tuple[int, List[Tuple[int, float]]]:
    """Returns (batch size, timing table).  benchmark_limit bounds the start-up time like the reference's
    max_batch_size_limit (:489-492)."""
    cfg = caller.cfg
    cap = max_batch_size_for_memory(caller, chunk_size, mem_limit_bytes, granularity, num_runners)
    short_chunk = cfg.normalise_chunk_size(288 * cfg.stride)
    table = caller.benchmark_batch_sizes(short_chunk, granularity, min(cap, benchmark_limit))
    return select_batch_size(table, cap, granularity, time_penalty), table


def gpu_name(caller) -> str:
    buf = C.create_string_buffer(256)
    L.check(L.load_library().b200_engine_gpu_name(caller.handle, buf, len(buf)))
    return buf.value.decode()


def lookup_chunk_benchmarks(gpu: str, model_name: str) -> List[Tuple[int, float]]:
    """CudaChunkBenchmarks::get_chunk_timings (benchmarks/CudaChunkBenchmarks.cpp:24-63): [(batch size, ms per chunk)] or []."""
    lib = L.load_library()
    cap = 512
    bs = (C.c_int32 * cap)()
    ms = (C.c_float * cap)()
    n = C.c_int32()
    L.check(lib.b200_chunk_benchmarks_lookup(gpu.encode(), model_name.encode(), bs, ms, cap, C.byref(n)))
    return [(int(bs[i]), float(ms[i])) for i in range(min(cap, n.value))]


def chunk_size_buckets(cfg, chunk_size: int, overlap: int = 500, pipeline: str = "simplex") -> List[int]:
    """The chunk sizes a caller serves (CudaCaller.cpp:379-414): the requested one and, for high-throughput simplex
    basecalling, half of it for short reads -- each rounded down to the chunk-size granularity and kept above the overlap."""
    gran = cfg.chunk_size_granularity()
    min_chunk = -(-(overlap + 1) // gran) * gran

    def t_out(x):
        return max(min_chunk, (x // gran) * gran) // cfg.stride

    outs = {t_out(chunk_size)}
    if pipeline == "simplex":
        outs.add(t_out(int(chunk_size * 0.5)))
    return [t * cfg.stride for t in sorted(outs, reverse=True)]


def determine_batch_dims(caller, model_name: str, chunk_size: int, mem_limit_bytes: int, time_penalty: float = 0.0,
                         pipeline: str = "simplex", num_runners: int = 2, run_benchmarks: bool = False,
                         benchmark_limit: int = 2048):
    """CudaCaller::determine_batch_dims (CudaCaller.cpp:372-632): [(batch size, chunk size)] for every chunk-size bucket.
    The timing table comes from the pre-computed ones when this GPU and model have one (unless run_benchmarks), else from
    the live loop; the memory cap is exact per bucket."""
    cfg = caller.cfg
    gran = batch_size_granularity(cfg)
    table = [] if run_benchmarks else lookup_chunk_benchmarks(gpu_name(caller), model_name)
    source = "table" if table else "measured"
    dims = []
    for T in chunk_size_buckets(cfg, chunk_size, pipeline=pipeline):
        cap = max_batch_size_for_memory(caller, T, mem_limit_bytes, gran, num_runners)
        if not table:
            short_chunk = cfg.normalise_chunk_size(288 * cfg.stride)
            table = caller.benchmark_batch_sizes(short_chunk, gran, min(cap, benchmark_limit))
        dims.append((select_batch_size(table, cap, gran, time_penalty), T))
    return dims, source

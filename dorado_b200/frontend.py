"""Front end of the hot path through the C ABI: chunk offsets and stitching (host-side integer/byte logic of
libb200call.so; the raw-signal scaling lives on the runner, see B200ModelRunner.accept_raw_chunk).

Mirrors dorado::utils::generate_chunks (read_pipeline/base/chunk.cpp:11-47) and dorado::utils::stitch_chunks
(read_pipeline/base/stitch.cpp:12-96).
"""
from __future__ import annotations

import ctypes as C
from typing import List, Sequence, Tuple

import numpy as np

from . import lib as L


def generate_chunks(num_samples: int, chunk_size: int, stride: int, overlap: int) -> List[int]:
    """Chunk start offsets; raises B200Error (B200_ERR_INVALID) where the reference throws."""
    lib = L.load_library()
    n = C.c_uint64()
    cap = max(1, num_samples // max(1, chunk_size - overlap) + 2) if chunk_size > overlap else 1
    buf = (C.c_uint64 * cap)()
    L.check(lib.b200_generate_chunks(num_samples, chunk_size, stride, overlap, buf, cap, C.byref(n)))
    if n.value > cap:  # cannot happen with the bound above; re-query rather than truncate
        cap = n.value
        buf = (C.c_uint64 * cap)()
        L.check(lib.b200_generate_chunks(num_samples, chunk_size, stride, overlap, buf, cap, C.byref(n)))
    return [int(buf[i]) for i in range(n.value)]


def generate_variable_chunks(num_samples: int, chunk_size: int, stride: int, overlap: int) -> List[Tuple[int, int]]:
    """[first, second) intervals of the variable-chunk-size mode (utils::generate_variable_chunks, chunk.cpp:49-113)."""
    lib = L.load_library()
    n = C.c_uint64()
    L.check(lib.b200_generate_variable_chunks(num_samples, chunk_size, stride, overlap, None, 0, C.byref(n)))
    buf = (C.c_uint64 * (2 * n.value))()
    L.check(lib.b200_generate_variable_chunks(num_samples, chunk_size, stride, overlap, buf, n.value, C.byref(n)))
    return [(int(buf[2 * i]), int(buf[2 * i + 1])) for i in range(n.value)]


def stitch_chunks(chunks: Sequence[Tuple[int, int, np.ndarray, str, str]], raw_samples: int, stride: int):
    """chunks: (input_offset, raw_chunk_size, moves uint8 [T_out], sequence, qstring) per called chunk, in read order.
    Returns (sequence, qstring, moves) of the stitched read."""
    lib = L.load_library()
    n = len(chunks)
    arr = (L.CalledChunk * max(1, n))()
    keep = []  # buffers must outlive the call
    tot_m = tot_b = 0
    for i, (off, size, moves, seq, qs) in enumerate(chunks):
        m = np.ascontiguousarray(moves, np.uint8)
        sb, qb = seq.encode("ascii"), qs.encode("ascii")
        if len(sb) != len(qb):
            raise ValueError("sequence and qstring lengths differ")
        sbuf, qbuf = C.create_string_buffer(sb, len(sb) + 1), C.create_string_buffer(qb, len(qb) + 1)
        keep += [m, sbuf, qbuf]
        arr[i] = L.CalledChunk(int(off), int(size), m.ctypes.data, m.size, C.addressof(sbuf), C.addressof(qbuf), len(sb))
        tot_m += m.size
        tot_b += len(sb)
    mo = np.zeros(max(1, tot_m), np.uint8)
    so = np.zeros(max(1, tot_b), np.uint8)
    qo = np.zeros(max(1, tot_b), np.uint8)
    nm, nb = C.c_uint64(), C.c_uint64()
    L.check(lib.b200_stitch_chunks(arr, n, int(raw_samples), int(stride), mo.ctypes.data, so.ctypes.data, qo.ctypes.data,
                                   C.byref(nm), C.byref(nb)))
    return bytes(so[: nb.value]).decode("ascii"), bytes(qo[: nb.value]).decode("ascii"), mo[: nm.value].copy()

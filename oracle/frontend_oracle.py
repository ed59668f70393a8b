"""TEST INFRASTRUCTURE ONLY: numpy restatement of the front end of the hot path (never imported by the product).

* chunk_input       what one model-input row holds in the reference: ScalerNode's
                    shift_scale_tensor_i16_to_f16_inplace (dorado/torch_utils/tensor_utils.cpp:100-143,
                    ScalerNode.cpp:226-229) followed by BasecallerNode's slice + repeat-padding
                    (dorado/read_pipeline/nodes/BasecallerNode.cpp:395-440)
* generate_chunks   dorado/read_pipeline/base/chunk.cpp:11-47
* stitch_chunks     dorado/read_pipeline/base/stitch.cpp:12-96

Pinned against the compiled reference (oracle/_ref, ref_driver.cpp: ref_make_chunk_input / ref_generate_chunks /
ref_stitch_chunks) and the golden vectors of tests/ChunkTest.cpp, tests/StitchTest.cpp, tests/TensorUtilsTest.cpp:121
by tests/test_frontend_cpu.py.
"""
from __future__ import annotations

import numpy as np


def scale_i16_to_f16(raw: np.ndarray, shift: float, scale: float) -> np.ndarray:
    """fp16((float(x) - shift) / scale): IEEE fp32 subtract and divide, round-to-nearest-even to half."""
    x = np.asarray(raw, np.int16).astype(np.float32)
    return ((x - np.float32(shift)) / np.float32(scale)).astype(np.float16)


def chunk_input(raw: np.ndarray, input_offset: int, chunk_size: int, shift: float, scale: float) -> np.ndarray:
    """fp16 [chunk_size]: scaled slice raw[offset : offset + chunk_size] (clamped at the read end), repeated to fill."""
    sl = scale_i16_to_f16(np.asarray(raw, np.int16)[input_offset: input_offset + chunk_size], shift, scale)
    if sl.size == 0:
        raise ValueError("chunk starts beyond the read")
    if sl.size == chunk_size:
        return sl
    n, overhang = divmod(chunk_size, sl.size)
    return np.concatenate([np.tile(sl, n), sl[:overhang]])


def generate_chunks(num_samples: int, chunk_size: int, stride: int, overlap: int):
    if num_samples == 0:
        raise RuntimeError("empty read")
    if stride == 0:
        raise ValueError("invalid stride")
    if chunk_size == 0 or chunk_size % stride or chunk_size <= overlap:
        raise ValueError("invalid chunk size")
    if overlap % stride:
        raise ValueError("invalid overlap")
    offsets = [0]
    offset = 0
    last_offset = num_samples - chunk_size if num_samples > chunk_size else 0
    if last_offset % stride:
        last_offset += stride - last_offset % stride
    step = chunk_size - overlap
    while offset + chunk_size < num_samples:
        offset = min(offset + step, last_offset)
        offsets.append(offset)
    return offsets


def generate_variable_chunks(num_samples: int, chunk_size: int, stride: int, overlap: int):
    """dorado/read_pipeline/base/chunk.cpp:49-113"""
    import math
    if num_samples == 0:
        raise RuntimeError("empty read")
    if stride == 0:
        raise ValueError("invalid stride")
    if chunk_size == 0 or chunk_size % stride or chunk_size == stride or chunk_size <= overlap:
        raise ValueError("invalid chunk size")
    if overlap % stride or (stride != 1 and overlap == 0):
        raise ValueError("invalid overlap")
    num_chunks = 1 + (math.ceil((num_samples - chunk_size) / float(chunk_size - overlap)) if num_samples > chunk_size else 0)
    with_overlaps = num_samples + (num_chunks - 1) * overlap
    longer, adjusted = with_overlaps % num_chunks, with_overlaps // num_chunks
    iv, start = [], 0
    for i in range(num_chunks):
        iv.append([start, start + adjusted + (1 if i < longer else 0)])
        start = iv[-1][1] - overlap
    for i in range(1, num_chunks):
        if iv[i][0] % stride:
            iv[i][0] += stride - iv[i][0] % stride
    for i in range(num_chunks - 1):
        iv[i][1] -= iv[i][1] % stride
    return [tuple(x) for x in iv]


def stitch_chunks(chunks, raw_samples: int, stride: int):
    """chunks: list of (input_offset, raw_chunk_size, moves, sequence, qstring); returns (sequence, qstring, moves)."""
    start_pos, mid_front = 0, 0
    moves, seqs, qs = [], [], []
    for i in range(len(chunks) - 1):
        off, size, mv, sq, q = chunks[i]
        mv = list(mv)
        overlap = (size + off - chunks[i + 1][0]) // stride
        mid_rear = overlap // 2
        trim = sum(mv[len(mv) - mid_rear:]) if mid_rear else 0
        end_pos = len(sq) - trim
        seqs.append(sq[start_pos:end_pos])
        qs.append(q[start_pos:end_pos])
        moves += mv[mid_front: len(mv) - mid_rear]
        mid_front = overlap - mid_rear
        start_pos = sum(list(chunks[i + 1][2])[:mid_front])
    off, size, mv, sq, q = chunks[-1]
    moves += list(mv)[mid_front:]
    if len(chunks) == 1:
        moves = moves[: raw_samples // stride]
        end = sum(moves)
        seqs.append(sq[start_pos: start_pos + end])
        qs.append(q[start_pos: start_pos + end])
    else:
        seqs.append(sq[start_pos:])
        qs.append(q[start_pos:])
    seq, qstr = "".join(seqs), "".join(qs)
    if len(moves) > raw_samples // stride:
        if moves[-1] == 1:
            seq, qstr = seq[:-1], qstr[:-1]
        moves.pop()
    return seq, qstr, np.array(moves, np.uint8)

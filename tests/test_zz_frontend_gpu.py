"""GPU parity of the front end: raw int16 chunks scaled, sliced and repeat-padded on the device
must give bit-identical model-input rows to what the reference's ScalerNode + BasecallerNode produce on the host
(golden fixture generated from the compiled reference, and the numpy oracle), and therefore identical calls.
(The file name keeps these tests after the hot-path parity tests in collection order.)
"""
import numpy as np
import pytest

from conftest import ROOT, model_dir
from oracle import frontend_oracle as fo

pytestmark = pytest.mark.gpu


def _caller(kind="fast"):
    from dorado_b200.config import load_model_config
    from dorado_b200.runner import B200Caller
    from dorado_b200.weights import synthetic_weights
    cfg = load_model_config(model_dir(kind))
    return cfg, B200Caller(cfg, synthetic_weights(cfg, 42))


def test_raw_chunk_rows_match_reference_fixture():
    from dorado_b200.frontend import generate_chunks
    from dorado_b200.runner import B200ModelRunner
    g = np.load(ROOT / "tests" / "golden" / "frontend.npz")
    chunk, stride, overlap = int(g["chunk_size"]), int(g["stride"]), int(g["overlap"])
    cfg, caller = _caller()
    runner = B200ModelRunner(caller, 32, chunk)
    assert runner.chunk_size() == chunk
    slot = 0
    for r, n in enumerate(g["read_lens"]):
        shift, scale = g[f"shift_scale_{r}"]
        for o in generate_chunks(int(n), chunk, stride, overlap):
            runner.accept_raw_chunk(slot, g[f"raw_{r}"], o, float(shift), float(scale))
            slot += 1
    want = g["input_rows_f16_bits"]
    assert slot == want.shape[0]
    got = runner.debug_read_input(slot).view(np.uint16)
    assert (got == want).all(), f"{int((got != want).sum())} input samples differ from the reference's rows"


@pytest.mark.parametrize("kind,N,T", [("fast", 48, 3000), ("hac", 32, 1200)])
def test_raw_chunks_call_like_host_scaled_chunks(kind, N, T):
    """Same reads fed (a) as fp16 chunks prepared the reference's way on the host, (b) raw, (c) half and half:
    identical input rows, identical sequence / qstring / moves."""
    from dorado_b200.runner import B200ModelRunner
    cfg, caller = _caller(kind)
    ra, rb, rc = (B200ModelRunner(caller, N, T) for _ in range(3))
    Tn = ra.chunk_size()
    rng = np.random.default_rng(3)
    rows = []
    for i in range(N):
        n = int(rng.integers(Tn // 7, 3 * Tn))              # short reads, ragged tails and interior chunks
        raw = np.clip(rng.standard_normal(n) * 170 + 400, -32768, 32767).astype(np.int16)
        offs = fo.generate_chunks(n, Tn, cfg.stride, cfg.stride * 20)
        o = offs[int(rng.integers(0, len(offs)))]
        shift, scale = float(rng.uniform(380, 420)), float(rng.uniform(150, 190))
        row = fo.chunk_input(raw, o, Tn, shift, scale)
        rows.append(row)
        ra.accept_chunk(i, row)
        rb.accept_raw_chunk(i, raw, o, shift, scale)
        if i % 2:
            rc.accept_raw_chunk(i, raw, o, shift, scale)
        else:
            rc.accept_raw_chunk(i, raw[::-1].copy(), 0, 1.0, 1.0)   # then overwritten by an fp16 chunk: slot reverts
            rc.accept_chunk(i, row)
    want_rows = np.stack(rows).view(np.uint16)
    for r in (rb, rc):
        assert (r.debug_read_input(N).view(np.uint16) == want_rows).all()
    base = [(c.sequence, c.qstring, bytes(c.moves)) for c in ra.call_chunks(N)]
    assert sum(len(s) for s, _, _ in base) > N * 10
    for r in (rb, rc):
        assert [(c.sequence, c.qstring, bytes(c.moves)) for c in r.call_chunks(N)] == base
    # partial batch through the raw path
    k = N // 3
    assert [(c.sequence, c.qstring, bytes(c.moves)) for c in rb.call_chunks(k)] == base[:k]


def test_read_to_stitched_call_roundtrip():
    """One long raw read: chunk (C ABI) -> raw accept -> call_chunks -> stitch (C ABI).  The stitched moves cover
    floor(samples / stride) blocks, every base has a move, and the result equals the oracle stitching the same calls."""
    from dorado_b200.frontend import generate_chunks, stitch_chunks
    from dorado_b200.runner import B200ModelRunner
    cfg, caller = _caller("fast")
    T, overlap = 1200, 120
    n = 11 * T + 517
    rng = np.random.default_rng(9)
    raw = np.clip(rng.standard_normal(n) * 170 + 400, -32768, 32767).astype(np.int16)
    offs = generate_chunks(n, T, cfg.stride, overlap)
    runner = B200ModelRunner(caller, 16, T)
    assert len(offs) <= 16
    for i, o in enumerate(offs):
        runner.accept_raw_chunk(i, raw, o, 400.0, 170.0)
    called = runner.call_chunks(len(offs))
    chunks = [(o, T, c.moves, c.sequence, c.qstring) for o, c in zip(offs, called)]
    seq, qstr, moves = stitch_chunks(chunks, n, cfg.stride)
    assert len(moves) == n // cfg.stride
    assert int(moves.sum()) == len(seq) == len(qstr) > 100
    want = fo.stitch_chunks(chunks, n, cfg.stride)
    assert (seq, qstr) == want[:2] and moves.tolist() == want[2].tolist()


def test_raw_chunk_argument_errors():
    from dorado_b200 import lib as L
    from dorado_b200.runner import B200ModelRunner
    cfg, caller = _caller()
    r = B200ModelRunner(caller, 16, 1200)
    raw = np.zeros(100, np.int16)
    with pytest.raises(L.B200Error):
        r.accept_raw_chunk(0, raw, 100, 0.0, 1.0)     # offset beyond the read
    with pytest.raises(L.B200Error):
        r.accept_raw_chunk(0, raw, 0, 0.0, 0.0)       # scale 0
    with pytest.raises(L.B200Error):
        r.accept_raw_chunk(16, raw, 0, 0.0, 1.0)      # slot out of range
    with pytest.raises(L.B200Error):
        r.accept_raw_chunk(0, np.zeros(0, np.int16), 0, 0.0, 1.0)


def test_runner_bytes_and_batch_size_selection():
    """b200_engine_runner_bytes is exactly what a runner allocates; the determine_batch_dims counterpart returns a batch
    size from its own timing table that respects the memory cap."""
    from dorado_b200.batching import determine_batch_size, max_batch_size_for_memory, select_batch_size
    from dorado_b200.runner import B200ModelRunner
    cfg, caller = _caller("fast")
    before = caller.stats()["arena_bytes"]
    r = B200ModelRunner(caller, 64, 1200)
    assert caller.stats()["arena_bytes"] - before == caller.runner_bytes(64, 1200)
    r.close()
    assert caller.stats()["arena_bytes"] == before
    assert caller.runner_bytes(128, 1200) > caller.runner_bytes(64, 1200)
    limit = 3 * caller.runner_bytes(128, 6000)
    cap = max_batch_size_for_memory(caller, 6000, limit, 64, num_runners=2)
    assert cap % 64 == 0 and caller.runner_bytes(cap, 6000) * 2 <= limit < caller.runner_bytes(cap + 64, 6000) * 2
    chosen, table = determine_batch_size(caller, 6000, limit, 64, time_penalty=0.05, num_runners=2, benchmark_limit=256)
    assert [b for b, _ in table] == list(range(64, min(cap, 256) + 1, 64)) and all(t > 0 for _, t in table)
    assert chosen == select_batch_size(table, cap, 64, 0.05) and 64 <= chosen <= cap

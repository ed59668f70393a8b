"""Generate tests/golden/full_*.npz and tests/golden/decode_full.npz: FULL-LENGTH chunks (the sizes bench.py times)
through the reference's own CPU code (oracle/_ref/libdorado_ref.so = the unmodified reference sources, oracle/Makefile).

Run where /root/reference exists:  python tools/make_golden_full.py [fast hac sup decode]

full_<kind>.npz   M chunks of T = normalise(10000) samples (fast/hac 9996, sup 9984), signal regenerated from a seed:
    ref_*      reference fp32 forward (CRFModel / TxModel, libtorch CPU) + CPUDecoder::beam_search_part_2 on ITS scores
               -- what `dorado basecaller --device cpu` emits for these chunks: sequence, qstring, moves (bit-packed)
    ref16_*    the same decoder on those scores rounded to fp16 (what a CUDA decoder is handed)
    rows       score rows (every ROW_STEP-th block) of the first ROW_CHUNKS chunks, reference fp32
    sha        sha256 of the signal bytes and of the complete fp32 score tensor
decode_full.npz   synthetic fp16 scores of full length for state_len 3/4/5 (regenerated from a seed by
    tests/conftest.synthetic_scores: i.i.d., and planted_scores: a planted best path, i.e. well conditioned) with the
    reference CPUDecoder's output on exactly those scores.
"""
import concurrent.futures as cf
import hashlib
import os
import pathlib
import sys
import tempfile
import time

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from conftest import model_dir, planted_scores, synthetic_scores  # noqa: E402
from dorado_b200.config import load_model_config  # noqa: E402
from dorado_b200.weights import save_b2w, synthetic_weights  # noqa: E402
from oracle.oracle import Reference  # noqa: E402

OUT = ROOT / "tests" / "golden"
M_CHUNKS = {"fast": 128, "hac": 128, "sup": 64}
SIGNAL_SEED = 20260923
ROW_CHUNKS, ROW_STEP = 16, 104      # 16 chunks x 17 rows of scores
DECODE_FULL = {3: (8, 1666), 4: (4, 1666), 5: (2, 1664)}


def full_signal(kind, M, T):
    """The fixture's chunks: i.i.d. N(0,1) fp16, one generator per model (tests regenerate it from the seed)."""
    return np.random.default_rng(SIGNAL_SEED + len(kind)).standard_normal((M, T)).astype(np.float16)


def pack_rows(buf, n_bases):
    """Concatenate the first n_bases[i] bytes of every row."""
    return np.concatenate([buf[i, : n_bases[i]] for i in range(buf.shape[0])]) if len(n_bases) else np.zeros(0, np.uint8)


def make_model(kind):
    ref = Reference()
    ref.set_num_threads(1)
    cfg = load_model_config(model_dir(kind))
    T = cfg.normalise_chunk_size(10000)
    M = M_CHUNKS[kind]
    sig = full_signal(kind, M, T)
    workers = min(os.cpu_count() or 1, 8)
    with tempfile.TemporaryDirectory() as td:
        save_b2w(f"{td}/w.b2w", synthetic_weights(cfg, 42))
        handles = [ref.load_model(model_dir(kind), f"{td}/w.b2w") for _ in range(workers)]
    T_out = T // cfg.stride
    res = {k: np.zeros((M, T_out), np.uint8) for k in ("seq", "qstr", "moves", "seq16", "qstr16", "moves16")}
    nb = np.zeros(M, np.int32)
    nb16 = np.zeros(M, np.int32)
    rows = np.zeros((ROW_CHUNKS, len(range(0, T_out, ROW_STEP)), cfg.outsize), np.float32)
    digests = [None] * M
    t0 = time.time()

    import queue
    free = queue.Queue()
    for h in handles:
        free.put(h)

    def work(i):
        h = free.get()                                                           # one model replica per job in flight
        try:
            s32 = ref.forward(h, sig[i:i + 1].astype(np.float32))                # [1, T_out, C] fp32 (Clamp applied inside)
        finally:
            free.put(h)
        r = ref.decode(s32, q_shift=cfg.qbias, q_scale=cfg.qscale)
        r16 = ref.decode(s32.astype(np.float16).astype(np.float32), q_shift=cfg.qbias, q_scale=cfg.qscale)
        res["seq"][i], res["qstr"][i], res["moves"][i], nb[i] = r.seq_buf[0], r.qstr_buf[0], r.moves[0], r.n_bases[0]
        res["seq16"][i], res["qstr16"][i], res["moves16"][i], nb16[i] = r16.seq_buf[0], r16.qstr_buf[0], r16.moves[0], r16.n_bases[0]
        if i < ROW_CHUNKS:
            rows[i] = s32[0, ::ROW_STEP]
        digests[i] = hashlib.sha256(s32.tobytes()).hexdigest()
        return i

    with cf.ThreadPoolExecutor(workers) as ex:
        futs = [ex.submit(work, i) for i in range(M)]
        for f in futs:
            f.result()
    for h in handles:
        ref.free_model(h)
    sha_scores = hashlib.sha256("".join(digests).encode()).hexdigest()
    np.savez_compressed(
        OUT / f"full_{kind}.npz",
        M=M, T=T, T_out=T_out, signal_seed=SIGNAL_SEED + len(kind), weights_seed=42, row_chunks=ROW_CHUNKS, row_step=ROW_STEP,
        sha_signal=hashlib.sha256(sig.tobytes()).hexdigest(), sha_scores=sha_scores,
        ref_n_bases=nb, ref_seq=pack_rows(res["seq"], nb), ref_qstr=pack_rows(res["qstr"], nb),
        ref_moves=np.packbits(res["moves"], axis=1),
        ref16_n_bases=nb16, ref16_seq=pack_rows(res["seq16"], nb16), ref16_qstr=pack_rows(res["qstr16"], nb16),
        ref16_moves=np.packbits(res["moves16"], axis=1),
        rows=rows)
    same16 = sum(bytes(res["seq"][i, :nb[i]]) == bytes(res["seq16"][i, :nb16[i]]) for i in range(M))
    print(f"{kind}: {M} chunks x {T} samples, {int(nb.sum())} bases, fp32-vs-fp16-scores identical sequences {same16}/{M}, "
          f"{time.time() - t0:.0f} s, file {os.path.getsize(OUT / f'full_{kind}.npz') / 1e3:.0f} kB")


def make_decode_full():
    ref = Reference()
    ref.set_num_threads(1)
    out = {}
    for sl, (N, T) in DECODE_FULL.items():
        s16 = synthetic_scores(N, T, sl, seed=900 + sl, scale=1.5)
        s = np.clip(s16.astype(np.float32), -5, 5)
        r = ref.decode(s, q_shift=-0.3, q_scale=0.95)
        out[f"sl{sl}_shape"] = np.array([N, T])
        out[f"sl{sl}_n_bases"] = r.n_bases
        out[f"sl{sl}_seq"] = pack_rows(r.seq_buf, r.n_bases)
        out[f"sl{sl}_qstr"] = pack_rows(r.qstr_buf, r.n_bases)
        out[f"sl{sl}_moves"] = np.packbits(r.moves, axis=1)
        print(f"decode_full state_len {sl}: {N} x {T}, {int(r.n_bases.sum())} bases")
        p16 = planted_scores(N, T, sl, seed=950 + sl)
        r = ref.decode(np.clip(p16.astype(np.float32), -5, 5), q_shift=-0.3, q_scale=0.95)
        out[f"pl{sl}_n_bases"] = r.n_bases
        out[f"pl{sl}_seq"] = pack_rows(r.seq_buf, r.n_bases)
        out[f"pl{sl}_qstr"] = pack_rows(r.qstr_buf, r.n_bases)
        out[f"pl{sl}_moves"] = np.packbits(r.moves, axis=1)
        print(f"decode_full planted state_len {sl}: {int(r.n_bases.sum())} bases, qstring head {bytes(r.qstr_buf[0, :40])}")
    np.savez_compressed(OUT / "decode_full.npz", q_shift=-0.3, q_scale=0.95, clamp=5.0, seed_base=900, planted_seed_base=950, scale=1.5, **out)


if __name__ == "__main__":
    OUT.mkdir(exist_ok=True)
    what = sys.argv[1:] or ["fast", "hac", "sup", "decode"]
    for k in what:
        if k == "decode":
            make_decode_full()
        else:
            make_model(k)

"""Generate tests/golden/*.npz: small fixed inputs with the outputs of (a) the reference's own CPU code
(oracle/_ref/libdorado_ref.so, built from /root/reference by oracle/Makefile) and (b) the C oracle.

Run where /root/reference exists:  python tools/make_golden.py
The fixtures let the GPU box (no /root/reference) check the oracle and the engine against reference outputs.
"""
import pathlib
import sys
import tempfile

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from conftest import model_dir  # noqa: E402
from dorado_b200.config import load_model_config  # noqa: E402
from dorado_b200.weights import save_b2w, synthetic_weights  # noqa: E402
from oracle.oracle import CrfOracle, Reference  # noqa: E402

OUT = ROOT / "tests" / "golden"
SHAPES = {"fast": (3, 1200), "hac": (2, 900), "sup": (2, 768)}


def main():
    ref, orc = Reference(), CrfOracle()
    OUT.mkdir(exist_ok=True)
    for kind, (N, T) in SHAPES.items():
        cfg = load_model_config(model_dir(kind))
        w = synthetic_weights(cfg, 42)
        with tempfile.TemporaryDirectory() as td:
            save_b2w(f"{td}/w.b2w", w)
            h = ref.load_model(model_dir(kind), f"{td}/w.b2w")
        sig = np.random.default_rng(2024).standard_normal((N, cfg.normalise_chunk_size(T))).astype(np.float16)
        scores32 = ref.forward(h, sig.astype(np.float32))           # reference CPU forward (fp32)
        scores16 = scores32.astype(np.float16)                      # what a CUDA decoder is handed
        r = ref.decode(scores16.astype(np.float32), q_shift=cfg.qbias, q_scale=cfg.qscale)  # reference CPUDecoder
        o = orc.decode(scores16, clamp_val=5.0 if cfg.clamp else 0.0, q_shift=cfg.qbias, q_scale=cfg.qscale)
        f, b, p = ref.scans(scores16[0].astype(np.float32))
        keep = slice(0, None, max(1, f.shape[0] // 8))              # a few rows of the guides keep the file small
        np.savez_compressed(
            OUT / f"decode_{kind}.npz", signal=sig, scores_f16=scores16,
            ref_seq=r.seq_buf, ref_qstr=r.qstr_buf, ref_moves=r.moves, ref_n_bases=r.n_bases,
            oracle_seq=o.seq_buf, oracle_qstr=o.qstr_buf, oracle_moves=o.moves, oracle_n_bases=o.n_bases,
            ref_fwd_rows=f[keep], ref_bwd_rows=b[keep], ref_posts_rows=p[keep], guide_row_step=max(1, f.shape[0] // 8),
            weights_seed=42, signal_seed=2024)
        qm = sum(a != b_ for s1, s2 in zip(r.qstrings, o.qstrings) for a, b_ in zip(s1, s2))
        print(kind, "scores", scores16.shape, "bases", r.n_bases.tolist(), "seq equal", r.sequences == o.sequences,
              "moves equal", bool((r.moves == o.moves).all()), "qstring mismatches", qm, "of", int(r.n_bases.sum()))
        ref.free_model(h)


if __name__ == "__main__":
    main()

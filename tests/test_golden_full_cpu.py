"""CPU checks of the full-length fixtures (tests/golden/full_*.npz, decode_full.npz; tools/make_golden_full.py).

The fixtures hold the outputs of the unmodified reference CPU sources on full-length chunks (T = 9996 / 9984 samples,
1666 / 1664 blocks).  Here the CPU oracle is held to them; tests/test_full_size_gpu.py holds the engine to both.
"""
import hashlib
import pathlib

import numpy as np
import pytest

from conftest import model_dir, planted_scores, synthetic_scores, unpack_rows

GOLD = pathlib.Path(__file__).resolve().parent / "golden"


@pytest.mark.parametrize("state_len", [3, 4, 5])
def test_oracle_full_length_decode_matches_reference(crf_oracle, state_len):
    g = np.load(GOLD / "decode_full.npz")
    N, T = (int(v) for v in g[f"sl{state_len}_shape"])
    for tag, s16 in (("sl", synthetic_scores(N, T, state_len, seed=int(g["seed_base"]) + state_len, scale=float(g["scale"]))),
                     ("pl", planted_scores(N, T, state_len, seed=int(g["planted_seed_base"]) + state_len))):
        o = crf_oracle.decode(s16, clamp_val=float(g["clamp"]), q_shift=float(g["q_shift"]), q_scale=float(g["q_scale"]))
        r_nb = g[f"{tag}{state_len}_n_bases"]
        r_seq, r_q = unpack_rows(g[f"{tag}{state_len}_seq"], r_nb), unpack_rows(g[f"{tag}{state_len}_qstr"], r_nb)
        r_mv = np.unpackbits(g[f"{tag}{state_len}_moves"], axis=1)[:, :T]
        same = [o.sequences[i].encode() == r_seq[i] for i in range(N)]
        q_bad = sum(sum(x != y for x, y in zip(o.qstrings[i].encode(), r_q[i])) for i in range(N) if same[i])
        q_tot = sum(len(r_q[i]) for i in range(N) if same[i])
        if tag == "pl":
            # well-conditioned scores: every base and move equals the reference's; the qstring differs only through the
            # last ulp of libtorch's vectorised scans (0.5 % of the characters of these confident calls, by one step)
            assert all(same) and (o.moves == r_mv).all()
            assert q_bad <= 0.015 * q_tot
            worst = max((abs(x - y) for i in range(N) for x, y in zip(o.qstrings[i].encode(), r_q[i])), default=0)
            assert worst <= 1
        else:
            assert sum(same) >= N // 2
            assert q_bad <= 0.002 * q_tot + 1


@pytest.mark.parametrize("kind", ["fast", "hac", "sup"])
def test_full_fixture_is_reproducible_from_its_seed(kind):
    from dorado_b200.config import load_model_config
    g = np.load(GOLD / f"full_{kind}.npz")
    cfg = load_model_config(model_dir(kind))
    M, T = int(g["M"]), int(g["T"])
    assert T == cfg.normalise_chunk_size(10000) and int(g["T_out"]) == T // cfg.stride
    sig = np.random.default_rng(int(g["signal_seed"])).standard_normal((M, T)).astype(np.float16)
    assert hashlib.sha256(sig.tobytes()).hexdigest() == str(g["sha_signal"])
    nb = g["ref_n_bases"]
    assert int(nb.sum()) == g["ref_seq"].size == g["ref_qstr"].size
    mv = np.unpackbits(g["ref_moves"], axis=1)[:, : int(g["T_out"])]
    np.testing.assert_array_equal(mv.sum(axis=1), nb)
    assert set(np.unique(g["ref_seq"]).tolist()) <= set(b"ACGT")
    assert g["ref_qstr"].min() >= 34 and g["ref_qstr"].max() <= 83


@pytest.mark.parametrize("kind", ["fast", "hac"])
def test_numpy_forward_matches_full_length_reference_rows(kind):
    """oracle/nn_oracle.py on one full-length chunk lands on the reference's committed score rows."""
    from dorado_b200.config import load_model_config
    from dorado_b200.weights import synthetic_weights
    from oracle import nn_oracle
    g = np.load(GOLD / f"full_{kind}.npz")
    cfg = load_model_config(model_dir(kind))
    M, T = int(g["M"]), int(g["T"])
    sig = np.random.default_rng(int(g["signal_seed"])).standard_normal((M, T)).astype(np.float16)
    mine = nn_oracle.forward(cfg, synthetic_weights(cfg, int(g["weights_seed"])), sig[:1].astype(np.float32))
    got = mine[0, :: int(g["row_step"])]
    if cfg.clamp:
        got = np.clip(got, -5.0, 5.0)
    np.testing.assert_allclose(got, g["rows"][0], rtol=0, atol=1e-4)

"""Pin the oracle to the reference itself (CPU only; needs oracle/_ref/libdorado_ref.so, which is the
UNMODIFIED reference CPU source compiled by oracle/Makefile -- it travels to the GPU box prebuilt).

  * scans: oracle/crf_oracle.c vs the reference's inner::forward_scores / backward_scores / softmax
    (dorado/basecall/decode/CPUDecoder.cpp:43-92,130).  libtorch's vectorised exp/log and reduction order are
    not reproducible bit for bit, and the guides reach magnitudes of ~2 per block (fp32 ulp 2.4e-4 at 4000), so the
    bound is 4 ulp of the largest guide.
  * beam search + sequence/qstring generation: the reference's own beam_search_decode
    (dorado/basecall/decode/beam_search.cpp:522-606) fed the ORACLE's guides must give bit-identical sequence and
    moves; the qstring may differ in isolated characters where libm's powf/log10f and the contract's differ in the
    last ulp (bound 1 %).
  * network forward: oracle/nn_oracle.py vs the reference's CRFModel / TxModel forward on the same weights.
"""
import numpy as np
import pytest

from conftest import model_dir, synthetic_scores


def _model(reference, tmp_path, kind, seed=42):
    from dorado_b200.config import load_model_config
    from dorado_b200.weights import save_b2w, synthetic_weights
    cfg = load_model_config(model_dir(kind))
    w = synthetic_weights(cfg, seed)
    save_b2w(tmp_path / "w.b2w", w)
    return cfg, w, reference.load_model(model_dir(kind), tmp_path / "w.b2w")


@pytest.mark.parametrize("state_len,T", [(3, 300), (4, 200), (5, 60)])
def test_scans_match_reference(crf_oracle, reference, state_len, T):
    s = synthetic_scores(1, T, state_len, seed=state_len, dtype=np.float32)[0]
    s = np.clip(s, -5, 5)
    f, b, p = crf_oracle.scans(s)
    rf, rb, rp = reference.scans(s)
    tol = 4 * np.spacing(np.float32(np.abs(rb).max()))
    assert np.abs(f - rf).max() <= tol and np.abs(b - rb).max() <= tol
    assert np.abs(p - rp).max() <= 2e-3 * rp.max() + 1e-6  # softmax of values carrying that absolute noise
    np.testing.assert_allclose(p.sum(axis=1), 1.0, rtol=1e-5)


@pytest.mark.parametrize("state_len,T", [(3, 400), (3, 1666), (4, 250), (5, 80)])
def test_beam_search_matches_reference_on_same_guides(crf_oracle, reference, state_len, T):
    """The reference is fed the fp16 scores widened to fp32, i.e. its CPU path (CPUDecoder hands beam_search_decode
    float tensors).  Its beam_search<c10::Half, float> instantiation is not a usable oracle: the beam-init threshold
    there memcpy()s the float back-guides into a vector<Half> (beam_search.cpp:168-171), so the initial beam is
    chosen from reinterpreted bytes."""
    mism = total = 0
    for seed in range(6):
        s16 = synthetic_scores(1, T, state_len, seed=50 + seed, scale=1.5)[0]
        s = np.clip(s16.astype(np.float32), -5, 5)
        _, b, p = crf_oracle.scans(s)
        rs, rq, rm = reference.beam_search_decode(s, b, p, q_shift=-1.1, q_scale=1.1)
        os_, oq, om = crf_oracle.beam_search(s, b, p, q_shift=-1.1, q_scale=1.1)
        assert rs == os_
        np.testing.assert_array_equal(rm, om)
        assert len(rq) == len(oq)
        mism += sum(a != c for a, c in zip(rq, oq))
        total += len(rq)
    # glibc's powf differs from the correctly rounded value on 0.06 % of its arguments; a character only moves when that
    # last ulp also crosses a quantiser edge, so at most an isolated character may differ
    assert mism <= 1, f"{mism}/{total} qstring characters differ from the reference"


@pytest.mark.parametrize("beam_width,beam_cut", [(32, 100.0), (8, 20.0), (32, 0.0)])
def test_beam_options_match_reference(crf_oracle, reference, beam_width, beam_cut):
    s = np.clip(synthetic_scores(1, 200, 3, seed=9, scale=1.0, dtype=np.float32)[0], -5, 5)
    _, b, p = crf_oracle.scans(s)
    rs, rq, rm = reference.beam_search_decode(s, b, p, beam_width=beam_width, beam_cut=beam_cut)
    os_, oq, om = crf_oracle.beam_search(s, b, p, beam_width=beam_width, beam_cut=beam_cut)
    assert rs == os_ and (rm == om).all()
    assert sum(a != c for a, c in zip(rq, oq)) <= 1


def test_full_decode_matches_reference_cpu_decoder(crf_oracle, reference):
    """CPUDecoder::beam_search_part_2 end to end (its own libtorch scans) vs the oracle: guides differ by a few
    ulp, so the bound is statistical: all move tables / sequences equal on these inputs, qstrings >= 98 %."""
    s16 = synthetic_scores(6, 300, 3, seed=3, scale=1.5)
    r = reference.decode(np.clip(s16.astype(np.float32), -5, 5))
    o = crf_oracle.decode(s16, clamp_val=5.0)
    same_seq = sum(a == b for a, b in zip(r.sequences, o.sequences))
    assert same_seq >= 5
    q_tot = q_bad = 0
    for a, b, sa, sb in zip(r.qstrings, o.qstrings, r.sequences, o.sequences):
        if sa == sb:
            q_tot += len(a)
            q_bad += sum(x != y for x, y in zip(a, b))
    assert q_bad <= 0.002 * q_tot + 1


@pytest.mark.parametrize("kind,N,T", [("fast", 2, 1200), ("hac", 2, 900), ("sup", 1, 1536)])
def test_forward_matches_reference(reference, tmp_path, kind, N, T):
    from oracle import nn_oracle
    cfg, w, h = _model(reference, tmp_path, kind)
    info = reference.model_info(h)
    assert info["stride"] == cfg.stride and info["outsize"] == cfg.outsize and info["state_len"] == cfg.state_len
    assert info["clamp"] == cfg.clamp and abs(info["qscale"] - cfg.qscale) < 1e-6 and abs(info["qbias"] - cfg.qbias) < 1e-6
    sig = np.random.default_rng(7).standard_normal((N, cfg.normalise_chunk_size(T))).astype(np.float32)
    ref = reference.forward(h, sig)
    mine = nn_oracle.forward(cfg, w, sig, cpu_split_quirk=True)  # the CPU fallback's 12-way key slicing
    assert ref.shape == mine.shape
    np.testing.assert_allclose(mine, ref, rtol=0, atol=5e-5)
    if kind == "sup":
        # the true window differs from the CPU fallback only through the one dropped key per split
        true_win = nn_oracle.forward(cfg, w, sig, cpu_split_quirk=False)
        d = np.abs(true_win - ref)
        assert d.max() <= 2e-2 and (d > 1e-3).mean() <= 5e-3
    reference.free_model(h)


@pytest.mark.parametrize("kind", ["fast", "hac"])
def test_reference_model_runner_reproduces_full_length_fixture(reference, kind):
    """dorado::basecall::ModelRunner (accept_chunk / call_chunks, ModelRunner.cpp:32-49) driven through the shim gives
    exactly the strings committed in tests/golden/full_<kind>.npz for the same chunk, and keeps its model_ms / decode_ms
    accounting (:51-57).  One chunk per call, as the fixture was generated: libtorch's fp32 kernels are not batch-invariant
    (a different GEMM blocking at batch 2 changes last bits of the scores, and these ill-conditioned synthetic models then
    flip a base), so only equal batch shapes are comparable bit for bit."""
    import pathlib
    from conftest import unpack_rows
    from dorado_b200.config import load_model_config
    from dorado_b200.weights import synthetic_weights
    from oracle.oracle import ReferenceRunner
    g = np.load(pathlib.Path(__file__).resolve().parent / "golden" / f"full_{kind}.npz")
    cfg = load_model_config(model_dir(kind))
    reference.set_num_threads(1)
    r = ReferenceRunner(reference, model_dir(kind), synthetic_weights(cfg, int(g["weights_seed"])), 1, 10000)
    assert r.chunk_size == int(g["T"]) and r.t_out == int(g["T_out"]) and r.batch_size == 1
    sig = np.random.default_rng(int(g["signal_seed"])).standard_normal((int(g["M"]), r.chunk_size)).astype(np.float16)
    nb = g["ref_n_bases"]
    seqs, qs = unpack_rows(g["ref_seq"], nb), unpack_rows(g["ref_qstr"], nb)
    mv = np.unpackbits(g["ref_moves"], axis=1)[:, : r.t_out]
    for i in (3, 0):
        r.accept_chunk(0, sig[i])
        out = r.call_chunks(1)
        assert out.sequences[0].encode() == seqs[i] and out.qstrings[0].encode() == qs[i]
        np.testing.assert_array_equal(out.moves[0], mv[i])
    st = r.sample_stats()
    assert st["batches_called"] == 2 and st["model_ms"] > 0 and st["decode_ms"] > 0
    r.close()

"""GPU parity of the CRF decoder: CUDA (through the C ABI) vs the CPU oracle, bit for bit.

Contract: on identical fp16 scores, sequence, qstring and move table are
bit-identical to the oracle's restatement of dorado/basecall/decode/{CPUDecoder,beam_search}.cpp.
"""
import numpy as np
import pytest

from conftest import synthetic_scores

pytestmark = pytest.mark.gpu


def _compare(crf_oracle, scores, clamp_val, **opt):
    from dorado_b200 import lib as L
    o = L.default_decoder_options()
    for k, v in opt.items():
        setattr(o, k, v)
    moves, seq, qstr, nb = L.decode_scores(scores, clamp_val=clamp_val, opts=o)
    ref = crf_oracle.decode(scores, clamp_val=clamp_val, beam_width=o.beam_width, beam_cut=o.beam_cut,
                            blank=o.blank_score, q_shift=o.q_shift, q_scale=o.q_scale)
    np.testing.assert_array_equal(nb, ref.n_bases)
    np.testing.assert_array_equal(moves, ref.moves)
    np.testing.assert_array_equal(seq, ref.seq_buf)
    np.testing.assert_array_equal(qstr, ref.qstr_buf)
    return nb


@pytest.mark.parametrize("state_len", [3, 4, 5])
@pytest.mark.parametrize("T", [1, 2, 37, 200])
def test_decode_bit_exact_small(crf_oracle, state_len, T):
    scores = synthetic_scores(5, T, state_len, seed=100 * state_len + T)
    _compare(crf_oracle, scores, clamp_val=5.0, q_shift=-1.1, q_scale=1.1)


@pytest.mark.parametrize("state_len,N,T", [(3, 33, 833), (4, 9, 833), (5, 3, 416)])
def test_decode_bit_exact_ragged_batch(crf_oracle, state_len, N, T):
    scores = synthetic_scores(N, T, state_len, seed=7 + state_len, scale=1.5)
    nb = _compare(crf_oracle, scores, clamp_val=0.0, q_shift=-3.0, q_scale=1.04)
    assert nb.min() >= 1 and nb.max() <= T


@pytest.mark.parametrize("beam_width,beam_cut", [(32, 100.0), (16, 100.0), (5, 10.0), (32, 0.0), (1, 100.0)])
def test_decode_beam_options(crf_oracle, beam_width, beam_cut):
    scores = synthetic_scores(4, 150, 3, seed=11, scale=1.0)
    _compare(crf_oracle, scores, clamp_val=5.0, beam_width=beam_width, beam_cut=beam_cut)


def test_decode_degenerate_scores(crf_oracle):
    """All-equal scores: every candidate ties, exercising the cutoff binary search / 80 % rule and the
    hash-merge path; zeros and saturated values."""
    for val in (0.0, 5.0, -5.0):
        scores = np.full((2, 64, 256), val, np.float16)
        _compare(crf_oracle, scores, clamp_val=5.0)
    scores = np.zeros((2, 64, 1024), np.float16)
    scores[:, ::2] = 4.0
    _compare(crf_oracle, scores, clamp_val=5.0)


def test_decode_full_size_properties(crf_oracle):
    """BASELINE-sized chunk count is too slow for the scalar oracle; check size-independent properties on a
    full-length batch (T=1666) and exact parity on a sample of its chunks."""
    from dorado_b200 import lib as L
    N, T = 64, 1666
    scores = synthetic_scores(N, T, 3, seed=5, scale=1.5)
    moves, seq, qstr, nb = L.decode_scores(scores, clamp_val=5.0)
    assert (moves[:, 0] == 1).all()
    np.testing.assert_array_equal(moves.sum(axis=1), nb)
    for i in range(N):
        assert set(seq[i, : nb[i]].tobytes()) <= set(b"ACGT")
        assert (seq[i, nb[i]:] == 0).all() and (qstr[i, nb[i]:] == 0).all()
        assert qstr[i, : nb[i]].min() >= 34 and qstr[i, : nb[i]].max() <= 83  # '!'+1 .. '!'+50
    # permutation invariance: chunks are independent
    perm = np.random.default_rng(0).permutation(N)
    m2, s2, q2, n2 = L.decode_scores(scores[perm], clamp_val=5.0)
    np.testing.assert_array_equal(m2, moves[perm])
    np.testing.assert_array_equal(s2, seq[perm])
    np.testing.assert_array_equal(q2, qstr[perm])
    sample = [0, 17, 63]
    ref = crf_oracle.decode(scores[sample], clamp_val=5.0)
    np.testing.assert_array_equal(moves[sample], ref.moves)
    np.testing.assert_array_equal(seq[sample], ref.seq_buf)
    np.testing.assert_array_equal(qstr[sample], ref.qstr_buf)


def test_decode_rejects_bad_arguments():
    from dorado_b200 import lib as L
    with pytest.raises(L.B200Error) as e:
        L.decode_scores(np.zeros((1, 4, 100), np.float16))
    assert e.value.status == L.B200_ERR_INVALID
    o = L.default_decoder_options()
    o.beam_width = 64
    with pytest.raises(L.B200Error):
        L.decode_scores(np.zeros((1, 4, 256), np.float16), opts=o)

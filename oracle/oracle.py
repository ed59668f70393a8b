"""TEST INFRASTRUCTURE ONLY: ctypes bindings for the two CPU checkers.

* ``CrfOracle``  -> oracle/libcrf_oracle.so   (plain-C restatement, crf_oracle.c)
* ``Reference``  -> oracle/_ref/libdorado_ref.so (the unmodified reference CPU sources + ref_driver.cpp)
"""
from __future__ import annotations

import ctypes as C
import os
import pathlib
import subprocess

import numpy as np

HERE = pathlib.Path(__file__).resolve().parent
_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
_u16p = np.ctypeslib.ndpointer(np.uint16, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")


def build(ref: bool = True) -> None:
    """Compile the checkers (building the checker is not using it)."""
    targets = ["oracle"] + (["ref"] if ref and os.path.isdir("/root/reference/dorado") else [])
    subprocess.run(["make", "-s", "-j8", "-C", str(HERE)] + targets, check=True)


def _rows_to_strings(buf: np.ndarray, n_bases: np.ndarray):
    return [bytes(buf[i, : n_bases[i]]).decode("ascii") for i in range(buf.shape[0])]


class DecodeResult:
    """Batch decode output in the reference's DecodedChunk terms (DecodedChunk.h:9-13)."""

    def __init__(self, seq, qstr, moves, n_bases):
        self.seq_buf, self.qstr_buf, self.moves, self.n_bases = seq, qstr, moves, n_bases

    @property
    def sequences(self):
        return _rows_to_strings(self.seq_buf, self.n_bases)

    @property
    def qstrings(self):
        return _rows_to_strings(self.qstr_buf, self.n_bases)


class CrfOracle:
    def __init__(self, path=None):
        path = pathlib.Path(path or HERE / "libcrf_oracle.so")
        if not path.exists():
            build(ref=False)
        self.lib = lib = C.CDLL(str(path))
        dec_tail = [C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, _u8p, _u8p, _u8p, _i32p]
        lib.crf_decode_f16.argtypes = [_u16p, C.c_int, C.c_int, C.c_int, C.c_float] + dec_tail
        lib.crf_decode_f32.argtypes = [_f32p, C.c_int, C.c_int, C.c_int, C.c_float] + dec_tail
        lib.crf_forward_scan.argtypes = [_f32p, C.c_int, C.c_int, C.c_float, _f32p]
        lib.crf_backward_scan.argtypes = [_f32p, C.c_int, C.c_int, C.c_float, _f32p]
        lib.crf_posts.argtypes = [_f32p, _f32p, C.c_int, C.c_int, _f32p]
        lib.crf_beam_search.argtypes = [_f32p, _f32p, _f32p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float,
                                        _i32p, _u8p, _f32p]
        lib.crf_beam_search.restype = C.c_float
        lib.crf_generate_sequence.argtypes = [_u8p, _i32p, _f32p, C.c_int, C.c_float, C.c_float, _u8p, _u8p]
        lib.crf_generate_sequence.restype = C.c_int
        for fn in ("crf_math_expf", "crf_math_expf_nonpos", "crf_math_logf", "crf_math_log1pf", "crf_math_pow0p4f"):
            getattr(lib, fn).argtypes = [C.c_float]
            getattr(lib, fn).restype = C.c_float
        lib.crf_math_lse2.argtypes = [C.c_float, C.c_float]
        lib.crf_math_lse2.restype = C.c_float
        lib.crf_half_to_float.argtypes = [C.c_uint16]
        lib.crf_half_to_float.restype = C.c_float

    def decode(self, scores: np.ndarray, clamp_val=0.0, beam_width=32, beam_cut=100.0, blank=2.0,
               q_shift=0.0, q_scale=1.0) -> DecodeResult:
        """scores [N,T,C] float16 or float32."""
        N, T, Cc = scores.shape
        seq = np.zeros((N, T), np.uint8)
        qstr = np.zeros((N, T), np.uint8)
        moves = np.zeros((N, T), np.uint8)
        nb = np.zeros(N, np.int32)
        if scores.dtype == np.float16:
            fn, arr = self.lib.crf_decode_f16, np.ascontiguousarray(scores).view(np.uint16)
        else:
            fn, arr = self.lib.crf_decode_f32, np.ascontiguousarray(scores, np.float32)
        fn(arr, N, T, Cc, clamp_val, beam_width, beam_cut, blank, q_shift, q_scale, seq, qstr, moves, nb)
        return DecodeResult(seq, qstr, moves, nb)

    def scans(self, scores_tc: np.ndarray, blank=2.0):
        """One chunk, fp32 scores [T,C] -> (fwd, bwd, posts) each [T+1, C/4]."""
        s = np.ascontiguousarray(scores_tc, np.float32)
        T, Cc = s.shape
        S = Cc // 4
        fwd = np.empty((T + 1, S), np.float32)
        bwd = np.empty((T + 1, S), np.float32)
        posts = np.empty((T + 1, S), np.float32)
        self.lib.crf_forward_scan(s, T, S, blank, fwd)
        self.lib.crf_backward_scan(s, T, S, blank, bwd)
        self.lib.crf_posts(fwd, bwd, T, S, posts)
        return fwd, bwd, posts

    def beam_search(self, scores_tc, bwd, posts, beam_width=32, beam_cut=100.0, blank=2.0,
                    q_shift=0.0, q_scale=1.0):
        s = np.ascontiguousarray(scores_tc, np.float32)
        T, Cc = s.shape
        S = Cc // 4
        states = np.zeros(T, np.int32)
        moves = np.zeros(T, np.uint8)
        qual = np.zeros(T * 4, np.float32)
        self.lib.crf_beam_search(s, np.ascontiguousarray(bwd), np.ascontiguousarray(posts), T,
                                 int(np.log2(S)), beam_width, beam_cut, blank, states, moves, qual)
        seq = np.zeros(T, np.uint8)
        qstr = np.zeros(T, np.uint8)
        n = self.lib.crf_generate_sequence(moves, states, qual, T, q_shift, q_scale, seq, qstr)
        return bytes(seq[:n]).decode(), bytes(qstr[:n]).decode(), moves


class Reference:
    """The reference's own CPU implementation (see ref_driver.cpp)."""

    def __init__(self, path=None):
        path = pathlib.Path(path or HERE / "_ref" / "libdorado_ref.so")
        if not path.exists():
            raise FileNotFoundError(f"{path} not built (run `make -C oracle ref` where /root/reference exists)")
        import torch  # noqa: F401  (loads libtorch's dependencies into the process)
        self.lib = lib = C.CDLL(str(path))
        lib.ref_last_error.restype = C.c_char_p
        lib.ref_model_create.restype = C.c_void_p
        lib.ref_model_create.argtypes = [C.c_char_p, C.c_char_p]
        lib.ref_model_destroy.argtypes = [C.c_void_p]
        lib.ref_model_info.argtypes = [C.c_void_p, _i32p, _f32p]
        lib.ref_forward.argtypes = [C.c_void_p, _f32p, C.c_int, C.c_int, C.c_void_p,
                                    C.POINTER(C.c_int), C.POINTER(C.c_int)]
        lib.ref_scans.argtypes = [_f32p, C.c_int, C.c_int, C.c_float, _f32p, _f32p, _f32p]
        lib.ref_beam_search_decode.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, _f32p, _f32p, C.c_int,
                                               C.c_float, C.c_float, C.c_float, C.c_float, _u8p, _u8p, _u8p,
                                               C.POINTER(C.c_int)]
        lib.ref_decode_chunks.argtypes = [_f32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float,
                                          C.c_float, C.c_float, _u8p, _u8p, _u8p, _i32p]
        lib.ref_set_num_threads.argtypes = [C.c_int]
        lib.ref_runner_create.restype = C.c_void_p
        lib.ref_runner_create.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_int, C.c_int]
        lib.ref_runner_destroy.argtypes = [C.c_void_p]
        lib.ref_runner_dims.argtypes = [C.c_void_p, _i32p]
        lib.ref_runner_accept_chunk.argtypes = [C.c_void_p, C.c_int, _f32p, C.c_int]
        lib.ref_runner_call_chunks.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.ref_runner_stats.argtypes = [C.c_void_p, np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")]
        u64 = C.c_uint64
        _u64p = np.ctypeslib.ndpointer(np.uint64, flags="C_CONTIGUOUS")
        lib.ref_generate_chunks.argtypes = [u64, u64, u64, u64, _u64p, u64]
        lib.ref_generate_chunks.restype = C.c_long
        lib.ref_generate_variable_chunks.argtypes = [u64, u64, u64, u64, _u64p, u64]
        lib.ref_generate_variable_chunks.restype = C.c_long
        lib.ref_make_chunk_input.argtypes = [np.ctypeslib.ndpointer(np.int16, flags="C_CONTIGUOUS"), u64, u64, u64,
                                             C.c_float, C.c_float, _u16p]
        lib.ref_stitch_chunks.argtypes = [u64, _u64p, _u64p, _u8p, _u64p, C.c_char_p, C.c_char_p, _u64p, u64, C.c_int,
                                          _u8p, _u8p, _u8p, C.POINTER(u64), C.POINTER(u64)]

    def _check(self, rc):
        if rc != 0:
            raise RuntimeError(self.lib.ref_last_error().decode())

    def set_num_threads(self, n):
        self.lib.ref_set_num_threads(n)

    # ---- front end (chunk.cpp, stitch.cpp, tensor_utils.cpp through ref_driver.cpp) --------------------------
    def generate_chunks(self, num_samples, chunk_size, stride, overlap):
        """utils::generate_chunks; raises RuntimeError where the reference throws."""
        cap = 1 << 16
        out = np.zeros(cap, np.uint64)
        n = self.lib.ref_generate_chunks(num_samples, chunk_size, stride, overlap, out, cap)
        if n < 0:
            raise RuntimeError(self.lib.ref_last_error().decode())
        return [int(v) for v in out[:n]]

    def generate_variable_chunks(self, num_samples, chunk_size, stride, overlap):
        """utils::generate_variable_chunks -> [(first, second), ...]; raises RuntimeError where the reference throws."""
        cap = 1 << 16
        out = np.zeros(2 * cap, np.uint64)
        n = self.lib.ref_generate_variable_chunks(num_samples, chunk_size, stride, overlap, out, cap)
        if n < 0:
            raise RuntimeError(self.lib.ref_last_error().decode())
        return [(int(out[2 * i]), int(out[2 * i + 1])) for i in range(n)]

    def make_chunk_input(self, raw, input_offset, chunk_size, shift, scale):
        """ScalerNode scaling + BasecallerNode slice/repeat-pad of one chunk -> fp16 [chunk_size]."""
        r = np.ascontiguousarray(raw, np.int16)
        out = np.zeros(chunk_size, np.uint16)
        self._check(self.lib.ref_make_chunk_input(r, r.size, input_offset, chunk_size, shift, scale, out))
        return out.view(np.float16)

    def stitch_chunks(self, chunks, raw_samples, stride):
        """chunks: (input_offset, raw_chunk_size, moves, sequence, qstring) -> (sequence, qstring, moves)."""
        offs = np.array([c[0] for c in chunks], np.uint64)
        sizes = np.array([c[1] for c in chunks], np.uint64)
        moves = np.ascontiguousarray(np.concatenate([np.asarray(c[2], np.uint8) for c in chunks]))
        mlen = np.array([len(c[2]) for c in chunks], np.uint64)
        seq = "".join(c[3] for c in chunks).encode("ascii")
        qs = "".join(c[4] for c in chunks).encode("ascii")
        slen = np.array([len(c[3]) for c in chunks], np.uint64)
        mo = np.zeros(max(1, moves.size), np.uint8)
        so = np.zeros(max(1, len(seq)), np.uint8)
        qo = np.zeros(max(1, len(seq)), np.uint8)
        nm, nb = C.c_uint64(), C.c_uint64()
        self._check(self.lib.ref_stitch_chunks(len(chunks), offs, sizes, moves, mlen, seq, qs, slen, raw_samples, stride,
                                               mo, so, qo, C.byref(nm), C.byref(nb)))
        return bytes(so[: nb.value]).decode("ascii"), bytes(qo[: nb.value]).decode("ascii"), mo[: nm.value].copy()

    def load_model(self, config_dir, weights_file):
        h = self.lib.ref_model_create(str(config_dir).encode(), str(weights_file).encode())
        if not h:
            raise RuntimeError(self.lib.ref_last_error().decode())
        return h

    def free_model(self, h):
        self.lib.ref_model_destroy(h)

    def model_info(self, h):
        info = np.zeros(6, np.int32)
        q = np.zeros(2, np.float32)
        self.lib.ref_model_info(h, info, q)
        return dict(stride=int(info[0]), outsize=int(info[1]), state_len=int(info[2]), is_tx=bool(info[3]),
                    clamp=bool(info[4]), num_features=int(info[5]), qscale=float(q[0]), qbias=float(q[1]))

    def forward(self, h, signal: np.ndarray) -> np.ndarray:
        """signal [N,T] or [N,1,T] fp32 -> scores [N,T_out,C] fp32."""
        sig = np.ascontiguousarray(signal, np.float32).reshape(signal.shape[0], -1)
        N, T = sig.shape
        to, co = C.c_int(), C.c_int()
        self._check(self.lib.ref_forward(h, sig, N, T, None, C.byref(to), C.byref(co)))
        out = np.empty((N, to.value, co.value), np.float32)
        self._check(self.lib.ref_forward(h, sig, N, T, out.ctypes.data_as(C.c_void_p), C.byref(to), C.byref(co)))
        return out

    def scans(self, scores_tc, blank=2.0):
        s = np.ascontiguousarray(scores_tc, np.float32)
        T, Cc = s.shape
        outs = [np.empty((T + 1, Cc // 4), np.float32) for _ in range(3)]
        self._check(self.lib.ref_scans(s, T, Cc, blank, *outs))
        return tuple(outs)

    def beam_search_decode(self, scores_tc, bwd, posts, beam_width=32, beam_cut=100.0, blank=2.0,
                           q_shift=0.0, q_scale=1.0):
        T, Cc = scores_tc.shape
        is_half = scores_tc.dtype == np.float16
        s = np.ascontiguousarray(scores_tc if is_half else scores_tc.astype(np.float32))
        seq, qstr, moves = (np.zeros(T, np.uint8) for _ in range(3))
        n = C.c_int()
        self._check(self.lib.ref_beam_search_decode(
            s.ctypes.data_as(C.c_void_p), int(is_half), T, Cc, np.ascontiguousarray(bwd, np.float32),
            np.ascontiguousarray(posts, np.float32), beam_width, beam_cut, blank, q_shift, q_scale, seq, qstr,
            moves, C.byref(n)))
        return bytes(seq[: n.value]).decode(), bytes(qstr[: n.value]).decode(), moves

    def decode(self, scores: np.ndarray, beam_width=32, beam_cut=100.0, blank=2.0, q_shift=0.0,
               q_scale=1.0) -> DecodeResult:
        s = np.ascontiguousarray(scores, np.float32)
        N, T, Cc = s.shape
        seq = np.zeros((N, T), np.uint8)
        qstr = np.zeros((N, T), np.uint8)
        moves = np.zeros((N, T), np.uint8)
        nb = np.zeros(N, np.int32)
        self._check(self.lib.ref_decode_chunks(s, N, T, Cc, beam_width, beam_cut, blank, q_shift, q_scale, seq,
                                               qstr, moves, nb))
        return DecodeResult(seq, qstr, moves, nb)


class ReferenceRunner:
    """The reference's CPU runner itself: dorado::basecall::ModelRunner (dorado/basecall/ModelRunner.cpp) with the
    ModelRunnerBase methods accept_chunk / call_chunks / sample_stats, on the synthetic weights written out as the
    reference's own *.tensor files."""

    def __init__(self, ref: "Reference", config_dir, weights, batch_size: int, chunk_size: int, share_with=None):
        """weights: name -> array; share_with: another ReferenceRunner of the same model whose *.tensor files are reused."""
        import tempfile
        from dorado_b200.weights import save_b2w
        self.ref = ref
        if share_with is not None:
            self._tmp = share_with._tmp
            self.h = ref.lib.ref_runner_create(str(config_dir).encode(), b"", self._tmp.name.encode(), batch_size, chunk_size)
        else:
            self._tmp = tempfile.TemporaryDirectory()
            wpath = os.path.join(self._tmp.name, "w.b2w")
            save_b2w(wpath, weights)
            self.h = ref.lib.ref_runner_create(str(config_dir).encode(), wpath.encode(), self._tmp.name.encode(),
                                               batch_size, chunk_size)
            os.remove(wpath)
        if not self.h:
            raise RuntimeError(ref.lib.ref_last_error().decode())
        self._owns_tmp = share_with is None
        d = np.zeros(3, np.int32)
        ref.lib.ref_runner_dims(self.h, d)
        self.batch_size, self.chunk_size, self.stride = int(d[0]), int(d[1]), int(d[2])
        self.t_out = self.chunk_size // self.stride

    def accept_chunk(self, idx: int, chunk: np.ndarray) -> None:
        c = np.ascontiguousarray(chunk, np.float32).reshape(-1)
        self.ref._check(self.ref.lib.ref_runner_accept_chunk(self.h, idx, c, c.size))

    def call_chunks(self, num_chunks: int, want_output: bool = True):
        if not want_output:
            self.ref._check(self.ref.lib.ref_runner_call_chunks(self.h, num_chunks, None, None, None, None))
            return None
        seq, qstr, moves = (np.zeros((num_chunks, self.t_out), np.uint8) for _ in range(3))
        nb = np.zeros(num_chunks, np.int32)
        self.ref._check(self.ref.lib.ref_runner_call_chunks(self.h, num_chunks, seq.ctypes.data, qstr.ctypes.data,
                                                            moves.ctypes.data, nb.ctypes.data))
        return DecodeResult(seq, qstr, moves, nb)

    def sample_stats(self) -> dict:
        st = np.zeros(3, np.float64)
        self.ref._check(self.ref.lib.ref_runner_stats(self.h, st))
        return {"batches_called": st[0], "model_ms": st[1], "decode_ms": st[2]}

    def close(self):
        if self.h:
            self.ref.lib.ref_runner_destroy(self.h)
            self.h = None
            if self._owns_tmp:
                self._tmp.cleanup()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def reference_available() -> bool:
    return (HERE / "_ref" / "libdorado_ref.so").exists()

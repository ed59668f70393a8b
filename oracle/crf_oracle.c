/* crf_oracle.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Plain-C restatement of the reference CRF decoder, the checker for dorado_b200/csrc/decode.cu.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this library.
 *
 * What it restates (reference = nanoporetech/dorado, paths relative to dorado/basecall/decode/):
 *   crf_forward_scan / crf_backward_scan   CPUDecoder.cpp:17-38 (scan), :43-67, :69-92
 *   crf_posts                              CPUDecoder.cpp:130   (softmax(fwd + bwd))
 *   crf_beam_search                        beam_search.cpp:125-520
 *   crf_generate_sequence                  beam_search.cpp:54-102
 *   crf_decode_f16 / crf_decode_f32        CUDADecoder.cpp:17-173 semantics (fp16 scores, clamp on
 *                                          read, fp32 scans) == CPUDecoder::beam_search_part_2 on the
 *                                          same scores widened to fp32.
 *
 * Transcendentals come from include/b200_crf_math.h (the numerics contract shared with the CUDA
 * kernels) instead of libm/libtorch, and reductions use the fixed orders documented there and in
 * crf_posts below; everything else follows the reference operation by operation.
 *
 * Pinning (tests/test_oracle_vs_reference.py): scans are checked against the reference's own
 * inner::forward_scores/backward_scores, and crf_beam_search against the reference's own
 * beam_search_decode fed the same guides, through oracle/_ref/libdorado_ref.so.
 */
#include "b200_crf_math.h"

#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define NUM_BASE_BITS 2
#define NUM_BASES 4
#define MAX_BEAM 256

static float half_to_float(uint16_t h) {
    const uint32_t sign = (uint32_t)(h & 0x8000u) << 16;
    uint32_t exp = (h >> 10) & 0x1fu;
    uint32_t man = h & 0x3ffu;
    uint32_t bits;
    if (exp == 0) {
        if (man == 0) {
            bits = sign;
        } else {
            /* subnormal half -> normal float */
            int e = -1;
            do {
                ++e;
                man <<= 1;
            } while ((man & 0x400u) == 0);
            man &= 0x3ffu;
            bits = sign | (uint32_t)(127 - 15 - e) << 23 | (man << 13);
        }
    } else if (exp == 31) {
        bits = sign | 0x7f800000u | (man << 13);
    } else {
        bits = sign | ((exp + 127 - 15) << 23) | (man << 13);
    }
    return B200_U2F(bits);
}

static float clamp_score(float v, float clamp_val) {
    if (clamp_val > 0.0f) {
        if (v < -clamp_val) v = -clamp_val;
        if (v > clamp_val) v = clamp_val;
    }
    return v;
}

/* ------------------------------------------------------------------------------------------- */
/* scans: scores [T, 4S] fp32, guides [T+1, S] fp32                                            */
/* ------------------------------------------------------------------------------------------- */

void crf_forward_scan(const float* scores, int T, int S, float blank, float* fwd) {
    const int C = S * 4;
    const int q4 = S / 4;
    for (int s = 0; s < S; ++s) fwd[s] = 0.0f;
    for (int t = 0; t < T; ++t) {
        const float* a = fwd + (size_t)t * S;
        float* o = fwd + (size_t)(t + 1) * S;
        const float* m = scores + (size_t)t * C;
        for (int s = 0; s < S; ++s) {
            const int p = s >> 2; /* predecessors: p + b * S/4, transition s*4 + b */
            o[s] = b200_lse5(B200_ADD(a[s], blank), B200_ADD(a[p], m[s * 4 + 0]),
                             B200_ADD(a[p + q4], m[s * 4 + 1]), B200_ADD(a[p + 2 * q4], m[s * 4 + 2]),
                             B200_ADD(a[p + 3 * q4], m[s * 4 + 3]));
        }
    }
}

void crf_backward_scan(const float* scores, int T, int S, float blank, float* bwd) {
    const int C = S * 4;
    const int q4 = S / 4;
    for (int s = 0; s < S; ++s) bwd[(size_t)T * S + s] = 0.0f;
    for (int t = T - 1; t >= 0; --t) {
        const float* a = bwd + (size_t)(t + 1) * S;
        float* o = bwd + (size_t)t * S;
        const float* m = scores + (size_t)t * C;
        for (int s = 0; s < S; ++s) {
            const int top = s / q4;          /* base dropped when stepping out of s */
            const int succ = (s % q4) * 4;   /* successors succ + b, transition (succ+b)*4 + top */
            o[s] = b200_lse5(B200_ADD(a[s], blank), B200_ADD(a[succ + 0], m[(succ + 0) * 4 + top]),
                             B200_ADD(a[succ + 1], m[(succ + 1) * 4 + top]),
                             B200_ADD(a[succ + 2], m[(succ + 2) * 4 + top]),
                             B200_ADD(a[succ + 3], m[(succ + 3) * 4 + top]));
        }
    }
}

/* softmax over states of fwd + bwd.  Reduction order (part of the numerics contract, mirrors the CUDA kernel's
 * one-thread-per-state layout): SPT = max(1, S/512) consecutive states are summed left to right, a xor-butterfly
 * runs over each aligned group of min(S/SPT, 32) such partials, and the groups are summed left to right. */
static void posts_row(const float* f, const float* b, int S, float* out) {
    float v[1024], e[1024], part[512];
    float mx = B200_ADD(f[0], b[0]);
    for (int s = 0; s < S; ++s) {
        v[s] = B200_ADD(f[s], b[s]);
        mx = b200_fmaxf(mx, v[s]);
    }
    const int SPT = S > 512 ? S / 512 : 1;
    const int P = S / SPT;
    for (int q = 0; q < P; ++q) {
        for (int j = 0; j < SPT; ++j) e[SPT * q + j] = b200_expf_nonpos(B200_SUB(v[SPT * q + j], mx));
        part[q] = e[SPT * q];
        for (int j = 1; j < SPT; ++j) part[q] = B200_ADD(part[q], e[SPT * q + j]);
    }
    const int G = P < 32 ? P : 32;
    for (int o = G / 2; o >= 1; o >>= 1) {
        float nxt[512];
        for (int q = 0; q < P; ++q) nxt[q] = B200_ADD(part[q], part[q ^ o]);
        memcpy(part, nxt, sizeof(float) * (size_t)P);
    }
    float z = part[0];
    for (int w = 1; w < P / G; ++w) z = B200_ADD(z, part[w * G]);
    for (int s = 0; s < S; ++s) out[s] = B200_DIV(e[s], z);
}

void crf_posts(const float* fwd, const float* bwd, int T, int S, float* posts) {
    for (int t = 0; t <= T; ++t) {
        posts_row(fwd + (size_t)t * S, bwd + (size_t)t * S, S, posts + (size_t)t * S);
    }
}

/* ------------------------------------------------------------------------------------------- */
/* beam search (beam_search.cpp:125-520)                                                        */
/* ------------------------------------------------------------------------------------------- */

static uint32_t crc32c_bits(uint32_t crc, uint32_t new_bits, int nbits) {
    /* Castagnoli polynomial, reflected (beam_search.cpp:106-119) */
    for (int i = 0; i < nbits; ++i) {
        const uint32_t b = (new_bits ^ crc) & 1u;
        crc >>= 1;
        if (b) crc ^= 0x82f63b78u;
        new_bits >>= 1;
    }
    return crc;
}

typedef struct {
    uint32_t hash;
    uint16_t state;
    uint8_t prev;
    uint8_t stay;
} front_t;

typedef struct {
    uint16_t state;
    uint8_t prev;
    uint8_t stay;
} beam_elem_t;

static int cmp_desc(const void* a, const void* b) {
    const float x = *(const float*)a, y = *(const float*)b;
    return (x < y) - (x > y);
}

static size_t count_ge(const float* v, size_t n, float cutoff) {
    size_t c = 0;
    for (size_t i = 0; i < n; ++i) c += (v[i] >= cutoff);
    return c;
}

/* scores [T, 4S] fp32 (already widened/clamped), bwd & posts [T+1, S].
 * Out: states[T] (full kmer state per block), moves[T], qual[T*4]. Returns final score. */
float crf_beam_search(const float* scores,
                      const float* bwd,
                      const float* posts,
                      int T,
                      int state_bits,
                      int W,
                      float beam_cut,
                      float blank,
                      int32_t* states,
                      uint8_t* moves,
                      float* qual) {
    const int S = 1 << state_bits;
    const uint32_t mask = (uint32_t)S - 1u;
    /* a per-call constant: libm (correctly rounded in glibc), as the host side of the engine does */
    const float log_beam_cut = (beam_cut > 0.0f) ? logf(beam_cut) : 3.402823466e+38f;
    if (W > MAX_BEAM) return 0.0f;

    beam_elem_t* beam = (beam_elem_t*)calloc((size_t)W * (size_t)(T + 1), sizeof(beam_elem_t));
    front_t cur[5 * MAX_BEAM], prev[5 * MAX_BEAM];
    float cur_sc[5 * MAX_BEAM], prev_sc[5 * MAX_BEAM];

    /* initial beam: the W best states by bwd[0], kept in state order (:166-189) */
    float thresh = B200_FLT_LOWEST;
    if (W < S) {
        float* tmp = (float*)malloc(sizeof(float) * (size_t)S);
        memcpy(tmp, bwd, sizeof(float) * (size_t)S);
        qsort(tmp, (size_t)S, sizeof(float), cmp_desc);
        thresh = tmp[W - 1];
        free(tmp);
    }
    int width = 0;
    for (int s = 0; s < S && width < W; ++s) {
        if (bwd[s] >= thresh) {
            prev[width].hash = crc32c_bits(0x12345678u, (uint32_t)s, 32);
            prev[width].state = (uint16_t)s;
            prev[width].prev = 0;
            prev[width].stay = 0;
            prev_sc[width] = 0.0f;
            ++width;
        }
    }
    width = W < S ? W : S;
    for (int i = 0; i < width; ++i) {
        beam[i].state = prev[i].state;
        beam[i].prev = prev[i].prev;
        beam[i].stay = prev[i].stay;
    }

    for (int t = 0; t < T; ++t) {
        const float* sc = scores + (size_t)t * (size_t)S * 4;
        const float* bg = bwd + (size_t)(t + 1) * (size_t)S;
        float max_score = B200_FLT_LOWEST;
        int n = 0;
        /* step candidates, index prev*4 + base (:225-252) */
        for (int i = 0; i < width; ++i) {
            const uint32_t shifted = (uint32_t)prev[i].state << NUM_BASE_BITS;
            for (uint32_t b = 0; b < NUM_BASES; ++b) {
                const uint16_t ns = (uint16_t)((shifted & mask) | b);
                const uint16_t move_idx = (uint16_t)(((uint32_t)ns << NUM_BASE_BITS) + (shifted >> state_bits));
                const float v = B200_ADD(B200_ADD(prev_sc[i], sc[move_idx]), bg[ns]);
                cur[n].hash = crc32c_bits(prev[i].hash, b, NUM_BASE_BITS);
                cur[n].state = ns;
                cur[n].prev = (uint8_t)i;
                cur[n].stay = 0;
                cur_sc[n] = v;
                max_score = b200_fmaxf(max_score, v);
                ++n;
            }
        }
        /* stay candidates, index width*4 + prev, merged with equal-hash steps (:254-308).
         * (The reference's 4096-bit "step_hash_present" filter only skips comparisons that cannot
         * match, so it is not restated.) */
        for (int i = 0; i < width; ++i) {
            const float v = B200_ADD(B200_ADD(prev_sc[i], blank), bg[prev[i].state]);
            cur[n].hash = prev[i].hash;
            cur[n].state = prev[i].state;
            cur[n].prev = (uint8_t)i;
            cur[n].stay = 1;
            cur_sc[n] = v;
            max_score = b200_fmaxf(max_score, v);
            const int stay_idx = (width << NUM_BASE_BITS) + i;
            const int latest = prev[i].state & 3;
            for (int j = 0; j < width; ++j) {
                const int step_idx = (j << NUM_BASE_BITS) | latest;
                if (cur[stay_idx].hash == cur[step_idx].hash) {
                    const float folded = b200_log_sum_exp(cur_sc[stay_idx], cur_sc[step_idx]);
                    if (cur_sc[stay_idx] > cur_sc[step_idx]) {
                        cur_sc[stay_idx] = folded;
                        cur_sc[step_idx] = B200_FLT_LOWEST;
                    } else {
                        cur_sc[step_idx] = folded;
                        cur_sc[stay_idx] = B200_FLT_LOWEST;
                    }
                    max_score = b200_fmaxf(max_score, folded);
                }
            }
            ++n;
        }

        /* cutoff selection (:310-396) */
        float cutoff = B200_SUB(max_score, log_beam_cut);
        size_t cnt = count_ge(cur_sc, (size_t)n, cutoff);
        if (cnt > (size_t)W) {
            const size_t min_w = ((size_t)W * 8) / 10;
            float lo = cutoff, hi = max_score;
            int guesses = 1;
            while ((cnt > (size_t)W || cnt < min_w) && guesses < 10) {
                if (cnt > (size_t)W) {
                    lo = cutoff;
                    cutoff = B200_DIV(B200_ADD(cutoff, hi), 2.0f);
                } else {
                    hi = cutoff;
                    cutoff = B200_DIV(B200_ADD(cutoff, lo), 2.0f);
                }
                cnt = count_ge(cur_sc, (size_t)n, cutoff);
                ++guesses;
            }
            if (guesses == 10) {
                cutoff = hi;
                cnt = count_ge(cur_sc, (size_t)n, cutoff);
            }
            if (cnt > (size_t)W) cnt = (size_t)W;
        }

        /* keep the first W candidates meeting the cutoff, in candidate order (:398-409) */
        int w = 0;
        for (int r = 0; r < n && w < W; ++r) {
            if (cur_sc[r] >= cutoff) {
                prev[w] = cur[r];
                prev_sc[w] = cur_sc[r];
                ++w;
            }
        }

        /* last block: best element to slot 0 (:413-424) */
        if (t == T - 1) {
            float best = B200_FLT_LOWEST;
            size_t bi = 0;
            for (size_t i = 0; i < cnt; ++i) {
                if (prev_sc[i] > best) {
                    best = prev_sc[i];
                    bi = i;
                }
            }
            const front_t tf = prev[0];
            prev[0] = prev[bi];
            prev[bi] = tf;
            const float ts = prev_sc[0];
            prev_sc[0] = prev_sc[bi];
            prev_sc[bi] = ts;
        }

        beam_elem_t* row = beam + (size_t)(t + 1) * (size_t)W;
        for (size_t i = 0; i < cnt; ++i) {
            prev_sc[i] = B200_SUB(prev_sc[i], bg[prev[i].state]);
            row[i].state = prev[i].state;
            row[i].prev = prev[i].prev;
            row[i].stay = prev[i].stay;
        }
        width = (int)cnt;
    }
    const float final_score = prev_sc[0];

    /* traceback (:447-455) */
    uint8_t ei = 0;
    for (int t = T; t != 0; --t) {
        const beam_elem_t* e = beam + (size_t)t * (size_t)W + ei;
        states[t - 1] = (int32_t)e->state;
        moves[t - 1] = e->stay ? 0 : 1;
        ei = e->prev;
    }
    if (T > 0) moves[0] = 1;
    free(beam);

    /* per-block quality (:459-517) */
    for (int t = 0; t < T; ++t) {
        const int state = states[t];
        const float* p = posts + (size_t)(t + 1) * (size_t)S;
        float prob = p[state];
        int shifted[2 * NUM_BASES];
        const int l = state >> NUM_BASE_BITS;
        const int r = (state << NUM_BASE_BITS) % S;
        const int msb = S >> NUM_BASE_BITS;
        for (int b = 0; b < NUM_BASES; ++b) {
            shifted[2 * b] = l + msb * b;
            shifted[2 * b + 1] = r + b;
        }
        for (int k = 0; k < 2 * NUM_BASES; ++k) {
            const int c = shifted[k];
            int count = (c != state);
            for (int j = 0; count && j < k; ++j) {
                if (shifted[j] == c) count = 0;
            }
            if (count) prob = B200_ADD(prob, p[c]);
        }
        prob = prob < 0.0f ? 0.0f : (prob > 1.0f ? 1.0f : prob);
        prob = b200_pow0p4f(prob);
        const float wrong = B200_DIV(B200_SUB(1.0f, prob), 3.0f);
        const int base = state % NUM_BASES;
        for (int b = 0; b < NUM_BASES; ++b) qual[t * NUM_BASES + b] = (b == base) ? prob : wrong;
    }
    return final_score;
}

/* beam_search.cpp:54-102. seq/qstr must hold T chars; returns number of bases. */
int crf_generate_sequence(const uint8_t* moves,
                          const int32_t* states,
                          const float* qual,
                          int T,
                          float shift,
                          float scale,
                          char* seq,
                          char* qstr) {
    static const char alphabet[4] = {'A', 'C', 'G', 'T'};
    int n_bases = 0;
    for (int t = 0; t < T; ++t) n_bases += moves[t];
    float* base_p = (float*)calloc((size_t)(n_bases > 0 ? n_bases : 1), sizeof(float));
    float* total_p = (float*)calloc((size_t)(n_bases > 0 ? n_bases : 1), sizeof(float));
    int pos = 0;
    for (int t = 0; t < T; ++t) {
        const int base = states[t] & 3;
        const int move = moves[t];
        const int pp = pos + ((t == 0) ? 0 : move - 1);
        base_p[pp] = B200_ADD(base_p[pp], qual[t * 4 + base]);
        for (int k = 0; k < 4; ++k) total_p[pp] = B200_ADD(total_p[pp], qual[t * 4 + k]);
        if (t == 0) {
            seq[pos++] = alphabet[base];
        } else {
            for (int j = 0; j < move; ++j) seq[pos++] = alphabet[base];
        }
    }
    /* the reference expression itself, on this host's libm (the engine reproduces it with b200_qtable) */
    for (int i = 0; i < n_bases; ++i) {
        const float err = B200_SUB(1.0f, B200_DIV(base_p[i], total_p[i]));
        qstr[i] = b200_qchar_libm(err, scale, shift);
    }
    free(base_p);
    free(total_p);
    return n_bases;
}

/* ------------------------------------------------------------------------------------------- */
/* whole-chunk decode                                                                           */
/* ------------------------------------------------------------------------------------------- */

static int ilog2(int v) {
    int b = 0;
    while ((1 << b) < v) ++b;
    return b;
}

/* One chunk, fp32 scores [T, C]. Optional outputs fwd/bwd/posts [T+1, S] (may be NULL). */
int crf_decode_chunk_f32(const float* scores,
                         int T,
                         int C,
                         int beam_width,
                         float beam_cut,
                         float blank,
                         float q_shift,
                         float q_scale,
                         char* seq,
                         char* qstr,
                         uint8_t* moves,
                         float* fwd_out,
                         float* bwd_out,
                         float* posts_out) {
    const int S = C / 4;
    const size_t g = (size_t)(T + 1) * (size_t)S;
    float* fwd = fwd_out ? fwd_out : (float*)malloc(g * sizeof(float));
    float* bwd = bwd_out ? bwd_out : (float*)malloc(g * sizeof(float));
    float* posts = posts_out ? posts_out : (float*)malloc(g * sizeof(float));
    int32_t* states = (int32_t*)malloc(sizeof(int32_t) * (size_t)(T > 0 ? T : 1));
    float* qual = (float*)malloc(sizeof(float) * 4 * (size_t)(T > 0 ? T : 1));
    crf_forward_scan(scores, T, S, blank, fwd);
    crf_backward_scan(scores, T, S, blank, bwd);
    crf_posts(fwd, bwd, T, S, posts);
    crf_beam_search(scores, bwd, posts, T, ilog2(S), beam_width, beam_cut, blank, states, moves, qual);
    const int n = crf_generate_sequence(moves, states, qual, T, q_shift, q_scale, seq, qstr);
    if (!fwd_out) free(fwd);
    if (!bwd_out) free(bwd);
    if (!posts_out) free(posts);
    free(states);
    free(qual);
    return n;
}

/* Batch decode of fp16 scores [N, T, C] (IEEE half bits), clamp applied on read when clamp_val > 0
 * (CUDADecoder.cpp:17-23 / Decoder.cpp:19). seq/qstr/moves are [N, T]; n_bases [N]. */
int crf_decode_f16(const uint16_t* scores,
                   int N,
                   int T,
                   int C,
                   float clamp_val,
                   int beam_width,
                   float beam_cut,
                   float blank,
                   float q_shift,
                   float q_scale,
                   char* seq,
                   char* qstr,
                   uint8_t* moves,
                   int32_t* n_bases) {
    float* buf = (float*)malloc(sizeof(float) * (size_t)T * (size_t)C + 4);
    for (int i = 0; i < N; ++i) {
        const uint16_t* s = scores + (size_t)i * (size_t)T * (size_t)C;
        for (size_t k = 0; k < (size_t)T * (size_t)C; ++k) buf[k] = clamp_score(half_to_float(s[k]), clamp_val);
        memset(seq + (size_t)i * T, 0, (size_t)T);
        memset(qstr + (size_t)i * T, 0, (size_t)T);
        n_bases[i] = crf_decode_chunk_f32(buf, T, C, beam_width, beam_cut, blank, q_shift, q_scale,
                                          seq + (size_t)i * T, qstr + (size_t)i * T, moves + (size_t)i * T,
                                          NULL, NULL, NULL);
    }
    free(buf);
    return 0;
}

int crf_decode_f32(const float* scores,
                   int N,
                   int T,
                   int C,
                   float clamp_val,
                   int beam_width,
                   float beam_cut,
                   float blank,
                   float q_shift,
                   float q_scale,
                   char* seq,
                   char* qstr,
                   uint8_t* moves,
                   int32_t* n_bases) {
    float* buf = (float*)malloc(sizeof(float) * (size_t)T * (size_t)C + 4);
    for (int i = 0; i < N; ++i) {
        const float* s = scores + (size_t)i * (size_t)T * (size_t)C;
        for (size_t k = 0; k < (size_t)T * (size_t)C; ++k) buf[k] = clamp_score(s[k], clamp_val);
        memset(seq + (size_t)i * T, 0, (size_t)T);
        memset(qstr + (size_t)i * T, 0, (size_t)T);
        n_bases[i] = crf_decode_chunk_f32(buf, T, C, beam_width, beam_cut, blank, q_shift, q_scale,
                                          seq + (size_t)i * T, qstr + (size_t)i * T, moves + (size_t)i * T,
                                          NULL, NULL, NULL);
    }
    free(buf);
    return 0;
}

/* scalar entry points so the tests can probe the numerics contract directly */
float crf_math_expf(float x) { return b200_expf(x); }
float crf_math_expf_nonpos(float x) { return b200_expf_nonpos(x); }
/* number of the n arguments (bit patterns, walked with the given stride from start) on which b200_expf_nonpos and b200_expf
 * differ, and on which the 2^n scale built by the bit trick differs from the (int) conversion form it replaced */
long crf_expf_audit(uint32_t start, uint32_t stride, long n, long* scale_mismatch) {
    long bad = 0, sbad = 0;
    uint32_t u = start;
    for (long i = 0; i < n; ++i, u += stride) {
        float x;
        memcpy(&x, &u, 4);
        if (!(x <= 0.0f)) continue; /* also skips NaN */
        const float a = b200_expf(x), b = b200_expf_nonpos(x);
        if (memcmp(&a, &b, 4) != 0) ++bad;
        const float xc = x < -86.0f ? -86.0f : x;
        const float t = xc * 1.44269504088896341f;
        const float tb = t + 12582912.0f;
        const float nn = tb - 12582912.0f;
        uint32_t bits;
        memcpy(&bits, &tb, 4);
        const uint32_t trick = (bits << 23) + 0x3f800000u;
        const uint32_t conv = (uint32_t)((int32_t)nn + 127) << 23;
        if (trick != conv) ++sbad;
    }
    *scale_mismatch = sbad;
    return bad;
}
float crf_math_logf(float x) { return b200_logf(x); }
float crf_math_log1pf(float x) { return b200_log1pf(x); }
float crf_math_pow0p4f(float x) { return b200_pow0p4f(x); }
float crf_math_lse2(float x, float y) { return b200_log_sum_exp(x, y); }
float crf_half_to_float(uint16_t h) { return half_to_float(h); }

/* quality-character quantiser (include/b200_crf_math.h): table construction and an audit of the table against the
 * reference expression on every `stride`-th positive float up to 1.0 plus a dense window around every edge. */
int crf_qtable_build(float scale, float shift, b200_qtable* tb) { return b200_qtable_build(scale, shift, tb); }
char crf_qtable_lookup(const b200_qtable* tb, float base_prob, float total_prob) {
    return b200_qtable_lookup(tb, base_prob, total_prob);
}
char crf_qchar_libm(float err, float scale, float shift) { return b200_qchar_libm(err, scale, shift); }
static char qtable_at(const b200_qtable* tb, uint32_t u) {
    uint32_t lo = 0, hi = tb->n;
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (tb->edge[mid] <= u) lo = mid + 1; else hi = mid;
    }
    return (char)tb->ch[lo];
}
long crf_qtable_audit(float scale, float shift, uint32_t stride, uint32_t window) {
    b200_qtable tb;
    if (b200_qtable_build(scale, shift, &tb) != 0) return -1;
    long bad = 0;
    for (uint64_t u = 1; u <= 0x3f800000u; u += stride) {
        bad += qtable_at(&tb, (uint32_t)u) != b200_qchar_libm(B200_U2F((uint32_t)u), scale, shift);
    }
    for (uint32_t i = 0; i < tb.n; ++i) {
        const uint32_t c = tb.edge[i];
        const uint32_t a = c > window ? c - window : 1, b = c + window < 0x3f800000u ? c + window : 0x3f800000u;
        for (uint32_t u = a; u <= b; ++u) bad += qtable_at(&tb, u) != b200_qchar_libm(B200_U2F(u), scale, shift);
    }
    return bad;
}
/* disagreements of the pinned pow0p4 with the host powf and with the correctly rounded value, on every stride-th float
 * in [2^-40, 1) */
void crf_pow0p4_audit(uint32_t stride, long* n, long* vs_libm, long* vs_exact, long* max_ulp_libm) {
    long c = 0, a = 0, b = 0, mu = 0;
    for (uint32_t u = 0x2b800000u; u < 0x3f800000u; u += stride) {
        const float x = B200_U2F(u);
        const float mine = b200_pow0p4f(x);
        const float lm = powf(x, 0.4f);
        const float ex = (float)pow((double)x, (double)0.4f);
        if (mine != lm) {
            ++a;
            long d = (long)B200_F2U(mine) - (long)B200_F2U(lm);
            if (d < 0) d = -d;
            if (d > mu) mu = d;
        }
        b += mine != ex;
        ++c;
    }
    *n = c; *vs_libm = a; *vs_exact = b; *max_ulp_libm = mu;
}

// CRF decode for sm_100a: backward scan -> fused forward scan / posteriors / beam search -> traceback.
//
// Replaces the closed Koi kernels behind dorado/basecall/decode/CUDADecoder.cpp:76-104
// (host_back_guide_step, host_beam_search_step, host_compute_posts_step, host_run_decode) and the CPU
// slicing of CUDADecoder.cpp:115-173.  Semantics are those of the reference's open CPU decoder:
//   scans          dorado/basecall/decode/CPUDecoder.cpp:17-92
//   posteriors     CPUDecoder.cpp:130
//   beam search    dorado/basecall/decode/beam_search.cpp:125-520
//   sequence/qual  beam_search.cpp:54-102
// All floating point goes through include/b200_crf_math.h so results are bit-identical to the CPU
// oracle (oracle/crf_oracle.c).
//
// Data layout in HBM (per batch of N chunks, T blocks, S = 4^state_len states, C = 4S):
//   scores  fp16  [N][T][C]        read twice (once per scan direction), coalesced 32 B per thread
//   bwd     fp32  [N][T+1][S]      written by kernel 1, read once by kernel 2
//   beam    8 B   [N][T][32]       {state:16, prev:8, stay:8, block_prob:f32} per kept element
//   out     u8    moves/seq/qstr [N][T], n_bases i32 [N]
// Thread mapping: S/4 threads per chunk; in the backward scan thread q owns states {q + k*S/4},
// in the forward scan states {4q..4q+3}; both touch the same 16 contiguous scores per block.
#include "decode.h"

#include "b200_crf_math.h"
#include "common.cuh"
#include "engine.h"

namespace b200 {

namespace {

constexpr int kBeamW = 32;
constexpr uint32_t kCrcPoly = 0x82f63b78u;
constexpr uint32_t kCrcSeed = 0x12345678u;

template <int SL>
struct Dims {
    static constexpr int S = 1 << (2 * SL);
    static constexpr int P = S / 4;
    static constexpr int C = 4 * S;
};

__device__ __forceinline__ void unpack16(const uint4& r0, const uint4& r1, float clamp_val, float* sc) {
    const uint32_t w[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const __half2 h = *reinterpret_cast<const __half2*>(&w[i]);
        const float2 f = __half22float2(h);
        sc[2 * i] = f.x;
        sc[2 * i + 1] = f.y;
    }
    if (clamp_val > 0.0f) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            sc[i] = sc[i] < -clamp_val ? -clamp_val : (sc[i] > clamp_val ? clamp_val : sc[i]);
        }
    }
}

template <int GT>
__device__ __forceinline__ void group_sync(int g) {
    if constexpr (GT <= 32) {
        __syncwarp();
    } else {
        named_bar_sync(1 + g, GT);
    }
}

// ------------------------------------------------------------------------------------------------
// Kernel 1: backward scan (CPUDecoder.cpp:69-92).
// ------------------------------------------------------------------------------------------------
template <int SL>
__global__ void __launch_bounds__(256) crf_bwd_scan_kernel(const __half* __restrict__ scores,
                                                           float* __restrict__ bwd,
                                                           int N,
                                                           int T,
                                                           float clamp_val,
                                                           float blank) {
    constexpr int S = Dims<SL>::S, P = Dims<SL>::P, C = Dims<SL>::C;
    constexpr int GROUPS = (256 / P) > 0 ? (256 / P) : 1;
    constexpr int GT = P < 32 ? 32 : P;  // sync granularity (sub-warp groups share a warp)
    constexpr int PF = 4;
    constexpr int RS = C / 8;  // row stride in uint4

    const int g = threadIdx.x / P;
    const int q = threadIdx.x % P;
    const int chunk = blockIdx.x * GROUPS + g;
    __shared__ __align__(16) float a[GROUPS][2][S];

    const bool active = chunk < N;
    const int chunk_c = active ? chunk : (N - 1);  // inactive groups shadow the last chunk, no stores
    const uint4* srow = reinterpret_cast<const uint4*>(scores + (size_t)chunk_c * T * C) + 2 * q;
    float* out = bwd + (size_t)chunk_c * (T + 1) * S;

    uint4 pf[PF][2];
#pragma unroll
    for (int k = 0; k < PF; ++k) {
        const int tt = T - 1 - k;
        if (tt >= 0) {
            pf[k][0] = ldg_nc_v4(srow + (size_t)tt * RS);
            pf[k][1] = ldg_nc_v4(srow + (size_t)tt * RS + 1);
        }
    }
    float own[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        own[k] = 0.0f;
        a[g][0][q + k * P] = 0.0f;
        if (active) out[(size_t)T * S + q + k * P] = 0.0f;
    }
    group_sync<GT>(g);

    int cur = 0;
    for (int t = T - 1; t >= 0; t -= PF) {
#pragma unroll
        for (int k = 0; k < PF; ++k) {
            const int tt = t - k;
            if (tt >= 0) {
                const uint4 r0 = pf[k][0], r1 = pf[k][1];
                if (tt - PF >= 0) {
                    pf[k][0] = ldg_nc_v4(srow + (size_t)(tt - PF) * RS);
                    pf[k][1] = ldg_nc_v4(srow + (size_t)(tt - PF) * RS + 1);
                }
                float sc[16];
                unpack16(r0, r1, clamp_val, sc);
                const float4 nx = *reinterpret_cast<const float4*>(&a[g][cur][4 * q]);
#pragma unroll
                for (int top = 0; top < 4; ++top) {
                    own[top] = b200_lse5(B200_ADD(own[top], blank), B200_ADD(nx.x, sc[0 + top]),
                                         B200_ADD(nx.y, sc[4 + top]), B200_ADD(nx.z, sc[8 + top]),
                                         B200_ADD(nx.w, sc[12 + top]));
                    a[g][cur ^ 1][q + top * P] = own[top];
                    if (active) out[(size_t)tt * S + q + top * P] = own[top];
                }
                group_sync<GT>(g);
                cur ^= 1;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Kernel 2: forward scan + posteriors + beam search.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t crc_bits(uint32_t crc, uint32_t nb, int nbits) {
    for (int i = 0; i < nbits; ++i) {
        const uint32_t b = (nb ^ crc) & 1u;
        crc >>= 1;
        if (b) crc ^= kCrcPoly;
        nb >>= 1;
    }
    return crc;
}

__device__ __forceinline__ uint32_t crc2(uint32_t crc, uint32_t nb) {
    uint32_t b = (nb ^ crc) & 1u;
    crc = (crc >> 1) ^ (b ? kCrcPoly : 0u);
    b = ((nb >> 1) ^ crc) & 1u;
    crc = (crc >> 1) ^ (b ? kCrcPoly : 0u);
    return crc;
}

struct BeamSmem {
    float cand_score[5 * kBeamW];
    uint32_t cand_hash[5 * kBeamW];
    float new_score[kBeamW];
    uint32_t new_hash[kBeamW];
    uint32_t new_meta[kBeamW];  // state | prev << 16 | stay << 24
};

struct BeamLane {
    uint32_t hash;
    uint32_t state;
    float score;
};

__device__ __forceinline__ uint32_t float_key(float f) {
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// Initial beam: the W best states of bwd[0] in state order (beam_search.cpp:166-200).
template <int SL>
__device__ int beam_init(const float* bw_row, BeamSmem& bs, BeamLane& me, int W, int lane) {
    constexpr int S = Dims<SL>::S;
    constexpr int VPL = S >= 32 ? S / 32 : 1;
    uint32_t keys[VPL];
#pragma unroll
    for (int j = 0; j < VPL; ++j) {
        const int s = lane * VPL + j;
        keys[j] = s < S ? float_key(bw_row[s]) : 0u;
    }
    uint32_t thr = 0u;  // key >= 0 always true: everything selected when W >= S
    if (W < S) {
        for (int bit = 31; bit >= 0; --bit) {
            const uint32_t cand = thr | (1u << bit);
            int c = 0;
#pragma unroll
            for (int j = 0; j < VPL; ++j) c += (lane * VPL + j < S) && (keys[j] >= cand);
            if (warp_sum_int(c) >= W) thr = cand;
        }
    }
    int c = 0;
#pragma unroll
    for (int j = 0; j < VPL; ++j) c += (lane * VPL + j < S) && (keys[j] >= thr);
    int incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    int pos = incl - c;
#pragma unroll
    for (int j = 0; j < VPL; ++j) {
        const int s = lane * VPL + j;
        if (s < S && keys[j] >= thr) {
            if (pos < W) bs.new_meta[pos] = (uint32_t)s;
            ++pos;
        }
    }
    __syncwarp();
    const int width = W < S ? W : S;
    if (lane < width) {
        me.state = bs.new_meta[lane] & 0xffffu;
        me.hash = crc_bits(kCrcSeed, me.state, 32);
        me.score = 0.0f;
    }
    __syncwarp();
    return width;
}

// One block of the beam search for one chunk, executed by one warp (lane = beam element).
// Returns the new beam width; writes the kept elements (and their block probabilities) to beam_row.
template <int SL>
__device__ int beam_step(const float* sc_row,
                         const float* bw_row,
                         const float* post_row,
                         BeamSmem& bs,
                         BeamLane& me,
                         int width,
                         int W,
                         float log_beam_cut,
                         float blank,
                         bool last_block,
                         uint2* beam_row,
                         int lane) {
    constexpr int S = Dims<SL>::S;
    constexpr int SB = 2 * SL;
    constexpr uint32_t mask = S - 1;
    const bool valid = lane < width;

    // --- candidates (beam_search.cpp:225-262) ---
    uint32_t ns[4], hs[4];
    float lmax = B200_FLT_LOWEST;
    if (valid) {
        const uint32_t shifted = me.state << 2;
#pragma unroll
        for (uint32_t b = 0; b < 4; ++b) {
            ns[b] = (shifted & mask) | b;
            const uint32_t move_idx = ((ns[b] << 2) + (shifted >> SB)) & 0xffffu;
            const float v = B200_ADD(B200_ADD(me.score, sc_row[move_idx]), bw_row[ns[b]]);
            hs[b] = crc2(me.hash, b);
            bs.cand_score[lane * 4 + b] = v;
            bs.cand_hash[lane * 4 + b] = hs[b];
            lmax = b200_fmaxf(lmax, v);
        }
        const float sv = B200_ADD(B200_ADD(me.score, blank), bw_row[me.state]);
        bs.cand_score[4 * width + lane] = sv;
        bs.cand_hash[4 * width + lane] = me.hash;
        lmax = b200_fmaxf(lmax, sv);
    }
    float max_score = warp_max(lmax);
    __syncwarp();

    // --- merge stays with equal-hash steps (beam_search.cpp:264-305) ---
    bool dup = false;
    const uint32_t vmask = __ballot_sync(0xffffffffu, valid);
    if (valid) {
        dup = __popc(__match_any_sync(vmask, me.hash)) > 1;
    }
    if (!__any_sync(0xffffffffu, dup)) {
        // all hashes distinct: each stay can match at most one step and no step is shared
        float folded = B200_FLT_LOWEST;
        if (valid) {
            const int latest = me.state & 3;
            int jm = -1;
            for (int j = 0; j < width; ++j) {
                if (bs.cand_hash[j * 4 + latest] == me.hash) jm = j;
            }
            if (jm >= 0) {
                const int si = 4 * width + lane, pi = jm * 4 + latest;
                const float st = bs.cand_score[si], sp = bs.cand_score[pi];
                folded = b200_log_sum_exp(st, sp);
                if (st > sp) {
                    bs.cand_score[si] = folded;
                    bs.cand_score[pi] = B200_FLT_LOWEST;
                } else {
                    bs.cand_score[pi] = folded;
                    bs.cand_score[si] = B200_FLT_LOWEST;
                }
            }
        }
        max_score = b200_fmaxf(max_score, warp_max(folded));
    } else {
        // hash collision between beam elements (rare): replay the reference's sequential order
        float m2 = max_score;
        // gather the latest bases through shared memory so lane 0 can see them
        bs.new_meta[lane] = valid ? (me.state & 3u) : 0u;
        __syncwarp();
        if (lane == 0) {
            for (int i = 0; i < width; ++i) {
                const int si = 4 * width + i;
                const int latest = (int)bs.new_meta[i];
                for (int j = 0; j < width; ++j) {
                    const int pi = j * 4 + latest;
                    if (bs.cand_hash[si] == bs.cand_hash[pi]) {
                        const float st = bs.cand_score[si], sp = bs.cand_score[pi];
                        const float folded = b200_log_sum_exp(st, sp);
                        if (st > sp) {
                            bs.cand_score[si] = folded;
                            bs.cand_score[pi] = B200_FLT_LOWEST;
                        } else {
                            bs.cand_score[pi] = folded;
                            bs.cand_score[si] = B200_FLT_LOWEST;
                        }
                        m2 = b200_fmaxf(m2, folded);
                    }
                }
            }
        }
        max_score = __shfl_sync(0xffffffffu, m2, 0);
    }
    __syncwarp();

    float s[5];
    if (valid) {
#pragma unroll
        for (int b = 0; b < 4; ++b) s[b] = bs.cand_score[lane * 4 + b];
        s[4] = bs.cand_score[4 * width + lane];
    }

    // --- cutoff (beam_search.cpp:310-396) ---
    float cutoff = B200_SUB(max_score, log_beam_cut);
    auto count_ge = [&](float c) {
        int n = 0;
        if (valid) {
#pragma unroll
            for (int k = 0; k < 5; ++k) n += (s[k] >= c);
        }
        return warp_sum_int(n);
    };
    int cnt = count_ge(cutoff);
    if (cnt > W) {
        const int min_w = (W * 8) / 10;
        float lo = cutoff, hi = max_score;
        int guesses = 1;
        while ((cnt > W || cnt < min_w) && guesses < 10) {
            if (cnt > W) {
                lo = cutoff;
                cutoff = B200_DIV(B200_ADD(cutoff, hi), 2.0f);
            } else {
                hi = cutoff;
                cutoff = B200_DIV(B200_ADD(cutoff, lo), 2.0f);
            }
            cnt = count_ge(cutoff);
            ++guesses;
        }
        if (guesses == 10) {
            cutoff = hi;
            cnt = count_ge(cutoff);
        }
        if (cnt > W) cnt = W;
    }

    // --- keep the first W candidates >= cutoff in candidate order (beam_search.cpp:398-409) ---
    bool f[5];
    int c = 0;
#pragma unroll
    for (int k = 0; k < 5; ++k) f[k] = valid && (s[k] >= cutoff);
#pragma unroll
    for (int k = 0; k < 4; ++k) c += f[k];
    int incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    const int steps_total = __shfl_sync(0xffffffffu, incl, 31);
    int pos = incl - c;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
        if (f[b]) {
            if (pos < W) {
                bs.new_score[pos] = s[b];
                bs.new_hash[pos] = hs[b];
                bs.new_meta[pos] = ns[b] | ((uint32_t)lane << 16);
            }
            ++pos;
        }
    }
    const uint32_t stay_mask = __ballot_sync(0xffffffffu, f[4]);
    if (f[4]) {
        const int sp = steps_total + __popc(stay_mask & ((1u << lane) - 1u));
        if (sp < W) {
            bs.new_score[sp] = s[4];
            bs.new_hash[sp] = me.hash;
            bs.new_meta[sp] = me.state | ((uint32_t)lane << 16) | (1u << 24);
        }
    }
    __syncwarp();

    const bool nvalid = lane < cnt;
    uint32_t meta = 0;
    if (nvalid) {
        me.score = bs.new_score[lane];
        me.hash = bs.new_hash[lane];
        meta = bs.new_meta[lane];
        me.state = meta & 0xffffu;
    }

    // --- last block: best element to slot 0 (beam_search.cpp:413-424) ---
    if (last_block) {
        const float best = warp_max(nvalid ? me.score : B200_FLT_LOWEST);
        // reference scans ascending with a strict '>' starting from lowest(): first index holding
        // the maximum, or index 0 if every score equals lowest()
        const uint32_t eq = __ballot_sync(0xffffffffu, nvalid && me.score == best && best > B200_FLT_LOWEST);
        const int bi = eq ? (__ffs(eq) - 1) : 0;
        const int src = lane == 0 ? bi : (lane == bi ? 0 : lane);
        const float sc2 = __shfl_sync(0xffffffffu, me.score, src);
        const uint32_t h2 = __shfl_sync(0xffffffffu, me.hash, src);
        const uint32_t m2 = __shfl_sync(0xffffffffu, meta, src);
        me.score = sc2;
        me.hash = h2;
        meta = m2;
        me.state = meta & 0xffffu;
    }

    if (nvalid) {
        me.score = B200_SUB(me.score, bw_row[me.state]);
        // block probability of this element's kmer (beam_search.cpp:459-503)
        const int state = (int)me.state;
        float prob = post_row[state];
        int sh[8];
        const int l = state >> 2;
        const int r = (state << 2) & (S - 1);
        const int msb = S >> 2;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            sh[2 * b] = l + msb * b;
            sh[2 * b + 1] = r + b;
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            bool count = sh[k] != state;
#pragma unroll
            for (int j = 0; j < k; ++j) count = count && (sh[j] != sh[k]);
            if (count) prob = B200_ADD(prob, post_row[sh[k]]);
        }
        prob = prob < 0.0f ? 0.0f : (prob > 1.0f ? 1.0f : prob);
        prob = b200_pow0p4f(prob);
        beam_row[lane] = make_uint2(meta, __float_as_uint(prob));
    }
    __syncwarp();
    return cnt;
}

template <int SL>
__global__ void __launch_bounds__(Dims<SL>::P < 32 ? 128 : (Dims<SL>::P < 256 ? 128 : 256))
        crf_fwd_beam_kernel(const __half* __restrict__ scores,
                            const float* __restrict__ bwd,
                            uint2* __restrict__ beam,
                            int N,
                            int T,
                            float clamp_val,
                            float blank,
                            int W,
                            float log_beam_cut) {
    constexpr int S = Dims<SL>::S, P = Dims<SL>::P, C = Dims<SL>::C;
    constexpr int GT = P < 32 ? 32 : P;
    constexpr int THREADS = P < 256 ? 128 : 256;
    constexpr int GROUPS = THREADS / GT;
    constexpr int NW = GT / 32;  // warps per group
    constexpr int PF = 4;
    constexpr int RS = C / 8;

    const int g = threadIdx.x / GT;
    const int q = threadIdx.x % GT;
    const int lane = threadIdx.x & 31;
    const int wg = q >> 5;  // warp within group
    const int chunk = blockIdx.x * GROUPS + g;
    const bool scan_thread = q < P;

    __shared__ __align__(16) float fa[GROUPS][2][S];
    __shared__ __align__(16) float sc_row[GROUPS][C];
    __shared__ __align__(16) float bw_row[GROUPS][S];
    __shared__ __align__(16) float post_row[GROUPS][S];
    __shared__ float red[GROUPS][2][NW];
    __shared__ BeamSmem bsm[GROUPS];

    if (chunk >= N) return;  // whole groups exit together; barriers are per group

    const int qs = scan_thread ? q : 0;
    const uint4* srow = reinterpret_cast<const uint4*>(scores + (size_t)chunk * T * C) + 2 * qs;
    const float4* brow = reinterpret_cast<const float4*>(bwd + (size_t)chunk * (T + 1) * S) + qs;
    constexpr int BRS = S / 4;
    uint2* beam_out = beam + (size_t)chunk * T * kBeamW;

    // beam init from bwd[0]
    if (scan_thread) {
        const float4 b0 = brow[0];
        *reinterpret_cast<float4*>(&bw_row[g][4 * q]) = b0;
#pragma unroll
        for (int j = 0; j < 4; ++j) fa[g][0][4 * q + j] = 0.0f;
    }
    group_sync<GT>(g);
    BeamLane me{0u, 0u, 0.0f};
    int width = 0;
    if (wg == 0) {
        width = beam_init<SL>(bw_row[g], bsm[g], me, W, lane);
    }
    group_sync<GT>(g);

    uint4 pf[PF][2];
    float4 pb[PF];
    if (scan_thread) {
#pragma unroll
        for (int k = 0; k < PF; ++k) {
            if (k < T) {
                pf[k][0] = ldg_nc_v4(srow + (size_t)k * RS);
                pf[k][1] = ldg_nc_v4(srow + (size_t)k * RS + 1);
                pb[k] = ldg_nc_f4(brow + (size_t)(k + 1) * BRS);
            }
        }
    }
    float f[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    int cur = 0;

    for (int t0 = 0; t0 < T; t0 += PF) {
#pragma unroll
        for (int k = 0; k < PF; ++k) {
            const int t = t0 + k;
            if (t < T) {
                float e[4];
                float lmax = B200_FLT_LOWEST;
                if (scan_thread) {
                    const uint4 r0 = pf[k][0], r1 = pf[k][1];
                    const float4 bw = pb[k];
                    if (t + PF < T) {
                        pf[k][0] = ldg_nc_v4(srow + (size_t)(t + PF) * RS);
                        pf[k][1] = ldg_nc_v4(srow + (size_t)(t + PF) * RS + 1);
                        pb[k] = ldg_nc_f4(brow + (size_t)(t + PF + 1) * BRS);
                    }
                    float sc[16];
                    unpack16(r0, r1, clamp_val, sc);
                    float4* dst = reinterpret_cast<float4*>(&sc_row[g][16 * q]);
                    dst[0] = make_float4(sc[0], sc[1], sc[2], sc[3]);
                    dst[1] = make_float4(sc[4], sc[5], sc[6], sc[7]);
                    dst[2] = make_float4(sc[8], sc[9], sc[10], sc[11]);
                    dst[3] = make_float4(sc[12], sc[13], sc[14], sc[15]);
                    *reinterpret_cast<float4*>(&bw_row[g][4 * q]) = bw;
                    const float p0 = fa[g][cur][q], p1 = fa[g][cur][q + P], p2 = fa[g][cur][q + 2 * P],
                                p3 = fa[g][cur][q + 3 * P];
                    const float bwv[4] = {bw.x, bw.y, bw.z, bw.w};
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        f[j] = b200_lse5(B200_ADD(f[j], blank), B200_ADD(p0, sc[4 * j + 0]),
                                         B200_ADD(p1, sc[4 * j + 1]), B200_ADD(p2, sc[4 * j + 2]),
                                         B200_ADD(p3, sc[4 * j + 3]));
                        e[j] = B200_ADD(f[j], bwv[j]);  // v = fwd + bwd
                        lmax = b200_fmaxf(lmax, e[j]);
                    }
                }
                // max over the group's states
                float mx = warp_max(lmax);
                if constexpr (NW > 1) {
                    if (lane == 0) red[g][0][wg] = mx;
                    group_sync<GT>(g);
                    mx = red[g][0][0];
#pragma unroll
                    for (int w = 1; w < NW; ++w) mx = b200_fmaxf(mx, red[g][0][w]);
                }
                // sum of exp in the contract's order (see oracle/crf_oracle.c posts_row)
                float part = 0.0f;
                if (scan_thread) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) e[j] = b200_expf(B200_SUB(e[j], mx));
                    part = B200_ADD(B200_ADD(B200_ADD(e[0], e[1]), e[2]), e[3]);
                }
#pragma unroll
                for (int o = (P < 32 ? P / 2 : 16); o >= 1; o >>= 1) {
                    part = B200_ADD(part, __shfl_xor_sync(0xffffffffu, part, o));
                }
                float z = part;
                if constexpr (NW > 1) {
                    if (lane == 0) red[g][1][wg] = part;
                    group_sync<GT>(g);
                    z = red[g][1][0];
#pragma unroll
                    for (int w = 1; w < NW; ++w) z = B200_ADD(z, red[g][1][w]);
                } else if constexpr (P < 32) {
                    z = __shfl_sync(0xffffffffu, part, 0);
                }
                if (scan_thread) {
                    *reinterpret_cast<float4*>(&post_row[g][4 * q]) =
                            make_float4(B200_DIV(e[0], z), B200_DIV(e[1], z), B200_DIV(e[2], z), B200_DIV(e[3], z));
                    *reinterpret_cast<float4*>(&fa[g][cur ^ 1][4 * q]) = make_float4(f[0], f[1], f[2], f[3]);
                }
                group_sync<GT>(g);
                if (wg == 0) {
                    width = beam_step<SL>(sc_row[g], bw_row[g], post_row[g], bsm[g], me, width, W, log_beam_cut,
                                          blank, t == T - 1, beam_out + (size_t)t * kBeamW, lane);
                }
                group_sync<GT>(g);
                cur ^= 1;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Kernel 3: traceback + sequence / qstring generation (beam_search.cpp:447-455, :54-102).
// ------------------------------------------------------------------------------------------------
constexpr int kTbTile = 32;
constexpr int kTbWarps = 2;

__global__ void __launch_bounds__(kTbWarps * 32) crf_traceback_kernel(const uint2* __restrict__ beam,
                                                                      int N,
                                                                      int T,
                                                                      float q_scale,
                                                                      float q_shift,
                                                                      uint8_t* __restrict__ moves_out,
                                                                      char* __restrict__ seq_out,
                                                                      char* __restrict__ qstr_out,
                                                                      int32_t* __restrict__ n_bases_out) {
    extern __shared__ __align__(16) unsigned char tb_smem[];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int chunk = blockIdx.x * kTbWarps + w;
    if (chunk >= N) return;
    const int Tp = (T + 3) & ~3;
    // per-warp carve-up
    const size_t per_warp = (size_t)kTbTile * kBeamW * sizeof(uint2) + (size_t)Tp * (4 + 2 + 2 + 1) + 16;
    unsigned char* base = tb_smem + (size_t)w * ((per_warp + 15) & ~(size_t)15);
    uint2* tile = reinterpret_cast<uint2*>(base);
    float* pprob = reinterpret_cast<float*>(base + (size_t)kTbTile * kBeamW * sizeof(uint2));
    uint16_t* pstate = reinterpret_cast<uint16_t*>(pprob + Tp);
    uint16_t* bstart = pstate + Tp;
    uint8_t* pmove = reinterpret_cast<uint8_t*>(bstart + Tp + 2);

    const uint2* brow = beam + (size_t)chunk * T * kBeamW;
    uint32_t ei = 0;
    for (int t_hi = T; t_hi > 0; t_hi -= kTbTile) {
        const int t_lo = t_hi - kTbTile > 0 ? t_hi - kTbTile : 0;
        const int rows = t_hi - t_lo;
        for (int r = 0; r < rows; ++r) tile[r * kBeamW + lane] = brow[(size_t)(t_lo + r) * kBeamW + lane];
        __syncwarp();
        if (lane == 0) {
            for (int r = rows - 1; r >= 0; --r) {
                const uint2 e = tile[r * kBeamW + ei];
                const int t = t_lo + r;
                pstate[t] = (uint16_t)(e.x & 0xffffu);
                pmove[t] = ((e.x >> 24) & 1u) ? 0 : 1;
                pprob[t] = __uint_as_float(e.y);
                ei = (e.x >> 16) & 0xffu;
            }
        }
        __syncwarp();
    }
    if (lane == 0) pmove[0] = 1;  // always step in the first block
    __syncwarp();

    // base start blocks
    int nb = 0;
    for (int t0 = 0; t0 < T; t0 += 32) {
        const int t = t0 + lane;
        const bool m = t < T && pmove[t];
        const uint32_t bal = __ballot_sync(0xffffffffu, m);
        if (m) bstart[nb + __popc(bal & ((1u << lane) - 1u))] = (uint16_t)t;
        nb += __popc(bal);
    }
    if (lane == 0) bstart[nb] = (uint16_t)T;
    __syncwarp();

    uint8_t* mo = moves_out + (size_t)chunk * T;
    char* so = seq_out + (size_t)chunk * T;
    char* qo = qstr_out + (size_t)chunk * T;
    for (int t = lane; t < T; t += 32) mo[t] = pmove[t];
    for (int p = lane; p < T; p += 32) {
        char sc = 0, qc = 0;
        if (p < nb) {
            const int b0 = bstart[p], b1 = bstart[p + 1];
            float bp = 0.0f, tp = 0.0f;
            int base = 0;
            for (int blk = b0; blk < b1; ++blk) {
                base = pstate[blk] & 3;
                const float prob = pprob[blk];
                const float wrong = B200_DIV(B200_SUB(1.0f, prob), 3.0f);
                bp = B200_ADD(bp, prob);
#pragma unroll
                for (int k = 0; k < 4; ++k) tp = B200_ADD(tp, k == base ? prob : wrong);
            }
            sc = "ACGT"[pstate[b0] & 3];
            qc = b200_qchar(bp, tp, q_scale, q_shift);
        }
        so[p] = sc;
        qo[p] = qc;
    }
    if (lane == 0) n_bases_out[chunk] = nb;
}

size_t traceback_smem_bytes(int T) {
    const int Tp = (T + 3) & ~3;
    const size_t per_warp = (size_t)kTbTile * kBeamW * sizeof(uint2) + (size_t)Tp * (4 + 2 + 2 + 1) + 16;
    return kTbWarps * ((per_warp + 15) & ~(size_t)15);
}

template <int SL>
void launch_decode(const DecodeArgs& a, cudaStream_t stream, ProfileSink* prof) {
    constexpr int P = Dims<SL>::P;
    {
        constexpr int GROUPS = (256 / P) > 0 ? (256 / P) : 1;
        const int grid = (a.N + GROUPS - 1) / GROUPS;
        crf_bwd_scan_kernel<SL><<<grid, P * GROUPS, 0, stream>>>(a.scores, a.bwd, a.N, a.T, a.clamp_val, a.blank);
        if (prof) prof->mark("crf_bwd_scan", stream);
    }
    {
        constexpr int GT = P < 32 ? 32 : P;
        constexpr int THREADS = P < 256 ? 128 : 256;
        constexpr int GROUPS = THREADS / GT;
        const int grid = (a.N + GROUPS - 1) / GROUPS;
        crf_fwd_beam_kernel<SL><<<grid, THREADS, 0, stream>>>(a.scores, a.bwd, a.beam, a.N, a.T, a.clamp_val,
                                                                a.blank, a.beam_width, a.log_beam_cut);
        if (prof) prof->mark("crf_fwd_beam", stream);
    }
    {
        const size_t smem = traceback_smem_bytes(a.T);
        static bool attr_set = false;
        if (!attr_set && smem > 48 * 1024) {
            B200_CUDA(cudaFuncSetAttribute(crf_traceback_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           200 * 1024));
            attr_set = true;
        }
        const int grid = (a.N + kTbWarps - 1) / kTbWarps;
        crf_traceback_kernel<<<grid, kTbWarps * 32, smem, stream>>>(a.beam, a.N, a.T, a.q_scale, a.q_shift, a.moves,
                                                                    a.sequence, a.qstring, a.n_bases);
        if (prof) prof->mark("crf_traceback", stream);
    }
    B200_CUDA(cudaGetLastError());
}

}  // namespace

size_t decode_scratch_bytes(int N, int T, int state_len, size_t* bwd_bytes, size_t* beam_bytes) {
    const size_t S = (size_t)1 << (2 * state_len);
    const size_t b1 = ((size_t)N * (T + 1) * S * sizeof(float) + 255) & ~(size_t)255;
    const size_t b2 = ((size_t)N * T * kBeamW * sizeof(uint2) + 255) & ~(size_t)255;
    if (bwd_bytes) *bwd_bytes = b1;
    if (beam_bytes) *beam_bytes = b2;
    return b1 + b2;
}

void decode_scores(const DecodeArgs& a, cudaStream_t stream, ProfileSink* prof) {
    if (a.beam_width < 1 || a.beam_width > kBeamW) {
        throw std::invalid_argument("b200 decode: beam_width must be in [1, 32]");
    }
    if (a.T < 1 || a.T > 65535 || a.N < 1) {
        throw std::invalid_argument("b200 decode: need 1 <= T <= 65535 and N >= 1");
    }
    switch (a.state_len) {
        case 3: launch_decode<3>(a, stream, prof); break;
        case 4: launch_decode<4>(a, stream, prof); break;
        case 5: launch_decode<5>(a, stream, prof); break;
        default: throw std::invalid_argument("b200 decode: state_len must be 3, 4 or 5");
    }
}

}  // namespace b200

// Conv -> LSTM -> linear CRF forward (fast / hac models) for sm_100a.
//
// Replaces, for the CUDA path of dorado/basecall/model/CRFModel.cpp:69-115:
//   ConvStack layers 1,2   host_convolution_f16              dorado/nn/ConvStack.cpp:216-232  -> conv12_tc_kernel (conv2 on
//                                                                                         tcgen05; v5 shapes), conv12_kernel (FMA pipe)
//   ConvStack layer 3      host_linear "cutlass conv"        dorado/nn/ConvStack.cpp:236-275  -> gemm.cu
//   LSTMStack              host_cutlass_lstm/host_small_lstm dorado/nn/LSTMStack.cpp:127-238 -> lstm_layer_kernel (C = 96),
//                                                                                         gx GEMM + lstm_cluster_kernel (C = 192, 384)
//   LinearCRF              host_linear                       dorado/nn/CRFModules.cpp:49-122   -> gemm.cu
// Semantics are those of the CPU modules (ConvStack.cpp:146-163, LSTMStack.cpp:29-41, CRFModules.cpp:24-34).
//
// Activation layouts in HBM (fp16):
//   signal  [N][T_in]
//   x2      [N][T_in + 2*pad3 + 8][16]      conv2 output, NTC, zero rows = conv3's padding (+ K padding)
//   seq     [T_out][N][C]                   time-major LSTM buffer, updated in place by every layer
//   scores  [N][T_out][outsize]
#include "engine.h"
#include "gemm.h"
#include "nvtx.h"
#include "tc.cuh"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace b200 {

namespace {

// ------------------------------------------------------------------------------------------------
// conv1 (1 -> C1, w1) + conv2 (C1 -> 16, w2), both stride 1, fused; fp32 math, fp16 NTC output.
// ------------------------------------------------------------------------------------------------
constexpr int CONV_TT = 256;  // output samples per CTA
constexpr int MAXW = 9;

struct Conv12Params {
    const __half* x;  // [N][T]
    __half* out;      // [N][T_pad][16], row r <-> sample r - front_pad
    const float* w;   // packed: w1 [c1][w1] | b1 [16] | w2 [k][ci][co] (16x16 per tap) | b2 [16]
    int N, T, T_pad, front_pad;
    int c1, w1, w2, act1, act2;
    const int32_t* lens;  // optional per-chunk length in samples (variable chunk sizes): the chunk is zero beyond it
    int* tile_counter;    // conv12_tc_kernel: next tile to hand out (zeroed on the stream before every launch)
};

constexpr int CONV_W_FLOATS = 16 * MAXW + 16 + MAXW * 16 * 16 + 16;

// swish(v) = v / (1 + e^-v), optionally capped at 3.5, or tanh(v) = 1 - 2 / (e^{2v} + 1): one branch-free path (ex2, rcp, select,
// fma, min) whose constants are picked once per thread -- a switch on the run-time activation would be if-converted and every
// element would pay for every variant's transcendentals
struct ConvAct {
    float k, cap;
    bool is_tanh;
};
__device__ __forceinline__ ConvAct conv_act_coef(int act) {
    ConvAct c;
    c.is_tanh = act != B200_ACT_SWISH && act != B200_ACT_SWISH_CLAMP;
    c.k = c.is_tanh ? 2.885390081777927f : -1.4426950408889634f;
    c.cap = act == B200_ACT_SWISH_CLAMP ? 3.5f : __int_as_float(0x7f800000);
    return c;
}
__device__ __forceinline__ float conv_act(float v, const ConvAct& c) {
    const float r = rcp_approx(1.0f + ex2_approx(c.k * v));
    return fminf(c.is_tanh ? fmaf(-2.0f, r, 1.0f) : v * r, c.cap);
}

__global__ void __launch_bounds__(CONV_TT) conv12_kernel(const Conv12Params p) {
    const ConvAct a1 = conv_act_coef(p.act1), a2 = conv_act_coef(p.act2);
    const int n = blockIdx.y;
    const int t0 = blockIdx.x * CONV_TT;
    const int p1 = p.w1 / 2, p2 = p.w2 / 2;
    const int L = p.lens ? min(p.T, __ldg(p.lens + n)) : p.T;  // samples of this chunk (zero padding starts at L)
    constexpr int Y1P = CONV_TT + 2 * MAXW;  // channel-major pitch: consecutive threads -> consecutive banks
    __shared__ float xs[CONV_TT + 4 * MAXW];
    __shared__ float y1[16 * Y1P];
    __shared__ __align__(16) float ws[CONV_W_FLOATS];
    const float* s_w1 = ws;
    const float* s_b1 = ws + 16 * MAXW;
    const float* s_w2 = s_b1 + 16;
    const float* s_b2 = s_w2 + MAXW * 16 * 16;
    for (int i = threadIdx.x; i < CONV_W_FLOATS; i += CONV_TT) ws[i] = __ldg(p.w + i);
    const int nx = CONV_TT + 2 * (p1 + p2);
    for (int i = threadIdx.x; i < nx; i += CONV_TT) {
        const int t = t0 - p1 - p2 + i;
        xs[i] = (t >= 0 && t < L) ? __half2float(p.x[(size_t)n * p.T + t]) : 0.0f;
    }
    __syncthreads();
    const int n1 = CONV_TT + 2 * p2;
    for (int i = threadIdx.x; i < n1; i += CONV_TT) {
        const int t = t0 - p2 + i;  // conv1 output position
        const bool inside = t >= 0 && t < L;
        for (int c = 0; c < p.c1; ++c) {
            float acc = s_b1[c];
            for (int k = 0; k < p.w1; ++k) acc += s_w1[c * p.w1 + k] * xs[i + k];
            y1[c * Y1P + i] = inside ? conv_act(acc, a1) : 0.0f;  // conv2 zero-pads conv1's *output*
        }
    }
    __syncthreads();
    const int t = t0 + threadIdx.x;
    if (t < p.T) {
        float acc[16];
#pragma unroll
        for (int co = 0; co < 16; ++co) acc[co] = s_b2[co];
        for (int k = 0; k < p.w2; ++k) {
            for (int ci = 0; ci < p.c1; ++ci) {
                const float v = y1[ci * Y1P + threadIdx.x + k];
                const float4* w = reinterpret_cast<const float4*>(&s_w2[(k * 16 + ci) * 16]);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float4 wv = w[q];
                    acc[4 * q + 0] += v * wv.x;
                    acc[4 * q + 1] += v * wv.y;
                    acc[4 * q + 2] += v * wv.z;
                    acc[4 * q + 3] += v * wv.w;
                }
            }
        }
        __half2 h[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) h[j] = __floats2half2_rn(conv_act(acc[2 * j], a2), conv_act(acc[2 * j + 1], a2));
        if (t >= L) {  // beyond a short chunk's end: the next convolution's zero padding
#pragma unroll
            for (int j = 0; j < 8; ++j) h[j] = __floats2half2_rn(0.0f, 0.0f);
        }
        uint4* dst = reinterpret_cast<uint4*>(p.out + ((size_t)n * p.T_pad + p.front_pad + t) * 16);
        dst[0] = *reinterpret_cast<uint4*>(&h[0]);
        dst[1] = *reinterpret_cast<uint4*>(&h[4]);
    }
}

// ------------------------------------------------------------------------------------------------
// conv1 + conv2 with conv2 on the tensor cores (the v5 shape: conv1 1 -> 16, conv2 16 -> 16 with 5 taps).
//
// conv2 is 1280 of the 1360 multiply-adds per sample; on the FMA pipe (conv12_kernel above) that is 1.4e10 FLOP per batch of
// 512 x 9996 samples and the kernel sat at 4 % of the HBM roofline it should be bound by (2 B in, 32 B out per sample).  Here a
// tile of 128 output samples is one M = 128, N = 16, K = 80 contraction on tcgen05:
//   * conv1 (+ activation) runs on the FMA pipe, one thread per conv1 position, and its 16 channels are rounded to fp16
//     (every activation tensor between layers is fp16, as in the reference's CUDA path);
//   * the A operand is conv2's im2col tile built directly in shared memory in the 128-byte-swizzled K-major layout:
//     K index = tap * 16 + channel, so the 32 bytes of conv1 position j go to row j - tap, 16-byte chunks 2 tap and 2 tap + 1
//     (five 2 x 16 B stores per thread; the XOR swizzle keeps the eight rows of a phase on distinct banks);
//   * B = the conv2 weights [16][80] fp16 in the same layout, written once per CTA; the accumulator is 16 TMEM columns;
//   * four warps read the accumulator back (thread = output sample), add the bias, apply the activation and store the
//     sample's 16 fp16 channels as one 32-byte row: a warp writes 1 KB contiguous.
// The phases of a tile are separated by CTA barriers; five CTAs fit an SM (37 KB shared memory, 32 TMEM columns, 12288
// registers each) and hide each other's barrier and MMA latency.  The grid is persistent and takes tiles from a counter.
// ------------------------------------------------------------------------------------------------
constexpr int C12_TILE = 128;
// warps 0-3: conv1 producers and epilogue (TMEM lane quarter = warp); warp 4: tail positions + MMA issue; warp 5 only pads the
// CTA to 192 threads x 64 registers = 12288 registers: more than the recurrence kernels leave free on their SMs (fast: 8192,
// hac: 10240).  Those kernels hold all 512 tensor-memory columns for a whole layer, and a CTA of this kernel placed beside one
// would sit in tcgen05.alloc until that layer ends.  Tiles are also handed out dynamically (atomic counter), so a CTA that
// does get stuck holds no work.
constexpr int C12_THREADS = 192;
constexpr int C12_SMEM = 2 * 16384 + 2 * 2048 + 1024;  // A (2 k-blocks of 128 rows x 128 B), B (2 k-blocks of 16 rows), alignment slack

__global__ void __launch_bounds__(C12_THREADS, 5) conv12_tc_kernel(const Conv12Params p, int tiles_per_chunk, int num_tiles) {
    extern __shared__ __align__(1024) uint8_t c12_smem_raw[];
    uint8_t* smem = c12_smem_raw + ((1024u - (tc::smem_u32(c12_smem_raw) & 1023u)) & 1023u);
    uint8_t* a_tile = smem;               // [2][128 rows][128 B]
    uint8_t* b_tile = smem + 2 * 16384;   // [2][16 rows][128 B]
    __shared__ float xs[C12_TILE + 4 + 2 * MAXW];
    __shared__ float s_w1[16 * MAXW], s_b1[16], s_b2[16];
    __shared__ uint64_t mma_done;
    __shared__ uint32_t tmem_holder;
    __shared__ int s_tile;
    const ConvAct a1 = conv_act_coef(p.act1), a2 = conv_act_coef(p.act2);
    const int tid = threadIdx.x, warp = tid >> 5;
    const int p1 = p.w1 / 2;
    constexpr int p2 = 2;

    // ---- once per CTA: conv1 weights / biases, conv2 weights as the fp16 B operand, barrier, tensor memory
    for (int i = tid; i < 16 * MAXW; i += C12_THREADS) s_w1[i] = __ldg(p.w + i);
    if (tid < 16) {
        s_b1[tid] = __ldg(p.w + 16 * MAXW + tid);
        s_b2[tid] = __ldg(p.w + 16 * MAXW + 16 + MAXW * 16 * 16 + tid);
    }
    for (int i = tid; i < 2 * 2048 / 16; i += C12_THREADS) reinterpret_cast<uint4*>(b_tile)[i] = make_uint4(0, 0, 0, 0);
    __syncthreads();
    {
        const float* w2 = p.w + 16 * MAXW + 16;  // [tap][ci][co]
        for (int i = tid; i < 5 * 16 * 16; i += C12_THREADS) {
            const int co = i & 15, ci = (i >> 4) & 15, tap = i >> 8;
            const int c = (tap * 2 + (ci >> 3)) & 7, kb = tap >> 2;
            *reinterpret_cast<__half*>(b_tile + kb * 2048 + tc::sw128_offset(co, c) + (ci & 7) * 2) =
                    __float2half_rn(__ldg(w2 + (tap * 16 + ci) * 16 + co));
        }
    }
    if (tid == 0) {
        tc::mbar_init(&mma_done, 1);
        tc::fence_barrier_init();
    }
    if (warp == 4) tc::tmem_alloc(&tmem_holder, 32);
    tc::fence_proxy_async_smem();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_acc = tmem_holder;
    const uint64_t adesc0 = tc::umma_desc_sw128(tc::smem_u32(a_tile));
    const uint64_t bdesc0 = tc::umma_desc_sw128(tc::smem_u32(b_tile));
    constexpr uint32_t idesc = tc::umma_idesc_f16(128, 16);

    uint32_t parity = 0;
    for (;;) {
        __syncthreads();  // the previous tile's accumulator has been read and its MMAs have read A: both may be overwritten
        if (tid == 0) s_tile = atomicAdd(p.tile_counter, 1);
        __syncthreads();
        const int tile = s_tile;
        if (tile >= num_tiles) break;
        const int n = tile / tiles_per_chunk;
        const int t0 = (tile - n * tiles_per_chunk) * C12_TILE;
        const int L = p.lens ? min(p.T, __ldg(p.lens + n)) : p.T;  // samples of this chunk (zero padding starts at L)
        // ---- signal window: xs[i] <-> sample t0 - p2 - p1 + i
        if (tid < C12_TILE + 2 * (p1 + p2)) {
            const int t = t0 - p2 - p1 + tid;
            xs[tid] = (t >= 0 && t < L) ? __half2float(p.x[(size_t)n * p.T + t]) : 0.0f;
        }
        __syncthreads();
        // ---- conv1 at position j <-> sample t0 - p2 + j, scattered into the im2col rows j - tap
        if (tid < C12_TILE + 2 * p2) {
            const int j = tid;
            const int t = t0 - p2 + j;
            const bool inside = t >= 0 && t < L;  // conv2 zero-pads conv1's *output*
            float xv[MAXW];
#pragma unroll
            for (int k = 0; k < MAXW; ++k) xv[k] = k < p.w1 ? xs[j + k] : 0.0f;
            __half2 h[8];
#pragma unroll
            for (int c2 = 0; c2 < 8; ++c2) {
                float acc0 = s_b1[2 * c2], acc1 = s_b1[2 * c2 + 1];
#pragma unroll
                for (int k = 0; k < MAXW; ++k) {
                    if (k < p.w1) {
                        acc0 = fmaf(s_w1[(2 * c2) * p.w1 + k], xv[k], acc0);
                        acc1 = fmaf(s_w1[(2 * c2 + 1) * p.w1 + k], xv[k], acc1);
                    }
                }
                h[c2] = inside ? __floats2half2_rn(conv_act(acc0, a1), conv_act(acc1, a1)) : __floats2half2_rn(0.0f, 0.0f);
            }
            const uint4 lo = *reinterpret_cast<const uint4*>(&h[0]), hi = *reinterpret_cast<const uint4*>(&h[4]);
#pragma unroll
            for (int tap = 0; tap < 5; ++tap) {
                const int i = j - tap;
                if (i >= 0 && i < C12_TILE) {
                    uint8_t* base = a_tile + (tap >> 2) * 16384;
                    const int c = (tap * 2) & 7;
                    *reinterpret_cast<uint4*>(base + tc::sw128_offset(i, c)) = lo;
                    *reinterpret_cast<uint4*>(base + tc::sw128_offset(i, c + 1)) = hi;
                }
            }
        }
        tc::fence_proxy_async_smem();
        __syncthreads();
        // ---- conv2: five K = 16 steps (one per tap), issued by one thread
        if (tid == 128) {
            tc::tc_fence_after();
#pragma unroll
            for (int tap = 0; tap < 5; ++tap) {
                const uint64_t off = (uint64_t)(((tap >> 2) * 16384) >> 4) + (uint64_t)(2 * (tap & 3));
                const uint64_t boff = (uint64_t)(((tap >> 2) * 2048) >> 4) + (uint64_t)(2 * (tap & 3));
                tc::umma_f16(tmem_acc, adesc0 + off, bdesc0 + boff, idesc, tap != 0);
            }
            tc::umma_commit(&mma_done);
        }
        // ---- epilogue: thread = output sample
        if (warp < 4) {
            tc::mbar_wait(&mma_done, parity);
            tc::tc_fence_after();
            uint32_t r[16];
            tc::tmem_ld_32x16(tmem_acc + ((uint32_t)(warp * 32) << 16), r);
            tc::tmem_ld_wait();
            const int t = t0 + tid;
            if (t < p.T) {
                __half2 o[8];
#pragma unroll
                for (int c2 = 0; c2 < 8; ++c2) {
                    o[c2] = __floats2half2_rn(conv_act(__uint_as_float(r[2 * c2]) + s_b2[2 * c2], a2),
                                              conv_act(__uint_as_float(r[2 * c2 + 1]) + s_b2[2 * c2 + 1], a2));
                }
                if (t >= L) {  // beyond a short chunk's end: the next convolution's zero padding
#pragma unroll
                    for (int c2 = 0; c2 < 8; ++c2) o[c2] = __floats2half2_rn(0.0f, 0.0f);
                }
                uint4* dst = reinterpret_cast<uint4*>(p.out + ((size_t)n * p.T_pad + p.front_pad + t) * 16);
                dst[0] = *reinterpret_cast<uint4*>(&o[0]);
                dst[1] = *reinterpret_cast<uint4*>(&o[4]);
            }
            tc::tc_fence_before();
        }
        parity ^= 1u;
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 4) tc::tmem_dealloc(tmem_acc, 32);
}

// ------------------------------------------------------------------------------------------------
// LSTM layer for a hidden size whose weights fit one SM's tensor memory (fast: C = 96): all T steps of one layer in one
// persistent launch, no inter-CTA traffic.
//
// Orientation is "weights as the M operand": per step, for each tile m of 32 hidden units,
//     D_m[128 gate rows][16] = Wp_m[128][2C] * Z_t[16][2C]^T ,  Z_t = [x_t ; h_{t-1}]  (rows = chunks, zero-padded to 16)
// where the 128 rows of Wp_m are (i | f | g | o) x 32 units, so one epilogue warp handles one gate type.  The permuted
// weights (4C x 2C fp16) are written ONCE into tensor memory (tcgen05.st) and every MMA of the sequence reads its A operand
// from there; the only shared-memory operand is the tiny Z_t: x_t arrives by TMA, h_t is written by the epilogue.
//
// A step is a dependent chain (MMA -> tcgen05.ld -> gates -> cell -> h_t -> MMA) of ~1000 cycles that leaves every pipe of
// the SM idle most of the time, and a batch of 512 chunks only has 512 / NBR such chains.  So one CTA runs NG = 2 chains
// ("groups") side by side: each group owns NBR chunks, its own Z / gate / cell buffers, barriers, accumulator columns and its
// own 16 warps (1 TMA producer, MT MMA issuers, MT x 4 epilogue warps); the groups share nothing but the weights in tensor
// memory, and one group's gate math runs under the other's MMAs and barrier hand-overs.
// Accumulators are double-buffered per group so the x_t half of step s+1 (which does not depend on the recurrence) is
// issued while the epilogue of step s still runs; only the h_{t-1} half sits on the critical path.
// ------------------------------------------------------------------------------------------------
constexpr int KBLK = 32;           // K elements per smem block (64 B rows, SWIZZLE_64B)
constexpr int UN = 16;             // UMMA N: rows of a Z block (chunks, zero-padded when a group owns fewer)
constexpr int ZBLK = UN * KBLK * 2;

struct LstmParams {
    __half* seq;        // [T][N][C] in place
    const float* bias;  // [4C] permuted like the weight rows (b_ih + b_hh)
    const __half* w;    // [4C][2C] permuted weights
    int T, N, C, reverse;
    long long* dbg;  // optional timeline (clock64 stamps of CTA 0 / group 0, steps 64..67); nullptr in production
};

__device__ __forceinline__ uint64_t umma_desc_sw64(uint32_t smem_addr) {
    // K-major, rows of 64 B, 8-row groups 512 B apart, SWIZZLE_64B (layout type 4)
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3ffffu) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(512 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)4 << 61;
    return d;
}

__device__ __forceinline__ uint32_t sw64_offset(int row, int col) {
    // byte offset of fp16 element (row, col<32) in a [rows x 32] SWIZZLE_64B tile
    return (uint32_t)(row * 64 + ((((col >> 3) ^ ((row >> 1) & 3))) << 4) + ((col & 7) << 1));
}

// Gate activations on one MUFU.TANH each: sigmoid(v) = 0.5 tanh(0.5 v) + 0.5 (am = 1), tanh(v) (am = 2); FMUL, MUFU.TANH, FFMA.
// A step of the recurrence applies 5 C activations per chunk, and with ex2 + rcp (two MUFU operations each, ~2^-22) the SFU
// was the busiest unit of the fast kernel (44 % at 2 x 8 chunks per CTA) and capped how many chunks a CTA could carry:
// 1.83 -> 1.60 ms per fast layer, 4.6 -> 4.1 ms per hac layer (profiles/r02_b19_*).  tanh.approx.f32 is good to ~2^-11 -- the
// rounding h_t gets anyway when it is stored as fp16; score errors against both oracles are unchanged
// (profiles/r02_b19_err_*.txt: the gates saturate and contract the error).  The same substitution in the swish of the
// convolutions was rejected, see common.cuh.  -DB200_LSTM_EX2_RCP restores the two-MUFU form.
#ifndef B200_LSTM_EX2_RCP
__device__ __forceinline__ float gate_act(float v, float am) {
    const float s = 0.5f * am;
    return fmaf(tanh_mufu(v * s), s, 1.0f - s);
}
#else
__device__ __forceinline__ float gate_act(float v, float am) {
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(v * (am * 1.4426950408889634f)));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(e + 1.0f));
    return fmaf(-am, r, 1.0f);
}
#endif
__device__ __forceinline__ float tanh_f(float v) { return gate_act(v, 2.0f); }

// C   = hidden size (compile time so the MMA issue loops unroll into immediate-offset descriptors).
// NBR = chunks owned by one group (2, 4, 8 or 16: small values spread a small batch over more SMs).
// NG  = groups (independent recurrences) per CTA.
template <int C, int NG>
struct LstmCfg {
    static constexpr int MT = C / 32;                 // gate tiles = MMA-issuing warps = epilogue warp quartets per group
    static constexpr int KB = 2 * C / KBLK;
    static constexpr int KBX = C / KBLK;
    static constexpr int GROUP_WARPS = 1 + MT + 4 * MT;
    static constexpr int THREADS = 32 * GROUP_WARPS * NG;
    static constexpr int ACC_COLS = NG * 2 * MT * UN;
    static constexpr int NEED_COLS = ACC_COLS + MT * C;   // + the weights: tile m, k-step ks at m*C + ks*8
    static constexpr uint32_t TMEM_COLS = NEED_COLS <= 256 ? 256 : 512;
    static constexpr int NBAR = 8 + 2 * MT;               // mbarriers per group
    static_assert(NEED_COLS <= 512, "weights + accumulators must fit tensor memory");
    static_assert(GROUP_WARPS % 4 == 0, "epilogue warps must keep warp % 4 == TMEM lane quarter");
    static_assert(THREADS <= 1024 && NG * MT + 1 <= 16, "too many warps / named barriers");
    static constexpr size_t smem_bytes(int nbr) {
        return 1024 + (size_t)NG * (2 * KB * ZBLK + (size_t)MT * 4 * nbr * 32 * 4 + (size_t)MT * nbr * 32 * 4 + NBAR * 8) + 64;
    }
};

template <int C, int NBR, int NG>
__global__ void __launch_bounds__(LstmCfg<C, NG>::THREADS, 1) lstm_layer_kernel(const __grid_constant__ CUtensorMap tma_x,
                                                                               const LstmParams p) {
    using Cfg = LstmCfg<C, NG>;
    constexpr int MT = Cfg::MT, KB = Cfg::KB, KBX = Cfg::KBX, GW = Cfg::GROUP_WARPS, NBAR = Cfg::NBAR;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // realign by an integer offset from the __shared__ symbol so the compiler keeps the shared address space
    uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int grp = warp / GW, lw = warp % GW;   // group and role within the group
    uint8_t* z_all = smem;                                                        // [NG][2][KB][ZBLK]
    float* g_all = reinterpret_cast<float*>(z_all + (size_t)NG * 2 * KB * ZBLK);  // [NG][MT][4][NBR][32]
    float* c_all = g_all + (size_t)NG * MT * 4 * NBR * 32;                        // [NG][MT][NBR][32] cell state
    uint64_t* bars_all = reinterpret_cast<uint64_t*>(c_all + (size_t)NG * MT * NBR * 32);
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars_all + NG * NBAR);
    uint8_t* z_s = z_all + (size_t)grp * 2 * KB * ZBLK;
    float* c_s = c_all + (size_t)grp * MT * NBR * 32;
    uint64_t* bars = bars_all + grp * NBAR;
    uint64_t* x_full = bars;          // [2]  TMA -> MMA
    uint64_t* z_free = bars + 2;      // [2]  MMAs done with Z[buf] -> TMA
    uint64_t* h_ready = bars + 4;     // [2]  epilogue wrote h into Z[buf] -> MMA
    uint64_t* acc_free = bars + 6;    // [2]  epilogue done reading accumulator set -> MMA
    uint64_t* acc_full = bars + 8;    // [2][MT] MMA -> epilogue
    const int n0 = (blockIdx.x * NG + grp) * NBR;

    // zero Z (padding rows, h_{-1}) and the cell state before anything asynchronous starts
    for (int i = threadIdx.x; i < NG * 2 * KB * ZBLK / 16; i += blockDim.x) reinterpret_cast<uint4*>(z_all)[i] = make_uint4(0, 0, 0, 0);
    for (int i = threadIdx.x; i < NG * MT * NBR * 32; i += blockDim.x) c_all[i] = 0.0f;
    tc::fence_proxy_async_smem();
    if (threadIdx.x == 0) {
        for (int gi = 0; gi < NG; ++gi) {
            uint64_t* b = bars_all + gi * NBAR;
            for (int i = 0; i < 2; ++i) {
                tc::mbar_init(&b[i], 1);
                tc::mbar_init(&b[2 + i], MT);        // one commit per MMA-issuing warp
                tc::mbar_init(&b[4 + i], MT * 4);    // one arrival per epilogue warp
                tc::mbar_init(&b[6 + i], MT * 4);
            }
            for (int i = 0; i < 2 * MT; ++i) tc::mbar_init(&b[8 + i], 1);
        }
        tc::fence_barrier_init();
        tc::prefetch_tmap(&tma_x);
    }
    if (warp == 1) tc::tmem_alloc(tmem_holder, Cfg::TMEM_COLS);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;
    const uint32_t tmem_acc = tmem_base + (uint32_t)(grp * 2 * MT * UN);
    const uint32_t tmem_w = tmem_base + Cfg::ACC_COLS;
    if (grp == 0 && lw > MT) {
        // weights into tensor memory: epilogue warp (tile m, lane quarter q) of group 0, one weight row per thread
        const int m = (lw - (1 + MT)) >> 2, q = warp & 3;
        const uint4* src = reinterpret_cast<const uint4*>(p.w + (size_t)(m * 128 + q * 32 + lane) * 2 * C);
#pragma unroll 1
        for (int cb = 0; cb < C / 32; ++cb) {
            uint32_t r[32];
#pragma unroll
            for (int v = 0; v < 8; ++v) {
                const uint4 x = __ldg(src + cb * 8 + v);
                r[4 * v] = x.x; r[4 * v + 1] = x.y; r[4 * v + 2] = x.z; r[4 * v + 3] = x.w;
            }
            tc::tmem_st_32x32(tmem_w + ((uint32_t)(q * 32) << 16) + (uint32_t)(m * C + cb * 32), r);
        }
        tc::tmem_st_wait();
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();

    if (lw == 0) {
        // ---------------- TMA producer ----------------
        if (tc::elect_one()) {
            for (int s = 0; s < p.T; ++s) {
                const int t = p.reverse ? p.T - 1 - s : s;
                const int buf = s & 1;
                tc::mbar_wait(&z_free[buf], ((s >> 1) & 1) ^ 1);
                tc::mbar_arrive_expect_tx(&x_full[buf], (uint32_t)(KBX * NBR * KBLK * 2));
                for (int kb = 0; kb < KBX; ++kb) {
                    tc::tma_load_2d(z_s + (size_t)(buf * KB + kb) * ZBLK, &tma_x, &x_full[buf], kb * KBLK, t * p.N + n0);
                }
            }
        }
    } else if (lw <= MT) {
        // ---------------- MMA issuers: warp m issues the MMAs of gate tile m ----------------
        const int m = lw - 1;
        if (tc::elect_one()) {
            constexpr uint32_t idesc = tc::umma_idesc_f16(128, UN);
            const uint64_t zdesc0 = umma_desc_sw64(tc::smem_u32(z_s));
            for (int s = 0; s < p.T; ++s) {
                const int buf = s & 1;
                const uint32_t par = (uint32_t)((s >> 1) & 1);
                const uint64_t zd = zdesc0 + (uint64_t)((buf * KB * ZBLK) >> 4);
                const uint32_t d_tmem = tmem_acc + (uint32_t)((buf * MT + m) * UN);
                const bool dbg = p.dbg && blockIdx.x == 0 && warp == 1 && s >= 64 && s < 68;
                tc::mbar_wait(&x_full[buf], par);
                tc::mbar_wait(&acc_free[buf], par ^ 1);
                tc::tc_fence_after();
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    if (half == 1) {
                        if (dbg) p.dbg[(s - 64) * 16 + 0] = clock64();
                        tc::mbar_wait(&h_ready[buf], par);
                        tc::tc_fence_after();
                        if (dbg) p.dbg[(s - 64) * 16 + 1] = clock64();
                    }
#pragma unroll
                    for (int kb = half * KBX; kb < (half + 1) * KBX; ++kb) {
                        const uint64_t bdesc = zd + (uint64_t)((kb * ZBLK) >> 4);
                        const uint32_t a_t = tmem_w + (uint32_t)(m * C + kb * 16);  // 16 columns per 32-element block
                        tc::umma_f16_ts(d_tmem, a_t, bdesc, idesc, kb != 0);
                        tc::umma_f16_ts(d_tmem, a_t + 8, bdesc + 2, idesc, true);
                    }
                }
                tc::umma_commit(&acc_full[buf * MT + m]);
                tc::umma_commit(&z_free[buf]);
                if (dbg) p.dbg[(s - 64) * 16 + 2] = clock64();
            }
        }
    } else {
        // ---------------- epilogue quartets: gates, cell update, h_t ----------------
        const int ewarp = lw - (1 + MT);
        const int m = ewarp >> 2;    // gate tile
        const int ew = ewarp & 3;    // warp within the quartet -> chunks ew, ew + 4, ...
        const int gate = warp & 3;   // TMEM lane quarter this warp may read == gate type (i,f,g,o)
        const int bar_id = 1 + grp * MT + m;
        float* gs = g_all + ((size_t)grp * MT + m) * 4 * NBR * 32;
        constexpr int CPT = (NBR + 3) / 4;
        const float b = __ldg(p.bias + m * 128 + gate * 32 + lane);
        const float am = gate == 2 ? 2.0f : 1.0f;  // tanh(v) = 1 - 2/(e^{2v}+1), sigmoid(v) = 1 - 1/(e^{v}+1)
        if (lane == 0) tc::mbar_arrive(&h_ready[0]);  // h_{-1} = 0 is in place (zeroed before the CTA-wide sync)

        for (int s = 0; s < p.T; ++s) {
            const int t = p.reverse ? p.T - 1 - s : s;
            const int buf = s & 1, nbuf = buf ^ 1;
            const uint32_t par = (uint32_t)((s >> 1) & 1);
            uint8_t* zh_next = z_s + (size_t)(nbuf * KB + KBX) * ZBLK;
            __half* y_t = p.seq + ((size_t)t * p.N + n0) * C;
            const bool dbg = p.dbg && blockIdx.x == 0 && grp == 0 && ewarp == 0 && lane == 0 && s >= 64 && s < 68;
            if (dbg) p.dbg[(s - 64) * 16 + 4] = clock64();
            tc::mbar_wait(&acc_full[buf * MT + m], par);
            tc::tc_fence_after();
            if (dbg) p.dbg[(s - 64) * 16 + 5] = clock64();
            uint32_t r[NBR];
            const uint32_t taddr = tmem_acc + ((uint32_t)(gate * 32) << 16) + (uint32_t)((buf * MT + m) * UN);
            if constexpr (NBR == 16) {
                tc::tmem_ld_32x16(taddr, r);
            } else if constexpr (NBR == 8) {
                tc::tmem_ld_32x8(taddr, r);
            } else if constexpr (NBR == 4) {
                tc::tmem_ld_32x4(taddr, r);
            } else {
                tc::tmem_ld_32x2(taddr, r);
            }
            tc::tmem_ld_wait();
            if (dbg) p.dbg[(s - 64) * 16 + 6] = clock64();
#pragma unroll
            for (int n = 0; n < NBR; ++n) gs[(gate * NBR + n) * 32 + lane] = gate_act(__uint_as_float(r[n]) + b, am);
            float c_old[CPT];
#pragma unroll
            for (int j = 0; j < CPT; ++j) {
                const int n = ew + 4 * j;
                c_old[j] = n < NBR ? c_s[((size_t)m * NBR + n) * 32 + lane] : 0.0f;
            }
            if (dbg) p.dbg[(s - 64) * 16 + 7] = clock64();
            named_bar_sync(bar_id, 128);
            if (dbg) p.dbg[(s - 64) * 16 + 8] = clock64();
#pragma unroll
            for (int j = 0; j < CPT; ++j) {
                const int n = ew + 4 * j;
                if (n < NBR) {
                    const float ig = gs[(0 * NBR + n) * 32 + lane];
                    const float fg = gs[(1 * NBR + n) * 32 + lane];
                    const float gg = gs[(2 * NBR + n) * 32 + lane];
                    const float og = gs[(3 * NBR + n) * 32 + lane];
                    const float cs = fg * c_old[j] + ig * gg;
                    c_s[((size_t)m * NBR + n) * 32 + lane] = cs;  // owned by this thread for all steps
                    const __half h = __float2half_rn(og * tanh_f(cs));
                    *reinterpret_cast<__half*>(zh_next + (size_t)m * ZBLK + sw64_offset(n, lane)) = h;
                    y_t[(size_t)n * C + m * 32 + lane] = h;
                }
            }
            if (dbg) p.dbg[(s - 64) * 16 + 9] = clock64();
            tc::tc_fence_before();
            tc::fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
                tc::mbar_arrive(&acc_free[buf]);
                tc::mbar_arrive(&h_ready[nbuf]);
            }
            if (dbg) p.dbg[(s - 64) * 16 + 10] = clock64();
            named_bar_sync(bar_id, 128);  // gs reuse across steps
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 1) tc::tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

// ------------------------------------------------------------------------------------------------
// LSTM layer for hidden sizes whose weights do not fit one SM (hac: C = 384): hoisted x-projection + cluster recurrence.
//
// The x_t half of the gate pre-activations does not depend on the recurrence, so it is hoisted into one large
// tcgen05 GEMM per layer (gemm.cu) that writes gx[t][chunk block of 32][gate row][chunk] (fp16, bias included).  Gate rows
// are permuted so that a warp's 32 TMEM lanes hold (8 units x 4 gates).
//
// Recurrence: W_hh (4C x C fp16 = 1.18 MB for C = 384) is split by gate tile across the CL = 6 CTAs of a thread-block
// cluster and stays in their TENSOR MEMORY for the whole sequence (A operand of tcgen05.mma read from TMEM, written once with
// tcgen05.st).  Every step each CTA computes the gates of its own 32 * TPC hidden units from the full h_{t-1} (its private
// copy in shared memory) and all-gathers its slice of h_t into every CTA's copy.  The kernel is built around what bounds a
// step -- the latency of that exchange and of the gate math -- not around barriers:
//   * the cluster's chunks form NG = 2 independent groups of GN = 16 or 32; the MMA issuer ping-pongs between them, so group
//     A's gate math and all-gather run under group B's MMAs (a single group has nothing to overlap with);
//   * h_t travels as BULK ASYNC COPIES (cp.async.bulk shared::cta -> shared::cluster): an epilogue tile stages its
//     32 units x GN chunks in the swizzled operand layout in local shared memory and one thread sends the block to
//     every CTA of the cluster; the bytes complete on the destination's mbarrier, which is what its MMA issuer waits on.
//     No cluster barrier, no remote scalar stores, and the operand is written and read by the async proxy (no
//     generic->async fence on the receiving side).  Measured alternatives, both slower (3.3 -> 5.7 ms per layer,
//     profiles/r02_b3_*, r02_b5_*): the exchange through L2 (TMA store + multicast TMA load) and remote 16-byte stores with
//     remote mbarrier arrivals;
//   * the same staged block leaves for HBM as one TMA store (no per-thread global stores);
//   * the four gates of a cell meet by quad transposes in registers (shuffles) instead of a shared-memory exchange.
// A step's cost hardly depends on GN (it is latency, ~12 small copies per CTA and group), so GN = 32 -- 64 chunks per
// cluster, 8 clusters = 48 SMs for a batch of 512 -- does the same work in fewer SM-cycles and leaves room for the
// recurrences of two more batches next to it; GN = 16 is the lower-latency shape for a lone runner.
// Safety of buffer reuse without extra barriers (Z and staging are double-buffered): a CTA starts step s+2 for a group
// only after every CTA's slice of h_{s+1} has landed, which those CTAs sent only after their step-(s+1) MMAs completed,
// which needed every slice of h_s to have landed everywhere -- so by then nobody reads Z[(s) & 1] of step s or the
// staging block sent at step s any more.
// ------------------------------------------------------------------------------------------------
struct LstmRecParams {
    __half* seq;          // [T][N][C] output h (in place over the layer input)
    const __half* gx;     // [T][N / 32][4C][32]
    int T, N, reverse;
    const int32_t* lens;  // optional per-chunk length in samples (variable chunk sizes); stride = samples per step
    int stride;
    long long* dbg;       // optional clock64 timeline of CTA 0, steps 64..67 (B200_DEBUG_LSTM_TIMELINE); nullptr in production
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t mapa_shared(uint32_t local_addr, uint32_t cta_rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(cta_rank));
    return r;
}
__device__ __forceinline__ void cluster_arrive_release() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait_acquire() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }

template <int C, int CL, int NG, int GN>
struct ClusterCfg {
    static constexpr int MT = C / 32;
    static constexpr int TPC = MT / CL;            // gate tiles per CTA
    static constexpr int KBH = C / KBLK;           // operand blocks (32 hidden units each)
    static constexpr int UNC = NG * GN;            // chunks per cluster
    static constexpr int NQ = GN / 4;              // chunk quads per group: a lane owns the cells (unit, chunk 4c + gate lane)
    static constexpr int EW = 4 * TPC * NG;        // epilogue warps: (group, tile, TMEM lane quarter)
    static constexpr int THREADS = 64 + 32 * EW;
    static constexpr int ZB = GN * KBLK * 2;       // bytes of one operand block (GN chunk rows x 32 fp16)
    static constexpr int WCOLS = C / 2;            // TMEM columns of one weight tile
    static constexpr int ACC0 = TPC * WCOLS;       // accumulators start behind the weights
    static constexpr int NEED = ACC0 + NG * TPC * GN;
    static constexpr uint32_t TMEM_COLS = NEED <= 64 ? 64 : NEED <= 128 ? 128 : NEED <= 256 ? 256 : 512;
    static constexpr size_t Z_BYTES = (size_t)NG * 2 * KBH * ZB;      // [group][buffer][block]
    static constexpr size_t ST_BYTES = (size_t)NG * 2 * TPC * ZB;     // [group][buffer][tile] staging
    static constexpr size_t SMEM = 1024 + Z_BYTES + ST_BYTES + 1024;
    static_assert(MT % CL == 0 && NEED <= 512, "tile split / tensor memory budget");
    static_assert(GN == 16 || GN == 32, "a group is 16 or 32 chunks (UMMA N, tcgen05.ld shape, gx block)");
    static_assert(UNC <= 64, "chunk-length table");
};

template <int C, int CL, int NG, int GN>
__global__ void __launch_bounds__(ClusterCfg<C, CL, NG, GN>::THREADS, 1) lstm_cluster_kernel(const __grid_constant__ CUtensorMap tma_y,
                                                                                            const __half* __restrict__ w_hh,
                                                                                            const LstmRecParams p) {
    using Cfg = ClusterCfg<C, CL, NG, GN>;
    constexpr int TPC = Cfg::TPC, KBH = Cfg::KBH, ZB = Cfg::ZB, NQ = Cfg::NQ;
    constexpr int GB = 32;  // chunk block of the gx layout
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* z_s = smem;                                   // [NG][2][KBH][ZB]  this CTA's copy of the full h
    uint8_t* st_s = z_s + Cfg::Z_BYTES;                    // [NG][2][TPC][ZB]  h_t of this CTA's units, staged for sending
    uint64_t* bars = reinterpret_cast<uint64_t*>(st_s + Cfg::ST_BYTES);
    uint64_t* h_full = bars;                               // [NG][2]   bytes of h landed -> MMA issuer
    uint64_t* acc_full = bars + NG * 2;                    // [NG][TPC] MMA -> epilogue
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(acc_full + NG * TPC);
    int* len_s = reinterpret_cast<int*>(tmem_holder + 2);   // [UNC] steps of every chunk of the cluster (variable chunk sizes)

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int cluster_id = blockIdx.x / CL;
    const int n0 = cluster_id * Cfg::UNC;
    if (threadIdx.x < Cfg::UNC) {
        len_s[threadIdx.x] = p.lens ? min(p.T, __ldg(p.lens + n0 + threadIdx.x) / p.stride) : p.T;
    }

    for (int i = threadIdx.x; i < (int)((Cfg::Z_BYTES + Cfg::ST_BYTES) / 16); i += blockDim.x) {
        reinterpret_cast<uint4*>(z_s)[i] = make_uint4(0, 0, 0, 0);
    }
    tc::fence_proxy_async_smem();
    if (threadIdx.x == 0) {
        for (int i = 0; i < NG * 2; ++i) tc::mbar_init(&h_full[i], 1);   // one armed arrival + the bytes of all slices
        for (int i = 0; i < NG * TPC; ++i) tc::mbar_init(&acc_full[i], 1);
        tc::fence_barrier_init();
        tc::prefetch_tmap(&tma_y);
    }
    if (warp == 1) tc::tmem_alloc(tmem_holder, Cfg::TMEM_COLS);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;

    const bool is_epi = warp >= 2;
    const int ew = warp - 2;
    const int eg = ew / (4 * TPC);        // group served by this epilogue warp
    const int ti = (ew >> 2) % TPC;       // tile of this CTA
    const int qt = warp & 3;              // TMEM lane quarter (the 4 warps of a (group, tile) cover all four)
    if (is_epi && eg == 0) {
        // this thread's weight row (tile rank*TPC + ti, row 32*qt + lane) -> tensor memory, once
        const int m = (int)rank * TPC + ti;
        const uint4* src = reinterpret_cast<const uint4*>(w_hh + (size_t)(m * 128 + qt * 32 + lane) * C);
#pragma unroll 1
        for (int cb = 0; cb < C / 64; ++cb) {
            uint32_t r[32];
#pragma unroll
            for (int v = 0; v < 8; ++v) {
                const uint4 x = __ldg(src + cb * 8 + v);
                r[4 * v] = x.x; r[4 * v + 1] = x.y; r[4 * v + 2] = x.z; r[4 * v + 3] = x.w;
            }
            tc::tmem_st_32x32(tmem_base + ((uint32_t)(qt * 32) << 16) + (uint32_t)(ti * Cfg::WCOLS + cb * 32), r);
        }
        tc::tmem_st_wait();
    }
    if (threadIdx.x == 0) {
        // h_{-1} = 0 is in place: complete phase 0 of the buffer-0 barriers without bytes
        for (int g = 0; g < NG; ++g) tc::mbar_arrive(&h_full[g * 2 + 0]);
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    // every CTA of the cluster has zeroed its operand buffers and initialised its barriers before any remote copy lands
    cluster_arrive_release();
    cluster_wait_acquire();

    const uint32_t z_local = tc::smem_u32(z_s);
    // A chunk shorter than the batch's chunk size is the sequence 0 .. len-1: outside it the state is zero (masked in the
    // epilogue), so the cluster only walks the steps some chunk of it is alive in -- the last ones backwards, the first
    // ones forwards -- and short chunks grouped in a cluster cost proportionally less.
    int steps = 0;
#pragma unroll 4
    for (int i = 0; i < Cfg::UNC; ++i) steps = max(steps, len_s[i]);
    if (warp == 1) {
        // ---------------- MMA issuer ----------------
        if (tc::elect_one()) {
            const uint64_t zdesc0 = umma_desc_sw64(z_local);
            constexpr uint32_t idesc = tc::umma_idesc_f16(128, GN);
            for (int s = 0; s < steps; ++s) {
                const int buf = s & 1, nbuf = buf ^ 1;
                const uint32_t par = (uint32_t)((s >> 1) & 1);
#pragma unroll
                for (int g = 0; g < NG; ++g) {
                    // arm the barrier the slices of h_s will complete on (all CL CTAs x TPC tiles, this group)
                    if (s + 1 < steps) tc::mbar_arrive_expect_tx(&h_full[g * 2 + nbuf], (uint32_t)(CL * TPC * ZB));
                    long long* d = (p.dbg && blockIdx.x == 0 && s >= 64 && s < 68) ? p.dbg + (s - 64) * 32 + g * 4 : nullptr;
                    if (d) d[0] = clock64();
                    tc::mbar_wait(&h_full[g * 2 + buf], par);
                    tc::tc_fence_after();
                    if (d) d[1] = clock64();
                    const uint64_t zd = zdesc0 + (uint64_t)(((g * 2 + buf) * KBH * ZB) >> 4);
#pragma unroll
                    for (int i = 0; i < TPC; ++i) {
                        const uint32_t d_tmem = tmem_base + (uint32_t)(Cfg::ACC0 + (g * TPC + i) * GN);
#pragma unroll
                        for (int kb = 0; kb < KBH; ++kb) {
                            const uint32_t a_t = tmem_base + (uint32_t)(i * Cfg::WCOLS + kb * 16);
                            const uint64_t bdesc = zd + (uint64_t)((kb * ZB) >> 4);
                            tc::umma_f16_ts(d_tmem, a_t, bdesc, idesc, kb != 0);
                            tc::umma_f16_ts(d_tmem, a_t + 8, bdesc + 2, idesc, true);
                        }
                        tc::umma_commit(&acc_full[g * TPC + i]);
                    }
                    if (d) d[2] = clock64();
                }
            }
        }
        __syncwarp();
    } else if (is_epi) {
        // ---------------- epilogue: gates, cell update, h_t out ----------------
        const int uk = lane >> 2, gj = lane & 3;   // unit within the warp's 8, gate type (i, f, g, o)
        const int g0 = gj & 1, g1 = gj >> 1;
        const int m = (int)rank * TPC + ti;        // global tile = operand block this warp's units belong to
        const float am = gj == 2 ? 2.0f : 1.0f;    // tanh(v) = 1 - 2/(e^{2v}+1), sigmoid(v) = 1 - 1/(e^{v}+1)
        const bool sender = qt == 0 && lane == 0;
        const int bar_id = 1 + eg * TPC + ti;
        float c_reg[NQ];                           // cells (unit uk, chunk 4c + gj)
#pragma unroll
        for (int c = 0; c < NQ; ++c) c_reg[c] = 0.0f;
        const uint32_t taddr = tmem_base + ((uint32_t)(qt * 32) << 16) + (uint32_t)(Cfg::ACC0 + (eg * TPC + ti) * GN);
        const int nc0 = n0 + eg * GN;              // first chunk of this group
        const size_t gx_step = (size_t)(p.N / GB) * (size_t)(4 * C) * GB;
        const __half* gx_base = p.gx + (size_t)(nc0 / GB) * (size_t)(4 * C) * GB + (size_t)(m * 128 + qt * 32 + lane) * GB + (size_t)(nc0 % GB);
        // remote addresses of this tile's operand block and of the barrier it completes on, per buffer
        uint32_t dst_z[2], dst_bar[2];
#pragma unroll
        for (int b = 0; b < 2; ++b) {
            dst_z[b] = z_local + (uint32_t)(((eg * 2 + b) * KBH + m) * ZB);
            dst_bar[b] = tc::smem_u32(&h_full[eg * 2 + b]);
        }
        int my_len[NQ];   // steps of the chunks this lane owns cells of
#pragma unroll
        for (int c = 0; c < NQ; ++c) my_len[c] = len_s[eg * GN + 4 * c + gj];
        for (int s = 0; s < steps; ++s) {
            const int t = p.reverse ? steps - 1 - s : s;
            const int nbuf = (s & 1) ^ 1;
            const uint4* gp = reinterpret_cast<const uint4*>(gx_base + (size_t)t * gx_step);
            uint4 gxv[GN / 8];
#pragma unroll
            for (int e = 0; e < GN / 8; ++e) gxv[e] = __ldg(gp + e);
            uint8_t* stage = st_s + (size_t)((eg * 2 + (s & 1)) * TPC + ti) * ZB;
            long long* d = (p.dbg && blockIdx.x == 0 && ew == 0 && lane == 0 && s >= 64 && s < 68) ? p.dbg + (s - 64) * 32 + 8 : nullptr;
            if (d) d[0] = clock64();
            tc::mbar_wait(&acc_full[eg * TPC + ti], (uint32_t)(s & 1));
            tc::tc_fence_after();
            if (d) d[1] = clock64();
            uint32_t r[GN];
            if constexpr (GN == 32) {
                tc::tmem_ld_32x32(taddr, r);
            } else {
                tc::tmem_ld_32x16(taddr, r);
            }
            tc::tmem_ld_wait();
            tc::tc_fence_before();
            if (d) d[2] = clock64();
            float a[GN];
#pragma unroll
            for (int e = 0; e < GN / 8; ++e) {
                const __half2* hx = reinterpret_cast<const __half2*>(&gxv[e]);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float2 f = __half22float2(hx[j]);
                    a[8 * e + 2 * j] = gate_act(__uint_as_float(r[8 * e + 2 * j]) + f.x, am);
                    a[8 * e + 2 * j + 1] = gate_act(__uint_as_float(r[8 * e + 2 * j + 1]) + f.y, am);
                }
            }
            if (d) d[3] = clock64();
            // 4x4 transposes inside the quad: afterwards a[4c + k] = gate k of (unit uk, chunk 4c + gj)
#pragma unroll
            for (int c = 0; c < NQ; ++c) {
                float s0 = g0 ? a[4 * c + 0] : a[4 * c + 1];
                float s1 = g0 ? a[4 * c + 2] : a[4 * c + 3];
                float r0 = __shfl_xor_sync(0xffffffffu, s0, 1), r1 = __shfl_xor_sync(0xffffffffu, s1, 1);
                if (g0) { a[4 * c + 0] = r0; a[4 * c + 2] = r1; } else { a[4 * c + 1] = r0; a[4 * c + 3] = r1; }
                s0 = g1 ? a[4 * c + 0] : a[4 * c + 2];
                s1 = g1 ? a[4 * c + 1] : a[4 * c + 3];
                r0 = __shfl_xor_sync(0xffffffffu, s0, 2);
                r1 = __shfl_xor_sync(0xffffffffu, s1, 2);
                if (g1) { a[4 * c + 0] = r0; a[4 * c + 1] = r1; } else { a[4 * c + 2] = r0; a[4 * c + 3] = r1; }
            }
            if (d) d[4] = clock64();
            float hv[NQ];
#pragma unroll
            for (int c = 0; c < NQ; ++c) {
                // outside the chunk (variable chunk sizes) the state is held at zero: a multiplicative mask, not a branch, so
                // that the cells' transcendentals stay interleaved
                const float alive = t < my_len[c] ? 1.0f : 0.0f;
                const float cs = (a[4 * c + 1] * c_reg[c] + a[4 * c + 0] * a[4 * c + 2]) * alive;
                c_reg[c] = cs;
                hv[c] = a[4 * c + 3] * tanh_f(cs) * alive;
            }
            // Pair the units (uk, uk ^ 1): of the NQ / 2 chunk-quad pairs the even unit's lane stores the first half for both
            // units and the odd unit's lane the second half, as 32-bit words (two adjacent units of one chunk row).
            {
                const bool odd = uk & 1;
                const int ucol = qt * 8 + (uk & ~1);
                constexpr int NP = NQ / 2;        // half2 pairs of chunk quads per lane: (hv[2i], hv[2i+1])
#pragma unroll
                for (int i = 0; i < NP / 2; ++i) {
                    const __half2 first = __floats2half2_rn(hv[2 * i], hv[2 * i + 1]);
                    const __half2 second = __floats2half2_rn(hv[2 * (NP / 2 + i)], hv[2 * (NP / 2 + i) + 1]);
                    const uint32_t fu = *reinterpret_cast<const uint32_t*>(&first), su = *reinterpret_cast<const uint32_t*>(&second);
                    const uint32_t got = __shfl_xor_sync(0xffffffffu, odd ? fu : su, 4);
                    const uint32_t keep = odd ? su : fu;
                    // keep = my h for quads (q0, q0+1), got = the partner unit's h for the same quads
                    const uint32_t even_unit = odd ? got : keep, odd_unit = odd ? keep : got;
                    const int q0 = 2 * (odd ? NP / 2 + i : i);
                    const uint32_t w0 = (even_unit & 0xffffu) | (odd_unit << 16);          // chunk 4*q0 + gj
                    const uint32_t w1 = (even_unit >> 16) | (odd_unit & 0xffff0000u);      // chunk 4*(q0+1) + gj
                    *reinterpret_cast<uint32_t*>(stage + sw64_offset(4 * q0 + gj, ucol)) = w0;
                    *reinterpret_cast<uint32_t*>(stage + sw64_offset(4 * (q0 + 1) + gj, ucol)) = w1;
                }
            }
            if (d) d[5] = clock64();
            tc::fence_proxy_async_smem();            // staged block -> visible to the bulk copy / TMA store
            if (d) d[6] = clock64();
            if (sender) tc::bulk_wait_group_read<0>();  // the store issued a step ago has read the other staging buffer
            if (d) d[7] = clock64();
            named_bar_sync(bar_id, 128);
            if (d) d[8] = clock64();
            if (sender) {
                // all-gather over distributed shared memory: one bulk copy per destination CTA
                if (s + 1 < steps) {
#pragma unroll
                    for (int rr = 0; rr < CL; ++rr) {
                        tc::bulk_copy_smem_to_cluster(mapa_shared(dst_z[nbuf], (uint32_t)rr), tc::smem_u32(stage), (uint32_t)ZB,
                                                      mapa_shared(dst_bar[nbuf], (uint32_t)rr));
                    }
                }
                tc::tma_store_2d(&tma_y, stage, m * 32, t * p.N + nc0);
                tc::bulk_commit_group();
            }
            if (d) d[9] = clock64();
        }
        if (sender) tc::bulk_wait_group<0>();
    }
    tc::tc_fence_before();
    __syncthreads();
    // nobody leaves while a peer may still address this CTA's shared memory
    cluster_arrive_release();
    cluster_wait_acquire();
    if (warp == 1) tc::tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
struct LstmLayerWeights {
    __half* w = nullptr;   // resident path: [4C][2C] permuted rows, [W_ih | W_hh]
    float* bias = nullptr; // [4C] permuted
    // hoisted path (lstm_cluster_kernel): rows permuted as (tile, quarter, unit-in-quarter, gate)
    __half* w_ih = nullptr;  // [4C][C]
    __half* w_hh = nullptr;  // [4C][C]
};

class LstmModel;

class LstmPlan final : public ForwardPlan {
public:
    void run(cudaStream_t stream, ProfileSink* prof) override;
    int launches() const override { return 1 + 1 + num_layers * (hoisted ? 2 : 1) + num_linear; }
    std::string info() const override {
        if (hoisted) return "lstm_rec.ctas=" + std::to_string(lstm_grid) + ";lstm_rec.chunks_per_cluster=" + std::to_string(rec_un);
        return "lstm_layer.ctas=" + std::to_string(lstm_grid) + ";lstm_layer.groups=" + std::to_string(lstm_ng) +
               ";lstm_layer.chunks_per_group=" + std::to_string(lstm_nbr);
    }
    void set_chunk_lengths(const int32_t* d_lens) override {
        if (!hoisted) return;  // the mode exists for the cluster recurrence only
        conv12.lens = d_lens;
        for (auto& rp : rec_p) rp.lens = d_lens;
    }

    Conv12Params conv12{};
    dim3 conv12_grid;
    bool conv12_tc = false;  // conv2 on tcgen05 (conv12_tc_kernel)
    int conv12_tiles_per_chunk = 0, conv12_tc_grid = 0;
    GemmPlan conv3;
    std::vector<CUtensorMap> lstm_x;
    std::vector<LstmParams> lstm_p;
    int lstm_grid = 0, lstm_nbr = 16, lstm_ng = 2;
    size_t lstm_smem = 0;
    void launch_lstm(int l, cudaStream_t stream) const;
    // hoisted path
    long long* dbg_timeline = nullptr;
    int debug_layers = -1;  // >= 0: stop after that many LSTM layers (B200_DEBUG_LSTM_LAYERS, read once at plan creation)
    bool hoisted = false;
    int rec_un = 32;       // chunks per cluster: 16 (one group of 16), 32 (two groups of 16) or 64 (two groups of 32)
    std::vector<GemmPlan> gx_gemm;
    std::vector<const __half*> rec_whh;
    std::vector<LstmRecParams> rec_p;
    CUtensorMap rec_y_map; // seq as [T_out * Np][C], box 32 units x (chunks of a group): the TMA store of h_t
    void launch_rec(int l, cudaStream_t stream) const;
    GemmPlan linear1, linear2;
    int num_layers = 0, num_linear = 1;
    const LstmModel* model = nullptr;
};

class LstmModel final : public Model {
public:
    LstmModel(const b200_model_desc& d, const b200_tensor* tensors, int n);
    ~LstmModel() override;
    size_t workspace_bytes(int N, int T_in) const override;
    std::unique_ptr<ForwardPlan> make_plan(int N, int T_in, const __half* signal, __half* scores, void* ws,
                                           size_t ws_bytes) override;
    bool variable_chunk_sizes() const override {
        // FLSTM models never run variable chunk sizes (api/runner_creation.cpp:28)
        return hoisted() && desc.lstm_inner_dim == 0;
    }

    b200_model_desc desc;
    float* conv_w = nullptr;  // packed conv1 / conv2 weights (see Conv12Params::w)
    __half* w3 = nullptr;               // [C][K3p]
    float* b3 = nullptr;
    int K3 = 0, K3p = 0;
    std::vector<LstmLayerWeights> layers;
    __half* wl1 = nullptr;  // [out1][Cp]
    float* bl1 = nullptr;
    __half* wl2 = nullptr;  // decomposition: [outsize][out_features_p]
    int Cp = 0, out1 = 0, out1p = 0;

private:
    int pad3() const { return desc.convs[2].winlen / 2; }
    int t_pad(int T_in) const { return T_in + 2 * pad3() + 8; }
    bool hoisted() const { return (size_t)4 * desc.lstm_size * 2 * desc.lstm_size * 2 > 150 * 1024; }
    int n_pad(int N) const { return hoisted() ? (N + 31) / 32 * 32 : (N + 15) / 16 * 16; }
};

LstmModel::LstmModel(const b200_model_desc& d, const b200_tensor* tensors, int n) : desc(d) {
    if (d.num_convs != 3) throw std::invalid_argument("Expected 3 convolution layers but found: " + std::to_string(d.num_convs));
    const auto &c1 = d.convs[0], &c2 = d.convs[1], &c3 = d.convs[2];
    if (c1.insize != 1 || c1.stride != 1 || c2.stride != 1 || c2.size != 16 || c1.size > 16 || c1.winlen > MAXW ||
        c2.winlen > MAXW || c2.insize != c1.size || c3.insize != 16) {
        throw Unsupported("conv stack shape outside what conv12_kernel implements");
    }
    const int C = d.lstm_size;
    if (C != c3.size) throw std::invalid_argument("last convolution size != lstm_size");
    if (C != 96 && C != 192 && C != 384) {
        // kernels are instantiated for the sizes of the reference's model zoo this engine covers
        throw Unsupported("lstm_size " + std::to_string(C) + " is not supported (96, 192 and 384 are)");
    }

    if (d.lstm_layers < 1 || d.lstm_layers > 8) throw std::invalid_argument("bad lstm_layers");

    // conv1: torch [c1][1][w] -> [c1][w]; conv2: torch [co][ci][k] -> [k][ci][co]
    {
        const auto& tw = find_tensor(tensors, n, "0.conv.weight.tensor");
        const auto& tb = find_tensor(tensors, n, "0.conv.bias.tensor");
        const auto& tw2 = find_tensor(tensors, n, "1.conv.weight.tensor");
        const auto& tb2 = find_tensor(tensors, n, "1.conv.bias.tensor");
        std::vector<float> pk(CONV_W_FLOATS, 0.0f);
        float* w1 = pk.data();
        float* b1 = w1 + 16 * MAXW;
        float* w2 = b1 + 16;
        float* b2 = w2 + MAXW * 16 * 16;
        std::memcpy(w1, tw.data, sizeof(float) * (size_t)c1.size * c1.winlen);
        std::memcpy(b1, tb.data, sizeof(float) * c1.size);
        for (int co = 0; co < 16; ++co)
            for (int ci = 0; ci < c2.insize; ++ci)
                for (int k = 0; k < c2.winlen; ++k)
                    w2[((size_t)k * 16 + ci) * 16 + co] = tw2.data[((size_t)co * c2.insize + ci) * c2.winlen + k];
        std::memcpy(b2, tb2.data, sizeof(float) * 16);
        conv_w = upload_f32(pk);
    }
    // conv3 as GEMM weights: [C][k*16 + ci], K padded to a multiple of 64
    {
        const auto& tw = find_tensor(tensors, n, "2.conv.weight.tensor");
        const auto& tb = find_tensor(tensors, n, "2.conv.bias.tensor");
        K3 = c3.winlen * 16;
        K3p = (K3 + 63) / 64 * 64;
        std::vector<float> w((size_t)C * K3p, 0.0f);
        for (int co = 0; co < C; ++co)
            for (int ci = 0; ci < 16; ++ci)
                for (int k = 0; k < c3.winlen; ++k)
                    w[(size_t)co * K3p + k * 16 + ci] = tw.data[((size_t)co * 16 + ci) * c3.winlen + k];
        w3 = upload_f16(w);
        b3 = upload_f32(std::vector<float>(tb.data, tb.data + C));
    }
    // LSTM layers: permute gate rows into tiles of (i|f|g|o) x 32 units, concatenate [W_ih | W_hh]
    for (int l = 0; l < d.lstm_layers; ++l) {
        const std::string pfx = std::to_string(d.num_convs + l + 1) + ".rnn.";
        // plain LSTM: the tensors as they are.  FLSTM (nn/FLSTMStack.cpp:18-25,108-124): gates = up_ih (dn_ih x_t) +
        // up_hh (dn_hh h_{t-1}) + bias, folded once into W = up x dn (double accumulation, rounded to fp32) so that the
        // recurrence kernels -- and their weights-in-tensor-memory residency -- serve both model kinds.
        struct View { const float* data; };
        std::vector<float> folded_ih, folded_hh;
        View wih{}, whh{}, bih{}, bhh{};
        if (d.lstm_inner_dim > 0) {
            const int K = d.lstm_inner_dim;
            auto fold = [&](const char* up_name, const char* dn_name, std::vector<float>& out) {
                const auto& up = find_tensor(tensors, n, pfx + up_name);
                const auto& dn = find_tensor(tensors, n, pfx + dn_name);
                if (up.ndim != 2 || dn.ndim != 2 || up.dims[0] != 4 * C || up.dims[1] != K || dn.dims[0] != K || dn.dims[1] != C) {
                    throw std::invalid_argument("FLSTM tensor " + pfx + up_name + " / " + dn_name + " has the wrong shape");
                }
                out.assign((size_t)4 * C * C, 0.0f);
                std::vector<double> acc((size_t)C);
                for (int r = 0; r < 4 * C; ++r) {
                    std::fill(acc.begin(), acc.end(), 0.0);
                    for (int k = 0; k < K; ++k) {
                        const double u = up.data[(size_t)r * K + k];
                        const float* drow = dn.data + (size_t)k * C;
                        for (int c2 = 0; c2 < C; ++c2) acc[c2] += u * (double)drow[c2];
                    }
                    for (int c2 = 0; c2 < C; ++c2) out[(size_t)r * C + c2] = (float)acc[c2];
                }
            };
            fold("up_weight_ih.tensor", "dn_weight_ih.tensor", folded_ih);
            fold("up_weight_hh.tensor", "dn_weight_hh.tensor", folded_hh);
            wih.data = folded_ih.data();
            whh.data = folded_hh.data();
            bih.data = find_tensor(tensors, n, pfx + "up_bias_ih.tensor").data;
            bhh.data = find_tensor(tensors, n, pfx + "up_bias_hh.tensor").data;
        } else {
            wih.data = find_tensor(tensors, n, pfx + "weight_ih_l0.tensor").data;
            whh.data = find_tensor(tensors, n, pfx + "weight_hh_l0.tensor").data;
            bih.data = find_tensor(tensors, n, pfx + "bias_ih_l0.tensor").data;
            bhh.data = find_tensor(tensors, n, pfx + "bias_hh_l0.tensor").data;
        }
        std::vector<float> w((size_t)4 * C * 2 * C), b((size_t)4 * C);
        for (int m = 0; m < C / 32; ++m)
            for (int g = 0; g < 4; ++g)
                for (int u = 0; u < 32; ++u) {
                    const int dst = m * 128 + g * 32 + u;
                    const int src = g * C + m * 32 + u;
                    std::memcpy(&w[(size_t)dst * 2 * C], &wih.data[(size_t)src * C], sizeof(float) * C);
                    std::memcpy(&w[(size_t)dst * 2 * C + C], &whh.data[(size_t)src * C], sizeof(float) * C);
                    b[dst] = bih.data[src] + bhh.data[src];
                }
        LstmLayerWeights lw;
        if ((size_t)4 * C * 2 * C * 2 <= 150 * 1024) {
            lw.w = upload_f16(w);
            lw.bias = upload_f32(b);
        } else {
            std::vector<float> wi((size_t)4 * C * C), wh((size_t)4 * C * C), b2((size_t)4 * C);
            for (int m = 0; m < C / 32; ++m)
                for (int r = 0; r < 128; ++r) {
                    const int g = r & 3, unit = 32 * m + 8 * (r >> 5) + ((r & 31) >> 2);
                    const int dst = m * 128 + r, src = g * C + unit;
                    std::memcpy(&wi[(size_t)dst * C], &wih.data[(size_t)src * C], sizeof(float) * C);
                    std::memcpy(&wh[(size_t)dst * C], &whh.data[(size_t)src * C], sizeof(float) * C);
                    b2[dst] = bih.data[src] + bhh.data[src];
                }
            lw.w_ih = upload_f16(wi);
            lw.w_hh = upload_f16(wh);
            lw.bias = upload_f32(b2);
        }
        layers.push_back(lw);
    }
    // linear(s)
    {
        const int layer = d.num_convs + d.lstm_layers + 1;
        const auto& tw = find_tensor(tensors, n, std::to_string(layer) + ".linear.weight.tensor");
        out1 = d.out_features > 0 ? d.out_features : d.outsize;
        Cp = (C + 63) / 64 * 64;
        std::vector<float> w((size_t)out1 * Cp, 0.0f);
        for (int o = 0; o < out1; ++o) std::memcpy(&w[(size_t)o * Cp], &tw.data[(size_t)o * C], sizeof(float) * C);
        wl1 = upload_f16(w);
        if (d.linear_bias) {
            const auto& tb = find_tensor(tensors, n, std::to_string(layer) + ".linear.bias.tensor");
            bl1 = upload_f32(std::vector<float>(tb.data, tb.data + out1));
        }
        if (d.out_features > 0) {
            if (out1 % 64 != 0) throw Unsupported("out_features must be a multiple of 64");
            const auto& tw2 = find_tensor(tensors, n, std::to_string(layer + 1) + ".linear.weight.tensor");
            wl2 = upload_f16(std::vector<float>(tw2.data, tw2.data + (size_t)d.outsize * out1));
        }
    }
}

LstmModel::~LstmModel() {
    cudaFree(conv_w);
    cudaFree(w3);
    cudaFree(b3);
    for (auto& l : layers) {
        cudaFree(l.w);
        cudaFree(l.bias);
        cudaFree(l.w_ih);
        cudaFree(l.w_hh);
    }
    cudaFree(wl1);
    cudaFree(bl1);
    cudaFree(wl2);
}

size_t LstmModel::workspace_bytes(int N, int T_in) const {
    const int T_out = T_in / desc.stride;
    const size_t x2 = (size_t)N * t_pad(T_in) * 16 * 2 + 4096;
    const size_t seq = (size_t)(T_out + 1) * n_pad(N) * desc.lstm_size * 2 + 4096;
    const size_t mid = desc.out_features > 0 ? (size_t)T_out * n_pad(N) * desc.out_features * 2 + 4096 : 0;
    const size_t gx = hoisted() ? (size_t)T_out * n_pad(N) * 4 * desc.lstm_size * 2 + 4096 : 0;
    return x2 + seq + mid + gx + 4096;  // + the tile counter of conv12_tc_kernel
}

std::unique_ptr<ForwardPlan> LstmModel::make_plan(int N, int T_in, const __half* signal, __half* scores, void* ws,
                                                  size_t ws_bytes) {
    const int C = desc.lstm_size;
    const int T_out = T_in / desc.stride;
    const int Np = n_pad(N);
    if (Np != N) {
        // the reference's tensor-core LSTM has the same kind of constraint (multiples of 64, CudaCaller.h:60-63)
        throw std::invalid_argument(hoisted() ? "batch_size must be a multiple of 32 for this LSTM size"
                                              : "batch_size must be a multiple of 16 for LSTM models");
    }
    auto plan = std::make_unique<LstmPlan>();
    plan->model = this;
    if (const char* dbg = std::getenv("B200_DEBUG_LSTM_LAYERS")) plan->debug_layers = std::atoi(dbg);
    uint8_t* base = static_cast<uint8_t*>(ws);
    auto take = [&](size_t bytes) {
        uint8_t* p = base;
        base += (bytes + 255) & ~size_t(255);
        if ((size_t)(base - static_cast<uint8_t*>(ws)) > ws_bytes) throw std::logic_error("LSTM workspace overflow");
        return p;
    };
    const int Tp = t_pad(T_in);
    __half* x2 = reinterpret_cast<__half*>(take((size_t)N * Tp * 16 * 2));
    __half* seq = reinterpret_cast<__half*>(take((size_t)(T_out + 1) * Np * C * 2));
    __half* mid = desc.out_features > 0 ? reinterpret_cast<__half*>(take((size_t)T_out * Np * desc.out_features * 2)) : nullptr;
    __half* gxbuf = hoisted() ? reinterpret_cast<__half*>(take((size_t)T_out * Np * 4 * C * 2)) : nullptr;
    int* tile_counter = reinterpret_cast<int*>(take(256));

    // conv1 + conv2
    plan->conv12 = Conv12Params{signal, x2, conv_w, N, T_in, Tp, pad3(), desc.convs[0].size, desc.convs[0].winlen,
                                desc.convs[1].winlen, desc.convs[0].activation, desc.convs[1].activation, nullptr, tile_counter};
    plan->conv12_grid = dim3((T_in + CONV_TT - 1) / CONV_TT, N, 1);
    // the v5 shape runs conv2 on the tensor cores (conv12_tc_kernel); anything else keeps the FMA-pipe kernel
    plan->conv12_tc = desc.convs[0].size == 16 && desc.convs[1].winlen == 5 && !std::getenv("B200_CONV12_FMA");
    plan->conv12_tiles_per_chunk = (T_in + C12_TILE - 1) / C12_TILE;
    {
        const long long tiles = (long long)plan->conv12_tiles_per_chunk * N;
        plan->conv12_tc_grid = (int)std::min<long long>(tiles, 5LL * kNumSMs);
        if (plan->conv12_tc) ensure_dynamic_smem(conv12_tc_kernel, C12_SMEM);
    }

    // conv3: rows (n, t) read K3p contiguous halfs starting at x2[n][stride * t]
    {
        if (T_in % desc.stride != 0) throw std::invalid_argument("chunk size must be a multiple of the model stride");
        GemmDesc g{};
        g.a = x2;
        g.batches = N;
        g.rows_per_batch = T_out;
        g.a_row_stride = (int64_t)desc.convs[2].stride * 16;
        g.a_batch_stride = (int64_t)Tp * 16;
        g.w = w3;
        g.N = C;
        g.K = K3p;
        g.bias = b3;
        g.act = desc.convs[2].activation;
        g.out = seq;
        g.out_m1 = T_out;        // g = n * T_out + t
        g.out_s0 = C;            // n
        g.out_s1 = (int64_t)Np * C;  // t
        plan->conv3 = make_gemm_plan(g);
    }
    // LSTM layers
    if (hoisted()) {
        plan->num_layers = desc.lstm_layers;
        plan->hoisted = true;
        if (C != 192 && C != 384) throw Unsupported("hoisted LSTM path is instantiated for lstm_size 192 and 384");
        // Weights-stationary cluster kernel (6 CTAs per cluster).  Chunks per cluster: a step costs about the same for 16, 32 or
        // 64 chunks, so fewer and fatter clusters do the same work in fewer SM-cycles at a higher latency per launch.  With R
        // runners in flight the recurrence takes ~1/R of the SMs (64 chunks per cluster: 8 clusters = 48 SMs at batch 512,
        // three batches' recurrences side by side); a lone runner spreads out (32 per cluster = 96 SMs, 16 for small batches).
        const int hint = num_runners_hint < 1 ? 1 : num_runners_hint;
        int un = Np > 256 ? 32 : 16;
        if (hint >= 2 && Np % 64 == 0 && (Np / 32) * 6 > 148 / hint) un = 64;
        if (const char* e = std::getenv("B200_CLUSTER_CHUNKS")) {   // tuning / A-B override
            const int v = std::atoi(e);
            if ((v == 16 || v == 32 || v == 64) && Np % v == 0) un = v;
            else throw std::invalid_argument("B200_CLUSTER_CHUNKS must be 16, 32 or 64 and divide the padded batch");
        }
        while (Np % un != 0) un /= 2;
        plan->rec_un = un;
        plan->rec_y_map = make_tmap_2d(seq, (uint64_t)C, (uint64_t)T_out * Np, (uint64_t)C * 2, 32, un == 64 ? 32 : 16);
        const int GB = 32;  // chunk block of the gx layout (n_pad = 32 on this path)
        // With three or more batches in flight the x-projection GEMMs keep off 32 SMs, so that another batch's recurrence
        // (48 SMs, several milliseconds of latency chain) can start beside them instead of queueing behind a GEMM that owns every SM
        // (measured at batch 512, 4 runners: 26.4 -> 24.0 ms per step in battery 13, whose files were lost; after the tile-order change
        // caps of 100-148 are within the run-to-run spread, profiles/r02_b26_bench_hac_cap*).
        const int gemm_cap = hint >= 3 ? 116 : 0;
        plan->lstm_grid = (Np / un) * 6;
        for (int l = 0; l < desc.lstm_layers; ++l) {
            GemmDesc g{};
            g.a = layers[l].w_ih;  // gate rows are the M dimension, (t, chunk) the N dimension
            g.batches = 1;
            g.rows_per_batch = 4 * C;
            g.a_row_stride = C;
            g.a_batch_stride = (int64_t)4 * C * C;
            g.w = seq;
            g.N = T_out * Np;
            g.K = C;
            g.bias = layers[l].bias;
            g.bias_per_row = 1;
            g.act = GEMM_ACT_NONE;
            g.out = gxbuf;
            g.out_m1 = 1;
            g.out_s0 = GB;
            g.out_col_m1 = GB;
            g.out_col_s0 = (int64_t)4 * C * GB;
            g.max_ctas = gemm_cap;
            plan->gx_gemm.push_back(make_gemm_plan(g));
            plan->rec_whh.push_back(layers[l].w_hh);
            LstmRecParams rp{};
            rp.seq = seq;
            rp.gx = gxbuf;
            rp.T = T_out;
            rp.N = Np;
            rp.reverse = (l % 2 == 0) ? 1 : 0;
            rp.lens = nullptr;
            rp.stride = desc.stride;
            rp.dbg = nullptr;
            if (l == 0 && getenv("B200_DEBUG_LSTM_TIMELINE")) {
                B200_CUDA(cudaMalloc(&rp.dbg, 128 * sizeof(long long)));
                B200_CUDA(cudaMemset(rp.dbg, 0, 128 * sizeof(long long)));
                plan->dbg_timeline = rp.dbg;
            }
            plan->rec_p.push_back(rp);
        }
    } else {
        plan->num_layers = desc.lstm_layers;
        if (C != 96) throw Unsupported("the single-CTA LSTM kernel is instantiated for lstm_size 96");
        // Grid shape.  A step costs a dependent chain of ~1000-2200 cycles whether a group holds 2 or 8 chunks, so SM-time per
        // chunk falls with chunks per CTA while the latency of one launch rises.  With R runners (batches) in flight the
        // recurrence is therefore sized for ~1/R of the SMs, two groups per CTA, and runs side by side with the other
        // batches' recurrences and decodes (measured at batch 512, profiles/r02_b10_*, r02_b11_*: 4 runners x 32 CTAs x
        // 2 x 8 chunks 5.76 ms/step; 2 runners x 64 CTAs x 2 x 4 chunks 7.27; 2 runners x 128 CTAs x 1 x 4 chunks 8.0-8.9).  A lone
        // runner has nothing to overlap with: one group of 4 chunks on every SM is fastest (0.98 vs 1.17 ms per layer).
        const int hint = num_runners_hint < 1 ? 1 : num_runners_hint;
        int ng = hint == 1 ? 1 : 2, nbr = 16;
        if (const char* e = std::getenv("B200_LSTM_GROUPS")) ng = std::atoi(e) == 1 ? 1 : 2;   // A/B comparisons
        if (Np % (2 * ng) != 0) ng = 1;
        const int target = 148 / hint > 0 ? 148 / hint : 1;
        const int min_nbr = ng == 1 ? 4 : 2;
        for (int v = min_nbr; v <= 16; v *= 2) {
            if (Np % (v * ng) == 0 && (Np / (v * ng) <= target || v == 16)) {
                nbr = v;
                break;
            }
        }
        if (Np % (nbr * ng) != 0) ng = 1;   // Np is a multiple of 16
        // tuning override: with several runners in flight, fatter CTAs (fewer SMs per kernel) let the recurrences of
        // different batches run side by side instead of queueing for the same SMs
        if (const char* e = std::getenv("B200_LSTM_CHUNKS_PER_CTA")) {
            const int v = std::atoi(e);
            if ((v == 2 || v == 4 || v == 8 || v == 16) && Np % (v * ng) == 0) nbr = v;
            else throw std::invalid_argument("B200_LSTM_CHUNKS_PER_CTA must be 2, 4, 8 or 16 and divide the padded batch");
        }
        plan->lstm_nbr = nbr;
        plan->lstm_ng = ng;
        plan->lstm_grid = Np / (nbr * ng);
        plan->lstm_smem = ng == 2 ? LstmCfg<96, 2>::smem_bytes(nbr) : LstmCfg<96, 1>::smem_bytes(nbr);
        for (int l = 0; l < desc.lstm_layers; ++l) {
            plan->lstm_x.push_back(make_tmap_2d(seq, (uint64_t)C, (uint64_t)T_out * Np, (uint64_t)C * 2, KBLK, nbr));
            LstmParams lp{};
            lp.seq = seq;
            lp.bias = layers[l].bias;
            lp.w = layers[l].w;
            lp.T = T_out;
            lp.N = Np;
            lp.C = C;
            lp.reverse = (l % 2 == 0) ? 1 : 0;  // reverse_first = true (CRFModel.cpp:40, LSTMStack.cpp:31-41)
            lp.dbg = nullptr;
            if (l == 0 && getenv("B200_DEBUG_LSTM_TIMELINE")) {
                B200_CUDA(cudaMalloc(&lp.dbg, 64 * sizeof(long long)));
                B200_CUDA(cudaMemset(lp.dbg, 0, 64 * sizeof(long long)));
                plan->dbg_timeline = lp.dbg;
            }
            plan->lstm_p.push_back(lp);
        }
        if (plan->lstm_smem > 227 * 1024) throw Unsupported("LSTM shared-memory plan does not fit");
    }
    // linear CRF (+ optional decomposition); rows g = t * Np + n  ->  scores[n][t][:]
    {
        GemmDesc g{};
        g.a = seq;
        g.batches = 1;
        g.rows_per_batch = T_out * Np;
        g.a_row_stride = C;
        g.a_batch_stride = (int64_t)T_out * Np * C;
        g.w = wl1;
        g.N = out1;
        g.K = Cp;
        g.a_inner = C;
        g.bias = bl1;
        if (desc.out_features > 0) {
            g.act = GEMM_ACT_NONE;
            g.out = mid;
            g.out_m1 = 1;
            g.out_s0 = out1;
            g.out_s1 = 0;
            plan->linear1 = make_gemm_plan(g);
            GemmDesc g2{};
            g2.a = mid;
            g2.batches = 1;
            g2.rows_per_batch = T_out * Np;
            g2.a_row_stride = out1;
            g2.a_batch_stride = (int64_t)T_out * Np * out1;
            g2.w = wl2;
            g2.N = desc.outsize;
            g2.K = out1;
            g2.act = desc.crf_scale == 5.0f ? GEMM_ACT_TANH_X5 : GEMM_ACT_NONE;
            g2.out = scores;
            g2.out_m1 = Np;
            g2.out_s0 = desc.outsize;
            g2.out_s1 = (int64_t)T_out * desc.outsize;
            plan->linear2 = make_gemm_plan(g2);
            plan->num_linear = 2;
        } else {
            g.act = desc.crf_scale == 5.0f ? GEMM_ACT_TANH_X5 : GEMM_ACT_NONE;
            g.out = scores;
            g.out_m1 = Np;                                  // g = t * Np + n
            g.out_s0 = desc.outsize;                        // t
            g.out_s1 = (int64_t)T_out * desc.outsize;       // n
            plan->linear1 = make_gemm_plan(g);
            plan->num_linear = 1;
        }
    }
    return plan;
}

template <int C, int NBR, int NG>
static void launch_lstm_t(const LstmPlan& pl, int l, cudaStream_t stream) {
    ensure_dynamic_smem(lstm_layer_kernel<C, NBR, NG>, (int)LstmCfg<C, NG>::smem_bytes(NBR));
    lstm_layer_kernel<C, NBR, NG><<<pl.lstm_grid, LstmCfg<C, NG>::THREADS, pl.lstm_smem, stream>>>(pl.lstm_x[l], pl.lstm_p[l]);
}

template <int C, int CL, int NG, int GN>
static void launch_cluster_t(const LstmPlan& pl, int l, cudaStream_t stream) {
    using Cfg = ClusterCfg<C, CL, NG, GN>;
    ensure_dynamic_smem(lstm_cluster_kernel<C, CL, NG, GN>, (int)Cfg::SMEM);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)pl.lstm_grid, 1, 1);
    cfg.blockDim = dim3(Cfg::THREADS, 1, 1);
    cfg.dynamicSmemBytes = Cfg::SMEM;
    cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CL;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    B200_CUDA(cudaLaunchKernelEx(&cfg, lstm_cluster_kernel<C, CL, NG, GN>, pl.rec_y_map, pl.rec_whh[l], pl.rec_p[l]));
}

template <int C>
static void launch_cluster_c(const LstmPlan& pl, int l, cudaStream_t stream) {
    switch (pl.rec_un) {
        case 64: launch_cluster_t<C, 6, 2, 32>(pl, l, stream); break;
        case 32: launch_cluster_t<C, 6, 2, 16>(pl, l, stream); break;
        default: launch_cluster_t<C, 6, 1, 16>(pl, l, stream); break;
    }
}

void LstmPlan::launch_rec(int l, cudaStream_t stream) const {
    const int C = model->desc.lstm_size;
    if (C == 384) launch_cluster_c<384>(*this, l, stream);
    else if (C == 192) launch_cluster_c<192>(*this, l, stream);
    else throw Unsupported("no cluster LSTM kernel instantiation for this lstm_size");
}

template <int C>
static void launch_lstm_c(const LstmPlan& pl, int l, cudaStream_t stream) {
    if (pl.lstm_ng == 2) {
        switch (pl.lstm_nbr) {
            case 2: launch_lstm_t<C, 2, 2>(pl, l, stream); break;
            case 4: launch_lstm_t<C, 4, 2>(pl, l, stream); break;
            case 8: launch_lstm_t<C, 8, 2>(pl, l, stream); break;
            default: launch_lstm_t<C, 16, 2>(pl, l, stream); break;
        }
    } else {
        switch (pl.lstm_nbr) {
            case 2: launch_lstm_t<C, 2, 1>(pl, l, stream); break;
            case 4: launch_lstm_t<C, 4, 1>(pl, l, stream); break;
            case 8: launch_lstm_t<C, 8, 1>(pl, l, stream); break;
            default: launch_lstm_t<C, 16, 1>(pl, l, stream); break;
        }
    }
}

void LstmPlan::launch_lstm(int l, cudaStream_t stream) const {
    switch (lstm_p[l].C) {
        case 96: launch_lstm_c<96>(*this, l, stream); break;
        default: throw Unsupported("no LSTM kernel instantiation for this lstm_size");
    }
}

void LstmPlan::run(cudaStream_t stream, ProfileSink* prof) {
    {
        NvtxRange r("conv");
        if (conv12_tc) {
            B200_CUDA(cudaMemsetAsync(conv12.tile_counter, 0, sizeof(int), stream));
            conv12_tc_kernel<<<conv12_tc_grid, C12_THREADS, C12_SMEM, stream>>>(conv12, conv12_tiles_per_chunk,
                                                                                 conv12_tiles_per_chunk * conv12.N);
        } else {
            conv12_kernel<<<conv12_grid, CONV_TT, 0, stream>>>(conv12);
        }
        if (prof) prof->mark("conv12", stream);
        run_gemm(conv3, stream);
        if (prof) prof->mark("conv3_gemm", stream);
    }
    {
        NvtxRange lstm_range("lstm_stack");
        const int nl = debug_layers >= 0 && debug_layers < num_layers ? debug_layers : num_layers;
        for (int l = 0; l < nl; ++l) {
            NvtxRange r("lstm_layer");
            if (hoisted) {
                run_gemm(gx_gemm[l], stream);
                if (prof) prof->mark("lstm_gx_gemm", stream);
                launch_rec(l, stream);
                if (prof) prof->mark("lstm_rec", stream);
            } else {
                launch_lstm(l, stream);
                if (prof) prof->mark("lstm_layer", stream);
            }
        }
        if (nl < num_layers) {  // debug: stop after nl layers (tools/debug_forward.py reads the sequence buffer)
            B200_CUDA(cudaGetLastError());
            return;
        }
    }
    if (dbg_timeline) {
        long long h[128];
        B200_CUDA(cudaStreamSynchronize(stream));
        B200_CUDA(cudaMemcpy(h, dbg_timeline, hoisted ? sizeof(h) : 64 * sizeof(long long), cudaMemcpyDeviceToHost));
        for (int s = 0; s < 4; ++s) {
            if (hoisted) {
                const long long* e = h + s * 32;
                const long long t0 = e[0];
                fprintf(stderr, "[cluster timeline step %d] mma g0: wait %lld..%lld issued %lld | g1: wait %lld..%lld issued %lld | epi(g0,t0): "
                                "wait_acc %lld..%lld ld %lld act %lld transpose %lld cells+stage %lld fence %lld wait_read %lld bar %lld sent %lld | "
                                "period %lld\n",
                        64 + s, e[0] - t0, e[1] - t0, e[2] - t0, e[4] - t0, e[5] - t0, e[6] - t0, e[8] - t0, e[9] - t0, e[10] - t0, e[11] - t0,
                        e[12] - t0, e[13] - t0, e[14] - t0, e[15] - t0, e[16] - t0, e[17] - t0, s > 0 ? e[0] - (e - 32)[0] : 0LL);
            } else {
                const long long* e = h + s * 16;
                const long long t0 = e[0];
                fprintf(stderr, "[lstm timeline step %d] mma: wait_h 0..%lld issued %lld | epi: wait_acc %lld..%lld ld %lld act %lld "
                                "transpose/bar %lld cells %lld fence %lld arrive %lld store %lld | period %lld\n",
                        64 + s, e[1] - t0, e[2] - t0, e[4] - t0, e[5] - t0, e[6] - t0, e[7] - t0, e[8] - t0, e[9] - t0, e[10] - t0, e[11] - t0,
                        e[12] - t0, s > 0 ? e[0] - (e - 16)[0] : 0LL);
            }
        }
    }
    NvtxRange r("linear");
    run_gemm(linear1, stream);
    if (prof) prof->mark("linear_gemm", stream);
    if (num_linear == 2) {
        run_gemm(linear2, stream);
        if (prof) prof->mark("linear2_gemm", stream);
    }
    B200_CUDA(cudaGetLastError());
}

}  // namespace

std::unique_ptr<Model> make_lstm_model(const b200_model_desc& desc, const b200_tensor* tensors, int n) {
    return std::make_unique<LstmModel>(desc, tensors, n);
}

}  // namespace b200

/* b200_crf_math.h -- the numerics contract of the CRF decoder.
 *
 * The reference decoder (dorado/basecall/decode/beam_search.cpp:42-45, :94-98, :503 and
 * CPUDecoder.cpp:29-34, :130) evaluates exp/log/log1p/pow/log10 through libm and libtorch, whose
 * last-bit behaviour is platform specific and cannot be reproduced on a GPU.  Because the beam
 * search is full of hard thresholds, "bit-identical" output needs bit-identical transcendentals, so
 * this header pins them: every function below is built only from IEEE-754 binary32 add / mul / fma /
 * div (all correctly rounded on x86-64 and on sm_100a) in a fixed order.  The CUDA kernels
 * (dorado_b200/csrc/decode.cu) and the CPU oracle (oracle/crf_oracle.c) both include this file, so
 * the two produce the same bits by construction.  Accuracy versus libm: <= 2 ulp over the ranges
 * the decoder uses (tests/test_host_cpu.py::test_numerics_contract_accuracy).
 *
 * Compile rules: C side with -ffp-contract=off (oracle/Makefile); CUDA side uses the explicit
 * __f*_rn intrinsics, which nvcc never contracts.
 */
#ifndef B200_CRF_MATH_H
#define B200_CRF_MATH_H

#include <stdint.h>

#if defined(__CUDA_ARCH__)
#define B200_MATH_FN static __device__ __forceinline__
#define B200_FMA(a, b, c) __fmaf_rn((a), (b), (c))
#define B200_MUL(a, b) __fmul_rn((a), (b))
#define B200_ADD(a, b) __fadd_rn((a), (b))
#define B200_SUB(a, b) __fsub_rn((a), (b))
#define B200_DIV(a, b) __fdiv_rn((a), (b))
#define B200_F2U(x) __float_as_uint(x)
#define B200_U2F(x) __uint_as_float(x)
#else
#include <math.h>
#include <string.h>
#define B200_MATH_FN static inline
#define B200_FMA(a, b, c) fmaf((a), (b), (c))
#define B200_MUL(a, b) ((a) * (b))
#define B200_ADD(a, b) ((a) + (b))
#define B200_SUB(a, b) ((a) - (b))
#define B200_DIV(a, b) ((a) / (b))
static inline uint32_t b200_f2u_(float x) {
    uint32_t u;
    memcpy(&u, &x, 4);
    return u;
}
static inline float b200_u2f_(uint32_t u) {
    float x;
    memcpy(&x, &u, 4);
    return x;
}
#define B200_F2U(x) b200_f2u_(x)
#define B200_U2F(x) b200_u2f_(x)
#endif

#define B200_FLT_LOWEST (-3.402823466e+38f)

B200_MATH_FN float b200_fmaxf(float a, float b) { return a > b ? a : b; }

/* exp(x) for any finite x <= ~88; returns 0 below -86 (results there would be subnormal). */
B200_MATH_FN float b200_expf(float x) {
    if (x < -86.0f) {
        return 0.0f;
    }
    if (x > 88.0f) {
        x = 88.0f;
    }
    /* n = round-to-nearest-even(x / ln2) via the 1.5*2^23 trick */
    const float t = B200_MUL(x, 1.44269504088896341f);
    const float big = 12582912.0f;
    const float n = B200_SUB(B200_ADD(t, big), big);
    float r = B200_FMA(n, -0.693359375f, x);
    r = B200_FMA(n, 2.12194440e-4f, r);
    float p = 1.9875691500e-4f;
    p = B200_FMA(p, r, 1.3981999507e-3f);
    p = B200_FMA(p, r, 8.3334519073e-3f);
    p = B200_FMA(p, r, 4.1665795894e-2f);
    p = B200_FMA(p, r, 1.6666665459e-1f);
    p = B200_FMA(p, r, 5.0000001201e-1f);
    const float r2 = B200_MUL(r, r);
    p = B200_FMA(p, r2, r);
    p = B200_ADD(p, 1.0f);
    const int32_t ni = (int32_t)n;
    const float scale = B200_U2F((uint32_t)(ni + 127) << 23);
    return B200_MUL(p, scale);
}

/* log(x) for normal x > 0. */
B200_MATH_FN float b200_logf(float x) {
    const uint32_t ix = B200_F2U(x);
    int32_t e = (int32_t)((ix >> 23) & 0xffu) - 126;
    float m = B200_U2F((ix & 0x007fffffu) | 0x3f000000u); /* [0.5, 1) */
    if (m < 0.707106781186547524f) {
        e -= 1;
        m = B200_SUB(B200_ADD(m, m), 1.0f);
    } else {
        m = B200_SUB(m, 1.0f);
    }
    const float z = B200_MUL(m, m);
    float y = 7.0376836292e-2f;
    y = B200_FMA(y, m, -1.1514610310e-1f);
    y = B200_FMA(y, m, 1.1676998740e-1f);
    y = B200_FMA(y, m, -1.2420140846e-1f);
    y = B200_FMA(y, m, 1.4249322787e-1f);
    y = B200_FMA(y, m, -1.6668057665e-1f);
    y = B200_FMA(y, m, 2.0000714765e-1f);
    y = B200_FMA(y, m, -2.4999993993e-1f);
    y = B200_FMA(y, m, 3.3333331174e-1f);
    y = B200_MUL(B200_MUL(y, m), z);
    const float fe = (float)e;
    y = B200_FMA(fe, -2.12194440e-4f, y);
    y = B200_FMA(z, -0.5f, y);
    float r = B200_ADD(m, y);
    r = B200_FMA(fe, 0.693359375f, r);
    return r;
}

/* log1p(x) for x in [0, 1]: log(u) corrected for the rounding of u = 1 + x. */
B200_MATH_FN float b200_log1pf(float x) {
    const float u = B200_ADD(1.0f, x);
    const float c = B200_SUB(x, B200_SUB(u, 1.0f)); /* exact rounding error of 1 + x */
    return B200_ADD(b200_logf(u), B200_DIV(c, u));
}

/* beam_search.cpp:42-45 */
B200_MATH_FN float b200_log_sum_exp(float x, float y) {
    const float d = B200_SUB(x, y);
    const float ad = d < 0.0f ? -d : d;
    const float mx = b200_fmaxf(x, y);
    if (ad < 17.0f) {
        return B200_ADD(mx, b200_log1pf(b200_expf(-ad)));
    }
    return mx;
}

/* CPUDecoder.cpp:29-34: logsumexp over {stay, step0..3}: max-shifted, summed in this order. */
B200_MATH_FN float b200_lse5(float v_stay, float v0, float v1, float v2, float v3) {
    float m = b200_fmaxf(v_stay, v0);
    m = b200_fmaxf(m, v1);
    m = b200_fmaxf(m, v2);
    m = b200_fmaxf(m, v3);
    float s = b200_expf(B200_SUB(v_stay, m));
    s = B200_ADD(s, b200_expf(B200_SUB(v0, m)));
    s = B200_ADD(s, b200_expf(B200_SUB(v1, m)));
    s = B200_ADD(s, b200_expf(B200_SUB(v2, m)));
    s = B200_ADD(s, b200_expf(B200_SUB(v3, m)));
    return B200_ADD(m, b200_logf(s));
}

/* beam_search.cpp:503: pow(p, 0.4f) for p in [0, 1]. */
B200_MATH_FN float b200_pow0p4f(float p) {
    if (p <= 0.0f) {
        return 0.0f;
    }
    if (p >= 1.0f) {
        return 1.0f;
    }
    if (p < 1.17549435e-38f) {
        return 0.0f; /* subnormal probabilities: 0 (pow would give < 2e-15) */
    }
    return b200_expf(B200_MUL(0.4f, b200_logf(p)));
}

/* beam_search.cpp:94-98: per-base quality character from the accumulated probabilities. */
B200_MATH_FN char b200_qchar(float base_prob, float total_prob, float scale, float shift) {
    float e = B200_SUB(1.0f, B200_DIV(base_prob, total_prob));
    float q;
    if (e <= 0.0f) {
        q = 50.0f; /* -10*log10(0) = +inf -> clamps to 50 (NaN from 0/0 never reaches here: see oracle) */
    } else {
        if (e < 1.17549435e-38f) {
            e = 1.17549435e-38f;
        }
        const float l10 = B200_MUL(b200_logf(e), 0.434294481903251828f);
        q = B200_ADD(B200_MUL(B200_MUL(-10.0f, l10), scale), shift);
        q = q < 1.0f ? 1.0f : (q > 50.0f ? 50.0f : q);
    }
    return (char)(int)B200_ADD(33.5f, q);
}

#endif /* B200_CRF_MATH_H */

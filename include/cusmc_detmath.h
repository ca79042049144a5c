/*
 * cusmc_detmath.h -- deterministic elementary functions and the fixed-point weight image.
 *
 * Everything here is built from IEEE-754 basic operations (add, mul, fma, div, sqrt,
 * conversions) in a FIXED order, so the same inputs give the same bits on the device
 * (nvcc -fmad=false) and on any host compiler that does not contract floating point
 * (gcc -ffp-contract=off).  That is what lets resampling decisions taken on the GPU be
 * reproduced bit-for-bit on a CPU: libm's and CUDA's exp/log differ in the last ulp,
 * these do not.  The CPU oracle carries an independent restatement (orc_det_exp,
 * orc_det_log, orc_fixed_weights) that the tests compare against.
 *
 * There is no counterpart in the reference (it never normalises weights and works in the
 * linear domain); semantics are defined here.
 */
#ifndef CUSMC_DETMATH_H
#define CUSMC_DETMATH_H

#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define CUSMC_HD __host__ __device__ __forceinline__
#else
#define CUSMC_HD static inline
#endif

CUSMC_HD double cusmc_bits_to_double(uint64_t b)
{
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)b);
#else
    double d;
    memcpy(&d, &b, 8);
    return d;
#endif
}

CUSMC_HD uint64_t cusmc_double_to_bits(double d)
{
#if defined(__CUDA_ARCH__)
    return (uint64_t)__double_as_longlong(d);
#else
    uint64_t b;
    memcpy(&b, &d, 8);
    return b;
#endif
}

/* 2^e for -1022 <= e <= 1023 */
CUSMC_HD double cusmc_pow2i(int e) { return cusmc_bits_to_double((uint64_t)(e + 1023) << 52); }

/* Taylor coefficients 1/13! .. 1/0! of exp.  On the device they sit in constant memory so each
 * DFMA of the Horner chain takes its coefficient as a constant-bank operand (no register or
 * uniform-register moves); on the host they are the same correctly rounded quotients. */
#define CUSMC_EXP_COEFS                                                                          \
    { 1.0 / 6227020800.0, 1.0 / 479001600.0, 1.0 / 39916800.0, 1.0 / 3628800.0, 1.0 / 362880.0,  \
      1.0 / 40320.0, 1.0 / 5040.0, 1.0 / 720.0, 1.0 / 120.0, 1.0 / 24.0, 1.0 / 6.0, 0.5, 1.0, 1.0 }
#if defined(__CUDACC__)
static __constant__ double cusmc_exp_coef_dev[14] = CUSMC_EXP_COEFS;
#endif
static const double cusmc_exp_coef_host[14] = CUSMC_EXP_COEFS;
#if defined(__CUDA_ARCH__)
#define CUSMC_EXP_C(i) cusmc_exp_coef_dev[i]
#else
#define CUSMC_EXP_C(i) cusmc_exp_coef_host[i]
#endif

/* The reduced exponential: x = k ln2 + r with a two-term ln2, degree-13 Taylor polynomial in
 * Horner/fma form.  Returns p = exp(r) and k. */
CUSMC_HD double cusmc_det_exp_core(double x, int *k_out)
{
    const double kf = rint(x * 1.4426950408889634074);
    double r = fma(kf, -6.93147180369123816490e-01, x);
    r = fma(kf, -1.90821492927058770002e-10, r);
    double p = CUSMC_EXP_C(0);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 1; i < 14; ++i) p = fma(p, r, CUSMC_EXP_C(i));
    *k_out = (int)kf;
    return p;
}

/* exp(x) with exact scaling by 2^k (gradual underflow handled in two steps). */
CUSMC_HD double cusmc_det_exp(double x)
{
    if (x != x) return x;
    if (x > 709.782712893384) return cusmc_bits_to_double(0x7FF0000000000000ull);
    if (x < -745.2) return 0.0;
    int k;
    const double p = cusmc_det_exp_core(x, &k);
    if (k >= -1021 && k <= 1023) return p * cusmc_pow2i(k);
    if (k > 1023) return (p * cusmc_pow2i(k - 1)) * 2.0;
    return (p * cusmc_pow2i(k + 1022)) * cusmc_pow2i(-1022);
}

/* exp(x) for the weight image: x <= 0, and anything below 2^-62 may be flushed to zero (it
 * truncates to q = 0 for every shift <= 61).  Bit-identical to cusmc_det_exp on [-43.5, 0];
 * NaN, -inf and x < -43.5 give 0.  One compare, one scale: the hot form. */
CUSMC_HD double cusmc_det_exp_unit(double x)
{
    /* branch-free on purpose: a thread that owns several weights gets its exp chains interleaved
     * and the polynomial coefficients shared between them */
    const int ok = x >= -43.5;
    int k;
    const double p = cusmc_det_exp_core(ok ? x : 0.0, &k);
    const double v = p * cusmc_pow2i(k);
    return ok ? v : 0.0;
}

/* log(x): x = m 2^e, m in [sqrt(1/2), sqrt 2), log m = 2 atanh(s), s = (m-1)/(m+1). */
CUSMC_HD double cusmc_det_log(double x)
{
    if (x != x || x < 0.0) return cusmc_bits_to_double(0x7FF8000000000000ull);
    if (x == 0.0) return cusmc_bits_to_double(0xFFF0000000000000ull);
    uint64_t b = cusmc_double_to_bits(x);
    if (b == 0x7FF0000000000000ull) return x;
    int e = 0;
    if ((b >> 52) == 0) {
        x = x * 18014398509481984.0; /* 2^54, exact */
        b = cusmc_double_to_bits(x);
        e = -54;
    }
    e += (int)(b >> 52) - 1023;
    double m = cusmc_bits_to_double((b & 0x000FFFFFFFFFFFFFull) | 0x3FF0000000000000ull);
    if (m > 1.4142135623730951) {
        m = m * 0.5;
        e += 1;
    }
    const double f = m - 1.0;
    const double s = f / (2.0 + f);
    const double s2 = s * s;
    double p = 1.0 / 23.0;
    p = fma(p, s2, 1.0 / 21.0);
    p = fma(p, s2, 1.0 / 19.0);
    p = fma(p, s2, 1.0 / 17.0);
    p = fma(p, s2, 1.0 / 15.0);
    p = fma(p, s2, 1.0 / 13.0);
    p = fma(p, s2, 1.0 / 11.0);
    p = fma(p, s2, 1.0 / 9.0);
    p = fma(p, s2, 1.0 / 7.0);
    p = fma(p, s2, 1.0 / 5.0);
    p = fma(p, s2, 1.0 / 3.0);
    p = p * s2;
    const double two_s = s + s;
    const double lo = fma(two_s, p, (double)e * 1.90821492927058770002e-10);
    return fma((double)e, 6.93147180369123816490e-01, two_s + lo);
}

/* sin(pi t), cos(pi t) for t in [0, 2): quadrant n = rint(2t), f = t - n/2 in [-1/4, 1/4]
 * (exact), Taylor polynomials in x = pi f. */
CUSMC_HD void cusmc_det_sincospi(double t, double *s_out, double *c_out)
{
    const double nf = rint(t + t);
    const double f = fma(nf, -0.5, t);
    const double x = fma(f, 3.141592653589793116, f * 1.2246467991473532e-16);
    const double x2 = x * x;
    double ps = -1.0 / 121645100408832000.0;          /* -1/19! */
    ps = fma(ps, x2, 1.0 / 355687428096000.0);        /*  1/17! */
    ps = fma(ps, x2, -1.0 / 1307674368000.0);         /* -1/15! */
    ps = fma(ps, x2, 1.0 / 6227020800.0);             /*  1/13! */
    ps = fma(ps, x2, -1.0 / 39916800.0);              /* -1/11! */
    ps = fma(ps, x2, 1.0 / 362880.0);                 /*  1/9!  */
    ps = fma(ps, x2, -1.0 / 5040.0);                  /* -1/7!  */
    ps = fma(ps, x2, 1.0 / 120.0);                    /*  1/5!  */
    ps = fma(ps, x2, -1.0 / 6.0);                     /* -1/3!  */
    const double sn = fma(x * x2, ps, x);
    double pc = 1.0 / 2432902008176640000.0;          /*  1/20! */
    pc = fma(pc, x2, -1.0 / 6402373705728000.0);      /* -1/18! */
    pc = fma(pc, x2, 1.0 / 20922789888000.0);         /*  1/16! */
    pc = fma(pc, x2, -1.0 / 87178291200.0);           /* -1/14! */
    pc = fma(pc, x2, 1.0 / 479001600.0);              /*  1/12! */
    pc = fma(pc, x2, -1.0 / 3628800.0);               /* -1/10! */
    pc = fma(pc, x2, 1.0 / 40320.0);                  /*  1/8!  */
    pc = fma(pc, x2, -1.0 / 720.0);                   /* -1/6!  */
    pc = fma(pc, x2, 1.0 / 24.0);                     /*  1/4!  */
    pc = fma(pc, x2, -0.5);                           /* -1/2!  */
    const double cs = fma(x2, pc, 1.0);
    const int n = ((int)nf) & 3;
    *s_out = (n == 0) ? sn : (n == 1) ? cs : (n == 2) ? -sn : -cs;
    *c_out = (n == 0) ? cs : (n == 1) ? -sn : (n == 2) ? -cs : sn;
}


/* ---- single-precision twins (FFMA + exact bit manipulation only) ---------------------------------
 * Used by the in-kernel normal generator: its draws need 24-bit, not 53-bit, resolution, and on
 * the device the fp64 versions above cost ~7x more issue slots (fp64 divide / sqrt / 64-bit
 * int->double conversions expand into long instruction sequences). */
CUSMC_HD float cusmc_bits_to_float(uint32_t b)
{
#if defined(__CUDA_ARCH__)
    return __uint_as_float(b);
#else
    float f;
    memcpy(&f, &b, 4);
    return f;
#endif
}

CUSMC_HD uint32_t cusmc_float_to_bits(float f)
{
#if defined(__CUDA_ARCH__)
    return __float_as_uint(f);
#else
    uint32_t b;
    memcpy(&b, &f, 4);
    return b;
#endif
}

/* log(x) for normal positive floats: x = m 2^e, m in [sqrt(1/2), sqrt 2), log m = f + f^2 P(f),
 * f = m - 1, P a degree-7 least-squares fit (max abs error 3.9e-8). */
CUSMC_HD float cusmc_det_logf(float x)
{
    const uint32_t b = cusmc_float_to_bits(x);
    int e = (int)(b >> 23) - 127;
    float m = cusmc_bits_to_float((b & 0x007FFFFFu) | 0x3F800000u);
    if (m > 1.41421354f) {
        m = m * 0.5f;
        e += 1;
    }
    const float f = m - 1.0f;
    float p = 0.09004202485084534f;
    p = fmaf(p, f, -0.14257794618606567f);
    p = fmaf(p, f, 0.14806459844112396f);
    p = fmaf(p, f, -0.16575047373771667f);
    p = fmaf(p, f, 0.19973105192184448f);
    p = fmaf(p, f, -0.25001609325408936f);
    p = fmaf(p, f, 0.33333659172058105f);
    p = fmaf(p, f, -0.4999999403953552f);
    const float r = fmaf(p * f, f, f);
    return fmaf((float)e, 0.693147182464599609375f, r);
}

/* sin(pi t), cos(pi t) for t in [0, 2), t a multiple of 2^-23. */
CUSMC_HD void cusmc_det_sincospif(float t, float *s_out, float *c_out)
{
    const float nf = rintf(t + t);
    const float f = fmaf(nf, -0.5f, t);                 /* exact, in [-1/4, 1/4] */
    const float x = f * 3.14159274101257324f;
    const float x2 = x * x;
    float ps = 2.75573192e-6f;                          /*  1/9! */
    ps = fmaf(ps, x2, -1.98412698e-4f);                 /* -1/7! */
    ps = fmaf(ps, x2, 8.33333377e-3f);                  /*  1/5! */
    ps = fmaf(ps, x2, -1.66666672e-1f);                 /* -1/3! */
    const float sn = fmaf(x * x2, ps, x);
    float pc = 2.48015876e-5f;                          /*  1/8! */
    pc = fmaf(pc, x2, -1.38888892e-3f);                 /* -1/6! */
    pc = fmaf(pc, x2, 4.16666679e-2f);                  /*  1/4! */
    pc = fmaf(pc, x2, -0.5f);
    const float cs = fmaf(x2, pc, 1.0f);
    /* quadrant n: (s, c) = (sn, cs), (cs, -sn), (-sn, -cs), (-cs, sn) -- a swap on odd n and two
     * sign flips, written on the bit patterns so it compiles to selects and XORs (no branches) */
    const uint32_t n = (uint32_t)(int)nf;
    const uint32_t sb = cusmc_float_to_bits((n & 1u) ? cs : sn);
    const uint32_t cb = cusmc_float_to_bits((n & 1u) ? sn : cs);
    *s_out = cusmc_bits_to_float(sb ^ ((n & 2u) << 30));
    *c_out = cusmc_bits_to_float(cb ^ (((n + 1u) & 2u) << 30));
}

/* ---- fixed-point weight image ------------------------------------------------------------
 * shift = 61 - ceil(log2(N_global)): N_global weights in [0, 2^shift] sum to at most 2^61, so
 * a prefix sum fits 62 bits and the two top bits of a 64-bit word stay free for the
 * decoupled look-back status flags. */
CUSMC_HD int cusmc_fixed_shift(int64_t n_global)
{
    int b = 0;
    while (((int64_t)1 << b) < n_global) ++b;
    return 61 - b;
}

/* wn in [0, 1] -> trunc(wn * 2^shift); anything else (NaN, negative) -> 0. */
CUSMC_HD uint64_t cusmc_fixed_from_unit(double wn, int shift)
{
    const double c = (wn > 0.0) ? (wn > 1.0 ? 1.0 : wn) : 0.0;   /* select form: no branches */
    return (uint64_t)(c * cusmc_pow2i(shift));
}

/* Linear-domain weight w relative to wmax. */
CUSMC_HD double cusmc_unit_from_linear(double w, double wmax)
{
    if (!(w > 0.0) || !(w <= wmax)) return 0.0;
    return w / wmax;
}

/* Log-domain weight lw relative to lmax. */
CUSMC_HD double cusmc_unit_from_log(double lw, double lmax)
{
    const double v = cusmc_det_exp_unit(lw - lmax);
    return (lw <= lmax) ? v : 0.0; /* NaN or above the max -> 0 */
}

/* ---- block-relative weight image (the filter's resampling image; oracle: orc_tile_image) ----------
 * A tile of particles is weighed against ITS OWN maximum m_b, so the pass that produces the log-weights
 * can finish the tile without a grid-wide dependency; once the global maximum M is known the tile is
 * rescaled by the 62-bit fixed-point factor F_b = trunc(exp(m_b - M) 2^62):  c -> (c F_b) >> 62. */
CUSMC_HD uint64_t cusmc_rescale_factor(double m_b, double M)
{
    const double v = cusmc_unit_from_log(m_b, M);          /* in [0, 1]; -inf / NaN -> 0 */
    return (uint64_t)(v * 4611686018427387904.0);          /* 2^62 */
}

/* floor(c F / 2^62) for c < 2^62, F <= 2^62 (128-bit product). */
CUSMC_HD uint64_t cusmc_mulshift62(uint64_t c, uint64_t F)
{
#if defined(__CUDA_ARCH__)
    return (__umul64hi(c, F) << 2) | ((c * F) >> 62);
#else
    return (uint64_t)(((unsigned __int128)c * F) >> 62);
#endif
}

#endif /* CUSMC_DETMATH_H */

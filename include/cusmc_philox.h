/*
 * cusmc_philox.h -- counter-based randomness (Philox4x32-10, Salmon et al., SC'11) and the
 * fixed (seed, stream, step, index) -> draw mapping the kernels use when the caller does not
 * inject pre-drawn numbers.
 *
 * Replaces the reference's one-XORWOW-state-per-scalar scheme (curand_init per thread,
 * src/mvn_dist.cu.cpp:15-31, src/mvt_dist.cu.cpp:63-82; 48 bytes of state per scalar, seeded
 * from srand(time), not reproducible) with a stateless generator: integer arithmetic only, so
 * a host can regenerate exactly the numbers a kernel consumed.  Normals come from Box-Muller
 * on cusmc_detmath.h functions and are therefore bit-reproducible too.
 *
 * Counter layout:  ctr = { index_lo, index_hi, (uint32) step, stream | (sub << 8) }
 *                  key = { seed_lo, seed_hi }
 */
#ifndef CUSMC_PHILOX_H
#define CUSMC_PHILOX_H

#include "cusmc_detmath.h"

enum cusmc_rng_stream {
    CUSMC_STREAM_METROPOLIS = 0, /* sub = n (0..B-1): out[0..1] -> u, out[2..3] -> j */
    CUSMC_STREAM_NORMAL = 1,     /* sub = component quad k/4: four normals per call */
    CUSMC_STREAM_CHI = 2,        /* sub = component k */
    CUSMC_STREAM_MULTINOMIAL = 3,
    CUSMC_STREAM_CHAIN_Z = 4,    /* index = chain, step = MH step, sub = component quad */
    CUSMC_STREAM_CHAIN_U = 5,
    CUSMC_STREAM_INIT = 6,
    CUSMC_STREAM_OFFSET = 7,     /* the systematic offset u0 of a step: index = sub = 0 */
    CUSMC_STREAM_SEGMENT = 8     /* Metropolis-C2: index = i / 32 (the warp), sub = n: out[0..1] -> the proposal segment */
};

typedef struct cusmc_u32x4 {
    uint32_t v[4];
} cusmc_u32x4;

CUSMC_HD cusmc_u32x4 cusmc_philox4x32_rounds(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                             uint32_t k0, uint32_t k1, int rounds)
{
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int round = 0; round < rounds; ++round) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        c1 = (uint32_t)p1;
        c3 = (uint32_t)p0;
        c0 = n0;
        c2 = n2;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    cusmc_u32x4 out;
    out.v[0] = c0;
    out.v[1] = c1;
    out.v[2] = c2;
    out.v[3] = c3;
    return out;
}

CUSMC_HD cusmc_u32x4 cusmc_philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                         uint32_t k0, uint32_t k1)
{
    return cusmc_philox4x32_rounds(c0, c1, c2, c3, k0, k1, 10);
}

CUSMC_HD cusmc_u32x4 cusmc_rng(uint64_t seed, int stream, uint64_t step, uint64_t index, uint32_t sub)
{
    return cusmc_philox4x32_10((uint32_t)index, (uint32_t)(index >> 32), (uint32_t)step,
                               (uint32_t)stream | (sub << 8), (uint32_t)seed, (uint32_t)(seed >> 32));
}

/* The same counter / key layout with 7 rounds: Philox4x32-7, the shortest variant Salmon et al. report
 * as passing BigCrush (10 rounds is their recommended safety margin, used everywhere a host may have
 * to reproduce the draws).  Only the throughput noise path uses it (cusmc_filter_config.reproducible_rng
 * = 0), where the 3 rounds are 22 of a d = 8 particle-step's ~440 instructions per block. */
CUSMC_HD cusmc_u32x4 cusmc_rng7(uint64_t seed, int stream, uint64_t step, uint64_t index, uint32_t sub)
{
    return cusmc_philox4x32_rounds((uint32_t)index, (uint32_t)(index >> 32), (uint32_t)step,
                                   (uint32_t)stream | (sub << 8), (uint32_t)seed, (uint32_t)(seed >> 32), 7);
}

/* 53-bit uniform in [0, 1). */
CUSMC_HD double cusmc_u01(uint32_t hi, uint32_t lo)
{
    const uint64_t bits = (((uint64_t)hi << 32) | lo) >> 11;
    return (double)bits * 1.1102230246251565e-16; /* 2^-53 */
}

/* 53-bit uniform in (0, 1] (safe argument for log). */
CUSMC_HD double cusmc_u01_open0(uint32_t hi, uint32_t lo)
{
    const uint64_t bits = ((((uint64_t)hi << 32) | lo) >> 11) + 1;
    return (double)bits * 1.1102230246251565e-16;
}

/* Uniform integer in [0, n): high word of a 64 x 64 -> 128 bit product (Lemire). */
CUSMC_HD uint64_t cusmc_uint_below(uint32_t hi, uint32_t lo, uint64_t n)
{
    const uint64_t bits = ((uint64_t)hi << 32) | lo;
#if defined(__CUDA_ARCH__)
    return __umul64hi(bits, n);
#else
    return (uint64_t)(((unsigned __int128)bits * n) >> 64);
#endif
}

/* Two independent standard normals from one Philox block, double-precision Box-Muller (53-bit
 * uniforms, fp64 log / sincos).  Kept for the chi-square sampler and for callers that want
 * full-resolution draws; the step kernels use cusmc_normal4 below. */
CUSMC_HD void cusmc_normal_pair(cusmc_u32x4 r, double *z0, double *z1)
{
    const double u1 = cusmc_u01_open0(r.v[0], r.v[1]);
    const double u2 = cusmc_u01(r.v[2], r.v[3]);
    const double rad = sqrt(-2.0 * cusmc_det_log(u1));
    double s, c;
    cusmc_det_sincospi(u2 + u2, &s, &c);
    *z0 = rad * c;
    *z1 = rad * s;
}


/* Single-precision Box-Muller on two 32-bit words: the kernels' default normal generator.
 * u1 = (a + 1/2) 2^-32 in (0, 1] (the tail keeps its full 2^-32 resolution: small a are exact
 * floats), angle = (b >> 8) 2^-23 half-turns.  Deterministic: FFMA, IEEE sqrt, bit operations.
 * The draws carry 24 significant bits and reach 6.7 sigma; every operation ON the state stays fp64. */
CUSMC_HD void cusmc_box_muller_f32(uint32_t a, uint32_t b, float *z0, float *z1)
{
    float u1 = fmaf((float)a, 2.3283064365386963e-10f, 1.16415321826934814e-10f);
    if (u1 > 1.0f) u1 = 1.0f;
    const float rad = sqrtf(-2.0f * cusmc_det_logf(u1));
    float s, c;
    cusmc_det_sincospif((float)(b >> 8) * 1.1920928955078125e-7f, &s, &c);
    *z0 = rad * c;
    *z1 = rad * s;
}

#if defined(__CUDACC__)
/* The throughput generator (device only): the same Philox words through the special-function unit --
 * MUFU.LG2, MUFU.SQRT, MUFU.SIN / COS -- instead of FFMA polynomials: ~12 instead of ~67 instructions
 * per pair of normals.  Same law and resolution as cusmc_box_muller_f32 (u1 in (0, 1] on a 2^-32
 * grid, the angle on a 2^-32 grid of [-pi, pi), where the fast sine / cosine are accurate to 2^-21),
 * but its last bits follow the hardware's approximations, so a host cannot mirror it: runs that must
 * be reproduced bit for bit on a CPU ask for the reproducible generator instead
 * (cusmc_filter_config.reproducible_rng). */
__device__ __forceinline__ void cusmc_box_muller_fast(uint32_t a, uint32_t b, float *z0, float *z1)
{
    const float u1 = fmaf((float)a, 2.3283064365386963e-10f, 1.16415321826934814e-10f);   /* <= 1 after rounding */
    const float r2 = -1.3862943611198906f * __log2f(u1);                                  /* -2 ln u1 >= 0 */
    float rad;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(rad) : "f"(r2));
    float s, c;
    __sincosf((float)(int32_t)b * 1.4629180792671596e-9f, &s, &c);                        /* 2 pi 2^-32 */
    *z0 = rad * c;
    *z1 = rad * s;
}
#endif

/* Four standard normals (as doubles) from one Philox block. */
CUSMC_HD void cusmc_normal4(cusmc_u32x4 r, double z[4])
{
    float a0, a1, b0, b1;
    cusmc_box_muller_f32(r.v[0], r.v[1], &a0, &a1);
    cusmc_box_muller_f32(r.v[2], r.v[3], &b0, &b1);
    z[0] = (double)a0;
    z[1] = (double)a1;
    z[2] = (double)b0;
    z[3] = (double)b1;
}

#endif /* CUSMC_PHILOX_H */

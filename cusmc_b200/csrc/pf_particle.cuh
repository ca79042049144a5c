// pf_particle.cuh -- what ONE thread does for ONE child particle of a filter step, shared by the
// one-particle-per-thread step kernel (pf_step_impl.cuh), the fused step kernel (pf_fused_impl.cuh)
// and the persistent whole-run kernel (pf_persist.cu):
//
//     x_new = mu + G x_parent + noise          noise = Q xi | chi (.) (Q xi)
//     lw    = log pdf_V(y_t - F x_new)         (or the density, reference mode)
//
// (propagate_K + sample, src/mcmc.cpp:90-160 and src/statistics.cc.cpp:224-259,355-412; reweight_G,
// src/mcmc.cpp:162-237.)  G, Q and the whitened observation operator ride in the kernel parameter
// bank; every operation on the state is fp64 in a fixed order (oracle: orc_step_det).
#pragma once

#include "pf_step.cuh"

#include "../../include/cusmc_detmath.h"
#include "../../include/cusmc_philox.h"

namespace pfstep {

template <int D, bool DIAG>
struct StepOp {
    static constexpr int NM = DIAG ? D : D * D;
    double G[NM];      // row-major transition (DIAG: the diagonal)
    double Q[NM];      // row-major noise factor (already multiplied by noise_scale)
    double M[NM];      // row-major whitened observation operator  L_V^-1 F
    double c[D];       // L_V^-1 y_t
    double mu[D];      // additive location (m0 at t = 0, otherwise 0)
};

// chi_k = sqrt(nu / X),  X ~ chi^2_nu = 2 Gamma(nu/2)  (the reference's curand_gamma /
// curand_chi_square, src/mvt_dist.cu.cpp:20-61): Marsaglia-Tsang in SINGLE precision, like the
// normals (the draws carry 24 significant bits; every operation on the state stays fp64), built from
// the reproducible fp32 functions of cusmc_detmath.h.  One Philox block serves TWO components: its
// Box-Muller pair gives their two normals, its other two words their two uniforms; the squeeze
// u < 1 - 0.0331 z^4 accepts ~92 % of the proposals without a logarithm and a proposal is rejected
// ~4 % of the time (the component then redraws from block `attempt + 1`).  A first version drew every
// component from two blocks with fp64 Box-Muller, log and sincos: 6.5x the cost of the whole Normal
// step at d = 8.
static __device__ __noinline__ void chi_pair(uint64_t seed, uint64_t step, uint64_t index, int kpair, float nu,
                                             float *chi0, float *chi1)
{
    const float a0 = 0.5f * nu;
    const float a = a0 < 1.0f ? a0 + 1.0f : a0;
    const float dd = a - 0.333333343f;
    const float cc = 1.0f / sqrtf(9.0f * dd);
    float g[2] = {dd, dd};
    bool done[2] = {false, false};
    for (uint32_t attempt = 0; attempt < 32 && !(done[0] && done[1]); ++attempt) {
        const cusmc_u32x4 r = cusmc_rng(seed, CUSMC_STREAM_CHI, step, index, ((uint32_t)kpair << 8) | attempt);
        float z[2];
        cusmc_box_muller_f32(r.v[0], r.v[1], &z[0], &z[1]);
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            if (done[e]) continue;
            const float t = fmaf(cc, z[e], 1.0f);
            if (!(t > 0.0f)) continue;
            const float v = t * t * t;
            const float u = fmaf((float)(r.v[2 + e] >> 8), 5.9604644775390625e-8f, 2.98023223876953125e-8f);  // (0, 1)
            const float z2 = z[e] * z[e];
            if (u < fmaf(-0.0331f * z2, z2, 1.0f) ||
                cusmc_det_logf(u) < fmaf(0.5f, z2, dd * (1.0f - v + cusmc_det_logf(v)))) {
                g[e] = dd * v;
                done[e] = true;
            }
        }
    }
    if (a0 < 1.0f) {
        // shape < 1 (nu < 2): Gamma(a0) = Gamma(a0 + 1) U^(1/a0), from a block of its own
        const cusmc_u32x4 r = cusmc_rng(seed, CUSMC_STREAM_CHI, step, index, ((uint32_t)kpair << 8) | 0x800000u);
#pragma unroll
        for (int e = 0; e < 2; ++e)
            g[e] = (float)((double)g[e] * cusmc_det_exp(cusmc_det_log(cusmc_u01_open0(r.v[2 * e], r.v[2 * e + 1])) / (double)a0));
    }
    *chi0 = sqrtf(nu / (2.0f * g[0]));
    *chi1 = sqrtf(nu / (2.0f * g[1]));
}

// The throughput variant of the same draw (reproducible_rng = 0, nu >= 2): Philox4x32-7, the
// special-function-unit Box-Muller and logarithm, and NO per-pair retry loop -- a warp almost always
// holds a lane whose proposal was rejected (4 % per component), so a loop per pair makes every warp pay
// ~2 rounds per pair.  Here all D components take their first proposal in straight-line code, leaving
// a mask of the rejected ones; then each lane retries its own rejected components one at a time (each
// retry block carries two proposals), which costs a warp the MAXIMUM over its lanes of the rejected
// count -- ~1.8 short rounds per particle instead of ~4 long ones.  Same law (Marsaglia-Tsang is exact
// whichever proposals are rejected); the draws differ from chi_pair's, like the normals of the two modes.
template <int D>
__device__ __forceinline__ void chi_fast(uint64_t seed, uint64_t step, uint64_t index, float nu, float (&chi)[D])
{
    const float dd = 0.5f * nu - 0.333333343f;
    const float cc = rsqrtf(9.0f * dd);
    auto propose = [&](float z, uint32_t ubits, float &g) {
        const float t = fmaf(cc, z, 1.0f);
        const float v = t * t * t;
        const float u = fmaf((float)(ubits >> 8), 5.9604644775390625e-8f, 2.98023223876953125e-8f);  // (0, 1)
        const float z2 = z * z;
        const bool ok = t > 0.0f && (u < fmaf(-0.0331f * z2, z2, 1.0f) ||
                                     __logf(u) < fmaf(0.5f, z2, dd * (1.0f - v + __logf(v))));
        if (ok) g = dd * v;
        return ok;
    };
    float g[D];
    uint32_t rejected = 0;
#pragma unroll
    for (int kp = 0; kp < (D + 1) / 2; ++kp) {
        const cusmc_u32x4 r = cusmc_rng7(seed, CUSMC_STREAM_CHI, step, index, (uint32_t)kp << 8);
        float z0, z1;
        cusmc_box_muller_fast(r.v[0], r.v[1], &z0, &z1);
        g[2 * kp] = dd;
        if (!propose(z0, r.v[2], g[2 * kp])) rejected |= 1u << (2 * kp);
        if (2 * kp + 1 < D) {
            g[2 * kp + 1] = dd;
            if (!propose(z1, r.v[3], g[2 * kp + 1])) rejected |= 1u << (2 * kp + 1);
        }
    }
    for (uint32_t attempt = 1; rejected; ++attempt) {
        const int e = __ffs((int)rejected) - 1;
        const cusmc_u32x4 r = cusmc_rng7(seed, CUSMC_STREAM_CHI, step, index, 0x8000u | ((uint32_t)e << 8) | attempt);
        float z0, z1, ge = dd;
        cusmc_box_muller_fast(r.v[0], r.v[1], &z0, &z1);
        if (propose(z0, r.v[2], ge) || propose(z1, r.v[3], ge) || attempt >= 200u) {
#pragma unroll
            for (int k = 0; k < D; ++k)
                if (k == e) g[k] = ge;
            rejected &= rejected - 1u;
        }
    }
    const float scale = 2.0f / nu;
#pragma unroll
    for (int k = 0; k < D; ++k) chi[k] = rsqrtf(scale * g[k]);
}

// Integer nu (throughput generator): chi^2_nu = 2 Gamma(nu / 2) without rejection.  Gamma(k + odd / 2) is the sum
// of k unit exponentials and, for odd nu, half a squared normal:  g = -ln(U_1 ... U_k) + odd Z^2 / 2.  Straight-line
// code, no retries and therefore no divergence (chi_fast pays ~1.8 retry rounds per particle because a warp almost
// always holds a rejected lane, and runs the squeeze's logarithms for the same reason): nu = 5 costs 6 Philox
// blocks and 16 logarithms per d = 8 particle -- ~58 instead of ~112 instructions per factor.  One block serves
// four components; k is uniform over the grid.  Used for integer nu <= kHalfIntMaxNu: its cost grows with nu and
// Marsaglia-Tsang's does not -- measured per d = 8 C5 step (chi_fast: 470-495 us at every nu): nu = 3: 374 us,
// 5: 412, 8: 450, 12: 536.  Same law as chi_fast / chi_pair (KS against t_nu, 10^6 draws x 6 seeds: p-values
// uniform, profiles/mvt_ks.py), other draws.
constexpr int kHalfIntMaxNu = 8;
template <int D>
__device__ __forceinline__ void chi_halfint(uint64_t seed, uint64_t step, uint64_t index, int nu_int, float (&chi)[D])
{
    const int k = nu_int >> 1;
    float acc[D];                       // sum of log2 U
#pragma unroll
    for (int c = 0; c < D; ++c) acc[c] = 0.0f;
    for (int i = 0; i < k; ++i) {
#pragma unroll
        for (int q = 0; q < (D + 3) / 4; ++q) {
            const cusmc_u32x4 r = cusmc_rng7(seed, CUSMC_STREAM_CHI, step, index, 0x4000u | ((uint32_t)q << 8) | (uint32_t)i);
#pragma unroll
            for (int e = 0; e < 4; ++e)
                if (4 * q + e < D)      // U = (word + 1/2) 2^-32 in (0, 1]
                    acc[4 * q + e] += __log2f(fmaf((float)r.v[e], 2.3283064365386963e-10f, 1.16415321826934814e-10f));
        }
    }
    float g[D];
#pragma unroll
    for (int c = 0; c < D; ++c) g[c] = -0.6931471805599453f * acc[c];
    if (nu_int & 1) {
#pragma unroll
        for (int q = 0; q < (D + 3) / 4; ++q) {
            const cusmc_u32x4 r = cusmc_rng7(seed, CUSMC_STREAM_CHI, step, index, 0x6000u | ((uint32_t)q << 8));
            float z[4];
            cusmc_box_muller_fast(r.v[0], r.v[1], &z[0], &z[1]);
            cusmc_box_muller_fast(r.v[2], r.v[3], &z[2], &z[3]);
#pragma unroll
            for (int e = 0; e < 4; ++e)
                if (4 * q + e < D) g[4 * q + e] = fmaf(0.5f * z[e], z[e], g[4 * q + e]);
        }
    }
    const float scale = 2.0f / (float)nu_int;
#pragma unroll
    for (int c = 0; c < D; ++c) chi[c] = rsqrtf(fmaxf(scale * g[c], 1e-37f));      // g = 0 has probability ~2^-32 per draw
}

// The block of (seed, stream, step, particle, quad): Philox4x32-10, or -7 on the throughput path.
template <bool FAST>
__device__ __forceinline__ cusmc_u32x4 step_rng(uint64_t seed, int stream, uint64_t step, uint64_t index, uint32_t sub)
{
    // (measured: 252 -> 245 us per C5 step, 27.1 -> 25.9 us per C4 step)
    if constexpr (FAST) return cusmc_rng7(seed, stream, step, index, sub);
    return cusmc_rng(seed, stream, step, index, sub);
}

// Throughput noise at D <= 2: a Philox block makes four normals and a particle needs two, so the
// neighbours 2p and 2p + 1 share the block of p (first pair | second pair): half the generator work of a
// d = 2 step.  Keyed by the GLOBAL slot, so shards and the persistent kernel draw the same normals.
#ifndef CUSMC_CHI_HALFINT
#define CUSMC_CHI_HALFINT 1
#endif
#ifndef CUSMC_PAIR_D2
#define CUSMC_PAIR_D2 1
#endif
template <bool FAST, int D>
struct PairBlocks {
    static constexpr bool value = FAST && D <= 2 && CUSMC_PAIR_D2 != 0;
};
// the child's first block (what callers compute while the parent index is still in flight)
template <bool FAST, int D>
__device__ __forceinline__ cusmc_u32x4 first_block(uint64_t seed, int stream, uint64_t step, uint64_t index)
{
    if constexpr (PairBlocks<FAST, D>::value) return cusmc_rng7(seed, stream, step, index >> 1, 0u);
    return step_rng<FAST>(seed, stream, step, index, 0u);
}

// One Philox block -> two Box-Muller pairs.  FAST: the special-function-unit transform
// (cusmc_box_muller_fast; the throughput default), otherwise the FFMA-only one a host reproduces.
template <bool FAST>
__device__ __forceinline__ void normals4(const cusmc_u32x4 &r, float (&z)[4])
{
    if constexpr (FAST) {
        cusmc_box_muller_fast(r.v[0], r.v[1], &z[0], &z[1]);
        cusmc_box_muller_fast(r.v[2], r.v[3], &z[2], &z[3]);
    } else {
        cusmc_box_muller_f32(r.v[0], r.v[1], &z[0], &z[1]);
        cusmc_box_muller_f32(r.v[2], r.v[3], &z[2], &z[3]);
    }
}

// The child `i` (local column; global slot a.i0 + i) of parent column `src` (already resolved to the
// owning rank's buffer; nullptr-free: callers pass has_prev = 0 for the initial draw).  r0 is the
// child's first Philox block (first_block()), computed by the caller while the parent index was still in flight.
// Stores x_new (and the history row) itself; returns the (log-)weight, not yet stored.
// cobs: the whitened observation L_V^-1 y_t -- op.c in the per-step kernels (a parameter-bank operand),
// a register array in the persistent kernel, whose observation changes inside the launch.  COH: the
// parent state was written earlier in the SAME launch by other blocks (persistent kernel): read it
// through L2, never through the non-coherent path.
// xp_in (optional): the parent state already gathered into registers by the caller (the persistent
// kernel issues a batch of gathers before it computes the batch: its rounds are latency-bound).
// zf_in (optional, PHILOX): the child's D normals already drawn (the persistent kernel draws the next
// step's normals while its block waits at the grid barrier).
// LEAN: the persistent kernel's main loop -- there is a parent, there is no history row, the weight is
// the Normal log-density: the run-time switches of the general step (has_prev, hist_x, skip_weight, the
// epilogue's kind / log branches) are compiled out.  Same arithmetic.
// SMOP (dense operators only): G, Q and M are read from a shared-memory copy `smop` = [G | Q | M], row-major
// D x D each, 16-byte aligned -- one broadcast LDS.128 per two DFMAs.  As parameter-bank operands the
// compiler hoists them out of the caller's loop over rounds into registers, spills those (d = 8: 344 bytes
// of stack per thread = 230 MB of extra DRAM writes per C5 step) and moves them back with R2UR before every
// use: 17 % of the dense kernel's stall samples sat on those reloads.  Same values, same order of operations.
template <int D, bool PHILOX, bool FAST, bool MVT, bool EXACT, bool DIAG, bool COH = false, bool LEAN = false,
          bool SMOP = false>
__device__ __forceinline__ double particle_step(const StepOp<D, DIAG> &op, const double (&cobs)[D], const Epilogue &ep,
                                                const StepArgs &a, int64_t i, const double *__restrict__ src,
                                                const cusmc_u32x4 &r0, const double *xp_in = nullptr,
                                                const float *zf_in = nullptr, const double *smop = nullptr)
{
    static_assert(!SMOP || (!DIAG && D % 2 == 0), "shared-memory operators: the dense step");
    // element `idx` of matrix `which` (0 = G, 1 = Q, 2 = M), two at a time from shared memory
    auto mat2 = [&](int which, int idx) {
        if constexpr (SMOP) {
            return reinterpret_cast<const double2 *>(smop)[(which * D * D + idx) >> 1];
        } else {
            const double *m = which == 0 ? op.G : which == 1 ? op.Q : op.M;
            return make_double2(m[DIAG ? 0 : idx], m[DIAG ? 0 : idx + 1]);
        }
    };
    const int d = EXACT ? D : a.d;
    const uint64_t idx = (uint64_t)(a.i0 + i);
    double xp[D], z[D], xn[D];
    constexpr bool kGathered = LEAN && COH && D <= 4;      // the persistent kernel's batched gathers: always
    if (kGathered || xp_in) {
#pragma unroll
        for (int j = 0; j < D; ++j) xp[j] = xp_in[j];
    } else if (LEAN || a.has_prev) {
#pragma unroll
        for (int j = 0; j < D; ++j)
            xp[j] = (EXACT || j < d) ? (COH ? __ldcg(src + (int64_t)j * a.ld_prev) : __ldg(src + (int64_t)j * a.ld_prev)) : 0.0;
    } else {
#pragma unroll
        for (int j = 0; j < D; ++j) xp[j] = 0.0;
    }
    // DIAG: component k needs only its own normal, so the draws stay in single precision (half the
    // registers) until the one FMA that consumes them
    constexpr bool kFloatNoise = PHILOX && DIAG;
    float zf[kFloatNoise ? D : 1];
    if (PHILOX && zf_in) {
#pragma unroll
        for (int k = 0; k < D; ++k) {
            if constexpr (kFloatNoise) zf[k] = zf_in[k];
            else z[k] = (double)zf_in[k];
        }
    } else if (PHILOX) {
        // one Philox block -> four single-precision Box-Muller normals (cusmc_philox.h)
#pragma unroll
        for (int jq = 0; jq < (D + 3) / 4; ++jq) {
            float zq[4] = {0.0f, 0.0f, 0.0f, 0.0f};
            if constexpr (PairBlocks<FAST, D>::value) {
                const bool odd = (idx & 1u) != 0;             // r0 = first_block(): the block of the pair idx >> 1
                cusmc_box_muller_fast(odd ? r0.v[2] : r0.v[0], odd ? r0.v[3] : r0.v[1], &zq[0], &zq[1]);
            } else if (EXACT || 4 * jq < d) {
                const cusmc_u32x4 rq = jq == 0 ? r0 : step_rng<FAST>(a.seed, a.rng_stream, a.step, idx, (uint32_t)jq);
                normals4<FAST>(rq, zq);
            }
#pragma unroll
            for (int e = 0; e < 4; ++e)
                if (4 * jq + e < D) {
                    const float v = (EXACT || 4 * jq + e < d) ? zq[e] : 0.0f;
                    if constexpr (kFloatNoise) zf[4 * jq + e] = v;
                    else z[4 * jq + e] = (double)v;
                }
        }
    } else {
        const double *nz = a.xi + i;
#pragma unroll
        for (int j = 0; j < D; ++j) z[j] = (EXACT || j < d) ? ld_stream(nz + (int64_t)j * a.ld_noise) : 0.0;
    }
    double chi[MVT ? D : 1];
    bool chi_drawn = false;
    if constexpr (MVT && FAST) {
        const int nu_int = (int)a.nu;
        if (!a.chi && CUSMC_CHI_HALFINT != 0 && (float)nu_int == a.nu && nu_int >= 1 && nu_int <= kHalfIntMaxNu) {
            float cf[D];
            chi_halfint<D>(a.seed, a.step, idx, nu_int, cf);
#pragma unroll
            for (int k = 0; k < D; ++k) chi[MVT ? k : 0] = (double)cf[k];
            chi_drawn = true;
        } else if (!a.chi && a.nu >= 2.0f) {
            float cf[D];
            chi_fast<D>(a.seed, a.step, idx, a.nu, cf);
#pragma unroll
            for (int k = 0; k < D; ++k) chi[MVT ? k : 0] = (double)cf[k];
            chi_drawn = true;
        }
    }
    if (MVT && !a.chi && !chi_drawn) {
#pragma unroll
        for (int kp = 0; kp < (D + 1) / 2; ++kp) {
            float c0 = 1.0f, c1 = 1.0f;
            if (EXACT || 2 * kp < d) chi_pair(a.seed, a.step, idx, kp, a.nu, &c0, &c1);
            chi[MVT ? 2 * kp : 0] = (double)c0;
            if (2 * kp + 1 < D) chi[MVT ? 2 * kp + 1 : 0] = (double)c1;
        }
    }
    double *dst_x = a.x_new + i;
#pragma unroll
    for (int k = 0; k < D; ++k) {
        double g = op.mu[k], s = 0.0;
        if constexpr (DIAG) {
            g = fma(op.G[k], xp[k], g);
            s = fma(op.Q[k], kFloatNoise ? (double)zf[kFloatNoise ? k : 0] : z[k], s);
        } else {
#pragma unroll
            for (int j = 0; j < D; j += 2) {
                const double2 m2 = mat2(0, k * D + j);
                g = fma(m2.x, xp[j], g);
                g = fma(m2.y, xp[j + 1], g);
            }
#pragma unroll
            for (int j = 0; j < D; j += 2) {
                const double2 m2 = mat2(1, k * D + j);
                s = fma(m2.x, z[j], s);
                s = fma(m2.y, z[j + 1], s);
            }
        }
        if (MVT && (EXACT || k < d))
            s = (a.chi ? ld_stream(a.chi + (int64_t)k * a.ld_noise + i) : chi[MVT ? k : 0]) * s;
        xn[k] = s + g;
        if (EXACT || k < d) {
            st_stream(dst_x + (int64_t)k * a.ld_new, xn[k]);
            // the history row goes out here too: stores keep their order, and one issued after the
            // weight would pin every xn[k] in a register until the end of the kernel
            if (!LEAN && a.hist_x) st_stream(a.hist_x + i * d + k, xn[k]);
        }
    }
    if (!LEAN && a.skip_weight) return a.const_weight;
    double q = 0.0;
#pragma unroll
    for (int k = 0; k < D; ++k) {
        double zk = cobs[k];
        if constexpr (DIAG) {
            zk = fma(-op.M[k], xn[k], zk);
        } else {
#pragma unroll
            for (int j = 0; j < D; j += 2) {
                const double2 m2 = mat2(2, k * D + j);
                zk = fma(-m2.x, xn[j], zk);
                zk = fma(-m2.y, xn[j + 1], zk);
            }
        }
        q = fma(zk, zk, q);
    }
    if constexpr (LEAN) return fma(-0.5, q, ep.lognorm);      // density_epilogue's MVN / log branch
    return density_epilogue(ep, q);
}

// Host-side view of the model matrices of one step (all optional, column-major like Eigen).
struct StepModel {
    int d, dy;
    const double *G, *Q;
    double qscale;
    const std::vector<double> *M;   // row-major dy x d
    const double *c, *mu;
};

template <int D, bool DIAG>
void fill_step_op(StepOp<D, DIAG> &op, const StepModel &m)
{
    std::memset(&op, 0, sizeof(op));
    const int d = m.d, dy = m.dy;
    if constexpr (DIAG) {
        for (int k = 0; k < d; ++k) {
            if (m.G) op.G[k] = m.G[(size_t)k * d + k];
            if (m.Q) op.Q[k] = m.Q[(size_t)k * d + k] * m.qscale;
            if (m.M) op.M[k] = (*m.M)[(size_t)k * d + k];
        }
    } else {
        for (int k = 0; k < d; ++k)
            for (int j = 0; j < d; ++j) {
                if (m.G) op.G[k * D + j] = m.G[(size_t)j * d + k];
                if (m.Q) op.Q[k * D + j] = m.Q[(size_t)j * d + k] * m.qscale;
            }
        if (m.M)
            for (int k = 0; k < dy; ++k)
                for (int j = 0; j < d; ++j) op.M[k * D + j] = (*m.M)[(size_t)k * d + j];
    }
    for (int k = 0; k < dy; ++k) op.c[k] = m.c ? m.c[k] : 0.0;
    for (int k = 0; k < d; ++k) op.mu[k] = m.mu ? m.mu[k] : 0.0;
}

}  // namespace pfstep

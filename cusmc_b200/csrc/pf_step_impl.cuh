// pf_step_impl.cuh -- the particle-filter step kernel (included by pf_step_mvn.cu / pf_step_mvt.cu,
// one translation unit per noise family so the 60 instantiations compile in parallel).
//
// One fused kernel per step replaces the reference's eight (Gmu, sample, y_minus_Fmu,
// Einv_alpha, pdf + RNG set-up kernels; src/mvn_dist.cu.cpp:33-172,455-668) and the host-side
// ancestor gather / AoS flattening / H2D+D2H of the whole particle cloud every step
// (src/mvn_dist.cu.cpp:194-205,231-251,300-302): a thread owns one child particle,
//
//     parent  = anc[i]
//     x_new   = mu + G x_prev[:, parent] + noise_i          noise = Q xi | chi (.) (Q xi)
//     lw[i]   = log pdf_V(y_t - F x_new)   (or the density, reference mode)
//     max     = atomic max over lw (for the max-shifted normalisation)
//
// State is SoA and stays on the device for the whole run; G, Q and the whitened observation
// operator ride in the kernel parameter bank.
//
// The kernel is ISSUE-bound, not HBM-bound (ncu r01: 975-1080 warp instructions per particle at
// d = 8, issue slots 84 % busy): the counter-based normals cost ~50 instructions each, the three
// dense d x d products 3 d^2 DFMA.  Two compile-time specialisations cut what can be cut:
//   EXACT : d == dy == D, every `k < d` predicate folds away;
//   DIAG  : G, Q and M = L_V^-1 F are all diagonal (random-walk / AR(1) models observed
//           component-wise -- both filter configurations of BASELINE.json), the products shrink
//           from 3 d^2 to 3 d DFMA.  Skipping the zero terms does not change a single bit:
//           fma(0, x, acc) == acc for finite x.
#pragma once

#include "pf_step.cuh"

#include "../../include/cusmc_detmath.h"
#include "../../include/cusmc_philox.h"

namespace pfstep {

constexpr int kThreads = 256;

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

#ifndef CUSMC_STEP_MINB8
#define CUSMC_STEP_MINB8 5
#endif
// d = 8, diagonal model, Normal noise (the C5 configuration): 48 registers hold the kernel without a
// spill, 5 blocks per SM instead of 4 took the step from 212 to 203 us; at 6 blocks (40 registers) it
// spills its store addresses and is back at 212 us
constexpr int min_blocks(int D, bool diag, bool mvt)
{
    return D >= 32 ? (diag ? 2 : 1) : (D >= 16 ? 2 : (D >= 8 ? (diag && !mvt ? CUSMC_STEP_MINB8 : 3) : 4));
}

// MVT is a template flag so the MVN kernel carries neither the chi branch nor the call to the
// (rejection-loop) chi-square sampler, whose calling convention alone costs ~30 registers.
template <int D, bool PHILOX, bool MVT, bool EXACT, bool DIAG>
__global__ void __launch_bounds__(kThreads, min_blocks(D, DIAG, MVT))
pf_step_kernel(const __grid_constant__ StepOp<D, DIAG> op, const Epilogue ep, const StepArgs a)
{
    const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    const bool active = i < a.n_out;
    const int d = EXACT ? D : a.d;
    double lw = -INFINITY;
    if (active) {
        double xp[D], z[D], xn[D];
        // Order matters for latency: the ancestor load is issued first, the first Philox block
        // (pure integer work) runs while it is in flight, then the parent gather goes out and the
        // Box-Muller transforms and the remaining blocks run under ITS latency.
        int64_t parent = i;
        if (a.anc) parent = (int64_t)a.anc[i] - a.parent_base;
        double *dst_x = a.x_new + i, *dst_lw = a.lw + i;
        const int64_t dst_ld = a.ld_new;
        const uint64_t idx = (uint64_t)(a.i0 + i);
        cusmc_u32x4 r0;
        if (PHILOX) r0 = cusmc_rng(a.seed, a.rng_stream, a.step, idx, 0u);
        if (a.has_prev) {
            const double *src = a.x_prev + parent;
            if (a.world > 1) {
                // ancestors are (nearly) sorted, so almost every parent is local: only a remote one
                // pays the dependent load of its owner's pointer from the table in device memory
                const uint32_t g = (uint32_t)parent, r = fast_div(g, a.per_rank);
                const uint32_t col = g - r * a.per_rank.d;
                src = (r == (uint32_t)a.rank ? a.x_prev : a.x_prev_peer[r]) + col;
            }
#pragma unroll
            for (int j = 0; j < D; ++j) xp[j] = (EXACT || j < d) ? __ldg(src + (int64_t)j * a.ld_prev) : 0.0;
        } else {
#pragma unroll
            for (int j = 0; j < D; ++j) xp[j] = 0.0;
        }
        // DIAG: component k needs only its own normal, so the draws stay in single precision (half the
        // registers) until the one FMA that consumes them
        constexpr bool kFloatNoise = PHILOX && DIAG;
        float zf[kFloatNoise ? D : 1];
        if (PHILOX) {
            // one Philox block -> four single-precision Box-Muller normals (cusmc_philox.h)
#pragma unroll
            for (int jq = 0; jq < (D + 3) / 4; ++jq) {
                float zq[4] = {0.0f, 0.0f, 0.0f, 0.0f};
                if (EXACT || 4 * jq < d) {
                    const cusmc_u32x4 rq = jq == 0 ? r0 : cusmc_rng(a.seed, a.rng_stream, a.step, idx, (uint32_t)jq);
                    cusmc_box_muller_f32(rq.v[0], rq.v[1], &zq[0], &zq[1]);
                    cusmc_box_muller_f32(rq.v[2], rq.v[3], &zq[2], &zq[3]);
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
            const double *src = a.xi + i;
#pragma unroll
            for (int j = 0; j < D; ++j) z[j] = (EXACT || j < d) ? ld_stream(src + (int64_t)j * a.ld_noise) : 0.0;
        }
        double chi[MVT ? D : 1];
        if (MVT && !a.chi) {
#pragma unroll
            for (int kp = 0; kp < (D + 1) / 2; ++kp) {
                float c0 = 1.0f, c1 = 1.0f;
                if (EXACT || 2 * kp < d) chi_pair(a.seed, a.step, (uint64_t)(a.i0 + i), kp, a.nu, &c0, &c1);
                chi[MVT ? 2 * kp : 0] = (double)c0;
                if (2 * kp + 1 < D) chi[MVT ? 2 * kp + 1 : 0] = (double)c1;
            }
        }
#pragma unroll
        for (int k = 0; k < D; ++k) {
            double g = op.mu[k], s = 0.0;
            if constexpr (DIAG) {
                g = fma(op.G[k], xp[k], g);
                s = fma(op.Q[k], kFloatNoise ? (double)zf[kFloatNoise ? k : 0] : z[k], s);
            } else {
#pragma unroll
                for (int j = 0; j < D; ++j) g = fma(op.G[k * D + j], xp[j], g);
#pragma unroll
                for (int j = 0; j < D; ++j) s = fma(op.Q[k * D + j], z[j], s);
            }
            if (MVT && (EXACT || k < d))
                s = (a.chi ? ld_stream(a.chi + (int64_t)k * a.ld_noise + i) : chi[MVT ? k : 0]) * s;
            xn[k] = s + g;
            if (EXACT || k < d) {
                st_stream(dst_x + (int64_t)k * dst_ld, xn[k]);
                // the history row goes out here too: stores keep their order, and one issued after the
                // weight would pin every xn[k] in a register until the end of the kernel
                if (a.hist_x) st_stream(a.hist_x + i * d + k, xn[k]);
            }
        }
        if (a.skip_weight) {
            lw = a.const_weight;
        } else {
            double q = 0.0;
#pragma unroll
            for (int k = 0; k < D; ++k) {
                double zk = op.c[k];
                if constexpr (DIAG) {
                    zk = fma(-op.M[k], xn[k], zk);
                } else {
#pragma unroll
                    for (int j = 0; j < D; ++j) zk = fma(-op.M[k * D + j], xn[j], zk);
                }
                q = fma(zk, zk, q);
            }
            lw = density_epilogue(ep, q);
            if (a.resampled && *a.resampled == 0) lw = *dst_lw + lw;   // no resampling: weights accumulate
        }
        st_stream(dst_lw, lw);
        if (a.hist_w) st_stream(a.hist_w + i, lw);
        if (a.hist_a) a.hist_a[i] = (uint32_t)(parent + a.parent_base);
    }
    if (a.lw_max) {
        double m = (lw == lw && lw < INFINITY) ? lw : -INFINITY;
        m = warp_max_double(m);
        __shared__ double sm[kThreads / 32];
        if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = m;
        __syncthreads();
        if (threadIdx.x < 32) {
            m = warp_max_double(threadIdx.x < kThreads / 32 ? sm[threadIdx.x] : -INFINITY);
            if (threadIdx.x == 0) atomic_max_double(a.lw_max, m);
        }
    }
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

template <int D, bool MVT, bool EXACT, bool DIAG>
int launch_one(cusmc_ctx *ctx, const StepModel &m, const Epilogue &ep, const StepArgs &a, bool philox)
{
    StepOp<D, DIAG> op;
    fill_step_op<D, DIAG>(op, m);
    const unsigned grid = (unsigned)((a.n_out + kThreads - 1) / kThreads);
    if (philox)
        pf_step_kernel<D, true, MVT, EXACT, DIAG><<<grid, kThreads, 0, ctx->stream>>>(op, ep, a);
    else
        pf_step_kernel<D, false, MVT, EXACT, DIAG><<<grid, kThreads, 0, ctx->stream>>>(op, ep, a);
    CUSMC_LAUNCHED(ctx);
    return CUSMC_OK;
}

template <bool MVT>
int launch_family(cusmc_ctx *ctx, const StepModel &m, const Epilogue &ep, const StepArgs &a, bool philox,
                  bool exact, bool diag)
{
    const int dm = m.d > m.dy ? m.d : m.dy;
#define CUSMC_STEP_CASE(DD)                                                                          \
    case DD:                                                                                         \
        if (exact && diag) return launch_one<DD, MVT, true, true>(ctx, m, ep, a, philox);            \
        if (exact) return launch_one<DD, MVT, true, false>(ctx, m, ep, a, philox);                   \
        return launch_one<DD, MVT, false, false>(ctx, m, ep, a, philox);
    switch (cusmc_pad_dim(dm)) {
        CUSMC_STEP_CASE(2)
        CUSMC_STEP_CASE(4)
        CUSMC_STEP_CASE(8)
        CUSMC_STEP_CASE(16)
        CUSMC_STEP_CASE(32)
    }
#undef CUSMC_STEP_CASE
    return cusmc_fail(ctx, CUSMC_ERR_UNSUPPORTED, "unreachable");
}

int launch_mvn(cusmc_ctx *ctx, const StepModel &m, const Epilogue &ep, const StepArgs &a, bool philox,
               bool exact, bool diag);
int launch_mvt(cusmc_ctx *ctx, const StepModel &m, const Epilogue &ep, const StepArgs &a, bool philox,
               bool exact, bool diag);

}  // namespace pfstep

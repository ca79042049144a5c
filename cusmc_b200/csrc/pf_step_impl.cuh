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

#include "pf_particle.cuh"

namespace pfstep {

constexpr int kThreads = 256;

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
template <int D, bool PHILOX, bool FAST, bool MVT, bool EXACT, bool DIAG>
__global__ void __launch_bounds__(kThreads, min_blocks(D, DIAG, MVT))
pf_step_kernel(const __grid_constant__ StepOp<D, DIAG> op, const Epilogue ep, const StepArgs a)
{
    cusmc_pdl_enter();
    const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    const bool active = i < a.n_out;
    double lw = -INFINITY;
    if (active) {
        // Order matters for latency: the ancestor load is issued first, the first Philox block
        // (pure integer work) runs while it is in flight, then the parent gather goes out and the
        // Box-Muller transforms and the remaining blocks run under ITS latency.
        int64_t parent = i;
        if (a.anc) parent = (int64_t)a.anc[i] - a.parent_base;
        cusmc_u32x4 r0{};
        if (PHILOX) r0 = pfstep::first_block<FAST, D>(a.seed, a.rng_stream, a.step, (uint64_t)(a.i0 + i));
        const double *src = a.x_prev + parent;
        if (a.has_prev && a.world > 1) {
            // ancestors are (nearly) sorted, so almost every parent is local: only a remote one
            // pays the dependent load of its owner's pointer from the table in device memory
            const uint32_t g = (uint32_t)parent, r = fast_div(g, a.per_rank);
            const uint32_t col = g - r * a.per_rank.d;
            src = (r == (uint32_t)a.rank ? a.x_prev : a.x_prev_peer[r]) + col;
        }
        lw = particle_step<D, PHILOX, FAST, MVT, EXACT, DIAG>(op, op.c, ep, a, i, src, r0);
        double *dst_lw = a.lw + i;
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

template <int D, bool MVT, bool EXACT, bool DIAG>
int launch_one(cusmc_ctx *ctx, const StepModel &m, const Epilogue &ep, const StepArgs &a, bool philox)
{
    StepOp<D, DIAG> op;
    fill_step_op<D, DIAG>(op, m);
    const unsigned grid = (unsigned)((a.n_out + kThreads - 1) / kThreads);
    if (philox && a.fast_noise)
        CUSMC_CUDA(ctx, cusmc_launch_pdl(pf_step_kernel<D, true, true, MVT, EXACT, DIAG>, grid, kThreads, 0, ctx->stream, op, ep, a));
    else if (philox)
        CUSMC_CUDA(ctx, cusmc_launch_pdl(pf_step_kernel<D, true, false, MVT, EXACT, DIAG>, grid, kThreads, 0, ctx->stream, op, ep, a));
    else
        CUSMC_CUDA(ctx, cusmc_launch_pdl(pf_step_kernel<D, false, false, MVT, EXACT, DIAG>, grid, kThreads, 0, ctx->stream, op, ep, a));
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

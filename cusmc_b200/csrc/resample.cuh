// resample.cuh -- internal launchers shared between resample.cu and filter.cu.
#pragma once

#include "common.cuh"

int cusmc_fill_double(cusmc_ctx *ctx, double *p, double v, int n);
int cusmc_launch_metropolis(cusmc_ctx *ctx, uint32_t *a, const double *w, const double *u,
                            const uint32_t *j, uint64_t seed, uint64_t step, int64_t N, int B,
                            int is_log, int64_t i0, int64_t n_out, const CusmcPeers *peers, bool c2 = false);
int cusmc_launch_weights_max(cusmc_ctx *ctx, const double *w, int64_t N, double *max_dev);
int cusmc_launch_rejection(cusmc_ctx *ctx, uint32_t *a, const double *w, const double *wmax_dev, uint64_t seed,
                           uint64_t step, int64_t N, int cap);
// Building-block weight image (global-max form; the filter itself lives on the block-relative image of
// image.cuh): cusmc_scan_state_bytes(N) bytes; receives the exclusive tile prefixes + tile-local CDF the
// scan / scatter pass consumes.  stats_dev may be NULL.
int cusmc_launch_weights_sum(cusmc_ctx *ctx, const double *w, int is_log, const double *max_dev,
                             int64_t N, int shift, uint64_t *stats_dev, void *image, bool full_stats);
size_t cusmc_scan_state_bytes(int64_t N);
int cusmc_launch_scan(cusmc_ctx *ctx, int64_t N, int64_t N_global, const uint64_t *total_dev,
                      const uint64_t *cdf_offset_dev, const void *image,
                      uint64_t *cdf_out, uint32_t *anc_out, int64_t j0, int64_t out_lo,
                      int64_t out_n, double u0);
int cusmc_launch_multinomial(cusmc_ctx *ctx, const uint64_t *cdf, int64_t N, const uint64_t *total_dev,
                             const double *u, uint64_t seed, uint64_t step, int64_t i0,
                             int64_t n_out, int64_t j0, uint32_t *a, uint64_t *degenerate_dev = nullptr);

// ---- shared with the persistent filter kernel (pf_persist.cu) -------------------------------------
constexpr int kResampleThreads = 256;
#ifndef CUSMC_TILE_ITEMS
#define CUSMC_TILE_ITEMS 8
#endif
constexpr int kTileItems = CUSMC_TILE_ITEMS;
constexpr int kTile = kResampleThreads * kTileItems;    // 2048 weights per tile


// Per-launch constants of the systematic offspring scatter (64 bytes at the head of the weight
// image).  Computed ONCE per step by one thread -- the tail of tile_scan_kernel or, for the building-
// block entry points, scatter_consts_kernel -- instead of once per block of the resampling pass: the
// dependent chain load(total) -> two fp64 divisions -> barrier was on every block's critical path.
struct ScatterConsts {
    unsigned long long T;        // global fixed-point mass
    unsigned long long r0;       // systematic offset in mass units: min(trunc(u0 T), T - 1)
    double ng_over_t, r0_over_t; // N_global / T, r0 / T (the floating estimate of offspring_below)
    unsigned long long resample; // adaptive resampling decision (1 = draw new ancestors)
    unsigned long long pad[3];
};
constexpr int kImageHead = sizeof(ScatterConsts) / 8;   // image words before the tile sums

#ifdef __CUDACC__
__device__ __forceinline__ ScatterConsts make_scatter_consts(unsigned long long Tt, double u0, uint32_t N_global,
                                                            double ess_bound, unsigned long long sum_q2)
{
    ScatterConsts c{};
    unsigned long long rr = (unsigned long long)(u0 * (double)Tt);
    if (Tt && rr > Tt - 1) rr = Tt - 1;
    c.T = Tt;
    c.r0 = rr;
    c.ng_over_t = (double)N_global / (double)Tt;
    c.r0_over_t = (double)rr / (double)Tt;
    // ESS = sum_q^2 / (sum_q2 2^shift) < threshold N  <=>  sum_q^2 < ess_bound sum_q2
    c.resample = 1;
    if (ess_bound > 0.0) c.resample = (double)Tt * (double)Tt < ess_bound * (double)sum_q2;
    return c;
}

// #{ i in [0, Ng) : i*T + r0 < C*Ng }  =  smallest k with k*T + r0 >= C*Ng, clamped to Ng.
// Floating-point estimate, then (rarely) an exact 128-bit correction.
__device__ __forceinline__ uint64_t offspring_below(uint64_t C, uint64_t Ng, uint64_t T, uint64_t r0,
                                                    double ng_over_t, double r0_over_t)
{
    // p = (C*Ng - r0) / T > -1; the answer is max(0, ceil(p)).  est is within 2^-19 of p (C, Ng/T and
    // r0/T each carry one rounding and p <= 2^32), so whenever est sits safely inside an open unit
    // interval (n, n + 1) the answer is n + 1 -- also for n = -1 -- and no 128-bit arithmetic is
    // needed.  With t = est - 1/2 that interval test is |t - rint(t)| < 1/2 - 1e-4 and n = rint(t):
    // one rounding, one subtraction, one compare.  Only boundaries within 1e-4 of an integer (2e-4 of
    // all cases) take the exact path below.
    const double est = fma((double)C, ng_over_t, -r0_over_t);
    const double t = est - 0.5;
    const double n = rint(t);
    if (fabs(t - n) < 0.5 - 1e-4) {
        const uint32_t kf = __double2uint_rz(n + 1.0);       // saturates at 2^32 - 1 >= Ng
        return kf > (uint32_t)Ng ? Ng : (uint64_t)kf;
    }
    const uint64_t rhs_lo = C * Ng, rhs_hi = __umul64hi(C, Ng);
    uint64_t k = est <= 0.0 ? 0 : (est >= (double)Ng ? Ng : (uint64_t)est);
    // lhs(k) = k*T + r0 as 128 bit
    auto lhs_less = [&](uint64_t kk) {
        uint64_t lo = kk * T, hi = __umul64hi(kk, T);
        const uint64_t lo2 = lo + r0;
        hi += lo2 < lo;
        return hi < rhs_hi || (hi == rhs_hi && lo2 < rhs_lo);
    };
    while (k < Ng && lhs_less(k)) ++k;
    while (k > 0 && !lhs_less(k - 1)) --k;
    return k;
}

#endif

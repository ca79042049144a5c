// pf_step.cuh -- arguments of the fused propagate + reweight kernel (pf_step.cu), shared with the
// host loops in filter.cu.
#pragma once

#include "density.cuh"

#include <vector>

struct StepArgs {
    double *x_new;
    const double *x_prev;
    const uint32_t *anc;
    const double *xi;
    const double *chi;
    double *lw;
    double *lw_max;              // optional
    // optional history rows of this step, written by the same threads (no extra launches):
    // hist_x [n][d] AoS, hist_w [n], hist_a [n] (global parent ids)
    double *hist_x, *hist_w;
    uint32_t *hist_a;
    int64_t n_out, ld_new, ld_prev, ld_noise;
    int64_t i0;                  // global index of child 0 (keys the counter-based draws)
    int64_t parent_base;         // global index of x_prev column 0
    uint64_t seed, step;
    double const_weight;         // used when skip_weight
    float nu;
    int d, dy, kind, has_prev, skip_weight, rng_stream;
    int fast_noise;              // device-drawn normals through the special-function unit (not host-reproducible)
    // peer form (world > 1): anc holds GLOBAL parent ids, parent g lives on rank g / per_rank at
    // column g % per_rank of that rank's state buffer (leading dimension ld_prev on every rank),
    // read through the peer-mapped pointer -- an 8d-byte gather over NVLink when it is remote.
    const double *const *x_prev_peer;   // device table of the ranks' state buffers
    FastDiv per_rank;
    int world, rank;             // rank: parents on this rank are read through x_prev directly
};

// G, Q column-major d x d host (either may be NULL = zero); M row-major dy x d (NULL = no
// weights); c dy; mu d (NULL = 0).  Picks the EXACT (d == dy == padded D) and DIAG (G, Q, M all
// diagonal) specialisations itself.
int cusmc_launch_step(cusmc_ctx *ctx, int d, int dy, const double *G, const double *Q, double qscale,
                      const std::vector<double> *M, const double *c, const double *mu, const Epilogue &ep,
                      const StepArgs &a, bool philox);

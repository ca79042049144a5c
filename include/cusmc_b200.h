/*
 * cusmc_b200.h -- C ABI of the B200-native CuSMC sampling hot path.
 *
 * This is the drop-in boundary: the five C++ wrappers the reference declares in
 * inst/include/distributions/mvn_dist.hpp:19-54 (reached through the virtuals
 * pdf_cu / sample_cu / sample_cu_init, inst/include/statistics.hpp:51-94) and the
 * resampler seam resampler_f (inst/include/types.hpp:32) are replaced by the
 * extern "C" entry points below.  Plain pointers and sizes only; no C++ or torch
 * types cross this line.  INTEGRATION.md shows the Rcpp-side binding.
 *
 * Conventions
 *   - All arithmetic is fp64; ancestor indices are uint32 (the reference's `unsigned`).
 *   - Small matrices (Sigma, F, G, Q, ...) are HOST pointers, column-major with
 *     leading dimension = rows, i.e. Eigen::MatrixXd::data().
 *   - Particle arrays: CUSMC_AOS = x[i*d + k] (what the reference's wrappers
 *     flatten VectorXd[N] into, src/mvn_dist.cu.cpp:202-205); CUSMC_SOA =
 *     x[k*ld + i], the device-resident layout the kernels stream with 128-bit loads.
 *   - Functions ending in _dev take DEVICE pointers for the particle-sized arrays
 *     and enqueue on the context's stream without synchronising; the others take
 *     HOST pointers, copy in and out, and return when the result is on the host.
 *   - Every function returns an int status (0 = CUSMC_OK); the message of the last
 *     failure is kept per context (cusmc_last_error).  Nothing throws or exits
 *     across this ABI (the reference's CUDA_CALL/FATAL -> Rcpp::stop convention,
 *     inst/include/support.cuh:9-32, is applied by the binding, not here).
 *   - One context = one device = one caller thread at a time; no global mutable
 *     state, contexts are independent.
 *   - There is no CPU fallback: without a usable CUDA device every call fails.
 */
#ifndef CUSMC_B200_H
#define CUSMC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CUSMC_VERSION 100

enum cusmc_status {
    CUSMC_OK = 0,
    CUSMC_ERR_INVALID = 1,      /* bad argument (NULL, size, unsupported d) */
    CUSMC_ERR_CUDA = 2,         /* CUDA runtime / launch failure */
    CUSMC_ERR_NOT_SPD = 3,      /* covariance not symmetric positive definite */
    CUSMC_ERR_DEGENERATE = 4,   /* all weights zero / non-finite: nothing to resample from */
    CUSMC_ERR_UNSUPPORTED = 5,
    CUSMC_ERR_TIMEOUT = 6       /* sharded run: a peer never arrived at a scalar exchange; results void */
};

enum cusmc_dist_kind { CUSMC_MVN = 0, CUSMC_MVT = 1 };   /* Distributions["mvn"|"mvt"], src/mcmc.cpp:53-58 */
enum cusmc_layout { CUSMC_SOA = 0, CUSMC_AOS = 1 };
enum cusmc_resampler {                                    /* Resamplers[...], src/mcmc.cpp:252-255 */
    CUSMC_RESAMPLE_METROPOLIS = 0,   /* the reference's only resampler */
    CUSMC_RESAMPLE_SYSTEMATIC = 1,
    CUSMC_RESAMPLE_MULTINOMIAL = 2,
    CUSMC_RESAMPLE_REJECTION = 3,    /* unbiased relative of the reference's resampler: see cusmc_rejection_resample_dev */
    CUSMC_RESAMPLE_METROPOLIS_C2 = 4 /* the reference's rule with warp-coalesced proposals: see cusmc_metropolis_c2_dev */
};

#define CUSMC_MAX_DIM 32            /* largest d with an unrolled kernel */
#define CUSMC_MAX_PEERS 8           /* ranks (GPUs of one NVLink domain) of a sharded filter */
#define CUSMC_IPC_HANDLE_BYTES 64   /* sizeof(cudaIpcMemHandle_t) */
#define CUSMC_FILTER_IPC_BUFFERS 7   /* state x 2, ancestors, weights, mailbox, weight image x 2 */

typedef struct cusmc_ctx cusmc_ctx;

/* ---- context ---------------------------------------------------------------- */
int cusmc_version(void);
int cusmc_ctx_create(cusmc_ctx **ctx, int device);
int cusmc_ctx_destroy(cusmc_ctx *ctx);
const char *cusmc_last_error(const cusmc_ctx *ctx);
/* Use an existing cudaStream_t (e.g. torch's current stream); NULL = the context's own
 * non-blocking stream.  To run on the legacy default stream pass cudaStreamLegacy (0x1). */
int cusmc_ctx_set_stream(cusmc_ctx *ctx, void *cuda_stream);
int cusmc_ctx_synchronize(cusmc_ctx *ctx);
/* Number of kernels this context has launched so far (bench.py's gpu_launches). */
uint64_t cusmc_ctx_launch_count(const cusmc_ctx *ctx);
/* Device time in ms of the last host-pointer call between its first and last enqueued operation
 * (CUDA events on the stream; cusmc_logpdf pipelines copies and kernels, so this is the whole call). */
double cusmc_ctx_last_kernel_ms(const cusmc_ctx *ctx);

/* ---- a1/a2: batched density, shared covariance ------------------------------- */
/*
 * out[i] = pdf(x_i) or log pdf(x_i) of MVN(mu, Sigma) / MVT(mu, Sigma, nu) for N points.
 * Replaces MultiVariateNormalDistribution::pdf (src/statistics.cc.cpp:171-196) and
 * MultiVariateTStudentDistribution::pdf (:295-324) evaluated per particle; nu is a
 * float and (nu + d) is a float sum, as in the reference (:302).
 * mu may be NULL (zero mean).  Sigma is factored once on the host (Cholesky; the
 * whitening operator W = L^-1 travels in the kernel's parameter bank).
 */
int cusmc_logpdf_dev(cusmc_ctx *ctx, int kind, int want_log,
                     const double *x_dev, int layout, int64_t N, int64_t ld, int d,
                     const double *mu, const double *sigma, float nu,
                     double *out_dev);
int cusmc_logpdf(cusmc_ctx *ctx, int kind, int want_log,
                 const double *x_host, int layout, int64_t N, int64_t ld, int d,
                 const double *mu, const double *sigma, float nu,
                 double *out_host);

/* Per-point covariance: L_dev holds N packed lower Cholesky factors, point-major,
 * row-packed (L00, L10, L11, L20, ...; d(d+1)/2 doubles each); mu_dev is N x d AoS or NULL;
 * x_dev is AoS N x d.  out[i] = log pdf (want_log) or pdf. */
int cusmc_logpdf_perpoint_dev(cusmc_ctx *ctx, int kind, int want_log,
                              const double *x_dev, const double *mu_dev, const double *L_dev,
                              int64_t N, int d, float nu, double *out_dev);

/* ---- drop-ins for the reference's L0 wrappers (host pointers) ------------------ */
/*
 * mvn_pdf_kernel_wrapper (inst/include/distributions/mvn_dist.hpp:19-26,
 * src/mvn_dist.cu.cpp:671-808): w[i] = norm * exp(-1/2 r^T E_inv r), r = y - F x_i.
 * x_aos is the flattened post_x_t[t]; E_inv is dy x dy, F is dy x d.
 */
int cusmc_mvn_pdf(cusmc_ctx *ctx, double *w, const double *y, const double *x_aos,
                  double norm, const double *E_inv, const double *F,
                  int64_t N, int d, int dy);
/* mvt_pdf_kernel_wrapper (mvn_dist.hpp:38-46, src/mvt_dist.cu.cpp:573-697):
 * w[i] = norm * (1 + q/df)^(-(df + dy)/2). */
int cusmc_mvt_pdf(cusmc_ctx *ctx, double *w, const double *y, const double *x_aos,
                  const double *E_inv, const double *F, double norm,
                  int64_t N, int d, int dy, float df);
/*
 * mvn_sample_kernel_wrapper, propagate overload (mvn_dist.hpp:28-31,
 * src/mvn_dist.cu.cpp:175-319): x_new[i] = G x_prev[a[i]] + Q xi_i.
 * xi (N x d AoS standard normals) may be NULL: they are then drawn on the device
 * from Philox4x32-10 keyed by (seed, step) -- see include/cusmc_philox.h.
 * a may be NULL (identity).  x_new may alias nothing.
 */
int cusmc_mvn_sample(cusmc_ctx *ctx, double *x_new_aos, const double *x_prev_aos,
                     const uint32_t *a, const double *G, const double *Q,
                     const double *xi, uint64_t seed, uint64_t step,
                     int64_t N, int d);
/* init overload (mvn_dist.hpp:33-36, src/mvn_dist.cu.cpp:321-453): x[i] = mu + Q xi_i. */
int cusmc_mvn_sample_init(cusmc_ctx *ctx, double *x_aos, const double *mu, const double *Q,
                          const double *xi, uint64_t seed, int64_t N, int d);
/* mvt_sample_kernel_wrapper (mvn_dist.hpp:48-54, src/mvt_dist.cu.cpp:225-354):
 * x_new[i] = G x_prev[a[i]] + chi_i (.) (Q xi_i); chi = sqrt(df / chi2_df) per component
 * (the reference's Q2 semantics).  chi/xi NULL -> drawn on the device. */
int cusmc_mvt_sample(cusmc_ctx *ctx, double *x_new_aos, const double *x_prev_aos,
                     const uint32_t *a, const double *G, const double *Q,
                     const double *xi, const double *chi, uint64_t seed, uint64_t step,
                     int64_t N, int d, float df);
/*
 * Sampler::metropolis_hastings / resampler_f (src/samplers.cpp:7-36,
 * inst/include/types.hpp:32): for each i, k = i; B times: if (u <= w[j]/w[k]) k = j.
 * u, j: N x B in the reference's consumption order (for i, for n), or both NULL to
 * draw them on the device from Philox keyed by (seed, step).  a receives a_t[t*N + i].
 */
int cusmc_metropolis_hastings(cusmc_ctx *ctx, uint32_t *a, const double *w,
                              const double *u, const uint32_t *j,
                              uint64_t seed, uint64_t step, int64_t N, int B);

/* ---- device-resident building blocks (SoA state) ------------------------------ */
/* is_log = 0: the reference rule u <= w[j]/w[k];  1: w holds log-weights, u <= exp(lw[j]-lw[k])
 * with the reproducible exp of cusmc_detmath.h (robust to the underflow of SURVEY.md Q8). */
int cusmc_metropolis_hastings_dev(cusmc_ctx *ctx, uint32_t *a_dev, const double *w_dev,
                                  const double *u_dev, const uint32_t *j_dev,
                                  uint64_t seed, uint64_t step, int64_t N, int B, int is_log);
/*
 * Metropolis-C2 (Dulger et al., "Memory coalescing for parallelised Metropolis resampling"; the same
 * `resampler_f` seam, inst/include/types.hpp:32): the rule of Sampler::metropolis_hastings
 * (src/samplers.cpp:21-35) with the proposals of the 32 particles i0 .. i0 + 31 of a warp confined, at each
 * iteration n, to ONE 32-particle segment of the weight vector -- the segment [first, first + len) holding a
 * uniform index drawn from Philox (seed, CUSMC_STREAM_SEGMENT, step, i / 32, n), so segments are picked in
 * proportion to their length and each proposal is still uniform over 0 .. N-1; j_n = first + (the lane's
 * own 64 Metropolis bits scaled to len), u_n the lane's own uniform.  A warp's 32 scattered 32-byte sectors
 * per iteration become 8 contiguous ones.  Draws are device-side only (a host reproduces them from the
 * counters; oracle: orc_rng_metropolis_c2).
 */
int cusmc_metropolis_c2_dev(cusmc_ctx *ctx, uint32_t *a_dev, const double *w_dev, uint64_t seed, uint64_t step,
                            int64_t N, int B, int is_log);
/*
 * Rejection resampler (Murray, Lee & Jacob 2016, the unbiased relative of the reference's B-step
 * Metropolis rule, same `resampler_f` seam, inst/include/types.hpp:32): for each i, k = i; attempt n =
 * 0, 1, ...: accept k if u_n <= w[k] / w_max, else k = j_n.  (u_n, j_n) is the Philox draw the Metropolis
 * resampler would use for (seed, step, i, n), so a host reproduces every ancestor.  w_max_dev: the
 * largest weight (device).  Attempts are capped at `cap` (the current k is kept: a bias of at most
 * (1 - mean w / w_max)^cap).  a_dev receives N ancestors.
 */
int cusmc_rejection_resample_dev(cusmc_ctx *ctx, uint32_t *a_dev, const double *w_dev, const double *w_max_dev,
                                 uint64_t seed, uint64_t step, int64_t N, int cap);
/*
 * Fused propagate + reweight (src/mcmc.cpp:90-160 + :162-237), SoA in and out:
 *   x_new[:, i] = G x_prev[:, a[i]] + noise_i,   noise = Q xi (mvn) | chi (.) (Q xi) (mvt)
 *   lw[i]       = log pdf_V(y - F x_new[:, i])   (want_log) or the density itself
 * a_dev NULL = identity; xi_dev/chi_dev are SoA [d][ld] or NULL (Philox).
 * lw_max_dev (optional, 1 double, must be initialised to -inf by the caller or
 * by cusmc_weights_begin) receives max_i lw[i] via an ordered atomic.
 */
int cusmc_propagate_reweight_dev(cusmc_ctx *ctx, int kind, int want_log,
                                 double *x_new_dev, const double *x_prev_dev,
                                 const uint32_t *a_dev, int64_t N, int64_t ld, int d, int dy,
                                 const double *G, const double *Q,
                                 const double *y, const double *F, const double *V, float nu,
                                 const double *xi_dev, const double *chi_dev,
                                 uint64_t seed, uint64_t step,
                                 double *lw_dev, double *lw_max_dev);

/*
 * Weight normalisation and resampling on the deterministic fixed-point image (DESIGN.md,
 * include/cusmc_detmath.h):
 *   wn_i = exp(lw_i - max)  (log weights, cusmc_det_exp)   or   w_i / max  (linear weights)
 *   q_i  = (uint64) trunc(wn_i * 2^shift),  shift = 61 - ceil(log2(N_global))
 * Sums and prefix sums of q are INTEGER, so they do not depend on the order blocks, tiles or
 * ranks are combined in; a CPU can reproduce every ancestor bit for bit.
 *
 * cusmc_weights_max_dev : *max_dev = max_i w_i over finite entries (set to -inf first).
 * cusmc_weights_sum_dev : stats_dev[0..2] = { sum q_i, sum trunc(wn_i^2 2^shift), #(q_i > 0) }
 *                         (4 words, zeroed first).  log-sum-exp = max + log(stats[0] / 2^shift),
 *                         ESS = stats[0]^2 / (stats[1] 2^shift).  The same pass leaves the WEIGHT
 *                         IMAGE in tile_prefix_dev: the exclusive prefix of the per-tile sums
 *                         (tile = 2048 weights) and every weight's inclusive prefix inside its
 *                         tile -- cusmc_tile_prefix_words(N) uint64 words (32-byte alignment lets
 *                         the pass use 256-bit stores; anything else falls back to 8-byte ones);
 *                         NULL = context scratch.  Words 0..7 are scratch of the resampling pass:
 *                         cusmc_resample_systematic_dev leaves its per-launch constants there.
 *                         exp() is evaluated once per weight, here.
 * cusmc_weights_scan_dev: cdf_dev[i] = *cdf_offset_dev + inclusive prefix sum of q.  Because the
 *                         total must be known before a single child can be assigned, the sum pass
 *                         above is mandatory anyway and hands this pass everything it needs:
 *                         CDF_i = offset + prefix[tile(i)] + local_i, one thread per weight, tiles
 *                         independent, nothing spins (a decoupled look-back would add a serial
 *                         dependency for information that is already in memory).  tile_prefix_dev:
 *                         what cusmc_weights_sum_dev left for the SAME (w, max, N, N_global), or
 *                         NULL to compute it here.  cdf_offset_dev may be NULL (0): it is the
 *                         fixed-point mass held by lower-ranked shards.
 * On several GPUs the caller all-reduces max (MAX) and stats (SUM) between these calls and
 * passes the exclusive prefix of the per-rank sums as cdf_offset (cusmc_b200/sharded.py).
 */
int64_t cusmc_tile_prefix_words(int64_t N);
int cusmc_weights_max_dev(cusmc_ctx *ctx, const double *w_dev, int64_t N, double *max_dev);
int cusmc_weights_sum_dev(cusmc_ctx *ctx, const double *w_dev, int is_log, const double *max_dev,
                          int64_t N, int64_t N_global, uint64_t *stats_dev,
                          uint64_t *tile_prefix_dev);
int cusmc_weights_scan_dev(cusmc_ctx *ctx, const double *w_dev, int is_log, const double *max_dev,
                           int64_t N, int64_t N_global, const uint64_t *cdf_offset_dev,
                           const uint64_t *tile_prefix_dev, uint64_t *cdf_dev);

/*
 * Systematic resampling, offspring-scatter form, fused into the scan: with T = *total_dev the
 * global fixed-point mass and r0 = min((uint64)(u0 * (double)T), T - 1), child slot i (global,
 * 0 <= i < N_global) belongs to the parent j with  C_{j-1} * N_global <= i * T + r0 < C_j * N_global
 * (C = global inclusive prefix; 128-bit integer compare).  Every local parent j writes
 *   a_dev[i - out_lo] = j0 + j      for its children i inside [out_lo, out_lo + out_n).
 * j0 = global index of w_dev[0].  One GPU: j0 = out_lo = 0, out_n = N_local = N_global.
 * tile_prefix_dev as for cusmc_weights_scan_dev (its words 0..7 are written: the pass keeps its
 * per-launch constants -- T, r0 and the two quotients of the offspring estimate -- there).
 */
int cusmc_resample_systematic_dev(cusmc_ctx *ctx, const double *w_dev, int is_log,
                                  const double *max_dev, int64_t N_local, int64_t N_global,
                                  const uint64_t *total_dev, const uint64_t *cdf_offset_dev,
                                  const uint64_t *tile_prefix_dev, int64_t j0, int64_t out_lo,
                                  int64_t out_n, double u0, uint32_t *a_dev);
/* Multinomial: a_dev[t] = j0 + #{ j : cdf_j <= p },  p = min((uint64)(u * (double)T), T - 1), for
 * children i0 .. i0 + n_out - 1; u = u_dev[t] or, if u_dev is NULL, the Philox draw keyed by
 * (seed, step, i0 + t). */
int cusmc_resample_multinomial_dev(cusmc_ctx *ctx, const uint64_t *cdf_dev, int64_t N,
                                   const uint64_t *total_dev, const double *u_dev,
                                   uint64_t seed, uint64_t step, int64_t i0, int64_t n_out,
                                   int64_t j0, uint32_t *a_dev);

/* Host-pointer conveniences: full resampling of N weights (linear domain) -> ancestors. */
int cusmc_resample_systematic(cusmc_ctx *ctx, const double *w, int64_t N, double u0, uint32_t *a);
int cusmc_resample_multinomial(cusmc_ctx *ctx, const double *w, int64_t N, const double *u,
                               uint32_t *a);
/* lse = lw_max + log(sum exp(lw - lw_max)),  ess = (sum wn)^2 / sum wn^2. */
int cusmc_normalize_ess(cusmc_ctx *ctx, const double *lw, int64_t N, double *lse, double *ess);

/* ---- independent Metropolis-Hastings chains (configs[2]) ------------------------ */
/*
 * C random-walk MH chains on an MVN / MVT target with per-chain (shared = 0) or
 * single (shared = 1) location mu and lower Cholesky factor L (column-major d x d,
 * strict upper ignored).  Proposal x' = x + step_size * (L z); because it uses the target's own
 * factor the chain is run in whitened coordinates v = L^-1 (x - mu): v' = v + step_size z,
 * q' = |v'|^2, x = mu + L v on output (a chain that never accepts keeps its x bit for bit).  Accept rule
 * (transcendental-free so decisions are bit-reproducible on the host):
 *   mvn: 0.5 (q' - q) < thr             thr = -log(u)
 *   mvt: (1 + q'/nu) < thr (1 + q/nu)   thr = exp(2 (-log u) / (nu + d))
 * z_dev [C][steps][d] and thr_dev [C][steps] pre-drawn, or both NULL: drawn in the
 * kernel from Philox keyed by (seed, chain, step).  x_dev is AoS C x d, updated in place.
 * n_accept_dev (C), accept_bits_dev (C x steps bytes), sum_x_dev / sum_xx_dev (C x d running
 * sums over steps) are optional.
 */
/* Device-drawn proposal normals of both chain entry points (z_dev == NULL): reproducible = 1 (the default)
 * draws them with Philox4x32-10 and the FFMA-only Box-Muller a host regenerates bit for bit (the oracle's
 * mirror; what the bit-exact tests use); 0 = the throughput generator of the filter kernels (Philox4x32-7,
 * Box-Muller on the special-function unit): same law and resolution, ~1/4 of the instructions per normal,
 * last bits not reproducible on a CPU.  Thresholds keep the exact logarithm either way; pre-drawn z / thr
 * are unaffected.  (The filter's switch is cusmc_filter_config.reproducible_rng.) */
int cusmc_ctx_set_chain_noise(cusmc_ctx *ctx, int reproducible);

int cusmc_mh_chains_dev(cusmc_ctx *ctx, int kind, int64_t C, int d, int steps,
                        double step_size, double nu, int shared,
                        const double *mu_dev, const double *L_dev, double *x_dev,
                        const double *z_dev, const double *thr_dev, uint64_t seed,
                        uint32_t *n_accept_dev, uint8_t *accept_bits_dev,
                        double *sum_x_dev, double *sum_xx_dev);

/*
 * The same chains with a proposal that does NOT use the target's factor:  x' = x + step_size * scale (.) z
 * (scale_dev: d per-component factors shared by all chains, or NULL = isotropic).  Nothing cancels, so
 * every step evaluates the target density in the kernel: r = x' - mu, v = L^-1 r by forward substitution
 * (the chain's factor stays in registers for the whole run), q' = |v|^2, and the accept rule above on
 * (q', q) -- proposal, acceptance ratio and accept / reject fused with the density evaluation
 * (src/statistics.cc.cpp:295-311 under the rule of src/samplers.cpp:30).  Other arguments as above.
 */
int cusmc_mh_chains_general_dev(cusmc_ctx *ctx, int kind, int64_t C, int d, int steps,
                                double step_size, const double *scale_dev, double nu, int shared,
                                const double *mu_dev, const double *L_dev, double *x_dev,
                                const double *z_dev, const double *thr_dev, uint64_t seed,
                                uint32_t *n_accept_dev, uint8_t *accept_bits_dev,
                                double *sum_x_dev, double *sum_xx_dev);

/* ---- the filter (a7): particle_filter() / MCMC() -------------------------------- */
/*
 * Device-resident bootstrap particle filter for x_t = G x_{t-1} + w_t, y_t = F x_t + v_t
 * (src/particle_filter.cpp:6-39, src/mcmc.cpp:44-88,239-309).  State stays on the
 * device in SoA double buffers owned by the filter object; history is optional.
 */
typedef struct cusmc_filter cusmc_filter;

typedef struct cusmc_filter_config {
    int64_t N;            /* particles */
    int d, dy, T;         /* state dim, observation dim, time steps (t = 0 .. T-1) */
    int kind;             /* cusmc_dist_kind for both noises (as the reference does) */
    int resampler;        /* cusmc_resampler */
    int B;                /* Metropolis steps per particle (reference hard-codes 10, src/mcmc.cpp:291) */
    float nu;             /* degrees of freedom (mvt) */
    double noise_scale;   /* multiplies Q_c0 and Q_w: sqrt(3) reproduces the reference CPU
                             build's CLT sampler moments, 1 the GPU build's */
    uint64_t seed;        /* Philox key for device-drawn randomness */
    const double *Y;      /* dy x T column-major observations (host) */
    const double *m0, *C0, *F, *G, *V, *W;   /* host, column-major */
    int keep_history;     /* 1: keep x (T x N x d), w (T x N), a (T x N) on the device for cusmc_filter_get_history
                             (cusmc_run needs no flag: it streams the history through a bounded ring) */
    int summary;          /* 1: per-step weighted posterior mean (one extra pass over the state) */
    /* Sharded runs (one process per GPU): N is the GLOBAL particle count; rank r of `world` owns the
     * global slots r*per .. min((r+1)*per, N) - 1, per = ceil(N / world).  world <= 1: one GPU. */
    int rank, world;
    /* cusmc_filter_run executes the whole run as ONE persistent cooperative kernel (one grid barrier
     * per step instead of two launches; every block keeps its tile of the weight image and runs the
     * tile update itself) when the configuration allows it: one GPU, systematic resampling at every
     * step, Normal noise, d == dy in {2, 4, 8}, device-drawn noise, no history, no posterior-mean
     * summary (the ESS and likelihood sums of every step are recorded either way), and a cloud that
     * fits one tile per resident block (cusmc_filter_tile_size() reports the tile).
     * Results are bit-identical to the per-step path run with that tile_size (19.6 vs 36 us per
     * 10^6-particle step).  0 = automatic, -1 = never. */
    int persistent;
    /* Adaptive resampling (systematic resampler only): 0 (default) = resample at every step, as the
     * reference does (src/mcmc.cpp:295); in (0, 1] = resample at step t only when the effective
     * sample size of step t - 1 is below ess_threshold * N.  A step that does not resample keeps
     * every particle's own ancestor (a_i = i) and ACCUMULATES the log-weights,
     * lw_t[i] = lw_{t-1}[i] + log p(y_t | x_t[i]).  The decision is taken on the device from the
     * integer weight sums, so it is identical on every rank of a sharded run and on the CPU oracle. */
    double ess_threshold;
    /* distribution "mvt": the reference draws x_0 from the SAME distribution object as the transition
     * noise (src/mcmc.cpp:73-79 -> MultiVariateTStudentDistribution::sample, src/statistics.cc.cpp:355-411),
     * i.e. x_0 = m0 + chi (.) (Q_c0 xi) -- the default here too.  1 = draw a Normal x_0 instead
     * (round 1's behaviour). */
    int mvt_normal_init;
    /* Device-drawn normals: 0 (default) = the throughput generator, Philox words through the special-
     * function unit (MUFU lg2 / sqrt / sin / cos, ~12 instructions per pair); 1 = the reproducible one
     * (FFMA-only polynomials, ~67 per pair) whose every bit a host can regenerate -- the oracle mirrors
     * it, so device-drawn runs can be checked bit for bit on a CPU.  Same law and resolution either way;
     * injected draws are unaffected. */
    int reproducible_rng;
    /* Particles per weight-image tile of the per-step path (a tile = the children one thread block finishes;
     * the weight image, hence every bit of a run, is defined per tile).  0 = 2048.  A multiple of 32 in
     * [32, 2048] (single GPU, systematic resampling) reproduces a run whose tile was something else -- the
     * persistent kernel's evenly spread tile, as reported by cusmc_filter_tile_size(). */
    int tile_size;
} cusmc_filter_config;

/* Injected randomness for one run (all DEVICE pointers, any may be NULL -> Philox):
 *   xi0 [d][N] SoA; per step t = 1..T-1: xi [(T-1)][d][N], chi idem,
 *   u [(T-1)][N][B], j [(T-1)][N][B] (metropolis), u0 [(T-1)] host doubles (systematic),
 *   um [(T-1)][N] (multinomial), chi0 [d][N] SoA: the factors of the initial draw (mvt). */
typedef struct cusmc_filter_draws {
    const double *xi0_dev, *xi_dev, *chi_dev, *u_dev;
    const uint32_t *j_dev;
    const double *u0_host;
    const double *um_dev;
    const double *chi0_dev;
} cusmc_filter_draws;

int cusmc_filter_create(cusmc_ctx *ctx, const cusmc_filter_config *cfg, cusmc_filter **out);
int cusmc_filter_destroy(cusmc_filter *f);
/* Particles per tile of the weight image cusmc_filter_run will use: 2048 for the per-step path, the
 * evenly spread tile of the persistent kernel when the run takes that path.  The resampling image is
 * defined per tile, so a CPU restatement needs this number to reproduce a run bit for bit. */
int64_t cusmc_filter_tile_size(cusmc_filter *f);
/* Runs t = 0 (initialize) then steps 1 .. T-1 on the stream; returns after enqueueing. */
int cusmc_filter_run(cusmc_filter *f, const cusmc_filter_draws *draws);
/*
 * The same run, phase by phase -- what cusmc_filter_run chains on one GPU:
 *     begin;  weigh(0);  for t = 1 .. T-1:  resample(t);  propagate(t);  weigh(t)
 * For the normalised resamplers propagate(t) is the FUSED step kernel (every block finds the parents of
 * its tile of children in the weight image of step t - 1, propagates, reweights and leaves its tile of
 * the new image; systematic resample(t) is a no-op, multinomial searches the images of step t - 1 -- its own
 * rank's or a peer's -- rank, then tile, then particle, the global CDF never materialised) and
 * weigh(t) is the one-block tile update (global maximum, rescaled tile prefixes, total mass, constants of
 * step t + 1), followed by the moment pass when cfg.summary is set.  Reference mode ("metropolis",
 * "metropolis_c2", "rejection") keeps densities: resample(t) runs the resampler, propagate(t) the one-particle-per-thread
 * step kernel.
 * A sharded filter (cfg.world > 1) is either run by cusmc_filter_run_sharded (scalar exchanges inside the
 * update kernel) or driven through these phases by a binding that carries the scalars itself
 * (cusmc_filter_weigh_phase below; cusmc_b200/sharded.py with NCCL).  Peers' weight images and states
 * are read by peer loads inside the step kernel; ancestors are GLOBAL indices, noise is keyed by the
 * global slot and a shard is a whole number of 2048-particle tiles, so a sharded run reproduces the
 * single-GPU run bit for bit.
 * Slot layout (8 x 8 bytes): { double lw_max; uint64 sum_q, sum_q2, n_pos, cdf_offset, resampled, degenerate; 1 spare }.
 * Injected draws of a sharded run are this rank's shard (leading dimension = its particle count).
 */
int cusmc_filter_begin(cusmc_filter *f, const cusmc_filter_draws *draws);
int cusmc_filter_weigh(cusmc_filter *f, int t);
/* weigh(t) in three phases for callers that carry the scalars themselves (NCCL formulation):
 *   phase 0: this rank's maximum log-weight -> slot[t] word 0            (then all-reduce MAX it)
 *   phase 1: rescale + scan against that maximum, this rank's sums -> slot[t] words 1..2
 *                                                                     (then all-gather words 1..3)
 *   phase 2: global totals, rank offsets, next step's constants from rank_sums_dev (3 words per rank,
 *            rank-major: what the all-gather produced), then the posterior moments. */
int cusmc_filter_weigh_phase(cusmc_filter *f, int t, int phase, const uint64_t *rank_sums_dev);
int cusmc_filter_resample(cusmc_filter *f, int t);
int cusmc_filter_propagate(cusmc_filter *f, int t);
/* Records the start (which = 0) / end (1) event cusmc_filter_last_ms measures between. */
int cusmc_filter_mark(cusmc_filter *f, int which);
int cusmc_filter_slot_dev(cusmc_filter *f, int t, void **slot_dev);
/* Per-step moment sums [T][2 + d] = { sum w, sum w^2, sum w x_k } of this rank's shard (device);
 * a sharded run all-reduces them (SUM) before cusmc_filter_get_summary. */
int cusmc_filter_moments_dev(cusmc_filter *f, double **moments_dev);
/* Peer mapping: export this rank's CUSMC_FILTER_IPC_BUFFERS buffers (state x 2, ancestors, weights,
 * mailbox, weight image x 2) as that many x CUSMC_IPC_HANDLE_BYTES bytes; after an all-gather of those, attach maps
 * every other rank's buffers (cudaIpcOpenMemHandle; all_handles is rank-major). */
int cusmc_filter_ipc_export(cusmc_filter *f, unsigned char *handles);
int cusmc_filter_ipc_attach(cusmc_filter *f, const unsigned char *all_handles);
/*
 * The whole sharded run enqueued by the library: two launches per step and rank, the two scalar
 * exchanges (maximum; per-rank sums) done INSIDE the tile-update kernel over PEER MEMORY (each rank
 * stores its scalars, flag included, into every peer's mailbox and spins, bounded, on its own) -- no
 * collective library and no host round trip inside the time loop.  Every rank calls it collectively.
 * cusmc_filter_exchange_status returns non-zero if a bounded spin timed out (a peer never arrived); the
 * run's results are then void and every getter returns CUSMC_ERR_TIMEOUT.
 */
int cusmc_filter_run_sharded(cusmc_filter *f, const cusmc_filter_draws *draws);
int cusmc_filter_exchange_status(cusmc_filter *f, uint64_t *status);
/* Bound of every spin-wait of the peer-memory exchanges (default 2 s).  Raise it when ranks share a
 * device or a peer's stream may still be busy with earlier work when the run is enqueued. */
int cusmc_filter_set_exchange_timeout(cusmc_filter *f, double seconds);
/* Waits for the last run and reports what went wrong inside it: CUSMC_ERR_TIMEOUT (a scalar exchange
 * of a sharded run timed out) or CUSMC_ERR_DEGENERATE (a step had no weight mass to resample from:
 * its ancestors are the identity).  cusmc_filter_get_summary / _get_history / cusmc_run return the
 * same status instead of handing out void results. */
int cusmc_filter_status(cusmc_filter *f);

/* Per-step outputs copied to the host (any pointer may be NULL):
 * mean [T][d] weighted posterior mean, ess [T], loglik [T] (log of the mean weight). */
int cusmc_filter_get_summary(cusmc_filter *f, double *mean, double *ess, double *loglik);
/* resampled[t] = 1 if step t drew new ancestors (always 1 without ess_threshold; resampled[0] = 0). */
int cusmc_filter_get_resampled(cusmc_filter *f, int *resampled);
/* History (needs keep_history = 1): x_aos [T][n][d], a [T][n] (global parent ids, row 0 = identity)
 * and w [T][n], the weights as the reference's R-level `weights` (src/run.rcpp.cpp:110-117):
 *   resampler metropolis (reference mode): the raw densities, w_0 = 1/N (src/mcmc.cpp:85,212);
 *   systematic / multinomial             : NORMALISED weights exp(lw - max) / sum (summing to one over
 *                                          the whole cloud; 1/N at t = 0; cumulative over the steps
 *                                          an ess_threshold run did not resample at).
 * cusmc_filter_get_log_weights returns the raw log-weights [T][n] of the normalised resamplers
 * (exactly what resampling consumed; row 0 is 0).  n = this rank's shard (N on one GPU). */
int cusmc_filter_get_history(cusmc_filter *f, double *x_aos, double *w, uint32_t *a);
int cusmc_filter_get_log_weights(cusmc_filter *f, double *lw);
/* The ancestor tree as paths (needs keep_history = 1, one GPU): lineage[t][i] = index at step t of the
 * ancestor of FINAL particle i (lineage[T-1][i] = i, lineage[t-1][i] = a_t[lineage[t][i]]), traced on
 * the device; n_unique[t] (optional, T ints) = distinct ancestors alive at step t -- the coalescence
 * profile of the genealogy.  x_aos[t][lineage[t][i]] is then the trajectory of particle i. */
int cusmc_filter_get_lineage(cusmc_filter *f, uint32_t *lineage, int *n_unique);
/* Device time of the last run's step loop (t = 1 .. T-1), ms, from CUDA events on the stream. */
double cusmc_filter_last_ms(const cusmc_filter *f);
/* Current device-resident state: x (SoA [d][N]), weights (N), ancestors of the last step (N).  The
 * normalised resamplers carry their weights in the weight image, so w and a are filled in by the LAST
 * step of a run (t = T - 1) -- and w by every step when the summary, the history or ess_threshold need
 * the log-weights anyway; in reference mode ("metropolis", "rejection") both are current after every step. */
int cusmc_filter_state_dev(cusmc_filter *f, double **x_soa_dev, double **w_dev, uint32_t **a_dev);

/* R-level run() (src/run.rcpp.cpp:58-126) on host pointers: runs the filter and returns weights
 * [T][N] (semantics as cusmc_filter_get_history) and posterior_x [T][N][d] exactly as the reference
 * shapes them; either may be NULL.  The history is STREAMED (the reference's io / return path,
 * src/run.rcpp.cpp:110-125, src/io.cpp:7-43): step kernels write their rows into a two-chunk ring on
 * the device, finished chunks travel device -> pinned host memory on a second stream while the next
 * chunk computes, and the calling thread (plus helpers) copies them into the caller's arrays.  Device
 * memory is bounded by the ring, whatever T.  Returns CUSMC_ERR_DEGENERATE if a step had no weight
 * mass.  cusmc_run_ancestors also returns the ancestor indices a_t [T][N] (row 0 = identity), the
 * genealogy needed to trace a particle's path through posterior_x. */
int cusmc_run(cusmc_ctx *ctx, const cusmc_filter_config *cfg, double *weights, double *posterior_x);
int cusmc_run_ancestors(cusmc_ctx *ctx, const cusmc_filter_config *cfg, double *weights,
                        double *posterior_x, uint32_t *ancestors);

/* ---- layout helpers -------------------------------------------------------------- */
int cusmc_aos_to_soa_dev(cusmc_ctx *ctx, const double *aos_dev, double *soa_dev,
                         int64_t N, int64_t ld, int d);
int cusmc_soa_to_aos_dev(cusmc_ctx *ctx, const double *soa_dev, double *aos_dev,
                         int64_t N, int64_t ld, int d);

#ifdef __cplusplus
}
#endif
#endif /* CUSMC_B200_H */

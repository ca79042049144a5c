/*
 * cusmc_oracle.h -- CPU ORACLE for the CuSMC sampling hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, the smoke check
 * in __graft_entry__.py and bench.py's cpu_baseline / --impl reference legs may
 * load it.  The product (cusmc_b200/) never links, imports or falls back to it.
 *
 * It is a plain-C restatement of the reference's CPU arithmetic (the files named
 * in SURVEY.md section 8c), with all randomness INJECTED as arguments (the
 * reference seeds std::mt19937 from std::random_device, so it has no
 * reproducible stream -- SURVEY.md Q7).  Paths cited as "ref:" are relative to
 * the upstream CuSMC tree.
 *
 * Parity status: the reference cannot be compiled in this image (every source
 * includes RcppEigen.h; no R / Rcpp / Eigen headers exist here), so the oracle
 * is pinned by (1) the reference's only known answers -- MVNPDF = 0.1591549,
 * MVTPDF = 0.07799708 (ref: CuSMC/CuSMC.tex:95-105,131-142) and the
 * metropolis_hastings(c(0,0),2,10) -> (0,1) edge case (ref:
 * man/metropolis_hastings.Rd:22-27 + src/samplers.cpp:30) -- and (2) independent
 * scipy / mpmath evaluations of the same formulas (tests/golden/make_golden.py).
 * Functions in the "extended" section (log-sum-exp, ESS, systematic/multinomial
 * resampling, MH chains, counter-based RNG) have NO reference counterpart:
 * PARITY UNPINNED at reference level; their semantics are defined here.
 *
 * Conventions: matrices are column-major (Eigen's default, so a binding can pass
 * MatrixXd::data()); particle arrays are "AoS": x[i*d + k] = component k of
 * particle i (what the reference's wrappers flatten VectorXd[N] into, ref:
 * src/mvn_dist.cu.cpp:202-205).
 */
#ifndef CUSMC_ORACLE_H
#define CUSMC_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- dense helpers mirroring what Eigen does for dynamic matrices ---------- */
/* PartialPivLU determinant / inverse (Eigen routes MatrixXd::determinant() and
 * ::inverse() through PartialPivLU for dynamic sizes; call sites ref:
 * src/statistics.cc.cpp:176-177,190,193,301,306,317,320). */
double orc_determinant(const double *A, int d);
int    orc_inverse(const double *A, int d, double *Ainv);
/* Lower Cholesky factor A = L L^T, column-major; returns 0 or k+1 if pivot k <= 0. */
int    orc_cholesky_lower(const double *A, int d, double *L);
/* Inverse of a lower-triangular matrix (column-major). */
void   orc_tri_inverse_lower(const double *L, int d, double *W);

/* ---- a1: MultiVariateNormalDistribution (ref: src/statistics.cc.cpp:171-211) */
double orc_mvn_norm(const double *sigma, int d);                       /* getNorm :205-211 */
double orc_mvn_pdf1(const double *y, const double *sigma, int d);      /* pdf(y)   :171-180 */
double orc_mvn_pdf2(const double *y, const double *F, const double *mu,
                    const double *sigma, int d);                       /* pdf(y,F) :183-196 */
/* ---- a2: MultiVariateTStudentDistribution (ref: :295-340); nu is float (Q9) */
double orc_mvt_norm(const double *sigma, int d, float nu);             /* getNorm :332-340 */
double orc_mvt_pdf1(const double *y, const double *sigma, int d, float nu);   /* :313-324 */
double orc_mvt_pdf2(const double *y, const double *F, const double *mu,
                    const double *sigma, int d, float nu);             /* :295-311 */

/* R helpers MVNPDF / MVTPDF (ref: src/mvn_dist.rcpp.cpp:52-58, src/mvt_dist.rcpp.cpp:60-66):
 * F = I, pdf(x, F). */
double orc_MVNPDF(const double *x, const double *mu, const double *sigma, int d);
double orc_MVTPDF(const double *x, const double *mu, const double *sigma, int d, float nu);

/* Batched density: out[i] = pdf(x_i - mu) (or its log) for N AoS points.
 * dist: 0 = mvn, 1 = mvt.  faithful != 0 recomputes determinant + inverse per
 * point exactly as the reference does (Q6); 0 hoists them (bit-identical result).
 * want_log != 0 returns log(pdf) computed in the log domain from the same q. */
void orc_pdf_batch(int dist, const double *x_aos, int64_t N, int d,
                   const double *mu, const double *sigma, float nu,
                   int faithful, int want_log, double *out);
/* Same, per-point covariance: sigma_all[i*d*d ...] column-major d x d each. */
void orc_pdf_batch_perpoint(int dist, const double *x_aos, int64_t N, int d,
                            const double *mu_all, const double *sigma_all, float nu,
                            int want_log, double *out);

/* ---- a4: Sampler::metropolis_hastings (ref: src/samplers.cpp:7-36) ---------- */
/* w = w_t[t-1] (N weights); u[i*B+n], j[i*B+n] are the injected U(0,1) and
 * U{0..N-1} draws in the order the reference consumes them (for i, for n);
 * a[i] = a_t[t*N + i].  Single-threaded semantics (the reference's OpenMP loop
 * is racy, Q7). */
void orc_metropolis_hastings(uint32_t *a, const double *w, const double *u,
                             const uint32_t *j, int64_t N, int B);

/* ---- a5/a6: propagate_K, initialize (ref: src/mcmc.cpp:44-88,90-160) -------- */
/* x_new[i] = G x_prev[a[i]] + Q xi_i                         (dist 0, ref: src/statistics.cc.cpp:258)
 *          = G x_prev[a[i]] + chi_i (.) (Q xi_i)             (dist 1, ref: :411)
 * xi = the effective standard draws (the reference's 200-term CLT sum, Q1), chi
 * = per-component sqrt(nu/chi2) (Q2); both injected, AoS N x d.  a may be NULL
 * (identity).  G may be NULL with mu0 != NULL for initialize: x = mu0 + Q xi. */
void orc_propagate(int dist, double *x_new, const double *x_prev, const uint32_t *a,
                   const double *G, const double *mu0, const double *Q,
                   const double *xi, const double *chi, int64_t N, int d);

/* ---- a3: reweight_G (ref: src/mcmc.cpp:185-215) ----------------------------- */
/* w[i] = pdf1(y - F x_i) with sigma = V (mu = 0).  F is dy x d column-major. */
void orc_reweight(int dist, double *w, const double *y, const double *x_aos,
                  const double *F, const double *V, float nu,
                  int64_t N, int d, int dy, int faithful, int want_log);

/* ---- a7: MCMC loop (ref: src/mcmc.cpp:239-309, src/particle_filter.cpp:6-39)
 * Runs t = 1..T-1: metropolis resample -> propagate -> reweight, keeping only
 * the running state (history optional).  Injected draws: xi0 [N*d] for
 * initialize, u/j [(T-1)*N*B], xi [(T-1)*N*d], chi idem (mvt only), chi0 [N*d] the
 * initial draw's per-component factors (mvt; NULL = Normal start).  faithful != 0
 * recomputes determinant() + inverse() per particle in the reweight (ref: src/mcmc.cpp:193-215).
 * Outputs (any may be NULL): x_hist [T*N*d], w_hist [T*N], a_hist [T*N]
 * (row t=0 of a_hist is left untouched, as in the reference), and per-step
 * weighted moments mean_hist [T*d].  Q_c0, Q_w are the eigen factors
 * (ref: src/linear_algebra.cpp:10-23) pre-scaled by the noise scale (Q1). */
void orc_filter_metropolis(int dist, int64_t N, int d, int dy, int T, int B,
                           const double *Y /* dy x T col-major */,
                           const double *m0, const double *Q_c0,
                           const double *F, const double *G,
                           const double *V, const double *Q_w, float nu,
                           const double *xi0, const double *u, const uint32_t *j,
                           const double *xi, const double *chi, const double *chi0,
                           double *x_hist, double *w_hist, uint32_t *a_hist,
                           double *mean_hist, int faithful);

/* ======================= extended (parity unpinned) ======================== */

/* Fixed-order fused-multiply-add quadratic form shared by the production
 * paths: z_k = c_k - sum_{j} M[k,j] v_j accumulated j ascending with fma,
 * q = sum_k z_k^2 accumulated k ascending with fma, starting from 0.
 * M is m x d ROW-major here (the packed "whitening operator"); if tri != 0
 * only j <= k is visited. */
double orc_quadform_fma(const double *M, const double *c, const double *v,
                        int m, int d, int tri);

/* Deterministic elementary functions (IEEE basic ops + fma only) mirrored
 * bit-for-bit by include/cusmc_detmath.h on the device. */
double orc_det_exp(double x);
double orc_det_log(double x);

/* Philox4x32-10 counter-based generator (Salmon et al. 2011), mirrored by
 * include/cusmc_philox.h. */
void orc_philox4x32(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);

/* Max-shifted log-sum-exp and ESS of N log-weights (textbook, libm, serial). */
void orc_logsumexp_ess(const double *lw, int64_t N, double *lse, double *ess, double *lmax);

/* Fixed-point weight image used by the integer scan: shift = 61 - ceil(log2(N_global)),
 * q[i] = (uint64) trunc((w[i] / wmax) * 2^shift) for finite w[i] > 0 else 0.
 * Returns the total (sum of q). */
int      orc_fixed_shift(int64_t N_global);
uint64_t orc_fixed_weights(const double *w, int64_t N, double wmax, int shift, uint64_t *q);

/* Systematic resampling on the fixed-point CDF: C_j = inclusive sum of q,
 * T = C_{N-1}, r0 = min((uint64)(u0 * (double)T), T-1),
 * a[i] = #{ j : C_j * N <= i*T + r0 } (128-bit compare).  Returns 0, or 1 if the
 * weights are degenerate (wmax <= 0 or not finite; a = identity). */
int orc_resample_systematic(const double *w, int64_t N, double u0, uint32_t *a);
/* Multinomial: p_i = min((uint64)(u[i] * (double)T), T-1), a[i] = #{ j : C_j <= p_i }. */
int orc_resample_multinomial(const double *w, int64_t N, const double *u, uint32_t *a);

/* Rejection resampler: k = i; attempt n: accept k if u_n <= w[k] / wmax, else k = j_n; (u_n, j_n) = the
 * counter-based draw of (seed, step, i, n) (orc_rng_metropolis); at most `cap` attempts. */
void orc_resample_rejection(uint32_t *a, const double *w, double wmax, int64_t N, uint64_t seed,
                            uint64_t step, int cap);

/* Independent random-walk Metropolis-Hastings chains.
 * Chain c (AoS): state x_c (d), target dist (0 mvn / 1 mvt) with location mu_c
 * and lower Cholesky factor L_c (column-major d x d; strict upper ignored),
 * proposal x' = x + step * (L_c z).  Whitening v = L_c^{-1}(x' - mu_c) by forward
 * substitution (fixed order, fma), q = sum v_k^2.
 * Accept rule (transcendental-free):
 *   mvn:  0.5*(q' - q) < e              with e = -log(u) pre-drawn
 *   mvt:  (1 + q'/nu) < r * (1 + q/nu)  with r = exp(2 e / (nu + d)) pre-drawn
 * z [C*steps*d] laid out z[(c*steps + s)*d + k]; thr [C*steps].
 * shared != 0: mu/L are single (not per chain).  Outputs: x_final [C*d],
 * n_accept [C], accept_bits (optional) [C*steps] bytes. */
void orc_mh_chains(int dist, int64_t C, int d, int steps, double step, double nu,
                   int shared, const double *mu, const double *L,
                   const double *x0, const double *z, const double *thr,
                   double *x_final, uint32_t *n_accept, uint8_t *accept_bits);


/* The same chains with the proposal x' = x + step * scale (.) z (scale: d factors or NULL): every step
 * evaluates q' = |L^-1 (x' - mu)|^2 by forward substitution (definition in the .c). */
void orc_mh_chains_general(int dist, int64_t C, int d, int steps, double step, const double *scale, double nu,
                           int shared, const double *mu, const double *L, const double *x0, const double *z,
                           const double *thr, double *x_final, uint32_t *n_accept, uint8_t *accept_bits);

/* ---- production-order ("det") restatements: same operation order as the kernels, so the
 * comparison is bit-for-bit (MVN; the MVT epilogue uses libm log1p and is compared to 1e-12). */

/* Host algebra of the observation model exactly as the library does it: Winv = L_V^-1
 * (row-major dy x dy), M = Winv F (row-major dy x d), lognorm of MVN/MVT(0, V). */
int orc_observation_operator(int dist, int d, int dy, const double *F, const double *V, float nu,
                             double *M, double *Winv, double *lognorm);
/* One fused step.  G, Q column-major d x d (Q already scaled), mu may be NULL (0), M row-major,
 * c = Winv y.  x arrays are AoS N x d.  want_log: lw = lognorm - q/2 | lognorm - h log1p(q/nu);
 * else the density scale * exp(-q/2) | scale * (1+q/nu)^-h with scale = exp(lognorm). */
void orc_step_det(int dist, int want_log, double *x_new, double *lw, const double *x_prev,
                  const uint32_t *a, const double *G, const double *Q, const double *mu,
                  const double *M, const double *c, double lognorm, float nu,
                  const double *xi, const double *chi, int64_t N, int d, int dy);
/* Counter-based draws, mirror of include/cusmc_philox.h. */
void orc_det_sincospi(double t, double *s, double *c);
void orc_rng_normal_pair(uint64_t seed, int stream, uint64_t step, uint64_t index, uint32_t sub, double z[2]);
/* The kernels' default generator: four single-precision Box-Muller normals per Philox block. */
float orc_det_logf(float x);
void orc_det_sincospif(float t, float *s, float *c);
void orc_rng_normal4(uint64_t seed, int stream, uint64_t step, uint64_t index, uint32_t sub, double z[4]);
double orc_rng_u01(uint64_t seed, int stream, uint64_t step, uint64_t index, uint32_t sub);
void orc_rng_metropolis(uint64_t seed, uint64_t step, uint64_t index, uint32_t n, uint64_t N,
                        double *u, uint32_t *j);
/* Metropolis-C2: the proposal of particle `index` at iteration n lies in the 32-particle segment its group
 * index / 32 drew for that iteration (extended; mirrors cusmc_metropolis_c2_dev). */
void orc_rng_metropolis_c2(uint64_t seed, uint64_t step, uint64_t index, uint32_t n, uint64_t N,
                           double *u, uint32_t *j);
/* Fills xi (AoS N x d) with the normals the step kernel draws for (seed, stream, step). */
void orc_rng_fill_normals(uint64_t seed, int stream, uint64_t step, int64_t i0, int64_t N, int d, double *xi);
/* Whole filter in production order.  resampler: 0 metropolis (linear weights, as the
 * reference), 1 systematic, 2 multinomial (log weights, block-relative fixed point), 3 rejection / 4 Metropolis-C2 (linear
 * weights, device-drawn only).  Draw arrays may
 * be NULL -> Philox mirror with `seed`.  Layouts as orc_filter_metropolis; u0 [(T-1)], um [(T-1)*N].
 * tile: tile size of the weight image (<= 0: 2048, the library's; the persistent kernel uses its own).
 * Outputs optional: x_hist [T*N*d], w_hist [T*N], a_hist [T*N], ess [T], loglik [T]. */
int orc_filter_det(int dist, int resampler, int64_t N, int d, int dy, int T, int B,
                   const double *Y, const double *m0, const double *Q_c0, const double *F,
                   const double *G, const double *V, const double *Q_w, float nu, uint64_t seed,
                   const double *xi0, const double *chi0, const double *xi, const double *chi, const double *u,
                   const uint32_t *j, const double *u0, const double *um,
                   double *x_hist, double *w_hist, uint32_t *a_hist, double *ess, double *loglik,
                   double ess_threshold, int *resampled, int64_t tile);
/* The block-relative weight image the filter's normalised resamplers read (definition in the .c):
 * C[i] = global inclusive integer CDF; returns the total T; optional T2 (sum of squares), M (max). */
uint64_t orc_tile_image(const double *lw, int64_t N, int64_t tile, int shift, uint64_t *C,
                        uint64_t *T2_out, double *M_out);

/* Threads the batched functions will use (OpenMP), for bench reporting. */
int orc_num_threads(void);
void orc_set_num_threads(int n);   /* launchers such as torchrun export OMP_NUM_THREADS=1 */

#ifdef __cplusplus
}
#endif
#endif

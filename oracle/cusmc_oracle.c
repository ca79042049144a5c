/*
 * cusmc_oracle.c -- CPU ORACLE (test infrastructure only; see cusmc_oracle.h).
 *
 * Plain-C restatement of the reference CPU path.  "ref:" citations are paths in
 * the upstream CuSMC tree.  Build: see oracle/Makefile (gcc -O2 -ffp-contract=off
 * so that a*b+c is never silently fused: every fused operation below is an
 * explicit fma(), which is what makes the fixed-order chains reproducible
 * bit-for-bit on the device).
 */
#include "cusmc_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

#define A_(M, r, c, ld) ((M)[(size_t)(c) * (size_t)(ld) + (size_t)(r)])

void orc_set_num_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int orc_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ------------------------------------------------------------------------- */
/* Dense helpers: partial-pivoting LU as Eigen::PartialPivLU does it          */
/* (row with the largest |entry| in the column is swapped up, Doolittle form).*/
/* ------------------------------------------------------------------------- */
static int lu_factor(double *A, int d, int *piv, int *sign)
{
    int singular = 0;
    *sign = 1;
    for (int k = 0; k < d; ++k) {
        int p = k;
        double best = fabs(A_(A, k, k, d));
        for (int r = k + 1; r < d; ++r) {
            double v = fabs(A_(A, r, k, d));
            if (v > best) { best = v; p = r; }
        }
        piv[k] = p;
        if (best == 0.0) { singular = 1; continue; }
        if (p != k) {
            for (int c = 0; c < d; ++c) {
                double tmp = A_(A, k, c, d);
                A_(A, k, c, d) = A_(A, p, c, d);
                A_(A, p, c, d) = tmp;
            }
            *sign = -*sign;
        }
        double pivot = A_(A, k, k, d);
        for (int r = k + 1; r < d; ++r) A_(A, r, k, d) /= pivot;
        for (int c = k + 1; c < d; ++c) {
            double ukc = A_(A, k, c, d);
            for (int r = k + 1; r < d; ++r) A_(A, r, c, d) -= A_(A, r, k, d) * ukc;
        }
    }
    return singular;
}

double orc_determinant(const double *A, int d)
{
    if (d == 0) return 1.0;
    double *LU = (double *)malloc(sizeof(double) * d * d);
    int *piv = (int *)malloc(sizeof(int) * d);
    int sign;
    memcpy(LU, A, sizeof(double) * d * d);
    lu_factor(LU, d, piv, &sign);
    double det = (double)sign;
    for (int k = 0; k < d; ++k) det *= A_(LU, k, k, d);
    free(LU); free(piv);
    return det;
}

int orc_inverse(const double *A, int d, double *Ainv)
{
    double *LU = (double *)malloc(sizeof(double) * d * d);
    int *piv = (int *)malloc(sizeof(int) * d);
    int sign;
    memcpy(LU, A, sizeof(double) * d * d);
    int singular = lu_factor(LU, d, piv, &sign);
    /* Solve A X = I one column at a time: P A = L U. */
    for (int c = 0; c < d; ++c) {
        double *x = Ainv + (size_t)c * d;
        for (int r = 0; r < d; ++r) x[r] = (r == c) ? 1.0 : 0.0;
        for (int k = 0; k < d; ++k) {
            if (piv[k] != k) { double t = x[k]; x[k] = x[piv[k]]; x[piv[k]] = t; }
        }
        for (int k = 0; k < d; ++k)           /* L y = P b (unit lower) */
            for (int r = k + 1; r < d; ++r) x[r] -= A_(LU, r, k, d) * x[k];
        for (int k = d - 1; k >= 0; --k) {    /* U x = y */
            x[k] /= A_(LU, k, k, d);
            for (int r = 0; r < k; ++r) x[r] -= A_(LU, r, k, d) * x[k];
        }
    }
    free(LU); free(piv);
    return singular;
}

int orc_cholesky_lower(const double *A, int d, double *L)
{
    memset(L, 0, sizeof(double) * d * d);
    for (int c = 0; c < d; ++c) {
        double s = A_(A, c, c, d);
        for (int k = 0; k < c; ++k) s -= A_(L, c, k, d) * A_(L, c, k, d);
        if (!(s > 0.0)) return c + 1;
        double lcc = sqrt(s);
        A_(L, c, c, d) = lcc;
        for (int r = c + 1; r < d; ++r) {
            double t = A_(A, r, c, d);
            for (int k = 0; k < c; ++k) t -= A_(L, r, k, d) * A_(L, c, k, d);
            A_(L, r, c, d) = t / lcc;
        }
    }
    return 0;
}

void orc_tri_inverse_lower(const double *L, int d, double *W)
{
    memset(W, 0, sizeof(double) * d * d);
    for (int c = 0; c < d; ++c) {
        A_(W, c, c, d) = 1.0 / A_(L, c, c, d);
        for (int r = c + 1; r < d; ++r) {
            double s = 0.0;
            for (int k = c; k < r; ++k) s -= A_(L, r, k, d) * A_(W, k, c, d);
            A_(W, r, c, d) = s / A_(L, r, r, d);
        }
    }
}

/* (r^T * P) * r, evaluated left to right as the Eigen expression
 * y.transpose() * sigma.inverse() * y is (ref: src/statistics.cc.cpp:177). */
static double quadform_left_to_right(const double *r, const double *P, int d)
{
    double q = 0.0;
    for (int c = 0; c < d; ++c) {
        double v = 0.0;
        for (int k = 0; k < d; ++k) v += r[k] * A_(P, k, c, d);
        q += v * r[c];
    }
    return q;
}

/* ------------------------------------------------------------------------- */
/* a1: multivariate Normal                                                   */
/* ------------------------------------------------------------------------- */
static double mvn_norm_from_det(double det, int d)
{
    /* ref: src/statistics.cc.cpp:173-176 -- 1 / (sqrt(2 pi)^n * det^0.5) */
    const double sqrt2pi = sqrt(2 * M_PI);
    return 1 / (pow(sqrt2pi, (double)(unsigned)d) * pow(det, 0.5));
}

double orc_mvn_norm(const double *sigma, int d)
{
    return mvn_norm_from_det(orc_determinant(sigma, d), d);
}

double orc_mvn_pdf1(const double *y, const double *sigma, int d)
{
    double *P = (double *)malloc(sizeof(double) * d * d);
    double norm = mvn_norm_from_det(orc_determinant(sigma, d), d);
    orc_inverse(sigma, d, P);
    double q = quadform_left_to_right(y, P, d);
    free(P);
    return norm * exp(-0.5 * q);
}

double orc_mvn_pdf2(const double *y, const double *F, const double *mu,
                    const double *sigma, int d)
{
    /* ref: src/statistics.cc.cpp:192 -- y_Fmu = y - F*mu */
    double *r = (double *)malloc(sizeof(double) * d);
    for (int k = 0; k < d; ++k) {
        double s = 0.0;
        for (int c = 0; c < d; ++c) s += A_(F, k, c, d) * mu[c];
        r[k] = y[k] - s;
    }
    double v = orc_mvn_pdf1(r, sigma, d);
    free(r);
    return v;
}

/* ------------------------------------------------------------------------- */
/* a2: multivariate Student t; nu is a float and nu + n is a float sum (Q9)  */
/* ------------------------------------------------------------------------- */
static double mvt_norm_from_det(double det, int d, float nu)
{
    unsigned n = (unsigned)d;
    double pixdf = M_PI * nu;                                   /* :300 */
    double norm1 = pow(pixdf, (-0.5 * n)) * pow(det, -0.5);     /* :301 */
    float nu_plus_n = nu + (float)n;                            /* float + unsigned -> float */
    double norm2 = tgamma(0.5 * nu_plus_n) / tgamma(0.5 * nu);  /* :302 */
    return norm1 * norm2;
}

double orc_mvt_norm(const double *sigma, int d, float nu)
{
    return mvt_norm_from_det(orc_determinant(sigma, d), d, nu);
}

static double mvt_pdf_from_q(double normc, double q, int d, float nu)
{
    float nu_plus_n = nu + (float)(unsigned)d;
    double quadform = 1.0f + pow((double)nu, -1.0) * q;         /* :308 */
    return normc * pow(quadform, (-0.5 * nu_plus_n));           /* :310 */
}

double orc_mvt_pdf1(const double *y, const double *sigma, int d, float nu)
{
    double *P = (double *)malloc(sizeof(double) * d * d);
    double normc = mvt_norm_from_det(orc_determinant(sigma, d), d, nu);
    orc_inverse(sigma, d, P);
    double q = quadform_left_to_right(y, P, d);
    free(P);
    return mvt_pdf_from_q(normc, q, d, nu);
}

double orc_mvt_pdf2(const double *y, const double *F, const double *mu,
                    const double *sigma, int d, float nu)
{
    double *r = (double *)malloc(sizeof(double) * d);
    for (int k = 0; k < d; ++k) {
        double s = 0.0;
        for (int c = 0; c < d; ++c) s += A_(F, k, c, d) * mu[c];
        r[k] = y[k] - s;
    }
    double v = orc_mvt_pdf1(r, sigma, d, nu);
    free(r);
    return v;
}

static void identity(double *I, int d)
{
    memset(I, 0, sizeof(double) * d * d);
    for (int k = 0; k < d; ++k) A_(I, k, k, d) = 1.0;
}

double orc_MVNPDF(const double *x, const double *mu, const double *sigma, int d)
{
    double *F = (double *)malloc(sizeof(double) * d * d);
    identity(F, d);
    double v = orc_mvn_pdf2(x, F, mu, sigma, d);
    free(F);
    return v;
}

double orc_MVTPDF(const double *x, const double *mu, const double *sigma, int d, float nu)
{
    double *F = (double *)malloc(sizeof(double) * d * d);
    identity(F, d);
    double v = orc_mvt_pdf2(x, F, mu, sigma, d, nu);
    free(F);
    return v;
}

/* Log-domain constants for the want_log variants (same q, log of the same norm). */
static double mvn_lognorm_from_det(double det, int d)
{
    return -((double)d * 0.5 * log(2 * M_PI) + 0.5 * log(det));
}

static double mvt_lognorm_from_det(double det, int d, float nu)
{
    float nu_plus_n = nu + (float)(unsigned)d;
    return -0.5 * d * log(M_PI * nu) - 0.5 * log(det)
           + lgamma(0.5 * nu_plus_n) - lgamma(0.5 * nu);
}

static double density_from_q(int dist, double q, double normc, double lognormc,
                             int d, float nu, int want_log)
{
    if (dist == 0)
        return want_log ? lognormc - 0.5 * q : normc * exp(-0.5 * q);
    if (want_log) {
        float nu_plus_n = nu + (float)(unsigned)d;
        return lognormc - 0.5 * nu_plus_n * log1p(q / (double)nu);
    }
    return mvt_pdf_from_q(normc, q, d, nu);
}

void orc_pdf_batch(int dist, const double *x_aos, int64_t N, int d,
                   const double *mu, const double *sigma, float nu,
                   int faithful, int want_log, double *out)
{
    double *P0 = (double *)malloc(sizeof(double) * d * d);
    double det0 = orc_determinant(sigma, d);
    orc_inverse(sigma, d, P0);
#pragma omp parallel
    {
        double *r = (double *)malloc(sizeof(double) * d);
        double *P = (double *)malloc(sizeof(double) * d * d);
#pragma omp for schedule(static)
        for (int64_t i = 0; i < N; ++i) {
            const double *xi = x_aos + (size_t)i * d;
            for (int k = 0; k < d; ++k) r[k] = mu ? xi[k] - mu[k] : xi[k];
            double det = det0;
            const double *Pi = P0;
            if (faithful) {      /* Q6: determinant() + inverse() on every call */
                det = orc_determinant(sigma, d);
                orc_inverse(sigma, d, P);
                Pi = P;
            }
            double q = quadform_left_to_right(r, Pi, d);
            double normc = dist == 0 ? mvn_norm_from_det(det, d) : mvt_norm_from_det(det, d, nu);
            double lognormc = 0.0;
            if (want_log)
                lognormc = dist == 0 ? mvn_lognorm_from_det(det, d) : mvt_lognorm_from_det(det, d, nu);
            out[i] = density_from_q(dist, q, normc, lognormc, d, nu, want_log);
        }
        free(r); free(P);
    }
    free(P0);
}

void orc_pdf_batch_perpoint(int dist, const double *x_aos, int64_t N, int d,
                            const double *mu_all, const double *sigma_all, float nu,
                            int want_log, double *out)
{
#pragma omp parallel
    {
        double *r = (double *)malloc(sizeof(double) * d);
        double *P = (double *)malloc(sizeof(double) * d * d);
#pragma omp for schedule(static)
        for (int64_t i = 0; i < N; ++i) {
            const double *xi = x_aos + (size_t)i * d;
            const double *S = sigma_all + (size_t)i * d * d;
            const double *m = mu_all ? mu_all + (size_t)i * d : NULL;
            for (int k = 0; k < d; ++k) r[k] = m ? xi[k] - m[k] : xi[k];
            double det = orc_determinant(S, d);
            orc_inverse(S, d, P);
            double q = quadform_left_to_right(r, P, d);
            double normc = dist == 0 ? mvn_norm_from_det(det, d) : mvt_norm_from_det(det, d, nu);
            double lognormc = dist == 0 ? mvn_lognorm_from_det(det, d) : mvt_lognorm_from_det(det, d, nu);
            out[i] = density_from_q(dist, q, normc, lognormc, d, nu, want_log);
        }
        free(r); free(P);
    }
}

/* ------------------------------------------------------------------------- */
/* a4: Metropolis ancestor resampler (ref: src/samplers.cpp:21-35)            */
/* ------------------------------------------------------------------------- */
void orc_metropolis_hastings(uint32_t *a, const double *w, const double *u,
                             const uint32_t *j, int64_t N, int B)
{
    for (int64_t i = 0; i < N; ++i) {
        uint32_t k = (uint32_t)i;
        for (int n = 0; n < B; ++n) {
            double un = u[(size_t)i * B + n];
            uint32_t jn = j[(size_t)i * B + n];
            if (un <= w[jn] / w[k])          /* :30 -- NaN compares false, inf accepts */
                k = jn;
        }
        a[i] = k;
    }
}

/* ------------------------------------------------------------------------- */
/* a5/a6: propagate / initialize                                             */
/* ------------------------------------------------------------------------- */
void orc_propagate(int dist, double *x_new, const double *x_prev, const uint32_t *a,
                   const double *G, const double *mu0, const double *Q,
                   const double *xi, const double *chi, int64_t N, int d)
{
#pragma omp parallel
    {
        double *mu = (double *)malloc(sizeof(double) * d);
#pragma omp for schedule(static)
        for (int64_t i = 0; i < N; ++i) {
            if (G) {                      /* ref: src/mcmc.cpp:133 -- mu = G * x_{t-1}[a] */
                const double *xp = x_prev + (size_t)(a ? a[i] : (uint32_t)i) * d;
                for (int k = 0; k < d; ++k) {
                    double s = 0.0;
                    for (int c = 0; c < d; ++c) s += A_(G, k, c, d) * xp[c];
                    mu[k] = s;
                }
            } else {                      /* ref: src/mcmc.cpp:66,79 -- mu = m0 */
                for (int k = 0; k < d; ++k) mu[k] = mu0 ? mu0[k] : 0.0;
            }
            const double *z = xi + (size_t)i * d;
            for (int k = 0; k < d; ++k) {
                double s = 0.0;           /* (Q * x) */
                for (int c = 0; c < d; ++c) s += A_(Q, k, c, d) * z[c];
                if (dist == 1) s = chi[(size_t)i * d + k] * s;   /* chi.asDiagonal() * (Q x), :411 */
                x_new[(size_t)i * d + k] = s + mu[k];            /* + mu, :258 / :411 */
            }
        }
        free(mu);
    }
}

/* ------------------------------------------------------------------------- */
/* a3: reweight (ref: src/mcmc.cpp:193-215): w_i = pdf(y - F x_i), sigma = V  */
/* ------------------------------------------------------------------------- */
void orc_reweight(int dist, double *w, const double *y, const double *x_aos,
                  const double *F, const double *V, float nu,
                  int64_t N, int d, int dy, int faithful, int want_log)
{
    double *P0 = (double *)malloc(sizeof(double) * dy * dy);
    double det0 = orc_determinant(V, dy);
    orc_inverse(V, dy, P0);
#pragma omp parallel
    {
        double *r = (double *)malloc(sizeof(double) * dy);
        double *P = (double *)malloc(sizeof(double) * dy * dy);
#pragma omp for schedule(static)
        for (int64_t i = 0; i < N; ++i) {
            const double *xi = x_aos + (size_t)i * d;
            for (int k = 0; k < dy; ++k) {
                double s = 0.0;
                for (int c = 0; c < d; ++c) s += A_(F, k, c, dy) * xi[c];
                r[k] = y[k] - s;
            }
            double det = det0;
            const double *Pi = P0;
            if (faithful) {
                det = orc_determinant(V, dy);
                orc_inverse(V, dy, P);
                Pi = P;
            }
            double q = quadform_left_to_right(r, Pi, dy);
            double normc = dist == 0 ? mvn_norm_from_det(det, dy) : mvt_norm_from_det(det, dy, nu);
            double lognormc = 0.0;
            if (want_log)
                lognormc = dist == 0 ? mvn_lognorm_from_det(det, dy) : mvt_lognorm_from_det(det, dy, nu);
            w[i] = density_from_q(dist, q, normc, lognormc, dy, nu, want_log);
        }
        free(r); free(P);
    }
    free(P0);
}

/* ------------------------------------------------------------------------- */
/* a7: the time loop (ref: src/mcmc.cpp:292-308)                             */
/* ------------------------------------------------------------------------- */
static void weighted_mean(const double *x, const double *w, int64_t N, int d, double *mean)
{
    double sw = 0.0;
    for (int k = 0; k < d; ++k) mean[k] = 0.0;
    for (int64_t i = 0; i < N; ++i) {
        sw += w[i];
        for (int k = 0; k < d; ++k) mean[k] += w[i] * x[(size_t)i * d + k];
    }
    for (int k = 0; k < d; ++k) mean[k] /= sw;
}

void orc_filter_metropolis(int dist, int64_t N, int d, int dy, int T, int B,
                           const double *Y, const double *m0, const double *Q_c0,
                           const double *F, const double *G,
                           const double *V, const double *Q_w, float nu,
                           const double *xi0, const double *u, const uint32_t *j,
                           const double *xi, const double *chi, const double *chi0,
                           double *x_hist, double *w_hist, uint32_t *a_hist,
                           double *mean_hist, int faithful)
{
    size_t Nd = (size_t)N * d;
    double *xa = (double *)malloc(sizeof(double) * Nd);
    double *xb = (double *)malloc(sizeof(double) * Nd);
    double *w = (double *)malloc(sizeof(double) * N);
    uint32_t *a = (uint32_t *)malloc(sizeof(uint32_t) * N);

    /* initialize (ref: src/mcmc.cpp:63-85): the SAME distribution object draws x_0, so for "mvt"
     * it is MultiVariateTStudentDistribution::sample (ref: src/statistics.cc.cpp:355-411):
     * x_0 = chi (.) (Q_c0 xi) + m0; for "mvn" x_0 = Q_c0 xi + m0 (:258).  w_0 = 1/N (:85).
     * chi0 == NULL with dist 1 keeps the Normal start (the library's mvt_normal_init switch). */
    orc_propagate(dist == 1 && chi0 ? 1 : 0, xa, NULL, NULL, NULL, m0, Q_c0, xi0, chi0, N, d);
    for (int64_t i = 0; i < N; ++i) w[i] = 1 / (double)N;
    if (x_hist) memcpy(x_hist, xa, sizeof(double) * Nd);
    if (w_hist) memcpy(w_hist, w, sizeof(double) * N);
    if (mean_hist) weighted_mean(xa, w, N, d, mean_hist);

    for (int t = 1; t < T; ++t) {
        size_t off = (size_t)(t - 1);
        orc_metropolis_hastings(a, w, u + off * N * B, j + off * N * B, N, B);
        orc_propagate(dist, xb, xa, a, G, NULL, Q_w, xi + off * Nd,
                      chi ? chi + off * Nd : NULL, N, d);
        /* faithful: determinant() + inverse() per particle, as ref: src/mcmc.cpp:193-215 does */
        orc_reweight(dist, w, Y + (size_t)t * dy, xb, F, V, nu, N, d, dy, faithful, 0);
        double *tmp = xa; xa = xb; xb = tmp;
        if (x_hist) memcpy(x_hist + (size_t)t * Nd, xa, sizeof(double) * Nd);
        if (w_hist) memcpy(w_hist + (size_t)t * N, w, sizeof(double) * N);
        if (a_hist) memcpy(a_hist + (size_t)t * N, a, sizeof(uint32_t) * N);
        if (mean_hist) weighted_mean(xa, w, N, d, mean_hist + (size_t)t * d);
    }
    free(xa); free(xb); free(w); free(a);
}

/* ========================================================================= */
/* extended section (no reference counterpart: parity unpinned)               */
/* ========================================================================= */

double orc_quadform_fma(const double *M, const double *c, const double *v,
                        int m, int d, int tri)
{
    double q = 0.0;
    for (int k = 0; k < m; ++k) {
        double z = c ? c[k] : 0.0;
        int jmax = tri ? (k < d - 1 ? k : d - 1) : d - 1;
        for (int j = 0; j <= jmax; ++j) z = fma(-M[(size_t)k * d + j], v[j], z);
        q = fma(z, z, q);
    }
    return q;
}

/* ---- deterministic exp / log -------------------------------------------- */
static double bits_to_double(uint64_t b) { double d; memcpy(&d, &b, 8); return d; }
static uint64_t double_to_bits(double d) { uint64_t b; memcpy(&b, &d, 8); return b; }

double orc_det_exp(double x)
{
    if (x != x) return x;
    if (x > 709.782712893384) return INFINITY;
    if (x < -745.2) return 0.0;
    const double LOG2E = 1.4426950408889634074;
    const double LN2_HI = 6.93147180369123816490e-01;   /* fdlibm split of ln 2 */
    const double LN2_LO = 1.90821492927058770002e-10;
    double kf = rint(x * LOG2E);
    double r = fma(kf, -LN2_HI, x);
    r = fma(kf, -LN2_LO, r);
    /* Taylor series of exp(r), |r| <= 0.347, degree 13, Horner with fma */
    double p = 1.0 / 6227020800.0;
    p = fma(p, r, 1.0 / 479001600.0);
    p = fma(p, r, 1.0 / 39916800.0);
    p = fma(p, r, 1.0 / 3628800.0);
    p = fma(p, r, 1.0 / 362880.0);
    p = fma(p, r, 1.0 / 40320.0);
    p = fma(p, r, 1.0 / 5040.0);
    p = fma(p, r, 1.0 / 720.0);
    p = fma(p, r, 1.0 / 120.0);
    p = fma(p, r, 1.0 / 24.0);
    p = fma(p, r, 1.0 / 6.0);
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    int k = (int)kf;
    if (k >= -1021 && k <= 1023)
        return p * bits_to_double((uint64_t)(k + 1023) << 52);
    if (k > 1023)      /* p * 2^k with k = 1024: split to stay finite when possible */
        return (p * bits_to_double((uint64_t)(k - 1 + 1023) << 52)) * 2.0;
    /* gradual underflow: two exact-then-rounded steps */
    double s2 = bits_to_double((uint64_t)(k + 1022 + 1023) << 52);   /* 2^(k+1022) >= 2^-54 */
    return (p * s2) * bits_to_double((uint64_t)1 << 52);             /* * 2^-1022 */
}

double orc_det_log(double x)
{
    if (x != x || x < 0.0) return NAN;
    if (x == 0.0) return -INFINITY;
    if (x == INFINITY) return x;
    const double LN2_HI = 6.93147180369123816490e-01;
    const double LN2_LO = 1.90821492927058770002e-10;
    int e = 0;
    uint64_t b = double_to_bits(x);
    if ((b >> 52) == 0) {                 /* subnormal: scale by 2^54 (exact) */
        x = x * 18014398509481984.0;
        b = double_to_bits(x);
        e = -54;
    }
    e += (int)(b >> 52) - 1023;
    uint64_t mant = b & 0x000FFFFFFFFFFFFFull;
    double m = bits_to_double(mant | 0x3FF0000000000000ull);   /* [1, 2) */
    if (m > 1.4142135623730951) { m = m * 0.5; e += 1; }       /* [sqrt2/2, sqrt2) */
    double f = m - 1.0;
    double s = f / (2.0 + f);
    double s2 = s * s;
    /* 2 atanh(s) = 2 s (1 + s^2/3 + s^4/5 + ... + s^22/23) */
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
    p = p * s2;                            /* series minus its leading 1 */
    double two_s = s + s;
    double lo = fma(two_s, p, (double)e * LN2_LO);
    return fma((double)e, LN2_HI, two_s + lo);
}

/* ---- Philox4x32-10 -------------------------------------------------------- */
void orc_philox4x32(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int round = 0; round < 10; ++round) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* ---- log-sum-exp / ESS ----------------------------------------------------- */
void orc_logsumexp_ess(const double *lw, int64_t N, double *lse, double *ess, double *lmax)
{
    double m = -INFINITY;
    for (int64_t i = 0; i < N; ++i) if (lw[i] > m) m = lw[i];
    double s1 = 0.0, s2 = 0.0;
    for (int64_t i = 0; i < N; ++i) {
        double e = exp(lw[i] - m);
        s1 += e;
        s2 += e * e;
    }
    if (lmax) *lmax = m;
    if (lse) *lse = m + log(s1);
    if (ess) *ess = s1 * s1 / s2;
}

/* ---- fixed-point weights ---------------------------------------------------- */
int orc_fixed_shift(int64_t N_global)
{
    int b = 0;
    while (((int64_t)1 << b) < N_global) ++b;
    return 61 - b;
}

static uint64_t fixed_one(double w, double wmax, int shift)
{
    if (!(w > 0.0) || !(w <= wmax)) return 0;   /* NaN, <= 0, +inf beyond wmax -> 0 */
    double ratio = w / wmax;                     /* (0, 1] */
    return (uint64_t)(ratio * bits_to_double((uint64_t)(shift + 1023) << 52));
}

uint64_t orc_fixed_weights(const double *w, int64_t N, double wmax, int shift, uint64_t *q)
{
    uint64_t total = 0;
    for (int64_t i = 0; i < N; ++i) {
        uint64_t v = fixed_one(w[i], wmax, shift);
        if (q) q[i] = v;
        total += v;
    }
    return total;
}

static double weights_max(const double *w, int64_t N)
{
    double m = 0.0;
    for (int64_t i = 0; i < N; ++i)
        if (w[i] > m && w[i] <= 1.7976931348623157e308) m = w[i];
    return m;
}

int orc_resample_systematic(const double *w, int64_t N, double u0, uint32_t *a)
{
    double wmax = weights_max(w, N);
    if (!(wmax > 0.0)) {
        for (int64_t i = 0; i < N; ++i) a[i] = (uint32_t)i;
        return 1;
    }
    int shift = orc_fixed_shift(N);
    uint64_t *C = (uint64_t *)malloc(sizeof(uint64_t) * N);
    uint64_t run = 0;
    for (int64_t i = 0; i < N; ++i) { run += fixed_one(w[i], wmax, shift); C[i] = run; }
    uint64_t T = run;
    uint64_t r0 = (uint64_t)(u0 * (double)T);
    if (r0 > T - 1) r0 = T - 1;
    int64_t jcur = 0;
    for (int64_t i = 0; i < N; ++i) {
        unsigned __int128 pos = (unsigned __int128)(uint64_t)i * T + r0;
        while ((unsigned __int128)C[jcur] * (uint64_t)N <= pos) ++jcur;
        a[i] = (uint32_t)jcur;
    }
    free(C);
    return 0;
}

int orc_resample_multinomial(const double *w, int64_t N, const double *u, uint32_t *a)
{
    double wmax = weights_max(w, N);
    if (!(wmax > 0.0)) {
        for (int64_t i = 0; i < N; ++i) a[i] = (uint32_t)i;
        return 1;
    }
    int shift = orc_fixed_shift(N);
    uint64_t *C = (uint64_t *)malloc(sizeof(uint64_t) * N);
    uint64_t run = 0;
    for (int64_t i = 0; i < N; ++i) { run += fixed_one(w[i], wmax, shift); C[i] = run; }
    uint64_t T = run;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < N; ++i) {
        uint64_t p = (uint64_t)(u[i] * (double)T);
        if (p > T - 1) p = T - 1;
        int64_t lo = 0, hi = N;            /* first j with C[j] > p */
        while (lo < hi) {
            int64_t mid = lo + ((hi - lo) >> 1);
            if (C[mid] <= p) lo = mid + 1; else hi = mid;
        }
        a[i] = (uint32_t)lo;
    }
    free(C);
    return 0;
}

/* Rejection resampler (extended; the unbiased relative of ref: src/samplers.cpp:21-35): k = i; attempt
 * n = 0, 1, ...: accept k if u_n <= w[k] / wmax, else k = j_n, with (u_n, j_n) the counter-based draw
 * the Metropolis resampler uses for (seed, step, i, n).  Capped at `cap` attempts. */
void orc_resample_rejection(uint32_t *a, const double *w, double wmax, int64_t N, uint64_t seed,
                            uint64_t step, int cap);

/* ---- independent MH chains -------------------------------------------------- */
static double whiten_q(const double *L, const double *rinv, const double *r, double *v, int d)
{
    double q = 0.0;
    for (int k = 0; k < d; ++k) {
        double acc = r[k];
        for (int j = 0; j < k; ++j) acc = fma(-A_(L, k, j, d), v[j], acc);
        v[k] = acc * rinv[k];
        q = fma(v[k], v[k], q);
    }
    return q;
}

/* Sum of 32 values by xor butterflies (16, 8, 4, 2, 1), the pairing of the kernel's warp sum. */
static double butterfly32(const double *a32)
{
    double a[32], b[32];
    memcpy(a, a32, sizeof a);
    for (int o = 16; o > 0; o >>= 1) {
        for (int l = 0; l < 32; ++l) b[l] = a[l] + a[l ^ o];
        memcpy(a, b, sizeof a);
    }
    return a[0];
}

/* Random-walk MH with the proposal x' = x + step L z, run in whitened coordinates
 * v = L^-1 (x - mu):  v' = v + step z,  q' = |v'|^2  (no counterpart in the reference; this
 * function DEFINES the arithmetic mh_chains_kernel reproduces bit for bit). */
void orc_mh_chains(int dist, int64_t C, int d, int steps, double step, double nu,
                   int shared, const double *mu, const double *L,
                   const double *x0, const double *z, const double *thr,
                   double *x_final, uint32_t *n_accept, uint8_t *accept_bits)
{
    double inv_nu = dist == 1 ? 1.0 / nu : 0.0;
#pragma omp parallel
    {
        double *r = (double *)malloc(sizeof(double) * d * 3);
        double *v = r + d, *rinv = r + 2 * d;
        double vv[32], vp[32], sq[32];
#pragma omp for schedule(static)
        for (int64_t c = 0; c < C; ++c) {
            const double *Lc = shared ? L : L + (size_t)c * d * d;
            const double *mc = shared ? mu : mu + (size_t)c * d;
            for (int k = 0; k < d; ++k) {
                rinv[k] = 1.0 / A_(Lc, k, k, d);
                r[k] = x0[(size_t)c * d + k] - mc[k];
            }
            (void)whiten_q(Lc, rinv, r, v, d);
            for (int k = 0; k < 32; ++k) {
                vv[k] = k < d ? v[k] : 0.0;
                sq[k] = vv[k] * vv[k];
            }
            double q = butterfly32(sq);
            uint32_t nacc = 0;
            for (int s = 0; s < steps; ++s) {
                const double *zs = z + ((size_t)c * steps + s) * d;
                for (int k = 0; k < 32; ++k) {
                    vp[k] = k < d ? fma(step, zs[k], vv[k]) : 0.0;
                    sq[k] = vp[k] * vp[k];
                }
                double qp = butterfly32(sq);
                double th = thr[(size_t)c * steps + s];
                int acc_flag;
                if (dist == 0) {
                    acc_flag = 0.5 * (qp - q) < th;
                } else {
                    double tp = fma(qp, inv_nu, 1.0);
                    double tc = fma(q, inv_nu, 1.0);
                    acc_flag = tp < th * tc;
                }
                if (acc_flag) {
                    memcpy(vv, vp, sizeof vv);
                    q = qp;
                    ++nacc;
                }
                if (accept_bits) accept_bits[(size_t)c * steps + s] = (uint8_t)acc_flag;
            }
            for (int k = 0; k < d; ++k) {          /* x = mu + L v, j ascending; untouched if it never moved */
                double acc = 0.0;
                for (int j = 0; j <= k; ++j) acc = fma(A_(Lc, k, j, d), vv[j], acc);
                x_final[(size_t)c * d + k] = nacc ? mc[k] + acc : x0[(size_t)c * d + k];
            }
            if (n_accept) n_accept[c] = nacc;
        }
        free(r);
    }
}

/* Random-walk MH whose proposal does not use the target's factor: x' = x + step * scale (.) z.  Every
 * step evaluates the target's quadratic form q' = |L^-1 (x' - mu)|^2 by forward substitution
 * (whiten_q_scaled: rows pre-scaled by 1 / L_kk, row k accumulates j ascending with fma, q k ascending) -- the
 * density arithmetic of ref: src/statistics.cc.cpp:295-311 in whitened form -- and applies the accept
 * rule of orc_mh_chains (ref: src/samplers.cpp:30 on the density ratio).  No counterpart in the
 * reference; defines what mh_general_kernel reproduces bit for bit. */
/* forward substitution with rows pre-scaled by 1 / L_kk: Ls[k][j] = L[k][j] * rinv[k], r~_k = r_k * rinv[k];
 * v_k = r~_k - sum_{j<k} Ls[k][j] v_j (fma, j ascending), q = sum v_k^2 (fma, k ascending). */
static double whiten_q_scaled(const double *Ls, const double *rs, double *v, int d)
{
    double q = 0.0;
    for (int k = 0; k < d; ++k) {
        double acc = rs[k];
        for (int j = 0; j < k; ++j) acc = fma(-Ls[(size_t)k * d + j], v[j], acc);
        v[k] = acc;
        q = fma(v[k], v[k], q);
    }
    return q;
}

void orc_mh_chains_general(int dist, int64_t C, int d, int steps, double step, const double *scale, double nu,
                           int shared, const double *mu, const double *L, const double *x0, const double *z,
                           const double *thr, double *x_final, uint32_t *n_accept, uint8_t *accept_bits)
{
    double inv_nu = dist == 1 ? 1.0 / nu : 0.0;
#pragma omp parallel
    {
        double *buf = (double *)malloc(sizeof(double) * ((size_t)d * 5 + (size_t)d * d));
        double *r = buf, *v = buf + d, *rinv = buf + 2 * d, *x = buf + 3 * d, *xp = buf + 4 * d, *Ls = buf + 5 * d;
#pragma omp for schedule(static)
        for (int64_t c = 0; c < C; ++c) {
            const double *Lc = shared ? L : L + (size_t)c * d * d;
            const double *mc = shared ? mu : mu + (size_t)c * d;
            for (int k = 0; k < d; ++k) {
                rinv[k] = 1.0 / A_(Lc, k, k, d);
                for (int j = 0; j < k; ++j) Ls[(size_t)k * d + j] = A_(Lc, k, j, d) * rinv[k];
                x[k] = x0[(size_t)c * d + k];
                r[k] = (x[k] - mc[k]) * rinv[k];
            }
            double q = whiten_q_scaled(Ls, r, v, d);
            uint32_t nacc = 0;
            for (int s = 0; s < steps; ++s) {
                const double *zs = z + ((size_t)c * steps + s) * d;
                for (int k = 0; k < d; ++k) {
                    double sk = step * (scale ? scale[k] : 1.0);
                    xp[k] = fma(sk, zs[k], x[k]);
                    r[k] = (xp[k] - mc[k]) * rinv[k];
                }
                double qp = whiten_q_scaled(Ls, r, v, d);
                double th = thr[(size_t)c * steps + s];
                int acc_flag;
                if (dist == 0) {
                    acc_flag = 0.5 * (qp - q) < th;
                } else {
                    double tp = fma(qp, inv_nu, 1.0);
                    double tc = fma(q, inv_nu, 1.0);
                    acc_flag = tp < th * tc;
                }
                if (acc_flag) {
                    memcpy(x, xp, sizeof(double) * d);
                    q = qp;
                    ++nacc;
                }
                if (accept_bits) accept_bits[(size_t)c * steps + s] = (uint8_t)acc_flag;
            }
            memcpy(x_final + (size_t)c * d, x, sizeof(double) * d);
            if (n_accept) n_accept[c] = nacc;
        }
        free(buf);
    }
}

/* ========================================================================= */
/* production-order restatements                                             */
/* ========================================================================= */
int orc_observation_operator(int dist, int d, int dy, const double *F, const double *V, float nu,
                             double *M, double *Winv, double *lognorm)
{
    double *L = (double *)malloc(sizeof(double) * dy * dy);
    double *W = (double *)malloc(sizeof(double) * dy * dy);
    int bad = orc_cholesky_lower(V, dy, L);
    if (bad) { free(L); free(W); return bad; }
    orc_tri_inverse_lower(L, dy, W);
    for (int k = 0; k < dy; ++k)
        for (int j = 0; j < dy; ++j) Winv[(size_t)k * dy + j] = j <= k ? A_(W, k, j, dy) : 0.0;
    for (int k = 0; k < dy; ++k)
        for (int j = 0; j < d; ++j) {
            double s = 0.0;
            for (int i = 0; i <= k; ++i) s += Winv[(size_t)k * dy + i] * A_(F, i, j, dy);
            M[(size_t)k * d + j] = s;
        }
    double sl = 0.0;
    for (int k = 0; k < dy; ++k) sl += log(A_(L, k, k, dy));
    double logdet = 2.0 * sl;
    if (dist == 0) {
        *lognorm = -(0.5 * dy * log(2.0 * M_PI) + 0.5 * logdet);
    } else {
        float s = nu + (float)(unsigned)dy;
        *lognorm = -0.5 * dy * log(M_PI * nu) - 0.5 * logdet + lgamma(0.5 * s) - lgamma(0.5 * nu);
    }
    free(L); free(W);
    return 0;
}

void orc_step_det(int dist, int want_log, double *x_new, double *lw, const double *x_prev,
                  const uint32_t *a, const double *G, const double *Q, const double *mu,
                  const double *M, const double *c, double lognorm, float nu,
                  const double *xi, const double *chi, int64_t N, int d, int dy)
{
    float nu_d = nu + (float)(unsigned)dy;
    double half_nu_d = 0.5 * nu_d, inv_nu = dist == 1 ? 1.0 / (double)nu : 0.0;
    double scale = exp(lognorm);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < N; ++i) {
        const double *xp = x_prev ? x_prev + (size_t)(a ? a[i] : (uint32_t)i) * d : NULL;
        const double *z = xi + (size_t)i * d;
        double *xn = x_new + (size_t)i * d;
        for (int k = 0; k < d; ++k) {
            double g = mu ? mu[k] : 0.0;
            if (G && xp) for (int j = 0; j < d; ++j) g = fma(A_(G, k, j, d), xp[j], g);
            double s = 0.0;
            for (int j = 0; j < d; ++j) s = fma(A_(Q, k, j, d), z[j], s);
            if (dist == 1 && chi) s = chi[(size_t)i * d + k] * s;
            xn[k] = s + g;
        }
        if (!lw) continue;
        double q = 0.0;
        for (int k = 0; k < dy; ++k) {
            double zk = c[k];
            for (int j = 0; j < d; ++j) zk = fma(-M[(size_t)k * d + j], xn[j], zk);
            q = fma(zk, zk, q);
        }
        if (dist == 0)
            lw[i] = want_log ? fma(-0.5, q, lognorm) : scale * exp(-0.5 * q);
        else
            lw[i] = want_log ? fma(-half_nu_d, log1p(q * inv_nu), lognorm)
                             : scale * pow(fma(q, inv_nu, 1.0), -half_nu_d);
    }
}

void orc_det_sincospi(double t, double *s_out, double *c_out)
{
    double nf = rint(t + t);
    double f = fma(nf, -0.5, t);
    double x = fma(f, 3.141592653589793116, f * 1.2246467991473532e-16);
    double x2 = x * x;
    static const double SC[9] = { -1.0 / 121645100408832000.0, 1.0 / 355687428096000.0,
        -1.0 / 1307674368000.0, 1.0 / 6227020800.0, -1.0 / 39916800.0, 1.0 / 362880.0,
        -1.0 / 5040.0, 1.0 / 120.0, -1.0 / 6.0 };
    static const double CC[10] = { 1.0 / 2432902008176640000.0, -1.0 / 6402373705728000.0,
        1.0 / 20922789888000.0, -1.0 / 87178291200.0, 1.0 / 479001600.0, -1.0 / 3628800.0,
        1.0 / 40320.0, -1.0 / 720.0, 1.0 / 24.0, -0.5 };
    double ps = SC[0];
    for (int i = 1; i < 9; ++i) ps = fma(ps, x2, SC[i]);
    double sn = fma(x * x2, ps, x);
    double pc = CC[0];
    for (int i = 1; i < 10; ++i) pc = fma(pc, x2, CC[i]);
    double cs = fma(x2, pc, 1.0);
    int n = ((int)nf) & 3;
    *s_out = n == 0 ? sn : n == 1 ? cs : n == 2 ? -sn : -cs;
    *c_out = n == 0 ? cs : n == 1 ? -sn : n == 2 ? -cs : sn;
}

static void rng_block(uint64_t seed, int stream, uint64_t step, uint64_t index, uint32_t sub, uint32_t out[4])
{
    uint32_t ctr[4] = { (uint32_t)index, (uint32_t)(index >> 32), (uint32_t)step,
                        (uint32_t)stream | (sub << 8) };
    uint32_t key[2] = { (uint32_t)seed, (uint32_t)(seed >> 32) };
    orc_philox4x32(ctr, key, out);
}

static double u01_from(uint32_t hi, uint32_t lo, int open0)
{
    uint64_t bits = ((((uint64_t)hi << 32) | lo) >> 11) + (open0 ? 1 : 0);
    return (double)bits * 1.1102230246251565e-16;
}

void orc_rng_normal_pair(uint64_t seed, int stream, uint64_t step, uint64_t index, uint32_t sub, double z[2])
{
    uint32_t r[4];
    rng_block(seed, stream, step, index, sub, r);
    double u1 = u01_from(r[0], r[1], 1), u2 = u01_from(r[2], r[3], 0);
    double rad = sqrt(-2.0 * orc_det_log(u1));
    double s, c;
    orc_det_sincospi(u2 + u2, &s, &c);
    z[0] = rad * c;
    z[1] = rad * s;
}

double orc_rng_u01(uint64_t seed, int stream, uint64_t step, uint64_t index, uint32_t sub)
{
    uint32_t r[4];
    rng_block(seed, stream, step, index, sub, r);
    return u01_from(r[0], r[1], 0);
}

void orc_rng_metropolis(uint64_t seed, uint64_t step, uint64_t index, uint32_t n, uint64_t N,
                        double *u, uint32_t *j)
{
    uint32_t r[4];
    rng_block(seed, 0, step, index, n, r);
    *u = u01_from(r[0], r[1], 0);
    uint64_t bits = ((uint64_t)r[2] << 32) | r[3];
    *j = (uint32_t)(((unsigned __int128)bits * N) >> 64);
}

/* Metropolis-C2 proposals (include/cusmc_b200.h, cusmc_metropolis_c2_dev): the segment comes from the
 * counter of the particle's group of 32, the slot inside it and the uniform from the particle's own. */
void orc_rng_metropolis_c2(uint64_t seed, uint64_t step, uint64_t index, uint32_t n, uint64_t N,
                           double *u, uint32_t *j)
{
    uint32_t r[4], rs[4];
    rng_block(seed, 0, step, index, n, r);
    rng_block(seed, 8, step, index >> 5, n, rs);
    *u = u01_from(r[0], r[1], 0);
    uint64_t sbits = ((uint64_t)rs[0] << 32) | rs[1];
    uint32_t first = (uint32_t)(((unsigned __int128)sbits * N) >> 64) & ~31u;
    uint64_t len = N - first < 32 ? N - first : 32;
    uint64_t bits = ((uint64_t)r[2] << 32) | r[3];
    *j = first + (uint32_t)(((unsigned __int128)bits * len) >> 64);
}

/* Single-precision deterministic log / sincospi / Box-Muller: mirror of the kernels' default
 * normal generator (include/cusmc_detmath.h, cusmc_philox.h). */
static float bits_to_float(uint32_t b) { float f; memcpy(&f, &b, 4); return f; }
static uint32_t float_to_bits(float f) { uint32_t b; memcpy(&b, &f, 4); return b; }

float orc_det_logf(float x)
{
    uint32_t b = float_to_bits(x);
    int e = (int)(b >> 23) - 127;
    float m = bits_to_float((b & 0x007FFFFFu) | 0x3F800000u);
    if (m > 1.41421354f) { m = m * 0.5f; e += 1; }
    float f = m - 1.0f;
    static const float P[8] = { 0.09004202485084534f, -0.14257794618606567f, 0.14806459844112396f,
        -0.16575047373771667f, 0.19973105192184448f, -0.25001609325408936f, 0.33333659172058105f,
        -0.4999999403953552f };
    float p = P[0];
    for (int i = 1; i < 8; ++i) p = fmaf(p, f, P[i]);
    float r = fmaf(p * f, f, f);
    return fmaf((float)e, 0.693147182464599609375f, r);
}

void orc_det_sincospif(float t, float *s_out, float *c_out)
{
    float nf = rintf(t + t);
    float f = fmaf(nf, -0.5f, t);
    float x = f * 3.14159274101257324f;
    float x2 = x * x;
    float ps = 2.75573192e-6f;
    ps = fmaf(ps, x2, -1.98412698e-4f);
    ps = fmaf(ps, x2, 8.33333377e-3f);
    ps = fmaf(ps, x2, -1.66666672e-1f);
    float sn = fmaf(x * x2, ps, x);
    float pc = 2.48015876e-5f;
    pc = fmaf(pc, x2, -1.38888892e-3f);
    pc = fmaf(pc, x2, 4.16666679e-2f);
    pc = fmaf(pc, x2, -0.5f);
    float cs = fmaf(x2, pc, 1.0f);
    int n = ((int)nf) & 3;
    *s_out = n == 0 ? sn : n == 1 ? cs : n == 2 ? -sn : -cs;
    *c_out = n == 0 ? cs : n == 1 ? -sn : n == 2 ? -cs : sn;
}

static void box_muller_f32(uint32_t a, uint32_t b, float *z0, float *z1)
{
    float u1 = fmaf((float)a, 2.3283064365386963e-10f, 1.16415321826934814e-10f);
    if (u1 > 1.0f) u1 = 1.0f;
    float rad = sqrtf(-2.0f * orc_det_logf(u1));
    float s, c;
    orc_det_sincospif((float)(b >> 8) * 1.1920928955078125e-7f, &s, &c);
    *z0 = rad * c;
    *z1 = rad * s;
}

void orc_rng_normal4(uint64_t seed, int stream, uint64_t step, uint64_t index, uint32_t sub, double z[4])
{
    uint32_t r[4];
    float a0, a1, b0, b1;
    rng_block(seed, stream, step, index, sub, r);
    box_muller_f32(r[0], r[1], &a0, &a1);
    box_muller_f32(r[2], r[3], &b0, &b1);
    z[0] = a0; z[1] = a1; z[2] = b0; z[3] = b1;
}

void orc_rng_fill_normals(uint64_t seed, int stream, uint64_t step, int64_t i0, int64_t N, int d, double *xi)
{
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < N; ++i)
        for (int jq = 0; 4 * jq < d; ++jq) {
            double z[4];
            orc_rng_normal4(seed, stream, step, (uint64_t)(i0 + i), (uint32_t)jq, z);
            for (int e = 0; e < 4 && 4 * jq + e < d; ++e) xi[(size_t)i * d + 4 * jq + e] = z[e];
        }
}

void orc_resample_rejection(uint32_t *a, const double *w, double wmax, int64_t N, uint64_t seed,
                            uint64_t step, int cap)
{
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t i = 0; i < N; ++i) {
        uint32_t k = (uint32_t)i;
        double wk = w[i];
        for (int n = 0; n < cap; ++n) {
            double u;
            uint32_t j;
            orc_rng_metropolis(seed, step, (uint64_t)i, (uint32_t)n, (uint64_t)N, &u, &j);
            if (u <= wk / wmax) break;
            k = j;
            wk = w[k];
        }
        a[i] = k;
    }
}

static uint64_t fixed_from_unit(double wn, int shift)
{
    if (!(wn > 0.0)) return 0;
    if (wn > 1.0) wn = 1.0;
    return (uint64_t)(wn * bits_to_double((uint64_t)(shift + 1023) << 52));
}

/* ---- block-relative weight image (extended; no reference counterpart) -----------------------
 * The filter's normalised resamplers take their weights from this image.  Particles are cut into
 * tiles of `tile` consecutive slots (global index space).  Inside tile b
 *     m_b = max finite lw,   q_i = trunc(exp(lw_i - m_b) 2^shift),   c_i = inclusive prefix of q,
 *     S_b = c_last,          S2_b = sum trunc(exp(lw_i - m_b)^2 2^shift),
 * so a tile is complete without knowing anything about the others.  With M = max_b m_b the tile is
 * rescaled by the 62-bit fixed-point factor F_b = trunc(exp(m_b - M) 2^62):
 *     C_i = P_b + (c_i F_b >> 62),   P_b = sum_{b' < b} (S_b' F_b' >> 62),   T = sum_b (S_b F_b >> 62),
 *     T2  = sum_b ((S2_b F_b >> 62) F_b >> 62)
 * (128-bit products).  C is the integer CDF systematic / multinomial resampling read; ESS =
 * T^2 / (T2 2^shift), log-likelihood = M + log(T / 2^shift / N).  exp is orc_det_exp, flushed to 0
 * below -43.5 (anything that small truncates to 0 anyway).  The kernels (pf_fused_kernel epilogue,
 * tile_update_kernel) reproduce every one of these integers. */
static uint64_t mulshift62(uint64_t c, uint64_t F) { return (uint64_t)(((unsigned __int128)c * F) >> 62); }

static double exp_unit(double x) { return (x >= -43.5) ? orc_det_exp(x) : 0.0; }   /* NaN -> 0 */

uint64_t orc_tile_image(const double *lw, int64_t N, int64_t tile, int shift, uint64_t *C,
                        uint64_t *T2_out, double *M_out)
{
    int64_t nt = (N + tile - 1) / tile;
    double *m = (double *)malloc(sizeof(double) * (nt > 0 ? nt : 1));
    uint64_t *S = (uint64_t *)malloc(sizeof(uint64_t) * (nt > 0 ? nt : 1));
    uint64_t *S2 = (uint64_t *)malloc(sizeof(uint64_t) * (nt > 0 ? nt : 1));
    double M = -INFINITY;
#pragma omp parallel for schedule(static) reduction(max : M)
    for (int64_t b = 0; b < nt; ++b) {
        int64_t lo = b * tile, hi = lo + tile < N ? lo + tile : N;
        double mb = -INFINITY;
        for (int64_t i = lo; i < hi; ++i) if (lw[i] > mb && lw[i] < INFINITY) mb = lw[i];
        uint64_t run = 0, s2 = 0;
        for (int64_t i = lo; i < hi; ++i) {
            double wn = (lw[i] <= mb) ? exp_unit(lw[i] - mb) : 0.0;
            run += fixed_from_unit(wn, shift);
            s2 += fixed_from_unit(wn * wn, shift);
            C[i] = run;                          /* tile-local for now */
        }
        m[b] = mb; S[b] = run; S2[b] = s2;
        if (mb > M) M = mb;
    }
    uint64_t P = 0, T2 = 0;
    const double two62 = 4611686018427387904.0;
    for (int64_t b = 0; b < nt; ++b) {
        int64_t lo = b * tile, hi = lo + tile < N ? lo + tile : N;
        double f = (m[b] <= M) ? exp_unit(m[b] - M) : 0.0;
        uint64_t F = (uint64_t)(f * two62);
        for (int64_t i = lo; i < hi; ++i) C[i] = P + mulshift62(C[i], F);
        P += mulshift62(S[b], F);
        T2 += mulshift62(mulshift62(S2[b], F), F);
    }
    free(m); free(S); free(S2);
    if (T2_out) *T2_out = T2;
    if (M_out) *M_out = M;
    return P;
}

int orc_filter_det(int dist, int resampler, int64_t N, int d, int dy, int T, int B,
                   const double *Y, const double *m0, const double *Q_c0, const double *F,
                   const double *G, const double *V, const double *Q_w, float nu, uint64_t seed,
                   const double *xi0, const double *chi0, const double *xi, const double *chi, const double *u,
                   const uint32_t *j, const double *u0, const double *um,
                   double *x_hist, double *w_hist, uint32_t *a_hist, double *ess, double *loglik,
                   double ess_threshold, int *resampled, int64_t tile)
{
    if (tile <= 0) tile = 2048;       /* the library's tile (kTile); the persistent kernel passes its own */
    size_t Nd = (size_t)N * d;
    double *w_old = (double *)malloc(sizeof(double) * N);
    int is_log = resampler != 0 && resampler != 3 && resampler != 4;
    double *xa = (double *)malloc(sizeof(double) * Nd), *xb = (double *)malloc(sizeof(double) * Nd);
    double *w = (double *)malloc(sizeof(double) * N), *noise = (double *)malloc(sizeof(double) * Nd);
    uint32_t *a = (uint32_t *)malloc(sizeof(uint32_t) * N);
    uint64_t *C = (uint64_t *)malloc(sizeof(uint64_t) * N);
    double *M = (double *)malloc(sizeof(double) * dy * d), *Winv = (double *)malloc(sizeof(double) * dy * dy);
    double *c = (double *)malloc(sizeof(double) * dy), *ub = (double *)malloc(sizeof(double) * N * (B > 0 ? B : 1));
    uint32_t *jb = (uint32_t *)malloc(sizeof(uint32_t) * N * (B > 0 ? B : 1));
    double lognorm;
    int rc = orc_observation_operator(dist, d, dy, F, V, nu, M, Winv, &lognorm);
    int shift = orc_fixed_shift(N);
    double scale2 = bits_to_double((uint64_t)(shift + 1023) << 52);
    if (rc) goto done;

    /* t = 0 */
    {
        const double *z0 = xi0;
        if (!z0) { orc_rng_fill_normals(seed, 6, 0, 0, N, d, noise); z0 = noise; }
        /* "mvt": the initial draw is chi (.) (Q_c0 xi) + m0 too (ref: src/mcmc.cpp:73-79 ->
         * src/statistics.cc.cpp:411); chi0 == NULL keeps the Normal start */
        orc_step_det(dist == 1 && chi0 ? 1 : 0, 0, xa, NULL, NULL, NULL, NULL, Q_c0, m0, NULL, NULL, 0.0, nu, z0,
                     chi0, N, d, dy);
        for (int64_t i = 0; i < N; ++i) w[i] = is_log ? 0.0 : 1.0 / (double)N;
    }
    for (int t = 0; t < T; ++t) {
        int do_resample = t > 0;
        if (t > 0) {
            size_t off = (size_t)(t - 1);
            if (resampler == 0 || resampler == 4) {
                const double *ut = u ? u + off * N * B : ub;
                const uint32_t *jt = j ? j + off * N * B : jb;
                if (!u)
                    for (int64_t i = 0; i < N; ++i)
                        for (int n = 0; n < B; ++n)
                            (resampler == 4 ? orc_rng_metropolis_c2 : orc_rng_metropolis)(
                                seed, (uint64_t)t, (uint64_t)i, (uint32_t)n, (uint64_t)N,
                                &ub[(size_t)i * B + n], &jb[(size_t)i * B + n]);
                orc_metropolis_hastings(a, w, ut, jt, N, B);
            } else if (resampler == 3) {
                double m = -INFINITY;           /* the kernels' atomic max over finite weights */
                for (int64_t i = 0; i < N; ++i) if (w[i] == w[i] && w[i] < INFINITY && w[i] > m) m = w[i];
                orc_resample_rejection(a, w, m, N, seed, (uint64_t)t, 4096);
            } else {
                /* weights of step t-1: the block-relative fixed-point image, integer CDF */
                uint64_t run2 = 0;
                uint64_t Tm = orc_tile_image(w, N, tile, shift, C, &run2, NULL);
                if (Tm == 0) { rc = -1; goto done; }
                /* adaptive resampling: ESS = sum_q^2 / (sum_q2 2^shift) < threshold N, evaluated as
                 * the kernels do (scan_resample_kernel) */
                do_resample = 1;
                if (ess_threshold > 0.0)
                    do_resample = (double)Tm * (double)Tm < (ess_threshold * (double)N * scale2) * (double)run2;
                if (!do_resample) {
                    for (int64_t i = 0; i < N; ++i) a[i] = (uint32_t)i;
                } else if (resampler == 1) {
                    double uu = u0 ? u0[off] : 0.0;
                    if (!u0) {
                        uint32_t r[4];
                        rng_block(seed, 7, (uint64_t)t, 0, 0, r);
                        uu = (double)((((uint64_t)r[0] << 32) | r[1]) >> 11) * 1.1102230246251565e-16;
                    }
                    uint64_t r0 = (uint64_t)(uu * (double)Tm);
                    if (r0 > Tm - 1) r0 = Tm - 1;
                    int64_t jc = 0;
                    for (int64_t i = 0; i < N; ++i) {
                        unsigned __int128 pos = (unsigned __int128)(uint64_t)i * Tm + r0;
                        while ((unsigned __int128)C[jc] * (uint64_t)N <= pos) ++jc;
                        a[i] = (uint32_t)jc;
                    }
                } else {
                    for (int64_t i = 0; i < N; ++i) {
                        double uu = um ? um[off * N + i] : orc_rng_u01(seed, 3, (uint64_t)t, (uint64_t)i, 0);
                        uint64_t p = (uint64_t)(uu * (double)Tm);
                        if (p > Tm - 1) p = Tm - 1;
                        int64_t lo = 0, hi = N;
                        while (lo < hi) { int64_t mid = lo + ((hi - lo) >> 1); if (C[mid] <= p) lo = mid + 1; else hi = mid; }
                        a[i] = (uint32_t)lo;
                    }
                }
            }
            const double *zt = xi ? xi + off * Nd : noise;
            if (!xi) orc_rng_fill_normals(seed, 1, (uint64_t)t, 0, N, d, noise);
            for (int k = 0; k < dy; ++k) {
                double s = 0.0;
                for (int i = 0; i <= k; ++i) s += Winv[(size_t)k * dy + i] * Y[(size_t)t * dy + i];
                c[k] = s;
            }
            if (!do_resample) memcpy(w_old, w, sizeof(double) * N);
            orc_step_det(dist, is_log, xb, w, xa, a, G, Q_w, NULL, M, c, lognorm, nu, zt,
                         chi ? chi + off * Nd : NULL, N, d, dy);
            if (!do_resample)            /* no resampling: the log-weights accumulate */
                for (int64_t i = 0; i < N; ++i) w[i] = w_old[i] + w[i];
            double *tmp = xa; xa = xb; xb = tmp;
        }
        if (resampled) resampled[t] = do_resample;
        if (x_hist) memcpy(x_hist + (size_t)t * Nd, xa, sizeof(double) * Nd);
        if (w_hist) memcpy(w_hist + (size_t)t * N, w, sizeof(double) * N);
        if (a_hist && t > 0) memcpy(a_hist + (size_t)t * N, a, sizeof(uint32_t) * N);
        if (is_log && (ess || loglik)) {
            double m;
            uint64_t s2 = 0;
            uint64_t s1 = orc_tile_image(w, N, tile, shift, C, &s2, &m);
            if (ess) ess[t] = ((double)s1 * (double)s1) / ((double)s2 * scale2);
            if (loglik) loglik[t] = m + log((double)s1 / scale2 / (double)N);
        }
    }
done:
    free(w_old);
    free(xa); free(xb); free(w); free(noise); free(a); free(C); free(M); free(Winv); free(c); free(ub); free(jb);
    return rc;
}

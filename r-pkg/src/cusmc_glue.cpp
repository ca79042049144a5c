// cusmc_glue.cpp -- the R-facing entry points of CuSMC over the C ABI of libcusmc_b200.so.
//
// Same six exports, argument order and return shapes as the reference package
// (ref: src/mvn_dist.rcpp.cpp:31-58, src/mvt_dist.rcpp.cpp:28-66, src/samplers.rcpp.cpp:35-55,
// src/run.rcpp.cpp:58-126, R/RcppExports.R:17-103); everything below the signature is a call into
// include/cusmc_b200.h.  Plain Rcpp types only: R matrices are column-major, which is what the ABI
// takes (a d x N matrix of points is the AoS layout, leading dimension d).
//
// Not built or tested in the development image (it has no R): tests/test_rpkg_cpu.py compiles this
// file against a minimal stand-in for <Rcpp.h> so that it at least parses and type-checks against
// the current cusmc_b200.h.
#include <Rcpp.h>

#include <cmath>
#include <cstdint>
#include <string>
#include <vector>

#include "cusmc_b200.h"

namespace {

// One context per R session (R is single threaded); device 0 unless CUSMC_DEVICE says otherwise.
cusmc_ctx *ctx()
{
    static cusmc_ctx *c = nullptr;
    if (!c) {
        const char *dev = std::getenv("CUSMC_DEVICE");
        if (cusmc_ctx_create(&c, dev ? std::atoi(dev) : 0) != CUSMC_OK)
            Rcpp::stop("CuSMC: no usable CUDA device (this build has no CPU fallback)");
    }
    return c;
}

// Non-zero status -> an R error carrying the library's message (ref: FATAL(), inst/include/support.cuh:22-32).
void check(int rc)
{
    if (rc != CUSMC_OK) Rcpp::stop("CuSMC: %s", cusmc_last_error(ctx()));
}

// 64 bits from R's generator, so set.seed() makes a session reproducible.
uint64_t session_seed()
{
    Rcpp::RNGScope scope;
    const uint64_t hi = (uint64_t)(R::unif_rand() * 4294967296.0);
    const uint64_t lo = (uint64_t)(R::unif_rand() * 4294967296.0);
    return (hi << 32) | lo;
}

// Q = V sqrt(Lambda) of a symmetric matrix (ref: eigenSolver, src/linear_algebra.cpp:10-23), cyclic
// Jacobi; column-major d x d in and out.  Host side by design: d <= 32, once per call.
std::vector<double> eigen_factor(const double *S, int d)
{
    std::vector<double> A(S, S + (size_t)d * d), V((size_t)d * d, 0.0);
    for (int i = 0; i < d; ++i) V[(size_t)i * d + i] = 1.0;
    auto a = [&](int r, int c) -> double & { return A[(size_t)c * d + r]; };
    auto v = [&](int r, int c) -> double & { return V[(size_t)c * d + r]; };
    for (int sweep = 0; sweep < 64; ++sweep) {
        double off = 0.0;
        for (int p = 0; p < d; ++p)
            for (int q = p + 1; q < d; ++q) off += a(p, q) * a(p, q);
        if (off < 1e-300) break;
        for (int p = 0; p < d; ++p)
            for (int q = p + 1; q < d; ++q) {
                if (a(p, q) == 0.0) continue;
                const double theta = (a(q, q) - a(p, p)) / (2.0 * a(p, q));
                const double t = (theta >= 0.0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
                const double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < d; ++k) {
                    const double akp = a(k, p), akq = a(k, q);
                    a(k, p) = c * akp - s * akq;
                    a(k, q) = s * akp + c * akq;
                }
                for (int k = 0; k < d; ++k) {
                    const double apk = a(p, k), aqk = a(q, k);
                    a(p, k) = c * apk - s * aqk;
                    a(q, k) = s * apk + c * aqk;
                }
                for (int k = 0; k < d; ++k) {
                    const double vkp = v(k, p), vkq = v(k, q);
                    v(k, p) = c * vkp - s * vkq;
                    v(k, q) = s * vkp + c * vkq;
                }
            }
    }
    for (int c = 0; c < d; ++c) {
        const double lam = a(c, c) > 0.0 ? std::sqrt(a(c, c)) : 0.0;
        for (int r = 0; r < d; ++r) v(r, c) *= lam;
    }
    return V;
}

void require_square(const Rcpp::NumericMatrix &m, int d, const char *name)
{
    if (m.nrow() != d || m.ncol() != d) Rcpp::stop("CuSMC: %s must be %d x %d", name, d, d);
}

// Density of one point (numeric vector -> scalar, as the reference) or of the columns of a d x N
// matrix (-> numeric vector of length N; batched extension).
SEXP density(int kind, SEXP x, const Rcpp::NumericVector &mu, const Rcpp::NumericMatrix &sigma, float nu)
{
    const int d = (int)mu.size();
    require_square(sigma, d, "sigma");
    if (Rf_isMatrix(x)) {
        const Rcpp::NumericMatrix X = Rcpp::as<Rcpp::NumericMatrix>(x);
        if (X.nrow() != d) Rcpp::stop("CuSMC: x must have %d rows (one point per column)", d);
        const int64_t N = X.ncol();
        Rcpp::NumericVector out((int)N);
        check(cusmc_logpdf(ctx(), kind, /*want_log=*/0, X.begin(), CUSMC_AOS, N, N, d, mu.begin(), sigma.begin(), nu,
                           out.begin()));
        return out;
    }
    const Rcpp::NumericVector xv = Rcpp::as<Rcpp::NumericVector>(x);
    if ((int)xv.size() != d) Rcpp::stop("CuSMC: length(x) != length(mu)");
    double out = 0.0;
    check(cusmc_logpdf(ctx(), kind, 0, xv.begin(), CUSMC_AOS, 1, 1, d, mu.begin(), sigma.begin(), nu, &out));
    return Rcpp::wrap(out);
}

}  // namespace

// One draw x = mu + Q xi.  As in the reference's CPU build the matrix handed in is used as the factor
// itself and its sampler's draws have variance 3: Q = sqrt(3) sigma.
// [[Rcpp::export]]
Rcpp::NumericVector MVN(Rcpp::NumericVector mu, Rcpp::NumericMatrix sigma)
{
    const int d = (int)mu.size();
    require_square(sigma, d, "sigma");
    std::vector<double> Q(sigma.begin(), sigma.end());
    for (double &q : Q) q *= std::sqrt(3.0);
    Rcpp::NumericVector x(d);
    check(cusmc_mvn_sample_init(ctx(), x.begin(), mu.begin(), Q.data(), /*xi=*/nullptr, session_seed(), 1, d));
    return x;
}

// [[Rcpp::export]]
SEXP MVNPDF(SEXP x, Rcpp::NumericVector mu, Rcpp::NumericMatrix sigma)
{
    return density(CUSMC_MVN, x, mu, sigma, 0.f);
}

// One draw mu + chi (.) (Q xi), chi_k = sqrt(nu / chi2_nu) per component, Q = V sqrt(Lambda).
// [[Rcpp::export]]
Rcpp::NumericVector MVT(Rcpp::NumericVector mu, Rcpp::NumericMatrix sigma, float nu)
{
    const int d = (int)mu.size();
    require_square(sigma, d, "sigma");
    std::vector<double> Q = eigen_factor(sigma.begin(), d);
    for (double &q : Q) q *= std::sqrt(3.0);
    const std::vector<double> zero_x((size_t)d, 0.0), zero_G((size_t)d * d, 0.0);
    Rcpp::NumericVector x(d);
    check(cusmc_mvt_sample(ctx(), x.begin(), zero_x.data(), /*a=*/nullptr, zero_G.data(), Q.data(), nullptr, nullptr,
                           session_seed(), /*step=*/1, 1, d, nu));
    for (int k = 0; k < d; ++k) x[k] += mu[k];
    return x;
}

// [[Rcpp::export]]
SEXP MVTPDF(SEXP x, Rcpp::NumericVector mu, Rcpp::NumericMatrix sigma, float nu)
{
    return density(CUSMC_MVT, x, mu, sigma, nu);
}

// Metropolis ancestor selection, B accept/reject steps per particle; 0-based ancestors as doubles.
// [[Rcpp::export]]
Rcpp::NumericVector metropolis_hastings(Rcpp::NumericVector w, int N, int B)
{
    if ((int)w.size() != N) Rcpp::stop("CuSMC: length(w) != N");
    std::vector<uint32_t> a((size_t)N);
    check(cusmc_metropolis_hastings(ctx(), a.data(), w.begin(), /*u=*/nullptr, /*j=*/nullptr, session_seed(),
                                    /*step=*/1, N, B));
    Rcpp::NumericVector out(N);
    for (int i = 0; i < N; ++i) out[i] = (double)a[(size_t)i];
    return out;
}

// The whole filter on the device; weights (timeSteps x N) and posterior_x (timeSteps x N x d) come back
// in the reference's shapes.  df reaches the Student-t noise as given (the reference passes it in the
// position of `runtime`); p (the particle whose path the reference writes to CSV) is accepted
// and ignored -- CSV output is not part of this package.
// [[Rcpp::export]]
Rcpp::List run(unsigned N, unsigned d, unsigned timeSteps, Rcpp::NumericMatrix Y, Rcpp::NumericVector m0,
               Rcpp::NumericMatrix C0, Rcpp::NumericMatrix F, Rcpp::NumericMatrix G, Rcpp::NumericMatrix V,
               Rcpp::NumericMatrix W, float df, std::string resampler, std::string distribution, unsigned p = 0)
{
    (void)p;
    const int dy = Y.nrow();
    if ((unsigned)Y.ncol() < timeSteps) Rcpp::stop("CuSMC: Y has fewer than timeSteps columns");
    if ((unsigned)m0.size() != d) Rcpp::stop("CuSMC: length(m0) != d");
    require_square(C0, (int)d, "C0");
    require_square(G, (int)d, "G");
    require_square(W, (int)d, "W");
    require_square(V, dy, "V");
    if (F.nrow() != dy || (unsigned)F.ncol() != d) Rcpp::stop("CuSMC: F must be nrow(Y) x d");
    cusmc_filter_config cfg{};
    cfg.N = N;
    cfg.d = (int)d;
    cfg.dy = dy;
    cfg.T = (int)timeSteps;
    if (distribution == "mvn" || distribution == "normal") cfg.kind = CUSMC_MVN;
    else if (distribution == "mvt") cfg.kind = CUSMC_MVT;
    else Rcpp::stop("CuSMC: unknown distribution '%s'", distribution.c_str());
    if (resampler == "metropolis") cfg.resampler = CUSMC_RESAMPLE_METROPOLIS;
    else if (resampler == "systematic") cfg.resampler = CUSMC_RESAMPLE_SYSTEMATIC;
    else if (resampler == "multinomial") cfg.resampler = CUSMC_RESAMPLE_MULTINOMIAL;
    else if (resampler == "rejection") cfg.resampler = CUSMC_RESAMPLE_REJECTION;
    else Rcpp::stop("CuSMC: unknown resampler '%s'", resampler.c_str());
    cfg.B = 10;                                   // ref: src/mcmc.cpp:252-255
    cfg.nu = df;
    cfg.noise_scale = 1.0;
    cfg.seed = session_seed();
    cfg.Y = Y.begin();
    cfg.m0 = m0.begin();
    cfg.C0 = C0.begin();
    cfg.F = F.begin();
    cfg.G = G.begin();
    cfg.V = V.begin();
    cfg.W = W.begin();
    const size_t T = timeSteps, TN = T * (size_t)N;
    std::vector<double> w(TN), x(TN * d);         // [T][N] and [T][N][d], row-major
    check(cusmc_run(ctx(), &cfg, w.data(), x.data()));
    // R arrays are column-major: element (t, i[, k]) lives at t + T i [+ T N k]
    Rcpp::NumericMatrix weights((int)T, (int)N);
    Rcpp::NumericVector theta((R_xlen_t)(TN * d));
    double *wo = weights.begin(), *xo = theta.begin();
    for (size_t t = 0; t < T; ++t)
        for (size_t i = 0; i < N; ++i) {
            wo[t + T * i] = w[t * N + i];
            for (size_t k = 0; k < d; ++k) xo[t + T * i + TN * k] = x[(t * N + i) * d + k];
        }
    theta.attr("dim") = Rcpp::Dimension((int)T, (int)N, (int)d);
    return Rcpp::List::create(Rcpp::Named("weights") = weights, Rcpp::Named("posterior_x") = theta);
}

// filter.cu -- the particle-filter step and the time loop.
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
#include "density.cuh"
#include "hostmath.h"
#include "filter_types.cuh"
#include "mailbox.cuh"
#include "pf_fused_impl.cuh"
#include "pf_step.cuh"
#include "resample.cuh"
#include "tile_update.cuh"

#include "../../include/cusmc_detmath.h"
#include "../../include/cusmc_philox.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdlib>
#include <cmath>
#include <new>
#include <thread>
#include <vector>

int cusmc_density_launch(cusmc_ctx *ctx, bool tri, int m, int d, const std::vector<double> &M_rowmajor,
                         const double *shift, const double *off, const Epilogue &ep, const double *x,
                         int layout, int64_t N, int64_t ld, double *out);
int cusmc_build_whitening(cusmc_ctx *ctx, int kind, int want_log, int d, const double *sigma, float nu,
                          std::vector<double> &W_rowmajor, Epilogue &ep);

namespace {

constexpr int kThreads = 256;

// ---- layout transposes through shared memory (coalesced on both sides) ----------------------
__global__ void __launch_bounds__(kThreads)
aos_to_soa_kernel(const double *__restrict__ aos, double *__restrict__ soa, int64_t N, int64_t ld, int d)
{
    extern __shared__ double tile[];
    const int pitch = d | 1;
    const int64_t base = (int64_t)blockIdx.x * kThreads;
    const int npts = (int)((N - base) < kThreads ? (N - base) : kThreads);
    const int n_el = npts * d;
    for (int e = threadIdx.x; e < n_el; e += kThreads) {
        const int row = e / d, col = e - row * d;
        tile[row * pitch + col] = aos[base * d + e];
    }
    __syncthreads();
    if ((int)threadIdx.x < npts)
        for (int j = 0; j < d; ++j) soa[(int64_t)j * ld + base + threadIdx.x] = tile[threadIdx.x * pitch + j];
}

__global__ void __launch_bounds__(kThreads)
soa_to_aos_kernel(const double *__restrict__ soa, double *__restrict__ aos, int64_t N, int64_t ld, int d)
{
    extern __shared__ double tile[];
    const int pitch = d | 1;
    const int64_t base = (int64_t)blockIdx.x * kThreads;
    const int npts = (int)((N - base) < kThreads ? (N - base) : kThreads);
    if ((int)threadIdx.x < npts)
        for (int j = 0; j < d; ++j) tile[threadIdx.x * pitch + j] = soa[(int64_t)j * ld + base + threadIdx.x];
    __syncthreads();
    const int n_el = npts * d;
    for (int e = threadIdx.x; e < n_el; e += kThreads) {
        const int row = e / d, col = e - row * d;
        aos[base * d + e] = tile[row * pitch + col];
    }
}

// Weighted first moments and weight sums of one step: out[0] += sum w, out[1] += sum w^2,
// out[2 + k] += sum w x_k, with w = exp(lw - max) (log mode) or the raw density.
__global__ void __launch_bounds__(kThreads)
moments_kernel(const double *__restrict__ x, const double *__restrict__ w, const double *__restrict__ wmax,
               int is_log, int64_t N, int64_t ld, int d, double *__restrict__ out)
{
    __shared__ double sm[kThreads / 32];
    const double m = (is_log && wmax) ? *wmax : 0.0;
    double acc[2 + CUSMC_MAX_DIM];
    for (int k = 0; k < 2 + d; ++k) acc[k] = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < N; i += (int64_t)gridDim.x * kThreads) {
        double wi = __ldg(w + i);
        if (is_log) wi = exp(wi - m);
        if (!(wi > 0.0) || wi == INFINITY) wi = 0.0;
        acc[0] += wi;
        acc[1] = fma(wi, wi, acc[1]);
        for (int k = 0; k < d; ++k) acc[2 + k] = fma(wi, __ldg(x + (int64_t)k * ld + i), acc[2 + k]);
    }
    for (int k = 0; k < 2 + d; ++k) {
        double v = acc[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
        __syncthreads();
        if (threadIdx.x == 0) {
            double t = 0.0;
            for (int q = 0; q < kThreads / 32; ++q) t += sm[q];
            atomicAdd(out + k, t);
        }
        __syncthreads();
    }
}

}  // namespace

// Observation model set-up: M = L_V^-1 F (row-major dy x d), Winv = L_V^-1 (row-major), epilogue.
static int build_observation(cusmc_ctx *ctx, int kind, int want_log, int d, int dy, const double *F,
                             const double *V, float nu, std::vector<double> &M, std::vector<double> &Winv,
                             Epilogue &ep)
{
    CUSMC_CHECK(cusmc_build_whitening(ctx, kind, want_log, dy, V, nu, Winv, ep));
    M.assign((size_t)dy * d, 0.0);
    for (int k = 0; k < dy; ++k)
        for (int j = 0; j < d; ++j) {
            double s = 0.0;
            for (int i = 0; i <= k; ++i) s += Winv[(size_t)k * dy + i] * F[(size_t)j * dy + i];
            M[(size_t)k * d + j] = s;
        }
    return CUSMC_OK;
}

// ---- extern "C": layout helpers -----------------------------------------------------------------
extern "C" int cusmc_aos_to_soa_dev(cusmc_ctx *ctx, const double *aos_dev, double *soa_dev, int64_t N,
                                    int64_t ld, int d)
{
    CUSMC_ENTER(ctx);
    CUSMC_REQUIRE(ctx, N >= 0 && d >= 1 && d <= 96 && ld >= N, "bad sizes (d must be in 1..96)");
    if (N == 0) return CUSMC_OK;
    const size_t smem = sizeof(double) * kThreads * (size_t)(d | 1);
    if (smem > 48 * 1024)
        CUSMC_CUDA(ctx, cudaFuncSetAttribute(aos_to_soa_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    aos_to_soa_kernel<<<(unsigned)((N + kThreads - 1) / kThreads), kThreads, smem, ctx->stream>>>(aos_dev, soa_dev, N, ld, d);
    CUSMC_LAUNCHED(ctx);
    return CUSMC_OK;
}

extern "C" int cusmc_soa_to_aos_dev(cusmc_ctx *ctx, const double *soa_dev, double *aos_dev, int64_t N,
                                    int64_t ld, int d)
{
    CUSMC_ENTER(ctx);
    CUSMC_REQUIRE(ctx, N >= 0 && d >= 1 && d <= 96 && ld >= N, "bad sizes (d must be in 1..96)");
    if (N == 0) return CUSMC_OK;
    const size_t smem = sizeof(double) * kThreads * (size_t)(d | 1);
    if (smem > 48 * 1024)
        CUSMC_CUDA(ctx, cudaFuncSetAttribute(soa_to_aos_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    soa_to_aos_kernel<<<(unsigned)((N + kThreads - 1) / kThreads), kThreads, smem, ctx->stream>>>(soa_dev, aos_dev, N, ld, d);
    CUSMC_LAUNCHED(ctx);
    return CUSMC_OK;
}

// ---- extern "C": fused propagate + reweight on device SoA state -----------------------------------
extern "C" int cusmc_propagate_reweight_dev(cusmc_ctx *ctx, int kind, int want_log, double *x_new_dev,
                                            const double *x_prev_dev, const uint32_t *a_dev, int64_t N,
                                            int64_t ld, int d, int dy, const double *G, const double *Q,
                                            const double *y, const double *F, const double *V, float nu,
                                            const double *xi_dev, const double *chi_dev, uint64_t seed,
                                            uint64_t step, double *lw_dev, double *lw_max_dev)
{
    CUSMC_ENTER(ctx);
    CUSMC_REQUIRE(ctx, N >= 0 && d >= 1 && dy >= 1 && ld >= N, "bad sizes");
    if (d > CUSMC_MAX_DIM || dy > CUSMC_MAX_DIM)       // before anything sized CUSMC_MAX_DIM is filled
        return cusmc_fail(ctx, CUSMC_ERR_UNSUPPORTED, "d/dy > %d", CUSMC_MAX_DIM);
    CUSMC_REQUIRE(ctx, G && Q && y && F && V, "model matrix is NULL");
    CUSMC_REQUIRE(ctx, N == 0 || (x_new_dev && x_prev_dev && lw_dev), "state pointer is NULL");
    CUSMC_REQUIRE(ctx, x_new_dev != x_prev_dev, "x_new must not alias x_prev (children read arbitrary parents)");
    std::vector<double> M, Winv;
    Epilogue ep;
    CUSMC_CHECK(build_observation(ctx, kind, want_log, d, dy, F, V, nu, M, Winv, ep));
    double c[CUSMC_MAX_DIM];
    cusmc_whiten_observation(Winv, dy, y, c);
    StepArgs a{};
    a.x_new = x_new_dev;
    a.x_prev = x_prev_dev;
    a.anc = a_dev;
    a.xi = xi_dev;
    a.chi = chi_dev;
    a.lw = lw_dev;
    a.lw_max = lw_max_dev;
    a.n_out = N;
    a.ld_new = a.ld_prev = a.ld_noise = ld;
    a.seed = seed;
    a.step = step;
    a.nu = nu;
    a.d = d;
    a.dy = dy;
    a.kind = kind;
    a.has_prev = 1;
    a.rng_stream = CUSMC_STREAM_NORMAL;
    return cusmc_launch_step(ctx, d, dy, G, Q, 1.0, &M, c, nullptr, ep, a, xi_dev == nullptr);
}

// ---- extern "C": drop-ins for the reference's pdf wrappers (host pointers, AoS) --------------------
static int pdf_dropin(cusmc_ctx *ctx, int kind, double *w, const double *y, const double *x_aos, double norm,
                      const double *E_inv, const double *F, int64_t N, int d, int dy, float df)
{
    CUSMC_ENTER(ctx);
    CUSMC_REQUIRE(ctx, N >= 0 && d >= 1 && dy >= 1, "bad sizes");
    CUSMC_REQUIRE(ctx, y && E_inv && F && (N == 0 || (w && x_aos)), "NULL pointer");
    if (d > CUSMC_MAX_DIM || dy > CUSMC_MAX_DIM)
        return cusmc_fail(ctx, CUSMC_ERR_UNSUPPORTED, "d/dy > %d", CUSMC_MAX_DIM);
    if (kind == CUSMC_MVT && !(df > 0.0f)) return cusmc_fail(ctx, CUSMC_ERR_INVALID, "mvt needs df > 0");
    if (N == 0) return CUSMC_OK;
    // q = r^T P r with P = E_inv = Lp Lp^T  ->  q = |Lp^T r|^2,  r = y - F x
    std::vector<double> Lp;
    const int bad = hostmath::cholesky_lower(E_inv, dy, Lp);
    if (bad) return cusmc_fail(ctx, CUSMC_ERR_NOT_SPD, "E_inv is not positive definite (pivot %d)", bad - 1);
    std::vector<double> M((size_t)dy * d, 0.0);
    double c[CUSMC_MAX_DIM];
    for (int k = 0; k < dy; ++k) {
        for (int j = 0; j < d; ++j) {
            double s = 0.0;
            for (int i = k; i < dy; ++i) s += Lp[(size_t)k * dy + i] * F[(size_t)j * dy + i];
            M[(size_t)k * d + j] = s;
        }
        double s = 0.0;
        for (int i = k; i < dy; ++i) s += Lp[(size_t)k * dy + i] * y[i];
        c[k] = s;
    }
    Epilogue ep{};
    ep.kind = kind;
    ep.want_log = 0;
    ep.scale = norm;
    ep.lognorm = std::log(norm);
    if (kind == CUSMC_MVT) {
        ep.half_nu_d = hostmath::mvt_half_nu_plus_d(df, dy);
        ep.inv_nu = 1.0 / (double)df;
    }
    void *xd = nullptr, *wd = nullptr;
    CUSMC_CHECK(cusmc_scratch(ctx, 0, sizeof(double) * (size_t)N * d, &xd));
    CUSMC_CHECK(cusmc_scratch(ctx, 1, sizeof(double) * (size_t)N, &wd));
    CUSMC_CUDA(ctx, cudaMemcpyAsync(xd, x_aos, sizeof(double) * (size_t)N * d, cudaMemcpyHostToDevice, ctx->stream));
    CUSMC_CUDA(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
    CUSMC_CHECK(cusmc_density_launch(ctx, false, dy, d, M, nullptr, c, ep, (const double *)xd, CUSMC_AOS, N, N, (double *)wd));
    CUSMC_CUDA(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
    CUSMC_CUDA(ctx, cudaMemcpyAsync(w, wd, sizeof(double) * (size_t)N, cudaMemcpyDeviceToHost, ctx->stream));
    CUSMC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    float ms = 0.f;
    CUSMC_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    ctx->last_ms = ms;
    return CUSMC_OK;
}

extern "C" int cusmc_mvn_pdf(cusmc_ctx *ctx, double *w, const double *y, const double *x_aos, double norm,
                             const double *E_inv, const double *F, int64_t N, int d, int dy)
{
    return pdf_dropin(ctx, CUSMC_MVN, w, y, x_aos, norm, E_inv, F, N, d, dy, 0.0f);
}

extern "C" int cusmc_mvt_pdf(cusmc_ctx *ctx, double *w, const double *y, const double *x_aos,
                             const double *E_inv, const double *F, double norm, int64_t N, int d, int dy,
                             float df)
{
    return pdf_dropin(ctx, CUSMC_MVT, w, y, x_aos, norm, E_inv, F, N, d, dy, df);
}

// ---- extern "C": drop-ins for the reference's sample wrappers (host pointers, AoS) -------------------
static int sample_dropin(cusmc_ctx *ctx, int kind, double *x_new_aos, const double *x_prev_aos,
                         const uint32_t *anc, const double *G, const double *mu, const double *Q,
                         const double *xi, const double *chi, uint64_t seed, uint64_t step, int stream_id,
                         int64_t N, int d, float df)
{
    CUSMC_ENTER(ctx);
    CUSMC_REQUIRE(ctx, N >= 0 && d >= 1 && d <= CUSMC_MAX_DIM, "bad sizes");
    CUSMC_REQUIRE(ctx, Q && (N == 0 || x_new_aos), "NULL pointer");
    CUSMC_REQUIRE(ctx, !G || N == 0 || x_prev_aos, "x_prev is NULL");
    if (kind == CUSMC_MVT && !(df > 0.0f)) return cusmc_fail(ctx, CUSMC_ERR_INVALID, "mvt needs df > 0");
    if (N == 0) return CUSMC_OK;
    const size_t nd = sizeof(double) * (size_t)N * d;
    void *stage = nullptr, *xprev = nullptr, *xnew = nullptr, *xis = nullptr, *chis = nullptr, *ad = nullptr, *lw = nullptr;
    CUSMC_CHECK(cusmc_scratch(ctx, 0, nd, &stage));
    CUSMC_CHECK(cusmc_scratch(ctx, 1, nd, &xprev));
    CUSMC_CHECK(cusmc_scratch(ctx, 2, nd, &xnew));
    CUSMC_CHECK(cusmc_scratch(ctx, 5, sizeof(double) * (size_t)N, &lw));
    if (G) {
        CUSMC_CUDA(ctx, cudaMemcpyAsync(stage, x_prev_aos, nd, cudaMemcpyHostToDevice, ctx->stream));
        CUSMC_CHECK(cusmc_aos_to_soa_dev(ctx, (const double *)stage, (double *)xprev, N, N, d));
    }
    if (xi) {
        CUSMC_CHECK(cusmc_scratch(ctx, 3, nd, &xis));
        CUSMC_CUDA(ctx, cudaMemcpyAsync(stage, xi, nd, cudaMemcpyHostToDevice, ctx->stream));
        CUSMC_CHECK(cusmc_aos_to_soa_dev(ctx, (const double *)stage, (double *)xis, N, N, d));
    }
    if (chi) {
        CUSMC_CHECK(cusmc_scratch(ctx, 4, nd, &chis));
        CUSMC_CUDA(ctx, cudaMemcpyAsync(stage, chi, nd, cudaMemcpyHostToDevice, ctx->stream));
        CUSMC_CHECK(cusmc_aos_to_soa_dev(ctx, (const double *)stage, (double *)chis, N, N, d));
    }
    if (anc) {
        CUSMC_CHECK(cusmc_scratch(ctx, 6, sizeof(uint32_t) * (size_t)N + 64, &ad));
        ad = (char *)ad + 64;   // the first 64 bytes of slot 6 hold reduction scalars elsewhere
        CUSMC_CUDA(ctx, cudaMemcpyAsync(ad, anc, sizeof(uint32_t) * (size_t)N, cudaMemcpyHostToDevice, ctx->stream));
    }
    StepArgs a{};
    a.x_new = (double *)xnew;
    a.x_prev = (const double *)xprev;
    a.anc = (const uint32_t *)ad;
    a.xi = (const double *)xis;
    a.chi = (const double *)chis;
    a.lw = (double *)lw;
    a.n_out = N;
    a.ld_new = a.ld_prev = a.ld_noise = N;
    a.seed = seed;
    a.step = step;
    a.nu = df;
    a.d = d;
    a.dy = d;
    a.kind = kind;
    a.has_prev = G ? 1 : 0;
    a.skip_weight = 1;
    a.rng_stream = stream_id;
    Epilogue ep{};
    CUSMC_CUDA(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
    CUSMC_CHECK(cusmc_launch_step(ctx, d, d, G, Q, 1.0, nullptr, nullptr, mu, ep, a, xi == nullptr));
    CUSMC_CUDA(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
    CUSMC_CHECK(cusmc_soa_to_aos_dev(ctx, (const double *)xnew, (double *)stage, N, N, d));
    CUSMC_CUDA(ctx, cudaMemcpyAsync(x_new_aos, stage, nd, cudaMemcpyDeviceToHost, ctx->stream));
    CUSMC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    float ms = 0.f;
    CUSMC_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    ctx->last_ms = ms;
    return CUSMC_OK;
}

extern "C" int cusmc_mvn_sample(cusmc_ctx *ctx, double *x_new_aos, const double *x_prev_aos, const uint32_t *a,
                                const double *G, const double *Q, const double *xi, uint64_t seed,
                                uint64_t step, int64_t N, int d)
{
    if (ctx && !G) return cusmc_fail(ctx, CUSMC_ERR_INVALID, "cusmc_mvn_sample: G is NULL");
    return sample_dropin(ctx, CUSMC_MVN, x_new_aos, x_prev_aos, a, G, nullptr, Q, xi, nullptr, seed, step,
                         CUSMC_STREAM_NORMAL, N, d, 0.0f);
}

extern "C" int cusmc_mvn_sample_init(cusmc_ctx *ctx, double *x_aos, const double *mu, const double *Q,
                                     const double *xi, uint64_t seed, int64_t N, int d)
{
    return sample_dropin(ctx, CUSMC_MVN, x_aos, nullptr, nullptr, nullptr, mu, Q, xi, nullptr, seed, 0,
                         CUSMC_STREAM_INIT, N, d, 0.0f);
}

extern "C" int cusmc_mvt_sample(cusmc_ctx *ctx, double *x_new_aos, const double *x_prev_aos, const uint32_t *a,
                                const double *G, const double *Q, const double *xi, const double *chi,
                                uint64_t seed, uint64_t step, int64_t N, int d, float df)
{
    if (ctx && !G) return cusmc_fail(ctx, CUSMC_ERR_INVALID, "cusmc_mvt_sample: G is NULL");
    return sample_dropin(ctx, CUSMC_MVT, x_new_aos, x_prev_aos, a, G, nullptr, Q, xi, chi, seed, step,
                         CUSMC_STREAM_NORMAL, N, d, df);
}

// ================================================================================================
// The filter object: particle_filter() / initialize() / MCMC() of the reference
// (src/particle_filter.cpp:6-39, src/mcmc.cpp:44-88,239-309) with device-resident state.
// ================================================================================================
// Scalar exchanges of a sharded run ride inside the kernels (mailbox.cuh) when `fused` is set.
static MailArgs filter_mail(const cusmc_filter *f)
{
    MailArgs m{};
    if (f->world > 1 && f->fused) {
        m.peer = (unsigned long long *const *)f->peer_mail.table_dev;
        m.err = f->mail_err;
        m.timeout_ns = f->mail_timeout_ns;
        m.epoch = f->epoch;
        m.rank = f->rank;
        m.world = f->world;
    }
    return m;
}

// Symmetric eigen factor Q = V sqrt(Lambda) (reference: eigenSolver, src/linear_algebra.cpp:10-23)
// by cyclic Jacobi; any Q with Q Q^T = Sigma gives the same law, the eigen form is kept so the
// noise a given xi produces matches the reference's convention up to eigenvector sign.
static int eigen_factor(cusmc_ctx *ctx, const double *S, int d, std::vector<double> &Q)
{
    std::vector<double> A(S, S + (size_t)d * d), Vv((size_t)d * d, 0.0);
    for (int i = 0; i < d; ++i) Vv[(size_t)i * d + i] = 1.0;
    for (int sweep = 0; sweep < 100; ++sweep) {
        double off = 0.0;
        for (int p = 0; p < d; ++p)
            for (int q = p + 1; q < d; ++q) off += A[(size_t)q * d + p] * A[(size_t)q * d + p];
        if (off < 1e-300) break;
        for (int p = 0; p < d; ++p)
            for (int q = p + 1; q < d; ++q) {
                const double apq = A[(size_t)q * d + p];
                if (std::fabs(apq) < 1e-300) continue;
                const double app = A[(size_t)p * d + p], aqq = A[(size_t)q * d + q];
                const double theta = (aqq - app) / (2.0 * apq);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
                const double cs = 1.0 / std::sqrt(t * t + 1.0), sn = t * cs;
                for (int k = 0; k < d; ++k) {
                    const double akp = A[(size_t)p * d + k], akq = A[(size_t)q * d + k];
                    A[(size_t)p * d + k] = cs * akp - sn * akq;
                    A[(size_t)q * d + k] = sn * akp + cs * akq;
                }
                for (int k = 0; k < d; ++k) {
                    const double apk = A[(size_t)k * d + p], aqk = A[(size_t)k * d + q];
                    A[(size_t)k * d + p] = cs * apk - sn * aqk;
                    A[(size_t)k * d + q] = sn * apk + cs * aqk;
                }
                for (int k = 0; k < d; ++k) {
                    const double vkp = Vv[(size_t)p * d + k], vkq = Vv[(size_t)q * d + k];
                    Vv[(size_t)p * d + k] = cs * vkp - sn * vkq;
                    Vv[(size_t)q * d + k] = sn * vkp + cs * vkq;
                }
            }
    }
    Q.assign((size_t)d * d, 0.0);
    for (int c = 0; c < d; ++c) {
        const double lam = A[(size_t)c * d + c];
        if (!(lam >= -1e-12)) return cusmc_fail(ctx, CUSMC_ERR_NOT_SPD, "covariance has a negative eigenvalue");
        const double s = std::sqrt(lam > 0 ? lam : 0.0);
        for (int r = 0; r < d; ++r) Q[(size_t)c * d + r] = Vv[(size_t)c * d + r] * s;
    }
    return CUSMC_OK;
}

// Exported buffers, in this order: x[0], x[1], ancestors, weights, mailbox, weight images 0 and 1
// (a reference-mode filter has no images: it exports its weights twice more, never read).
enum { kIpcBuffers = CUSMC_FILTER_IPC_BUFFERS };

static void detach_peers(cusmc_filter *f)
{
    if (!f->attached) return;
    CusmcPeers *tabs[kIpcBuffers] = {&f->peer_x[0], &f->peer_x[1], &f->peer_anc, &f->peer_lw, &f->peer_mail,
                                     &f->peer_img[0], &f->peer_img[1]};
    for (CusmcPeers *t : tabs)
        for (int r = 0; r < f->world; ++r)
            if (r != f->rank && t->ptr[r]) cudaIpcCloseMemHandle(t->ptr[r]);
    f->attached = false;
}

extern "C" int cusmc_filter_destroy(cusmc_filter *f)
{
    if (!f) return CUSMC_OK;
    cudaSetDevice(f->ctx->device);
    cudaStreamSynchronize(f->ctx->stream);
    detach_peers(f);
    // pooled (single-GPU) filters hand their buffers back to the stream-ordered pool; sharded ones own
    // cudaMalloc memory (CUDA IPC cannot export pool allocations)
    cudaStream_t st = f->ctx->stream;
    auto release = [&](void *p) {
        if (!p) return;
        if (f->pooled) cudaFreeAsync(p, st); else cudaFree(p);
    };
    release(f->x[0]);
    release(f->x[1]);
    release(f->lw);
    release(f->anc);
    release(f->slots);
    release(f->moments);
    release(f->img[0]);
    release(f->img[1]);
    release(f->rank_sums);
    cudaFree(f->persist);
    release(f->mail);
    release(f->mail_err);
    cudaFree(f->peer_tables);
    release(f->hist_x);
    release(f->hist_w);
    release(f->hist_a);
    if (f->ev0) cudaEventDestroy(f->ev0);
    if (f->ev1) cudaEventDestroy(f->ev1);
    delete f;
    return CUSMC_OK;
}

// Particles per tile of the per-step path: kTile unless the caller fixes it (cfg.tile_size).  An automatic
// "whole waves of resident blocks" choice was measured and dropped: a C5 shard as 4370 tiles of 1920 (5.9 waves
// of 740 blocks) instead of 4096 tiles of 2048 (5.5 waves) runs at 240 us per step instead of 233 (1792: 243;
// dense model 370 vs 357; 16 Mi particles 460 vs 447) -- blocks finish staggered, so the last wave costs less
// than its emptiness suggests, and every extra tile pays its own lookup, scans and a half-empty last round.
static uint32_t pick_step_tile(const cusmc_filter *f)
{
    return f->cfg.tile_size ? (uint32_t)f->cfg.tile_size : (uint32_t)kTile;
}

extern "C" int cusmc_filter_create(cusmc_ctx *ctx, const cusmc_filter_config *cfg, cusmc_filter **out)
{
    if (!ctx || !out) return CUSMC_ERR_INVALID;
    *out = nullptr;
    CUSMC_REQUIRE(ctx, cfg != nullptr, "config is NULL");
    CUSMC_REQUIRE(ctx, cfg->N >= 1 && cfg->N <= 0xFFFFFFFFll, "N must be in 1..2^32-1");
    CUSMC_REQUIRE(ctx, cfg->d >= 1 && cfg->dy >= 1 && cfg->T >= 1, "d, dy, T must be positive");
    CUSMC_REQUIRE(ctx, cfg->d <= CUSMC_MAX_DIM && cfg->dy <= CUSMC_MAX_DIM, "d/dy exceed CUSMC_MAX_DIM");
    CUSMC_REQUIRE(ctx, cfg->Y && cfg->m0 && cfg->C0 && cfg->F && cfg->G && cfg->V && cfg->W, "model pointer is NULL");
    CUSMC_REQUIRE(ctx, cfg->kind == CUSMC_MVN || cfg->kind == CUSMC_MVT, "unknown distribution");
    CUSMC_REQUIRE(ctx, cfg->resampler >= 0 && cfg->resampler <= 4, "unknown resampler");
    CUSMC_REQUIRE(ctx, cfg->kind == CUSMC_MVN || cfg->nu > 0.0f, "mvt needs nu > 0");
    CUSMC_REQUIRE(ctx, cfg->ess_threshold >= 0.0 && cfg->ess_threshold <= 1.0, "ess_threshold must lie in [0, 1]");
    CUSMC_REQUIRE(ctx, cfg->tile_size == 0 || (cfg->tile_size >= 32 && cfg->tile_size <= kTile && cfg->tile_size % 32 == 0),
                  "tile_size must be 0 or a multiple of 32 in [32, 2048]");
    CUSMC_REQUIRE(ctx, cfg->tile_size == 0 || cfg->tile_size == kTile ||
                           (cfg->world <= 1 && cfg->resampler == CUSMC_RESAMPLE_SYSTEMATIC),
                  "tile_size other than 2048: single-GPU systematic resampling only");
    CUSMC_REQUIRE(ctx, cfg->ess_threshold == 0.0 || cfg->resampler == CUSMC_RESAMPLE_SYSTEMATIC,
                  "adaptive resampling needs the systematic resampler");
    const int world = cfg->world <= 1 ? 1 : cfg->world;
    CUSMC_REQUIRE(ctx, world <= CUSMC_MAX_PEERS, "world exceeds CUSMC_MAX_PEERS");
    CUSMC_REQUIRE(ctx, world == 1 || (cfg->rank >= 0 && cfg->rank < world), "rank outside 0..world-1");
    if (world > 1 && cfg->resampler == CUSMC_RESAMPLE_REJECTION)
        return cusmc_fail(ctx, CUSMC_ERR_UNSUPPORTED, "the rejection resampler is single-GPU only");
    cusmc_filter *f = new (std::nothrow) cusmc_filter();
    if (!f) return cusmc_fail(ctx, CUSMC_ERR_CUDA, "out of host memory");
    f->ctx = ctx;
    f->cfg = *cfg;
    const int d = cfg->d, dy = cfg->dy, T = cfg->T;
    const int64_t N = cfg->N;
    f->world = world;
    f->rank = world == 1 ? 0 : cfg->rank;
    // a shard is a whole number of weight-image tiles, so the tiles of a sharded run are the tiles of the
    // single-GPU run and both produce the same bits
    f->per = (N + world - 1) / world;
    if (world > 1) f->per = (f->per + kTile - 1) / kTile * kTile;
    f->lo = std::min<int64_t>((int64_t)f->rank * f->per, N);
    f->n = std::max<int64_t>(0, std::min<int64_t>(f->per, N - f->lo));
    f->Y.assign(cfg->Y, cfg->Y + (size_t)dy * T);
    f->m0.assign(cfg->m0, cfg->m0 + d);
    f->C0.assign(cfg->C0, cfg->C0 + (size_t)d * d);
    f->F.assign(cfg->F, cfg->F + (size_t)dy * d);
    f->G.assign(cfg->G, cfg->G + (size_t)d * d);
    f->V.assign(cfg->V, cfg->V + (size_t)dy * dy);
    f->W.assign(cfg->W, cfg->W + (size_t)d * d);
    f->cfg.Y = f->Y.data();
    f->cfg.m0 = f->m0.data();
    f->cfg.C0 = f->C0.data();
    f->cfg.F = f->F.data();
    f->cfg.G = f->G.data();
    f->cfg.V = f->V.data();
    f->cfg.W = f->W.data();
    if (f->cfg.noise_scale == 0.0) f->cfg.noise_scale = 1.0;
    if (f->cfg.B <= 0) f->cfg.B = 10;   // the reference hard-codes B = 10 (src/mcmc.cpp:291)
    // Reference mode (metropolis) keeps raw densities as weights like src/mcmc.cpp:212; the
    // normalised resamplers work on log-weights.
    f->is_log = (cfg->resampler == CUSMC_RESAMPLE_METROPOLIS || cfg->resampler == CUSMC_RESAMPLE_REJECTION ||
                 cfg->resampler == CUSMC_RESAMPLE_METROPOLIS_C2) ? 0 : 1;
    f->shift = cusmc_fixed_shift(N);
    int rc = eigen_factor(ctx, f->C0.data(), d, f->Qc0);
    if (rc == CUSMC_OK) rc = eigen_factor(ctx, f->W.data(), d, f->Qw);
    if (rc == CUSMC_OK)
        rc = build_observation(ctx, cfg->kind, f->is_log, d, dy, f->F.data(), f->V.data(), cfg->nu, f->M, f->Winv, f->ep);
    if (rc != CUSMC_OK) {
        delete f;
        return rc;
    }
    f->tile = pick_step_tile(f);
    cudaError_t e = cudaSetDevice(ctx->device);
    f->pooled = world == 1;
    auto alloc = [&](void **p, size_t bytes) {
        if (e != cudaSuccess) return;
        e = f->pooled ? cudaMallocAsync(p, bytes ? bytes : 8, ctx->stream) : cudaMalloc(p, bytes ? bytes : 8);
    };
    const size_t P = (size_t)f->per;          // columns allocated on every rank
    alloc((void **)&f->x[0], sizeof(double) * P * d);
    alloc((void **)&f->x[1], sizeof(double) * P * d);
    alloc((void **)&f->lw, sizeof(double) * P);
    alloc((void **)&f->anc, sizeof(uint32_t) * P);
    alloc((void **)&f->slots, sizeof(StepSlot) * (size_t)T);
    alloc((void **)&f->moments, sizeof(double) * (size_t)T * (2 + d));
    if (f->is_log) {
        // two weight images (step parity): step t reads image t - 1 while it writes image t.  A persistent
        // run spreads the cloud over more, smaller tiles (one per resident block): the layout holds both.
        f->persist_tile = cusmc_filter_persistent_tile(f);
        f->img_n = f->per;
        if (f->persist_tile) f->img_n = std::max<int64_t>(f->img_n, (N + f->persist_tile - 1) / f->persist_tile * (int64_t)kTile);
        if (f->tile != (uint32_t)kTile) f->img_n = std::max<int64_t>(f->img_n, (N + f->tile - 1) / f->tile * (int64_t)kTile);
        const size_t img_bytes = sizeof(unsigned long long) * (size_t)fimage_words(f->img_n);
        for (int b = 0; b < 2; ++b) {
            alloc((void **)&f->img[b], img_bytes);
            if (e == cudaSuccess)
                e = cudaMemsetAsync(f->img[b], 0, sizeof(unsigned long long) * (size_t)fimage_header_words(f->img_n), ctx->stream);
        }
        if (world > 1) alloc((void **)&f->rank_sums, sizeof(unsigned long long) * 3 * CUSMC_MAX_PEERS);
    }
    if (world > 1) {
        const size_t mail_bytes = sizeof(unsigned long long) * kMailWords * kMailCells * (size_t)world * (size_t)T;
        alloc((void **)&f->mail, mail_bytes);
        alloc((void **)&f->mail_err, 8);
        if (e == cudaSuccess) e = cudaMemsetAsync(f->mail, 0, mail_bytes, ctx->stream);
        if (e == cudaSuccess) e = cudaMemsetAsync(f->mail_err, 0, 8, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);     // peers may map these buffers right away
    }
    if (cfg->keep_history) {
        // keep_history = 1: all T steps stay on the device (cusmc_filter_get_history);
        // keep_history < 0: a ring of 2 chunks of -keep_history steps (cusmc_run, world == 1)
        const size_t rows = cfg->keep_history < 0 ? (size_t)2 * (size_t)(-cfg->keep_history) : (size_t)T;
        f->ring_K = cfg->keep_history < 0 ? -cfg->keep_history : 0;
        alloc((void **)&f->hist_x, sizeof(double) * rows * P * d);
        alloc((void **)&f->hist_w, sizeof(double) * rows * P);
        alloc((void **)&f->hist_a, sizeof(uint32_t) * rows * P);
    }
    if (e == cudaSuccess) e = cudaEventCreate(&f->ev0);
    if (e == cudaSuccess) e = cudaEventCreate(&f->ev1);
    if (e != cudaSuccess) {
        cusmc_fail(ctx, CUSMC_ERR_CUDA, "filter allocation failed: %s", cudaGetErrorString(e));
        cusmc_filter_destroy(f);
        return CUSMC_ERR_CUDA;
    }
    *out = f;
    return CUSMC_OK;
}

// ---- sharded runs: peer mapping of the state ---------------------------------------------------

extern "C" int cusmc_filter_ipc_export(cusmc_filter *f, unsigned char *handles)
{
    if (!f) return CUSMC_ERR_INVALID;
    cusmc_ctx *ctx = f->ctx;
    CUSMC_REQUIRE(ctx, handles != nullptr, "handles is NULL");
    static_assert(sizeof(cudaIpcMemHandle_t) == CUSMC_IPC_HANDLE_BYTES, "IPC handle size");
    CUSMC_REQUIRE(ctx, f->world > 1, "not a sharded filter");
    void *bufs[kIpcBuffers] = {f->x[0], f->x[1], f->anc, f->lw, f->mail, f->img[0] ? (void *)f->img[0] : (void *)f->lw,
                               f->img[1] ? (void *)f->img[1] : (void *)f->lw};
    for (int b = 0; b < kIpcBuffers; ++b) {
        cudaIpcMemHandle_t h;
        CUSMC_CUDA(ctx, cudaIpcGetMemHandle(&h, bufs[b]));
        std::memcpy(handles + (size_t)b * CUSMC_IPC_HANDLE_BYTES, &h, sizeof h);
    }
    return CUSMC_OK;
}

extern "C" int cusmc_filter_ipc_attach(cusmc_filter *f, const unsigned char *all_handles)
{
    if (!f) return CUSMC_ERR_INVALID;
    cusmc_ctx *ctx = f->ctx;
    CUSMC_REQUIRE(ctx, all_handles != nullptr, "handles is NULL");
    CUSMC_REQUIRE(ctx, !f->attached, "peers are already attached");
    CUSMC_CUDA(ctx, cudaSetDevice(ctx->device));
    CUSMC_REQUIRE(ctx, f->world > 1, "not a sharded filter");
    CusmcPeers *tabs[kIpcBuffers] = {&f->peer_x[0], &f->peer_x[1], &f->peer_anc, &f->peer_lw, &f->peer_mail,
                                     &f->peer_img[0], &f->peer_img[1]};
    void *mine[kIpcBuffers] = {f->x[0], f->x[1], f->anc, f->lw, f->mail, f->img[0] ? (void *)f->img[0] : (void *)f->lw,
                               f->img[1] ? (void *)f->img[1] : (void *)f->lw};
    for (int b = 0; b < kIpcBuffers; ++b) {
        *tabs[b] = CusmcPeers{};
        tabs[b]->per_rank = f->per;
        tabs[b]->world = f->world;
    }
    f->attached = true;   // so a failure below closes what was opened
    for (int r = 0; r < f->world; ++r)
        for (int b = 0; b < kIpcBuffers; ++b) {
            if (r == f->rank) {
                tabs[b]->ptr[r] = mine[b];
                continue;
            }
            cudaIpcMemHandle_t h;
            std::memcpy(&h, all_handles + ((size_t)r * kIpcBuffers + b) * CUSMC_IPC_HANDLE_BYTES, sizeof h);
            cudaError_t e = cudaIpcOpenMemHandle(&tabs[b]->ptr[r], h, cudaIpcMemLazyEnablePeerAccess);
            if (e != cudaSuccess) {
                tabs[b]->ptr[r] = nullptr;
                detach_peers(f);
                return cusmc_fail(ctx, CUSMC_ERR_CUDA, "cudaIpcOpenMemHandle(rank %d, buffer %d): %s", r, b,
                                  cudaGetErrorString(e));
            }
        }
    // the pointer tables the kernels index live in device memory
    void *host_tab[kIpcBuffers][CUSMC_MAX_PEERS] = {};
    for (int b = 0; b < kIpcBuffers; ++b)
        for (int r = 0; r < f->world; ++r) host_tab[b][r] = tabs[b]->ptr[r];
    if (!f->peer_tables) CUSMC_CUDA(ctx, cudaMalloc((void **)&f->peer_tables, sizeof host_tab));
    CUSMC_CUDA(ctx, cudaMemcpy(f->peer_tables, host_tab, sizeof host_tab, cudaMemcpyHostToDevice));
    for (int b = 0; b < kIpcBuffers; ++b) tabs[b]->table_dev = f->peer_tables + (size_t)b * CUSMC_MAX_PEERS;
    return CUSMC_OK;
}

extern "C" int64_t cusmc_filter_tile_size(cusmc_filter *f)
{
    if (!f) return 0;
    // what cusmc_filter_run (without injected per-particle draws) will use
    return cusmc_filter_persistent_eligible(f, nullptr) ? (int64_t)f->persist_tile : (int64_t)f->tile;
}

extern "C" int cusmc_filter_slot_dev(cusmc_filter *f, int t, void **slot_dev)
{
    if (!f) return CUSMC_ERR_INVALID;
    CUSMC_REQUIRE(f->ctx, slot_dev && t >= 0 && t < f->cfg.T, "bad step");
    *slot_dev = f->slots + t;
    return CUSMC_OK;
}

extern "C" int cusmc_filter_moments_dev(cusmc_filter *f, double **moments_dev)
{
    if (!f) return CUSMC_ERR_INVALID;
    CUSMC_REQUIRE(f->ctx, moments_dev != nullptr, "NULL pointer");
    *moments_dev = f->moments;
    return CUSMC_OK;
}

__global__ void init_slots_kernel(StepSlot *slots, int T)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < T) {
        StepSlot s;
        s.lw_max = -INFINITY;
        s.sum_q = s.sum_q2 = s.n_pos = s.cdf_offset = 0;
        s.resampled = t > 0 ? 1 : 0;
        s.degenerate = 0;
        s.reserved = 0.0;
        slots[t] = s;
    }
}

// ---- the step, phase by phase ------------------------------------------------------------------
// One GPU: cusmc_filter_run chains them.  Sharded: the binding (cusmc_b200/sharded.py) puts the
// scalar exchanges between them -- MAX of slot[t].lw_max after propagate, the per-rank sums after
// weigh, a barrier after resample (ancestors land in peers' memory).

int cusmc_filter_init_slots(cusmc_filter *f)
{
    const int T = f->cfg.T;
    init_slots_kernel<<<(T + 255) / 256, 256, 0, f->ctx->stream>>>(f->slots, T);
    CUSMC_LAUNCHED(f->ctx);
    return CUSMC_OK;
}

constexpr int kRejectionCap = 4096;      // attempts per particle of the rejection resampler

// Does step t leave its log-weights in f->lw?  Nothing in a fused systematic step reads them (the weight
// image carries the weights), so they are stored only where somebody needs them: the summary's
// moment pass, the history, adaptive resampling (weights accumulate), the multinomial search's CDF
// comes from the image too -- and the LAST step, so that cusmc_filter_state_dev hands out the final
// log-weights.  Reference mode (metropolis) always keeps its densities: the resampler reads them.
static bool filter_stores_weights(const cusmc_filter *f, int t)
{
    const cusmc_filter_config &cfg = f->cfg;
    return !f->is_log || cfg.summary || cfg.keep_history || cfg.ess_threshold > 0.0 || t == cfg.T - 1;
}

// The step arguments common to t = 0 and t >= 1.
static StepArgs filter_step_args(cusmc_filter *f, int t)
{
    const cusmc_filter_config &cfg = f->cfg;
    StepArgs a{};
    a.lw = filter_stores_weights(f, t) ? f->lw : nullptr;
    a.n_out = f->n;
    a.ld_new = a.ld_prev = f->per;
    a.ld_noise = f->n;
    a.i0 = f->lo;
    a.seed = cfg.seed;
    a.step = (uint64_t)t;
    a.nu = cfg.nu;
    a.d = cfg.d;
    a.dy = cfg.dy;
    a.fast_noise = cfg.reproducible_rng ? 0 : 1;
    if (f->world > 1) {
        a.world = f->world;
        a.rank = f->rank;
        a.per_rank = make_fast_div((uint32_t)f->per);
    }
    return a;
}

static pffused::FusedArgs filter_fused_args(cusmc_filter *f, const StepArgs &a, int t)
{
    pffused::FusedArgs fa{};
    fa.s = a;
    fa.img_new = f->img[t & 1];
    fa.img_prev = f->img[(t & 1) ^ 1];
    fa.img_prev_peer = f->world > 1 ? (const unsigned long long *const *)f->peer_img[(t & 1) ^ 1].table_dev : nullptr;
    fa.img_hdr_words = fimage_header_words(f->img_n);
    fa.tiles_alloc = (uint32_t)fimage_tiles(f->img_n);
    fa.N_global = (uint32_t)f->cfg.N;
    fa.tiles_per_rank = (uint32_t)(f->per / f->tile > 0 ? (f->per + f->tile - 1) / f->tile : 1);
    fa.tile_n = f->tile;
    fa.shift = f->shift;
    return fa;
}

extern "C" int cusmc_filter_begin(cusmc_filter *f, const cusmc_filter_draws *draws)
{
    if (!f) return CUSMC_ERR_INVALID;
    cusmc_ctx *ctx = f->ctx;
    const cusmc_filter_config &cfg = f->cfg;
    const int d = cfg.d, dy = cfg.dy, T = cfg.T;
    CUSMC_REQUIRE(ctx, f->world == 1 || f->attached, "sharded filter: attach the peers first");
    f->draws = draws ? *draws : cusmc_filter_draws{};
    CUSMC_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    CUSMC_CHECK(cusmc_filter_init_slots(f));
    CUSMC_CUDA(ctx, cudaMemsetAsync(f->moments, 0, sizeof(double) * (size_t)T * (2 + d), st));
    // t = 0: initialize (src/mcmc.cpp:63-85): x_0 = m0 + Q_c0 xi, w_0 = 1/N
    f->cur = 0;
    StepArgs a = filter_step_args(f, 0);
    a.x_new = f->x[0];
    a.x_prev = f->x[1];
    a.xi = f->draws.xi0_dev;
    a.lw_max = cfg.resampler == CUSMC_RESAMPLE_REJECTION ? &f->slots[0].lw_max : nullptr;   // the largest density
    // "mvt": x_0 = m0 + chi (.) (Q_c0 xi), the reference's initialize() draws from the same distribution
    // object as the transition noise (src/mcmc.cpp:73-79 -> src/statistics.cc.cpp:379-411)
    a.kind = (cfg.kind == CUSMC_MVT && !cfg.mvt_normal_init) ? CUSMC_MVT : CUSMC_MVN;
    a.chi = a.kind == CUSMC_MVT ? f->draws.chi0_dev : nullptr;
    a.has_prev = 0;
    a.skip_weight = 1;
    a.const_weight = f->is_log ? 0.0 : 1.0 / (double)cfg.N;
    a.rng_stream = CUSMC_STREAM_INIT;
    if (cfg.keep_history) {       // row 0 (of chunk 0 in ring mode)
        a.hist_x = f->hist_x;
        a.hist_w = f->hist_w;
    }
    if (f->is_log) {
        // the fused kernel also leaves the weight image of step 0 (constant log-weight 0)
        pffused::FusedArgs fa = filter_fused_args(f, a, 0);
        fa.mode = pffused::kParentSelf;
        CUSMC_CHECK(cusmc_launch_fused(ctx, d, dy, nullptr, f->Qc0.data(), cfg.noise_scale, nullptr, nullptr,
                                       f->m0.data(), f->ep, fa, f->draws.xi0_dev == nullptr));
    } else {
        CUSMC_CHECK(cusmc_launch_step(ctx, d, dy, nullptr, f->Qc0.data(), cfg.noise_scale, nullptr, nullptr,
                                      f->m0.data(), f->ep, a, f->draws.xi0_dev == nullptr));
    }
    f->next_t = 1;
    f->ran = false;
    return CUSMC_OK;
}

// The systematic offset of step t: injected, or 53 Philox bits keyed by (seed, t).
static double filter_u0(const cusmc_filter *f, int t)
{
    return f->draws.u0_host ? f->draws.u0_host[t - 1]
                            : (double)(cusmc_u0_bits(f->cfg.seed, (uint64_t)t) >> 11) * 1.1102230246251565e-16;
}

static double filter_ess_bound(const cusmc_filter *f)
{
    return f->cfg.ess_threshold * (double)f->cfg.N * std::ldexp(1.0, f->shift);
}

// The tile update of step t (tile_update.cu), all of it or the phases named: global maximum, rescaled
// tile prefixes and totals, the constants / start tiles of step t + 1.
static int filter_tile_update(cusmc_filter *f, int t, int phases)
{
    const cusmc_filter_config &cfg = f->cfg;
    UpdateArgs u{};
    u.img = f->img[t & 1];
    u.slot = f->slots + t;
    u.slot_next = t + 1 < cfg.T ? f->slots + t + 1 : nullptr;
    u.rank_sums = f->rank_sums;
    u.tiles = (f->n + f->tile - 1) / f->tile;
    u.tiles_alloc = fimage_tiles(f->img_n);
    u.lo = (unsigned)f->lo;
    u.N_global = (unsigned)cfg.N;
    u.u0_next = (t + 1 < cfg.T && cfg.resampler == CUSMC_RESAMPLE_SYSTEMATIC) ? filter_u0(f, t + 1) : 0.0;
    u.ess_bound = cfg.ess_threshold > 0.0 ? filter_ess_bound(f) : 0.0;
    u.rank = f->rank;
    u.world = f->world;
    u.phases = phases;
    u.mail = filter_mail(f);
    u.cell_max = mail_cell(t, kCellMax, f->world);
    u.cell_sums = mail_cell(t, kCellSums, f->world);
    return cusmc_launch_tile_update(f->ctx, u);
}

static int filter_moments(cusmc_filter *f, int t)
{
    cusmc_ctx *ctx = f->ctx;
    const cusmc_filter_config &cfg = f->cfg;
    if (!cfg.summary || f->n == 0) return CUSMC_OK;
    const int mom_grid = (int)std::min<int64_t>((f->n + kThreads * 4 - 1) / (kThreads * 4), (int64_t)ctx->sm_count * 8);
    moments_kernel<<<mom_grid, kThreads, 0, ctx->stream>>>(f->x[f->cur], f->lw, &f->slots[t].lw_max, f->is_log, f->n,
                                                           f->per, cfg.d, f->moments + (size_t)t * (2 + cfg.d));
    CUSMC_LAUNCHED(ctx);
    return CUSMC_OK;
}

// After step t's particles exist: the weight image becomes global (log modes), posterior moments.
extern "C" int cusmc_filter_weigh(cusmc_filter *f, int t)
{
    if (!f) return CUSMC_ERR_INVALID;
    cusmc_ctx *ctx = f->ctx;
    const cusmc_filter_config &cfg = f->cfg;
    CUSMC_REQUIRE(ctx, t >= 0 && t < cfg.T && f->next_t == t + 1, "weigh(t) follows begin / propagate(t)");
    CUSMC_REQUIRE(ctx, f->world == 1 || f->fused || !f->is_log,
                  "a sharded filter driven phase by phase calls cusmc_filter_weigh_phase (scalar exchanges in between)");
    CUSMC_CUDA(ctx, cudaSetDevice(ctx->device));
    if (f->is_log) CUSMC_CHECK(filter_tile_update(f, t, kUpdAll));
    return filter_moments(f, t);
}

// The same in three phases, for callers that carry the scalars themselves (the NCCL formulation,
// cusmc_b200/sharded.py): phase 0 leaves this rank's maximum in slot[t] word 0 (all-reduce MAX it);
// phase 1 rescales and scans against that maximum and leaves this rank's sums in words 1..2 (all-gather
// words 1..3 of every rank into rank_sums_dev, 3 words per rank); phase 2 derives the global totals,
// rank offsets and the next step's constants from rank_sums_dev, then the moments.
extern "C" int cusmc_filter_weigh_phase(cusmc_filter *f, int t, int phase, const uint64_t *rank_sums_dev)
{
    if (!f) return CUSMC_ERR_INVALID;
    cusmc_ctx *ctx = f->ctx;
    const cusmc_filter_config &cfg = f->cfg;
    CUSMC_REQUIRE(ctx, t >= 0 && t < cfg.T && f->next_t == t + 1, "weigh(t) follows begin / propagate(t)");
    CUSMC_REQUIRE(ctx, phase >= 0 && phase <= 2, "phase must be 0, 1 or 2");
    CUSMC_CUDA(ctx, cudaSetDevice(ctx->device));
    if (!f->is_log) return phase == 2 ? filter_moments(f, t) : CUSMC_OK;
    if (phase == 2 && f->world > 1) {
        CUSMC_REQUIRE(ctx, rank_sums_dev != nullptr, "phase 2 needs the all-gathered sums");
        CUSMC_CUDA(ctx, cudaMemcpyAsync(f->rank_sums, rank_sums_dev, sizeof(uint64_t) * 3 * (size_t)f->world,
                                        cudaMemcpyDeviceToDevice, ctx->stream));
    }
    CUSMC_CHECK(filter_tile_update(f, t, phase == 0 ? kUpdMax : phase == 1 ? kUpdScan : kUpdConsts));
    return phase == 2 ? filter_moments(f, t) : CUSMC_OK;
}

// Ancestors of step t (src/mcmc.cpp:295) from the weights of step t - 1.  Systematic resampling has no
// pass of its own any more: every block of the fused step kernel looks its children's parents up in
// the weight image (pf_fused_impl.cuh), so this is a no-op for it.
extern "C" int cusmc_filter_resample(cusmc_filter *f, int t)
{
    if (!f) return CUSMC_ERR_INVALID;
    cusmc_ctx *ctx = f->ctx;
    const cusmc_filter_config &cfg = f->cfg;
    CUSMC_REQUIRE(ctx, t >= 1 && t < cfg.T && f->next_t == t, "resample(t) follows weigh(t - 1)");
    CUSMC_CUDA(ctx, cudaSetDevice(ctx->device));
    const cusmc_filter_draws &dr = f->draws;
    const int64_t n = f->n, N = cfg.N;
    const size_t off = (size_t)(t - 1);
    const bool sharded = f->world > 1;
    if (cfg.resampler == CUSMC_RESAMPLE_METROPOLIS || cfg.resampler == CUSMC_RESAMPLE_METROPOLIS_C2) {
        const double *u = dr.u_dev ? dr.u_dev + off * n * cfg.B : nullptr;
        const uint32_t *j = dr.j_dev ? dr.j_dev + off * n * cfg.B : nullptr;
        return cusmc_launch_metropolis(ctx, f->anc, f->lw, u, j, cfg.seed, (uint64_t)t, N, cfg.B, f->is_log,
                                       f->lo, n, sharded ? &f->peer_lw : nullptr,
                                       cfg.resampler == CUSMC_RESAMPLE_METROPOLIS_C2);
    }
    if (cfg.resampler == CUSMC_RESAMPLE_SYSTEMATIC) return CUSMC_OK;
    if (cfg.resampler == CUSMC_RESAMPLE_REJECTION)
        return cusmc_launch_rejection(ctx, f->anc, f->lw, &f->slots[t - 1].lw_max, cfg.seed, (uint64_t)t, N,
                                      kRejectionCap);
    // multinomial: every child searches the global integer CDF on the weight images themselves (its own
    // rank's or a peer's), tile_update.cu
    const double *um = dr.um_dev ? dr.um_dev + off * n : nullptr;
    const unsigned long long *const *peers =
        sharded ? (const unsigned long long *const *)f->peer_img[(t - 1) & 1].table_dev : nullptr;
    return cusmc_launch_multinomial_image(ctx, f->img[(t - 1) & 1], peers, f->img_n, n, f->lo, N, f->per, f->rank,
                                          f->world, um, cfg.seed, (uint64_t)t, f->anc,
                                          (unsigned long long *)&f->slots[t].degenerate);
}

// Propagate and reweight, fused (src/mcmc.cpp:298-307) -- and, for the normalised resamplers, the
// parent lookup before and the tile's weight image after, in the same kernel.  Sharded: every rank's
// image and state of step t - 1 must be complete (the exchanges inside weigh(t - 1) guarantee it).
extern "C" int cusmc_filter_propagate(cusmc_filter *f, int t)
{
    if (!f) return CUSMC_ERR_INVALID;
    cusmc_ctx *ctx = f->ctx;
    const cusmc_filter_config &cfg = f->cfg;
    CUSMC_REQUIRE(ctx, t >= 1 && t < cfg.T && f->next_t == t, "propagate(t) follows resample(t)");
    CUSMC_CUDA(ctx, cudaSetDevice(ctx->device));
    const cusmc_filter_draws &dr = f->draws;
    const int d = cfg.d, dy = cfg.dy;
    const int64_t n = f->n;
    const size_t off = (size_t)(t - 1);
    double c[CUSMC_MAX_DIM];
    cusmc_whiten_observation(f->Winv, dy, f->Y.data() + (size_t)t * dy, c);
    StepArgs a = filter_step_args(f, t);
    a.x_new = f->x[f->cur ^ 1];
    a.x_prev = f->x[f->cur];
    a.anc = f->anc;
    a.xi = dr.xi_dev ? dr.xi_dev + off * n * d : nullptr;
    a.chi = dr.chi_dev ? dr.chi_dev + off * n * d : nullptr;
    a.kind = cfg.kind;
    a.has_prev = 1;
    a.rng_stream = CUSMC_STREAM_NORMAL;
    if (cfg.keep_history) {       // the step's history rows are written by the step kernel itself
        const size_t row = f->hist_row(t);
        a.hist_x = f->hist_x + row * n * d;
        a.hist_w = f->hist_w + row * n;
        a.hist_a = f->hist_a + row * n;
    }
    if (f->world > 1) a.x_prev_peer = (const double *const *)f->peer_x[f->cur].table_dev;
    if (cfg.resampler == CUSMC_RESAMPLE_REJECTION) a.lw_max = &f->slots[t].lw_max;     // the next step's w_max
    if (f->is_log) {
        pffused::FusedArgs fa = filter_fused_args(f, a, t);
        if (cfg.resampler == CUSMC_RESAMPLE_SYSTEMATIC) {
            fa.mode = pffused::kParentLookup;
            fa.pdl = cfg.summary ? 0 : 1;        // the kernel right before it on the stream is the tile update
            // the parents it found: an OUTPUT only (nothing downstream reads them), so they are stored where
            // somebody can ask for them -- the last step (cusmc_filter_state_dev) and the history rows
            // (written through hist_a)
            fa.anc_out = t == cfg.T - 1 ? f->anc : nullptr;
            fa.s.anc = nullptr;
            fa.accumulate = cfg.ess_threshold > 0.0;
        } else {
            fa.mode = pffused::kParentArray;     // multinomial: ancestors from the search kernel
        }
        CUSMC_CHECK(cusmc_launch_fused(ctx, d, dy, f->G.data(), f->Qw.data(), cfg.noise_scale, &f->M, c, nullptr,
                                       f->ep, fa, a.xi == nullptr));
    } else {
        CUSMC_CHECK(cusmc_launch_step(ctx, d, dy, f->G.data(), f->Qw.data(), cfg.noise_scale, &f->M, c, nullptr,
                                      f->ep, a, a.xi == nullptr));
    }
    f->cur ^= 1;
    f->next_t = t + 1;
    return CUSMC_OK;
}

extern "C" int cusmc_filter_mark(cusmc_filter *f, int which)
{
    if (!f) return CUSMC_ERR_INVALID;
    CUSMC_CUDA(f->ctx, cudaEventRecord(which ? f->ev1 : f->ev0, f->ctx->stream));
    if (which) f->ran = true;
    return CUSMC_OK;
}

// ---- the sharded run ------------------------------------------------------------------------------
// One-warp barrier over the mailboxes (reference-mode runs: the peers read each other's densities and
// state buffers between launches; the normalised resamplers exchange inside tile_update_kernel).
__global__ void __launch_bounds__(32) barrier_kernel(const MailArgs m, size_t cell0)
{
    const int lane = threadIdx.x;
    mail_publish(m, cell0, lane, 0ull, 0ull, 0ull);
    unsigned long long w0, w1, w2;
    mail_wait(m, cell0, lane, w0, w1, w2);
}

static int launch_barrier(cusmc_filter *f, int cell, int t)
{
    cusmc_ctx *ctx = f->ctx;
    barrier_kernel<<<1, 32, 0, ctx->stream>>>(filter_mail(f), mail_cell(t, cell, f->world));
    CUSMC_LAUNCHED(ctx);
    return CUSMC_OK;
}

// The whole sharded run enqueued from C++: two launches per step (the fused step kernel, the tile
// update with its two scalar exchanges over the peer-memory mailboxes inside).  Every rank calls it
// the same number of times (the epoch must agree); returns after enqueueing.
//
// Ordering between ranks: rank A's step kernel of t + 1 reads the peers' images and states of step t
// and overwrites buffers the peers read during step t.  A's tile update of step t passes its first
// exchange only when every peer has published its maximum, i.e. finished ITS step kernel of t, and its
// second only when every peer's tile records of step t are complete -- so both hazards are covered
// without a barrier of their own.
extern "C" int cusmc_filter_run_sharded(cusmc_filter *f, const cusmc_filter_draws *draws)
{
    if (!f) return CUSMC_ERR_INVALID;
    CUSMC_REQUIRE(f->ctx, f->world > 1 && f->attached, "needs a sharded filter with attached peers");
    const bool is_log = f->is_log != 0;
    ++f->epoch;
    f->fused = true;
    auto after_weights = [&](int t) -> int {
        // reference mode: just a barrier (peers read these densities next)
        if (!is_log) CUSMC_CHECK(launch_barrier(f, kCellMax, t));
        return cusmc_filter_weigh(f, t);
    };
    int rc = cudaMemsetAsync(f->mail_err, 0, 8, f->ctx->stream) == cudaSuccess ? CUSMC_OK : CUSMC_ERR_CUDA;
    if (rc == CUSMC_OK) rc = cusmc_filter_begin(f, draws);
    if (rc == CUSMC_OK) rc = after_weights(0);
    if (rc == CUSMC_OK) rc = cusmc_filter_mark(f, 0);
    // CUSMC_SHARD_TRACE=1: device time of the step kernel and of the update (exchanges included) of steps
    // 5 .. 14 of this rank, on stderr (profiling aid; events on the stream, read after the run)
    const bool trace = std::getenv("CUSMC_SHARD_TRACE") != nullptr && f->cfg.T > 16;
    cudaEvent_t tev[10][3] = {};
    if (trace)
        for (auto &row : tev)
            for (auto &e : row) cudaEventCreate(&e);
    for (int t = 1; t < f->cfg.T && rc == CUSMC_OK; ++t) {
        const bool stamp = trace && t >= 5 && t < 15;
        rc = cusmc_filter_resample(f, t);
        // reference mode: the ancestors are local, but the step kernel overwrites the state buffer the
        // peers gathered from during the previous step
        if (rc == CUSMC_OK && !is_log) rc = launch_barrier(f, kCellBarrier, t);
        if (stamp) cudaEventRecord(tev[t - 5][0], f->ctx->stream);
        if (rc == CUSMC_OK) rc = cusmc_filter_propagate(f, t);
        if (stamp) cudaEventRecord(tev[t - 5][1], f->ctx->stream);
        if (rc == CUSMC_OK) rc = after_weights(t);
        if (stamp) cudaEventRecord(tev[t - 5][2], f->ctx->stream);
    }
    if (rc == CUSMC_OK) rc = cusmc_filter_mark(f, 1);
    if (trace) {
        cudaStreamSynchronize(f->ctx->stream);
        float k1 = 0.f, k2 = 0.f, ms = 0.f;
        for (auto &row : tev) {
            cudaEventElapsedTime(&ms, row[0], row[1]);
            k1 += ms;
            cudaEventElapsedTime(&ms, row[1], row[2]);
            k2 += ms;
            for (auto &e : row) cudaEventDestroy(e);
        }
        std::fprintf(stderr, "rank %d: step kernel %.1f us, update + exchanges %.1f us (mean of steps 5..14)\n", f->rank,
                     k1 * 100.0, k2 * 100.0);
    }
    f->fused = false;
    return rc;
}

// Non-zero if a spin-wait of the last sharded run timed out (a peer never arrived).
extern "C" int cusmc_filter_exchange_status(cusmc_filter *f, uint64_t *status)
{
    if (!f) return CUSMC_ERR_INVALID;
    CUSMC_REQUIRE(f->ctx, status != nullptr, "NULL pointer");
    *status = 0;
    if (!f->mail_err) return CUSMC_OK;
    CUSMC_CUDA(f->ctx, cudaStreamSynchronize(f->ctx->stream));
    CUSMC_CUDA(f->ctx, cudaMemcpy(status, f->mail_err, 8, cudaMemcpyDeviceToHost));
    return CUSMC_OK;
}

extern "C" int cusmc_filter_run(cusmc_filter *f, const cusmc_filter_draws *draws)
{
    if (!f) return CUSMC_ERR_INVALID;
    CUSMC_REQUIRE(f->ctx, f->world == 1, "a sharded filter is driven phase by phase (cusmc_b200/sharded.py)");
    CUSMC_CUDA(f->ctx, cudaSetDevice(f->ctx->device));
    if (cusmc_filter_persistent_eligible(f, draws)) {
        const int rc = cusmc_filter_run_persistent(f, draws);
        if (rc != CUSMC_ERR_UNSUPPORTED) return rc;      // unsupported after all (occupancy): fall through
    }
    CUSMC_CHECK(cusmc_filter_begin(f, draws));
    CUSMC_CHECK(cusmc_filter_weigh(f, 0));
    CUSMC_CHECK(cusmc_filter_mark(f, 0));
    for (int t = 1; t < f->cfg.T; ++t) {
        CUSMC_CHECK(cusmc_filter_resample(f, t));
        CUSMC_CHECK(cusmc_filter_propagate(f, t));
        CUSMC_CHECK(cusmc_filter_weigh(f, t));
    }
    return cusmc_filter_mark(f, 1);
}

// What went wrong inside the last run, read back from the device: a scalar exchange that timed out
// (sharded runs) or a step whose weights had no mass to resample from.  Every getter that hands
// results to the caller goes through this, so a void run cannot be consumed silently.
static int filter_run_status(cusmc_filter *f)
{
    cusmc_ctx *ctx = f->ctx;
    CUSMC_CUDA(ctx, cudaSetDevice(ctx->device));
    CUSMC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (f->mail_err) {
        unsigned long long err = 0;
        CUSMC_CUDA(ctx, cudaMemcpy(&err, f->mail_err, 8, cudaMemcpyDeviceToHost));
        if (err)
            return cusmc_fail(ctx, CUSMC_ERR_TIMEOUT, "sharded run: a scalar exchange timed out after %.1f s "
                              "(a peer never arrived); the results are void", f->mail_timeout_ns * 1e-9);
    }
    const int T = f->cfg.T;
    std::vector<StepSlot> slots(T);
    CUSMC_CUDA(ctx, cudaMemcpy(slots.data(), f->slots, sizeof(StepSlot) * T, cudaMemcpyDeviceToHost));
    for (int t = 1; t < T; ++t)
        if (slots[t].degenerate)
            return cusmc_fail(ctx, CUSMC_ERR_DEGENERATE, "step %d: every weight of step %d is zero or non-finite "
                              "(nothing to resample from); ancestors were kept as the identity", t, t - 1);
    return CUSMC_OK;
}

extern "C" int cusmc_filter_status(cusmc_filter *f)
{
    if (!f) return CUSMC_ERR_INVALID;
    CUSMC_REQUIRE(f->ctx, f->ran, "filter has not run");
    return filter_run_status(f);
}

extern "C" int cusmc_filter_set_exchange_timeout(cusmc_filter *f, double seconds)
{
    if (!f) return CUSMC_ERR_INVALID;
    CUSMC_REQUIRE(f->ctx, seconds > 0.0 && seconds < 3600.0, "timeout must lie in (0, 3600) seconds");
    f->mail_timeout_ns = (unsigned long long)(seconds * 1e9);
    return CUSMC_OK;
}

extern "C" double cusmc_filter_last_ms(const cusmc_filter *f)
{
    if (!f || !f->ran) return 0.0;
    cudaEventSynchronize(f->ev1);
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, f->ev0, f->ev1) != cudaSuccess) return 0.0;
    return ms;
}

extern "C" int cusmc_filter_get_summary(cusmc_filter *f, double *mean, double *ess, double *loglik)
{
    if (!f) return CUSMC_ERR_INVALID;
    cusmc_ctx *ctx = f->ctx;
    CUSMC_REQUIRE(ctx, f->ran, "filter has not run");
    CUSMC_CHECK(filter_run_status(f));
    const int d = f->cfg.d, T = f->cfg.T;
    std::vector<StepSlot> slots(T);
    std::vector<double> mom((size_t)T * (2 + d));
    CUSMC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    CUSMC_CUDA(ctx, cudaMemcpy(slots.data(), f->slots, sizeof(StepSlot) * T, cudaMemcpyDeviceToHost));
    CUSMC_CUDA(ctx, cudaMemcpy(mom.data(), f->moments, sizeof(double) * mom.size(), cudaMemcpyDeviceToHost));
    const double scale = std::ldexp(1.0, f->shift);
    for (int t = 0; t < T; ++t) {
        const double *m = &mom[(size_t)t * (2 + d)];
        if (mean)
            for (int k = 0; k < d; ++k) mean[(size_t)t * d + k] = f->cfg.summary ? m[2 + k] / m[0] : NAN;
        if (f->is_log) {
            const double sq = (double)slots[t].sum_q, sq2 = (double)slots[t].sum_q2;
            if (ess) ess[t] = sq * sq / (sq2 * scale);
            if (loglik) loglik[t] = slots[t].lw_max + std::log(sq / scale / (double)f->cfg.N);
        } else {
            if (ess) ess[t] = f->cfg.summary ? m[0] * m[0] / m[1] : NAN;
            if (loglik) loglik[t] = f->cfg.summary ? std::log(m[0] / (double)f->cfg.N) : NAN;
        }
    }
    return CUSMC_OK;
}

extern "C" int cusmc_filter_get_resampled(cusmc_filter *f, int *resampled)
{
    if (!f) return CUSMC_ERR_INVALID;
    cusmc_ctx *ctx = f->ctx;
    CUSMC_REQUIRE(ctx, f->ran && resampled, "filter has not run");
    const int T = f->cfg.T;
    std::vector<StepSlot> slots(T);
    CUSMC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    CUSMC_CUDA(ctx, cudaMemcpy(slots.data(), f->slots, sizeof(StepSlot) * T, cudaMemcpyDeviceToHost));
    for (int t = 0; t < T; ++t) resampled[t] = (int)slots[t].resampled;
    return CUSMC_OK;
}

// History weights as the caller sees them (the reference's w_t, src/mcmc.cpp:85,212; R-level
// `weights`, src/run.rcpp.cpp:110-117): reference mode ("metropolis") keeps the raw densities with
// w_0 = 1/N exactly as the reference does; the normalised resamplers keep log-weights internally and
// hand out NORMALISED weights  w_i = exp(lw_i - max) 2^shift / sum_q  (they sum to one over the whole
// cloud; 1/N at t = 0, and cumulative over steps that did not resample).  Rows [row0, row0 + rows) of
// `lw` (n columns each) belong to steps t0 .. t0 + rows - 1.
__global__ void __launch_bounds__(kThreads)
normalise_rows_kernel(const double *__restrict__ lw, double *__restrict__ out, const StepSlot *__restrict__ slots,
                      int t0, int64_t n, int64_t total, double two_shift)
{
    for (int64_t e = (int64_t)blockIdx.x * kThreads + threadIdx.x; e < total; e += (int64_t)gridDim.x * kThreads) {
        const int r = (int)(e / n);
        const StepSlot s = slots[t0 + r];
        const double v = exp(lw[e] - s.lw_max) * (two_shift / (double)s.sum_q);
        out[e] = (v == v) ? v : 0.0;
    }
}

static int launch_normalise_rows(cusmc_filter *f, const double *lw, double *out, int t0, int rows, cudaStream_t st)
{
    const int64_t total = (int64_t)rows * f->n;
    if (total == 0) return CUSMC_OK;
    const int grid = (int)std::min<int64_t>((total + kThreads - 1) / kThreads, (int64_t)f->ctx->sm_count * 16);
    normalise_rows_kernel<<<grid, kThreads, 0, st>>>(lw, out, f->slots, t0, f->n, total, std::ldexp(1.0, f->shift));
    CUSMC_LAUNCHED(f->ctx);
    return CUSMC_OK;
}

extern "C" int cusmc_filter_get_history(cusmc_filter *f, double *x_aos, double *w, uint32_t *a)
{
    if (!f) return CUSMC_ERR_INVALID;
    cusmc_ctx *ctx = f->ctx;
    CUSMC_REQUIRE(ctx, f->ran && f->cfg.keep_history > 0, "history was not kept (keep_history = 1)");
    CUSMC_CHECK(filter_run_status(f));
    const int T = f->cfg.T;
    const size_t TN = (size_t)T * (size_t)f->n;   // this rank's shard
    if (x_aos) CUSMC_CHECK(cusmc_d2h_staged(ctx, x_aos, f->hist_x, sizeof(double) * TN * f->cfg.d));
    if (w && !f->is_log) CUSMC_CHECK(cusmc_d2h_staged(ctx, w, f->hist_w, sizeof(double) * TN));
    if (w && f->is_log) {
        // normalised on the device, a bounded number of rows at a time
        const int rows_max = (int)std::max<int64_t>(1, std::min<int64_t>(T, ((int64_t)4 << 20) / std::max<int64_t>(1, f->n)));
        void *tmp = nullptr;
        CUSMC_CHECK(cusmc_scratch(ctx, 0, sizeof(double) * (size_t)rows_max * (size_t)f->n, &tmp));
        for (int t0 = 0; t0 < T; t0 += rows_max) {
            const int rows = std::min(rows_max, T - t0);
            CUSMC_CHECK(launch_normalise_rows(f, f->hist_w + (size_t)t0 * f->n, (double *)tmp, t0, rows, ctx->stream));
            CUSMC_CHECK(cusmc_d2h_staged(ctx, w + (size_t)t0 * f->n, tmp, sizeof(double) * (size_t)rows * f->n));
        }
    }
    if (a) {
        CUSMC_CHECK(cusmc_d2h_staged(ctx, a, f->hist_a, sizeof(uint32_t) * TN));
        for (int64_t i = 0; i < f->n; ++i) a[i] = (uint32_t)(f->lo + i);   // row t = 0: identity
    }
    return CUSMC_OK;
}

// The raw per-step log-weights [T][n] of the normalised resamplers (what resampling consumed; row 0 is 0).
extern "C" int cusmc_filter_get_log_weights(cusmc_filter *f, double *lw)
{
    if (!f) return CUSMC_ERR_INVALID;
    cusmc_ctx *ctx = f->ctx;
    CUSMC_REQUIRE(ctx, f->ran && f->cfg.keep_history > 0 && lw, "history was not kept (keep_history = 1)");
    CUSMC_REQUIRE(ctx, f->is_log, "the metropolis (reference) mode keeps densities, not log-weights");
    CUSMC_CHECK(filter_run_status(f));
    return cusmc_d2h_staged(ctx, lw, f->hist_w, sizeof(double) * (size_t)f->cfg.T * (size_t)f->n);
}

// Genealogy: thread i walks the ancestor rows back from the final step.
__global__ void __launch_bounds__(kThreads)
lineage_kernel(const uint32_t *__restrict__ hist_a, int T, int64_t n, uint32_t *__restrict__ lineage)
{
    const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (i >= n) return;
    uint32_t k = (uint32_t)i;
    lineage[(size_t)(T - 1) * n + i] = k;
    for (int t = T - 1; t >= 1; --t) {
        k = __ldg(hist_a + (size_t)t * n + k);
        lineage[(size_t)(t - 1) * n + i] = k;
    }
}

extern "C" int cusmc_filter_get_lineage(cusmc_filter *f, uint32_t *lineage, int *n_unique)
{
    if (!f) return CUSMC_ERR_INVALID;
    cusmc_ctx *ctx = f->ctx;
    CUSMC_REQUIRE(ctx, f->ran && f->cfg.keep_history > 0 && lineage, "history was not kept (keep_history = 1)");
    CUSMC_REQUIRE(ctx, f->world == 1, "the lineage is traced on one GPU");
    CUSMC_CHECK(filter_run_status(f));
    const int T = f->cfg.T;
    const int64_t n = f->n;
    void *tmp = nullptr;
    CUSMC_CHECK(cusmc_scratch(ctx, 0, sizeof(uint32_t) * (size_t)T * (size_t)n, &tmp));
    lineage_kernel<<<(unsigned)((n + kThreads - 1) / kThreads), kThreads, 0, ctx->stream>>>(f->hist_a, T, n, (uint32_t *)tmp);
    CUSMC_LAUNCHED(ctx);
    CUSMC_CHECK(cusmc_d2h_staged(ctx, lineage, tmp, sizeof(uint32_t) * (size_t)T * (size_t)n));
    if (n_unique)                       // rows are non-decreasing in i (ancestors are monotone): count the steps
        for (int t = 0; t < T; ++t) {
            const uint32_t *row = lineage + (size_t)t * n;
            int u = n > 0 ? 1 : 0;
            bool sorted = true;
            for (int64_t i = 1; i < n; ++i) {
                u += row[i] != row[i - 1];
                sorted = sorted && row[i] >= row[i - 1];
            }
            if (!sorted) {              // resamplers without monotone ancestors (metropolis, rejection, multinomial)
                std::vector<uint32_t> v(row, row + n);
                std::sort(v.begin(), v.end());
                u = (int)(std::unique(v.begin(), v.end()) - v.begin());
            }
            n_unique[t] = u;
        }
    return CUSMC_OK;
}

extern "C" int cusmc_filter_state_dev(cusmc_filter *f, double **x_soa, double **lw, uint32_t **anc)
{
    if (!f) return CUSMC_ERR_INVALID;
    if (x_soa) *x_soa = f->x[f->cur];
    if (lw) *lw = f->lw;
    if (anc) *anc = f->anc;
    return CUSMC_OK;
}

// R-level run() (src/run.rcpp.cpp:58-126): weights [T][N], posterior_x [T][N][d] (+ optionally the
// ancestors [T][N]) on HOST pointers.  The reference keeps every step of every particle alive on the
// host as separately allocated vectors (src/run.rcpp.cpp:82-97) and its GPU build ships the whole
// cloud both ways every step; here the history STREAMS: the step kernels write their rows into a
// two-chunk ring on the device, a finished chunk goes device -> pinned host memory on a second
// stream while the next chunk computes, and this thread (with helpers) copies drained chunks into
// the caller's arrays.  Device memory is bounded by the ring (2 chunks of ~8 MB of rows), whatever T.
namespace {

struct RunRing {
    cusmc_filter *f = nullptr;
    int K = 1, chunks = 0;
    size_t n = 0, d = 0;
    void *pin[2] = {nullptr, nullptr};
    size_t pin_bytes = 0;
    cudaEvent_t ready[2] = {nullptr, nullptr}, copied[2] = {nullptr, nullptr};
    double *weights = nullptr, *posterior_x = nullptr;
    uint32_t *ancestors = nullptr;
    // the drain crew: `workers` host threads, alive for the whole run, each copying ITS slice of every
    // landed chunk into the caller's arrays while the calling thread keeps enqueueing steps.  (Spawning
    // helpers per chunk and draining from the enqueueing thread made the host side the critical path:
    // 51 ms for the C1 run, of which the GPU needed 12.)
    int workers = 1;
    std::atomic<int> shipped{0};      // chunks whose device -> pinned copy has been ENQUEUED (by the caller's thread)
    std::atomic<int> landed{0};       // chunks whose copy has COMPLETED (worker 0 waits on the event)
    std::atomic<long long> drained{0};   // worker-chunks finished: chunk c is out of its slot when drained >= workers (c + 1)
    std::atomic<int> abort{0};
    std::vector<std::thread> crew;

    size_t off_w() const { return sizeof(double) * (size_t)K * n * d; }
    size_t off_a() const { return off_w() + sizeof(double) * (size_t)K * n; }
    int rows_of(int c) const { return std::min(K, f->cfg.T - c * K); }

    // chunk c is complete on the compute stream: normalise its weights (log modes), ship it
    int ship(int c)
    {
        cusmc_ctx *ctx = f->ctx;
        const int s = c & 1, rows = rows_of(c);
        double *dx = f->hist_x + (size_t)s * K * n * d, *dw = f->hist_w + (size_t)s * K * n;
        uint32_t *da = f->hist_a + (size_t)s * K * n;
        if (f->is_log && weights) CUSMC_CHECK(launch_normalise_rows(f, dw, dw, c * K, rows, ctx->stream));
        CUSMC_CUDA(ctx, cudaEventRecord(ready[s], ctx->stream));
        CUSMC_CUDA(ctx, cudaStreamWaitEvent(ctx->aux_stream, ready[s], 0));
        char *p = (char *)pin[s];
        if (posterior_x)
            CUSMC_CUDA(ctx, cudaMemcpyAsync(p, dx, sizeof(double) * (size_t)rows * n * d, cudaMemcpyDeviceToHost, ctx->aux_stream));
        if (weights)
            CUSMC_CUDA(ctx, cudaMemcpyAsync(p + off_w(), dw, sizeof(double) * (size_t)rows * n, cudaMemcpyDeviceToHost, ctx->aux_stream));
        if (ancestors)
            CUSMC_CUDA(ctx, cudaMemcpyAsync(p + off_a(), da, sizeof(uint32_t) * (size_t)rows * n, cudaMemcpyDeviceToHost, ctx->aux_stream));
        CUSMC_CUDA(ctx, cudaEventRecord(copied[s], ctx->aux_stream));
        shipped.store(c + 1, std::memory_order_release);
        return CUSMC_OK;
    }

    static void slice_copy(char *dst, const char *src, size_t bytes, int w, int W)
    {
        const size_t part = ((bytes + W - 1) / W + 4095) & ~(size_t)4095;      // whole pages per worker
        const size_t lo = (size_t)w * part;
        if (lo < bytes) std::memcpy(dst + lo, src + lo, std::min(part, bytes - lo));
    }

    void work(int w)
    {
        for (int c = 0; c < chunks; ++c) {
            if (w == 0) {
                while (shipped.load(std::memory_order_acquire) <= c && !abort.load()) std::this_thread::yield();
                if (abort.load()) return;
                if (cudaEventSynchronize(copied[c & 1]) != cudaSuccess) abort.store(1);
                landed.store(c + 1, std::memory_order_release);
            } else {
                while (landed.load(std::memory_order_acquire) <= c && !abort.load()) std::this_thread::yield();
            }
            if (abort.load()) return;
            const int rows = rows_of(c);
            const char *p = (const char *)pin[c & 1];
            const size_t t0 = (size_t)c * K;
            if (posterior_x) slice_copy((char *)(posterior_x + t0 * n * d), p, sizeof(double) * (size_t)rows * n * d, w, workers);
            if (weights) slice_copy((char *)(weights + t0 * n), p + off_w(), sizeof(double) * (size_t)rows * n, w, workers);
            if (ancestors) slice_copy((char *)(ancestors + t0 * n), p + off_a(), sizeof(uint32_t) * (size_t)rows * n, w, workers);
            drained.fetch_add(1, std::memory_order_release);
        }
    }

    // blocks the caller's thread until chunk c has left its ring slot (every worker is done with it)
    bool wait_drained(int c)
    {
        while (drained.load(std::memory_order_acquire) < (long long)workers * (c + 1)) {
            if (abort.load()) return false;
            std::this_thread::yield();
        }
        return true;
    }

    void start()
    {
        const unsigned hw = std::thread::hardware_concurrency();
        workers = (int)std::max(1u, std::min(8u, hw ? hw / 2 : 4u));
        try {                                            // nothing may unwind across the C ABI
            for (int w = 0; w < workers; ++w) crew.emplace_back([this, w] { work(w); });
        } catch (...) {
            // fewer threads than hoped for: the slices are fixed, so finish what exists and fall back to one
            abort.store(1);
            for (auto &t : crew) t.join();
            crew.clear();
            abort.store(0);
            shipped.store(0);
            landed.store(0);
            drained.store(0);
            workers = 1;
            crew.emplace_back([this] { work(0); });
        }
    }

    void release()
    {
        for (auto &t : crew) t.join();
        for (int s = 0; s < 2; ++s) {          // (the pinned slots are the context's)
            if (ready[s]) cudaEventDestroy(ready[s]);
            if (copied[s]) cudaEventDestroy(copied[s]);
        }
    }
};

int run_streamed(cusmc_ctx *ctx, const cusmc_filter_config *cfg, double *weights, double *posterior_x,
                 uint32_t *ancestors)
{
    cusmc_filter_config c = *cfg;
    // CUSMC_RUN_TRACE=1: wall-clock of the phases of this call on stderr (profiles/c1_breakdown.py)
    const bool trace = std::getenv("CUSMC_RUN_TRACE") != nullptr;
    const auto t_start = std::chrono::steady_clock::now();
    auto lap = [&](const char *what) {
        if (trace)
            std::fprintf(stderr, "cusmc_run %-28s %8.2f ms\n", what,
                         std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_start).count());
    };
    CUSMC_REQUIRE(ctx, c.world <= 1, "cusmc_run is single-GPU (a sharded run is driven through cusmc_filter_run_sharded)");
    CUSMC_REQUIRE(ctx, c.N >= 1 && c.d >= 1 && c.T >= 1, "N, d, T must be positive");
    const size_t row_bytes = (size_t)c.N * (sizeof(double) * (size_t)c.d + sizeof(double) + sizeof(uint32_t));
    const int K = (int)std::max<size_t>(1, std::min<size_t>((size_t)c.T, ((size_t)8 << 20) / std::max<size_t>(1, row_bytes)));
    c.keep_history = -K;
    c.persistent = -1;            // the history rows are written by the per-step kernels
    cusmc_filter *f = nullptr;
    CUSMC_CHECK(cusmc_filter_create(ctx, &c, &f));
    lap("filter created");
    RunRing ring;
    ring.f = f;
    ring.K = K;
    ring.chunks = (c.T + K - 1) / K;
    ring.n = (size_t)c.N;
    ring.d = (size_t)c.d;
    ring.weights = weights;
    ring.posterior_x = posterior_x;
    ring.ancestors = ancestors;
    ring.pin_bytes = (size_t)K * row_bytes;
    int rc = cusmc_aux_stream(ctx);
    // the two pinned slots belong to the CONTEXT (grow-only): page-locking 2 x 8 MB costs ~20 ms, more than
    // the whole C1 run, so only the first call on a context pays for it
    void *pin_base = nullptr;
    const size_t slot_bytes = (ring.pin_bytes + 4095) & ~(size_t)4095;
    if (rc == CUSMC_OK) rc = cusmc_pinned(ctx, 2 * slot_bytes, &pin_base);
    for (int s = 0; s < 2 && rc == CUSMC_OK; ++s) {
        ring.pin[s] = (char *)pin_base + (size_t)s * slot_bytes;
        if (cudaEventCreateWithFlags(&ring.ready[s], cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&ring.copied[s], cudaEventDisableTiming) != cudaSuccess)
            rc = cusmc_fail(ctx, CUSMC_ERR_CUDA, "cusmc_run: event creation failed: %s", cudaGetErrorString(cudaGetLastError()));
    }
    lap("pinned ring allocated");
    if (rc == CUSMC_OK) ring.start();
    const int T = c.T;
    for (int t = 0; t < T && rc == CUSMC_OK; ++t) {
        const int ch = t / K;
        // the ring slot of this chunk was last used by chunk ch - 2: it must have been drained
        if (t % K == 0 && ch >= 2 && !ring.wait_drained(ch - 2))
            rc = cusmc_fail(ctx, CUSMC_ERR_CUDA, "cusmc_run: the device -> host copy of the history failed");
        if (rc != CUSMC_OK) break;
        if (t == 0) {
            rc = cusmc_filter_begin(f, nullptr);
            if (rc == CUSMC_OK) rc = cusmc_filter_weigh(f, 0);
            if (rc == CUSMC_OK) rc = cusmc_filter_mark(f, 0);
        } else {
            rc = cusmc_filter_resample(f, t);
            if (rc == CUSMC_OK) rc = cusmc_filter_propagate(f, t);
            if (rc == CUSMC_OK) rc = cusmc_filter_weigh(f, t);
        }
        if (rc == CUSMC_OK && (t % K == K - 1 || t == T - 1)) rc = ring.ship(ch);
    }
    if (rc == CUSMC_OK) rc = cusmc_filter_mark(f, 1);
    lap("all steps enqueued");
    if (rc == CUSMC_OK && !ring.wait_drained(ring.chunks - 1))
        rc = cusmc_fail(ctx, CUSMC_ERR_CUDA, "cusmc_run: the device -> host copy of the history failed");
    if (rc != CUSMC_OK) ring.abort.store(1);
    if (rc == CUSMC_OK && ancestors)
        for (size_t i = 0; i < ring.n; ++i) ancestors[i] = (uint32_t)i;      // row t = 0: identity
    lap("history drained");
    if (rc == CUSMC_OK) rc = filter_run_status(f);
    cudaStreamSynchronize(ctx->aux_stream);
    cudaStreamSynchronize(ctx->stream);
    ring.release();
    lap("ring released");
    cusmc_filter_destroy(f);
    lap("filter destroyed");
    return rc;
}

}  // namespace

extern "C" int cusmc_run(cusmc_ctx *ctx, const cusmc_filter_config *cfg, double *weights, double *posterior_x)
{
    CUSMC_ENTER(ctx);
    CUSMC_REQUIRE(ctx, cfg != nullptr, "config is NULL");
    return run_streamed(ctx, cfg, weights, posterior_x, nullptr);
}

// The same run, also returning the ancestor indices a_t [T][N] (row 0 is the identity): with them the
// caller can trace any particle's genealogy through posterior_x (the reference keeps a_t as well,
// src/run.rcpp.cpp:97, but never returns it).
extern "C" int cusmc_run_ancestors(cusmc_ctx *ctx, const cusmc_filter_config *cfg, double *weights,
                                   double *posterior_x, uint32_t *ancestors)
{
    CUSMC_ENTER(ctx);
    CUSMC_REQUIRE(ctx, cfg != nullptr, "config is NULL");
    return run_streamed(ctx, cfg, weights, posterior_x, ancestors);
}

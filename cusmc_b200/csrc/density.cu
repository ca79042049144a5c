// density.cu -- batched MVN / MVT (log-)density kernels, shared covariance.
//
// Replaces mvn_pdf_kernel_y_minus_Fmu / _Einv_alpha / mvn_pdf_kernel and the mvt_ twins
// (src/mvn_dist.cu.cpp:455-668, src/mvt_dist.cu.cpp:356-571: three launches, one block
// per particle, Sigma^-1 re-read from global by every block) with ONE streaming kernel:
// a thread owns one or two whole points, the whitening operator sits in the parameter
// bank, HBM traffic is exactly 8d bytes in and 8 bytes out per point.
#include "density.cuh"
#include "hostmath.h"

#include <algorithm>
#include <new>
#include <vector>

namespace {

constexpr int kThreads = 256;
#ifndef CUSMC_DENSITY_PDL
#define CUSMC_DENSITY_PDL 1
#endif

// ---- SoA: x[j*ld + i].  VEC = 2 -> 128-bit loads, a thread owns points 2u and 2u+1. ----
// One unit per thread and no grid-stride loop ON PURPOSE: with a loop (grid-stride or a
// persistent TMA/mbarrier ring -- both tried, profiles/micro/density_variants.cu) LICM hoists the
// whole operator out of it into registers (96-170 regs, spills).  Straight-line code lets ptxas
// feed each DFMA its coefficient from the constant bank through a uniform register
// (LDCU -> UR).  Shapes were picked on the B200 with that microbenchmark: at d = 16 one point
// per thread at 56 registers (4 blocks/SM) reaches the read-only streaming ceiling of a
// 142 MB pass (0.91 of the measured copy bandwidth); two points per thread pay off for d <= 8
// where a point is only a few loads.  EXACT (d == D) drops every j < d predicate.
template <int D, bool TRI, int VEC, bool EXACT>
__global__ void __launch_bounds__(kThreads, (D >= 32 ? 2 : 4))
density_soa_kernel(const __grid_constant__ AffineOp<D, TRI> op, const Epilogue ep,
                   const double *__restrict__ x, int64_t n_units, int64_t ld, int d,
                   double *__restrict__ out)
{
#if CUSMC_DENSITY_PDL
    cusmc_pdl_enter();      // back-to-back density calls: the block dispatch of call n + 1 overlaps the tail of call n
#endif
    const int64_t u = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (u >= n_units) return;
    if constexpr (VEC == 2) {
        const int64_t i = 2 * u;
        double ra[D], rb[D];
#pragma unroll
        for (int j = 0; j < D; ++j) {
            if (EXACT || j < d) {
                const double2 v = ld_stream2(x + (int64_t)j * ld + i);
                ra[j] = v.x - op.shift[j];
                rb[j] = v.y - op.shift[j];
            } else {
                ra[j] = 0.0;
                rb[j] = 0.0;
            }
        }
        double2 res;
        res.x = density_epilogue(ep, affine_quadform<D, TRI>(op, ra));
        res.y = density_epilogue(ep, affine_quadform<D, TRI>(op, rb));
        st_stream2(out + i, res);
    } else {
        double r[D];
#pragma unroll
        for (int j = 0; j < D; ++j)
            r[j] = (EXACT || j < d) ? ld_stream(x + (int64_t)j * ld + u) - op.shift[j] : 0.0;
        st_stream(out + u, density_epilogue(ep, affine_quadform<D, TRI>(op, r)));
    }
}

// ---- AoS: x[i*d + j] (the layout the reference's host code hands over).  A block stages one
// tile of 256 points through shared memory with fully coalesced loads; the odd row pitch
// makes the per-thread row reads bank-conflict free. ------------------------------------
template <int D, bool TRI>
__global__ void __launch_bounds__(kThreads, (D >= 16 ? 3 : 4))
density_aos_kernel(const __grid_constant__ AffineOp<D, TRI> op, const Epilogue ep,
                   const double *__restrict__ x, int64_t N, int d, int vec_ok,
                   double *__restrict__ out)
{
    extern __shared__ double tile[];
    const int pitch = d | 1;
    const int64_t base = (int64_t)blockIdx.x * kThreads;
    const int npts = (int)((N - base) < kThreads ? (N - base) : kThreads);
    const double *src = x + base * d;
    const int n_el = npts * d;
    if (vec_ok) {   // d even and x 16-byte aligned: the tile start is 16-byte aligned too
        for (int e = 2 * threadIdx.x; e < n_el; e += 2 * kThreads) {
            const double2 v = ld_stream2(src + e);
            const int row = e / d, col = e - row * d;   // col even, col + 1 < d
            tile[row * pitch + col] = v.x;
            tile[row * pitch + col + 1] = v.y;
        }
    } else {
        for (int e = threadIdx.x; e < n_el; e += kThreads) {
            const int row = e / d, col = e - row * d;
            tile[row * pitch + col] = ld_stream(src + e);
        }
    }
    __syncthreads();
    if ((int)threadIdx.x < npts) {
        double r[D];
#pragma unroll
        for (int j = 0; j < D; ++j)
            r[j] = (j < d) ? tile[threadIdx.x * pitch + j] - op.shift[j] : 0.0;
        st_stream(out + base + threadIdx.x, density_epilogue(ep, affine_quadform<D, TRI>(op, r)));
    }
}

// SoA launch.  CUSMC_DENSITY_PDL: launched as a programmatic dependent of whatever precedes it on the
// stream -- back-to-back density calls overlap the block dispatch of call n + 1 with the tail of call n.
template <int D, bool TRI, int VEC, bool EXACT>
cudaError_t launch_soa(cusmc_ctx *ctx, unsigned grid, const AffineOp<D, TRI> &op, const Epilogue &ep, const double *x,
                       int64_t units, int64_t ld, int d, double *out)
{
#if CUSMC_DENSITY_PDL
    return cusmc_launch_pdl(density_soa_kernel<D, TRI, VEC, EXACT>, grid, kThreads, 0, ctx->stream, op, ep, x, units, ld, d, out);
#else
    density_soa_kernel<D, TRI, VEC, EXACT><<<grid, kThreads, 0, ctx->stream>>>(op, ep, x, units, ld, d, out);
    return cudaSuccess;
#endif
}

template <int D, bool TRI>
int launch_density(cusmc_ctx *ctx, const AffineOp<D, TRI> &op, const Epilogue &ep,
                   const double *x, int layout, int64_t N, int64_t ld, int d, double *out)
{
    if (N == 0) return CUSMC_OK;
    if (layout == CUSMC_SOA) {
        const bool exact = d == D;
        const bool vec2 = D <= 8 && (N % 2 == 0) && (ld % 2 == 0) &&
                          ((uintptr_t)x % 16 == 0) && ((uintptr_t)out % 16 == 0);
        if constexpr (D <= 8) if (vec2) {
            const int64_t units = N / 2;
            const unsigned grid = (unsigned)((units + kThreads - 1) / kThreads);
            if (exact)
                CUSMC_CUDA(ctx, (launch_soa<D, TRI, 2, true>(ctx, grid, op, ep, x, units, ld, d, out)));
            else
                CUSMC_CUDA(ctx, (launch_soa<D, TRI, 2, false>(ctx, grid, op, ep, x, units, ld, d, out)));
            CUSMC_LAUNCHED(ctx);
            return CUSMC_OK;
        }
        const unsigned grid = (unsigned)((N + kThreads - 1) / kThreads);
        if (exact)
            CUSMC_CUDA(ctx, (launch_soa<D, TRI, 1, true>(ctx, grid, op, ep, x, N, ld, d, out)));
        else
            CUSMC_CUDA(ctx, (launch_soa<D, TRI, 1, false>(ctx, grid, op, ep, x, N, ld, d, out)));
    } else {
        const size_t smem = sizeof(double) * kThreads * (size_t)(d | 1);
        const int64_t grid = (N + kThreads - 1) / kThreads;
        const int vec_ok = (d % 2 == 0) && ((uintptr_t)x % 16 == 0);
        if (smem > 48 * 1024)     // d > 23: the tile needs the opt-in shared-memory carve-out.  The attribute
                                  // is per device, so it is set on every such launch (no process-wide cache)
            CUSMC_CUDA(ctx, cudaFuncSetAttribute(density_aos_kernel<D, TRI>,
                                                 cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024));
        density_aos_kernel<D, TRI><<<(unsigned)grid, kThreads, smem, ctx->stream>>>(
            op, ep, x, N, d, vec_ok, out);
    }
    CUSMC_LAUNCHED(ctx);
    return CUSMC_OK;
}

// Zero-padded operator from a dense row-accessor.
template <int D, bool TRI, typename FM>
void fill_op(AffineOp<D, TRI> &op, int m, int d, FM Mkj, const double *shift, const double *off)
{
    std::memset(&op, 0, sizeof(op));
    for (int k = 0; k < m; ++k)
        for (int j = 0; j < (TRI ? k + 1 : d); ++j)
            op.M[TRI ? k * (k + 1) / 2 + j : k * D + j] = Mkj(k, j);
    for (int j = 0; j < d; ++j) op.shift[j] = shift ? shift[j] : 0.0;
    for (int k = 0; k < m; ++k) op.off[k] = off ? off[k] : 0.0;
}

}  // namespace

// Shared by filter.cu / api.cu: run "z = off - M (x - shift)" with a dense or lower
// triangular M (row accessor Mkj), rows m <= 32, cols d <= 32.
int cusmc_density_launch(cusmc_ctx *ctx, bool tri, int m, int d,
                         const std::vector<double> &M_rowmajor /* m x d */, const double *shift,
                         const double *off, const Epilogue &ep, const double *x, int layout,
                         int64_t N, int64_t ld, double *out)
{
    const int dm = m > d ? m : d;
    if (dm > CUSMC_MAX_DIM || d < 1 || m < 1)
        return cusmc_fail(ctx, CUSMC_ERR_UNSUPPORTED, "dimension %d not in 1..%d", dm, CUSMC_MAX_DIM);
    auto Mkj = [&](int k, int j) { return M_rowmajor[(size_t)k * d + j]; };
#define CUSMC_DENSITY_CASE(DD)                                                   \
    case DD:                                                                     \
        if (tri) {                                                               \
            AffineOp<DD, true> op;                                               \
            fill_op<DD, true>(op, m, d, Mkj, shift, off);                        \
            return launch_density<DD, true>(ctx, op, ep, x, layout, N, ld, d, out);  \
        } else {                                                                 \
            AffineOp<DD, false> op;                                              \
            fill_op<DD, false>(op, m, d, Mkj, shift, off);                       \
            return launch_density<DD, false>(ctx, op, ep, x, layout, N, ld, d, out); \
        }
    switch (cusmc_pad_dim(dm)) {
        CUSMC_DENSITY_CASE(2)
        CUSMC_DENSITY_CASE(4)
        CUSMC_DENSITY_CASE(8)
        CUSMC_DENSITY_CASE(16)
        CUSMC_DENSITY_CASE(32)
    }
#undef CUSMC_DENSITY_CASE
    return cusmc_fail(ctx, CUSMC_ERR_UNSUPPORTED, "unreachable");
}

// Builds W = L^-1 (row-major, lower) and the epilogue constants for MVN/MVT(mu, Sigma, nu).
int cusmc_build_whitening(cusmc_ctx *ctx, int kind, int want_log, int d, const double *sigma,
                          float nu, std::vector<double> &W_rowmajor, Epilogue &ep)
{
    if (kind != CUSMC_MVN && kind != CUSMC_MVT)
        return cusmc_fail(ctx, CUSMC_ERR_INVALID, "unknown distribution kind %d", kind);
    if (kind == CUSMC_MVT && !(nu > 0.0f))
        return cusmc_fail(ctx, CUSMC_ERR_INVALID, "mvt needs nu > 0 (got %g)", (double)nu);
    std::vector<double> L, W;
    const int bad = hostmath::cholesky_lower(sigma, d, L);
    if (bad)
        return cusmc_fail(ctx, CUSMC_ERR_NOT_SPD, "covariance is not positive definite (pivot %d)", bad - 1);
    hostmath::tri_inverse_lower(L, d, W);
    W_rowmajor.assign((size_t)d * d, 0.0);
    for (int k = 0; k < d; ++k)
        for (int j = 0; j <= k; ++j) W_rowmajor[(size_t)k * d + j] = W[(size_t)j * d + k];
    const double logdet = hostmath::logdet_from_cholesky(L, d);
    ep.kind = kind;
    ep.want_log = want_log;
    if (kind == CUSMC_MVN) {
        ep.lognorm = hostmath::mvn_lognorm(logdet, d);
        ep.half_nu_d = 0.0;
        ep.inv_nu = 0.0;
    } else {
        ep.lognorm = hostmath::mvt_lognorm(logdet, d, nu);
        ep.half_nu_d = hostmath::mvt_half_nu_plus_d(nu, d);
        ep.inv_nu = 1.0 / (double)nu;
    }
    ep.scale = std::exp(ep.lognorm);
    return CUSMC_OK;
}

struct cusmc_density_cache {
    int kind = -1, want_log = -1, d = 0;
    float nu = 0.f;
    std::vector<double> sigma, W;
    Epilogue ep{};
};

void cusmc_density_cache_free(cusmc_ctx *ctx)
{
    delete ctx->dcache;
    ctx->dcache = nullptr;
}

// cusmc_build_whitening through the context's one-entry cache (keyed on every input bit).
static int cached_whitening(cusmc_ctx *ctx, int kind, int want_log, int d, const double *sigma, float nu,
                            const std::vector<double> **W, Epilogue *ep)
{
    if (!ctx->dcache) ctx->dcache = new (std::nothrow) cusmc_density_cache();
    cusmc_density_cache *c = ctx->dcache;
    if (!c) return cusmc_fail(ctx, CUSMC_ERR_CUDA, "out of host memory");
    const size_t n = (size_t)d * d;
    const bool hit = c->kind == kind && c->want_log == want_log && c->d == d &&
                     std::memcmp(&c->nu, &nu, sizeof nu) == 0 && c->sigma.size() == n &&
                     std::memcmp(c->sigma.data(), sigma, n * sizeof(double)) == 0;
    if (!hit) {
        c->kind = -1;   // stays invalid if the factorisation fails
        CUSMC_CHECK(cusmc_build_whitening(ctx, kind, want_log, d, sigma, nu, c->W, c->ep));
        c->sigma.assign(sigma, sigma + n);
        c->kind = kind;
        c->want_log = want_log;
        c->d = d;
        c->nu = nu;
    }
    *W = &c->W;
    *ep = c->ep;
    return CUSMC_OK;
}

extern "C" int cusmc_logpdf_dev(cusmc_ctx *ctx, int kind, int want_log, const double *x_dev,
                                int layout, int64_t N, int64_t ld, int d, const double *mu,
                                const double *sigma, float nu, double *out_dev)
{
    CUSMC_ENTER(ctx);
    CUSMC_REQUIRE(ctx, N >= 0 && d >= 1, "N >= 0 and d >= 1 required");
    CUSMC_REQUIRE(ctx, sigma != nullptr, "sigma is NULL");
    CUSMC_REQUIRE(ctx, N == 0 || (x_dev && out_dev), "x/out is NULL");
    CUSMC_REQUIRE(ctx, layout == CUSMC_SOA || layout == CUSMC_AOS, "bad layout");
    CUSMC_REQUIRE(ctx, layout == CUSMC_AOS || ld >= N, "ld < N");
    if (d > CUSMC_MAX_DIM)
        return cusmc_fail(ctx, CUSMC_ERR_UNSUPPORTED, "d = %d > %d", d, CUSMC_MAX_DIM);
    const std::vector<double> *W = nullptr;
    Epilogue ep;
    CUSMC_CHECK(cached_whitening(ctx, kind, want_log, d, sigma, nu, &W, &ep));
    return cusmc_density_launch(ctx, true, d, d, *W, mu, nullptr, ep, x_dev, layout, N, ld, out_dev);
}

// Host-pointer entry point.  The call is PCIe-bound (8d bytes in, 8 out per point, a kernel that
// needs ~1 % of the copy time), so it is pipelined: the batch is cut into chunks that alternate
// between two streams, and a chunk's kernel and device->host copy run under the next chunk's
// host->device copy (the two directions use different copy engines).  With pinned host buffers
// the call takes the time of the host->device copy alone.
extern "C" int cusmc_logpdf(cusmc_ctx *ctx, int kind, int want_log, const double *x_host,
                            int layout, int64_t N, int64_t ld, int d, const double *mu,
                            const double *sigma, float nu, double *out_host)
{
    CUSMC_ENTER(ctx);
    CUSMC_REQUIRE(ctx, N >= 0 && d >= 1, "N >= 0 and d >= 1 required");
    CUSMC_REQUIRE(ctx, N == 0 || (x_host && out_host), "x/out is NULL");
    CUSMC_REQUIRE(ctx, layout == CUSMC_SOA || layout == CUSMC_AOS, "bad layout");
    if (N == 0) return CUSMC_OK;
    if (layout == CUSMC_AOS) ld = N;
    CUSMC_REQUIRE(ctx, ld >= N, "ld < N");
    constexpr int64_t kChunk = 1 << 17;                      // points per chunk (16 MiB at d = 16)
    const int64_t n_chunks = (N + kChunk - 1) / kChunk;
    const int64_t cpts = n_chunks > 1 ? kChunk : N;          // device columns per buffer
    void *xd = nullptr, *od = nullptr;
    CUSMC_CHECK(cusmc_scratch(ctx, 0, sizeof(double) * (size_t)cpts * d * 2, &xd));
    CUSMC_CHECK(cusmc_scratch(ctx, 1, sizeof(double) * (size_t)cpts * 2, &od));
    CUSMC_CHECK(cusmc_aux_stream(ctx));
    cudaStream_t main_stream = ctx->stream, lanes[2] = {ctx->stream, ctx->aux_stream};
    CUSMC_CUDA(ctx, cudaEventRecord(ctx->ev0, main_stream));
    if (n_chunks > 1) CUSMC_CUDA(ctx, cudaStreamWaitEvent(ctx->aux_stream, ctx->ev0, 0));
    int rc = CUSMC_OK;
    for (int64_t c = 0; c < n_chunks && rc == CUSMC_OK; ++c) {
        const int lane = (int)(c & 1);
        const int64_t i0 = c * kChunk, n = std::min<int64_t>(kChunk, N - i0);
        double *xb = (double *)xd + (size_t)lane * cpts * d, *ob = (double *)od + (size_t)lane * cpts;
        cudaStream_t st = lanes[lane];
        cudaError_t e;
        if (layout == CUSMC_AOS)
            e = cudaMemcpyAsync(xb, x_host + (size_t)i0 * d, sizeof(double) * (size_t)n * d, cudaMemcpyHostToDevice, st);
        else
            e = cudaMemcpy2DAsync(xb, sizeof(double) * (size_t)cpts, x_host + i0, sizeof(double) * (size_t)ld,
                                  sizeof(double) * (size_t)n, (size_t)d, cudaMemcpyHostToDevice, st);
        if (e != cudaSuccess) {
            rc = cusmc_fail(ctx, CUSMC_ERR_CUDA, "host->device copy failed: %s", cudaGetErrorString(e));
            break;
        }
        ctx->stream = st;                                      // launch this chunk on its lane
        rc = cusmc_logpdf_dev(ctx, kind, want_log, xb, layout, n, cpts, d, mu, sigma, nu, ob);
        ctx->stream = main_stream;
        if (rc != CUSMC_OK) break;
        e = cudaMemcpyAsync(out_host + i0, ob, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, st);
        if (e != cudaSuccess) rc = cusmc_fail(ctx, CUSMC_ERR_CUDA, "device->host copy failed: %s", cudaGetErrorString(e));
    }
    if (n_chunks > 1) {                                        // join the second lane
        cudaEventRecord(ctx->ev_aux, ctx->aux_stream);
        cudaStreamWaitEvent(main_stream, ctx->ev_aux, 0);
    }
    cudaEventRecord(ctx->ev1, main_stream);
    cudaError_t e = cudaStreamSynchronize(main_stream);
    if (rc != CUSMC_OK) return rc;
    if (e != cudaSuccess) return cusmc_fail(ctx, CUSMC_ERR_CUDA, "cusmc_logpdf: %s", cudaGetErrorString(e));
    float ms = 0.f;
    CUSMC_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    ctx->last_ms = ms;   // whole pipelined call: copies and kernels overlap
    return CUSMC_OK;
}

// resample.cu -- ancestor selection.
//
//  * metropolis_kernel: the reference's resampler (Sampler::metropolis_hastings,
//    src/samplers.cpp:7-36) -- B accept/reject steps per particle on the rule
//    u <= w[j] / w[k] (IEEE division, so decisions match the CPU bit for bit).
//  * weights_max / weights_sum / scan_resample: max-shifted normalisation, ESS and
//    systematic / multinomial resampling on the deterministic fixed-point weight image
//    (include/cusmc_detmath.h).  No counterpart in the reference.
//
// All of it is integer / byte / 8-byte-gather work bound by HBM or L2, not by math.
#include "common.cuh"
#include "resample.cuh"

#include "../../include/cusmc_detmath.h"
#include "../../include/cusmc_philox.h"

namespace {

constexpr int kThreads = 256;

// ------------------------------------------------------------------------------------------
// Metropolis ancestor resampler
// ------------------------------------------------------------------------------------------
template <bool PREDRAWN>
__global__ void __launch_bounds__(kThreads)
metropolis_kernel(uint32_t *__restrict__ a, const double *__restrict__ w,
                  const double *__restrict__ u, const uint32_t *__restrict__ j, uint64_t seed,
                  uint64_t step, int64_t N, int B, int is_log, int64_t i0, int64_t n_out)
{
    const int64_t t = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (t >= n_out) return;
    const int64_t i = i0 + t;              // global particle index (i0 = 0 on one GPU)
    uint32_t k = (uint32_t)i;
    double wk = __ldg(w + i);
    for (int n = 0; n < B; ++n) {
        double un;
        uint32_t jn;
        if (PREDRAWN) {
            un = __ldg(u + t * B + n);
            jn = __ldg(j + t * B + n);
        } else {
            const cusmc_u32x4 r = cusmc_rng(seed, CUSMC_STREAM_METROPOLIS, step, (uint64_t)i, (uint32_t)n);
            un = cusmc_u01(r.v[0], r.v[1]);
            jn = (uint32_t)cusmc_uint_below(r.v[2], r.v[3], (uint64_t)N);
        }
        const double wj = __ldg(w + jn);
        // linear: u <= w_j / w_k (0/0 = NaN rejects, x/0 = inf accepts, as on the CPU);
        // log   : u <= exp(lw_j - lw_k) with the reproducible exp.
        const double ratio = is_log ? cusmc_det_exp(wj - wk) : wj / wk;
        if (un <= ratio) {
            k = jn;
            wk = wj;
        }
    }
    a[t] = k;
}

// ------------------------------------------------------------------------------------------
// max and fixed-point sums
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void atomic_max_double(double *addr, double v)
{
    // slot must start at -inf.  Non-negative doubles order like signed ints, negative ones
    // like unsigned ints reversed.
    if (v != v) return;
    if (v >= 0.0)
        atomicMax(reinterpret_cast<long long *>(addr), __double_as_longlong(v));
    else
        atomicMin(reinterpret_cast<unsigned long long *>(addr),
                  (unsigned long long)__double_as_longlong(v));
}

__global__ void fill_double_kernel(double *p, double v, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

__global__ void __launch_bounds__(kThreads)
weights_max_kernel(const double *__restrict__ w, int64_t N, double *__restrict__ out)
{
    double m = -INFINITY;
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < N; i += (int64_t)gridDim.x * kThreads) {
        const double v = __ldg(w + i);
        if (v > m && v < INFINITY) m = v;   // NaN and +inf never become the reference weight
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    __shared__ double sm[kThreads / 32];
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x < 32) {
        m = threadIdx.x < kThreads / 32 ? sm[threadIdx.x] : -INFINITY;
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
        if (threadIdx.x == 0) atomic_max_double(out, m);
    }
}

__device__ __forceinline__ double unit_weight(double w, double wmax, int is_log)
{
    return is_log ? cusmc_unit_from_log(w, wmax) : cusmc_unit_from_linear(w, wmax);
}

// stats[0] += sum q, stats[1] += sum q2, stats[2] += #positive.  Integer atomics: the totals
// do not depend on the order the blocks arrive in.
__global__ void __launch_bounds__(kThreads)
weights_sum_kernel(const double *__restrict__ w, int is_log, const double *__restrict__ wmax_p,
                   int64_t N, int shift, unsigned long long *__restrict__ stats)
{
    const double wmax = *wmax_p;
    unsigned long long s1 = 0, s2 = 0, np = 0;
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < N; i += (int64_t)gridDim.x * kThreads) {
        const double wn = unit_weight(__ldg(w + i), wmax, is_log);
        const uint64_t q = cusmc_fixed_from_unit(wn, shift);
        s1 += q;
        s2 += cusmc_fixed_from_unit(wn * wn, shift);
        np += q > 0;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
        np += __shfl_xor_sync(0xffffffffu, np, o);
    }
    __shared__ unsigned long long sm[3][kThreads / 32];
    if ((threadIdx.x & 31) == 0) {
        sm[0][threadIdx.x >> 5] = s1;
        sm[1][threadIdx.x >> 5] = s2;
        sm[2][threadIdx.x >> 5] = np;
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        unsigned long long t = 0;
        for (int k = 0; k < kThreads / 32; ++k) t += sm[threadIdx.x][k];
        atomicAdd(stats + threadIdx.x, t);
    }
}

// ------------------------------------------------------------------------------------------
// Single-pass inclusive scan (decoupled look-back) of the fixed-point weights, with the
// systematic offspring scatter fused in.
// ------------------------------------------------------------------------------------------
constexpr int kScanItems = 8;                       // per thread: 4 rounds of one 128-bit load
constexpr int kScanTile = kThreads * kScanItems;    // 2048 weights per tile
constexpr uint64_t kFlagAgg = 1ull << 62, kFlagPrefix = 2ull << 62, kValMask = (1ull << 62) - 1;

// #{ i in [0, Ng) : i*T + r0 < C*Ng }  =  smallest k with k*T + r0 >= C*Ng, clamped to Ng.
// Floating-point estimate, then an exact 128-bit correction.
__device__ __forceinline__ uint64_t offspring_below(uint64_t C, uint64_t Ng, uint64_t T, uint64_t r0,
                                                    double ng_over_t, double r0_over_t)
{
    // p = (C*Ng - r0) / T; the answer is ceil(p) for p > 0.  est is within 2^-19 of p (C, Ng/T and
    // r0/T each carry one rounding and p <= 2^32), so whenever est sits safely inside an open unit
    // interval the answer is floor(est) + 1 and no 128-bit arithmetic is needed.  Only boundaries
    // within 1e-4 of an integer (2e-4 of all cases) take the exact path below.
    const double est = fma((double)C, ng_over_t, -r0_over_t);
    const double fl = floor(est);
    const double frac = est - fl;
    if (est > 1e-4 && frac > 1e-4 && frac < 1.0 - 1e-4) {
        const uint64_t kf = (uint64_t)fl + 1;
        return kf > Ng ? Ng : kf;
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

struct ScanArgs {
    const double *w;
    const double *wmax;
    const unsigned long long *total;       // global fixed-point mass (device)
    const unsigned long long *cdf_offset;  // mass on lower shards, or NULL
    unsigned long long *desc;              // tile descriptors, zeroed before the launch
    unsigned int *ticket;                  // zeroed before the launch
    unsigned long long *cdf_out;           // optional inclusive global CDF
    uint32_t *anc_out;                     // optional systematic ancestors for children
    int64_t N, N_global, j0, out_lo, out_n;
    double u0;
    int shift, is_log;
};

__global__ void __launch_bounds__(kThreads)
scan_resample_kernel(const ScanArgs p)
{
    __shared__ unsigned int s_tile;
    __shared__ unsigned long long s_warp[kThreads / 32];
    __shared__ unsigned long long s_prefix;
    if (threadIdx.x == 0) s_tile = atomicAdd(p.ticket, 1u);
    __syncthreads();
    const int64_t tile = s_tile;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const double wmax = *p.wmax;

    // each warp owns 256 consecutive weights; round r: lane loads weights 64 r + 2 lane, +1
    const int64_t wbase = tile * kScanTile + (int64_t)warp * (kScanItems * 32);
    uint64_t q[kScanItems];
#pragma unroll
    for (int r = 0; r < kScanItems / 2; ++r) {
        const int64_t idx = wbase + r * 64 + lane * 2;
        double w0 = 0.0, w1 = 0.0;
        bool have0 = idx < p.N, have1 = idx + 1 < p.N;
        if (have1 && (((uintptr_t)(p.w + idx)) & 15) == 0) {
            const double2 v = __ldg(reinterpret_cast<const double2 *>(p.w + idx));
            w0 = v.x;
            w1 = v.y;
        } else {
            if (have0) w0 = __ldg(p.w + idx);
            if (have1) w1 = __ldg(p.w + idx + 1);
        }
        q[2 * r] = have0 ? cusmc_fixed_from_unit(unit_weight(w0, wmax, p.is_log), p.shift) : 0;
        q[2 * r + 1] = have1 ? cusmc_fixed_from_unit(unit_weight(w1, wmax, p.is_log), p.shift) : 0;
    }

    // warp-level inclusive scan, round after round
    uint64_t excl[kScanItems / 2];   // exclusive prefix of each pair within the warp
    uint64_t running = 0;
#pragma unroll
    for (int r = 0; r < kScanItems / 2; ++r) {
        const uint64_t s = q[2 * r] + q[2 * r + 1];
        uint64_t inc = s;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint64_t n = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += n;
        }
        excl[r] = running + inc - s;
        running += __shfl_sync(0xffffffffu, inc, 31);
    }
    if (lane == 0) s_warp[warp] = running;
    __syncthreads();
    uint64_t warp_off = 0, aggregate = 0;
#pragma unroll
    for (int k = 0; k < kThreads / 32; ++k) {
        const uint64_t v = s_warp[k];
        if (k < warp) warp_off += v;
        aggregate += v;
    }

    // decoupled look-back by warp 0
    if (warp == 0) {
        uint64_t exclusive = 0;
        if (tile == 0) {
            exclusive = p.cdf_offset ? *p.cdf_offset : 0;
        } else {
            if (lane == 0) {
                __threadfence();
                atomicExch(p.desc + tile, kFlagAgg | aggregate);
            }
            int64_t look = tile - 1;
            while (true) {
                const int64_t mine = look - lane;
                unsigned long long d = kFlagPrefix;   // lanes before tile 0 contribute nothing
                if (mine >= 0) {
                    do {
                        d = *reinterpret_cast<volatile unsigned long long *>(p.desc + mine);
                    } while ((d >> 62) == 0);
                }
                const unsigned has_prefix = __ballot_sync(0xffffffffu, (d >> 62) == 2);
                const int first = has_prefix ? __ffs(has_prefix) - 1 : 32;
                uint64_t v = (lane <= first && mine >= 0) ? (d & kValMask) : 0;
                if (mine < 0) v = 0;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                exclusive += v;
                if (has_prefix) {
                    // tile 0's published prefix already contains cdf_offset; if the window ran
                    // past tile 0 the offset came with it.
                    break;
                }
                look -= 32;
            }
        }
        if (lane == 0) {
            __threadfence();
            atomicExch(p.desc + tile, kFlagPrefix | (exclusive + aggregate));
            s_prefix = exclusive;
        }
    }
    __syncthreads();
    const uint64_t base = s_prefix + warp_off;

    // outputs
    uint64_t T = 0, r0 = 0;
    double ng_over_t = 0.0, r0_over_t = 0.0;
    if (p.anc_out) {
        T = *p.total;
        if (T == 0) return;                           // degenerate: host reports it
        r0 = (uint64_t)(p.u0 * (double)T);
        if (r0 > T - 1) r0 = T - 1;
        ng_over_t = (double)p.N_global / (double)T;
        r0_over_t = (double)r0 / (double)T;
    }
    uint64_t carry_k = 0;   // k2 of lane 31 in the previous round
#pragma unroll
    for (int r = 0; r < kScanItems / 2; ++r) {
        const int64_t idx = wbase + r * 64 + lane * 2;
        const uint64_t c_before = base + excl[r];
        const uint64_t c0 = c_before + q[2 * r];
        const uint64_t c1 = c0 + q[2 * r + 1];
        if (p.cdf_out) {
            if (idx + 1 < p.N && (((uintptr_t)(p.cdf_out + idx)) & 15) == 0) {
                ulonglong2 v;
                v.x = c0;
                v.y = c1;
                *reinterpret_cast<ulonglong2 *>(p.cdf_out + idx) = v;
            } else {
                if (idx < p.N) p.cdf_out[idx] = c0;
                if (idx + 1 < p.N) p.cdf_out[idx + 1] = c1;
            }
        }
        if (p.anc_out) {
            // children of particle idx : [k0, k1),  of idx+1 : [k1, k2).  The count is a pure
            // function of the CDF value, so k0 is the left neighbour's k2: two evaluations per
            // pair, plus one per warp for the first pair of the warp's chunk.
            const uint64_t k1 = offspring_below(c0, (uint64_t)p.N_global, T, r0, ng_over_t, r0_over_t);
            const uint64_t k2 = offspring_below(c1, (uint64_t)p.N_global, T, r0, ng_over_t, r0_over_t);
            uint64_t k0 = __shfl_up_sync(0xffffffffu, k2, 1);
            if (lane == 0)
                k0 = (r == 0) ? offspring_below(c_before, (uint64_t)p.N_global, T, r0, ng_over_t, r0_over_t)
                              : carry_k;
            carry_k = __shfl_sync(0xffffffffu, k2, 31);
            const uint64_t lo_lim = (uint64_t)p.out_lo, hi_lim = (uint64_t)(p.out_lo + p.out_n);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                uint64_t a = h ? k1 : k0, b = h ? k2 : k1;
                if (a < lo_lim) a = lo_lim;
                if (b > hi_lim) b = hi_lim;
                const uint32_t parent = (uint32_t)(p.j0 + idx + h);
                uint64_t cnt = b > a ? b - a : 0;
                // small families: the owning thread writes them; large ones: the whole warp helps
                const bool big = cnt > 8;
                if (!big)
                    for (uint64_t i = a; i < b; ++i) p.anc_out[i - lo_lim] = parent;
                unsigned bigmask = __ballot_sync(0xffffffffu, big);
                while (bigmask) {
                    const int src = __ffs(bigmask) - 1;
                    bigmask &= bigmask - 1;
                    const uint64_t sa = __shfl_sync(0xffffffffu, a, src);
                    const uint64_t sb = __shfl_sync(0xffffffffu, b, src);
                    const uint32_t sp = __shfl_sync(0xffffffffu, parent, src);
                    for (uint64_t i = sa + lane; i < sb; i += 32) p.anc_out[i - lo_lim] = sp;
                }
            }
        }
    }
}

// Multinomial: a[i] = j0 + #{ j : cdf_j <= p_i },  p_i = min((uint64)(u_i * T), T - 1).
template <bool PREDRAWN>
__global__ void __launch_bounds__(kThreads)
multinomial_kernel(const unsigned long long *__restrict__ cdf, int64_t N,
                   const unsigned long long *__restrict__ total_p, const double *__restrict__ u,
                   uint64_t seed, uint64_t step, int64_t i0, int64_t n_out, int64_t j0,
                   uint32_t *__restrict__ a)
{
    const int64_t t = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (t >= n_out) return;
    const uint64_t T = *total_p;
    if (T == 0) return;
    double ui;
    if (PREDRAWN) {
        ui = __ldg(u + t);
    } else {
        const cusmc_u32x4 r = cusmc_rng(seed, CUSMC_STREAM_MULTINOMIAL, step, (uint64_t)(i0 + t), 0);
        ui = cusmc_u01(r.v[0], r.v[1]);
    }
    uint64_t pos = (uint64_t)(ui * (double)T);
    if (pos > T - 1) pos = T - 1;
    int64_t lo = 0, hi = N;
    while (lo < hi) {
        const int64_t mid = lo + ((hi - lo) >> 1);
        if (__ldg(cdf + mid) <= pos) lo = mid + 1; else hi = mid;
    }
    a[t] = (uint32_t)(j0 + lo);
}

int grid_for(cusmc_ctx *ctx, int64_t n, int per_block)
{
    int64_t g = (n + per_block - 1) / per_block;
    const int64_t cap = (int64_t)ctx->sm_count * 16;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

}  // namespace

// ---- internal launchers (device pointers, no synchronisation) --------------------------------
int cusmc_fill_double(cusmc_ctx *ctx, double *p, double v, int n)
{
    fill_double_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(p, v, n);
    CUSMC_LAUNCHED(ctx);
    return CUSMC_OK;
}

int cusmc_launch_metropolis(cusmc_ctx *ctx, uint32_t *a, const double *w, const double *u,
                            const uint32_t *j, uint64_t seed, uint64_t step, int64_t N, int B,
                            int is_log, int64_t i0, int64_t n_out)
{
    if (n_out == 0) return CUSMC_OK;
    const unsigned grid = (unsigned)((n_out + kThreads - 1) / kThreads);
    if (u)
        metropolis_kernel<true><<<grid, kThreads, 0, ctx->stream>>>(a, w, u, j, seed, step, N, B, is_log, i0, n_out);
    else
        metropolis_kernel<false><<<grid, kThreads, 0, ctx->stream>>>(a, w, u, j, seed, step, N, B, is_log, i0, n_out);
    CUSMC_LAUNCHED(ctx);
    return CUSMC_OK;
}

int cusmc_launch_weights_max(cusmc_ctx *ctx, const double *w, int64_t N, double *max_dev)
{
    if (N == 0) return CUSMC_OK;
    weights_max_kernel<<<grid_for(ctx, N, kThreads * 4), kThreads, 0, ctx->stream>>>(w, N, max_dev);
    CUSMC_LAUNCHED(ctx);
    return CUSMC_OK;
}

int cusmc_launch_weights_sum(cusmc_ctx *ctx, const double *w, int is_log, const double *max_dev,
                             int64_t N, int shift, uint64_t *stats_dev)
{
    if (N == 0) return CUSMC_OK;
    weights_sum_kernel<<<grid_for(ctx, N, kThreads * 4), kThreads, 0, ctx->stream>>>(
        w, is_log, max_dev, N, shift, (unsigned long long *)stats_dev);
    CUSMC_LAUNCHED(ctx);
    return CUSMC_OK;
}

size_t cusmc_scan_state_bytes(int64_t N)
{
    const int64_t tiles = (N + kScanTile - 1) / kScanTile;
    return sizeof(unsigned long long) * (size_t)(tiles + 2);
}

// state: [ticket (8 bytes)] [descriptors ...]; must be zero when the kernel starts.
int cusmc_launch_scan(cusmc_ctx *ctx, const double *w, int is_log, const double *max_dev, int64_t N,
                      int64_t N_global, int shift, const uint64_t *total_dev,
                      const uint64_t *cdf_offset_dev, void *state_dev, bool zero_state,
                      uint64_t *cdf_out, uint32_t *anc_out, int64_t j0, int64_t out_lo,
                      int64_t out_n, double u0)
{
    if (N == 0) return CUSMC_OK;
    if (zero_state)
        CUSMC_CUDA(ctx, cudaMemsetAsync(state_dev, 0, cusmc_scan_state_bytes(N), ctx->stream));
    ScanArgs p;
    p.w = w;
    p.wmax = max_dev;
    p.total = (const unsigned long long *)total_dev;
    p.cdf_offset = (const unsigned long long *)cdf_offset_dev;
    p.ticket = (unsigned int *)state_dev;
    p.desc = (unsigned long long *)state_dev + 1;
    p.cdf_out = (unsigned long long *)cdf_out;
    p.anc_out = anc_out;
    p.N = N;
    p.N_global = N_global;
    p.j0 = j0;
    p.out_lo = out_lo;
    p.out_n = out_n;
    p.u0 = u0;
    p.shift = shift;
    p.is_log = is_log;
    const unsigned tiles = (unsigned)((N + kScanTile - 1) / kScanTile);
    scan_resample_kernel<<<tiles, kThreads, 0, ctx->stream>>>(p);
    CUSMC_LAUNCHED(ctx);
    return CUSMC_OK;
}

int cusmc_launch_multinomial(cusmc_ctx *ctx, const uint64_t *cdf, int64_t N, const uint64_t *total_dev,
                             const double *u, uint64_t seed, uint64_t step, int64_t i0,
                             int64_t n_out, int64_t j0, uint32_t *a)
{
    if (n_out == 0) return CUSMC_OK;
    const unsigned grid = (unsigned)((n_out + kThreads - 1) / kThreads);
    if (u)
        multinomial_kernel<true><<<grid, kThreads, 0, ctx->stream>>>(
            (const unsigned long long *)cdf, N, (const unsigned long long *)total_dev, u, seed, step, i0, n_out, j0, a);
    else
        multinomial_kernel<false><<<grid, kThreads, 0, ctx->stream>>>(
            (const unsigned long long *)cdf, N, (const unsigned long long *)total_dev, u, seed, step, i0, n_out, j0, a);
    CUSMC_LAUNCHED(ctx);
    return CUSMC_OK;
}

// ---- extern "C": device-pointer API -----------------------------------------------------------
extern "C" int cusmc_metropolis_hastings_dev(cusmc_ctx *ctx, uint32_t *a_dev, const double *w_dev,
                                             const double *u_dev, const uint32_t *j_dev,
                                             uint64_t seed, uint64_t step, int64_t N, int B, int is_log)
{
    if (!ctx) return CUSMC_ERR_INVALID;
    CUSMC_REQUIRE(ctx, N >= 0 && B >= 0, "N, B must be non-negative");
    CUSMC_REQUIRE(ctx, N == 0 || (a_dev && w_dev), "a/w is NULL");
    CUSMC_REQUIRE(ctx, (u_dev == nullptr) == (j_dev == nullptr), "u and j must both be given or both NULL");
    CUSMC_REQUIRE(ctx, N <= 0xFFFFFFFFll, "N exceeds the 32-bit ancestor range");
    return cusmc_launch_metropolis(ctx, a_dev, w_dev, u_dev, j_dev, seed, step, N, B, is_log, 0, N);
}

extern "C" int cusmc_weights_max_dev(cusmc_ctx *ctx, const double *w_dev, int64_t N, double *max_dev)
{
    if (!ctx) return CUSMC_ERR_INVALID;
    CUSMC_REQUIRE(ctx, max_dev && (N == 0 || w_dev), "NULL pointer");
    CUSMC_CHECK(cusmc_fill_double(ctx, max_dev, -INFINITY, 1));
    return cusmc_launch_weights_max(ctx, w_dev, N, max_dev);
}

extern "C" int cusmc_weights_sum_dev(cusmc_ctx *ctx, const double *w_dev, int is_log,
                                     const double *max_dev, int64_t N, int64_t N_global,
                                     uint64_t *stats_dev)
{
    if (!ctx) return CUSMC_ERR_INVALID;
    CUSMC_REQUIRE(ctx, stats_dev && max_dev && (N == 0 || w_dev), "NULL pointer");
    CUSMC_REQUIRE(ctx, N_global >= N && N_global >= 1, "N_global < N");
    CUSMC_CUDA(ctx, cudaMemsetAsync(stats_dev, 0, 4 * sizeof(uint64_t), ctx->stream));
    return cusmc_launch_weights_sum(ctx, w_dev, is_log, max_dev, N, cusmc_fixed_shift(N_global), stats_dev);
}

extern "C" int cusmc_weights_scan_dev(cusmc_ctx *ctx, const double *w_dev, int is_log,
                                      const double *max_dev, int64_t N, int64_t N_global,
                                      const uint64_t *cdf_offset_dev, uint64_t *cdf_dev)
{
    if (!ctx) return CUSMC_ERR_INVALID;
    CUSMC_REQUIRE(ctx, max_dev && (N == 0 || (w_dev && cdf_dev)), "NULL pointer");
    CUSMC_REQUIRE(ctx, N_global >= N && N_global >= 1, "N_global < N");
    void *state = nullptr;
    CUSMC_CHECK(cusmc_scratch(ctx, 7, cusmc_scan_state_bytes(N), &state));
    return cusmc_launch_scan(ctx, w_dev, is_log, max_dev, N, N_global, cusmc_fixed_shift(N_global),
                             nullptr, cdf_offset_dev, state, true, cdf_dev, nullptr, 0, 0, 0, 0.0);
}

extern "C" int cusmc_resample_systematic_dev(cusmc_ctx *ctx, const double *w_dev, int is_log,
                                             const double *max_dev, int64_t N_local, int64_t N_global,
                                             const uint64_t *total_dev, const uint64_t *cdf_offset_dev,
                                             int64_t j0, int64_t out_lo, int64_t out_n, double u0,
                                             uint32_t *a_dev)
{
    if (!ctx) return CUSMC_ERR_INVALID;
    CUSMC_REQUIRE(ctx, max_dev && total_dev && (N_local == 0 || w_dev) && (out_n == 0 || a_dev), "NULL pointer");
    CUSMC_REQUIRE(ctx, N_global >= N_local && N_global >= 1, "N_global < N_local");
    CUSMC_REQUIRE(ctx, N_global <= 0xFFFFFFFFll, "N exceeds the 32-bit ancestor range");
    CUSMC_REQUIRE(ctx, u0 >= 0.0 && u0 < 1.0, "u0 must lie in [0, 1)");
    void *state = nullptr;
    CUSMC_CHECK(cusmc_scratch(ctx, 7, cusmc_scan_state_bytes(N_local), &state));
    return cusmc_launch_scan(ctx, w_dev, is_log, max_dev, N_local, N_global, cusmc_fixed_shift(N_global),
                             total_dev, cdf_offset_dev, state, true, nullptr, a_dev, j0, out_lo, out_n, u0);
}

extern "C" int cusmc_resample_multinomial_dev(cusmc_ctx *ctx, const uint64_t *cdf_dev, int64_t N,
                                              const uint64_t *total_dev, const double *u_dev,
                                              uint64_t seed, uint64_t step, int64_t i0, int64_t n_out,
                                              int64_t j0, uint32_t *a_dev)
{
    if (!ctx) return CUSMC_ERR_INVALID;
    CUSMC_REQUIRE(ctx, total_dev && (N == 0 || cdf_dev) && (n_out == 0 || a_dev), "NULL pointer");
    return cusmc_launch_multinomial(ctx, cdf_dev, N, total_dev, u_dev, seed, step, i0, n_out, j0, a_dev);
}

// ---- extern "C": host-pointer conveniences -------------------------------------------------------
namespace {

struct HostWeights {
    double *w_dev = nullptr;
    double *max_dev = nullptr;
    uint64_t *stats_dev = nullptr;
};

// Uploads N weights, leaves max and {sum q, sum q2, n_pos} on the device and in `host` (8 doubles).
int upload_and_reduce(cusmc_ctx *ctx, const double *w, int64_t N, int is_log, HostWeights &hw,
                      double *max_host, uint64_t stats_host[4])
{
    void *wd = nullptr, *small = nullptr, *pin = nullptr;
    CUSMC_CHECK(cusmc_scratch(ctx, 0, sizeof(double) * (size_t)N, &wd));
    CUSMC_CHECK(cusmc_scratch(ctx, 6, 64, &small));
    CUSMC_CHECK(cusmc_pinned(ctx, 64, &pin));
    hw.w_dev = (double *)wd;
    hw.max_dev = (double *)small;
    hw.stats_dev = (uint64_t *)small + 1;
    CUSMC_CUDA(ctx, cudaMemcpyAsync(wd, w, sizeof(double) * (size_t)N, cudaMemcpyHostToDevice, ctx->stream));
    CUSMC_CHECK(cusmc_weights_max_dev(ctx, hw.w_dev, N, hw.max_dev));
    CUSMC_CHECK(cusmc_weights_sum_dev(ctx, hw.w_dev, is_log, hw.max_dev, N, N, hw.stats_dev));
    CUSMC_CUDA(ctx, cudaMemcpyAsync(pin, small, 40, cudaMemcpyDeviceToHost, ctx->stream));
    CUSMC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    std::memcpy(max_host, pin, 8);
    std::memcpy(stats_host, (char *)pin + 8, 32);
    return CUSMC_OK;
}

}  // namespace

extern "C" int cusmc_metropolis_hastings(cusmc_ctx *ctx, uint32_t *a, const double *w, const double *u,
                                         const uint32_t *j, uint64_t seed, uint64_t step, int64_t N, int B)
{
    if (!ctx) return CUSMC_ERR_INVALID;
    CUSMC_REQUIRE(ctx, N >= 0 && B >= 0, "N, B must be non-negative");
    CUSMC_REQUIRE(ctx, N == 0 || (a && w), "a/w is NULL");
    CUSMC_REQUIRE(ctx, (u == nullptr) == (j == nullptr), "u and j must both be given or both NULL");
    if (N == 0) return CUSMC_OK;
    void *wd = nullptr, *ad = nullptr, *ud = nullptr, *jd = nullptr;
    CUSMC_CHECK(cusmc_scratch(ctx, 0, sizeof(double) * (size_t)N, &wd));
    CUSMC_CHECK(cusmc_scratch(ctx, 1, sizeof(uint32_t) * (size_t)N, &ad));
    CUSMC_CUDA(ctx, cudaMemcpyAsync(wd, w, sizeof(double) * (size_t)N, cudaMemcpyHostToDevice, ctx->stream));
    if (u) {
        CUSMC_CHECK(cusmc_scratch(ctx, 2, sizeof(double) * (size_t)N * B, &ud));
        CUSMC_CHECK(cusmc_scratch(ctx, 3, sizeof(uint32_t) * (size_t)N * B, &jd));
        CUSMC_CUDA(ctx, cudaMemcpyAsync(ud, u, sizeof(double) * (size_t)N * B, cudaMemcpyHostToDevice, ctx->stream));
        CUSMC_CUDA(ctx, cudaMemcpyAsync(jd, j, sizeof(uint32_t) * (size_t)N * B, cudaMemcpyHostToDevice, ctx->stream));
    }
    CUSMC_CUDA(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
    CUSMC_CHECK(cusmc_metropolis_hastings_dev(ctx, (uint32_t *)ad, (const double *)wd, (const double *)ud,
                                              (const uint32_t *)jd, seed, step, N, B, 0));
    CUSMC_CUDA(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
    CUSMC_CUDA(ctx, cudaMemcpyAsync(a, ad, sizeof(uint32_t) * (size_t)N, cudaMemcpyDeviceToHost, ctx->stream));
    CUSMC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    float ms = 0.f;
    CUSMC_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    ctx->last_ms = ms;
    return CUSMC_OK;
}

extern "C" int cusmc_resample_systematic(cusmc_ctx *ctx, const double *w, int64_t N, double u0, uint32_t *a)
{
    if (!ctx) return CUSMC_ERR_INVALID;
    CUSMC_REQUIRE(ctx, N >= 0 && (N == 0 || (w && a)), "bad arguments");
    CUSMC_REQUIRE(ctx, u0 >= 0.0 && u0 < 1.0, "u0 must lie in [0, 1)");
    if (N == 0) return CUSMC_OK;
    HostWeights hw;
    double mx;
    uint64_t st[4];
    CUSMC_CHECK(upload_and_reduce(ctx, w, N, 0, hw, &mx, st));
    if (!(mx > 0.0) || st[0] == 0) {
        for (int64_t i = 0; i < N; ++i) a[i] = (uint32_t)i;   // nothing to resample from
        return cusmc_fail(ctx, CUSMC_ERR_DEGENERATE, "all weights are zero or non-finite");
    }
    void *ad = nullptr;
    CUSMC_CHECK(cusmc_scratch(ctx, 1, sizeof(uint32_t) * (size_t)N, &ad));
    CUSMC_CHECK(cusmc_resample_systematic_dev(ctx, hw.w_dev, 0, hw.max_dev, N, N, hw.stats_dev, nullptr, 0, 0,
                                              N, u0, (uint32_t *)ad));
    CUSMC_CUDA(ctx, cudaMemcpyAsync(a, ad, sizeof(uint32_t) * (size_t)N, cudaMemcpyDeviceToHost, ctx->stream));
    CUSMC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return CUSMC_OK;
}

extern "C" int cusmc_resample_multinomial(cusmc_ctx *ctx, const double *w, int64_t N, const double *u,
                                          uint32_t *a)
{
    if (!ctx) return CUSMC_ERR_INVALID;
    CUSMC_REQUIRE(ctx, N >= 0 && (N == 0 || (w && a && u)), "bad arguments");
    if (N == 0) return CUSMC_OK;
    HostWeights hw;
    double mx;
    uint64_t st[4];
    CUSMC_CHECK(upload_and_reduce(ctx, w, N, 0, hw, &mx, st));
    if (!(mx > 0.0) || st[0] == 0) {
        for (int64_t i = 0; i < N; ++i) a[i] = (uint32_t)i;
        return cusmc_fail(ctx, CUSMC_ERR_DEGENERATE, "all weights are zero or non-finite");
    }
    void *ad = nullptr, *cd = nullptr, *ud = nullptr;
    CUSMC_CHECK(cusmc_scratch(ctx, 1, sizeof(uint32_t) * (size_t)N, &ad));
    CUSMC_CHECK(cusmc_scratch(ctx, 2, sizeof(uint64_t) * (size_t)N, &cd));
    CUSMC_CHECK(cusmc_scratch(ctx, 3, sizeof(double) * (size_t)N, &ud));
    CUSMC_CUDA(ctx, cudaMemcpyAsync(ud, u, sizeof(double) * (size_t)N, cudaMemcpyHostToDevice, ctx->stream));
    CUSMC_CHECK(cusmc_weights_scan_dev(ctx, hw.w_dev, 0, hw.max_dev, N, N, nullptr, (uint64_t *)cd));
    CUSMC_CHECK(cusmc_resample_multinomial_dev(ctx, (const uint64_t *)cd, N, hw.stats_dev, (const double *)ud,
                                               0, 0, 0, N, 0, (uint32_t *)ad));
    CUSMC_CUDA(ctx, cudaMemcpyAsync(a, ad, sizeof(uint32_t) * (size_t)N, cudaMemcpyDeviceToHost, ctx->stream));
    CUSMC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return CUSMC_OK;
}

extern "C" int cusmc_normalize_ess(cusmc_ctx *ctx, const double *lw, int64_t N, double *lse, double *ess)
{
    if (!ctx) return CUSMC_ERR_INVALID;
    CUSMC_REQUIRE(ctx, N >= 1 && lw, "bad arguments");
    HostWeights hw;
    double mx;
    uint64_t st[4];
    CUSMC_CHECK(upload_and_reduce(ctx, lw, N, 1, hw, &mx, st));
    if (st[0] == 0) return cusmc_fail(ctx, CUSMC_ERR_DEGENERATE, "all log-weights are -inf or NaN");
    const double scale = std::ldexp(1.0, cusmc_fixed_shift(N));
    if (lse) *lse = mx + std::log((double)st[0] / scale);
    if (ess) *ess = ((double)st[0] * (double)st[0]) / ((double)st[1] * scale);
    return CUSMC_OK;
}

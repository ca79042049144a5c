// common.cuh -- context object, error plumbing and launch helpers shared by every
// translation unit of libcusmc_b200.so.  sm_100a only; no CPU fallback anywhere.
#pragma once

#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>

#include "../../include/cusmc_b200.h"

#define CUSMC_NUM_SCRATCH 8

// One peer-mapped pointer per rank of a sharded run (rank r owns global slots
// r * per_rank .. ), see cusmc_ipc_open.
struct CusmcPeers {
    void *ptr[CUSMC_MAX_PEERS];   // host copy (lifetime management)
    void **table_dev;             // the same world pointers in DEVICE memory: kernels index this with a
                                  // load -- indexing a kernel-parameter array by a runtime value would
                                  // make every thread copy the whole parameter block to local memory
    int64_t per_rank;
    int world;
};

// n / d for any 32-bit n by multiply-high and shifts (Granlund-Montgomery round-up form): slot ->
// owning rank on every peer access, without a hardware-emulated division per particle.
struct FastDiv {
    uint32_t d, M, s1, s2;
};
inline FastDiv make_fast_div(uint32_t d)
{
    uint32_t l = 0;
    while (((uint64_t)1 << l) < d) ++l;
    FastDiv f;
    f.d = d;
    f.M = (uint32_t)((((uint64_t)1 << 32) * (((uint64_t)1 << l) - d)) / d + 1);
    f.s1 = l < 1 ? l : 1;
    f.s2 = l > 0 ? l - 1 : 0;
    return f;
}
#ifdef __CUDACC__
__device__ __forceinline__ uint32_t fast_div(uint32_t n, const FastDiv &f)
{
    const uint32_t t = __umulhi(f.M, n);
    return (t + ((n - t) >> f.s1)) >> f.s2;
}
#endif

struct cusmc_density_cache;   // density.cu: last factored (kind, mu, Sigma, nu)

struct cusmc_ctx {
    int device = 0;
    int sm_count = 148;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    // second lane of the pipelined host-pointer calls (created on first use)
    cudaStream_t aux_stream = nullptr;
    cudaEvent_t ev_aux = nullptr;
    uint64_t launches = 0;
    double last_ms = 0.0;
    std::string err;
    // grow-only device scratch used by the host-pointer entry points
    void *scratch[CUSMC_NUM_SCRATCH] = {};
    size_t scratch_cap[CUSMC_NUM_SCRATCH] = {};
    // pinned staging for small device->host reads (status words, stats)
    void *pinned = nullptr;
    size_t pinned_cap = 0;
    // set-up algebra of the last density call (Cholesky, L^-1, log norm): a caller that evaluates
    // batch after batch under one distribution -- the reference builds its distribution object once
    // and calls pdf() many times, src/mcmc.cpp:53-58 -- pays for it once, not per 25 us kernel
    cusmc_density_cache *dcache = nullptr;
    // cusmc_ctx_set_chain_noise: 1 = the MH chain kernels draw their proposal normals with the throughput
    // generator (Philox4x32-7, special-function-unit Box-Muller) instead of the host-reproducible one
    int chain_fast_noise = 0;
};

void cusmc_density_cache_free(cusmc_ctx *ctx);

inline int cusmc_fail(cusmc_ctx *ctx, int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (ctx) ctx->err = buf;
    return code;
}

// The reference's CUDA_CALL (inst/include/support.cuh:9-20) prints and throws; here a
// failed runtime call becomes a status code plus a message on the context.
#define CUSMC_CUDA(ctx, call)                                                          \
    do {                                                                                \
        cudaError_t e__ = (call);                                                       \
        if (e__ != cudaSuccess)                                                         \
            return cusmc_fail((ctx), CUSMC_ERR_CUDA, "%s:%d: %s failed: %s", __FILE__,  \
                              __LINE__, #call, cudaGetErrorString(e__));                \
    } while (0)

// Programmatic dependent launch for the short kernels that follow each other on a stream (step /
// resample pairs of the reference-mode filter, back-to-back density calls): the grid may become resident
// while its predecessor drains, which hides its launch latency and block dispatch.  A kernel launched
// this way executes cusmc_pdl_enter() BEFORE its first access to memory: it lets ITS successor do the same
// and then waits until the predecessor has completed and its writes are visible (a no-op after an
// ordinary launch or a copy).
#ifdef __CUDACC__
__device__ __forceinline__ void cusmc_pdl_enter()
{
    asm volatile("griddepcontrol.launch_dependents;");
    asm volatile("griddepcontrol.wait;" ::: "memory");
}
template <typename... P, typename... A>
inline cudaError_t cusmc_launch_pdl(void (*kernel)(P...), unsigned grid, unsigned block, size_t smem, cudaStream_t st,
                                    A &&...args)
{
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cudaLaunchConfig_t lc{};
    lc.gridDim = dim3(grid);
    lc.blockDim = dim3(block);
    lc.dynamicSmemBytes = smem;
    lc.stream = st;
    lc.attrs = attr;
    lc.numAttrs = 1;
    return cudaLaunchKernelEx(&lc, kernel, P(args)...);
}
#endif

#define CUSMC_CHECK(expr)                 \
    do {                                  \
        int rc__ = (expr);                \
        if (rc__ != CUSMC_OK) return rc__; \
    } while (0)

#define CUSMC_REQUIRE(ctx, cond, msg)                                             \
    do {                                                                          \
        if (!(cond)) return cusmc_fail((ctx), CUSMC_ERR_INVALID, "%s: %s", __func__, (msg)); \
    } while (0)

// First statement of every extern "C" entry point that launches, allocates or copies: contexts on
// different devices may be used from one thread, and per-device state must follow the context.
#define CUSMC_ENTER(ctx)                                   \
    do {                                                   \
        if (!(ctx)) return CUSMC_ERR_INVALID;              \
        CUSMC_CUDA((ctx), cudaSetDevice((ctx)->device));   \
    } while (0)

// Checked after every launch; counts the launch for bench.py's gpu_launches.
#define CUSMC_LAUNCHED(ctx)                                                          \
    do {                                                                             \
        cudaError_t e__ = cudaGetLastError();                                        \
        if (e__ != cudaSuccess)                                                      \
            return cusmc_fail((ctx), CUSMC_ERR_CUDA, "%s:%d: kernel launch failed: %s", \
                              __FILE__, __LINE__, cudaGetErrorString(e__));          \
        (ctx)->launches++;                                                           \
    } while (0)

int cusmc_scratch(cusmc_ctx *ctx, int slot, size_t bytes, void **out);
int cusmc_pinned(cusmc_ctx *ctx, size_t bytes, void **out);
int cusmc_aux_stream(cusmc_ctx *ctx);
// Large device -> pageable-host copy: pinned double-buffered staging, the host side of every chunk
// copied (and first-touched) by several threads while the next chunk is in flight.  Synchronous.
int cusmc_d2h_staged(cusmc_ctx *ctx, void *dst_host, const void *src_dev, size_t bytes);

// ---- device helpers ----------------------------------------------------------------
// Streaming (read-once / write-once) accesses: keep them out of L1 and mark evict-first.
__device__ __forceinline__ double ld_stream(const double *p) { return __ldcs(p); }
__device__ __forceinline__ double2 ld_stream2(const double *p)
{
    return __ldcs(reinterpret_cast<const double2 *>(p));
}
__device__ __forceinline__ void st_stream(double *p, double v) { __stcs(p, v); }
__device__ __forceinline__ void st_stream2(double *p, double2 v)
{
    __stcs(reinterpret_cast<double2 *>(p), v);
}

// Order-preserving map double <-> uint64 so atomicMax on integers gives the fp64 max
// (deterministic: max is associative and commutative).
__device__ __forceinline__ unsigned long long ordered_from_double(double v)
{
    unsigned long long b = (unsigned long long)__double_as_longlong(v);
    return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double double_from_ordered(unsigned long long o)
{
    unsigned long long b = (o & 0x8000000000000000ull) ? (o & 0x7FFFFFFFFFFFFFFFull) : ~o;
    return __longlong_as_double((long long)b);
}

// atomicMax on a double slot that starts at -inf: non-negative doubles order like signed ints,
// negative ones like unsigned ints reversed.  NaN is ignored.
__device__ __forceinline__ void atomic_max_double(double *addr, double v)
{
    if (v != v) return;
    // -0.0 compares >= 0 but its bit pattern is INT64_MIN as a signed integer, which could never
    // replace the initial -inf: it goes through the negative branch (where it beats every negative)
    if (v > 0.0 || __double_as_longlong(v) == 0)
        atomicMax(reinterpret_cast<long long *>(addr), __double_as_longlong(v));
    else
        atomicMin(reinterpret_cast<unsigned long long *>(addr), (unsigned long long)__double_as_longlong(v));
}

// Warp-wide maximum of finite-or--inf doubles (callers map NaN / +inf to -inf first): two
// REDUX.MAX on the halves of the order-preserving integer image instead of five rounds of
// 64-bit shuffles + compares.  Every lane gets the result.
__device__ __forceinline__ double warp_max_double(double v)
{
    const unsigned long long o = ordered_from_double(v);
    const unsigned hi = (unsigned)(o >> 32), lo = (unsigned)o;
    const unsigned mhi = __reduce_max_sync(0xffffffffu, hi);
    const unsigned mlo = __reduce_max_sync(0xffffffffu, hi == mhi ? lo : 0u);
    return double_from_ordered(((unsigned long long)mhi << 32) | mlo);
}

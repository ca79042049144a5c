// context.cu -- context lifetime, stream binding, grow-only scratch.
//
// Ownership model: the context owns device scratch, stream and
// events; the caller owns host buffers.  The reference instead cudaMalloc/cudaFree-s
// 8-10 buffers and calls cudaDeviceReset() on every time step
// (src/mvn_dist.cu.cpp:231-240,305-314,788) -- none of that survives here.
#include "common.cuh"

#include <algorithm>
#include <new>
#include <thread>
#include <vector>

extern "C" int cusmc_version(void) { return CUSMC_VERSION; }

extern "C" int cusmc_ctx_create(cusmc_ctx **out, int device)
{
    if (!out) return CUSMC_ERR_INVALID;
    *out = nullptr;
    int n_dev = 0;
    cudaError_t e = cudaGetDeviceCount(&n_dev);
    if (e != cudaSuccess || n_dev == 0) return CUSMC_ERR_CUDA;   // no CPU fallback, by design
    if (device < 0 || device >= n_dev) return CUSMC_ERR_INVALID;
    cusmc_ctx *ctx = new (std::nothrow) cusmc_ctx();
    if (!ctx) return CUSMC_ERR_CUDA;
    ctx->device = device;
    if (cudaSetDevice(device) != cudaSuccess ||
        cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device) != cudaSuccess ||
        cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreate(&ctx->ev0) != cudaSuccess || cudaEventCreate(&ctx->ev1) != cudaSuccess) {
        delete ctx;
        return CUSMC_ERR_CUDA;
    }
    ctx->stream = ctx->own_stream;
    // single-GPU filters take their device buffers from the device's stream-ordered pool: keep what they
    // free, so that a filter per call (cusmc_run) costs microseconds of allocation, not ~15 ms of
    // cudaMalloc / cudaFree
    cudaMemPool_t pool = nullptr;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
        uint64_t keep = UINT64_MAX;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    *out = ctx;
    return CUSMC_OK;
}

extern "C" int cusmc_ctx_destroy(cusmc_ctx *ctx)
{
    if (!ctx) return CUSMC_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (int s = 0; s < CUSMC_NUM_SCRATCH; ++s)
        if (ctx->scratch[s]) cudaFree(ctx->scratch[s]);
    if (ctx->pinned) cudaFreeHost(ctx->pinned);
    cusmc_density_cache_free(ctx);
    if (ctx->ev_aux) cudaEventDestroy(ctx->ev_aux);
    if (ctx->aux_stream) cudaStreamDestroy(ctx->aux_stream);
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    delete ctx;
    return CUSMC_OK;
}

extern "C" const char *cusmc_last_error(const cusmc_ctx *ctx)
{
    return ctx ? ctx->err.c_str() : "cusmc: NULL context (no CUDA device, or creation failed)";
}

extern "C" int cusmc_ctx_set_stream(cusmc_ctx *ctx, void *cuda_stream)
{
    if (!ctx) return CUSMC_ERR_INVALID;
    // NULL selects the context's own stream; the legacy default stream is addressed by its
    // explicit handle cudaStreamLegacy (0x1), which is what bindings pass for "stream 0".
    ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
    return CUSMC_OK;
}

extern "C" int cusmc_ctx_set_chain_noise(cusmc_ctx *ctx, int reproducible)
{
    if (!ctx) return CUSMC_ERR_INVALID;
    ctx->chain_fast_noise = reproducible ? 0 : 1;
    return CUSMC_OK;
}

extern "C" int cusmc_ctx_synchronize(cusmc_ctx *ctx)
{
    if (!ctx) return CUSMC_ERR_INVALID;
    CUSMC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return CUSMC_OK;
}

extern "C" uint64_t cusmc_ctx_launch_count(const cusmc_ctx *ctx) { return ctx ? ctx->launches : 0; }
extern "C" double cusmc_ctx_last_kernel_ms(const cusmc_ctx *ctx) { return ctx ? ctx->last_ms : 0.0; }

int cusmc_scratch(cusmc_ctx *ctx, int slot, size_t bytes, void **out)
{
    if (slot < 0 || slot >= CUSMC_NUM_SCRATCH) return cusmc_fail(ctx, CUSMC_ERR_INVALID, "bad scratch slot");
    if (bytes > ctx->scratch_cap[slot]) {
        CUSMC_CUDA(ctx, cudaSetDevice(ctx->device));
        if (ctx->scratch[slot]) {
            CUSMC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            CUSMC_CUDA(ctx, cudaFree(ctx->scratch[slot]));
            ctx->scratch[slot] = nullptr;
            ctx->scratch_cap[slot] = 0;
        }
        const size_t cap = (bytes + ((size_t)1 << 20) - 1) & ~(((size_t)1 << 20) - 1);
        CUSMC_CUDA(ctx, cudaMalloc(&ctx->scratch[slot], cap));
        ctx->scratch_cap[slot] = cap;
    }
    *out = ctx->scratch[slot];
    return CUSMC_OK;
}

int cusmc_aux_stream(cusmc_ctx *ctx)
{
    if (ctx->aux_stream) return CUSMC_OK;
    CUSMC_CUDA(ctx, cudaSetDevice(ctx->device));
    CUSMC_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->aux_stream, cudaStreamNonBlocking));
    CUSMC_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_aux, cudaEventDisableTiming));
    return CUSMC_OK;
}

int cusmc_pinned(cusmc_ctx *ctx, size_t bytes, void **out)
{
    if (bytes > ctx->pinned_cap) {
        if (ctx->pinned) {
            CUSMC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            CUSMC_CUDA(ctx, cudaFreeHost(ctx->pinned));
            ctx->pinned = nullptr;
            ctx->pinned_cap = 0;
        }
        const size_t cap = bytes < 4096 ? 4096 : bytes;
        CUSMC_CUDA(ctx, cudaMallocHost(&ctx->pinned, cap));
        ctx->pinned_cap = cap;
    }
    *out = ctx->pinned;
    return CUSMC_OK;
}

// A plain cudaMemcpy into pageable memory runs at ~4 GB/s when the destination has never been
// touched (the R / numpy result arrays of run(): 240 MB for the C1 model): the driver's staging copy
// is single-threaded and takes every page fault itself.  Here the device side streams into two
// pinned 16 MiB slots at PCIe speed and up to 8 host threads copy a finished slot out in parallel.
int cusmc_d2h_staged(cusmc_ctx *ctx, void *dst_host, const void *src_dev, size_t bytes)
{
    constexpr size_t kSlot = (size_t)16 << 20;
    if (bytes == 0) return CUSMC_OK;
    if (bytes < kSlot / 4) {
        CUSMC_CUDA(ctx, cudaMemcpyAsync(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost, ctx->stream));
        CUSMC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        return CUSMC_OK;
    }
    void *pin = nullptr;
    CUSMC_CHECK(cusmc_pinned(ctx, 2 * kSlot, &pin));
    cudaEvent_t done[2] = {nullptr, nullptr};
    for (auto &e : done) CUSMC_CUDA(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    const unsigned hw = std::thread::hardware_concurrency();
    const int workers = (int)std::max(1u, std::min(8u, hw ? hw / 2 : 4u));
    const size_t chunks = (bytes + kSlot - 1) / kSlot;
    auto issue = [&](size_t c) {
        const size_t off = c * kSlot, n = std::min(kSlot, bytes - off);
        cudaMemcpyAsync((char *)pin + (c & 1) * kSlot, (const char *)src_dev + off, n, cudaMemcpyDeviceToHost, ctx->stream);
        cudaEventRecord(done[c & 1], ctx->stream);
    };
    issue(0);
    int rc = CUSMC_OK;
    for (size_t c = 0; c < chunks; ++c) {
        const size_t off = c * kSlot, n = std::min(kSlot, bytes - off);
        if (cudaEventSynchronize(done[c & 1]) != cudaSuccess) {
            rc = cusmc_fail(ctx, CUSMC_ERR_CUDA, "staged device-to-host copy failed: %s", cudaGetErrorString(cudaGetLastError()));
            break;
        }
        if (c + 1 < chunks) issue(c + 1);                 // the other slot: drained one iteration ago
        const char *src = (const char *)pin + (c & 1) * kSlot;
        char *dst = (char *)dst_host + off;
        const size_t part = ((n + workers - 1) / workers + 4095) & ~(size_t)4095;
        std::vector<std::thread> pool;
        size_t done_to = std::min(part, n);              // [0, part) is this thread's; helpers take the rest
        try {                                            // nothing may unwind across the C ABI
            pool.reserve(workers);
            for (int w = 1; w < workers; ++w) {
                const size_t lo = (size_t)w * part;
                if (lo >= n) break;
                pool.emplace_back([=] { std::memcpy(dst + lo, src + lo, std::min(part, n - lo)); });
                done_to = std::min(lo + part, n);
            }
        } catch (...) {
            // no more threads to be had: this thread copies what has not been handed out
        }
        std::memcpy(dst, src, std::min(part, n));
        if (done_to < n) std::memcpy(dst + done_to, src + done_to, n - done_to);
        for (auto &t : pool) t.join();
    }
    cudaStreamSynchronize(ctx->stream);
    for (auto &e : done) cudaEventDestroy(e);
    return rc;
}

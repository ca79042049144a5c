// pf_persist.cu -- the whole filter run as ONE cooperative kernel, for particle clouds that live in L2.
//
// At 10^6 particles (BASELINE configs[3]: the bootstrap filter on data_raw/y_t.csv) a step touches
// 44 MB that never leave the 126 MB L2, and the four-launch step of filter.cu is bound by the
// latency of its dependent ~9 us kernels.  A step has three grid-wide dependencies
//     max of the log-weights  ->  total fixed-point mass  ->  ancestors scattered to the children
// and here they are three grid barriers inside a persistent kernel instead of kernel boundaries:
// two blocks of 512 threads per SM, the cloud spread evenly over all of them (a tile of <= 4096
// particles per block), and everything a tile carries from one phase to the next (log-weights,
// tile-local CDF) stays in shared memory -- neither the log-weights nor the weight image are ever
// re-read from global memory.  (A first version kept them in registers, 16 particles per thread: 3.3
// resident warps per scheduler, latency-bound, slower than the four launches.)  Between its arrival
// at a barrier and the last block's, a block draws the next step's normals (pregen_noise below).
//
//   scatter(t) :  C_j = prefix(tile sums of t-1) + c_j  ->  children [k(C_{j-1}), k(C_j)) get ancestor j
//   ---- grid barrier ----
//   propagate(t): x_t[i] = G x_{t-1}[a_i] + Q z_i,  lw_i = log p(y_t | x_t[i]),  atomic max
//   ---- grid barrier ----
//   weigh(t)   :  q_i = fixed(exp(lw_i - max)), c_i = tile-local inclusive prefix, tile sum
//   ---- grid barrier ----
//
// The arithmetic is the four-launch path's, operation for operation (same helpers, same Philox
// counters): a persistent run reproduces it bit for bit (tests/test_gpu_parity.py).
// Same semantics as cusmc_filter_run (src/mcmc.cpp:239-309 of the reference) restricted to:
// one GPU, systematic resampling, Normal noise, d == dy in {2, 4}, device-drawn noise, no history,
// N <= one tile per resident block.
#include "filter_types.cuh"
#include "pf_step_impl.cuh"
#include "resample.cuh"

#include "../../include/cusmc_detmath.h"
#include "../../include/cusmc_philox.h"

#include <algorithm>
#include <type_traits>

namespace {

#ifndef CUSMC_PERSIST_UNROLL
#define CUSMC_PERSIST_UNROLL 1
#endif
constexpr int kPropagateUnroll = CUSMC_PERSIST_UNROLL;   // particles of a thread in flight in the propagate phase
constexpr int kThreads = 512;
// particles per thread: a template parameter P <= 8, picked so that N spreads evenly over the resident blocks
constexpr int kBlocksPerSM = 2;                 // 2 x 512 threads at <= 64 registers, 2 x 72 KB of shared memory
// shared-memory index of tile offset j: one pad word per 8, so both the striped (j = r*512 + tid)
// and the blocked (j = P*tid + r) access patterns stay (almost) conflict-free
__device__ __forceinline__ int pad(int j) { return j + (j >> 3); }
__host__ __device__ constexpr int padded_words(int P) { return kThreads * P + kThreads * P / 8 + 8; }
// Noise in the barrier shadow: the normals of step t + 1 depend on nothing but (seed, t + 1, slot),
// so a block generates them -- a third of its tile at each of the three grid barriers that precede
// propagate(t + 1) -- BETWEEN its arrival at the barrier and the moment the last block arrives, and
// parks them in shared memory as the single-precision values the generator produces (8 bytes per
// d = 2 particle).  The 150 of ~400 instructions per particle-step that the Philox block and the
// Box-Muller transform cost then run while the SM would otherwise spin.  d = 4 would need 2 x 64 KB
// more shared memory than two resident blocks have: it keeps drawing inside propagate.
#ifndef CUSMC_PERSIST_PREGEN
#define CUSMC_PERSIST_PREGEN 1
#endif
__host__ __device__ constexpr bool pregen_noise(int D) { return CUSMC_PERSIST_PREGEN && D == 2; }
__host__ __device__ constexpr size_t persist_smem_bytes(int D, int P)
{
    return 2 * sizeof(double) * (size_t)padded_words(P) + (pregen_noise(D) ? sizeof(float2) * (size_t)kThreads * P : 0);
}

struct PersistArgs {
    double *x[2];                   // SoA [d][ld] double buffer
    double *lw;                     // [N] log-weights (kept for cusmc_filter_state_dev)
    uint32_t *anc;                  // [N]
    StepSlot *slots;                // [T]
    unsigned long long *tile_sums;  // [2][gridDim.x]
    const double *obs;              // [T][D]: L_V^-1 y_t
    const double *u0;               // [T]: systematic offsets (entry t used by step t)
    double *moments;                // [T][2 + D] or NULL: sum w, -, sum w x_k (summary)
    unsigned *barrier;              // grid barrier counter, zero at launch
    uint64_t seed;
    int64_t ld;
    uint32_t N;
    uint32_t tile_n;                // particles per block (<= 512 P): N spread evenly over the resident blocks
    int T, shift;
};

// x = mu + G xp + Q z and the whitened residual norm, in pf_step_kernel's operation order.
template <int D, bool DIAG>
__device__ __forceinline__ void propagate_one(const pfstep::StepOp<D, DIAG> &op, const double (&c)[D], const double (&xp)[D],
                                              const double (&z)[D], double (&xn)[D], double &q)
{
#pragma unroll
    for (int k = 0; k < D; ++k) {
        double g = op.mu[k], s = 0.0;
        if constexpr (DIAG) {
            g = fma(op.G[k], xp[k], g);
            s = fma(op.Q[k], z[k], s);
        } else {
#pragma unroll
            for (int j = 0; j < D; ++j) g = fma(op.G[k * D + j], xp[j], g);
#pragma unroll
            for (int j = 0; j < D; ++j) s = fma(op.Q[k * D + j], z[j], s);
        }
        xn[k] = s + g;
    }
    q = 0.0;
#pragma unroll
    for (int k = 0; k < D; ++k) {
        double zk = c[k];
        if constexpr (DIAG) {
            zk = fma(-op.M[k], xn[k], zk);
        } else {
#pragma unroll
            for (int j = 0; j < D; ++j) zk = fma(-op.M[k * D + j], xn[j], zk);
        }
        q = fma(zk, zk, q);
    }
}

template <int D>
__device__ __forceinline__ void draw_normals(uint64_t seed, int stream, uint64_t step, uint64_t idx, double (&z)[D])
{
#pragma unroll
    for (int jq = 0; jq < (D + 3) / 4; ++jq) {
        double zq[4];
        cusmc_normal4(cusmc_rng(seed, stream, step, idx, (uint32_t)jq), zq);
#pragma unroll
        for (int e = 0; e < 4; ++e)
            if (4 * jq + e < D) z[4 * jq + e] = zq[e];
    }
}

// Grid barrier on a monotonic arrival counter (the kernel is launched cooperatively, so every block
// is resident).  Everything that crosses a barrier is read with L2 loads (__ldcg), so unlike
// cooperative_groups' grid.sync() the wait loop does not have to invalidate L1 on every poll.
// Split in two so that work which needs nothing from the other blocks runs between them.
__device__ __forceinline__ void grid_barrier_arrive(unsigned *bar, unsigned &target)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        target += gridDim.x;
        __threadfence();                        // release: the block's writes, cumulative through the bar.sync
        atomicAdd(bar, 1u);
    }
}
__device__ __forceinline__ void grid_barrier_wait(unsigned *bar, unsigned target)
{
    if (threadIdx.x == 0) {
        unsigned seen;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(bar) : "memory");
        } while ((int)(seen - target) < 0);
    }
    __syncthreads();
}

// Block-wide sums of two uint64 per thread (every thread gets both totals).
__device__ __forceinline__ void block_sum2(unsigned long long &a, unsigned long long &b, unsigned long long *sm)
{
    constexpr int kWarps = kThreads / 32;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        b += __shfl_xor_sync(0xffffffffu, b, o);
    }
    __syncthreads();
    if ((threadIdx.x & 31) == 0) {
        sm[threadIdx.x >> 5] = a;
        sm[kWarps + (threadIdx.x >> 5)] = b;
    }
    __syncthreads();
    a = b = 0;
#pragma unroll
    for (int k = 0; k < kWarps; ++k) {
        a += sm[k];
        b += sm[kWarps + k];
    }
}

// Phase mappings of a block's tile of 512 P particles:
//   propagate : STRIPED, particle tile + r*512 + tid (round r) -- every global access of a warp is one
//               contiguous line, one particle in flight per thread at a time (small register footprint,
//               12 resident warps per scheduler hide the gather latency);
//   weigh / scatter : BLOCKED, particles tile + P*tid .. +P-1 -- the tile-local CDF is a thread-local
//               running sum plus one block scan.
// The log-weights and the CDF cross between the two mappings, and between phases, through shared
// memory (2 x 36 KB per block): neither is ever re-read from global memory.
template <int D, bool DIAG, int P, bool SUMMARY>
__global__ void __launch_bounds__(kThreads, kBlocksPerSM)
pf_persistent_kernel(const __grid_constant__ pfstep::StepOp<D, DIAG> op_init,
                     const __grid_constant__ pfstep::StepOp<D, DIAG> op, const Epilogue ep, const PersistArgs a)
{
    unsigned bar_target = 0;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int kItems = P, kPadded = padded_words(P);
    double *s_lw = reinterpret_cast<double *>(smem_raw);                               // [kPadded]
    unsigned long long *s_c = reinterpret_cast<unsigned long long *>(smem_raw) + kPadded;   // [kPadded]
    constexpr bool kPregen = pregen_noise(D);
#ifndef CUSMC_PERSIST_CHUNK_DIV
#define CUSMC_PERSIST_CHUNK_DIV 3
#endif
    constexpr int kPregenChunk = (P + CUSMC_PERSIST_CHUNK_DIV - 1) / CUSMC_PERSIST_CHUNK_DIV;   // rounds generated per barrier: three barriers cover the tile
    static_assert(3 * kPregenChunk >= P, "the three barriers before a propagate must cover the tile");
    float2 *s_z = reinterpret_cast<float2 *>(smem_raw + 2 * sizeof(double) * kPadded);     // [kThreads * P], striped
    int z_done = 0;                                    // rounds of the NEXT propagate whose normals sit in s_z
    __shared__ unsigned long long s_u64[2 * (kThreads / 32)];
    __shared__ double s_dbl[kThreads / 32];
    __shared__ uint32_t s_k[kThreads];
    __shared__ unsigned long long s_T, s_r0;
    __shared__ double s_ng_over_t, s_r0_over_t;

    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t tile0 = blockIdx.x * a.tile_n;
    const uint32_t tile_n = min(a.tile_n, a.N - tile0);     // particles of this tile
    int cur = 0;

    // one particle: gather (t > 0), noise, propagate, reweight; striped round r
    auto particle = [&](auto is_init, const pfstep::StepOp<D, DIAG> &o, const double (&cobs)[D], int t_step, int r, double &m) {
        constexpr bool kInit = decltype(is_init)::value;
        const int t = kInit ? 0 : t_step;
        const int j = r * kThreads + (int)tid;
        const uint32_t i = tile0 + (uint32_t)j;
        double lw = -INFINITY;
        if ((uint32_t)j < tile_n) {
            double xp[D], z[D], xn[D], q;
            if (!kInit) {
                const uint32_t par = __ldcg(a.anc + i);
                const double *src = a.x[cur] + par;
#pragma unroll
                for (int k = 0; k < D; ++k) xp[k] = __ldcg(src + (int64_t)k * a.ld);
            } else {
#pragma unroll
                for (int k = 0; k < D; ++k) xp[k] = 0.0;
            }
            if (kPregen && !kInit) {
                // drawn in the shadow of the three barriers since the last propagate (all kItems rounds:
                // 3 x kPregenChunk >= kItems)
                const float2 zz = s_z[j];
                z[0] = (double)zz.x;
                z[D > 1 ? 1 : 0] = (double)zz.y;
            } else {
                draw_normals<D>(a.seed, kInit ? CUSMC_STREAM_INIT : CUSMC_STREAM_NORMAL, (uint64_t)t, (uint64_t)i, z);
            }
            propagate_one<D, DIAG>(o, cobs, xp, z, xn, q);
            double *dst = a.x[kInit ? 0 : cur ^ 1] + i;
#pragma unroll
            for (int k = 0; k < D; ++k) st_stream(dst + (int64_t)k * a.ld, xn[k]);
            lw = kInit ? 0.0 : density_epilogue(ep, q);
            st_stream(a.lw + i, lw);
            if (lw == lw && lw < INFINITY && lw > m) m = lw;
        }
        s_lw[pad(j)] = lw;
    };

    // grid barrier; between arrival and release the block draws the next chunk of step t_next's normals
    auto grid_barrier = [&](int t_next) {
        grid_barrier_arrive(a.barrier, bar_target);
        if (kPregen && t_next < a.T && z_done < kItems) {
            const int hi = min(kItems, z_done + kPregenChunk);
#pragma unroll 1
            for (int r = z_done; r < hi; ++r) {
                const int j = r * kThreads + (int)tid;
                if ((uint32_t)j < tile_n) {
                    // the first Box-Muller pair of the particle's block 0 (cusmc_normal4 with D = 2)
                    const cusmc_u32x4 rb = cusmc_rng(a.seed, CUSMC_STREAM_NORMAL, (uint64_t)t_next, (uint64_t)(tile0 + (uint32_t)j), 0u);
                    float2 zz;
                    cusmc_box_muller_f32(rb.v[0], rb.v[1], &zz.x, &zz.y);
                    s_z[j] = zz;
                }
            }
            z_done = hi;
        }
        grid_barrier_wait(a.barrier, bar_target);
    };

    // block max -> atomic max into the step's slot
    auto publish_max = [&](int t, double m) {
        m = warp_max_double(m);
        if (lane == 0) s_dbl[warp] = m;
        __syncthreads();
        if (tid < 32) {
            m = warp_max_double(tid < kThreads / 32 ? s_dbl[tid] : -INFINITY);
            if (tid == 0) atomic_max_double(&a.slots[t].lw_max, m);
        }
    };

    // weigh(t): fixed-point weights against the global max, tile-local CDF into shared memory, tile sum;
    // with the summary on, also the ESS sum and the weighted first moments of the step
    auto weigh = [&](int t) {
        const double wmax = __ldcg(&a.slots[t].lw_max);
        constexpr bool summary = SUMMARY;
        unsigned long long c[kItems], run = 0, s2 = 0;
#pragma unroll
        for (int r = 0; r < kItems; ++r) {
            const double wn = cusmc_unit_from_log(s_lw[pad(kItems * (int)tid + r)], wmax);
            run += cusmc_fixed_from_unit(wn, a.shift);
            c[r] = run;
            if (summary) s2 += cusmc_fixed_from_unit(wn * wn, a.shift);
        }
        unsigned long long inc = run;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long v = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += v;
        }
        __syncthreads();
        if (lane == 31) s_u64[warp] = inc;
        __syncthreads();
        unsigned long long before = inc - run, tile_total = 0;
#pragma unroll
        for (int k = 0; k < kThreads / 32; ++k) {
            const unsigned long long v = s_u64[k];
            if (k < (int)warp) before += v;
            tile_total += v;
        }
#pragma unroll
        for (int r = 0; r < kItems; ++r) s_c[pad(kItems * (int)tid + r)] = before + c[r];
        if (tid == 0) a.tile_sums[(size_t)(t & 1) * gridDim.x + blockIdx.x] = tile_total;
        if constexpr (!SUMMARY) return;
        // ESS: integer sum of the squared weights (order-independent)
        unsigned long long dummy = 0;
        block_sum2(s2, dummy, s_u64);
        if (tid == 0 && s2) atomicAdd((unsigned long long *)&a.slots[t].sum_q2, s2);
        // weighted first moments, striped (coalesced reads of the state); the weight of particle j is
        // the difference of neighbouring CDF entries -- exactly the fixed-point weight resampling uses
        __syncthreads();
        const double scale = cusmc_pow2i(-a.shift);
        const double *xc = a.x[cur];
        double acc[1 + D];
#pragma unroll
        for (int k = 0; k <= D; ++k) acc[k] = 0.0;
#pragma unroll 2
        for (int r = 0; r < kItems; ++r) {
            const int j = r * kThreads + (int)tid;
            if ((uint32_t)j < tile_n) {
                const unsigned long long qj = s_c[pad(j)] - (j ? s_c[pad(j - 1)] : 0ull);
                const double w = (double)qj * scale;
                acc[0] += w;
#pragma unroll
                for (int k = 0; k < D; ++k) acc[1 + k] = fma(w, __ldcg(xc + (int64_t)k * a.ld + tile0 + j), acc[1 + k]);
            }
        }
#pragma unroll
        for (int k = 0; k <= D; ++k) {
            double v = acc[k];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            __syncthreads();
            if (lane == 0) s_dbl[warp] = v;
            __syncthreads();
            if (tid == 0) {
                double tsum = 0.0;
                for (int q = 0; q < kThreads / 32; ++q) tsum += s_dbl[q];
                atomicAdd(a.moments + (size_t)t * (2 + D) + (k ? 1 + k : 0), tsum);
            }
        }
    };

    // ---- t = 0: x_0 = m0 + Q_c0 z, constant log-weight 0 (src/mcmc.cpp:63-85) -----------------
    {
        const double zero_c[D] = {};
        double m = -INFINITY;
#pragma unroll 1
        for (int r = 0; r < kItems; ++r) particle(std::true_type{}, op_init, zero_c, 0, r, m);
        publish_max(0, m);
    }
    grid_barrier(1);
    weigh(0);
    grid_barrier(1);

    for (int t = 1; t < a.T; ++t) {
        // ---- scatter(t): ancestors of step t from the weight image of step t - 1 ----------------
        {
            const unsigned long long *ts = a.tile_sums + (size_t)((t - 1) & 1) * gridDim.x;
            unsigned long long pre = 0, tot = 0;
            for (uint32_t b = tid; b < gridDim.x; b += kThreads) {
                const unsigned long long v = __ldcg(ts + b);
                tot += v;
                if (b < blockIdx.x) pre += v;
            }
            block_sum2(pre, tot, s_u64);
            if (tid == 0) {
                if (blockIdx.x == 0) a.slots[t - 1].sum_q = tot;
                uint64_t rr = (uint64_t)(__ldg(a.u0 + t) * (double)tot);
                if (tot && rr > tot - 1) rr = tot - 1;
                s_T = tot;
                s_r0 = rr;
                s_ng_over_t = (double)a.N / (double)tot;
                s_r0_over_t = (double)rr / (double)tot;
            }
            __syncthreads();
            const uint64_t T = s_T;
            if (T != 0) {
                const uint64_t r0 = s_r0, Ng = a.N;
                const double ng_over_t = s_ng_over_t, r0_over_t = s_r0_over_t;
                uint32_t k[kItems];
                unsigned long long c_prev = ~0ull;
                uint32_t k_last = 0;
#pragma unroll
                for (int r = 0; r < kItems; ++r) {
                    // the count is a pure function of the CDF value: zero weights repeat it for free
                    const unsigned long long cr = s_c[pad(kItems * (int)tid + r)];
                    k[r] = cr != c_prev ? (uint32_t)offspring_below(pre + cr, Ng, T, r0, ng_over_t, r0_over_t) : k_last;
                    c_prev = cr;
                    k_last = k[r];
                }
                s_k[tid] = k[kItems - 1];
                __syncthreads();
                uint32_t k_prev = tid ? s_k[tid - 1] : (uint32_t)offspring_below(pre, Ng, T, r0, ng_over_t, r0_over_t);
#pragma unroll
                for (int r = 0; r < kItems; ++r) {
                    uint32_t lo = k_prev;
                    const uint32_t hi = k[r];
                    const uint32_t parent = tile0 + kItems * tid + r;      // (beyond the tile: zero weight, no children)
                    const uint32_t n = hi > lo ? hi - lo : 0u;
                    const bool big = n > 8;
                    const uint32_t ns = big ? 0u : n;
                    // small families: predicated store slots, as many as the warp's largest needs
                    const uint32_t slots = __reduce_max_sync(0xffffffffu, ns);
                    uint32_t *dst = a.anc + lo;
#pragma unroll
                    for (uint32_t q = 0; q < 8; ++q) {
                        if (q >= slots) break;
                        if (q < ns) dst[q] = parent;
                    }
                    unsigned bigmask = __ballot_sync(0xffffffffu, big);
                    while (bigmask) {
                        const int src = __ffs(bigmask) - 1;
                        bigmask &= bigmask - 1;
                        const uint32_t sa = __shfl_sync(0xffffffffu, lo, src);
                        const uint32_t sb = __shfl_sync(0xffffffffu, hi, src);
                        const uint32_t sp = __shfl_sync(0xffffffffu, parent, src);
#pragma unroll 1
                        for (uint32_t ch = sa + lane; ch < sb; ch += 32) a.anc[ch] = sp;
                    }
                    k_prev = hi;
                }
            } else {
                // no mass to resample from: identity ancestors, reported by the host getters
#pragma unroll
                for (int r = 0; r < kItems; ++r)
                    if ((uint32_t)(kItems * (int)tid + r) < tile_n) a.anc[tile0 + kItems * tid + r] = tile0 + kItems * tid + r;
                if (blockIdx.x == 0 && tid == 0) a.slots[t].degenerate = 1;
            }
        }
        grid_barrier(t);

        // ---- propagate(t) + reweight(t) (src/mcmc.cpp:298-307), max of the log-weights ----------
        {
            double cobs[D];
#pragma unroll
            for (int k = 0; k < D; ++k) cobs[k] = __ldg(a.obs + (size_t)t * D + k);
            double m = -INFINITY;
#pragma unroll kPropagateUnroll
            for (int r = 0; r < kItems; ++r) particle(std::false_type{}, op, cobs, t, r, m);
            publish_max(t, m);
            cur ^= 1;
            z_done = 0;                                 // s_z is free: it refills with step t + 1's normals
        }
        grid_barrier(t + 1);

        // ---- weigh(t) ---------------------------------------------------------------------------
        weigh(t);
        grid_barrier(t + 1);
    }
    // total mass of the last step (log-likelihood of the summary)
    if (blockIdx.x == 0) {
        const unsigned long long *ts = a.tile_sums + (size_t)((a.T - 1) & 1) * gridDim.x;
        unsigned long long pre = 0, tot = 0;
        for (uint32_t b = tid; b < gridDim.x; b += kThreads) tot += __ldcg(ts + b);
        block_sum2(pre, tot, s_u64);
        if (tid == 0) a.slots[a.T - 1].sum_q = tot;
    }
}

template <int D, bool DIAG, int P, bool SUMMARY>
int launch_persistent(cusmc_filter *f, const PersistArgs &args, bool probe_only)
{
    cusmc_ctx *ctx = f->ctx;
    const cusmc_filter_config &cfg = f->cfg;
    auto kernel = pf_persistent_kernel<D, DIAG, P, SUMMARY>;
    constexpr size_t kSmem = persist_smem_bytes(D, P);     // log-weights + CDF of one tile (+ parked normals)
    const unsigned grid = (unsigned)((cfg.N + args.tile_n - 1) / args.tile_n);
    CUSMC_CUDA(ctx, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem));
    int per_sm = 0;
    CUSMC_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kThreads, kSmem));
    if ((int64_t)per_sm * ctx->sm_count < (int64_t)grid)
        return cusmc_fail(ctx, CUSMC_ERR_UNSUPPORTED, "persistent run: %u tiles exceed the %d resident blocks", grid,
                          per_sm * ctx->sm_count);
    if (probe_only) return CUSMC_OK;
    pfstep::StepOp<D, DIAG> op0, op;
    const pfstep::StepModel m0{cfg.d, cfg.dy, nullptr, f->Qc0.data(), cfg.noise_scale, nullptr, nullptr, f->m0.data()};
    const pfstep::StepModel m1{cfg.d, cfg.dy, f->G.data(), f->Qw.data(), cfg.noise_scale, &f->M, nullptr, nullptr};
    pfstep::fill_step_op<D, DIAG>(op0, m0);
    pfstep::fill_step_op<D, DIAG>(op, m1);
    Epilogue ep = f->ep;
    PersistArgs a = args;
    void *params[] = {&op0, &op, &ep, &a};
    const cudaError_t e = cudaLaunchCooperativeKernel((const void *)kernel, dim3(grid), dim3(kThreads), params, kSmem,
                                                      ctx->stream);
    if (e == cudaErrorCooperativeLaunchTooLarge || e == cudaErrorNotSupported) {
        cudaGetLastError();
        return cusmc_fail(ctx, CUSMC_ERR_UNSUPPORTED, "cooperative launch refused: %s", cudaGetErrorString(e));
    }
    CUSMC_CUDA(ctx, e);
    ctx->launches++;
    return CUSMC_OK;
}

// Particles per tile: N spread evenly over ALL resident block slots (every SM gets the same work; a
// fixed tile size would leave some SMs with half the work of their neighbours), and the smallest
// particles-per-thread count P that holds such a tile.
uint32_t pick_tile(const cusmc_filter *f)
{
    const int64_t slots = (int64_t)f->ctx->sm_count * kBlocksPerSM;
    int64_t n = (f->cfg.N + slots - 1) / slots;
    n = (n + 31) & ~(int64_t)31;                       // whole warps in the striped rounds
    return (uint32_t)std::max<int64_t>(n, 32);
}
int pick_items(const cusmc_filter *f)
{
    const uint32_t n = pick_tile(f);
    for (int P : {2, 4, 6, 7, 8})
        if (n <= (uint32_t)(kThreads * P)) return P;
    return 0;
}

template <int D, bool DIAG, bool SUMMARY>
int launch_persistent_p(cusmc_filter *f, const PersistArgs &args, int P, bool probe_only)
{
    switch (P) {
        case 2: return launch_persistent<D, DIAG, 2, SUMMARY>(f, args, probe_only);
        case 4: return launch_persistent<D, DIAG, 4, SUMMARY>(f, args, probe_only);
        case 6: return launch_persistent<D, DIAG, 6, SUMMARY>(f, args, probe_only);
        case 7: return launch_persistent<D, DIAG, 7, SUMMARY>(f, args, probe_only);
        case 8: return launch_persistent<D, DIAG, 8, SUMMARY>(f, args, probe_only);
    }
    return cusmc_fail(f->ctx, CUSMC_ERR_UNSUPPORTED, "persistent run: too many particles for one tile per resident block");
}

template <int D, bool DIAG>
int launch_persistent_s(cusmc_filter *f, const PersistArgs &args, int P, bool probe_only)
{
    return f->cfg.summary ? launch_persistent_p<D, DIAG, true>(f, args, P, probe_only)
                          : launch_persistent_p<D, DIAG, false>(f, args, P, probe_only);
}

int launch_persistent_any(cusmc_filter *f, const PersistArgs &args, int d, bool diag, int P, bool probe_only)
{
    if (d == 2) return diag ? launch_persistent_s<2, true>(f, args, P, probe_only) : launch_persistent_s<2, false>(f, args, P, probe_only);
    return diag ? launch_persistent_s<4, true>(f, args, P, probe_only) : launch_persistent_s<4, false>(f, args, P, probe_only);
}

}  // namespace

bool cusmc_filter_persistent_eligible(const cusmc_filter *f, const cusmc_filter_draws *draws)
{
    const cusmc_filter_config &cfg = f->cfg;
    if (true) return false;   // TODO(round 2): being rewritten on the block-relative weight image
    if (cfg.persistent < 0 || f->world != 1) return false;
    if (cfg.resampler != CUSMC_RESAMPLE_SYSTEMATIC || cfg.kind != CUSMC_MVN) return false;
    if (cfg.keep_history || cfg.ess_threshold > 0.0) return false;
    if (cfg.d != cfg.dy || (cfg.d != 2 && cfg.d != 4)) return false;
    if (cfg.T < 2) return false;
    if (draws && (draws->xi0_dev || draws->xi_dev || draws->chi_dev || draws->u_dev || draws->j_dev || draws->um_dev))
        return false;
    return pick_items(f) != 0;                        // refined by the occupancy query at launch
}

int cusmc_filter_run_persistent(cusmc_filter *f, const cusmc_filter_draws *draws)
{
    cusmc_ctx *ctx = f->ctx;
    const cusmc_filter_config &cfg = f->cfg;
    if (!cusmc_filter_persistent_eligible(f, draws))
        return cusmc_fail(ctx, CUSMC_ERR_UNSUPPORTED, "configuration not covered by the persistent kernel");
    const int d = cfg.d, T = cfg.T;
    const int P = pick_items(f);
    const uint32_t tile_n = pick_tile(f);
    const unsigned grid = (unsigned)((cfg.N + tile_n - 1) / tile_n);
    bool diag = cusmc_is_diag_colmajor(f->G.data(), d) && cusmc_is_diag_colmajor(f->Qw.data(), d) && cusmc_is_diag_colmajor(f->Qc0.data(), d);
    for (int k = 0; k < d && diag; ++k)
        for (int j = 0; j < d; ++j)
            if (j != k && f->M[(size_t)k * d + j] != 0.0) diag = false;
    PersistArgs a{};
    a.tile_n = tile_n;
    CUSMC_CUDA(ctx, cudaSetDevice(ctx->device));
    // resident-block check before anything is enqueued (so the caller can still fall back)
    int rc = launch_persistent_any(f, a, d, diag, P, true);
    if (rc != CUSMC_OK) return rc;

    cudaStream_t st = ctx->stream;
    // per-run scratch: tile sums [2][grid], whitened observations [T][d], systematic offsets [T]
    const size_t n_sum = 2 * (size_t)grid, n_obs = (size_t)T * d, n_u0 = (size_t)T;
    const size_t bytes = 8 * (n_sum + n_obs + n_u0 + 1);      // + the grid barrier counter
    if (bytes > f->persist_bytes) {
        CUSMC_CUDA(ctx, cudaStreamSynchronize(st));
        cudaFree(f->persist);
        f->persist = nullptr;
        f->persist_bytes = 0;
        CUSMC_CUDA(ctx, cudaMalloc(&f->persist, bytes));
        f->persist_bytes = bytes;
    }
    std::vector<double> host(n_obs + n_u0, 0.0);
    for (int t = 0; t < T; ++t) {
        cusmc_whiten_observation(f->Winv, d, f->Y.data() + (size_t)t * d, &host[(size_t)t * d]);
        if (t >= 1)
            host[n_obs + t] = (draws && draws->u0_host) ? draws->u0_host[t - 1]
                                                        : (double)(cusmc_u0_bits(cfg.seed, (uint64_t)t) >> 11) * 1.1102230246251565e-16;
    }
    unsigned long long *sums = (unsigned long long *)f->persist;
    double *obs = (double *)(sums + n_sum), *u0 = obs + n_obs;
    unsigned *barrier = (unsigned *)(u0 + n_u0);
    CUSMC_CUDA(ctx, cudaMemsetAsync(barrier, 0, 8, st));
    // the host vector dies with this call: a synchronous copy (pageable memory) is what we want
    CUSMC_CUDA(ctx, cudaMemcpyAsync(obs, host.data(), 8 * host.size(), cudaMemcpyHostToDevice, st));
    CUSMC_CUDA(ctx, cudaStreamSynchronize(st));
    CUSMC_CHECK(cusmc_filter_init_slots(f));
    if (cfg.summary) CUSMC_CUDA(ctx, cudaMemsetAsync(f->moments, 0, sizeof(double) * (size_t)T * (2 + d), st));
    a.x[0] = f->x[0];
    a.x[1] = f->x[1];
    a.lw = f->lw;
    a.anc = f->anc;
    a.slots = f->slots;
    a.tile_sums = sums;
    a.obs = obs;
    a.u0 = u0;
    a.moments = cfg.summary ? f->moments : nullptr;
    a.barrier = barrier;
    a.seed = cfg.seed;
    a.ld = f->per;
    a.N = (uint32_t)cfg.N;
    a.T = T;
    a.shift = f->shift;
    CUSMC_CUDA(ctx, cudaEventRecord(f->ev0, st));
    rc = launch_persistent_any(f, a, d, diag, P, false);
    if (rc != CUSMC_OK) return rc;
    CUSMC_CUDA(ctx, cudaEventRecord(f->ev1, st));
    f->cur = (T - 1) & 1;
    f->next_t = T;
    f->ran = true;
    return CUSMC_OK;
}

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
// Sharded runs: the weight vector is the concatenation of every rank's shard (per_rank slots
// each), read through peer-mapped pointers over NVLink.
struct PeerWeights {
    const double *const *w;      // device table of the ranks' weight arrays
    FastDiv per_rank;
};

template <bool PEERS>
__device__ __forceinline__ double weight_at(const double *__restrict__ w, const PeerWeights &pw, uint32_t idx)
{
    if (PEERS) {
        const uint32_t r = fast_div(idx, pw.per_rank);
        return __ldg(pw.w[r] + (idx - r * pw.per_rank.d));
    }
    return __ldg(w + idx);
}

// C2 = true: the Metropolis-C2 variant (Dulger et al., "Memory coalescing for parallelised Metropolis
// resampling"): at every iteration the 32 particles of a warp draw their proposals from ONE
// common 32-particle segment of the weight vector (256 bytes: 8 sectors instead of 32 scattered ones), a
// fresh segment per iteration.  The segment is the one holding a uniform index drawn from a counter keyed
// by (i / 32, n) -- so it is picked with probability proportional to its length and every proposal is
// still uniform over 0 .. N-1 -- and the lane's own draw picks the slot inside it.  oracle:
// orc_rng_metropolis_c2.
template <bool PREDRAWN, bool PEERS, bool C2>
__global__ void __launch_bounds__(kThreads)
metropolis_kernel(uint32_t *__restrict__ a, const double *__restrict__ w, const PeerWeights pw,
                  const double *__restrict__ u, const uint32_t *__restrict__ j, uint64_t seed,
                  uint64_t step, int64_t N, int B, int is_log, int64_t i0, int64_t n_out)
{
    cusmc_pdl_enter();
    const int64_t t = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    const bool live = t < n_out;
    if (!C2 && !live) return;              // (C2 keeps whole warps alive: lanes draw segments for each other)
    const int64_t i = i0 + t;              // global particle index (i0 = 0 on one GPU)
    uint32_t k = (uint32_t)i;
    double wk = live ? weight_at<PEERS>(w, pw, (uint32_t)i) : 0.0;
    // C2: the segment of (group, n) is the same for the 32 particles of a group.  When the warp IS the
    // group (i0 a multiple of 32), lane l draws the segments of iterations nb + l for everybody, 32
    // iterations per Philox block per lane, handed round by shuffles; otherwise every lane draws its own.
    const uint32_t lane = threadIdx.x & 31;
    const bool shared_draw = C2 && (i0 & 31) == 0;
    uint32_t seg_mine = 0;
    auto segment_of = [&](int n) {
        const cusmc_u32x4 rs = cusmc_rng(seed, CUSMC_STREAM_SEGMENT, step, (uint64_t)i >> 5, (uint32_t)n);
        return (uint32_t)cusmc_uint_below(rs.v[0], rs.v[1], (uint64_t)N) & ~31u;
    };
    for (int n = 0; n < B; ++n) {
        double un = 0.0;
        uint32_t jn = 0;
        if (PREDRAWN) {
            un = __ldg(u + t * B + n);
            jn = __ldg(j + t * B + n);
        } else {
            uint32_t first = 0;
            if (C2) {
                if (shared_draw) {
                    if ((n & 31) == 0 && n + (int)lane < B) seg_mine = segment_of(n + (int)lane);
                    first = __shfl_sync(0xffffffffu, seg_mine, n & 31);
                } else {
                    first = segment_of(n);
                }
            }
            const cusmc_u32x4 r = cusmc_rng(seed, CUSMC_STREAM_METROPOLIS, step, (uint64_t)i, (uint32_t)n);
            un = cusmc_u01(r.v[0], r.v[1]);
            if (C2) {
                const uint64_t len = (uint64_t)N - first < 32u ? (uint64_t)N - first : 32u;
                jn = first + (uint32_t)cusmc_uint_below(r.v[2], r.v[3], len);
            } else {
                jn = (uint32_t)cusmc_uint_below(r.v[2], r.v[3], (uint64_t)N);
            }
        }
        if (C2 && !live) continue;
        const double wj = weight_at<PEERS>(w, pw, jn);
        // linear: u <= w_j / w_k (0/0 = NaN rejects, x/0 = inf accepts, as on the CPU);
        // log   : u <= exp(lw_j - lw_k) with the reproducible exp.
        // (An exact multiply-and-compare shortcut for the division was tried: no gain -- the kernel is
        // bound by the 32-byte sectors its random 8-byte reads pull out of L2, ~6 TB/s at N = 10^6.)
        const bool accept = un <= (is_log ? cusmc_det_exp(wj - wk) : wj / wk);
        if (accept) {
            k = jn;
            wk = wj;
        }
    }
    if (live) a[t] = k;
}

// Rejection resampler: the unbiased relative of the rule above (no counterpart in the reference; the
// resampler_f seam of inst/include/types.hpp:32 would register it).  Attempt n consumes the Philox block
// the Metropolis resampler uses for (seed, step, i, n): its uniform decides, its index is the next proposal.
__global__ void __launch_bounds__(kThreads)
rejection_kernel(uint32_t *__restrict__ a, const double *__restrict__ w, const double *__restrict__ wmax_p,
                 uint64_t seed, uint64_t step, int64_t N, int cap)
{
    const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (i >= N) return;
    const double wmax = *wmax_p;
    uint32_t k = (uint32_t)i;
    double wk = __ldg(w + i);
    for (int n = 0; n < cap; ++n) {
        const cusmc_u32x4 r = cusmc_rng(seed, CUSMC_STREAM_METROPOLIS, step, (uint64_t)i, (uint32_t)n);
        if (cusmc_u01(r.v[0], r.v[1]) <= wk / wmax) break;          // NaN (0 / 0) never accepts, as in the reference's rule
        k = (uint32_t)cusmc_uint_below(r.v[2], r.v[3], (uint64_t)N);
        wk = __ldg(w + k);
    }
    a[i] = k;
}

// ------------------------------------------------------------------------------------------
// max and fixed-point sums
// ------------------------------------------------------------------------------------------
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
    m = warp_max_double(m);
    __shared__ double sm[kThreads / 32];
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x < 32) {
        m = warp_max_double(threadIdx.x < kThreads / 32 ? sm[threadIdx.x] : -INFINITY);
        if (threadIdx.x == 0) atomic_max_double(out, m);
    }
}

// ------------------------------------------------------------------------------------------
// The weight image: normalisation sums, tile prefixes and tile-local CDF of the fixed-point weights.
//
// A tile is kTile = 2048 consecutive weights; thread t of the block owns the 8 consecutive weights
// 8t .. 8t+7 (four 128-bit loads).  Workspace layout, uint64 words (cusmc_scan_state_bytes):
//     [0 .. 7]            ScatterConsts: per-launch constants of the resampling pass (resample.cuh)
//     [8 + b]             sum of tile b, turned IN PLACE into the exclusive prefix over tiles by
//                         tile_scan_kernel (one block; ~2 us at N = 8 Mi)
//     [8 + tiles + b]     sum of squared weights of tile b   } FULL only (ESS); reduced by
//     [8 + 2 tiles + b]   positive weights in tile b         } tile_scan_kernel
//     [H + i]             inclusive prefix of weight i INSIDE its tile (H = header words, even)
// exp() is evaluated once per weight, here.  The resampling pass that follows needs neither the
// weights nor a scan of its own: global CDF_i = offset + prefix[tile(i)] + local_i, one thread per
// particle, tiles independent -- no decoupled look-back, no spinning, no atomics, no state to clear
// between steps.  (A first version had every block atomicAdd its sums into one word and the last
// block to arrive scan the tile sums: 2 same-address atomics per tile cost 18 us at N = 8 Mi.)
// Integer sums: every order gives the same bits.
// ------------------------------------------------------------------------------------------
__host__ __device__ inline int64_t image_tiles(int64_t N) { return (N + kTile - 1) / kTile; }
__host__ __device__ inline int64_t image_header_words(int64_t N)
{
    return (kImageHead + 3 * image_tiles(N) + 3) & ~(int64_t)3;       // rounded up to 32 bytes
}

// 256-bit global accesses (sm_100: LDG/STG.E.ENL2.256): a thread's 8 weights are two loads, its 8
// prefixes two stores, every request a whole 32-byte sector.
__device__ __forceinline__ void ldg256(const double *p, double &a, double &b, double &c, double &d)
{
    asm volatile("ld.global.nc.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p));
}
__device__ __forceinline__ void stg256(unsigned long long *p, unsigned long long a, unsigned long long b,
                                       unsigned long long c, unsigned long long d)
{
    asm volatile("st.global.v4.u64 [%0], {%1, %2, %3, %4};" ::"l"(p), "l"(a), "l"(b), "l"(c), "l"(d) : "memory");
}

__device__ __forceinline__ void load_tile_items(const double *__restrict__ w, int64_t base, int64_t N,
                                                double (&v)[kTileItems])
{
    static_assert(kTileItems % 4 == 0, "tile items come in 256-bit groups");
    if (base + kTileItems <= N && (((uintptr_t)(w + base)) & 31) == 0) {
#pragma unroll
        for (int r = 0; r < kTileItems / 4; ++r) ldg256(w + base + 4 * r, v[4 * r], v[4 * r + 1], v[4 * r + 2], v[4 * r + 3]);
    } else {
#pragma unroll
        for (int r = 0; r < kTileItems; ++r) v[r] = base + r < N ? __ldg(w + base + r) : -INFINITY;
    }
}

// Block-wide sum of one uint64 per thread; every thread gets the total.
__device__ __forceinline__ unsigned long long block_sum_u64(unsigned long long v, unsigned long long *sm)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
    __syncthreads();
    unsigned long long t = 0;
#pragma unroll
    for (int k = 0; k < kThreads / 32; ++k) t += sm[k];
    return t;
}

// FULL: also sum of squares and positive count (ESS); otherwise only what resampling needs.
#ifndef CUSMC_WEIGH_MINB
#define CUSMC_WEIGH_MINB 6
#endif
#ifndef CUSMC_SCAN_MINB
#define CUSMC_SCAN_MINB 8
#endif
template <bool FULL, bool LOG>
__global__ void __launch_bounds__(kThreads, LOG ? CUSMC_WEIGH_MINB : 4)
weigh_kernel(const double *__restrict__ w, const double *__restrict__ wmax_p, int64_t N,
             int shift, unsigned long long *__restrict__ image)
{
    __shared__ unsigned long long sm[kThreads / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const double wmax = *wmax_p;
    const int64_t base = (int64_t)blockIdx.x * kTile + (int64_t)threadIdx.x * kTileItems;
    double v[kTileItems];
    load_tile_items(w, base, N, v);
    unsigned long long c[kTileItems];   // inclusive prefix within the thread
    unsigned long long run = 0, s2 = 0, np = 0;
#pragma unroll
    for (int r = 0; r < kTileItems; ++r) {
        // -inf padding -> 0
        const double wn = LOG ? cusmc_unit_from_log(v[r], wmax) : cusmc_unit_from_linear(v[r], wmax);
        const uint64_t q = cusmc_fixed_from_unit(wn, shift);
        run += q;
        c[r] = run;
        if (FULL) {
            s2 += cusmc_fixed_from_unit(wn * wn, shift);
            np += q > 0;
        }
    }
    // block-wide exclusive prefix of the thread totals
    unsigned long long inc = run;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) sm[warp] = inc;
    __syncthreads();
    unsigned long long before = inc - run, tile_total = 0;
#pragma unroll
    for (int k = 0; k < kThreads / 32; ++k) {
        const unsigned long long t = sm[k];
        if (k < warp) before += t;
        tile_total += t;
    }
    // tile-local inclusive CDF (the workspace is padded to whole tiles: no bounds checks)
    unsigned long long *local = image + image_header_words(N) + base;
    if ((((uintptr_t)local) & 31) == 0) {        // always, unless the caller supplied an unaligned image
#pragma unroll
        for (int r = 0; r < kTileItems / 4; ++r)
            stg256(local + 4 * r, before + c[4 * r], before + c[4 * r + 1], before + c[4 * r + 2], before + c[4 * r + 3]);
    } else {
#pragma unroll
        for (int r = 0; r < kTileItems; ++r) local[r] = before + c[r];
    }
    if (FULL) {
        s2 = block_sum_u64(s2, sm);
        np = block_sum_u64(np, sm);
    }
    if (threadIdx.x == 0) {
        const int64_t tiles = gridDim.x;
        image[kImageHead + blockIdx.x] = tile_total;
        if (FULL) {
            image[kImageHead + tiles + blockIdx.x] = s2;
            image[kImageHead + 2 * tiles + blockIdx.x] = np;
        }
    }
}

// One block: exclusive prefix of the tile sums, in place, and the totals
// stats[0..2] = { sum q, sum q2, #positive } (plain stores: nothing to zero beforehand).  A thread
// owns kScanItems consecutive tiles, all loads of a chunk are issued before the first dependent
// instruction: 4096 tiles (N = 8 Mi) are one chunk, one block-wide scan.
constexpr int kScanThreads = 1024;
constexpr int kScanItems = 4;
__global__ void __launch_bounds__(kScanThreads)
tile_scan_kernel(unsigned long long *__restrict__ image, int64_t tiles, unsigned long long *__restrict__ stats,
                 int full)
{
    __shared__ unsigned long long sm[kScanThreads / 32];
    __shared__ unsigned long long s_chunk;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned long long *sums = image + kImageHead;
    unsigned long long carry = 0, s2 = 0, np = 0;
    for (int64_t base = 0; base < tiles; base += kScanThreads * kScanItems) {
        const int64_t i0 = base + (int64_t)threadIdx.x * kScanItems;
        unsigned long long v[kScanItems], run = 0;
#pragma unroll
        for (int r = 0; r < kScanItems; ++r) v[r] = i0 + r < tiles ? sums[i0 + r] : 0ull;
        if (full) {
#pragma unroll
            for (int r = 0; r < kScanItems; ++r)
                if (i0 + r < tiles) {
                    s2 += sums[tiles + i0 + r];
                    np += sums[2 * tiles + i0 + r];
                }
        }
#pragma unroll
        for (int r = 0; r < kScanItems; ++r) run += v[r];
        unsigned long long inc = run;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) sm[warp] = inc;
        __syncthreads();
        if (warp == 0) {                       // exclusive scan of the 32 warp totals
            const unsigned long long w = sm[lane];
            unsigned long long wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned long long t = __shfl_up_sync(0xffffffffu, wi, o);
                if (lane >= o) wi += t;
            }
            sm[lane] = wi - w;
            if (lane == 31) s_chunk = wi;
        }
        __syncthreads();
        unsigned long long excl = carry + sm[warp] + inc - run;
#pragma unroll
        for (int r = 0; r < kScanItems; ++r) {
            if (i0 + r < tiles) sums[i0 + r] = excl;
            excl += v[r];
        }
        carry += s_chunk;
        __syncthreads();
    }
    if (!stats) return;
    if (full) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            s2 += __shfl_xor_sync(0xffffffffu, s2, o);
            np += __shfl_xor_sync(0xffffffffu, np, o);
        }
        __shared__ unsigned long long sm2[kScanThreads / 32];
        if (lane == 0) {
            sm[warp] = s2;
            sm2[warp] = np;
        }
        __syncthreads();
        if (warp == 0) {
            s2 = sm[lane];
            np = sm2[lane];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                s2 += __shfl_xor_sync(0xffffffffu, s2, o);
                np += __shfl_xor_sync(0xffffffffu, np, o);
            }
        }
    }
    if (threadIdx.x == 0) {                 // thread 0 holds the running total and warp 0's reductions
        stats[0] = carry;
        if (full) {
            stats[1] = s2;
            stats[2] = np;
        }
    }
}

struct ScanArgs {
    const unsigned long long *total;       // global fixed-point mass (device)
    const unsigned long long *cdf_offset;  // mass on lower shards, or NULL
    const unsigned long long *tile_prefix; // weight image: exclusive prefix of tile b at [b]
    const unsigned long long *local;       // weight image: inclusive prefix of weight i inside its tile
    ScatterConsts *consts;                 // weight image: the per-launch constants of the scatter
    unsigned long long *cdf_out;           // optional inclusive global CDF
    uint32_t *anc_out;                     // optional systematic ancestors for the children out_lo .. out_hi - 1
    uint32_t N, N_global, j0, out_lo, out_hi;   // all < 2^32 (ancestors are 32-bit)
    double u0;
};

// The per-launch constants of the scatter, computed once by one thread (the caller supplies the total).
__global__ void scatter_consts_kernel(const ScanArgs p)
{
    if (threadIdx.x == 0) *p.consts = make_scatter_consts(*p.total, p.u0, p.N_global, 0.0, 0ull);
}

__device__ __forceinline__ void put_ancestor(const ScanArgs &p, uint32_t child, uint32_t parent)
{
    p.anc_out[child - p.out_lo] = parent;
}

// Global CDF from the weight image and, fused in, the systematic offspring scatter: parent j owns
// the child slots [k(C_{j-1}), k(C_j)) and writes its own index into them.  A warp owns 32 kPar
// consecutive parents, lane l the parents l, l + 32, ... of them: in round r the lanes hold 32
// CONSECUTIVE parents, so their loads of the tile-local CDF and -- since neighbouring parents own
// neighbouring child ranges -- their ancestor stores coalesce (a thread owning 4 consecutive parents
// spread every store instruction of the warp over 16 sectors).  Indices and offspring counts are 32-bit
// throughout (N_global < 2^32).  Warps are independent: the per-launch constants come precomputed
// (ScatterConsts), a parent's left neighbour's count arrives by shuffle (lane 31's of the round before
// for lane 0) and lane 0 evaluates the count below the warp's first parent itself -- no shared
// memory, no barrier, nothing between a warp's loads and its stores but its own arithmetic.
constexpr int kPar = 4;
static_assert(kTile % (32 * kPar) == 0, "a warp of the resampling pass must lie inside one tile");

// The rounds of one warp.  FAST: all 32 kPar parents exist and the children are not clipped to a
// sub-range (every filter step but a warp at the end of the cloud): no per-parent predicates.
template <bool FAST>
__device__ __forceinline__ void scatter_rounds(const ScanArgs &p, const ScatterConsts &c, const uint64_t (&C)[kPar],
                                               uint64_t Cbelow, uint32_t i0, uint32_t lane)
{
    const uint64_t T = c.T, r0 = c.r0, Ng = p.N_global;
    const double ng_over_t = c.ng_over_t, r0_over_t = c.r0_over_t;
    // the offspring count is a pure function of the CDF value
    uint32_t k[kPar];
#pragma unroll
    for (int r = 0; r < kPar; ++r)
        k[r] = (FAST || i0 + 32 * r < p.N) ? (uint32_t)offspring_below(C[r], Ng, T, r0, ng_over_t, r0_over_t) : 0u;
    uint32_t k_carry = 0;                              // count below the round's first parent (lane 0)
    if (lane == 0) k_carry = (uint32_t)offspring_below(Cbelow, Ng, T, r0, ng_over_t, r0_over_t);
#pragma unroll
    for (int r = 0; r < kPar; ++r) {
        const bool active = FAST || i0 + 32 * r < p.N;
        uint32_t k_prev = __shfl_up_sync(0xffffffffu, k[r], 1);
        if (lane == 0) k_prev = k_carry;
        k_carry = __shfl_sync(0xffffffffu, k[r], 31);  // lane 0 uses it as the next round's k_prev; a full round is all active
        // parents past the end own the empty range; their threads stay to help with large families
        uint32_t a = FAST ? k_prev : (active ? max(k_prev, p.out_lo) : 0u);
        const uint32_t b = FAST ? k[r] : (active ? min(k[r], p.out_hi) : 0u);
        const uint32_t parent = p.j0 + i0 + 32 * r;
        // small families (<= 8 children): the owning thread writes them, the warp running as many
        // predicated store slots as its largest small family needs (a uniform trip count: two
        // instructions per slot instead of a divergent six-instruction loop); large ones: the whole
        // warp helps
        const uint32_t n = b > a ? b - a : 0u;
        const bool big = n > 8;
        const uint32_t ns = big ? 0u : n;
        const uint32_t slots = __reduce_max_sync(0xffffffffu, ns);
        {
            uint32_t *dst = p.anc_out + (a - p.out_lo);       // slot q: one compare, one store at dst + 4 q
#pragma unroll
            for (uint32_t q = 0; q < 8; ++q) {
                if (q >= slots) break;
                if (q < ns) dst[q] = parent;
            }
        }
        unsigned bigmask = __ballot_sync(0xffffffffu, big);
        while (bigmask) {
            const int src = __ffs(bigmask) - 1;
            bigmask &= bigmask - 1;
            const uint32_t sa = __shfl_sync(0xffffffffu, a, src);
            const uint32_t sb = __shfl_sync(0xffffffffu, b, src);
            const uint32_t sp = __shfl_sync(0xffffffffu, parent, src);
#pragma unroll 1
            for (uint32_t cc = sa + lane; cc < sb; cc += 32) put_ancestor(p, cc, sp);
        }
    }
}

__global__ void __launch_bounds__(kThreads, CUSMC_SCAN_MINB)
scan_resample_kernel(const ScanArgs p)
{
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t w0 = ((blockIdx.x * kThreads + threadIdx.x) >> 5) * (32 * kPar);   // the warp's first local parent
    if (w0 >= p.N) return;                                                 // whole warp past the end
    const uint32_t i0 = w0 + lane;                                        // parents i0 + 32 r
    const uint32_t tile = w0 / kTile;                                     // a warp lies inside one tile
    const bool scatter = p.anc_out != nullptr;
    const bool full = w0 + 32 * kPar <= p.N;
    // every load goes out before the first use
    const uint64_t prefix = __ldg(p.tile_prefix + tile);
    const uint64_t offset = p.cdf_offset ? __ldg(p.cdf_offset) : 0ull;
    uint64_t C[kPar], Cprev = 0;
#pragma unroll
    for (int r = 0; r < kPar; ++r) C[r] = (full || i0 + 32 * r < p.N) ? __ldg(p.local + i0 + 32 * r) : 0ull;
    if (scatter && lane == 0 && (w0 % kTile)) Cprev = __ldg(p.local + w0 - 1);
    ScatterConsts c{};
    if (scatter) {
        c.T = __ldg(&p.consts->T);
        c.r0 = __ldg(&p.consts->r0);
        c.ng_over_t = __ldg(&p.consts->ng_over_t);
        c.r0_over_t = __ldg(&p.consts->r0_over_t);
    }
    const uint64_t base = prefix + offset;
#pragma unroll
    for (int r = 0; r < kPar; ++r) C[r] += base;
    if (p.cdf_out) {
#pragma unroll
        for (int r = 0; r < kPar; ++r)
            if (i0 + 32 * r < p.N) p.cdf_out[i0 + 32 * r] = C[r];
    }
    if (!scatter) return;
    if (c.T == 0) {                                   // no mass: identity ancestors (the host entry points report it)
#pragma unroll
        for (int r = 0; r < kPar; ++r)
            if (i0 + 32 * r < p.N) {
                const uint32_t g = p.j0 + i0 + 32 * r;
                if (g >= p.out_lo && g < p.out_hi) put_ancestor(p, g, g);
            }
        return;
    }
    if (full && p.out_lo == 0 && p.out_hi >= p.N_global)
        scatter_rounds<true>(p, c, C, base + Cprev, i0, lane);
    else
        scatter_rounds<false>(p, c, C, base + Cprev, i0, lane);
}

// Multinomial: a[i] = j0 + #{ j : cdf_j <= p_i },  p_i = min((uint64)(u_i * T), T - 1).
template <bool PREDRAWN>
__global__ void __launch_bounds__(kThreads)
multinomial_kernel(const unsigned long long *__restrict__ cdf, int64_t N,
                   const unsigned long long *__restrict__ total_p, const double *__restrict__ u,
                   uint64_t seed, uint64_t step, int64_t i0, int64_t n_out, int64_t j0,
                   uint32_t *__restrict__ a, unsigned long long *degenerate_out)
{
    const int64_t t = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (t >= n_out) return;
    const uint64_t T = *total_p;
    if (T == 0) {                                     // no mass: identity ancestors, flagged
        a[t] = (uint32_t)(j0 + i0 + t);
        if (t == 0 && degenerate_out) *degenerate_out = 1;
        return;
    }
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
                            int is_log, int64_t i0, int64_t n_out, const CusmcPeers *peers, bool c2)
{
    if (n_out == 0) return CUSMC_OK;
    const unsigned grid = (unsigned)((n_out + kThreads - 1) / kThreads);
    PeerWeights pw{};
#define CUSMC_METRO_GO(PR, PE, C2)                                                                                       \
    CUSMC_CUDA(ctx, cusmc_launch_pdl(metropolis_kernel<PR, PE, C2>, grid, kThreads, 0, ctx->stream, a, w, pw, u, j, seed, step, \
                                     N, B, is_log, i0, n_out))
    if (peers) {
        pw.w = (const double *const *)peers->table_dev;
        pw.per_rank = make_fast_div((uint32_t)peers->per_rank);
        if (u) CUSMC_METRO_GO(true, true, false);          // injected proposals are the caller's, whatever the variant
        else if (c2) CUSMC_METRO_GO(false, true, true);
        else CUSMC_METRO_GO(false, true, false);
    } else {
        if (u) CUSMC_METRO_GO(true, false, false);
        else if (c2) CUSMC_METRO_GO(false, false, true);
        else CUSMC_METRO_GO(false, false, false);
    }
#undef CUSMC_METRO_GO
    CUSMC_LAUNCHED(ctx);
    return CUSMC_OK;
}

int cusmc_launch_rejection(cusmc_ctx *ctx, uint32_t *a, const double *w, const double *wmax_dev, uint64_t seed,
                           uint64_t step, int64_t N, int cap)
{
    if (N == 0) return CUSMC_OK;
    rejection_kernel<<<(unsigned)((N + kThreads - 1) / kThreads), kThreads, 0, ctx->stream>>>(a, w, wmax_dev, seed, step, N, cap);
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

// image: cusmc_scan_state_bytes(N) bytes (see weigh_kernel).  stats_dev may be NULL (image only);
// full_stats adds the sum of squares and the positive count (ESS).  Two launches: the tiles, then
// one block that turns the tile sums into prefixes and totals.
int cusmc_launch_weights_sum(cusmc_ctx *ctx, const double *w, int is_log, const double *max_dev,
                             int64_t N, int shift, uint64_t *stats_dev, void *image, bool full_stats)
{
    if (N == 0) return CUSMC_OK;
    const unsigned tiles = (unsigned)image_tiles(N);
    const bool full = full_stats && stats_dev;
    unsigned long long *img = (unsigned long long *)image;
    if (full && is_log)
        weigh_kernel<true, true><<<tiles, kThreads, 0, ctx->stream>>>(w, max_dev, N, shift, img);
    else if (full)
        weigh_kernel<true, false><<<tiles, kThreads, 0, ctx->stream>>>(w, max_dev, N, shift, img);
    else if (is_log)
        weigh_kernel<false, true><<<tiles, kThreads, 0, ctx->stream>>>(w, max_dev, N, shift, img);
    else
        weigh_kernel<false, false><<<tiles, kThreads, 0, ctx->stream>>>(w, max_dev, N, shift, img);
    CUSMC_LAUNCHED(ctx);
    tile_scan_kernel<<<1, kScanThreads, 0, ctx->stream>>>(img, (int64_t)tiles, (unsigned long long *)stats_dev, full ? 1 : 0);
    CUSMC_LAUNCHED(ctx);
    return CUSMC_OK;
}

size_t cusmc_scan_state_bytes(int64_t N)
{
    return sizeof(unsigned long long) * (size_t)(image_header_words(N) + image_tiles(N) * kTile);
}

// image must hold the weight image of exactly these weights (same w, max, N, shift).
int cusmc_launch_scan(cusmc_ctx *ctx, int64_t N, int64_t N_global, const uint64_t *total_dev,
                      const uint64_t *cdf_offset_dev, const void *image,
                      uint64_t *cdf_out, uint32_t *anc_out, int64_t j0, int64_t out_lo,
                      int64_t out_n, double u0)
{
    if (N == 0) return CUSMC_OK;
    if (N_global > 0xFFFFFFFFll || out_lo < 0 || out_lo + out_n > N_global)
        return cusmc_fail(ctx, CUSMC_ERR_INVALID, "resampling range outside 0..N_global (< 2^32)");
    ScanArgs p{};
    p.total = (const unsigned long long *)total_dev;
    p.cdf_offset = (const unsigned long long *)cdf_offset_dev;
    p.tile_prefix = (const unsigned long long *)image + kImageHead;
    p.local = (const unsigned long long *)image + image_header_words(N);
    p.consts = (ScatterConsts *)const_cast<void *>(image);
    p.cdf_out = (unsigned long long *)cdf_out;
    p.anc_out = anc_out;
    p.N = (uint32_t)N;
    p.N_global = (uint32_t)N_global;
    p.j0 = (uint32_t)j0;
    p.out_lo = (uint32_t)out_lo;
    p.out_hi = (uint32_t)(out_lo + out_n);
    p.u0 = u0;
    const unsigned grid = (unsigned)((N + kThreads * kPar - 1) / (kThreads * kPar));
    if (anc_out) {
        scatter_consts_kernel<<<1, 32, 0, ctx->stream>>>(p);
        CUSMC_LAUNCHED(ctx);
    }
    scan_resample_kernel<<<grid, kThreads, 0, ctx->stream>>>(p);
    CUSMC_LAUNCHED(ctx);
    return CUSMC_OK;
}

int cusmc_launch_multinomial(cusmc_ctx *ctx, const uint64_t *cdf, int64_t N, const uint64_t *total_dev,
                             const double *u, uint64_t seed, uint64_t step, int64_t i0,
                             int64_t n_out, int64_t j0, uint32_t *a, uint64_t *degenerate_dev)
{
    if (n_out == 0) return CUSMC_OK;
    const unsigned grid = (unsigned)((n_out + kThreads - 1) / kThreads);
    if (u)
        multinomial_kernel<true><<<grid, kThreads, 0, ctx->stream>>>(
            (const unsigned long long *)cdf, N, (const unsigned long long *)total_dev, u, seed, step, i0, n_out, j0, a,
            (unsigned long long *)degenerate_dev);
    else
        multinomial_kernel<false><<<grid, kThreads, 0, ctx->stream>>>(
            (const unsigned long long *)cdf, N, (const unsigned long long *)total_dev, u, seed, step, i0, n_out, j0, a,
            (unsigned long long *)degenerate_dev);
    CUSMC_LAUNCHED(ctx);
    return CUSMC_OK;
}

// ---- extern "C": device-pointer API -----------------------------------------------------------
extern "C" int cusmc_metropolis_hastings_dev(cusmc_ctx *ctx, uint32_t *a_dev, const double *w_dev,
                                             const double *u_dev, const uint32_t *j_dev,
                                             uint64_t seed, uint64_t step, int64_t N, int B, int is_log)
{
    CUSMC_ENTER(ctx);
    CUSMC_REQUIRE(ctx, N >= 0 && B >= 0, "N, B must be non-negative");
    CUSMC_REQUIRE(ctx, N == 0 || (a_dev && w_dev), "a/w is NULL");
    CUSMC_REQUIRE(ctx, (u_dev == nullptr) == (j_dev == nullptr), "u and j must both be given or both NULL");
    CUSMC_REQUIRE(ctx, N <= 0xFFFFFFFFll, "N exceeds the 32-bit ancestor range");
    return cusmc_launch_metropolis(ctx, a_dev, w_dev, u_dev, j_dev, seed, step, N, B, is_log, 0, N, nullptr, false);
}

extern "C" int cusmc_metropolis_c2_dev(cusmc_ctx *ctx, uint32_t *a_dev, const double *w_dev, uint64_t seed, uint64_t step,
                                       int64_t N, int B, int is_log)
{
    CUSMC_ENTER(ctx);
    CUSMC_REQUIRE(ctx, N >= 0 && B >= 0, "N, B must be non-negative");
    CUSMC_REQUIRE(ctx, N == 0 || (a_dev && w_dev), "a/w is NULL");
    CUSMC_REQUIRE(ctx, N <= 0xFFFFFFFFll, "N exceeds the 32-bit ancestor range");
    return cusmc_launch_metropolis(ctx, a_dev, w_dev, nullptr, nullptr, seed, step, N, B, is_log, 0, N, nullptr, true);
}

extern "C" int cusmc_rejection_resample_dev(cusmc_ctx *ctx, uint32_t *a_dev, const double *w_dev,
                                            const double *w_max_dev, uint64_t seed, uint64_t step, int64_t N, int cap)
{
    CUSMC_ENTER(ctx);
    CUSMC_REQUIRE(ctx, N >= 0 && cap >= 1, "N >= 0 and cap >= 1 required");
    CUSMC_REQUIRE(ctx, N == 0 || (a_dev && w_dev && w_max_dev), "NULL pointer");
    CUSMC_REQUIRE(ctx, N <= 0xFFFFFFFFll, "N exceeds the 32-bit ancestor range");
    return cusmc_launch_rejection(ctx, a_dev, w_dev, w_max_dev, seed, step, N, cap);
}

extern "C" int cusmc_weights_max_dev(cusmc_ctx *ctx, const double *w_dev, int64_t N, double *max_dev)
{
    CUSMC_ENTER(ctx);
    CUSMC_REQUIRE(ctx, max_dev && (N == 0 || w_dev), "NULL pointer");
    CUSMC_CHECK(cusmc_fill_double(ctx, max_dev, -INFINITY, 1));
    return cusmc_launch_weights_max(ctx, w_dev, N, max_dev);
}

// Tile-prefix storage for the building-block entry points: the caller's buffer
// (cusmc_tile_prefix_words(N) words, word 0 zero before first use) or context scratch.
static int tile_state_for(cusmc_ctx *ctx, int64_t N, uint64_t *user, void **out)
{
    if (user) {
        *out = user;
        return CUSMC_OK;
    }
    return cusmc_scratch(ctx, 7, cusmc_scan_state_bytes(N), out);
}

extern "C" int64_t cusmc_tile_prefix_words(int64_t N)
{
    return N < 0 ? 0 : (int64_t)(cusmc_scan_state_bytes(N) / sizeof(uint64_t));
}

extern "C" int cusmc_weights_sum_dev(cusmc_ctx *ctx, const double *w_dev, int is_log,
                                     const double *max_dev, int64_t N, int64_t N_global,
                                     uint64_t *stats_dev, uint64_t *tile_prefix_dev)
{
    CUSMC_ENTER(ctx);
    CUSMC_REQUIRE(ctx, stats_dev && max_dev && (N == 0 || w_dev), "NULL pointer");
    CUSMC_REQUIRE(ctx, N_global >= N && N_global >= 1, "N_global < N");
    if (N == 0) {
        CUSMC_CUDA(ctx, cudaMemsetAsync(stats_dev, 0, 3 * sizeof(uint64_t), ctx->stream));
        return CUSMC_OK;
    }
    void *state = nullptr;
    CUSMC_CHECK(tile_state_for(ctx, N, tile_prefix_dev, &state));
    return cusmc_launch_weights_sum(ctx, w_dev, is_log, max_dev, N, cusmc_fixed_shift(N_global), stats_dev, state, true);
}

// Tile prefixes for a scan: the caller's (from cusmc_weights_sum_dev on the same weights) or a
// fresh reduction into context scratch.
static int scan_prefixes(cusmc_ctx *ctx, const double *w_dev, int is_log, const double *max_dev, int64_t N,
                         int64_t N_global, const uint64_t *tile_prefix_dev, const void **state)
{
    if (tile_prefix_dev) {
        *state = tile_prefix_dev;
        return CUSMC_OK;
    }
    void *mine = nullptr;
    CUSMC_CHECK(tile_state_for(ctx, N, nullptr, &mine));
    CUSMC_CHECK(cusmc_launch_weights_sum(ctx, w_dev, is_log, max_dev, N, cusmc_fixed_shift(N_global), nullptr, mine, false));
    *state = mine;
    return CUSMC_OK;
}

extern "C" int cusmc_weights_scan_dev(cusmc_ctx *ctx, const double *w_dev, int is_log,
                                      const double *max_dev, int64_t N, int64_t N_global,
                                      const uint64_t *cdf_offset_dev, const uint64_t *tile_prefix_dev,
                                      uint64_t *cdf_dev)
{
    CUSMC_ENTER(ctx);
    CUSMC_REQUIRE(ctx, max_dev && (N == 0 || (w_dev && cdf_dev)), "NULL pointer");
    CUSMC_REQUIRE(ctx, N_global >= N && N_global >= 1, "N_global < N");
    if (N == 0) return CUSMC_OK;
    const void *state = nullptr;
    CUSMC_CHECK(scan_prefixes(ctx, w_dev, is_log, max_dev, N, N_global, tile_prefix_dev, &state));
    return cusmc_launch_scan(ctx, N, N_global, nullptr, cdf_offset_dev, state, cdf_dev, nullptr, 0, 0, 0, 0.0);
}

extern "C" int cusmc_resample_systematic_dev(cusmc_ctx *ctx, const double *w_dev, int is_log,
                                             const double *max_dev, int64_t N_local, int64_t N_global,
                                             const uint64_t *total_dev, const uint64_t *cdf_offset_dev,
                                             const uint64_t *tile_prefix_dev, int64_t j0, int64_t out_lo,
                                             int64_t out_n, double u0, uint32_t *a_dev)
{
    CUSMC_ENTER(ctx);
    CUSMC_REQUIRE(ctx, max_dev && total_dev && (N_local == 0 || w_dev) && (out_n == 0 || a_dev), "NULL pointer");
    CUSMC_REQUIRE(ctx, N_global >= N_local && N_global >= 1, "N_global < N_local");
    CUSMC_REQUIRE(ctx, N_global <= 0xFFFFFFFFll, "N exceeds the 32-bit ancestor range");
    CUSMC_REQUIRE(ctx, u0 >= 0.0 && u0 < 1.0, "u0 must lie in [0, 1)");
    if (N_local == 0) return CUSMC_OK;
    const void *state = nullptr;
    CUSMC_CHECK(scan_prefixes(ctx, w_dev, is_log, max_dev, N_local, N_global, tile_prefix_dev, &state));
    return cusmc_launch_scan(ctx, N_local, N_global, total_dev, cdf_offset_dev, state, nullptr, a_dev, j0, out_lo,
                             out_n, u0);
}

extern "C" int cusmc_resample_multinomial_dev(cusmc_ctx *ctx, const uint64_t *cdf_dev, int64_t N,
                                              const uint64_t *total_dev, const double *u_dev,
                                              uint64_t seed, uint64_t step, int64_t i0, int64_t n_out,
                                              int64_t j0, uint32_t *a_dev)
{
    CUSMC_ENTER(ctx);
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
    CUSMC_CHECK(cusmc_weights_sum_dev(ctx, hw.w_dev, is_log, hw.max_dev, N, N, hw.stats_dev, nullptr));
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
    CUSMC_ENTER(ctx);
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
    CUSMC_ENTER(ctx);
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
    // the reduction above left this weight vector's tile prefixes in scratch slot 7
    CUSMC_CHECK(cusmc_resample_systematic_dev(ctx, hw.w_dev, 0, hw.max_dev, N, N, hw.stats_dev, nullptr,
                                              (const uint64_t *)ctx->scratch[7], 0, 0, N, u0, (uint32_t *)ad));
    CUSMC_CUDA(ctx, cudaMemcpyAsync(a, ad, sizeof(uint32_t) * (size_t)N, cudaMemcpyDeviceToHost, ctx->stream));
    CUSMC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return CUSMC_OK;
}

extern "C" int cusmc_resample_multinomial(cusmc_ctx *ctx, const double *w, int64_t N, const double *u,
                                          uint32_t *a)
{
    CUSMC_ENTER(ctx);
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
    CUSMC_CHECK(cusmc_weights_scan_dev(ctx, hw.w_dev, 0, hw.max_dev, N, N, nullptr,
                                       (const uint64_t *)ctx->scratch[7], (uint64_t *)cd));
    CUSMC_CHECK(cusmc_resample_multinomial_dev(ctx, (const uint64_t *)cd, N, hw.stats_dev, (const double *)ud,
                                               0, 0, 0, N, 0, (uint32_t *)ad));
    CUSMC_CUDA(ctx, cudaMemcpyAsync(a, ad, sizeof(uint32_t) * (size_t)N, cudaMemcpyDeviceToHost, ctx->stream));
    CUSMC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return CUSMC_OK;
}

extern "C" int cusmc_normalize_ess(cusmc_ctx *ctx, const double *lw, int64_t N, double *lse, double *ess)
{
    CUSMC_ENTER(ctx);
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

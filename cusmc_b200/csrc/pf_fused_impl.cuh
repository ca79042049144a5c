// pf_fused_impl.cuh -- the fused filter step: resample + propagate + reweight + weigh in ONE kernel
// (included by pf_fused_mvn.cu / pf_fused_mvt.cu, one translation unit per noise family).
//
// A block owns one tile of kTile = 2048 consecutive CHILD slots and does, for that tile,
//
//   1. lookup    : systematic resampling, child-centric.  The children of a tile descend from a window
//                  of consecutive parents.  Starting at the parent tile tile_update_kernel left in the
//                  tile record, the block walks parent tiles: thread t evaluates the offspring counts
//                  k(C_j) of its 8 parents from the weight image of step t - 1 (C_j = rank_off + P_b +
//                  (c_j F_b >> 62), 128-bit exact where the floating estimate is not safe) and writes
//                  j into the shared-memory table for the children [k(C_{j-1}), k(C_j)) that fall in
//                  the block's tile.  No ancestor array round trip, no scatter pass of its own.
//   2. propagate : 8 striped rounds of one child per thread (pf_particle.cuh): gather the parent
//                  (a window of ~2048 consecutive columns: the gather is nearly sequential), noise,
//                  x_new = G x + noise, lw = log p(y_t | x_new); coalesced stores.
//   3. weigh     : the tile's log-weights never leave shared memory: block maximum m_b, fixed-point
//                  weights q_i = trunc(exp(lw_i - m_b) 2^shift), tile-local inclusive prefix c_i
//                  (256-bit stores, 8 bytes per particle) and the tile record {m_b, sum q, sum q^2}.
//
// Replaces scan_resample_kernel + pf_step_kernel + weigh_kernel of round 1 (three passes, 168 bytes
// per particle-step at d = 8, 274 us per 8 Mi particles) by one pass of 4 + 8 + 8d + 8d + 8 = 148 bytes.
// The per-step grid-wide dependency (global maximum, tile prefix, total mass) is the one-block
// tile_update_kernel between two launches of this kernel.
//
// Reference: resampler seam inst/include/types.hpp:32 (systematic has no counterpart upstream, SURVEY
// a11), propagate_K src/mcmc.cpp:90-160, reweight_G src/mcmc.cpp:162-237.  oracle: orc_filter_det.
#pragma once

#include "image.cuh"
#include "pf_particle.cuh"

namespace pffused {

using pfstep::StepOp;

constexpr int kThreads = kResampleThreads;                 // 256
constexpr int kItems = kTileItems;                         // 8 children per thread
constexpr int kPadded = kTile + kTile / 8;                 // shared-memory words of one padded tile
static_assert(kThreads * kItems == kTile, "a block owns exactly one tile");

enum ParentMode { kParentSelf = 0, kParentArray = 1, kParentLookup = 2 };

struct FusedArgs {
    StepArgs s;                                   // x, noise, history rows, lw (optional store), rng keys
    const unsigned long long *img_prev;           // weight image of step t - 1 (this rank)
    unsigned long long *img_new;                  // weight image of step t (this rank)
    const unsigned long long *const *img_prev_peer;   // world > 1: device table of every rank's image t - 1
    uint32_t *anc_out;                            // optional: the parents this kernel used (global ids)
    int64_t img_hdr_words;                        // words before c[] in an image (same on every rank)
    uint32_t tiles_alloc;                         // tiles per rank as laid out: field f of tile b is word 16 + f tiles_alloc + b
    uint32_t N_global;
    uint32_t tiles_per_rank;                      // parent tile tau lives on rank tau / tiles_per_rank
    uint32_t tile_n;                              // particles per tile (kTile; the persistent kernel spreads N
                                                  // evenly over its resident blocks: tile_n <= kTile, c stride kTile)
    int mode;                                     // ParentMode
    int accumulate;                               // adaptive resampling: add the old log-weight when not resampled
    int shift;
    int pdl;                                      // launch as a programmatic dependent of the preceding tile update
    double *trace;                                // profiling builds: 8 time stamps per step (else NULL)
};

// shared-memory index of tile offset j: one pad word per 8, so both the striped (j = r*256 + tid)
// and the blocked (j = 8*tid + r) access patterns stay conflict-free
__device__ __forceinline__ int pad(int j) { return j + (j >> 3); }

__device__ __forceinline__ void ldg256u(const unsigned long long *p, unsigned long long &a, unsigned long long &b,
                                        unsigned long long &c, unsigned long long &d)
{
    asm volatile("ld.global.nc.v4.u64 {%0, %1, %2, %3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
}
// the same through L2 only: data other blocks wrote earlier in the SAME launch (persistent kernel)
__device__ __forceinline__ void ldcg256u(const unsigned long long *p, unsigned long long &a, unsigned long long &b,
                                         unsigned long long &c, unsigned long long &d)
{
    asm volatile("ld.global.cg.v4.u64 {%0, %1, %2, %3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p) : "memory");
}
template <bool COH>
__device__ __forceinline__ unsigned long long ld_word(const unsigned long long *p)
{
    return COH ? __ldcg(p) : __ldg(p);
}
__device__ __forceinline__ void stg256u(unsigned long long *p, unsigned long long a, unsigned long long b,
                                        unsigned long long c, unsigned long long d)
{
    asm volatile("st.global.v4.u64 [%0], {%1, %2, %3, %4};" ::"l"(p), "l"(a), "l"(b), "l"(c), "l"(d) : "memory");
}

// ---- 1. lookup -------------------------------------------------------------------------------------
// Fills s_anc[pad(j)] = global parent of child i_a + j for j < n_tile.
//
// Everything that can be decided in the integer mass domain is: with Q_i = floor((i T + r0) / N)
// (mass_quotient, two per block) "the CDF value C reaches past child i" is the 64-bit compare C > Q_i,
// so which parent tiles to walk, when to stop and which threads hold parents of this tile's children
// cost no offspring count at all.  Only those threads evaluate k(C_j) for their 8 parents, and each
// parent with children leaves ONE marker -- its index at the slot of its first child inside the tile;
// a block-wide running maximum then spreads the markers over the slots (parents and slots both
// ascend), so family sizes never matter: no per-child loops, no divergence on heavy parents.
//
// TAB: the tile fields F / P / Sp of the parents' image sit in SHARED memory (the persistent kernel, where
// every block runs the tile update itself after the grid barrier): the prefix search and the fields of a
// walked tile cost no L2 round trip; only the CDF words of the walked tiles are fetched.
struct TileTab {
    const unsigned long long *F, *P, *Sp;
};

template <bool PEERS, bool COH, bool TAB = false>
__device__ __forceinline__ void lookup_parents(const FusedArgs &fa, const StepConsts &sc, uint32_t i_a, uint32_t n_tile,
                                               uint32_t *__restrict__ s_anc, unsigned long long *__restrict__ s_q,
                                               uint32_t *__restrict__ s_warp, const TileTab &tab = TileTab{})
{
    static_assert(!(TAB && PEERS), "tile tables in shared memory: one rank");
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint64_t T = sc.T, r0 = sc.r0, Ng = fa.N_global;
    const double ngt = sc.ng_over_t, r0t = sc.r0_over_t;
    const uint32_t i_b = i_a + n_tile;
    auto kof = [&](uint64_t C) { return (uint32_t)offspring_below(C, Ng, T, r0, ngt, r0t); };

    if constexpr (!TAB) {         // (TAB: the block's own tile update left the quotients and the zeroed table)
        if (tid == 0) s_q[0] = mass_quotient(i_a, T, r0, Ng);
        if (tid == 32) s_q[1] = mass_quotient(i_b - 1, T, r0, Ng);
#pragma unroll
        for (int r = 0; r < kItems; ++r) s_anc[pad(r * kThreads + (int)tid)] = 0u;
        __syncthreads();
    }
    const uint64_t Qa = s_q[0], Qb = s_q[1];          // C > Qa: reaches past the first child; C > Qb: past the last

    // the parent tile of the first child: largest tile whose exclusive prefix is <= Qa -- first the rank
    // (block-uniform), then two block-wide rounds over that rank's compact prefix array, every thread
    // one compare, the block barrier counting the hits
    uint32_t rk = 0;
    if (PEERS)
        for (int r = 1; r < fa.s.world; ++r)
            if (sc.rank_off[r] <= Qa) rk = (uint32_t)r;
    uint32_t tau;
    {
        const unsigned long long *img = (PEERS && rk != (uint32_t)fa.s.rank) ? fa.img_prev_peer[rk] : fa.img_prev;
        const unsigned long long *Pc = img + kConstWords + (size_t)kTileP * fa.tiles_alloc;
        const uint64_t q = Qa - sc.rank_off[rk];
        const uint32_t tiles = fa.tiles_per_rank, stride = (tiles + kThreads - 1) / kThreads;
        uint32_t lt;
        if (TAB) {
            // prefixes in shared memory (at most 4 per thread): every thread counts the hits h among its
            // `stride` consecutive prefixes; they ascend, so the total is the sum over k of #{threads with
            // h >= k}: `stride` barrier counts.  (With the prefixes in L2 this single-round form was slower
            // than the two dependent rounds below: 25.7 vs 24.4 us per C4 step.)
            uint32_t h = 0;
            unsigned long long pv[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint32_t idx = tid * stride + k;
                pv[k] = (k < (int)stride && idx < tiles) ? tab.P[idx] : ~0ull;
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) h += pv[k] <= q;
            uint32_t total = 0;
            for (uint32_t k = 1; k <= stride; ++k) total += (uint32_t)__syncthreads_count(h >= k);
            lt = total - 1u;                                                  // P_0 = 0 always qualifies
        } else {
            uint32_t idx = tid * stride;
            int hit = idx < tiles && ld_word<COH>(Pc + idx) <= q;
            const uint32_t bucket = (uint32_t)__syncthreads_count(hit) - 1u;  // P_0 = 0 always qualifies
            lt = bucket * stride;
            if (stride > 1) {
                idx = lt + tid;
                hit = tid < stride && idx < tiles && ld_word<COH>(Pc + idx) <= q;
                lt += (uint32_t)__syncthreads_count(hit) - 1u;
            }
        }
        tau = rk * fa.tiles_per_rank + lt;
    }
    // (Software-pipelining this walk -- the loads of tile tau + 1 in flight while tile tau is processed --
    // was tried: the speculative tile per block and the registers it pins cost more than the shorter
    // dependency chain gains: 241 -> 272 us per C5 step, 27.3 -> 30.1 us per C4 step.)
    for (;; ++tau) {
        uint32_t lt = tau;
        const unsigned long long *img = fa.img_prev;
        rk = 0;
        if (PEERS) {
            rk = tau / fa.tiles_per_rank;
            lt = tau - rk * fa.tiles_per_rank;
            if (rk != (uint32_t)fa.s.rank) img = fa.img_prev_peer[rk];
        }
        const unsigned long long *fld = img + kConstWords + lt;
        // the tile's three fields AND this thread's CDF words go out together: the words' address does not
        // depend on the fields, and a tile on the walk is (almost) always needed -- one round trip per tile
        // instead of two dependent ones
        const unsigned long long *cp = img + fa.img_hdr_words + (size_t)lt * kTile + kItems * tid;
        unsigned long long c[kItems];
        static_assert(kItems == 8, "two 256-bit loads per thread");
        if (COH) {
            ldcg256u(cp, c[0], c[1], c[2], c[3]);
            ldcg256u(cp + 4, c[4], c[5], c[6], c[7]);
        } else {
            ldg256u(cp, c[0], c[1], c[2], c[3]);
            ldg256u(cp + 4, c[4], c[5], c[6], c[7]);
        }
        const unsigned long long c_left = (lane == 0 && tid != 0) ? ld_word<COH>(cp - 1) : 0ull;
        const uint64_t F = TAB ? tab.F[lt] : ld_word<COH>(fld + (size_t)kTileF * fa.tiles_alloc),
                       P = TAB ? tab.P[lt] : sc.rank_off[rk] + ld_word<COH>(fld + (size_t)kTileP * fa.tiles_alloc),
                       Sp = TAB ? tab.Sp[lt] : ld_word<COH>(fld + (size_t)kTileSp * fa.tiles_alloc);
        const uint64_t Chi = P + Sp;                           // CDF at the end of this parent tile
        if (Sp != 0 && Chi > Qa) {
            const uint64_t C_last = P + cusmc_mulshift62(c[kItems - 1], F);
            uint64_t C_prev = __shfl_up_sync(0xffffffffu, C_last, 1);
            if (lane == 0) C_prev = tid == 0 ? P : P + cusmc_mulshift62(c_left, F);
            // my parents matter iff their CDF span (C_prev, C_last] is non-empty, starts at or before the
            // last child and ends past the first
            if (C_last != C_prev && C_prev <= Qb && C_last > Qa) {
                const uint32_t parent0 = tau * fa.tile_n + kItems * tid;
                uint32_t k_prev = kof(C_prev);
                uint64_t Cp = C_prev;
#pragma unroll
                for (int r = 0; r < kItems; ++r) {
                    const uint64_t C = r == kItems - 1 ? C_last : P + cusmc_mulshift62(c[r], F);
                    if (C != Cp) {                              // zero weight: no children, same count
                        const uint32_t k = kof(C);
                        if (k > k_prev && k > i_a && k_prev < i_b) s_anc[pad((int)(max(k_prev, i_a) - i_a))] = parent0 + r;
                        k_prev = k;
                        Cp = C;
                    }
                }
            }
        }
        if (Chi > Qb) break;                                    // the tile reaches past the last child
    }
    __syncthreads();
    // running maximum over the slots: a marker holds until the next one
    uint32_t v[kItems], run = 0;
#pragma unroll
    for (int r = 0; r < kItems; ++r) {
        run = max(run, s_anc[pad(kItems * (int)tid + r)]);
        v[r] = run;
    }
    uint32_t inc = run;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc = max(inc, t);
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    uint32_t before = __shfl_up_sync(0xffffffffu, inc, 1);
    if (lane == 0) before = 0;
#pragma unroll
    for (int k = 0; k < kThreads / 32; ++k)
        if (k < (int)warp) before = max(before, s_warp[k]);
#pragma unroll
    for (int r = 0; r < kItems; ++r) s_anc[pad(kItems * (int)tid + r)] = max(before, v[r]);
}

#ifndef CUSMC_FUSED_MINB8
#define CUSMC_FUSED_MINB8 5
#endif
#ifndef CUSMC_DENSE_SMOP
#define CUSMC_DENSE_SMOP 1
#endif
#ifndef CUSMC_DENSE_MINB8
#define CUSMC_DENSE_MINB8 3
#endif
// dense operators up to d = 16 are staged in shared memory (3 d^2 doubles: 6 KB at d = 16)
__host__ __device__ constexpr bool dense_smop(int D, bool diag) { return !diag && D <= 16 && CUSMC_DENSE_SMOP != 0; }
__host__ __device__ constexpr int min_blocks(int D, bool diag, bool mvt)
{
    return D >= 32 ? (diag ? 2 : 1) : (D >= 16 ? 2 : (D >= 8 ? (diag && !mvt ? CUSMC_FUSED_MINB8 : (diag ? 3 : CUSMC_DENSE_MINB8)) : 4));
}

// Phase trace (profiling builds only, -DCUSMC_TRACE): block 0 / thread 0 stamps the global timer.
#ifdef CUSMC_TRACE
__device__ __forceinline__ void trace_stamp(double *buf, int slot)
{
    if (buf && (blockIdx.x == 0 || buf[-1] == 1.0) && threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        buf[slot] = (double)t;
    }
}
#define CUSMC_STAMP(buf, slot) pffused::trace_stamp(buf, slot)
#else
#define CUSMC_STAMP(buf, slot) ((void)0)
#endif

struct FusedSmem {
    double lw[kPadded];
    uint32_t anc[kPadded];
    unsigned long long u64[2 * (kThreads / 32)];
    double dbl[kThreads / 32];
    StepConsts c;
    uint32_t warp[kThreads / 32];
};

// One tile of children through the three phases.  `tile` = the block's tile on this rank.
// LEAN (the persistent kernel's main loop: systematic lookup every step, Normal log-weights, no history,
// no accumulation): the run-time switches of the general step are compiled out.  TAB (with LEAN): the
// step constants are already in sm.c and the parents' tile fields in shared memory (`tab`).
// SMOP (dense operators, per-step kernel): the matrices are read from the shared-memory copy `smop`
// (pf_particle.cuh).
template <int D, bool PHILOX, bool FAST, bool MVT, bool EXACT, bool DIAG, bool PEERS, bool COH, bool LEAN = false,
          bool TAB = false, bool SMOP = false>
__device__ __forceinline__ void fused_block_step(const StepOp<D, DIAG> &op, const double (&cobs)[D], const Epilogue &ep,
                                                 const FusedArgs &fa, uint32_t tile, FusedSmem &sm,
                                                 const float *z_ready = nullptr, const TileTab &tab = TileTab{},
                                                 const double *smop = nullptr)
{
    static_assert(!TAB || LEAN, "tile tables: the persistent main loop");
    const StepArgs &a = fa.s;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t j0 = tile * fa.tile_n;                             // local column of the tile's first child
    const uint32_t n_tile = min(fa.tile_n, (uint32_t)a.n_out - j0);
    const uint32_t i_a = (uint32_t)a.i0 + j0;                         // its global slot

    // ---- 1. parents ---------------------------------------------------------------------------------
    const bool lookup = LEAN || fa.mode == kParentLookup;
    bool resample = true;
    if (lookup) {
        if constexpr (!TAB) {
            if (tid < kConstWords) reinterpret_cast<unsigned long long *>(&sm.c)[tid] = ld_word<COH>(fa.img_prev + tid);
            __syncthreads();
        }
        resample = sm.c.resample != 0 && sm.c.T != 0;                 // no mass: identity (flagged by the update)
        if (resample) lookup_parents<PEERS, COH, TAB>(fa, sm.c, i_a, n_tile, sm.anc, sm.u64, sm.warp, tab);
        __syncthreads();
    }
    CUSMC_STAMP(fa.trace, 1);
    const bool accumulate = !LEAN && fa.accumulate && lookup && sm.c.resample == 0;

    // ---- 2. propagate + reweight, striped ------------------------------------------------------------
    // one child: everything after the parent's state is known
    auto child = [&](uint32_t j, uint32_t parent, const double *src, const double *xp_in) {
        const int64_t i = (int64_t)j0 + j;
        cusmc_u32x4 r0{};
        const float *zin = z_ready ? z_ready + (size_t)j * D : nullptr;       // drawn in the barrier shadow
        if (PHILOX && !zin) r0 = pfstep::first_block<FAST, D>(a.seed, a.rng_stream, a.step, (uint64_t)(a.i0 + i));
        double lw = pfstep::particle_step<D, PHILOX, FAST, MVT, EXACT, DIAG, COH, LEAN, SMOP>(op, cobs, ep, a, i, src, r0, xp_in,
                                                                                              zin, smop);
        if (accumulate) lw = a.lw[i] + lw;                        // no resampling: the log-weights accumulate
        if (a.lw) st_stream(a.lw + i, lw);
        if (!LEAN && a.hist_w) st_stream(a.hist_w + i, lw);
        if (!LEAN && a.hist_a) a.hist_a[i] = parent;
        if (fa.anc_out) fa.anc_out[i] = parent;
        return lw;
    };
    auto parent_of = [&](uint32_t j) {
        uint32_t parent = i_a + j;                                // global id
        if (lookup) {
            if (resample) parent = sm.anc[pad((int)j)];
        } else if (fa.mode == kParentArray) {
            parent = __ldg(a.anc + (int64_t)j0 + j);
        }
        return parent;
    };
    // Sharded: a tile whose parents all live on this rank (ancestors are monotone: look at the first and
    // the last) skips the per-child slot -> rank division -- almost every tile of a shard.
    bool all_local = !PEERS;
    const int64_t local_base = PEERS ? (int64_t)a.rank * (int64_t)a.per_rank.d : a.parent_base;
    if (PEERS && lookup) {
        all_local = true;
        if (resample) {
            const uint32_t lo = (uint32_t)local_base, first = sm.anc[pad(0)], last = sm.anc[pad((int)n_tile - 1)];
            all_local = first >= lo && last - lo < a.per_rank.d;
        }
    }
    auto column_of = [&](uint32_t parent) {
        const double *src = a.x_prev + ((int64_t)parent - local_base);
        if (PEERS && a.has_prev && !all_local) {
            const uint32_t rk = fast_div(parent, a.per_rank);
            const uint32_t col = parent - rk * a.per_rank.d;
            src = (rk == (uint32_t)a.rank ? a.x_prev : a.x_prev_peer[rk]) + col;
        }
        return src;
    };
    if constexpr (COH && D <= 4) {
        // persistent kernel: the cloud lives in L2 and a round is one dependent gather -- bound by its
        // latency, not by bandwidth.  A batch of rounds issues all its gathers first, then computes.
        constexpr int kBatch = 4;
#pragma unroll 1
        for (int rb = 0; rb < kItems; rb += kBatch) {
            double xpre[kBatch][D];
            uint32_t par[kBatch];
#pragma unroll
            for (int b = 0; b < kBatch; ++b) {
                const uint32_t j = (uint32_t)(rb + b) * kThreads + tid;
                par[b] = j < n_tile ? parent_of(j) : 0u;
                const double *src = column_of(par[b]);
#pragma unroll
                for (int k = 0; k < D; ++k)
                    xpre[b][k] = (j < n_tile && (LEAN || a.has_prev)) ? __ldcg(src + (int64_t)k * a.ld_prev) : 0.0;
            }
#pragma unroll
            for (int b = 0; b < kBatch; ++b) {
                const uint32_t j = (uint32_t)(rb + b) * kThreads + tid;
                double lw = -INFINITY;
                if (j < n_tile) lw = child(j, par[b], nullptr, xpre[b]);
                sm.lw[pad((int)j)] = lw;
            }
        }
    } else {
        // Dense operators (SMOP): 3 blocks per SM.  What was tried against the latency of the parent gather
        // there (17 % of the stall samples on its first use): a register-held prefetch of the next round
        // (spills: 358 -> 401 us per dense C5 step), prefetch.global.L1 / .L2 hints (362 us), cp.async into a
        // double-buffered per-thread slot of shared memory (long-scoreboard stalls 2.2 -> 0.8 per issue, but
        // the LDGSTS traffic delays the operator LDS: short-scoreboard 1.7 -> 2.7, kernel time unchanged).
#pragma unroll 1
        for (int r = 0; r < kItems; ++r) {
            const uint32_t j = (uint32_t)r * kThreads + tid;
            double lw = -INFINITY;
            if (j < n_tile) {
                const uint32_t parent = parent_of(j);
                lw = child(j, parent, column_of(parent), nullptr);
            }
            sm.lw[pad((int)j)] = lw;
        }
    }
    __syncthreads();
    CUSMC_STAMP(fa.trace, 2);

    // ---- 3. weigh: block-relative fixed-point image of the tile ----------------------------------------
    double v[kItems], m = -INFINITY;
#pragma unroll
    for (int r = 0; r < kItems; ++r) {
        v[r] = sm.lw[pad(kItems * (int)tid + r)];
        if (v[r] == v[r] && v[r] < INFINITY && v[r] > m) m = v[r];
    }
    m = warp_max_double(m);
    if (lane == 0) sm.dbl[warp] = m;
    __syncthreads();
    m = warp_max_double(lane < kThreads / 32 ? sm.dbl[lane] : -INFINITY);          // the tile's maximum, every thread
    // (Letting the persistent kernel's all-padding warps -- its tiles are shorter than kTile -- skip the
    // exponentials was tried: 20.2 -> 21.2 us per C4 step.)
    unsigned long long c[kItems], run = 0, s2 = 0;
#pragma unroll
    for (int r = 0; r < kItems; ++r) {
        unsigned long long q, q2;
        weigh_fixed(v[r], m, fa.shift, q, q2);                                     // -inf padding -> 0
        run += q;
        c[r] = run;
        s2 += q2;
    }
    unsigned long long inc = run;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    if (lane == 31) sm.u64[warp] = inc;
    if (lane == 0) sm.u64[kThreads / 32 + warp] = s2;
    __syncthreads();
    unsigned long long before = inc - run, tile_total = 0, s2_total = 0;
#pragma unroll
    for (int k = 0; k < kThreads / 32; ++k) {
        const unsigned long long t = sm.u64[k];
        if (k < (int)warp) before += t;
        tile_total += t;
        s2_total += sm.u64[kThreads / 32 + k];
    }
    unsigned long long *local = fa.img_new + fa.img_hdr_words + (size_t)tile * kTile + kItems * tid;   // 32-byte aligned
    stg256u(local, before + c[0], before + c[1], before + c[2], before + c[3]);
    stg256u(local + 4, before + c[4], before + c[5], before + c[6], before + c[7]);
    if (tid < 3) {
        unsigned long long *fld = fa.img_new + kConstWords + tile;
        fld[(size_t)tid * fa.tiles_alloc] = tid == 0 ? (unsigned long long)__double_as_longlong(m) : tid == 1 ? tile_total : s2_total;
    }
}

template <int D, bool PHILOX, bool FAST, bool MVT, bool EXACT, bool DIAG, bool PEERS>
__global__ void __launch_bounds__(kThreads, min_blocks(D, DIAG, MVT))
pf_fused_kernel(const __grid_constant__ StepOp<D, DIAG> op, const Epilogue ep, const FusedArgs fa)
{
    __shared__ FusedSmem sm;
    constexpr bool SMOP = dense_smop(D, DIAG);
    __shared__ alignas(16) double s_op[SMOP ? 3 * D * D : 2];
    if constexpr (SMOP) {
        // the operators are launch parameters: staged before the wait on the previous grid
        for (int e = threadIdx.x; e < D * D; e += kThreads) {
            s_op[e] = op.G[DIAG ? 0 : e];
            s_op[D * D + e] = op.Q[DIAG ? 0 : e];
            s_op[2 * D * D + e] = op.M[DIAG ? 0 : e];
        }
        __syncthreads();
    }
    // Programmatic dependent launch: this grid is launched while the tile update of the previous step is
    // still running (its blocks become resident, its launch latency is hidden) and waits HERE until that
    // grid has completed and its writes are visible.  A no-op when launched the ordinary way.
    asm volatile("griddepcontrol.wait;" ::: "memory");
    // ... and the tile update that follows THIS grid may be made resident as soon as every block here has
    // started (it waits the same way), which hides its launch latency behind the last wave.
    asm volatile("griddepcontrol.launch_dependents;");
    fused_block_step<D, PHILOX, FAST, MVT, EXACT, DIAG, PEERS, false, false, false, SMOP>(op, op.c, ep, fa, blockIdx.x, sm, nullptr,
                                                                                          TileTab{}, s_op);
}

// ---- launch ---------------------------------------------------------------------------------------------
template <int D, bool MVT, bool EXACT, bool DIAG>
int launch_one(cusmc_ctx *ctx, const pfstep::StepModel &m, const Epilogue &ep, const FusedArgs &fa, bool philox)
{
    StepOp<D, DIAG> op;
    pfstep::fill_step_op<D, DIAG>(op, m);
    const unsigned grid = (unsigned)((fa.s.n_out + fa.tile_n - 1) / fa.tile_n);
    const bool peers = fa.s.world > 1;
    // launched as a programmatic dependent of the kernel before it on the stream (the tile update, which
    // releases its dependents as soon as it starts): see pf_fused_kernel
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = fa.pdl ? 1 : 0;
    cudaLaunchConfig_t lc{};
    lc.gridDim = dim3(grid);
    lc.blockDim = dim3(kThreads);
    lc.dynamicSmemBytes = 0;
    lc.stream = ctx->stream;
    lc.attrs = attr;
    lc.numAttrs = 1;
#define CUSMC_FUSED_GO(PH, FA, PE) \
    CUSMC_CUDA(ctx, cudaLaunchKernelEx(&lc, pf_fused_kernel<D, PH, FA, MVT, EXACT, DIAG, PE>, op, ep, fa))
    if (philox && fa.s.fast_noise) {
        if (peers) CUSMC_FUSED_GO(true, true, true); else CUSMC_FUSED_GO(true, true, false);
    } else if (philox) {
        if (peers) CUSMC_FUSED_GO(true, false, true); else CUSMC_FUSED_GO(true, false, false);
    } else {
        if (peers) CUSMC_FUSED_GO(false, false, true); else CUSMC_FUSED_GO(false, false, false);
    }
#undef CUSMC_FUSED_GO
    CUSMC_LAUNCHED(ctx);
    return CUSMC_OK;
}

// One translation unit per (D, noise family) instantiates its three kernels' worth of launch_one
// (pf_fused_inst.cu, compiled once per pair by the Makefile so the build parallelises); everybody else
// sees extern templates.
#define CUSMC_FUSED_VARIANTS(X, DD, MV)                  \
    X int launch_one<DD, MV, true, true>(cusmc_ctx *, const pfstep::StepModel &, const Epilogue &, const FusedArgs &, bool);  \
    X int launch_one<DD, MV, true, false>(cusmc_ctx *, const pfstep::StepModel &, const Epilogue &, const FusedArgs &, bool); \
    X int launch_one<DD, MV, false, false>(cusmc_ctx *, const pfstep::StepModel &, const Epilogue &, const FusedArgs &, bool);
#ifndef CUSMC_INST_D
#define CUSMC_FUSED_EXTERN(DD) CUSMC_FUSED_VARIANTS(extern template, DD, false) CUSMC_FUSED_VARIANTS(extern template, DD, true)
CUSMC_FUSED_EXTERN(2)
CUSMC_FUSED_EXTERN(4)
CUSMC_FUSED_EXTERN(8)
CUSMC_FUSED_EXTERN(16)
CUSMC_FUSED_EXTERN(32)
#undef CUSMC_FUSED_EXTERN
#endif

template <bool MVT>
int launch_family(cusmc_ctx *ctx, const pfstep::StepModel &m, const Epilogue &ep, const FusedArgs &fa, bool philox,
                  bool exact, bool diag)
{
    const int dm = m.d > m.dy ? m.d : m.dy;
#define CUSMC_FUSED_CASE(DD)                                                                         \
    case DD:                                                                                         \
        if (exact && diag) return launch_one<DD, MVT, true, true>(ctx, m, ep, fa, philox);           \
        if (exact) return launch_one<DD, MVT, true, false>(ctx, m, ep, fa, philox);                  \
        return launch_one<DD, MVT, false, false>(ctx, m, ep, fa, philox);
    switch (cusmc_pad_dim(dm)) {
        CUSMC_FUSED_CASE(2)
        CUSMC_FUSED_CASE(4)
        CUSMC_FUSED_CASE(8)
        CUSMC_FUSED_CASE(16)
        CUSMC_FUSED_CASE(32)
    }
#undef CUSMC_FUSED_CASE
    return cusmc_fail(ctx, CUSMC_ERR_UNSUPPORTED, "unreachable");
}

}  // namespace pffused

// G, Q column-major d x d host (either may be NULL = zero); M row-major dy x d (NULL = no weights);
// c dy; mu d (NULL = 0).  Picks the EXACT / DIAG specialisations itself (as cusmc_launch_step).
int cusmc_launch_fused(cusmc_ctx *ctx, int d, int dy, const double *G, const double *Q, double qscale,
                       const std::vector<double> *M, const double *c, const double *mu, const Epilogue &ep,
                       const pffused::FusedArgs &fa, bool philox);

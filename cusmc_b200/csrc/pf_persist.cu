// pf_persist.cu -- the whole filter run as ONE cooperative kernel, for particle clouds that live in L2.
//
// At 10^6 particles (BASELINE configs[3]: the bootstrap filter on data_raw/y_t.csv) a step touches
// ~40 MB that never leave the 126 MB L2, and a step made of separate launches is bound by launch
// latency and by its grid-wide dependencies.  Since the fused step (pf_fused_impl.cuh) needs exactly
// ONE grid-wide dependency per step -- the tile update between two steps -- the whole run is a loop
//
//     for t:   fused_block_step(t)          every block: lookup + propagate + reweight + weigh of its tile
//              grid barrier                 the LAST block to arrive runs tile_update_block(t) before
//                                           it releases the others
//
// inside one cooperative launch: one barrier per step (round 1's persistent kernel needed three: max,
// total, scatter), no launch gaps, and the cloud is spread EVENLY over all resident blocks (tile_n
// particles per block instead of a fixed 2048).
//
// CUSMC_PERSIST_SELFUPD (default): nobody runs the update FOR the others.  Once every block has arrived,
// EVERY block runs the (tiny: <= 768 tiles) update itself, redundantly, from the tile fields in L2 into its
// own shared memory: the serial update + release hop behind the barrier disappears, and the next step's
// lookup finds the step constants, the tile prefix array and the fields of every parent tile in shared
// memory instead of two dependent L2 round trips.  Same integers in every block (the update is integer /
// IEEE arithmetic in a fixed order); ONE block per step (the first to arrive -- the one with the most
// slack) also writes the step record, the image header and the tile fields to global memory, so the
// filter's state after the run is what the per-step path leaves.  The code is the per-step path's, instantiated with
// L2-coherent loads for everything another block wrote earlier in the same launch.
//
// Same arithmetic as the per-step path operation for operation (same device functions, same Philox
// counters) with tile = tile_n: the oracle reproduces a persistent run bit for bit when told that tile
// size (orc_filter_det's `tile`; tests/test_gpu_parity.py).
// Same semantics as cusmc_filter_run (src/mcmc.cpp:239-309 of the reference) restricted to: one GPU,
// systematic resampling every step, Normal noise, d == dy in {2, 4, 8}, device-drawn noise, no history,
// no per-step means, N small enough for one tile per resident block.
#include "filter_types.cuh"
#include "pf_fused_impl.cuh"
#include "tile_update_impl.cuh"

#include <algorithm>
#include <cstdlib>

namespace {

using pffused::FusedArgs;
using pffused::FusedSmem;
using pffused::kThreads;
using pfstep::StepOp;

struct PersistArgs {
    FusedArgs fa;                   // the per-step arguments that do not change (sizes, keys, layout)
    double *x[2];                   // SoA [d][ld] double buffer
    unsigned long long *img[2];     // weight images by step parity
    double *lw;                     // [N] final log-weights (cusmc_filter_state_dev)
    uint32_t *anc;                  // [N] ancestors of the final step
    StepSlot *slots;                // [T]
    const double *obs;              // [T][D]: L_V^-1 y_t
    const double *u0;               // [T]: systematic offsets (entry t used by step t)
    unsigned *barrier;              // [0] arrival counter, [1] release generation; zero at launch
    int64_t tiles_alloc;
    int T;
    double *trace;                  // profiling builds (-DCUSMC_TRACE): [T][8] time stamps of block 0
};

// Grid barrier whose last arriver runs the tile update of step t before releasing the others.
// Monotonic arrival counter + release generation (the kernel is launched cooperatively, so every block
// is resident); everything that crosses it is read through L2 (ld.global.cg / ld.acquire).
// While a block waits -- between its arrival and the release -- it draws the NEXT step's normals for its
// own tile into shared memory (`shadow`, called by every block but the last to arrive, which has the
// update to run and draws its normals inside its rounds instead).  Returns whether the shadow ran.
template <typename Shadow>
__device__ __forceinline__ bool barrier_with_update(const PersistArgs &pa, int t, unsigned &gen, int *s_last,
                                                    UpdateSmem<kThreads> &us, Shadow shadow)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        // one acq_rel atomic: releases this block's tile (the bar.sync above makes the release cumulative
        // over the whole block's writes) and, for the last arrival, acquires everybody else's
        unsigned old;
        asm volatile("atom.add.acq_rel.gpu.global.u32 %0, [%1], 1;" : "=r"(old) : "l"(pa.barrier) : "memory");
        *s_last = old == (gen + 1u) * gridDim.x - 1u;
#ifdef CUSMC_TRACE
        if (pa.trace && t > 0) {
            unsigned long long now;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            if (blockIdx.x == 0) pa.trace[1 + (size_t)t * 9 + 7] = (double)now;       // block 0 has arrived
            if (*s_last) pa.trace[1 + (size_t)t * 9 + 5] = (double)now;               // the last arrival
            if (t == 50) pa.trace[1 + ((size_t)pa.T + blockIdx.x) * 9 + 7] = (double)now;
        }
#endif
    }
    __syncthreads();
    if (*s_last) {
        UpdateArgs u{};
        u.img = pa.img[t & 1];
        u.slot = pa.slots + t;
        u.slot_next = t + 1 < pa.T ? pa.slots + t + 1 : nullptr;
        u.tiles = gridDim.x;
        u.tiles_alloc = pa.tiles_alloc;
        u.N_global = pa.fa.N_global;
        u.u0_next = t + 1 < pa.T ? __ldg(pa.u0 + t + 1) : 0.0;
        u.world = 1;
        u.phases = kUpdAll;
        tile_update_block<kThreads>(u, us);
        __syncthreads();
        if (threadIdx.x == 0) {
#ifdef CUSMC_TRACE
            if (pa.trace && t > 0) {
                unsigned long long now;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
                pa.trace[1 + (size_t)t * 9 + 6] = (double)now;                          // update done
            }
#endif
            asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(pa.barrier + 1), "r"(gen + 1u) : "memory");
        }
    } else {
        shadow();
        if (threadIdx.x == 0) {
            unsigned seen;
            do {
                asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(pa.barrier + 1) : "memory");
            } while ((int)(seen - (gen + 1u)) < 0);
        }
    }
    const bool ran = !*s_last;
    __syncthreads();
    ++gen;
    return ran;
}

// ---- the self-served update -------------------------------------------------------------------------
#ifndef CUSMC_PERSIST_SELFUPD
#define CUSMC_PERSIST_SELFUPD 1
#endif
#ifndef CUSMC_PERSIST_LEAN
#define CUSMC_PERSIST_LEAN 1
#endif
// the three tile tables alias FusedSmem::lw (dead between the weigh phase and the next step's rounds)
constexpr int kSelfMaxTiles = 768;
static_assert(3 * kSelfMaxTiles * sizeof(unsigned long long) <= sizeof(FusedSmem::lw), "tile tables alias the log-weight tile");
static_assert(kSelfMaxTiles <= kThreads * kUpdItems, "one chunk: a thread owns four consecutive tiles");

// tile_update_block (tile_update_impl.cuh) for one rank, no adaptive decision, at most kThreads * 4 tiles,
// results into shared memory: tables F / P / Sp and the step constants sm.c.  `writer`: this block also
// computes the sum of squares and writes what tile_update_block writes to global memory.
// It also leaves what the next step's lookup starts from: the marker table zeroed and the mass quotients of
// the block's first and last child in sm.u64[0..1] (lookup_parents with TAB).  The serial tail is spread
// over four threads (constants and Q_a | Q_b | the two quotients N / T, r0 / T | the step record), and
// u0_next was fetched before the barrier wait.
__device__ __forceinline__ void self_update(const PersistArgs &pa, int t, bool writer, FusedSmem &sm,
                                            UpdateSmem<kThreads> &us, unsigned long long *tabF,
                                            unsigned long long *tabP, unsigned long long *tabSp, double u0_next,
                                            uint32_t i_a, uint32_t n_tile)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t tiles = gridDim.x, tiles_all = pa.tiles_alloc;
    unsigned long long *img = pa.img[t & 1];
    unsigned long long *fld[kTileFields];
#pragma unroll
    for (int f = 0; f < kTileFields; ++f) fld[f] = img + kConstWords + (int64_t)f * tiles_all;
    const int64_t b0 = (int64_t)tid * kUpdItems;
    unsigned long long mv[kUpdItems] = {0, 0, 0, 0}, S[kUpdItems] = {0, 0, 0, 0}, S2[kUpdItems] = {0, 0, 0, 0};
    if (b0 < tiles) {
        ld4(fld[kTileM] + b0, mv);
        ld4(fld[kTileS] + b0, S);
        if (writer) ld4(fld[kTileS2] + b0, S2);
    }
#pragma unroll
    for (int r = 0; r < pffused::kItems; ++r) sm.anc[pffused::pad(r * kThreads + tid)] = 0u;      // the lookup's marker table
    double m = -INFINITY;
#pragma unroll
    for (int r = 0; r < kUpdItems; ++r) {
        const double v = __longlong_as_double((long long)mv[r]);
        if (b0 + r < tiles && v > m) m = v;                        // records hold finite values or -inf
    }
    m = warp_max_double(m);
    if (lane == 0) us.dbl[warp] = m;
    __syncthreads();
    const double M = warp_max_double(lane < kThreads / 32 ? us.dbl[lane] : -INFINITY);
    unsigned long long F[kUpdItems] = {0, 0, 0, 0}, sp[kUpdItems] = {0, 0, 0, 0}, run = 0, t2 = 0, inc = 0;
    const bool warp_live = (int64_t)(warp * 32) * kUpdItems < tiles;       // (a warp past the last tile has nothing to scan)
    if (warp_live) {
#pragma unroll
        for (int r = 0; r < kUpdItems; ++r)
            if (b0 + r < tiles) {
                F[r] = cusmc_rescale_factor(__longlong_as_double((long long)mv[r]), M);
                sp[r] = cusmc_mulshift62(S[r], F[r]);
                if (writer) t2 += cusmc_mulshift62(cusmc_mulshift62(S2[r], F[r]), F[r]);
                run += sp[r];
            }
        inc = run;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long v = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += v;
        }
    }
    if (lane == 31) us.sm[warp] = inc;
    if (writer) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t2 += __shfl_xor_sync(0xffffffffu, t2, o);
        if (lane == 0) us.sm2[warp] = t2;
    }
    __syncthreads();
    unsigned long long excl = inc - run, T = 0, T2 = 0;
#pragma unroll
    for (int k = 0; k < kThreads / 32; ++k) {
        const unsigned long long w = us.sm[k];
        if (k < warp) excl += w;
        T += w;
        if (writer) T2 += us.sm2[k];
    }
    unsigned long long P[kUpdItems];
#pragma unroll
    for (int r = 0; r < kUpdItems; ++r) {
        P[r] = excl;
        excl += sp[r];
    }
    if (b0 < tiles_all) {                                          // tiles past the end of the cloud are empty
#pragma unroll
        for (int r = 0; r < kUpdItems; ++r) {
            tabF[b0 + r] = F[r];
            tabP[b0 + r] = P[r];
            tabSp[b0 + r] = sp[r];
        }
        if (writer) {
            st4(fld[kTileF] + b0, F);
            st4(fld[kTileP] + b0, P);
            st4(fld[kTileSp] + b0, sp);
        }
    }
    if ((tid & 31) == 0 && tid < 128) {
        unsigned long long rr = (unsigned long long)(u0_next * (double)T);
        if (T && rr > T - 1) rr = T - 1;
        StepConsts &c = sm.c;
        const unsigned long long Ng = pa.fa.N_global;
        if (tid == 0) {
            c.T = T;
            c.r0 = rr;
            c.resample = 1;
            c.T2 = T2;
            c.M = M;
            c.reserved = 0;
            c.rank_off[0] = 0;
#pragma unroll
            for (int k = 1; k < CUSMC_MAX_PEERS; ++k) c.rank_off[k] = T;
            sm.u64[0] = mass_quotient(i_a, T, rr, Ng);
        } else if (tid == 32) {
            sm.u64[1] = mass_quotient(i_a + n_tile - 1, T, rr, Ng);
        } else if (tid == 64) {
            c.ng_over_t = (double)pa.fa.N_global / (double)T;
            c.r0_over_t = (double)rr / (double)T;
        } else if (writer) {
            StepSlot *slot = pa.slots + t;
            slot->lw_max = M;
            slot->sum_q = T;
            slot->sum_q2 = T2;
            slot->n_pos = 0;
            slot->cdf_offset = 0;
            if (t + 1 < pa.T) {
                pa.slots[t + 1].resampled = 1;
                pa.slots[t + 1].degenerate = T == 0 ? 1 : 0;
            }
        }
    }
    __syncthreads();
    if (writer && tid < kConstWords) img[tid] = reinterpret_cast<const unsigned long long *>(&sm.c)[tid];
}

// The grid barrier of the self-served update: arrive, draw the next step's normals (`shadow`), wait until
// everybody has arrived, update.
template <typename Shadow>
__device__ __forceinline__ void barrier_self_update(const PersistArgs &pa, int t, unsigned &gen, int *s_first, FusedSmem &sm,
                                                    UpdateSmem<kThreads> &us, unsigned long long *tabF,
                                                    unsigned long long *tabP, unsigned long long *tabSp, double *trace,
                                                    Shadow shadow)
{
    const double u0_next = t + 1 < pa.T ? __ldg(pa.u0 + t + 1) : 0.0;      // in flight while the block waits
    __syncthreads();
    if (threadIdx.x == 0) {
        // releases this block's tile (the bar.sync above makes the release cumulative over the block's writes)
        unsigned old;
        asm volatile("atom.add.release.gpu.global.u32 %0, [%1], 1;" : "=r"(old) : "l"(pa.barrier) : "memory");
        *s_first = old == gen * gridDim.x;
    }
    CUSMC_STAMP(trace, 7);                                         // arrived
    shadow();
    if (threadIdx.x == 0) {
        const unsigned want = (gen + 1u) * gridDim.x;
        unsigned seen;
        // (relaxed polls + one acquire fence at the end measured no faster: 20.4 vs 20.2 us per C4 step)
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(pa.barrier) : "memory");
        } while ((int)(seen - want) < 0);
    }
    CUSMC_STAMP(trace, 5);                                         // everybody has arrived
    __syncthreads();
    const uint32_t tile_lo = blockIdx.x * pa.fa.tile_n;
    self_update(pa, t, *s_first != 0, sm, us, tabF, tabP, tabSp, u0_next, (uint32_t)pa.fa.s.i0 + tile_lo,
                min(pa.fa.tile_n, (uint32_t)pa.fa.s.n_out - tile_lo));
    CUSMC_STAMP(trace, 6);                                         // update done
    ++gen;
}

#ifndef CUSMC_PERSIST_MINB
#define CUSMC_PERSIST_MINB 4
#endif
constexpr int persist_min_blocks(int D, bool diag) { return D >= 8 ? (diag ? 4 : 2) : CUSMC_PERSIST_MINB; }

template <int D, bool FAST, bool DIAG>
__global__ void __launch_bounds__(kThreads, persist_min_blocks(D, DIAG))
pf_persistent_kernel(const __grid_constant__ StepOp<D, DIAG> op_init, const __grid_constant__ StepOp<D, DIAG> op,
                     const Epilogue ep, const __grid_constant__ PersistArgs pa)
{
    __shared__ FusedSmem sm;
    __shared__ UpdateSmem<kThreads> us;
    __shared__ int s_last;
    // normals of the next step, drawn in the barrier shadow (d = 2: 16 KB; wider states would cost a
    // resident block per SM, they draw inside their rounds)
    constexpr bool kShadowNoise = D == 2;
    __shared__ __align__(16) float s_z[kShadowNoise ? kTile * D : 1];
    unsigned gen = 0;
    FusedArgs fa = pa.fa;
    const uint32_t tile_lo = blockIdx.x * fa.tile_n;
    const uint32_t tile_cnt = min(fa.tile_n, (uint32_t)fa.s.n_out - tile_lo);
    // the normals the step kernel would draw for (step, slot): same counters, same transform
    auto draw_next = [&](int t_next) {
        if constexpr (kShadowNoise) {
            if (t_next >= pa.T) return;
            if constexpr (pfstep::PairBlocks<FAST, D>::value) {
                // one block per PAIR of neighbours (i0 = 0 and tile_lo is a multiple of 32: pairs never straddle tiles)
                for (uint32_t p = threadIdx.x; 2 * p < tile_cnt; p += kThreads) {
                    const cusmc_u32x4 r = pfstep::first_block<FAST, D>(fa.s.seed, CUSMC_STREAM_NORMAL, (uint64_t)t_next,
                                                                       (uint64_t)(fa.s.i0 + tile_lo + 2 * p));
                    float zq[4];
                    pfstep::normals4<FAST>(r, zq);
                    *reinterpret_cast<float4 *>(s_z + 4 * (size_t)p) = make_float4(zq[0], zq[1], zq[2], zq[3]);
                }
                return;
            }
            for (uint32_t j = threadIdx.x; j < tile_cnt; j += kThreads) {
                const cusmc_u32x4 r = pfstep::step_rng<FAST>(fa.s.seed, CUSMC_STREAM_NORMAL, (uint64_t)t_next,
                                                             (uint64_t)(fa.s.i0 + tile_lo + j), 0u);
                float zq[4];
                pfstep::normals4<FAST>(r, zq);
#pragma unroll
                for (int k = 0; k < D; ++k) s_z[(size_t)j * D + k] = zq[k];
            }
        }
    };

    // ---- t = 0: x_0 = m0 + Q_c0 z, constant log-weight 0 (src/mcmc.cpp:63-85) -----------------------
    {
        const double zero_c[D] = {};
        fa.s.x_new = pa.x[0];
        fa.s.x_prev = pa.x[1];
        fa.img_new = pa.img[0];
        fa.img_prev = pa.img[1];
        fa.mode = pffused::kParentSelf;
        fa.s.has_prev = 0;
        fa.s.skip_weight = 1;
        fa.s.const_weight = 0.0;
        fa.s.rng_stream = CUSMC_STREAM_INIT;
        fa.s.step = 0;
        pffused::fused_block_step<D, true, FAST, false, true, DIAG, false, true>(op_init, zero_c, ep, fa, blockIdx.x, sm);
    }
    // dense operators of the main loop: read from shared memory (pf_particle.cuh, SMOP) -- as parameter-bank
    // operands they are hoisted out of the step loop and spilled (272-352 bytes of stack at d = 4, 8)
    constexpr bool kSmop = pffused::dense_smop(D, DIAG);
    __shared__ alignas(16) double s_op[kSmop ? 3 * D * D : 2];
    if constexpr (kSmop) {
        for (int e = threadIdx.x; e < D * D; e += kThreads) {
            s_op[e] = op.G[DIAG ? 0 : e];
            s_op[D * D + e] = op.Q[DIAG ? 0 : e];
            s_op[2 * D * D + e] = op.M[DIAG ? 0 : e];
        }
        __syncthreads();
    }
    constexpr bool kLean = CUSMC_PERSIST_LEAN != 0, kSelf = kLean && CUSMC_PERSIST_SELFUPD != 0;
    unsigned long long *tabF = reinterpret_cast<unsigned long long *>(sm.lw), *tabP = tabF + kSelfMaxTiles,
                       *tabSp = tabP + kSelfMaxTiles;
    const pffused::TileTab tab{tabF, tabP, tabSp};
    bool z_ready;
    if constexpr (kSelf) {
        barrier_self_update(pa, 0, gen, &s_last, sm, us, tabF, tabP, tabSp, nullptr, [&] { draw_next(1); });
        z_ready = kShadowNoise;
    } else {
        z_ready = barrier_with_update(pa, 0, gen, &s_last, us, [&] { draw_next(1); }) && kShadowNoise;
    }

    fa.mode = pffused::kParentLookup;
    fa.s.has_prev = 1;
    fa.s.skip_weight = 0;
    fa.s.rng_stream = CUSMC_STREAM_NORMAL;
#pragma unroll 1
    for (int t = 1; t < pa.T; ++t) {
        double cobs[D];
#pragma unroll
        for (int k = 0; k < D; ++k) cobs[k] = __ldg(pa.obs + (size_t)t * D + k);
        fa.s.x_new = pa.x[t & 1];
        fa.s.x_prev = pa.x[(t & 1) ^ 1];
        fa.img_new = pa.img[t & 1];
        fa.img_prev = pa.img[(t & 1) ^ 1];
        fa.s.step = (uint64_t)t;
        if (t == pa.T - 1) {                      // the final state: log-weights and ancestors for the host
            fa.s.lw = pa.lw;
            fa.anc_out = pa.anc;
        }
#ifdef CUSMC_TRACE
        // block 0 stamps every step; at step 50 EVERY block stamps its own row (after the T rows of block 0;
        // the word before a row says "everybody": rows are 9 words apart)
        fa.trace = pa.trace ? pa.trace + 1 + (size_t)t * 9 : nullptr;
        if (pa.trace && t == 50) fa.trace = pa.trace + 1 + ((size_t)pa.T + blockIdx.x) * 9;
#endif
        CUSMC_STAMP(fa.trace, 0);
        pffused::fused_block_step<D, true, FAST, false, true, DIAG, false, true, kLean, kSelf, kSmop>(
            op, cobs, ep, fa, blockIdx.x, sm, z_ready ? s_z : nullptr, tab, s_op);
        CUSMC_STAMP(fa.trace, 3);
        if constexpr (kSelf)
            barrier_self_update(pa, t, gen, &s_last, sm, us, tabF, tabP, tabSp, fa.trace, [&] { draw_next(t + 1); });
        else
            z_ready = barrier_with_update(pa, t, gen, &s_last, us, [&] { draw_next(t + 1); }) && kShadowNoise;
        CUSMC_STAMP(fa.trace, 4);
    }
}

// Particles per tile: N spread evenly over ALL resident block slots (every SM gets the same work), in
// whole warps; 0 if such a tile exceeds what one block holds.
uint32_t pick_tile(const cusmc_filter *f, int per_sm)
{
    int64_t slots = (int64_t)f->ctx->sm_count * per_sm;
    if (CUSMC_PERSIST_LEAN && CUSMC_PERSIST_SELFUPD) slots = std::min<int64_t>(slots, kSelfMaxTiles);   // tables in shared memory
    int64_t n = (f->cfg.N + slots - 1) / slots;
    n = std::max<int64_t>((n + 31) & ~(int64_t)31, 32);
    return n <= kTile ? (uint32_t)n : 0u;
}

template <int D, bool FAST, bool DIAG>
int launch_persistent(cusmc_filter *f, PersistArgs &pa, bool probe_only, uint32_t *tile_out)
{
    cusmc_ctx *ctx = f->ctx;
    const cusmc_filter_config &cfg = f->cfg;
    auto kernel = pf_persistent_kernel<D, FAST, DIAG>;
    int per_sm = 0;
    CUSMC_CUDA(ctx, cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    CUSMC_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kThreads, 0));
    const uint32_t tile_n = per_sm > 0 ? pick_tile(f, per_sm) : 0u;
    if (!tile_n)
        return cusmc_fail(ctx, CUSMC_ERR_UNSUPPORTED, "persistent run: %lld particles exceed one tile per resident block (%d per SM)",
                          (long long)cfg.N, per_sm);
    *tile_out = tile_n;
    if (probe_only) return CUSMC_OK;
    const unsigned grid = (unsigned)((cfg.N + tile_n - 1) / tile_n);
    if (CUSMC_PERSIST_SELFUPD && pa.tiles_alloc > kSelfMaxTiles)
        return cusmc_fail(ctx, CUSMC_ERR_UNSUPPORTED, "persistent run: %lld tiles exceed the %d the blocks keep in shared memory",
                          (long long)pa.tiles_alloc, kSelfMaxTiles);
    pa.fa.tile_n = tile_n;
    StepOp<D, DIAG> op0, op;
    const pfstep::StepModel m0{cfg.d, cfg.dy, nullptr, f->Qc0.data(), cfg.noise_scale, nullptr, nullptr, f->m0.data()};
    const pfstep::StepModel m1{cfg.d, cfg.dy, f->G.data(), f->Qw.data(), cfg.noise_scale, &f->M, nullptr, nullptr};
    pfstep::fill_step_op<D, DIAG>(op0, m0);
    pfstep::fill_step_op<D, DIAG>(op, m1);
    Epilogue ep = f->ep;
    void *params[] = {&op0, &op, &ep, &pa};
    const cudaError_t e = cudaLaunchCooperativeKernel((const void *)kernel, dim3(grid), dim3(kThreads), params, 0, ctx->stream);
    if (e == cudaErrorCooperativeLaunchTooLarge || e == cudaErrorNotSupported) {
        cudaGetLastError();
        return cusmc_fail(ctx, CUSMC_ERR_UNSUPPORTED, "cooperative launch refused: %s", cudaGetErrorString(e));
    }
    CUSMC_CUDA(ctx, e);
    ctx->launches++;
    return CUSMC_OK;
}

template <int D>
int launch_persistent_d(cusmc_filter *f, PersistArgs &pa, bool fast, bool diag, bool probe_only, uint32_t *tile_out)
{
    if (fast) return diag ? launch_persistent<D, true, true>(f, pa, probe_only, tile_out)
                          : launch_persistent<D, true, false>(f, pa, probe_only, tile_out);
    return diag ? launch_persistent<D, false, true>(f, pa, probe_only, tile_out)
                : launch_persistent<D, false, false>(f, pa, probe_only, tile_out);
}

int launch_persistent_any(cusmc_filter *f, PersistArgs &pa, bool fast, bool diag, bool probe_only, uint32_t *tile_out)
{
    switch (f->cfg.d) {
        case 2: return launch_persistent_d<2>(f, pa, fast, diag, probe_only, tile_out);
        case 4: return launch_persistent_d<4>(f, pa, fast, diag, probe_only, tile_out);
        case 8: return launch_persistent_d<8>(f, pa, fast, diag, probe_only, tile_out);
    }
    return cusmc_fail(f->ctx, CUSMC_ERR_UNSUPPORTED, "persistent run: d must be 2, 4 or 8");
}

bool model_is_diag(const cusmc_filter *f)
{
    const int d = f->cfg.d;
    bool diag = cusmc_is_diag_colmajor(f->G.data(), d) && cusmc_is_diag_colmajor(f->Qw.data(), d) &&
                cusmc_is_diag_colmajor(f->Qc0.data(), d);
    for (int k = 0; k < d && diag; ++k)
        for (int j = 0; j < d; ++j)
            if (j != k && f->M[(size_t)k * d + j] != 0.0) diag = false;
    return diag;
}

bool config_allows(const cusmc_filter *f)
{
    const cusmc_filter_config &cfg = f->cfg;
    return cfg.persistent >= 0 && f->world == 1 && cfg.resampler == CUSMC_RESAMPLE_SYSTEMATIC && cfg.kind == CUSMC_MVN &&
           !cfg.keep_history && !cfg.summary && cfg.ess_threshold == 0.0 && cfg.d == cfg.dy &&
           (cfg.d == 2 || cfg.d == 4 || cfg.d == 8) && cfg.T >= 2 &&
           f->is_log && f->ep.kind == CUSMC_MVN && f->ep.want_log;      // the lean main loop's epilogue
}

}  // namespace

bool cusmc_filter_persistent_eligible(const cusmc_filter *f, const cusmc_filter_draws *draws)
{
    if (!config_allows(f) || !f->persist_tile) return false;
    if (draws && (draws->xi0_dev || draws->xi_dev || draws->chi_dev || draws->chi0_dev || draws->u_dev || draws->j_dev ||
                  draws->um_dev))
        return false;
    return true;
}

// Tile size a persistent run of this filter would use (0: configuration not covered, or the cloud does
// not fit one tile per resident block); the weight images are sized for it at creation.
uint32_t cusmc_filter_persistent_tile(cusmc_filter *f)
{
    if (!config_allows(f)) return 0;
    if (cudaSetDevice(f->ctx->device) != cudaSuccess) return 0;
    PersistArgs pa{};
    uint32_t tile_n = 0;
    const std::string keep = f->ctx->err;
    if (launch_persistent_any(f, pa, !f->cfg.reproducible_rng, model_is_diag(f), true, &tile_n) != CUSMC_OK) {
        f->ctx->err = keep;                               // a probe is not an error of the caller's
        return 0;
    }
    return tile_n;
}

int cusmc_filter_run_persistent(cusmc_filter *f, const cusmc_filter_draws *draws)
{
    cusmc_ctx *ctx = f->ctx;
    const cusmc_filter_config &cfg = f->cfg;
    if (!cusmc_filter_persistent_eligible(f, draws))
        return cusmc_fail(ctx, CUSMC_ERR_UNSUPPORTED, "configuration not covered by the persistent kernel");
    const int d = cfg.d, T = cfg.T;
    CUSMC_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    // per-run scratch: whitened observations [T][d], systematic offsets [T], the barrier words
    const size_t n_obs = (size_t)T * d, n_u0 = (size_t)T;
    const size_t bytes = 8 * (n_obs + n_u0 + 1);
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
    double *obs = (double *)f->persist, *u0 = obs + n_obs;
    unsigned *barrier = (unsigned *)(u0 + n_u0);
    CUSMC_CUDA(ctx, cudaMemsetAsync(barrier, 0, 8, st));
    // the host vector dies with this call: a synchronous copy (pageable memory) is what we want
    CUSMC_CUDA(ctx, cudaMemcpyAsync(obs, host.data(), 8 * host.size(), cudaMemcpyHostToDevice, st));
    CUSMC_CUDA(ctx, cudaStreamSynchronize(st));
    CUSMC_CHECK(cusmc_filter_init_slots(f));
    PersistArgs pa{};
    StepArgs &a = pa.fa.s;
    a.n_out = cfg.N;
    a.ld_new = a.ld_prev = f->per;
    a.ld_noise = cfg.N;
    a.seed = cfg.seed;
    a.nu = cfg.nu;
    a.d = a.dy = d;
    a.kind = CUSMC_MVN;
    a.fast_noise = cfg.reproducible_rng ? 0 : 1;
    pa.fa.img_hdr_words = fimage_header_words(f->img_n);
    pa.fa.tiles_alloc = (uint32_t)fimage_tiles(f->img_n);
    pa.fa.N_global = (uint32_t)cfg.N;
    pa.fa.tiles_per_rank = pa.fa.tiles_alloc;
    pa.fa.shift = f->shift;
    pa.x[0] = f->x[0];
    pa.x[1] = f->x[1];
    pa.img[0] = f->img[0];
    pa.img[1] = f->img[1];
    pa.lw = f->lw;
    pa.anc = f->anc;
    pa.slots = f->slots;
    pa.obs = obs;
    pa.u0 = u0;
    pa.barrier = barrier;
    pa.tiles_alloc = fimage_tiles(f->img_n);
    pa.T = T;
#ifdef CUSMC_TRACE
    pa.trace = nullptr;      // a device buffer of T x 8 doubles, handed in by the profiling script
    if (getenv("CUSMC_TRACE_BUF")) pa.trace = (double *)strtoull(getenv("CUSMC_TRACE_BUF"), nullptr, 0);
#endif
    CUSMC_CUDA(ctx, cudaEventRecord(f->ev0, st));
    uint32_t tile_n = 0;
    const int rc = launch_persistent_any(f, pa, !cfg.reproducible_rng, model_is_diag(f), false, &tile_n);
    if (rc != CUSMC_OK) return rc;
    CUSMC_CUDA(ctx, cudaEventRecord(f->ev1, st));
    f->cur = (T - 1) & 1;
    f->next_t = T;
    f->ran = true;
    return CUSMC_OK;
}

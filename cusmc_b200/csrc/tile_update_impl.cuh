// tile_update_impl.cuh -- the per-step tile update as a block-wide device function (see tile_update.cu).
#pragma once

#include "filter_types.cuh"
#include "image.cuh"
#include "mailbox.cuh"
#include "tile_update.cuh"

#include "../../include/cusmc_detmath.h"

constexpr int kUpdItems = 4;

__device__ __forceinline__ void ld4(const unsigned long long *p, unsigned long long (&v)[4])
{
    asm volatile("ld.global.cg.v4.u64 {%0, %1, %2, %3}, [%4];" : "=l"(v[0]), "=l"(v[1]), "=l"(v[2]), "=l"(v[3]) : "l"(p) : "memory");
}
__device__ __forceinline__ void st4(unsigned long long *p, const unsigned long long (&v)[4])
{
    asm volatile("st.global.v4.u64 [%0], {%1, %2, %3, %4};" ::"l"(p), "l"(v[0]), "l"(v[1]), "l"(v[2]), "l"(v[3]) : "memory");
}

// Shared memory of one update (NT threads).
template <int NT>
struct UpdateSmem {
    unsigned long long sm[32], sm2[32];
    double dbl[32];
    unsigned long long chunk, tot[2];
    StepConsts c;
};

// The update as a block-wide device function: the body of tile_update_kernel (one block of 1024 threads
// per step) and of the persistent whole-run kernel's grid barrier (the last block to arrive runs it
// with its 256 threads).  All cross-block data is read through L2 (ld.global.cg).
template <int NT>
__device__ __forceinline__ void tile_update_block(const UpdateArgs &u, UpdateSmem<NT> &us)
{
    constexpr int kUpdThreads = NT, kUpdChunk = NT * kUpdItems;
    unsigned long long *sm = us.sm, *sm2 = us.sm2;
    double *s_dbl = us.dbl;
    unsigned long long &s_chunk = us.chunk;
    unsigned long long *s_tot = us.tot;
    StepConsts &s_c = us.c;
    static_assert(kUpdItems == 4, "a thread owns four consecutive tiles: one 256-bit access per field");
    static_assert(NT % 32 == 0 && NT <= 1024, "whole warps");
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t tiles = u.tiles, tiles_all = u.tiles_alloc;
    unsigned long long *img = u.img;
    unsigned long long *fld[kTileFields];
#pragma unroll
    for (int f = 0; f < kTileFields; ++f) fld[f] = img + kConstWords + (int64_t)f * tiles_all;

    // ---- A: the maximum log-weight --------------------------------------------------------------------
    // (the first chunk's maxima stay in registers for phase B: one pass over the records, not two)
    unsigned long long m0[kUpdItems] = {0, 0, 0, 0};
    double M;
    if (u.phases & kUpdMax) {
        double m = -INFINITY;
        for (int64_t base = 0; base < tiles; base += kUpdChunk) {
            const int64_t b0 = base + (int64_t)tid * kUpdItems;
            if (b0 >= tiles) continue;
            unsigned long long mv[kUpdItems];
            ld4(fld[kTileM] + b0, mv);
#pragma unroll
            for (int r = 0; r < kUpdItems; ++r) {
                const double v = __longlong_as_double((long long)mv[r]);
                if (b0 + r < tiles && v > m) m = v;            // records hold finite values or -inf
                if (base == 0) m0[r] = mv[r];
            }
        }
        m = warp_max_double(m);
        if (lane == 0) s_dbl[warp] = m;
        __syncthreads();
        if (warp == 0) {
            m = warp_max_double(lane < kUpdThreads / 32 ? s_dbl[lane] : -INFINITY);
            if (u.mail.world > 1) {
                unsigned long long w0, w1, w2;
                mail_publish(u.mail, u.cell_max, lane, (unsigned long long)__double_as_longlong(m), 0, 0);
                mail_wait(u.mail, u.cell_max, lane, w0, w1, w2);
                double v = lane < u.mail.world ? __longlong_as_double((long long)w0) : -INFINITY;
                if (!(v == v)) v = -INFINITY;
                m = warp_max_double(v);
            }
            if (lane == 0) u.slot->lw_max = m;
            if (lane == 0) s_dbl[0] = m;
        }
        __syncthreads();
        M = s_dbl[0];
        __syncthreads();
    } else {
        M = __ldcg(&u.slot->lw_max);                      // made global by the caller (NCCL formulation)
    }
    if (!(u.phases & (kUpdScan | kUpdConsts))) return;

    // ---- B: rescale, scan, totals ---------------------------------------------------------------------
    unsigned long long T_loc = 0, T2_loc = 0;
    if (u.phases & kUpdScan) {
        unsigned long long carry = 0, t2 = 0;
        // every tile up to tiles_alloc gets its F / P / Sp: tiles past the end of the shard are empty
        for (int64_t base = 0; base < tiles_all; base += kUpdChunk) {
            const int64_t b0 = base + (int64_t)tid * kUpdItems;
            unsigned long long F[kUpdItems] = {0, 0, 0, 0}, sp[kUpdItems] = {0, 0, 0, 0}, run = 0;
            if (b0 < tiles) {
                unsigned long long mv[kUpdItems], S[kUpdItems], S2[kUpdItems];
                if (base == 0 && (u.phases & kUpdMax)) {
#pragma unroll
                    for (int r = 0; r < kUpdItems; ++r) mv[r] = m0[r];
                } else {
                    ld4(fld[kTileM] + b0, mv);
                }
                ld4(fld[kTileS] + b0, S);
                ld4(fld[kTileS2] + b0, S2);
#pragma unroll
                for (int r = 0; r < kUpdItems; ++r)
                    if (b0 + r < tiles) {
                        F[r] = cusmc_rescale_factor(__longlong_as_double((long long)mv[r]), M);
                        sp[r] = cusmc_mulshift62(S[r], F[r]);
                        t2 += cusmc_mulshift62(cusmc_mulshift62(S2[r], F[r]), F[r]);
                        run += sp[r];
                    }
            }
            unsigned long long inc = run;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned long long t = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += t;
            }
            if (lane == 31) sm[warp] = inc;
            __syncthreads();
            if (warp == 0) {                               // exclusive scan of the 32 warp totals
                const unsigned long long w = lane < kUpdThreads / 32 ? sm[lane] : 0ull;
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
            if (b0 < tiles_all) {
                unsigned long long P[kUpdItems];
                unsigned long long excl = carry + sm[warp] + inc - run;
#pragma unroll
                for (int r = 0; r < kUpdItems; ++r) {
                    P[r] = excl;
                    excl += sp[r];
                }
                st4(fld[kTileF] + b0, F);
                st4(fld[kTileP] + b0, P);
                st4(fld[kTileSp] + b0, sp);
            }
            carry += s_chunk;
            __syncthreads();
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t2 += __shfl_xor_sync(0xffffffffu, t2, o);
        if (lane == 0) sm2[warp] = t2;
        __syncthreads();
        if (tid == 0) {
            unsigned long long s = 0;
            for (int k = 0; k < kUpdThreads / 32; ++k) s += sm2[k];
            s_tot[0] = carry;
            s_tot[1] = s;
            // this rank's sums: what the NCCL formulation all-gathers (words 1..2 of the slot)
            u.slot->sum_q = carry;
            u.slot->sum_q2 = s;
        }
        __syncthreads();
        T_loc = s_tot[0];
        T2_loc = s_tot[1];
    }
    if (!(u.phases & kUpdConsts)) return;

    // ---- C: global totals, rank offsets, constants of the next step -----------------------------------
    if (warp == 0) {
        unsigned long long t_r = 0, t2_r = 0;             // lane r: rank r's totals
        if (u.mail.world > 1) {
            unsigned long long w2;
            mail_publish(u.mail, u.cell_sums, lane, T_loc, T2_loc, 0);
            mail_wait(u.mail, u.cell_sums, lane, t_r, t2_r, w2);
        } else if (u.world > 1) {
            // NCCL formulation: the caller all-gathered every rank's (sum_q, sum_q2, n_pos) into rank_sums
            if (lane < u.world) {
                t_r = u.rank_sums[3 * lane];
                t2_r = u.rank_sums[3 * lane + 1];
            }
        } else if (lane == 0) {
            if (u.phases & kUpdScan) {
                t_r = T_loc;
                t2_r = T2_loc;
            } else {
                t_r = u.slot->sum_q;
                t2_r = u.slot->sum_q2;
            }
        }
        // exclusive prefix over ranks (lanes), totals
        unsigned long long inc = t_r;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        const unsigned long long T = __shfl_sync(0xffffffffu, inc, 31);
        unsigned long long T2 = t2_r;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) T2 += __shfl_xor_sync(0xffffffffu, T2, o);
        if (lane < CUSMC_MAX_PEERS) s_c.rank_off[lane] = inc - t_r;
        if (lane == 0) {
            unsigned long long rr = (unsigned long long)(u.u0_next * (double)T);
            if (T && rr > T - 1) rr = T - 1;
            s_c.T = T;
            s_c.r0 = rr;
            s_c.ng_over_t = (double)u.N_global / (double)T;
            s_c.r0_over_t = (double)rr / (double)T;
            // ESS = T^2 / (T2 2^shift) < threshold N  <=>  T^2 < ess_bound T2   (ess_bound = threshold N 2^shift)
            s_c.resample = 1;
            if (u.ess_bound > 0.0) s_c.resample = (double)T * (double)T < u.ess_bound * (double)T2;
            s_c.T2 = T2;
            s_c.M = M;
            s_c.reserved = 0;
            u.slot->sum_q = T;
            u.slot->sum_q2 = T2;
            u.slot->n_pos = 0;
            if (u.slot_next) {
                u.slot_next->resampled = s_c.resample;
                u.slot_next->degenerate = T == 0 ? 1 : 0;
            }
        }
        __syncwarp();
        if (lane == 0) u.slot->cdf_offset = s_c.rank_off[u.rank];
    }
    __syncthreads();
    if (tid < kConstWords) img[tid] = reinterpret_cast<const unsigned long long *>(&s_c)[tid];
}

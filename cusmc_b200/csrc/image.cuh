// image.cuh -- the block-relative weight image the filter's normalised resamplers live on.
//
// The fused step kernel (pf_fused_impl.cuh) finishes a tile of kTile consecutive children on its own:
// it weighs them against the TILE's maximum log-weight m_b (no grid-wide dependency), leaves the
// tile-local inclusive prefix c_i of the fixed-point weights (8 bytes per particle) and a 32-byte
// tile record {m_b, S_b, S2_b}.  One small kernel per step (tile_update.cu) then takes the global
// maximum M, rescales every tile by F_b = trunc(exp(m_b - M) 2^62) in integer arithmetic, scans the
// tile masses, and leaves in the image header everything the NEXT step's kernel needs to find each
// child's parent by itself:
//
//     C_i = rank_off + P_b + (c_i F_b >> 62)        global integer CDF, never materialised
//
// so a systematic step is ONE big launch: no ancestor scatter pass, no separate weigh pass, and the
// parent gather of a tile reads a window of ~kTile consecutive parents.  oracle: orc_tile_image.
//
// Layout of one image (uint64 words), n local particles, tiles = ceil(n / kTile) rounded up to 4:
//     [0 .. 15]              StepConsts  (128 bytes), written by tile_update_kernel
//     [16 + f tiles + b]     field f of tile b, structure of arrays (so the one-block update kernel reads
//                            and writes them with coalesced 256-bit accesses; as 64-byte records its
//                            strided accesses alone cost 15 us per step):
//                               f = 0 m   max finite log-weight of the tile (-inf: none)   } written by the
//                               f = 1 S   sum of fixed-point weights relative to m         } step kernel that
//                               f = 2 S2  sum of squared weights relative to m             } produced the tile
//                               f = 3 F   rescale factor trunc(exp(m - M) 2^62)            } written by
//                               f = 4 P   exclusive prefix of the rescaled masses, per rank } tile_update_kernel
//                               f = 5 Sp  rescaled mass (S F) >> 62                        }
//     [H + i]                c_i, padded to whole tiles   (H = 16 + 6 tiles)
// Two images per filter (step parity): step t reads image t - 1 while it writes image t.
#pragma once

#include "common.cuh"
#include "resample.cuh"

#include "../../include/cusmc_detmath.h"

struct StepConsts {
    unsigned long long T;            // global fixed-point mass of the step this image belongs to
    unsigned long long r0;           // systematic offset of the NEXT step in mass units
    double ng_over_t, r0_over_t;     // N_global / T, r0 / T (the floating estimate of offspring_below)
    unsigned long long resample;     // 1: the next step draws new ancestors (adaptive resampling decision)
    unsigned long long T2;           // global sum of squared weights (ESS)
    double M;                        // global maximum log-weight
    unsigned long long reserved;
    unsigned long long rank_off[CUSMC_MAX_PEERS];   // mass held by lower-ranked shards, per rank
};
static_assert(sizeof(StepConsts) == 128, "StepConsts is 16 words");

enum TileField { kTileM = 0, kTileS = 1, kTileS2 = 2, kTileF = 3, kTileP = 4, kTileSp = 5, kTileFields = 6 };
constexpr int kConstWords = sizeof(StepConsts) / 8;

// tiles of an n-particle shard, as laid out (a multiple of 4: 256-bit accesses stay aligned)
__host__ __device__ inline int64_t fimage_tiles(int64_t n) { return n > 0 ? (((n + kTile - 1) / kTile) + 3) & ~(int64_t)3 : 4; }
__host__ __device__ inline int64_t fimage_field_word(int64_t n, int f) { return kConstWords + (int64_t)f * fimage_tiles(n); }
__host__ __device__ inline int64_t fimage_header_words(int64_t n) { return kConstWords + kTileFields * fimage_tiles(n); }
__host__ __device__ inline int64_t fimage_words(int64_t n) { return fimage_header_words(n) + fimage_tiles(n) * kTile; }

#ifdef __CUDACC__
__device__ __forceinline__ const StepConsts *fimage_consts(const unsigned long long *img)
{
    return reinterpret_cast<const StepConsts *>(img);
}

// The fixed-point weight of a log-weight against its tile's maximum m and the image of its square:
//     q = cusmc_fixed_from_unit(w, shift),  q2 = cusmc_fixed_from_unit(w w, shift),  w = cusmc_unit_from_log(lw, m)
// bit for bit, with the selects those three calls repeat folded into two: w is exp of a number in [-43.5, 0]
// or nothing, so it lies in (0, 1] (the last Horner step of cusmc_det_exp_core is fma(p, r, 1) with r <= 0
// when k = 0, and p 2^k <= 0.71 when k < 0) and the clamps of cusmc_fixed_from_unit never act; "nothing" (NaN,
// +inf, below 2^-62, above the maximum) zeroes the SCALE 2^shift -- a power of two, whose low word is zero
// either way: one 32-bit select.
__device__ __forceinline__ void weigh_fixed(double lw, double m, int shift, unsigned long long &q, unsigned long long &q2)
{
    const double x = lw - m;
    const bool ok = x >= -43.5 && x <= 0.0;
    int k;
    const double p = cusmc_det_exp_core(ok ? x : 0.0, &k);
    const double w = p * cusmc_pow2i(k);
    const double scale = __hiloint2double(ok ? (shift + 1023) << 20 : 0, 0);
    q = (unsigned long long)(w * scale);
    q2 = (unsigned long long)((w * w) * scale);
}

// Q_i = floor((i T + r0) / N): the largest CDF value NOT above child i, i.e. for any integer C
//     offspring_below(C) <= i   <=>   C N <= i T + r0   <=>   C <= Q_i.
// Every "does this parent / tile / rank reach child i" question of the lookup becomes one 64-bit
// compare against a Q computed once per block.  128-by-32-bit division in two 64-bit steps
// (i T + r0 < N T, so the high word is below N and the quotient fits 64 bits).
__device__ __forceinline__ unsigned long long mass_quotient(unsigned long long i, unsigned long long T,
                                                            unsigned long long r0, unsigned long long Ng)
{
    unsigned long long lo = i * T, hi = __umul64hi(i, T);
    const unsigned long long lo2 = lo + r0;
    hi += lo2 < lo;
    // Each step divides a number below N 2^32 by N < 2^32 (the quotient fits 32 bits): an fp64 quotient is
    // within one of it (both roundings are relative 2^-53, the quotient is below 2^32), one exact
    // remainder in wrap-around 64-bit arithmetic corrects it.  A third of the generic 64-bit division's
    // instructions, on the serial path between the tile update and the lookup.
    auto div_step = [&](unsigned long long cur, unsigned long long &rem) {
        unsigned long long q = (unsigned long long)((double)cur / (double)Ng);
        long long r = (long long)(cur - q * Ng);
        if (r < 0) {
            q -= 1;
            r += (long long)Ng;
        } else if (r >= (long long)Ng) {
            q += 1;
            r -= (long long)Ng;
        }
        rem = (unsigned long long)r;
        return q;
    };
    unsigned long long rem;
    const unsigned long long q1 = div_step((hi << 32) | (lo2 >> 32), rem);      // hi < N < 2^32
    const unsigned long long q0 = div_step((rem << 32) | (lo2 & 0xffffffffull), rem);
    return (q1 << 32) | q0;
}
#endif

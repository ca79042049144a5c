// tile_update.cuh -- arguments of the per-step tile update (tile_update.cu), shared with filter.cu.
#pragma once

#include "common.cuh"
#include "mailbox.cuh"

struct StepSlot;

enum { kUpdMax = 1, kUpdScan = 2, kUpdConsts = 4, kUpdAll = 7 };

struct UpdateArgs {
    unsigned long long *img;          // weight image of step t on this rank
    StepSlot *slot;                   // slot[t]: receives M, T, T2 and this rank's CDF offset
    StepSlot *slot_next;              // slot[t + 1] or NULL: receives the resampling decision / degenerate flag
    const unsigned long long *rank_sums;   // NCCL formulation, phase C: all-gathered (sum_q, sum_q2, n_pos) per rank
    int64_t tiles;                    // tiles of this rank's shard (0: empty shard)
    int64_t tiles_alloc;              // tiles per rank as laid out (image.cuh: a multiple of 4, the same on every rank)
    unsigned lo;                      // global slot of this rank's first particle
    unsigned N_global;
    double u0_next;                   // systematic offset of step t + 1, [0, 1)
    double ess_bound;                 // adaptive resampling bound (0: always resample)
    int rank, world;
    int phases;                       // kUpdMax | kUpdScan | kUpdConsts
    MailArgs mail;                    // world > 1 inside cusmc_filter_run_sharded: exchanges ride in the kernel
    size_t cell_max, cell_sums;
};

int cusmc_launch_tile_update(cusmc_ctx *ctx, const UpdateArgs &u);
// multinomial ancestors of n local children (global slots i0 ..) searched on the weight images themselves, peers' included
// (img_peer: device table of every rank's image, NULL on one GPU; n_alloc: the particle count the images were sized for)
int cusmc_launch_multinomial_image(cusmc_ctx *ctx, const unsigned long long *img, const unsigned long long *const *img_peer,
                                   int64_t n_alloc, int64_t n, int64_t i0, int64_t N_global, int64_t per_rank, int rank,
                                   int world, const double *u, uint64_t seed, uint64_t step, uint32_t *a,
                                   unsigned long long *degenerate_out);

// resample.cuh -- internal launchers shared between resample.cu and filter.cu.
#pragma once

#include "common.cuh"
#include "mailbox.cuh"

int cusmc_fill_double(cusmc_ctx *ctx, double *p, double v, int n);
int cusmc_launch_metropolis(cusmc_ctx *ctx, uint32_t *a, const double *w, const double *u,
                            const uint32_t *j, uint64_t seed, uint64_t step, int64_t N, int B,
                            int is_log, int64_t i0, int64_t n_out, const CusmcPeers *peers);
int cusmc_launch_weights_max(cusmc_ctx *ctx, const double *w, int64_t N, double *max_dev);
// image: cusmc_scan_state_bytes(N) bytes whose first word is zero; receives the weight image
// (exclusive tile prefixes + tile-local CDF) the resampling pass consumes.  stats_dev may be NULL.
int cusmc_launch_weights_sum(cusmc_ctx *ctx, const double *w, int is_log, const double *max_dev,
                             int64_t N, int shift, uint64_t *stats_dev, void *image, bool full_stats,
                             const MailArgs *mail = nullptr, int t = 0);
size_t cusmc_scan_state_bytes(int64_t N);
int cusmc_launch_scan(cusmc_ctx *ctx, int64_t N, int64_t N_global, const uint64_t *total_dev,
                      const uint64_t *cdf_offset_dev, const void *image,
                      uint64_t *cdf_out, uint32_t *anc_out, int64_t j0, int64_t out_lo,
                      int64_t out_n, double u0, const CusmcPeers *peers);
int cusmc_launch_multinomial(cusmc_ctx *ctx, const uint64_t *cdf, int64_t N, const uint64_t *total_dev,
                             const double *u, uint64_t seed, uint64_t step, int64_t i0,
                             int64_t n_out, int64_t j0, uint32_t *a);

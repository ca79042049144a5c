// tile_update.cu -- the one grid-wide dependency of a fused filter step, as ONE small kernel.
//
// pf_fused_kernel leaves a record {m_b, S_b, S2_b} per tile (the tile's maximum log-weight and its
// fixed-point sums RELATIVE to that maximum).  This kernel turns them into the weight image the next
// step reads (image.cuh; oracle: orc_tile_image):
//
//   A  M = max_b m_b                                    (sharded: all-reduce over the mailboxes)
//   B  F_b = trunc(exp(m_b - M) 2^62),  Sp_b = (S_b F_b) >> 62,  P_b = exclusive prefix of Sp inside the
//      rank,  T = sum Sp,  T2 = sum ((S2_b F_b >> 62) F_b >> 62)
//                                                       (sharded: all-gather of (T, T2) -> rank offsets)
//   C  the constants of the NEXT step's systematic resampling (r0, N/T, r0/T, the adaptive decision,
//      the degenerate flag).  (Which parent tile a child tile starts at is found by the step kernel's
//      blocks themselves -- two block-wide compare rounds over the compact prefix array; doing it here
//      for all tiles cost this one-block kernel 25 of its 38 us.)
//
// Integer arithmetic throughout: the result does not depend on block, tile or rank order.  This is the
// max-shifted log-sum-exp normalisation + ESS of the north star (no counterpart in the reference,
// SURVEY a11): log-sum-exp = M + log(T / 2^shift), ESS = T^2 / (T2 2^shift).
#include "tile_update_impl.cuh"

#include <algorithm>

namespace {

constexpr int kUpdThreads = 1024;

__global__ void __launch_bounds__(kUpdThreads)
tile_update_kernel(const UpdateArgs u)
{
    __shared__ UpdateSmem<kUpdThreads> us;
    // let the next step kernel (a programmatic dependent, pf_fused_kernel) be launched now: its blocks wait
    // at their griddepcontrol.wait until this grid has completed
    asm volatile("griddepcontrol.launch_dependents;");
    // launched itself as a programmatic dependent of the step kernel: resident during that kernel's last
    // wave, it starts the moment the grid has completed (a no-op after any other kernel)
    asm volatile("griddepcontrol.wait;" ::: "memory");
    tile_update_block<kUpdThreads>(u, us);
}

// Multinomial resampling straight on the weight images (src/mcmc.cpp:295 with the registry's multinomial entry,
// SURVEY N4): child i draws p_i = min((uint64)(u_i T), T - 1) and takes a_i = #{ j : C_j <= p_i } over the GLOBAL
// integer CDF C_j = rank_off + P_b + (c_j F_b >> 62) -- which is never materialised: the count is found by three
// nested searches, rank (<= 8 compares on the step constants), tile (binary search over that rank's compact
// prefix array), particle (binary search inside the tile, the rescale applied per probe).  Sharded, the rank's
// image is a peer's, read over NVLink; the result is the same integer whichever GPU computes it.
// oracle: orc_filter_det, resampler 2.
template <bool PEERS, bool PREDRAWN>
__global__ void __launch_bounds__(256)
multinomial_image_kernel(const unsigned long long *__restrict__ img, const unsigned long long *const *__restrict__ img_peer,
                         int64_t hdr_words, int64_t tiles_alloc, int64_t n_out, int64_t i0, int64_t N_global,
                         int64_t per_rank, int rank, int world, const double *__restrict__ u, uint64_t seed, uint64_t step,
                         uint32_t *__restrict__ a, unsigned long long *degenerate_out)
{
    const int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (t >= n_out) return;
    const StepConsts *sc = fimage_consts(img);
    const uint64_t T = sc->T;
    if (T == 0) {                                     // no mass: identity ancestors, flagged
        a[t] = (uint32_t)(i0 + t);
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
    int rk = 0;
    if (PEERS)
        for (int r = 1; r < world; ++r)
            if (sc->rank_off[r] <= pos) rk = r;       // offsets ascend: the last rank starting at or below p
    const unsigned long long *im = (PEERS && rk != rank) ? img_peer[rk] : img;
    const uint64_t rem = pos - sc->rank_off[rk];
    const int64_t n_r = min(per_rank, N_global - (int64_t)rk * per_rank);
    const unsigned long long *P = im + kConstWords + (int64_t)kTileP * tiles_alloc;
    int64_t lo = 1, hi = (n_r + kTile - 1) / kTile;  // tiles b with P_b <= rem form a prefix; P_0 = 0 is one
    while (lo < hi) {
        const int64_t mid = lo + ((hi - lo) >> 1);
        if (__ldg(P + mid) <= rem) lo = mid + 1; else hi = mid;
    }
    const int64_t b = lo - 1;
    const uint64_t rem_b = rem - __ldg(P + b);
    const uint64_t F = __ldg(im + kConstWords + (int64_t)kTileF * tiles_alloc + b);
    const unsigned long long *c = im + hdr_words + b * kTile;
    lo = 0;
    hi = min((int64_t)kTile, n_r - b * kTile);
    while (lo < hi) {
        const int64_t mid = lo + ((hi - lo) >> 1);
        if (cusmc_mulshift62(__ldg(c + mid), F) <= rem_b) lo = mid + 1; else hi = mid;
    }
    a[t] = (uint32_t)((int64_t)rk * per_rank + b * kTile + lo);
}

}  // namespace

int cusmc_launch_tile_update(cusmc_ctx *ctx, const UpdateArgs &u)
{
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cudaLaunchConfig_t lc{};
    lc.gridDim = dim3(1);
    lc.blockDim = dim3(kUpdThreads);
    lc.stream = ctx->stream;
    lc.attrs = attr;
    lc.numAttrs = 1;
    CUSMC_CUDA(ctx, cudaLaunchKernelEx(&lc, tile_update_kernel, u));
    CUSMC_LAUNCHED(ctx);
    return CUSMC_OK;
}

int cusmc_launch_multinomial_image(cusmc_ctx *ctx, const unsigned long long *img, const unsigned long long *const *img_peer,
                                   int64_t n_alloc, int64_t n, int64_t i0, int64_t N_global, int64_t per_rank, int rank,
                                   int world, const double *u, uint64_t seed, uint64_t step, uint32_t *a,
                                   unsigned long long *degenerate_out)
{
    if (n == 0) return CUSMC_OK;
    const int grid = (int)((n + 255) / 256);
    const int64_t hdr = fimage_header_words(n_alloc), tiles = fimage_tiles(n_alloc);
#define CUSMC_MULTI_GO(PE, PR)                                                                                         \
    multinomial_image_kernel<PE, PR><<<grid, 256, 0, ctx->stream>>>(img, img_peer, hdr, tiles, n, i0, N_global, per_rank, \
                                                                    rank, world, u, seed, step, a, degenerate_out)
    if (world > 1 && img_peer) {
        if (u) CUSMC_MULTI_GO(true, true); else CUSMC_MULTI_GO(true, false);
    } else {
        if (u) CUSMC_MULTI_GO(false, true); else CUSMC_MULTI_GO(false, false);
    }
#undef CUSMC_MULTI_GO
    CUSMC_LAUNCHED(ctx);
    return CUSMC_OK;
}

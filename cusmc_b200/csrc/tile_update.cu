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

// Global inclusive CDF from the image (multinomial resampling searches it): C_i = off + P_b + (c_i F_b >> 62).
__global__ void __launch_bounds__(256)
image_cdf_kernel(const unsigned long long *__restrict__ img, int64_t hdr_words, int64_t tiles_alloc, int64_t n, int rank,
                 unsigned long long *__restrict__ cdf)
{
    const StepConsts *sc = fimage_consts(img);
    const unsigned long long off = sc->rank_off[rank];
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
        const unsigned long long *fld = img + kConstWords + i / kTile;
        cdf[i] = off + fld[(int64_t)kTileP * tiles_alloc] + cusmc_mulshift62(img[hdr_words + i], fld[(int64_t)kTileF * tiles_alloc]);
    }
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

int cusmc_launch_image_cdf(cusmc_ctx *ctx, const unsigned long long *img, int64_t n_alloc, int64_t n, int rank,
                           uint64_t *cdf)
{
    if (n == 0) return CUSMC_OK;
    const int grid = (int)std::min<int64_t>((n + 255) / 256, (int64_t)ctx->sm_count * 16);
    image_cdf_kernel<<<grid, 256, 0, ctx->stream>>>(img, fimage_header_words(n_alloc), fimage_tiles(n_alloc), n, rank,
                                                    (unsigned long long *)cdf);
    CUSMC_LAUNCHED(ctx);
    return CUSMC_OK;
}

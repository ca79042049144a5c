// density_variants.cu -- microbenchmark used to choose the shape of the headline kernel
// (batched MVN log-density, d = 16, 2^20 points, SoA).  Not part of the library: it includes the
// library's quadratic-form header so every variant computes the same bits, times each variant
// with CUDA events over inputs that rotate through a pool larger than L2, and prints one line
// per variant.  Build + run:  make -C profiles/micro run   (on a B200).
#include "../../cusmc_b200/csrc/density.cuh"

#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x)                                                                              \
    do {                                                                                   \
        cudaError_t e = (x);                                                               \
        if (e != cudaSuccess) {                                                            \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); \
            exit(1);                                                                       \
        }                                                                                  \
    } while (0)

constexpr int D = 16;
using Op = AffineOp<D, true>;

// ---- variant A: the library's r01 kernel (thread owns two points, 128-bit loads) --------------
template <int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB)
k_vec2(const __grid_constant__ Op op, const Epilogue ep, const double *__restrict__ x, int64_t n_units,
       int64_t ld, double *__restrict__ out)
{
    const int64_t u = (int64_t)blockIdx.x * THREADS + threadIdx.x;
    if (u >= n_units) return;
    const int64_t i = 2 * u;
    double ra[D], rb[D];
#pragma unroll
    for (int j = 0; j < D; ++j) {
        const double2 v = ld_stream2(x + (int64_t)j * ld + i);
        ra[j] = v.x - op.shift[j];
        rb[j] = v.y - op.shift[j];
    }
    double2 res;
    res.x = density_epilogue(ep, affine_quadform<D, true>(op, ra));
    res.y = density_epilogue(ep, affine_quadform<D, true>(op, rb));
    st_stream2(out + i, res);
}

// ---- variant B: one point per thread, 64-bit loads, more resident warps ------------------------
template <int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB)
k_vec1(const __grid_constant__ Op op, const Epilogue ep, const double *__restrict__ x, int64_t N,
       int64_t ld, double *__restrict__ out)
{
    const int64_t u = (int64_t)blockIdx.x * THREADS + threadIdx.x;
    if (u >= N) return;
    double r[D];
#pragma unroll
    for (int j = 0; j < D; ++j) r[j] = ld_stream(x + (int64_t)j * ld + u) - op.shift[j];
    st_stream(out + u, density_epilogue(ep, affine_quadform<D, true>(op, r)));
}

// ---- variant C: persistent CTAs, TMA (1-D bulk copy) producer warp, mbarrier ring ---------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

template <int P, int S, int MINB, int PPT /* points per consumer thread */>
__global__ void __launch_bounds__(P / PPT + 32, MINB)
k_tma(const __grid_constant__ Op op, const Epilogue ep, const double *__restrict__ x, int64_t N, int64_t ld,
      double *__restrict__ out, int n_tiles)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *tile = reinterpret_cast<double *>(smem_raw);                         // [S][D][P]
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + sizeof(double) * S * D * P);
    uint64_t *empty = full + S;
    constexpr int NC = P / PPT;   // consumer threads
    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], NC / 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid >= NC) {
        if (tid == NC) {   // producer: one elected thread
            int it = 0;
            for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
                const int s = it % S;
                const uint32_t ph = (it / S) & 1;
                mbar_wait(&empty[s], ph ^ 1);
                const int64_t base = (int64_t)t * P;
                const int npts = (int)((N - base) < P ? (N - base) : P);
                const uint32_t row_bytes = (uint32_t)npts * 8u;
                mbar_expect_tx(&full[s], row_bytes * D);
#pragma unroll
                for (int j = 0; j < D; ++j)
                    tma_load_1d(tile + ((size_t)s * D + j) * P, x + (int64_t)j * ld + base, row_bytes, &full[s]);
            }
        }
        return;
    }
    int it = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
        const int s = it % S;
        const uint32_t ph = (it / S) & 1;
        const int64_t base = (int64_t)t * P;
        const int npts = (int)((N - base) < P ? (N - base) : P);
        mbar_wait(&full[s], ph);
        double r[PPT][D];
        const double *ts = tile + (size_t)s * D * P;
#pragma unroll
        for (int j = 0; j < D; ++j) {
            if constexpr (PPT == 2) {
                const double2 v = *reinterpret_cast<const double2 *>(ts + j * P + 2 * tid);
                r[0][j] = v.x - op.shift[j];
                r[1][j] = v.y - op.shift[j];
            } else {
                r[0][j] = ts[j * P + tid] - op.shift[j];
            }
        }
        __syncwarp();
        if ((tid & 31) == 0) mbar_arrive(&empty[s]);
        if constexpr (PPT == 2) {
            if (2 * tid + 1 < npts) {
                double2 res;
                res.x = density_epilogue(ep, affine_quadform<D, true>(op, r[0]));
                res.y = density_epilogue(ep, affine_quadform<D, true>(op, r[1]));
                st_stream2(out + base + 2 * tid, res);
            }
        } else {
            if (tid < npts) st_stream(out + base + tid, density_epilogue(ep, affine_quadform<D, true>(op, r[0])));
        }
    }
}

// ---- calibration: pure streaming read of the same 16 rows (no math), writes one double/point ----
template <int THREADS>
__global__ void __launch_bounds__(THREADS)
k_readonly(const double *__restrict__ x, int64_t n_units, int64_t ld, double *__restrict__ out)
{
    const int64_t u = (int64_t)blockIdx.x * THREADS + threadIdx.x;
    if (u >= n_units) return;
    double2 acc = {0.0, 0.0};
#pragma unroll
    for (int j = 0; j < D; ++j) {
        const double2 v = ld_stream2(x + (int64_t)j * ld + 2 * u);
        acc.x += v.x;
        acc.y += v.y;
    }
    st_stream2(out + 2 * u, acc);
}

struct Bench {
    int64_t N = 1 << 20;
    int pool = 8;
    std::vector<double *> xs;
    double *out = nullptr, *ref = nullptr;
    Op op;
    Epilogue ep;
    cudaEvent_t e0, e1;
    int steps = 200;
};

template <typename F>
void run(Bench &b, const char *name, F launch, bool check = true)
{
    for (int i = 0; i < 10; ++i) launch(b.xs[i % b.pool], b.out);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(b.e0));
    for (int i = 0; i < b.steps; ++i) launch(b.xs[i % b.pool], b.out);
    CK(cudaEventRecord(b.e1));
    CK(cudaDeviceSynchronize());
    float ms;
    CK(cudaEventElapsedTime(&ms, b.e0, b.e1));
    const double us = 1e3 * ms / b.steps;
    const double gbs = 136.0 * b.N / (us * 1e-6) / 1e9;
    int bad = -1;
    if (check) {
        launch(b.xs[0], b.out);
        std::vector<double> h(b.N), r(b.N);
        CK(cudaMemcpy(h.data(), b.out, 8 * b.N, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(r.data(), b.ref, 8 * b.N, cudaMemcpyDeviceToHost));
        bad = 0;
        for (int64_t i = 0; i < b.N; ++i) bad += (h[i] != r[i]);
    }
    printf("%-34s %8.2f us  %8.1f GB/s  frac(6543.4)=%.3f  mismatches=%d\n", name, us, gbs, gbs / 6543.4, bad);
    fflush(stdout);
}

template <int P, int S, int MINB, int PPT>
void run_tma(Bench &b, const char *name, int ctas_per_sm)
{
    const size_t smem = sizeof(double) * S * D * P + 2 * S * sizeof(uint64_t);
    CK(cudaFuncSetAttribute(k_tma<P, S, MINB, PPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int n_tiles = (int)((b.N + P - 1) / P);
    int grid = 148 * ctas_per_sm;
    if (grid > n_tiles) grid = n_tiles;
    run(b, name, [&](const double *x, double *o) {
        k_tma<P, S, MINB, PPT><<<grid, P / PPT + 32, smem>>>(b.op, b.ep, x, b.N, b.N, o, n_tiles);
    });
}

int main(int argc, char **argv)
{
    Bench b;
    if (argc > 1) b.steps = atoi(argv[1]);
    CK(cudaEventCreate(&b.e0));
    CK(cudaEventCreate(&b.e1));
    std::vector<double> h((size_t)D * b.N);
    uint64_t s = 88172645463325252ull;
    for (auto &v : h) {
        s ^= s << 13; s ^= s >> 7; s ^= s << 17;
        v = (double)(int64_t)(s >> 11) * (1.0 / 9007199254740992.0) * 4.0 - 2.0;
    }
    for (int p = 0; p < b.pool; ++p) {
        double *x;
        CK(cudaMalloc(&x, 8 * h.size()));
        CK(cudaMemcpy(x, h.data(), 8 * h.size(), cudaMemcpyHostToDevice));
        b.xs.push_back(x);
    }
    CK(cudaMalloc(&b.out, 8 * b.N));
    CK(cudaMalloc(&b.ref, 8 * b.N));
    memset(&b.op, 0, sizeof(b.op));
    for (int k = 0; k < D; ++k) {
        for (int j = 0; j <= k; ++j) b.op.M[k * (k + 1) / 2 + j] = (j == k ? 1.0 : 0.01 * (k - j)) / (1.0 + 0.1 * k);
        b.op.shift[k] = 0.1 * k;
    }
    b.ep.kind = CUSMC_MVN;
    b.ep.want_log = 1;
    b.ep.lognorm = -14.7;
    b.ep.scale = exp(-14.7);
    b.ep.half_nu_d = b.ep.inv_nu = 0.0;

    const int64_t units = b.N / 2;
    // reference bits
    k_vec2<256, 3><<<(unsigned)((units + 255) / 256), 256>>>(b.op, b.ep, b.xs[0], units, b.N, b.ref);
    CK(cudaDeviceSynchronize());

    run(b, "readonly 256thr", [&](const double *x, double *o) {
        k_readonly<256><<<(unsigned)((units + 255) / 256), 256>>>(x, units, b.N, o); }, false);
    run(b, "memcpy d2d 71 MB (r+w = 142 MB)", [&](const double *x, double *o) {
        CK(cudaMemcpyAsync(b.xs[b.pool - 1], x, 71303168, cudaMemcpyDeviceToDevice)); }, false);
    CK(cudaMemcpy(b.xs[b.pool - 1], h.data(), 8 * h.size(), cudaMemcpyHostToDevice));
    run(b, "A vec2 256thr x3 (r01)", [&](const double *x, double *o) {
        k_vec2<256, 3><<<(unsigned)((units + 255) / 256), 256>>>(b.op, b.ep, x, units, b.N, o); });
    run(b, "A vec2 128thr x6", [&](const double *x, double *o) {
        k_vec2<128, 6><<<(unsigned)((units + 127) / 128), 128>>>(b.op, b.ep, x, units, b.N, o); });
    run(b, "A vec2 64thr x12", [&](const double *x, double *o) {
        k_vec2<64, 12><<<(unsigned)((units + 63) / 64), 64>>>(b.op, b.ep, x, units, b.N, o); });
    run(b, "B vec1 256thr x4", [&](const double *x, double *o) {
        k_vec1<256, 4><<<(unsigned)((b.N + 255) / 256), 256>>>(b.op, b.ep, x, b.N, b.N, o); });
    run(b, "B vec1 256thr x5", [&](const double *x, double *o) {
        k_vec1<256, 5><<<(unsigned)((b.N + 255) / 256), 256>>>(b.op, b.ep, x, b.N, b.N, o); });
    run(b, "B vec1 128thr x10", [&](const double *x, double *o) {
        k_vec1<128, 10><<<(unsigned)((b.N + 127) / 128), 128>>>(b.op, b.ep, x, b.N, b.N, o); });
    run(b, "B vec1 128thr x12", [&](const double *x, double *o) {
        k_vec1<128, 12><<<(unsigned)((b.N + 127) / 128), 128>>>(b.op, b.ep, x, b.N, b.N, o); });
    run_tma<128, 4, 3, 1>(b, "C tma P128 S4 x3 ppt1", 3);
    run_tma<128, 3, 4, 1>(b, "C tma P128 S3 x4 ppt1", 4);
    run_tma<128, 6, 2, 1>(b, "C tma P128 S6 x2 ppt1", 2);
    run_tma<256, 3, 2, 1>(b, "C tma P256 S3 x2 ppt1", 2);
    run_tma<256, 2, 3, 1>(b, "C tma P256 S2 x3 ppt1", 3);
    run_tma<256, 3, 2, 2>(b, "C tma P256 S3 x2 ppt2", 2);
    run_tma<256, 2, 3, 2>(b, "C tma P256 S2 x3 ppt2", 3);
    run_tma<64, 4, 6, 1>(b, "C tma P64 S4 x6 ppt1", 6);
    run_tma<64, 8, 3, 1>(b, "C tma P64 S8 x3 ppt1", 3);
    run_tma<512, 2, 1, 2>(b, "C tma P512 S2 x1 ppt2", 1);
    run_tma<512, 3, 1, 1>(b, "C tma P512 S3 x1 ppt1", 1);
    return 0;
}

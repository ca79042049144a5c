// fp64_peak.cu -- register-resident DFMA loop: the fp64 FMA peak of this B200 (the denominator for
// the compute-bound fractions quoted for the MH-chain kernel; MEASURED_PEAKS.json has no fp64 entry).
#include <cstdio>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(256) dfma_loop(double *out, int iters, double a, double b)
{
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; ++i) {
        x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
        x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

int main()
{
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int blocks = sms * 8, threads = 256, iters = 20000;
    double *out;
    cudaMalloc(&out, sizeof(double) * blocks * threads);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        dfma_loop<<<blocks, threads>>>(out, iters, 0.999999, 1e-9);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        const double flops = 2.0 * 8 * (double)iters * blocks * threads;
        printf("fp64 DFMA loop: %d SMs, %.3f ms, %.2f TFLOP/s (%.1f DFMA/clk/SM at 1.965 GHz)\n", sms, ms,
               flops / ms / 1e9, flops / 2 / (ms * 1e-3) / sms / 1.965e9);
    }
    return 0;
}

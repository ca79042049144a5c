// l2_random.cu -- peak rate of RANDOM 8-byte reads served by L2: the bound of the Metropolis resampler
// (metropolis_kernel: B dependent random reads of the weight vector per particle, src/samplers.cpp:21-35).
// Each read pulls one 32-byte sector out of L2; the weight vector (8 MB at N = 10^6) is L2-resident.
// The loop is the resampler's access pattern without its arithmetic: an LCG picks the next index, the
// loaded value feeds the following index (dependent, like k = j after an accept), 4 chains per thread.
#include <cstdio>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(256) chase(const double *__restrict__ w, unsigned n, int iters, double *out)
{
    unsigned s[4];
    double acc[4] = {0, 0, 0, 0};
#pragma unroll
    for (int c = 0; c < 4; ++c) s[c] = (blockIdx.x * 256 + threadIdx.x) * 4 + c + 1;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            s[c] = s[c] * 1664525u + 1013904223u;
            const unsigned idx = (unsigned)(((unsigned long long)s[c] * n) >> 32);
            const double v = __ldg(w + idx);
            acc[c] += v;
            s[c] += (unsigned)(v > 2.0);          // data dependence without changing the stream (v in [0, 1))
        }
    }
    out[blockIdx.x * 256 + threadIdx.x] = acc[0] + acc[1] + acc[2] + acc[3];
}

int main()
{
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    for (unsigned n : {1000000u, 8u << 20}) {
        double *w, *out;
        cudaMalloc(&w, sizeof(double) * n);
        cudaMemset(w, 0, sizeof(double) * n);
        const int blocks = sms * 8, iters = 400;
        cudaMalloc(&out, sizeof(double) * blocks * 256);
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        float best = 1e30f;
        for (int rep = 0; rep < 4; ++rep) {
            cudaEventRecord(e0);
            chase<<<blocks, 256>>>(w, n, iters, out);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            if (rep) best = ms < best ? ms : best;
        }
        const double reads = 4.0 * iters * blocks * 256;
        printf("random 8-byte reads over %u doubles (%.0f MB): %.3f ms, %.3e reads/s = %.1f GB/s of 32-byte sectors\n", n,
               n * 8 / 1e6, best, reads / (best * 1e-3), reads * 32 / (best * 1e-3) / 1e9);
        cudaFree(w);
        cudaFree(out);
    }
    return 0;
}

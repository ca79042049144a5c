// mailbox.cuh -- scalar exchange between the ranks of a sharded filter through PEER MEMORY.
//
// Every rank owns a mailbox [T][3 cells][world] of 4-word entries (words 0..2 payload, word 3 flag)
// that its peers map with CUDA IPC.  Publishing = lane r of one warp stores this rank's payload
// into entry [cell][rank] of rank r's mailbox (a peer store over NVLink), fences at system scope,
// then stores the flag; waiting = lane r spins on entry [cell][r] of the rank's OWN mailbox (local
// memory the peers write).  The exchanges are FUSED into the compute kernels on either side of
// them -- block 0 publishes in its prologue / epilogue, every block that needs the result gates on
// the flags -- so a sharded step launches exactly the kernels a single-GPU step does: no collective
// library, no extra launches, no host round trip.
//
// Flags carry the run epoch, so a mailbox is never cleared.  Spins are bounded (2 s): on a timeout
// the error word is set and the kernel goes on (results void) instead of hanging the GPU.
#pragma once

#include "common.cuh"

enum { kCellMax = 0, kCellSums = 1, kCellBarrier = 2, kMailCells = 3 };

struct MailArgs {
    unsigned long long *peer[CUSMC_MAX_PEERS];   // every rank's mailbox (own one included)
    unsigned long long *err;
    unsigned long long epoch;
    int rank, world;                              // world <= 1: no exchange
};

__host__ __device__ inline size_t mail_cell(int t, int cell, int world)
{
    return ((size_t)t * kMailCells + (size_t)cell) * (size_t)world;
}

#ifdef __CUDACC__
__device__ __forceinline__ unsigned long long mail_now_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// Lanes 0 .. world-1 of ONE warp: store (w0, w1, w2) + flag into every rank's mailbox.
__device__ __forceinline__ void mail_publish(const MailArgs &m, size_t cell0, int lane, unsigned long long w0,
                                             unsigned long long w1, unsigned long long w2)
{
    if (lane < m.world) {
        volatile unsigned long long *dst = m.peer[lane] + (cell0 + m.rank) * 4;
        dst[0] = w0;
        dst[1] = w1;
        dst[2] = w2;
        __threadfence_system();
        dst[3] = m.epoch;
    }
}

// Lanes 0 .. world-1 of a warp: lane r returns rank r's payload once its flag is up (0 for the
// other lanes).
__device__ __forceinline__ void mail_wait(const MailArgs &m, size_t cell0, int lane, unsigned long long &w0,
                                          unsigned long long &w1, unsigned long long &w2)
{
    w0 = w1 = w2 = 0;
    if (lane < m.world) {
        volatile unsigned long long *src = m.peer[m.rank] + (cell0 + lane) * 4;
        if (src[3] != m.epoch) {
            const unsigned long long t0 = mail_now_ns();
            while (src[3] != m.epoch) {
                __nanosleep(32);
                if (mail_now_ns() - t0 > 2000000000ull) {
                    *m.err = 1;
                    break;
                }
            }
        }
        __threadfence_system();
        w0 = src[0];
        w1 = src[1];
        w2 = src[2];
    }
}

// Barrier fused into a kernel prologue: block 0 announces "everything this rank enqueued before
// this kernel is complete" (stream order guarantees it), every block waits for all ranks' flags.
// Call from all threads of the block; ends with __syncthreads().
__device__ __forceinline__ void mail_gate(const MailArgs &m, size_t cell0)
{
    if (m.world <= 1) return;
    if (threadIdx.x < 32) {
        if (blockIdx.x == 0) mail_publish(m, cell0, threadIdx.x, 0, 0, 0);
        unsigned long long a, b, c;
        mail_wait(m, cell0, threadIdx.x, a, b, c);
    }
    __syncthreads();
}
#endif

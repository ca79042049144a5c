// mailbox.cuh -- scalar exchange between the ranks of a sharded filter through PEER MEMORY.
//
// Every rank owns a mailbox [T][3 cells][world] of 4-word entries (words 0..2 payload, word 3 flag)
// that its peers map with CUDA IPC.  Publishing = lane r of one warp stores this rank's payload
// into entry [cell][rank] of rank r's mailbox (a peer store over NVLink), fences at system scope,
// then stores the flag; waiting = lane r spins on entry [cell][r] of the rank's OWN mailbox (local
// memory the peers write).  No collective library and no host round trip inside the time loop.
// The sums exchange is fused into the tail of the (single-block) tile-scan kernel; the max exchange
// and the barrier are one-warp kernels.  (Gating every block of the big kernels on the flags was
// tried: an extra dependent L2 round trip + system fence per block cost 100 us per step at 32 Ki
// blocks.)
//
// Flags carry the run epoch, so a mailbox is never cleared.  Spins are bounded (2 s by default,
// cusmc_filter_set_exchange_timeout): on a timeout the error word is set and the kernel goes on instead
// of hanging the GPU; every getter of the filter then returns CUSMC_ERR_TIMEOUT (results void).
#pragma once

#include "common.cuh"

enum { kCellMax = 0, kCellSums = 1, kCellBarrier = 2, kMailCells = 3 };

struct MailArgs {
    unsigned long long *const *peer;              // device table: every rank's mailbox (own one included)
    unsigned long long *err;
    unsigned long long timeout_ns;                // bound of a spin-wait (default 2 s)
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
                if (mail_now_ns() - t0 > m.timeout_ns) {
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

#endif

// mailbox.cuh -- scalar exchange between the ranks of a sharded filter through PEER MEMORY.
//
// Every rank owns a mailbox [T][3 cells][world] of 64-byte entries that its peers map with CUDA IPC.
// An entry carries three 64-bit payload words as SIX 8-byte units, each unit = 32 bits of data + the
// 32-bit run epoch (the flag travels INSIDE every store, the scheme of NCCL's low-latency protocol):
// publishing = lane r of one warp stores the six units into entry [cell][rank] of rank r's mailbox (peer
// stores over NVLink); waiting = lane r polls entry [cell][r] of the rank's OWN mailbox until all six
// epochs match.  8-byte stores are single-copy atomic, so there is no payload / flag ordering to
// enforce and no fence between them; a device-scope fence BEFORE the stores puts the publisher's earlier
// writes (its tile fields) into L2, where the peers' NVLink loads are served from.  Measured on 2 GPUs
// (profiles/shard_trace.py): update + both exchanges 28 us with payload / system fence / flag, 26 with
// flag-in-data units, 19 without the two system-scope fences, against 8 us for the update alone.
// No collective library and no host round trip inside the time loop.  The exchanges ride inside the
// one-block tile-update kernel; reference-mode runs use one-warp exchange kernels as barriers.
//
// Epochs make clearing unnecessary.  Spins are bounded (2 s by default,
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

constexpr int kMailWords = 8;     // uint64 words per mailbox entry (six used)

__device__ __forceinline__ unsigned long long mail_unit(unsigned int data, unsigned long long epoch)
{
    return (unsigned long long)data | (epoch << 32);
}

// Lanes 0 .. world-1 of ONE warp: store (w0, w1, w2), flag included, into every rank's mailbox.
__device__ __forceinline__ void mail_publish(const MailArgs &m, size_t cell0, int lane, unsigned long long w0,
                                             unsigned long long w1, unsigned long long w2)
{
    if (lane < m.world) {
        // This rank's earlier writes (its tile fields) must be in L2 -- where the peers' NVLink loads are
        // served from -- before a unit is even issued: a DEVICE-scope fence.  (A system-scope one costs
        // 1.8 us per exchange here and buys nothing: the only remote stores are the units themselves.)
        __threadfence();
        volatile unsigned long long *dst = m.peer[lane] + (cell0 + m.rank) * kMailWords;
        const unsigned long long w[3] = {w0, w1, w2};
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            dst[2 * k] = mail_unit((unsigned int)w[k], m.epoch);
            dst[2 * k + 1] = mail_unit((unsigned int)(w[k] >> 32), m.epoch);
        }
    }
}

// Lanes 0 .. world-1 of a warp: lane r returns rank r's payload once all its units carry the epoch (0
// for the other lanes).
__device__ __forceinline__ void mail_wait(const MailArgs &m, size_t cell0, int lane, unsigned long long &w0,
                                          unsigned long long &w1, unsigned long long &w2)
{
    w0 = w1 = w2 = 0;
    if (lane < m.world) {
        volatile unsigned long long *src = m.peer[m.rank] + (cell0 + lane) * kMailWords;
        const unsigned long long want = m.epoch & 0xffffffffull;
        unsigned long long u[6];
        unsigned long long t0 = 0;
        unsigned spins = 0;
        for (;;) {
            bool ok = true;
#pragma unroll
            for (int k = 0; k < 6; ++k) {
                u[k] = src[k];
                ok = ok && (u[k] >> 32) == want;
            }
            if (ok) break;
            if ((++spins & 15u) != 0) continue;          // the global timer is slow to read: every 16th poll
            if (t0 == 0) t0 = mail_now_ns();
            if (mail_now_ns() - t0 > m.timeout_ns) {
                *m.err = 1;
                break;
            }
        }
        // no acquire fence: the waiting warp only consumes the payload it has just read; the peers' bulk
        // data is read by the NEXT kernel, which the kernel boundary orders after this one
        w0 = (u[0] & 0xffffffffull) | (u[1] << 32);
        w1 = (u[2] & 0xffffffffull) | (u[3] << 32);
        w2 = (u[4] & 0xffffffffull) | (u[5] << 32);
    }
}

#endif

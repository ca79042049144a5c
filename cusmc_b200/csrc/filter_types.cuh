// filter_types.cuh -- the filter object and its per-step device record, shared by the host loop
// (filter.cu) and the persistent whole-run kernel (pf_persist.cu).
#pragma once

#include "density.cuh"

#include "../../include/cusmc_philox.h"

#include <vector>

struct StepSlot {            // one per time step, on the device (64 bytes = 8 words)
    double lw_max;           // [0] max log-weight (log modes), -inf initialised
    uint64_t sum_q, sum_q2, n_pos;   // [1..3] fixed-point sums (weigh_kernel); global after the exchange
    uint64_t cdf_offset;     // [4] fixed-point mass held by lower-ranked shards (0 on one GPU)
    uint64_t resampled;      // [5] 1 if this step drew new ancestors (adaptive resampling), else 0
    uint64_t degenerate;     // [6] 1 if the weights this step resampled FROM had no mass (all -inf / NaN):
                             //     its ancestors are the identity and the run reports CUSMC_ERR_DEGENERATE
    double reserved;
};

struct cusmc_filter {
    cusmc_ctx *ctx = nullptr;
    cusmc_filter_config cfg{};
    cusmc_filter_draws draws{};
    std::vector<double> Y, m0, C0, F, G, V, W;       // host copies (column-major)
    std::vector<double> Qc0, Qw;                      // noise factors
    std::vector<double> M, Winv;                      // observation operator
    Epilogue ep{};
    int is_log = 1;
    int shift = 0;
    // sharding: this rank owns the global slots lo .. lo + n - 1; every rank allocates `per` columns
    int world = 1, rank = 0;
    int64_t per = 0, lo = 0, n = 0;
    bool attached = false;
    bool pooled = false;                  // device buffers come from the stream-ordered pool (single-GPU filters)
    CusmcPeers peer_x[2]{}, peer_anc{}, peer_lw{}, peer_mail{}, peer_img[2]{};
    unsigned long long *mail = nullptr;   // [T][3 phases][world] x 4 words, written by the peers
    unsigned long long *mail_err = nullptr;   // 1 word: a spin-wait timed out
    unsigned long long mail_timeout_ns = 2000000000ull;   // bound of every spin-wait (cusmc_filter_set_exchange_timeout)
    void **peer_tables = nullptr;             // device: [CUSMC_FILTER_IPC_BUFFERS][CUSMC_MAX_PEERS] peer pointers
    unsigned long long *img[2] = {nullptr, nullptr};   // weight images by step parity (image.cuh; log modes)
    unsigned long long *rank_sums = nullptr;  // NCCL formulation: all-gathered per-rank sums of the current step
    unsigned long long epoch = 0;         // flag value of the current run (mail is never cleared)
    bool fused = false;                   // inside cusmc_filter_run_sharded: exchanges ride in the kernels
    double *x[2] = {nullptr, nullptr};
    double *lw = nullptr;
    uint32_t *anc = nullptr;
    StepSlot *slots = nullptr;
    double *moments = nullptr;        // T x (2 + d)
    void *persist = nullptr;          // scratch of the persistent-kernel run (pf_persist.cu)
    size_t persist_bytes = 0;
    uint32_t persist_tile = 0;        // particles per block of a persistent run (0: not covered / does not fit)
    uint32_t tile = 2048;             // particles per tile of the per-step path (cfg.tile_size; kTile unless balanced)
    int64_t img_n = 0;                // particle count the weight images are laid out for (image.cuh)
    double *hist_x = nullptr, *hist_w = nullptr;
    uint32_t *hist_a = nullptr;
    // ring_K > 0: the history buffers hold a RING of 2 chunks of ring_K steps each instead of all T
    // steps; the chunk of step t is (t / ring_K) & 1.  cusmc_run streams finished chunks to the host
    // while the next ones compute, so a run with full history needs bounded device memory.
    int ring_K = 0;
    size_t hist_row(int t) const
    {
        return ring_K > 0 ? (size_t)((t / ring_K) & 1) * (size_t)ring_K + (size_t)(t % ring_K) : (size_t)t;
    }
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    double last_ms = 0.0;
    int cur = 0;
    int next_t = -1;                  // phase bookkeeping: the step cusmc_filter_propagate expects
    bool ran = false;
};


// ---- host helpers shared by filter.cu and pf_persist.cu --------------------------------------------
// c = L_V^-1 y (Winv row-major lower triangular)
inline void cusmc_whiten_observation(const std::vector<double> &Winv, int dy, const double *y, double *c)
{
    for (int k = 0; k < dy; ++k) {
        double s = 0.0;
        for (int i = 0; i <= k; ++i) s += Winv[(size_t)k * dy + i] * y[i];
        c[k] = s;
    }
}

inline bool cusmc_is_diag_colmajor(const double *A, int d)
{
    if (!A) return true;
    for (int c = 0; c < d; ++c)
        for (int r = 0; r < d; ++r)
            if (r != c && A[(size_t)c * d + r] != 0.0) return false;
    return true;
}

// The systematic offset of step t when the caller injects none: 64 Philox bits keyed by (seed, step).
inline uint64_t cusmc_u0_bits(uint64_t seed, uint64_t step)
{
    const cusmc_u32x4 r = cusmc_rng(seed, 7 /* systematic offset */, step, 0, 0);
    return ((uint64_t)r.v[0] << 32) | r.v[1];
}

// filter.cu: every StepSlot back to { -inf, 0, 0, ... } on the stream
int cusmc_filter_init_slots(cusmc_filter *f);

// pf_persist.cu: the whole run as ONE cooperative kernel when the configuration allows it
// (returns CUSMC_ERR_UNSUPPORTED otherwise, without side effects).
int cusmc_filter_run_persistent(cusmc_filter *f, const cusmc_filter_draws *draws);
bool cusmc_filter_persistent_eligible(const cusmc_filter *f, const cusmc_filter_draws *draws);
uint32_t cusmc_filter_persistent_tile(cusmc_filter *f);

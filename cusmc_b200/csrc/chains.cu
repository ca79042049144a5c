// chains.cu -- per-chain / per-point covariance paths (BASELINE.json configs[2]):
//
//  * mh_chains_kernel: independent random-walk Metropolis-Hastings chains on an MVN / MVT
//    target, one warp per chain, lane k owns component k.  The proposal x' = x + s L z uses the
//    target's own factor, so the chain runs in whitened coordinates v = L^-1 (x - mu): a step is
//    v' = v + s z, q' = |v'|^2 (one warp sum) and a transcendental-free accept comparison; the
//    factor is used to whiten the start and un-whiten the result (and for running sums of x).
//  * perpoint_kernel: log-density with one covariance per point; each warp pulls its point's
//    packed factor into shared memory with a 1-D TMA bulk copy (cp.async.bulk + mbarrier,
//    double buffered) and runs the same forward substitution.
//
// The density arithmetic restates src/statistics.cc.cpp:171-196,295-324 in whitened form; the
// chains themselves have no counterpart in the reference -- the oracle
// (orc_mh_chains) fixes the operation order these kernels reproduce bit for bit.
#include "common.cuh"
#include "density.cuh"
#include "hostmath.h"

#include "../../include/cusmc_detmath.h"
#include "../../include/cusmc_philox.h"

namespace {

constexpr int kWarpsPerBlock = 4;
#ifndef CUSMC_CHAINS_DUAL
#define CUSMC_CHAINS_DUAL 1
#endif
#ifndef CUSMC_GENERAL_DUAL
#define CUSMC_GENERAL_DUAL 1
#endif

struct ChainArgs {
    const double *mu, *L, *z, *thr;
    double *x, *sum_x, *sum_xx;
    uint32_t *n_accept;
    uint8_t *accept_bits;
    int64_t C;
    uint64_t seed;
    double step_size, nu;
    int d, steps, kind, shared;
};

// Sum over the W lanes of a chain by xor butterflies (W/2, ..., 2, 1): every lane of the chain ends
// with the same bits.  Lanes that hold no component contribute +0, so the value equals the 32-lane
// butterfly (16, 8, 4, 2, 1) a host reproduces with the same pairing (oracle: butterfly32).
template <int W>
__device__ __forceinline__ double chain_sum_butterfly(double v)
{
#pragma unroll
    for (int o = W / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// The chain runs in WHITENED coordinates.  The proposal x' = x + s L z uses the target's own factor,
// so with v = L^-1 (x - mu) it is simply v' = v + s z and the target's quadratic form is |v'|^2:
// a step is one FMA and one sum over the chain's lanes -- no mat-vec, no forward substitution, and
// the factor is needed only to whiten the start, to un-whiten the result and (MOMENTS) for the
// running sums of x, so the step loop does not keep it in registers.  Same law as the x-space
// chain, a fraction of the work; the oracle (orc_mh_chains) restates exactly this arithmetic.
//
// A chain owns W = max(D, 4) lanes (D = d padded to a power of two; a quad at least, because four
// lanes share a Philox block of normals), a warp G = 32 / W chains: at d = 8 four chains advance in
// the instruction stream one used to take, and no lane idles.
// FAST (with PHILOX): the throughput generator of the filter kernels for the proposal normals -- Philox4x32-7 and
// the special-function-unit Box-Muller (cusmc_box_muller_fast: ~12 instead of ~67 instructions per pair); the
// thresholds keep the exact logarithm.  Same law; a host cannot mirror its last bits (cusmc_ctx_set_chain_noise).
template <int D, bool PHILOX, bool MOMENTS, bool FAST = false>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
mh_chains_kernel(const ChainArgs a)
{
    constexpr int W = D < 4 ? 4 : D;                 // lanes per chain
    constexpr int G = 32 / W;                        // chains per warp
    __shared__ __align__(16) double s_v[kWarpsPerBlock][32];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int grp = lane / W, sub = lane % W;
    const int64_t c0 = ((int64_t)blockIdx.x * kWarpsPerBlock + wib) * G;
    if (c0 >= a.C) return;                           // the whole warp
    const int d = a.d;
    const bool active = c0 + grp < a.C;              // chains past the end ride along on zeros
    const int64_t c = active ? c0 + grp : a.C - 1;   // (a valid address for their guarded loads)
    const bool live = active && sub < d;
    const double *Lc = a.shared ? a.L : a.L + (size_t)c * d * d;
    const double *mc = a.shared ? a.mu : a.mu + (size_t)c * d;
    const double mu = live ? __ldg(mc + sub) : 0.0;
    // Row k of the factor (column-major storage, so element j of every lane's row is one coalesced
    // line).  Without MOMENTS the rows are streamed from memory where they are used -- when the start
    // is whitened and when the result is un-whitened -- in chunks of 8 columns, so neither the step
    // loop nor those two passes hold 2 D registers of factor; MOMENTS needs x every step and keeps
    // the row in registers.
    auto Lkj = [&](const double *Lp, int j) { return (live && j <= sub) ? __ldg(Lp + (size_t)j * d + sub) : 0.0; };
    double row_m[MOMENTS ? D : 1];
    if (MOMENTS) {
#pragma unroll
        for (int j = 0; j < D; ++j) row_m[j] = Lkj(Lc, j);
    }
    // x = mu + L v: v broadcast through shared memory, row k = sum_j L[k][j] v_j, j ascending
    auto unwhiten = [&](const double *Lp, double v) {
        __syncwarp();
        s_v[wib][lane] = v;
        __syncwarp();
        const double *vc = &s_v[wib][grp * W];
        double acc = 0.0;
        if constexpr (MOMENTS) {
#pragma unroll
            for (int j = 0; j < D; ++j) acc = fma(row_m[j], vc[j], acc);
        } else {
#pragma unroll 8
            for (int j = 0; j < D; ++j) acc = fma(Lkj(Lp, j), vc[j], acc);
        }
        return mu + acc;
    };

    // whiten the start: column-oriented forward substitution, lane k keeps v_k
    double v = 0.0;
    const double x_start = live ? a.x[(size_t)c * d + sub] : 0.0;
    {
        const double rinv = live ? 1.0 / __ldg(Lc + (size_t)sub * d + sub) : 0.0;
        double r = live ? x_start - mu : 0.0;
#pragma unroll 8
        for (int j = 0; j < D; ++j) {
            const double vj = __shfl_sync(0xffffffffu, r * rinv, j, W);
            if (j == sub) v = vj;
            r = fma(-Lkj(Lc, j), vj, r);   // no-op for lanes k < j (L[k][j] == 0); lane j is done with r
        }
    }
    double q = chain_sum_butterfly<W>(v * v);
    const double inv_nu = a.kind == CUSMC_MVT ? 1.0 / a.nu : 0.0;
    double sx = 0.0, sxx = 0.0;
    uint32_t nacc = 0;

    const double *zc = a.z ? a.z + (size_t)c * a.steps * d : nullptr;
    double z_next = (zc && live) ? ld_stream(zc + sub) : 0.0;
    // In-kernel randomness is generated in batches so that no Philox block is computed twice and
    // none of its output is thrown away; the (seed, chain, step, component) -> draw mapping is the
    // one of cusmc_philox.h, unchanged:
    //  * thresholds: every W steps lane l of the chain computes the threshold of step s + l (one
    //    Philox block, one log, one exp per lane per W steps instead of per step), broadcast per step;
    //  * normals: lanes 4m .. 4m+3 of the chain share the block (step, m).  Every 4 steps lane 4m+k
    //    computes the block of step s + k -- four normals, one for each lane of its quad -- and a
    //    4 x 4 exchange inside the quad hands every lane its own normal of steps s .. s + 3.
    double thr_batch = 0.0;
    float zq[4] = {0.f, 0.f, 0.f, 0.f};    // zq[r]: this lane's normal of step s0 + ((lane & 3) ^ r)
    for (int s = 0; s < a.steps; ++s) {
        double z, thr;
        if (PHILOX) {
            if ((s & (W - 1)) == 0) {
                const cusmc_u32x4 r = cusmc_rng(a.seed, CUSMC_STREAM_CHAIN_U, (uint64_t)(s + sub), (uint64_t)c, 0);
                const double e = -cusmc_det_log(cusmc_u01_open0(r.v[0], r.v[1]));
                thr_batch = a.kind == CUSMC_MVT ? cusmc_det_exp((e + e) / (a.nu + (double)d)) : e;
            }
            thr = __shfl_sync(0xffffffffu, thr_batch, s & (W - 1), W);
            if ((s & 3) == 0) {
                float n[4];
                if constexpr (FAST) {
                    const cusmc_u32x4 rz = cusmc_rng7(a.seed, CUSMC_STREAM_CHAIN_Z, (uint64_t)(s + (sub & 3)), (uint64_t)c,
                                                      (uint32_t)(sub >> 2));
                    cusmc_box_muller_fast(rz.v[0], rz.v[1], &n[0], &n[1]);
                    cusmc_box_muller_fast(rz.v[2], rz.v[3], &n[2], &n[3]);
                } else {
                    const cusmc_u32x4 rz = cusmc_rng(a.seed, CUSMC_STREAM_CHAIN_Z, (uint64_t)(s + (sub & 3)), (uint64_t)c,
                                                     (uint32_t)(sub >> 2));
                    cusmc_box_muller_f32(rz.v[0], rz.v[1], &n[0], &n[1]);
                    cusmc_box_muller_f32(rz.v[2], rz.v[3], &n[2], &n[3]);
                }
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const int pick = (lane & 3) ^ r;        // the normal owed to lane ^ r
                    const float val = pick == 0 ? n[0] : pick == 1 ? n[1] : pick == 2 ? n[2] : n[3];
                    zq[r] = __shfl_xor_sync(0xffffffffu, val, r);
                }
            }
            const int r = (lane & 3) ^ (s & 3);
            const float zf = r == 0 ? zq[0] : r == 1 ? zq[1] : r == 2 ? zq[2] : zq[3];
            z = live ? (double)zf : 0.0;
        } else {
            z = z_next;
            if (s + 1 < a.steps && live) z_next = ld_stream(zc + (size_t)(s + 1) * d + sub);   // prefetch
            thr = __ldg(a.thr + (size_t)c * a.steps + s);
        }
        // proposal and its quadratic form, in whitened coordinates
        const double vp = fma(a.step_size, z, v);
        const double qp = chain_sum_butterfly<W>(vp * vp);
        bool accept;
        if (a.kind == CUSMC_MVT)
            accept = fma(qp, inv_nu, 1.0) < thr * fma(q, inv_nu, 1.0);
        else
            accept = 0.5 * (qp - q) < thr;
        if (accept) {
            v = vp;
            q = qp;
            ++nacc;
        }
        if (MOMENTS) {
            // (the un-whitening synchronises the warp: every chain of it takes this path every step)
            const double xu = unwhiten(Lc, v);
            const double x = nacc ? xu : x_start;
            sx += x;
            sxx = fma(x, x, sxx);
        }
        if (a.accept_bits && sub == 0 && active) a.accept_bits[(size_t)c * a.steps + s] = (uint8_t)accept;
    }
    const double xu = unwhiten(Lc, v);
    const double x = nacc ? xu : x_start;                  // a chain that never moved is left untouched
    if (live) {
        a.x[(size_t)c * d + sub] = x;
        if (MOMENTS) {
            if (a.sum_x) a.sum_x[(size_t)c * d + sub] = sx;
            if (a.sum_xx) a.sum_xx[(size_t)c * d + sub] = sxx;
        }
    }
    if (a.n_accept && sub == 0 && active) a.n_accept[c] = nacc;
}

// ---- the whitened chains at 16 < d <= 32: two components per lane, two chains per warp ----------------------
// mh_chains_kernel<32> spends a warp on one chain and most of a step on the 32-lane butterfly of the quadratic
// form (five double shuffles) and the accept logic -- work per WARP, not per component.  With components l and
// l + 16 in lane l of a 16-lane group the butterfly's first stage (lanes l and l ^ 16) becomes a local addition
// of the same two squares, four shuffle stages remain, and every warp-wide instruction of the step serves two
// chains.  Same pairing, same (seed, chain, step, component) -> draw mapping: bit-identical to the one-row kernel
// and to the oracle (butterfly32).  Running moments need the factor's rows in registers every step and stay on
// the one-row kernel.
template <bool PHILOX, bool FAST>
__global__ void __launch_bounds__(kWarpsPerBlock * 32, 8)
mh_chains_dual_kernel(const ChainArgs a)
{
    constexpr int W = 16, H = 16, G = 2;
    __shared__ __align__(16) double s_v[kWarpsPerBlock][G][2 * H];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int grp = lane / W, sub = lane % W;
    const int64_t c0 = ((int64_t)blockIdx.x * kWarpsPerBlock + wib) * G;
    if (c0 >= a.C) return;
    const int d = a.d;                               // 16 < d <= 32
    const bool active = c0 + grp < a.C;
    const int64_t c = active ? c0 + grp : a.C - 1;
    const bool live0 = active, live1 = active && sub + H < d;
    const double *Lc = a.shared ? a.L : a.L + (size_t)c * d * d;
    const double *mc = a.shared ? a.mu : a.mu + (size_t)c * d;
    const double mu0 = live0 ? __ldg(mc + sub) : 0.0, mu1 = live1 ? __ldg(mc + sub + H) : 0.0;
    // L[k][j], j <= k, of this lane's two rows (column-major storage), streamed where they are used
    auto LA = [&](int j) { return (live0 && j <= sub) ? __ldg(Lc + (size_t)j * d + sub) : 0.0; };
    auto LB = [&](int j) { return (live1 && j <= sub + H) ? __ldg(Lc + (size_t)j * d + sub + H) : 0.0; };
    // x = mu + L v, row k = sum_j L[k][j] v_j with j ascending (the rows' zeros beyond the diagonal are skipped:
    // they add exact zeros in the one-row kernel)
    auto unwhiten = [&](double va, double vb, double &xa, double &xb) {
        __syncwarp();
        s_v[wib][grp][sub] = va;
        s_v[wib][grp][sub + H] = vb;
        __syncwarp();
        const double *vc = s_v[wib][grp];
        double acc0 = 0.0, acc1 = 0.0;
#pragma unroll 8
        for (int j = 0; j < H; ++j) {
            acc0 = fma(LA(j), vc[j], acc0);
            acc1 = fma(LB(j), vc[j], acc1);
        }
#pragma unroll 8
        for (int j = H; j < 2 * H; ++j) acc1 = fma(LB(j), vc[j], acc1);
        xa = mu0 + acc0;
        xb = mu1 + acc1;
    };

    // whiten the start: column-oriented forward substitution over the 32 columns
    double v0 = 0.0, v1 = 0.0;
    const double xs0 = live0 ? a.x[(size_t)c * d + sub] : 0.0, xs1 = live1 ? a.x[(size_t)c * d + sub + H] : 0.0;
    {
        const double rinv0 = live0 ? 1.0 / __ldg(Lc + (size_t)sub * d + sub) : 0.0;
        const double rinv1 = live1 ? 1.0 / __ldg(Lc + (size_t)(sub + H) * d + sub + H) : 0.0;
        double ra = live0 ? xs0 - mu0 : 0.0, rb = live1 ? xs1 - mu1 : 0.0;
#pragma unroll 8
        for (int j = 0; j < H; ++j) {
            const double vj = __shfl_sync(0xffffffffu, ra * rinv0, j, W);
            if (j == sub) v0 = vj;
            ra = fma(-LA(j), vj, ra);
            rb = fma(-LB(j), vj, rb);
        }
#pragma unroll 8
        for (int j = 0; j < H; ++j) {
            const double vj = __shfl_sync(0xffffffffu, rb * rinv1, j, W);
            if (j == sub) v1 = vj;
            rb = fma(-LB(H + j), vj, rb);
        }
    }
    // the 32-lane butterfly (16, 8, 4, 2, 1) of the one-row kernel: stage 16 pairs components l and l + 16
    auto sumsq = [&](double a0, double a1) { return chain_sum_butterfly<W>(a0 * a0 + a1 * a1); };
    double q = sumsq(v0, v1);
    const double inv_nu = a.kind == CUSMC_MVT ? 1.0 / a.nu : 0.0;
    uint32_t nacc = 0;

    const double *zc = a.z ? a.z + (size_t)c * a.steps * d : nullptr;
    double zn0 = (zc && live0) ? ld_stream(zc + sub) : 0.0, zn1 = (zc && live1) ? ld_stream(zc + sub + H) : 0.0;
    double thr_batch = 0.0;
    float zq0[4] = {0.f, 0.f, 0.f, 0.f}, zq1[4] = {0.f, 0.f, 0.f, 0.f};
    auto quad_normals = [&](int s, uint32_t m, float (&zq)[4]) {
        float n[4];
        if constexpr (FAST) {
            const cusmc_u32x4 rz = cusmc_rng7(a.seed, CUSMC_STREAM_CHAIN_Z, (uint64_t)(s + (sub & 3)), (uint64_t)c, m);
            cusmc_box_muller_fast(rz.v[0], rz.v[1], &n[0], &n[1]);
            cusmc_box_muller_fast(rz.v[2], rz.v[3], &n[2], &n[3]);
        } else {
            const cusmc_u32x4 rz = cusmc_rng(a.seed, CUSMC_STREAM_CHAIN_Z, (uint64_t)(s + (sub & 3)), (uint64_t)c, m);
            cusmc_box_muller_f32(rz.v[0], rz.v[1], &n[0], &n[1]);
            cusmc_box_muller_f32(rz.v[2], rz.v[3], &n[2], &n[3]);
        }
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int pick = (lane & 3) ^ r;
            const float val = pick == 0 ? n[0] : pick == 1 ? n[1] : pick == 2 ? n[2] : n[3];
            zq[r] = __shfl_xor_sync(0xffffffffu, val, r);
        }
    };
    for (int s = 0; s < a.steps; ++s) {
        double z0, z1, thr;
        if (PHILOX) {
            if ((s & (W - 1)) == 0) {
                const cusmc_u32x4 r = cusmc_rng(a.seed, CUSMC_STREAM_CHAIN_U, (uint64_t)(s + sub), (uint64_t)c, 0);
                const double e = -cusmc_det_log(cusmc_u01_open0(r.v[0], r.v[1]));
                thr_batch = a.kind == CUSMC_MVT ? cusmc_det_exp((e + e) / (a.nu + (double)d)) : e;
            }
            thr = __shfl_sync(0xffffffffu, thr_batch, s & (W - 1), W);
            if ((s & 3) == 0) {
                quad_normals(s, (uint32_t)(sub >> 2), zq0);
                quad_normals(s, (uint32_t)(sub >> 2) + (uint32_t)(H / 4), zq1);
            }
            const int r = (lane & 3) ^ (s & 3);
            const float f0 = r == 0 ? zq0[0] : r == 1 ? zq0[1] : r == 2 ? zq0[2] : zq0[3];
            const float f1 = r == 0 ? zq1[0] : r == 1 ? zq1[1] : r == 2 ? zq1[2] : zq1[3];
            z0 = live0 ? (double)f0 : 0.0;
            z1 = live1 ? (double)f1 : 0.0;
        } else {
            z0 = zn0;
            z1 = zn1;
            if (s + 1 < a.steps) {
                if (live0) zn0 = ld_stream(zc + (size_t)(s + 1) * d + sub);
                if (live1) zn1 = ld_stream(zc + (size_t)(s + 1) * d + sub + H);
            }
            thr = __ldg(a.thr + (size_t)c * a.steps + s);
        }
        const double vp0 = fma(a.step_size, z0, v0), vp1 = fma(a.step_size, z1, v1);
        const double qp = sumsq(vp0, vp1);
        bool accept;
        if (a.kind == CUSMC_MVT)
            accept = fma(qp, inv_nu, 1.0) < thr * fma(q, inv_nu, 1.0);
        else
            accept = 0.5 * (qp - q) < thr;
        if (accept) {
            v0 = vp0;
            v1 = vp1;
            q = qp;
            ++nacc;
        }
        if (a.accept_bits && sub == 0 && active) a.accept_bits[(size_t)c * a.steps + s] = (uint8_t)accept;
    }
    double xu0, xu1;
    unwhiten(v0, v1, xu0, xu1);
    if (live0) a.x[(size_t)c * d + sub] = nacc ? xu0 : xs0;            // a chain that never moved is left untouched
    if (live1) a.x[(size_t)c * d + sub + H] = nacc ? xu1 : xs1;
    if (a.n_accept && sub == 0 && active) a.n_accept[c] = nacc;
}

// ---- general random walk: the proposal does NOT use the target's factor ---------------------------
// x' = x + s (.) z  (isotropic step, optionally a per-component scale): nothing cancels any more, so
// EVERY step evaluates the target density -- the north star's "proposal, log-acceptance ratio and
// accept / reject fused into the same kernel as the density evaluation" (a2's arithmetic,
// src/statistics.cc.cpp:295-311, under the accept rule of src/samplers.cpp:30):
//
//     r = x' - mu,   v = L_c^-1 r  (forward substitution),   q' = |v|^2,   accept on (q', q)
//
// Row k of the chain's factor stays in lane k's REGISTERS for the whole run (d doubles per lane; the
// factor is read from memory once per chain, not once per step: 264 B instead of 4.4 KB per step at
// d = 32).  The substitution is column oriented inside the chain's W lanes: step j broadcasts
// v_j by shuffle, every lane k > j subtracts (L_kj / L_kk) v_j; every lane accumulates the same
// q' = sum_j v_j^2 in the same order, so no reduction follows.  oracle: orc_mh_chains_general (same
// operation order: bit-exact decisions and states).
struct GeneralArgs {
    ChainArgs c;
    const double *scale;          // optional per-component proposal scale (d doubles, shared)
};

template <int D, bool PHILOX, bool MOMENTS, bool FAST = false>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
mh_general_kernel(const GeneralArgs ga)
{
    const ChainArgs &a = ga.c;
    constexpr int W = D < 4 ? 4 : D;                 // lanes per chain
    constexpr int G = 32 / W;                        // chains per warp
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int grp = lane / W, sub = lane % W;
    const int64_t c0 = ((int64_t)blockIdx.x * kWarpsPerBlock + wib) * G;
    if (c0 >= a.C) return;                           // the whole warp
    const int d = a.d;
    const bool active = c0 + grp < a.C;              // chains past the end ride along on zeros
    const int64_t c = active ? c0 + grp : a.C - 1;
    const bool live = active && sub < d;
    const double *Lc = a.shared ? a.L : a.L + (size_t)c * d * d;
    const double *mc = a.shared ? a.mu : a.mu + (size_t)c * d;
    const double mu = live ? __ldg(mc + sub) : 0.0;
    // row k of the factor, PRE-SCALED by 1 / L_kk (column-major storage: element j of every lane's row is
    // one coalesced line).  With r~_k = r_k / L_kk carried instead of r_k the substitution's dependent
    // chain is shuffle -> FMA per column, the division's multiply is off it (1.9e9 -> 2.1e9 chain-steps/s).
    const double rinv = live ? 1.0 / __ldg(Lc + (size_t)sub * d + sub) : 0.0;
    // (Rows in shared memory instead -- column j of all lanes contiguous, 72 registers, 24 instead of 16 warps per
    // SM -- were measured: 2.05e9 against 2.22e9 chain-steps/s.  The step is not short of warps: its 64 SHFL per
    // step (two per broadcast double) keep the shuffle / shared-memory pipe half busy by themselves, and the
    // extra LDS go through the same pipe.)
    double row[D];
#pragma unroll
    for (int j = 0; j < D; ++j) row[j] = (live && j < sub) ? __ldg(Lc + (size_t)j * d + sub) * rinv : 0.0;
    const double s_k = a.step_size * ((ga.scale && live) ? __ldg(ga.scale + sub) : 1.0);

    // q = |L^-1 (x - mu)|^2, identical in every lane of the chain
    auto quadform = [&](double xk) {
        double r = live ? (xk - mu) * rinv : 0.0, q = 0.0;
#pragma unroll
        for (int j = 0; j < D; ++j) {
            const double vj = __shfl_sync(0xffffffffu, r, j, W);
            q = fma(vj, vj, q);
            r = fma(-row[j], vj, r);          // no-op for lanes k <= j (row[j] == 0); lane j is done with r
        }
        return q;
    };
    double x = live ? a.x[(size_t)c * d + sub] : 0.0;
    double q = quadform(x);
    const double inv_nu = a.kind == CUSMC_MVT ? 1.0 / a.nu : 0.0;
    double sx = 0.0, sxx = 0.0;
    uint32_t nacc = 0;

    const double *zc = a.z ? a.z + (size_t)c * a.steps * d : nullptr;
    double z_next = (zc && live) ? ld_stream(zc + sub) : 0.0;
    // in-kernel randomness: batched exactly as in mh_chains_kernel (same (seed, chain, step, component)
    // -> draw mapping of cusmc_philox.h)
    double thr_batch = 0.0;
    float zq[4] = {0.f, 0.f, 0.f, 0.f};
    for (int s = 0; s < a.steps; ++s) {
        double z, thr;
        if (PHILOX) {
            if ((s & (W - 1)) == 0) {
                const cusmc_u32x4 r = cusmc_rng(a.seed, CUSMC_STREAM_CHAIN_U, (uint64_t)(s + sub), (uint64_t)c, 0);
                const double e = -cusmc_det_log(cusmc_u01_open0(r.v[0], r.v[1]));
                thr_batch = a.kind == CUSMC_MVT ? cusmc_det_exp((e + e) / (a.nu + (double)d)) : e;
            }
            thr = __shfl_sync(0xffffffffu, thr_batch, s & (W - 1), W);
            if ((s & 3) == 0) {
                float n[4];
                if constexpr (FAST) {
                    const cusmc_u32x4 rz = cusmc_rng7(a.seed, CUSMC_STREAM_CHAIN_Z, (uint64_t)(s + (sub & 3)), (uint64_t)c,
                                                      (uint32_t)(sub >> 2));
                    cusmc_box_muller_fast(rz.v[0], rz.v[1], &n[0], &n[1]);
                    cusmc_box_muller_fast(rz.v[2], rz.v[3], &n[2], &n[3]);
                } else {
                    const cusmc_u32x4 rz = cusmc_rng(a.seed, CUSMC_STREAM_CHAIN_Z, (uint64_t)(s + (sub & 3)), (uint64_t)c,
                                                     (uint32_t)(sub >> 2));
                    cusmc_box_muller_f32(rz.v[0], rz.v[1], &n[0], &n[1]);
                    cusmc_box_muller_f32(rz.v[2], rz.v[3], &n[2], &n[3]);
                }
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const int pick = (lane & 3) ^ r;
                    const float val = pick == 0 ? n[0] : pick == 1 ? n[1] : pick == 2 ? n[2] : n[3];
                    zq[r] = __shfl_xor_sync(0xffffffffu, val, r);
                }
            }
            const int r = (lane & 3) ^ (s & 3);
            const float zf = r == 0 ? zq[0] : r == 1 ? zq[1] : r == 2 ? zq[2] : zq[3];
            z = live ? (double)zf : 0.0;
        } else {
            z = z_next;
            if (s + 1 < a.steps && live) z_next = ld_stream(zc + (size_t)(s + 1) * d + sub);   // prefetch
            thr = __ldg(a.thr + (size_t)c * a.steps + s);
        }
        const double xp = fma(s_k, z, x);
        const double qp = quadform(xp);
        bool accept;
        if (a.kind == CUSMC_MVT)
            accept = fma(qp, inv_nu, 1.0) < thr * fma(q, inv_nu, 1.0);
        else
            accept = 0.5 * (qp - q) < thr;
        if (accept) {
            x = xp;
            q = qp;
            ++nacc;
        }
        if (MOMENTS) {
            sx += x;
            sxx = fma(x, x, sxx);
        }
        if (a.accept_bits && sub == 0 && active) a.accept_bits[(size_t)c * a.steps + s] = (uint8_t)accept;
    }
    if (live) {
        a.x[(size_t)c * d + sub] = x;
        if (MOMENTS) {
            if (a.sum_x) a.sum_x[(size_t)c * d + sub] = sx;
            if (a.sum_xx) a.sum_xx[(size_t)c * d + sub] = sxx;
        }
    }
    if (a.n_accept && sub == 0 && active) a.n_accept[c] = nacc;
}

// ---- the same chains at 16 < d <= 32: TWO rows per lane, two chains per warp ---------------------------------
// mh_general_kernel<32> gives a chain a whole warp: 32 broadcast rounds of two SHFL and two DFMA in which, on
// average, half the lanes multiply by a zero of the triangle -- its shuffle pipe is half busy with 16 warps per SM
// and more warps do not help (see the note there).  Here lane l of a 16-lane group owns components l and l + 16:
// rounds 0..15 broadcast the first residual (both rows updated), rounds 16..31 the second (only the second row has
// entries there), and ONE warp-wide shuffle serves two chains: half the shuffles and 5/8 of the FMAs per
// chain-step.  Same operations on the same values in the same order for every component (the skipped updates
// multiplied by an exact zero), same (seed, chain, step, component) -> draw mapping: bit-identical to the one-row
// kernel and to the oracle (orc_mh_chains_general).
template <bool PHILOX, bool MOMENTS, bool FAST>
__global__ void __launch_bounds__(kWarpsPerBlock * 32, 3)
mh_general_dual_kernel(const GeneralArgs ga)
{
    const ChainArgs &a = ga.c;
    constexpr int W = 16, H = 16;                    // lanes per chain; lane l: components l and l + H
    constexpr int G = 32 / W;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int grp = lane / W, sub = lane % W;
    const int64_t c0 = ((int64_t)blockIdx.x * kWarpsPerBlock + wib) * G;
    if (c0 >= a.C) return;
    const int d = a.d;                               // 16 < d <= 32
    const bool active = c0 + grp < a.C;
    const int64_t c = active ? c0 + grp : a.C - 1;
    const bool live0 = active, live1 = active && sub + H < d;
    const double *Lc = a.shared ? a.L : a.L + (size_t)c * d * d;
    const double *mc = a.shared ? a.mu : a.mu + (size_t)c * d;
    const double mu0 = live0 ? __ldg(mc + sub) : 0.0, mu1 = live1 ? __ldg(mc + sub + H) : 0.0;
    const double rinv0 = live0 ? 1.0 / __ldg(Lc + (size_t)sub * d + sub) : 0.0;
    const double rinv1 = live1 ? 1.0 / __ldg(Lc + (size_t)(sub + H) * d + sub + H) : 0.0;
    double rowA[H], rowB[2 * H];                     // rows sub and sub + H, pre-scaled by 1 / L_kk (column-major L)
#pragma unroll
    for (int j = 0; j < H; ++j) rowA[j] = (live0 && j < sub) ? __ldg(Lc + (size_t)j * d + sub) * rinv0 : 0.0;
#pragma unroll
    for (int j = 0; j < 2 * H; ++j) rowB[j] = (live1 && j < sub + H) ? __ldg(Lc + (size_t)j * d + sub + H) * rinv1 : 0.0;
    const double s0 = a.step_size * ((ga.scale && live0) ? __ldg(ga.scale + sub) : 1.0);
    const double s1 = a.step_size * ((ga.scale && live1) ? __ldg(ga.scale + sub + H) : 1.0);

    auto quadform = [&](double xa, double xb) {
        double ra = live0 ? (xa - mu0) * rinv0 : 0.0, rb = live1 ? (xb - mu1) * rinv1 : 0.0, q = 0.0;
#pragma unroll
        for (int j = 0; j < H; ++j) {
            const double vj = __shfl_sync(0xffffffffu, ra, j, W);
            q = fma(vj, vj, q);
            ra = fma(-rowA[j], vj, ra);
            rb = fma(-rowB[j], vj, rb);
        }
#pragma unroll
        for (int j = 0; j < H; ++j) {
            const double vj = __shfl_sync(0xffffffffu, rb, j, W);
            q = fma(vj, vj, q);
            rb = fma(-rowB[H + j], vj, rb);
        }
        return q;
    };
    double x0 = live0 ? a.x[(size_t)c * d + sub] : 0.0, x1 = live1 ? a.x[(size_t)c * d + sub + H] : 0.0;
    double q = quadform(x0, x1);
    const double inv_nu = a.kind == CUSMC_MVT ? 1.0 / a.nu : 0.0;
    double sx0 = 0.0, sxx0 = 0.0, sx1 = 0.0, sxx1 = 0.0;
    uint32_t nacc = 0;

    const double *zc = a.z ? a.z + (size_t)c * a.steps * d : nullptr;
    double zn0 = (zc && live0) ? ld_stream(zc + sub) : 0.0, zn1 = (zc && live1) ? ld_stream(zc + sub + H) : 0.0;
    // randomness batched as in mh_general_kernel; the block (step, m) holds the normals of components 4m .. 4m + 3,
    // so this lane's quad computes blocks m = sub / 4 and m + 4
    double thr_batch = 0.0;
    float zq0[4] = {0.f, 0.f, 0.f, 0.f}, zq1[4] = {0.f, 0.f, 0.f, 0.f};
    auto quad_normals = [&](int s, uint32_t m, float (&zq)[4]) {
        float n[4];
        if constexpr (FAST) {
            const cusmc_u32x4 rz = cusmc_rng7(a.seed, CUSMC_STREAM_CHAIN_Z, (uint64_t)(s + (sub & 3)), (uint64_t)c, m);
            cusmc_box_muller_fast(rz.v[0], rz.v[1], &n[0], &n[1]);
            cusmc_box_muller_fast(rz.v[2], rz.v[3], &n[2], &n[3]);
        } else {
            const cusmc_u32x4 rz = cusmc_rng(a.seed, CUSMC_STREAM_CHAIN_Z, (uint64_t)(s + (sub & 3)), (uint64_t)c, m);
            cusmc_box_muller_f32(rz.v[0], rz.v[1], &n[0], &n[1]);
            cusmc_box_muller_f32(rz.v[2], rz.v[3], &n[2], &n[3]);
        }
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int pick = (lane & 3) ^ r;
            const float val = pick == 0 ? n[0] : pick == 1 ? n[1] : pick == 2 ? n[2] : n[3];
            zq[r] = __shfl_xor_sync(0xffffffffu, val, r);
        }
    };
    for (int s = 0; s < a.steps; ++s) {
        double z0, z1, thr;
        if (PHILOX) {
            if ((s & (W - 1)) == 0) {
                const cusmc_u32x4 r = cusmc_rng(a.seed, CUSMC_STREAM_CHAIN_U, (uint64_t)(s + sub), (uint64_t)c, 0);
                const double e = -cusmc_det_log(cusmc_u01_open0(r.v[0], r.v[1]));
                thr_batch = a.kind == CUSMC_MVT ? cusmc_det_exp((e + e) / (a.nu + (double)d)) : e;
            }
            thr = __shfl_sync(0xffffffffu, thr_batch, s & (W - 1), W);
            if ((s & 3) == 0) {
                quad_normals(s, (uint32_t)(sub >> 2), zq0);
                quad_normals(s, (uint32_t)(sub >> 2) + (uint32_t)(H / 4), zq1);
            }
            const int r = (lane & 3) ^ (s & 3);
            const float f0 = r == 0 ? zq0[0] : r == 1 ? zq0[1] : r == 2 ? zq0[2] : zq0[3];
            const float f1 = r == 0 ? zq1[0] : r == 1 ? zq1[1] : r == 2 ? zq1[2] : zq1[3];
            z0 = live0 ? (double)f0 : 0.0;
            z1 = live1 ? (double)f1 : 0.0;
        } else {
            z0 = zn0;
            z1 = zn1;
            if (s + 1 < a.steps) {
                if (live0) zn0 = ld_stream(zc + (size_t)(s + 1) * d + sub);
                if (live1) zn1 = ld_stream(zc + (size_t)(s + 1) * d + sub + H);
            }
            thr = __ldg(a.thr + (size_t)c * a.steps + s);
        }
        const double xp0 = fma(s0, z0, x0), xp1 = fma(s1, z1, x1);
        const double qp = quadform(xp0, xp1);
        bool accept;
        if (a.kind == CUSMC_MVT)
            accept = fma(qp, inv_nu, 1.0) < thr * fma(q, inv_nu, 1.0);
        else
            accept = 0.5 * (qp - q) < thr;
        if (accept) {
            x0 = xp0;
            x1 = xp1;
            q = qp;
            ++nacc;
        }
        if (MOMENTS) {
            sx0 += x0;
            sxx0 = fma(x0, x0, sxx0);
            sx1 += x1;
            sxx1 = fma(x1, x1, sxx1);
        }
        if (a.accept_bits && sub == 0 && active) a.accept_bits[(size_t)c * a.steps + s] = (uint8_t)accept;
    }
    if (live0) {
        a.x[(size_t)c * d + sub] = x0;
        if (MOMENTS) {
            if (a.sum_x) a.sum_x[(size_t)c * d + sub] = sx0;
            if (a.sum_xx) a.sum_xx[(size_t)c * d + sub] = sxx0;
        }
    }
    if (live1) {
        a.x[(size_t)c * d + sub + H] = x1;
        if (MOMENTS) {
            if (a.sum_x) a.sum_x[(size_t)c * d + sub + H] = sx1;
            if (a.sum_xx) a.sum_xx[(size_t)c * d + sub + H] = sxx1;
        }
    }
    if (a.n_accept && sub == 0 && active) a.n_accept[c] = nacc;
}

// ---- per-point covariance log-density -------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
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
// 1-D TMA: global -> shared bulk copy that completes on an mbarrier (SASS: UBLKCP).
__device__ __forceinline__ void tma_load_1d(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

struct PerPointArgs {
    const double *x, *mu, *L;
    double *out;
    int64_t N;
    int d, use_tma;
};

// A warp serves G = 32 / D consecutive points per iteration, D lanes each (D = d padded to a power
// of two): at d = 8 four points, so no lane idles and the warp's loads of x / mu are one contiguous
// line.  The G packed factors are contiguous too (point-major): ONE 1-D TMA bulk copy per iteration
// stages them, double buffered, completing on the warp's mbarrier.  d = 32: rows are read from
// shared memory as the substitution needs them instead of being copied to 64 registers first -- 80
// registers instead of 108, and shared memory (6 blocks of 4 warps), not the register file, sets the
// residency: 544 -> 482 us for 2^19 points = 0.79 of the HBM roofline.
template <int D>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
perpoint_kernel(const PerPointArgs a, const Epilogue ep, const double lognorm_base)
{
    constexpr int G = 32 / D;                        // points per warp and iteration
    constexpr int kPackedMax = D * (D + 1) / 2;
    constexpr int kBuf = G * kPackedMax + ((G * kPackedMax) & 1);
    constexpr bool kRowsInSmem = D >= 32;
    __shared__ __align__(128) double s_L[kWarpsPerBlock][2][kBuf];
    __shared__ __align__(8) uint64_t s_bar[kWarpsPerBlock][2];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int grp = lane / D, sub = lane % D;        // point within the iteration, component
    const int d = a.d;
    const int packed = d * (d + 1) / 2;
    const int64_t n_warps = (int64_t)gridDim.x * kWarpsPerBlock;
    const int64_t w0 = ((int64_t)blockIdx.x * kWarpsPerBlock + wib) * G;   // first point of the warp's first iteration
    const int64_t stride = n_warps * G;
    const bool comp = sub < d;

    if (a.use_tma && lane == 0) {
        mbar_init(&s_bar[wib][0], 1);
        mbar_init(&s_bar[wib][1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    auto issue = [&](int64_t pt0, int buf) {
        const int64_t left = a.N - pt0;
        const int n_pts = left < G ? (int)left : G;
        if (a.use_tma) {
            if (lane == 0) {
                const uint32_t bytes = (uint32_t)(n_pts * packed * sizeof(double));
                mbar_expect_tx(&s_bar[wib][buf], bytes);
                tma_load_1d(&s_L[wib][buf][0], a.L + (size_t)pt0 * packed, bytes, &s_bar[wib][buf]);
            }
        } else {
            const double *src = a.L + (size_t)pt0 * packed;
            for (int e = lane; e < n_pts * packed; e += 32) s_L[wib][buf][e] = ld_stream(src + e);
        }
    };
    // the point and its mean are prefetched one iteration ahead as well (registers): without it every
    // iteration started by waiting a full memory latency on them (the kernel's top stall site)
    auto load_x = [&](int64_t pt) { return (comp && pt < a.N) ? ld_stream(a.x + (size_t)pt * d + sub) : 0.0; };
    auto load_m = [&](int64_t pt) { return (comp && pt < a.N && a.mu) ? ld_stream(a.mu + (size_t)pt * d + sub) : 0.0; };
    double x_next = 0.0, m_next = 0.0;
    if (w0 < a.N) {
        issue(w0, 0);
        x_next = load_x(w0 + grp);
        m_next = load_m(w0 + grp);
    }
    uint32_t phase[2] = {0, 0};
    int buf = 0;
    for (int64_t pt0 = w0; pt0 < a.N; pt0 += stride, buf ^= 1) {
        const int64_t nxt = pt0 + stride;
        const int64_t pt = pt0 + grp;
        const bool live = comp && pt < a.N;
        const double x_cur = x_next, m_cur = m_next;
        if (nxt < a.N) {
            issue(nxt, buf ^ 1);                     // prefetch the next factors
            x_next = load_x(nxt + grp);
            m_next = load_m(nxt + grp);
        }
        double r = x_cur - m_cur;
        if (a.use_tma) {
            mbar_wait(&s_bar[wib][buf], phase[buf]);
            phase[buf] ^= 1;
        } else {
            __syncwarp();
        }
        const double *rowp = &s_L[wib][buf][grp * packed + sub * (sub + 1) / 2];
        const double diag = live ? rowp[sub] : 1.0;
        const double rinv = live ? 1.0 / diag : 0.0;
        // forward substitution, column oriented, inside the point's D lanes: lane k holds r_k; step j
        // broadcasts v_j = r_j / L_jj and every lane k > j subtracts L_kj v_j
        double q = 0.0;
        if constexpr (kRowsInSmem) {
#pragma unroll
            for (int j = 0; j < D; ++j) {
                const double vj = __shfl_sync(0xffffffffu, r * rinv, j, D);
                q = fma(vj, vj, q);
                const double lkj = (live && j <= sub) ? rowp[j] : 0.0;
                r = fma(-lkj, vj, r);
            }
            __syncwarp();                             // everyone has read the buffer before it is refilled
        } else {
            double row[D];
#pragma unroll
            for (int j = 0; j < D; ++j) row[j] = (live && j <= sub) ? rowp[j] : 0.0;
            __syncwarp();                             // everyone has read the buffer before it is refilled
#pragma unroll
            for (int j = 0; j < D; ++j) {
                const double vj = __shfl_sync(0xffffffffu, r * rinv, j, D);
                q = fma(vj, vj, q);
                r = fma(-row[j], vj, r);              // no-op for lanes k < j (row[j] == 0); lane j is done with r
            }
        }
        // log det Sigma = 2 sum log L_kk
        double ld = live ? log(diag) : 0.0;
#pragma unroll
        for (int o = D / 2; o > 0; o >>= 1) ld += __shfl_xor_sync(0xffffffffu, ld, o);
        if (sub == 0 && pt < a.N) {
            Epilogue e2 = ep;
            e2.lognorm = lognorm_base - ld;           // lognorm_base excludes -1/2 log det
            e2.scale = exp(e2.lognorm);
            st_stream(a.out + pt, density_epilogue(e2, q));
        }
    }
}

}  // namespace

extern "C" int cusmc_mh_chains_dev(cusmc_ctx *ctx, int kind, int64_t C, int d, int steps, double step_size,
                                   double nu, int shared, const double *mu_dev, const double *L_dev,
                                   double *x_dev, const double *z_dev, const double *thr_dev, uint64_t seed,
                                   uint32_t *n_accept_dev, uint8_t *accept_bits_dev, double *sum_x_dev,
                                   double *sum_xx_dev)
{
    CUSMC_ENTER(ctx);
    CUSMC_REQUIRE(ctx, C >= 0 && steps >= 0 && d >= 1, "bad sizes");
    CUSMC_REQUIRE(ctx, kind == CUSMC_MVN || kind == CUSMC_MVT, "unknown distribution");
    CUSMC_REQUIRE(ctx, kind == CUSMC_MVN || nu > 0.0, "mvt needs nu > 0");
    CUSMC_REQUIRE(ctx, (z_dev == nullptr) == (thr_dev == nullptr), "z and thr must both be given or both NULL");
    CUSMC_REQUIRE(ctx, C == 0 || (mu_dev && L_dev && x_dev), "NULL pointer");
    if (d > 32) return cusmc_fail(ctx, CUSMC_ERR_UNSUPPORTED, "d = %d > 32", d);
    if (C == 0) return CUSMC_OK;
    ChainArgs a;
    a.mu = mu_dev; a.L = L_dev; a.z = z_dev; a.thr = thr_dev;
    a.x = x_dev; a.sum_x = sum_x_dev; a.sum_xx = sum_xx_dev;
    a.n_accept = n_accept_dev; a.accept_bits = accept_bits_dev;
    a.C = C; a.seed = seed; a.step_size = step_size; a.nu = nu;
    a.d = d; a.steps = steps; a.kind = kind; a.shared = shared;
    const int pad = cusmc_pad_dim(d);
    const int per_block = kWarpsPerBlock * (32 / (pad < 4 ? 4 : pad));    // chains per block
    const unsigned grid = (unsigned)((C + per_block - 1) / per_block);
    const bool philox = z_dev == nullptr;
    const bool moments = sum_x_dev != nullptr || sum_xx_dev != nullptr;
    const bool fast = philox && ctx->chain_fast_noise;
    if (pad == 32 && !moments && CUSMC_CHAINS_DUAL != 0) {
        // two chains per warp (mh_chains_dual_kernel)
        const unsigned grid2 = (unsigned)((C + kWarpsPerBlock * 2 - 1) / (kWarpsPerBlock * 2));
        if (fast) mh_chains_dual_kernel<true, true><<<grid2, kWarpsPerBlock * 32, 0, ctx->stream>>>(a);
        else if (philox) mh_chains_dual_kernel<true, false><<<grid2, kWarpsPerBlock * 32, 0, ctx->stream>>>(a);
        else mh_chains_dual_kernel<false, false><<<grid2, kWarpsPerBlock * 32, 0, ctx->stream>>>(a);
        CUSMC_LAUNCHED(ctx);
        return CUSMC_OK;
    }
#define CUSMC_CHAIN_LAUNCH(DD, PH, MO) \
    mh_chains_kernel<DD, PH, MO><<<grid, kWarpsPerBlock * 32, 0, ctx->stream>>>(a)
#define CUSMC_CHAIN_CASE(DD)                                                                         \
    case DD:                                                                                         \
        if (fast && moments) mh_chains_kernel<DD, true, true, true><<<grid, kWarpsPerBlock * 32, 0, ctx->stream>>>(a); \
        else if (fast) mh_chains_kernel<DD, true, false, true><<<grid, kWarpsPerBlock * 32, 0, ctx->stream>>>(a);      \
        else if (philox && moments) CUSMC_CHAIN_LAUNCH(DD, true, true);                              \
        else if (philox) CUSMC_CHAIN_LAUNCH(DD, true, false);                                        \
        else if (moments) CUSMC_CHAIN_LAUNCH(DD, false, true);                                       \
        else CUSMC_CHAIN_LAUNCH(DD, false, false);                                                   \
        break;
    switch (cusmc_pad_dim(d)) {
        CUSMC_CHAIN_CASE(2)
        CUSMC_CHAIN_CASE(4)
        CUSMC_CHAIN_CASE(8)
        CUSMC_CHAIN_CASE(16)
        CUSMC_CHAIN_CASE(32)
    }
#undef CUSMC_CHAIN_LAUNCH
#undef CUSMC_CHAIN_CASE
    CUSMC_LAUNCHED(ctx);
    return CUSMC_OK;
}

extern "C" int cusmc_mh_chains_general_dev(cusmc_ctx *ctx, int kind, int64_t C, int d, int steps, double step_size,
                                           const double *scale_dev, double nu, int shared, const double *mu_dev,
                                           const double *L_dev, double *x_dev, const double *z_dev,
                                           const double *thr_dev, uint64_t seed, uint32_t *n_accept_dev,
                                           uint8_t *accept_bits_dev, double *sum_x_dev, double *sum_xx_dev)
{
    CUSMC_ENTER(ctx);
    CUSMC_REQUIRE(ctx, C >= 0 && steps >= 0 && d >= 1, "bad sizes");
    CUSMC_REQUIRE(ctx, kind == CUSMC_MVN || kind == CUSMC_MVT, "unknown distribution");
    CUSMC_REQUIRE(ctx, kind == CUSMC_MVN || nu > 0.0, "mvt needs nu > 0");
    CUSMC_REQUIRE(ctx, (z_dev == nullptr) == (thr_dev == nullptr), "z and thr must both be given or both NULL");
    CUSMC_REQUIRE(ctx, C == 0 || (mu_dev && L_dev && x_dev), "NULL pointer");
    if (d > 32) return cusmc_fail(ctx, CUSMC_ERR_UNSUPPORTED, "d = %d > 32", d);
    if (C == 0) return CUSMC_OK;
    GeneralArgs ga;
    ChainArgs &a = ga.c;
    a.mu = mu_dev; a.L = L_dev; a.z = z_dev; a.thr = thr_dev;
    a.x = x_dev; a.sum_x = sum_x_dev; a.sum_xx = sum_xx_dev;
    a.n_accept = n_accept_dev; a.accept_bits = accept_bits_dev;
    a.C = C; a.seed = seed; a.step_size = step_size; a.nu = nu;
    a.d = d; a.steps = steps; a.kind = kind; a.shared = shared;
    ga.scale = scale_dev;
    const int pad = cusmc_pad_dim(d);
    const int per_block = kWarpsPerBlock * (32 / (pad < 4 ? 4 : pad));    // chains per block
    const unsigned grid = (unsigned)((C + per_block - 1) / per_block);
    const bool philox = z_dev == nullptr;
    const bool moments = sum_x_dev != nullptr || sum_xx_dev != nullptr;
    const bool fast = philox && ctx->chain_fast_noise;
    if (pad == 32 && CUSMC_GENERAL_DUAL != 0) {
        // two chains per warp (mh_general_dual_kernel)
        const unsigned grid2 = (unsigned)((C + kWarpsPerBlock * 2 - 1) / (kWarpsPerBlock * 2));
#define CUSMC_DUAL_GO(PH, MO, FA) mh_general_dual_kernel<PH, MO, FA><<<grid2, kWarpsPerBlock * 32, 0, ctx->stream>>>(ga)
        if (fast && moments) CUSMC_DUAL_GO(true, true, true);
        else if (fast) CUSMC_DUAL_GO(true, false, true);
        else if (philox && moments) CUSMC_DUAL_GO(true, true, false);
        else if (philox) CUSMC_DUAL_GO(true, false, false);
        else if (moments) CUSMC_DUAL_GO(false, true, false);
        else CUSMC_DUAL_GO(false, false, false);
#undef CUSMC_DUAL_GO
        CUSMC_LAUNCHED(ctx);
        return CUSMC_OK;
    }
#define CUSMC_GEN_LAUNCH(DD, PH, MO) \
    mh_general_kernel<DD, PH, MO><<<grid, kWarpsPerBlock * 32, 0, ctx->stream>>>(ga)
#define CUSMC_GEN_CASE(DD)                                                                           \
    case DD:                                                                                         \
        if (fast && moments) mh_general_kernel<DD, true, true, true><<<grid, kWarpsPerBlock * 32, 0, ctx->stream>>>(ga); \
        else if (fast) mh_general_kernel<DD, true, false, true><<<grid, kWarpsPerBlock * 32, 0, ctx->stream>>>(ga);      \
        else if (philox && moments) CUSMC_GEN_LAUNCH(DD, true, true);                                \
        else if (philox) CUSMC_GEN_LAUNCH(DD, true, false);                                          \
        else if (moments) CUSMC_GEN_LAUNCH(DD, false, true);                                         \
        else CUSMC_GEN_LAUNCH(DD, false, false);                                                     \
        break;
    switch (pad) {
        CUSMC_GEN_CASE(2)
        CUSMC_GEN_CASE(4)
        CUSMC_GEN_CASE(8)
        CUSMC_GEN_CASE(16)
        CUSMC_GEN_CASE(32)
    }
#undef CUSMC_GEN_LAUNCH
#undef CUSMC_GEN_CASE
    CUSMC_LAUNCHED(ctx);
    return CUSMC_OK;
}

extern "C" int cusmc_logpdf_perpoint_dev(cusmc_ctx *ctx, int kind, int want_log, const double *x_dev,
                                         const double *mu_dev, const double *L_dev, int64_t N, int d,
                                         float nu, double *out_dev)
{
    CUSMC_ENTER(ctx);
    CUSMC_REQUIRE(ctx, N >= 0 && d >= 1, "bad sizes");
    CUSMC_REQUIRE(ctx, kind == CUSMC_MVN || kind == CUSMC_MVT, "unknown distribution");
    CUSMC_REQUIRE(ctx, kind == CUSMC_MVN || nu > 0.0f, "mvt needs nu > 0");
    CUSMC_REQUIRE(ctx, N == 0 || (x_dev && L_dev && out_dev), "NULL pointer");
    if (d > 32) return cusmc_fail(ctx, CUSMC_ERR_UNSUPPORTED, "d = %d > 32", d);
    if (N == 0) return CUSMC_OK;
    Epilogue ep{};
    ep.kind = kind;
    ep.want_log = want_log;
    double base;   // log normalising constant without the -1/2 log det term
    if (kind == CUSMC_MVN) {
        base = hostmath::mvn_lognorm(0.0, d);
    } else {
        base = hostmath::mvt_lognorm(0.0, d, nu);
        ep.half_nu_d = hostmath::mvt_half_nu_plus_d(nu, d);
        ep.inv_nu = 1.0 / (double)nu;
    }
    PerPointArgs a;
    a.x = x_dev; a.mu = mu_dev; a.L = L_dev; a.out = out_dev; a.N = N; a.d = d;
    const size_t bytes = sizeof(double) * (size_t)d * (d + 1) / 2;
    a.use_tma = (bytes % 16 == 0) && ((uintptr_t)L_dev % 16 == 0);
    const int per_warp = 32 / cusmc_pad_dim(d);        // points per warp and iteration
    int64_t grid = (N + (int64_t)kWarpsPerBlock * per_warp - 1) / ((int64_t)kWarpsPerBlock * per_warp);
    // one wave of resident blocks (every warp strides over the points)
    int per_sm = 0;
    const void *fn = nullptr;
    switch (cusmc_pad_dim(d)) {
        case 2: fn = (const void *)perpoint_kernel<2>; break;
        case 4: fn = (const void *)perpoint_kernel<4>; break;
        case 8: fn = (const void *)perpoint_kernel<8>; break;
        case 16: fn = (const void *)perpoint_kernel<16>; break;
        default: fn = (const void *)perpoint_kernel<32>; break;
    }
    CUSMC_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, kWarpsPerBlock * 32, 0));
    const int64_t cap = (int64_t)ctx->sm_count * (per_sm > 0 ? per_sm : 4);
    if (grid > cap) grid = cap;
    switch (cusmc_pad_dim(d)) {
        case 2: perpoint_kernel<2><<<(unsigned)grid, kWarpsPerBlock * 32, 0, ctx->stream>>>(a, ep, base); break;
        case 4: perpoint_kernel<4><<<(unsigned)grid, kWarpsPerBlock * 32, 0, ctx->stream>>>(a, ep, base); break;
        case 8: perpoint_kernel<8><<<(unsigned)grid, kWarpsPerBlock * 32, 0, ctx->stream>>>(a, ep, base); break;
        case 16: perpoint_kernel<16><<<(unsigned)grid, kWarpsPerBlock * 32, 0, ctx->stream>>>(a, ep, base); break;
        default: perpoint_kernel<32><<<(unsigned)grid, kWarpsPerBlock * 32, 0, ctx->stream>>>(a, ep, base); break;
    }
    CUSMC_LAUNCHED(ctx);
    return CUSMC_OK;
}

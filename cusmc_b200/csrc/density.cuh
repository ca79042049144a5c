// density.cuh -- the "affine whitening" quadratic form every density path shares.
//
//   z = off - M (x - shift),   q = |z|^2,   value = epilogue(q)
//
// Shared-covariance MVN/MVT log-density: M = L^-1 (lower triangular, Sigma = L L^T),
// shift = mu, off = 0 (replaces the per-particle sigma.inverse()/determinant() of
// src/statistics.cc.cpp:171-196,295-324).  Reweight with an observation matrix:
// M = L_V^-1 F, off = L_V^-1 y, shift = 0 (src/mcmc.cpp:212).
//
// The operator lives in the kernel's parameter bank (__grid_constant__): every
// thread of a warp reads the same coefficient at the same time, so each DFMA takes
// its matrix operand straight from the constant bank -- no shared-memory staging,
// no extra global traffic, nothing to copy to the device per call.
//
// Summation order is part of the contract (oracle: orc_quadform_fma): row k
// accumulates j ascending from off[k] with fma, q accumulates k ascending from 0 with
// fma.  The library is compiled with -fmad=false, so only the fma() written here fuses.
#pragma once

#include "common.cuh"

template <int D, bool TRI>
struct AffineOp {
    static constexpr int NM = TRI ? D * (D + 1) / 2 : D * D;
    double M[NM];      // row-major; TRI: rows packed, row k starts at k(k+1)/2
    double shift[D];
    double off[D];
};

struct Epilogue {
    double scale;       // density mode: normalising constant (the reference's `norm`)
    double lognorm;     // log mode: log of it
    double half_nu_d;   // 0.5 * (float)(nu + d)      (mvt)
    double inv_nu;      // 1 / nu                     (mvt)
    int kind;           // cusmc_dist_kind
    int want_log;
};

template <int D, bool TRI>
__device__ __forceinline__ double affine_quadform(const AffineOp<D, TRI> &op, const double (&r)[D])
{
    double q = 0.0;
#pragma unroll
    for (int k = 0; k < D; ++k) {
        double z = op.off[k];
#pragma unroll
        for (int j = 0; j < (TRI ? k + 1 : D); ++j)
            z = fma(-op.M[TRI ? k * (k + 1) / 2 + j : k * D + j], r[j], z);
        q = fma(z, z, q);
    }
    return q;
}

__device__ __forceinline__ double density_epilogue(const Epilogue &ep, double q)
{
    if (ep.kind == CUSMC_MVN) {
        if (ep.want_log) return fma(-0.5, q, ep.lognorm);
        return ep.scale * exp(-0.5 * q);                       // norm * exp(-0.5 * quadform)
    }
    if (ep.want_log) return fma(-ep.half_nu_d, log1p(q * ep.inv_nu), ep.lognorm);
    return ep.scale * pow(fma(q, ep.inv_nu, 1.0), -ep.half_nu_d);   // norm * quadform^(-(nu+n)/2)
}

// Smallest instantiated D that holds d (operators are zero-padded up to it).
inline int cusmc_pad_dim(int d)
{
    if (d <= 2) return 2;
    if (d <= 4) return 4;
    if (d <= 8) return 8;
    if (d <= 16) return 16;
    return 32;
}

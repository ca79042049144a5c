// pf_fused.cu -- dispatcher of the fused step kernel (instantiations: pf_fused_inst.cu).
#include "filter_types.cuh"
#include "pf_fused_impl.cuh"

int cusmc_launch_fused(cusmc_ctx *ctx, int d, int dy, const double *G, const double *Q, double qscale,
                       const std::vector<double> *M, const double *c, const double *mu, const Epilogue &ep,
                       const pffused::FusedArgs &fa, bool philox)
{
    const int dm = d > dy ? d : dy;
    if (dm > CUSMC_MAX_DIM || d < 1)
        return cusmc_fail(ctx, CUSMC_ERR_UNSUPPORTED, "dimension %d not in 1..%d", dm, CUSMC_MAX_DIM);
    if (fa.s.n_out == 0) return CUSMC_OK;
    const bool exact = d == dy && d == cusmc_pad_dim(dm);
    bool diag = exact && cusmc_is_diag_colmajor(G, d) && cusmc_is_diag_colmajor(Q, d);
    if (diag && M)
        for (int k = 0; k < d && diag; ++k)
            for (int j = 0; j < d; ++j)
                if (j != k && (*M)[(size_t)k * d + j] != 0.0) {
                    diag = false;
                    break;
                }
    const pfstep::StepModel m{d, dy, G, Q, qscale, M, c, mu};
    return fa.s.kind == CUSMC_MVT ? pffused::launch_family<true>(ctx, m, ep, fa, philox, exact, diag)
                                  : pffused::launch_family<false>(ctx, m, ep, fa, philox, exact, diag);
}

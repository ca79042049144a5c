// pf_step_mvn.cu -- Normal-noise instantiations of the step kernel + the dispatcher.
#include "filter_types.cuh"
#include "pf_step_impl.cuh"

namespace pfstep {
int launch_mvn(cusmc_ctx *ctx, const StepModel &m, const Epilogue &ep, const StepArgs &a, bool philox,
               bool exact, bool diag)
{
    return launch_family<false>(ctx, m, ep, a, philox, exact, diag);
}
}  // namespace pfstep

int cusmc_launch_step(cusmc_ctx *ctx, int d, int dy, const double *G, const double *Q, double qscale,
                      const std::vector<double> *M, const double *c, const double *mu, const Epilogue &ep,
                      const StepArgs &a, bool philox)
{
    const int dm = d > dy ? d : dy;
    if (dm > CUSMC_MAX_DIM || d < 1)
        return cusmc_fail(ctx, CUSMC_ERR_UNSUPPORTED, "dimension %d not in 1..%d", dm, CUSMC_MAX_DIM);
    if (a.n_out == 0) return CUSMC_OK;
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
    return a.kind == CUSMC_MVT ? pfstep::launch_mvt(ctx, m, ep, a, philox, exact, diag)
                               : pfstep::launch_mvn(ctx, m, ep, a, philox, exact, diag);
}

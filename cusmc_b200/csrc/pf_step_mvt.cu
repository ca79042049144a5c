// pf_step_mvt.cu -- Student-t-noise instantiations of the step kernel (per-component chi factors,
// the reference's Q2 semantics: src/statistics.cc.cpp:381-387,411).
#include "pf_step_impl.cuh"

namespace pfstep {
int launch_mvt(cusmc_ctx *ctx, const StepModel &m, const Epilogue &ep, const StepArgs &a, bool philox,
               bool exact, bool diag)
{
    return launch_family<true>(ctx, m, ep, a, philox, exact, diag);
}
}  // namespace pfstep

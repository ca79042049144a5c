// pf_fused_inst.cu -- the fused step kernels of ONE (dimension, noise family) pair; the Makefile
// compiles this file once per pair (-DCUSMC_INST_D=... -DCUSMC_INST_MVT=0|1) so the build parallelises.
#ifndef CUSMC_INST_D
#error "compile with -DCUSMC_INST_D=<2|4|8|16|32> -DCUSMC_INST_MVT=<0|1>"
#endif
#include "pf_fused_impl.cuh"

namespace pffused {
CUSMC_FUSED_VARIANTS(template, CUSMC_INST_D, (CUSMC_INST_MVT != 0))
}  // namespace pffused

"""cusmc_b200 -- B200-native implementation of the CuSMC sampling hot path.

Hand-written sm_100a CUDA kernels (cusmc_b200/csrc) behind the C ABI of include/cusmc_b200.h,
plus this thin host-side mirror of the reference's interface.  No CPU fallback: importing the
package without the built library, or using it without a CUDA device, fails loudly.
"""
from ._lib import (AOS, SOA, CusmcError, LIB_PATH, MVN as KIND_MVN, MVT as KIND_MVT,  # noqa: F401
                   RESAMPLE_METROPOLIS, RESAMPLE_MULTINOMIAL, RESAMPLE_SYSTEMATIC, load)
from .api import (Context, MVN, MVNPDF, MVT, MVTPDF, ParticleFilter, default_context,  # noqa: F401
                  metropolis_hastings, run)

from .sharded import ShardPlan, ShardedParticleFilter  # noqa: F401,E402

load()   # a missing or incomplete libcusmc_b200.so is an import error, never a silent fallback

__all__ = ["Context", "ParticleFilter", "MVN", "MVNPDF", "MVT", "MVTPDF", "metropolis_hastings", "run",
           "CusmcError", "AOS", "SOA", "ShardPlan", "ShardedParticleFilter"]

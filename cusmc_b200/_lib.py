"""ctypes binding of libcusmc_b200.so (the extern "C" layer in include/cusmc_b200.h).

The library is built in-tree by ``__graft_entry__.build()`` (or ``make -C cusmc_b200/csrc``).
There is no fallback of any kind: a missing library, a missing symbol or a machine without a
CUDA device raises immediately.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# CUSMC_B200_LIB: another build of the same library (kernel-variant experiments under profiles/)
LIB_PATH = os.environ.get("CUSMC_B200_LIB") or os.path.join(_HERE, "libcusmc_b200.so")

c_double_p = C.POINTER(C.c_double)
c_u32_p = C.POINTER(C.c_uint32)
c_u64_p = C.POINTER(C.c_uint64)
c_u8_p = C.POINTER(C.c_uint8)
vp = C.c_void_p
i64 = C.c_int64
u64 = C.c_uint64
ci = C.c_int
dbl = C.c_double
flt = C.c_float

OK, ERR_INVALID, ERR_CUDA, ERR_NOT_SPD, ERR_DEGENERATE, ERR_UNSUPPORTED, ERR_TIMEOUT = range(7)
MVN, MVT = 0, 1
SOA, AOS = 0, 1
RESAMPLE_METROPOLIS, RESAMPLE_SYSTEMATIC, RESAMPLE_MULTINOMIAL, RESAMPLE_REJECTION, RESAMPLE_METROPOLIS_C2 = 0, 1, 2, 3, 4
MAX_DIM = 32
MAX_PEERS = 8
IPC_HANDLE_BYTES = 64
FILTER_IPC_BUFFERS = 7


class FilterConfig(C.Structure):
    _fields_ = [
        ("N", i64), ("d", ci), ("dy", ci), ("T", ci), ("kind", ci), ("resampler", ci), ("B", ci),
        ("nu", flt), ("noise_scale", dbl), ("seed", u64),
        ("Y", vp), ("m0", vp), ("C0", vp), ("F", vp), ("G", vp), ("V", vp), ("W", vp),
        ("keep_history", ci), ("summary", ci), ("rank", ci), ("world", ci), ("persistent", ci), ("ess_threshold", dbl),
        ("mvt_normal_init", ci), ("reproducible_rng", ci), ("tile_size", ci),
    ]


class FilterDraws(C.Structure):
    _fields_ = [
        ("xi0_dev", vp), ("xi_dev", vp), ("chi_dev", vp), ("u_dev", vp), ("j_dev", vp),
        ("u0_host", vp), ("um_dev", vp), ("chi0_dev", vp),
    ]


# name -> (restype, argtypes); every symbol include/cusmc_b200.h declares.
PROTOTYPES = {
    "cusmc_version": (ci, []),
    "cusmc_ctx_create": (ci, [C.POINTER(vp), ci]),
    "cusmc_ctx_destroy": (ci, [vp]),
    "cusmc_last_error": (C.c_char_p, [vp]),
    "cusmc_ctx_set_stream": (ci, [vp, vp]),
    "cusmc_ctx_set_chain_noise": (ci, [vp, ci]),
    "cusmc_ctx_synchronize": (ci, [vp]),
    "cusmc_ctx_launch_count": (u64, [vp]),
    "cusmc_ctx_last_kernel_ms": (dbl, [vp]),
    "cusmc_logpdf_dev": (ci, [vp, ci, ci, vp, ci, i64, i64, ci, vp, vp, flt, vp]),
    "cusmc_logpdf": (ci, [vp, ci, ci, vp, ci, i64, i64, ci, vp, vp, flt, vp]),
    "cusmc_logpdf_perpoint_dev": (ci, [vp, ci, ci, vp, vp, vp, i64, ci, flt, vp]),
    "cusmc_mvn_pdf": (ci, [vp, vp, vp, vp, dbl, vp, vp, i64, ci, ci]),
    "cusmc_mvt_pdf": (ci, [vp, vp, vp, vp, vp, vp, dbl, i64, ci, ci, flt]),
    "cusmc_mvn_sample": (ci, [vp, vp, vp, vp, vp, vp, vp, u64, u64, i64, ci]),
    "cusmc_mvn_sample_init": (ci, [vp, vp, vp, vp, vp, u64, i64, ci]),
    "cusmc_mvt_sample": (ci, [vp, vp, vp, vp, vp, vp, vp, vp, u64, u64, i64, ci, flt]),
    "cusmc_metropolis_hastings": (ci, [vp, vp, vp, vp, vp, u64, u64, i64, ci]),
    "cusmc_metropolis_hastings_dev": (ci, [vp, vp, vp, vp, vp, u64, u64, i64, ci, ci]),
    "cusmc_rejection_resample_dev": (ci, [vp, vp, vp, vp, u64, u64, i64, ci]),
    "cusmc_metropolis_c2_dev": (ci, [vp, vp, vp, u64, u64, i64, ci, ci]),
    "cusmc_propagate_reweight_dev": (ci, [vp, ci, ci, vp, vp, vp, i64, i64, ci, ci, vp, vp, vp, vp,
                                          vp, flt, vp, vp, u64, u64, vp, vp]),
    "cusmc_weights_max_dev": (ci, [vp, vp, i64, vp]),
    "cusmc_tile_prefix_words": (i64, [i64]),
    "cusmc_weights_sum_dev": (ci, [vp, vp, ci, vp, i64, i64, vp, vp]),
    "cusmc_weights_scan_dev": (ci, [vp, vp, ci, vp, i64, i64, vp, vp, vp]),
    "cusmc_resample_systematic_dev": (ci, [vp, vp, ci, vp, i64, i64, vp, vp, vp, i64, i64, i64, dbl, vp]),
    "cusmc_resample_multinomial_dev": (ci, [vp, vp, i64, vp, vp, u64, u64, i64, i64, i64, vp]),
    "cusmc_resample_systematic": (ci, [vp, vp, i64, dbl, vp]),
    "cusmc_resample_multinomial": (ci, [vp, vp, i64, vp, vp]),
    "cusmc_normalize_ess": (ci, [vp, vp, i64, c_double_p, c_double_p]),
    "cusmc_mh_chains_dev": (ci, [vp, ci, i64, ci, ci, dbl, dbl, ci, vp, vp, vp, vp, vp, u64, vp, vp,
                                 vp, vp]),
    "cusmc_mh_chains_general_dev": (ci, [vp, ci, i64, ci, ci, dbl, vp, dbl, ci, vp, vp, vp, vp, vp, u64, vp, vp,
                                         vp, vp]),
    "cusmc_filter_create": (ci, [vp, C.POINTER(FilterConfig), C.POINTER(vp)]),
    "cusmc_filter_destroy": (ci, [vp]),
    "cusmc_filter_tile_size": (i64, [vp]),
    "cusmc_filter_run": (ci, [vp, C.POINTER(FilterDraws)]),
    "cusmc_filter_begin": (ci, [vp, C.POINTER(FilterDraws)]),
    "cusmc_filter_weigh": (ci, [vp, ci]),
    "cusmc_filter_weigh_phase": (ci, [vp, ci, ci, vp]),
    "cusmc_filter_resample": (ci, [vp, ci]),
    "cusmc_filter_propagate": (ci, [vp, ci]),
    "cusmc_filter_mark": (ci, [vp, ci]),
    "cusmc_filter_slot_dev": (ci, [vp, ci, C.POINTER(vp)]),
    "cusmc_filter_moments_dev": (ci, [vp, C.POINTER(vp)]),
    "cusmc_filter_ipc_export": (ci, [vp, vp]),
    "cusmc_filter_ipc_attach": (ci, [vp, vp]),
    "cusmc_filter_run_sharded": (ci, [vp, C.POINTER(FilterDraws)]),
    "cusmc_filter_exchange_status": (ci, [vp, C.POINTER(u64)]),
    "cusmc_filter_set_exchange_timeout": (ci, [vp, dbl]),
    "cusmc_filter_status": (ci, [vp]),
    "cusmc_filter_get_summary": (ci, [vp, vp, vp, vp]),
    "cusmc_filter_get_history": (ci, [vp, vp, vp, vp]),
    "cusmc_filter_get_resampled": (ci, [vp, vp]),
    "cusmc_filter_last_ms": (dbl, [vp]),
    "cusmc_filter_state_dev": (ci, [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)]),
    "cusmc_filter_get_log_weights": (ci, [vp, vp]),
    "cusmc_filter_get_lineage": (ci, [vp, vp, vp]),
    "cusmc_run": (ci, [vp, C.POINTER(FilterConfig), vp, vp]),
    "cusmc_run_ancestors": (ci, [vp, C.POINTER(FilterConfig), vp, vp, vp]),
    "cusmc_aos_to_soa_dev": (ci, [vp, vp, vp, i64, i64, ci]),
    "cusmc_soa_to_aos_dev": (ci, [vp, vp, vp, i64, i64, ci]),
}

_lib = None


class CusmcError(RuntimeError):
    def __init__(self, code, message):
        super().__init__("cusmc status %d: %s" % (code, message))
        self.code = code


def load():
    """Loads the in-tree shared library and binds every declared entry point."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "%s is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU fallback)" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (restype, argtypes) in PROTOTYPES.items():
        fn = getattr(lib, name)   # AttributeError if the .so does not export it
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def exported_symbols():
    return sorted(PROTOTYPES)

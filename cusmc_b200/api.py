"""Host-side mirror of the reference's interface for the sampling hot path, over the C ABI.

Two levels, both thin:

* ``Context`` -- one method per ``extern "C"`` entry point of include/cusmc_b200.h.  Arguments
  are numpy arrays (host-pointer entry points) or torch CUDA tensors (``*_dev`` entry points;
  PyTorch is only the owner of device memory and streams here).
* the R-facing names of the reference package (NAMESPACE:3-8): ``MVN``, ``MVNPDF``, ``MVT``,
  ``MVTPDF``, ``metropolis_hastings``, ``run`` -- same argument order and meaning as
  R/RcppExports.R:17-103, returning numpy arrays shaped like the R values.

Nothing here computes: every number comes out of libcusmc_b200.so, and every failure of the
library is raised as ``CusmcError`` (the binding-side equivalent of the reference's
``Rcpp::stop``, inst/include/support.cuh:9-32).
"""
import ctypes as C
import math

import numpy as np

from . import _lib
from ._lib import (AOS, MVN as KIND_MVN, MVT as KIND_MVT, RESAMPLE_METROPOLIS, RESAMPLE_METROPOLIS_C2,
                   RESAMPLE_MULTINOMIAL, RESAMPLE_REJECTION, RESAMPLE_SYSTEMATIC, SOA, CusmcError,
                   FilterConfig, FilterDraws)

_KINDS = {"mvn": KIND_MVN, "mvt": KIND_MVT}
_RESAMPLERS = {"metropolis": RESAMPLE_METROPOLIS, "systematic": RESAMPLE_SYSTEMATIC,
               "multinomial": RESAMPLE_MULTINOMIAL, "rejection": RESAMPLE_REJECTION,
               "metropolis_c2": RESAMPLE_METROPOLIS_C2}


def _kind(name):
    # the reference looks the name up in a std::map and dies with std::bad_function_call on a
    # miss (src/mcmc.cpp:53-58); here it is an immediate error
    try:
        return _KINDS[name] if isinstance(name, str) else int(name)
    except KeyError:
        raise ValueError("unknown distribution %r (expected 'mvn' or 'mvt')" % (name,))


def _resampler(name):
    try:
        return _RESAMPLERS[name] if isinstance(name, str) else int(name)
    except KeyError:
        raise ValueError("unknown resampler %r (expected one of %s)" % (name, sorted(_RESAMPLERS)))


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _colmajor(M):
    M = np.asarray(M, dtype=np.float64)
    if M.ndim != 2:
        raise ValueError("expected a matrix")
    return np.ascontiguousarray(M.T).ravel()


def _hp(a):
    """Host pointer of a numpy array the CALLER keeps alive (None -> NULL)."""
    return None if a is None else a.ctypes.data


class _Args:
    """Keeps converted temporaries alive for the duration of one C call: taking .ctypes.data of
    an unnamed temporary would leave a dangling pointer by the time the call executes."""

    def __init__(self):
        self.keep = []

    def mat(self, M):
        if M is None:
            return None
        a = _colmajor(M)
        self.keep.append(a)
        return a.ctypes.data

    def vec(self, v, dtype=np.float64):
        if v is None:
            return None
        a = np.ascontiguousarray(v, dtype=dtype)
        self.keep.append(a)
        return a.ctypes.data


def _dp(t):
    """Device pointer of a torch CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda or not t.is_contiguous():
        raise ValueError("expected a contiguous CUDA tensor")
    return t.data_ptr()


class Context:
    """One device, one stream, one caller thread (include/cusmc_b200.h)."""

    def __init__(self, device=0, stream=None):
        self.lib = _lib.load()
        h = C.c_void_p()
        rc = self.lib.cusmc_ctx_create(C.byref(h), int(device))
        if rc != _lib.OK:
            raise CusmcError(rc, "cusmc_ctx_create(device=%d) failed: no usable CUDA device "
                                 "(this library has no CPU fallback)" % device)
        self.h = h
        self.device = int(device)
        if stream is not None:
            self.set_stream(stream)

    def close(self):
        if getattr(self, "h", None):
            self.lib.cusmc_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != _lib.OK:
            raise CusmcError(rc, self.lib.cusmc_last_error(self.h).decode())

    # ---- stream / bookkeeping --------------------------------------------------------------
    def set_chain_noise(self, reproducible=True):
        """Device-drawn proposal normals of mh_chains*_dev: the host-reproducible generator (default) or the
        throughput one (Philox4x32-7 + special-function-unit Box-Muller); see cusmc_ctx_set_chain_noise."""
        self._check(self.lib.cusmc_ctx_set_chain_noise(self.h, int(bool(reproducible))))

    def set_stream(self, stream):
        """stream: a raw cudaStream_t integer, a torch.cuda.Stream, or None (context's own)."""
        ptr = None
        if stream is not None:
            # torch's default stream has handle 0, which the C ABI reads as "the context's own
            # stream": address the legacy default stream by its explicit handle cudaStreamLegacy.
            ptr = getattr(stream, "cuda_stream", stream) or 1
        self._check(self.lib.cusmc_ctx_set_stream(self.h, ptr))

    def use_torch_stream(self):
        import torch
        self.set_stream(torch.cuda.current_stream(self.device))

    def synchronize(self):
        self._check(self.lib.cusmc_ctx_synchronize(self.h))

    @property
    def launch_count(self):
        return int(self.lib.cusmc_ctx_launch_count(self.h))

    @property
    def last_kernel_ms(self):
        return float(self.lib.cusmc_ctx_last_kernel_ms(self.h))

    # ---- densities ------------------------------------------------------------------------
    def logpdf(self, dist, x, mu, sigma, nu=0.0, log=True, layout=AOS):
        """Host arrays.  x: (N, d) for AOS, (d, N) for SOA.  Returns (N,) float64."""
        x = _f64(x)
        N, d = (x.shape if layout == AOS else x.shape[::-1])
        out = np.empty(N)
        mu_ = None if mu is None else _f64(mu)
        sg = _colmajor(sigma)
        self._check(self.lib.cusmc_logpdf(self.h, _kind(dist), int(log), _hp(x), layout, N, N, d,
                                          _hp(mu_), _hp(sg), float(nu), _hp(out)))
        return out

    def logpdf_dev(self, dist, x, mu, sigma, out, nu=0.0, log=True, layout=SOA, N=None, d=None,
                   ld=None):
        """torch CUDA tensors.  SOA: x is (d, ld); AOS: x is (N, d)."""
        if layout == SOA:
            d_ = x.shape[0] if d is None else d
            ld_ = x.shape[1] if ld is None else ld
            N_ = ld_ if N is None else N
        else:
            N_ = x.shape[0] if N is None else N
            d_ = x.shape[1] if d is None else d
            ld_ = N_
        mu_ = None if mu is None else _f64(mu)
        sg = _colmajor(sigma)
        self._check(self.lib.cusmc_logpdf_dev(self.h, _kind(dist), int(log), _dp(x), layout, N_, ld_,
                                              d_, _hp(mu_), _hp(sg), float(nu), _dp(out)))
        return out

    def prepare_density(self, dist, mu, sigma, nu=0.0, log=True):
        """The distribution object of the reference (built once by the Distributions factory,
        src/mcmc.cpp:53-58, then asked for pdf() many times): converts mu / sigma once; use with
        ``logpdf_prepared_dev`` when batches are small enough for per-call conversions to show."""
        return PreparedDensity(dist, mu, sigma, nu, log)

    def logpdf_prepared_dev(self, prep, x, out, layout=SOA, N=None, ld=None):
        """x: torch CUDA tensor, (d, ld) for SOA or (N, d) for AOS; out: (N,)."""
        if layout == SOA:
            ld_ = x.shape[1] if ld is None else ld
            N_ = ld_ if N is None else N
        else:
            N_ = x.shape[0] if N is None else N
            ld_ = N_
        rc = self.lib.cusmc_logpdf_dev(self.h, prep.kind, prep.log, x.data_ptr(), layout, N_, ld_, prep.d,
                                       prep.mu_p, prep.sigma_p, prep.nu, out.data_ptr())
        if rc != _lib.OK:
            self._check(rc)
        return out

    def logpdf_perpoint_dev(self, dist, x, mu, L_packed, out, nu=0.0, log=True):
        N, d = x.shape
        self._check(self.lib.cusmc_logpdf_perpoint_dev(self.h, _kind(dist), int(log), _dp(x), _dp(mu),
                                                       _dp(L_packed), N, d, float(nu), _dp(out)))
        return out

    # ---- drop-ins for the reference's wrappers ------------------------------------------------
    def mvn_pdf(self, y, x_aos, norm, E_inv, F):
        """mvn_pdf_kernel_wrapper: w_i = norm exp(-1/2 r' E_inv r), r = y - F x_i."""
        x = _f64(x_aos)
        N, d = x.shape
        F = np.asarray(F, dtype=np.float64)
        dy = F.shape[0]
        w = np.empty(N)
        A = _Args()
        self._check(self.lib.cusmc_mvn_pdf(self.h, _hp(w), A.vec(y), _hp(x), float(norm),
                                           A.mat(E_inv), A.mat(F), N, d, dy))
        return w

    def mvt_pdf(self, y, x_aos, E_inv, F, norm, df):
        x = _f64(x_aos)
        N, d = x.shape
        F = np.asarray(F, dtype=np.float64)
        dy = F.shape[0]
        w = np.empty(N)
        A = _Args()
        self._check(self.lib.cusmc_mvt_pdf(self.h, _hp(w), A.vec(y), _hp(x), A.mat(E_inv),
                                           A.mat(F), float(norm), N, d, dy, float(df)))
        return w

    def mvn_sample(self, x_prev_aos, a, G, Q, xi=None, seed=0, step=0):
        xp = _f64(x_prev_aos)
        N, d = xp.shape
        out = np.empty((N, d))
        a_ = None if a is None else np.ascontiguousarray(a, dtype=np.uint32)
        xi_ = None if xi is None else _f64(xi)
        A = _Args()
        self._check(self.lib.cusmc_mvn_sample(self.h, _hp(out), _hp(xp), _hp(a_), A.mat(G),
                                              A.mat(Q), _hp(xi_), int(seed), int(step), N, d))
        return out

    def mvn_sample_init(self, mu, Q, N, xi=None, seed=0):
        mu = _f64(mu)
        d = mu.size
        out = np.empty((N, d))
        xi_ = None if xi is None else _f64(xi)
        A = _Args()
        self._check(self.lib.cusmc_mvn_sample_init(self.h, _hp(out), _hp(mu), A.mat(Q),
                                                   _hp(xi_), int(seed), N, d))
        return out

    def mvt_sample(self, x_prev_aos, a, G, Q, df, xi=None, chi=None, seed=0, step=0):
        xp = _f64(x_prev_aos)
        N, d = xp.shape
        out = np.empty((N, d))
        a_ = None if a is None else np.ascontiguousarray(a, dtype=np.uint32)
        xi_ = None if xi is None else _f64(xi)
        chi_ = None if chi is None else _f64(chi)
        A = _Args()
        self._check(self.lib.cusmc_mvt_sample(self.h, _hp(out), _hp(xp), _hp(a_), A.mat(G),
                                              A.mat(Q), _hp(xi_), _hp(chi_), int(seed),
                                              int(step), N, d, float(df)))
        return out

    def metropolis_hastings(self, w, B, u=None, j=None, seed=0, step=1):
        """Sampler::metropolis_hastings.  u, j: (N, B) pre-drawn, or None for device Philox."""
        w = _f64(w)
        N = w.size
        a = np.empty(N, dtype=np.uint32)
        u_ = None if u is None else _f64(u)
        j_ = None if j is None else np.ascontiguousarray(j, dtype=np.uint32)
        self._check(self.lib.cusmc_metropolis_hastings(self.h, _hp(a), _hp(w), _hp(u_), _hp(j_),
                                                       int(seed), int(step), N, int(B)))
        return a

    # ---- normalisation / resampling --------------------------------------------------------
    def resample_systematic(self, w, u0):
        w = _f64(w)
        a = np.empty(w.size, dtype=np.uint32)
        self._check(self.lib.cusmc_resample_systematic(self.h, _hp(w), w.size, float(u0), _hp(a)))
        return a

    def resample_multinomial(self, w, u):
        w, u = _f64(w), _f64(u)
        a = np.empty(w.size, dtype=np.uint32)
        self._check(self.lib.cusmc_resample_multinomial(self.h, _hp(w), w.size, _hp(u), _hp(a)))
        return a

    def normalize_ess(self, lw):
        lw = _f64(lw)
        lse, ess = C.c_double(), C.c_double()
        self._check(self.lib.cusmc_normalize_ess(self.h, _hp(lw), lw.size, C.byref(lse), C.byref(ess)))
        return lse.value, ess.value

    # device-pointer building blocks (torch CUDA tensors)
    def metropolis_hastings_dev(self, a, w, B, u=None, j=None, seed=0, step=1, is_log=False, N=None):
        N = w.numel() if N is None else N
        self._check(self.lib.cusmc_metropolis_hastings_dev(self.h, _dp(a), _dp(w), _dp(u), _dp(j),
                                                           int(seed), int(step), N, int(B), int(is_log)))

    def metropolis_c2_dev(self, a, w, B, seed=0, step=1, is_log=False, N=None):
        """Metropolis-C2: the same rule, a warp's proposals confined to one 32-particle segment per iteration."""
        self._check(self.lib.cusmc_metropolis_c2_dev(self.h, _dp(a), _dp(w), int(seed), int(step),
                                                     w.numel() if N is None else N, int(B), int(is_log)))

    def rejection_resample_dev(self, a, w, w_max, seed=0, step=1, cap=4096, N=None):
        self._check(self.lib.cusmc_rejection_resample_dev(self.h, _dp(a), _dp(w), _dp(w_max), int(seed), int(step),
                                                          w.numel() if N is None else N, int(cap)))

    def weights_max_dev(self, w, max_out, N=None):
        self._check(self.lib.cusmc_weights_max_dev(self.h, _dp(w), w.numel() if N is None else N,
                                                   _dp(max_out)))

    def tile_prefix_words(self, N):
        return int(self.lib.cusmc_tile_prefix_words(int(N)))

    def weights_sum_dev(self, w, is_log, max_dev, N_global, stats, N=None, tile_prefix=None):
        """stats: 4 uint64 words (torch int64).  tile_prefix: int64 tensor of tile_prefix_words(N)
        words with word 0 zero, or None (context scratch)."""
        self._check(self.lib.cusmc_weights_sum_dev(self.h, _dp(w), int(is_log), _dp(max_dev),
                                                   w.numel() if N is None else N, int(N_global),
                                                   _dp(stats), _dp(tile_prefix)))

    def weights_scan_dev(self, w, is_log, max_dev, N_global, cdf, cdf_offset=None, N=None,
                         tile_prefix=None):
        self._check(self.lib.cusmc_weights_scan_dev(self.h, _dp(w), int(is_log), _dp(max_dev),
                                                    w.numel() if N is None else N, int(N_global),
                                                    _dp(cdf_offset), _dp(tile_prefix), _dp(cdf)))

    def resample_systematic_dev(self, w, is_log, max_dev, N_global, total, a, u0, cdf_offset=None,
                                j0=0, out_lo=0, out_n=None, N=None, tile_prefix=None):
        N = w.numel() if N is None else N
        out_n = a.numel() if out_n is None else out_n
        self._check(self.lib.cusmc_resample_systematic_dev(
            self.h, _dp(w), int(is_log), _dp(max_dev), N, int(N_global), _dp(total), _dp(cdf_offset),
            _dp(tile_prefix), int(j0), int(out_lo), int(out_n), float(u0), _dp(a)))

    def resample_multinomial_dev(self, cdf, total, a, u=None, seed=0, step=0, i0=0, j0=0, N=None,
                                 n_out=None):
        N = cdf.numel() if N is None else N
        n_out = a.numel() if n_out is None else n_out
        self._check(self.lib.cusmc_resample_multinomial_dev(self.h, _dp(cdf), N, _dp(total), _dp(u),
                                                            int(seed), int(step), int(i0), int(n_out),
                                                            int(j0), _dp(a)))

    def propagate_reweight_dev(self, dist, x_new, x_prev, a, G, Q, y, F, V, lw, nu=0.0, log=True,
                               xi=None, chi=None, seed=0, step=0, lw_max=None, N=None):
        d, ld = x_prev.shape
        N = ld if N is None else N
        F = np.asarray(F, dtype=np.float64)
        dy = F.shape[0]
        A = _Args()
        self._check(self.lib.cusmc_propagate_reweight_dev(
            self.h, _kind(dist), int(log), _dp(x_new), _dp(x_prev), _dp(a), N, ld, d, dy,
            A.mat(G), A.mat(Q), A.vec(y), A.mat(F), A.mat(V),
            float(nu), _dp(xi), _dp(chi), int(seed), int(step), _dp(lw), _dp(lw_max)))

    def mh_chains_dev(self, dist, mu, L, x, steps, step_size, nu=0.0, shared=False, z=None, thr=None,
                      seed=0, n_accept=None, accept_bits=None, sum_x=None, sum_xx=None):
        """mu: (C, d) or (d,), L: (C, d, d) per-chain COLUMN-major or (d, d); x: (C, d) in/out."""
        Cn, d = x.shape
        self._check(self.lib.cusmc_mh_chains_dev(
            self.h, _kind(dist), Cn, d, int(steps), float(step_size), float(nu), int(shared), _dp(mu),
            _dp(L), _dp(x), _dp(z), _dp(thr), int(seed), _dp(n_accept), _dp(accept_bits), _dp(sum_x),
            _dp(sum_xx)))

    def mh_chains_general_dev(self, dist, mu, L, x, steps, step_size, nu=0.0, shared=False, scale=None, z=None,
                              thr=None, seed=0, n_accept=None, accept_bits=None, sum_x=None, sum_xx=None):
        """Random-walk chains whose proposal x' = x + step_size * scale (.) z does not use the target's
        factor: every step evaluates the target density in the kernel (forward substitution with the
        chain's factor resident in registers).  Arguments as mh_chains_dev; scale: (d,) or None."""
        Cn, d = x.shape
        self._check(self.lib.cusmc_mh_chains_general_dev(
            self.h, _kind(dist), Cn, d, int(steps), float(step_size), _dp(scale), float(nu), int(shared), _dp(mu),
            _dp(L), _dp(x), _dp(z), _dp(thr), int(seed), _dp(n_accept), _dp(accept_bits), _dp(sum_x), _dp(sum_xx)))

    def aos_to_soa_dev(self, aos, soa):
        N, d = aos.shape
        self._check(self.lib.cusmc_aos_to_soa_dev(self.h, _dp(aos), _dp(soa), N, soa.shape[1], d))

    def soa_to_aos_dev(self, soa, aos):
        N, d = aos.shape
        self._check(self.lib.cusmc_soa_to_aos_dev(self.h, _dp(soa), _dp(aos), N, soa.shape[1], d))

    # ---- the filter --------------------------------------------------------------------------
    def filter(self, **kw):
        return ParticleFilter(self, **kw)


class PreparedDensity:
    """(kind, mu, sigma, nu) converted to the C ABI's layout once (see Context.prepare_density)."""

    def __init__(self, dist, mu, sigma, nu=0.0, log=True):
        self.kind = _kind(dist)
        self.log = int(log)
        self.nu = float(nu)
        self._sigma = _colmajor(sigma)
        self.d = int(round(math.sqrt(self._sigma.size)))
        self._mu = None if mu is None else _f64(mu)
        if self._mu is not None and self._mu.size != self.d:
            raise ValueError("mu has %d entries, sigma is %d x %d" % (self._mu.size, self.d, self.d))
        self.mu_p = _hp(self._mu)
        self.sigma_p = _hp(self._sigma)


TILE = 2048     # particles per weight-image tile (csrc/resample.cuh: kTile)


def shard_size(N, world):
    """Slots per rank of a sharded filter (the rule of cusmc_filter_create): ceil(N / world) rounded up to
    whole weight-image tiles, so a sharded run has the tiles -- hence the bits -- of the single-GPU run."""
    per = -(-int(N) // max(1, int(world)))
    return per if world <= 1 else -(-per // TILE) * TILE


def _filter_config(N, Y, m0, C0, F, G, V, W, distribution="mvn", resampler="metropolis", B=10, df=0.0,
                   noise_scale=1.0, seed=0, keep_history=False, summary=True, rank=0, world=1,
                   persistent=True, ess_threshold=0.0, mvt_normal_init=False, reproducible_rng=False, tile_size=0):
    """cusmc_filter_config from numpy inputs; returns (config, arrays to keep alive while it is used)."""
    Y = np.asarray(Y, dtype=np.float64)
    if Y.ndim != 2:
        raise ValueError("Y must be a (dy, T) matrix")
    keep = [_colmajor(Y), _f64(m0), _colmajor(C0), _colmajor(F), _colmajor(G), _colmajor(V), _colmajor(W)]
    cfg = FilterConfig()
    cfg.N, cfg.d, cfg.dy, cfg.T = int(N), np.asarray(G).shape[0], Y.shape[0], Y.shape[1]
    cfg.kind = _kind(distribution)
    cfg.resampler = _resampler(resampler)
    cfg.B = int(B)
    cfg.nu = float(df)
    cfg.noise_scale = float(noise_scale)
    cfg.seed = int(seed)
    (cfg.Y, cfg.m0, cfg.C0, cfg.F, cfg.G, cfg.V, cfg.W) = [a.ctypes.data for a in keep]
    cfg.keep_history = int(keep_history)
    cfg.summary = int(summary)
    cfg.rank, cfg.world = int(rank), int(world)   # world > 1: see cusmc_b200/sharded.py
    cfg.persistent = 0 if persistent else -1      # one cooperative kernel per run when eligible
    cfg.ess_threshold = float(ess_threshold)      # 0: resample every step (the reference's behaviour)
    cfg.mvt_normal_init = int(mvt_normal_init)    # "mvt": x_0 = m0 + chi (.) (Q xi) unless set
    cfg.reproducible_rng = int(reproducible_rng)  # device-drawn normals a host can regenerate bit for bit
    cfg.tile_size = int(tile_size)                # 0: automatic (see cusmc_filter_config.tile_size)
    return cfg, keep


class ParticleFilter:
    """particle_filter() of the reference (src/particle_filter.cpp:6-39) with device-resident state."""

    def __init__(self, ctx, N, Y, m0, C0, F, G, V, W, distribution="mvn", resampler="metropolis", B=10,
                 df=0.0, noise_scale=1.0, seed=0, keep_history=False, summary=True, rank=0, world=1,
                 persistent=True, ess_threshold=0.0, mvt_normal_init=False, reproducible_rng=False, tile_size=0):
        self.ctx = ctx
        cfg, self._keep = _filter_config(N, Y, m0, C0, F, G, V, W, distribution, resampler, B, df, noise_scale,
                                         seed, keep_history, summary, rank, world, persistent, ess_threshold,
                                         mvt_normal_init, reproducible_rng, tile_size)
        self.dy, self.T, self.d, self.N = cfg.dy, cfg.T, cfg.d, int(N)
        self._world = int(world)
        per = shard_size(self.N, int(world))
        self.n_local = self.N if world <= 1 else max(0, min(per, self.N - int(rank) * per))
        self.keep_history = bool(keep_history)
        self.is_log = cfg.resampler not in (RESAMPLE_METROPOLIS, RESAMPLE_REJECTION, RESAMPLE_METROPOLIS_C2)
        h = C.c_void_p()
        ctx._check(ctx.lib.cusmc_filter_create(ctx.h, C.byref(cfg), C.byref(h)))
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.ctx.lib.cusmc_filter_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def run(self, xi0=None, xi=None, chi=None, u=None, j=None, u0=None, um=None, chi0=None):
        """Injected draws are torch CUDA tensors (see cusmc_filter_draws), u0 a host array; any
        omitted stream of randomness is drawn on the device from Philox."""
        dr = self._make_draws(xi0, xi, chi, u, j, u0, um, chi0)
        self.ctx._check(self.ctx.lib.cusmc_filter_run(self.h, C.byref(dr)))
        return self

    def _make_draws(self, xi0=None, xi=None, chi=None, u=None, j=None, u0=None, um=None, chi0=None):
        dr = FilterDraws()
        dr.xi0_dev, dr.xi_dev, dr.chi_dev = _dp(xi0), _dp(xi), _dp(chi)
        dr.chi0_dev = _dp(chi0)
        dr.u_dev, dr.j_dev, dr.um_dev = _dp(u), _dp(j), _dp(um)
        self._u0 = None if u0 is None else _f64(u0)
        dr.u0_host = _hp(self._u0)
        self._draws = (xi0, xi, chi, u, j, um, chi0)   # keep alive until the stream has consumed them
        return dr

    @property
    def last_ms(self):
        return float(self.ctx.lib.cusmc_filter_last_ms(self.h))

    @property
    def tile_size(self):
        """Particles per weight-image tile of the path run() takes (2048, or the persistent kernel's)."""
        return int(self.ctx.lib.cusmc_filter_tile_size(self.h))

    def summary(self):
        mean = np.empty((self.T, self.d))
        ess = np.empty(self.T)
        ll = np.empty(self.T)
        self.ctx._check(self.ctx.lib.cusmc_filter_get_summary(self.h, _hp(mean), _hp(ess), _hp(ll)))
        return dict(mean=mean, ess=ess, loglik=ll)

    def resampled(self):
        """(T,) int array: 1 where the step drew new ancestors (adaptive resampling), else 0."""
        r = np.zeros(self.T, dtype=np.int32)
        self.ctx._check(self.ctx.lib.cusmc_filter_get_resampled(self.h, _hp(r)))
        return r

    def state(self):
        """Current device-resident state copied to the host: x (d, n) SoA, weights (n,) (log-weights
        for the normalised resamplers, densities for "metropolis"), ancestors of the last step (n,)."""
        import torch
        from .sharded import _device_view
        px, pw, pa = C.c_void_p(), C.c_void_p(), C.c_void_p()
        self.ctx._check(self.ctx.lib.cusmc_filter_state_dev(self.h, C.byref(px), C.byref(pw), C.byref(pa)))
        self.ctx.synchronize()
        torch.cuda.synchronize()
        n, dev = self.n_local, self.ctx.device
        per = self.N if self._world <= 1 else shard_size(self.N, self._world)
        x = _device_view(px.value, (self.d, per), "<f8", dev)[:, :n].cpu().numpy()
        w = _device_view(pw.value, (per,), "<f8", dev)[:n].cpu().numpy()
        a = _device_view(pa.value, (per,), "<i4", dev)[:n].cpu().numpy().view(np.uint32)
        return x, w, a

    def history(self):
        """x (T, n, d), a (T, n) and w (T, n): the raw densities in reference mode ("metropolis", w_0 =
        1/N), NORMALISED weights for the other resamplers -- plus their raw log-weights as "lw"."""
        n = self.n_local            # a sharded filter returns its own shard
        x = np.empty((self.T, n, self.d))
        w = np.empty((self.T, n))
        a = np.empty((self.T, n), dtype=np.uint32)
        self.ctx._check(self.ctx.lib.cusmc_filter_get_history(self.h, _hp(x), _hp(w), _hp(a)))
        out = dict(x=x, w=w, a=a)
        if self.is_log:
            lw = np.empty((self.T, n))
            self.ctx._check(self.ctx.lib.cusmc_filter_get_log_weights(self.h, _hp(lw)))
            out["lw"] = lw
        return out

    def lineage(self):
        """The ancestor tree as paths: (lineage (T, n), n_unique (T,)) -- lineage[t, i] is the index at step t
        of the ancestor of final particle i, n_unique[t] the distinct ancestors alive at step t."""
        lin = np.empty((self.T, self.n_local), dtype=np.uint32)
        nu = np.empty(self.T, dtype=np.int32)
        self.ctx._check(self.ctx.lib.cusmc_filter_get_lineage(self.h, _hp(lin), _hp(nu)))
        return lin, nu

    def status(self):
        """Waits for the last run; raises CusmcError (TIMEOUT / DEGENERATE) if it went wrong inside."""
        self.ctx._check(self.ctx.lib.cusmc_filter_status(self.h))


# ==============================================================================================
# R-facing names (NAMESPACE:3-8).  A process-wide default context is created on first use.
# ==============================================================================================
_default_ctx = None


def default_context():
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context(0)
    return _default_ctx


def MVNPDF(x, mu, sigma):
    """MVNPDF(x, mu, sigma) -> density (src/mvn_dist.rcpp.cpp:52-58).  x may also be an (N, d)
    matrix of points (batched extension); a single vector returns a float like the reference."""
    x = np.asarray(x, dtype=np.float64)
    out = default_context().logpdf("mvn", np.atleast_2d(x), mu, sigma, log=False)
    return float(out[0]) if x.ndim == 1 else out


def MVTPDF(x, mu, sigma, nu):
    """MVTPDF(x, mu, sigma, nu) -> density (src/mvt_dist.rcpp.cpp:60-66)."""
    x = np.asarray(x, dtype=np.float64)
    out = default_context().logpdf("mvt", np.atleast_2d(x), mu, sigma, nu=nu, log=False)
    return float(out[0]) if x.ndim == 1 else out


def _eigen_factor(sigma):
    lam, vec = np.linalg.eigh(np.asarray(sigma, dtype=np.float64))
    return vec * np.sqrt(np.clip(lam, 0.0, None))


def MVN(mu, sigma, seed=0, reference_quirks=True):
    """MVN(mu, sigma) -> one draw (src/mvn_dist.rcpp.cpp:31-37).  With reference_quirks the call
    reproduces what the reference's CPU build does: sigma itself is used as the factor
    and the CLT sampler's draws have variance 3 (Q1); otherwise x = mu + V sqrt(Lambda) z."""
    mu = _f64(mu)
    Q = np.asarray(sigma, dtype=np.float64) * math.sqrt(3.0) if reference_quirks else _eigen_factor(sigma)
    return default_context().mvn_sample_init(mu, Q, 1, seed=seed)[0]


def MVT(mu, sigma, nu, seed=0, reference_quirks=True):
    """MVT(mu, sigma, nu) -> one draw (src/mvt_dist.rcpp.cpp:28-49): mu + chi (.) (Q z) with an
    independent chi per component (Q2); Q = V sqrt(Lambda)."""
    mu = _f64(mu)
    d = mu.size
    Q = _eigen_factor(sigma) * (math.sqrt(3.0) if reference_quirks else 1.0)
    ctx = default_context()
    x = ctx.mvt_sample(np.zeros((1, d)), None, np.zeros((d, d)), Q, nu, seed=seed)
    return x[0] + mu


def metropolis_hastings(w, N, B, seed=0):
    """metropolis_hastings(w, N, B) -> 0-based ancestors as doubles (src/samplers.rcpp.cpp:35-55)."""
    w = _f64(w)
    if w.size != int(N):
        raise ValueError("length(w) != N")
    return default_context().metropolis_hastings(w, B, seed=seed, step=1).astype(np.float64)


def run(N, d, timeSteps, Y, m0, C0, F, G, V, W, df, resampler, distribution, p=0, seed=0,
        noise_scale=1.0, ancestors=False, **kw):
    """run(...) (src/run.rcpp.cpp:58-126) -> {'weights': (T, N), 'posterior_x': (T, N, d)} through
    cusmc_run: the history streams to the host while the filter runs (bounded device memory).
    df is passed through correctly (the reference swaps it with `runtime`).
    ancestors=True adds 'ancestors' (T, N) (cusmc_run_ancestors)."""
    Y = np.asarray(Y, dtype=np.float64)
    if Y.shape[1] < timeSteps:
        raise ValueError("Y has fewer than timeSteps columns")
    if np.asarray(G).shape[0] != int(d):
        raise ValueError("G is not d x d")
    ctx = default_context()
    cfg, keep = _filter_config(N, Y[:, :timeSteps], m0, C0, F, G, V, W, distribution, resampler, df=df, seed=seed,
                               noise_scale=noise_scale, summary=False, **kw)
    T = int(timeSteps)
    w = np.empty((T, int(N)))
    x = np.empty((T, int(N), int(d)))
    if ancestors:
        a = np.empty((T, int(N)), dtype=np.uint32)
        ctx._check(ctx.lib.cusmc_run_ancestors(ctx.h, C.byref(cfg), _hp(w), _hp(x), _hp(a)))
        return {"weights": w, "posterior_x": x, "ancestors": a}
    ctx._check(ctx.lib.cusmc_run(ctx.h, C.byref(cfg), _hp(w), _hp(x)))
    return {"weights": w, "posterior_x": x}

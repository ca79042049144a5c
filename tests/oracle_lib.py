"""ctypes view of the CPU oracle (oracle/libcusmc_oracle.so).

Test infrastructure only: the product package never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_DIR = os.path.join(os.path.dirname(_HERE), "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "libcusmc_oracle.so")

_dp = C.POINTER(C.c_double)
_u32p = C.POINTER(C.c_uint32)
_u64p = C.POINTER(C.c_uint64)
_u8p = C.POINTER(C.c_uint8)


def _p(a, typ=_dp):
    if a is None:
        return typ()
    return a.ctypes.data_as(typ)


def f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def colmajor(M):
    """Flat column-major copy of a 2-D matrix (Eigen's default storage)."""
    return np.asfortranarray(np.asarray(M, dtype=np.float64)).ravel(order="F").copy()


def build():
    if not os.path.exists(ORACLE_SO) or (
        os.path.getmtime(ORACLE_SO) < os.path.getmtime(os.path.join(ORACLE_DIR, "cusmc_oracle.c"))
    ):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "-s"])
    return ORACLE_SO


class Oracle:
    def __init__(self):
        self.lib = C.CDLL(build())
        L = self.lib
        L.orc_determinant.restype = C.c_double
        L.orc_determinant.argtypes = [_dp, C.c_int]
        L.orc_inverse.restype = C.c_int
        L.orc_inverse.argtypes = [_dp, C.c_int, _dp]
        L.orc_cholesky_lower.restype = C.c_int
        L.orc_cholesky_lower.argtypes = [_dp, C.c_int, _dp]
        L.orc_tri_inverse_lower.restype = None
        L.orc_tri_inverse_lower.argtypes = [_dp, C.c_int, _dp]
        L.orc_mvn_norm.restype = C.c_double
        L.orc_mvn_norm.argtypes = [_dp, C.c_int]
        L.orc_mvt_norm.restype = C.c_double
        L.orc_mvt_norm.argtypes = [_dp, C.c_int, C.c_float]
        L.orc_mvn_pdf1.restype = C.c_double
        L.orc_mvn_pdf1.argtypes = [_dp, _dp, C.c_int]
        L.orc_mvt_pdf1.restype = C.c_double
        L.orc_mvt_pdf1.argtypes = [_dp, _dp, C.c_int, C.c_float]
        L.orc_MVNPDF.restype = C.c_double
        L.orc_MVNPDF.argtypes = [_dp, _dp, _dp, C.c_int]
        L.orc_MVTPDF.restype = C.c_double
        L.orc_MVTPDF.argtypes = [_dp, _dp, _dp, C.c_int, C.c_float]
        L.orc_pdf_batch.restype = None
        L.orc_pdf_batch.argtypes = [C.c_int, _dp, C.c_int64, C.c_int, _dp, _dp, C.c_float,
                                    C.c_int, C.c_int, _dp]
        L.orc_pdf_batch_perpoint.restype = None
        L.orc_pdf_batch_perpoint.argtypes = [C.c_int, _dp, C.c_int64, C.c_int, _dp, _dp,
                                             C.c_float, C.c_int, _dp]
        L.orc_metropolis_hastings.restype = None
        L.orc_metropolis_hastings.argtypes = [_u32p, _dp, _dp, _u32p, C.c_int64, C.c_int]
        L.orc_propagate.restype = None
        L.orc_propagate.argtypes = [C.c_int, _dp, _dp, _u32p, _dp, _dp, _dp, _dp, _dp,
                                    C.c_int64, C.c_int]
        L.orc_reweight.restype = None
        L.orc_reweight.argtypes = [C.c_int, _dp, _dp, _dp, _dp, _dp, C.c_float, C.c_int64,
                                   C.c_int, C.c_int, C.c_int, C.c_int]
        L.orc_filter_metropolis.restype = None
        L.orc_filter_metropolis.argtypes = [C.c_int, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int,
                                            _dp, _dp, _dp, _dp, _dp, _dp, _dp, C.c_float,
                                            _dp, _dp, _u32p, _dp, _dp, _dp,
                                            _dp, _dp, _u32p, _dp, C.c_int]
        L.orc_quadform_fma.restype = C.c_double
        L.orc_quadform_fma.argtypes = [_dp, _dp, _dp, C.c_int, C.c_int, C.c_int]
        L.orc_det_exp.restype = C.c_double
        L.orc_det_exp.argtypes = [C.c_double]
        L.orc_det_log.restype = C.c_double
        L.orc_det_log.argtypes = [C.c_double]
        L.orc_philox4x32.restype = None
        L.orc_philox4x32.argtypes = [_u32p, _u32p, _u32p]
        L.orc_logsumexp_ess.restype = None
        L.orc_logsumexp_ess.argtypes = [_dp, C.c_int64, _dp, _dp, _dp]
        L.orc_fixed_shift.restype = C.c_int
        L.orc_fixed_shift.argtypes = [C.c_int64]
        L.orc_fixed_weights.restype = C.c_uint64
        L.orc_fixed_weights.argtypes = [_dp, C.c_int64, C.c_double, C.c_int, _u64p]
        L.orc_resample_systematic.restype = C.c_int
        L.orc_resample_systematic.argtypes = [_dp, C.c_int64, C.c_double, _u32p]
        L.orc_resample_multinomial.restype = C.c_int
        L.orc_resample_multinomial.argtypes = [_dp, C.c_int64, _dp, _u32p]
        L.orc_mh_chains.restype = None
        L.orc_mh_chains.argtypes = [C.c_int, C.c_int64, C.c_int, C.c_int, C.c_double, C.c_double,
                                    C.c_int, _dp, _dp, _dp, _dp, _dp, _dp, _u32p, _u8p]
        L.orc_resample_rejection.restype = None
        L.orc_resample_rejection.argtypes = [_u32p, _dp, C.c_double, C.c_int64, C.c_uint64, C.c_uint64, C.c_int]
        L.orc_mh_chains_general.restype = None
        L.orc_mh_chains_general.argtypes = [C.c_int, C.c_int64, C.c_int, C.c_int, C.c_double, _dp, C.c_double,
                                            C.c_int, _dp, _dp, _dp, _dp, _dp, _dp, _u32p, _u8p]
        L.orc_num_threads.restype = C.c_int
        L.orc_observation_operator.restype = C.c_int
        L.orc_observation_operator.argtypes = [C.c_int, C.c_int, C.c_int, _dp, _dp, C.c_float, _dp, _dp, _dp]
        L.orc_step_det.restype = None
        L.orc_step_det.argtypes = [C.c_int, C.c_int, _dp, _dp, _dp, _u32p, _dp, _dp, _dp, _dp, _dp,
                                   C.c_double, C.c_float, _dp, _dp, C.c_int64, C.c_int, C.c_int]
        L.orc_det_sincospi.restype = None
        L.orc_det_sincospi.argtypes = [C.c_double, _dp, _dp]
        L.orc_rng_normal_pair.restype = None
        L.orc_rng_normal_pair.argtypes = [C.c_uint64, C.c_int, C.c_uint64, C.c_uint64, C.c_uint32, _dp]
        L.orc_rng_u01.restype = C.c_double
        L.orc_rng_u01.argtypes = [C.c_uint64, C.c_int, C.c_uint64, C.c_uint64, C.c_uint32]
        L.orc_rng_metropolis.restype = None
        L.orc_rng_metropolis.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint64, _dp, _u32p]
        L.orc_rng_metropolis_c2.restype = None
        L.orc_rng_metropolis_c2.argtypes = L.orc_rng_metropolis.argtypes
        L.orc_rng_fill_normals.restype = None
        L.orc_rng_fill_normals.argtypes = [C.c_uint64, C.c_int, C.c_uint64, C.c_int64, C.c_int64, C.c_int, _dp]
        L.orc_filter_det.restype = C.c_int
        L.orc_filter_det.argtypes = [C.c_int, C.c_int, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int,
                                     _dp, _dp, _dp, _dp, _dp, _dp, _dp, C.c_float, C.c_uint64,
                                     _dp, _dp, _dp, _dp, _dp, _u32p, _dp, _dp,
                                     _dp, _dp, _u32p, _dp, _dp, C.c_double, C.POINTER(C.c_int), C.c_int64]
        L.orc_tile_image.restype = C.c_uint64
        L.orc_tile_image.argtypes = [_dp, C.c_int64, C.c_int64, C.c_int, _u64p, _u64p, _dp]

    # ---- dense helpers ----------------------------------------------------
    def determinant(self, A):
        A = np.asarray(A, dtype=np.float64)
        return self.lib.orc_determinant(_p(colmajor(A)), A.shape[0])

    def inverse(self, A):
        A = np.asarray(A, dtype=np.float64)
        d = A.shape[0]
        out = np.empty(d * d)
        self.lib.orc_inverse(_p(colmajor(A)), d, _p(out))
        return out.reshape(d, d, order="F")

    def cholesky_lower(self, A):
        A = np.asarray(A, dtype=np.float64)
        d = A.shape[0]
        out = np.empty(d * d)
        rc = self.lib.orc_cholesky_lower(_p(colmajor(A)), d, _p(out))
        if rc:
            raise np.linalg.LinAlgError("not positive definite at pivot %d" % (rc - 1))
        return out.reshape(d, d, order="F")

    def tri_inverse_lower(self, L):
        L = np.asarray(L, dtype=np.float64)
        d = L.shape[0]
        out = np.empty(d * d)
        self.lib.orc_tri_inverse_lower(_p(colmajor(L)), d, _p(out))
        return out.reshape(d, d, order="F")

    # ---- densities ----------------------------------------------------------
    def MVNPDF(self, x, mu, sigma):
        x, mu = f64(x), f64(mu)
        return self.lib.orc_MVNPDF(_p(x), _p(mu), _p(colmajor(sigma)), x.size)

    def MVTPDF(self, x, mu, sigma, nu):
        x, mu = f64(x), f64(mu)
        return self.lib.orc_MVTPDF(_p(x), _p(mu), _p(colmajor(sigma)), x.size, float(nu))

    def mvn_norm(self, sigma):
        sigma = np.asarray(sigma, dtype=np.float64)
        return self.lib.orc_mvn_norm(_p(colmajor(sigma)), sigma.shape[0])

    def mvt_norm(self, sigma, nu):
        sigma = np.asarray(sigma, dtype=np.float64)
        return self.lib.orc_mvt_norm(_p(colmajor(sigma)), sigma.shape[0], float(nu))

    def pdf_batch(self, dist, x, mu, sigma, nu=0.0, faithful=False, log=False):
        """x: (N, d) AoS.  dist: 'mvn' | 'mvt'."""
        x = f64(x)
        N, d = x.shape
        out = np.empty(N)
        mu_ = None if mu is None else f64(mu)
        self.lib.orc_pdf_batch(0 if dist == "mvn" else 1, _p(x), N, d, _p(mu_),
                               _p(colmajor(sigma)), float(nu), int(faithful), int(log), _p(out))
        return out

    def pdf_batch_perpoint(self, dist, x, mu_all, sigma_all, nu=0.0, log=False):
        """sigma_all: (N, d, d) symmetric; mu_all: (N, d) or None."""
        x = f64(x)
        N, d = x.shape
        out = np.empty(N)
        S = f64(np.transpose(np.asarray(sigma_all), (0, 2, 1)))  # per-point column-major
        mu_ = None if mu_all is None else f64(mu_all)
        self.lib.orc_pdf_batch_perpoint(0 if dist == "mvn" else 1, _p(x), N, d, _p(mu_), _p(S),
                                        float(nu), int(log), _p(out))
        return out

    # ---- resamplers ---------------------------------------------------------
    def metropolis_hastings(self, w, u, j):
        """u, j: (N, B) in the reference's consumption order."""
        w = f64(w)
        u = f64(u)
        j = np.ascontiguousarray(j, dtype=np.uint32)
        N, B = u.shape
        a = np.empty(N, dtype=np.uint32)
        self.lib.orc_metropolis_hastings(_p(a, _u32p), _p(w), _p(u), _p(j, _u32p), N, B)
        return a

    def resample_systematic(self, w, u0):
        w = f64(w)
        a = np.empty(w.size, dtype=np.uint32)
        rc = self.lib.orc_resample_systematic(_p(w), w.size, float(u0), _p(a, _u32p))
        return a, rc

    def resample_multinomial(self, w, u):
        w, u = f64(w), f64(u)
        a = np.empty(w.size, dtype=np.uint32)
        rc = self.lib.orc_resample_multinomial(_p(w), w.size, _p(u), _p(a, _u32p))
        return a, rc

    def resample_rejection(self, w, wmax, seed, step, cap=4096):
        w = f64(w)
        a = np.empty(w.size, dtype=np.uint32)
        self.lib.orc_resample_rejection(_p(a, _u32p), _p(w), float(wmax), w.size, int(seed), int(step), int(cap))
        return a

    def fixed_shift(self, n_global):
        return self.lib.orc_fixed_shift(int(n_global))

    def fixed_weights(self, w, wmax, shift):
        w = f64(w)
        q = np.empty(w.size, dtype=np.uint64)
        tot = self.lib.orc_fixed_weights(_p(w), w.size, float(wmax), int(shift), _p(q, _u64p))
        return q, tot

    def tile_image(self, lw, tile, shift=None):
        """Block-relative weight image: (C, T, T2, M)."""
        lw = f64(lw)
        shift = self.fixed_shift(lw.size) if shift is None else shift
        Cd = np.empty(lw.size, dtype=np.uint64)
        T2, M = C.c_uint64(), C.c_double()
        T = self.lib.orc_tile_image(_p(lw), lw.size, int(tile), int(shift), _p(Cd, _u64p), C.byref(T2), C.byref(M))
        return Cd, int(T), int(T2.value), M.value

    def logsumexp_ess(self, lw):
        lw = f64(lw)
        lse, ess, lmax = C.c_double(), C.c_double(), C.c_double()
        self.lib.orc_logsumexp_ess(_p(lw), lw.size, C.byref(lse), C.byref(ess), C.byref(lmax))
        return lse.value, ess.value, lmax.value

    # ---- propagate / reweight / filter ---------------------------------------
    def propagate(self, dist, x_prev, a, G, mu0, Q, xi, chi=None):
        xi = f64(xi)
        N, d = xi.shape
        out = np.empty((N, d))
        xp = None if x_prev is None else f64(x_prev)
        a_ = None if a is None else np.ascontiguousarray(a, dtype=np.uint32)
        G_ = None if G is None else colmajor(G)
        m_ = None if mu0 is None else f64(mu0)
        chi_ = None if chi is None else f64(chi)
        self.lib.orc_propagate(0 if dist == "mvn" else 1, _p(out), _p(xp), _p(a_, _u32p), _p(G_),
                               _p(m_), _p(colmajor(Q)), _p(xi), _p(chi_), N, d)
        return out

    def reweight(self, dist, y, x, F, V, nu=0.0, faithful=False, log=False):
        x = f64(x)
        N, d = x.shape
        F = np.asarray(F, dtype=np.float64)
        dy = F.shape[0]
        w = np.empty(N)
        self.lib.orc_reweight(0 if dist == "mvn" else 1, _p(w), _p(f64(y)), _p(x), _p(colmajor(F)),
                              _p(colmajor(V)), float(nu), N, d, dy, int(faithful), int(log))
        return w

    def filter_metropolis(self, dist, Y, m0, Q_c0, F, G, V, Q_w, nu, xi0, u, j, xi, chi=None,
                          history=True, chi0=None, faithful=False):
        """Y: (dy, T).  xi0: (N, d); u, j: (T-1, N, B); xi (chi): (T-1, N, d)."""
        Y = np.asarray(Y, dtype=np.float64)
        dy, T = Y.shape
        xi0 = f64(xi0)
        N, d = xi0.shape
        u = f64(u)
        B = u.shape[2]
        j = np.ascontiguousarray(j, dtype=np.uint32)
        xi = f64(xi)
        chi_ = None if chi is None else f64(chi)
        chi0_ = None if chi0 is None else f64(chi0)
        xh = np.zeros((T, N, d)) if history else None
        wh = np.zeros((T, N)) if history else None
        ah = np.zeros((T, N), dtype=np.uint32) if history else None
        mh = np.zeros((T, d))
        self.lib.orc_filter_metropolis(
            0 if dist == "mvn" else 1, N, d, dy, T, B, _p(colmajor(Y)), _p(f64(m0)),
            _p(colmajor(Q_c0)), _p(colmajor(F)), _p(colmajor(G)), _p(colmajor(V)),
            _p(colmajor(Q_w)), float(nu), _p(xi0), _p(u), _p(j, _u32p), _p(xi), _p(chi_), _p(chi0_),
            _p(xh), _p(wh), _p(ah, _u32p), _p(mh), int(faithful))
        return dict(x=xh, w=wh, a=ah, mean=mh)

    # ---- extended -------------------------------------------------------------
    def quadform_fma(self, M_rowmajor, c, v, tri):
        M = f64(M_rowmajor)
        m, d = M.shape
        c_ = None if c is None else f64(c)
        return self.lib.orc_quadform_fma(_p(M), _p(c_), _p(f64(v)), m, d, int(tri))

    def det_exp(self, x):
        return np.array([self.lib.orc_det_exp(float(v)) for v in np.atleast_1d(x)])

    def det_log(self, x):
        return np.array([self.lib.orc_det_log(float(v)) for v in np.atleast_1d(x)])

    def philox4x32(self, ctr, key):
        ctr = np.ascontiguousarray(ctr, dtype=np.uint32)
        key = np.ascontiguousarray(key, dtype=np.uint32)
        out = np.empty(4, dtype=np.uint32)
        self.lib.orc_philox4x32(_p(ctr, _u32p), _p(key, _u32p), _p(out, _u32p))
        return out

    def mh_chains(self, dist, mu, L, x0, z, thr, step, nu=0.0, shared=False, want_bits=True):
        """x0: (C, d); z: (C, steps, d); thr: (C, steps); L: (C, d, d) or (d, d) lower."""
        x0, z, thr = f64(x0), f64(z), f64(thr)
        Cn, d = x0.shape
        steps = z.shape[1]
        if shared:
            Lf = colmajor(L)
        else:
            Lf = f64(np.transpose(np.asarray(L), (0, 2, 1)))
        xf = np.empty((Cn, d))
        nacc = np.empty(Cn, dtype=np.uint32)
        bits = np.empty((Cn, steps), dtype=np.uint8) if want_bits else None
        self.lib.orc_mh_chains(0 if dist == "mvn" else 1, Cn, d, steps, float(step), float(nu),
                               int(shared), _p(f64(mu)), _p(Lf), _p(x0), _p(z), _p(thr),
                               _p(xf), _p(nacc, _u32p), _p(bits, _u8p))
        return xf, nacc, bits

    def mh_chains_general(self, dist, mu, L, x0, z, thr, step, nu=0.0, shared=False, scale=None, want_bits=True):
        """Isotropic / per-component-scaled random walk: every step evaluates the target density."""
        x0, z, thr = f64(x0), f64(z), f64(thr)
        Cn, d = x0.shape
        steps = z.shape[1]
        Lf = colmajor(L) if shared else f64(np.transpose(np.asarray(L), (0, 2, 1)))
        sc = None if scale is None else f64(scale)
        xf = np.empty((Cn, d))
        nacc = np.empty(Cn, dtype=np.uint32)
        bits = np.empty((Cn, steps), dtype=np.uint8) if want_bits else None
        self.lib.orc_mh_chains_general(0 if dist == "mvn" else 1, Cn, d, steps, float(step), _p(sc), float(nu),
                                       int(shared), _p(f64(mu)), _p(Lf), _p(x0), _p(z), _p(thr), _p(xf),
                                       _p(nacc, _u32p), _p(bits, _u8p))
        return xf, nacc, bits

    # ---- production-order restatements --------------------------------------------
    def observation_operator(self, dist, F, V, nu=0.0):
        F = np.asarray(F, dtype=np.float64)
        dy, d = F.shape
        M = np.empty((dy, d))
        Winv = np.empty((dy, dy))
        ln = C.c_double()
        rc = self.lib.orc_observation_operator(0 if dist == "mvn" else 1, d, dy, _p(colmajor(F)),
                                               _p(colmajor(V)), float(nu), _p(M), _p(Winv), C.byref(ln))
        if rc:
            raise np.linalg.LinAlgError("V not SPD")
        return M, Winv, ln.value

    def step_det(self, dist, x_prev, a, G, Q, y, F, V, xi, nu=0.0, chi=None, log=True, mu=None):
        """Fused propagate + reweight in production order.  Arrays AoS (N, d).  Returns x_new, lw."""
        xi = f64(xi)
        N, d = xi.shape
        M, Winv, ln = self.observation_operator(dist, F, V, nu)
        dy = M.shape[0]
        c = np.array([sum(Winv[k, i] * y[i] for i in range(k + 1)) for k in range(dy)])
        c = self._whiten_y(Winv, f64(y))
        xn = np.empty((N, d))
        lw = np.empty(N)
        xp = None if x_prev is None else f64(x_prev)
        a_ = None if a is None else np.ascontiguousarray(a, dtype=np.uint32)
        G_ = None if G is None else colmajor(G)
        mu_ = None if mu is None else f64(mu)
        chi_ = None if chi is None else f64(chi)
        self.lib.orc_step_det(0 if dist == "mvn" else 1, int(log), _p(xn), _p(lw), _p(xp), _p(a_, _u32p),
                              _p(G_), _p(colmajor(Q)), _p(mu_), _p(f64(M)), _p(c), ln, float(nu), _p(xi),
                              _p(chi_), N, d, dy)
        return xn, lw

    @staticmethod
    def _whiten_y(Winv, y):
        dy = Winv.shape[0]
        c = np.zeros(dy)
        for k in range(dy):
            s = 0.0
            for i in range(k + 1):
                s += Winv[k, i] * y[i]      # python floats: IEEE double, no contraction
            c[k] = s
        return c

    def det_sincospi(self, t):
        s, c = C.c_double(), C.c_double()
        self.lib.orc_det_sincospi(float(t), C.byref(s), C.byref(c))
        return s.value, c.value

    def rng_fill_normals(self, seed, stream, step, i0, N, d):
        xi = np.empty((N, d))
        self.lib.orc_rng_fill_normals(int(seed), int(stream), int(step), int(i0), int(N), int(d), _p(xi))
        return xi

    def rng_u01(self, seed, stream, step, index, sub=0):
        return self.lib.orc_rng_u01(int(seed), int(stream), int(step), int(index), int(sub))

    def rng_metropolis(self, seed, step, N, B, c2=False):
        """(u, j) [N][B] the Metropolis resampler draws on the device; c2: the Metropolis-C2 proposals."""
        u = np.empty((N, B))
        j = np.empty((N, B), dtype=np.uint32)
        uu, jj = C.c_double(), C.c_uint32()
        draw = self.lib.orc_rng_metropolis_c2 if c2 else self.lib.orc_rng_metropolis
        for i in range(N):
            for n in range(B):
                draw(int(seed), int(step), i, n, N, C.byref(uu), C.byref(jj))
                u[i, n], j[i, n] = uu.value, jj.value
        return u, j

    def filter_det(self, dist, resampler, Y, m0, Q_c0, F, G, V, Q_w, N, nu=0.0, seed=0, B=10,
                   xi0=None, xi=None, chi=None, u=None, j=None, u0=None, um=None, ess_threshold=0.0,
                   chi0=None, tile=0):
        """Production-order filter.  resampler: 'metropolis' | 'systematic' | 'multinomial'."""
        Y = np.asarray(Y, dtype=np.float64)
        dy, T = Y.shape
        d = np.asarray(G).shape[0]
        rs = {"metropolis": 0, "systematic": 1, "multinomial": 2, "rejection": 3, "metropolis_c2": 4}[resampler]
        opt = lambda a: None if a is None else f64(a)
        j_ = None if j is None else np.ascontiguousarray(j, dtype=np.uint32)
        xh = np.zeros((T, N, d))
        wh = np.zeros((T, N))
        ah = np.zeros((T, N), dtype=np.uint32)
        ah[0] = np.arange(N)
        ess = np.full(T, np.nan)
        ll = np.full(T, np.nan)
        res = np.zeros(T, dtype=np.int32)
        rc = self.lib.orc_filter_det(
            0 if dist == "mvn" else 1, rs, N, d, dy, T, B, _p(colmajor(Y)), _p(f64(m0)),
            _p(colmajor(Q_c0)), _p(colmajor(F)), _p(colmajor(G)), _p(colmajor(V)), _p(colmajor(Q_w)),
            float(nu), int(seed), _p(opt(xi0)), _p(opt(chi0)), _p(opt(xi)), _p(opt(chi)), _p(opt(u)), _p(j_, _u32p),
            _p(opt(u0)), _p(opt(um)), _p(xh), _p(wh), _p(ah, _u32p), _p(ess), _p(ll),
            float(ess_threshold), res.ctypes.data_as(C.POINTER(C.c_int)), int(tile))
        if rc:
            raise RuntimeError("orc_filter_det failed: %d" % rc)
        return dict(x=xh, w=wh, a=ah, ess=ess, loglik=ll, resampled=res)

    def num_threads(self):
        return self.lib.orc_num_threads()

    def use_all_cores(self):
        """torchrun exports OMP_NUM_THREADS=1; a CPU baseline is timed on every host core."""
        import os
        n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
        self.lib.orc_set_num_threads(int(n))
        return self.num_threads()


_ORACLE = None


def oracle():
    global _ORACLE
    if _ORACLE is None:
        _ORACLE = Oracle()
    return _ORACLE

"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle.

Tolerances (BASELINE.json north_star): fp64 log-densities within 1e-10 relative; accept/reject
decisions and ancestor indices bit-exact on identical inputs; posterior moments within a stated
Monte Carlo tolerance.
"""
import json
import math
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = json.load(open(os.path.join(HERE, "golden", "golden.json")))
RTOL_LOGPDF = 1e-10


def spd(rng, d):
    A = rng.standard_normal((d, d))
    return A @ A.T / d + np.eye(d)


def relerr(got, want):
    return np.max(np.abs(got - want) / np.maximum(np.abs(want), 1e-300))


def torch_dev(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


# ------------------------------------------------------------------------------------------------
# a1 / a2: densities
# ------------------------------------------------------------------------------------------------
def test_reference_known_answers(ctx):
    import cusmc_b200
    ref = GOLDEN["reference"]
    v = cusmc_b200.MVNPDF(ref["MVNPDF"]["x"], ref["MVNPDF"]["mu"], np.array(ref["MVNPDF"]["sigma"], float))
    assert abs(v - ref["MVNPDF"]["value"]) < 5e-8          # 0.1591549, 7 significant digits
    assert abs(v - 1 / (2 * math.pi)) < 1e-15
    v = cusmc_b200.MVTPDF(ref["MVTPDF"]["x"], ref["MVTPDF"]["mu"], np.array(ref["MVTPDF"]["sigma"], float),
                          ref["MVTPDF"]["nu"])
    assert abs(v - ref["MVTPDF"]["value"]) < 5e-9          # 0.07799708
    a = cusmc_b200.metropolis_hastings(ref["metropolis_hastings"]["w"], 2, 10)
    assert a.dtype == np.float64 and a.tolist() == [0.0, 1.0]   # 0/0 = NaN never accepts


def test_independent_golden_cases(ctx):
    for c in GOLDEN["independent"]:
        x = np.array(c["x"])
        sigma = np.array(c["sigma"])
        got = ctx.logpdf(c["kind"], x, c["mu"], sigma, nu=c["nu"], log=True)
        assert relerr(got, np.array(c["logpdf"])) < RTOL_LOGPDF, (c["kind"], c["d"])
        got = ctx.logpdf(c["kind"], x, c["mu"], sigma, nu=c["nu"], log=False)
        assert relerr(got, np.array(c["pdf"])) < 1e-10, (c["kind"], c["d"])


@pytest.mark.parametrize("kind,nu", [("mvn", 0.0), ("mvt", 5.0), ("mvt", 0.7)])
@pytest.mark.parametrize("d", [1, 2, 3, 5, 8, 11, 16, 24, 32])
def test_logpdf_shared_vs_oracle(ctx, orc, kind, nu, d):
    rng = np.random.default_rng(100 + d)
    N = 3001                               # odd: exercises the scalar (VEC = 1) SoA path and tile tails
    sigma, mu = spd(rng, d), rng.standard_normal(d)
    x = rng.standard_normal((N, d)) * 2.0 + mu
    want = orc.pdf_batch(kind, x, mu, sigma, nu, faithful=True, log=False)   # the reference's own form
    want_log = orc.pdf_batch(kind, x, mu, sigma, nu, log=True)
    got = ctx.logpdf(kind, x, mu, sigma, nu=nu, log=True)                    # AoS host entry point
    assert relerr(got, want_log) < RTOL_LOGPDF
    assert relerr(np.exp(got), want) < 1e-10
    assert relerr(ctx.logpdf(kind, x, mu, sigma, nu=nu, log=False), want) < 1e-10
    # SoA device entry point, both vector widths
    import torch
    for n in (N, N - 1):
        xs = torch_dev(x[:n].T)
        out = torch.empty(n, dtype=torch.float64, device="cuda")
        ctx.logpdf_dev(kind, xs, mu, sigma, out, nu=nu, log=True)
        ctx.synchronize()
        assert relerr(out.cpu().numpy(), want_log[:n]) < RTOL_LOGPDF


def test_logpdf_layouts_agree_bitwise(ctx):
    import torch
    rng = np.random.default_rng(5)
    d, N = 16, 8192
    sigma, mu = spd(rng, d), rng.standard_normal(d)
    x = rng.standard_normal((N, d))
    aos = ctx.logpdf("mvn", x, mu, sigma)
    soa = ctx.logpdf("mvn", np.ascontiguousarray(x.T), mu, sigma, layout=0)
    assert np.array_equal(aos, soa)        # same operation order in both kernels


def test_logpdf_edge_cases(ctx, orc):
    import cusmc_b200
    rng = np.random.default_rng(6)
    sigma = spd(rng, 4)
    assert ctx.logpdf("mvn", np.zeros((0, 4)), None, sigma).shape == (0,)
    x1 = rng.standard_normal((1, 4))
    assert relerr(ctx.logpdf("mvn", x1, None, sigma), orc.pdf_batch("mvn", x1, None, sigma, log=True)) < 1e-12
    with pytest.raises(cusmc_b200.CusmcError) as e:
        ctx.logpdf("mvn", x1, None, -np.eye(4))
    assert e.value.code == 3               # CUSMC_ERR_NOT_SPD, not NaNs
    with pytest.raises(cusmc_b200.CusmcError):
        ctx.logpdf("mvt", x1, None, sigma, nu=0.0)
    with pytest.raises(cusmc_b200.CusmcError) as e:
        ctx.logpdf("mvn", np.zeros((2, 33)), None, np.eye(33))
    assert e.value.code == 5               # d > CUSMC_MAX_DIM
    with pytest.raises(ValueError):
        ctx.logpdf("normal", x1, None, sigma)
    # far tail: log-density stays finite where the reference's density underflows to 0 (Q8)
    far = np.full((2, 4), 60.0)
    assert np.all(np.isfinite(ctx.logpdf("mvn", far, None, sigma)))
    assert np.all(ctx.logpdf("mvn", far, None, sigma, log=False) == 0.0)


@pytest.mark.parametrize("N", [(1 << 17) + 1, 3 * (1 << 17) + 777])
def test_logpdf_host_call_is_chunk_invariant(ctx, orc, N):
    """cusmc_logpdf pipelines 2^17-point chunks over two streams: ragged last chunk, both layouts,
    every point against the oracle and AoS == SoA bit for bit."""
    rng = np.random.default_rng(N)
    d = 5
    sigma, mu = spd(rng, d), rng.standard_normal(d)
    x = rng.standard_normal((N, d))
    got_aos = ctx.logpdf("mvt", x, mu, sigma, nu=4.0)
    got_soa = ctx.logpdf("mvt", np.ascontiguousarray(x.T), mu, sigma, nu=4.0, layout=0)
    assert np.array_equal(got_aos, got_soa)
    want = orc.pdf_batch("mvt", x, mu, sigma, nu=4.0, log=True)
    assert relerr(got_aos, want) < 1e-10


def test_logpdf_full_size_properties(ctx, orc):
    """BASELINE config: N = 2^20 points, d = 16, shared covariance.  Size-independent checks:
    translation invariance, a checksum against the oracle on a strided sample, agreement of the
    SoA and AoS kernels."""
    import torch
    rng = np.random.default_rng(1234)
    N, d = 1 << 20, 16
    sigma, mu = spd(rng, d), rng.standard_normal(d)
    x = rng.standard_normal((N, d))
    xs = torch_dev(x.T)
    out = torch.empty(N, dtype=torch.float64, device="cuda")
    ctx.logpdf_dev("mvn", xs, mu, sigma, out)
    ctx.synchronize()
    got = out.cpu().numpy()
    idx = np.arange(0, N, 997)
    want = orc.pdf_batch("mvn", x[idx], mu, sigma, log=True)
    assert relerr(got[idx], want) < RTOL_LOGPDF
    shift = rng.standard_normal(d)
    out2 = torch.empty_like(out)
    ctx.logpdf_dev("mvn", xs + torch_dev(shift)[:, None], mu + shift, sigma, out2)
    ctx.synchronize()
    assert relerr(out2.cpu().numpy(), got) < 1e-9
    xa = torch_dev(x)
    out3 = torch.empty_like(out)
    ctx.logpdf_dev("mvn", xa, mu, sigma, out3, layout=1)
    ctx.synchronize()
    assert np.array_equal(out3.cpu().numpy(), got)


@pytest.mark.parametrize("kind,nu", [("mvn", 0.0), ("mvt", 4.0)])
@pytest.mark.parametrize("d", [2, 3, 8, 32])
def test_logpdf_perpoint(ctx, orc, kind, nu, d):
    import torch
    rng = np.random.default_rng(40 + d)
    N = 777
    sig = np.stack([spd(rng, d) for _ in range(N)])
    mu = rng.standard_normal((N, d))
    x = mu + rng.standard_normal((N, d))
    Ls = np.linalg.cholesky(sig)
    tril = np.tril_indices(d)
    packed = np.ascontiguousarray(Ls[:, tril[0], tril[1]])       # row-packed lower factors
    out = torch.empty(N, dtype=torch.float64, device="cuda")
    ctx.logpdf_perpoint_dev(kind, torch_dev(x), torch_dev(mu), torch_dev(packed), out, nu=nu, log=True)
    ctx.synchronize()
    want = orc.pdf_batch_perpoint(kind, x, mu, sig, nu, log=True)
    assert relerr(out.cpu().numpy(), want) < RTOL_LOGPDF


# ------------------------------------------------------------------------------------------------
# drop-ins for the reference wrappers
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("d,dy", [(2, 2), (4, 2), (3, 5), (8, 8)])
def test_pdf_wrappers_vs_reweight(ctx, orc, d, dy):
    rng = np.random.default_rng(d * 10 + dy)
    N = 2500
    F = rng.standard_normal((dy, d))
    V = spd(rng, dy)
    y = rng.standard_normal(dy)
    x = rng.standard_normal((N, d))
    Vinv = orc.inverse(V)
    got = ctx.mvn_pdf(y, x, orc.mvn_norm(V), Vinv, F)
    want = orc.reweight("mvn", y, x, F, V, faithful=True)        # src/mcmc.cpp:193-215 as written
    assert relerr(got, want) < 1e-10
    got = ctx.mvt_pdf(y, x, Vinv, F, orc.mvt_norm(V, 3.5), 3.5)
    want = orc.reweight("mvt", y, x, F, V, nu=3.5, faithful=True)
    assert relerr(got, want) < 1e-10


@pytest.mark.parametrize("d", [2, 5, 8])
def test_sample_wrappers(ctx, orc, d):
    rng = np.random.default_rng(70 + d)
    N = 1500
    G, Q = rng.standard_normal((d, d)), rng.standard_normal((d, d))
    xp, xi = rng.standard_normal((N, d)), rng.standard_normal((N, d))
    chi = np.sqrt(4.0 / rng.chisquare(4.0, (N, d)))
    a = rng.integers(0, N, N, dtype=np.uint32)
    got = ctx.mvn_sample(xp, a, G, Q, xi=xi)
    assert np.allclose(got, orc.propagate("mvn", xp, a, G, None, Q, xi), rtol=1e-13, atol=1e-13)
    want_det, _ = orc.step_det("mvn", xp, a, G, Q, np.zeros(d), np.eye(d), np.eye(d), xi)
    assert np.array_equal(got, want_det)                         # production order: bit for bit
    got = ctx.mvt_sample(xp, a, G, Q, 4.0, xi=xi, chi=chi)
    assert np.allclose(got, orc.propagate("mvt", xp, a, G, None, Q, xi, chi), rtol=1e-13, atol=1e-13)
    mu = rng.standard_normal(d)
    got = ctx.mvn_sample_init(mu, Q, N, xi=xi)
    assert np.allclose(got, orc.propagate("mvn", None, None, None, mu, Q, xi), rtol=1e-13, atol=1e-13)
    # device-drawn noise is the Philox stream the oracle mirrors
    got = ctx.mvn_sample(xp, a, G, Q, seed=99, step=3)
    xi_dev = orc.rng_fill_normals(99, 1, 3, 0, N, d)
    want_det, _ = orc.step_det("mvn", xp, a, G, Q, np.zeros(d), np.eye(d), np.eye(d), xi_dev)
    assert np.array_equal(got, want_det)
    # chi-square factors drawn on the device: right law (E[chi^-2] = 1, i.e. chi2/nu has mean 1)
    got = ctx.mvt_sample(np.zeros((20000, 1)), None, np.zeros((1, 1)), np.ones((1, 1)), 6.0, xi=np.ones((20000, 1)), seed=5)
    inv = 1.0 / got[:, 0] ** 2
    assert abs(inv.mean() - 1.0) < 0.02 and abs(inv.var() - 2.0 / 6.0) < 0.03
    # ... and the whole law, for a shape above and one below 1 (nu = 1.5 takes the U^(1/a) boost) and
    # for both components of a pair: Kolmogorov-Smirnov against chi^2_nu
    from scipy import stats
    for nu in (6.0, 1.5, 30.0):
        n = 100000
        got = ctx.mvt_sample(np.zeros((n, 2)), None, np.zeros((2, 2)), np.eye(2), nu, xi=np.ones((n, 2)), seed=int(nu * 10))
        for k in range(2):
            x2 = nu / got[:, k] ** 2
            assert stats.kstest(x2, "chi2", args=(nu,)).pvalue > 1e-3, (nu, k)
        assert abs(np.corrcoef(got[:, 0], got[:, 1])[0, 1]) < 0.02      # the pair's factors are independent


# ------------------------------------------------------------------------------------------------
# a4: Metropolis resampler
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("N,B", [(1, 3), (2, 10), (1000, 10), (65537, 7)])
def test_metropolis_bit_exact(ctx, orc, N, B):
    rng = np.random.default_rng(N + B)
    w = rng.random(N) ** 4
    w[rng.random(N) < 0.1] = 0.0                                # zero weights: 0/0, x/0 branches
    u, j = rng.random((N, B)), rng.integers(0, N, (N, B), dtype=np.uint32)
    assert np.array_equal(ctx.metropolis_hastings(w, B, u=u, j=j), orc.metropolis_hastings(w, u, j))


def test_metropolis_invariants(ctx, orc):
    rng = np.random.default_rng(3)
    N, B = 512, 10
    u, j = rng.random((N, B)) * 0.999 + 0.001, rng.integers(0, N, (N, B), dtype=np.uint32)
    assert np.array_equal(ctx.metropolis_hastings(np.ones(N), B, u=u, j=j), j[:, -1])   # constant w
    m = 17
    onehot = np.zeros(N)
    onehot[m] = 1.0
    a = ctx.metropolis_hastings(onehot, B, u=u, j=j)
    hit = (j == m).any(axis=1)
    assert np.array_equal(a[hit], np.full(hit.sum(), m))
    others = ~hit
    others[m] = False
    # particles that never draw m keep k = i ... unless k = i has weight 0 and j has weight 0: 0/0 rejects
    assert np.array_equal(a[others], np.arange(N, dtype=np.uint32)[others])
    assert np.array_equal(ctx.metropolis_hastings(np.zeros(2), 10), [0, 1])


def test_metropolis_near_ties_take_the_reference_decision(ctx, orc):
    """u <= w_j / w_k with u built ON the boundary (equal to the IEEE quotient and its neighbours) and
    zero / denormal / huge / infinite / NaN weights in play: the reference's decision, bit for bit."""
    rng = np.random.default_rng(8)
    N, B = 4096, 6
    w = rng.random(N) ** 3
    w[rng.random(N) < 0.05] = 0.0
    w[5], w[6], w[7], w[8], w[9] = 1e-310, 3e-300, 1e305, np.inf, np.nan
    j = rng.integers(0, N, (N, B), dtype=np.uint32)
    j[::7, 0] = rng.integers(5, 10, j[::7, 0].shape)           # hit the special weights often
    u = rng.random((N, B))
    # walk the chain on the host to know w_k at every test, then put u on / next to the quotient
    k = np.arange(N)
    with np.errstate(all="ignore"):
        for n in range(B):
            q = w[j[:, n]] / w[k]
            mode = rng.integers(0, 4, N)
            un = u[:, n].copy()
            ok = np.isfinite(q) & (q > 0) & (q < 1)
            un[ok & (mode == 0)] = q[ok & (mode == 0)]
            un[ok & (mode == 1)] = np.nextafter(q[ok & (mode == 1)], 2.0)
            un[ok & (mode == 2)] = np.nextafter(q[ok & (mode == 2)], 0.0)
            u[:, n] = un
            acc = un <= q
            k = np.where(acc, j[:, n], k)
    want = orc.metropolis_hastings(w, u, j)
    assert np.mean(want == k.astype(np.uint32)) > 0.99           # the host walk follows the same rule
    assert np.array_equal(ctx.metropolis_hastings(w, B, u=u, j=j), want)


def test_metropolis_philox_mirror(ctx, orc):
    N, B = 300, 10
    w = np.random.default_rng(8).random(N)
    u, j = orc.rng_metropolis(1234, 5, N, B)
    assert np.array_equal(ctx.metropolis_hastings(w, B, seed=1234, step=5), orc.metropolis_hastings(w, u, j))


# ------------------------------------------------------------------------------------------------
# a11: normalisation, ESS, systematic / multinomial resampling
# ------------------------------------------------------------------------------------------------
def _weight_cases(rng, N):
    yield "uniform", np.ones(N)
    yield "random", rng.random(N)
    yield "heavy", rng.random(N) ** 20
    w = np.zeros(N)
    w[N // 3] = 1.0
    yield "onehot", w
    w = rng.random(N)
    w[rng.random(N) < 0.9] = 0.0
    w[-1] = 0.5
    yield "sparse", w
    yield "tiny", rng.random(N) * 1e-300
    if N >= 3:
        w = rng.random(N)
        w[0] = np.nan
        w[1] = -1.0
        yield "nan_neg", w


@pytest.mark.parametrize("N", [1, 2, 255, 2048, 2049, 10000, 300001])
def test_systematic_bit_exact(ctx, orc, N):
    rng = np.random.default_rng(N)
    for name, w in _weight_cases(rng, N):
        for u0 in (0.0, 0.37, 0.999999999):
            want, rc = orc.resample_systematic(w, u0)
            assert rc == 0
            got = ctx.resample_systematic(w, u0)
            assert np.array_equal(got, want), (name, N, u0)


def test_systematic_properties_large(ctx):
    rng = np.random.default_rng(11)
    N = (1 << 22) + 12345                   # many tiles: look-back windows beyond 32 tiles
    w = rng.random(N) ** 3
    a = ctx.resample_systematic(w, 0.5).astype(np.int64)
    assert np.all(np.diff(a) >= 0)          # sorted
    counts = np.bincount(a, minlength=N)
    expect = N * w / w.sum()
    assert counts.sum() == N
    assert np.all(np.abs(counts - expect) < 1.0 + 1e-6)   # systematic: floor or ceil of N w / sum w


@pytest.mark.parametrize("N", [1 << 20, 1000003])
def test_systematic_equal_weights_is_the_identity(ctx, N):
    """Equal weights put EVERY offspring boundary on (or within rounding of) an integer -- the case the
    floating estimate of the scatter pass must hand to its exact 128-bit path: a_i = i for any offset."""
    w = np.full(N, 0.125)
    for u0 in (0.0, 1e-12, 0.5, 1.0 - 1e-12):
        a = ctx.resample_systematic(w, u0)
        assert np.array_equal(a, np.arange(N, dtype=np.uint32)), u0


def test_resample_degenerate_is_an_error(ctx):
    import cusmc_b200
    with pytest.raises(cusmc_b200.CusmcError) as e:
        ctx.resample_systematic(np.zeros(100), 0.3)
    assert e.value.code == 4
    with pytest.raises(cusmc_b200.CusmcError):
        ctx.resample_multinomial(np.full(10, np.nan), np.full(10, 0.5))
    with pytest.raises(cusmc_b200.CusmcError):
        ctx.resample_systematic(np.ones(4), 1.0)      # u0 outside [0, 1)


@pytest.mark.parametrize("N", [1, 100, 5000, 200003])
def test_multinomial_bit_exact(ctx, orc, N):
    rng = np.random.default_rng(N + 1)
    for name, w in _weight_cases(rng, N):
        u = rng.random(N)
        u[0] = 0.0
        want, rc = orc.resample_multinomial(w, u)
        assert np.array_equal(ctx.resample_multinomial(w, u), want), (name, N)


def test_normalize_ess(ctx, orc):
    rng = np.random.default_rng(21)
    for N in (1, 1000, 100000):
        lw = rng.standard_normal(N) * 30 - 700.0     # the linear domain would underflow here
        lse, ess = ctx.normalize_ess(lw)
        lse0, ess0, _ = orc.logsumexp_ess(lw)
        assert abs(lse - lse0) < 1e-9 * abs(lse0)
        assert abs(ess - ess0) < 1e-6 * ess0 + 1e-9
    lse, ess = ctx.normalize_ess(np.array([-np.inf, 0.0, -np.inf]))
    assert abs(lse) < 1e-12 and abs(ess - 1.0) < 1e-9


# ------------------------------------------------------------------------------------------------
# fused propagate + reweight, and the whole filter
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("kind,nu", [("mvn", 0.0), ("mvt", 5.0)])
@pytest.mark.parametrize("d,dy", [(2, 2), (8, 8), (5, 3), (16, 16)])
def test_propagate_reweight_dev(ctx, orc, kind, nu, d, dy):
    import torch
    rng = np.random.default_rng(d + dy)
    N = 4099
    G, Q, F, V = rng.standard_normal((d, d)) * 0.5, rng.standard_normal((d, d)) * 0.3, rng.standard_normal((dy, d)), spd(rng, dy)
    y = rng.standard_normal(dy)
    xp, xi = rng.standard_normal((N, d)), rng.standard_normal((N, d))
    chi = np.sqrt(nu / rng.chisquare(nu, (N, d))) if kind == "mvt" else None
    a = rng.integers(0, N, N, dtype=np.uint32)
    x_new = torch.empty((d, N), dtype=torch.float64, device="cuda")
    lw = torch.empty(N, dtype=torch.float64, device="cuda")
    mx = torch.full((1,), -np.inf, dtype=torch.float64, device="cuda")
    ctx.propagate_reweight_dev(kind, x_new, torch_dev(xp.T), torch_dev(a), G, Q, y, F, V, lw, nu=nu,
                               xi=torch_dev(xi.T), chi=None if chi is None else torch_dev(chi.T), lw_max=mx)
    ctx.synchronize()
    want_x, want_lw = orc.step_det(kind, xp, a, G, Q, y, F, V, xi, nu=nu, chi=chi)
    assert np.array_equal(x_new.cpu().numpy().T, want_x)
    if kind == "mvn":
        assert np.array_equal(lw.cpu().numpy(), want_lw)
    else:
        assert relerr(lw.cpu().numpy(), want_lw) < 1e-12
    assert mx.item() == lw.max().item()
    # and against the reference-form arithmetic (LU inverse, left-to-right quadratic form)
    ref_x = orc.propagate(kind, xp, a, G, None, Q, xi, chi)
    ref_lw = orc.reweight(kind, y, ref_x, F, V, nu=nu, log=True)
    assert relerr(lw.cpu().numpy(), ref_lw) < RTOL_LOGPDF


def _model(d, rng=None):
    I = np.eye(d)
    return dict(m0=np.zeros(d), C0=I, F=I, G=0.9 * I, V=0.5 * I, W=0.3 * I)


def _eig_factor(S):
    lam, vec = np.linalg.eigh(S)
    return vec * np.sqrt(lam)


def _raw_weights(h, resampler):
    """What resampling consumed: densities in reference mode, log-weights otherwise."""
    return h["w"] if resampler == "metropolis" else h["lw"]


@pytest.mark.parametrize("kind,nu", [("mvn", 0.0), ("mvt", 5.0)])
@pytest.mark.parametrize("resampler", ["systematic", "multinomial", "metropolis"])
@pytest.mark.parametrize("d", [2, 8])
def test_filter_bit_exact_vs_oracle(ctx, orc, resampler, d, kind, nu):
    """Same pre-drawn normals / uniforms (and, for "mvt", chi factors -- the initial draw's included,
    ref: src/mcmc.cpp:73-79) on both sides: ancestors and particle states must agree bit for bit at
    every step; log-weights too (MVN), ESS and log-likelihood to rounding."""
    rng = np.random.default_rng(500 + d)
    N, T, B = 3000, 12, 10
    md = _model(d)
    Y = rng.standard_normal((d, T))
    xi0, xi = rng.standard_normal((N, d)), rng.standard_normal((T - 1, N, d))
    u, j = rng.random((T - 1, N, B)), rng.integers(0, N, (T - 1, N, B), dtype=np.uint32)
    u0, um = rng.random(T - 1), rng.random((T - 1, N))
    chi0 = chi = None
    kw = {}
    if kind == "mvt":
        chi0 = np.sqrt(nu / rng.chisquare(nu, (N, d)))
        chi = np.sqrt(nu / rng.chisquare(nu, (T - 1, N, d)))
        kw = dict(chi0=torch_dev(chi0.T), chi=torch_dev(chi.transpose(0, 2, 1)))
    pf = ctx.filter(N=N, Y=Y, resampler=resampler, B=B, keep_history=True, distribution=kind, df=nu, **md)
    pf.run(xi0=torch_dev(xi0.T), xi=torch_dev(xi.transpose(0, 2, 1)), u=torch_dev(u), j=torch_dev(j),
           u0=u0, um=torch_dev(um), **kw)
    h, s = pf.history(), pf.summary()
    pf.close()
    # the library uses an eigen factor of C0 / W; for scalar covariances it is sqrt(c) I up to
    # sign conventions of the Jacobi sweep (none applied to a diagonal matrix)
    ref = orc.filter_det(kind, resampler, Y, md["m0"], _eig_factor(md["C0"]), md["F"], md["G"], md["V"],
                         _eig_factor(md["W"]), N, nu=nu, B=B, xi0=xi0, xi=xi, u=u, j=j, u0=u0, um=um,
                         chi0=chi0, chi=chi)
    assert np.array_equal(h["a"], ref["a"])
    assert np.array_equal(h["x"], ref["x"])
    if kind == "mvt":
        # x_0 really carries the chi factors: m0 + chi0 (.) (Q_c0 xi0)
        assert np.array_equal(h["x"][0], md["m0"] + chi0 * (xi0 @ _eig_factor(md["C0"]).T))
    raw = _raw_weights(h, resampler)
    if resampler == "metropolis":
        assert relerr(raw, ref["w"]) < 1e-13          # densities: CUDA exp / pow vs libm
        assert np.all(raw[0] == 1.0 / N)              # w_0 = 1/N (src/mcmc.cpp:85)
    else:
        if kind == "mvn":
            assert np.array_equal(raw, ref["w"])
        else:
            assert np.allclose(raw, ref["w"], rtol=1e-12, atol=1e-12)   # log1p: CUDA vs libm
        assert np.allclose(s["ess"], ref["ess"], rtol=1e-9 if kind == "mvt" else 1e-12)
        assert np.allclose(s["loglik"], ref["loglik"], rtol=1e-12, atol=1e-12)
        # the history hands out NORMALISED weights (1/N at t = 0)
        wn = np.exp(raw - raw.max(axis=1, keepdims=True))
        wn /= wn.sum(axis=1, keepdims=True)
        assert np.allclose(h["w"], wn, rtol=1e-9, atol=1e-300)
        assert np.allclose(h["w"][0], 1.0 / N, rtol=1e-12)
    wn = np.exp(raw - raw.max(axis=1, keepdims=True)) if resampler != "metropolis" else raw
    mean = (wn[:, :, None] * h["x"]).sum(1) / wn.sum(1)[:, None]
    assert np.allclose(s["mean"], mean, rtol=1e-10, atol=1e-12)


@pytest.mark.parametrize("reproducible", [False, True])
@pytest.mark.parametrize("nu,d", [(5.0, 8), (2.5, 3), (1.5, 2), (1.0, 4), (2.0, 2), (4.0, 8), (8.0, 3), (13.0, 4), (7.5, 8)])
def test_mvt_device_noise_is_student_t(ctx, reproducible, nu, d):
    """Device-drawn "mvt" noise inside the filter, both generators (the throughput one draws its chi factors
    without rejection for integer nu <= 8 -- sums of exponentials and half a squared normal, chi_halfint -- and
    with chi_fast for the other nu >= 2): with G = 0 and W = C0 = I every component of x_0 and x_1 is chi z ~ t_nu.
    Kolmogorov-Smirnov per component, and the chi factors of different components are independent."""
    from scipy import stats
    N, T = 120000, 2
    I = np.eye(d)
    Y = np.zeros((d, T))
    pf = ctx.filter(N=N, Y=Y, m0=np.zeros(d), C0=I, F=I, G=np.zeros((d, d)), V=100.0 * I, W=I, resampler="systematic",
                    distribution="mvt", df=nu, seed=77, keep_history=True, reproducible_rng=reproducible)
    h = pf.run().history()
    pf.close()
    for t in range(T):
        x = h["x"][t]
        assert np.all(np.isfinite(x))
        for k in range(d):
            # ~170 KS tests over the parametrisation: 1e-4 keeps the family-wise false-alarm rate below 2 %
            assert stats.kstest(x[:, k], "t", args=(nu,)).pvalue > 1e-4, (t, k)
        # |x_k| of two components would correlate if they shared a chi factor
        r = np.corrcoef(np.log(np.abs(x) + 1e-300).T)
        assert np.all(np.abs(r - np.eye(d)) < 0.02), (t, r)
    assert not np.array_equal(h["x"][0], h["x"][1])


@pytest.mark.parametrize("persistent", [False, True])
def test_fast_normal_noise_d2_pairs(ctx, persistent):
    """Throughput noise at d = 2: the neighbours 2p, 2p + 1 share one Philox block (first_block(), pf_particle.cuh).
    With G = 0, W = C0 = I every x_t is the raw noise: standard Normal per component (KS), components and
    NEIGHBOURS uncorrelated (also in squares), steps differ, and the one-kernel run draws the same normals
    as the per-step path."""
    from scipy import stats
    N, T, d = 200000, 3, 2
    I = np.eye(d)
    Y = np.zeros((d, T))
    pf = ctx.filter(N=N, Y=Y, m0=np.zeros(d), C0=I, F=I, G=np.zeros((d, d)), V=1e6 * I, W=I, resampler="systematic",
                    seed=123, summary=False, persistent=persistent)
    pf.run()
    x, _, _ = pf.state()
    pf.close()
    x = x.T                                                        # [N, d], the last step's cloud
    assert np.all(np.isfinite(x))
    for k in range(d):
        assert stats.kstest(x[:, k], "norm").pvalue > 1e-3, k
    cols = np.stack([x[0::2, 0], x[0::2, 1], x[1::2, 0], x[1::2, 1]])
    for v in (cols, cols ** 2):
        r = np.corrcoef(v)
        assert np.all(np.abs(r - np.eye(4)) < 0.015), r
    test_fast_normal_noise_d2_pairs.seen = getattr(test_fast_normal_noise_d2_pairs, "seen", {})
    test_fast_normal_noise_d2_pairs.seen[persistent] = x
    if len(test_fast_normal_noise_d2_pairs.seen) == 2:
        assert np.array_equal(test_fast_normal_noise_d2_pairs.seen[False], test_fast_normal_noise_d2_pairs.seen[True])


def test_mvt_normal_init_switch(ctx, orc):
    """mvt_normal_init = 1 keeps round 1's Normal start; the default draws x_0 with chi factors."""
    rng = np.random.default_rng(3)
    d, N, T = 2, 512, 3
    md = _model(d)
    Y = rng.standard_normal((d, T))
    xi0 = rng.standard_normal((N, d))
    out = {}
    for flag in (False, True):
        pf = ctx.filter(N=N, Y=Y, resampler="systematic", keep_history=True, distribution="mvt", df=4.0, seed=8,
                        mvt_normal_init=flag, **md)
        out[flag] = pf.run(xi0=torch_dev(xi0.T)).history()["x"][0]
        pf.close()
    base = xi0 @ _eig_factor(md["C0"]).T
    assert np.array_equal(out[True], base)
    ratio = out[False] / base                       # the device-drawn chi factors: positive, not all one
    assert np.all(ratio > 0) and np.std(ratio) > 0.1
    # their law: nu / chi^2 has mean nu / (nu - 2) = 2 at nu = 4 (heavy tailed: loose bound)
    assert 1.5 < np.mean(ratio ** 2) < 3.5


def test_eigen_factor_through_the_abi(ctx):
    """a8 (ref: src/linear_algebra.cpp:10-23): the factor the library builds from C0 and W must satisfy
    Q Q^T = Sigma for NON-diagonal covariances.  Q never crosses the ABI, so it is read off the
    particles: unit-vector draws make x_0[k] - m0 = Q_c0[:, k] and, with G = 0, x_1[k] = Q_w[:, k]."""
    rng = np.random.default_rng(2024)
    for d in (2, 3, 5, 8):
        C0, W = spd(rng, d), spd(rng, d) * 0.7
        m0 = rng.standard_normal(d)
        N, T = 64, 2
        xi = np.zeros((N, d))
        xi[:d] = np.eye(d)
        pf = ctx.filter(N=N, Y=np.zeros((d, T)), m0=m0, C0=C0, F=np.eye(d), G=np.zeros((d, d)), V=np.eye(d), W=W,
                        resampler="systematic", keep_history=True)
        h = pf.run(xi0=torch_dev(xi.T), xi=torch_dev(xi.T[None]), u0=np.array([0.5])).history()
        pf.close()
        Qc0 = (h["x"][0, :d] - m0).T
        # uniform weights at t = 0 and u0 = 0.5: the ancestors of step 1 are the identity
        assert np.array_equal(h["a"][1], np.arange(N))
        Qw = h["x"][1, :d].T
        assert np.allclose(Qc0 @ Qc0.T, C0, rtol=0, atol=1e-13 * np.abs(C0).max() * d)
        assert np.allclose(Qw @ Qw.T, W, rtol=0, atol=1e-13 * np.abs(W).max() * d)
        # the eigen form, not just any factor: columns are orthogonal (Q^T Q = Lambda)
        G_ = Qc0.T @ Qc0
        assert np.allclose(G_ - np.diag(np.diag(G_)), 0.0, atol=1e-12)
        assert np.allclose(np.sort(np.diag(G_)), np.linalg.eigvalsh(C0), rtol=1e-12)


@pytest.mark.parametrize("resampler", ["systematic", "metropolis"])
def test_filter_device_rng_matches_oracle_mirror(ctx, orc, resampler):
    rng = np.random.default_rng(77)
    d, N, T = 2, 2048, 8
    md = _model(d)
    Y = rng.standard_normal((d, T))
    pf = ctx.filter(N=N, Y=Y, resampler=resampler, seed=4242, keep_history=True, reproducible_rng=True, **md)
    h = pf.run().history()
    pf.close()
    ref = orc.filter_det("mvn", resampler, Y, md["m0"], _eig_factor(md["C0"]), md["F"], md["G"], md["V"],
                         _eig_factor(md["W"]), N, seed=4242)
    assert np.array_equal(h["a"], ref["a"])
    assert np.array_equal(h["x"], ref["x"])


@pytest.mark.parametrize("d,diag,N", [(2, True, 20000), (2, False, 8192), (4, True, 12346), (4, False, 4100),
                                      (8, True, 50000), (8, False, 9001), (2, True, 300000), (2, True, 1000000)])
def test_persistent_kernel_bit_exact_vs_oracle(ctx, orc, d, diag, N):
    """cusmc_filter_run as ONE cooperative kernel (pf_persist.cu; the C4 path) against the ORACLE: device-
    drawn noise on one side, the oracle's Philox mirror on the other, no history.  The weight image is
    defined per tile and the persistent kernel spreads the cloud evenly over its resident blocks, so
    the oracle is told the run's tile size: final states, log-weights and ancestors bit for bit, ESS and
    log-likelihood of every step.  The per-step path of the same filter (tile 2048) is checked too."""
    rng = np.random.default_rng(31 * d + N)
    T = 9 if N >= 300000 else 14
    I = np.eye(d)
    if diag:
        md = dict(m0=np.zeros(d), C0=I, F=I, G=0.9 * I, V=0.5 * I, W=0.3 * I)
    else:
        A = rng.standard_normal((d, d)) * 0.2
        # dense F, G, V (the general kernel path); C0 and W stay scalar so that the library's Jacobi factor
        # and numpy's eigh factor are the same matrix (eigenvector signs are not unique otherwise)
        md = dict(m0=rng.standard_normal(d), C0=1.3 * I, F=I + A, G=0.8 * I + A.T, V=spd(rng, d), W=0.4 * I)
    Y = rng.standard_normal((d, T))
    tiles = []
    for persistent in (True, False):
        pf = ctx.filter(N=N, Y=Y, resampler="systematic", seed=99, summary=False, persistent=persistent,
                        reproducible_rng=True, **md)
        tile = pf.tile_size
        l0 = ctx.launch_count
        pf.run()
        x, w, a = pf.state()
        launches = ctx.launch_count - l0
        s = pf.summary()
        pf.close()
        # the persistent run really took the one-kernel path: init_slots + ONE cooperative kernel
        assert launches == (2 if persistent else 1 + 2 * T)
        assert (tile != 2048 or N > 500000) if persistent else tile == 2048
        tiles.append(tile)
        ref = orc.filter_det("mvn", "systematic", Y, md["m0"], _eig_factor(md["C0"]), md["F"], md["G"], md["V"],
                             _eig_factor(md["W"]), N, seed=99, tile=tile)
        assert np.array_equal(a, ref["a"][-1])
        assert np.array_equal(x.T, ref["x"][-1])
        assert np.array_equal(w, ref["w"][-1])
        assert np.allclose(s["ess"], ref["ess"], rtol=1e-12)
        assert np.allclose(s["loglik"], ref["loglik"], rtol=1e-12, atol=1e-12)
        assert len(np.unique(a)) > 10 and np.all(np.diff(a.astype(np.int64)) >= 0)
    assert tiles[0] % 32 == 0 and tiles[0] <= 2048


def _library_factors(ctx, C0, W):
    """The eigen factors the library builds from C0 and W, bit for bit, read off the particles as in
    test_eigen_factor_through_the_abi (m0 = 0, G = 0 and unit-vector draws: x_0[k] = Q_c0[:, k], x_1[k] = Q_w[:, k])."""
    d = C0.shape[0]
    N = 64
    xi = np.zeros((N, d))
    xi[:d] = np.eye(d)
    pf = ctx.filter(N=N, Y=np.zeros((d, 2)), m0=np.zeros(d), C0=C0, F=np.eye(d), G=np.zeros((d, d)), V=np.eye(d), W=W,
                    resampler="systematic", keep_history=True)
    h = pf.run(xi0=torch_dev(xi.T), xi=torch_dev(xi.T[None]), u0=np.array([0.5])).history()
    pf.close()
    assert np.array_equal(h["a"][1], np.arange(N))
    return h["x"][0, :d].T.copy(), h["x"][1, :d].T.copy()


@pytest.mark.parametrize("kind,nu", [("mvn", 0.0), ("mvt", 5.0)])
@pytest.mark.parametrize("d", [3, 4, 8, 16])
def test_filter_dense_model_bit_exact_vs_oracle(ctx, orc, d, kind, nu):
    """The GENERAL kernel path (every operator dense: F, G, V, W, C0 full; d = 3 is the padded, non-EXACT
    instantiation) of the per-step fused kernel -- operators staged in shared memory, parent gathers
    issued one round ahead -- against the oracle on the same pre-drawn normals / uniforms / chi factors:
    ancestors, states and log-weights bit for bit at every step.  N spans several tiles, the last ragged."""
    rng = np.random.default_rng(900 + d)
    N, T = 2048 * 3 + 517, 7
    A = rng.standard_normal((d, d)) * (0.4 / np.sqrt(d))
    C0, W, V = spd(rng, d), spd(rng, d) * 0.6, spd(rng, d)
    md = dict(m0=rng.standard_normal(d), C0=C0, F=np.eye(d) + A, G=0.8 * np.eye(d) + A.T, V=V, W=W)
    Qc0, Qw = _library_factors(ctx, C0, W)
    assert np.abs(Qw - np.diag(np.diag(Qw))).max() > 1e-3            # really dense
    Y = rng.standard_normal((d, T))
    xi0, xi = rng.standard_normal((N, d)), rng.standard_normal((T - 1, N, d))
    u0 = rng.random(T - 1)
    chi0 = chi = None
    kw = {}
    if kind == "mvt":
        chi0 = np.sqrt(nu / rng.chisquare(nu, (N, d)))
        chi = np.sqrt(nu / rng.chisquare(nu, (T - 1, N, d)))
        kw = dict(chi0=torch_dev(chi0.T), chi=torch_dev(chi.transpose(0, 2, 1)))
    pf = ctx.filter(N=N, Y=Y, resampler="systematic", keep_history=True, distribution=kind, df=nu, **md)
    pf.run(xi0=torch_dev(xi0.T), xi=torch_dev(xi.transpose(0, 2, 1)), u0=u0, **kw)
    h = pf.history()
    pf.close()
    ref = orc.filter_det(kind, "systematic", Y, md["m0"], Qc0, md["F"], md["G"], md["V"], Qw, N, nu=nu,
                         xi0=xi0, xi=xi, u0=u0, chi0=chi0, chi=chi)
    assert np.array_equal(h["a"], ref["a"])
    assert np.array_equal(h["x"], ref["x"])
    if kind == "mvn":
        assert np.array_equal(h["lw"], ref["w"])
    else:
        assert np.allclose(h["lw"], ref["w"], rtol=1e-12, atol=1e-12)
    assert len(np.unique(h["a"][-1])) > 100


@pytest.mark.parametrize("d", [8, 16])
def test_filter_dense_model_device_rng_vs_oracle_mirror(ctx, orc, d):
    """Same path with device-drawn (reproducible) noise and no history -- the configuration the dense
    bench line runs, minus the throughput generator: final state against the oracle's Philox mirror."""
    rng = np.random.default_rng(950 + d)
    N, T = 2048 * 5 + 33, 6
    A = rng.standard_normal((d, d)) * (0.4 / np.sqrt(d))
    C0, W = spd(rng, d), spd(rng, d) * 0.6
    md = dict(m0=rng.standard_normal(d), C0=C0, F=np.eye(d) + A, G=0.8 * np.eye(d) + A.T, V=spd(rng, d), W=W)
    Qc0, Qw = _library_factors(ctx, C0, W)
    Y = rng.standard_normal((d, T))
    pf = ctx.filter(N=N, Y=Y, resampler="systematic", seed=31, summary=False, persistent=False,
                    reproducible_rng=True, **md)
    pf.run()
    x, w, a = pf.state()
    pf.close()
    ref = orc.filter_det("mvn", "systematic", Y, md["m0"], Qc0, md["F"], md["G"], md["V"], Qw, N, seed=31)
    assert np.array_equal(a, ref["a"][-1])
    assert np.array_equal(x.T, ref["x"][-1])
    assert np.array_equal(w, ref["w"][-1])


@pytest.mark.parametrize("d,tile,N", [(8, 1920, 2 * 1920 + 777), (8, 96, 5000), (2, 1696, 9000), (4, 2048, 7000)])
def test_filter_tile_size_bit_exact_vs_oracle(ctx, orc, d, tile, N):
    """cfg.tile_size: the per-step path with a caller-chosen tile (e.g. to reproduce a persistent run's evenly
    spread tile on the per-step path).  The weight image is defined per tile, so the oracle is told the tile:
    final states, log-weights and ancestors bit for bit, for tiles that are not a multiple of the block size,
    with a ragged last tile."""
    rng = np.random.default_rng(17 * d + tile)
    T = 9
    md = _model(d)
    Y = rng.standard_normal((d, T))
    pf = ctx.filter(N=N, Y=Y, resampler="systematic", seed=123, summary=False, persistent=False, reproducible_rng=True,
                    tile_size=tile, **md)
    assert pf.tile_size == tile
    pf.run()
    x, w, a = pf.state()
    s = pf.summary()
    pf.close()
    ref = orc.filter_det("mvn", "systematic", Y, md["m0"], _eig_factor(md["C0"]), md["F"], md["G"], md["V"],
                         _eig_factor(md["W"]), N, seed=123, tile=tile)
    assert np.array_equal(a, ref["a"][-1])
    assert np.array_equal(x.T, ref["x"][-1])
    assert np.array_equal(w, ref["w"][-1])
    assert np.allclose(s["ess"], ref["ess"], rtol=1e-12)
    assert np.allclose(s["loglik"], ref["loglik"], rtol=1e-12, atol=1e-12)


def test_filter_tile_size_default_and_validation(ctx):
    """The tile is 2048 unless the caller fixes it; bad values are refused."""
    import cusmc_b200
    d = 8
    md = _model(d)
    Y = np.random.default_rng(3).standard_normal((d, 4))
    for N, resampler in ((100000, "systematic"), (8 << 20, "systematic"), (50000, "multinomial")):
        pf = ctx.filter(N=N, Y=Y, resampler=resampler, summary=False, persistent=False, **md)
        assert pf.tile_size == 2048
        pf.close()
    with pytest.raises(cusmc_b200.CusmcError):
        ctx.filter(N=5000, Y=Y, resampler="systematic", tile_size=100, **md)          # not a multiple of 32
    with pytest.raises(cusmc_b200.CusmcError):
        ctx.filter(N=5000, Y=Y, resampler="multinomial", tile_size=1024, **md)        # systematic only


@pytest.mark.parametrize("d,thr", [(2, 0.5), (8, 0.05)])
def test_adaptive_resampling_bit_exact_vs_oracle(ctx, orc, d, thr):
    """ess_threshold: resample only when ESS < threshold N, otherwise keep a_i = i and accumulate the
    log-weights.  The decision comes from the integer weight sums, so the oracle takes the same one:
    ancestors, states and (cumulative) log-weights agree bit for bit, and both kinds of step occur."""
    rng = np.random.default_rng(77 + d)
    N, T = 4000, 30
    md = _model(d)
    Y = rng.standard_normal((d, T)) * 0.5
    pf = ctx.filter(N=N, Y=Y, resampler="systematic", seed=5, keep_history=True, ess_threshold=thr,
                    reproducible_rng=True, **md)
    h, s, res = pf.run().history(), pf.summary(), pf.resampled()
    pf.close()
    ref = orc.filter_det("mvn", "systematic", Y, md["m0"], _eig_factor(md["C0"]), md["F"], md["G"], md["V"],
                         _eig_factor(md["W"]), N, seed=5, ess_threshold=thr)
    assert np.array_equal(res, ref["resampled"])
    assert 0 < res[1:].sum() < T - 1                              # both kinds of step happened
    assert np.array_equal(h["a"], ref["a"])
    assert np.array_equal(h["x"], ref["x"])
    assert np.array_equal(h["lw"], ref["w"])
    assert np.allclose(s["ess"], ref["ess"], rtol=1e-12)
    kept = np.where(res[1:] == 0)[0] + 1
    assert all(np.array_equal(h["a"][t], np.arange(N)) for t in kept)
    # a resampling step is triggered exactly by the ESS of the step before
    assert np.array_equal(res[1:], (s["ess"][:-1] < thr * N).astype(np.int32))
    with pytest.raises(Exception):
        ctx.filter(N=N, Y=Y, resampler="multinomial", ess_threshold=thr, **md)


def test_rejection_resampler(ctx, orc):
    """The rejection resampler (N4): bit-exact against the oracle's mirror of the same counter-based
    draws, unbiased offspring counts (its point over the B-step Metropolis rule), and a whole filter run
    bit for bit."""
    import torch
    rng = np.random.default_rng(606)
    N = 50000
    w = rng.random(N) ** 4
    w[rng.random(N) < 0.05] = 0.0
    wd = torch_dev(w)
    wmax = torch_dev(np.array([w.max()]))
    a = torch.empty(N, dtype=torch.int32, device="cuda")
    ctx.rejection_resample_dev(a, wd, wmax, seed=11, step=3)
    ctx.synchronize()
    got = a.cpu().numpy().view(np.uint32)
    assert np.array_equal(got, orc.resample_rejection(w, w.max(), 11, 3))
    # offspring counts: E[#children of j] = N w_j / sum(w), checked on coarse bins (exact in law)
    bins = np.add.reduceat(np.bincount(got, minlength=N), np.arange(0, N, 500))
    want = N * np.add.reduceat(w, np.arange(0, N, 500)) / w.sum()
    assert np.all(np.abs(bins - want) < 6 * np.sqrt(want + 1) + 3)
    assert np.all(w[got] > 0)                              # nobody descends from a zero-weight particle
    # inside the filter (reference-mode densities, device-drawn everything)
    d, Nf, T = 2, 4000, 8
    md = _model(d)
    Y = rng.standard_normal((d, T))
    pf = ctx.filter(N=Nf, Y=Y, resampler="rejection", seed=21, keep_history=True, reproducible_rng=True, **md)
    h = pf.run().history()
    lin, nuq = pf.lineage()
    pf.close()
    ref = orc.filter_det("mvn", "rejection", Y, md["m0"], _eig_factor(md["C0"]), md["F"], md["G"], md["V"],
                         _eig_factor(md["W"]), Nf, seed=21)
    assert np.array_equal(h["a"], ref["a"])
    assert np.array_equal(h["x"], ref["x"])
    assert relerr(h["w"], ref["w"]) < 1e-13
    # the ancestor tree as paths: lineage[t - 1] = a_t[lineage[t]], and the genealogy coalesces backwards
    want_lin = np.empty((T, Nf), dtype=np.uint32)
    want_lin[T - 1] = np.arange(Nf)
    for t in range(T - 1, 0, -1):
        want_lin[t - 1] = h["a"][t][want_lin[t]]
    assert np.array_equal(lin, want_lin)
    assert np.array_equal(nuq, [len(np.unique(r)) for r in want_lin])
    assert np.all(np.diff(nuq) >= 0) and nuq[-1] == Nf and nuq[0] < Nf


def test_metropolis_c2_resampler(ctx, orc):
    """Metropolis-C2 (N4): the reference's accept rule with a warp's proposals confined to one 32-particle
    segment per iteration.  Bit-exact against the oracle's mirror of the counters (ragged N: the last
    segment is short), proposals uniform over 0 .. N-1, and a whole filter run bit for bit."""
    import torch
    rng = np.random.default_rng(707)
    for N, B in ((5000, 10), (4099, 7), (33, 4)):
        w = rng.random(N) ** 4
        w[rng.random(N) < 0.05] = 0.0
        u, j = orc.rng_metropolis(13, 2, N, B, c2=True)
        # the structure: the 32 particles of a group share their segment at every iteration
        seg = j // 32
        for g0 in range(0, N, 32):
            assert np.all(seg[g0:g0 + 32] == seg[g0]), (N, g0)
        assert j.max() < N
        a = torch.empty(N, dtype=torch.int32, device="cuda")
        ctx.metropolis_c2_dev(a, torch_dev(w), B, seed=13, step=2)
        ctx.synchronize()
        assert np.array_equal(a.cpu().numpy().view(np.uint32), orc.metropolis_hastings(w, u, j)), (N, B)
    # in law: ancestors follow the weights as well as the plain B-step rule's do (same bias allowance)
    N, B = 200000, 30
    w = rng.random(N)
    a = torch.empty(N, dtype=torch.int32, device="cuda")
    ctx.metropolis_c2_dev(a, torch_dev(w), B, seed=5, step=1)
    ctx.synchronize()
    got = a.cpu().numpy().view(np.uint32)
    bins = np.add.reduceat(np.bincount(got, minlength=N), np.arange(0, N, 2000))
    want = N * np.add.reduceat(w, np.arange(0, N, 2000)) / w.sum()
    # (the 32 particles of a group propose from common segments: bin counts up to 32 x Poisson)
    assert np.all(np.abs(bins - want) < 6 * np.sqrt(32 * want) + 0.02 * want)
    # inside the filter (reference-mode densities)
    d, Nf, T = 2, 4000, 8
    md = _model(d)
    Y = rng.standard_normal((d, T))
    pf = ctx.filter(N=Nf, Y=Y, resampler="metropolis_c2", seed=22, keep_history=True, reproducible_rng=True, **md)
    h = pf.run().history()
    pf.close()
    ref = orc.filter_det("mvn", "metropolis_c2", Y, md["m0"], _eig_factor(md["C0"]), md["F"], md["G"], md["V"],
                         _eig_factor(md["W"]), Nf, seed=22)
    assert np.array_equal(h["a"], ref["a"])
    assert np.array_equal(h["x"], ref["x"])
    assert relerr(h["w"], ref["w"]) < 1e-13


def test_lineage_of_a_systematic_run(ctx):
    d, N, T = 2, 20000, 12
    rng = np.random.default_rng(5)
    md = _model(d)
    pf = ctx.filter(N=N, Y=rng.standard_normal((d, T)), resampler="systematic", seed=4, keep_history=True, **md)
    h = pf.run().history()
    lin, nuq = pf.lineage()
    pf.close()
    k = np.arange(N)
    for t in range(T - 1, 0, -1):
        assert np.array_equal(lin[t], k)
        k = h["a"][t][k]
    assert np.array_equal(lin[0], k)
    assert nuq[-1] == N and np.all(np.diff(nuq) >= 0) and nuq[0] < N // 2
    traj = h["x"][np.arange(T)[:, None], lin]             # (T, N, d): the path of every final particle
    assert traj.shape == (T, N, d) and np.array_equal(traj[-1], h["x"][-1])


def test_filter_c5_size_properties(ctx):
    """BASELINE configs[4] at its per-GPU size (d = 8, 8 Mi particles), checked through size-independent
    properties of systematic resampling on the library's own fixed-point weights: ancestors are sorted,
    every parent's offspring count is floor or ceil of N w_j (never off by more), zero-weight parents
    have no children, the normalised weights sum to one, ESS agrees with the weights."""
    d, N, T = 8, 8 << 20, 3
    I = np.eye(d)
    Y = np.random.default_rng(5000).standard_normal((d, T))
    pf = ctx.filter(N=N, Y=Y, m0=np.zeros(d), C0=I, F=I, G=0.9 * I, V=I, W=I, resampler="systematic", seed=2,
                    keep_history=True, summary=False)
    pf.run()
    h, s = pf.history(), pf.summary()
    pf.close()
    for t in (1, 2):
        a = h["a"][t].astype(np.int64)
        assert np.all(np.diff(a) >= 0) and a[0] >= 0 and a[-1] < N
        w = h["w"][t - 1]                                   # normalised weights of the parents
        assert abs(w.sum() - 1.0) < 1e-9
        counts = np.bincount(a, minlength=N)
        expect = N * w
        assert np.all(counts >= np.floor(expect - 1e-6)) and np.all(counts <= np.ceil(expect + 1e-6))
        assert np.all(counts[w == 0.0] == 0)
        assert np.isclose(s["ess"][t - 1], 1.0 / np.sum(w * w), rtol=1e-6)
    assert np.all(np.isfinite(h["x"][2]))


def kalman_means(Y, m0, C0, F, G, V, W):
    m, P = m0.copy(), C0.copy()
    out = [m.copy()]
    for t in range(1, Y.shape[1]):
        m, P = G @ m, G @ P @ G.T + W
        S = F @ P @ F.T + V
        K = P @ F.T @ np.linalg.inv(S)
        m = m + K @ (Y[:, t] - F @ m)
        P = P - K @ F @ P
        out.append(m.copy())
    return np.array(out), P


@pytest.mark.parametrize("resampler", ["systematic", "multinomial", "metropolis"])
def test_filter_posterior_moments_vs_kalman(ctx, resampler):
    """Linear-Gaussian model on the reference's data series: the exact filtering mean is the Kalman
    mean.  Monte Carlo tolerance: 6 sigma / sqrt(ESS) per component (ESS ~ N/3 here), plus the
    O(1/B)-type bias of the Metropolis resampler (B = 30)."""
    Y = np.loadtxt(os.path.join(HERE, "golden", "y_t.csv"), delimiter=",", skiprows=1).T[:, :60]
    I = np.eye(2)
    md = dict(m0=np.zeros(2), C0=I, F=I, G=I, V=0.1 * I, W=0.1 * I)
    N = 200000
    pf = ctx.filter(N=N, Y=Y, resampler=resampler, B=30, seed=9, **md)
    s = pf.run().summary()
    pf.close()
    km, P = kalman_means(Y, **md)
    sd = math.sqrt(P[0, 0])
    # the first observations sit ~2 prior standard deviations out: ESS is ~1% of N there, so the
    # tolerance is scaled by the ESS the filter itself reports (N/4 floor for the Metropolis mode,
    # whose B-step chains are biased while the weights are that uneven: compare from t = 5 on)
    if resampler == "metropolis":
        assert np.max(np.abs(s["mean"][5:] - km[5:])) < 6 * sd / math.sqrt(N / 4) + 0.02
    else:
        # multinomial resampling adds its own O(1/N) variance every step on top of the weight
        # degeneracy the ESS measures: twice the allowance
        tol = (12 if resampler == "multinomial" else 6) * sd / np.sqrt(s["ess"][1:, None])
        assert np.all(np.abs(s["mean"][1:] - km[1:]) < tol)
        assert np.all(s["ess"][5:] > N / 10) and np.all(s["ess"] <= N * (1 + 1e-9))


def test_filter_posterior_moments_vs_kalman_correlated(ctx):
    """The same check on a model whose W, C0, V are full and F, G non-diagonal: the dense kernel path
    and the eigen factors of non-diagonal covariances (a wrong rotation of Q_w would bias the mean)."""
    rng = np.random.default_rng(12)
    d, T, N = 3, 40, 400000
    A = rng.standard_normal((d, d)) * 0.15
    md = dict(m0=rng.standard_normal(d) * 0.3, C0=spd(rng, d), F=np.eye(d) + A, G=0.6 * np.eye(d) + A.T,   # spectral radius 0.97
              V=spd(rng, d) * 0.4, W=spd(rng, d) * 0.3)
    # data simulated from the model itself, so the filter stays in its typical regime
    x = md["m0"] + np.linalg.cholesky(md["C0"]) @ rng.standard_normal(d)
    Y = np.zeros((d, T))
    for t in range(1, T):
        x = md["G"] @ x + np.linalg.cholesky(md["W"]) @ rng.standard_normal(d)
        Y[:, t] = md["F"] @ x + np.linalg.cholesky(md["V"]) @ rng.standard_normal(d)
    pf = ctx.filter(N=N, Y=Y, resampler="systematic", seed=21, **md)
    s = pf.run().summary()
    pf.close()
    km, P = kalman_means(Y, **md)
    sd = np.sqrt(np.diag(P))
    tol = 6 * sd[None, :] / np.sqrt(s["ess"][1:, None])
    assert np.all(np.abs(s["mean"][1:] - km[1:]) < tol)
    assert np.all(s["ess"][1:] > N / 500)


def test_cusmc_run_streams_the_history(ctx):
    """cusmc_run (the entry the R glue calls) through ctypes: its streamed history -- several ring
    chunks -- equals the device-resident history of a filter object run with the same seed."""
    import cusmc_b200
    Y = np.loadtxt(os.path.join(HERE, "golden", "y_t.csv"), delimiter=",", skiprows=1).T
    I = np.eye(2)
    md = dict(m0=np.zeros(2), C0=I, F=I, G=I, V=0.1 * I, W=0.1 * I)
    N, T = 20000, 61                      # a row is 560 KB: chunks of 14 steps, 5 chunks
    for resampler in ("metropolis", "systematic"):
        out = cusmc_b200.run(N, 2, T, Y, md["m0"], I, I, I, md["V"], md["W"], 0.0, resampler, "mvn", seed=77,
                             ancestors=True)
        pf = cusmc_b200.ParticleFilter(cusmc_b200.default_context(), N, Y[:, :T], resampler=resampler, seed=77,
                                       keep_history=True, summary=False, persistent=False, **md)
        h = pf.run().history()
        pf.close()
        assert out["posterior_x"].shape == (T, N, 2) and out["weights"].shape == (T, N)
        assert np.array_equal(out["posterior_x"], h["x"])
        assert np.array_equal(out["ancestors"], h["a"])
        assert np.array_equal(out["weights"], h["w"])
        if resampler == "systematic":
            assert np.allclose(out["weights"].sum(axis=1), 1.0, rtol=1e-9)
    # the plain entry (no ancestors) gives the same arrays
    out2 = cusmc_b200.run(N, 2, T, Y, md["m0"], I, I, I, md["V"], md["W"], 0.0, "systematic", "mvn", seed=77)
    assert np.array_equal(out2["posterior_x"], out["posterior_x"]) and np.array_equal(out2["weights"], out["weights"])


def test_degenerate_step_is_reported(ctx):
    """A step whose weights have no mass (NaN observation): identity ancestors, and every getter that
    hands results out returns CUSMC_ERR_DEGENERATE instead of stale ancestors and CUSMC_OK."""
    import cusmc_b200
    from cusmc_b200 import _lib
    I = np.eye(2)
    md = dict(m0=np.zeros(2), C0=I, F=I, G=I, V=0.1 * I, W=0.1 * I)
    Y = np.zeros((2, 6))
    Y[:, 2] = np.nan                      # every log-weight of step 2 is NaN -> step 3 cannot resample
    for resampler in ("systematic", "multinomial"):
        for persistent in ((True, False) if resampler == "systematic" else (False,)):
            pf = ctx.filter(N=5000, Y=Y, resampler=resampler, seed=3, keep_history=not persistent,
                            persistent=persistent, **md)
            pf.run()
            with pytest.raises(cusmc_b200.CusmcError) as e:
                pf.summary()
            assert e.value.code == _lib.ERR_DEGENERATE and "step 3" in str(e.value)
            with pytest.raises(cusmc_b200.CusmcError):
                pf.status()
            pf.close()
    with pytest.raises(cusmc_b200.CusmcError) as e:
        cusmc_b200.run(1000, 2, 6, Y, md["m0"], I, I, I, md["V"], md["W"], 0.0, "systematic", "mvn")
    assert e.value.code == _lib.ERR_DEGENERATE
    # a healthy run reports nothing
    pf = ctx.filter(N=5000, Y=np.zeros((2, 6)), resampler="systematic", seed=3, **md)
    pf.run().status()
    pf.close()


def test_run_r_api(ctx):
    import cusmc_b200
    Y = np.loadtxt(os.path.join(HERE, "golden", "y_t.csv"), delimiter=",", skiprows=1).T
    I = np.eye(2)
    out = cusmc_b200.run(1000, 2, 25, Y, np.zeros(2), I, I, I, 0.1 * I, 0.1 * I, 0.0, "metropolis", "mvn")
    assert out["weights"].shape == (25, 1000) and out["posterior_x"].shape == (25, 1000, 2)
    assert np.allclose(out["weights"][0], 1.0 / 1000)           # w_0 = 1/N (src/mcmc.cpp:85)
    assert np.all(out["weights"][1:] >= 0) and np.all(np.isfinite(out["posterior_x"]))
    out = cusmc_b200.run(500, 2, 10, Y, np.zeros(2), I, I, I, 0.1 * I, 0.1 * I, 4.0, "metropolis", "mvt")
    assert np.all(np.isfinite(out["posterior_x"]))
    with pytest.raises(ValueError):
        cusmc_b200.run(10, 2, 5, Y, np.zeros(2), I, I, I, I, I, 0.0, "gibbs", "mvn")
    d1 = cusmc_b200.MVN(np.zeros(2), I, seed=1)
    d2 = cusmc_b200.MVT(np.zeros(3), np.eye(3), 3.0, seed=1)
    assert d1.shape == (2,) and d2.shape == (3,) and np.all(np.isfinite(d1)) and np.all(np.isfinite(d2))


# ------------------------------------------------------------------------------------------------
# independent MH chains
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("kind,nu", [("mvn", 0.0), ("mvt", 5.0)])
@pytest.mark.parametrize("d,shared", [(2, False), (8, True), (32, False), (17, False)])
def test_mh_chains_bit_exact(ctx, orc, kind, nu, d, shared):
    import torch
    rng = np.random.default_rng(900 + d)
    Cn, steps, step = 257, 40, 0.3
    if shared:
        L, mu = np.linalg.cholesky(spd(rng, d)), rng.standard_normal(d)
        Ldev, mudev = torch_dev(L.T), torch_dev(mu)              # column-major
    else:
        L = np.stack([np.linalg.cholesky(spd(rng, d)) for _ in range(Cn)])
        mu = rng.standard_normal((Cn, d))
        Ldev, mudev = torch_dev(L.transpose(0, 2, 1)), torch_dev(mu)
    x0 = rng.standard_normal((Cn, d))
    z = rng.standard_normal((Cn, steps, d))
    e = -np.log(rng.random((Cn, steps)))
    thr = e if kind == "mvn" else np.exp(2 * e / (nu + d))
    want_x, want_n, want_bits = orc.mh_chains(kind, mu, L, x0, z, thr, step, nu=nu, shared=shared)
    x = torch_dev(x0)
    nacc = torch.zeros(Cn, dtype=torch.int32, device="cuda")
    bits = torch.zeros((Cn, steps), dtype=torch.uint8, device="cuda")
    ctx.mh_chains_dev(kind, mudev, Ldev, x, steps, step, nu=nu, shared=shared, z=torch_dev(z), thr=torch_dev(thr),
                      n_accept=nacc, accept_bits=bits)
    ctx.synchronize()
    assert np.array_equal(bits.cpu().numpy(), want_bits)          # accept / reject decisions
    assert np.array_equal(nacc.cpu().numpy().astype(np.uint32), want_n)
    assert np.array_equal(x.cpu().numpy(), want_x)
    assert 0.05 < want_bits.mean() < 0.99


@pytest.mark.parametrize("kind,nu,d", [("mvn", 0.0, 8), ("mvt", 5.0, 32), ("mvt", 3.0, 5)])
def test_mh_chains_device_rng_matches_oracle_mirror(ctx, orc, kind, nu, d):
    """Device-drawn chains: the kernel batches its Philox blocks (thresholds per 32 steps, normals per
    4 steps and quad of lanes) but must consume exactly the (seed, chain, step, component) -> draw
    mapping of cusmc_philox.h.  The oracle regenerates those draws on the host and runs the chains
    with them: accept bits and final states agree bit for bit."""
    import torch
    rng = np.random.default_rng(4100 + d)
    Cn, steps, step, seed = 48, 70, 0.3, 991
    L = np.stack([np.linalg.cholesky(spd(rng, d)) for _ in range(Cn)])
    mu = rng.standard_normal((Cn, d))
    x0 = rng.standard_normal((Cn, d))
    z = np.stack([orc.rng_fill_normals(seed, 4, s, 0, Cn, d) for s in range(steps)], axis=1)   # CHAIN_Z
    thr = np.empty((Cn, steps))
    for c in range(Cn):
        for s in range(steps):
            e = -float(orc.det_log(orc.rng_u01(seed, 5, s, c) + 2.0 ** -53)[0])             # CHAIN_U, (0, 1]
            thr[c, s] = e if kind == "mvn" else float(orc.det_exp((e + e) / (nu + d))[0])
    want_x, want_n, want_bits = orc.mh_chains(kind, mu, L, x0, z, thr, step, nu=nu, shared=False)
    x = torch_dev(x0)
    nacc = torch.zeros(Cn, dtype=torch.int32, device="cuda")
    bits = torch.zeros((Cn, steps), dtype=torch.uint8, device="cuda")
    ctx.mh_chains_dev(kind, torch_dev(mu), torch_dev(L.transpose(0, 2, 1)), x, steps, step, nu=nu, seed=seed,
                      n_accept=nacc, accept_bits=bits)
    ctx.synchronize()
    assert np.array_equal(bits.cpu().numpy(), want_bits)
    assert np.array_equal(x.cpu().numpy(), want_x)
    assert 0.05 < want_bits.mean() < 0.99


@pytest.mark.parametrize("kind,nu,d,shared,Cn,steps", [("mvn", 0.0, 2, False, 257, 40), ("mvt", 5.0, 8, True, 257, 40),
                                                       ("mvn", 0.0, 17, False, 100, 40),
                                                       ("mvn", 0.0, 25, True, 101, 30),         # two-rows-per-lane kernel,
                                                       ("mvt", 5.0, 16, False, 65, 30),         # odd chain counts
                                                       ("mvt", 5.0, 32, False, 1024, 100)])     # C3's shape, 1024 x 100
def test_mh_chains_general_proposal_bit_exact(ctx, orc, kind, nu, d, shared, Cn, steps):
    """The random walk whose proposal does not use the target's factor (x' = x + s z, optionally scaled per
    component): every step evaluates the target density in the kernel.  Same pre-drawn normals and
    thresholds on both sides: accept / reject decisions and final states bit for bit."""
    import torch
    rng = np.random.default_rng(1900 + d)
    step = 0.9 / math.sqrt(d)
    if shared:
        L, mu = np.linalg.cholesky(spd(rng, d)), rng.standard_normal(d)
        Ldev, mudev = torch_dev(L.T), torch_dev(mu)              # column-major
    else:
        L = np.stack([np.linalg.cholesky(spd(rng, d)) for _ in range(Cn)])
        mu = rng.standard_normal((Cn, d))
        Ldev, mudev = torch_dev(L.transpose(0, 2, 1)), torch_dev(mu)
    scale = None if d == 32 else 0.5 + rng.random(d)
    x0 = mu + rng.standard_normal((Cn, d))
    z = rng.standard_normal((Cn, steps, d))
    e = -np.log(rng.random((Cn, steps)))
    thr = e if kind == "mvn" else np.exp(2 * e / (nu + d))
    want_x, want_n, want_bits = orc.mh_chains_general(kind, mu, L, x0, z, thr, step, nu=nu, shared=shared, scale=scale)
    x = torch_dev(x0)
    nacc = torch.zeros(Cn, dtype=torch.int32, device="cuda")
    bits = torch.zeros((Cn, steps), dtype=torch.uint8, device="cuda")
    ctx.mh_chains_general_dev(kind, mudev, Ldev, x, steps, step, nu=nu, shared=shared,
                              scale=None if scale is None else torch_dev(scale), z=torch_dev(z), thr=torch_dev(thr),
                              n_accept=nacc, accept_bits=bits)
    ctx.synchronize()
    assert np.array_equal(bits.cpu().numpy(), want_bits)          # accept / reject decisions
    assert np.array_equal(nacc.cpu().numpy().astype(np.uint32), want_n)
    assert np.array_equal(x.cpu().numpy(), want_x)
    assert 0.05 < want_bits.mean() < 0.95


@pytest.mark.parametrize("general", [False, True])
@pytest.mark.parametrize("kind,nu", [("mvn", 0.0), ("mvt", 7.0)])
def test_mh_chains_throughput_noise_law(ctx, general, kind, nu):
    """cusmc_ctx_set_chain_noise(0): proposal normals from Philox4x32-7 + the special-function-unit Box-Muller.
    Not reproducible on a host, so the check is the law: many chains on one target, time-averaged mean and
    variance against the target's (MVT: covariance nu / (nu - 2) Sigma), both kernels; and the switch really
    changes the draws while the default stays the mirrored generator."""
    import torch
    rng = np.random.default_rng(555)
    d, Cn, steps = 4, 4096, 3000
    S = spd(rng, d)
    Ls, mus = np.linalg.cholesky(S), rng.standard_normal(d)
    run = ctx.mh_chains_general_dev if general else ctx.mh_chains_dev
    step = 1.1 if general else 0.9

    def chains(seed, n_steps):
        x = torch_dev(np.tile(mus, (Cn, 1)))
        sx = torch.zeros((Cn, d), dtype=torch.float64, device="cuda")
        sxx = torch.zeros_like(sx)
        nacc = torch.zeros(Cn, dtype=torch.int32, device="cuda")
        run(kind, torch_dev(mus), torch_dev(Ls.T), x, n_steps, step, nu=nu, shared=True, seed=seed, n_accept=nacc,
            sum_x=sx, sum_xx=sxx)
        ctx.synchronize()
        return x.cpu().numpy(), sx.cpu().numpy(), sxx.cpu().numpy(), nacc.cpu().numpy()

    x_rep, _, _, _ = chains(78, 20)
    ctx.set_chain_noise(reproducible=False)
    try:
        x_fast, _, _, _ = chains(78, 20)
        _, sx, sxx, nacc = chains(79, steps)
    finally:
        ctx.set_chain_noise(reproducible=True)
    x_rep2, _, _, _ = chains(78, 20)
    assert np.array_equal(x_rep, x_rep2) and not np.array_equal(x_rep, x_fast)
    mean = sx.mean(0) / steps
    var = sxx.mean(0) / steps - mean ** 2
    target_var = np.diag(S) * (nu / (nu - 2.0) if kind == "mvt" else 1.0)
    assert 0.1 < nacc.mean() / steps < 0.7
    tol = 6 * np.sqrt(target_var * 60 / (Cn * steps))
    assert np.all(np.abs(mean - mus) < tol)
    assert np.all(np.abs(var / target_var - 1.0) < (0.1 if kind == "mvt" else 0.05))


@pytest.mark.parametrize("d", [5, 21, 32])
def test_mh_chains_general_device_rng(ctx, orc, d):
    """Device-drawn general chains: the oracle regenerates the kernel's Philox draws (bit-exact decisions; d > 16
    runs the two-rows-per-lane kernel, whose lanes draw the blocks of two components), and the chains'
    time-averaged moments match the MVN target's."""
    import torch
    rng = np.random.default_rng(4242 + d)
    Cn, steps, seed = 47, 70, 313
    L = np.stack([np.linalg.cholesky(spd(rng, d)) for _ in range(Cn)])
    mu = rng.standard_normal((Cn, d))
    x0 = mu + rng.standard_normal((Cn, d))
    z = np.stack([orc.rng_fill_normals(seed, 4, s, 0, Cn, d) for s in range(steps)], axis=1)   # CHAIN_Z
    thr = np.empty((Cn, steps))
    for c in range(Cn):
        for s in range(steps):
            thr[c, s] = -float(orc.det_log(orc.rng_u01(seed, 5, s, c) + 2.0 ** -53)[0])       # CHAIN_U, (0, 1]
    step = 1.1 / math.sqrt(d)
    want_x, _, want_bits = orc.mh_chains_general("mvn", mu, L, x0, z, thr, step, shared=False)
    x = torch_dev(x0)
    bits = torch.zeros((Cn, steps), dtype=torch.uint8, device="cuda")
    ctx.mh_chains_general_dev("mvn", torch_dev(mu), torch_dev(L.transpose(0, 2, 1)), x, steps, step, seed=seed,
                              accept_bits=bits)
    ctx.synchronize()
    assert np.array_equal(bits.cpu().numpy(), want_bits)
    assert np.array_equal(x.cpu().numpy(), want_x)
    assert 0.05 < want_bits.mean() < 0.95
    if d != 5:
        return
    # the law: many chains on one MVN target, running sums of x and x^2
    d, Cn, steps = 4, 4096, 3000
    S = spd(rng, d)
    Ls, mus = np.linalg.cholesky(S), rng.standard_normal(d)
    x = torch_dev(np.tile(mus, (Cn, 1)))
    sx = torch.zeros((Cn, d), dtype=torch.float64, device="cuda")
    sxx = torch.zeros_like(sx)
    nacc = torch.zeros(Cn, dtype=torch.int32, device="cuda")
    ctx.mh_chains_general_dev("mvn", torch_dev(mus), torch_dev(Ls.T), x, steps, 1.1, shared=True, seed=78, n_accept=nacc,
                              sum_x=sx, sum_xx=sxx)
    ctx.synchronize()
    mean = sx.cpu().numpy().mean(0) / steps
    var = sxx.cpu().numpy().mean(0) / steps - mean ** 2
    acc = nacc.cpu().numpy().mean() / steps
    assert 0.15 < acc < 0.6
    # ~Cn * steps / (integrated autocorrelation ~ 40) effective draws; the start at the mode adds O(1 / steps) bias
    tol = 6 * np.sqrt(np.diag(S) * 40 / (Cn * steps))
    assert np.all(np.abs(mean - mus) < tol)
    assert np.all(np.abs(var / np.diag(S) - 1.0) < 0.05)


def test_mh_chains_device_rng_moments(ctx):
    """Philox-driven chains on an MVN target: time-averaged first and second moments of many chains
    against the target's, 5 sigma Monte Carlo tolerance with the chain autocorrelation folded in."""
    import torch
    rng = np.random.default_rng(31)
    d, Cn, steps = 4, 4096, 2000
    S = spd(rng, d)
    L, mu = np.linalg.cholesky(S), rng.standard_normal(d)
    x = torch_dev(np.tile(mu, (Cn, 1)))
    sx = torch.zeros((Cn, d), dtype=torch.float64, device="cuda")
    sxx = torch.zeros_like(sx)
    nacc = torch.zeros(Cn, dtype=torch.int32, device="cuda")
    ctx.mh_chains_dev("mvn", torch_dev(mu), torch_dev(L.T), x, steps, 1.0, shared=True, seed=77, n_accept=nacc,
                      sum_x=sx, sum_xx=sxx)
    ctx.synchronize()
    mean = sx.cpu().numpy().mean(0) / steps
    var = sxx.cpu().numpy().mean(0) / steps - mean ** 2
    acc = nacc.cpu().numpy().mean() / steps
    assert 0.1 < acc < 0.6
    tol = 5 * np.sqrt(np.diag(S) * 40.0 / (Cn * steps))           # integrated autocorrelation <~ 40
    assert np.all(np.abs(mean - mu) < tol + 0.01)
    assert np.all(np.abs(var / np.diag(S) - 1.0) < 0.05)


def test_layout_helpers(ctx):
    import torch
    x = torch.randn((1000, 7), dtype=torch.float64, device="cuda")
    s = torch.empty((7, 1000), dtype=torch.float64, device="cuda")
    b = torch.empty_like(x)
    ctx.aos_to_soa_dev(x, s)
    ctx.soa_to_aos_dev(s, b)
    ctx.synchronize()
    assert torch.equal(s, x.T.contiguous()) and torch.equal(b, x)

"""CPU tests: pin the oracle against the reference's known answers, the committed golden vectors
and textbook invariants.  No GPU needed."""
import json
import math
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = json.load(open(os.path.join(HERE, "golden", "golden.json")))


def spd(rng, d):
    A = rng.standard_normal((d, d))
    return A @ A.T / d + np.eye(d)


# ---- the reference's own known answers ----------------------------------
def test_reference_known_answers(orc):
    ref = GOLDEN["reference"]
    v = orc.MVNPDF(ref["MVNPDF"]["x"], ref["MVNPDF"]["mu"], np.array(ref["MVNPDF"]["sigma"], float))
    assert round(v, 7) == ref["MVNPDF"]["value"]                  # CuSMC/CuSMC.tex:104
    v = orc.MVTPDF(ref["MVTPDF"]["x"], ref["MVTPDF"]["mu"], np.array(ref["MVTPDF"]["sigma"], float), 3.0)
    assert round(v, 8) == ref["MVTPDF"]["value"]                  # CuSMC/CuSMC.tex:141
    rng = np.random.default_rng(0)
    u, j = rng.random((2, 10)), rng.integers(0, 2, (2, 10), dtype=np.uint32)
    assert orc.metropolis_hastings([0.0, 0.0], u, j).tolist() == [0, 1]   # man/metropolis_hastings.Rd:22-27


def test_independent_golden_vectors(orc):
    worst = 0.0
    for c in GOLDEN["independent"]:
        x, sigma = np.array(c["x"]), np.array(c["sigma"])
        for faithful in (False, True):
            pdf = orc.pdf_batch(c["kind"], x, c["mu"], sigma, c["nu"], faithful=faithful)
            worst = max(worst, np.max(np.abs(pdf - c["pdf"]) / np.abs(c["pdf"])))
        lg = orc.pdf_batch(c["kind"], x, c["mu"], sigma, c["nu"], log=True)
        worst = max(worst, np.max(np.abs(lg - c["logpdf"]) / np.abs(c["logpdf"])))
        # the single-vector R helpers agree with the batch form
        one = orc.MVNPDF(x[0], c["mu"], sigma) if c["kind"] == "mvn" else orc.MVTPDF(x[0], c["mu"], sigma, c["nu"])
        assert abs(one - c["pdf"][0]) <= 1e-12 * c["pdf"][0]
    assert worst < 1e-12


def test_philox_known_answers(orc):
    for k in GOLDEN["philox4x32_10"]:
        assert orc.philox4x32(k["ctr"], k["key"]).tolist() == k["out"]


def test_data_fixture():
    y = np.loadtxt(os.path.join(HERE, "golden", "y_t.csv"), delimiter=",", skiprows=1)
    assert y.shape == (1000, 2) and y[0].tolist() == [0.0, 0.0]      # data_raw/y_t.csv:2
    assert y[:, 1].max() == 0.0 and -2.61 < y[:, 0].min() < -2.59


# ---- dense helpers --------------------------------------------------------------------------------
@pytest.mark.parametrize("d", [1, 2, 5, 16, 32])
def test_lu_det_inverse_cholesky(orc, d):
    rng = np.random.default_rng(d)
    A = rng.standard_normal((d, d)) + 3 * np.eye(d)
    assert abs(orc.determinant(A) - np.linalg.det(A)) <= 1e-10 * abs(np.linalg.det(A))
    assert np.allclose(orc.inverse(A) @ A, np.eye(d), atol=1e-10)
    S = spd(rng, d)
    L = orc.cholesky_lower(S)
    assert np.allclose(L @ L.T, S, atol=1e-12) and np.allclose(L, np.tril(L))
    assert np.allclose(orc.tri_inverse_lower(L) @ L, np.eye(d), atol=1e-12)
    with pytest.raises(np.linalg.LinAlgError):
        orc.cholesky_lower(-np.eye(d))


def test_mvt_float_nu_semantics(orc):
    """nu is a float and nu + n is a float sum: nu = 0.1 is not a double 0.1."""
    S = np.eye(3)
    x = np.array([0.3, -0.2, 0.5])
    nu32 = float(np.float32(0.1))
    q = float(x @ x)
    nun = float(np.float32(0.1) + np.float32(3))
    want = ((math.pi * nu32) ** -1.5) * math.gamma(0.5 * nun) / math.gamma(0.5 * nu32) * (1 + q / nu32) ** (-0.5 * nun)
    assert abs(orc.MVTPDF(x, np.zeros(3), S, 0.1) - want) <= 1e-13 * want
    wrong = ((math.pi * 0.1) ** -1.5) * math.gamma(1.55) / math.gamma(0.05) * (1 + q / 0.1) ** (-1.55)
    assert abs(orc.MVTPDF(x, np.zeros(3), S, 0.1) - wrong) > 1e-9 * wrong


# ---- Metropolis resampler invariants ---------------------------------------
def test_metropolis_invariants(orc):
    rng = np.random.default_rng(1)
    N, B = 200, 10
    u = rng.random((N, B)) * 0.99 + 0.005
    j = rng.integers(0, N, (N, B), dtype=np.uint32)
    assert np.array_equal(orc.metropolis_hastings(np.ones(N), u, j), j[:, -1])   # constant w: last j drawn
    m = 7
    w = np.zeros(N)
    w[m] = 1.0
    a = orc.metropolis_hastings(w, u, j)
    hit = (j == m).any(axis=1)
    hit[m] = True
    assert np.all(a[hit] == m) and np.array_equal(a[~hit], np.arange(N)[~hit])
    assert np.array_equal(orc.metropolis_hastings(rng.random(N), u[:, :0], j[:, :0]), np.arange(N))   # B = 0


# ---- deterministic math -----------------------------------------------------------------------------
def test_det_exp_log_accuracy(orc):
    x = np.concatenate([np.linspace(-700, 0, 20001), np.linspace(0, 700, 2001)])
    rel = np.abs(orc.det_exp(x) - np.exp(x)) / np.exp(x)
    assert rel.max() < 4e-16 * 4                      # a few ulp
    assert orc.det_exp([0.0])[0] == 1.0 and orc.det_exp([-800.0])[0] == 0.0 and np.isinf(orc.det_exp([710.0])[0])
    sub = orc.det_exp([-740.0])[0]                    # subnormal result
    assert 0 < sub < 1e-320 and abs(sub - math.exp(-740.0)) <= 5e-324 * 4
    y = np.concatenate([np.logspace(-307, 307, 20001), np.linspace(0.5, 2.0, 5001), [5e-324, 1e-310]])
    err = np.abs(orc.det_log(y) - np.log(y))
    assert np.all(err <= 4e-16 * np.maximum(np.abs(np.log(y)), 1.0))
    assert orc.det_log([1.0])[0] == 0.0 and np.isneginf(orc.det_log([0.0])[0]) and np.isnan(orc.det_log([-1.0])[0])
    for t in np.linspace(0, 1.9999, 2001):
        s, c = orc.det_sincospi(t)
        assert abs(s - math.sin(math.pi * t)) < 2e-15 and abs(c - math.cos(math.pi * t)) < 2e-15   # pi*t itself rounds


def test_counter_based_normals(orc):
    z = orc.rng_fill_normals(123, 1, 7, 0, 200000, 3)
    assert abs(z.mean()) < 0.01 and abs(z.var() - 1) < 0.01
    assert abs(np.corrcoef(z[:, 0], z[:, 1])[0, 1]) < 0.01
    assert np.array_equal(z[1000:1010], orc.rng_fill_normals(123, 1, 7, 1000, 10, 3))   # keyed by global index
    assert not np.array_equal(z[:10], orc.rng_fill_normals(123, 1, 8, 0, 10, 3))
    u, j = orc.rng_metropolis(5, 2, 64, 10)
    assert np.all((u >= 0) & (u < 1)) and np.all(j < 64)


# ---- fixed-point normalisation / resampling ---------------------------------------------------------
def test_fixed_point_weights(orc):
    assert orc.fixed_shift(1) == 61 and orc.fixed_shift(2) == 60 and orc.fixed_shift(3) == 59
    assert orc.fixed_shift(1 << 20) == 41 and orc.fixed_shift((1 << 20) + 1) == 40
    w = np.array([1.0, 0.5, 0.0, -1.0, np.nan, 0.25, np.inf])
    q, tot = orc.fixed_weights(w, 1.0, 10)
    assert q.tolist() == [1024, 512, 0, 0, 0, 256, 0] and tot == 1792


@pytest.mark.parametrize("N", [1, 2, 3, 17, 1000, 65536])
def test_systematic_matches_textbook(orc, N):
    """Against the floating-point textbook algorithm: equal except where a boundary falls within
    fixed-point resolution (none for these sizes), sorted, counts = floor/ceil of N w / sum w."""
    rng = np.random.default_rng(N)
    w = rng.random(N) + 1e-3
    u0 = rng.random()
    a, rc = orc.resample_systematic(w, u0)
    assert rc == 0
    cdf = np.cumsum(w / w.sum())
    cdf[-1] = 1.0
    tb = np.searchsorted(cdf, (np.arange(N) + u0) / N, side="right")
    assert np.mean(a == tb) > 0.999
    assert np.all(np.diff(a.astype(np.int64)) >= 0)
    cnt = np.bincount(a, minlength=N)
    assert cnt.sum() == N and np.all(np.abs(cnt - N * w / w.sum()) < 1 + 1e-6)


def test_systematic_degenerate(orc):
    a, rc = orc.resample_systematic(np.zeros(5), 0.5)
    assert rc == 1 and a.tolist() == [0, 1, 2, 3, 4]
    w = np.zeros(100)
    w[42] = 3.0
    a, rc = orc.resample_systematic(w, 0.9)
    assert rc == 0 and np.all(a == 42)


def test_multinomial_matches_textbook(orc):
    rng = np.random.default_rng(3)
    N = 5000
    w, u = rng.random(N), rng.random(N)
    a, rc = orc.resample_multinomial(w, u)
    tb = np.searchsorted(np.cumsum(w / w.sum()), u, side="right")
    assert rc == 0 and np.mean(a == np.minimum(tb, N - 1)) > 0.999
    freq = np.bincount(a, minlength=N) / N
    assert abs(freq @ np.arange(N) - (w / w.sum()) @ np.arange(N)) < 5 * N / math.sqrt(N)


def test_logsumexp_ess(orc):
    lw = np.log(np.array([1.0, 2.0, 3.0, 4.0])) - 1000.0
    lse, ess, m = orc.logsumexp_ess(lw)
    assert abs(lse - (math.log(10.0) - 1000.0)) < 1e-12 and abs(ess - 100.0 / 30.0) < 1e-12
    assert m == lw.max()


# ---- propagate / reweight / filter --------------------------------------------------------------------
def test_propagate_reweight_reference_form(orc):
    rng = np.random.default_rng(4)
    N, d, dy = 300, 3, 2
    G, Q, F, V = rng.standard_normal((d, d)), rng.standard_normal((d, d)), rng.standard_normal((dy, d)), spd(rng, dy)
    xp, xi, y = rng.standard_normal((N, d)), rng.standard_normal((N, d)), rng.standard_normal(dy)
    a = rng.integers(0, N, N, dtype=np.uint32)
    x = orc.propagate("mvn", xp, a, G, None, Q, xi)
    assert np.allclose(x, xp[a] @ G.T + xi @ Q.T, atol=1e-13)
    chi = rng.random((N, d)) + 0.5
    xt = orc.propagate("mvt", xp, a, G, None, Q, xi, chi)
    assert np.allclose(xt, xp[a] @ G.T + chi * (xi @ Q.T), atol=1e-13)                   # Q2: chi per component
    from scipy import stats
    w = orc.reweight("mvn", y, x, F, V, faithful=True)
    assert np.allclose(w, stats.multivariate_normal(mean=np.zeros(dy), cov=V).pdf(y - x @ F.T), rtol=1e-11)
    wt = orc.reweight("mvt", y, x, F, V, nu=4.0)
    assert np.allclose(wt, stats.multivariate_t(loc=np.zeros(dy), shape=V, df=4.0).pdf(y - x @ F.T), rtol=1e-11)
    # production-order restatement agrees with the reference form to rounding
    xd, lwd = orc.step_det("mvn", xp, a, G, Q, y, F, V, xi)
    assert np.allclose(xd, x, rtol=1e-13, atol=1e-13)
    assert np.allclose(lwd, np.log(w), rtol=1e-11)
    assert orc.quadform_fma(np.eye(3), np.array([1.0, 2.0, 3.0]), np.array([1.0, 1.0, 1.0]), tri=True) == 0 + 1 + 4


def kalman_means(Y, m0, C0, F, G, V, W):
    m, P = m0.copy(), C0.copy()
    out = [m.copy()]
    for t in range(1, Y.shape[1]):
        m, P = G @ m, G @ P @ G.T + W
        K = P @ F.T @ np.linalg.inv(F @ P @ F.T + V)
        m = m + K @ (Y[:, t] - F @ m)
        P = P - K @ F @ P
        out.append(m.copy())
    return np.array(out), P


@pytest.mark.parametrize("resampler", ["metropolis", "systematic", "multinomial"])
def test_filter_tracks_kalman(orc, resampler):
    """configs[0] in small: the reference model on y_t.csv, against the exact Kalman mean."""
    Y = np.loadtxt(os.path.join(HERE, "golden", "y_t.csv"), delimiter=",", skiprows=1).T[:, :40]
    I = np.eye(2)
    md = dict(m0=np.zeros(2), C0=I, F=I, G=I, V=0.1 * I, W=0.1 * I)
    N = 20000
    s10 = math.sqrt(0.1)
    r = orc.filter_det("mvn", resampler, Y, md["m0"], I, I, I, md["V"], s10 * I, N, seed=3, B=30)
    w = r["w"] if resampler == "metropolis" else np.exp(r["w"] - r["w"].max(axis=1, keepdims=True))
    mean = (w[:, :, None] * r["x"]).sum(1) / w.sum(1)[:, None]
    km, P = kalman_means(Y, **md)
    assert np.max(np.abs(mean[8:] - km[8:])) < 6 * math.sqrt(P[0, 0]) / math.sqrt(N / 4) + 0.02
    if resampler != "metropolis":
        assert np.all(r["ess"][8:] > N / 10)


def test_adaptive_resampling_oracle(orc):
    """ess_threshold in the oracle's filter: steps resample exactly when the previous ESS is below the
    bound, skipped steps keep a_i = i and accumulate the log-weights, and the filter still tracks the
    Kalman mean."""
    Y = np.loadtxt(os.path.join(os.path.dirname(__file__), "golden", "y_t.csv"), delimiter=",", skiprows=1).T[:, :40]
    I2 = np.eye(2)
    N, thr = 20000, 0.5
    r = orc.filter_det("mvn", "systematic", Y, np.zeros(2), I2, I2, I2, 0.1 * I2, np.sqrt(0.1) * I2, N, seed=11,
                       ess_threshold=thr)
    res = r["resampled"]
    assert res[0] == 0 and 0 < res[1:].sum() < len(res) - 1
    assert np.array_equal(res[1:], (r["ess"][:-1] < thr * N).astype(np.int32))
    for t in np.where(res[1:] == 0)[0] + 1:
        assert np.array_equal(r["a"][t], np.arange(N))
    w = np.exp(r["w"] - r["w"].max(axis=1, keepdims=True))
    mean = (w[:, :, None] * r["x"]).sum(1) / w.sum(1)[:, None]
    km, _ = kalman_means(Y, np.zeros(2), I2, I2, I2, 0.1 * I2, 0.1 * I2)
    assert np.max(np.abs(mean[8:] - km[8:])) < 0.05
    always = orc.filter_det("mvn", "systematic", Y, np.zeros(2), I2, I2, I2, 0.1 * I2, np.sqrt(0.1) * I2, N, seed=11)
    assert np.all(always["resampled"][1:] == 1)


def test_filter_reference_loop_equals_det_loop(orc):
    """orc_filter_metropolis (reference-form arithmetic) and orc_filter_det (production order) are two
    statements of src/mcmc.cpp:292-308: same ancestors, states equal to rounding."""
    rng = np.random.default_rng(8)
    N, d, T, B = 400, 2, 10, 10
    I = np.eye(d)
    Y = rng.standard_normal((d, T))
    xi0, xi = rng.standard_normal((N, d)), rng.standard_normal((T - 1, N, d))
    u, j = rng.random((T - 1, N, B)), rng.integers(0, N, (T - 1, N, B), dtype=np.uint32)
    G, Qw, V = 0.9 * I, 0.6 * I, 0.5 * I
    r1 = orc.filter_metropolis("mvn", Y, np.zeros(d), I, I, G, V, Qw, 0.0, xi0, u, j, xi)
    r2 = orc.filter_det("mvn", "metropolis", Y, np.zeros(d), I, I, G, V, Qw, N, B=B, xi0=xi0, xi=xi, u=u, j=j)
    assert np.array_equal(r1["a"][1:], r2["a"][1:])
    assert np.allclose(r1["x"], r2["x"], rtol=1e-12, atol=1e-12)
    assert np.allclose(r1["w"], r2["w"], rtol=1e-11)


def test_filter_mvt_initial_draw_carries_chi(orc):
    """ref: src/mcmc.cpp:73-79 builds ONE distribution object and uses it for the initial draw as well,
    so an "mvt" run starts from x_0 = m0 + chi (.) (Q_c0 xi) (src/statistics.cc.cpp:411).  Both
    statements of the loop take the factors of the initial draw and agree; without them the start
    is the Normal one (the library's mvt_normal_init switch)."""
    rng = np.random.default_rng(9)
    N, d, T, B, nu = 300, 3, 6, 10, 4.0
    I = np.eye(d)
    Y = rng.standard_normal((d, T))
    m0 = rng.standard_normal(d)
    Qc0 = np.linalg.cholesky(spd(rng, d))
    xi0, xi = rng.standard_normal((N, d)), rng.standard_normal((T - 1, N, d))
    chi0, chi = np.sqrt(nu / rng.chisquare(nu, (N, d))), np.sqrt(nu / rng.chisquare(nu, (T - 1, N, d)))
    u, j = rng.random((T - 1, N, B)), rng.integers(0, N, (T - 1, N, B), dtype=np.uint32)
    G, Qw, V = 0.9 * I, 0.6 * I, 0.5 * I
    r1 = orc.filter_metropolis("mvt", Y, m0, Qc0, I, G, V, Qw, nu, xi0, u, j, xi, chi=chi, chi0=chi0)
    r2 = orc.filter_det("mvt", "metropolis", Y, m0, Qc0, I, G, V, Qw, N, nu=nu, B=B, xi0=xi0, xi=xi, u=u, j=j,
                        chi=chi, chi0=chi0)
    want0 = m0 + chi0 * (xi0 @ Qc0.T)
    assert np.allclose(r1["x"][0], want0, rtol=1e-14, atol=1e-14)
    assert np.allclose(r2["x"][0], want0, rtol=1e-14, atol=1e-14)
    assert np.array_equal(r1["a"][1:], r2["a"][1:])
    assert np.allclose(r1["x"], r2["x"], rtol=1e-12, atol=1e-12)
    assert np.allclose(r1["w"], r2["w"], rtol=1e-11)
    r3 = orc.filter_det("mvt", "metropolis", Y, m0, Qc0, I, G, V, Qw, N, nu=nu, B=B, xi0=xi0, xi=xi, u=u, j=j, chi=chi)
    assert np.allclose(r3["x"][0], m0 + xi0 @ Qc0.T, rtol=1e-14, atol=1e-14)
    # faithful reweight (per-particle determinant + inverse, src/mcmc.cpp:193-215) changes no bit
    r4 = orc.filter_metropolis("mvt", Y, m0, Qc0, I, G, V, Qw, nu, xi0, u, j, xi, chi=chi, chi0=chi0, faithful=True)
    assert np.array_equal(r4["w"], r1["w"]) and np.array_equal(r4["a"], r1["a"])


# ---- MH chains ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("kind,nu", [("mvn", 0.0), ("mvt", 6.0)])
def test_mh_chains_oracle_targets_the_right_law(orc, kind, nu):
    rng = np.random.default_rng(12)
    d, Cn, steps = 3, 400, 1500
    S = spd(rng, d)
    L, mu = np.linalg.cholesky(S), rng.standard_normal(d)
    x0 = mu + rng.standard_normal((Cn, d)) @ L.T
    z = rng.standard_normal((Cn, steps, d))
    e = -np.log(rng.random((Cn, steps)))
    thr = e if kind == "mvn" else np.exp(2 * e / (nu + d))
    xf, nacc, bits = orc.mh_chains(kind, mu, L, x0, z, thr, 1.0, nu=nu, shared=True)
    assert 0.15 < bits.mean() < 0.7 and np.array_equal(bits.sum(1), nacc)
    cov = S if kind == "mvn" else S * nu / (nu - 2)
    assert np.all(np.abs(xf.mean(0) - mu) < 5 * np.sqrt(np.diag(cov) / Cn))
    assert np.all(np.abs(np.var(xf, axis=0) / np.diag(cov) - 1) < 0.45)
    # a step with thr = 0 never accepts (mvn: 0.5 (q' - q) < 0 only if q' < q -> still possible), so
    # use thr = -inf for "never" and +inf for "always"
    xa, na, _ = orc.mh_chains(kind, mu, L, x0, z[:, :5], np.full((Cn, 5), np.inf), 1.0, nu=nu, shared=True)
    xn, nn, _ = orc.mh_chains(kind, mu, L, x0, z[:, :5], np.full((Cn, 5), -np.inf), 1.0, nu=nu, shared=True)
    assert np.all(na == 5) and np.all(nn == 0) and np.array_equal(xn, x0)


def test_mh_chains_general_oracle_targets_the_right_law(orc):
    """The general random walk (proposal independent of the target's factor) keeps the MVT target
    invariant: the q of draws from the chains follows d * F(d, nu)."""
    import scipy.stats as st
    rng = np.random.default_rng(14)
    d, Cn, steps, nu = 3, 600, 1500, 6.0
    S = spd(rng, d)
    L, mu = np.linalg.cholesky(S), rng.standard_normal(d)
    x0 = mu + rng.standard_normal((Cn, d)) @ L.T
    z = rng.standard_normal((Cn, steps, d))
    e = -np.log(rng.random((Cn, steps)))
    thr = np.exp(2 * e / (nu + d))
    xf, nacc, bits = orc.mh_chains_general("mvt", mu, L, x0, z, thr, 0.9, nu=nu, shared=True)
    assert 0.2 < bits.mean() < 0.7
    v = np.linalg.solve(L, (xf - mu).T).T
    q = (v * v).sum(1)
    assert st.kstest(q / d, st.f(d, nu).cdf).pvalue > 1e-3
    # scaled proposal: same law
    xf2, _, _ = orc.mh_chains_general("mvt", mu, L, x0, z, thr, 0.9, nu=nu, shared=True, scale=np.array([0.5, 1.0, 1.5]))
    v = np.linalg.solve(L, (xf2 - mu).T).T
    assert st.kstest((v * v).sum(1) / d, st.f(d, nu).cdf).pvalue > 1e-3


def test_tile_image_definition(orc):
    """orc_tile_image DEFINES the block-relative weight image the fused step and the tile update produce.
    Checked here against what it must mean: a monotone integer CDF whose increments are the max-shifted
    weights in fixed point (to one unit of the rescale), the same mass whatever the tile size (to
    rounding), ESS from its integer sums = textbook ESS, exactness on constant weights, zero mass for
    -inf / NaN weights and for tiles far below the global maximum."""
    rng = np.random.default_rng(77)
    N = 10000
    lw = rng.standard_normal(N) * 3.0 - 7.0
    lw[rng.random(N) < 0.02] = -np.inf
    lw[5] = np.nan
    shift = orc.fixed_shift(N)
    lse, ess, lmax = orc.logsumexp_ess(np.where(np.isfinite(lw), lw, -np.inf))
    totals = []
    for tile in (64, 2048, 1 << 20):
        C, T, T2, M = orc.tile_image(lw, tile)
        assert M == lmax
        Ci = C.astype(object)
        assert all(Ci[i] <= Ci[i + 1] for i in range(N - 1)) and int(C[-1]) == T
        q = np.diff(np.concatenate([[0], C.astype(np.float64)]))
        want = np.where(np.isfinite(lw), np.exp(lw - M), 0.0) * 2.0 ** shift
        # one unit from truncating q, one from each of the two rescales of a prefix difference
        assert np.all(np.abs(q - want) <= 3.0 + 1e-12 * want)
        assert q[5] == 0.0 and np.all(q[~np.isfinite(lw)] == 0.0)
        assert np.isclose(T / 2.0 ** shift, np.exp(lse - M), rtol=1e-9)
        assert np.isclose(T * T / (T2 * 2.0 ** shift), ess, rtol=1e-6)
        totals.append(T)
    assert max(totals) - min(totals) <= 2 * N                  # tile sizes differ by rounding only
    # one tile = the global-max image: no rescale at all, increments are trunc(exp(lw - M) 2^shift) exactly
    C, T, _, M = orc.tile_image(lw, 1 << 20)
    q = np.diff(np.concatenate([[0], C.astype(np.float64)]))
    exact = np.array([int(orc.det_exp(v - M)[0] * 2.0 ** shift) if np.isfinite(v) and v - M >= -43.5 else 0 for v in lw[:200]])
    assert np.array_equal(q[:200], exact)
    # constant weights: exact whatever the tile
    for tile in (7, 100, 4096):
        C, T, T2, _ = orc.tile_image(np.full(1000, -3.25), tile)
        assert np.array_equal(C, (np.arange(1000, dtype=np.uint64) + 1) << np.uint64(orc.fixed_shift(1000)))
    # a tile 60 nats below the maximum carries no mass
    lw2 = np.concatenate([np.full(64, -60.0), np.zeros(64)])
    C, T, _, _ = orc.tile_image(lw2, 64)
    assert int(C[63]) == 0 and T == 64 << orc.fixed_shift(128)


def test_metropolis_c2_proposals(orc):
    """Metropolis-C2 (include/cusmc_b200.h): a group of 32 particles shares its proposal segment at every
    iteration, segments are picked in proportion to their length, so proposals are uniform over 0 .. N-1."""
    N, B = 4099, 40                       # ragged: the last segment holds 3 particles
    u, j = orc.rng_metropolis(3, 1, N, B, c2=True)
    assert j.max() < N and u.min() >= 0.0 and u.max() < 1.0
    seg = j // 32
    for g0 in range(0, N, 32):
        assert np.all(seg[g0:g0 + 32] == seg[g0])
    assert len(np.unique(seg[::32])) > 100                 # a fresh segment per (group, iteration)
    counts = np.bincount(j.ravel(), minlength=N)
    bins = np.add.reduceat(counts, np.arange(0, N, 128))
    want = np.diff(np.append(np.arange(0, N, 128), N)) * (B * N / N)
    # proposals inside a group are correlated (common segment): the bin variance is up to 32 x Poisson
    assert np.all(np.abs(bins - want) < 6 * np.sqrt(32 * want))
    u1, j1 = orc.rng_metropolis(3, 1, N, B)
    assert np.array_equal(u, u1) and not np.array_equal(j, j1)      # the lane's own uniform is the plain rule's


def test_rejection_resampler_law(orc):
    """Rejection resampler (include/cusmc_b200.h; the unbiased relative of ref: src/samplers.cpp:21-35):
    a particle keeps itself with probability w_i / wmax, otherwise the accepted proposal is a draw from
    w / sum(w), so E[#children of k] = w_k / wmax + (N - sum(w) / wmax) w_k / sum(w)."""
    rng = np.random.default_rng(11)
    N = 1 << 15
    w = rng.random(N) ** 2 + 1e-3
    w[::97] = 0.0                                                 # dead particles never survive
    wmax = float(w.max())
    a = orc.resample_rejection(w, wmax, seed=5, step=3)
    assert a.dtype == np.uint32 and a.max() < N
    assert np.all(w[a] > 0.0)
    assert np.array_equal(a, orc.resample_rejection(w, wmax, seed=5, step=3))     # counter-based: reproducible
    assert not np.array_equal(a, orc.resample_rejection(w, wmax, seed=5, step=4))
    keep = a == np.arange(N)
    p_keep = w / wmax
    # #self-survivors vs its expectation (proposals that land on i itself add at most 1/N each)
    assert abs(keep.sum() - p_keep.sum()) < 6 * np.sqrt((p_keep * (1 - p_keep)).sum()) + 4
    counts = np.bincount(a, minlength=N).astype(float)
    want = p_keep + (N - p_keep.sum()) * w / w.sum()
    order = np.argsort(w)
    bins = np.add.reduceat(counts[order], np.arange(0, N, 1024))
    wbin = np.add.reduceat(want[order], np.arange(0, N, 1024))
    assert np.all(np.abs(bins - wbin) < 6 * np.sqrt(wbin) + 6)
    # equal weights: everybody is accepted at the first attempt
    assert np.array_equal(orc.resample_rejection(np.full(257, 0.25), 0.25, 1, 0), np.arange(257))
    # cap = 1: one attempt, then the particle holds the first proposal
    a1 = orc.resample_rejection(w, wmax, seed=5, step=3, cap=1)
    u, j = orc.rng_metropolis(5, 3, N, 1)
    assert np.array_equal(a1, np.where(u[:, 0] <= w / wmax, np.arange(N), j[:, 0]))


@pytest.mark.parametrize("kind,nu", [("mvn", 0.0), ("mvt", 5.0)])
def test_perpoint_batch_equals_shared_batch(orc, kind, nu):
    """Per-point parameters (C3's per-chain covariance): the per-point form with every point carrying the
    same (mu, Sigma) is the shared form; with its own Sigma a point equals a one-point shared call."""
    rng = np.random.default_rng(12)
    N, d = 64, 5
    x, mu, S = rng.standard_normal((N, d)), rng.standard_normal(d), spd(rng, d)
    shared = orc.pdf_batch(kind, x, mu, S, nu, log=True)
    pp = orc.pdf_batch_perpoint(kind, x, np.tile(mu, (N, 1)), np.tile(S, (N, 1, 1)), nu, log=True)
    assert np.max(np.abs(pp - shared)) <= 1e-12 * np.max(np.abs(shared))
    S_all = np.stack([spd(rng, d) for _ in range(N)])
    mu_all = rng.standard_normal((N, d))
    pp = orc.pdf_batch_perpoint(kind, x, mu_all, S_all, nu, log=True)
    for i in (0, 17, N - 1):
        one = orc.pdf_batch(kind, x[i:i + 1], mu_all[i], S_all[i], nu, log=True)[0]
        assert abs(pp[i] - one) <= 1e-12 * abs(one)
    lin = orc.pdf_batch_perpoint(kind, x, mu_all, S_all, nu)
    assert np.max(np.abs(np.log(lin) - pp)) < 1e-12 * np.max(np.abs(pp))


def test_normalising_constants(orc):
    """getNorm (ref: src/statistics.cc.cpp:205-211 MVN, :332-340 MVT incl. the float nu + n of Q9)
    against the closed forms; the density at the mean IS the constant."""
    rng = np.random.default_rng(13)
    for d in (1, 2, 3, 8, 16):
        S = spd(rng, d)
        det = np.linalg.det(S)
        want = (2 * math.pi) ** (-d / 2) / math.sqrt(det)
        assert abs(orc.mvn_norm(S) - want) <= 1e-13 * want
        assert abs(orc.pdf_batch("mvn", np.zeros((1, d)), None, S, faithful=True)[0] - want) <= 1e-13 * want
        for nu in (1.0, 3.0, 5.0, 7.5):
            lg = math.lgamma((nu + d) / 2) - math.lgamma(nu / 2) - d / 2 * math.log(math.pi * nu) - 0.5 * math.log(det)
            got = orc.mvt_norm(S, nu)
            assert abs(got - math.exp(lg)) <= 1e-12 * got
            assert abs(orc.pdf_batch("mvt", np.zeros((1, d)), None, S, nu, faithful=True)[0] - got) <= 1e-13 * got
    assert round(orc.mvn_norm(np.eye(2)), 7) == 0.1591549          # CuSMC/CuSMC.tex:104
    assert round(orc.mvt_norm(np.eye(3), 3.0), 8) == 0.07799708    # CuSMC/CuSMC.tex:141


def test_reference_clt_sampler_is_a_scaled_normal():
    """Quirk Q1 (ref: src/statistics.cc.cpp:245-256): the CPU sampler sums n = 200 draws of (z + 1) / 2 with
    z ~ N(0, 1), subtracts n / 2 and divides by sqrt(n / 12).  Algebraically that is sqrt(3) * (sum z / sqrt n):
    a N(0, 3) draw -- what cfg.noise_scale = sqrt(3) reproduces with one standard normal per component
    (and why injected `xi` are "effective standard draws", oracle/cusmc_oracle.h)."""
    rng = np.random.default_rng(14)
    n_iter, d, reps = 200, 3, 20000
    z = rng.standard_normal((reps, n_iter, d))
    s = np.zeros((reps, d))
    for i in range(n_iter):                                       # the reference's loop, line by line
        x = 0.5 * (z[:, i, :] + 1.0)
        s += x
    s = s - n_iter / 2
    x = s / math.sqrt(n_iter / 12)
    xi = z.sum(axis=1) / math.sqrt(n_iter)                        # one N(0, 1) draw per component
    assert np.max(np.abs(x - math.sqrt(3.0) * xi)) < 1e-11
    assert abs(x.var() - 3.0) < 0.06 and abs(x.mean()) < 0.03

import sys, os, time
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np, torch
import cusmc_b200
from oracle_lib import oracle
orc = oracle(); orc.use_all_cores()
ctx = cusmc_b200.Context(0)
def eig(S):
    lam, vec = np.linalg.eigh(S); return vec*np.sqrt(lam)
def td(a): return torch.from_numpy(np.ascontiguousarray(a)).cuda()
ok = True
for (d, N, T, resampler) in [(2, 3000, 6, "systematic"), (8, 5000, 8, "systematic"), (2, 70001, 5, "systematic"), (8, 3000, 6, "multinomial"), (4, 300000, 4, "systematic")]:
    rng = np.random.default_rng(d+N)
    I = np.eye(d)
    md = dict(m0=np.zeros(d), C0=I, F=I, G=0.9*I, V=0.5*I, W=0.3*I)
    Y = rng.standard_normal((d, T))
    pf = ctx.filter(N=N, Y=Y, resampler=resampler, seed=77, keep_history=True, reproducible_rng=True, **md)
    pf.run(); h = pf.history(); s = pf.summary(); pf.close()
    ref = orc.filter_det("mvn", resampler, Y, md["m0"], eig(md["C0"]), md["F"], md["G"], md["V"], eig(md["W"]), N, seed=77)
    ea = np.array_equal(h["a"], ref["a"]); ex = np.array_equal(h["x"], ref["x"]); ew = np.array_equal(h["lw"], ref["w"])
    print(d, N, T, resampler, "a", ea, "x", ex, "lw", ew, "ess", np.allclose(s["ess"], ref["ess"], rtol=1e-12), "ll", np.allclose(s["loglik"], ref["loglik"], rtol=1e-12, atol=1e-12))
    if not ea:
        bad = np.argwhere(h["a"] != ref["a"]); print("  first mismatches", bad[:5], h["a"][tuple(bad[0])], ref["a"][tuple(bad[0])])
    ok &= ea and ex and ew
# timing C5 shard
d, N, T = 8, 8 << 20, 21
I = np.eye(d); Y = np.random.default_rng(5000).standard_normal((d, T))
for fast in (True, False):
    pf = ctx.filter(N=N, Y=Y, m0=np.zeros(d), C0=I, F=I, G=0.9*I, V=I, W=I, resampler="systematic", seed=2, summary=False, reproducible_rng=not fast)
    pf.run(); ctx.synchronize(); pf.run(); ms = pf.last_ms; ess = pf.summary()["ess"]; pf.close()
    print("C5 shard fast=%s: %.1f us/step, roofline %.3f, ess/N %.3f" % (fast, ms/(T-1)*1e3, N*160/(ms/(T-1)*1e-3)/6543.4e9, ess[1:].mean()/N))
Yc = np.loadtxt("tests/golden/y_t.csv", delimiter=",", skiprows=1).T[:, :101]
I2 = np.eye(2)
pf = ctx.filter(N=1000000, Y=Yc, m0=np.zeros(2), C0=I2, F=I2, G=I2, V=0.1*I2, W=0.1*I2, resampler="systematic", seed=1, summary=False)
pf.run(); ctx.synchronize(); pf.run(); ms = pf.last_ms; pf.close()
print("C4 (two launches/step): %.1f us/step" % (ms/100*1e3))
print("ALL OK" if ok else "MISMATCH")

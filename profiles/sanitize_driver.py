"""Small invocations of the kernels touched late in round 2, written for compute-sanitizer (memcheck / racecheck;
the tool is closed on this GPU pool -- gpurun refuses it -- so it has only been run plain, as a smoke driver):
dense fused step (shared-memory operators; d = 8 exact, d = 3 padded), persistent dense kernel, back-to-back
dependent-launch density calls, the dependent-launch step / Metropolis pair, rejection-free chi factors, MH chains
with the throughput generator.  usage: compute-sanitizer --tool memcheck python profiles/sanitize_driver.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cusmc_b200  # noqa: E402

rng = np.random.default_rng(1)
ctx = cusmc_b200.Context(0)
ctx.use_torch_stream()


def spd(d):
    A = rng.standard_normal((d, d))
    return A @ A.T / d + np.eye(d)


for d, N, persistent in ((8, 5000, False), (3, 4500, False), (4, 6000, True), (16, 3000, False)):
    A = rng.standard_normal((d, d)) * 0.1
    md = dict(m0=np.zeros(d), C0=spd(d), F=np.eye(d) + A, G=0.8 * np.eye(d) + A.T, V=spd(d), W=spd(d))
    pf = ctx.filter(N=N, Y=rng.standard_normal((d, 5)), resampler="systematic", seed=3, summary=False,
                    persistent=persistent, **md)
    pf.run()
    x, w, a = pf.state()
    assert np.all(np.isfinite(x))
    pf.close()
d, N = 16, 20000
x = torch.randn((d, N), dtype=torch.float64, device="cuda")
out = torch.empty(N, dtype=torch.float64, device="cuda")
for _ in range(4):
    ctx.logpdf_dev("mvn", x, np.zeros(d), spd(d), out)
ctx.synchronize()
I2 = np.eye(2)
pf = ctx.filter(N=3000, Y=rng.standard_normal((2, 6)), m0=np.zeros(2), C0=I2, F=I2, G=I2, V=I2, W=I2,
                resampler="metropolis", seed=4)
pf.run()
pf.summary()
pf.close()
for nu in (5.0, 2.0, 5.5):
    I = np.eye(8)
    pf = ctx.filter(N=4000, Y=rng.standard_normal((8, 4)), m0=np.zeros(8), C0=I, F=I, G=0.9 * I, V=I, W=I,
                    resampler="systematic", distribution="mvt", df=nu, seed=5, summary=False)
    pf.run()
    assert np.all(np.isfinite(pf.state()[0]))
    pf.close()
Cn, d = 64, 32
L = np.stack([np.linalg.cholesky(spd(d)) for _ in range(Cn)])
xs = torch.tensor(rng.standard_normal((Cn, d)), device="cuda")
ctx.set_chain_noise(reproducible=False)
ctx.mh_chains_dev("mvt", torch.zeros((Cn, d), dtype=torch.float64, device="cuda"),
                  torch.tensor(L.transpose(0, 2, 1).copy(), device="cuda"), xs, 20, 0.3, nu=5.0, seed=6)
ctx.mh_chains_general_dev("mvt", torch.zeros((Cn, d), dtype=torch.float64, device="cuda"),
                          torch.tensor(L.transpose(0, 2, 1).copy(), device="cuda"), xs, 20, 0.2, nu=5.0, seed=7)
ctx.synchronize()
print("sanitize_driver done")

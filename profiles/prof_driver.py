"""Small driver profiled under ncu (see profiles/README.md): one or two launches of every hot kernel
at the BASELINE shapes, nothing else.  Not a benchmark -- numbers printed under a profiler are
never reported."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cusmc_b200  # noqa: E402

ctx = cusmc_b200.Context(0)
ctx.use_torch_stream()
which = sys.argv[1] if len(sys.argv) > 1 else "all"

if which in ("all", "logpdf"):
    d, N = 16, 1 << 20
    A = np.random.default_rng(1235).standard_normal((d, d))
    sigma, mu = A @ A.T / d + np.eye(d), np.random.default_rng(1236).standard_normal(d)
    xs = [torch.randn((d, N), dtype=torch.float64, device="cuda") for _ in range(3)]
    out = torch.empty(N, dtype=torch.float64, device="cuda")
    for x in xs:
        ctx.logpdf_dev("mvn", x, mu, sigma, out)
    xa = torch.randn((N, d), dtype=torch.float64, device="cuda")
    ctx.logpdf_dev("mvn", xa, mu, sigma, out, layout=1)
    torch.cuda.synchronize()

if which in ("all", "pf"):
    I2 = np.eye(2)
    Y = np.loadtxt(os.path.join(ROOT, "tests", "golden", "y_t.csv"), delimiter=",", skiprows=1).T[:, :4]
    pf = ctx.filter(N=1000000, Y=Y, m0=np.zeros(2), C0=I2, F=I2, G=I2, V=0.1 * I2, W=0.1 * I2,
                    resampler="systematic", seed=1, summary=False)
    pf.run()
    ctx.synchronize()
    pf.close()
    d = 8
    I = np.eye(d)
    Y = np.random.default_rng(5000).standard_normal((d, 3))
    pf = ctx.filter(N=8 << 20, Y=Y, m0=np.zeros(d), C0=I, F=I, G=0.9 * I, V=I, W=I, resampler="systematic",
                    seed=2, summary=False)
    pf.run()
    ctx.synchronize()
    pf.close()

if which in ("all", "pfmore"):
    # the same shard with Student-t noise / a dense transition / the multinomial search (T = 3: two steps each)
    d = 8
    I = np.eye(d)
    Y = np.random.default_rng(5000).standard_normal((d, 3))
    Gd = 0.9 * np.linalg.qr(np.random.default_rng(11).standard_normal((d, d)))[0]
    for kw in (dict(G=0.9 * I, distribution="mvt", df=5.0), dict(G=Gd), dict(G=0.9 * I, resampler="multinomial")):
        kw.setdefault("resampler", "systematic")
        pf = ctx.filter(N=8 << 20, Y=Y, m0=np.zeros(d), C0=I, F=I, V=I, W=I, seed=2, summary=False, **kw)
        pf.run()
        ctx.synchronize()
        pf.close()

if which in ("all", "mh"):
    Cn, d, steps = 65536, 32, 200
    A = torch.randn((Cn, d, d), dtype=torch.float64, device="cuda")
    L = torch.linalg.cholesky(A @ A.transpose(1, 2) / d + torch.eye(d, dtype=torch.float64, device="cuda"))
    Lcm = L.transpose(1, 2).contiguous()
    mu = torch.zeros((Cn, d), dtype=torch.float64, device="cuda")
    x = (L @ torch.randn((Cn, d, 1), dtype=torch.float64, device="cuda")).squeeze(-1).contiguous()
    for rep in (True, False):        # the host-reproducible generator, then the throughput one
        ctx.set_chain_noise(reproducible=rep)
        ctx.mh_chains_dev("mvt", mu, Lcm, x, steps, 0.3, nu=5.0, seed=3)
        ctx.mh_chains_general_dev("mvt", mu, Lcm, x, steps, 1.2 / np.sqrt(d), nu=5.0, seed=4)
    ctx.set_chain_noise(reproducible=True)
    torch.cuda.synchronize()

if which in ("all", "metropolis"):
    N, B = 1000000, 10
    w = torch.rand(N, dtype=torch.float64, device="cuda")
    a = torch.empty(N, dtype=torch.int32, device="cuda")
    for r in range(2):
        ctx.metropolis_hastings_dev(a, w, B, seed=5, step=1 + r)
    ctx.metropolis_c2_dev(a, w, B, seed=5, step=3)
    torch.cuda.synchronize()

if which in ("all", "perpoint"):
    N, d = 1 << 19, 32
    packed = d * (d + 1) // 2
    Lp = torch.randn((N, packed), dtype=torch.float64, device="cuda") * 0.1
    diag_idx = torch.tensor([k * (k + 1) // 2 + k for k in range(d)], device="cuda")
    Lp[:, diag_idx] = Lp[:, diag_idx].abs() + 1.0
    x = torch.randn((N, d), dtype=torch.float64, device="cuda")
    mu = torch.zeros((N, d), dtype=torch.float64, device="cuda")
    o = torch.empty(N, dtype=torch.float64, device="cuda")
    for _ in range(2):
        ctx.logpdf_perpoint_dev("mvt", x, mu, Lp, o, nu=5.0)
    torch.cuda.synchronize()
print("prof_driver done", ctx.launch_count, "launches")

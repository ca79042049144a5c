"""Live rate of the MH-chain kernel (65 536 chains, MVT nu = 5 target, per-chain factor, in-kernel noise).
usage: python profiles/mh_micro.py [d ...]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cusmc_b200  # noqa: E402

ctx = cusmc_b200.Context(0)
ctx.use_torch_stream()
Cn, steps = 65536, 200
for d in [int(v) for v in sys.argv[1:]] or [32, 8]:
    g = torch.Generator(device="cuda").manual_seed(2000)
    A = torch.randn((Cn, d, d), dtype=torch.float64, device="cuda", generator=g)
    L = torch.linalg.cholesky(A @ A.transpose(1, 2) / d + torch.eye(d, dtype=torch.float64, device="cuda"))
    Lcm = L.transpose(1, 2).contiguous()
    mu = torch.zeros((Cn, d), dtype=torch.float64, device="cuda")
    x = (L @ torch.randn((Cn, d, 1), dtype=torch.float64, device="cuda", generator=g)).squeeze(-1).contiguous()
    nacc = torch.zeros(Cn, dtype=torch.int32, device="cuda")
    ctx.mh_chains_dev("mvt", mu, Lcm, x, 5, 0.3, nu=5.0, seed=3, n_accept=nacc)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ctx.mh_chains_dev("mvt", mu, Lcm, x, steps, 0.3, nu=5.0, seed=4, n_accept=nacc)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print("d=%d: %.3f ms, %.2e chain-steps/s, accept %.3f" % (d, ms, Cn * steps / (ms * 1e-3), float(nacc.double().mean()) / steps))

"""Live time of the per-point-covariance density kernel (N = 2^19, packed lower factors, MVT nu = 5).
usage: python profiles/perpoint_micro.py [d ...]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cusmc_b200  # noqa: E402

ctx = cusmc_b200.Context(0)
ctx.use_torch_stream()
N = 1 << 19
for d in [int(v) for v in sys.argv[1:]] or [32, 16]:
    packed = d * (d + 1) // 2
    Lp = torch.randn((N, packed), dtype=torch.float64, device="cuda") * 0.1
    diag_idx = torch.tensor([k * (k + 1) // 2 + k for k in range(d)], device="cuda")
    Lp[:, diag_idx] = Lp[:, diag_idx].abs() + 1.0
    x = torch.randn((N, d), dtype=torch.float64, device="cuda")
    mu = torch.zeros((N, d), dtype=torch.float64, device="cuda")
    o = torch.empty(N, dtype=torch.float64, device="cuda")
    for _ in range(3):
        ctx.logpdf_perpoint_dev("mvt", x, mu, Lp, o, nu=5.0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 20
    e0.record()
    for _ in range(reps):
        ctx.logpdf_perpoint_dev("mvt", x, mu, Lp, o, nu=5.0)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / reps
    bytes_per = 8 * d + 8 * d + 8 * packed + 8
    print("d=%d: %.1f us, %.2e evals/s, %.0f GB/s" % (d, us, N / (us * 1e-6), N * bytes_per / (us * 1e-6) / 1e9))

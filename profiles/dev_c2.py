"""Metropolis vs Metropolis-C2 at C4 size (10^6 weights, B = 10): device time per call."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cusmc_b200
ctx = cusmc_b200.Context(0)
ctx.use_torch_stream()
for N in (1000000, 8388608):
    w = torch.rand(N, dtype=torch.float64, device="cuda")
    a = torch.empty(N, dtype=torch.int32, device="cuda")
    for name, fn in (("metropolis", lambda s: ctx.metropolis_hastings_dev(a, w, 10, seed=5, step=s)),
                     ("metropolis_c2", lambda s: ctx.metropolis_c2_dev(a, w, 10, seed=5, step=s))):
        for s in range(5):
            fn(s + 1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for s in range(50):
            fn(s + 10)
        e1.record()
        torch.cuda.synchronize()
        print(f"N={N} {name}: {e0.elapsed_time(e1) / 50 * 1e3:.1f} us/call")

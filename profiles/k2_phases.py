"""Times the three phases of tile_update_kernel separately (single GPU, C5 shard shape) with CUDA events.
usage: python profiles/k2_phases.py"""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cusmc_b200  # noqa: E402

ctx = cusmc_b200.Context(0)
ctx.use_torch_stream()
lib = ctx.lib
d, N, T = 8, 8 << 20, 40
I = np.eye(d)
Y = np.random.default_rng(5000).standard_normal((d, T))
pf = ctx.filter(N=N, Y=Y, m0=np.zeros(d), C0=I, F=I, G=0.9 * I, V=I, W=I, resampler="systematic", seed=2, summary=False)
h = pf.h
dr = pf._make_draws()
ck = ctx._check
ck(lib.cusmc_filter_begin(h, C.byref(dr)))
ck(lib.cusmc_filter_weigh(h, 0))
ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
acc = np.zeros(5)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for t in range(1, T):
    ck(lib.cusmc_filter_resample(h, t))
    ev[0].record()
    ck(lib.cusmc_filter_propagate(h, t))
    ev[1].record()
    ck(lib.cusmc_filter_weigh_phase(h, t, 0, None))
    ev[2].record()
    ck(lib.cusmc_filter_weigh_phase(h, t, 1, None))
    ev[3].record()
    ck(lib.cusmc_filter_weigh_phase(h, t, 2, None))
    ev[4].record()
    torch.cuda.synchronize()
    if t > 5:
        acc += [ev[i].elapsed_time(ev[i + 1]) * 1e3 for i in range(4)] + [ev[0].elapsed_time(ev[4]) * 1e3]
n = T - 6
print("step kernel %.1f us | K2 phase A %.1f | B %.1f | C %.1f | total %.1f us" % tuple(acc / n))
pf.close()

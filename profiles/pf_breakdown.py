"""Live per-phase device time of the filter step (CUDA events on the launching stream, clocks as in
a normal run -- ncu's per-launch times are serialised and taken at whatever clock the idle GPU
sits at, so they only give the SHARE).  usage: python profiles/pf_breakdown.py [c4|c5] [T]"""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cusmc_b200  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "c4"
T = int(sys.argv[2]) if len(sys.argv) > 2 else 201
ctx = cusmc_b200.Context(0)
ctx.use_torch_stream()
if which == "c4":
    d, N = 2, 1000000
    Y = np.loadtxt(os.path.join(ROOT, "tests", "golden", "y_t.csv"), delimiter=",", skiprows=1).T[:, :T]
    I = np.eye(d)
    kw = dict(m0=np.zeros(d), C0=I, F=I, G=I, V=0.1 * I, W=0.1 * I)
else:
    d, N = 8, 8 << 20
    Y = np.random.default_rng(5000).standard_normal((d, T))
    I = np.eye(d)
    kw = dict(m0=np.zeros(d), C0=I, F=I, G=0.9 * I, V=I, W=I)
pf = ctx.filter(N=N, Y=Y, resampler="systematic", seed=1, summary=False, **kw)
lib, h = ctx.lib, pf.h
pf.run()                      # warm-up: whole run through the normal entry point
torch.cuda.synchronize()
whole = pf.last_ms / (T - 1)
dr = pf._make_draws()
ctx._check(lib.cusmc_filter_begin(h, C.byref(dr)))
ctx._check(lib.cusmc_filter_weigh(h, 0))
names = ("resample", "propagate", "weigh")
fns = (lib.cusmc_filter_resample, lib.cusmc_filter_propagate, lib.cusmc_filter_weigh)
ev = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(T)]
for t in range(1, T):
    ev[t][0].record()
    for k, fn in enumerate(fns):
        ctx._check(fn(h, t))
        ev[t][k + 1].record()
torch.cuda.synchronize()
tot = np.zeros(3)
for t in range(1, T):
    for k in range(3):
        tot[k] += ev[t][k].elapsed_time(ev[t][k + 1])
tot *= 1e3 / (T - 1)
print("%s N=%d d=%d T=%d: whole-run %.2f us/step; per phase (with event gaps) " % (which, N, d, T, whole * 1e3) +
      ", ".join("%s %.2f us" % (n, v) for n, v in zip(names, tot)) + "; sum %.2f us" % tot.sum())
pf.close()

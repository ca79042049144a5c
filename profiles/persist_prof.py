import os, sys
import numpy as np
ROOT = "/root/repo"
sys.path.insert(0, ROOT)
import cusmc_b200, torch
ctx = cusmc_b200.Context(0); ctx.use_torch_stream()
d, N, T = 2, 1000000, 41
Y = np.loadtxt(os.path.join(ROOT, "tests", "golden", "y_t.csv"), delimiter=",", skiprows=1).T[:, :T]
I = np.eye(d)
pf = ctx.filter(N=N, Y=Y, m0=np.zeros(d), C0=I, F=I, G=I, V=0.1 * I, W=0.1 * I, resampler="systematic", seed=1, summary=False)
pf.run(); ctx.synchronize()
print("us/step", pf.last_ms * 1e3 / (T - 1))

"""One C4-shaped run (10^6 particles, d = 2) through the persistent kernel and through the four-launch
step; prints us/step of both.  usage: python profiles/persist_prof.py [T]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cusmc_b200  # noqa: E402

T = int(sys.argv[1]) if len(sys.argv) > 1 else 41
ctx = cusmc_b200.Context(0)
ctx.use_torch_stream()
d, N = 2, 1000000
Y = np.loadtxt(os.path.join(ROOT, "tests", "golden", "y_t.csv"), delimiter=",", skiprows=1).T[:, :T]
I = np.eye(d)
for persistent in (True, False):
    pf = ctx.filter(N=N, Y=Y, m0=np.zeros(d), C0=I, F=I, G=I, V=0.1 * I, W=0.1 * I, resampler="systematic", seed=1,
                    summary=False, persistent=persistent)
    pf.run()
    ctx.synchronize()
    pf.run()
    ctx.synchronize()
    print("persistent=%s: %.2f us/step" % (persistent, pf.last_ms * 1e3 / (T - 1)))
    pf.close()

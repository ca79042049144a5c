"""C4-shaped persistent runs under different weight skews (how much of a step is lookup-span imbalance?).
usage: python profiles/c4_variants.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cusmc_b200  # noqa: E402

ctx = cusmc_b200.Context(0)
T = 201
Y = np.loadtxt(os.path.join(ROOT, "tests", "golden", "y_t.csv"), delimiter=",", skiprows=1).T[:, :T]
I = np.eye(2)
for name, V, persistent, N in (("C4 V=0.1", 0.1, True, 1000000), ("V=100 (flat weights)", 100.0, True, 1000000),
                               ("V=0.01 (skewed)", 0.01, True, 1000000), ("C4 per-step path", 0.1, False, 1000000),
                               ("C4 N=500k", 0.1, True, 500000), ("C4 N=250k", 0.1, True, 250000)):
    pf = ctx.filter(N=N, Y=Y, m0=np.zeros(2), C0=I, F=I, G=I, V=V * I, W=0.1 * I, resampler="systematic", seed=2,
                    summary=False, persistent=persistent)
    pf.run()
    ctx.synchronize()
    pf.run()
    ms = pf.last_ms
    ess = pf.summary()["ess"]
    print("%-24s tile %4d  %.1f us/step  ESS/N %.3f" % (name, pf.tile_size, ms / (T - 1) * 1e3, ess[1:].mean() / N))
    pf.close()

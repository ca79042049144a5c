"""C5-shard filter step with Student-t noise (d = 8, 8 Mi particles) at several nu: live CUDA-event time.
usage: python profiles/mvt_c5.py [nu ...]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cusmc_b200  # noqa: E402

ctx = cusmc_b200.Context(0)
d, N, T = 8, 8 << 20, 21
I = np.eye(d)
Y = np.random.default_rng(5000).standard_normal((d, T))
for nu in [float(v) for v in sys.argv[1:]] or [5.0, 5.5, 3.0, 8.0, 12.0]:
    pf = ctx.filter(N=N, Y=Y, m0=np.zeros(d), C0=I, F=I, G=0.9 * I, V=I, W=I, resampler="systematic", seed=2,
                    summary=False, distribution="mvt", df=nu)
    pf.run()
    ctx.synchronize()
    pf.run()
    print("mvt nu=%g: %.1f us/step" % (nu, pf.last_ms / (T - 1) * 1e3))
    pf.close()

import os, sys
import numpy as np
sys.path.insert(0, "/root/repo")
import cusmc_b200
ctx = cusmc_b200.Context(0); ctx.use_torch_stream()
for d, N, T in ((2, 1000000, 51), (8, 1 << 20, 21)):
    I = np.eye(d)
    Y = np.random.default_rng(1).standard_normal((d, T))
    for dist in ("mvn", "mvt"):
        pf = ctx.filter(N=N, Y=Y, m0=np.zeros(d), C0=I, F=I, G=0.9 * I, V=I, W=I, resampler="systematic", seed=1,
                        summary=False, distribution=dist, df=5.0, persistent=False)
        pf.run(); ctx.synchronize(); pf.run(); ctx.synchronize()
        print("d=%d N=%d %s: %.1f us/step" % (d, N, dist, pf.last_ms * 1e3 / (T - 1)))
        pf.close()

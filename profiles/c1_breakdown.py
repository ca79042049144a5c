import os, sys, time
import numpy as np
sys.path.insert(0, "/root/repo")
import cusmc_b200
ctx = cusmc_b200.default_context()
Y = np.loadtxt("/root/repo/tests/golden/y_t.csv", delimiter=",", skiprows=1).T
I2 = np.eye(2)
N, T = 10000, 1000
for rs in ("metropolis", "systematic"):
    pf = cusmc_b200.ParticleFilter(ctx, N, Y[:, :T], np.zeros(2), I2, I2, I2, 0.1 * I2, 0.1 * I2, resampler=rs, seed=1,
                                   keep_history=True, summary=False)
    for rep in range(2):
        t0 = time.perf_counter(); pf.run(); t1 = time.perf_counter(); ctx.synchronize(); t2 = time.perf_counter()
        h = pf.history(); t3 = time.perf_counter()
        print(rs, "enqueue %.1f ms, sync %.1f ms, device loop %.1f ms, history D2H %.1f ms" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3, pf.last_ms, (t3 - t2) * 1e3))
    pf.close()

"""Where the wall time of the C1 reference-model run() goes (host side included).
usage: python profiles/c1_breakdown.py"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cusmc_b200  # noqa: E402

ctx = cusmc_b200.default_context()
Y = np.loadtxt(os.path.join(ROOT, "tests", "golden", "y_t.csv"), delimiter=",", skiprows=1).T
I2 = np.eye(2)
N, T = 10000, 1000
for rs in ("metropolis", "systematic"):
    for rep in range(3):
        t0 = time.perf_counter()
        pf = cusmc_b200.ParticleFilter(ctx, N, Y[:, :T], np.zeros(2), I2, I2, I2, 0.1 * I2, 0.1 * I2, resampler=rs,
                                       seed=1, keep_history=True, summary=False)
        t1 = time.perf_counter()
        pf.run()
        t2 = time.perf_counter()
        ctx.synchronize()
        t3 = time.perf_counter()
        h = pf.history()
        t4 = time.perf_counter()
        pf.close()
        t5 = time.perf_counter()
        print(rs, "create %.1f ms, enqueue %.1f ms, sync %.1f ms (device loop %.1f ms), history D2H %.1f ms, close %.1f ms"
              % ((t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, pf.last_ms if False else 0.0, (t4 - t3) * 1e3, (t5 - t4) * 1e3))
    for rep in range(3):
        os.environ["CUSMC_RUN_TRACE"] = "1" if rep == 2 else ""
        if rep < 2:
            os.environ.pop("CUSMC_RUN_TRACE")
        t0 = time.perf_counter()
        cusmc_b200.run(N, 2, T, Y, np.zeros(2), I2, I2, I2, 0.1 * I2, 0.1 * I2, 0.0, rs, "mvn", seed=1)
        print(rs, "run() end to end %.1f ms" % ((time.perf_counter() - t0) * 1e3))

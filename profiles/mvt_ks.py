"""KS p-values of the device-drawn Student-t noise (throughput generator) against t_nu, several seeds.
usage: python profiles/mvt_ks.py nu d N"""
import os
import sys

import numpy as np
from scipy import stats

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cusmc_b200  # noqa: E402

nu, d, N = float(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
ctx = cusmc_b200.Context(0)
I = np.eye(d)
ps = []
for seed in range(1, 7):
    pf = ctx.filter(N=N, Y=np.zeros((d, 2)), m0=np.zeros(d), C0=I, F=I, G=np.zeros((d, d)), V=100.0 * I, W=I,
                    resampler="systematic", distribution="mvt", df=nu, seed=seed, keep_history=True)
    h = pf.run().history()
    pf.close()
    for t in range(2):
        for k in range(d):
            ps.append(stats.kstest(h["x"][t][:, k], "t", args=(nu,)).pvalue)
ps = np.array(ps)
print("nu=%g d=%d N=%d: %d KS tests, min p %.4g, median %.3f, fraction < 0.05: %.3f; KS of the p-values vs uniform: %.3f"
      % (nu, d, N, len(ps), ps.min(), np.median(ps), (ps < 0.05).mean(), stats.kstest(ps, "uniform").pvalue))

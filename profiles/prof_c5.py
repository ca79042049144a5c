"""One short C5-shard (d = 8, 8 Mi particles) or C4 (d = 2, 10^6) filter run for ncu captures.
usage: python profiles/prof_c5.py [c5|c4|c5dense] [T] [N]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cusmc_b200  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "c5"
T = int(sys.argv[2]) if len(sys.argv) > 2 else 6
ctx = cusmc_b200.Context(0)
if which.startswith("c5"):
    d, N = 8, 8 << 20
    I = np.eye(d)
    G = 0.9 * I
    if which == "c5dense":
        G = 0.9 * np.linalg.qr(np.random.default_rng(5001).standard_normal((d, d)))[0]
    Y = np.random.default_rng(5000).standard_normal((d, T))
    md = dict(m0=np.zeros(d), C0=I, F=I, G=G, V=I, W=I)
else:
    d, N = 2, 1000000
    I = np.eye(d)
    Y = np.loadtxt(os.path.join(ROOT, "tests", "golden", "y_t.csv"), delimiter=",", skiprows=1).T[:, :T]
    md = dict(m0=np.zeros(d), C0=I, F=I, G=I, V=0.1 * I, W=0.1 * I)
if len(sys.argv) > 3:
    N = int(sys.argv[3])
tile = int(os.environ.get("TILE", "0"))           # 0: the library's automatic tile
pf = ctx.filter(N=N, Y=Y, resampler="systematic", seed=2, summary=False, tile_size=tile, **md)
pf.run()
ctx.synchronize()
pf.run()
ms = pf.last_ms
print("tile %d " % pf.tile_size, end="")
print("%s N=%d: %.1f us/step, %.3f ns/particle-step" % (which, N, ms / (T - 1) * 1e3, ms / (T - 1) * 1e6 / N))
pf.close()

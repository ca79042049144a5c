"""Per-phase time line of the persistent kernel (block 0), from a -DCUSMC_TRACE build:
   make -C cusmc_b200/csrc VARIANT=trace EXTRA=-DCUSMC_TRACE
   CUSMC_B200_LIB=cusmc_b200/libcusmc_b200_trace.so python profiles/persist_trace.py [N]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cusmc_b200  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
T = 101
ctx = cusmc_b200.Context(0)
Y = np.loadtxt(os.path.join(ROOT, "tests", "golden", "y_t.csv"), delimiter=",", skiprows=1).T[:, :T]
I = np.eye(2)
GRID = 1024
buf = torch.zeros((T + GRID, 9), dtype=torch.float64, device="cuda")      # word 0 of a row: 1 = every block stamps
buf[T:, 0] = 1.0
os.environ["CUSMC_TRACE_BUF"] = str(buf.data_ptr())
pf = ctx.filter(N=N, Y=Y, m0=np.zeros(2), C0=I, F=I, G=I, V=0.1 * I, W=0.1 * I, resampler="systematic", seed=2, summary=False)
pf.run()
ctx.synchronize()
pf.run()
ms = pf.last_ms
ctx.synchronize()
allb = buf.cpu().numpy()[T:, 1:]
allb = allb[allb[:, 0] > 0]
b = buf.cpu().numpy()[5:T, 1:]
d = np.diff(b[:, :5], axis=1)
nxt = b[1:, 0] - b[:-1, 4]
# (self-served update, the default build: columns 5 / 6 are block 0's own "everybody has arrived" / "my update is
# done"; with -DCUSMC_PERSIST_SELFUPD=0 they are the last arriver's arrival / end of its update)
arrive0 = (b[:, 7] - b[:, 3]).mean() / 1e3          # block 0: end of its tile -> its arrival registered
wait_last = (b[:, 5] - b[:, 7]).mean() / 1e3        # block 0's arrival -> the last block's arrival (seen by block 0)
update = (b[:, 6] - b[:, 5]).mean() / 1e3           # the update
release = (b[:, 4] - b[:, 6]).mean() / 1e3          # update done -> block 0 goes on
print("   barrier: arrive %.2f  wait for the last block %.2f  update %.2f  release seen %.2f us" % (arrive0, wait_last, update, release))
print("N %d tile %d: %.1f us/step | lookup %.2f  rounds %.2f  weigh %.2f  barrier+update %.2f  loop-around %.2f us (block 0, mean over steps)"
      % (N, pf.tile_size, ms / (T - 1) * 1e3, *(d.mean(axis=0) / 1e3), nxt.mean() / 1e3))
# step 50, every block: when did it start, finish its phases, arrive (relative to the earliest start)
t0 = allb[:, 0].min()
for name, col in (("start", 0), ("lookup done", 1), ("rounds done", 2), ("weigh done", 3), ("arrived", 7), ("all arrived", 5),
                  ("update done", 6)):
    v = (allb[:, col] - t0) / 1e3
    print("   step 50, %3d blocks: %-12s min %.2f  median %.2f  p90 %.2f  max %.2f us" % (len(v), name, v.min(), np.median(v), np.percentile(v, 90), v.max()))
dur = (allb[:, 3] - allb[:, 0]) / 1e3
slow = np.argsort(dur)[-5:]
print("   slowest blocks", slow, dur[slow].round(2), "lookup", ((allb[slow, 1] - allb[slow, 0]) / 1e3).round(2))
pf.close()

"""Per-phase device time of the SHARDED filter step on rank 0 (torchrun, one rank per GPU): the
Python-driven phase loop with CUDA events around every phase and every exchange.
usage: torchrun --nproc-per-node N profiles/sharded_breakdown.py [T]"""
import ctypes as C
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cusmc_b200  # noqa: E402
from cusmc_b200 import sharded  # noqa: E402

T = int(sys.argv[1]) if len(sys.argv) > 1 else 41
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = cusmc_b200.Context(local)
ctx.use_torch_stream()
d, per = 8, 8 << 20
N = per * world
I = np.eye(d)
Y = np.random.default_rng(5000).standard_normal((d, T))
pf = cusmc_b200.ShardedParticleFilter(ctx, N, Y, np.zeros(d), I, I, 0.9 * I, I, I, resampler="systematic",
                                      seed=2, summary=False)
for ex in ("p2p", "nccl"):
    pf.run(exchange=ex)
    torch.cuda.synchronize()
    dist.barrier()
    pf.run(exchange=ex)
    torch.cuda.synchronize()
    if rank == 0:
        print("%s whole run: %.2f us/step" % (ex, pf.last_ms * 1e3 / (T - 1)))
lib, h, ck = ctx.lib, pf.pf.h, ctx._check
names = ["resample", "barrier", "propagate", "xmax", "weigh", "xsums"]
ev = [[torch.cuda.Event(enable_timing=True) for _ in range(7)] for _ in range(T)]
dr = pf.pf._make_draws()
ck(lib.cusmc_filter_begin(h, C.byref(dr)))
pf._after_weights(0)
for t in range(1, T):
    e = ev[t]
    e[0].record()
    ck(lib.cusmc_filter_resample(h, t)); e[1].record()
    sharded.rank_barrier(pf._token); e[2].record()
    ck(lib.cusmc_filter_propagate(h, t)); e[3].record()
    sharded.exchange_max(pf.slots_f64[t]); e[4].record()
    ck(lib.cusmc_filter_weigh(h, t)); e[5].record()
    sharded.exchange_sums(pf.slots_i64[t], rank, world, pf._scratch); e[6].record()
torch.cuda.synchronize()
tot = np.zeros(6)
for t in range(2, T):
    for k in range(6):
        tot[k] += ev[t][k].elapsed_time(ev[t][k + 1])
tot *= 1e3 / (T - 2)
if rank == 0:
    print("rank 0 phases (us/step, event gaps included): " + ", ".join("%s %.1f" % kv for kv in zip(names, tot)) +
          "; sum %.1f" % tot.sum())
dist.barrier()
pf.close()
dist.destroy_process_group()

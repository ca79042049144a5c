"""Per-rank device time of the sharded step's two launches (torchrun; CUSMC_SHARD_TRACE build-in trace).
usage: CUSMC_SHARD_TRACE=1 python -m torch.distributed.run --nproc-per-node N profiles/shard_trace.py"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cusmc_b200  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = cusmc_b200.Context(local)
ctx.use_torch_stream()
d, per, T = 8, 8 << 20, 21
I = np.eye(d)
Y = np.random.default_rng(5000).standard_normal((d, T))
pf = cusmc_b200.ShardedParticleFilter(ctx, per * world, Y, np.zeros(d), I, I, 0.9 * I, I, I, resampler="systematic", seed=2,
                                      summary=False)
for _ in range(2):
    pf.run()
    torch.cuda.synchronize()
    dist.barrier()
print("rank %d: %.1f us/step" % (rank, pf.last_ms / (T - 1) * 1e3))
pf.close()
dist.destroy_process_group()

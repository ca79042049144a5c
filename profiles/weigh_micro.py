"""Live timing of the weight-image pass alone (cusmc_weights_sum_dev) and of the resampling pass
(cusmc_resample_systematic_dev on a prepared image).  usage: python profiles/weigh_micro.py [N]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cusmc_b200  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
ctx = cusmc_b200.Context(0)
ctx.use_torch_stream()
g = torch.Generator(device="cuda").manual_seed(1)
lw = -0.5 * torch.randn(N, dtype=torch.float64, device="cuda", generator=g) ** 2
mx = lw.max().reshape(1).clone()
stats = torch.zeros(4, dtype=torch.int64, device="cuda")
image = torch.zeros(ctx.tile_prefix_words(N), dtype=torch.int64, device="cuda")
anc = torch.empty(N, dtype=torch.int32, device="cuda")
reps = 200


def timed(fn):
    for _ in range(10):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps


t_w = timed(lambda: ctx.weights_sum_dev(lw, True, mx, N, stats, tile_prefix=image))
t_r = timed(lambda: ctx.resample_systematic_dev(lw, True, mx, N, stats[0:1], anc, 0.37, tile_prefix=image))
print("N=%d  weights_sum_dev (memset + weigh) %.2f us   resample_systematic_dev %.2f us" % (N, t_w, t_r))

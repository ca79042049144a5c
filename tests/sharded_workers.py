"""Worker bodies of the multi-process tests (spawned with torch.multiprocessing, one per rank)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def _init(rank, world, port, backend):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group(backend, rank=rank, world_size=world)
    return dist


def offspring_below(C, N, T, r0):
    """#{ i in [0, N) : i T + r0 < C N } in exact integer arithmetic (include/cusmc_b200.h)."""
    rhs = C * N
    if rhs <= r0:
        return 0
    return min(N, -(-(rhs - r0) // T))


def exchange_protocol_worker(rank, world, port, N, out_dir):
    """CPU / gloo: the scalar exchange of cusmc_b200/sharded.py carries a sharded systematic
    resampling to exactly the single-rank oracle's ancestors.  The per-rank "kernels" here are
    the oracle's fixed-point weights and Python integers -- test infrastructure only."""
    import torch
    from oracle_lib import oracle
    from cusmc_b200 import sharded
    dist = _init(rank, world, port, "gloo")
    orc = oracle()
    rng = np.random.default_rng(99)
    w = rng.random(N) ** 6
    w[rng.random(N) < 0.1] = 0.0
    u0 = 0.6180339887
    plan = sharded.ShardPlan(N, world, rank)
    mine = w[plan.lo:plan.hi]
    slot = torch.zeros(sharded.SLOT_WORDS, dtype=torch.int64)
    slot_f = slot.view(torch.float64)
    slot_f[sharded.W_MAX] = mine.max() if mine.size else -np.inf
    sharded.exchange_max(slot_f)
    wmax = float(slot_f[sharded.W_MAX])
    assert wmax == w.max()
    shift = orc.fixed_shift(N)
    q, tot = orc.fixed_weights(mine, wmax, shift)
    slot[sharded.W_SUM] = int(tot)
    slot[sharded.W_SUM2] = int(tot) // 3
    slot[sharded.W_NPOS] = int(np.count_nonzero(q))
    sharded.exchange_sums(slot, rank, world)
    q_all, tot_all = orc.fixed_weights(w, wmax, shift)
    assert int(slot[sharded.W_SUM]) == int(tot_all)
    assert int(slot[sharded.W_NPOS]) == int(np.count_nonzero(q_all))
    assert int(slot[sharded.W_OFFSET]) == int(q_all[:plan.lo].astype(object).sum()) if plan.lo else int(slot[sharded.W_OFFSET]) == 0
    # scatter form: every local parent claims its children's global slots
    T, off = int(slot[sharded.W_SUM]), int(slot[sharded.W_OFFSET])
    r0 = min(int(u0 * float(T)), T - 1)
    pairs, C = [], off
    k_prev = offspring_below(C, N, T, r0)
    for jl, qj in enumerate(q):
        C += int(qj)
        k = offspring_below(C, N, T, r0)
        pairs.extend((i, plan.lo + jl) for i in range(k_prev, k))
        k_prev = k
    gathered = [None] * world
    dist.all_gather_object(gathered, pairs)
    a = np.full(N, -1, dtype=np.int64)
    for part in gathered:
        for i, j in part:
            assert a[i] == -1, "a child slot was claimed twice"
            a[i] = j
    want, rc = orc.resample_systematic(w, u0)
    assert rc == 0 and np.array_equal(a, want.astype(np.int64))
    # every child is owned by exactly one rank of the plan
    owners = np.array([plan.owner(i) for i in range(N)])
    bounds = plan.bounds()
    assert all(bounds[r][0] <= i < bounds[r][1] for i, r in enumerate(owners))
    sharded.rank_barrier(torch.zeros(1, dtype=torch.int32))
    open(os.path.join(out_dir, "ok%d" % rank), "w").write("ok")
    dist.destroy_process_group()


def gpu_filter_worker(rank, world, port, backend, devices, cfg, out_dir):
    """GPU: a sharded filter run (device-drawn noise) writes its shard of the final state."""
    import torch
    import cusmc_b200
    dev = devices[rank]
    torch.cuda.set_device(dev)
    dist = _init(rank, world, port, backend)
    ctx = cusmc_b200.Context(dev)
    ctx.use_torch_stream()
    d, T, N = cfg["d"], cfg["T"], cfg["N"]
    I = np.eye(d)
    Y = np.random.default_rng(cfg["yseed"]).standard_normal((d, T))
    pf = cusmc_b200.ShardedParticleFilter(ctx, N, Y, np.zeros(d), I, I, 0.9 * I, 0.5 * I, 0.3 * I,
                                          resampler=cfg["resampler"], distribution=cfg.get("dist", "mvn"),
                                          df=cfg.get("df", 0.0), seed=cfg["seed"], summary=True,
                                          ess_threshold=cfg.get("ess_threshold", 0.0))
    pf.run(exchange=cfg.get("exchange", "p2p"))
    x, w, a = pf.local_state()
    s = pf.summary()
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), x=x, w=w, a=a, mean=s["mean"], ess=s["ess"],
             loglik=s["loglik"], lo=pf.plan.lo, n=pf.plan.n, status=pf.exchange_status())
    pf.close()
    ctx.close()
    dist.destroy_process_group()

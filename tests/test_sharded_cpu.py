"""World-size-2 `gloo` tests of the sharded filter's host side (no GPU): the shard plan and the
scalar exchange protocol of cusmc_b200/sharded.py."""
import socket

import numpy as np
import pytest

from cusmc_b200 import sharded


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


@pytest.mark.parametrize("N,world", [(10, 1), (10, 3), (1 << 20, 8), (7, 8), (1000003, 4)])
def test_shard_plan_partitions_the_slots(N, world):
    plans = [sharded.ShardPlan(N, world, r) for r in range(world)]
    assert sum(p.n for p in plans) == N
    assert plans[0].lo == 0 and plans[-1].hi == N
    for a, b in zip(plans, plans[1:]):
        assert a.hi == b.lo
    for p in plans:
        assert p.per == plans[0].per and (p.n == 0 or p.owner(p.lo) == p.rank and p.owner(p.hi - 1) == p.rank)
    with pytest.raises(ValueError):
        sharded.ShardPlan(N, world, world)
    with pytest.raises(ValueError):
        sharded.ShardPlan(N, 9, 0)


@pytest.mark.parametrize("N", [5003, 64])
def test_exchange_protocol_world2_gloo(tmp_path, N):
    import torch.multiprocessing as mp
    from sharded_workers import exchange_protocol_worker
    mp.spawn(exchange_protocol_worker, args=(2, free_port(), N, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok0").exists() and (tmp_path / "ok1").exists()


@pytest.mark.parametrize("N,world", [(5003, 3), (5, 4)])
def test_exchange_protocol_ragged_worlds_gloo(tmp_path, N, world):
    """Three ranks with a ragged last shard, and more ranks than a tiny cloud fills (one rank empty)."""
    import torch.multiprocessing as mp
    from sharded_workers import exchange_protocol_worker
    mp.spawn(exchange_protocol_worker, args=(world, free_port(), N, str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / ("ok%d" % r)).exists() for r in range(world))


def test_sharded_filter_needs_a_process_group():
    with pytest.raises(RuntimeError):
        sharded.ShardedParticleFilter(None, 8, np.zeros((2, 3)), None, None, None, None, None, None)

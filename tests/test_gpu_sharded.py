"""Sharded filter on the GPU: `world` processes, peer-mapped state (CUDA IPC), scalar exchange over
torch.distributed.  The sharded run must equal the single-GPU run BIT FOR BIT (ancestors, states,
log-weights): weights are integer fixed point, noise is keyed by the global slot.

With >= 2 visible GPUs the ranks take one GPU each and talk NCCL; on a one-GPU box the two ranks
share cuda:0 (CUDA IPC works across processes on one device, NCCL does not) and the scalars are
staged through gloo -- the kernels and the peer loads / stores are the same."""
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def run_sharded(tmp_path, world, cfg):
    import torch
    import torch.multiprocessing as mp
    from sharded_workers import gpu_filter_worker
    n_gpu = torch.cuda.device_count()
    if n_gpu >= world:
        backend, devices = "nccl", list(range(world))
    else:
        backend, devices = "gloo", [0] * world
    mp.spawn(gpu_filter_worker, args=(world, free_port(), backend, devices, cfg, str(tmp_path)),
             nprocs=world, join=True)
    parts = [np.load(tmp_path / ("rank%d.npz" % r)) for r in range(world)]
    assert [int(p["lo"]) for p in parts] == list(np.cumsum([0] + [int(p["n"]) for p in parts[:-1]]))
    return parts


def run_single(ctx, cfg):
    d, T, N = cfg["d"], cfg["T"], cfg["N"]
    I = np.eye(d)
    Y = np.random.default_rng(cfg["yseed"]).standard_normal((d, T))
    pf = ctx.filter(N=N, Y=Y, m0=np.zeros(d), C0=I, F=I, G=0.9 * I, V=0.5 * I, W=0.3 * I,
                    resampler=cfg["resampler"], distribution=cfg.get("dist", "mvn"), df=cfg.get("df", 0.0),
                    seed=cfg["seed"], keep_history=True, summary=True, ess_threshold=cfg.get("ess_threshold", 0.0))
    h, s = pf.run().history(), pf.summary()
    pf.close()
    return h, s


@pytest.mark.timeout(600)
@pytest.mark.parametrize("exchange", ["p2p", "nccl"])
@pytest.mark.parametrize("resampler,world,N,d", [
    ("systematic", 2, 6000, 2),
    ("systematic", 2, 5001, 8),      # ragged: the last rank owns one slot fewer
    ("systematic", 3, 4097, 4),
    ("metropolis", 2, 4096, 2),
    ("metropolis_c2", 2, 4096, 2),
    ("multinomial", 2, 6000, 2),     # every child searches the global CDF on its own or a peer's weight image
    ("multinomial", 3, 4097, 4),
])
def test_sharded_equals_single_gpu_bitwise(ctx, tmp_path, resampler, world, N, d, exchange):
    """exchange = p2p: scalars through peer-memory mailboxes, whole run enqueued by the library;
    nccl: the same phases with torch.distributed collectives in between."""
    cfg = dict(d=d, T=9, N=N, resampler=resampler, seed=2024, yseed=31, exchange=exchange)
    h, s = run_single(ctx, cfg)
    parts = run_sharded(tmp_path, world, cfg)
    x = np.concatenate([p["x"] for p in parts], axis=1)           # [d][N]
    w = np.concatenate([p["w"] for p in parts])
    a = np.concatenate([p["a"] for p in parts])
    assert np.array_equal(a, h["a"][-1])
    assert np.array_equal(x.T, h["x"][-1])
    assert np.array_equal(w, h.get("lw", h["w"])[-1])      # raw weights: log-weights / densities
    assert all(int(p["status"]) == 0 for p in parts)      # no spin-wait timed out
    for p in parts:          # every rank reports the GLOBAL summary
        if not resampler.startswith("metropolis"):
            assert np.allclose(p["ess"], s["ess"], rtol=1e-12)
            assert np.allclose(p["loglik"], s["loglik"], rtol=1e-12, atol=1e-12)
        assert np.allclose(p["mean"], s["mean"], rtol=1e-9, atol=1e-12)


@pytest.mark.timeout(600)
def test_sharded_adaptive_resampling(ctx, tmp_path):
    """The resample / keep decision comes from the GLOBAL integer sums: every rank takes the same one
    and the sharded run still equals the single-GPU run bit for bit."""
    cfg = dict(d=2, T=25, N=5000, resampler="systematic", seed=12, yseed=33, ess_threshold=0.5)
    h, _ = run_single(ctx, cfg)
    parts = run_sharded(tmp_path, 2, cfg)
    assert np.array_equal(np.concatenate([p["a"] for p in parts]), h["a"][-1])
    assert np.array_equal(np.concatenate([p["x"] for p in parts], axis=1).T, h["x"][-1])
    assert np.array_equal(np.concatenate([p["w"] for p in parts]), h["lw"][-1])
    kept = [t for t in range(1, cfg["T"]) if np.array_equal(h["a"][t], np.arange(cfg["N"]))]
    assert 0 < len(kept) < cfg["T"] - 1


@pytest.mark.timeout(600)
def test_sharded_mvt_noise(ctx, tmp_path):
    cfg = dict(d=4, T=6, N=3000, resampler="systematic", seed=7, yseed=32, dist="mvt", df=5.0)
    h, _ = run_single(ctx, cfg)
    parts = run_sharded(tmp_path, 2, cfg)
    assert np.array_equal(np.concatenate([p["a"] for p in parts]), h["a"][-1])
    assert np.array_equal(np.concatenate([p["x"] for p in parts], axis=1).T, h["x"][-1])

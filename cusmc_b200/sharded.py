"""Particle filter sharded across the GPUs of one NVLink domain -- one process per GPU.

The reference is single-device; this is the north star's "particles shard
naturally across the 8 GPUs" path.  Rank r owns the global particle slots
``[r * per, min((r + 1) * per, N))``, ``per = ceil(N / world)``.  Per step:

    step(t)        peer LOADS : ONE fused kernel -- every block looks its children's parents up in the
                                weight image of step t - 1 (the owning rank's tile records and tile-local
                                CDF, read over NVLink when the parents are remote), gathers the parent
                                states from the owning rank's buffer, propagates, reweights and leaves
                                its tile of the new weight image
    all-reduce MAX            : 8 bytes  (log-weight max)
    tile update(t)            : rescale the tiles against the global max, scan, local sums
    all-gather                : 2 x 8 bytes per rank -> global mass, every rank's CDF offset

Two launches per step.  By default the scalars travel through peer-memory mailboxes inside the tile
update kernel (cusmc_filter_run_sharded); ``exchange="nccl"`` carries them with ``torch.distributed``
collectives between the kernel's three phases.  The particle data never goes through a collective -- it moves by peer loads / stores inside the kernels
(`cudaIpcOpenMemHandle`-mapped buffers, cusmc_filter_ipc_export / _attach).  Weights are integer
fixed point and noise is keyed by the global slot, so a sharded run equals the single-GPU run bit
for bit, whatever the world size.

The exchange helpers below work on any tensor device, which is how the world-size-2 `gloo` test
covers them on CPU (tests/test_sharded_cpu.py).
"""
import ctypes as C

import numpy as np

from . import _lib
from .api import ParticleFilter, shard_size

SLOT_WORDS = 8          # struct StepSlot (csrc/filter.cu), in 8-byte words
W_MAX, W_SUM, W_SUM2, W_NPOS, W_OFFSET = 0, 1, 2, 3, 4


class ShardPlan:
    """Contiguous ownership of N global slots by `world` ranks (the rule cusmc_filter_create applies):
    ceil(N / world) slots per rank, rounded up to whole weight-image tiles of 2048."""

    def __init__(self, N, world, rank):
        if not (0 <= rank < world):
            raise ValueError("rank %d outside 0..%d" % (rank, world - 1))
        if world > _lib.MAX_PEERS:
            raise ValueError("world %d exceeds CUSMC_MAX_PEERS = %d" % (world, _lib.MAX_PEERS))
        self.N, self.world, self.rank = int(N), int(world), int(rank)
        self.per = shard_size(self.N, self.world)
        self.lo = min(self.rank * self.per, self.N)
        self.n = max(0, min(self.per, self.N - self.rank * self.per))

    @property
    def hi(self):
        return self.lo + self.n

    def owner(self, slot):
        return slot // self.per

    def bounds(self):
        return [(min(r * self.per, self.N), min((r + 1) * self.per, self.N)) for r in range(self.world)]


def _staged(t, group):
    """NCCL works on device memory in stream order.  Under `gloo` (CPU test runs, or two ranks
    sharing ONE GPU, which NCCL refuses) device scalars are staged through the host."""
    import torch.distributed as dist
    return t.is_cuda and dist.get_backend(group) == "gloo"


def exchange_max(slot_f64, group=None):
    """slot_f64: the 8-word slot viewed as float64.  Word 0 becomes the max over ranks."""
    import torch.distributed as dist
    word = slot_f64[W_MAX:W_MAX + 1]
    if _staged(word, group):
        h = word.cpu()
        dist.all_reduce(h, op=dist.ReduceOp.MAX, group=group)
        word.copy_(h)
    else:
        dist.all_reduce(word, op=dist.ReduceOp.MAX, group=group)


def exchange_sums(slot_i64, rank, world, scratch=None, group=None):
    """slot_i64: the slot viewed as int64, words 1..3 = this rank's fixed-point sums.  Afterwards
    words 1..3 hold the sums over ranks and word 4 the mass on lower ranks.  Integer arithmetic:
    the result does not depend on the order ranks are combined in."""
    import torch
    import torch.distributed as dist
    mine = slot_i64[W_SUM:W_NPOS + 1]
    if _staged(mine, group):
        h = torch.empty(world * 3, dtype=torch.int64)
        dist.all_gather_into_tensor(h, mine.cpu(), group=group)
        scratch = h.to(mine.device)
    else:
        if scratch is None:
            scratch = torch.empty(world * 3, dtype=torch.int64, device=mine.device)
        dist.all_gather_into_tensor(scratch, mine.contiguous(), group=group)   # flat: rank-major
    per_rank = scratch.view(world, 3)
    slot_i64[W_SUM:W_NPOS + 1] = per_rank.sum(0)
    slot_i64[W_OFFSET] = per_rank[:rank, 0].sum()
    return scratch


def rank_barrier(token, group=None):
    """Stream-ordered barrier: every rank's earlier kernels are complete before any rank's later
    ones start (a one-element all-reduce on the compute stream)."""
    import torch
    import torch.distributed as dist
    if _staged(token, group):
        torch.cuda.synchronize()
        dist.barrier(group=group)
    else:
        dist.all_reduce(token, group=group)


class _DeviceWords:
    """Zero-copy torch view of device memory the C library owns (CUDA array interface)."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr,
                                         "data": (int(ptr), False), "version": 3}


def _device_view(ptr, shape, typestr, device):
    import torch
    return torch.as_tensor(_DeviceWords(ptr, shape, typestr), device=torch.device("cuda", device))


class ShardedParticleFilter:
    """particle_filter() of the reference (src/particle_filter.cpp:6-39) over `world` GPUs.

    Same arguments as ``ParticleFilter``; N is the GLOBAL particle count.  Needs an initialised
    ``torch.distributed`` process group whose ranks each hold one GPU of the same node (NCCL), and
    the context's stream must be torch's current stream so collectives and kernels stay ordered.
    """

    def __init__(self, ctx, N, Y, m0, C0, F, G, V, W, group=None, **kw):
        import torch
        import torch.distributed as dist
        if not dist.is_initialized():
            raise RuntimeError("ShardedParticleFilter needs torch.distributed to be initialised")
        self.ctx, self.group = ctx, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.plan = ShardPlan(N, self.world, self.rank)
        if kw.get("resampler", "metropolis") == "rejection" and self.world > 1:
            raise ValueError("the rejection resampler is single-GPU only")
        ctx.use_torch_stream()
        exchange_timeout = kw.pop("exchange_timeout", None)
        self.pf = ParticleFilter(ctx, N, Y, m0, C0, F, G, V, W, rank=self.rank, world=self.world, **kw)
        if exchange_timeout is None:
            # ranks that time-slice one device (test rigs) wait for each other's kernels to be scheduled
            exchange_timeout = 2.0 if torch.cuda.device_count() >= self.world else 10.0
        ctx._check(ctx.lib.cusmc_filter_set_exchange_timeout(self.pf.h, float(exchange_timeout)))
        self.T, self.d, self.N = self.pf.T, self.pf.d, int(N)
        self.is_log = kw.get("resampler", "metropolis") not in ("metropolis", "rejection", "metropolis_c2")
        self.summary_on = bool(kw.get("summary", True))
        lib, h = ctx.lib, self.pf.h
        if self.world > 1:
            # one-off: exchange the IPC handles of (x[0], x[1], ancestors, weights) and map the peers'
            mine = (C.c_ubyte * (_lib.FILTER_IPC_BUFFERS * _lib.IPC_HANDLE_BYTES))()
            ctx._check(lib.cusmc_filter_ipc_export(h, mine))
            gathered = [None] * self.world
            dist.all_gather_object(gathered, bytes(mine), group=group)
            blob = b"".join(gathered)
            buf = (C.c_ubyte * len(blob)).from_buffer_copy(blob)
            ctx._check(lib.cusmc_filter_ipc_attach(h, buf))
        p = C.c_void_p()
        ctx._check(lib.cusmc_filter_slot_dev(h, 0, C.byref(p)))
        self.slots_i64 = _device_view(p.value, (self.T, SLOT_WORDS), "<i8", ctx.device)
        self.slots_f64 = self.slots_i64.view(torch.float64)
        ctx._check(lib.cusmc_filter_moments_dev(h, C.byref(p)))
        self.moments = _device_view(p.value, (self.T, 2 + self.d), "<f8", ctx.device)
        self._scratch = torch.empty(self.world * 3, dtype=torch.int64, device=self.slots_i64.device)
        self._token = torch.zeros(1, dtype=torch.int32, device=self.slots_i64.device)

    def close(self):
        if getattr(self, "pf", None) is not None:
            import torch.distributed as dist
            if self.world > 1 and dist.is_initialized():
                # nobody unmaps a buffer a peer may still be reading
                import torch
                torch.cuda.synchronize()
                dist.barrier(group=self.group)
            self.pf.close()
            self.pf = None

    def _after_weights(self, t):
        """NCCL formulation of weigh(t): the three phases of the tile update with the scalars carried by
        torch.distributed in between -- all-reduce MAX of the log-weight maximum, all-gather of the
        per-rank fixed-point sums (from which every rank derives the global mass and its CDF offset)."""
        lib, h, ck = self.ctx.lib, self.pf.h, self.ctx._check
        if self.world == 1 or not self.is_log:
            if self.world > 1:
                rank_barrier(self._token, self.group)     # reference mode: peers read these densities next
            ck(lib.cusmc_filter_weigh(h, t))
            return
        ck(lib.cusmc_filter_weigh_phase(h, t, 0, None))
        exchange_max(self.slots_f64[t], self.group)
        ck(lib.cusmc_filter_weigh_phase(h, t, 1, None))
        gathered = exchange_sums(self.slots_i64[t], self.rank, self.world, self._scratch, self.group)
        ck(lib.cusmc_filter_weigh_phase(h, t, 2, gathered.data_ptr()))

    def run(self, xi0=None, xi=None, chi=None, u=None, j=None, u0=None, um=None, exchange="p2p"):
        """Injected draws (optional) are this rank's shard; omitted ones come from Philox keyed by
        the global slot.  Returns self; everything is enqueued on the stream.

        exchange="p2p" (default): the library enqueues the whole run and the per-step scalars
        travel through peer-memory mailboxes (cusmc_filter_run_sharded) -- no collective and no
        Python inside the time loop.  exchange="nccl": the same phases driven from here with
        torch.distributed collectives in between (the reference formulation of the exchange)."""
        import torch.distributed as dist
        lib, h, ck = self.ctx.lib, self.pf.h, self.ctx._check
        dr = self.pf._make_draws(xi0=xi0, xi=xi, chi=chi, u=u, j=j, u0=u0, um=um)
        if exchange not in ("p2p", "nccl"):
            raise ValueError("exchange must be 'p2p' or 'nccl'")
        if exchange == "p2p" and self.world > 1:
            ck(lib.cusmc_filter_run_sharded(h, C.byref(dr)))
            self._finish_moments()
            return self
        ck(lib.cusmc_filter_begin(h, C.byref(dr)))
        self._after_weights(0)
        ck(lib.cusmc_filter_mark(h, 0))
        for t in range(1, self.T):
            ck(lib.cusmc_filter_resample(h, t))
            if self.world > 1 and not self.is_log:
                rank_barrier(self._token, self.group)     # the state buffer about to be overwritten was read by peers
            ck(lib.cusmc_filter_propagate(h, t))
            self._after_weights(t)
        ck(lib.cusmc_filter_mark(h, 1))
        self._finish_moments()
        return self

    def exchange_status(self):
        """0, or which bounded spin-wait of the last p2p run timed out (results are then void)."""
        st = C.c_uint64(0)
        self.ctx._check(self.ctx.lib.cusmc_filter_exchange_status(self.pf.h, C.byref(st)))
        return int(st.value)

    def _finish_moments(self):
        import torch.distributed as dist
        if self.world > 1 and self.summary_on:
            if _staged(self.moments, self.group):
                hm = self.moments.cpu()
                dist.all_reduce(hm, group=self.group)
                self.moments.copy_(hm)
            else:
                dist.all_reduce(self.moments, group=self.group)
        return self

    @property
    def last_ms(self):
        return self.pf.last_ms

    def status(self):
        """Waits for the last run; raises CusmcError if an exchange timed out (results void) or a step
        had no weight mass.  summary() and local_state() check it too."""
        self.pf.status()

    def summary(self):
        return self.pf.summary()      # cusmc_filter_get_summary reports TIMEOUT / DEGENERATE itself

    def local_state(self):
        """(x [d][n] SoA, weights [n], ancestors [n]) of this rank's shard as numpy arrays."""
        import torch
        lib, h = self.ctx.lib, self.pf.h
        px, pw, pa = C.c_void_p(), C.c_void_p(), C.c_void_p()
        self.pf.status()
        self.ctx._check(lib.cusmc_filter_state_dev(h, C.byref(px), C.byref(pw), C.byref(pa)))
        torch.cuda.synchronize()
        n, per = self.plan.n, self.plan.per
        x = _device_view(px.value, (self.d, per), "<f8", self.ctx.device)[:, :n].cpu().numpy()
        w = _device_view(pw.value, (per,), "<f8", self.ctx.device)[:n].cpu().numpy()
        a = _device_view(pa.value, (per,), "<i4", self.ctx.device)[:n].cpu().numpy()
        return x, w, a.view(np.uint32)

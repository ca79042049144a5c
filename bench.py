#!/usr/bin/env python
"""bench.py -- headline benchmark of the CuSMC hot path on B200 (contract: see the task brief).

Workload (BASELINE.json configs[1]): batched MVN log-density, d = 16, N = 2^20 points per step,
shared covariance, fp64.  A "step" is one pass of the density kernel over one batch; batches
rotate through a pool larger than L2 so no step re-reads cached input.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

N > 1 is launched by torchrun (one rank per GPU); points are sharded across ranks with no
data-path collective (weak scaling: every rank evaluates its own 2^20-point batches).
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

N_POINTS = 1 << 20
DIM = 16
BYTES_PER_EVAL = 8 * DIM + 8          # read the point, write the result
POOL_BATCHES = 8                      # 8 x 128 MiB = 1 GiB of resident input > 126 MB L2
METRIC = "mvn_logpdf_evals_per_sec"
UNIT = "evals/s"


def workload_inputs():
    """Synthetic inputs of SURVEY.md section 8d (C2)."""
    A = np.random.default_rng(1235).standard_normal((DIM, DIM))
    sigma = A @ A.T / DIM + np.eye(DIM)
    mu = np.random.default_rng(1236).standard_normal(DIM)
    return mu, sigma


def bench_config():
    """The workload both arms (ours and --impl reference) are quoted on."""
    return {"workload": "batched MVN log-density, d=16, N=2^20 points/step/GPU, shared covariance, "
                        "fp64 (BASELINE configs[1]); SoA device-resident input",
            "points_per_step_per_gpu": N_POINTS, "dim": DIM,
            "l2_policy": "inputs rotate through a %d-batch pool (%.0f MiB) larger than the 126 MB L2"
                         % (POOL_BATCHES, POOL_BATCHES * N_POINTS * DIM * 8 / 2 ** 20),
            "parallelism": "points sharded across ranks, no data-path collective"}


def ncu_traffic():
    """DRAM bytes per launch of the headline kernel from the committed ncu --set full capture."""
    try:
        for name in ("r02_traffic.json", "r01_traffic.json"):      # regenerated from the final build each round
            path = os.path.join(ROOT, "profiles", name)
            if os.path.exists(path):
                return float(json.load(open(path))["traffic_bytes_per_launch"])
        return None
    except Exception:
        return None


def bind_to_gpu_numa_node(torch, device):
    """The end-to-end leg is bound by host->device copies; with one rank per GPU the pinned staging
    buffers must live on the NUMA node the GPU hangs off (first-touch), or the ranks' DMA traffic
    crosses the socket interconnect.  Best effort: silently does nothing where sysfs says nothing."""
    try:
        bus = torch.cuda.get_device_properties(device).pci_bus_id
        dom = torch.cuda.get_device_properties(device).pci_domain_id
        dev = torch.cuda.get_device_properties(device).pci_device_id
        path = "/sys/bus/pci/devices/%04x:%02x:%02x.0/numa_node" % (dom, bus, dev)
        node = int(open(path).read().strip())
        if node < 0:
            return None
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return node
    except Exception:
        pass
    return None


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "50"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        for line in self.f.read().strip().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1]))
                mx.append(float(c[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.f.name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# CPU arms (the oracle is test infrastructure; bench.py may time it as the reported CPU baseline)
# ------------------------------------------------------------------------------------------------
def cpu_faithful_evals_per_sec(n_sample, repeats=1):
    """The reference's own CPU arithmetic for this path: per-point determinant() + inverse() (LU)
    and the left-to-right quadratic form (src/statistics.cc.cpp:171-180 called per particle,
    src/mcmc.cpp:193-215), OpenMP over points on all host cores."""
    from oracle_lib import oracle
    orc = oracle()
    orc.use_all_cores()
    mu, sigma = workload_inputs()
    x = np.random.default_rng(1234).standard_normal((n_sample, DIM))
    orc.pdf_batch("mvn", x[:4096], mu, sigma, faithful=True)      # warm the thread pool
    best = float("inf")
    for _ in range(repeats):
        t0 = time.perf_counter()
        orc.pdf_batch("mvn", x, mu, sigma, faithful=True)
        best = min(best, time.perf_counter() - t0)
    return n_sample / best, orc.num_threads(), best


def cpu_hoisted_evals_per_sec(n_sample):
    from oracle_lib import oracle
    orc = oracle()
    mu, sigma = workload_inputs()
    x = np.random.default_rng(1234).standard_normal((n_sample, DIM))
    orc.pdf_batch("mvn", x[:4096], mu, sigma, log=True)
    t0 = time.perf_counter()
    orc.pdf_batch("mvn", x, mu, sigma, log=True)
    return n_sample / (time.perf_counter() - t0)


def run_reference(args, rank):
    """--impl reference: the reference's CPU implementation of the path (restated oracle, faithful
    mode; the reference itself needs R + Rcpp + Eigen and cannot be built in this image).  Every step
    is one FULL batch of the workload (2^20 points, ~0.35 s on 16 cores) unless the requested step
    count would make the run last more than a few minutes."""
    if rank != 0:
        return
    n_sample = N_POINTS if (args.steps + args.warmup) * 0.4 <= 240.0 else N_POINTS // 8
    times = []
    from oracle_lib import oracle
    orc = oracle()
    orc.use_all_cores()
    mu, sigma = workload_inputs()
    x = np.random.default_rng(1234).standard_normal((n_sample, DIM))
    for s in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        orc.pdf_batch("mvn", x, mu, sigma, faithful=True)
        dt = time.perf_counter() - t0
        if s >= args.warmup:
            times.append(dt)
    total = sum(times)
    value = n_sample * len(times) / total
    sample = ("one full step = all 2^20 points of the workload" if n_sample == N_POINTS else
              "%d of the workload's 2^20 points per step (bounded: %d steps requested)" % (n_sample, args.steps))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": bench_config(),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": orc.num_threads(), "kind": "port",
                         "sample": sample + "; CPU restatement of the reference (oracle, faithful mode: per-point "
                                            "LU determinant + inverse as src/mcmc.cpp:211-212), OpenMP over points"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# secondary workloads (reported inside the same JSON line; they do not affect `value`)
# ------------------------------------------------------------------------------------------------
NOISE_NOTE = ("in-kernel, counter-based: philox4x32-7 + single-precision Box-Muller on the special-function unit "
              "(24-bit normals, +-6.7 sigma; the throughput generator -- reproducible_rng=1 selects philox4x32-10 + "
              "FFMA-only polynomials, which the oracle mirrors bit for bit)")
CHAIN_NOISE_NOTE = "in-kernel philox4x32-10 + reproducible single-precision Box-Muller (mirrored bit for bit by the oracle)"
CHAIN_FAST_NOISE_NOTE = ("in-kernel, counter-based: proposal normals from philox4x32-7 + single-precision Box-Muller on the "
                         "special-function unit (cusmc_ctx_set_chain_noise(0), the throughput generator of the filter "
                         "kernels); thresholds -log u with the exact fp64 logarithm; 'reproducible_noise' = the same run with "
                         "the default generator the oracle mirrors bit for bit")


def _timed(torch, fn, reps):
    """Mean device time of fn() in ms over reps calls (CUDA events on torch's current stream)."""
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def _pf_line(ctx, N, d, T, model, Y, bytes_per, hbm_gbs, **kw):
    """One filter run (after a warm-up run): device time of the step loop, ESS computed every step."""
    pf = ctx.filter(N=N, Y=Y, resampler="systematic", **model, **kw)
    pf.run()
    ctx.synchronize()
    l0 = ctx.launch_count
    pf.run()
    ms = pf.last_ms
    launches = ctx.launch_count - l0
    ess = pf.summary()["ess"]
    pf.close()
    rate = N * (T - 1) / (ms * 1e-3)
    return {"value": rate, "N": N, "d": d, "T": T, "ms_per_step": ms / (T - 1), "resampler": "systematic, every step",
            "noise": NOISE_NOTE, "ess": True, "ess_mean_over_N": float(np.mean(ess[1:]) / N),
            "bytes_per_particle_step": bytes_per, "roofline_frac": rate * bytes_per / (hbm_gbs * 1e9),
            "launches_per_step": launches / T}


def secondary_benchmarks(ctx, torch, hbm_gbs, quick):
    import cusmc_b200
    out = {}
    I2 = np.eye(2)

    def guarded(name, fn):
        try:
            out[name] = fn()
        except Exception as e:   # a secondary failure must not hide the headline
            out[name] = {"error": repr(e)}

    def c4():
        # C4: bootstrap PF on data_raw/y_t.csv, 1M particles, systematic resampling every step
        Y = np.loadtxt(os.path.join(ROOT, "tests", "golden", "y_t.csv"), delimiter=",", skiprows=1).T
        T = 101 if quick else 1000
        return _pf_line(ctx, 1000000, 2, T, dict(m0=np.zeros(2), C0=I2, F=I2, G=I2, V=0.1 * I2, W=0.1 * I2),
                        Y[:, :T], 64, hbm_gbs, seed=1, summary=False)
    guarded("pf_c4_particle_steps_per_sec", c4)

    def c5(dense):
        # C5 per-GPU shard: d = 8, 8 Mi particles, synthetic state-space model (T = 100)
        d, N, T = 8, 8 << 20, (11 if quick else 100)
        I = np.eye(d)
        Y = np.random.default_rng(5000).standard_normal((d, T))
        G = 0.9 * I
        if dense:                      # a dense transition: the general (non-diagonal) kernel path
            R = np.linalg.qr(np.random.default_rng(5001).standard_normal((d, d)))[0]
            G = 0.9 * R
        line = _pf_line(ctx, N, d, T, dict(m0=np.zeros(d), C0=I, F=I, G=G, V=I, W=I), Y, 160, hbm_gbs, seed=2,
                        summary=False)
        line["model"] = "dense G (0.9 x rotation)" if dense else "diagonal G, F, V, W (BASELINE configs[4])"
        return line
    guarded("pf_c5_shard_particle_steps_per_sec", lambda: c5(False))
    guarded("pf_c5_shard_dense_G_particle_steps_per_sec", lambda: c5(True))

    def c5_mvt():
        # the same shard with Student-t noise and observation density (the reference's "mvt": per-component chi
        # factors, src/statistics.cc.cpp:381-411), nu = 5
        d, N, T = 8, 8 << 20, (11 if quick else 41)
        I = np.eye(d)
        Y = np.random.default_rng(5000).standard_normal((d, T))
        line = _pf_line(ctx, N, d, T, dict(m0=np.zeros(d), C0=I, F=I, G=0.9 * I, V=I, W=I), Y, 160, hbm_gbs, seed=2,
                        summary=False, distribution="mvt", df=5.0)
        line["noise"] = ("normals as above; chi factors sqrt(nu / chi2_nu), nu = 5: chi2_nu = 2 Gamma(nu / 2) as a sum of two unit "
                         "exponentials plus half a squared normal (the rejection-free path of integer nu <= 8, philox4x32-7; "
                         "other nu: Marsaglia-Tsang gammas), single precision")
        return line
    guarded("pf_c5_shard_mvt_particle_steps_per_sec", c5_mvt)

    def c3():
        # C3: 65 536 independent MH chains, MVT target d = 32, per-chain covariance, 1k steps
        Cn, d, steps = 65536, 32, (100 if quick else 1000)
        g = torch.Generator(device="cuda").manual_seed(2000)
        A = torch.randn((Cn, d, d), dtype=torch.float64, device="cuda", generator=g)
        S = A @ A.transpose(1, 2) / d + torch.eye(d, dtype=torch.float64, device="cuda")
        L = torch.linalg.cholesky(S)
        Lcm = L.transpose(1, 2).contiguous()
        del A, S
        mu = torch.zeros((Cn, d), dtype=torch.float64, device="cuda")
        # start in the typical set (x0 = L z): at d = 32 a chain started AT the mode cannot leave it
        x = (L @ torch.randn((Cn, d, 1), dtype=torch.float64, device="cuda", generator=g)).squeeze(-1).contiguous()
        del L
        nacc = torch.zeros(Cn, dtype=torch.int32, device="cuda")
        ctx.use_torch_stream()
        res = {}
        for name, kw, fast in (("target_factor_proposal", {}, True), ("isotropic_proposal", {"proposal": "isotropic"}, True),
                               ("target_factor_proposal_reproducible_noise", {}, False),
                               ("isotropic_proposal_reproducible_noise", {"proposal": "isotropic"}, False)):
            if kw and not hasattr(ctx, "mh_chains_general_dev"):
                continue
            ctx.set_chain_noise(reproducible=not fast)
            x0 = x.clone()
            run = (lambda st, sd: ctx.mh_chains_dev("mvt", mu, Lcm, x0, st, 0.3, nu=5.0, seed=sd, n_accept=nacc)) if not kw \
                else (lambda st, sd: ctx.mh_chains_general_dev("mvt", mu, Lcm, x0, st, 1.2 / math.sqrt(d), nu=5.0, seed=sd,
                                                               n_accept=nacc))
            run(5, 3)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            run(steps, 4)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            res[name] = {"value": Cn * steps / (ms * 1e-3), "ms": ms,
                         "accept_rate": float(nacc.double().mean().item() / steps)}
        ctx.set_chain_noise(reproducible=True)
        first = res["target_factor_proposal"]
        line = {"value": first["value"], "chains": Cn, "d": d, "steps": steps, "ms": first["ms"],
                "target": "mvt nu=5 per-chain L", "noise": CHAIN_FAST_NOISE_NOTE, "accept_rate": first["accept_rate"],
                "reproducible_noise": dict(res["target_factor_proposal_reproducible_noise"], noise=CHAIN_NOISE_NOTE),
                "formulation": "proposal x' = x + s L_c z: whitened coordinates v' = v + s z, q' = |v'|^2; the factor "
                               "whitens the start and un-whitens the result"}
        if "isotropic_proposal" in res:
            g_ = res["isotropic_proposal"]
            flops = 2 * d * (d + 1) / 2 + 4 * d          # forward substitution + residual + sum of squares
            line["general_proposal"] = {
                "value": g_["value"], "ms": g_["ms"], "accept_rate": g_["accept_rate"],
                "reproducible_noise_value": res["isotropic_proposal_reproducible_noise"]["value"],
                "formulation": "proposal x' = x + s z (isotropic random walk): every step evaluates the target "
                               "density, q' = |L_c^-1 (x' - mu_c)|^2 by forward substitution with L_c resident on chip",
                "fp64_flop_per_step": flops, "fp64_tflops": g_["value"] * flops / 1e12,
                "frac_of_fp64_fma_bound_1.7e10": g_["value"] / 1.7e10, "frac_of_hbm_bound_2.5e10": g_["value"] / 2.5e10}
        return line
    guarded("mh_c3_chain_steps_per_sec", c3)

    def metropolis():
        # a4: the reference's own resampler at C4 size (10^6 weights, B = 10), device-drawn (u, j)
        N, B = 1000000, 10
        w = torch.rand(N, dtype=torch.float64, device="cuda")
        a = torch.empty(N, dtype=torch.int32, device="cuda")
        step = [1]

        def call():
            step[0] += 1
            ctx.metropolis_hastings_dev(a, w, B, seed=5, step=step[0])
        us = _timed(torch, call, 50) * 1e3
        line = {"us_per_call": us, "N": N, "B": B, "accept_tests_per_sec": N * B / (us * 1e-6),
                "bytes_per_particle": 12 + 8 * B, "noise": "philox4x32-10 in-kernel (u: 53 bits, j: Lemire)",
                "l2_sector_gbs": N * B * 32 / (us * 1e-6) / 1e9}
        try:
            peak = json.load(open(os.path.join(ROOT, "profiles", "r02_l2_random_sector_peak.json")))["gbs"]
            line["l2_random_sector_peak_gbs"] = peak
            line["frac_of_l2_random_sector_peak"] = line["l2_sector_gbs"] / peak
            line["bound"] = "32-byte L2 sectors pulled by random 8-byte weight reads (peak: profiles/micro/l2_random.cu)"
        except Exception:
            line["bound"] = "32-byte L2 sectors pulled by random 8-byte weight reads (no measured peak committed)"

        # N4: Metropolis-C2, the same rule with a warp's proposals confined to one 256-byte segment per iteration
        def call_c2():
            step[0] += 1
            ctx.metropolis_c2_dev(a, w, B, seed=5, step=step[0])
        us2 = _timed(torch, call_c2, 50) * 1e3
        line["metropolis_c2"] = {"us_per_call": us2, "accept_tests_per_sec": N * B / (us2 * 1e-6),
                                 "speedup_over_scattered_proposals": us / us2,
                                 "note": "8 contiguous sectors per warp and iteration instead of 32 scattered ones (the "
                                         "warp's segments are drawn by its lanes for each other, one Philox block per lane "
                                         "and 32 iterations).  On B200 the whole weight vector sits in the 126 MB L2 and the "
                                         "scattered kernel is issue-bound (Philox-10, 64-bit Lemire, fp64 division: ~190 "
                                         "issue slots per warp-iteration), not sector-bound, so coalescing the proposals "
                                         "does not pay here; the variant is kept for parity with the registry (SURVEY N4)"}
        return line
    guarded("metropolis_c4_resample", metropolis)

    def mvt_logpdf():
        # C2's other half: batched MVT log-density, same shape as the headline
        mu, sigma = workload_inputs()
        x = torch.randn((DIM, N_POINTS), dtype=torch.float64, device="cuda")
        o = torch.empty(N_POINTS, dtype=torch.float64, device="cuda")
        prep = ctx.prepare_density("mvt", mu, sigma, nu=5.0, log=True)
        ms = _timed(torch, lambda: ctx.logpdf_prepared_dev(prep, x, o), 100)
        return {"value": N_POINTS / (ms * 1e-3), "N": N_POINTS, "d": DIM, "nu": 5.0, "us": ms * 1e3,
                "bytes_per_eval": BYTES_PER_EVAL, "roofline_frac": N_POINTS * BYTES_PER_EVAL / (ms * 1e-3) / (hbm_gbs * 1e9),
                "note": "input re-used (142 MB > 126 MB L2)"}
    guarded("mvt_logpdf_evals_per_sec", mvt_logpdf)

    def logpdf_d32():
        # shared covariance at d = 32: 4.4 flop/B, where the fp64 pipe rather than HBM becomes the bound
        d, N = 32, 1 << 20
        rng = np.random.default_rng(3200)
        A = rng.standard_normal((d, d))
        sigma, mu = A @ A.T / d + np.eye(d), rng.standard_normal(d)
        x = torch.randn((d, N), dtype=torch.float64, device="cuda")
        o = torch.empty(N, dtype=torch.float64, device="cuda")
        prep = ctx.prepare_density("mvn", mu, sigma, log=True)
        ms = _timed(torch, lambda: ctx.logpdf_prepared_dev(prep, x, o), 50)
        flops = d * d + 4 * d
        return {"value": N / (ms * 1e-3), "N": N, "d": d, "us": ms * 1e3, "bytes_per_eval": 8 * d + 8,
                "roofline_frac": N * (8 * d + 8) / (ms * 1e-3) / (hbm_gbs * 1e9),
                "fp64_tflops": N * flops / (ms * 1e-3) / 1e12,
                "frac_of_fp64_fma_peak_33.8": N * flops / (ms * 1e-3) / 33.8e12}
    guarded("mvn_logpdf_d32_evals_per_sec", logpdf_d32)

    def perpoint():
        # a1/a2 with one covariance per point (the C3 shape as a batched density): d = 32, packed factors
        N, d = 1 << 19, 32
        packed = d * (d + 1) // 2
        g = torch.Generator(device="cuda").manual_seed(9)
        Lp = torch.randn((N, packed), dtype=torch.float64, device="cuda", generator=g) * 0.1
        diag_idx = torch.tensor([k * (k + 1) // 2 + k for k in range(d)], device="cuda")
        Lp[:, diag_idx] = Lp[:, diag_idx].abs() + 1.0
        x = torch.randn((N, d), dtype=torch.float64, device="cuda", generator=g)
        mu = torch.zeros((N, d), dtype=torch.float64, device="cuda")
        o = torch.empty(N, dtype=torch.float64, device="cuda")
        ms = _timed(torch, lambda: ctx.logpdf_perpoint_dev("mvt", x, mu, Lp, o, nu=5.0), 10)
        bytes_per = 8 * d + 8 * d + 8 * packed + 8
        return {"value": N / (ms * 1e-3), "N": N, "d": d, "ms": ms, "bytes_per_eval": bytes_per,
                "roofline_frac": N * bytes_per / (ms * 1e-3) / (hbm_gbs * 1e9),
                "note": "per-point packed Cholesky factor staged per warp by 1-D TMA bulk copies"}
    guarded("perpoint_logpdf_evals_per_sec", perpoint)

    def c1():
        # C1, the reference's own model: run(N = 10 000, d = 2, T = 1000, Y = y_sim, "metropolis", "mvn")
        # through the R-facing API (host arrays in, weights [T][N] and posterior_x [T][N][d] out),
        # next to the reference-form CPU loop of the oracle on a bounded sample of the same run
        # (T = 101; its draws are pre-generated and not timed -- the reference itself spends most of
        # its time in its 200-draw CLT sampler, src/statistics.cc.cpp:245-258)
        Y = np.loadtxt(os.path.join(ROOT, "tests", "golden", "y_t.csv"), delimiter=",", skiprows=1).T
        N, d, T = 10000, 2, 1000
        kw = dict(m0=np.zeros(2), C0=I2, F=I2, G=I2, V=0.1 * I2, W=0.1 * I2)
        cusmc_b200.run(N, d, 50, Y, df=0.0, resampler="metropolis", distribution="mvn", seed=1, **kw)   # warm-up
        runs = []
        for _ in range(3):
            t0 = time.perf_counter()
            res = cusmc_b200.run(N, d, T, Y, df=0.0, resampler="metropolis", distribution="mvn", seed=1, **kw)
            runs.append(time.perf_counter() - t0)
        ours_s = min(runs)
        entry = {"ours_seconds": ours_s, "ours_seconds_all": runs, "N": N, "d": d, "T": T, "resampler": "metropolis (B = 10)",
                 "particle_steps_per_sec": N * (T - 1) / ours_s,
                 "path": "cusmc_b200.run -> cusmc_run (host model in; %.0f MB of history streamed to the host through "
                         "a two-chunk device ring while the filter runs)"
                         % ((res["weights"].nbytes + res["posterior_x"].nbytes) / 1e6)}
        try:
            from oracle_lib import oracle
            orc = oracle()
            orc.use_all_cores()
            Ts, B = 101, 10
            rng = np.random.default_rng(0)
            xi0, xi = rng.standard_normal((N, d)), rng.standard_normal((Ts - 1, N, d))
            u, j = rng.random((Ts - 1, N, B)), rng.integers(0, N, (Ts - 1, N, B), dtype=np.uint32)
            s10 = np.sqrt(0.1)
            t0 = time.perf_counter()
            orc.filter_metropolis("mvn", Y[:, :Ts], kw["m0"], I2, I2, I2, kw["V"], s10 * I2, 0.0, xi0, u, j, xi,
                                  history=True, faithful=True)
            cpu_s = time.perf_counter() - t0
            entry["cpu_port"] = {"seconds": cpu_s, "T_sample": Ts, "cores": orc.num_threads(),
                                 "particle_steps_per_sec": N * (Ts - 1) / cpu_s,
                                 "note": "oracle reference-form loop, faithful reweight (per-particle LU determinant + "
                                         "inverse, OpenMP over particles); the Metropolis resampling loop is serial "
                                         "(the reference's OpenMP version of it is racy, SURVEY Q7); draws pre-generated, "
                                         "not timed"}
        except Exception as e:
            entry["cpu_port"] = {"error": repr(e)}
        return entry
    guarded("run_c1_reference_model", c1)

    def filter_e2e():
        # the filter end to end: host model + observations in, per-step summary (mean, ESS, log-likelihood)
        # out, wall clock including the construction of the filter -- next to the oracle's loop on the host
        Y = np.loadtxt(os.path.join(ROOT, "tests", "golden", "y_t.csv"), delimiter=",", skiprows=1).T
        N, T = 1000000, (101 if quick else 1000)
        md = dict(m0=np.zeros(2), C0=I2, F=I2, G=I2, V=0.1 * I2, W=0.1 * I2)

        def once():
            t0 = time.perf_counter()
            pf = ctx.filter(N=N, Y=Y[:, :T], resampler="systematic", seed=11, summary=True, **md)
            s_ = pf.run().summary()
            dt = time.perf_counter() - t0
            pf.close()
            return dt, s_
        once()
        ours_s, summ = min((once() for _ in range(3)), key=lambda r: r[0])
        entry = {"ours_seconds": ours_s, "N": N, "d": 2, "T": T, "particle_steps_per_sec": N * (T - 1) / ours_s,
                 "path": "Context.filter(host model) -> run -> summary (mean, ESS, log-likelihood per step on the host); "
                         "wall clock including filter construction and teardown of nothing else",
                 "h2d_bytes": int(Y[:, :T].nbytes + 6 * 32), "d2h_bytes": int(T * (2 + 2) * 8 + T * 64)}
        try:
            from oracle_lib import oracle
            orc = oracle()
            orc.use_all_cores()
            Ts = 6
            s10 = np.sqrt(0.1)
            t0 = time.perf_counter()
            orc.filter_det("mvn", "systematic", Y[:, :Ts], md["m0"], I2, I2, I2, md["V"], s10 * I2, N, seed=11)
            cpu_s = time.perf_counter() - t0
            entry["cpu_port"] = {"seconds": cpu_s, "T_sample": Ts, "cores": orc.num_threads(),
                                 "particle_steps_per_sec": N * (Ts - 1) / cpu_s,
                                 "note": "oracle production-order loop (hoisted observation algebra, counter-based noise "
                                         "mirrored on the host), bounded sample of %d steps" % Ts}
            entry["speedup_vs_cpu_port"] = entry["particle_steps_per_sec"] / entry["cpu_port"]["particle_steps_per_sec"]
        except Exception as e:
            entry["cpu_port"] = {"error": repr(e)}
        return entry
    guarded("filter_e2e_c4", filter_e2e)
    return out


def _sharded_run(pf, torch, dist, exchange):
    pf.run(exchange=exchange)
    torch.cuda.synchronize()
    dist.barrier()
    pf.run(exchange=exchange)
    torch.cuda.synchronize()
    t = torch.tensor([pf.last_ms], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sharded_filter_benchmark(ctx, torch, dist, world, hbm_gbs, quick):
    """C5: synthetic state-space SMC, d = 8, 8 Mi particles per GPU (64 Mi on 8), systematic resampling
    every step: two launches per step and rank; scalars through peer-memory mailboxes (or NCCL), parents'
    weight images and states by peer loads over NVLink."""
    import cusmc_b200
    out = {}
    rank = dist.get_rank()
    d, per = 8, 8 << 20
    N = per * world
    I = np.eye(d)
    try:
        T = 11 if quick else 100
        Y = np.random.default_rng(5000).standard_normal((d, T))
        pf = cusmc_b200.ShardedParticleFilter(ctx, N, Y, np.zeros(d), I, I, 0.9 * I, I, I,
                                              resampler="systematic", seed=2, summary=False)
        res = {ex: _sharded_run(pf, torch, dist, ex) for ex in ("p2p", "nccl")}
        ess = pf.summary()["ess"]
        status = pf.exchange_status()
        pf.close()
        ms = res["p2p"]
        rate = N * (T - 1) / (ms * 1e-3)
        out["pf_c5_sharded_particle_steps_per_sec"] = {
            "value": rate, "N_global": N, "n_gpus": world, "d": d, "T": T, "ms_per_step": ms / (T - 1),
            "resampler": "systematic, every step", "noise": NOISE_NOTE, "ess": True,
            "ess_mean_over_N": float(np.mean(ess[1:]) / N), "bytes_per_particle_step": 160,
            "roofline_frac_per_gpu": rate / world * 160 / (hbm_gbs * 1e9), "launches_per_step_per_rank": 2,
            "exchange": "p2p: the per-step maximum and sums travel through peer-memory mailboxes INSIDE the one-block "
                        "tile-update kernel (CUDA IPC over NVLink), whole run enqueued by the library; parents' weight "
                        "images and states read by peer loads in the fused step kernel",
            "exchange_status": status,
            "ms_per_step_nccl_exchange": res["nccl"] / (T - 1),
            "value_nccl_exchange": N * (T - 1) / (res["nccl"] * 1e-3)}
    except Exception as e:
        out["pf_c5_sharded_particle_steps_per_sec"] = {"error": repr(e)}
    try:
        # the same model with informative observations (V = 0.5 I): the weights are uneven enough for the mass
        # to move BETWEEN shards, so children really descend from parents on other GPUs
        T = 11 if quick else 21
        rng = np.random.default_rng(5002)
        xs, Y = np.zeros(d), np.zeros((d, T))
        for t in range(1, T):
            xs = 0.9 * xs + rng.standard_normal(d)
            Y[:, t] = xs + np.sqrt(0.5) * rng.standard_normal(d)
        pf = cusmc_b200.ShardedParticleFilter(ctx, N, Y, np.zeros(d), I, I, 0.9 * I, 0.5 * I, I,
                                              resampler="systematic", seed=3, summary=False)
        ms = _sharded_run(pf, torch, dist, "p2p")
        ess = pf.summary()["ess"]
        _, _, a = pf.local_state()
        remote = torch.tensor([float(np.count_nonzero(a // pf.plan.per != rank)), float(a.size)], dtype=torch.float64,
                              device="cuda")
        dist.all_reduce(remote)
        frac = float(remote[0] / remote[1])
        pf.close()
        out["pf_c5_sharded_skewed_weights"] = {
            "value": N * (T - 1) / (ms * 1e-3), "N_global": N, "n_gpus": world, "d": d, "T": T, "ms_per_step": ms / (T - 1),
            "model": "V = 0.5 I (informative observations), data simulated from the model",
            "ess_mean_over_N": float(np.mean(ess[1:]) / N), "remote_parent_fraction_last_step": frac,
            "nvlink_gather_bytes_per_step_estimate": frac * N * 8 * d,
            "note": "remote parents are read over NVLink by the fused step kernel (8 d bytes per child whose parent "
                    "lives on another GPU, plus the parents' tile-local CDF words during the lookup)"}
    except Exception as e:
        out["pf_c5_sharded_skewed_weights"] = {"error": repr(e)}
    try:
        # real-NVLink bit-exactness: a 2^20 x world particle filter sharded over the ranks against the SAME
        # filter on rank 0 alone (final states, log-weights and ancestors must be identical bit for bit)
        n1, T = (1 << 20) * world, 6
        Y = np.random.default_rng(5003).standard_normal((d, T))
        md = (np.zeros(d), I, I, 0.9 * I, 0.5 * I, 0.3 * I)
        pf = cusmc_b200.ShardedParticleFilter(ctx, n1, Y, *md, resampler="systematic", seed=77, summary=False)
        pf.run(exchange="p2p")
        xs_, ws_, as_ = pf.local_state()
        st = pf.exchange_status()
        pf.close()
        parts = [None] * world
        dist.gather_object((xs_, ws_, as_, st), parts if rank == 0 else None, dst=0)
        if rank == 0:
            single = ctx.filter(N=n1, Y=Y, m0=md[0], C0=md[1], F=md[2], G=md[3], V=md[4], W=md[5], resampler="systematic",
                                seed=77, summary=False, persistent=False, tile_size=2048)   # the shards' tiles
            single.run()
            x1, w1, a1 = single.state()
            single.close()
            same = (np.array_equal(np.concatenate([p[0] for p in parts], axis=1), x1) and
                    np.array_equal(np.concatenate([p[1] for p in parts]), w1) and
                    np.array_equal(np.concatenate([p[2] for p in parts]), a1))
            out["sharded_equals_single"] = bool(same) and all(p[3] == 0 for p in parts)
            out["sharded_equals_single_config"] = {"N_global": n1, "d": d, "T": T, "compared": "final x, log-weights, ancestors"}
        dist.barrier()
    except Exception as e:
        out["sharded_equals_single"] = {"error": repr(e)}
    try:
        # C3 over the ranks: the 65 536 chains are independent, each rank advances its contiguous share
        # (strong scaling, no data-path collective); one all-reduce of the posterior moments at the end
        Cn_all, d, steps = 65536, 32, (50 if quick else 200)
        Cn = Cn_all // world
        g = torch.Generator(device="cuda").manual_seed(77 + rank)
        A = torch.randn((Cn, d, d), dtype=torch.float64, device="cuda", generator=g)
        L = torch.linalg.cholesky(A @ A.transpose(1, 2) / d + torch.eye(d, dtype=torch.float64, device="cuda"))
        Lcm = L.transpose(1, 2).contiguous()
        mu = torch.zeros((Cn, d), dtype=torch.float64, device="cuda")
        x = (L @ torch.randn((Cn, d, 1), dtype=torch.float64, device="cuda", generator=g)).squeeze(-1).contiguous()
        del A, L
        nacc = torch.zeros(Cn, dtype=torch.int32, device="cuda")
        ctx.use_torch_stream()
        ctx.set_chain_noise(reproducible=False)                  # the throughput generator, as in the 1-GPU line
        ctx.mh_chains_dev("mvt", mu, Lcm, x, 5, 0.3, nu=5.0, seed=3 + rank, n_accept=nacc)
        dist.all_reduce(torch.cat([x.sum(0), (x * x).sum(0)]))   # warm-up of the reduction too
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ctx.mh_chains_dev("mvt", mu, Lcm, x, steps, 0.3, nu=5.0, seed=40 + rank, n_accept=nacc)
        mom = torch.cat([x.sum(0), (x * x).sum(0)])
        dist.all_reduce(mom)                                   # the final moment reduction (2 d doubles)
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        ctx.set_chain_noise(reproducible=True)
        out["mh_c3_sharded_chain_steps_per_sec"] = {
            "value": Cn * world * steps / (ms * 1e-3), "chains": Cn * world, "chains_per_gpu": Cn, "n_gpus": world,
            "d": d, "steps": steps, "ms": ms, "scaling": "strong", "target": "mvt nu=5 per-chain L",
            "noise": CHAIN_FAST_NOISE_NOTE, "includes": "all-reduce of the 2 d posterior moment sums"}
    except Exception as e:
        out["mh_c3_sharded_chain_steps_per_sec"] = {"error": repr(e)}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=400)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-secondary", action="store_true")
    ap.add_argument("--quick", action="store_true", help="shorter secondary workloads")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank)
        return 0

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        print(json.dumps({"error": "no CUDA device: cusmc_b200 has no CPU fallback"}))
        return 1
    torch.cuda.set_device(local_rank)
    numa_node = bind_to_gpu_numa_node(torch, local_rank) if world > 1 else None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # stdout carries ONE JSON line: whatever NCCL logs (its version banner at NCCL_DEBUG=VERSION / WARN, more at
        # INFO) goes to stderr unless the caller chose a file
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    import cusmc_b200
    ctx = cusmc_b200.Context(local_rank)
    ctx.use_torch_stream()
    hbm_gbs, peak_src = measured_peaks()

    mu, sigma = workload_inputs()
    g = torch.Generator(device="cuda").manual_seed(1234 + rank)
    pool = [torch.randn((DIM, N_POINTS), dtype=torch.float64, device="cuda", generator=g) for _ in range(POOL_BATCHES)]
    out = torch.empty(N_POINTS, dtype=torch.float64, device="cuda")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    prep = ctx.prepare_density("mvn", mu, sigma, log=True)

    def step(i):
        ctx.logpdf_prepared_dev(prep, pool[i % POOL_BATCHES], out)

    for i in range(args.warmup):
        step(i)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.1)
    l0 = ctx.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(args.steps):
        step(i)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = ctx.launch_count - l0
    clocks = sampler.stop()
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = world * N_POINTS * args.steps / (ms_max * 1e-3)
    kernel_ms = ms / args.steps
    achieved = BYTES_PER_EVAL * N_POINTS / (kernel_ms * 1e-3) / 1e9

    # ---- end to end through the host-pointer C ABI: pinned host AoS in, host result out --------
    e2e_steps = max(3, min(args.steps, 20))
    host_x = [torch.randn((N_POINTS, DIM), dtype=torch.float64).pin_memory() for _ in range(2)]
    host_out = torch.empty(N_POINTS, dtype=torch.float64).pin_memory()
    lib = ctx.lib
    mu_c = np.ascontiguousarray(mu)
    sg_c = np.ascontiguousarray(sigma.T).ravel()

    def e2e_step(i):
        rc = lib.cusmc_logpdf(ctx.h, 0, 1, host_x[i % 2].data_ptr(), 1, N_POINTS, N_POINTS, DIM,
                              mu_c.ctypes.data, sg_c.ctypes.data, 0.0, host_out.data_ptr())
        if rc:
            raise RuntimeError(lib.cusmc_last_error(ctx.h).decode())

    for i in range(3):
        e2e_step(i)
    barrier()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        e2e_step(i)
    barrier()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * N_POINTS * e2e_steps / float(te.item())

    # the ceiling of that leg: every rank copying one step's input from pinned host memory at the same
    # time, nothing else (what the box's PCIe links give N concurrent host->device streams)
    dev_buf = torch.empty((N_POINTS, DIM), dtype=torch.float64, device="cuda")
    for _ in range(2):
        dev_buf.copy_(host_x[0], non_blocking=True)
    barrier()
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    c0.record()
    for i in range(5):
        dev_buf.copy_(host_x[i % 2], non_blocking=True)
    c1.record()
    barrier()
    tc = torch.tensor([c0.elapsed_time(c1) / 5.0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tc, op=dist.ReduceOp.MAX)
    h2d_gbs_per_rank = N_POINTS * DIM * 8 / (float(tc.item()) * 1e-3) / 1e9
    h2d_ceiling_evals = world * N_POINTS / (float(tc.item()) * 1e-3)
    del dev_buf

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": bench_config(),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_gbs, "unit": "GB/s",
                     "frac": achieved / hbm_gbs, "traffic": ncu_traffic(), "peak_source": peak_src,
                     "kernel": "density_soa_kernel<16,true,1,true>", "algorithmic_bytes_per_launch": BYTES_PER_EVAL * N_POINTS,
                     "kernel_ms": kernel_ms,
                     "launch": "each call is launched as a programmatic dependent of the one before it on the stream "
                               "(griddepcontrol): its blocks are dispatched while the predecessor drains and touch memory "
                               "only after it has completed; kernel_ms = timed region / launches in that stream of calls. "
                               "One launch alone (ncu launch list, profiles/r02m_launches_summary.md): 22.4 us = 0.97; the "
                               "same stream launched the ordinary way: 23.8 us = 0.92. The peak is a COPY bandwidth "
                               "(read + write); this kernel reads 16 words per word it writes"},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": N_POINTS * DIM * 8,
                "d2h_bytes_per_step": N_POINTS * 8, "steps": e2e_steps,
                "path": "cusmc_logpdf (host pointers, pinned AoS in, pinned result out)",
                "numa_node": numa_node,
                "h2d_ceiling": {"gbs_per_rank_all_ranks_concurrent": h2d_gbs_per_rank, "evals_per_sec": h2d_ceiling_evals,
                                "how": "all ranks copy one step's 128 MiB input from pinned host memory at once, 5 times"},
                "frac_of_h2d_ceiling": e2e_value / h2d_ceiling_evals},
        "gpu_launches": int(launches),
        "clocks": clocks,
    }
    if rank == 0 and world == 1:      # the CPU baseline is reported at N = 1 only
        try:
            v, cores, secs = cpu_faithful_evals_per_sec(N_POINTS, repeats=3)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": "one full step (2^20 points), best of 3, faithful mode (per-point LU "
                                              "det+inverse as src/mcmc.cpp:211-212), %.2f s each" % secs,
                                    "hoisted_value": cpu_hoisted_evals_per_sec(1 << 20)}
        except Exception as e:
            line["cpu_baseline"] = {"error": repr(e)}
    if rank == 0:
        if not args.no_secondary and world == 1:
            line["secondary"] = secondary_benchmarks(ctx, torch, hbm_gbs, args.quick)
    if not args.no_secondary and world > 1:
        sec = sharded_filter_benchmark(ctx, torch, dist, world, hbm_gbs, args.quick)   # collective: all ranks
        if rank == 0:
            line["secondary"] = sec
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())

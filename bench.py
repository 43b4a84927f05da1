#!/usr/bin/env python
"""bench.py -- particle-steps/sec of the 1M-particle linear-Gaussian SMC filter
(BASELINE.json configs[1]) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--dim D] [--particles P] [--T T]

A "step" is ONE full pass of the hot path over one batch of synthetic input:
a T=100-step bootstrap particle filter (propose + weight + resample each step)
over P=1,048,576 particles per GPU, i.e. 1.048e8 particle-steps.
  value : whole-job particle-steps/s, inputs resident in HBM, CUDA-event timed,
          max over ranks; L2 flushed between timed steps.
  e2e   : same metric through the public API (ParticleFilter.run) with HOST
          buffers: pinned x0 / observations / keys copied H2D and the log
          marginal likelihood read back D2H inside the timed region.
  roofline : the fused gather+propose+logpdf model kernel, algorithmic bytes per
          launch / average launch duration (CUDA events, back-to-back launches).
  cpu_baseline : the C/OpenMP restatement of the oracle filter (oracle/c/pf_port.c;
          NumPy oracle if it did not build) timed on the host cores on a bounded
          sample (the reference itself -- GenJAX on jax[cpu] -- is not installable
          in this image: no jax/tfp wheels, no network).
--impl reference times that same port as the reference arm; if `import jax, genjax`
ever succeeds (baseline/_ref populated) it times baseline/run_genjax_cpu.py instead
(the same filter written against GenJAX's own API) and reports kind "reference".
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--dim", type=int, default=1, help="state width d (1 = configs[1] shape 2a; 32 = HBM-bound shape 2b)")
    ap.add_argument("--particles", type=int, default=1 << 20, help="particles per GPU")
    ap.add_argument("--T", type=int, default=100)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--obs-sd", type=float, default=None,
                    help="observation noise sd r (default: 0.5 at d = 1, the BASELINE shape; 0.5 * sqrt(d) at d > 1, which keeps the "
                         "bootstrap filter's effective sample size up so that the ancestor gathers really stream from HBM instead of "
                         "collapsing onto a few L2-resident rows)")
    ap.add_argument("--eager", action="store_true", help="enqueue the launches every run instead of replaying a CUDA graph")
    ap.add_argument("--multi-gpu", default="global", choices=["islands", "global"],
                    help="N>1: global (default, the north star's split) = ONE filter over N*particles with global systematic "
                         "resampling every step (tile records mailed over NVLink peer memory, parents read from peers); the line "
                         "also carries an 'islands' measurement.  islands = every rank filters its own block of particles with "
                         "local resampling and the per-shard log-marginal-likelihood terms are combined by ONE NCCL all-gather per run")
    ap.add_argument("--reference-max", default="running", choices=["running", "analytic"],
                    help="running (default): max + mass passes; analytic: masses relative to an analytic bound, accumulated "
                         "in the model kernel (graph mode, d = 1; DESIGN.md section 10 -- not yet measured on a device)")
    ap.add_argument("--single-pass", action="store_true",
                    help="with --reference-max analytic: ONE launch per step (output-slot resampling of the previous step fused "
                         "into the model kernel, model_kernel_static_pull; DESIGN.md section 10 -- not yet measured on a device)")
    ap.add_argument("--mode", default="step", choices=["persistent", "graph", "step", "steps"],
                    help="step: ONE launch per filter step (pf_step_kernel: output-slot resampling of the previous step + gather + "
                         "propose + logpdf + tile-exponent masses), captured in a CUDA graph; graph: round 1's 2 launches per step; "
                         "persistent: one cooperative launch per filter")
    return ap.parse_args()


# ------------------------------------------------------------------ helpers


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f).get("hbm_gbs", 6650.0), "measured"
    return 6650.0, "fallback"


class ClockSampler:
    """SM clocks / throttle reasons sampled DURING the timed region: in-process NVML polling every ~2 ms (the timed
    region of the default run is only tens of milliseconds), nvidia-smi as a fallback."""

    NAMES = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index: int):
        self.index = index
        self.sm: list[float] = []
        self.mx = None
        self.bits = 0
        self.stop_flag = False
        self.thread = None
        self.how = None

    def _nvml_loop(self):
        import pynvml

        try:
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[self.index]) if vis and all(v.strip().isdigit() for v in vis.split(",")) else self.index
            h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            reasons = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or pynvml.nvmlDeviceGetCurrentClocksThrottleReasons
            self.how = "nvml"
            while not self.stop_flag:
                self.sm.append(float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
                try:
                    self.bits |= int(reasons(h))
                except Exception:
                    pass
                time.sleep(0.002)
        except Exception:
            self.how = None

    def start(self):
        try:
            import pynvml  # noqa: F401

            self.thread = threading.Thread(target=self._nvml_loop, daemon=True)
            self.thread.start()
            time.sleep(0.02)
        except Exception:
            self.thread = None

    def stop(self) -> dict:
        self.stop_flag = True
        if self.thread is not None:
            self.thread.join(timeout=2)
        if self.how == "nvml" and self.sm:
            return {"sm_mhz": float(np.median(self.sm)), "sm_max_mhz": self.mx,
                    "reasons": sorted(n for b, n in self.NAMES.items() if self.bits & b), "samples": len(self.sm), "source": "nvml"}
        return self._smi_once()

    def _smi_once(self) -> dict:
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                 capture_output=True, text=True, timeout=10).stdout.strip().splitlines()[0]
            parts = [p.strip() for p in out.split(",")]
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            return {"sm_mhz": float(parts[0]), "sm_max_mhz": float(parts[1]),
                    "reasons": [n for n, v in zip(names, parts[2:6]) if v.lower().startswith("active")], "samples": 1,
                    "source": "nvidia-smi (one sample right after the timed region)"}
        except Exception:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"], "samples": 0}


def obs_sd(args_or_none, d):
    from genjax_b200.workloads import LG_R

    v = getattr(args_or_none, "obs_sd", None) if args_or_none is not None else None
    return float(v) if v is not None else (LG_R if d == 1 else LG_R * float(np.sqrt(d)))


def synth_obs(T, d, seed=0, r=None):
    """Synthetic observations y_1:T of the LGSSM (NumPy PCG64; no oracle import on the product arm)."""
    from genjax_b200.workloads import LG_A, LG_C, LG_Q, LG_R

    LG_R = LG_R if r is None else r

    g = np.random.default_rng(seed)
    x = g.standard_normal(d)
    ys = np.empty((T, d), dtype=np.float32)
    for t in range(T):
        x = LG_A * x + LG_Q * g.standard_normal(d)
        ys[t] = LG_C * x + LG_R * g.standard_normal(d)
    return ys if d > 1 else ys[:, 0]


def workload_string(n, T, d):
    """config.workload: the SAME string on both arms (the driver compares them)."""
    return (f"linear-Gaussian SSM bootstrap SMC (BASELINE configs[1]): T={T}, N={n} particles/GPU, d={d}, "
            "systematic resampling every step")


def kalman_logz(ys, d, r=None):
    """Exact log p(y_1:T) of the (diagonal) linear-Gaussian model by the Kalman filter, float64 -- the ground truth the
    filter's estimate is checked against inside this script (accept within 4 sigma)."""
    from genjax_b200.workloads import LG_A, LG_C, LG_Q, LG_R

    LG_R = LG_R if r is None else r
    ys = np.asarray(ys, dtype=np.float64).reshape(len(ys), -1)
    total = 0.0
    for j in range(ys.shape[1]):
        m, p = 0.0, 1.0
        for y in ys[:, j]:
            m, p = LG_A * m, LG_A * LG_A * p + LG_Q * LG_Q
            s = LG_C * LG_C * p + LG_R * LG_R
            total += -0.5 * (np.log(2 * np.pi * s) + (y - LG_C * m) ** 2 / s)
            k = p * LG_C / s
            m, p = m + k * (y - LG_C * m), (1 - k * LG_C) * p
    return float(total)


# ---------------------------------------------------------------- CPU legs


def cpu_port_rate(n, T_sample, d, seed=314159, threads=None):
    """particle-steps/s of the NumPy oracle port on `T_sample` filter steps over n particles, using `threads` host
    threads: the propose + weight pass (Philox, Box-Muller, logpdf: >90 % of the work) runs per particle chunk in a
    thread pool (NumPy releases the GIL inside its loops; lanes are global indices, so chunking does not change a
    single number); max / integer mass / CDF / systematic ancestors / gather stay serial NumPy."""
    from concurrent.futures import ThreadPoolExecutor

    from genjax_b200.workloads import LG_A, LG_C, LG_Q, LG_R
    from oracle import gfi as ogfi
    from oracle import rng as orng
    from oracle import smc as osmc

    threads = threads or os.cpu_count() or 1
    ys = synth_obs(T_sample, d)
    g = np.random.default_rng(1)
    if d == 1:
        x0 = g.standard_normal(n).astype(np.float32)

        def step(h, x_prev):
            x = h.normal("x", np.float32(LG_A) * x_prev, np.float32(LG_Q))
            h.normal("y", np.float32(LG_C) * x, np.float32(LG_R))
            return x

        shared = ()
    else:
        x0 = g.standard_normal((n, d)).astype(np.float32)
        q = np.full(d, LG_Q, np.float32)
        r = np.full(d, LG_R, np.float32)

        def step(h, x_prev, q, r):
            x = h.mv_normal_diag("x", np.float32(LG_A) * x_prev, q)
            h.mv_normal_diag("y", np.float32(LG_C) * x, r)
            return x

        shared = (q, r)
    obs = [{"y": (np.float32(y) if d == 1 else y)} for y in ys]
    key = orng.key(seed)
    chunk = max(4, ((n + threads - 1) // threads + 3) // 4 * 4)  # quad-aligned chunks
    bounds = [(a, min(n, a + chunk)) for a in range(0, n, chunk)]
    pool = ThreadPoolExecutor(max_workers=threads)
    t0 = time.perf_counter()
    x = x0
    logz = 0.0
    for t, ob in enumerate(obs):
        k_prop, k_res = osmc.pf_step_keys(key, t)
        keys = orng.split(k_prop, n)

        def part(ab, keys=keys, x=x, ob=ob):
            a, b = ab
            tr, w = ogfi.generate(step, keys[a:b], ob, (x[a:b],) + tuple(shared))
            return tr.retval, w

        outs = list(pool.map(part, bounds))
        xs = np.concatenate([o[0] for o in outs])
        w = np.concatenate([o[1] for o in outs])
        logz += osmc.log_mean_exp(w)
        x = xs[osmc.resample_systematic(w, k_res)]
    dt = time.perf_counter() - t0
    pool.shutdown()
    return n * T_sample / dt, dt, logz, threads


def cpu_c_port_rate(n, T_sample, d, seed=314159, fast=True):
    """particle-steps/s of the C / OpenMP restatement of the same filter on all host threads; None when gcc is
    unavailable.  fast=True (the CPU arm): oracle/c/pf_port_fast.c, written for speed (one Philox block per quad, float
    libm, -O3 -march=native, gather fused into the resampling); fast=False: oracle/c/pf_port.c, the bit-compatible
    restatement the tests use (tests/test_oracle_c_port.py), several times slower by construction."""
    from genjax_b200.core.key import key as pkey, pf_key_table
    from genjax_b200.workloads import LG_A, LG_C, LG_Q, LG_R
    from oracle import cport

    if (cport.fast_lib() if fast else cport.lib()) is None:
        return None
    threads = os.cpu_count() or 1  # torchrun exports OMP_NUM_THREADS=1: ask for the cores explicitly
    ys = synth_obs(T_sample, d)
    g = np.random.default_rng(1)
    x0 = g.standard_normal(n if d == 1 else (n, d)).astype(np.float32)
    tab = pf_key_table(pkey(seed), T_sample)
    t0 = time.perf_counter()
    out = cport.pf_lgssm(x0, ys, LG_A, LG_Q, LG_C, LG_R, tab, fast=fast, threads_=threads)
    dt = time.perf_counter() - t0
    return n * T_sample / dt, dt, float(out["logz_inc"].sum()), threads


def genjax_reference_rate(args):
    """The real reference (GenJAX on jax[cpu]) through baseline/run_genjax_cpu.py, or None where it cannot be
    imported -- which is the case in this image (no jax / tfp / genjax wheels, no network)."""
    ref_dir = os.path.join(ROOT, "baseline", "_ref")
    if os.path.isdir(ref_dir) and ref_dir not in sys.path:
        sys.path.insert(0, ref_dir)
    try:
        import jax  # noqa: F401
        import genjax  # noqa: F401
    except Exception:
        return None
    try:
        sys.path.insert(0, os.path.join(ROOT, "baseline"))
        import run_genjax_cpu

        return run_genjax_cpu.measure(1, args.particles, args.T, args.dim, repeats=max(1, min(args.steps, 3)))
    except Exception as e:  # an installed but broken reference must not take the arm down
        return {"error": f"{type(e).__name__}: {e}"}


def run_reference(args):
    """Reference arm: the CPU restatement of the path (oracle port; the real reference is not installable)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n, d = args.particles, args.dim
    real = genjax_reference_rate(args)
    if real is not None and "value" in real:
        line = {
            "impl": "reference", "metric": "particle-steps/sec", "value": real["value"], "unit": "particle-steps/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * real["seconds"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_string(n, args.T, d), "particles_per_gpu": n, "T": args.T, "d": d},
            "cpu_baseline": {"value": real["value"], "unit": "particle-steps/s", "cores": real["cores"], "kind": "reference",
                             "sample": f"whole {args.T}-step filter, GenJAX {real['genjax']} on jax {real['jax']} [cpu], "
                                       "baseline/run_genjax_cpu.py"},
            "e2e": {"value": real["value"], "unit": "particle-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
        }
        print(json.dumps(line), flush=True)
        return
    use_c = cpu_c_port_rate(min(n, 4096), 1, d) is not None
    T_sample = args.T if use_c else 4  # the C port runs the whole T-step filter per timed step; NumPy a 4-step sample
    port = cpu_c_port_rate if use_c else cpu_port_rate
    impl_name = ("performance C/OpenMP restatement of the filter (oracle/c/pf_port_fast.c: one Philox block per quad, float libm, "
                 "-O3 -march=native)" if use_c else "NumPy float32 oracle port")
    for _ in range(min(args.warmup, 1)):
        port(n, 1, d)
    rates, times = [], []
    threads = os.cpu_count() or 1
    for _ in range(args.steps):
        r, dt, _, threads = port(n, T_sample, d)
        rates.append(r)
        times.append(dt)
    total = n * T_sample * args.steps
    value = total / sum(times)
    line = {
        "impl": "reference",
        "metric": "particle-steps/sec",
        "value": value,
        "unit": "particle-steps/s",
        "n_gpus": args.gpus,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": 1e3 * float(np.mean(times)),
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": workload_string(n, args.T, d), "particles_per_gpu": n, "T": args.T, "d": d,
                   "sample": f"{T_sample} of {args.T} filter steps per timed step"},
        "cpu_baseline": {"value": value, "unit": "particle-steps/s", "cores": threads, "kind": "port",
                         "sample": f"{T_sample} filter steps x {n} particles per timed step, {impl_name}, {threads} threads "
                                   f"({os.cpu_count()} cores visible); the reference itself (GenJAX on jax[cpu]) is not installable "
                                   "in this image"},
        "e2e": {"value": value, "unit": "particle-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------- GPU legs


def run_ours(args):
    import torch
    import torch.distributed as dist

    import genjax_b200 as gj
    from genjax_b200.inference.pf import ParticleFilter
    from genjax_b200.runtime import cabi
    from genjax_b200.workloads import LG_Q, LG_R, lgssm_step, lgssm_step_vec

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)

    n, d, T = args.particles, args.dim, args.T
    r_sd = obs_sd(args, d)
    ys_np = synth_obs(T, d, r=r_sd)
    g = np.random.default_rng(1 + rank)
    x0_np = g.standard_normal(n if d == 1 else (n, d)).astype(np.float32)
    # host buffers (pinned) for the e2e leg, device-resident copies for `value`
    x0_host = torch.from_numpy(x0_np).pin_memory()
    ys_host = torch.from_numpy(ys_np).pin_memory()
    x0_dev = x0_host.to(device)
    ys_dev = ys_host.to(device)
    state_host = torch.empty_like(x0_host).pin_memory()
    if d == 1:
        model, shared = lgssm_step, ()
    else:
        model = lgssm_step_vec
        shared = (torch.full((d,), LG_Q, device=device), torch.full((d,), r_sd, device=device))
    # weak scaling: every rank filters its own block of n particles; lanes are
    # global particle indices so the streams of different ranks never overlap
    global_resample = world > 1 and args.multi_gpu == "global"
    if global_resample:
        from genjax_b200.inference.pf_dist import DistributedParticleFilter

        pf = DistributedParticleFilter(model, n)
    else:
        pf = ParticleFilter(model, n, idx_offset=0, mode=args.mode, reference_max=args.reference_max,
                            single_pass=args.single_pass)
    obs_dev = gj.C["y"].set(ys_dev)
    obs_host = gj.C["y"].set(ys_host)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)  # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(device)

    log_world = float(np.log(world))

    gathered = torch.empty(max(world, 1), dtype=torch.float64, device=device)

    def combine(logz, is_global):
        """islands: log-mean-exp over ranks of the per-shard estimates from ONE collective (an all-gather of 8 bytes per
        rank) -- the single NCCL reduction of per-shard log-sum-exp terms of the north star."""
        if world == 1 or is_global:
            return logz
        dist.all_gather_into_tensor(gathered, logz.detach().reshape(1))
        return torch.logsumexp(gathered, 0) - log_world

    def one(step_idx, e2e, pf_=None, is_global=None):
        pf_ = pf if pf_ is None else pf_
        is_global = global_resample if is_global is None else is_global
        key = gj.fold_in(gj.key(314159 + (0 if is_global else rank)), step_idx)
        if e2e:
            res = pf_.run(key, x0_host.to(device, non_blocking=True), obs_host, shared_args=shared, use_graph=not args.eager)
            state_host.copy_(res.state[0], non_blocking=True)  # the filter's product: final particle cloud, pinned D2H
            return combine(res.log_marginal_likelihood, is_global).item()  # (all-gather of the shard terms +) D2H read + sync
        res = pf_.run(key, x0_dev, obs_dev, shared_args=shared, use_graph=not args.eager)
        res.combined = combine(res.log_marginal_likelihood, is_global)
        return res

    # ---- warm-up
    for w in range(max(args.warmup, 3)):
        one(w, False)
    torch.cuda.synchronize(device)
    for w in range(2):
        one(w, True)

    sampler = ClockSampler(local)
    sampler.start()

    # ---- device-resident timing: one event pair per step, L2 flushed between steps
    barrier()
    times = []
    for k in range(args.steps):
        flush.fill_(k & 0xFF)
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        # every timed step starts from a barrier + synchronize on all ranks: with global resampling a rank's step cannot
        # finish before the slowest rank has STARTED it, so without the barrier the host-side skew between the ranks'
        # Python loops (L2 flush, synchronize, launch) would be counted as device time of the step
        barrier()
        e0.record()
        res = one(100 + k, False)
        e1.record()
        torch.cuda.synchronize(device)
        times.append(e0.elapsed_time(e1))
    barrier()
    logz = res.combined.item()
    t_dev = torch.tensor([sum(times)], dtype=torch.float64, device=device)

    # ---- end to end through the public API with host buffers
    barrier()
    t0 = time.perf_counter()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(args.steps):
        logz_e2e = one(200 + k, True)
    e1.record()
    torch.cuda.synchronize(device)
    t_e2e_ms = max(e0.elapsed_time(e1), 1e3 * (time.perf_counter() - t0))
    t_e2e = torch.tensor([t_e2e_ms], dtype=torch.float64, device=device)
    clocks = sampler.stop()

    # ---- N > 1, global default: the islands mode beside it (device-resident timing only, same protocol)
    islands = None
    if global_resample:
        pf_i = ParticleFilter(model, n, idx_offset=0, mode=args.mode)
        for w in range(3):
            one(w, False, pf_i, False)
        barrier()
        ti = []
        for k in range(args.steps):
            flush.fill_(k & 0xFF)
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            barrier()  # (the all-gather of the shard terms couples the ranks: same reason as above)
            e0.record()
            r_i = one(300 + k, False, pf_i, False)
            e1.record()
            torch.cuda.synchronize(device)
            ti.append(e0.elapsed_time(e1))
        barrier()
        t_isl = torch.tensor([sum(ti)], dtype=torch.float64, device=device)
        dist.all_reduce(t_isl, op=dist.ReduceOp.MAX)
        islands = {"value": float(n) * T * args.steps * world / (t_isl.item() * 1e-3), "unit": "particle-steps/s",
                   "ms_per_step": t_isl.item() / args.steps, "logZ_last": r_i.combined.item(),
                   "what": "every rank filters its own block (local resampling); the shard log-marginal-likelihood terms are combined "
                           "by ONE NCCL all-gather per run, inside the timed region"}

    if world > 1:
        dist.all_reduce(t_dev, op=dist.ReduceOp.MAX)
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    total_units = float(n) * T * args.steps * world
    value = total_units / (t_dev.item() * 1e-3)
    e2e_value = total_units / (t_e2e.item() * 1e-3)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel
    plan = next(iter(pf._plans.values()))
    import ctypes as C

    stream = cabi.stream_ptr(device)
    peak, how = peaks()
    ms_per_step = t_dev.item() / args.steps
    step_bytes = (8 * d + 24) * n * T  # whole bootstrap-PF step with fused gather

    def time_launches(fn, reps, warm):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize(device)
        k0 = torch.cuda.Event(enable_timing=True)
        k1 = torch.cuda.Event(enable_timing=True)
        k0.record()
        for _ in range(reps):
            fn()
        k1.record()
        torch.cuda.synchronize(device)
        return k0.elapsed_time(k1) / reps

    # (i) the stand-alone fused gather + propose + logpdf launch (gjb_model_launch), as ImportanceK / ChangeTarget use it
    MA = cabi.ModelArgs()
    MA.n = n
    MA.key0, MA.key1 = 1, 2
    anc_id = torch.arange(n, dtype=torch.int32, device=device)
    x_out = torch.empty_like(x0_dev)
    lw_out = torch.empty(n, dtype=torch.float32, device=device)
    wmax = torch.zeros(1, dtype=torch.int32, device=device)
    y_one = ys_dev[0].contiguous() if d > 1 else ys_dev[:1].contiguous()
    MA.args[0] = x0_dev.data_ptr()
    for k_, s_ in enumerate(shared):
        MA.args[1 + k_] = s_.data_ptr()
    MA.gather = anc_id.data_ptr()
    MA.site_flags[0] = cabi.SITE_SAMPLE
    MA.site_out[0] = x_out.data_ptr()
    MA.site_in[1] = y_one.data_ptr()
    MA.site_flags[1] = cabi.SITE_WEIGHT | cabi.SITE_BCAST
    MA.weight_out = lw_out.data_ptr()
    MA.wmax = wmax.data_ptr()
    model_ms = time_launches(lambda: plan.cm.lib.gjb_model_launch(C.byref(MA), stream), 200, 20)
    model_bytes = (8 * d + 12) * n  # read ancestor 4 + x_prev 4d, write x 4d + logw 4
    model_gbs = model_bytes / (model_ms * 1e-3) / 1e9

    kernels = {"model_kernel": {"kernel_us": model_ms * 1e3, "algorithmic_bytes_per_launch": model_bytes,
                                "achieved": model_gbs, "frac": model_gbs / peak,
                                "what": "fused ancestor-gather + propose + logpdf + running max (gjb_model_launch)"}}
    if global_resample or getattr(plan, "te_table", False):
        # table form (several devices): a step kernel waits for the table kernel of the step before, fed by every rank's tile
        # records, so it cannot be re-launched in isolation; the roofline is taken from the step time of the timed region
        # itself (max over ranks)
        st_ms = ms_per_step / T
        st_bytes = (8 * d + 24) * n
        st_gbs = st_bytes / (st_ms * 1e-3) / 1e9
        kernels["pf_step_kernel"] = {
            "kernel_us": st_ms * 1e3, "algorithmic_bytes_per_launch": st_bytes, "achieved": st_gbs, "frac": st_gbs / peak,
            "what": "one filter step per GPU in one launch (global resampling: parent CDF rows / states read over NVLink, tile "
                    "records mailed to every rank, prefix table built by each rank's last CTA); duration = timed region / T"}
        roofline = {
            "bound": "hbm", "kernel": "pf_step_kernel: " + kernels["pf_step_kernel"]["what"],
            "achieved": st_gbs, "peak": peak, "peak_source": how, "unit": "GB/s", "frac": st_gbs / peak, "traffic": None,
            "kernel_us": st_ms * 1e3, "algorithmic_bytes_per_launch": st_bytes, "algorithmic_bytes_per_particle_step": 8 * d + 24,
            "note": "per GPU; algorithmic bytes = SURVEY 8d's whole-step figure (8d + 24 per particle-step)",
        }
    elif getattr(plan, "stepsmode", False):
        # ALL steps are one cooperative launch: its duration / T is the step
        ss_ms = time_launches(lambda: plan.cm.lib.gjb_model_pf_steps(C.byref(plan.steps_args), stream), 10, 2) / T
        st_bytes = (8 * d + 24) * n
        st_gbs = st_bytes / (ss_ms * 1e-3) / 1e9
        kernels["pf_steps_kernel"] = {
            "kernel_us": ss_ms * 1e3, "algorithmic_bytes_per_launch": st_bytes * T, "achieved": st_gbs, "frac": st_gbs / peak,
            "what": "all T filter steps in ONE cooperative launch (one grid barrier per step): per step output-slot systematic "
                    "resampling of the previous step + ancestor gather + propose + logpdf + within-tile CDF / tile record "
                    "(gjb_model_pf_steps); kernel_us = launch duration / T"}
        roofline = {
            "bound": "hbm", "kernel": "pf_steps_kernel: " + kernels["pf_steps_kernel"]["what"],
            "achieved": st_gbs, "peak": peak, "peak_source": how, "unit": "GB/s", "frac": st_gbs / peak, "traffic": None,
            "kernel_us": ss_ms * 1e3, "algorithmic_bytes_per_launch": st_bytes * T, "algorithmic_bytes_per_particle_step": 8 * d + 24,
            "note": "algorithmic bytes = SURVEY 8d's whole-step figure (8d + 24 per particle-step) x T steps per launch",
        }
    elif getattr(plan, "stepmode", False):
        # the step IS one kernel: launch t = 1 (reads step 0's CDF / tile records, writes its own) is what every later
        # step repeats; re-launching it is idempotent
        plan.cm.lib.gjb_model_pf_step(C.byref(plan.sargs[0]), stream)
        st_ms = time_launches(lambda: plan.cm.lib.gjb_model_pf_step(C.byref(plan.sargs[1]), stream), 200, 20)
        st_bytes = (8 * d + 24) * n
        st_gbs = st_bytes / (st_ms * 1e-3) / 1e9
        kernels["pf_step_kernel"] = {
            "kernel_us": st_ms * 1e3, "algorithmic_bytes_per_launch": st_bytes, "achieved": st_gbs, "frac": st_gbs / peak,
            "what": "the whole filter step in one launch: output-slot systematic resampling of the previous step (tile-exponent "
                    "CDF) + ancestor gather + propose + logpdf + within-tile CDF / tile record of the new weights (gjb_model_pf_step)"}
        roofline = {
            "bound": "hbm", "kernel": "pf_step_kernel: " + kernels["pf_step_kernel"]["what"],
            "achieved": st_gbs, "peak": peak, "peak_source": how, "unit": "GB/s", "frac": st_gbs / peak, "traffic": None,
            "kernel_us": st_ms * 1e3, "algorithmic_bytes_per_launch": st_bytes, "algorithmic_bytes_per_particle_step": 8 * d + 24,
            "note": "algorithmic bytes = SURVEY 8d's whole-step figure (8d + 24 per particle-step): this kernel is the whole step",
        }
    elif getattr(plan, "persistent", False):
        pf_ms = time_launches(lambda: plan.cm.lib.gjb_model_pf_run(C.byref(plan.pf_args), stream), 10, 2)
        achieved = step_bytes / (pf_ms * 1e-3) / 1e9
        roofline = {
            "bound": "hbm",
            "kernel": "pf_kernel (persistent cooperative filter: per step gather+propose+logpdf+max | integer mass | CDF scan+systematic ancestors)",
            "achieved": achieved, "peak": peak, "peak_source": how, "unit": "GB/s", "frac": achieved / peak,
            "traffic": None, "kernel_us": pf_ms * 1e3, "algorithmic_bytes_per_launch": step_bytes,
            "algorithmic_bytes_per_particle_step": 8 * d + 24,
        }
    else:
        dominant = "model_kernel"
        if not global_resample and getattr(plan, "fuse_mass_resample", False):
            # the other kernel of the step: integer mass + CDF scan + systematic ancestors in one cooperative launch;
            # algorithmic bytes: log-weights read once (4) + ancestors written (4) per particle
            core = cabi.core()
            core.gjb_wmax_reset(plan.wmax2.data_ptr(), stream)
            plan.cm.lib.gjb_model_launch(C.byref(plan.margs[0]), stream)  # a valid max + weights to resample
            mr_ms = time_launches(lambda: core.gjb_mass_resample_systematic(C.byref(plan.rargs[0][1]), stream), 200, 20)
            mr_bytes = 8 * n
            mr_gbs = mr_bytes / (mr_ms * 1e-3) / 1e9
            kernels["mass_resample_kernel"] = {
                "kernel_us": mr_ms * 1e3, "algorithmic_bytes_per_launch": mr_bytes, "achieved": mr_gbs, "frac": mr_gbs / peak,
                "what": "exact integer weight mass + grid barrier + CDF scan + systematic offspring ranges (gjb_mass_resample_systematic)"}
            if mr_ms > model_ms:
                dominant = "mass_resample_kernel"
        if not global_resample and getattr(plan, "single_pass", False):
            # single-pass step: step 0 leaves weights + tile masses; step 1's launch (pull-resample them into the CTA's
            # own slots, gather, propose, score, masses) is the kernel every later step repeats.  Re-launching it adds
            # to the same tile masses again, which nothing reads here.
            plan.tm3.zero_()
            plan.cm.lib.gjb_model_launch(C.byref(plan.margs[0]), stream)
            sp_ms = time_launches(lambda: plan.cm.lib.gjb_model_launch(C.byref(plan.margs[1]), stream), 200, 20)
            sp_bytes = model_bytes + 4 * n  # + the previous step's log-weights; the ancestors are written AND read back
            kernels["model_kernel_static_pull"] = {
                "kernel_us": sp_ms * 1e3, "algorithmic_bytes_per_launch": sp_bytes, "achieved": sp_bytes / (sp_ms * 1e-3) / 1e9,
                "frac": sp_bytes / (sp_ms * 1e-3) / 1e9 / peak,
                "what": "output-slot systematic resampling of the previous step + gather + propose + logpdf + exact integer "
                        "masses relative to the analytic bound, one launch per filter step (gjb_model_launch)"}
            dominant = "model_kernel_static_pull"
        elif not global_resample and getattr(plan, "analytic", False):
            # reference-maximum step: the model kernel also accumulates the masses; the resampler takes them as given
            core = cabi.core()
            plan.tm2.zero_()
            mm_ms = time_launches(lambda: plan.cm.lib.gjb_model_launch(C.byref(plan.margs[0]), stream), 200, 20)
            kernels["model_kernel_static_mass"] = {
                "kernel_us": mm_ms * 1e3, "algorithmic_bytes_per_launch": model_bytes, "achieved": model_bytes / (mm_ms * 1e-3) / 1e9,
                "frac": model_bytes / (mm_ms * 1e-3) / 1e9 / peak,
                "what": "gather + propose + logpdf + exact integer masses relative to the analytic bound (gjb_model_launch)"}
            plan.tm2.zero_()
            plan.cm.lib.gjb_model_launch(C.byref(plan.margs[0]), stream)
            rs_ms = time_launches(lambda: core.gjb_resample_systematic(C.byref(plan.rargs[0][1]), stream), 200, 20)
            kernels["resample_systematic_kernel"] = {
                "kernel_us": rs_ms * 1e3, "algorithmic_bytes_per_launch": 8 * n, "achieved": 8 * n / (rs_ms * 1e-3) / 1e9,
                "frac": 8 * n / (rs_ms * 1e-3) / 1e9 / peak,
                "what": "CDF scan + systematic offspring ranges on given tile masses (gjb_resample_systematic)"}
            dominant = "resample_systematic_kernel" if rs_ms > mm_ms else "model_kernel_static_mass"
        k = kernels[dominant]
        roofline = {
            "bound": "hbm", "kernel": f"{dominant}: {k['what']}",
            "achieved": k["achieved"], "peak": peak, "peak_source": how, "unit": "GB/s", "frac": k["frac"],
            "traffic": None, "kernel_us": k["kernel_us"], "algorithmic_bytes_per_launch": k["algorithmic_bytes_per_launch"],
            "note": "dominant = the kernel with the larger measured launch duration of the step; both kernels are listed under "
                    "'kernels'. At d=1 neither is HBM-bound (working set L2-resident, instruction / barrier-latency bound, see DESIGN.md 6)",
        }
    roofline["kernels"] = kernels
    # DRAM traffic of the dominant kernel per launch from the committed `ncu --set full` capture (profiles/)
    try:
        with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
            tr = json.load(f).get(f"d{d}", {})
        # keyed by the kernel the roofline names: a capture of another kernel never stands in (round 1's constant went stale)
        kname = roofline["kernel"].split(":")[0].split(" ")[0]
        roofline["traffic"] = tr.get(kname + "_dram_bytes_per_launch")
        roofline["traffic_source"] = tr.get(kname + "_source", tr.get("source")) if roofline["traffic"] is not None else \
            f"no committed ncu capture of {kname} at d={d}"
        if roofline["traffic"] is not None and global_resample:
            roofline["traffic_source"] += " (single-device capture)"
    except Exception:
        pass
    roofline["working_set_note"] = (
        "L2-resident: %d MB of particle arrays < 126 MB L2, DRAM traffic is only the first touch" % ((8 * d + 12) * n >> 20)
        if (8 * d + 12) * n < (100 << 20) else "HBM-resident: particle arrays exceed the 126 MB L2")
    roofline["whole_step_GBps"] = step_bytes / (ms_per_step * 1e-3) / 1e9
    roofline["model_kernel_alone"] = {"kernel_us": model_ms * 1e3, "algorithmic_bytes_per_launch": model_bytes,
                                      "achieved": model_gbs, "frac": model_gbs / peak}

    cpu = None
    if not args.no_cpu_baseline and world == 1:  # the CPU arm is timed at N = 1 only
        T_s = 20 if d == 1 else 2
        c_res = cpu_c_port_rate(n, T, d)
        exact_res = None
        if c_res is not None:
            cpu_c_port_rate(n, 2, d)  # (first touch of the buffers / thread pool outside the timed sample)
            c_res = cpu_c_port_rate(n, T, d)
            rate, dt, cpu_logz, thr = c_res
            what = (f"all {T} filter steps x {n} particles in {dt:.2f} s, performance C/OpenMP restatement of the filter "
                    "(oracle/c/pf_port_fast.c: one Philox block per quad, float libm, -O3 -march=native)")
            exact_res = cpu_c_port_rate(n, max(T // 5, 1), d, fast=False)
        else:
            rate, dt, _, thr = cpu_port_rate(n, T_s, d)
            what = f"{T_s} of {T} filter steps x {n} particles in {dt:.1f} s, NumPy float32 oracle port"
        cpu = {"value": rate, "unit": "particle-steps/s", "cores": thr, "kind": "port",
               "sample": f"{what}, {thr} threads ({os.cpu_count()} cores visible); GenJAX jax[cpu] itself is not installable here"}
        if exact_res is not None:
            cpu["bit_compatible_port"] = {"value": exact_res[0], "unit": "particle-steps/s", "cores": exact_res[3],
                                          "what": "oracle/c/pf_port.c, the operation-for-operation restatement the parity tests use "
                                                  f"({max(T // 5, 1)} filter steps)"}

    # the estimate of the last timed run against the exact Kalman log-likelihood (sd of one d = 1 run measured at N = 2^18,
    # T = 50: 0.029, tests/test_pf_gpu.py; variance scales with T / N; R island estimates average)
    logz_exact = kalman_logz(ys_np, d, r_sd)
    logz_sigma = 0.029 * float(np.sqrt((T / 50.0) * ((1 << 18) / float(n)) / world)) if d == 1 else None
    if logz_sigma is None:
        logz_check = f"not gated at d > 1 (no sigma on file); estimate - exact = {logz - logz_exact:+.3f}"
    else:
        ok = abs(logz - logz_exact) <= 4 * logz_sigma and abs(logz_e2e - logz_exact) <= 4 * logz_sigma
        logz_check = "pass: |logZ - exact| <= 4 sigma (device-resident and e2e runs)" if ok else \
            f"FAIL: logZ {logz} / {logz_e2e} vs exact {logz_exact}, 4 sigma = {4 * logz_sigma}"
    h2d = x0_host.numel() * 4 + ys_host.numel() * 4 + T * 8 * 4
    line = {
        "metric": "particle-steps/sec",
        "value": value,
        "unit": "particle-steps/s",
        "n_gpus": world,
        "steps": args.steps,
        "warmup": max(args.warmup, 3),
        "ms_per_step": ms_per_step,
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "f32",
        "data": "synthetic",
        "config": {
            "workload": workload_string(n, T, d),
            "particles_per_gpu": n, "T": T, "d": d, "obs_sd": r_sd,
            "l2": "flushed between timed steps (256 MiB write)",
            "mode": "step (1 launch/step/GPU, one cross-rank hand-off per step inside the kernel)" if global_resample else args.mode,
            "multi_gpu": ("one filter over all ranks' particles, global systematic resampling: per step every CTA mails its tile record "
                          "to every rank, each rank's last CTA waits for all of them (the step's one hand-off) and builds the prefix "
                          "table; parent CDF rows / states are read over NVLink peer memory; weak scaling"
                          if global_resample else "islands: each rank filters its own block (local resampling, global RNG lanes "
                          "differ by key), the shard log-marginal-likelihood terms are combined by one NCCL all-gather (8 bytes per "
                          "rank) per run, inside the timed region; weak scaling. --multi-gpu global (the default for N > 1) times the "
                          "global-resampling filter"),
            "logZ_last": logz, "logZ_exact_kalman": logz_exact, "logZ_sigma": logz_sigma, "logZ_check": logz_check,
            "reference_max": args.reference_max, "single_pass": bool(args.single_pass),
        },
        "e2e": {"value": e2e_value, "unit": "particle-steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 8 + state_host.numel() * 4,
                "logZ_last": logz_e2e, "d2h": "log-marginal-likelihood estimate (8 B) + the final particle state"},
        "gpu_launches": plan.launches_per_run() * args.steps,
        "islands": islands,
        "roofline": roofline,
        "cpu_baseline": cpu,
        "clocks": clocks,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    if logz_check.startswith("FAIL"):
        raise SystemExit("bench.py: the filter's log-marginal-likelihood estimate is outside 4 sigma of the exact value: " + logz_check)


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

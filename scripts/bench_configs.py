#!/usr/bin/env python
"""The other BASELINE.json configs at full size (they are parity-test cases, not the bench line):
    python scripts/bench_configs.py mh      # configs[2]: 8-comp 8-D GMM, 256K-chain MH, 1000 steps, 1 GPU
    python scripts/bench_configs.py hmc     # configs[4] per-GPU share: 8-schools HMC, 8K chains x 200 x L=10 (64K over 8 GPUs)
    torchrun --nproc-per-node 8 scripts/bench_configs.py hmm   # configs[3]: 16-state HMM PF, T=1000, 4M particles, global resample
Each prints one JSON line (CUDA-event timed, after a warm-up run)."""
import json
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import torch

import genjax_b200 as gj


def _time(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out


def _shard():
    """(world, rank, device): chains are sharded by global index; no collective on the data path."""
    import torch.distributed as dist

    world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0"))))
    return world, rank


def _fingerprint(t):
    """sha1 of a tensor's bytes: chain i's result depends on its GLOBAL index only, so the same chains give the same
    fingerprint on 1 and on R GPUs (R-independence check)."""
    import hashlib

    return hashlib.sha1(t.detach().cpu().numpy().tobytes()).hexdigest()[:16]


def _finish(world, ms):
    import torch.distributed as dist

    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()


def mh():
    from genjax_b200.inference.mcmc import mh_chain
    from genjax_b200.workloads import gmm_target

    world, rank = _shard()
    K, D, n_total, steps = 8, 8, 262_144, 1000
    n = n_total // world
    g = np.random.default_rng(1)
    mu = torch.from_numpy(g.uniform(-4, 4, size=(K, D)).astype(np.float32))
    args = (torch.zeros(K), mu, torch.full((K,), 0.7))
    lanes = slice(rank * n, (rank + 1) * n)  # this rank's chains = lanes of the GLOBAL key batches
    tr = gmm_target.simulate(gj.split(gj.key(2), n_total)[lanes], args)
    kb = gj.split(gj.key(3), n_total)[lanes]
    ms, res = _time(lambda: mh_chain(kb, tr, gj.S["x"], step_size=0.5, n_steps=steps, rebuild_trace=False))
    ms = _finish(world, ms)
    x = res.trace.state.cpu().numpy()
    occ = np.bincount(((x[:, None, :] - mu.numpy()[None]) ** 2).sum(-1).argmin(1), minlength=K) / n
    if rank == 0:
        print(json.dumps({"config": "configs[2] GMM MH", "n_gpus": world, "chains": n_total, "steps": steps, "ms": ms,
                          "chain_steps_per_s": n_total * steps / (ms * 1e-3), "accept_rate": float(res.accept_rate),
                          "component_occupancy_rank0": [round(float(o), 4) for o in occ],
                          "fingerprint_chains_0_1023": _fingerprint(res.trace.state[:1024]),
                          "hbm_bytes_per_launch": 2 * n * (4 * D + 4), "note": "state in registers; 1 launch = 1000 transitions"}))


def hmc():
    from genjax_b200.inference.mcmc import hmc_chain
    from genjax_b200.workloads import EIGHT_SCHOOLS_SIGMA, EIGHT_SCHOOLS_Y, eight_schools

    world, rank = _shard()
    n_total = int(os.environ.get("GJB_HMC_CHAINS", "65536" if world > 1 else "8192"))  # configs[4]: 64 K chains over 8 GPUs
    n, iters, L = n_total // world, 200, 10
    lanes = slice(rank * n, (rank + 1) * n)
    y, sig = torch.tensor(EIGHT_SCHOOLS_Y), torch.tensor(EIGHT_SCHOOLS_SIGMA)
    tr, _ = eight_schools.importance(gj.split(gj.key(4), n_total)[lanes], gj.C["y"].set(y), (sig,))
    sel = gj.S["mu"] | gj.S["log_tau"] | gj.S["theta"]
    kb = gj.split(gj.key(5), n_total)[lanes]
    out = {}
    for compat in (False, True):
        ms, res = _time(lambda: hmc_chain(kb, tr, sel, eps=0.05, L=L, n_iters=iters, compat_stale_grad=compat, rebuild_trace=False))
        ms = _finish(world, ms)
        st = res.trace.state
        out["compat_hmc_py_186" if compat else "textbook"] = {
            "ms": ms, "leapfrogs_per_s": n_total * iters * L / (ms * 1e-3), "accept_rate": float(res.accept_rate),
            "mean_mu_rank0": float(st[:, 0].mean()), "mean_log_tau_rank0": float(st[:, 1].mean()),
            "fingerprint_chains_0_1023": _fingerprint(st[:1024])}
    if rank == 0:
        print(json.dumps({"config": "configs[4] 8-schools HMC", "n_gpus": world, "chains": n_total,
                          "iters": iters, "L": L, "eps": 0.05, **out}))


def hmm():
    import torch.distributed as dist

    from genjax_b200.inference.pf import ParticleFilter
    from genjax_b200.inference.pf_dist import DistributedParticleFilter
    from genjax_b200.workloads import hmm_step

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    K, T, n_total = 16, 1000, 4_194_304
    n = n_total // world
    i = np.arange(K)
    d = np.minimum((i[:, None] - i[None, :]) % K, (i[None, :] - i[:, None]) % K).astype(np.float64)
    trans = (-0.5 * (d / 0.5) ** 2).astype(np.float32)
    obs = (-0.5 * (d / 0.5) ** 2).astype(np.float32)
    pt = np.exp(trans - trans.max(1, keepdims=True)); pt /= pt.sum(1, keepdims=True)
    po = np.exp(obs - obs.max(1, keepdims=True)); po /= po.sum(1, keepdims=True)
    g = np.random.default_rng(3)
    z, ys = 0, np.empty(T, dtype=np.int32)
    for t in range(T):
        z = g.choice(K, p=pt[z]); ys[t] = g.choice(K, p=po[z])
    z0_all = np.random.default_rng(4).integers(0, K, n_total).astype(np.int32)
    z0 = torch.from_numpy(z0_all[rank * n:(rank + 1) * n]).to(dev)
    shared = (torch.from_numpy(trans).to(dev), torch.from_numpy(obs).to(dev))
    obs_chm = gj.C["y"].set(torch.from_numpy(ys).to(dev))
    pf = DistributedParticleFilter(hmm_step, n) if world > 1 else ParticleFilter(hmm_step, n)  # both: the single-launch step kernel
    ms, res = _time(lambda: pf.run(gj.key(11), z0, obs_chm, shared_args=shared), reps=2)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    alpha = np.bincount(z0_all, minlength=K) / n_total
    ll = 0.0
    for k in range(T):
        alpha = (alpha @ pt) * po[:, ys[k]]; s = alpha.sum(); ll += math.log(s); alpha /= s
    if rank == 0:
        print(json.dumps({"config": "configs[3] 16-state HMM bootstrap PF, global systematic resample every step", "n_gpus": world,
                          "particles": n_total, "T": T, "ms": t.item(), "particle_steps_per_s": n_total * T / (t.item() * 1e-3),
                          "logZ": float(res.log_marginal_likelihood), "exact_forward_logZ": ll, "state_dtype": "int32"}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    {"mh": mh, "hmc": hmc, "hmm": hmm}[sys.argv[1]]()

"""Worker of the multi-GPU particle-filter test: launched with
``python -m torch.distributed.run --nproc-per-node R tests/dist_pf_worker.py``.
Checks that the R-rank filter with global resampling equals the single-GPU
filter of R*n particles BIT FOR BIT (states, ancestors, log-weights, logZ)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch
import torch.distributed as dist

import genjax_b200 as gj
from genjax_b200.inference.pf import ParticleFilter
from genjax_b200.inference.pf_dist import DistributedParticleFilter
from genjax_b200.workloads import LG_Q, LG_R, lgssm_step, lgssm_step_vec


def main():
    rank = int(os.environ["RANK"])
    world = int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    ndev = torch.cuda.device_count()
    same_device = ndev < world
    dev = torch.device("cuda", local % ndev)
    torch.cuda.set_device(dev)
    backend = "gloo" if same_device else "nccl"
    dist.init_process_group(backend, device_id=None if same_device else dev)
    n = int(os.environ.get("GJB_TEST_N", "50000"))
    n -= n % 4
    T = 7
    g = np.random.default_rng(5)
    for d in (1, 8):
        ys = g.standard_normal((T, d) if d > 1 else T).astype(np.float32)
        x0 = g.standard_normal((world * n, d) if d > 1 else world * n).astype(np.float32)
        if d == 1:
            model, shared = lgssm_step, ()
        else:
            model, shared = lgssm_step_vec, (torch.full((d,), LG_Q), torch.full((d,), LG_R))
        obs = gj.C["y"].set(torch.from_numpy(ys))
        mine = torch.from_numpy(x0[rank * n:(rank + 1) * n])
        # ---- the single-launch step kernel (default): R ranks == one device with R * n particles, bit for bit
        ns = max(2048, (n // 2048) * 2048)
        x0s = g.standard_normal((world * ns, d) if d > 1 else world * ns).astype(np.float32)
        mine_s = torch.from_numpy(x0s[rank * ns:(rank + 1) * ns])
        ref = ParticleFilter(model, world * ns, mode="step").run(gj.key(21), torch.from_numpy(x0s), obs, shared_args=shared,
                                                                 record=True, use_graph=False)
        torch.cuda.synchronize()
        lo, hi = rank * ns, (rank + 1) * ns
        for use_graph in (False, True):
            dpf = DistributedParticleFilter(model, ns, mode="step")
            runs = [dpf.run(gj.key(21), mine_s, obs, shared_args=shared, record=True, use_graph=use_graph) for _ in range(2)]  # (replay: tags advance)
            torch.cuda.synchronize()
            for r in runs:
                assert torch.equal(r.ancestors, ref.ancestors[:, lo:hi]), f"step: ancestors differ (d={d}, rank={rank})"
                assert torch.equal(r.history["log_weights"], ref.history["log_weights"][:, lo:hi]), "step: log-weights differ"
                assert torch.equal(r.history["state"][0], ref.history["state"][0][:, lo:hi]), "step: states differ"
                assert torch.equal(r.log_increments, ref.log_increments), "step: logZ increments differ"
                assert torch.equal(r.state[0], ref.state[0][lo:hi]), "step: final state differs"
            nr = DistributedParticleFilter(model, ns, mode="step").run(gj.key(21), mine_s, obs, shared_args=shared, use_graph=use_graph)
            assert torch.equal(nr.log_increments, ref.log_increments) and torch.equal(nr.state[0], ref.state[0][lo:hi]), "step: non-record run differs"
            a = runs[0].ancestors[-1]
            print(f"[rank {rank}] d={d} graph={use_graph} mode=step: OK, logZ={runs[0].log_marginal_likelihood.item():.4f}, "
                  f"{int(((a < lo) | (a >= hi)).sum())} of {ns} last-step ancestors remote", flush=True)
            del dpf
        if os.environ.get("GJB_TEST_STEP_ONLY") == "1":
            continue
        for use_graph, fused, mode in ((False, True, "pull"), (True, True, "pull"), (True, True, "push"), (True, False, "push")):
            dpf = DistributedParticleFilter(model, n, fused=fused, mode=mode)
            res = dpf.run(gj.key(21), mine, obs, shared_args=shared, record=True, use_graph=use_graph)
            res2 = dpf.run(gj.key(21), mine, obs, shared_args=shared, record=True, use_graph=use_graph)  # replay: epoch tags advance
            torch.cuda.synchronize()
            ref = ParticleFilter(model, world * n, mode="graph").run(gj.key(21), torch.from_numpy(x0), obs, shared_args=shared,
                                                                      record=True, use_graph=False)
            torch.cuda.synchronize()
            lo, hi = rank * n, (rank + 1) * n
            for r in (res, res2):
                assert torch.equal(r.ancestors, ref.ancestors[:, lo:hi]), f"ancestors differ (d={d}, rank={rank})"
                assert torch.equal(r.history["log_weights"], ref.history["log_weights"][:, lo:hi]), "log-weights differ"
                assert torch.equal(r.history["state"][0], ref.history["state"][0][:, lo:hi]), "states differ"
                assert torch.equal(r.log_increments, ref.log_increments), "logZ increments differ"
                assert torch.equal(r.state[0], ref.state[0][lo:hi]), "final state differs"
            # cross-rank traffic really happened: some ancestors of my slots live on the other rank(s)
            a = res.ancestors[-1]
            remote = int(((a < lo) | (a >= hi)).sum())
            print(f"[rank {rank}] d={d} graph={use_graph} mode={mode} fused={fused}: OK, logZ={res.log_marginal_likelihood.item():.4f}, "
                  f"{remote} of {n} last-step ancestors remote", flush=True)
            del dpf
    dist.barrier()
    if rank == 0:
        print("DIST_PF_OK", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

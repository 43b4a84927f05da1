"""Per-kernel CUDA-event timing of the multi-GPU filter step on every rank (eager launches)."""
import os, sys
sys.path.insert(0, os.getcwd())
import ctypes as C
import numpy as np, torch, torch.distributed as dist
import genjax_b200 as gj
from genjax_b200.inference.pf_dist import DistributedParticleFilter
from genjax_b200.runtime import cabi
from genjax_b200.workloads import lgssm_step

rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n, T = 1 << 20, 40
g = np.random.default_rng(rank)
x0 = torch.from_numpy(g.standard_normal(n).astype(np.float32)).to(dev)
ys = torch.from_numpy(np.random.default_rng(0).standard_normal(T).astype(np.float32)).to(dev)
for mode in ("pull", "push"):
    pf = DistributedParticleFilter(lgssm_step, n, mode=mode)
    pf.run(gj.key(1), x0, gj.C["y"].set(ys), use_graph=False)
    torch.cuda.synchronize()
    plan = next(iter(pf._plans.values()))
    core = cabi.core(); lib = plan.cm.lib; stream = cabi.stream_ptr(dev)
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(T)]
    dist.barrier(); torch.cuda.synchronize()
    core.gjb_wmax_reset(plan.wmax2.data_ptr(), stream)
    for t, st in enumerate(plan.steps):
        ev[t][0].record()
        lib.gjb_model_launch(C.byref(st["A"]), stream)
        ev[t][1].record()
        if mode == "pull":
            core.gjb_weight_mass_prefix_linked(st["lw"].data_ptr(), st["lw"].numel(), plan.ws.tile_mass.data_ptr(), plan.tpre[t & 1].data_ptr(), plan.link.data_ptr(), 2*t+1, 2*t+2, stream)
            ev[t][2].record()
            core.gjb_resample_systematic_pull(C.byref(st["R"]), C.byref(st["lw_peers"]), C.byref(st["pre_peers"]), plan.link.data_ptr(), 2*t+1, 2*t+2, stream)
        else:
            core.gjb_weight_mass_linked(st["lw"].data_ptr(), st["lw"].numel(), plan.ws.tile_mass.data_ptr(), plan.link.data_ptr(), 3*t+1, 3*t+2, stream)
            ev[t][2].record()
            core.gjb_resample_systematic_linked(C.byref(st["R"]), C.byref(st["anc_peers"]), plan.link.data_ptr(), 3*t+1, 3*t+2, 3*t+3, stream)
        ev[t][3].record()
    if mode == "pull":
        core.gjb_exchange(C.byref(plan.x_end), stream)
    else:
        core.gjb_exchange(C.byref(plan.x_final), stream)
    core.gjb_epoch_bump(plan.epoch.data_ptr(), stream)
    torch.cuda.synchronize()
    m = np.array([[ev[t][k].elapsed_time(ev[t][k+1]) * 1e3 for k in range(3)] for t in range(5, T)])
    tot = ev[5][0].elapsed_time(ev[T-1][3]) * 1e3 / (T - 5)
    print(f"[rank {rank}] {mode}: model {m[:,0].mean():.1f} us  mass {m[:,1].mean():.1f} us  resample {m[:,2].mean():.1f} us  | step {tot:.1f} us (eager launches)", flush=True)
    del pf
dist.destroy_process_group()

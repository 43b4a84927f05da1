"""world=1 run of the distributed filter vs the plain filter: isolates symmetric-memory / link overhead."""
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch, torch.distributed as dist
import genjax_b200 as gj
from genjax_b200.inference.pf import ParticleFilter
from genjax_b200.inference.pf_dist import DistributedParticleFilter
from genjax_b200.workloads import lgssm_step
os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29555")
torch.cuda.set_device(0); dev = torch.device("cuda", 0)
n, T = 1 << 20, 100
x0 = torch.randn(n, device=dev); ys = torch.randn(T, device=dev)
def timeit(pf, reps=10):
    for _ in range(3): pf.run(gj.key(1), x0, gj.C["y"].set(ys))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): pf.run(gj.key(1), x0, gj.C["y"].set(ys))
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps / T * 1e3
plain = ParticleFilter(lgssm_step, n, mode="graph")
print("plain graph (before any process group)   us/step", round(timeit(plain), 2), flush=True)
dist.init_process_group("nccl", rank=0, world_size=1, device_id=dev)
print("plain graph (after nccl init)            us/step", round(timeit(plain), 2), flush=True)
dpf = DistributedParticleFilter(lgssm_step, n, mode="pull")
print("dist world=1 pull                        us/step", round(timeit(dpf), 2), flush=True)
print("plain graph (symmetric memory allocated) us/step", round(timeit(plain), 2), flush=True)
dpf2 = DistributedParticleFilter(lgssm_step, n, mode="push", fused=False)
print("dist world=1 push unfused                us/step", round(timeit(dpf2), 2), flush=True)
dist.destroy_process_group()

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
dist.init_process_group("nccl", rank=0, world_size=1, device_id=dev)
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
print("plain graph      us/step", round(timeit(ParticleFilter(lgssm_step, n, mode="graph")), 2))
for mode, fused in (("pull", True),):
    print(f"dist world=1 {mode} fused={fused} us/step", round(timeit(DistributedParticleFilter(lgssm_step, n, mode=mode, fused=fused)), 2))
dist.destroy_process_group()

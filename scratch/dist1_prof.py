import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch, torch.distributed as dist
import genjax_b200 as gj
from genjax_b200.inference.pf_dist import DistributedParticleFilter
from genjax_b200.workloads import lgssm_step
os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29556")
torch.cuda.set_device(0); dev = torch.device("cuda", 0)
dist.init_process_group("nccl", rank=0, world_size=1, device_id=dev)
n, T = 1 << 20, 12
x0 = torch.randn(n, device=dev); ys = torch.randn(T, device=dev)
pf = DistributedParticleFilter(lgssm_step, n, mode="pull")
for _ in range(2): pf.run(gj.key(1), x0, gj.C["y"].set(ys), use_graph=False)
torch.cuda.synchronize()
dist.destroy_process_group()

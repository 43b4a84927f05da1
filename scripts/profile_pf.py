#!/usr/bin/env python
"""Short particle-filter run for ncu captures (not a benchmark).

    ncu --set full --clock-control none --import-source on -k regex:model_kernel -s 4 -c 3 \
        -o gpurun_out/prof_model_d1 python scripts/profile_pf.py --dim 1
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import torch

import genjax_b200 as gj
from genjax_b200.inference.pf import ParticleFilter
from genjax_b200.workloads import LG_Q, LG_R, lgssm_step, lgssm_step_vec

ap = argparse.ArgumentParser()
ap.add_argument("--dim", type=int, default=1)
ap.add_argument("--particles", type=int, default=1 << 20)
ap.add_argument("--T", type=int, default=8)
ap.add_argument("--graph", action="store_true")
ap.add_argument("--mode", default="step")
ap.add_argument("--reference-max", default="running")
ap.add_argument("--single-pass", action="store_true")
ap.add_argument("--obs-sd", type=float, default=None)
a = ap.parse_args()
dev = torch.device("cuda", 0)
n, d, T = a.particles, a.dim, a.T
g = np.random.default_rng(0)
ys = g.standard_normal((T, d) if d > 1 else T).astype(np.float32)
x0 = torch.from_numpy(g.standard_normal((n, d) if d > 1 else n).astype(np.float32)).to(dev)
if d == 1:
    model, shared = lgssm_step, ()
else:
    model, shared = lgssm_step_vec, (torch.full((d,), LG_Q, device=dev), torch.full((d,), LG_R if a.obs_sd is None else a.obs_sd, device=dev))
pf = ParticleFilter(model, n, mode=a.mode, reference_max=a.reference_max, single_pass=a.single_pass)
res = pf.run(gj.key(1), x0, gj.C["y"].set(torch.from_numpy(ys).to(dev)), shared_args=shared, use_graph=a.graph)
torch.cuda.synchronize()
print("logZ", res.log_marginal_likelihood.item())

"""What does the last CTA's 4096-tile table build cost when NO peer is involved?  One GPU, 8 388 608 particles (4096 tiles: the
table of an 8-rank run at 1 M particles per rank), the step kernel in its table form (GJB_STEP_TABLE=1: records through the local
mailbox, last CTA builds the per-tile table) against the table-free form."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import genjax_b200 as gj
from genjax_b200.inference.pf import ParticleFilter
from genjax_b200.workloads import lgssm_step

dev = torch.device("cuda", 0)
T = 20
ys = torch.from_numpy(np.random.default_rng(0).standard_normal(T).astype(np.float32)).to(dev)
for n in [int(a) for a in sys.argv[1:]] or [1 << 23, 1 << 21]:
    x0 = torch.randn(n, device=dev)
    out = {}
    for form in ("0", "1"):
        os.environ["GJB_STEP_TABLE"] = form
        pf = ParticleFilter(lgssm_step, n, mode="step")
        for _ in range(3):
            res = pf.run(gj.key(1), x0, gj.C["y"].set(ys))
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for k in range(5):
            res = pf.run(gj.key(2 + k), x0, gj.C["y"].set(ys))
        e1.record(); torch.cuda.synchronize()
        out[form] = (e0.elapsed_time(e1) * 1e3 / (5 * T), float(res.log_marginal_likelihood))
    print(f"n={n} tiles={n // 2048}: table-free {out['0'][0]:.2f} us/step, table form {out['1'][0]:.2f} us/step, "
          f"difference {out['1'][0] - out['0'][0]:.2f}; logZ equal: {out['0'][1] == out['1'][1]}", flush=True)

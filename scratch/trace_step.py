"""Phase timeline of pf_step_kernel (needs GJB_NVCC_EXTRA=-DGJB_TRACE): SM-clock deltas between the trace points of one
launch, averaged over CTAs, and the spread of CTA start / end times (globaltimer).
    GJB_NVCC_EXTRA=-DGJB_TRACE python scratch/trace_step.py [--dim 1]"""
import argparse, ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import genjax_b200 as gj
from genjax_b200.inference.pf import ParticleFilter
from genjax_b200.workloads import LG_Q, LG_R, lgssm_step, lgssm_step_vec

ap = argparse.ArgumentParser(); ap.add_argument("--dim", type=int, default=1); ap.add_argument("--particles", type=int, default=1 << 20)
ap.add_argument("--obs-sd", type=float, default=None)
a = ap.parse_args()
dev = torch.device("cuda", 0); n, d, T = a.particles, a.dim, 6
g = np.random.default_rng(0)
ys = torch.from_numpy(g.standard_normal(T if d == 1 else (T, d)).astype(np.float32))
x0 = torch.from_numpy(g.standard_normal(n if d == 1 else (n, d)).astype(np.float32))
r_sd = a.obs_sd if a.obs_sd is not None else LG_R
shared = () if d == 1 else (torch.full((d,), LG_Q), torch.full((d,), r_sd))
pf = ParticleFilter(lgssm_step if d == 1 else lgssm_step_vec, n, mode="step")
for rep in range(4):
    res = pf.run(gj.key(rep), x0, gj.C["y"].set(ys), shared_args=shared, use_graph=os.environ.get("GJB_TRACE_GRAPH", "1") == "1")
torch.cuda.synchronize()
plan = next(iter(pf._plans.values()))
lib = plan.cm.lib
tiles = min((n + 2047) // 2048, 1024)
buf = (C.c_ulonglong * (tiles * 16))()
lib.gjb_model_trace_read.argtypes = [C.c_void_p, C.c_int]
assert lib.gjb_model_trace_read(buf, tiles * 16) == 0
t = np.frombuffer(buf, dtype=np.uint64).reshape(tiles, 16).astype(np.int64)
names = {0: "entry", 1: "recs requested + RNG drawn", 2: "E (sync 1)", 3: "tile prefix (sync 2)", 4: "parent range (sync 3)",
         5: "parent loop (sync 4)", 6: "max-scan (sync 5)", 8: "gather + body", 9: "tile max (sync 6)", 10: "masses + scan (sync 7)", 11: "cdf / rec stores"}
order = [0, 1, 2, 3, 4, 5, 6, 8, 9, 10, 11]
print(f"pf_step_kernel d={d} n={n}: SM-clock cycles per phase, mean over {tiles} CTAs (median) [p95]")
for a_, b_ in zip(order[:-1], order[1:]):
    dlt = t[:, b_] - t[:, a_]
    print(f"  {names[b_]:34s} {dlt.mean():9.0f} ({np.median(dlt):7.0f}) [{np.percentile(dlt, 95):7.0f}]")
tot = t[:, 11] - t[:, 0]
print(f"  {'CTA total':34s} {tot.mean():9.0f} ({np.median(tot):7.0f}) [{np.percentile(tot, 95):7.0f}]")
st, en = t[:, 14] - t[:, 14].min(), t[:, 15] - t[:, 14].min()
print(f"globaltimer (ns): CTA starts spread {st.max()}, median start {np.median(st):.0f}; first end {en.min()}, last end {en.max()}")

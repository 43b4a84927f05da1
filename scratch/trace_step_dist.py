"""Phase timeline of pf_step_kernel in the multi-GPU global-resampling filter (GJB_NVCC_EXTRA=-DGJB_TRACE):
    GJB_NVCC_EXTRA=-DGJB_TRACE torchrun --nproc-per-node 2 scratch/trace_step_dist.py
Every rank prints mean / p95 / max SM-clock cycles per phase over its CTAs and the slowest CTAs of the gather + body and
parent-loop phases (the ones that read peer memory)."""
import ctypes as C, os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import genjax_b200 as gj
from genjax_b200.inference.pf_dist import DistributedParticleFilter
from genjax_b200.workloads import lgssm_step

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n, T = 1 << 20, int(os.environ.get("GJB_TRACE_T", "32"))
g = np.random.default_rng(0)
ys = torch.from_numpy(g.standard_normal(T).astype(np.float32))
x0 = torch.from_numpy(np.random.default_rng(1 + rank).standard_normal(n).astype(np.float32))
pf = DistributedParticleFilter(lgssm_step, n)
for rep in range(4):
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    res = pf.run(gj.key(rep), x0, gj.C["y"].set(ys), use_graph=os.environ.get("GJB_TRACE_GRAPH", "1") == "1")
    e1.record()
    torch.cuda.synchronize()
    run_us = e0.elapsed_time(e1) * 1e3 / T
torch.cuda.synchronize(); dist.barrier()
plan = next(iter(pf._plans.values()))
lib = plan.cm.lib
tiles = 512
buf = (C.c_ulonglong * (tiles * 16))()
lib.gjb_model_trace_read.argtypes = [C.c_void_p, C.c_int]
assert lib.gjb_model_trace_read(buf, tiles * 16) == 0
t = np.frombuffer(buf, dtype=np.uint64).reshape(tiles, 16).astype(np.int64)
light = os.environ.get("GJB_STEP_LIGHT", "1") != "0" and os.environ.get("GJB_STEP_TABLE_KERNEL", "0") != "1"
names = {0: "entry", 1: "RNG drawn", 12: "wait for the previous launch", 2: "table S/E + rank range", 3: "records + rank prefix",
         4: "tile range" if light else "table + first rows", 5: "parent loop", 6: "max-scan", 8: "gather + body",
         9: "tile max", 10: "masses + scan", 11: "cdf stores", 15: "mail records"}
order = [0, 1, 12, 2, 3, 4, 5, 6, 8, 9, 10, 11] if light else [0, 1, 12, 4, 5, 6, 8, 9, 10, 11]
for r in range(world):
    if r == rank:
        print(f"[rank {rank}] pf_step_kernel (global resampling, {world} GPUs): cycles per phase, mean [p95] max(CTA)")
        for a_, b_ in zip(order[:-1], order[1:]):
            d = t[:, b_] - t[:, a_]
            print(f"  {names[b_]:30s} {d.mean():8.0f} [{np.percentile(d, 95):8.0f}] {d.max():8d} (CTA {d.argmax()})")
        tot = t[:, 11] - t[:, 0]
        print(f"  {'CTA total':30s} {tot.mean():8.0f} [{np.percentile(tot, 95):8.0f}] {tot.max():8d} (CTA {tot.argmax()})")
        st, en = t[:, 14] - t[:, 14].min(), t[:, 15] - t[:, 14].min()
        print(f"  globaltimer ns: starts spread {st.max()}, first end {en.min()}, last end {en.max()} (CTA {en.argmax()})")
        pe = np.percentile(en, [50, 90, 99])
        print(f"  CTA end times ns after the first start: p50 {pe[0]:.0f} p90 {pe[1]:.0f} p99 {pe[2]:.0f} max {en.max()}; "
              f"the 4 latest CTAs {list(np.argsort(en)[-4:][::-1])}")
        work_ns = (t[:, 11] - t[:, 12]) / 1.965  # cycles from the release (end of the wait) to the end of the stores
        pw = np.percentile(work_ns, [50, 90, 99])
        print(f"  CTA work after its release, ns: p50 {pw[0]:.0f} p90 {pw[1]:.0f} p99 {pw[2]:.0f} max {work_ns.max():.0f} (CTA {work_ns.argmax()})")
        last = np.nonzero(t[:, 7])[0]
        if len(last):
            b = int(last[0])
            others = np.delete(t[:, 15], b)
            print(f"  last CTA {b}: other CTAs of this rank all done at {others.max() - t[:, 14].min()} ns; all records of all ranks seen at "
                  f"{t[b, 7] - t[:, 14].min()} ns; table written at {t[b, 13] - t[:, 14].min()} ns; CTA exit {t[b, 15] - t[:, 14].min()} ns", flush=True)
        print(f"  run: {run_us:.1f} us per step over the last run (T = {T}, events)", flush=True)
    dist.barrier()
dist.destroy_process_group()

"""The scoring half of the path on its own: ``model.assess`` (every site constrained) over a batch of particles, the
streaming shape of the fused model kernel -- per particle it reads the state row and the choice row and writes one score.
No RNG work, so this is where the kernel meets HBM: reports algorithmic GB/s against MEASURED_PEAKS.json.

  python scripts/bench_assess.py [--dim 32] [--particles 4194304] [--iters 50]
"""
import argparse
import ctypes as C
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import genjax_b200 as gj  # noqa: E402
from genjax_b200.runtime import cabi  # noqa: E402
from genjax_b200.workloads import lgssm_step, lgssm_step_vec  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dim", type=int, default=32)
    ap.add_argument("--particles", type=int, default=1 << 22)
    ap.add_argument("--iters", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    n, d = a.particles, a.dim
    g = torch.Generator(device=dev).manual_seed(1)
    if d == 1:
        model = lgssm_step
        x_prev = torch.randn(n, device=dev, generator=g)
        x = torch.randn(n, device=dev, generator=g)
        chm = gj.C["x"].set(gj.Batched(x)) | gj.C["y"].set(0.3)
        args, axes = (x_prev,), (0,)
    else:
        model = lgssm_step_vec
        x_prev = torch.randn(n, d, device=dev, generator=g)
        x = torch.randn(n, d, device=dev, generator=g)
        y = torch.randn(d, device=dev, generator=g)
        q = torch.ones(d, device=dev)
        r = torch.full((d,), 0.5 * d ** 0.5, device=dev)
        chm = gj.C["x"].set(gj.Batched(x)) | gj.C["y"].set(y)
        args, axes = (x_prev, q, r), (0, None, None)

    def call():
        return gj.vmap(model.assess, in_axes=(None, axes))(chm, args)

    score, _ = call()  # compiles / loads the model library
    cm = list(model._cache.values())[-1]
    saved = []
    orig = cm.lib.gjb_model_launch

    def spy(a_ref, stream):
        saved.append(cabi.ModelArgs.from_buffer_copy(a_ref._obj))
        return orig(a_ref, stream)

    cm.lib.gjb_model_launch = spy
    score2, _ = call()
    cm.lib.gjb_model_launch = orig
    A = saved[-1]
    keep = (x_prev, x, chm, args, score2)  # the buffers the captured launch points at
    st = cabi.stream_ptr(dev)
    for _ in range(a.warmup):
        cabi.check(orig(C.byref(A), st), "launch")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(a.iters):
        orig(C.byref(A), st)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / a.iters
    # end to end through the public call (argument binding, output allocation, launch), device-timed
    torch.cuda.synchronize()
    e0.record()
    for _ in range(a.iters):
        call()
    e1.record()
    torch.cuda.synchronize()
    us_api = e0.elapsed_time(e1) * 1e3 / a.iters

    bytes_per_particle = (2 * d + 1) * 4  # state row + choice row read, one score written
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6550.1)) if isinstance(peaks, dict) else 6550.1
    gbs = bytes_per_particle * n / us / 1e3
    ref = score.double().sum().item()
    print(json.dumps({
        "workload": f"lgssm d={d}: model.assess over {n} particles (all sites constrained; scoring only, no RNG)",
        "kernel": "model_kernel (assess flags)", "us_per_launch": us, "us_per_call_public_api": us_api,
        "particles_per_s": n / us * 1e6, "algorithmic_bytes_per_particle": bytes_per_particle,
        "roofline": {"bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak},
        "working_set_mb": 2 * d * 4 * n / 2 ** 20, "score_checksum": ref, "keep": len(keep)}))


if __name__ == "__main__":
    main()

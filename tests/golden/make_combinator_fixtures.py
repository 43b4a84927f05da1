#!/usr/bin/env python
"""Freezes oracle outputs for the Scan / Vmap combinators and the output-slot resampler into
tests/golden/combinator_fixtures.npz (same status as oracle_fixtures.npz: regression vectors of the oracle
restatement, NOT reference outputs -- the reference cannot be imported in this image).
Regenerate:  python tests/golden/make_combinator_fixtures.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import gfi, rng, smc  # noqa: E402

F32 = np.float32
STDS = np.array([2.0, 4.0, 3.0, 5.0, 1.0], dtype=F32)
YOBS = np.array([0.5, -1.0, 2.0, 0.0, 1.0], dtype=F32)


def walk(h, x, std):
    nx = h.normal("x", x, std)
    y = h.normal("y", (F32(2.0) * nx).astype(F32), F32(0.5))
    return nx, (nx + y).astype(F32)


def cell(h, x):
    return h.normal("z", x, F32(1.0))


def build():
    out = {"scan_stds": STDS, "scan_yobs": YOBS}
    n = 6
    trs, carry, ys, score = gfi.scan_simulate(walk, rng.split(rng.key(314159), n), F32(0.25), STDS)
    out["scan_x"] = np.stack([t.choices["x"] for t in trs], 1)
    out["scan_y"] = np.stack([t.choices["y"] for t in trs], 1)
    out["scan_score"] = score
    out["scan_carry"] = np.broadcast_to(carry, (n,)).astype(F32)
    trs, _, _, score, w = gfi.scan_generate(walk, rng.split(rng.key(2), n), lambda t: {"y": YOBS[t]}, F32(0.1), STDS)
    out["scan_imp_x"] = np.stack([t.choices["x"] for t in trs], 1)
    out["scan_imp_weight"], out["scan_imp_score"] = w, score
    tr = gfi.simulate(cell, rng.split(rng.key(314159), 50), (np.arange(50, dtype=F32),))
    out["vmap_z"], out["vmap_score_lanes"] = tr.choices["z"], tr.get_score()
    g = np.random.default_rng(3)
    z = g.standard_normal(5000).astype(F32)
    lc = F32(0.5 * np.log(2 * np.pi) + np.log(0.5))
    lw = (F32(-0.5) * z * z - lc).astype(F32)
    out["pull_logw"], out["pull_bound"] = lw, np.array([-lc], dtype=F32)
    key = rng.split(rng.key(11))[1]
    out["pull_ancestors_true_max"] = smc.resample_systematic_pull(lw, key)
    out["pull_ancestors_bounded"] = smc.resample_systematic_pull(lw, key, M=F32(-lc))
    out["pull_lme_bounded"] = np.array([smc.log_mean_exp_ref(lw, F32(-lc))], dtype=np.float64)
    return out


if __name__ == "__main__":
    out = build()
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "combinator_fixtures.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items()})

#!/usr/bin/env python
"""Freezes outputs of the CPU oracle into tests/golden/oracle_fixtures.npz.

The reference (GenJAX on jax + tfp) cannot be imported in this image, so these are NOT reference outputs:
they are regression vectors of the oracle restatement (itself pinned to the reference's known answers in
tests/golden/reference_kats.json).  They (a) detect silent drift of the oracle and (b) let the GPU tests
compare the CUDA path against committed numbers.   Regenerate:  python tests/golden/make_oracle_fixtures.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import dists, gfi, mcmc, rng, smc  # noqa: E402

F32 = np.float32


def lg_step(h, x_prev):
    x = h.normal("x", F32(0.9) * x_prev, F32(1.0))
    h.normal("y", F32(1.0) * x, F32(0.5))
    return x


def build():
    out = {}
    words = (0x12345678, 0x9ABCDEF0)
    idx = np.arange(64, dtype=np.uint64) + np.uint64(1000)
    out["philox_words"] = np.stack(rng.site_words(words, idx, 3, 1), 1)
    out["normal_vec"] = rng.normal_vec(words, idx, 2, 8)
    out["quad_normal"] = rng.quad_normal(words, idx, 1)
    out["quad_u01"] = rng.quad_u01(words, idx, 1)
    v = np.linspace(0.05, 0.95, 19).astype(F32)
    out["logpdf_inputs"] = v
    out["logpdf_normal"] = dists.normal_logpdf(v, F32(0.3), F32(0.7))
    out["logpdf_beta"] = dists.beta_logpdf(v, F32(2.0), F32(3.5))
    out["logpdf_gamma"] = dists.gamma_logpdf(v, F32(2.5), F32(1.5))
    out["logpdf_exponential"] = dists.exponential_logpdf(v, F32(2.0))
    out["logpdf_flip"] = dists.flip_logpdf(np.array([0, 1, 1, 0]), np.array([0.2, 0.2, 0.9, 0.9], dtype=F32))
    out["logpdf_categorical"] = dists.categorical_logpdf(np.arange(4), np.array([0.1, -0.4, 1.3, 0.0], dtype=F32))
    x = -np.linspace(0, 40, 257).astype(F32)
    out["det_exp_q_in"] = x
    out["det_exp_q"] = smc.det_exp_q(x)
    g = np.random.default_rng(7)
    lw = (g.standard_normal(4099) * 2.0).astype(F32)
    out["resample_logw"] = lw
    out["resample_systematic"] = smc.resample_systematic(lw, rng.key(5))
    out["resample_multinomial"] = smc.resample_multinomial(lw, rng.split(rng.key(6), 4099))
    M, S = smc.lse_terms(lw)
    out["lse_M_S"] = np.array([float(M), float(S)], dtype=np.float64)
    # 3 steps of the linear-Gaussian bootstrap filter, 512 particles
    ys = np.array([0.3, -0.8, 1.1], dtype=F32)
    x0 = g.standard_normal(512).astype(F32)
    res = smc.particle_filter(lg_step, rng.key(99), x0, [{"y": F32(y)} for y in ys], record=True)
    out["pf_ys"], out["pf_x0"] = ys, x0
    out["pf_states"] = np.stack([h["pre_state"][0] for h in res["history"]])
    out["pf_logw"] = np.stack([h["logw"] for h in res["history"]])
    out["pf_ancestors"] = np.stack([h["ancestors"] for h in res["history"]])
    out["pf_logz_inc"] = np.array(res["logz_inc"], dtype=np.float64)
    # 3 MH transitions and one HMC edit on x ~ N(0, 3), y ~ N(x, 0.5) | y = 3
    def logp(q):
        return (dists.normal_logpdf(q[:, 0], F32(0), F32(3)) + dists.normal_logpdf(F32(3), q[:, 0], F32(0.5))).astype(F32)

    def logp_grad(q):
        xx = q[:, 0].astype(F32)
        return logp(q), ((-xx / F32(9.0)) + (F32(3) - xx) / F32(0.25)).astype(F32)[:, None]

    q0 = g.standard_normal((64, 1)).astype(F32)
    out["chain_q0"] = q0
    q, lp, acc, alpha = mcmc.mh_chain(logp, q0, rng.split(rng.key(7), 64), 3, step_size=0.3)
    out["mh_q"], out["mh_acc"], out["mh_alpha"] = q, acc, alpha
    q, lp, acc, alpha = mcmc.hmc_chain(logp_grad, q0, rng.split(rng.key(8), 64), 1, 0.05, 10, compat_stale_grad=True, accept=False)
    out["hmc_q"], out["hmc_alpha"] = q, alpha
    return out


if __name__ == "__main__":
    out = build()
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_fixtures.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")

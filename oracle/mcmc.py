"""MCMC restatement (oracle; test infrastructure only): the Rejuvenate / HMC edit
requests plus the user-side accept step, vectorised over chains in NumPy float32.

Reference:
  * ``Rejuvenate.edit`` (inference/requests/rejuvenate.py:70-94):
        proposed, fwd = proposal.propose(key, mapping(cur))
        new_tr, w, _, bwd_request = Update(proposed).edit(...)      # w = logp_new - logp_old
        bwd = proposal.assess(old, mapping(proposed))
        final_weight = w + bwd - fwd                                  (:88)
  * accept idiom (tests/inference/test_requests.py:136-137, 190-191):
        check = log(uniform(key)) < w ; tr = where(check, new, old)
  * ``HMC.edit`` (inference/requests/hmc.py:156-211): momenta ~ N(0,1) per
    selected leaf; L steps of
        p += eps/2 * gradient(carried); q += eps * p; (q, g) = value_and_grad(q); p += eps/2 * g
    returning ``(new_trace, values, gradient, momenta)`` -- i.e. the CARRIED
    gradient is never refreshed (hmc.py:186), so every step's first half-kick
    uses the gradient at the INITIAL position (``compat_stale_grad=True``);
    alpha = logp_L - logp_0 + logN(-p_L) - logN(p_0)                  (:196-203).

RNG stream shared with genjax_b200/gen/codegen_chain.py: chain lane = global
chain index; transition t uses Philox ``site`` word t + 1; chunks 0.. hold the
state-width normals (proposal noise / momenta), chunk 0xFFFF word 0 the accept
uniform.  "parity unpinned" against the reference's threefry streams.
"""

from __future__ import annotations

import numpy as np

from . import dists, rng

F32 = np.float32


def chain_normals(words, idx, t1, d):
    return rng.normal_vec(words, idx, t1, d)


def chain_uniform(words, idx, t1):
    w0, _, _, _ = rng.site_words(words, idx, t1, 0xFFFF)
    return rng.u01(w0)


def _std_normal_sum(p):
    lp = dists.normal_logpdf(p, F32(0.0), F32(1.0))
    out = np.zeros(lp.shape[0], dtype=F32)
    for k in range(lp.shape[1]):
        out = (out + lp[:, k]).astype(F32)
    return out


def _normal_sum(v, loc, scale):
    lp = dists.normal_logpdf(v, loc, scale)
    out = np.zeros(lp.shape[0], dtype=F32)
    for k in range(lp.shape[1]):
        out = (out + lp[:, k]).astype(F32)
    return out


def mh_chain(logp, q, key, n_steps, step_size=1.0, proposal=None, accept=True, step0=0, bwd_at_old=False):
    """``logp(q[n, D]) -> float32[n]``; ``proposal(q) -> (loc, scale)`` (default: random walk).
    ``bwd_at_old=True``: the backward proposal arguments are taken at the OLD state, as rejuvenate.py:84-86 does
    (``argument_mapping(bwd_chm)``, bwd_chm = the discarded choices); False: Metropolis-Hastings.
    Returns (q, logp(q), accept_count, last alpha)."""
    words, idx = rng.lanes(key)
    q = np.asarray(q, dtype=F32).copy()
    n, d = q.shape
    if proposal is None:
        proposal = lambda x: (x, np.full_like(x, F32(step_size)))  # noqa: E731
    lp = logp(q).astype(F32)
    acc = np.zeros(n, dtype=np.int32)
    alpha = np.zeros(n, dtype=F32)
    for s in range(n_steps):
        t1 = step0 + s + 1
        z = chain_normals(words, idx, t1, d)
        loc, scale = proposal(q)
        loc = np.broadcast_to(np.asarray(loc, dtype=F32), q.shape)
        scale = np.broadcast_to(np.asarray(scale, dtype=F32), q.shape)
        prop = (loc + scale * z).astype(F32)
        fwd = _normal_sum(prop, loc, scale)
        loc_b, scale_b = proposal(q if bwd_at_old else prop)
        loc_b = np.broadcast_to(np.asarray(loc_b, dtype=F32), q.shape)
        scale_b = np.broadcast_to(np.asarray(scale_b, dtype=F32), q.shape)
        bwd = _normal_sum(q, loc_b, scale_b)
        lp_new = logp(prop).astype(F32)
        alpha = (((lp_new - lp).astype(F32) + bwd).astype(F32) - fwd).astype(F32)
        if accept:
            with np.errstate(divide="ignore"):
                ok = np.log(chain_uniform(words, idx, t1).astype(np.float64)).astype(F32) < alpha
        else:
            ok = np.ones(n, dtype=bool)
        q[ok] = prop[ok]
        lp = np.where(ok, lp_new, lp).astype(F32)
        acc += ok
    return q, lp, acc, alpha


def hmc_chain(logp_grad, q, key, n_iters, eps, L, compat_stale_grad=True, accept=True, step0=0):
    """``logp_grad(q[n, D]) -> (float32[n], float32[n, D])``.  Returns (q, logp, accept_count, last alpha)."""
    words, idx = rng.lanes(key)
    q0 = np.asarray(q, dtype=F32).copy()
    n, d = q0.shape
    eps = F32(eps)
    half = (eps * F32(0.5)).astype(F32)
    lp0, g0 = logp_grad(q0)
    lp0, g0 = lp0.astype(F32), g0.astype(F32)
    acc = np.zeros(n, dtype=np.int32)
    alpha = np.zeros(n, dtype=F32)
    for s in range(n_iters):
        t1 = step0 + s + 1
        p = chain_normals(words, idx, t1, d)
        k0 = _std_normal_sum(p)
        qc, gc, lp = q0.copy(), g0.copy(), lp0.copy()
        g = g0.copy()
        for _ in range(L):
            p = (p + half * gc).astype(F32)
            qc = (qc + eps * p).astype(F32)
            lp, g = logp_grad(qc)
            lp, g = lp.astype(F32), g.astype(F32)
            p = (p + half * g).astype(F32)
            if not compat_stale_grad:
                gc = g
        k1 = _std_normal_sum((-p).astype(F32))
        alpha = (((lp - lp0).astype(F32) + k1).astype(F32) - k0).astype(F32)
        if accept:
            with np.errstate(divide="ignore"):
                ok = np.log(chain_uniform(words, idx, t1).astype(np.float64)).astype(F32) < alpha
        else:
            ok = np.ones(n, dtype=bool)
        q0[ok] = qc[ok]
        g0[ok] = g[ok]
        lp0 = np.where(ok, lp, lp0).astype(F32)
        acc += ok
    return q0, lp0, acc, alpha

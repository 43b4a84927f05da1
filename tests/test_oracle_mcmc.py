"""Pins the oracle's MCMC restatement (oracle/mcmc.py) to the outcomes the reference's own tests assert
(/root/reference/tests/inference/test_requests.py): Rejuvenate with the prior as proposal has weight 0 (:141-166),
Rejuvenate + accept converges to the posterior mean 3.0 (:168-193), and 20 bare HMC edits with eps = 1e-2, L = 10
-- integrated exactly as inference/requests/hmc.py:170-186 does, carried gradient included -- reach x = 3.0 within
5e-3 (:197-235)."""
import numpy as np
import pytest

from oracle import dists as od
from oracle import mcmc, rng

F32 = np.float32


def test_rejuvenate_with_prior_proposal_has_zero_weight():
    n = 512
    q0 = rng.normal_vec((1, 2), np.arange(n, dtype=np.uint64), 1, 1)

    def logp(q):
        return od.normal_logpdf(q[:, 0], F32(0.0), F32(1.0))

    def prior_proposal(q):
        return np.zeros_like(q), np.ones_like(q)

    q, lp, acc, alpha = mcmc.mh_chain(logp, q0, rng.split(rng.key(314159), n), 1, proposal=prior_proposal, accept=False)
    assert np.all(q != q0)
    assert np.abs(alpha).max() < 2e-6  # (lp_new - lp_old) + bwd - fwd == 0 up to fp32 rounding


def test_rejuvenate_random_walk_converges_to_posterior_mean():
    """y1 ~ N(0, 3), y2 ~ N(y1, 0.001) | y2 = 3; RW(0.3) proposal + accept, 100 steps; many chains."""
    n = 2048

    def logp(q):
        return (od.normal_logpdf(q[:, 0], F32(0.0), F32(3.0)) + od.normal_logpdf(F32(3.0), q[:, 0], F32(0.001))).astype(F32)

    q0 = (3.0 * rng.normal_vec((5, 6), np.arange(n, dtype=np.uint64), 1, 1)).astype(F32)
    q, lp, acc, _ = mcmc.mh_chain(logp, q0, rng.split(rng.key(0), n), 100, step_size=0.3)
    assert np.median(np.abs(q[:, 0] - 3.0)) < 5e-3
    # a finer random walk started there samples the N(3, 0.001) posterior itself (robust statistics: the few chains
    # the coarse walk left far out in the tail need more than 2000 fine steps to walk back)
    q, lp, acc, _ = mcmc.mh_chain(logp, q, rng.split(rng.key(1), n), 2000, step_size=0.002)
    d = q[:, 0].astype(np.float64) - 3.0
    assert abs(np.median(d)) < 1e-4
    assert 1.4826 * np.median(np.abs(d - np.median(d))) == pytest.approx(0.001, rel=0.15)


def test_hmc_edit_reference_integrator_converges():
    """x ~ N(0, 1), y ~ N(x, 0.01) | y = 3: 20 x HMC(eps=1e-2, L=10).edit without accept (test_requests.py:197-235)."""
    n = 256

    def logp_grad(q):
        x = q[:, 0].astype(F32)
        lp = (od.normal_logpdf(x, F32(0.0), F32(1.0)) + od.normal_logpdf(F32(3.0), x, F32(0.01))).astype(F32)
        return lp, ((-x + (F32(3.0) - x) / F32(1e-4)).astype(F32))[:, None]

    q = rng.normal_vec((7, 8), np.arange(n, dtype=np.uint64), 1, 1)
    lp0, _ = logp_grad(q)
    q1, lp1, _, alpha = mcmc.hmc_chain(logp_grad, q, rng.split(rng.key(0), n), 1, 1e-2, 10, compat_stale_grad=True, accept=False)
    assert np.all(alpha != 0.0)
    assert np.all((alpha - (lp1 - lp0)) != 0.0)  # the momenta terms contribute (test_requests.py:226-227)
    for i in range(20):
        q, _, _, _ = mcmc.hmc_chain(logp_grad, q, rng.split(rng.fold_in(rng.key(9), i), n), 1, 1e-2, 10,
                                    compat_stale_grad=True, accept=False)
    assert np.abs(q[:, 0] - 3.0).max() < 0.05
    assert q[:, 0].mean() == pytest.approx(3.0, abs=5e-3)


def test_textbook_hmc_samples_the_posterior_and_compat_does_not():
    """With an accept step at eps = 0.15 the reference integrator (stale first half-kick) is biased; the textbook
    leapfrog is exact: x ~ N(0, 1), y ~ N(x, 0.5) | y = 1 has posterior N(0.8, 0.2)."""
    n = 8192

    def logp_grad(q):
        x = q[:, 0].astype(F32)
        lp = (od.normal_logpdf(x, F32(0.0), F32(1.0)) + od.normal_logpdf(F32(1.0), x, F32(0.5))).astype(F32)
        return lp, ((-x + (F32(1.0) - x) / F32(0.25)).astype(F32))[:, None]

    q0 = rng.normal_vec((3, 4), np.arange(n, dtype=np.uint64), 1, 1)
    q, _, acc, _ = mcmc.hmc_chain(logp_grad, q0, rng.split(rng.key(4), n), 60, 0.15, 5, compat_stale_grad=False)
    assert q.mean() == pytest.approx(0.8, abs=0.03) and q.var() == pytest.approx(0.2, rel=0.1) and acc.mean() / 60 > 0.9
    q, _, _, _ = mcmc.hmc_chain(logp_grad, q0, rng.split(rng.key(4), n), 60, 0.15, 5, compat_stale_grad=True)
    assert abs(q.mean() - 0.8) > 0.1

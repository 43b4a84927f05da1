"""GPU parity of the batched MCMC kernels (gjb_model_mh_chain / gjb_model_hmc_chain)
against the NumPy oracle, plus the reference's own MCMC tests run over many chains
(/root/reference/tests/inference/test_requests.py)."""
import math

import numpy as np
import pytest
import torch

from oracle import dists as od
from oracle import mcmc as omcmc
from oracle import rng as orng

pytestmark = pytest.mark.gpu
F32 = np.float32


def _gj():
    import genjax_b200 as gj

    return gj


def _linked(sd1, sd2):
    gj = _gj()

    @gj.gen
    def linked_normal():
        y1 = gj.normal(0.0, sd1) @ "y1"
        gj.normal(y1, sd2) @ "y2"

    return linked_normal


def o_linked_logp(sd1, sd2, y2):
    def f(q):
        a = od.normal_logpdf(q[:, 0], F32(0.0), F32(sd1))
        b = od.normal_logpdf(F32(y2), q[:, 0], F32(sd2))
        return (a + b).astype(F32)

    return f


def test_mh_random_walk_matches_oracle(device):
    gj = _gj()
    from genjax_b200.inference.mcmc import mh_chain

    n = 4096
    model = _linked(3.0, 0.5)
    tr, _ = model.importance(gj.split(gj.key(1), n), gj.C.kw(y2=3.0), ())
    q0 = tr.get_choices()["y1"].cpu().numpy().reshape(n, 1)
    logp = o_linked_logp(3.0, 0.5, 3.0)
    kb = gj.split(gj.key(7), n)
    okb = orng.split(orng.key(7), n)
    # one transition: proposal, weight and accept decision chain by chain
    res = mh_chain(kb, tr, gj.S["y1"], step_size=0.3, n_steps=1)
    oq, olp, oacc, oalpha = omcmc.mh_chain(logp, q0, okb, 1, step_size=0.3)
    np.testing.assert_allclose(res.alpha.cpu().numpy(), oalpha, rtol=2e-4, atol=2e-4)
    got = res.trace.get_choices()["y1"].cpu().numpy()
    same = res.accept_count.cpu().numpy() == oacc
    assert same.mean() > 0.999
    np.testing.assert_allclose(got[same], oq[same, 0], rtol=1e-5, atol=2e-6)
    np.testing.assert_allclose(res.trace.get_score().cpu().numpy()[same], olp[same], rtol=1e-5, atol=1e-5)
    # 25 transitions in one launch == the oracle's 25-step loop (a rare 1-ulp accept flip aside)
    res = mh_chain(kb, tr, gj.S["y1"], step_size=0.3, n_steps=25)
    oq, _, oacc, _ = omcmc.mh_chain(logp, q0, okb, 25, step_size=0.3)
    got = res.trace.get_choices()["y1"].cpu().numpy()
    close = np.isclose(got, oq[:, 0], rtol=1e-4, atol=1e-5)
    assert close.mean() > 0.995
    # and a launch split in two (step0 continues the stream) equals one launch
    r1 = mh_chain(kb, tr, gj.S["y1"], step_size=0.3, n_steps=10)
    r2 = mh_chain(kb, r1.trace, gj.S["y1"], step_size=0.3, n_steps=15, step0=10)
    assert torch.equal(r2.trace.get_choices()["y1"], res.trace.get_choices()["y1"])


def test_rejuvenate_prior_proposal_has_zero_weight(device):
    """tests/inference/test_requests.py:141-166."""
    gj = _gj()
    from genjax_b200.inference.requests import Rejuvenate, StaticRequest

    @gj.gen
    def simple_normal():
        gj.normal(0.0, 1.0) @ "y1"

    n = 2000
    kb = gj.split(gj.key(314159), n)
    tr = simple_normal.simulate(kb, ())
    old_v = tr.get_choices()["y1"].clone()
    request = StaticRequest({"y1": Rejuvenate(gj.normal, lambda chm: (0.0, 1.0))})
    new_tr, w, _, bwd = request.edit(gj.split(gj.key(2), n), tr, ())
    new_v = new_tr.get_choices()["y1"]
    assert (old_v != new_v).all()
    assert w.abs().max().item() < 2e-6  # == 0.0 up to fp32 (a - b) + b - a
    assert torch.equal(bwd.constraint["y1"], old_v)


def test_rejuvenate_convergence(device):
    """tests/inference/test_requests.py:168-193 over 4096 chains: request.edit + accept, 100 steps."""
    gj = _gj()
    from genjax_b200.inference.mcmc import mh_accept, mh_chain
    from genjax_b200.inference.requests import Rejuvenate, StaticRequest

    n = 4096
    model = _linked(3.0, 0.001)
    tr, _ = model.importance(gj.split(gj.key(314159), n), gj.C.kw(y2=3.0), ())
    request = StaticRequest({"y1": Rejuvenate(gj.normal, lambda chm: (chm.get_value(), 0.3))})
    key = gj.key(0)
    for i in range(100):
        k1, k2 = gj.split(gj.fold_in(key, i))
        new_tr, w, _, _ = request.edit(gj.split(k1, n), tr, ())
        tr, _ = mh_accept(gj.split(k2, n), new_tr, tr, w)
    y1 = tr.get_choices()["y1"]
    # the RW(0.3) chain on a 0.001-wide posterior mixes slowly: most chains have arrived, like the reference's one
    assert (y1 - 3.0).abs().median().item() < 5e-3
    # the fused equivalent: 2000 transitions in ONE launch; every chain is now in the posterior bulk
    res = mh_chain(gj.split(gj.key(5), n), tr, gj.S["y1"], step_size=0.01, n_steps=2000)
    y1 = res.trace.get_choices()["y1"]
    assert y1.mean().item() == pytest.approx(3.0, abs=2e-4)
    assert y1.std().item() == pytest.approx(0.001, rel=0.15)


def test_regenerate_mh_convergence(device):
    """tests/inference/test_requests.py:120-139: Regenerate(S['y1']) + accept, 200 steps."""
    gj = _gj()
    from genjax_b200.inference.mcmc import mh_accept

    n = 8192
    model = _linked(3.0, 0.01)
    tr, _ = model.importance(gj.split(gj.key(314159), n), gj.C.kw(y2=3.0), ())
    request = gj.Regenerate(gj.S["y1"])
    key = gj.key(1)
    for i in range(200):
        k1, k2 = gj.split(gj.fold_in(key, i))
        new_tr, w, _, _ = request.edit(gj.split(k1, n), tr, ())
        tr, _ = mh_accept(gj.split(k2, n), new_tr, tr, w)
    # independence sampler from the N(0, 3) prior against a 0.01-wide posterior at 3.0: the per-chain
    # acceptance probability is ~0.5 % a step; after 200 steps most chains sit within 0.03 of 3.0
    y1 = tr.get_choices()["y1"]
    assert (y1 - 3.0).abs().median().item() < 3e-2


def o_simple_logp_grad(q):
    """x ~ N(0,1), y ~ N(x, 0.01) | y = 3: value and gradient in float32."""
    x = q[:, 0].astype(F32)
    lp = (od.normal_logpdf(x, F32(0.0), F32(1.0)) + od.normal_logpdf(F32(3.0), x, F32(0.01))).astype(F32)
    g = (-x + (F32(3.0) - x) / F32(0.01 * 0.01)).astype(F32)
    return lp, g[:, None]


def test_hmc_edit_matches_oracle_and_reference_semantics(device):
    """tests/inference/test_requests.py:197-235 + step-level parity with the oracle's restatement of
    hmc.py:156-211 (including the carried gradient of hmc.py:186)."""
    gj = _gj()
    from genjax_b200.inference.requests import HMC

    @gj.gen
    def model():
        x = gj.normal(0.0, 1.0) @ "x"
        y = gj.normal(x, 0.01) @ "y"
        return y

    n = 2048
    tr, _ = model.importance(gj.split(gj.key(0), n), gj.ChoiceMap.kw(y=3.0), ())
    request = HMC(gj.Selection.at["x"], 1e-2)
    kb = gj.split(gj.key(3), n)
    new_tr, fwd_w, _, _ = request.edit(kb, tr, ())
    old_x, new_x = tr.get_choices()["x"], new_tr.get_choices()["x"]
    old_d = gj.normal.logpdf(old_x, 0.0, 1.0) + gj.normal.logpdf(torch.full_like(old_x, 3.0), old_x, 0.01)
    new_d = gj.normal.logpdf(new_x, 0.0, 1.0) + gj.normal.logpdf(torch.full_like(new_x, 3.0), new_x, 0.01)
    assert (fwd_w != 0).all()
    torch.testing.assert_close(new_tr.get_score() - tr.get_score(), new_d - old_d, rtol=1e-4, atol=0.5)
    assert ((fwd_w - (new_tr.get_score() - tr.get_score())).abs() > 0).all()
    # oracle parity of one edit (L = 10 leapfrog steps, stale carried gradient, no accept)
    q0 = old_x.cpu().numpy().reshape(n, 1)
    oq, olp, _, oalpha = omcmc.hmc_chain(o_simple_logp_grad, q0, orng.split(orng.key(3), n), 1, 1e-2, 10,
                                         compat_stale_grad=True, accept=False)
    np.testing.assert_allclose(new_x.cpu().numpy(), oq[:, 0], rtol=2e-4, atol=2e-4)
    scale = np.maximum(1.0, np.abs(olp))
    assert np.max(np.abs(fwd_w.cpu().numpy() - oalpha) / scale) < 2e-3
    # gradient convergence, as the reference test: 20 bare edits
    key = gj.key(9)
    cur = tr
    for i in range(20):
        cur, *_ = request.edit(gj.split(gj.fold_in(key, i), n), cur, ())
    assert cur.get_choices()["x"].mean().item() == pytest.approx(3.0, abs=5e-3)


def test_hmc_chain_textbook_and_compat_sample_the_posterior(device):
    gj = _gj()
    from genjax_b200.inference.mcmc import hmc_chain

    @gj.gen
    def model():
        x = gj.normal(0.0, 1.0) @ "x"
        gj.normal(x, 0.5) @ "y"

    n = 1 << 15
    tr, _ = model.importance(gj.split(gj.key(0), n), gj.C.kw(y=1.0), ())
    post_var = 1.0 / (1.0 + 4.0)
    post_mean = post_var * 4.0 * 1.0
    # textbook leapfrog: a valid sampler of the exact posterior N(0.8, 0.2)
    res = hmc_chain(gj.split(gj.key(4), n), tr, gj.S["x"], eps=0.15, L=5, n_iters=60, compat_stale_grad=False)
    x = res.trace.get_choices()["x"]
    assert x.mean().item() == pytest.approx(post_mean, abs=0.02)
    assert x.var().item() == pytest.approx(post_var, rel=0.08)
    assert 0.9 < res.accept_rate.item() <= 1.0
    # reference-compatible integrator (hmc.py:186 carries the initial gradient): NOT reversible, so with an
    # accept step at this step size it is biased (mean ~0.66, variance ~0.64) -- the kernel must reproduce
    # exactly that, i.e. agree with the oracle's restatement of the reference, not with the true posterior
    res = hmc_chain(gj.split(gj.key(4), n), tr, gj.S["x"], eps=0.15, L=5, n_iters=60, compat_stale_grad=True)
    x = res.trace.get_choices()["x"].cpu().numpy()
    q0 = tr.get_choices()["x"].cpu().numpy().reshape(n, 1)

    def lpg(q):
        v = q[:, 0].astype(F32)
        lp = (od.normal_logpdf(v, F32(0), F32(1)) + od.normal_logpdf(F32(1.0), v, F32(0.5))).astype(F32)
        return lp, ((-v + (F32(1.0) - v) / F32(0.25)).astype(F32))[:, None]

    oq, _, oacc, _ = omcmc.hmc_chain(lpg, q0, orng.split(orng.key(4), n), 60, 0.15, 5, compat_stale_grad=True)
    assert np.isclose(x, oq[:, 0], rtol=1e-3, atol=1e-4).mean() > 0.99
    assert abs(x.mean() - post_mean) > 0.1  # documents the reference integrator's bias


def test_gmm_mh_config3_small(device):
    """BASELINE configs[2] at reduced size: 8-component 8-D mixture target, random-walk MH."""
    gj = _gj()
    from genjax_b200.inference.mcmc import mh_chain
    from genjax_b200.workloads import gmm_target

    K, D, n = 8, 8, 1 << 15
    g = np.random.default_rng(1)
    mu = g.uniform(-4, 4, size=(K, D)).astype(F32)
    logits = np.zeros(K, dtype=F32)
    sigma = np.full(K, 0.7, dtype=F32)
    args = (torch.from_numpy(logits), torch.from_numpy(mu), torch.from_numpy(sigma))
    tr = gmm_target.simulate(gj.split(gj.key(2), n), args)  # exact draws from the mixture
    x0 = tr.get_choices()["x"].cpu().numpy()
    # the model kernel's mixture logpdf agrees with the chain kernel's symbolic one
    def ologp(q):
        comp = np.stack([od.mv_normal_diag_logpdf(q, mu[k], np.full(D, sigma[k], F32)) for k in range(K)], 1).astype(np.float64)
        comp += -math.log(K)
        m = comp.max(1, keepdims=True)
        return (m[:, 0] + np.log(np.exp(comp - m).sum(1))).astype(F32)

    np.testing.assert_allclose(tr.get_score().cpu().numpy(), ologp(x0), rtol=2e-5, atol=2e-5)
    kb = gj.split(gj.key(8), n)
    res = mh_chain(kb, tr, gj.S["x"], step_size=0.5, n_steps=1)
    oq, olp, oacc, oalpha = omcmc.mh_chain(ologp, x0, orng.split(orng.key(8), n), 1, step_size=0.5)
    np.testing.assert_allclose(res.alpha.cpu().numpy(), oalpha, rtol=2e-4, atol=3e-4)
    assert (res.accept_count.cpu().numpy() == oacc).mean() > 0.999
    # stationarity: started from exact mixture draws, 300 MH steps leave the component occupancy uniform
    res = mh_chain(kb, tr, gj.S["x"], step_size=0.5, n_steps=300)
    x = res.trace.get_choices()["x"].cpu().numpy()
    d2 = ((x[:, None, :] - mu[None]) ** 2).sum(-1)
    occ = np.bincount(d2.argmin(1), minlength=K) / n
    np.testing.assert_allclose(occ, np.full(K, 1 / K), atol=0.012)
    np.testing.assert_allclose(x.mean(0), mu.mean(0), atol=0.08)
    assert 0.05 < res.accept_rate.item() < 0.9


def test_eight_schools_hmc_config5_small(device):
    """BASELINE configs[4] at reduced size: fused HMC on the hierarchical-normal model vs a float64 NumPy HMC."""
    gj = _gj()
    from genjax_b200.inference.mcmc import hmc_chain
    from genjax_b200.workloads import EIGHT_SCHOOLS_SIGMA, EIGHT_SCHOOLS_Y, eight_schools

    n = 1 << 14
    y = torch.tensor(EIGHT_SCHOOLS_Y)
    sig = torch.tensor(EIGHT_SCHOOLS_SIGMA)
    tr, _ = eight_schools.importance(gj.split(gj.key(4), n), gj.C["y"].set(y), (sig,))
    sel = gj.S["mu"] | gj.S["log_tau"] | gj.S["theta"]
    res = hmc_chain(gj.split(gj.key(5), n), tr, sel, eps=0.05, L=10, n_iters=300)
    ch = res.trace.get_choices()
    mu, lt = ch["mu"].cpu().numpy(), ch["log_tau"].cpu().numpy()
    assert 0.5 < res.accept_rate.item() <= 1.0

    # float64 reference sampler (textbook HMC, same eps / L), 4096 chains
    yy, ss = np.array(EIGHT_SCHOOLS_Y), np.array(EIGHT_SCHOOLS_SIGMA)

    def lpg(q):
        m, l, th = q[:, 0], q[:, 1], q[:, 2:]
        tau = np.exp(l)
        r = (th - m[:, None]) / tau[:, None]
        lp = -0.5 * (m / 5) ** 2 - 0.5 * l**2 - 0.5 * (r**2).sum(1) - 8 * l - 0.5 * (((yy - th) / ss) ** 2).sum(1)
        g = np.empty_like(q)
        g[:, 0] = -m / 25 + (r / tau[:, None]).sum(1)
        g[:, 1] = -l + (r**2).sum(1) - 8
        g[:, 2:] = -r / tau[:, None] + (yy - th) / ss**2
        return lp, g

    rs = np.random.default_rng(0)
    m_ = 4096
    q = np.concatenate([rs.normal(0, 5, (m_, 1)), rs.normal(0, 1, (m_, 1)), np.zeros((m_, 8))], 1)
    q[:, 2:] = q[:, :1] + np.exp(q[:, 1:2]) * rs.normal(size=(m_, 8))
    lp, g = lpg(q)
    for _ in range(300):
        p = rs.normal(size=q.shape)
        qn, gn, pn = q.copy(), g.copy(), p.copy()
        for _ in range(10):
            pn = pn + 0.025 * gn
            qn = qn + 0.05 * pn
            lpn, gn = lpg(qn)
            pn = pn + 0.025 * gn
        a = lpn - lp - 0.5 * (pn**2).sum(1) + 0.5 * (p**2).sum(1)
        ok = np.log(rs.random(m_)) < a
        q[ok], g[ok], lp[ok] = qn[ok], gn[ok], lpn[ok]
    assert mu.mean() == pytest.approx(q[:, 0].mean(), abs=0.25)
    assert lt.mean() == pytest.approx(q[:, 1].mean(), abs=0.08)
    assert mu.std() == pytest.approx(q[:, 0].std(), rel=0.1)


def test_rejuvenate_backward_score_reference_compat(device):
    """rejuvenate.py:84-88: the reference scores the backward move with ``argument_mapping(bwd_chm)``, bwd_chm = the OLD
    choice, i.e. log q(old; mapping(old)); ``Rejuvenate(..., reference_compat=False)`` is Metropolis-Hastings'
    log q(old; mapping(new)).  With a random-walk proposal N(chm, s) the two differ by z^2 / 2."""
    gj = _gj()
    from genjax_b200.inference.requests import Rejuvenate, StaticRequest

    @gj.gen
    def model():
        gj.normal(0.0, 3.0) @ "y1"

    n, s = 4000, 0.3
    tr = model.simulate(gj.split(gj.key(7), n), ())
    old = tr.get_choices()["y1"]
    kb = gj.split(gj.key(8), n)
    outs = {}
    for compat in (True, False):
        req = StaticRequest({"y1": Rejuvenate(gj.normal, lambda chm: (chm.get_value(), s), reference_compat=compat)})
        new_tr, w, _, _ = req.edit(kb, tr, ())
        outs[compat] = (new_tr.get_choices()["y1"], w)
    new = outs[True][0]
    assert torch.equal(new, outs[False][0])  # same proposal draw
    dlp = gj.normal.logpdf(new, 0.0, 3.0) - gj.normal.logpdf(old, 0.0, 3.0)
    fwd = gj.normal.logpdf(new, old, s)
    torch.testing.assert_close(outs[True][1], dlp + gj.normal.logpdf(old, old, s) - fwd, rtol=2e-4, atol=2e-4)   # reference
    torch.testing.assert_close(outs[False][1], dlp + gj.normal.logpdf(old, new, s) - fwd, rtol=2e-4, atol=2e-4)  # MH: = dlp
    assert Rejuvenate(gj.normal, lambda chm: (0.0, 1.0)).reference_compat

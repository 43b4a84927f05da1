"""GPU parity of the generative-function interface on the fused kernels: every device
primitive against the oracle (bit-shared RNG streams), the reference's algebraic
identities (tests/generative_functions/test_distributions.py, test_static_gen_fn.py),
ImportanceK / ChangeTarget / sample_particle (tests/inference/test_smc.py) and the README
quickstart (configs[0])."""
import json
import math
import os

import numpy as np
import pytest
import torch

from oracle import dists as od
from oracle import gfi as ogfi
from oracle import rng as orng
from oracle import smc as osmc

pytestmark = pytest.mark.gpu
F32 = np.float32
GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_kats.json")))


def _gj():
    import genjax_b200 as gj

    return gj


@pytest.mark.parametrize(
    "name,args,tol",
    [
        ("normal", (1.5, 2.0), 1e-5),
        ("uniform", (-1.0, 3.0), 1e-6),
        ("exponential", (2.5,), 2e-6),
        ("half_normal", (1.7,), 1e-5),
        ("gamma", (2.5, 1.5), 5e-5),
        ("gamma", (0.4, 1.0), 5e-5),
        ("beta", (2.0, 3.0), 5e-5),
        ("flip", (0.3,), 0),
        ("bernoulli", (-0.8,), 0),
    ],
)
def test_primitive_sample_and_logpdf_match_oracle(device, name, args, tol):
    """dist.simulate over a KeyBatch == the oracle's sampler on the same Philox lanes; score == oracle logpdf."""
    gj = _gj()
    n = 50_001
    dist = getattr(gj, name)
    kb = gj.split(gj.key(11), n)
    with pytest.warns(DeprecationWarning) if name == "bernoulli" else _nullcontext():
        tr = dist.simulate(kb, args)
    v = tr.get_retval().cpu().numpy()
    okb = orng.split(orng.key(11), n)
    words, idx = orng.lanes(okb)
    ov = od.DISTS[name][0](words, idx, 1, *[F32(a) for a in args])
    if tol == 0:
        assert (v.astype(np.int32) != np.asarray(ov).astype(np.int32)).mean() < 1e-4
    else:
        bad = ~np.isclose(v, ov, rtol=tol, atol=tol)
        assert bad.mean() < (2e-4 if name in ("gamma", "beta") else 1e-9 + 0), (name, bad.sum())
    lp = od.DISTS[name][1](v if tol else v.astype(np.int32), *[F32(a) for a in args])
    np.testing.assert_allclose(tr.get_score().cpu().numpy(), lp, rtol=2e-5, atol=2e-5)


class _nullcontext:
    def __enter__(self):
        return None

    def __exit__(self, *a):
        return False


def test_categorical_sample_and_logpdf(device):
    gj = _gj()
    n = 40_000
    logits = np.log(np.array([0.1, 0.2, 0.3, 0.4], dtype=F32))
    tr = gj.categorical(logits=torch.from_numpy(logits)).simulate(gj.split(gj.key(5), n), ())
    k = tr.get_retval().cpu().numpy()
    words, idx = orng.lanes(orng.split(orng.key(5), n))
    ok = od.categorical_sample(words, idx, 1, logits)
    assert (k != ok).mean() < 1e-4
    np.testing.assert_allclose(tr.get_score().cpu().numpy(), od.categorical_logpdf(k, logits), rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(np.bincount(k, minlength=4) / n, [0.1, 0.2, 0.3, 0.4], atol=0.01)


def test_scalar_key_equals_lane_of_batch(device):
    """vmap-over-split-keys property: lane i of a batched call == the scalar call with split(key, n)[i]."""
    gj = _gj()
    kb = gj.split(gj.key(3), 64)
    batch = gj.normal.sample(kb, 0.0, 1.0)
    for i in (0, 1, 5, 63):
        one = gj.normal.sample(kb[i], 0.0, 1.0)
        assert one.ndim == 0 and one.item() == batch[i].item()


def _model():
    gj = _gj()

    @gj.gen
    def model(mu):
        x = gj.normal(mu, 2.0) @ "x"
        s = gj.exponential(1.5) @ "s"
        y = gj.normal(x, 0.5 + s) @ "y"
        return x + y

    def o_model(h, mu):
        x = h.normal("x", mu, F32(2.0))
        s = h.exponential("s", F32(1.5))
        y = h.normal("y", x, (F32(0.5) + s).astype(F32))
        return (x + y).astype(F32)

    return model, o_model


def test_simulate_importance_assess_identities(device):
    """test_distributions.py:38-40 / test_static_gen_fn.py:196-330: importance weight == logpdf of the constrained
    sites; assess of a trace's choices == its score; retval recomputed."""
    gj = _gj()
    model, o_model = _model()
    n = 20_000
    kb = gj.split(gj.key(1), n)
    okb = orng.split(orng.key(1), n)
    tr = model.simulate(kb, (0.3,))
    otr = ogfi.simulate(o_model, okb, (F32(0.3),))
    ch = tr.get_choices()
    for a in ("x", "s", "y"):
        np.testing.assert_allclose(ch[a].cpu().numpy(), otr.choices[a], rtol=1e-5, atol=3e-6)
    np.testing.assert_allclose(tr.get_score().cpu().numpy(), otr.get_score(), rtol=1e-5, atol=2e-5)
    np.testing.assert_allclose(tr.get_retval().cpu().numpy(), otr.retval, rtol=1e-5, atol=1e-5)
    # importance with y constrained (shared scalar constraint)
    tr2, w = model.importance(kb, gj.C["y"].set(1.25), (0.3,))
    otr2, ow = ogfi.generate(o_model, okb, {"y": F32(1.25)}, (F32(0.3),))
    np.testing.assert_allclose(w.cpu().numpy(), ow, rtol=1e-5, atol=2e-5)
    ch2 = tr2.get_choices()
    lp_y = gj.normal.logpdf(torch.full((n,), 1.25, device=device), ch2["x"], 0.5 + ch2["s"])
    torch.testing.assert_close(w, lp_y, rtol=1e-5, atol=2e-5)
    # assess == score
    score, rv = model.assess(gj.vmap(lambda c: c, in_axes=0)(ch2), (0.3,))
    torch.testing.assert_close(score, tr2.get_score(), rtol=1e-6, atol=1e-6)
    with pytest.raises(gj.MissingAddress):
        model.assess(gj.C["x"].set(0.1), (0.3,))


def test_update_and_regenerate_weights(device):
    """test_static_gen_fn.py:623-667 (update weight = new - old density, discard = old values) and
    tests/inference/test_requests.py:37-61 (regenerate fwd/bwd weights cancel)."""
    gj = _gj()
    model, o_model = _model()
    n = 10_000
    kb = gj.split(gj.key(2), n)
    tr = model.simulate(kb, (0.0,))
    old = tr.get_choices()
    new_x = torch.linspace(-1, 1, n, device=device)
    tr2, w, retdiff, discard = model.update(gj.split(gj.key(3), n), tr, gj.vmap(lambda v: gj.C["x"].set(v), in_axes=0)(new_x))
    torch.testing.assert_close(w, tr2.get_score() - tr.get_score(), rtol=1e-4, atol=1e-4)
    assert torch.equal(discard["x"], old["x"]) and torch.equal(tr2.get_choices()["s"], old["s"])
    d_old = gj.normal.logpdf(old["x"], 0.0, 2.0) + gj.normal.logpdf(old["y"], old["x"], 0.5 + old["s"])
    d_new = gj.normal.logpdf(new_x, 0.0, 2.0) + gj.normal.logpdf(old["y"], new_x, 0.5 + old["s"])
    torch.testing.assert_close(w, d_new - d_old, rtol=1e-4, atol=1e-4)
    # regenerate x: forward weight, then the backward update cancels it
    req = gj.Regenerate(gj.S["x"])
    tr3, fwd, _, bwd = req.edit(gj.split(gj.key(4), n), tr, (0.0,))
    assert (tr3.get_choices()["x"] != old["x"]).all()
    torch.testing.assert_close(fwd, tr3.get_score() - tr.get_score(), rtol=1e-4, atol=1e-4)
    _, bwd_w, _, _ = bwd.edit(gj.split(gj.key(5), n), tr3, (0.0,))
    assert (fwd + bwd_w).abs().max().item() < 2e-4
    # EmptyRequest with unchanged args: zero weight, same trace
    tr4, w0, _, _ = gj.EmptyRequest().edit(gj.key(0), tr, gj.Diff.no_change((0.0,)))
    assert tr4 is tr and float(w0.abs().max()) == 0.0


def test_importance_k_reference_kats(device):
    """tests/inference/test_smc.py:32-87 on the GPU path."""
    gj = _gj()
    from genjax_b200.inference.smc import Importance, ImportanceK

    @gj.gen
    def flip_flip_trivial():
        gj.flip(0.5) @ "x"
        gj.flip(0.7) @ "y"

    g1 = GOLD["importance_k_flip"]
    target = gj.Target(flip_flip_trivial, (), gj.C["y"].set(True))
    z = ImportanceK(target, k_particles=g1["k"]).run_smc(gj.key(314159)).get_log_marginal_likelihood_estimate()
    assert z.item() == pytest.approx(g1["exact_logz"], rel=g1["rel"])
    z1 = Importance(target).run_smc(gj.key(314159)).get_log_marginal_likelihood_estimate()
    assert z1.item() == pytest.approx(g1["exact_logz"], rel=1e-3)

    @gj.gen
    def flip_flip():
        v1 = gj.flip(0.5) @ "x"
        p = gj.numpy.where(v1, 0.9, 0.3)
        gj.flip(p) @ "y"

    g2 = GOLD["importance_k_flip_flip"]
    target = gj.Target(flip_flip, (), gj.C["y"].set(True))
    pc = ImportanceK(target, k_particles=g2["k"]).run_smc(gj.key(314159))
    assert pc.get_log_marginal_likelihood_estimate().item() == pytest.approx(g2["exact_logz"], rel=g2["rel"])
    # bit-level parity with the oracle's ImportanceK (same key tree, same lanes)
    def o_ff(h):
        v1 = h.flip("x", F32(0.5))
        h.flip("y", np.where(v1 == 1, F32(0.9), F32(0.3)).astype(F32))

    opc = osmc.importance_k(o_ff, (), {"y": np.int32(1)}, orng.key(314159), g2["k"])
    np.testing.assert_allclose(pc.get_log_weights().cpu().numpy(), opc.log_weights, rtol=1e-5, atol=1e-6)
    assert pc.get_log_marginal_likelihood_estimate().item() == pytest.approx(opc.log_marginal_likelihood_estimate(), abs=1e-6)
    with pytest.raises(TypeError):
        gj.Target(ImportanceK(target, k_particles=2), (), gj.C.n())


def test_change_target_and_sample_particle(device):
    """smc.py:370-396 (reweight identity) and :102-109 (categorical draw of one particle)."""
    gj = _gj()
    from genjax_b200.inference.smc import ChangeTarget, ImportanceK

    @gj.gen
    def model():
        x = gj.normal(0.0, 2.0) @ "x"
        gj.normal(x, 1.0) @ "y"

    k = 4096
    t1 = gj.Target(model, (), gj.C["y"].set(1.0))
    t2 = gj.Target(model, (), gj.C["y"].set(2.0))
    alg = ImportanceK(t1, k_particles=k)
    pc1 = alg.run_smc(gj.key(7))
    same = ChangeTarget(alg, t1).run_smc(gj.key(7))
    torch.testing.assert_close(same.get_log_weights(), pc1.get_log_weights(), rtol=0, atol=3e-6)
    moved = ChangeTarget(alg, t2).run_smc(gj.key(7))
    x = pc1.get_particles().get_choices()["x"]
    torch.testing.assert_close(moved.get_log_weights(), gj.normal.logpdf(torch.full_like(x, 2.0), x, 1.0), rtol=1e-5, atol=1e-5)
    # sample_particle: index == the oracle's draw from the same weights; GenSP random_weighted runs end to end
    idx = pc1.sample_particle_index(gj.key(9)).item()
    assert idx == osmc.sample_particle_index(pc1.get_log_weights().cpu().numpy(), orng.key(9))
    w, chm = alg.random_weighted(gj.key(10), t1)
    assert "x" in chm and "y" not in chm and math.isfinite(w.item())
    # resample(): systematic ancestors bit-exact vs the oracle, weights reset to the log-mean-exp
    pc2, anc = pc1.resample(gj.key(12))
    assert np.array_equal(anc.cpu().numpy(), osmc.resample_systematic(pc1.get_log_weights().cpu().numpy(), orng.key(12)))
    assert pc2.get_log_weights()[0].item() == pytest.approx(osmc.log_mean_exp(pc1.get_log_weights().cpu().numpy()), abs=1e-6)


def test_readme_quickstart_config0(device):
    """BASELINE configs[0] / README.md:81-123: beta-bernoulli, ImportanceK k=50, 50 trials (SIR)."""
    gj = _gj()
    from genjax_b200.inference.smc import ImportanceK
    from genjax_b200.workloads import beta_bernoulli

    g = GOLD["readme_quickstart"]
    for obs, exact in ((True, g["exact_true"]), (False, g["exact_false"])):
        target = gj.Target(beta_bernoulli, (2.0, 2.0), gj.C["v"].set(obs))
        alg = ImportanceK(target, k_particles=50)
        ps = []
        keys = gj.split(gj.key(314159), 50)
        for t in range(50):
            _, chm = alg.random_weighted(keys[t], target)
            ps.append(chm["p"].item())
        se = 0.2 / math.sqrt(50)
        assert abs(np.mean(ps) - exact) < 4 * se


def test_custom_proposal_marginal_and_conditional_smc(device):
    """data-driven proposals ``ImportanceK(target, q=proposal.marginal())`` (smc.py:301-305), conditional
    SMC ``run_csmc`` (smc.py:268-279, 317-351, 398-425), ``Marginal`` (sp.py:208-252) and the GenSP density
    estimators.  With the EXACT posterior as proposal every importance weight equals log p(y): zero variance."""
    gj = _gj()
    from genjax_b200.inference.smc import ChangeTarget, Importance, ImportanceK

    @gj.gen
    def model():
        x = gj.normal(0.0, 1.0) @ "x"
        gj.normal(x, 1.0) @ "y"

    @gj.gen
    def proposal(target):
        y = target["y"]  # the proposal reads the observation out of the target (custom_proposal.ipynb)
        gj.normal(y / 2.0, math.sqrt(0.5)) @ "x"

    yv = 1.3
    target = gj.Target(model, (), gj.C["y"].set(yv))
    exact_logz = float(od.normal_logpdf(F32(yv), F32(0.0), F32(math.sqrt(2.0))))
    k = 512
    # the reference's own weight (sp.py:227-228: projection on the complement of the selection, 0 for a full
    # selection) is the default: ImportanceK's weights are then the joint score of each particle (smc.py:301-315)
    pc_ref = ImportanceK(target, q=proposal.marginal(), k_particles=k).run_smc(gj.key(1))
    x_ref = pc_ref.get_particles().get_choices()["x"]
    joint = gj.normal.logpdf(x_ref, 0.0, 1.0) + gj.normal.logpdf(torch.full_like(x_ref, yv), x_ref, 1.0)
    torch.testing.assert_close(pc_ref.get_log_weights(), joint, rtol=1e-5, atol=2e-5)
    # the corrected estimator is the opt-in: q's own density enters the weights
    alg = ImportanceK(target, q=proposal.marginal(reference_compat=False), k_particles=k)
    pc = alg.run_smc(gj.key(1))
    lw = pc.get_log_weights().cpu().numpy()
    assert lw.shape == (k,)
    np.testing.assert_allclose(lw, exact_logz, atol=2e-5)
    x = pc.get_particles().get_choices()["x"].cpu().numpy()
    assert abs(x.mean() - yv / 2) < 4 * math.sqrt(0.5 / k) and abs(x.std() - math.sqrt(0.5)) < 0.08
    assert alg.log_marginal_likelihood_estimate(gj.key(2)).item() == pytest.approx(exact_logz, abs=2e-5)
    # conditional SMC: K - 1 fresh particles + the retained one in the last slot
    retained = gj.C["x"].set(0.4)
    pc2 = alg.run_csmc(gj.key(3), retained)
    assert len(pc2) == k and pc2.get_particles().get_choices()["x"][-1].item() == pytest.approx(0.4)
    np.testing.assert_allclose(pc2.get_log_weights().cpu().numpy(), exact_logz, atol=2e-5)
    # GenSP density estimate of the retained value == the exact posterior density N(0.4; y/2, sqrt(1/2))
    alg.reference_compat = False
    est = alg.estimate_logpdf(gj.key(4), retained, target)
    assert est.item() == pytest.approx(float(od.normal_logpdf(F32(0.4), F32(yv / 2), F32(math.sqrt(0.5)))), abs=1e-4)
    # the default is smc.py:181-198 to the letter: the particle scored is ``sample_particle(sub_key)`` of the collection
    alg.reference_compat = True
    kb4 = gj.split(gj.key(4))
    pc_c = ChangeTarget(alg, target).run_csmc(kb4[0], retained)
    drawn = pc_c.sample_particle(kb4[1])
    est_c = alg.estimate_logpdf(gj.key(4), retained, target)
    assert est_c.item() == pytest.approx((drawn.get_score() - pc_c.get_log_marginal_likelihood_estimate()).item(), abs=1e-6)
    # without a proposal: prior particles, the retained one scored by the likelihood
    alg0 = ImportanceK(target, k_particles=64)
    pc3 = alg0.run_csmc(gj.key(5), retained)
    assert len(pc3) == 64 and pc3.get_particles().get_choices()["x"][-1].item() == pytest.approx(0.4)
    # (importance with x AND y constrained: both sites are weighted, smc.py:340-342)
    assert pc3.get_log_weights()[-1].item() == pytest.approx(
        float(od.normal_logpdf(F32(0.4), F32(0.0), F32(1.0)) + od.normal_logpdf(F32(yv), F32(0.4), F32(1.0))), abs=1e-5)
    one = Importance(target, q=proposal.marginal(reference_compat=False)).run_csmc(gj.key(6), retained)
    assert one.get_log_weights()[0].item() == pytest.approx(exact_logz, abs=2e-5)
    # ChangeTarget on top of conditional SMC keeps the retained particle and reweights everyone
    t2 = gj.Target(model, (), gj.C["y"].set(2.0))
    pc4 = ChangeTarget(alg, t2).run_csmc(gj.key(7), retained)
    x4 = pc4.get_particles().get_choices()["x"]
    assert x4[-1].item() == pytest.approx(0.4)
    want = exact_logz + (gj.normal.logpdf(torch.full_like(x4, 2.0), x4, 1.0) - gj.normal.logpdf(torch.full_like(x4, yv), x4, 1.0))
    torch.testing.assert_close(pc4.get_log_weights(), want, rtol=1e-4, atol=1e-4)
    # VI hooks (smc.py:204-230)
    assert alg.estimate_normalizing_constant(gj.key(8), target).item() == pytest.approx(exact_logz, abs=2e-5)
    w_ret = float(od.normal_logpdf(F32(0.4), F32(0), F32(1)) + od.normal_logpdf(F32(yv), F32(0.4), F32(1.0))
                  - od.normal_logpdf(F32(0.4), F32(yv / 2), F32(math.sqrt(0.5))))
    rz = alg.estimate_reciprocal_normalizing_constant(gj.key(9), target, retained, w_ret)
    score_ret = float(od.normal_logpdf(F32(0.4), F32(0), F32(1)) + od.normal_logpdf(F32(yv), F32(0.4), F32(1.0)))
    # smc.py:432-465: K - 1 reweighted particles (each exactly log p(y) here) + the retained one at w - score + weight
    last = w_ret - score_ret + exact_logz
    total = math.log(((k - 1) * math.exp(exact_logz) + math.exp(last)) / k)
    assert rz.item() == pytest.approx(score_ret - total, abs=1e-4)


def test_marginal_of_a_generative_function(device):
    """sp.py:208-252: random_weighted keeps the selected choices and weights by the projection on the rest;
    estimate_logpdf is the importance weight of the given choices."""
    gj = _gj()

    @gj.gen
    def model():
        x = gj.normal(0.0, 2.0) @ "x"
        gj.normal(x, 0.5) @ "y"

    n = 4096
    m = model.marginal(selection=gj.S["y"])
    assert m.reference_compat  # the reference's weight is the default
    kb = gj.split(gj.key(1), n)
    wc, chm = m.random_weighted(kb)
    assert "y" in chm and "x" not in chm and wc.shape == (n,)
    # sp.py:227-228: the weight is the projection on the COMPLEMENT of the selection, the prior density of the
    # discarded x: a Normal(0, 2) log-density, bounded above by -log(2) - 0.5 log(2 pi).  The opt-in corrected form
    # weights by the density of the kept choice y given the simulated x, a Normal(x, 0.5) log-density
    assert (wc <= -math.log(2.0) - 0.5 * math.log(2 * math.pi) + 1e-5).all()
    w, _ = model.marginal(selection=gj.S["y"], reference_compat=False).random_weighted(kb)
    assert (w <= -math.log(0.5) - 0.5 * math.log(2 * math.pi) + 1e-5).all() and not torch.equal(wc, w)
    yv = chm["y"]
    est = m.estimate_logpdf(gj.split(gj.key(2), n), gj.vmap(lambda v: gj.C["y"].set(v), in_axes=0)(yv))
    assert est.shape == (n,) and torch.isfinite(est).all()
    # importance weight of y given a fresh prior draw of x is log N(y; x', 0.5): on average exp(est) ~ p(y)
    marg = torch.exp(est.double()).mean().item()
    exact = torch.exp(gj.normal.logpdf(yv, 0.0, math.sqrt(4.25)).double()).mean().item()
    assert marg == pytest.approx(exact, rel=0.1)
    with pytest.raises(TypeError):
        gj.Target(m, (), gj.C.n())
    dec = gj.marginal(selection=gj.S["x"])(model)
    assert isinstance(dec, gj.Marginal)

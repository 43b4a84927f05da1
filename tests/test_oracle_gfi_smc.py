"""Oracle GFI / SMC restatement against the reference's algebraic identities and
statistical known answers."""
import json
import math
import os

import numpy as np
import pytest

from oracle import dists, gfi, rng
from oracle import smc as osmc

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_kats.json")))
F32 = np.float32


def linked(h):
    y1 = h.normal("y1", F32(0.0), F32(1.0))
    h.normal("y2", y1, F32(1.0))
    return y1


def test_importance_weight_equals_assess_score():
    """tests/generative_functions/test_distributions.py:38-40."""
    keys = rng.split(rng.key(314159), 64)
    v = np.linspace(-2, 2, 64).astype(F32)
    tr, w = gfi.generate(linked, keys, {"y2": v}, ())
    lp = dists.normal_logpdf(v, tr.choices["y1"], F32(1.0))
    assert np.array_equal(w, lp)
    full, _ = gfi.assess(linked, {"y1": tr.choices["y1"], "y2": v}, (), n=64)
    np.testing.assert_allclose(full, tr.get_score(), rtol=0, atol=0)


def test_missing_address_and_reuse():
    with pytest.raises(gfi.MissingAddress):
        gfi.assess(linked, {"y1": F32(0.0)}, ())

    def bad(h):
        h.normal("x", F32(0), F32(1))
        h.normal("x", F32(0), F32(1))

    with pytest.raises(gfi.AddressReuse):
        gfi.simulate(bad, rng.key(0), ())


def test_update_weight_identity():
    """tests/generative_functions/test_static_gen_fn.py:623-667: w = logp(new) - logp(old)."""
    keys = rng.split(rng.key(1), 32)
    tr = gfi.simulate(linked, keys, ())
    new_y1 = np.full(32, 0.25, dtype=F32)
    tr2, w, discard = gfi.update(linked, keys, tr, {"y1": new_y1})
    old = dists.normal_logpdf(tr.choices["y1"], 0, 1) + dists.normal_logpdf(tr.choices["y2"], tr.choices["y1"], 1)
    new = dists.normal_logpdf(new_y1, 0, 1) + dists.normal_logpdf(tr.choices["y2"], new_y1, 1)
    np.testing.assert_allclose(w, new - old, rtol=1e-5, atol=1e-5)
    assert np.array_equal(discard["y1"], tr.choices["y1"])
    assert np.array_equal(tr2.choices["y2"], tr.choices["y2"])


def test_regenerate_fwd_bwd_cancel():
    """tests/inference/test_requests.py:52-61: fwd_w == new - old density, fwd_w + bwd_w == 0."""
    keys = rng.split(rng.key(2), 32)
    tr = gfi.simulate(linked, keys, ())
    tr2, fwd, discard = gfi.regenerate(linked, rng.split(rng.key(3), 32), tr, {"y1"})
    assert not np.array_equal(tr2.choices["y1"], tr.choices["y1"])
    np.testing.assert_allclose(fwd, tr2.get_score() - tr.get_score(), rtol=1e-6, atol=1e-6)
    _, bwd, _ = gfi.update(linked, keys, tr2, discard)
    np.testing.assert_allclose(fwd + bwd, 0.0, atol=1e-5)


def test_importance_k_flip_kat():
    """tests/inference/test_smc.py:32-57."""
    g = GOLD["importance_k_flip"]

    def model(h):
        h.flip("x", F32(0.5))
        h.flip("y", F32(0.7))

    pc = osmc.importance_k(model, (), {"y": np.int32(1)}, rng.key(314159), g["k"])
    assert pc.log_marginal_likelihood_estimate() == pytest.approx(g["exact_logz"], rel=g["rel"])


def test_importance_k_flip_flip_kat():
    """tests/inference/test_smc.py:59-87."""
    g = GOLD["importance_k_flip_flip"]

    def model(h):
        v1 = h.flip("x", F32(0.5))
        p = np.where(v1 == 1, F32(0.9), F32(0.3)).astype(F32)
        h.flip("y", p)

    pc = osmc.importance_k(model, (), {"y": np.int32(1)}, rng.key(314159), g["k"])
    assert pc.log_marginal_likelihood_estimate() == pytest.approx(g["exact_logz"], rel=g["rel"])


def test_readme_quickstart_band():
    """README.md:81-123: 50 SIR trials of ImportanceK(k=50) on beta(2,2)/flip."""
    g = GOLD["readme_quickstart"]

    def model(h, a, b):
        p = h.beta("p", a, b)
        return h.flip("v", p)

    for obs, exact, pub in ((1, g["exact_true"], g["obs_true"]), (0, g["exact_false"], g["obs_false"])):
        ps = []
        keys = rng.split(rng.key(314159), 50)
        for t in range(50):
            kb = rng.split(keys[t])
            pc = osmc.importance_k(model, (F32(2.0), F32(2.0)), {"v": np.int32(obs)}, kb[0], 50)
            i = osmc.sample_particle_index(pc.log_weights, kb[1])
            ps.append(pc.trace.choices["p"][i])
        m = float(np.mean(ps))
        se = 0.2 / math.sqrt(50)  # posterior sd = 0.2
        assert abs(m - exact) < 4 * se
        assert abs(pub - exact) < 4 * se  # the published value sits in the same band


def test_det_exp_q_accuracy_and_monotone():
    x = -np.abs(np.random.default_rng(0).standard_normal(100_000).astype(F32)) * 10
    q = osmc.det_exp_q(x).astype(np.float64) / 2.0**36
    np.testing.assert_allclose(q, np.exp(x.astype(np.float64)), rtol=3e-6, atol=2.0**-36)
    assert osmc.det_exp_q(F32(0.0)) == np.uint64(1) << np.uint64(36) or abs(int(osmc.det_exp_q(F32(0.0))) - 2**36) < 2**14
    xs = np.sort(x)
    assert np.all(np.diff(osmc.det_exp_q(xs).astype(np.int64)) >= -2**13)  # monotone up to rounding of the polynomial
    assert osmc.det_exp_q(F32(-np.inf)) == 0 and osmc.det_exp_q(F32(np.nan)) == 0 and osmc.det_exp_q(F32(-200.0)) == 0


def test_log_mean_exp_matches_float64():
    lw = (np.random.default_rng(1).standard_normal(10_000) * 5).astype(F32)
    ref = np.log(np.mean(np.exp(lw.astype(np.float64))))
    assert osmc.log_mean_exp(lw) == pytest.approx(ref, abs=1e-5)
    assert osmc.log_mean_exp(np.full(7, -np.inf, dtype=F32)) == -math.inf


@pytest.mark.parametrize("n", [1, 2, 17, 2048, 5001])
def test_systematic_properties(n):
    g = np.random.default_rng(n)
    lw = (g.standard_normal(n) * 2).astype(F32)
    anc = osmc.resample_systematic(lw, rng.key(5))
    assert anc.shape == (n,) and anc.dtype == np.int32
    assert np.all(np.diff(anc) >= 0)  # sorted ancestors
    w = np.exp(lw.astype(np.float64))
    w /= w.sum()
    cnt = np.bincount(anc, minlength=n)
    assert np.all(np.abs(cnt - n * w) < 1.0 + 1e-3 * n * w + 1e-6)  # floor/ceil of the expected count
    # shard-combination: two halves with the global (M, S, offset) reproduce the same counts
    if n >= 2:
        M, S = osmc.lse_terms(lw)
        u0 = osmc.resample_u0(rng.key(5))
        h = n // 2
        q = osmc.det_exp_q((lw[:h] - M).astype(F32))
        c0, _ = osmc.systematic_counts(lw[:h], u0, n_out=n, M=M, S=S, c_offset=0)
        c1, _ = osmc.systematic_counts(lw[h:], u0, n_out=n, M=M, S=S, c_offset=int(q.sum(dtype=np.uint64)))
        full, _ = osmc.systematic_counts(lw, u0)
        assert np.array_equal(np.concatenate([c0, c1]), full)


def test_systematic_all_zero_weights_identity():
    anc = osmc.resample_systematic(np.full(9, -np.inf, dtype=F32), rng.key(0))
    assert np.array_equal(anc, np.arange(9))


def test_multinomial_frequencies():
    n = 20_000
    lw = np.log(np.tile(np.array([0.1, 0.4, 0.2, 0.3], dtype=np.float64), n // 4)).astype(F32)
    anc = osmc.resample_multinomial(lw, rng.split(rng.key(8), n))
    f = np.bincount(anc % 4, minlength=4) / n
    np.testing.assert_allclose(f, [0.1, 0.4, 0.2, 0.3], atol=0.012)


def test_bootstrap_pf_matches_kalman():
    a, q, c, r = 0.9, 1.0, 1.0, 0.5
    T, n = 25, 20_000
    ys = osmc.simulate_lgssm(0, T, 1, a, q, c, r)[:, 0]

    def step(h, x_prev):
        x = h.normal("x", F32(a) * x_prev, F32(q))
        h.normal("y", F32(c) * x, F32(r))
        return x

    x0 = np.random.default_rng(3).standard_normal(n).astype(F32)
    out = osmc.particle_filter(step, rng.key(314159), x0, [{"y": F32(y)} for y in ys])
    exact = osmc.kalman_logz(ys, a, q, c, r)
    assert out["logz"] == pytest.approx(exact, abs=0.25)


def test_change_target_reweight_identity():
    """smc.py:378-384: w' = new_score - old_score + w; unchanged target => w' == w (fp32 noise)."""

    def model(h):
        x = h.normal("x", F32(0.0), F32(2.0))
        h.normal("y", x, F32(1.0))

    pc = osmc.importance_k(model, (), {"y": F32(1.0)}, rng.key(7), 256)
    pc2 = osmc.change_target(pc, rng.key(7), model, (), {"y": F32(1.0)}, {"y"})
    np.testing.assert_allclose(pc2.log_weights, pc.log_weights, atol=2e-6)
    pc3 = osmc.change_target(pc, rng.key(7), model, (), {"y": F32(2.0)}, {"y"})
    x = pc.trace.choices["x"]
    np.testing.assert_allclose(pc3.log_weights, dists.normal_logpdf(F32(2.0), x, F32(1.0)), atol=1e-5)

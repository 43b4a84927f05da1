"""Oracle distributions: the reference KAT for Normal, and every restated TFP
log_prob against an independent float64 implementation (scipy.stats)."""
import json
import math
import os

import numpy as np
import pytest
from scipy import stats

from oracle import dists, gfi, rng

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_kats.json")))
RNG = np.random.default_rng(0)


def test_reference_kat_assess_two_normals():
    """tests/generative_functions/test_static_gen_fn.py:317-318."""
    g = GOLD["assess_two_normals"]

    def model(h):
        y1 = h.normal("y1", np.float32(0.0), np.float32(1.0))
        y2 = h.normal("y2", np.float32(0.0), np.float32(1.0))
        return y1 + y2

    score, rv = gfi.assess(model, {k: np.float32(v) for k, v in g["choices"].items()}, ())
    assert score[0] == pytest.approx(g["score"], abs=1e-6)
    assert float(rv) == 0.0


def _close(a, b, rtol=2e-6, atol=2e-6):
    np.testing.assert_allclose(np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64), rtol=rtol, atol=atol)


def test_logpdfs_against_scipy():
    v = RNG.standard_normal(1000).astype(np.float32) * 3
    loc = RNG.standard_normal(1000).astype(np.float32)
    sc = (0.1 + RNG.random(1000) * 4).astype(np.float32)
    _close(dists.normal_logpdf(v, loc, sc), stats.norm.logpdf(v.astype(np.float64), loc, sc), rtol=1e-5, atol=1e-5)
    u = RNG.random(1000).astype(np.float32)
    _close(dists.uniform_logpdf(u * 3 - 1, -1.0, 2.0), np.full(1000, -math.log(3.0)))
    assert dists.uniform_logpdf(np.float32(2.5), -1.0, 2.0) == -np.inf
    p = (0.01 + 0.98 * RNG.random(1000)).astype(np.float32)
    x = (RNG.random(1000) < 0.5).astype(np.int32)
    _close(dists.flip_logpdf(x, p), stats.bernoulli.logpmf(x, p.astype(np.float64)), rtol=1e-5)
    lg = RNG.standard_normal(1000).astype(np.float32) * 3
    _close(dists.bernoulli_logpdf(x, lg), stats.bernoulli.logpmf(x, 1 / (1 + np.exp(-lg.astype(np.float64)))), rtol=1e-5, atol=1e-6)
    r = (0.1 + 5 * RNG.random(1000)).astype(np.float32)
    e = (RNG.exponential(1.0, 1000) / r).astype(np.float32)
    _close(dists.exponential_logpdf(e, r), stats.expon.logpdf(e.astype(np.float64), scale=1 / r.astype(np.float64)), rtol=1e-5, atol=1e-5)
    a = (0.2 + 5 * RNG.random(1000)).astype(np.float32)
    gv = (RNG.gamma(a) / r).astype(np.float32) + np.float32(1e-3)
    _close(dists.gamma_logpdf(gv, a, r), stats.gamma.logpdf(gv.astype(np.float64), a.astype(np.float64), scale=1 / r.astype(np.float64)), rtol=2e-5, atol=2e-5)
    b = (0.2 + 5 * RNG.random(1000)).astype(np.float32)
    bv = np.clip(RNG.beta(a, b), 1e-3, 1 - 1e-3).astype(np.float32)
    _close(dists.beta_logpdf(bv, a, b), stats.beta.logpdf(bv.astype(np.float64), a.astype(np.float64), b.astype(np.float64)), rtol=2e-5, atol=2e-5)
    hv = np.abs(v)
    _close(dists.half_normal_logpdf(hv, sc), stats.halfnorm.logpdf(hv.astype(np.float64), scale=sc.astype(np.float64)), rtol=1e-5, atol=1e-5)
    assert dists.half_normal_logpdf(np.float32(-1.0), 1.0) == -np.inf


def test_long_tail_scalar_logpdfs_match_scipy():
    """cauchy, half_cauchy, laplace, log_normal, gumbel, weibull (tensorflow_probability/__init__.py:110-309) against
    scipy's float64 densities."""
    n = 2000
    v = (RNG.standard_normal(n) * 3).astype(np.float32)
    loc = RNG.standard_normal(n).astype(np.float32)
    sc = (0.3 + 2 * RNG.random(n)).astype(np.float32)
    k = (0.5 + 3 * RNG.random(n)).astype(np.float32)
    pv = np.exp(v / 3).astype(np.float32)
    hv = (loc + np.abs(v)).astype(np.float32)
    f = lambda x: x.astype(np.float64)  # noqa: E731
    _close(dists.cauchy_logpdf(v, loc, sc), stats.cauchy.logpdf(f(v), f(loc), f(sc)), rtol=1e-5, atol=1e-5)
    _close(dists.half_cauchy_logpdf(hv, loc, sc), stats.halfcauchy.logpdf(f(hv), f(loc), f(sc)), rtol=1e-5, atol=1e-5)
    _close(dists.laplace_logpdf(v, loc, sc), stats.laplace.logpdf(f(v), f(loc), f(sc)), rtol=1e-5, atol=1e-5)
    _close(dists.log_normal_logpdf(pv, loc, sc), stats.lognorm.logpdf(f(pv), f(sc), scale=np.exp(f(loc))), rtol=1e-5, atol=1e-5)
    _close(dists.gumbel_logpdf(v, loc, sc), stats.gumbel_r.logpdf(f(v), f(loc), f(sc)), rtol=1e-5, atol=1e-5)
    _close(dists.weibull_logpdf(pv, k, sc), stats.weibull_min.logpdf(f(pv), f(k), scale=f(sc)), rtol=1e-5, atol=1e-5)
    for fn, args in ((dists.half_cauchy_logpdf, (0.0, 1.0)), (dists.log_normal_logpdf, (0.0, 1.0)), (dists.weibull_logpdf, (1.5, 1.0))):
        assert fn(np.float32(-1.0), *args) == -np.inf  # outside the support


def test_second_slice_logpdfs_match_scipy():
    """kumaraswamy, logit_normal, geometric, inverse_gamma, chi2 (tensorflow_probability/__init__.py:204, 224, 169, 194,
    120) against float64 closed forms / scipy; the geometric sampler's frequencies."""
    n = 2000
    f = lambda x: x.astype(np.float64)  # noqa: E731
    a = (0.4 + 3 * RNG.random(n)).astype(np.float32)
    b = (0.4 + 3 * RNG.random(n)).astype(np.float32)
    unit = (0.01 + 0.98 * RNG.random(n)).astype(np.float32)
    pos = (0.05 + 4 * RNG.random(n)).astype(np.float32)
    loc = RNG.standard_normal(n).astype(np.float32)
    kref = np.log(f(a)) + np.log(f(b)) + (f(a) - 1) * np.log(f(unit)) + (f(b) - 1) * np.log1p(-f(unit) ** f(a))
    _close(dists.kumaraswamy_logpdf(unit, a, b), kref, rtol=1e-5, atol=1e-5)
    lref = stats.norm.logpdf(np.log(f(unit)) - np.log1p(-f(unit)), f(loc), f(b)) - np.log(f(unit)) - np.log1p(-f(unit))
    _close(dists.logit_normal_logpdf(unit, loc, b), lref, rtol=1e-5, atol=1e-5)
    k = RNG.integers(0, 20, n).astype(np.float32)
    _close(dists.geometric_logpdf(k, unit), stats.geom.logpmf(f(k) + 1, f(unit)), rtol=1e-5, atol=1e-5)  # scipy counts trials
    _close(dists.inverse_gamma_logpdf(pos, a, b), stats.invgamma.logpdf(f(pos), f(a), scale=f(b)), rtol=1e-5, atol=1e-5)
    _close(dists.chi2_logpdf(pos, 2 * a), stats.chi2.logpdf(f(pos), f(2 * a)), rtol=1e-5, atol=1e-5)
    tv = (3 * RNG.standard_normal(n)).astype(np.float32)
    _close(dists.student_t_logpdf(tv, 2 * a, loc, b), stats.t.logpdf(f(tv), f(2 * a), f(loc), f(b)), rtol=1e-5, atol=1e-5)
    for fn, v, args in ((dists.kumaraswamy_logpdf, 1.5, (2.0, 2.0)), (dists.logit_normal_logpdf, 1.0, (0.0, 1.0)),
                        (dists.geometric_logpdf, -1.0, (0.3,)), (dists.inverse_gamma_logpdf, 0.0, (2.0, 1.0))):
        assert fn(np.float32(v), *args) == -np.inf
    x = dists.geometric_sample((11, 22), np.arange(100_000, dtype=np.uint64), 1, np.float32(0.3))
    assert x.dtype == np.float32 and x.min() == 0 and (x == np.floor(x)).all()
    np.testing.assert_allclose(np.bincount(x.astype(int))[:4] / x.size, 0.3 * 0.7 ** np.arange(4), atol=5e-3)


@pytest.mark.parametrize("rate", [0.3, 4.0, 9.9, 10.0, 37.5, 400.0])
def test_poisson_sampler_and_logpmf(rate):
    """tfd.Poisson (tensorflow_probability/__init__.py:264): inversion below rate 10, PTRS above; chi-square of the
    draws against the exact pmf, log-pmf against scipy."""
    n = 100_000
    x = dists.poisson_sample((11, 22), np.arange(n, dtype=np.uint64), 2, np.float32(rate))
    assert x.dtype == np.float32 and (x == np.floor(x)).all() and x.min() >= 0
    m = int(x.max()) + 1
    obs = np.bincount(x.astype(int), minlength=m)
    exp = stats.poisson.pmf(np.arange(m), rate) * n
    keep = exp > 5
    chi = ((obs[keep] - exp[keep]) ** 2 / exp[keep]).sum()
    assert stats.chi2.sf(chi, keep.sum() - 1) > 1e-3
    k = RNG.integers(0, int(3 * rate) + 5, 500).astype(np.float32)
    # float32 cancellation of k log(rate) - lgamma(k + 1) (TFP's own formula): a few ulps of the larger term
    _close(dists.poisson_logpdf(k, rate), stats.poisson.logpmf(k.astype(np.float64), rate), rtol=2e-5,
           atol=2e-5 + 2e-6 * rate * max(1.0, math.log(rate)))
    assert dists.poisson_logpdf(np.float32(-1.0), rate) == -np.inf


def test_categorical_and_mvn_logpdf():
    logits = RNG.standard_normal((50, 16)).astype(np.float32) * 2
    k = RNG.integers(0, 16, 50)
    ref = logits.astype(np.float64) - np.log(np.exp(logits.astype(np.float64)).sum(-1, keepdims=True))
    _close(dists.categorical_logpdf(k, logits), ref[np.arange(50), k], rtol=1e-5, atol=1e-5)
    v = RNG.standard_normal((100, 8)).astype(np.float32)
    loc = RNG.standard_normal(8).astype(np.float32)
    sc = (0.5 + RNG.random(8)).astype(np.float32)
    ref = stats.norm.logpdf(v.astype(np.float64), loc, sc).sum(-1)
    _close(dists.mv_normal_diag_logpdf(v, loc, sc), ref, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize(
    "name,args,cdf",
    [
        ("normal", (1.5, 2.0), lambda x: stats.norm.cdf(x, 1.5, 2.0)),
        ("uniform", (-1.0, 3.0), lambda x: stats.uniform.cdf(x, -1.0, 4.0)),
        ("exponential", (2.5,), lambda x: stats.expon.cdf(x, scale=1 / 2.5)),
        ("half_normal", (1.7,), lambda x: stats.halfnorm.cdf(x, scale=1.7)),
        ("gamma", (2.5, 1.5), lambda x: stats.gamma.cdf(x, 2.5, scale=1 / 1.5)),
        ("gamma", (0.4, 1.0), lambda x: stats.gamma.cdf(x, 0.4)),
        ("beta", (2.0, 2.0), lambda x: stats.beta.cdf(x, 2.0, 2.0)),
        ("beta", (0.5, 3.0), lambda x: stats.beta.cdf(x, 0.5, 3.0)),
        ("cauchy", (1.0, 2.0), lambda x: stats.cauchy.cdf(x, 1.0, 2.0)),
        ("half_cauchy", (1.0, 2.0), lambda x: stats.halfcauchy.cdf(x, 1.0, 2.0)),
        ("laplace", (-1.0, 0.5), lambda x: stats.laplace.cdf(x, -1.0, 0.5)),
        ("log_normal", (0.3, 0.8), lambda x: stats.lognorm.cdf(x, 0.8, scale=math.exp(0.3))),
        ("gumbel", (0.5, 1.5), lambda x: stats.gumbel_r.cdf(x, 0.5, 1.5)),
        ("weibull", (1.7, 2.0), lambda x: stats.weibull_min.cdf(x, 1.7, scale=2.0)),
        ("kumaraswamy", (2.0, 3.0), lambda x: 1 - (1 - x**2.0) ** 3.0),
        ("logit_normal", (0.3, 0.8), lambda x: stats.norm.cdf(np.log(x) - np.log1p(-x), 0.3, 0.8)),
        ("inverse_gamma", (3.0, 2.0), lambda x: stats.invgamma.cdf(x, 3.0, scale=2.0)),
        ("chi2", (3.5,), lambda x: stats.chi2.cdf(x, 3.5)),
        ("chi2", (0.8,), lambda x: stats.chi2.cdf(x, 0.8)),
        ("student_t", (4.0, 1.0, 2.0), lambda x: stats.t.cdf(x, 4.0, 1.0, 2.0)),
        ("student_t", (0.9, 0.0, 1.0), lambda x: stats.t.cdf(x, 0.9)),
    ],
)
def test_continuous_samplers_ks(name, args, cdf):
    n = 40_000
    idx = np.arange(n, dtype=np.uint64)
    x = dists.DISTS[name][0]((11, 22), idx, 1, *[np.float32(a) for a in args])
    assert x.dtype == np.float32 and x.shape == (n,)
    d, p = stats.kstest(x.astype(np.float64), cdf)
    assert p > 1e-3, (name, d, p)


def test_discrete_samplers_frequencies():
    n = 100_000
    idx = np.arange(n, dtype=np.uint64)
    x = dists.flip_sample((1, 2), idx, 1, np.float32(0.3))
    assert abs(x.mean() - 0.3) < 5e-3
    x = dists.bernoulli_sample((1, 2), idx, 2, np.float32(-1.0))
    assert abs(x.mean() - 1 / (1 + math.e)) < 5e-3
    logits = np.log(np.array([0.1, 0.2, 0.3, 0.4], dtype=np.float32))
    k = dists.categorical_sample((1, 2), idx, 3, logits)
    f = np.bincount(k, minlength=4) / n
    np.testing.assert_allclose(f, [0.1, 0.2, 0.3, 0.4], atol=6e-3)
    # per-particle logits rows
    rows = np.tile(logits, (n, 1))
    k2 = dists.categorical_sample((1, 2), idx, 3, rows)
    assert np.array_equal(k, k2)


def test_mv_normal_logpdf_and_sampler_against_scipy():
    """tfd.MultivariateNormalFullCovariance restated through a float32 Cholesky + forward substitution."""
    g = np.random.default_rng(0)
    for d in (2, 6, 16):
        A = g.standard_normal((d, d))
        cov = (A @ A.T + d * np.eye(d)).astype(np.float32)
        loc = g.standard_normal(d).astype(np.float32)
        v = (g.standard_normal((200, d)) * 2).astype(np.float32)
        ref = stats.multivariate_normal.logpdf(v.astype(np.float64), loc.astype(np.float64), cov.astype(np.float64))
        _close(dists.mv_normal_logpdf(v, loc, cov), ref, rtol=2e-5, atol=2e-5)
    x = dists.mv_normal_sample((1, 2), np.arange(200_000, dtype=np.uint64), 1, loc, cov)
    assert np.abs(np.cov(x.T) - cov).max() < 0.02 * np.abs(cov).max() + 0.15
    assert np.abs(x.mean(0) - loc).max() < 0.05

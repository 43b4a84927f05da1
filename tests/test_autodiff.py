"""gen/autodiff.py on the CPU: symbolic log-densities equal the oracle's, and the
reverse-mode gradients equal central finite differences (float64 interpreter)."""
import numpy as np
import pytest

from genjax_b200.gen import autodiff as AD
from genjax_b200.gen import capture as cap
from genjax_b200.gen.capture import ArgSpec
from genjax_b200.workloads import EIGHT_SCHOOLS_SIGMA, EIGHT_SCHOOLS_Y, eight_schools, gmm_target
from oracle import dists
from expr_eval import evaluate


def _fd(f, x, h=1e-6):
    g = np.zeros_like(x)
    for k in range(x.size):
        d = np.zeros_like(x)
        d.flat[k] = h
        g.flat[k] = (f(x + d) - f(x - d)) / (2 * h)
    return g


def test_eight_schools_logp_and_grad():
    ir = cap.capture(eight_schools.source, "es", [ArgSpec("shared", "f32", (8,))], ("tuple", [("leaf", 0)]))
    logp = AD.model_logp(ir)
    vals = [s.value for s in ir.sites[:3]]
    grads = AD.grad(logp, vals)
    g = np.random.default_rng(0)
    mu, lt, th = g.normal(), 0.3 * g.normal(), g.normal(size=8) * 3
    y, sig = np.array(EIGHT_SCHOOLS_Y), np.array(EIGHT_SCHOOLS_SIGMA)

    def env(mu, lt, th):
        return {0: mu, 1: lt, 2: th, 3: y, ("arg", 0): sig}

    got = float(evaluate(logp, env(mu, lt, th)))
    want = (dists.normal_logpdf(mu, 0, 5).astype(np.float64) + dists.normal_logpdf(lt, 0, 1)
            + dists.mv_normal_diag_logpdf(th[None], np.full(8, mu), np.full(8, np.exp(lt)))[0]
            + dists.mv_normal_diag_logpdf(y[None], th, sig)[0])
    assert got == pytest.approx(float(want), rel=1e-5)
    e = env(mu, lt, th)
    cache = {}
    gm, gl, gt = (np.asarray(evaluate(x, e, cache), dtype=np.float64) for x in grads)
    assert gm == pytest.approx(_fd(lambda v: float(evaluate(logp, env(v[0], lt, th))), np.array([mu]))[0], rel=1e-5)
    assert gl == pytest.approx(_fd(lambda v: float(evaluate(logp, env(mu, v[0], th))), np.array([lt]))[0], rel=1e-5)
    np.testing.assert_allclose(gt, _fd(lambda v: float(evaluate(logp, env(mu, lt, v))), th), rtol=1e-5, atol=1e-7)


def test_gmm_logp_and_grad():
    K, D = 8, 8
    specs = [ArgSpec("shared", "f32", (K,)), ArgSpec("shared", "f32", (K, D)), ArgSpec("shared", "f32", (K,))]
    ir = cap.capture(gmm_target.source, "gmm", specs, ("tuple", [("leaf", i) for i in range(3)]))
    logp = AD.model_logp(ir)
    (gx,) = AD.grad(logp, [ir.sites[0].value])
    g = np.random.default_rng(1)
    logits = g.normal(size=K)
    mu = g.uniform(-4, 4, size=(K, D))
    sigma = 0.5 + g.random(K)
    x = mu[3] + 0.4 * g.normal(size=D)

    def env(x):
        return {0: x, ("arg", 0): logits, ("arg", 1): mu, ("arg", 2): sigma}

    def ref(x):
        lw = logits - np.log(np.sum(np.exp(logits)))
        comp = lw + np.sum(-0.5 * ((x - mu) / sigma[:, None]) ** 2 - np.log(sigma[:, None]) - 0.5 * np.log(2 * np.pi), axis=1)
        return np.log(np.sum(np.exp(comp)))

    assert float(evaluate(logp, env(x))) == pytest.approx(ref(x), rel=1e-9)
    np.testing.assert_allclose(np.asarray(evaluate(gx, env(x))), _fd(ref, x), rtol=1e-5, atol=1e-7)


def test_grad_rules_elementwise():
    from genjax_b200.gen import expr as E

    x = E.Expr("site", (), "f32", (), 0)
    f = (E.unary("tanh", x) * E.unary("exp", -x) + E.unary("softplus", x) / (1.0 + E.unary("square", x))
         + E.unary("log1p", E.unary("sigmoid", x)) - E.unary("sqrt", 2.0 + E.unary("cos", x)) + E.binary("pow", x, 3.0)
         + E.binary("max", x, 0.5) * E.binary("min", x, 2.0) + E.where(x > 0.2, x * x, -x) + abs(x))
    (g,) = AD.grad(f, [x])
    for v in (-1.3, 0.4, 1.7):
        got = float(evaluate(g, {0: v}))
        want = _fd(lambda t: float(evaluate(f, {0: t[0]})), np.array([v]))[0]
        assert got == pytest.approx(want, rel=1e-5, abs=1e-7)


def test_long_tail_logp_and_grad():
    """cauchy, half_cauchy, laplace, log_normal, gumbel, weibull: symbolic log-densities equal the oracle's, gradients
    (what HMC / the analytic MH ratio consume) equal central finite differences."""
    import genjax_b200 as gj

    def model(s):
        a = gj.cauchy(0.0, s) @ "a"
        b = gj.half_cauchy(a, 1.5) @ "b"
        c = gj.laplace(a, s) @ "c"
        d = gj.log_normal(0.1 * c, 0.5) @ "d"
        e = gj.gumbel(c, d) @ "e"
        f = gj.weibull(1.0 + d, s) @ "f"
        return e + f

    ir = cap.capture(model, "lt", [ArgSpec("scalar", "f32", ())], ("tuple", [("leaf", 0)]))
    logp = AD.model_logp(ir)
    vals = [s.value for s in ir.sites]
    grads = AD.grad(logp, vals)
    s = 0.7
    x = np.array([0.4, 1.3, -0.2, 0.8, 0.1, 1.7])  # a, b >= a, c, d > 0, e, f >= 0

    def env(x):
        return {**{j: x[j] for j in range(6)}, ("arg", 0): s}

    a, b, c, d, e, f = x
    want = (dists.cauchy_logpdf(a, 0, s).astype(np.float64) + dists.half_cauchy_logpdf(b, a, 1.5) + dists.laplace_logpdf(c, a, s)
            + dists.log_normal_logpdf(d, 0.1 * c, 0.5) + dists.gumbel_logpdf(e, c, d) + dists.weibull_logpdf(f, 1.0 + d, s))
    assert float(evaluate(logp, env(x))) == pytest.approx(float(want), rel=2e-6)
    cache = {}
    got = np.array([float(evaluate(g, env(x), cache)) for g in grads])
    np.testing.assert_allclose(got, _fd(lambda v: float(evaluate(logp, env(v))), x), rtol=1e-5, atol=1e-7)
    # outside the support: -inf, as the device structs return (d feeds the scales of e and f, so it is left alone)
    for j, bad in ((1, 0.0), (5, -1.0)):
        y = x.copy()
        y[j] = bad
        assert float(evaluate(logp, env(y))) == -np.inf


def test_second_slice_logp_and_grad():
    """kumaraswamy, logit_normal, geometric, inverse_gamma, chi2: symbolic log-densities against the oracle, gradients
    in the continuous sites against central finite differences."""
    import genjax_b200 as gj

    def model(s):
        a = gj.kumaraswamy(2.0, s) @ "a"
        b = gj.logit_normal(a, 0.5) @ "b"
        k = gj.geometric(0.1 + 0.8 * b) @ "k"
        g = gj.inverse_gamma(2.0 + k, s + a) @ "g"
        c = gj.chi2(1.0 + s) @ "c"  # a df that depends on a continuous site would need digamma
        return g + c

    ir = cap.capture(model, "s2", [ArgSpec("scalar", "f32", ())], ("tuple", [("leaf", 0)]))
    logp = AD.model_logp(ir)
    cont = [0, 1, 3, 4]
    grads = AD.grad(logp, [ir.sites[j].value for j in cont])
    s, k = 1.5, 3.0
    x = np.array([0.35, 0.6, 0.9, 2.5])

    def env(x):
        return {0: x[0], 1: x[1], 2: k, 3: x[2], 4: x[3], ("arg", 0): s}

    a, b, g, c = x
    want = (dists.kumaraswamy_logpdf(a, 2.0, s).astype(np.float64) + dists.logit_normal_logpdf(b, a, 0.5)
            + dists.geometric_logpdf(k, 0.1 + 0.8 * b) + dists.inverse_gamma_logpdf(g, 2.0 + k, s + a) + dists.chi2_logpdf(c, 1.0 + s))
    assert float(evaluate(logp, env(x))) == pytest.approx(float(want), rel=2e-6)
    cache = {}
    got = np.array([float(evaluate(gr, env(x), cache)) for gr in grads])
    np.testing.assert_allclose(got, _fd(lambda v: float(evaluate(logp, env(v))), x), rtol=1e-5, atol=1e-7)


def test_student_t_logp_and_grad():
    import genjax_b200 as gj

    def model(s):
        m = gj.normal(0.0, 2.0) @ "m"
        t = gj.student_t(3.0 + s, m, s) @ "t"
        return t

    ir = cap.capture(model, "st", [ArgSpec("scalar", "f32", ())], ("tuple", [("leaf", 0)]))
    logp = AD.model_logp(ir)
    grads = AD.grad(logp, [s_.value for s_ in ir.sites])
    s, x = 0.7, np.array([0.4, -1.3])
    env = lambda x: {0: x[0], 1: x[1], ("arg", 0): s}  # noqa: E731
    want = dists.normal_logpdf(x[0], 0.0, 2.0).astype(np.float64) + dists.student_t_logpdf(x[1], 3.0 + s, x[0], s)
    assert float(evaluate(logp, env(x))) == pytest.approx(float(want), rel=2e-6)
    cache = {}
    got = np.array([float(evaluate(g, env(x), cache)) for g in grads])
    np.testing.assert_allclose(got, _fd(lambda v: float(evaluate(logp, env(v))), x), rtol=1e-5, atol=1e-7)

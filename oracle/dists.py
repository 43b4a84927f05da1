"""Per-primitive sample / logpdf restatement (oracle; test infrastructure only).

Reference call sites: ``ExactDensity.random_weighted / estimate_logpdf``
(src/genjax/_src/generative_functions/distributions/distribution.py:371-396,
non-scalar logpdf is sum-reduced at :393-394) which call
``tfd.<Dist>(*args).sample(seed=key)`` / ``.log_prob(v)``
(distributions/tensorflow_probability/__init__.py:52-62).

tensorflow-probability 0.23.0 (poetry.lock:5015-5016) is NOT vendored and not
installable here; the log_prob formulas below restate its published source in
TFP's float32 operation order.  PINNED: Normal (reference KAT
tests/generative_functions/test_static_gen_fn.py:317-318).  UNPINNED: every
other logpdf and every sampler ("parity unpinned").  Samplers are this build's
own bits->variate maps (see oracle/rng.py) -- same distributions as TFP's,
different streams.

Every function is vectorised over a leading particle axis and computes in
float32 with float64 only inside transcendental calls (rounded once).
"""

from __future__ import annotations

import math

import numpy as np

from . import rng

F32 = np.float32
_HALF_LOG_2PI = F32(0.5 * math.log(2.0 * math.pi))


def _f(x):
    return np.asarray(x, dtype=F32)


def _log(x):
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.log(_f(x).astype(np.float64)).astype(F32)


def _exp(x):
    with np.errstate(over="ignore"):
        return np.exp(_f(x).astype(np.float64)).astype(F32)


def _log1p(x):
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.log1p(_f(x).astype(np.float64)).astype(F32)


def _softplus(x):
    x = _f(x).astype(np.float64)
    return (np.maximum(x, 0.0) + np.log1p(np.exp(-np.abs(x)))).astype(F32)


def _lgamma(x):
    from scipy.special import gammaln

    return gammaln(_f(x).astype(np.float64)).astype(F32)


# ----------------------------------------------------------------- logpdfs


def normal_logpdf(v, loc, scale):
    """tfd.Normal._log_prob: -0.5*squared_difference(x/s, m/s) - (0.5 log 2pi + log s)."""
    v, loc, scale = _f(v), _f(loc), _f(scale)
    z = (v / scale - loc / scale).astype(F32)
    return (F32(-0.5) * (z * z) - (_HALF_LOG_2PI + _log(scale))).astype(F32)


def uniform_logpdf(v, low, high):
    """tfd.Uniform._log_prob: -log(high-low) inside [low, high], -inf outside."""
    v, low, high = _f(v), _f(low), _f(high)
    inside = (v >= low) & (v <= high)
    return np.where(inside, -_log(high - low), F32(-np.inf)).astype(F32)


def flip_logpdf(v, p):
    """tfd.Bernoulli(probs=p)._log_prob: xlogy(x, p) + xlog1py(1-x, -p)."""
    x = _f(np.asarray(v).astype(F32))
    p = _f(p)
    a = np.where(x == 0, F32(0), x * _log(p))
    b = np.where((F32(1) - x) == 0, F32(0), (F32(1) - x) * _log1p(-p))
    return (a + b).astype(F32)


def bernoulli_logpdf(v, logits):
    """tfd.Bernoulli(logits=l)._log_prob: -softplus(-l)*x - softplus(l)*(1-x)."""
    x = _f(np.asarray(v).astype(F32))
    l = _f(logits)
    return (-_softplus(-l) * x - _softplus(l) * (F32(1) - x)).astype(F32)


def log_softmax(logits):
    l = _f(logits)
    m = np.max(l, axis=-1, keepdims=True)
    sh = (l - m).astype(F32)
    # sequential left-to-right float32 sum, the order the device loop uses
    e = _exp(sh)
    s = np.zeros(e.shape[:-1], dtype=F32)
    for j in range(e.shape[-1]):
        s = (s + e[..., j]).astype(F32)
    return (sh - _log(s)[..., None]).astype(F32)


def categorical_logpdf(v, logits):
    """tfd.Categorical(logits)._log_prob: log_softmax(logits)[k]."""
    ls = log_softmax(logits)
    k = np.asarray(v).astype(np.int64)
    if ls.ndim == 1:
        return ls[k].astype(F32)
    k = np.broadcast_to(k, ls.shape[:-1])
    return np.take_along_axis(ls, k[..., None], axis=-1)[..., 0].astype(F32)


def exponential_logpdf(v, rate):
    """tfd.Exponential._log_prob: log(rate) - rate*x (x >= 0)."""
    v, rate = _f(v), _f(rate)
    lp = (_log(rate) - rate * v).astype(F32)
    return np.where(v < 0, F32(-np.inf), lp).astype(F32)


def mv_normal_diag_logpdf(v, loc, scale_diag):
    """tfd.MultivariateNormalDiag: sum of independent Normal log_probs
    (left-to-right over the event axis in float32)."""
    lp = normal_logpdf(v, loc, scale_diag)
    out = np.zeros(lp.shape[:-1], dtype=F32)
    for j in range(lp.shape[-1]):
        out = (out + lp[..., j]).astype(F32)
    return out


def _cholesky_f32(cov):
    """Cholesky factor in float32 with the device's operation order (csrc/gjb_dist.cuh MvNormal::cholesky)."""
    cov = _f(cov)
    d = cov.shape[0]
    L = np.zeros((d, d), dtype=F32)
    logdet = F32(0.0)
    for i in range(d):
        for j in range(i + 1):
            s = cov[i, j]
            for m in range(j):
                s = F32(s - F32(L[i, m] * L[j, m]))
            if i == j:
                L[i, i] = np.sqrt(s).astype(F32)
                logdet = F32(logdet + _log(L[i, i]))
            else:
                L[i, j] = F32(s / L[j, j])
    return L, logdet


def mv_normal_logpdf(v, loc, cov):
    """tfd.MultivariateNormalFullCovariance._log_prob via the Cholesky factor:
    -0.5 |L^-1 (x - loc)|^2 - sum log L_kk - d/2 log 2pi (forward substitution, float32)."""
    L, logdet = _cholesky_f32(cov)
    d = L.shape[0]
    v = _f(v)
    if v.ndim == 1:
        v = v[None]
    loc = np.broadcast_to(_f(loc), v.shape)
    z = np.zeros(v.shape, dtype=F32)
    q = np.zeros(v.shape[0], dtype=F32)
    for i in range(d):
        s = (v[:, i] - loc[:, i]).astype(F32)
        for m in range(i):
            s = (s - L[i, m] * z[:, m]).astype(F32)
        z[:, i] = (s / L[i, i]).astype(F32)
        q = (q + z[:, i] * z[:, i]).astype(F32)
    return (F32(-0.5) * q - logdet - F32(d) * _HALF_LOG_2PI).astype(F32)


def mv_normal_sample(words, idx, site, loc, cov):
    """loc + L eps with eps_k from chunk k // 4 of the particle's own stream (like mv_normal_diag)."""
    L, _ = _cholesky_f32(cov)
    d = L.shape[0]
    eps = rng.normal_vec(words, idx, site, d)
    loc = np.broadcast_to(_f(loc), eps.shape)
    out = np.empty_like(eps)
    for i in range(d):
        s = loc[:, i].astype(F32)
        for m in range(i + 1):
            s = (s + L[i, m] * eps[:, m]).astype(F32)
        out[:, i] = s
    return out


def gamma_logpdf(v, concentration, rate):
    """tfd.Gamma._log_prob: xlogy(a-1, x) - rate*x - (lgamma(a) - a*log(rate))."""
    v, a, b = _f(v), _f(concentration), _f(rate)
    t = np.where((a - F32(1)) == 0, F32(0), (a - F32(1)) * _log(v))
    return (t - b * v - (_lgamma(a) - a * _log(b))).astype(F32)


def beta_logpdf(v, a, b):
    """tfd.Beta._log_prob: xlogy(a-1,x) + xlog1py(b-1,-x) - lbeta(a,b)."""
    v, a, b = _f(v), _f(a), _f(b)
    t1 = np.where((a - F32(1)) == 0, F32(0), (a - F32(1)) * _log(v))
    t2 = np.where((b - F32(1)) == 0, F32(0), (b - F32(1)) * _log1p(-v))
    lbeta = (_lgamma(a) + _lgamma(b) - _lgamma(a + b)).astype(F32)
    return (t1 + t2 - lbeta).astype(F32)


def half_normal_logpdf(v, scale):
    """tfd.HalfNormal._log_prob: 0.5 log(2/pi) - log s - 0.5 (x/s)^2, x >= 0."""
    v, s = _f(v), _f(scale)
    z = (v / s).astype(F32)
    lp = (F32(0.5 * math.log(2.0 / math.pi)) - _log(s) - F32(0.5) * z * z).astype(F32)
    return np.where(v < 0, F32(-np.inf), lp).astype(F32)


def cauchy_logpdf(v, loc, scale):
    """tfd.Cauchy._log_prob: -log1p(z^2) - (log pi + log s), z = (x - loc) / s
    (tensorflow_probability/__init__.py:110)."""
    v, m, s = _f(v), _f(loc), _f(scale)
    z = ((v - m) / s).astype(F32)
    return (-_log1p((z * z).astype(F32)) - (F32(math.log(math.pi)) + _log(s))).astype(F32)


def half_cauchy_logpdf(v, loc, scale):
    """tfd.HalfCauchy._log_prob: log(2/pi) - log s - log1p(z^2) for x >= loc
    (tensorflow_probability/__init__.py:179)."""
    v, m, s = _f(v), _f(loc), _f(scale)
    z = ((v - m) / s).astype(F32)
    lp = (F32(math.log(2.0 / math.pi)) - _log(s) - _log1p((z * z).astype(F32))).astype(F32)
    return np.where(v < m, F32(-np.inf), lp).astype(F32)


def laplace_logpdf(v, loc, scale):
    """tfd.Laplace._log_prob: -|z| - log 2 - log s (tensorflow_probability/__init__.py:214)."""
    v, m, s = _f(v), _f(loc), _f(scale)
    return (-np.abs(((v - m) / s).astype(F32)) - F32(math.log(2.0)) - _log(s)).astype(F32)


def log_normal_logpdf(v, loc, scale):
    """tfd.LogNormal = Exp(Normal): Normal log-density of log x minus log x, x > 0
    (tensorflow_probability/__init__.py:219)."""
    v = _f(v)
    with np.errstate(divide="ignore", invalid="ignore"):
        lv = _log(v)
        lp = (normal_logpdf(lv, loc, scale) - lv).astype(F32)
    return np.where(v > 0, lp, F32(-np.inf)).astype(F32)


def gumbel_logpdf(v, loc, scale):
    """tfd.Gumbel._log_prob: -(z + exp(-z)) - log s (tensorflow_probability/__init__.py:174)."""
    v, m, s = _f(v), _f(loc), _f(scale)
    z = ((v - m) / s).astype(F32)
    return (-(z + _exp(-z)) - _log(s)).astype(F32)


def weibull_logpdf(v, concentration, scale):
    """tfd.Weibull (inverse WeibullCDF bijector on Uniform): log k - log s + (k-1) t - exp(k t), t = log x - log s,
    x >= 0 (tensorflow_probability/__init__.py:309)."""
    v, k, s = _f(v), _f(concentration), _f(scale)
    with np.errstate(divide="ignore", invalid="ignore"):
        t = (_log(v) - _log(s)).astype(F32)
        lp = (_log(k) - _log(s) + (k - F32(1)) * t - _exp((k * t).astype(F32))).astype(F32)
    return np.where(v < 0, F32(-np.inf), lp).astype(F32)


def kumaraswamy_logpdf(v, concentration1, concentration0):
    """tfd.Kumaraswamy (inverse KumaraswamyCDF bijector on Uniform): log a + log b + xlogy(a-1, x) + xlog1py(b-1, -x^a),
    0 <= x <= 1 (tensorflow_probability/__init__.py:204)."""
    v, a, b = _f(v), _f(concentration1), _f(concentration0)
    with np.errstate(divide="ignore", invalid="ignore"):
        lv = _log(v)
        t1 = np.where((a - F32(1)) == 0, F32(0), (a - F32(1)) * lv).astype(F32)
        t2 = np.where((b - F32(1)) == 0, F32(0), (b - F32(1)) * _log1p(-_exp((a * lv).astype(F32)))).astype(F32)
        lp = (_log(a) + _log(b) + t1 + t2).astype(F32)
    return np.where((v < 0) | (v > 1), F32(-np.inf), lp).astype(F32)


def logit_normal_logpdf(v, loc, scale):
    """tfd.LogitNormal = Sigmoid(Normal): Normal log-density of logit x minus log x + log(1 - x), 0 < x < 1
    (tensorflow_probability/__init__.py:224)."""
    v = _f(v)
    with np.errstate(divide="ignore", invalid="ignore"):
        lv, l1 = _log(v), _log1p(-v)
        lp = (normal_logpdf((lv - l1).astype(F32), loc, scale) - lv - l1).astype(F32)
    return np.where((v > 0) & (v < 1), lp, F32(-np.inf)).astype(F32)


def geometric_logpdf(v, probs):
    """tfd.Geometric._log_prob: xlog1py(x, -p) + log p, x >= 0 (tensorflow_probability/__init__.py:169)."""
    v, p = _f(v), _f(probs)
    t = np.where(v == 0, F32(0), v * _log1p(-p)).astype(F32)
    return np.where(v < 0, F32(-np.inf), (t + _log(p)).astype(F32)).astype(F32)


def inverse_gamma_logpdf(v, concentration, scale):
    """tfd.InverseGamma._log_prob: a log b - lgamma(a) - (a + 1) log x - b / x, x > 0
    (tensorflow_probability/__init__.py:194)."""
    v, a, b = _f(v), _f(concentration), _f(scale)
    with np.errstate(divide="ignore", invalid="ignore"):
        lp = (a * _log(b) - _lgamma(a) - (a + F32(1)) * _log(v) - b / v).astype(F32)
    return np.where(v > 0, lp, F32(-np.inf)).astype(F32)


def chi2_logpdf(v, df):
    """tfd.Chi2(df) = Gamma(df / 2, rate 1/2) (tensorflow_probability/__init__.py:120)."""
    return gamma_logpdf(v, (F32(0.5) * _f(df)).astype(F32), F32(0.5))


def student_t_logpdf(v, df, loc, scale):
    """tfd.StudentT._log_prob: -0.5 (df + 1) log1p(y^2 / df) - (log|s| + 0.5 log df + 0.5 log pi + lgamma(df / 2)
    - lgamma((df + 1) / 2)), y = (x - loc) / s (tensorflow_probability/__init__.py:279)."""
    v, df, m, s = _f(v), _f(df), _f(loc), _f(scale)
    y = ((v - m) / s).astype(F32)
    norm = (_log(np.abs(s)) + F32(0.5) * _log(df) + F32(0.5) * F32(math.log(math.pi)) + _lgamma((F32(0.5) * df).astype(F32))
            - _lgamma((F32(0.5) * (df + F32(1))).astype(F32))).astype(F32)
    return (F32(-0.5) * (df + F32(1)) * _log1p((y * y / df).astype(F32)) - norm).astype(F32)


def poisson_logpdf(v, rate):
    """tfd.Poisson._log_prob: xlogy(x, rate) - lgamma(x + 1) - rate, x >= 0 (tensorflow_probability/__init__.py:264)."""
    v, r = _f(v), _f(rate)
    t = np.where(v == 0, F32(0), v * _log(r)).astype(F32)
    with np.errstate(invalid="ignore"):
        lp = (t - _lgamma((v + F32(1)).astype(F32)) - r).astype(F32)
    return np.where(v < 0, F32(-np.inf), lp).astype(F32)


# ---------------------------------------------------------------- samplers
# sampler(words, idx, site, *args) -> values for the lanes in idx


def _bc(a, n):
    a = _f(a)
    return np.broadcast_to(a, (n,) + a.shape[1:]) if a.ndim >= 1 and a.shape[0] == n else np.broadcast_to(a, (n,) + a.shape)


def normal_sample(words, idx, site, loc, scale):
    z = rng.quad_normal(words, idx, site)
    return rng.fma32(_f(scale), z, _f(loc))  # one fused multiply-add, as Normal::sample on the device


def uniform_sample(words, idx, site, low, high):
    u = rng.quad_u01(words, idx, site)
    return (_f(low) + (_f(high) - _f(low)) * u).astype(F32)


def flip_sample(words, idx, site, p):
    return rng.quad_u01(words, idx, site) < _f(p)


def bernoulli_sample(words, idx, site, logits):
    l = _f(logits).astype(np.float64)
    p = (1.0 / (1.0 + np.exp(-l))).astype(F32)
    return rng.quad_u01(words, idx, site) < p


def categorical_sample(words, idx, site, logits):
    """Inverse-CDF over exp(l - max l): first k with cumsum_k > u * total.
    (TFP draws argmax(logits + Gumbel); same distribution, different stream.)"""
    u = rng.quad_u01(words, idx, site)
    l = _f(logits)
    if l.ndim == 1:
        l = np.broadcast_to(l, (u.shape[0], l.shape[0]))
    m = np.max(l, axis=-1, keepdims=True)
    e = _exp((l - m).astype(F32))
    K = e.shape[-1]
    tot = np.zeros(u.shape, dtype=F32)
    for j in range(K):
        tot = (tot + e[:, j]).astype(F32)
    t = (u * tot).astype(F32)
    acc = np.zeros(u.shape, dtype=F32)
    k = np.full(u.shape, K - 1, dtype=np.int32)
    done = np.zeros(u.shape, dtype=bool)
    for j in range(K):
        acc = (acc + e[:, j]).astype(F32)
        hit = (~done) & (acc > t)
        k[hit] = j
        done |= hit
    return k


def exponential_sample(words, idx, site, rate):
    return (-_log(rng.quad_u01(words, idx, site)) / _f(rate)).astype(F32)


def _tan_centered(words, idx, site):
    u = rng.quad_u01(words, idx, site)
    return np.tan((F32(math.pi) * (u - F32(0.5)).astype(F32)).astype(F32).astype(np.float64)).astype(F32)


def cauchy_sample(words, idx, site, loc, scale):
    return (_f(loc) + _f(scale) * _tan_centered(words, idx, site)).astype(F32)


def half_cauchy_sample(words, idx, site, loc, scale):
    return (_f(loc) + _f(scale) * np.abs(_tan_centered(words, idx, site))).astype(F32)


def laplace_sample(words, idx, site, loc, scale):
    w = (F32(2) * rng.quad_u01(words, idx, site) - F32(1)).astype(F32)
    return (_f(loc) - _f(scale) * np.copysign(_log1p(-np.abs(w)), w).astype(F32)).astype(F32)


def log_normal_sample(words, idx, site, loc, scale):
    return _exp((_f(loc) + _f(scale) * rng.quad_normal(words, idx, site)).astype(F32))


def gumbel_sample(words, idx, site, loc, scale):
    return (_f(loc) - _f(scale) * _log(-_log(rng.quad_u01(words, idx, site)))).astype(F32)


def weibull_sample(words, idx, site, concentration, scale):
    e = (-_log1p(-rng.quad_u01(words, idx, site))).astype(F32)
    return (_f(scale) * _exp((_log(e) / _f(concentration)).astype(F32))).astype(F32)


def kumaraswamy_sample(words, idx, site, concentration1, concentration0):
    u = rng.quad_u01(words, idx, site)
    t = (-np.expm1((_log1p(-u) / _f(concentration0)).astype(F32).astype(np.float64))).astype(F32)
    return _exp((_log(t) / _f(concentration1)).astype(F32))


def logit_normal_sample(words, idx, site, loc, scale):
    x = (_f(loc) + _f(scale) * rng.quad_normal(words, idx, site)).astype(F32)
    return (F32(1) / (F32(1) + _exp(-x))).astype(F32)


def geometric_sample(words, idx, site, probs):
    return np.floor((_log(rng.quad_u01(words, idx, site)) / _log1p(-_f(probs))).astype(F32)).astype(F32)


def mv_normal_diag_sample(words, idx, site, loc, scale_diag):
    loc = _f(loc)
    scale_diag = _f(scale_diag)
    d = max(loc.shape[-1] if loc.ndim else 1, scale_diag.shape[-1] if scale_diag.ndim else 1)
    z = rng.normal_vec(words, idx, site, d)
    return rng.fma32(np.broadcast_to(scale_diag, z.shape), z, np.broadcast_to(loc, z.shape))  # one FMA per element (mvn_diag_sample)


def half_normal_sample(words, idx, site, scale):
    z = rng.quad_normal(words, idx, site)
    return (np.abs(z) * _f(scale)).astype(F32)


def _gamma_mt(words, idx, site, a, chunk0=0, max_iter=64):
    """Marsaglia-Tsang (2000) Gamma(a,1) for a >= 1; a < 1 boosted by u^(1/a).

    Attempt t uses Philox chunk ``chunk0 + t``: words (0,1) -> normal x,
    word 2 -> acceptance uniform, word 3 (attempt 0 only) -> boost uniform.
    """
    idx = np.asarray(idx, dtype=np.uint64)
    n = idx.shape[0]
    a = np.broadcast_to(_f(a), (n,)).astype(F32)
    boost = a < F32(1)
    a_eff = np.where(boost, a + F32(1), a).astype(F32)
    d = (a_eff - F32(1.0 / 3.0)).astype(F32)
    c = (F32(1) / np.sqrt(F32(9) * d)).astype(F32)
    out = np.zeros(n, dtype=F32)
    done = np.zeros(n, dtype=bool)
    ub = None
    for t in range(max_iter):
        w0, w1, w2, w3 = rng.site_words(words, idx, site, chunk0 + t)
        if t == 0:
            ub = rng.u01(w3)
        x, _ = rng.box_muller(w0, w1)
        u = rng.u01(w2)
        v = (F32(1) + c * x).astype(F32)
        v3 = (v * v * v).astype(F32)
        ok = v > 0
        with np.errstate(invalid="ignore", divide="ignore"):
            lhs = _log(u)
            rhs = (F32(0.5) * x * x + d - d * v3 + d * _log(v3)).astype(F32)
        acc = ok & (lhs < rhs) & (~done)
        out[acc] = (d * v3)[acc]
        done |= acc
        if done.all():
            break
    g = out
    with np.errstate(divide="ignore"):
        bf = _exp(_log(ub) / a)
    return np.where(boost, g * bf, g).astype(F32)


def gamma_sample(words, idx, site, concentration, rate):
    g = _gamma_mt(words, idx, site, concentration, 0)
    return (g / _f(rate)).astype(F32)


def inverse_gamma_sample(words, idx, site, concentration, scale):
    return (_f(scale) / _gamma_mt(words, idx, site, concentration, 0)).astype(F32)


def chi2_sample(words, idx, site, df):
    return (_gamma_mt(words, idx, site, (F32(0.5) * _f(df)).astype(F32), 0) / F32(0.5)).astype(F32)


def student_t_sample(words, idx, site, df, loc, scale):
    """loc + scale * z / sqrt(g / df): g ~ Gamma(df / 2, rate 1/2) on chunks [0, 64), z from chunk 128 of the lane."""
    df = _f(df)
    g = (_gamma_mt(words, idx, site, (F32(0.5) * df).astype(F32), 0) / F32(0.5)).astype(F32)
    w = rng.site_words(words, idx, site, 128)
    z = rng.box_muller(w[0], w[1])[0]
    return (_f(loc) + _f(scale) * (z / np.sqrt((g / df).astype(F32)).astype(F32)).astype(F32)).astype(F32)


def poisson_sample(words, idx, site, rate):
    """rate < 10: inversion by sequential search on one uniform (chunk 0, word 0).  Otherwise PTRS (Hormann 1993):
    attempt t draws (U, V) from words (0, 1) of chunk t; float32 operation order of gjb_dist.cuh: Poisson."""
    idx = np.asarray(idx, dtype=np.uint64)
    n = idx.shape[0]
    r = np.broadcast_to(_f(rate), (n,)).astype(F32)
    out = np.zeros(n, dtype=F32)
    small = r < F32(10)
    if small.any():
        rs = r[small]
        u = rng.u01(rng.site_words(words, idx[small], site, 0)[0])
        p = _exp(-rs)
        s = p.copy()
        k = np.zeros(rs.shape, dtype=F32)
        for _ in range(128):
            go = (u > s) & (k < F32(128))
            if not go.any():
                break
            k = np.where(go, k + F32(1), k).astype(F32)
            with np.errstate(divide="ignore", invalid="ignore"):
                p = np.where(go, (p * (rs / k).astype(F32)).astype(F32), p)
            s = np.where(go, (s + p).astype(F32), s)
        out[small] = k
    big = ~small
    if big.any():
        rb, ib = r[big], idx[big]
        b = (F32(0.931) + F32(2.53) * np.sqrt(rb)).astype(F32)
        a = (F32(-0.059) + F32(0.02483) * b).astype(F32)
        lia = _log((F32(1.1239) + F32(1.1328) / (b - F32(3.4))).astype(F32))
        vr = (F32(0.9277) - F32(3.6224) / (b - F32(2))).astype(F32)
        llam = _log(rb)
        k = np.zeros(rb.shape, dtype=F32)
        done = np.zeros(rb.shape, dtype=bool)
        for t in range(64):
            if done.all():
                break
            w = rng.site_words(words, ib, site, t)
            U = (rng.u01(w[0]) - F32(0.5)).astype(F32)
            V = rng.u01(w[1])
            us = (F32(0.5) - np.abs(U)).astype(F32)
            kt = np.floor(((F32(2) * a / us + b) * U + rb + F32(0.43)).astype(F32)).astype(F32)
            k = np.where(done, k, kt)
            acc1 = (us >= F32(0.07)) & (V <= vr)
            rej = (kt < 0) | ((us < F32(0.013)) & (V > us))
            with np.errstate(invalid="ignore"):
                lhs = (_log(V) + lia - _log((a / (us * us).astype(F32) + b).astype(F32))).astype(F32)
                rhs = (-rb + kt * llam - _lgamma((kt + F32(1)).astype(F32))).astype(F32)
            acc2 = ~rej & (lhs <= rhs)
            done |= ~done & (acc1 | acc2)
        out[big] = np.maximum(k, F32(0))
    return out


def beta_sample(words, idx, site, a, b):
    """X = Ga / (Ga + Gb) with Ga ~ Gamma(a,1) on chunks [0,64), Gb on [64,128)."""
    ga = _gamma_mt(words, idx, site, a, 0)
    gb = _gamma_mt(words, idx, site, b, 64)
    return (ga / (ga + gb)).astype(F32)


# registry: name -> (sampler, logpdf, n_args, value kind)
DISTS = {
    "normal": (normal_sample, normal_logpdf),
    "uniform": (uniform_sample, uniform_logpdf),
    "flip": (flip_sample, flip_logpdf),
    "bernoulli": (bernoulli_sample, bernoulli_logpdf),
    "categorical": (categorical_sample, categorical_logpdf),
    "exponential": (exponential_sample, exponential_logpdf),
    "mv_normal_diag": (mv_normal_diag_sample, mv_normal_diag_logpdf),
    "half_normal": (half_normal_sample, half_normal_logpdf),
    "gamma": (gamma_sample, gamma_logpdf),
    "mv_normal": (mv_normal_sample, mv_normal_logpdf),
    "beta": (beta_sample, beta_logpdf),
    "cauchy": (cauchy_sample, cauchy_logpdf),
    "half_cauchy": (half_cauchy_sample, half_cauchy_logpdf),
    "laplace": (laplace_sample, laplace_logpdf),
    "log_normal": (log_normal_sample, log_normal_logpdf),
    "gumbel": (gumbel_sample, gumbel_logpdf),
    "weibull": (weibull_sample, weibull_logpdf),
    "kumaraswamy": (kumaraswamy_sample, kumaraswamy_logpdf),
    "logit_normal": (logit_normal_sample, logit_normal_logpdf),
    "geometric": (geometric_sample, geometric_logpdf),
    "inverse_gamma": (inverse_gamma_sample, inverse_gamma_logpdf),
    "chi2": (chi2_sample, chi2_logpdf),
    "student_t": (student_t_sample, student_t_logpdf),
    "poisson": (poisson_sample, poisson_logpdf),
}


# ---------------------------------------------------------------- dist.repeat / dist.vmap (combinators/repeat.py, vmap.py)
# N independent draws of a scalar primitive as ONE vector site: element k takes word k % 4 of chunk k // 4 of the lane's own
# stream (the uniform, or the Box-Muller normal of that slot) and goes through the scalar sampler; the log-density is the
# sum of the elements' (vmap.py:180-218: score = sum of the inner scores), accumulated in element order.


class _FixedDraws:
    """Make the scalar samplers above read given draws instead of the quad stream."""

    def __init__(self, u, z):
        self.u, self.z = u, z

    def __enter__(self):
        self.saved = (rng.quad_u01, rng.quad_normal)
        rng.quad_u01 = lambda *a: self.u
        rng.quad_normal = lambda *a: self.z

    def __exit__(self, *exc):
        rng.quad_u01, rng.quad_normal = self.saved
        return False


def _elem(a, k, n):
    a = _f(a)
    if a.ndim == 2 or (a.ndim == 1 and a.shape[0] != n):
        return a[..., k]
    return a


def repeated(base: str, width: int):
    """(sampler, logpdf) of ``base.repeat(n=width)``."""
    b_sample, b_logpdf = DISTS[base]

    def sample(words, idx, site, *args):
        idx_ = np.asarray(idx, dtype=np.uint64)
        n = idx_.shape[0]
        out = np.empty((n, width), dtype=F32)
        for c in range((width + 3) // 4):
            w = rng.site_words(words, idx_, site, c)
            z = rng.normal4(words, idx_, site, c)
            for t in range(4):
                k = 4 * c + t
                if k < width:
                    with _FixedDraws(rng.u01(w[t]), z[t]):
                        out[:, k] = b_sample(words, idx_, site, *[_elem(a, k, n) for a in args])
        return out

    def logpdf(v, *args):
        v = _f(v)
        n = v.shape[0] if v.ndim == 2 else 1
        tot = None
        for k in range(width):
            lp = _f(b_logpdf(v[..., k], *[_elem(a, k, n) for a in args]))
            tot = lp if tot is None else (tot + lp).astype(F32)
        return tot

    return sample, logpdf


class _Dists(dict):
    """``DISTS["repeat_<base>"]`` resolves lazily; the width comes from the arguments at call time."""

    def __contains__(self, name):
        return dict.__contains__(self, name) or (isinstance(name, str) and name.startswith("repeat_") and dict.__contains__(self, name[7:]))

    def __missing__(self, name):
        if isinstance(name, str) and name.startswith("repeat_") and dict.__contains__(self, name[7:]):
            base = name[7:]

            def sample(words, idx, site, *args, width=None):
                w = width or max(_f(a).shape[-1] for a in args if _f(a).ndim >= 1)
                return repeated(base, w)[0](words, idx, site, *args)

            def logpdf(v, *args):
                return repeated(base, _f(v).shape[-1])[1](v, *args)

            return sample, logpdf
        raise KeyError(name)


DISTS = _Dists(DISTS)

"""Primitive distributions: host objects with a device-side (sampler, logpdf)
pair in csrc/gjb_dist.cuh.

API mirror of ``Distribution`` / ``ExactDensity`` / ``exact_density``
(generative_functions/distributions/distribution.py:90-106, 359-396, 436-476)
and of the TFP wrappers the hot path uses
(distributions/tensorflow_probability/__init__.py:72-294).  In the reference a
distribution's ``sample`` / ``logpdf`` are arbitrary JAX-traceable Python; here
a primitive can be fused only if it has device code registered under its name
-- anything else raises ``NotFusable`` instead of falling back to the CPU.
"""

from __future__ import annotations

import warnings

from . import expr as E
from .expr import Expr, F32, I32
from .gfi import GenerativeFunction, GenerativeFunctionClosure


class NotFusable(NotImplementedError):
    pass


def implicit_logit_warning(name: str = "distribution"):
    """distribution.py:479-500: a bare positional argument to categorical / bernoulli means logits, and warns."""
    warnings.warn(
        f"The use of a bare argument to genjax.{name} is deprecated. Please specify `logits=` or `probs=` for the "
        "parameters. The default, which will be used in this case, is logits.",
        DeprecationWarning,
        stacklevel=3,
    )


class Distribution(GenerativeFunction):
    """A single random choice.  Subclasses define the canonical argument list
    and the CUDA snippets the code generator splices into the fused kernel."""

    name: str = "?"
    cuda: str = "?"  # struct in gjb_dist.cuh
    n_args: int = 0
    value_dtype: str = F32
    vector: bool = False  # value has an event axis

    # -- closure syntax: dist(args) @ "addr"
    def __call__(self, *args, **kwargs) -> GenerativeFunctionClosure:
        return GenerativeFunctionClosure(self, args, kwargs)

    def canonical_args(self, args) -> list[Expr]:
        args, kwargs = _split_kwargs(args)
        return [E.lift(a) for a in self._canonical(args, kwargs)]

    def _canonical(self, args, kwargs) -> list:
        if kwargs:
            raise TypeError(f"{self.name} takes positional arguments only")
        if len(args) != self.n_args:
            raise TypeError(f"{self.name} expects {self.n_args} arguments, got {len(args)}")
        return list(args)

    def value_type(self, cargs: list[Expr]) -> tuple[str, tuple]:
        return self.value_dtype, ()

    # -- codegen hooks (scalar distributions) ---------------------------
    # rng_kind: what the sampler consumes from the site's Philox block
    #   "uniform": one u01 word of the quad block W{site} (slot `sub`)
    #   "normal" : one component of normal4_of(W{site}) = Z{site}
    #   "lane"   : its own per-particle stream (rejection samplers)
    rng_kind: str = "uniform"

    def draw_expr(self, site: int) -> str:
        if self.rng_kind == "normal":
            return f"gjb::pick(Z{site}, sub)"
        if self.rng_kind == "uniform":
            return f"gjb::u01(gjb::pick(W{site}, sub))"
        return f"rng, {site + 1}u"

    def emit_sample(self, site: int, a: list[str], cg) -> str:
        return f"gjb::{self.cuda}::sample({self.draw_expr(site)}, {', '.join(a)})"

    def emit_logpdf(self, v: str, a: list[str], cg) -> str:
        return f"gjb::{self.cuda}::logpdf({v}, {', '.join(a)})"

    def __repr__(self):
        return f"genjax.{self.name}()"  # what the reference prints (test_distributions.py:476-489)


def _check_kwargs(name: str, kwargs: dict, allowed: tuple) -> dict:
    """Reject what this build cannot honour instead of dropping it: unknown keywords, and ``sample_shape=n``
    (tfp draws n iid values into one choice, tensorflow_probability/__init__.py:53-55; here one choice is one draw)."""
    shape = kwargs.get("sample_shape", ())
    shape = getattr(shape, "value", shape)  # genjax.Const((...)) wraps the static shape
    if shape not in ((), None):
        raise NotImplementedError(f"{name}(..., sample_shape=...) is not supported: one choice holds one draw")
    rest = {k: v for k, v in kwargs.items() if k != "sample_shape"}
    unknown = [k for k in rest if k not in allowed]
    if unknown:
        raise TypeError(f"{name} got unexpected keyword argument(s) {unknown}; accepted: {list(allowed)}")
    return rest


def _split_kwargs(args):
    """Unpack the reference's ``(args, kwargs_dict)`` convention (distribution.py:448-462)."""
    if isinstance(args, tuple) and len(args) == 2 and isinstance(args[1], dict) and isinstance(args[0], tuple):
        return args[0], args[1]
    return tuple(args), {}


class _Normal(Distribution):
    name, cuda, n_args = "normal", "Normal", 2
    rng_kind = "normal"

    def _canonical(self, args, kwargs):
        kwargs = _check_kwargs("normal", kwargs, ("loc", "scale"))
        if kwargs:
            args = tuple(args) + tuple(kwargs[k] for k in ("loc", "scale") if k in kwargs)
        return super()._canonical(args, {})


class _Uniform(Distribution):
    name, cuda, n_args = "uniform", "Uniform", 2

    def _canonical(self, args, kwargs):
        kwargs = _check_kwargs("uniform", kwargs, ("low", "high"))
        if kwargs:
            args = tuple(args) + tuple(kwargs[k] for k in ("low", "high") if k in kwargs)
        if len(args) == 0:
            args = (0.0, 1.0)
        return super()._canonical(args, {})


class _Exponential(Distribution):
    name, cuda, n_args = "exponential", "Exponential", 1


class _HalfNormal(Distribution):
    name, cuda, n_args = "half_normal", "HalfNormal", 1
    rng_kind = "normal"


class _LocScale(Distribution):
    """Long-tail scalar wrappers with ``(loc, scale)`` parameters (tensorflow_probability/__init__.py:110, 174, 179,
    214, 219): one inverse-CDF draw per site, keyword spelling as TFP's."""

    n_args = 2
    kw_names = ("loc", "scale")

    def _canonical(self, args, kwargs):
        kwargs = _check_kwargs(self.name, kwargs, self.kw_names)
        if kwargs:
            args = tuple(args) + tuple(kwargs[k] for k in self.kw_names if k in kwargs)
        return super()._canonical(args, {})


class _Cauchy(_LocScale):
    name, cuda = "cauchy", "Cauchy"


class _HalfCauchy(_LocScale):
    name, cuda = "half_cauchy", "HalfCauchy"


class _Laplace(_LocScale):
    name, cuda = "laplace", "Laplace"


class _LogNormal(_LocScale):
    name, cuda = "log_normal", "LogNormal"
    rng_kind = "normal"


class _Gumbel(_LocScale):
    name, cuda = "gumbel", "Gumbel"


class _Weibull(_LocScale):
    """tfd.Weibull(concentration, scale) (tensorflow_probability/__init__.py:309)."""

    name, cuda = "weibull", "Weibull"
    kw_names = ("concentration", "scale")


class _Kumaraswamy(_LocScale):
    """tfd.Kumaraswamy(concentration1, concentration0) (tensorflow_probability/__init__.py:204)."""

    name, cuda = "kumaraswamy", "Kumaraswamy"
    kw_names = ("concentration1", "concentration0")


class _LogitNormal(_LocScale):
    name, cuda = "logit_normal", "LogitNormal"
    rng_kind = "normal"


class _Geometric(_LocScale):
    """tfd.Geometric(probs=p) (tensorflow_probability/__init__.py:169): float-valued count of failures."""

    name, cuda, n_args = "geometric", "Geometric", 1
    kw_names = ("probs",)


class _InverseGamma(_LocScale):
    """tfd.InverseGamma(concentration, scale) (tensorflow_probability/__init__.py:194)."""

    name, cuda = "inverse_gamma", "InverseGamma"
    kw_names = ("concentration", "scale")
    rng_kind = "lane"


class _Chi2(_LocScale):
    """tfd.Chi2(df) (tensorflow_probability/__init__.py:120)."""

    name, cuda, n_args = "chi2", "Chi2", 1
    kw_names = ("df",)
    rng_kind = "lane"


class _StudentT(_LocScale):
    """tfd.StudentT(df, loc, scale) (tensorflow_probability/__init__.py:279)."""

    name, cuda, n_args = "student_t", "StudentT", 3
    kw_names = ("df", "loc", "scale")
    rng_kind = "lane"


class _Poisson(_LocScale):
    """tfd.Poisson(rate) (tensorflow_probability/__init__.py:264): float-valued count."""

    name, cuda, n_args = "poisson", "Poisson", 1
    kw_names = ("rate",)
    rng_kind = "lane"


class _Gamma(Distribution):
    name, cuda, n_args = "gamma", "Gamma", 2
    rng_kind = "lane"


class _Beta(Distribution):
    name, cuda, n_args = "beta", "Beta", 2
    rng_kind = "lane"


class _Flip(Distribution):
    """tfd.Bernoulli(probs=p, dtype=bool) (tensorflow_probability/__init__.py:155)."""

    name, cuda, n_args, value_dtype = "flip", "Flip", 1, I32
    bool_valued = True


class _Bernoulli(Distribution):
    """tfd.Bernoulli(logits=...) or (probs=...) (tensorflow_probability/__init__.py:72)."""

    name, cuda, n_args, value_dtype = "bernoulli", "Bernoulli", 1, I32

    def _canonical(self, args, kwargs):
        kwargs = _check_kwargs("bernoulli", kwargs, ("probs", "logits"))
        if "probs" in kwargs:
            p = E.lift(kwargs["probs"])
            return [E.unary("log", p) - E.unary("log1p", -p)]
        if "logits" in kwargs:
            return [kwargs["logits"]]
        if len(args) == 1:
            implicit_logit_warning("bernoulli")
            return [args[0]]
        raise TypeError("bernoulli expects logits= or probs=")


class _Categorical(Distribution):
    """tfd.Categorical(logits=...) (tensorflow_probability/__init__.py:102)."""

    name, cuda, n_args, value_dtype = "categorical", "Categorical", 1, I32

    def _canonical(self, args, kwargs):
        kwargs = _check_kwargs("categorical", kwargs, ("probs", "logits"))
        if "probs" in kwargs:
            return [E.unary("log", E.lift(kwargs["probs"]))]
        if "logits" in kwargs:
            return [kwargs["logits"]]
        if len(args) == 1:
            implicit_logit_warning("categorical")
            return [args[0]]
        raise TypeError("categorical expects logits= or probs=")

    def value_type(self, cargs):
        if cargs[0].ndim != 1:
            raise TypeError("categorical logits must be a vector")
        return I32, ()

    def emit_sample(self, site, a, cg):
        ptr, k = a[0]
        return f"gjb::Categorical::sample({self.draw_expr(site)}, {ptr}, {k})"

    def emit_logpdf(self, v, a, cg):
        ptr, k = a[0]
        return f"gjb::Categorical::logpdf({v}, {ptr}, {k})"


class _MvNormalDiag(Distribution):
    """tfd.MultivariateNormalDiag(loc, scale_diag) (tensorflow_probability/__init__.py:239)."""

    name, cuda, n_args, vector = "mv_normal_diag", "MvNormalDiag", 2, True
    rng_kind = "lane"

    def value_type(self, cargs):
        d = max((c.shape[0] for c in cargs if c.ndim == 1), default=0)
        if d == 0:
            raise TypeError("mv_normal_diag needs a vector loc or scale_diag")
        return F32, (d,)


class _RepeatedNormal(_MvNormalDiag):
    """``normal.repeat(n=N)`` / ``normal.vmap(...)`` over N iid draws sharing (or
    zipping) loc and scale: one vector site scored like mv_normal_diag
    (combinators/repeat.py:37-41, vmap.py:180-218: score = sum of the inner scores)."""

    def __init__(self, n: int):
        self.n = int(n)

    def _canonical(self, args, kwargs):
        loc, scale = _Normal._canonical(normal, args, kwargs)
        zeros = Expr("constvec", (), F32, (self.n,), tuple(0.0 for _ in range(self.n)))
        out = []
        for a in (loc, scale):
            a = E.lift(a)
            if a.ndim == 0:
                a = a + zeros
            elif a.shape[0] != self.n:
                raise TypeError(f"argument of width {a.shape[0]} under repeat/vmap of n={self.n}")
            out.append(a)
        return out


def _normal_repeat(self, n: int):
    return _RepeatedNormal(n)


def _normal_vmap(self, in_axes=0, axis_size: int | None = None):
    if axis_size is None:
        raise TypeError("normal.vmap needs axis_size=N here (shapes are static at capture time)")
    return _RepeatedNormal(axis_size)


_Normal.repeat = _normal_repeat
_Normal.vmap = _normal_vmap


class _Repeated(Distribution):
    """``dist.repeat(n=N)`` / ``dist.vmap(in_axes=...)`` for a scalar primitive: ONE vector site of N independent draws
    (combinators/repeat.py:37-41; vmap.py:180-218: the score is the sum of the inner scores).  Arguments shared by all
    draws stay scalars, mapped ones are width-N vectors.  Element k draws from word ``k % 4`` of chunk ``k // 4`` of the
    particle's own Philox stream (as mv_normal_diag does), through the base primitive's device sampler and log-density."""

    vector = True
    rng_kind = "lane"

    def __init__(self, base: Distribution, n: int | None, in_axes=0):
        if base.vector or base.value_dtype != F32 or base.rng_kind not in ("uniform", "normal") \
                or type(base).emit_sample is not Distribution.emit_sample:
            raise NotFusable(f"{base.name}.repeat / .vmap: only float-valued scalar primitives drawn from one uniform or "
                             "one normal word can be mapped into a vector site")
        self.base = base
        self.n = None if n is None else int(n)
        self.in_axes = in_axes
        self.name = f"repeat_{base.name}"
        self.cuda = base.cuda
        self.n_args = base.n_args

    def _canonical(self, args, kwargs):
        cargs = [E.lift(a) for a in self.base._canonical(args, kwargs)]
        axes = self.in_axes if isinstance(self.in_axes, (tuple, list)) else (self.in_axes,) * len(cargs)
        if len(axes) != len(cargs):
            raise ValueError("vmap in_axes specification must be a tree prefix of the corresponding value")
        widths = {a.shape[0] for a in cargs if a.ndim == 1} | ({self.n} if self.n is not None else set())
        if len(widths) != 1:
            raise TypeError(f"{self.name}: cannot infer one axis size from the arguments (got {sorted(widths)}); pass "
                            "n= / axis_size=")
        for a, ax in zip(cargs, axes):
            if a.ndim > 1 or (ax is None and a.ndim == 1):
                raise TypeError(f"{self.name}: argument of shape {a.shape} under in_axes={ax}")
        return cargs

    def value_type(self, cargs):
        widths = {a.shape[0] for a in cargs if a.ndim == 1} | ({self.n} if self.n is not None else set())
        return F32, (widths.pop(),)

    def logpdf_expr(self, v, args):
        from . import autodiff as AD

        lp = AD.logpdf_expr(self.base, v, list(args))
        return E.vsum(lp) if lp.ndim == 1 else lp * float(v.shape[0])

    def __repr__(self):
        return f"genjax.{self.base.name}.repeat(n={self.n})"


def _dist_repeat(self, n: int):
    """``dist.repeat(n=N)`` (generative_function.py ``repeat``; combinators/repeat.py:25-79)."""
    return _Repeated(self, n)


def _dist_vmap(self, in_axes=0, axis_size: int | None = None):
    """``dist.vmap(in_axes=...)`` (generative_function.py ``vmap``; combinators/vmap.py:384): the axis size comes from
    the mapped arguments, or from ``axis_size=`` when every argument is shared."""
    return _Repeated(self, axis_size, in_axes)


Distribution.repeat = _dist_repeat
Distribution.vmap = _dist_vmap
_Normal.repeat = _normal_repeat


def _normal_vmap2(self, in_axes=0, axis_size: int | None = None):
    # the lane-group fast path needs the width up front; without it the generic vector site infers it from the arguments
    return _RepeatedNormal(axis_size) if axis_size is not None else _Repeated(self, None, in_axes)


_Normal.vmap = _normal_vmap2


class _GmmDiag(Distribution):
    """Mixture of K diagonal Gaussians with one scale per component:
    ``gmm_diag(logits[K], mu[K, D], sigma[K])`` -- the fusable form of the
    cookbook's custom ``GaussianMixture`` ExactDensity
    (docs/cookbook/inactive/expressivity/custom_distribution.ipynb cell 9):
    sample = ancestral (component, then N(mu_k, sigma_k)), logpdf =
    logsumexp_k(log_softmax(logits)_k + sum_d logN(x_d; mu_kd, sigma_k))."""

    name, cuda, n_args, vector = "gmm_diag", "GmmDiag", 3, True
    rng_kind = "lane"

    def value_type(self, cargs):
        logits, mu, sigma = cargs
        if not (mu.op == "arg" and mu.attr["kind"] == "shared" and mu.ndim == 2):
            raise TypeError("gmm_diag needs mu as a shared [K, D] argument")
        if logits.ndim != 1 or sigma.ndim != 1 or logits.shape[0] != mu.shape[0] or sigma.shape[0] != mu.shape[0]:
            raise TypeError("gmm_diag needs logits [K], mu [K, D], sigma [K]")
        return F32, (mu.shape[1],)

    def logpdf_expr(self, v, args):
        import math

        logits, mu, sigma = args
        K = logits.shape[0]
        m = logits[0]
        for k in range(1, K):
            m = E.binary("max", m, logits[k])
        tot = None
        for k in range(K):
            t = E.unary("exp", logits[k] - m)
            tot = t if tot is None else tot + t
        lse = m + E.unary("log", tot)
        comps = []
        for k in range(K):
            sk = sigma[k]
            z = v / sk - mu[k] / sk
            lpk = E.vsum(E.const(-0.5) * E.unary("square", z) - (E.const(0.5 * math.log(2.0 * math.pi)) + E.unary("log", sk)))
            comps.append((logits[k] - lse) + lpk)
        M = comps[0]
        for c in comps[1:]:
            M = E.binary("max", M, c)
        tot = None
        for c in comps:
            t = E.unary("exp", c - M)
            tot = t if tot is None else tot + t
        return M + E.unary("log", tot)


class _MvNormal(Distribution):
    """tfd.MultivariateNormalFullCovariance(loc, covariance_matrix) (tensorflow_probability/__init__.py:244).
    ``covariance_matrix`` must be a shared [D, D] argument (D <= 16); its Cholesky factor is computed once per
    thread per launch, each particle pays one forward substitution."""

    name, cuda, n_args, vector = "mv_normal", "MvNormal", 2, True
    rng_kind = "lane"

    def value_type(self, cargs):
        loc, cov = cargs
        if not (cov.op == "arg" and cov.attr["kind"] == "shared" and cov.ndim == 2 and cov.shape[0] == cov.shape[1]):
            raise TypeError("mv_normal needs the covariance as a shared [D, D] argument")
        d = cov.shape[0]
        if d > 16:
            raise NotImplementedError("mv_normal on FMA handles D <= 16")
        if loc.ndim == 1 and loc.shape[0] != d:
            raise TypeError("mv_normal: loc and covariance disagree on D")
        return F32, (d,)

    def _canonical(self, args, kwargs):
        if kwargs:
            args = tuple(args) + tuple(kwargs[k] for k in ("loc", "covariance_matrix") if k in kwargs)
        loc, cov = super()._canonical(args, {})
        loc, cov = E.lift(loc), E.lift(cov)
        if loc.ndim == 0 and cov.ndim == 2:
            loc = loc + Expr("constvec", (), F32, (cov.shape[0],), tuple(0.0 for _ in range(cov.shape[0])))
        return [loc, cov]


normal = _Normal()
uniform = _Uniform()
exponential = _Exponential()
half_normal = _HalfNormal()
cauchy = _Cauchy()
half_cauchy = _HalfCauchy()
laplace = _Laplace()
log_normal = _LogNormal()
gumbel = _Gumbel()
weibull = _Weibull()
kumaraswamy = _Kumaraswamy()
logit_normal = _LogitNormal()
geometric = _Geometric()
inverse_gamma = _InverseGamma()
chi2 = _Chi2()
student_t = _StudentT()
poisson = _Poisson()
gamma = _Gamma()
beta = _Beta()
flip = _Flip()
bernoulli = _Bernoulli()
categorical = _Categorical()
mv_normal_diag = _MvNormalDiag()
gmm_diag = _GmmDiag()
mv_normal = _MvNormal()

REGISTRY: dict[str, Distribution] = {
    d.name: d
    for d in (normal, uniform, exponential, half_normal, cauchy, half_cauchy, laplace, log_normal, gumbel, weibull, kumaraswamy, logit_normal, geometric,
              inverse_gamma, chi2, student_t, poisson, gamma, beta, flip, bernoulli, categorical, mv_normal_diag, gmm_diag, mv_normal)
}


class ExactDensity(Distribution):
    """Host-only distribution defined by Python ``sample`` / ``logpdf``
    (distribution.py:359-396).  It keeps the reference's extension API but is
    NOT fusable: using it inside a ``@gen`` model on the GPU path raises."""

    def sample(self, key, *args):  # pragma: no cover - interface
        raise NotImplementedError

    def logpdf(self, v, *args):  # pragma: no cover - interface
        raise NotImplementedError

    def canonical_args(self, args):
        raise NotFusable(
            f"distribution {self.name!r} has no device-side (sampler, logpdf) pair; register one with "
            "genjax_b200.register_primitive(...) -- there is no CPU fallback on the fused path"
        )


def exact_density(sample, logpdf, name: str = "custom") -> ExactDensity:
    """``exact_density(sample, logpdf, name)`` (distribution.py:436-476)."""

    class _Custom(ExactDensity):
        pass

    c = _Custom()
    c.name = name
    c.sample = lambda key, *a: sample(key, *a)  # type: ignore[assignment]
    c.logpdf = lambda v, *a: logpdf(v, *a)  # type: ignore[assignment]
    return c


def register_primitive(name: str, cuda_struct: str, n_args: int, value_dtype: str = F32) -> Distribution:
    """Register a primitive whose ``sample(rng, site, args...)`` / ``logpdf(v, args...)``
    live in a CUDA struct visible to the generated kernels (gjb_dist.cuh or a
    user header on the include path) -- the fusable form of ``exact_density``."""

    class _P(Distribution):
        pass

    p = _P()
    p.name, p.cuda, p.n_args, p.value_dtype = name, cuda_struct, n_args, value_dtype
    p.rng_kind = "lane"  # user primitives get the particle's own stream: sample(rng, site, args...)
    REGISTRY[name] = p
    return p


# ---------------------------------------------------------------------------
# Direct GFI use of a distribution (``genjax.normal.sample(key, 0., 1.)``,
# ``normal(0., 1.).simulate(key, ())``, ``categorical.random_weighted(key, logits)``;
# distribution.py:92-161, 360-419): a one-site static model on the same fused path.


class DistributionTrace:
    """distribution.py:60-87: trace of a single random choice."""

    def __init__(self, gen_fn, inner, args):
        self.gen_fn = gen_fn
        self.inner = inner
        self.args = args

    def get_gen_fn(self):
        return self.gen_fn

    def get_args(self):
        return self.args

    def get_retval(self):
        return self.inner.get_retval()

    def get_score(self):
        return self.inner.get_score()

    def get_choices(self):
        from ..core.choice_map import ChoiceMap

        return ChoiceMap.choice(self.inner.get_retval())

    get_sample = get_choices

    def get_value(self):
        return self.inner.get_retval()


def _wrapped_model(dist: Distribution, kw_names: tuple):
    cache = dist.__dict__.setdefault("_wrapped", {})
    m = cache.get(kw_names)
    if m is None:
        from .static import gen

        n_kw = len(kw_names)

        def body(*flat):
            pos = flat[: len(flat) - n_kw] if n_kw else flat
            kw = dict(zip(kw_names, flat[len(flat) - n_kw:])) if n_kw else {}
            return dist(*pos, **kw) @ "_v"

        body.__name__ = f"dist_{dist.name}"
        m = gen(body)
        cache[kw_names] = m
    return m


def _flat_call(dist, args):
    pos, kw = _split_kwargs(args)
    names = tuple(kw.keys())
    return _wrapped_model(dist, names), tuple(pos) + tuple(kw[k] for k in names)


def _d_simulate(self, key, args):
    m, flat = _flat_call(self, args)
    return DistributionTrace(self, m.simulate(key, flat), args)


def _d_generate(self, key, constraint, args):
    from ..core.choice_map import ChoiceMap

    m, flat = _flat_call(self, args)
    if constraint is not None and constraint.has_value():
        chm = ChoiceMap.entry(constraint.get_value(), "_v")
    else:
        chm = ChoiceMap.empty()
    tr, w = m.generate(key, chm, flat)
    return DistributionTrace(self, tr, args), w


def _d_assess(self, sample, args):
    from ..core.choice_map import ChoiceMap

    m, flat = _flat_call(self, args)
    if not sample.has_value():
        raise ValueError("assess on a distribution needs a value choice map")
    return m.assess(ChoiceMap.entry(sample.get_value(), "_v"), flat)


def _d_sample(self, key, *args, **kwargs):
    a = (args, kwargs) if kwargs else args
    return _d_simulate(self, key, a).get_retval()


def _auto_batch(self, v, args):
    """``dist.logpdf(values[n], ...)``: jax broadcasting over a leading axis becomes one batched launch."""
    import torch

    from .static import Batched

    ev = 1 if self.vector else 0
    if not (isinstance(v, torch.Tensor) and v.ndim == ev + 1):
        return v, args
    n = v.shape[0]
    out = []
    for a in args:
        if isinstance(a, torch.Tensor) and a.ndim >= 1 and a.shape[0] == n and (a.ndim == ev + 1 or not self.vector):
            a = Batched(a)
        out.append(a)
    return Batched(v), tuple(out)


def _d_logpdf(self, v, *args, **kwargs):
    from ..core.choice_map import ChoiceMap

    v, args = _auto_batch(self, v, args)
    a = (args, kwargs) if kwargs else args
    score, _ = _d_assess(self, ChoiceMap.choice(v), a)
    return score


def _d_random_weighted(self, key, *args, **kwargs):
    a = (args, kwargs) if kwargs else args
    tr = _d_simulate(self, key, a)
    return tr.get_score(), tr.get_retval()


def _d_estimate_logpdf(self, key, v, *args, **kwargs):
    return _d_logpdf(self, v, *args, **kwargs)


def _d_edit(self, key, trace, request, argdiffs):
    from ..core.choice_map import ChoiceMap
    from .gfi import Diff, EmptyRequest, Regenerate, Update

    new_args = Diff.tree_primal(argdiffs) if argdiffs not in (None, ()) else trace.get_args()
    m, flat = _flat_call(self, new_args)
    inner = trace.inner
    if isinstance(request, Update):
        c = request.constraint
        chm = ChoiceMap.entry(c.get_value(), "_v") if c.has_value() else ChoiceMap.empty()
        tr, w, rd, bwd = m.edit(key, inner, Update(chm), Diff.unknown_change(flat))
        disc = bwd.constraint
        back = ChoiceMap.choice(disc["_v"]) if "_v" in disc else ChoiceMap.empty()
        return DistributionTrace(self, tr, new_args), w, rd, Update(back)
    if isinstance(request, Regenerate):
        from ..core.choice_map import Selection

        sel = Selection.at["_v"] if request.selection.check() else Selection.none()
        tr, w, rd, bwd = m.edit(key, inner, Regenerate(sel), Diff.unknown_change(flat))
        disc = bwd.constraint
        back = ChoiceMap.choice(disc["_v"]) if "_v" in disc else ChoiceMap.empty()
        return DistributionTrace(self, tr, new_args), w, rd, Update(back)
    if isinstance(request, EmptyRequest):
        return _d_edit(self, key, trace, Update(ChoiceMap.empty()), argdiffs)
    from .gfi import NotSupportedEditRequest

    raise NotSupportedEditRequest(request)


Distribution.simulate = _d_simulate
Distribution.generate = _d_generate
Distribution.assess = _d_assess
Distribution.sample = _d_sample
Distribution.logpdf = _d_logpdf
Distribution.random_weighted = _d_random_weighted
Distribution.estimate_logpdf = _d_estimate_logpdf
Distribution.edit = _d_edit

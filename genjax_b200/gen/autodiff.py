"""Symbolic log-densities and reverse-mode differentiation over the captured
model IR -- what the batched MCMC kernels are generated from.

The reference gets ``grad log p`` from ``jax.grad(gen_fn.assess)`` over the
selected float leaves (inference/requests/hmc.py:70-96).  There is no tracing
AD here: every primitive's ``log_prob`` is restated as an expression over the
same ``Expr`` DAG the model body was captured into (TFP operation order, as in
gjb_dist.cuh / oracle/dists.py), the model's total log-density is their sum,
and ``grad`` walks that DAG backwards once, at code-generation time.  The
generated chain kernels (gen/codegen_chain.py) then evaluate value and
gradient in registers.
"""

from __future__ import annotations

import math

from . import expr as E
from .expr import Expr, F32

_HALF_LOG_2PI = 0.5 * math.log(2.0 * math.pi)


class NotDifferentiable(NotImplementedError):
    pass


# ---------------------------------------------------------------- log-densities


def _sum(x: Expr) -> Expr:
    return E.vsum(x) if x.ndim else x


def logpdf_expr(dist, v: Expr, args: list) -> Expr:
    """Scalar ``Expr`` of ``dist.log_prob(v)`` (sum-reduced over an event axis,
    distribution.py:393-394)."""
    name = dist.name
    if name in ("normal", "mv_normal_diag"):
        loc, scale = args
        z = v / scale - loc / scale
        lp = E.const(-0.5) * E.unary("square", z) - (E.const(_HALF_LOG_2PI) + E.unary("log", scale))
        if name == "mv_normal_diag" and lp.ndim == 0:
            raise ValueError("mv_normal_diag value must be a vector")
        return _sum(lp)
    if name == "half_normal":
        (scale,) = args
        z = v / scale
        lp = E.const(0.5 * math.log(2.0 / math.pi)) - E.unary("log", scale) - E.const(0.5) * E.unary("square", z)
        return E.where(v < 0.0, E.const(-math.inf), lp)
    if name in ("cauchy", "half_cauchy"):
        loc, scale = args
        z = (v - loc) / scale
        if name == "cauchy":
            return -E.unary("log1p", E.unary("square", z)) - (E.const(math.log(math.pi)) + E.unary("log", scale))
        lp = E.const(math.log(2.0 / math.pi)) - E.unary("log", scale) - E.unary("log1p", E.unary("square", z))
        return E.where(v < loc, E.const(-math.inf), lp)
    if name == "laplace":
        loc, scale = args
        return -E.unary("abs", (v - loc) / scale) - E.const(math.log(2.0)) - E.unary("log", scale)
    if name == "log_normal":
        loc, scale = args
        lv = E.unary("log", v)
        z = lv / scale - loc / scale
        lp = E.const(-0.5) * E.unary("square", z) - (E.const(_HALF_LOG_2PI) + E.unary("log", scale)) - lv
        return E.where(v > 0.0, lp, E.const(-math.inf))
    if name == "gumbel":
        loc, scale = args
        z = (v - loc) / scale
        return -(z + E.unary("exp", -z)) - E.unary("log", scale)
    if name == "weibull":
        k, scale = args
        t = E.unary("log", v) - E.unary("log", scale)
        lp = E.unary("log", k) - E.unary("log", scale) + (k - 1.0) * t - E.unary("exp", k * t)
        return E.where(v < 0.0, E.const(-math.inf), lp)
    if name == "kumaraswamy":
        a, b = args
        lv = E.unary("log", v)
        lp = E.unary("log", a) + E.unary("log", b) + (a - 1.0) * lv + (b - 1.0) * E.unary("log1p", -E.unary("exp", a * lv))
        return E.where((v < 0.0) | (v > 1.0), E.const(-math.inf), lp)
    if name == "logit_normal":
        loc, scale = args
        lv, l1 = E.unary("log", v), E.unary("log1p", -v)
        z = (lv - l1) / scale - loc / scale
        lp = E.const(-0.5) * E.unary("square", z) - (E.const(_HALF_LOG_2PI) + E.unary("log", scale)) - lv - l1
        return E.where((v > 0.0) & (v < 1.0), lp, E.const(-math.inf))
    if name == "geometric":
        (p,) = args
        return E.where(v < 0.0, E.const(-math.inf), v * E.unary("log1p", -p) + E.unary("log", p))
    if name == "inverse_gamma":
        a, b = args
        lp = a * E.unary("log", b) - E.unary("lgamma", a) - (a + 1.0) * E.unary("log", v) - b / v
        return E.where(v > 0.0, lp, E.const(-math.inf))
    if name == "chi2":
        (df,) = args
        a = 0.5 * df
        return (a - 1.0) * E.unary("log", v) - 0.5 * v - (E.unary("lgamma", a) - a * E.const(math.log(0.5)))
    if name == "student_t":
        df, loc, scale = args
        y = (v - loc) / scale
        norm = (E.unary("log", E.unary("abs", scale)) + 0.5 * E.unary("log", df) + E.const(0.5 * math.log(math.pi))
                + E.unary("lgamma", 0.5 * df) - E.unary("lgamma", 0.5 * (df + 1.0)))
        return -0.5 * (df + 1.0) * E.unary("log1p", E.unary("square", y) / df) - norm
    if name == "poisson":
        (rate,) = args
        return E.where(v < 0.0, E.const(-math.inf), v * E.unary("log", rate) - E.unary("lgamma", v + 1.0) - rate)
    if name == "exponential":
        (rate,) = args
        return E.where(v < 0.0, E.const(-math.inf), E.unary("log", rate) - rate * v)
    if name == "uniform":
        lo, hi = args
        inside = (v >= lo) & (v <= hi)
        return E.where(inside, -E.unary("log", hi - lo), E.const(-math.inf))
    if name == "gamma":
        a, rate = args
        return (a - 1.0) * E.unary("log", v) - rate * v - (E.unary("lgamma", a) - a * E.unary("log", rate))
    if name == "beta":
        a, b = args
        lbeta = E.unary("lgamma", a) + E.unary("lgamma", b) - E.unary("lgamma", a + b)
        return (a - 1.0) * E.unary("log", v) + (b - 1.0) * E.unary("log1p", -v) - lbeta
    if name == "flip":
        (p,) = args
        x = E.cast(v, F32)
        return x * E.unary("log", p) + (1.0 - x) * E.unary("log1p", -p)
    if name == "bernoulli":
        (logit,) = args
        x = E.cast(v, F32)
        return -E.unary("softplus", -logit) * x - E.unary("softplus", logit) * (1.0 - x)
    if name == "gmm_diag":
        return dist.logpdf_expr(v, args)
    if hasattr(dist, "logpdf_expr"):
        return dist.logpdf_expr(v, args)
    raise NotDifferentiable(f"no symbolic log-density for distribution {name!r}")


def model_logp(ir, site_values: dict | None = None) -> Expr:
    """Total log-density of the model = sum over sites of their log-densities
    (StaticTrace.get_score, static.py:102-105), as one scalar Expr over the
    site-value nodes of ``ir``."""
    total = None
    for s in ir.sites:
        lp = logpdf_expr(s.dist, s.value, s.args)
        flag = s.flag() if hasattr(s, "flag") else None
        if flag is not None:  # a site under a Switch branch / Mask counts only where it is valid (switch.py:171, mask.py:84)
            lp = E.where(flag, lp, E.const(0.0))
        total = lp if total is None else total + lp
    if total is None:
        return E.const(0.0)
    return total


# ------------------------------------------------------------------ reverse mode


def _unbroadcast(g: Expr, like: Expr) -> Expr:
    """Reduce a cotangent to the shape of the primal it belongs to."""
    if like.ndim == 0 and g.ndim == 1:
        return E.vsum(g)
    return g


def _zeros_like(x: Expr) -> Expr:
    if x.ndim == 0:
        return E.const(0.0)
    return Expr("constvec", (), F32, x.shape, tuple(0.0 for _ in range(x.shape[0])))


def grad(out: Expr, wrt: list) -> list:
    """d out / d w for each w in ``wrt`` (``out`` scalar).  Returns Exprs of the
    shapes of the ``wrt`` nodes (zeros when ``out`` does not depend on w)."""
    if out.ndim != 0:
        raise ValueError("grad needs a scalar output")
    order = E.topo([out])
    adj: dict[int, Expr] = {out._id: E.const(1.0)}
    dep = {w._id for w in wrt}  # nodes whose value moves with a differentiated variable
    for e in order:
        if any(x._id in dep for x in e.ins):
            dep.add(e._id)

    def acc(node: Expr, g):
        if node.dtype != F32 or node.op in ("const", "constvec"):
            return
        g = _unbroadcast(E.lift(g), node)
        cur = adj.get(node._id)
        adj[node._id] = g if cur is None else cur + g

    for e in reversed(order):
        g = adj.get(e._id)
        if g is None or not e.ins:
            continue
        op = e.op
        i = e.ins
        if op == "add":
            acc(i[0], g); acc(i[1], g)
        elif op == "sub":
            acc(i[0], g); acc(i[1], -g)
        elif op == "mul":
            acc(i[0], g * i[1]); acc(i[1], g * i[0])
        elif op == "div":
            acc(i[0], g / i[1]); acc(i[1], -(g * e) / i[1])
        elif op == "neg":
            acc(i[0], -g)
        elif op == "exp":
            acc(i[0], g * e)
        elif op == "log":
            acc(i[0], g / i[0])
        elif op == "log1p":
            acc(i[0], g / (1.0 + i[0]))
        elif op == "expm1":
            acc(i[0], g * (e + 1.0))
        elif op == "sqrt":
            acc(i[0], g / (2.0 * e))
        elif op == "square":
            acc(i[0], g * (2.0 * i[0]))
        elif op == "reciprocal":
            acc(i[0], -g * E.unary("square", e))
        elif op == "abs":
            acc(i[0], g * E.where(i[0] < 0.0, E.const(-1.0), E.const(1.0)))
        elif op == "tanh":
            acc(i[0], g * (1.0 - E.unary("square", e)))
        elif op == "sigmoid":
            acc(i[0], g * e * (1.0 - e))
        elif op == "softplus":
            acc(i[0], g * E.unary("sigmoid", i[0]))
        elif op == "sin":
            acc(i[0], g * E.unary("cos", i[0]))
        elif op == "cos":
            acc(i[0], -g * E.unary("sin", i[0]))
        elif op == "pow":
            a, b = i
            acc(a, g * b * E.binary("pow", a, b - 1.0))
            if b.op != "const":
                acc(b, g * e * E.unary("log", a))
        elif op == "min":
            acc(i[0], E.where(i[0] <= i[1], g, 0.0)); acc(i[1], E.where(i[0] <= i[1], 0.0, g))
        elif op == "max":
            acc(i[0], E.where(i[0] >= i[1], g, 0.0)); acc(i[1], E.where(i[0] >= i[1], 0.0, g))
        elif op == "where":
            acc(i[1], E.where(i[0], g, 0.0)); acc(i[2], E.where(i[0], 0.0, g))
        elif op == "sum":
            acc(i[0], g + _zeros_like(i[0]))  # broadcast the scalar cotangent
        elif op == "elem":
            (vec,) = i
            D = vec.shape[0]
            onehot = Expr("constvec", (), F32, (D,), tuple(1.0 if k == int(e.attr) else 0.0 for k in range(D)))
            acc(vec, onehot * g)
        elif op == "cast":
            if i[0].dtype == F32:
                acc(i[0], g)
        elif op in ("floor", "lt", "le", "gt", "ge", "eq", "ne", "and", "or", "logical_not", "row", "gather1"):
            pass  # piecewise constant / integer / table lookups of non-differentiated data
        elif op == "lgamma":
            if i[0]._id not in dep:
                continue  # a shape parameter fed by arguments / non-differentiated sites only: constant here
            raise NotDifferentiable("gradient through lgamma (needs digamma) is not available on the device")
        else:
            raise NotDifferentiable(f"no derivative rule for op {op!r}")

    outs = []
    for w in wrt:
        g = adj.get(w._id)
        if g is None:
            g = _zeros_like(w)
        elif g.ndim == 0 and w.ndim == 1:
            g = g + _zeros_like(w)
        outs.append(g)
    return outs

"""Analytic upper bounds of incremental log-weights, from the captured model.

A bootstrap filter's incremental weight is the sum of the log-densities of the observed sites.  When the scale
(or the support) of every observed site does not depend on particle data, the supremum of that sum over particles
is a function of the shared / scalar arguments alone:

    normal, mv_normal_diag :  -(1/2 log 2 pi + log scale)  per element   (attained at value == loc)
    half_normal            :   1/2 log(2 / pi) - log scale
    exponential            :   log rate
    uniform                :  -log(high - low)
    cauchy / half_cauchy / laplace / gumbel :  -log(pi s) / log(2 / (pi s)) / -log(2 s) / -1 - log s
    flip, bernoulli, categorical : 0

Such a bound is a valid reference maximum for the exact integer weight masses (any ``M >= max w`` that every CTA and
every rank agrees on): it removes the running-max pass and its grid-wide synchronisation from the filter step
(properties pinned in tests/test_oracle_pull_resampling.py).  This module only DERIVES and
EVALUATES the bound on the host; ``ParticleFilter.weight_upper_bound`` reports it, no kernel consumes it yet.
"""

from __future__ import annotations

import math

import numpy as np

from . import expr as E
from .capture import ModelIR
from .expr import Expr

_HALF_LOG_2PI = 0.5 * math.log(2.0 * math.pi)


def is_particle_invariant(e: Expr) -> bool:
    """True when ``e`` reads no random choice and no per-particle argument."""
    for node in E.topo([e]):
        if node.op == "site":
            return False
        if node.op == "arg" and node.attr["kind"] == "particle":
            return False
    return True


def _site_bound(site, width: int) -> Expr | None:
    name = site.dist.name
    a = site.args
    if name in ("flip", "bernoulli", "categorical"):
        return E.const(0.0)
    if name in ("normal", "mv_normal_diag"):
        scale = a[1]
        if not is_particle_invariant(scale):
            return None
        per = E.const(-_HALF_LOG_2PI) - E.unary("log", scale)
        if per.ndim:
            return E.vsum(per)
        return per * float(width) if width else per  # a scalar scale shared by every element of a vector site
    if name == "half_normal":
        return E.const(0.5 * math.log(2.0 / math.pi)) - E.unary("log", a[0]) if is_particle_invariant(a[0]) else None
    if name in ("cauchy", "half_cauchy", "laplace"):  # the mode sits at loc: -log(pi s), log(2 / (pi s)), -log(2 s)
        if not is_particle_invariant(a[1]):
            return None
        c = {"cauchy": -math.log(math.pi), "half_cauchy": math.log(2.0 / math.pi), "laplace": -math.log(2.0)}[name]
        return E.const(c) - E.unary("log", a[1])
    if name == "gumbel":  # mode at loc: -1 - log s
        return E.const(-1.0) - E.unary("log", a[1]) if is_particle_invariant(a[1]) else None
    if name == "exponential":
        return E.unary("log", a[0]) if is_particle_invariant(a[0]) else None
    if name == "uniform":
        if is_particle_invariant(a[0]) and is_particle_invariant(a[1]):
            return -E.unary("log", a[1] - a[0])
        return None
    return None


def log_weight_upper_bound(ir: ModelIR, weighted_sites) -> Expr | None:
    """Scalar ``Expr`` over shared / scalar arguments bounding ``sum_j logpdf_j`` over the given sites, or None."""
    total = None
    for j in weighted_sites:
        s = ir.sites[j]
        width = s.value.shape[0] if s.value.ndim else 0
        b = _site_bound(s, width)
        if b is None:
            return None
        total = b if total is None else total + b
    return total if total is not None else E.const(0.0)


def evaluate_invariant(e: Expr, arg_values: dict):
    """Evaluate a particle-invariant expression on the host in float32; ``arg_values[i]`` = value of argument leaf i
    (Python number or array)."""
    F = np.float32
    cache: dict = {}

    def ev(x: Expr):
        if x._id in cache:
            return cache[x._id]
        ins = [ev(i) for i in x.ins]
        op = x.op
        if op == "const":
            v = F(x.attr)
        elif op == "constvec":
            v = np.asarray(x.attr, dtype=F)
        elif op == "arg":
            v = np.asarray(arg_values[x.attr["index"]], dtype=F)
        elif op == "sum":
            v = np.sum(ins[0], dtype=F)
        elif op == "elem":
            v = ins[0][..., int(x.attr)]
        elif op == "cast":
            v = np.asarray(ins[0], dtype=F)
        elif op in ("add", "sub", "mul", "div", "pow", "min", "max"):
            f = {"add": np.add, "sub": np.subtract, "mul": np.multiply, "div": np.divide, "pow": np.power,
                 "min": np.minimum, "max": np.maximum}[op]
            v = f(ins[0], ins[1]).astype(F)
        elif op in ("neg", "exp", "log", "sqrt", "abs", "square", "reciprocal", "log1p", "expm1", "tanh"):
            f = {"neg": np.negative, "exp": np.exp, "log": np.log, "sqrt": np.sqrt, "abs": np.abs, "square": np.square,
                 "reciprocal": lambda t: F(1) / t, "log1p": np.log1p, "expm1": np.expm1, "tanh": np.tanh}[op]
            v = np.asarray(f(ins[0]), dtype=F)
        else:
            raise NotImplementedError(f"evaluate_invariant: operation {op}")
        cache[x._id] = v
        return v

    return float(ev(e))

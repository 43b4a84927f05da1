"""``jnp``-style math for ``@gen`` bodies: works on traced values (``Expr``),
Python numbers and torch tensors.  (The reference's bodies use ``jax.numpy``;
import this module as ``jnp`` to port a model unchanged.)"""

from __future__ import annotations

import math as _math

import torch as _torch

from . import expr as _E
from .expr import Expr as _Expr


def _un(op, tfn, mfn):
    def f(x):
        if isinstance(x, _Expr):
            return _E.unary(op, x)
        if isinstance(x, _torch.Tensor):
            return tfn(x)
        return mfn(x)

    f.__name__ = op
    return f


exp = _un("exp", _torch.exp, _math.exp)
log = _un("log", _torch.log, _math.log)
sqrt = _un("sqrt", _torch.sqrt, _math.sqrt)
abs = _un("abs", _torch.abs, _math.fabs)  # noqa: A001
tanh = _un("tanh", _torch.tanh, _math.tanh)
log1p = _un("log1p", _torch.log1p, _math.log1p)
expm1 = _un("expm1", _torch.expm1, _math.expm1)
floor = _un("floor", _torch.floor, _math.floor)
sin = _un("sin", _torch.sin, _math.sin)
cos = _un("cos", _torch.cos, _math.cos)
square = _un("square", _torch.square, lambda x: x * x)
sigmoid = _un("sigmoid", _torch.sigmoid, lambda x: 1.0 / (1.0 + _math.exp(-x)))
softplus = _un("softplus", _torch.nn.functional.softplus, lambda x: _math.log1p(_math.exp(x)))
lgamma = _un("lgamma", _torch.lgamma, _math.lgamma)


def _bin(op, tfn):
    def f(a, b):
        if isinstance(a, _Expr) or isinstance(b, _Expr):
            return _E.binary(op, a, b)
        return tfn(_torch.as_tensor(a), _torch.as_tensor(b))

    f.__name__ = op
    return f


minimum = _bin("min", _torch.minimum)
maximum = _bin("max", _torch.maximum)
power = _bin("pow", _torch.pow)


def where(c, a, b):
    if any(isinstance(x, _Expr) for x in (c, a, b)):
        return _E.where(c, a, b)
    return _torch.where(_torch.as_tensor(c), _torch.as_tensor(a), _torch.as_tensor(b))


def sum(x, axis=None):  # noqa: A001
    if isinstance(x, _Expr):
        return _E.vsum(x)
    return _torch.sum(_torch.as_tensor(x)) if axis is None else _torch.sum(_torch.as_tensor(x), dim=axis)


class _ScalarType:
    """``jnp.int32`` / ``jnp.float32``: a dtype that also converts when called (``jnp.int32(flag)`` on a traced value)."""

    def __init__(self, name: str, torch_dtype):
        self.name, self.torch = name, torch_dtype

    def __call__(self, x):
        if isinstance(x, _Expr):
            return _E.cast(x, self.name)
        if isinstance(x, _torch.Tensor):
            return x.to(self.torch)
        return int(x) if self.name == "int32" else float(x)

    def __repr__(self):
        return self.name

    def __eq__(self, other):
        return other is self or other == self.torch

    __hash__ = object.__hash__


def _tdt(dtype):
    return getattr(dtype, "torch", dtype)


def array(x, dtype=None):
    if isinstance(x, _Expr):
        return x if dtype is None else _E.cast(x, str(dtype))
    return _torch.as_tensor(x, dtype=_tdt(dtype))


asarray = array
float32 = _ScalarType("float32", _torch.float32)
int32 = _ScalarType("int32", _torch.int32)
pi = _math.pi
inf = _math.inf


# ---- composites of the traced operations above (no new kernel-side operations) --------------------------------


def _is_traced(*xs) -> bool:
    return any(isinstance(x, _Expr) for x in xs)


def negative(x):
    return -x


def reciprocal(x):
    if isinstance(x, _Expr):
        return _E.unary("reciprocal", x)
    return 1.0 / x


def add(a, b):
    return a + b


def subtract(a, b):
    return a - b


def multiply(a, b):
    return a * b


def divide(a, b):
    return a / b


def clip(x, a_min=None, a_max=None):
    """``jnp.clip``: ``minimum(maximum(x, a_min), a_max)``."""
    if a_min is not None:
        x = maximum(x, a_min)
    if a_max is not None:
        x = minimum(x, a_max)
    return x


def logaddexp(a, b):
    """``log(exp(a) + exp(b))`` without overflow: ``max(a, b) + log1p(exp(-|a - b|))``."""
    if not _is_traced(a, b):
        return _torch.logaddexp(_torch.as_tensor(a, dtype=_torch.float32), _torch.as_tensor(b, dtype=_torch.float32))
    return maximum(a, b) + log1p(exp(-abs(a - b)))


def mean(x, axis=None):
    if isinstance(x, _Expr):
        return _E.vsum(x) / float(x.shape[0]) if x.ndim else x
    t = _torch.as_tensor(x, dtype=_torch.float32)
    return _torch.mean(t) if axis is None else _torch.mean(t, dim=axis)


def dot(a, b):
    """Inner product of two vectors (``sum(a * b)``)."""
    if _is_traced(a, b):
        return _E.vsum(a * b)
    return _torch.dot(_torch.as_tensor(a, dtype=_torch.float32), _torch.as_tensor(b, dtype=_torch.float32))


def logical_and(a, b):
    if _is_traced(a, b):
        return _E.binary("and", a, b)
    return _torch.logical_and(_torch.as_tensor(a), _torch.as_tensor(b))


def logical_or(a, b):
    if _is_traced(a, b):
        return _E.binary("or", a, b)
    return _torch.logical_or(_torch.as_tensor(a), _torch.as_tensor(b))


def logical_not(a):
    if isinstance(a, _Expr):
        return _E.unary("logical_not", a)
    return _torch.logical_not(_torch.as_tensor(a))


def zeros(shape, dtype=None):
    return _torch.zeros(shape, dtype=_tdt(dtype) or _torch.float32)


def ones(shape, dtype=None):
    return _torch.ones(shape, dtype=_tdt(dtype) or _torch.float32)


def arange(*args, dtype=None):
    return _torch.arange(*args, dtype=_tdt(dtype))


e = _math.e

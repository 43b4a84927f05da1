"""Symbolic values recorded while a ``@gen`` body runs once (model capture).

This replaces the reference's jaxpr staging (core/compiler/staging.py:286,
initial_style_primitive.py:31-57; ``trace_p`` sites at static.py:156-193): the
body is executed with ``Expr`` arguments, every ``dist(args) @ "addr"`` records
a site and returns the site value as another ``Expr``; arithmetic builds a
small SSA DAG that ``codegen.py`` turns into one fused CUDA kernel.

Like the static language itself (static.py:737-739) Python control flow must
not depend on traced values: ``bool(expr)`` raises.
"""

from __future__ import annotations

import math
from typing import Any

F32 = "f32"
I32 = "i32"

_UNARY = {
    "neg", "exp", "log", "sqrt", "abs", "tanh", "sigmoid", "log1p", "expm1", "square", "floor",
    "sin", "cos", "softplus", "lgamma", "logical_not", "reciprocal",
}
_BINARY = {"add", "sub", "mul", "div", "pow", "min", "max", "lt", "le", "gt", "ge", "eq", "ne", "and", "or"}
_CMP = {"lt", "le", "gt", "ge", "eq", "ne", "and", "or"}


class TracedControlFlow(TypeError):
    pass


class Expr:
    """One node of the captured DAG."""

    __slots__ = ("op", "ins", "dtype", "shape", "attr", "_id")
    _counter = 0

    def __init__(self, op: str, ins: tuple = (), dtype: str = F32, shape: tuple = (), attr: Any = None):
        self.op = op
        self.ins = tuple(ins)
        self.dtype = dtype
        self.shape = tuple(shape)
        self.attr = attr
        Expr._counter += 1
        self._id = Expr._counter

    # -------------------------------------------------------------- helpers
    @property
    def ndim(self):
        return len(self.shape)

    def __repr__(self):
        return f"Expr<{self.op}#{self._id} {self.dtype}{list(self.shape)}>"

    def __bool__(self):
        raise TracedControlFlow(
            "Python control flow on a traced value inside @gen: the static language needs "
            "deterministic control flow (reference static.py:737-739); use where()/select instead"
        )

    def __hash__(self):
        return self._id

    def __iter__(self):
        if self.ndim == 0:
            raise TypeError("cannot iterate a scalar traced value")
        return (self[i] for i in range(self.shape[0]))

    def __len__(self):
        if self.ndim == 0:
            raise TypeError("len() of a scalar traced value")
        return self.shape[0]

    # ----------------------------------------------------------- arithmetic
    def __add__(self, o): return binary("add", self, o)
    def __radd__(self, o): return binary("add", o, self)
    def __sub__(self, o): return binary("sub", self, o)
    def __rsub__(self, o): return binary("sub", o, self)
    def __mul__(self, o): return binary("mul", self, o)
    def __rmul__(self, o): return binary("mul", o, self)
    def __truediv__(self, o): return binary("div", self, o)
    def __rtruediv__(self, o): return binary("div", o, self)
    def __pow__(self, o): return _pow(self, o)
    def __rpow__(self, o): return binary("pow", o, self)
    def __neg__(self): return unary("neg", self)
    def __pos__(self): return self
    def __abs__(self): return unary("abs", self)
    def __lt__(self, o): return binary("lt", self, o)
    def __le__(self, o): return binary("le", self, o)
    def __gt__(self, o): return binary("gt", self, o)
    def __ge__(self, o): return binary("ge", self, o)
    def __eq__(self, o): return binary("eq", self, o)  # type: ignore[override]
    def __ne__(self, o): return binary("ne", self, o)  # type: ignore[override]
    def __and__(self, o): return binary("and", self, o)
    def __or__(self, o): return binary("or", self, o)
    def __invert__(self): return unary("logical_not", self)

    def astype(self, dtype):
        return cast(self, dtype)

    def sum(self, axis=None):
        return vsum(self)

    def __getitem__(self, idx):
        return index(self, idx)


def const(v) -> Expr:
    if isinstance(v, bool):
        return Expr("const", (), I32, (), int(v))
    if isinstance(v, int):
        return Expr("const", (), I32, (), int(v))
    if isinstance(v, float):
        return Expr("const", (), F32, (), float(v))
    raise TypeError(f"cannot lift {type(v).__name__} into a traced constant")


def lift(v) -> Expr:
    """Python number / 0-d or 1-d tensor-like -> Expr."""
    if isinstance(v, Expr):
        return v
    if isinstance(v, (bool, int, float)):
        return const(v)
    # numpy / torch scalars and small vectors become literals
    try:
        import numpy as np

        a = np.asarray(v.detach().cpu().numpy() if hasattr(v, "detach") else v)
    except Exception as e:  # pragma: no cover
        raise TypeError(f"cannot lift {type(v).__name__} into the captured model") from e
    if a.ndim == 0:
        return const(a.item())
    if a.ndim == 1 and a.size <= 64:
        kind = I32 if a.dtype.kind in "iub" else F32
        vals = tuple(int(x) if kind == I32 else float(x) for x in a.tolist())
        return Expr("constvec", (), kind, (a.size,), vals)
    raise TypeError(
        "tensors closed over by a @gen body must be passed as model arguments "
        f"(got a constant of shape {a.shape})"
    )


def _bshape(a: Expr, b: Expr) -> tuple:
    if a.shape == b.shape:
        return a.shape
    if a.shape == ():
        return b.shape
    if b.shape == ():
        return a.shape
    raise ValueError(f"shape mismatch in traced arithmetic: {a.shape} vs {b.shape}")


def _tofloat(x: Expr) -> Expr:
    return x if x.dtype == F32 else cast(x, F32)


def unary(op: str, x) -> Expr:
    assert op in _UNARY, op
    x = lift(x)
    if op == "logical_not":
        return Expr("logical_not", (x,), I32, x.shape)
    if op in ("neg", "abs") and x.dtype == I32:
        return Expr(op, (x,), I32, x.shape)
    x = _tofloat(x)
    return Expr(op, (x,), F32, x.shape)


def binary(op: str, a, b) -> Expr:
    assert op in _BINARY, op
    a, b = lift(a), lift(b)
    shape = _bshape(a, b)
    if op in _CMP:
        if a.dtype != b.dtype:
            a, b = _tofloat(a), _tofloat(b)
        return Expr(op, (a, b), I32, shape)
    if a.dtype == I32 and b.dtype == I32 and op in ("add", "sub", "mul", "min", "max"):
        return Expr(op, (a, b), I32, shape)
    return Expr(op, (_tofloat(a), _tofloat(b)), F32, shape)


def _pow(a, b) -> Expr:
    if isinstance(b, (int, float)) and float(b) == 2.0:
        return unary("square", a)
    return binary("pow", a, b)


def where(c, a, b) -> Expr:
    c, a, b = lift(c), lift(a), lift(b)
    if a.dtype != b.dtype:
        a, b = _tofloat(a), _tofloat(b)
    shape = _bshape(Expr("tmp", (), F32, _bshape(c, a)), b)
    return Expr("where", (c, a, b), a.dtype, shape)


def cast(x, dtype) -> Expr:
    x = lift(x)
    d = {float: F32, int: I32, bool: I32}.get(dtype, dtype)
    if str(d) in ("torch.float32", "float32", "f32"):
        d = F32
    elif str(d) in ("torch.int32", "int32", "i32", "torch.int64", "int64", "torch.bool", "bool"):
        d = I32
    if d not in (F32, I32):
        raise TypeError(f"unsupported dtype {dtype!r}")
    if x.dtype == d:
        return x
    return Expr("cast", (x,), d, x.shape)


def vsum(x) -> Expr:
    x = lift(x)
    if x.ndim == 0:
        return x
    return Expr("sum", (_tofloat(x),), F32, ())


def index(x: Expr, idx) -> Expr:
    """``table[i]``: row of a shared matrix / element of a shared vector (dynamic i),
    or a static element of any vector value."""
    if x.ndim == 0:
        raise IndexError("cannot index a scalar traced value")
    if isinstance(idx, tuple):
        out = x
        for i in idx:
            out = index(out, i)
        return out
    if isinstance(idx, int):
        if x.op == "arg" and x.attr["kind"] == "shared" and x.ndim == 2:
            return Expr("row", (x, const(idx)), x.dtype, (x.shape[1],))
        n = x.shape[0]
        if not -n <= idx < n:
            raise IndexError(idx)
        return Expr("elem", (x,), x.dtype, (), idx % n)
    idx = lift(idx)
    if idx.dtype != I32 or idx.ndim != 0:
        raise TypeError("dynamic index must be a scalar integer traced value")
    if not (x.op in ("arg",) and x.attr["kind"] == "shared"):
        raise NotImplementedError("dynamic indexing is supported on shared (un-batched) arguments only")
    if x.ndim == 2:
        return Expr("row", (x, idx), x.dtype, (x.shape[1],))
    return Expr("gather1", (x, idx), x.dtype, ())


def is_expr(x) -> bool:
    return isinstance(x, Expr)


def topo(roots) -> list[Expr]:
    """Dependency-ordered unique nodes reachable from ``roots``."""
    seen: dict[int, Expr] = {}
    order: list[Expr] = []

    def visit(e: Expr):
        if e._id in seen:
            return
        seen[e._id] = e
        for i in e.ins:
            visit(i)
        order.append(e)

    for r in roots:
        visit(r)
    return order


def fmt_float(v: float) -> str:
    if math.isinf(v):
        return "INFINITY" if v > 0 else "(-INFINITY)"
    if math.isnan(v):
        return "NAN"
    import numpy as np

    return float(np.float32(v)).hex() + "f"

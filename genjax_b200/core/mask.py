"""``Mask``: a value with a validity flag.

API mirror of src/genjax/_src/core/generative/functional_types.py:43-365
(``Mask(value, flag)``, ``build`` :145, ``maybe_mask`` :172, ``flatten`` :211,
``unmask`` :233, ``|`` :309, ``^`` :321, ``~`` :340, ``or_n`` / ``xor_n``).
Flags are Python bools, torch bool / int tensors carrying the particle axis, or
-- while a ``@gen`` body is being captured -- traced ``Expr`` predicates.
"""

from __future__ import annotations

import functools

import torch

__all__ = ["Mask"]


def _and(a, b):
    if isinstance(a, bool):
        return b if a else False
    if isinstance(b, bool):
        return a if b else False
    if isinstance(a, torch.Tensor) and isinstance(b, torch.Tensor):
        return a.to(torch.bool) & b.to(torch.bool)
    return a & b  # traced predicates


def _not(a):
    if isinstance(a, bool):
        return not a
    if isinstance(a, torch.Tensor):
        return ~a.to(torch.bool)
    return ~a


def _shape(v):
    return tuple(v.shape) if hasattr(v, "shape") else ()


def _expand(flag: torch.Tensor, like) -> torch.Tensor:
    f = flag.to(torch.bool)
    if isinstance(like, torch.Tensor):
        f = f.to(like.device)
        while f.ndim < like.ndim:
            f = f.unsqueeze(-1)
    return f


def _where(flag, a, b):
    """Leaf-wise ``where(flag, a, b)`` over matching pytrees of tensors / numbers."""
    if isinstance(a, (tuple, list)):
        return type(a)(_where(flag, x, y) for x, y in zip(a, b))
    if isinstance(a, dict):
        return {k: _where(flag, a[k], b[k]) for k in a}
    ta = a if isinstance(a, torch.Tensor) else torch.as_tensor(a)
    tb = b if isinstance(b, torch.Tensor) else torch.as_tensor(b)
    f = flag if isinstance(flag, torch.Tensor) else torch.as_tensor(flag)
    dev = next((t.device for t in (ta, tb, f) if t.device.type != "cpu"), ta.device)
    ta, tb, f = ta.to(dev), tb.to(dev), f.to(dev)
    return torch.where(_expand(f, ta if ta.ndim >= tb.ndim else tb), ta, tb)


class Mask:
    """functional_types.py:43: ``value`` is only meaningful where ``flag`` holds."""

    __slots__ = ("value", "flag")

    def __init__(self, value, flag=True):
        assert not isinstance(value, Mask), f"Mask should not be instantiated with another Mask! found {value}"
        self.value = value
        self.flag = flag

    # -- constructors ---------------------------------------------------
    @staticmethod
    def build(v, f=True) -> "Mask":
        """functional_types.py:145-169: masking a Mask ANDs the flags."""
        if isinstance(v, Mask):
            return Mask(v.value, _and(f, v.flag))
        return Mask(v, f)

    @staticmethod
    def maybe_mask(v, f):
        return Mask.build(v, f).flatten()

    # -- accessors ------------------------------------------------------
    def primal_flag(self):
        return getattr(self.flag, "primal", self.flag) if type(self.flag).__name__ == "Diff" else self.flag

    def flatten(self):
        """functional_types.py:211-231: a concrete flag dissolves the mask."""
        f = self.primal_flag()
        if isinstance(f, bool):
            return self.value if (f and self.value is not None) else None
        return self

    def unmask(self, default=None):
        """functional_types.py:233-260.  Without ``default`` the value is handed back as is (the reference checks the
        flag only under ``checkify``); with one, invalid entries read ``default``."""
        if default is None:
            return self.value
        return _where(self.primal_flag(), self.value, default)

    def __getitem__(self, path) -> "Mask":
        path = path if isinstance(path, tuple) else (path,)
        f = self.primal_flag()
        if isinstance(f, torch.Tensor) and f.ndim:
            f = f[path[: f.ndim]]
        v = self.value[path] if not isinstance(self.value, (tuple, list)) else type(self.value)(x[path] for x in self.value)
        return Mask.build(v, f)

    # -- combinators -----------------------------------------------------
    def _check_shapes(self, other: "Mask"):
        a, b = _shape(self.value), _shape(other.value)
        fa, fb = _shape(self.primal_flag()), _shape(other.primal_flag())
        # value shapes past the flag's (vectorised) prefix must agree (functional_types.py:110-141)
        if a[len(fa):] != b[len(fb):]:
            raise ValueError(f"Cannot combine masks with different array shapes: {a[len(fa):]} vs {b[len(fb):]}")

    def __or__(self, other: "Mask") -> "Mask":
        self._check_shapes(other)
        f, g = self.primal_flag(), other.primal_flag()
        if isinstance(f, bool):
            return self if f else other
        if isinstance(g, bool):
            g = torch.full_like(f.to(torch.bool), g)
        f = f.to(torch.bool)
        g = g.to(torch.bool).to(f.device)
        return Mask(_where(f, self.value, other.value), f | g)

    def __xor__(self, other: "Mask") -> "Mask":
        self._check_shapes(other)
        f, g = self.primal_flag(), other.primal_flag()
        if isinstance(f, bool) and isinstance(g, bool):
            if f == g:
                return Mask.build(self, False)
            return self if f else other
        ft = f if isinstance(f, torch.Tensor) else torch.full_like(g.to(torch.bool), f)
        gt = g if isinstance(g, torch.Tensor) else torch.full_like(ft.to(torch.bool), g)
        ft = ft.to(torch.bool)
        gt = gt.to(torch.bool).to(ft.device)
        return Mask(_where(ft, self.value, other.value), ft ^ gt)

    def __invert__(self) -> "Mask":
        return Mask(self.value, _not(self.flag))

    @staticmethod
    def or_n(mask: "Mask", *masks: "Mask") -> "Mask":
        return functools.reduce(lambda a, b: a | b, masks, mask)

    @staticmethod
    def xor_n(mask: "Mask", *masks: "Mask") -> "Mask":
        return functools.reduce(lambda a, b: a ^ b, masks, mask)

    # -- comparison / printing ---------------------------------------------
    def __eq__(self, other):
        if not isinstance(other, Mask):
            return NotImplemented
        return _tree_equal(self.value, other.value) and _tree_equal(self.primal_flag(), other.primal_flag())

    __hash__ = object.__hash__

    def __repr__(self):
        return f"Mask({self.value!r}, {self.flag!r})"


def _tree_equal(a, b) -> bool:
    if isinstance(a, (tuple, list)) and isinstance(b, (tuple, list)):
        return len(a) == len(b) and all(_tree_equal(x, y) for x, y in zip(a, b))
    if isinstance(a, torch.Tensor) or isinstance(b, torch.Tensor):
        ta = a if isinstance(a, torch.Tensor) else torch.as_tensor(a)
        tb = b if isinstance(b, torch.Tensor) else torch.as_tensor(b)
        if ta.shape != tb.shape:
            return False
        return bool((ta.cpu().to(torch.float64) == tb.cpu().to(torch.float64)).all())
    try:
        return bool(a == b)
    except Exception:
        return False

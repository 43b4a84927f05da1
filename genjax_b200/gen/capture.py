"""Model capture: run a ``@gen`` body once with symbolic arguments and record
its random-choice sites and dataflow as a ``ModelIR``.

Stands in for the reference's ``stage`` + ``trace_p`` handler machinery
(core/compiler/staging.py:286; generative_functions/static.py:156-193,
209-246): the reference re-interprets a jaxpr once per GFI method, this build
captures one IR per argument signature and lets the per-site flags of the
launch select simulate / assess / generate / update / regenerate behaviour.
"""

from __future__ import annotations

import contextvars
import hashlib
import json
import dataclasses
from dataclasses import dataclass
from typing import Any, Callable

from . import expr as E
from .expr import Expr


class AddressReuse(Exception):
    """Attempt to re-write an address in a static model (static.py:139-144, 213-216)."""

    def __init__(self, addr):
        self.addr = addr
        super().__init__(addr)


class MissingAddress(Exception):
    """``assess`` found no value for a site (static.py:147-152, 317-318)."""

    def __init__(self, addr):
        self.addr = addr
        super().__init__(addr)


@dataclass(frozen=True)
class ArgSpec:
    """One flattened model-argument leaf.

    kind: "scalar" (host number passed by value), "particle" (tensor with a
    leading particle axis) or "shared" (tensor shared by all particles).
    shape: event shape (without the particle axis)."""

    kind: str
    dtype: str
    shape: tuple

    def key(self):
        return (self.kind, self.dtype, self.shape)


@dataclass
class SiteSpec:
    index: int
    addr: tuple
    dist: Any  # Distribution
    args: list  # Expr per canonical distribution argument
    value: Expr  # the site's value node


@dataclass
class ModelIR:
    name: str
    args: list  # ArgSpec
    arg_exprs: list  # Expr per arg leaf
    sites: list  # SiteSpec
    ret_leaves: list  # Expr | python constant
    ret_tree: Any  # structure to rebuild the retval
    width: int = 0  # vector event width D (0 = all scalar)
    digest: str = ""
    subcalls: dict = dataclasses.field(default_factory=dict)  # address prefix of a nested @gen call -> (arg leaves, ret leaves)

    def site_index(self, addr: tuple) -> int:
        for s in self.sites:
            if s.addr == addr:
                return s.index
        raise KeyError(addr)

    def addresses(self) -> list:
        return [s.addr for s in self.sites]


_CAPTURE: contextvars.ContextVar = contextvars.ContextVar("genjax_b200_capture", default=None)


class _Capture:
    def __init__(self):
        self.sites: list[SiteSpec] = []
        self.prefix: tuple = ()
        self.addrs: set = set()
        self.subcalls: dict = {}

    def record(self, addr, dist, args) -> Expr:
        full = self.prefix + addr
        if full in self.addrs:
            raise AddressReuse(full if len(full) > 1 else full[0])
        self.addrs.add(full)
        cargs = dist.canonical_args(args)
        dtype, shape = dist.value_type(cargs)
        idx = len(self.sites)
        v = Expr("site", (), dtype, shape, idx)
        self.sites.append(SiteSpec(idx, full, dist, cargs, v))
        return v


def current_capture() -> _Capture | None:
    return _CAPTURE.get()


def trace_site(addr, gen_fn, args):
    """``gen_fn(*args) @ addr`` inside a ``@gen`` body (static.py:175-193)."""
    cap = _CAPTURE.get()
    if cap is None:
        raise RuntimeError("`gen_fn(*args) @ addr` can only be used inside a @gen function body")
    from ..core.choice_map import _norm_addr

    addr = _norm_addr(addr)
    if hasattr(gen_fn, "capture_inline"):
        # nested @gen call: inline its sites under the address prefix
        old = cap.prefix
        cap.prefix = old + addr
        try:
            ret = gen_fn.capture_inline(args)
        finally:
            cap.prefix = old
        # what flows into and out of the callee: change propagation for edit requests addressed at it
        cap.subcalls[old + addr] = (flatten(args)[0], flatten(ret)[0])
        return ret
    return cap.record(addr, gen_fn, args)


# ----------------------------------------------------------------- pytrees


def flatten(tree) -> tuple[list, Any]:
    """Minimal pytree flatten over tuple / list / dict."""
    leaves: list = []

    def go(t):
        from ..core.choice_map import ChoiceMap

        if isinstance(t, ChoiceMap):
            # a choice map is a pytree of its leaves (the reference's ChoiceMap is a Pytree, choice_map.py:847)
            return ("chm", [(addr, go(v)) for addr, v in t.leaves()])
        if type(t).__name__ == "Target" and hasattr(t, "constraint") and hasattr(t, "p"):
            # Target(p, args, constraint): p is static, args and constraint are traced (sp.py:53-81)
            return ("target", (t.p, go(t.args), go(t.constraint)))
        if dataclasses.is_dataclass(t) and not isinstance(t, type):
            # user pytrees (the reference's Pytree.dataclass): fields are children, the class is static
            return ("dataclass", (type(t), [(f.name, go(getattr(t, f.name))) for f in dataclasses.fields(t)]))
        if isinstance(t, tuple):
            return ("tuple", [go(x) for x in t])
        if isinstance(t, list):
            return ("list", [go(x) for x in t])
        if isinstance(t, dict):
            return ("dict", [(k, go(v)) for k, v in t.items()])
        if t is None:
            return ("none", None)
        leaves.append(t)
        return ("leaf", len(leaves) - 1)

    return leaves, go(tree)


def unflatten(tree, leaves):
    kind, payload = tree
    if kind == "chm":
        from ..core.choice_map import ChoiceMap

        out = ChoiceMap.empty()
        for addr, sub in payload:
            out = out | ChoiceMap.entry(unflatten(sub, leaves), *addr)
        return out
    if kind == "target":
        from ..inference.sp import Target

        p, a, c = payload
        return Target(p, unflatten(a, leaves), unflatten(c, leaves))
    if kind == "dataclass":
        cls, fields = payload
        return cls(**{name: unflatten(sub, leaves) for name, sub in fields})
    if kind == "tuple":
        return tuple(unflatten(x, leaves) for x in payload)
    if kind == "list":
        return [unflatten(x, leaves) for x in payload]
    if kind == "dict":
        return {k: unflatten(v, leaves) for k, v in payload}
    if kind == "none":
        return None
    return leaves[payload]


def capture(source: Callable, name: str, arg_specs: list, arg_tree) -> ModelIR:
    """Run ``source`` on symbolic args built from ``arg_specs``."""
    arg_exprs = [
        Expr("arg", (), spec.dtype, spec.shape, {"index": i, "kind": spec.kind}) for i, spec in enumerate(arg_specs)
    ]
    sym_args = unflatten(arg_tree, arg_exprs)
    cap = _Capture()
    tok = _CAPTURE.set(cap)
    try:
        ret = source(*sym_args)
    finally:
        _CAPTURE.reset(tok)
    ret_leaves, ret_tree = flatten(ret)
    widths = set()
    for s in cap.sites:
        if s.value.ndim == 1:
            widths.add(s.value.shape[0])
    for spec in arg_specs:
        if spec.kind == "particle" and len(spec.shape) == 1:
            widths.add(spec.shape[0])
    for r in ret_leaves:
        if isinstance(r, Expr) and r.ndim == 1:
            widths.add(r.shape[0])
    if len(widths) > 1:
        raise NotImplementedError(f"all vector-valued choices/arguments of one model must share one width, got {widths}")
    ir = ModelIR(name, list(arg_specs), arg_exprs, cap.sites, ret_leaves, ret_tree, width=(widths.pop() if widths else 0),
                 subcalls=cap.subcalls)
    if len(ir.sites) > 16:
        raise NotImplementedError("more than 16 random-choice sites in one static model")
    return ir


def ir_fingerprint(ir: ModelIR) -> str:
    """Structural hash (used to key compiled kernels)."""
    roots = []
    for s in ir.sites:
        roots.extend(s.args)
    roots.extend(r for r in ir.ret_leaves if isinstance(r, Expr))
    order = E.topo(roots)
    ids = {e._id: i for i, e in enumerate(order)}
    desc = {
        "args": [a.key() for a in ir.args],
        "nodes": [
            (e.op, [ids[i._id] for i in e.ins], e.dtype, e.shape, e.attr if e.op != "arg" else e.attr["index"])
            for e in order
        ],
        "sites": [(s.addr, s.dist.name, [ids[a._id] for a in s.args]) for s in ir.sites],
        "rets": [ids[r._id] if isinstance(r, Expr) else ("const", repr(r)) for r in ir.ret_leaves],
    }
    return hashlib.sha256(json.dumps(desc, default=str).encode()).hexdigest()[:16]

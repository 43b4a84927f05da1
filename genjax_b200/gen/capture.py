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
from .expr import I32 as I32_


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
    # dynamic structure (combinators/switch.py, mask.py): a site inside a Switch branch exists only where ``live`` holds
    # (elsewhere its value reads 0, as the zero-filled sub-traces of the unselected branches, switch.py:171-180); a site
    # under a MaskCombinator is always visited but contributes its log-density only where ``scored`` holds (mask.py:84).
    live: Expr | None = None
    scored: Expr | None = None
    excl: tuple = ()  # ((switch id, branch), ...) -- two sites may share an address iff they sit in different branches
    stack: tuple = ()  # positions in ``addr`` that are the step index of an unrolled Scan: the choice map shows the
    #                    steps stacked along an axis under the address without them (``chm["tracks", :, "x"]``)
    cmask: Expr | None = None  # hidden per-particle argument: bit 0 = take the supplied value (else sample), bit 1 = the
    #                            site's log-density stays out of the weight (Mask-ed constraints, distribution.py:129-142)

    @property
    def sel_addr(self) -> tuple:
        """The address a Selection sees: the step index of an unrolled Scan is transparent to it (Scan hands the same
        selection to every step, scan.py:417-507)."""
        return tuple(a for i, a in enumerate(self.addr) if i not in self.stack) if self.stack else self.addr

    def flag(self) -> Expr | None:
        """Where the choice is valid: what ``chm.mask(flag)`` / ``ChoiceMap.switch`` report for it."""
        if self.live is None:
            return self.scored
        return self.live if self.scored is None else E.binary("and", self.live, self.scored)


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
    flag_leaves: list = dataclasses.field(default_factory=list)  # validity predicates written out after the return leaves
    flag_of: dict = dataclasses.field(default_factory=dict)  # site index -> position in flag_leaves
    n_user_args: int = -1  # model arguments proper; the rest are hidden constraint-mask arguments (SiteSpec.cmask)

    def site_index(self, addr: tuple) -> int:
        hits = [s.index for s in self.sites if s.addr == addr]
        if len(hits) == 1:
            return hits[0]
        if hits:
            raise NotImplementedError(f"address {addr} names a site in several Switch branches")
        raise KeyError(addr)

    @property
    def dynamic(self) -> bool:
        """Does the model hold sites under a Switch / MaskCombinator or Mask-ed constraints?"""
        return bool(self.flag_of) or any(s.cmask is not None for s in self.sites)

    def addresses(self) -> list:
        return [s.addr for s in self.sites]


_CAPTURE: contextvars.ContextVar = contextvars.ContextVar("genjax_b200_capture", default=None)


class _Capture:
    def __init__(self, cmask_addrs=(), n_args: int = 0):
        self.sites: list[SiteSpec] = []
        self.prefix: tuple = ()
        self.addrs: dict = {}  # full address -> exclusivity tags of the sites recorded there
        self.subcalls: dict = {}
        self.frames: list = []  # (kind "switch" | "mask", id, branch, predicate Expr)
        self.cmask_addrs = set(cmask_addrs)
        self.scan_positions: tuple = ()  # address positions holding the step index of the enclosing unrolled Scans
        self.extra_args: list = []  # (ArgSpec, Expr) hidden arguments appended after the model's own
        self.n_args = n_args
        self._ids = 0

    def new_id(self) -> int:
        self._ids += 1
        return self._ids

    def frame(self, kind: str, ident: int, branch: int, pred: Expr):
        cap = self

        class _Frame:
            def __enter__(self):
                cap.frames.append((kind, ident, branch, pred))

            def __exit__(self, *exc):
                cap.frames.pop()
                return False

        return _Frame()

    def _preds(self):
        live = scored = None
        for kind, _i, _b, pred in self.frames:
            if kind == "switch":
                live = pred if live is None else E.binary("and", live, pred)
            else:
                scored = pred if scored is None else E.binary("and", scored, pred)
        return live, scored

    def record(self, addr, dist, args) -> Expr:
        full = self.prefix + addr
        excl = tuple((i, b) for kind, i, b, _p in self.frames if kind == "switch")
        for other in self.addrs.get(full, ()):
            # the same address may be visited once per branch of one Switch (ChoiceMap.switch merges them)
            if not any(i == j and b != c for i, b in excl for j, c in other):
                raise AddressReuse(full if len(full) > 1 else full[0])
        self.addrs.setdefault(full, []).append(excl)
        cargs = dist.canonical_args(args)
        dtype, shape = dist.value_type(cargs)
        idx = len(self.sites)
        v = Expr("site", (), dtype, shape, idx)
        live, scored = self._preds()
        cmask = None
        if full in self.cmask_addrs:
            k = self.n_args + len(self.extra_args)
            cmask = Expr("arg", (), I32_, (), {"index": k, "kind": "particle"})
            self.extra_args.append((ArgSpec("particle", I32_, ()), cmask))
        self.sites.append(SiteSpec(idx, full, dist, cargs, v, live, scored, excl, self.scan_positions, cmask))
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
    if isinstance(gen_fn, GenerativeFunctionClosure_()):
        # ``f(a)(b) @ addr`` / ``normal(0., 1.).or_else(...)(...)``: a closure used as a callee
        args = tuple(gen_fn.args) + tuple(args)
        gen_fn = gen_fn.gen_fn
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


def GenerativeFunctionClosure_():
    from .gfi import GenerativeFunctionClosure

    return GenerativeFunctionClosure


def inline_call(gen_fn, args):
    """Run ``gen_fn(*args)`` at the CURRENT address prefix of the capture: a nested ``@gen`` function / combinator inlines
    its sites, a distribution records one site at the prefix itself (its address is the callee's, ``()`` below it)."""
    cap = _CAPTURE.get()
    if isinstance(gen_fn, GenerativeFunctionClosure_()):
        args = tuple(gen_fn.args) + tuple(args)
        gen_fn = gen_fn.gen_fn
    if hasattr(gen_fn, "capture_inline"):
        return gen_fn.capture_inline(tuple(args))
    return cap.record((), gen_fn, tuple(args))


# ----------------------------------------------------------------- pytrees


class StackedList(list):
    """Per-element values of an unrolled Scan / Vmap: a Python list while the body is captured (``ys[0]``, ``sum(ys)``);
    as a return value its tensor leaves come back stacked along the mapped axis (after the particle axis), which is what
    the reference's vectorised traces hold."""


_STACK_DIM: contextvars.ContextVar = contextvars.ContextVar("genjax_b200_stack_dim", default=0)


def flatten(tree, is_leaf=None) -> tuple[list, Any]:
    """Minimal pytree flatten over tuple / list / dict (``is_leaf(node)`` stops the descent, as in jax.tree_util)."""
    leaves: list = []

    def go(t):
        from ..core.choice_map import ChoiceMap
        from ..core.mask import Mask

        if is_leaf is not None and is_leaf(t):
            leaves.append(t)
            return ("leaf", len(leaves) - 1)

        if isinstance(t, ChoiceMap):
            # a choice map is a pytree of its leaves (the reference's ChoiceMap is a Pytree, choice_map.py:847)
            return ("chm", [(addr, go(v)) for addr, v in t.leaves()])
        if isinstance(t, Mask):
            return ("mask", (go(t.value), go(t.flag)))
        if type(t).__name__ == "Target" and hasattr(t, "constraint") and hasattr(t, "p"):
            # Target(p, args, constraint): p is static, args and constraint are traced (sp.py:53-81)
            return ("target", (t.p, go(t.args), go(t.constraint)))
        if dataclasses.is_dataclass(t) and not isinstance(t, type):
            # user pytrees (the reference's Pytree.dataclass): fields are children, the class is static
            return ("dataclass", (type(t), [(f.name, go(getattr(t, f.name))) for f in dataclasses.fields(t)]))
        if isinstance(t, tuple):
            return ("tuple", [go(x) for x in t])
        if isinstance(t, StackedList):
            return ("stack", [go(x) for x in t])
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
    if kind == "mask":
        from ..core.mask import Mask

        v, f = payload
        f = unflatten(f, leaves)
        if hasattr(f, "dtype") and hasattr(f, "to") and not isinstance(f, Expr):
            import torch

            f = f.to(torch.bool)
        return Mask(unflatten(v, leaves), f)
    if kind == "target":
        from ..inference.sp import Target

        p, a, c = payload
        return Target(p, unflatten(a, leaves), unflatten(c, leaves))
    if kind == "dataclass":
        cls, fields = payload
        return cls(**{name: unflatten(sub, leaves) for name, sub in fields})
    if kind == "stack":
        items = [unflatten(x, leaves) for x in payload]
        if items and all(hasattr(v, "dim") and hasattr(v, "device") for v in items):
            import torch

            return torch.stack(items, dim=min(_STACK_DIM.get(), items[0].dim()))
        return StackedList(items)
    if kind == "tuple":
        return tuple(unflatten(x, leaves) for x in payload)
    if kind == "list":
        return [unflatten(x, leaves) for x in payload]
    if kind == "dict":
        return {k: unflatten(v, leaves) for k, v in payload}
    if kind == "none":
        return None
    return leaves[payload]


def capture(source: Callable, name: str, arg_specs: list, arg_tree, cmask_addrs=()) -> ModelIR:
    """Run ``source`` on symbolic args built from ``arg_specs``.  ``cmask_addrs``: addresses whose constraint carries a
    per-particle validity flag (a hidden int32 argument per such site is appended to the model's own)."""
    arg_specs = list(arg_specs)
    arg_exprs = [
        Expr("arg", (), spec.dtype, spec.shape, {"index": i, "kind": spec.kind}) for i, spec in enumerate(arg_specs)
    ]
    sym_args = unflatten(arg_tree, arg_exprs)
    cap = _Capture(cmask_addrs, len(arg_specs))
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
    n_user = len(arg_specs)
    for spec, e in cap.extra_args:
        arg_specs.append(spec)
        arg_exprs.append(e)
    missing = set(cmask_addrs) - {s.addr for s in cap.sites}
    if missing:
        raise KeyError(f"masked constraint at {sorted(missing)}: no such site")
    ir = ModelIR(name, list(arg_specs), arg_exprs, cap.sites, ret_leaves, ret_tree, width=(widths.pop() if widths else 0),
                 subcalls=cap.subcalls, n_user_args=n_user)
    # validity flags of the sites under a Switch / MaskCombinator leave the kernel as extra (hidden) return leaves
    # (``flag_of`` indexes the combined list [return leaves..., hidden flag leaves...]; a flag the model itself returns --
    # the ``Mask(retval, check)`` of a MaskCombinator -- is not written twice)
    returned = {r._id: k for k, r in enumerate(ret_leaves) if isinstance(r, Expr) and r.ndim == 0}
    seen: dict = {}
    for s in cap.sites:
        f = s.flag()
        if f is None:
            continue
        key = (s.live._id if s.live is not None else 0, s.scored._id if s.scored is not None else 0)
        if key not in seen:
            if f._id in returned:
                seen[key] = returned[f._id]
            else:
                seen[key] = len(ret_leaves) + len(ir.flag_leaves)
                ir.flag_leaves.append(f)
        ir.flag_of[s.index] = seen[key]
    if len(ir.sites) > 32:
        raise NotImplementedError("more than 32 random-choice sites in one static model (GJB_MAX_SITES)")
    if len(ir.args) > 16:
        raise NotImplementedError("more than 16 model arguments (hidden constraint masks included)")
    if len(ir.ret_leaves) + len(ir.flag_leaves) > 16:
        raise NotImplementedError("more than 16 return leaves (validity flags of Switch / Mask sites included)")
    return ir


def ir_fingerprint(ir: ModelIR) -> str:
    """Structural hash (used to key compiled kernels)."""
    roots = []
    for s in ir.sites:
        roots.extend(s.args)
        roots.extend(p for p in (s.live, s.scored, s.cmask) if p is not None)
    roots.extend(r for r in ir.ret_leaves if isinstance(r, Expr))
    roots.extend(ir.flag_leaves)
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
    if ir.dynamic:
        desc["dynamic"] = [(ids[s.live._id] if s.live is not None else None, ids[s.scored._id] if s.scored is not None else None,
                            s.cmask.attr["index"] if s.cmask is not None else None) for s in ir.sites]
        desc["flags"] = [ids[f._id] for f in ir.flag_leaves]
    return hashlib.sha256(json.dumps(desc, default=str).encode()).hexdigest()[:16]

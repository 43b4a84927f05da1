"""``@gen``: the static modeling language on fused CUDA kernels.

API mirror of src/genjax/_src/generative_functions/static.py
(``StaticGenerativeFunction:726``, ``gen:1044``, ``StaticTrace:81``,
``simulate:787``, ``generate:795-810``, ``assess:983-989``, ``edit:965-981``,
``project``).  Where the reference re-interprets a jaxpr with a handler per GFI
method under ``jax.vmap``, this class captures the body once per argument
signature (gen/capture.py), compiles ONE fused kernel (gen/codegen.py ->
nvcc -> genjax_b200/_lib/model_<digest>.so) and drives it through the C-ABI
with per-site flags.  Batching: pass a ``KeyBatch`` (``split(key, n)``) as the
key, or use ``genjax_b200.vmap`` with ``in_axes`` exactly like ``jax.vmap``.

There is no CPU path: calling any GFI method without a CUDA device raises.
"""

from __future__ import annotations

import ctypes as C
import json
from typing import Callable

import numpy as np
import torch

from ..core.choice_map import ChoiceMap, Selection
from ..core.key import KeyBatch, PRNGKey, lanes_of
from ..core.mask import Mask
from ..runtime import build, cabi
from . import capture as cap
from . import codegen
from .capture import AddressReuse, ArgSpec, MissingAddress, ModelIR
from . import expr as E
from .expr import Expr, F32, I32
from .gfi import (
    Diff,
    DiffAnnotate,
    EditRequest,
    EmptyRequest,
    GenerativeFunction,
    NoChange,
    NotSupportedEditRequest,
    Regenerate,
    StaticRequest,
    Trace,
    UnknownChange,
    Update,
)

__all__ = ["gen", "StaticGenerativeFunction", "StaticTrace", "SubTrace", "vmap", "AddressReuse", "MissingAddress"]


# --------------------------------------------------------------- batch specs


class Batched:
    """Marks a value that carries a leading particle axis (``in_axes=0``)."""

    __slots__ = ("value",)

    def __init__(self, value):
        self.value = value


def _is_scalar_number(v) -> bool:
    return isinstance(v, (bool, int, float, np.integer, np.floating))


def _mark(tree, axes):
    """Wrap leaves with ``Batched`` where the in_axes tree says 0."""
    if isinstance(tree, ChoiceMap):
        def wrap(v):
            if isinstance(v, Mask):  # a masked constraint: value and flag both carry the particle axis
                return Mask(wrap(v.value), v.flag)
            if isinstance(v, torch.Tensor) and v.ndim == 0:
                return v  # a 0-d leaf has no particle axis: shared by all particles
            return Batched(v)

        return tree.map_leaves(wrap) if axes == 0 else tree
    if isinstance(axes, (tuple, list)) and isinstance(tree, (tuple, list)):
        if len(axes) != len(tree):
            raise ValueError("in_axes structure does not match the argument structure")
        return type(tree)(_mark(t, a) for t, a in zip(tree, axes))
    if isinstance(tree, (tuple, list)):
        return type(tree)(_mark(t, axes) for t in tree)
    if isinstance(tree, dict):
        return {k: _mark(v, axes[k] if isinstance(axes, dict) else axes) for k, v in tree.items()}
    if axes is None or tree is None:
        return tree
    if axes == 0:
        if isinstance(tree, (KeyBatch, PRNGKey, Trace)):
            return tree
        return Batched(tree)
    raise NotImplementedError("only in_axes of 0 / None are supported (particles live on the leading axis)")


def vmap(fn: Callable | None = None, in_axes=0):
    """``jax.vmap`` for GFI calls: ``vmap(model.importance, in_axes=(0, None, (0,)))(keys, chm, (x_prev,))``.

    The batch is not unrolled: the call runs ONE fused kernel over all lanes
    of the ``KeyBatch`` / all rows of the ``in_axes=0`` arguments.

    ``vmap(in_axes=...)`` without a function is the reference's combinator decorator ``@genjax.vmap(in_axes=...)``
    (vmap.py:384): applied to an ``@gen`` function it returns a ``Vmap`` generative function."""
    if fn is None or isinstance(fn, StaticGenerativeFunction):
        from .vmap_combinator import Vmap

        if fn is not None:
            return Vmap(fn, in_axes)
        return lambda f: Vmap(f, in_axes) if isinstance(f, StaticGenerativeFunction) else vmap(f, in_axes)

    def wrapped(*args):
        axes = in_axes if isinstance(in_axes, (tuple, list)) else (in_axes,) * len(args)
        if len(axes) != len(args):
            raise ValueError("vmap in_axes must match the number of positional arguments")
        return fn(*[_mark(a, ax) for a, ax in zip(args, axes)])

    return wrapped


# ----------------------------------------------------------------- compiled


class CompiledModel:
    def __init__(self, ir: ModelIR, lib, path):
        self.ir = ir
        self.lib = lib
        self.path = path
        self.info = json.loads(lib.gjb_model_info().decode())


_COMPILED: dict[str, CompiledModel] = {}


def compile_ir(ir: ModelIR, pf_obs: tuple | None = None, chain=None) -> CompiledModel:
    ir.digest = cap.ir_fingerprint(ir)
    source = codegen.generate(ir, pf_obs, chain)
    key = build.model_digest(source)
    cm = _COMPILED.get(key)
    if cm is None:
        path = build.build_model(source)
        cm = CompiledModel(ir, cabi.load_model_library(path), path)
        _COMPILED[key] = cm
    return cm


def cap_norm(addr) -> tuple:
    from ..core.choice_map import _norm_addr

    return _norm_addr(addr)


def _dev_tensor(v, device, want_int=None) -> torch.Tensor:
    """Value -> contiguous float32 / int32 CUDA tensor."""
    if isinstance(v, torch.Tensor):
        t = v
    else:
        t = torch.as_tensor(np.asarray(v))
    if t.dtype in (torch.float32,):
        pass
    elif t.dtype in (torch.float64, torch.float16, torch.bfloat16):
        t = t.to(torch.float32)
    elif t.dtype in (torch.int32,):
        pass
    elif t.dtype in (torch.int64, torch.int16, torch.int8, torch.uint8, torch.bool):
        t = t.to(torch.int32)
    else:
        raise TypeError(f"unsupported dtype {t.dtype}")
    if want_int is True and t.dtype != torch.int32:
        t = t.to(torch.int32)
    if want_int is False and t.dtype != torch.float32:
        t = t.to(torch.float32)
    if t.device != device:
        t = t.to(device)
    return t.contiguous()


def _dtype_of(t: torch.Tensor) -> str:
    return F32 if t.dtype == torch.float32 else I32


class _BoundArgs:
    """Model args classified into the capture signature + launch payload."""

    def __init__(self, args, device):
        leaves, self.tree = cap.flatten(args)
        self.specs: list[ArgSpec] = []
        self.payload: list = []  # python scalar | tensor
        self.n_batched = None
        for leaf in leaves:
            batched = isinstance(leaf, Batched)
            v = leaf.value if batched else leaf
            if isinstance(v, Diff):
                v = v.primal
            if _is_scalar_number(v) and not batched:
                dt = I32 if isinstance(v, (bool, int, np.integer)) else F32
                self.specs.append(ArgSpec("scalar", dt, ()))
                self.payload.append(float(v))
                continue
            t = _dev_tensor(v, device)
            if batched:
                if t.ndim == 0:
                    raise ValueError("in_axes=0 argument has no leading axis")
                n = t.shape[0]
                if self.n_batched is not None and self.n_batched != n:
                    raise ValueError("batched arguments disagree on the particle-axis size")
                self.n_batched = n
                if t.ndim > 2:
                    raise NotImplementedError("per-particle arguments must be [n] or [n, d]")
                self.specs.append(ArgSpec("particle", _dtype_of(t), tuple(t.shape[1:])))
            else:
                if t.ndim > 2:
                    raise NotImplementedError("shared arguments must have at most 2 axes")
                self.specs.append(ArgSpec("shared", _dtype_of(t), tuple(t.shape)))
            self.payload.append(t)

    def signature(self):
        return (repr(self.tree), tuple(s.key() for s in self.specs))


# -------------------------------------------------------------------- trace


class StaticTrace(Trace):
    """Trace of a static model over ``n`` particles (struct-of-arrays: every
    leaf has a leading [n] axis, as under ``jax.vmap`` in the reference).
    ``batched=False`` traces expose 0-d / event-shaped views."""

    def __init__(self, gen_fn, cm: CompiledModel, bound: _BoundArgs, args, n, batched, values, score, ret_leaves,
                 bcast, flags=None):
        self.flags = flags or {}  # site index -> int32 [n]: where a site under a Switch branch / Mask is valid
        self.gen_fn = gen_fn
        self.cm = cm
        self.bound = bound
        self.args = args
        self.n = n
        self.batched = batched
        self.values = values  # site index -> tensor [n(, d)] or broadcast value tensor
        self.bcast = bcast  # site index -> bool (value shared by all particles)
        self.score = score  # [n]
        self.ret_leaves = ret_leaves
        self._site_score_cache = None

    def _view(self, t, is_bcast=False):
        if t is None or not isinstance(t, torch.Tensor):
            return t
        if is_bcast:
            return t
        return t if self.batched else t[0]

    def get_gen_fn(self):
        return self.gen_fn

    def get_args(self):
        return _unmark(self.args)

    def get_score(self):
        return self._view(self.score)

    def get_retval(self):
        leaves = [self._view(r) for r in self.ret_leaves]
        tok = cap._STACK_DIM.set(1 if self.batched else 0)  # unrolled Scan / Vmap outputs stack after the particle axis
        try:
            return cap.unflatten(self.cm.ir.ret_tree, leaves)
        finally:
            cap._STACK_DIM.reset(tok)

    def _site_value(self, s):
        v = self.values[s.index]
        v = self._view(v, self.bcast[s.index])
        if self.bcast[s.index] and self.batched and isinstance(v, torch.Tensor):
            # a constraint shared by all particles reads back with the particle axis, like every
            # leaf of a vmapped trace in the reference (zero-copy expand)
            ev = tuple(s.value.shape)
            v = v.reshape(ev).expand((self.n,) + ev)
        if getattr(s.dist, "bool_valued", False) and isinstance(v, torch.Tensor):
            v = v.to(torch.bool)
        f = self.flags.get(s.index)
        if f is not None:
            # ``inner.get_choices().mask(check)`` (mask.py:83) / ``ChoiceMap.switch(idx, sub_chms)`` (switch.py:75-78)
            v = Mask(v, self._view(f).to(torch.bool))
        return v

    def get_choices(self) -> ChoiceMap:
        return _choices_from_sites(self, self.cm.ir.sites)

    @property
    def inner(self):
        """``MaskTrace.inner`` (mask.py:55): the trace of the masked callee -- the same choices, scored without the mask."""
        mc = getattr(self.gen_fn, "_mask_of", None)
        if mc is None:
            raise AttributeError("inner")
        if getattr(self, "_inner", None) is None:
            self._inner = mc._inner_trace(self)
        return self._inner

    def get_subtrace(self, *addr):
        """``tr.get_subtrace("x")`` / ``tr.get_subtrace("f", "x")`` (generative_function.py:141-175,
        static.py:118-127).  The sites are fused into one kernel, so a sub-trace is a VIEW: the value of the site
        (or the choices under a nested call's prefix) plus its log-density from ``site_scores()``."""
        a = cap_norm(addr)
        sites = [s for s in self.cm.ir.sites if s.addr[: len(a)] == a]
        if not a or not sites:
            from ..core.choice_map import ChoiceMapNoValueAtAddress

            raise ChoiceMapNoValueAtAddress(a[0] if len(a) == 1 else a)
        return SubTrace(self, a, sites)

    def site_scores(self) -> dict:
        """Per-site log-densities ``{addr: [n]}`` (one weight-only launch per site, computed once per trace)."""
        if self._site_score_cache is None:
            self._site_score_cache = self.gen_fn._site_scores(self)
        return self._site_score_cache

    def take(self, idx) -> "StaticTrace":
        """``tree_map(lambda v: v[idx])`` over the particle axis (smc.py:90-91)."""
        if not self.batched:
            raise ValueError("take() needs a batched trace")
        if isinstance(idx, int):
            sel = torch.tensor([idx], device=self.score.device)
            out_batched = False
        else:
            sel = idx.to(self.score.device).long()
            out_batched = True

        sel32 = sel.to(torch.int32).contiguous()

        def g(t, b):
            if b or not isinstance(t, torch.Tensor):
                return t
            if t.element_size() == 4 and t.shape[0] == self.n:  # float32 / int32 leaves: the resampling gather kernel
                from ..runtime import smc_ops

                return smc_ops.gather_rows(t.contiguous(), sel32)
            return t.index_select(0, sel)

        values = {k: g(v, self.bcast[k]) for k, v in self.values.items()}
        rets = [g(r, False) for r in self.ret_leaves]
        args = _take_args(self.args, sel)
        return StaticTrace(self.gen_fn, self.cm, None, args, int(sel.numel()), out_batched, values,
                           g(self.score, False), rets, dict(self.bcast), {k: g(f, False) for k, f in self.flags.items()})


def _choices_from_sites(trace: StaticTrace, sites) -> ChoiceMap:
    """Choice map of ``sites``.  An address visited in several branches of one Switch holds the OR of the branches'
    masked values (``ChoiceMap.switch``, choice_map.py ``switch``; ``Mask.__or__``, functional_types.py:309)."""
    by_addr: dict = {}
    stacks: dict = {}
    for s in sites:
        by_addr.setdefault(s.addr, []).append(trace._site_value(s))
        if s.stack:
            stacks[s.addr] = s.stack
    merged = {addr: (vals[0] if len(vals) == 1 else Mask.or_n(*vals)) for addr, vals in by_addr.items()}
    chm = ChoiceMap.empty()
    # the steps of an unrolled Scan read back stacked along a time axis (after the particle axis), under the address
    # without the step index -- what ScanTrace.get_choices shows (scan.py:81-84: the inner trace is vectorised)
    groups: dict = {}
    for addr, v in merged.items():
        st = stacks.get(addr)
        if st and len(st) == 1 and isinstance(v, (torch.Tensor, Mask)):
            groups.setdefault(addr[: st[0]] + addr[st[0] + 1:], []).append((addr[st[0]], v, addr))
        else:
            chm = chm | ChoiceMap.entry(v, *addr)
    dim = 1 if trace.batched else 0
    totals: dict = {}
    for s in trace.cm.ir.sites:
        if s.stack and len(s.stack) == 1:
            k = s.addr[: s.stack[0]] + s.addr[s.stack[0] + 1:]
            totals.setdefault(k, set()).add(s.addr[s.stack[0]])
    for addr, steps in groups.items():
        steps.sort(key=lambda p: p[0])
        vals = [v for _, v, _ in steps]
        if len({i for i, _, _ in steps}) < len(totals.get(addr, ())):
            # only some elements (the discard of an update at one index): they stay under their own indices
            for _, v, full in steps:
                chm = chm | ChoiceMap.entry(v, *full)
        elif all(isinstance(v, Mask) for v in vals):  # masked elements: Mask(stacked values, stacked flags)
            chm = chm | ChoiceMap.entry(Mask(torch.stack([v.value for v in vals], dim=dim),
                                             torch.stack([v.flag for v in vals], dim=dim)), *addr)
        elif any(isinstance(v, Mask) for v in vals):  # (some elements masked, some not: kept apart under their indices)
            for _, v, full in steps:
                chm = chm | ChoiceMap.entry(v, *full)
        else:
            chm = chm | ChoiceMap.entry(torch.stack(vals, dim=dim), *addr)
    return chm


class ZeroTrace(Trace):
    """``gen_fn.get_zero_trace(*args)``: addresses, shapes and dtypes of a trace, all values zero."""

    def __init__(self, gen_fn, args, ir: ModelIR):
        self.gen_fn, self.args, self.ir = gen_fn, args, ir

    @staticmethod
    def _zero(e):
        if not isinstance(e, Expr):
            return e
        dt = torch.int32 if e.dtype == I32 else torch.float32
        return torch.zeros(tuple(e.shape), dtype=dt) if e.shape else (0 if e.dtype == I32 else 0.0)

    def get_gen_fn(self):
        return self.gen_fn

    def get_args(self):
        return self.args

    def get_score(self):
        return 0.0

    def get_retval(self):
        return cap.unflatten(self.ir.ret_tree, [self._zero(r) for r in self.ir.ret_leaves])

    def get_choices(self) -> ChoiceMap:
        chm = ChoiceMap.empty()
        for s in self.ir.sites:
            chm = chm | ChoiceMap.entry(self._zero(s.value), *s.addr)
        return chm


class SubTrace(Trace):
    """View of one site (``DistributionTrace``, distribution.py:60-87) or of the sites under a nested call's
    address prefix (the inlined callee's ``StaticTrace``) inside a fused ``StaticTrace``."""

    def __init__(self, parent: StaticTrace, prefix: tuple, sites: list):
        self.parent = parent
        self.prefix = prefix
        self.sites = sites
        self.is_site = len(sites) == 1 and sites[0].addr == prefix

    def get_gen_fn(self):
        if self.is_site:
            return self.sites[0].dist
        raise NotImplementedError("the callee of an inlined nested call is not kept in the fused trace")

    def get_args(self):
        raise NotImplementedError("per-site arguments are fused away (they live in registers of the model kernel)")

    def get_score(self):
        scores = self.parent.site_scores()
        tot = None
        for s in self.sites:
            tot = scores[s.addr] if tot is None else tot + scores[s.addr]
        return tot if self.parent.batched else tot[0]

    def get_choices(self) -> ChoiceMap:
        return self.parent.get_choices().get_submap(*self.prefix)

    get_sample = get_choices

    def get_retval(self):
        if self.is_site:
            return self.parent._site_value(self.sites[0])
        raise NotImplementedError("the return value of an inlined nested call is not kept in the fused trace")

    get_value = get_retval

    def get_subtrace(self, *addr):
        return self.parent.get_subtrace(*self.prefix, *addr)

    def project(self, key, selection: Selection):
        return self.parent.project(key, selection.extend(*self.prefix))


def _unmark(tree):
    if isinstance(tree, Batched):
        return tree.value
    if isinstance(tree, tuple):
        return tuple(_unmark(t) for t in tree)
    if isinstance(tree, list):
        return [_unmark(t) for t in tree]
    if isinstance(tree, dict):
        return {k: _unmark(v) for k, v in tree.items()}
    return tree


def _take_args(tree, sel):
    if isinstance(tree, Batched):
        return Batched(tree.value.index_select(0, sel.to(tree.value.device)))
    if isinstance(tree, tuple):
        return tuple(_take_args(t, sel) for t in tree)
    if isinstance(tree, list):
        return [_take_args(t, sel) for t in tree]
    if isinstance(tree, dict):
        return {k: _take_args(v, sel) for k, v in tree.items()}
    return tree


# ----------------------------------------------------- the generative function


class StaticGenerativeFunction(GenerativeFunction):
    def __init__(self, source: Callable, partial_args: tuple = ()):
        self.source = source
        # functools.wraps-style metadata (static.py:1044-1062; test_static_gen_fn.py:38-79)
        self.__name__ = getattr(source, "__name__", "model")
        self.__doc__ = getattr(source, "__doc__", None)
        self.__module__ = getattr(source, "__module__", self.__module__)
        self.__qualname__ = getattr(source, "__qualname__", self.__name__)
        self.__annotations__ = dict(getattr(source, "__annotations__", {}))
        self.__wrapped__ = source
        self.partial_args = tuple(partial_args)  # arguments fixed by partial_apply / method binding
        self._cache: dict = {}
        self._kwarged = None

    def __repr__(self):
        return f"StaticGenerativeFunction({self.__name__})"

    def __get__(self, instance, owner):
        # @gen on methods (static.py:757-763): bind `self` as the first argument
        if instance is None:
            return self
        bound = StaticGenerativeFunction(lambda *a, **kw: self.source(instance, *a, **kw), self.partial_args + (instance,))
        bound.__name__ = self.__name__
        return bound

    def partial_apply(self, *bound):
        """Same model with its first arguments fixed (test_static_gen_fn.py:1116-1163): addresses unchanged, the
        fixed values stay readable as ``partial_args``."""
        src = self.source
        out = StaticGenerativeFunction(lambda *rest, **kw: src(*bound, *rest, **kw), self.partial_args + tuple(bound))
        out.__name__ = f"{self.__name__}_partial"
        return out

    def handle_kwargs(self) -> "StaticGenerativeFunction":
        """The same model taking ``(args, kwargs)`` as its two arguments (static.py ``handle_kwargs``;
        generative_function.py:1563-1565): what a closure with keyword arguments runs."""
        if self._kwarged is None:
            src = self.source
            out = StaticGenerativeFunction(lambda args, kwargs: src(*args, **kwargs), self.partial_args)
            out.__name__ = f"{self.__name__}_kwargs"
            out._kwarged = out
            self._kwarged = out
        return self._kwarged

    def dimap(self, *, pre: Callable = lambda *args: args, post: Callable = lambda _args, _xformed, retval: retval):
        """``gen_fn.dimap(pre=, post=)`` (generative_function.py:1339-1389; combinators/dimap.py): ``pre`` maps the
        arguments (and must return a tuple), ``post(args, transformed_args, retval)`` maps the return value.  Both
        are traced into the fused kernel with the body; the choices keep their addresses."""
        src = self.source

        def wrapped(*args):
            xformed = tuple(pre(*args))
            return post(args, xformed, src(*xformed))

        out = StaticGenerativeFunction(wrapped, self.partial_args)
        out.__name__ = f"{self.__name__}_dimap"
        return out

    def map(self, f: Callable):
        """``gen_fn.map(f)``: post-process the return value (generative_function.py:1391-1427)."""
        return self.dimap(post=lambda _args, _xformed, retval: f(retval))

    def contramap(self, f: Callable):
        """``gen_fn.contramap(f)``: pre-process the arguments; ``f`` returns a tuple (generative_function.py:1429-1467)."""
        return self.dimap(pre=f)

    def inline(self, *args):
        """``callee.inline(*args)`` inside an ``@gen`` body: the callee's choices are recorded at the CALLER's
        address level (static.py ``inline``; test_static_gen_fn.py:949-1087)."""
        if cap.current_capture() is None:
            raise RuntimeError("`gen_fn.inline(*args)` can only be used inside a @gen function body")
        return self.source(*args)

    def get_zero_trace(self, *args):
        """A trace-shaped object with zeros everywhere (generative_function.py ``get_zero_trace``): the SHAPE of what
        the model visits for these arguments.  Host-only, nothing is launched."""
        bound = _BoundArgs(args, torch.device("cpu"))
        return ZeroTrace(self, args, cap.capture(self.source, self.__name__, bound.specs, bound.tree))

    # -- capture ---------------------------------------------------------
    def capture_inline(self, args):
        """Nested call inside another @gen body: inline the sites."""
        if isinstance(args, tuple) and len(args) == 2 and isinstance(args[1], dict) and isinstance(args[0], tuple):
            return self.source(*args[0], **args[1])
        return self.source(*args)

    def compiled_for(self, bound: _BoundArgs, cmask: frozenset = frozenset()) -> CompiledModel:
        """``cmask``: addresses whose constraint is a ``Mask`` with a per-particle flag -- a variant of the model with
        one hidden flag argument per such site (gen/capture.py ``SiteSpec.cmask``)."""
        sig = bound.signature() + ((tuple(sorted(cmask, key=repr)),) if cmask else ())
        cm = self._cache.get(sig)
        if cm is None:
            ir = cap.capture(self.source, self.__name__, bound.specs, bound.tree, cmask_addrs=cmask)
            cm = compile_ir(ir)
            self._cache[sig] = cm
        return cm

    def get_site_addresses(self, args) -> list[tuple]:
        """Addresses the body visits for these arguments, in program order -- the shape information the
        reference takes from ``get_zero_trace(*args).get_choices()`` (choice_map.py:1372).  Host-only: the body is
        captured symbolically, nothing is compiled or launched."""
        bound = _BoundArgs(args if isinstance(args, tuple) else tuple(args), torch.device("cpu"))
        ir = cap.capture(self.source, self.__name__, bound.specs, bound.tree)
        return [s.addr for s in ir.sites]

    def prebuild(self, specs: list, tree=None, pf_obs: tuple | None = None) -> CompiledModel:
        """Capture + compile for an explicit argument signature (no GPU needed):
        ``specs`` is a list of ``ArgSpec``; used by ``__graft_entry__.build()``.
        ``pf_obs`` (addresses observed at every filter step) selects the
        variant whose persistent filter kernel has its site flags baked in."""
        if tree is None:
            tree = ("tuple", [("leaf", i) for i in range(len(specs))])
        obs_key = None if pf_obs is None else tuple(sorted(cap_norm(a) for a in pf_obs))
        sig = (repr(tree), tuple(s.key() for s in specs), obs_key)
        cm = self._cache.get(sig)
        if cm is None:
            ir = cap.capture(self.source, self.__name__, list(specs), tree)
            idx = None if obs_key is None else tuple(ir.site_index(a) for a in obs_key)
            cm = compile_ir(ir, idx)
            self._cache[sig] = cm
        return cm

    # -- engine ------------------------------------------------------------
    def _run(self, key, args, constraints: ChoiceMap | None, *, sample_addrs=None, prev: StaticTrace | None = None,
             weight_mode: str = "generate", weight_in=None, score_in=None, n=None, batched=None, want_score=True,
             gather=None, wmax=None, weight_sites=None, revive=False):
        """One fused launch.  ``weight_mode``:
        "generate": weight = sum logpdf over constrained sites;
        "delta":    weight = new score - prev score (update / regenerate);
        "none":     no weight."""
        device = cabi.require_cuda()
        args = args if isinstance(args, tuple) else tuple(args)
        bound = _BoundArgs(args, device)
        cm = self.compiled_for(bound)
        ir = cm.ir
        if key is not None:
            words, lane0, n_key = lanes_of(key)
            key_batched = isinstance(key, KeyBatch)
        else:
            words, lane0, n_key, key_batched = (0, 0), 0, None, False

        constraints = constraints if constraints is not None else ChoiceMap.empty()
        # resolve the batch size
        sizes = [s for s in (n_key if key_batched else None, bound.n_batched, n, prev.n if prev is not None and prev.batched else None) if s is not None]
        cvals = {}
        cmasks = {}  # site index -> int32 [n] flag words of a Mask-ed constraint (see SiteSpec.cmask)
        mask_discard = {}  # site index -> flag of an update's Mask-ed constraint (the discard is masked by it)
        for s in ir.sites:
            sub = constraints.get_submap(*s.addr)
            if sub.has_value():
                v = sub.get_value()
                if isinstance(v, Batched) and isinstance(v.value, Mask):
                    v = Mask(Batched(v.value.value), v.value.flag)
                if isinstance(v, Mask):
                    # distribution.py:129-142 (generate) / :190-226 (update): where the flag holds the site is
                    # constrained, elsewhere it is drawn (generate) or keeps its value (update)
                    flag, v = v.primal_flag(), v.value
                    if key is None and prev is None:
                        pass  # assess scores the wrapped value whatever the flag says (distribution.py:404-417)
                    elif isinstance(flag, bool) or (isinstance(flag, torch.Tensor) and flag.ndim == 0):
                        if not bool(flag):
                            continue
                    else:
                        f = _dev_tensor(flag, device, want_int=True)
                        sizes.append(f.shape[0])
                        if prev is not None and not (sample_addrs is not None and s.addr in sample_addrs):
                            raw = v.value if isinstance(v, Batched) else v
                            new_v = _dev_tensor(raw, device, want_int=(s.value.dtype == I32))
                            old_v = prev.values[s.index]
                            fb = f.to(torch.bool).reshape((-1,) + (1,) * (old_v.ndim - 1))
                            v = Batched(torch.where(fb, new_v.expand_as(old_v) if new_v.ndim <= old_v.ndim else new_v, old_v))
                            mask_discard[s.index] = f.to(torch.bool)
                        else:
                            cmasks[s.index] = torch.where(f != 0, 1, 2).to(torch.int32)
                            if not isinstance(v, Batched):
                                raw = _dev_tensor(v, device, want_int=(s.value.dtype == I32))
                                if tuple(raw.shape) != (f.shape[0],) + tuple(s.value.shape):
                                    raw = raw.expand((f.shape[0],) + tuple(raw.shape)).contiguous()
                                v = Batched(raw)
                if isinstance(v, Batched):
                    t = _dev_tensor(v.value, device, want_int=(s.value.dtype == I32))
                    sizes.append(t.shape[0])
                    cvals[s.index] = (t, False)
                else:
                    cvals[s.index] = (v, True)
        if revive and prev is not None:
            # Switch.edit with a changed index (switch.py:226-246): a site whose branch was not the selected one has
            # no value to keep -- it is drawn afresh where it comes alive.  (A site under BOTH a Switch and a Mask is
            # also redrawn where only its mask was off: the two flags leave the kernel as one.)
            for s in ir.sites:
                if s.live is not None and s.index not in cvals and s.index in prev.flags \
                        and not (sample_addrs is not None and s.addr in sample_addrs):
                    if key is None:
                        raise ValueError("an update that may change a Switch index needs a key")
                    cmasks[s.index] = (prev.flags[s.index] != 0).to(torch.int32)
                    cvals[s.index] = (prev.values[s.index], bool(prev.bcast[s.index]))
        if cmasks:
            if gather is not None:
                raise NotImplementedError("Mask-ed constraints together with an ancestor gather")
            cm = self.compiled_for(bound, frozenset(ir.sites[j].addr for j in cmasks))
            ir = cm.ir
        if sizes:
            n_run = sizes[0]
            if any(x != n_run for x in sizes):
                raise ValueError(f"inconsistent particle-axis sizes {sizes}")
            is_batched = True
        else:
            n_run, is_batched = 1, False
        if batched is not None:
            is_batched = batched

        A = cabi.ModelArgs()
        A.n = n_run
        A.idx_offset = lane0
        A.key0, A.key1 = words
        keep = []
        for i, (spec, pl) in enumerate(zip(bound.specs, bound.payload)):
            if spec.kind == "scalar":
                A.scalars[i] = pl
            else:
                if spec.kind == "particle" and pl.shape[0] != n_run and gather is None:
                    raise ValueError("batched argument size does not match the batch")
                A.args[i] = pl.data_ptr()
                keep.append(pl)
        if gather is not None:
            A.gather = cabi.ptr(gather)

        values, bcast = {}, {}
        for s in ir.sites:
            j = s.index
            ev = tuple(s.value.shape)
            tdt = torch.int32 if s.value.dtype == I32 else torch.float32
            flags = 0
            if j in cvals:
                v, is_b = cvals[j]
                if is_b:
                    t = _dev_tensor(v, device, want_int=(s.value.dtype == I32))
                    if tuple(t.shape) == (n_run,) + ev and is_batched and n_run > 1 and not isinstance(v, (int, float, bool)):
                        is_b = False  # a full-size tensor constrains per particle
                    elif tuple(t.shape) != ev:
                        raise ValueError(f"constraint at {s.addr} has shape {tuple(t.shape)}, expected {ev}")
                else:
                    t = v
                    if tuple(t.shape) != (n_run,) + ev:
                        raise ValueError(f"batched constraint at {s.addr} has shape {tuple(t.shape)}")
                A.site_in[j] = t.data_ptr()
                flags |= cabi.SITE_BCAST if is_b else 0
                if weight_mode in ("generate", "delta") and (weight_sites is None or j in weight_sites):
                    flags |= cabi.SITE_WEIGHT
                values[j], bcast[j] = t, is_b
                keep.append(t)
                if j in cmasks:
                    # per particle the kernel keeps the supplied value or draws one: the outcome goes to a fresh buffer
                    cmt = cmasks[j].contiguous()
                    A.args[s.cmask.attr["index"]] = cmt.data_ptr()
                    out = torch.empty((n_run,) + ev, dtype=tdt, device=device)
                    A.site_out[j] = out.data_ptr()
                    values[j], bcast[j] = out, False
                    keep.append(cmt)
            elif prev is not None and not (sample_addrs is not None and s.addr in sample_addrs):
                t = prev.values[j]
                A.site_in[j] = t.data_ptr()
                flags |= cabi.SITE_BCAST if prev.bcast[j] else 0
                if weight_mode == "delta" or (weight_sites is not None and j in weight_sites):
                    flags |= cabi.SITE_WEIGHT
                values[j], bcast[j] = t, prev.bcast[j]
            else:
                if key is None:
                    raise MissingAddress(s.addr[0] if len(s.addr) == 1 else s.addr)
                flags |= cabi.SITE_SAMPLE
                if weight_mode == "delta":
                    flags |= cabi.SITE_WEIGHT
                out = torch.empty((n_run,) + ev, dtype=tdt, device=device)
                A.site_out[j] = out.data_ptr()
                values[j], bcast[j] = out, False
            A.site_flags[j] = flags

        score = torch.empty(n_run, dtype=torch.float32, device=device) if want_score else None
        if score is not None:
            A.score_out = score.data_ptr()
        weight = None
        if weight_mode != "none":
            weight = torch.empty(n_run, dtype=torch.float32, device=device)
            A.weight_out = weight.data_ptr()
        if weight_mode == "delta":
            assert prev is not None
            A.score_in = prev.score.data_ptr()
        if score_in is not None:
            A.score_in = cabi.ptr(score_in)
        if weight_in is not None:
            A.weight_in = cabi.ptr(weight_in)
        if wmax is not None:
            A.wmax = cabi.ptr(wmax)

        ret_leaves = []
        for k, r in enumerate(ir.ret_leaves):
            if isinstance(r, Expr) and r.op == "site" and not bcast[r.attr]:
                ret_leaves.append(values[r.attr])
            elif isinstance(r, Expr):
                tdt = torch.int32 if r.dtype == I32 else torch.float32
                out = torch.empty((n_run,) + tuple(r.shape), dtype=tdt, device=device)
                A.ret_out[k] = out.data_ptr()
                ret_leaves.append(out)
            else:
                ret_leaves.append(r)

        flag_bufs = []
        for m in range(len(ir.flag_leaves)):
            fb = torch.empty(n_run, dtype=torch.int32, device=device)
            A.ret_out[len(ir.ret_leaves) + m] = fb.data_ptr()
            flag_bufs.append(fb)

        cabi.check(cm.lib.gjb_model_launch(C.byref(A), cabi.stream_ptr(device)), f"gjb_model_launch({self.__name__})")
        n_ret = len(ir.ret_leaves)
        tr = StaticTrace(self, cm, bound, args, n_run, is_batched, values, score, ret_leaves, bcast,
                         {j: (ret_leaves[m] if m < n_ret else flag_bufs[m - n_ret]) for j, m in ir.flag_of.items()})
        tr._mask_discard = mask_discard
        return tr, weight

    def _w(self, tr: StaticTrace, w):
        return w if tr.batched else w[0]

    # -- GFI -------------------------------------------------------------
    def simulate(self, key, args: tuple) -> StaticTrace:
        tr, _ = self._run(key, args, None, weight_mode="none")
        return tr

    def generate(self, key, constraint: ChoiceMap, args: tuple):
        tr, w = self._run(key, args, constraint, weight_mode="generate")
        return tr, self._w(tr, w)

    def assess(self, sample: ChoiceMap, args: tuple):
        tr, _ = self._run(None, args, sample, weight_mode="none")
        return tr.get_score(), tr.get_retval()

    def project(self, key, trace: StaticTrace, selection: Selection):
        sel_chm = trace.get_choices().filter(selection)
        if sel_chm.static_is_empty():
            return torch.zeros_like(trace.get_score())
        scores = self._site_scores(trace)
        tot = torch.zeros_like(trace.score)
        for addr, sel_addr in dict.fromkeys((s.addr, s.sel_addr) for s in trace.cm.ir.sites):
            if selection(sel_addr).check():
                tot = tot + scores[addr]
        return tot if trace.batched else tot[0]

    def _site_scores(self, trace: StaticTrace) -> dict:
        """addr -> [n] logpdf: one launch per site, every value read back from the trace, only that site weighted (an
        address visited in several Switch branches sums its sites: at most one of them is alive per particle)."""
        out = {}
        for s in trace.cm.ir.sites:
            _, w = self._run(None, trace.args, None, prev=trace, weight_mode="generate", weight_sites={s.index},
                             n=trace.n, batched=trace.batched, want_score=False)
            out[s.addr] = w if s.addr not in out else out[s.addr] + w
        return out

    def edit(self, key, trace: StaticTrace, request: EditRequest, argdiffs):
        new_args = Diff.tree_primal(argdiffs) if argdiffs is not None and argdiffs != () else trace.args
        if new_args == () and trace.args != ():
            new_args = trace.args
        new_args = _carry_batch_marks(new_args, trace.args)
        # a Switch index may move with the arguments OR with an upstream choice: sites of a branch that comes alive are
        # drawn afresh (switch.py:226-246), so every edit of a model holding Switch sites runs the per-particle form
        revive = key is not None and any(s.live is not None for s in trace.cm.ir.sites)
        if isinstance(request, Update):
            tr, w = self._run(key, new_args, _rebatch_constraint(request.constraint, trace), prev=trace,
                              weight_mode="delta", n=trace.n, batched=trace.batched, revive=revive)
            discard = _choices_from_sites(trace, [s for s in trace.cm.ir.sites
                                                  if request.constraint.get_submap(*s.addr).has_value()])
            if tr._mask_discard:  # ``old_choices.mask(flag)`` (distribution.py:222-224)
                flags = {trace.cm.ir.sites[j].addr: f for j, f in tr._mask_discard.items()}
                discard = _mask_leaves(discard, flags, trace.batched)
            retdiff = Diff.unknown_change(tr.get_retval())
            return tr, self._w(tr, w), retdiff, Update(discard)
        if isinstance(request, Regenerate):
            sel = {s.addr for s in trace.cm.ir.sites if request.selection(s.sel_addr).check()}
            tr, w = self._run(key, new_args, None, prev=trace, sample_addrs=sel, weight_mode="delta", n=trace.n,
                              batched=trace.batched, revive=revive)
            discard = _choices_from_sites(trace, [s for s in trace.cm.ir.sites if s.addr in sel])
            return tr, self._w(tr, w), Diff.unknown_change(tr.get_retval()), Update(discard)
        if isinstance(request, EmptyRequest):
            return request.edit(key, trace, argdiffs)
        if isinstance(request, StaticRequest):
            return _edit_static_request(self, key, trace, request, argdiffs)
        if isinstance(request, DiffAnnotate):
            return request.edit(key, trace, argdiffs)
        if hasattr(request, "edit") and type(request).edit is not EditRequest.edit:
            return request.edit(key, trace, argdiffs)
        raise NotSupportedEditRequest(request)


def _carry_batch_marks(new_args, old_args):
    """Primal args from argdiffs lose their ``Batched`` marks; restore them by position."""
    if isinstance(old_args, Batched) and not isinstance(new_args, Batched):
        return Batched(new_args)
    if isinstance(old_args, tuple) and isinstance(new_args, tuple) and len(old_args) == len(new_args):
        return tuple(_carry_batch_marks(n, o) for n, o in zip(new_args, old_args))
    return new_args


def _mask_leaves(chm: ChoiceMap, flags: dict, batched: bool) -> ChoiceMap:
    """Wrap the leaves at the addresses of ``flags`` in ``Mask(value, flag)``."""
    out = ChoiceMap.empty()
    for addr, v in chm.leaves():
        f = flags.get(addr)
        if f is not None:
            v = Mask.build(v, f if batched else f[0])
        out = out | ChoiceMap.entry(v, *addr)
    return out


def _rebatch(trace: StaticTrace) -> ChoiceMap:
    """Choice map of a trace with batched leaves marked for re-launch."""
    chm = ChoiceMap.empty()
    for s in trace.cm.ir.sites:
        v = trace.values[s.index]
        chm = chm | ChoiceMap.entry(v if trace.bcast[s.index] else Batched(v), *s.addr)
    return chm


def _rebatch_constraint(chm: ChoiceMap, trace: StaticTrace) -> ChoiceMap:
    """Leaves of an update constraint that already carry the particle axis are batched."""
    if not trace.batched:
        return chm

    def mark(v):
        if isinstance(v, Batched):
            return v
        if isinstance(v, Mask):
            return Mask(mark(v.value), v.flag)
        if isinstance(v, torch.Tensor) and v.ndim >= 1 and v.shape[0] == trace.n:
            return Batched(v)
        return v

    return chm.map_leaves(mark)


def _depends(exprs, moved_sites: set, changed_args: set) -> bool:
    """Does any of ``exprs`` read a site whose value moves or a model argument that changed?"""
    roots = [e for e in exprs if isinstance(e, Expr)]
    for node in E.topo(roots):
        if node.op == "site" and node.attr in moved_sites:
            return True
        if node.op == "arg" and node.attr["index"] in changed_args:
            return True
    return False


def _changed_arg_leaves(argdiffs) -> set:
    """Indices of the flattened model-argument leaves whose ``Diff`` tangent is not NoChange."""
    if argdiffs is None or argdiffs == ():
        return set()
    leaves, _ = cap.flatten(Diff._map(argdiffs, lambda v: v if isinstance(v, Diff) else Diff(v, UnknownChange)))
    return {i for i, v in enumerate(leaves) if isinstance(v, Diff) and v.tangent is not NoChange}


def _callee_diffs(ir: ModelIR, addr: tuple, moved: set, changed_args: set):
    """(argdiffs, retdiff) of the callee at ``addr`` -- a distribution site or a nested @gen call -- as trees of
    ``Diff(None, tangent)``: the static change propagation the reference does with ``Diff`` values
    (core/compiler/interpreters/incremental.py) restricted to changed / unchanged."""

    def d(flag):
        return Diff(None, UnknownChange if flag else NoChange)

    for s in ir.sites:
        if s.addr == addr:
            return tuple(d(_depends([a], moved, changed_args)) for a in s.args), d(s.index in moved)
    if addr in ir.subcalls:
        args, rets = ir.subcalls[addr]
        return (tuple(d(_depends([a], moved, changed_args)) for a in args),
                tuple(d(_depends([r], moved, changed_args)) for r in rets))
    raise MissingAddress(addr[0] if len(addr) == 1 else addr)


def _collect_static_request(request: StaticRequest, prefix: tuple, out: dict) -> None:
    """Flatten (possibly nested) per-address sub-requests into one constraint, one selection list, the custom
    requests and the ``DiffAnnotate`` callbacks addressed along the way."""
    from ..core.choice_map import _norm_addr

    for addr, sub in request.addressed.items():
        a = prefix + _norm_addr(addr)
        while isinstance(sub, DiffAnnotate):
            out["annot"].append((a, sub.argdiff_fn, sub.retdiff_fn))
            sub = sub.request
        if isinstance(sub, Update):
            out["constraint"] = out["constraint"] | ChoiceMap.entry(sub.constraint, *a)
        elif isinstance(sub, Regenerate):
            out["selected"].append(Selection.all().extend(*a) if sub.selection.check() else sub.selection.extend(*a))
        elif isinstance(sub, EmptyRequest):
            pass
        elif isinstance(sub, StaticRequest):
            _collect_static_request(sub, a, out)
        else:
            out["custom"].append((a, sub))


def _edit_static_request(gf, key, trace, request: StaticRequest, argdiffs):
    """StaticRequest (static.py:512-566, 867-904): per-address sub-requests, nested ones included.
    Update / Regenerate sub-requests fuse into ONE launch (constrained sites read, selected sites resampled, every
    site re-scored); Rejuvenate / HMC sub-requests run on their own batched drivers, in the order given, before
    it; ``DiffAnnotate`` callbacks see the changed / unchanged pattern of their callee's arguments and return value."""
    out = {"constraint": ChoiceMap.empty(), "selected": [], "custom": [], "annot": []}
    _collect_static_request(request, (), out)
    constraint, selected, custom, annot = out["constraint"], out["selected"], out["custom"], out["annot"]
    ir = trace.cm.ir
    sel = Selection.none()
    for s in selected:
        sel = sel | s

    if annot:
        moved = {s.index for s in ir.sites if sel(s.sel_addr).check() or constraint.get_submap(*s.addr).has_value()}
        for a, sub in custom:
            moved |= set(sub.moved_sites(trace, a))
        changed_args = _changed_arg_leaves(argdiffs)
        for a, argdiff_fn, retdiff_fn in annot:
            adiff, rdiff = _callee_diffs(ir, a, moved, changed_args)
            argdiff_fn(adiff)
            retdiff_fn(rdiff[0] if isinstance(rdiff, tuple) and len(rdiff) == 1 else rdiff)

    weight, bwd = None, ChoiceMap.empty()
    for a, sub in custom:
        if not hasattr(sub, "edit_at"):
            raise NotSupportedEditRequest(request)
        trace, w, _, b = sub.edit_at(key, trace, a, argdiffs)
        weight = w if weight is None else weight + w
        bwd = bwd | b.constraint  # the oldest discarded value wins: undoing goes back to the start
        argdiffs = ()  # the following steps start from the arguments the first one installed
    if custom and not selected and constraint.static_is_empty():
        return trace, weight, Diff.unknown_change(trace.get_retval()), Update(bwd)

    if not selected:
        tr, w, rd, b = gf.edit(key, trace, Update(constraint), argdiffs)
    elif constraint.static_is_empty():
        tr, w, rd, b = gf.edit(key, trace, Regenerate(sel), argdiffs)
    else:
        # both kinds at once: the backward request restores every touched site (distribution.py:179-300)
        new_args = Diff.tree_primal(argdiffs) if argdiffs is not None and argdiffs != () else trace.args
        if new_args == () and trace.args != ():
            new_args = trace.args
        new_args = _carry_batch_marks(new_args, trace.args)
        sel_addrs = {s.addr for s in ir.sites if sel(s.sel_addr).check() and not constraint.get_submap(*s.addr).has_value()}
        tr, w = gf._run(key, new_args, _rebatch_constraint(constraint, trace), prev=trace, sample_addrs=sel_addrs,
                        weight_mode="delta", n=trace.n, batched=trace.batched)
        discard = _choices_from_sites(trace, [s for s in ir.sites
                                              if s.addr in sel_addrs or constraint.get_submap(*s.addr).has_value()])
        w, rd, b = gf._w(tr, w), Diff.unknown_change(tr.get_retval()), Update(discard)
    if weight is not None:
        w = weight + w
        b = Update(bwd | b.constraint)
    return tr, w, rd, b


def gen(source: Callable) -> StaticGenerativeFunction:
    """``@gen`` (static.py:1044-1062)."""
    return StaticGenerativeFunction(source)

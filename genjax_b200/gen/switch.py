"""Dynamic structure: ``Switch``, ``MaskCombinator``, ``mix``, ``or_else`` inside the fused model kernel.

API mirror of src/genjax/_src/generative_functions/combinators/switch.py
(``Switch:98``, ``simulate:160-181``, ``assess:183-195``, ``generate:197-216``,
``edit:262-306``, ``switch:313``), mask.py (``MaskCombinator:111``,
``simulate:158``, ``generate:167``, ``edit:186-276``, ``assess:278-288``,
``mask:296``), mixture.py:27-81 and or_else.py:23-84.  SURVEY.md row 8f-3.

The reference runs every branch under ``jax.lax.switch`` / ``jnp.where`` and
keeps the selected results.  Here a combinator is INLINED into the enclosing
``@gen`` body when the body is captured: every site of a branch is recorded
with the predicate under which it exists (``SiteSpec.live`` for Switch
branches, ``SiteSpec.scored`` for masks, gen/capture.py), and the generated
kernel evaluates the predicates per particle -- one launch, no divergence
beyond the predicated statements, branch index free to differ between
particles.  Unselected Switch branches read back zeros (the reference's
zero-filled sub-traces); masked calls are still sampled, only their score is
dropped (``MaskTrace.build``: ``score = check * inner.score``).

Used directly (``switch.simulate(key, (idx, args0, args1))``) a combinator is
a one-call static model, so every GFI method, batching over a ``KeyBatch`` and
``genjax_b200.vmap`` work as for ``@gen`` functions.
"""

from __future__ import annotations

from ..core.choice_map import ChoiceMap
from ..core.mask import Mask
from . import capture as cap
from . import expr as E
from .expr import Expr, F32, I32
from .gfi import GenerativeFunction

__all__ = ["Switch", "MaskCombinator", "switch", "mask", "mix", "or_else", "tree_choose"]


def _as_index(idx) -> Expr:
    e = E.lift(idx)
    if e.ndim != 0:
        raise TypeError("the Switch index must be a scalar per particle")
    return e if e.dtype == I32 else E.cast(e, I32)


def tree_choose(idx: Expr, trees: list):
    """Leaf-wise selection among same-shaped return values (core/compiler/staging.py ``tree_choose``): dtypes are
    promoted (int32 with bool -> int32, anything with float -> float32), shapes must agree."""
    flat = [cap.flatten(t) for t in trees]
    leaves0, tree0 = flat[0]
    for leaves, tree in flat[1:]:
        if len(leaves) != len(leaves0) or _shape_of_tree(tree) != _shape_of_tree(tree0):
            raise ValueError("Switch branches must return values of the same structure")
    out = []
    for k in range(len(leaves0)):
        cands = [E.lift(fl[0][k]) for fl in flat]
        shapes = {c.shape for c in cands}
        if len(shapes) > 1:
            raise ValueError(f"Incompatible shapes for broadcasting: {sorted(shapes)}")
        if len({c.dtype for c in cands}) > 1:
            cands = [c if c.dtype == F32 else E.cast(c, F32) for c in cands]
        r = cands[-1]
        for i in range(len(cands) - 2, -1, -1):
            r = E.where(E.binary("eq", idx, i), cands[i], r)
        out.append(r)
    return cap.unflatten(tree0, out)


def _shape_of_tree(tree):
    kind, payload = tree
    if kind in ("tuple", "list"):
        return (kind, tuple(_shape_of_tree(p) for p in payload))
    if kind == "dict":
        return (kind, tuple((k, _shape_of_tree(v)) for k, v in payload))
    if kind == "mask":
        return (kind, tuple(_shape_of_tree(p) for p in payload))
    return (kind,)


class _InlinedCombinator(GenerativeFunction):
    """A combinator whose body is captured into the caller's kernel; on its own it runs as a one-call static model."""

    _static_fn = None

    def capture_inline(self, args):  # pragma: no cover - abstract
        raise NotImplementedError

    def _static(self):
        if self._static_fn is None:
            from .static import StaticGenerativeFunction

            fn = StaticGenerativeFunction(lambda *args: self.capture_inline(args))
            fn.__name__ = type(self).__name__.lower()
            self._decorate(fn)
            self._static_fn = fn
        return self._static_fn

    def _decorate(self, fn):
        pass

    def simulate(self, key, args):
        return self._static().simulate(key, tuple(args))

    def generate(self, key, constraint: ChoiceMap, args):
        return self._static().generate(key, constraint, tuple(args))

    def assess(self, sample: ChoiceMap, args):
        return self._static().assess(sample, tuple(args))

    def project(self, key, trace, selection):
        return self._static().project(key, trace, selection)

    def edit(self, key, trace, request, argdiffs):
        return self._static().edit(key, trace, request, argdiffs)

    def get_site_addresses(self, args):
        return self._static().get_site_addresses(tuple(args))

    def get_zero_trace(self, *args):
        return self._static().get_zero_trace(*args)

    # the reference's combinator methods compose (generative_function.py:1339-1467)
    def dimap(self, **kw):
        return self._static().dimap(**kw)

    def map(self, f):
        return self._static().map(f)

    def contramap(self, f):
        return self._static().contramap(f)


class Switch(_InlinedCombinator):
    """``Switch(*branches)(idx, args_0, ..., args_{n-1})``: run branch ``clamp(idx)`` with its argument tuple
    (switch.py:98-306).  Branches may be ``@gen`` functions, distributions, closures or other combinators, and need
    not share addresses; an address that several branches visit holds the selected branch's value."""

    def __init__(self, *branches):
        if not branches:
            raise ValueError("Switch needs at least one branch")
        self.branches = tuple(branches)

    def __repr__(self):
        return f"Switch({', '.join(map(repr, self.branches))})"

    def capture_inline(self, args):
        c = cap.current_capture()
        if c is None:
            raise RuntimeError("a Switch can only be traced inside a @gen function body")
        idx, branch_args = args[0], tuple(args[1:])
        if len(branch_args) != len(self.branches):  # switch.py:155-156
            raise AssertionError(f"Switch over {len(self.branches)} branches got {len(branch_args)} argument tuples")
        n = len(self.branches)
        k = E.binary("min", E.binary("max", _as_index(idx), 0), n - 1)  # out-of-bounds indices are clamped (:108)
        sid = c.new_id()
        rets = []
        for i, (f, a) in enumerate(zip(self.branches, branch_args)):
            a = tuple(a) if isinstance(a, (tuple, list)) else (a,)
            with c.frame("switch", sid, i, E.binary("eq", k, i)):
                rets.append(cap.inline_call(f, a))
        return tree_choose(k, rets)


class MaskCombinator(_InlinedCombinator):
    """``MaskCombinator(f)(check, *args)``: ``f`` runs either way; its score counts, and its choices and return value
    are valid, only where ``check`` holds (mask.py:111-288).  The return value is ``Mask(retval, check)``."""

    def __init__(self, gen_fn):
        self.gen_fn = gen_fn

    def __repr__(self):
        return f"MaskCombinator({self.gen_fn!r})"

    def _decorate(self, fn):
        fn._mask_of = self

    def capture_inline(self, args):
        c = cap.current_capture()
        if c is None:
            raise RuntimeError("a MaskCombinator can only be traced inside a @gen function body")
        check, inner_args = args[0], tuple(args[1:])
        flag = E.lift(check)
        if flag.ndim != 0:
            # mask.py ``ScalarFlag``: a vector of flags needs ``.vmap()`` (test_mask_combinator.py:226-244)
            raise TypeError("MaskCombinator takes a scalar flag per particle; map it over an axis of flags with vmap")
        pred = E.binary("ne", flag, 0) if flag.dtype == I32 else E.binary("ne", flag, 0.0)
        with c.frame("mask", c.new_id(), 0, pred):
            ret = cap.inline_call(self.gen_fn, inner_args)
        return Mask.build(ret, pred)

    def _inner_trace(self, trace):
        """Trace of the callee alone over the choices of ``trace`` (what ``MaskTrace.inner`` holds, mask.py:55-62)."""
        from .distributions import Distribution
        from .static import _rebatch

        chm = ChoiceMap.empty()
        for addr, v in _rebatch(trace).leaves():
            chm = chm | ChoiceMap.entry(v, *addr)
        inner_args = tuple(trace.args[1:])
        f = self.gen_fn
        if isinstance(f, Distribution):
            from .distributions import DistributionTrace, _flat_call

            m, flat = _flat_call(f, inner_args)
            v = chm.get_value() if chm.has_value() else None
            tr, _ = m._run(None, flat, ChoiceMap.entry(v, "_v"), weight_mode="none", n=trace.n, batched=trace.batched)
            return DistributionTrace(f, tr, inner_args)
        fn = f._static() if isinstance(f, _InlinedCombinator) else f
        tr, _ = fn._run(None, inner_args, chm, weight_mode="none", n=trace.n, batched=trace.batched)
        return tr


def _mapped_size(tree, axes, sizes: list):
    """Leading sizes of the leaves mapped along axis 0 (``axes``: a tree prefix of 0 / None, as jax.vmap's in_axes)."""
    if isinstance(axes, (tuple, list)) and isinstance(tree, (tuple, list)):
        if len(axes) != len(tree):
            raise ValueError("vmap in_axes specification must be a tree prefix of the corresponding value")
        for t, a in zip(tree, axes):
            _mapped_size(t, a, sizes)
        return
    if axes is None:
        return
    if axes != 0:
        raise NotImplementedError("only in_axes of 0 / None are supported")
    for leaf in cap.flatten(tree)[0]:
        if hasattr(leaf, "shape") and len(leaf.shape) >= 1:
            sizes.append(int(leaf.shape[0]))
        elif isinstance(leaf, (list, tuple)):
            sizes.append(len(leaf))
        else:
            raise ValueError("vmap in_axes=0 on a value without a leading axis")


def _index_mapped(tree, axes, i: int):
    if isinstance(axes, (tuple, list)) and isinstance(tree, (tuple, list)):
        return type(tree)(_index_mapped(t, a, i) for t, a in zip(tree, axes))
    if axes is None:
        return tree
    leaves, shape = cap.flatten(tree)
    return cap.unflatten(shape, [v[i] for v in leaves])


def unrolled_vmap(gen_fn, in_axes, axis_size, args):
    """``gen_fn.vmap(in_axes=...)(*args) @ addr`` inside an ``@gen`` body: the mapped axis is UNROLLED into the caller's
    fused kernel (vmap.py:180-218: the inner call per element, scores summed) -- element i's sites are recorded under
    ``(..., i, addr)`` and read back stacked (``chm[addr, :, sub]``), the per-element return values come back stacked.
    This is also how a vmapped function runs under an outer particle batch (the two axes the top-level ``Vmap`` cannot
    hold).  ``n x sites per element`` has to fit the kernel's site table."""
    c = cap.current_capture()
    if c is None:
        raise RuntimeError("a vmapped generative function can only be traced inside a @gen function body")
    args = tuple(args)
    axes = tuple(in_axes) if isinstance(in_axes, (tuple, list)) else (in_axes,) * len(args)
    if len(axes) != len(args):
        raise ValueError("vmap in_axes specification must be a tree prefix of the corresponding value")
    sizes: list = []
    for a, ax in zip(args, axes):
        _mapped_size(a, ax, sizes)
    if sizes and any(s != sizes[0] for s in sizes):
        raise IndexError(f"vmap got inconsistent sizes for the mapped axis: {sizes}")
    n = sizes[0] if sizes else axis_size
    if n is None:
        raise ValueError("vmap has nothing to map over (every in_axes entry is None); use repeat(n=...)")
    prefix, positions = c.prefix, c.scan_positions
    rets = []
    try:
        c.scan_positions = positions + (len(prefix),)
        for i in range(int(n)):
            c.prefix = prefix + (i,)
            rets.append(cap.inline_call(gen_fn, tuple(_index_mapped(a, ax, i) for a, ax in zip(args, axes))))
    finally:
        c.prefix, c.scan_positions = prefix, positions
    from .scan import _stack_lists

    return _stack_lists(rets)


class UnrolledVmap(_InlinedCombinator):
    """``combinator.vmap(in_axes=...)`` for the inlined combinators (``f.mask().vmap()``, ``f.switch(g).vmap()``)."""

    def __init__(self, gen_fn, in_axes=0, axis_size: int | None = None):
        self.gen_fn, self.in_axes, self.axis_size = gen_fn, in_axes, axis_size

    def __repr__(self):
        return f"Vmap({self.gen_fn!r}, in_axes={self.in_axes})"

    def capture_inline(self, args):
        return unrolled_vmap(self.gen_fn, self.in_axes, self.axis_size, args)


def _combinator_vmap(self, in_axes=0, axis_size: int | None = None):
    return UnrolledVmap(self, in_axes, axis_size)


_InlinedCombinator.vmap = _combinator_vmap
_InlinedCombinator.repeat = lambda self, n: UnrolledVmap(self, None, int(n))


def switch(*gen_fns) -> Switch:
    """``genjax.switch(f, g, ...)`` (switch.py:313-354)."""
    return Switch(*gen_fns)


def mask(f) -> MaskCombinator:
    """``@genjax.mask`` (mask.py:296-322)."""
    return MaskCombinator(f)


def mix(*gen_fns):
    """``genjax.mix(f, g, ...)(mixture_logits, args_f, args_g, ...)``: a categorical draw at ``"mixture_component"``
    picks the component traced at ``"component_sample"`` (mixture.py:27-81)."""
    from .distributions import categorical
    from .static import gen

    inner = Switch(*gen_fns)

    def mixture_model(mixture_logits, *args):
        mix_idx = categorical(logits=mixture_logits) @ "mixture_component"
        return inner(mix_idx, *args) @ "component_sample"

    return gen(mixture_model)


def or_else(if_gen_fn, else_gen_fn):
    """``f.or_else(g)(flag, args_f, args_g)``: ``f`` where the flag holds, ``g`` elsewhere (or_else.py:23-84:
    ``True`` maps to branch 0)."""

    def argument_mapping(b, if_args, else_args):
        idx = E.cast(E.unary("logical_not", E.lift(b)), I32)
        return (idx, if_args, else_args)

    return Switch(if_gen_fn, else_gen_fn).contramap(argument_mapping)


# method spellings (generative_function.py ``switch`` / ``mask`` / ``or_else`` / ``mix``)
def _m_switch(self, *others):
    return Switch(self, *others)


def _m_mask(self):
    return MaskCombinator(self)


def _m_or_else(self, other):
    return or_else(self, other)


def _m_mix(self, *others):
    return mix(self, *others)


def _install():
    from .gfi import GenerativeFunctionClosure

    for cls in (GenerativeFunction, GenerativeFunctionClosure):
        cls.switch = _m_switch
        cls.mask = _m_mask
        cls.or_else = _m_or_else
        cls.mix = _m_mix


_install()

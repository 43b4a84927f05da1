"""``Vmap`` / ``repeat``: a generative function mapped over an axis of its arguments.

API mirror of src/genjax/_src/generative_functions/combinators/vmap.py
(``Vmap:79``, ``VmapTrace:44``, ``simulate:180-189``, ``generate:191-218``,
``project:220-239``, ``edit_choice_map:241-275``, ``assess:363-375``,
``vmap:384``) and combinators/repeat.py:25-41.  SURVEY.md rows a16 / 8f-1.

The reference maps the inner GFI method with ``jax.vmap`` over ``split(key, n)``
and sums scores / weights over the axis.  Here the mapped axis IS the lane axis
of the fused model kernel: one launch over ``split(key, n)`` with the mapped
arguments passed per lane.  Choices are addressed ``[i, addr]`` / ``[:, addr]``.
A constraint that names only SOME indices (``C[0, "z"].set(3.0)``) splits the
lanes into contiguous runs with the same constrained addresses, one launch per
run (per-site flags are uniform within a launch).

With a single key the mapped axis is the launch's lane axis.  Under an OUTER
particle batch (a ``KeyBatch``, per-particle arguments) and inside an ``@gen``
body (``f.vmap(in_axes=...)(*args) @ "addr"``) the mapped axis is UNROLLED into
the fused kernel (``capture_inline`` -> gen/switch.py ``unrolled_vmap``): element
i's sites are ``(..., i, addr)``, read back stacked; ``n x sites per element``
must fit the site table.  For primitives, ``dist.vmap(...)`` / ``dist.repeat(n=)``
(gen/distributions.py) make ONE vector site instead.
"""

from __future__ import annotations

import torch

from ..core.choice_map import ChoiceMap, Selection
from ..core.key import KeyBatch, PRNGKey, split
from ..runtime import cabi
from . import capture as cap
from .gfi import Diff, EditRequest, GenerativeFunction, IndexRequest, NotSupportedEditRequest, Regenerate, Trace, Update
from .static import Batched, StaticGenerativeFunction, StaticTrace, _dev_tensor

__all__ = ["Vmap", "VmapTrace", "vmap_combinator", "repeat"]


def _axes_tuple(in_axes, n_args: int) -> tuple:
    if isinstance(in_axes, (tuple, list)):
        if len(in_axes) != n_args:
            raise ValueError("vmap in_axes specification must be a tree prefix of the corresponding value")
        return tuple(in_axes)
    return (in_axes,) * n_args


def _check_prefix(tree, axes) -> None:
    """``in_axes`` must be a tree prefix of the arguments (checked before any leaf is looked at, like jax.vmap)."""
    if isinstance(axes, (tuple, list)):
        if not isinstance(tree, (tuple, list)) or len(tree) != len(axes):
            raise ValueError("vmap in_axes specification must be a tree prefix of the corresponding value")
        for t, a in zip(tree, axes):
            _check_prefix(t, a)


def _mark_arg(tree, axes, sizes: list, device):
    """Wrap the leaves mapped along axis 0 in ``Batched`` (device tensors); collect their leading sizes."""
    if isinstance(axes, (tuple, list)):
        if not isinstance(tree, (tuple, list)) or len(tree) != len(axes):
            raise ValueError("vmap in_axes specification must be a tree prefix of the corresponding value")
        return type(tree)(_mark_arg(t, a, sizes, device) for t, a in zip(tree, axes))
    if axes is None:
        return tree
    if axes != 0:
        raise NotImplementedError("only in_axes of 0 / None are supported (the mapped axis is the lane axis)")

    def one(v):
        v = v.primal if isinstance(v, Diff) else v
        t = _dev_tensor(v, device)
        if t.ndim == 0:
            raise ValueError("vmap was requested to map its argument along axis 0, which implies that its rank "
                             "should be at least 1, but is only 0")
        sizes.append(int(t.shape[0]))
        return Batched(t)

    leaves, shape = cap.flatten(tree)
    return cap.unflatten(shape, [one(v) for v in leaves])


def _slice_args(tree, lo: int, hi: int):
    leaves, shape = cap.flatten(tree)
    return cap.unflatten(shape, [Batched(v.value[lo:hi].contiguous()) if isinstance(v, Batched) else v for v in leaves])


def replace_lane(inner: StaticTrace, idx: int, one: StaticTrace) -> StaticTrace:
    """``tree_map(lambda v, v_: v.at[idx].set(v_), inner, one)`` (vmap.py:330-332): a copy of ``inner`` whose lane
    ``idx`` holds the one-lane trace ``one``; traces are immutable, so every leaf is cloned."""
    ir = inner.cm.ir
    n = inner.n
    values = {}
    for s in ir.sites:
        j = s.index
        ev = tuple(s.value.shape)
        full = inner.values[j]
        full = full.reshape(ev).expand((n,) + ev) if inner.bcast[j] else full
        full = full.clone()
        full[idx] = one.values[j].reshape((-1,) + ev)[0]
        values[j] = full
    rets = []
    for ra, rb in zip(inner.ret_leaves, one.ret_leaves):
        if isinstance(ra, torch.Tensor):
            r = ra.clone()
            r[idx] = rb.reshape((-1,) + tuple(ra.shape[1:]))[0]
            rets.append(r)
        else:
            rets.append(ra)
    score = inner.score.clone()
    score[idx] = one.score.reshape(-1)[0]
    return StaticTrace(inner.gen_fn, inner.cm, inner.bound, inner.args, n, True, values, score, rets,
                       {j: False for j in values})


class VmapTrace(Trace):
    """vmap.py:44-76: ``inner`` keeps the mapped axis; the score is summed over it."""

    def __init__(self, gen_fn: "Vmap", inner, args, dim_length: int):
        self.gen_fn = gen_fn
        self.inner = inner
        self.args = args
        self.dim_length = dim_length

    def get_gen_fn(self):
        return self.gen_fn

    def get_args(self):
        return self.args

    def get_retval(self):
        return self.inner.get_retval()

    def get_score(self):
        return self.inner.get_score().sum()

    def get_choices(self) -> ChoiceMap:
        return self.inner.get_choices()

    def get_subtrace(self, *addr):
        return self.inner.get_subtrace(*addr)


class Vmap(GenerativeFunction):
    """``Vmap(gen_fn, in_axes)`` / ``gen_fn.vmap(in_axes=...)``: type ``[a] -> [b]``."""

    def __init__(self, gen_fn: StaticGenerativeFunction, in_axes=0, axis_size: int | None = None):
        if not isinstance(gen_fn, StaticGenerativeFunction):
            raise TypeError("Vmap needs an @gen function")
        self.gen_fn = gen_fn
        self.in_axes = in_axes
        self.axis_size = axis_size  # repeat(n=...): nothing is mapped, the axis has this length
        self.__name__ = f"vmap({gen_fn.__name__})"

    # -- nested use: ``f.vmap(in_axes=...)(*args) @ "addr"`` inside an @gen body: unrolled into the caller's kernel
    def capture_inline(self, args):
        from .switch import unrolled_vmap

        return unrolled_vmap(self.gen_fn, self.in_axes if self.in_axes is not None else None, self.axis_size, args)

    # -- helpers -----------------------------------------------------------
    def _bind(self, key, args):
        if isinstance(key, KeyBatch):
            raise NotImplementedError("a vmapped generative function under a particle batch needs two batch axes")
        device = cabi.require_cuda()
        args = tuple(args)
        sizes: list = []
        for a, ax in zip(args, _axes_tuple(self.in_axes, len(args))):
            _check_prefix(a, ax)
        marked = tuple(_mark_arg(a, ax, sizes, device) for a, ax in zip(args, _axes_tuple(self.in_axes, len(args))))
        if sizes and any(s != sizes[0] for s in sizes):
            raise IndexError(f"vmap got inconsistent sizes for the mapped axis: {sizes}")
        n = sizes[0] if sizes else self.axis_size
        if n is None:
            raise ValueError("vmap has nothing to map over (every in_axes entry is None); use repeat(n=...)")
        if self.axis_size is not None and n != self.axis_size:
            raise ValueError("mapped axis size disagrees with n=")
        return device, marked, n

    @staticmethod
    def _lane_constraints(constraint: ChoiceMap, n: int):
        """Per lane i the addresses constrained there, and a getter for the lane values of a run [lo, hi)."""
        if constraint.has_value():
            raise ValueError("a vmapped generative function needs addressed constraints")
        ints = {k for k in constraint.keys() if isinstance(k, int)}
        base = ChoiceMap(children={k: c for k, c in constraint._children.items() if not isinstance(k, int)})
        if not ints:  # vector constraints only: one run, the leaves already carry the mapped axis
            whole = base.map_leaves(lambda v: v if isinstance(v, Batched) else Batched(torch.as_tensor(v)))
            return [(0, n)], lambda lo, hi: whole
        sigs = []
        for i in range(n):
            addrs = {a for a, _ in base.leaves()}
            if i in ints:
                addrs |= {a for a, _ in constraint._children[i].leaves()}
            sigs.append(frozenset(addrs))

        def run_chm(lo: int, hi: int) -> ChoiceMap:
            out = ChoiceMap.empty()
            for addr in sorted(sigs[lo], key=repr):
                rows = []
                for i in range(lo, hi):
                    sub = constraint.get_submap(i, *addr)
                    rows.append(torch.as_tensor(sub.get_value()))
                out = out | ChoiceMap.entry(Batched(torch.stack(rows)), *addr)
            return out

        runs, lo = [], 0
        for i in range(1, n + 1):
            if i == n or sigs[i] != sigs[lo]:
                runs.append((lo, i))
                lo = i
        return runs, run_chm

    def _launch_runs(self, key, marked, n, constraint, prev, request_kind, selection=None):
        """One launch per run of lanes with the same constrained addresses; traces concatenated in lane order."""
        from ..inference.smc import _concat_traces

        kb = split(key, n) if key is not None else None
        if constraint is None or constraint.static_is_empty():
            runs, run_chm = [(0, n)], lambda lo, hi: ChoiceMap.empty()
        else:
            runs, run_chm = self._lane_constraints(constraint, n)
        inner, weight, discard = None, [], ChoiceMap.empty()
        for lo, hi in runs:
            args = marked if (lo, hi) == (0, n) else _slice_args(marked, lo, hi)
            k = None if kb is None else kb[lo:hi]
            chm = run_chm(lo, hi)
            if request_kind == "simulate":
                tr, w = self.gen_fn._run(k, args, None, weight_mode="none", n=hi - lo, batched=True)
            elif request_kind == "generate":
                tr, w = self.gen_fn._run(k, args, chm, weight_mode="generate", n=hi - lo, batched=True)
            elif request_kind == "assess":
                tr, w = self.gen_fn._run(None, args, chm, weight_mode="none", n=hi - lo, batched=True)
            else:
                old = prev if (lo, hi) == (0, n) else prev.take(torch.arange(lo, hi, device=prev.score.device))
                sub = Update(chm) if request_kind == "update" else Regenerate(selection)
                tr, w, _, bwd = self.gen_fn.edit(k, old, sub, Diff.unknown_change(args))
                if (lo, hi) == (0, n):
                    discard = bwd.constraint  # leaves keep the mapped axis: C[:, addr]
                else:
                    for addr, v in bwd.constraint.leaves():
                        for i in range(lo, hi):
                            discard = discard | ChoiceMap.entry(v[i - lo], i, *addr)
            inner = tr if inner is None else _concat_traces(inner, tr)
            if w is not None:
                weight.append(w)
        if len(runs) > 1:
            inner.args = marked  # the concatenation kept the first run's slice; the trace spans all lanes
        return inner, (torch.cat(weight).sum() if weight else None), discard

    def _unrolled(self) -> StaticGenerativeFunction:
        """The same function with the mapped axis unrolled into ONE static model (``capture_inline``): the form that runs
        under an OUTER particle batch -- a batched key, or per-particle arguments (vmap.py:180-275 under ``jax.vmap``)."""
        if getattr(self, "_unrolled_fn", None) is None:
            fn = StaticGenerativeFunction(lambda *args: self.capture_inline(args))
            fn.__name__ = f"{self.gen_fn.__name__}_vmap_unrolled"
            self._unrolled_fn = fn
        return self._unrolled_fn

    @staticmethod
    def _outer_batch(key, args) -> bool:
        if isinstance(key, KeyBatch):
            return True
        return any(isinstance(v, Batched) for v in cap.flatten(tuple(args))[0])

    # -- GFI ---------------------------------------------------------------
    def simulate(self, key: PRNGKey, args: tuple) -> VmapTrace:
        if self._outer_batch(key, args):
            return self._unrolled().simulate(key, tuple(args))
        _, marked, n = self._bind(key, args)
        inner, _, _ = self._launch_runs(key, marked, n, None, None, "simulate")
        return VmapTrace(self, inner, args, n)

    def generate(self, key: PRNGKey, constraint: ChoiceMap, args: tuple):
        if self._outer_batch(key, args):
            return self._unrolled().generate(key, constraint, tuple(args))
        device, marked, n = self._bind(key, args)
        inner, w, _ = self._launch_runs(key, marked, n, constraint, None, "generate")
        return VmapTrace(self, inner, args, n), w

    def assess(self, sample: ChoiceMap, args: tuple):
        if self._outer_batch(None, args) or any(isinstance(v, Batched) for _, v in sample.leaves()):
            return self._unrolled().assess(sample, tuple(args))
        _, marked, n = self._bind(None, args)
        inner, _, _ = self._launch_runs(None, marked, n, sample, None, "assess")
        tr = VmapTrace(self, inner, args, n)
        return tr.get_score(), tr.get_retval()

    def project(self, key, trace: VmapTrace, selection: Selection):
        if not isinstance(trace, VmapTrace):  # a trace of the unrolled form (outer particle batch)
            return self._unrolled().project(key, trace, selection)
        return self.gen_fn.project(key, trace.inner, selection).sum()

    def _edit_index(self, key, trace: VmapTrace, request: IndexRequest, argdiffs):
        """vmap.py:277-336: the sub-request runs on lane ``index`` alone, with ``key`` itself."""
        if not Diff.static_check_no_change(argdiffs if argdiffs not in (None, ()) else ()):
            raise AssertionError("IndexRequest needs unchanged arguments (vmap.py:285)")
        idx = request.index
        if not 0 <= idx < trace.dim_length:
            raise AssertionError(f"index {idx} is outside the mapped axis of length {trace.dim_length}")
        _, marked, n = self._bind(None, trace.args)
        lane = trace.inner.take(torch.tensor([idx], device=trace.inner.score.device))
        new_lane, w, _, bwd = self.gen_fn.edit(key, lane, request.request, Diff.no_change(_slice_args(marked, idx, idx + 1)))
        inner = replace_lane(trace.inner, idx, new_lane)
        inner.args = marked
        new = VmapTrace(self, inner, trace.args, n)
        return new, w.reshape(-1)[0], Diff.unknown_change(new.get_retval()), IndexRequest(idx, bwd)

    def edit(self, key, trace: VmapTrace, request: EditRequest, argdiffs):
        if not isinstance(trace, VmapTrace):  # a trace of the unrolled form (outer particle batch)
            if isinstance(request, IndexRequest):
                raise NotImplementedError("IndexRequest on a vmapped trace under an outer particle batch")
            return self._unrolled().edit(key, trace, request, argdiffs)
        if isinstance(request, IndexRequest):
            return self._edit_index(key, trace, request, argdiffs)
        if not isinstance(request, (Update, Regenerate)):
            if hasattr(request, "edit") and type(request).edit is not EditRequest.edit:
                return request.edit(key, trace, argdiffs)
            raise NotSupportedEditRequest(request)
        args = Diff.tree_primal(argdiffs) if argdiffs is not None and argdiffs != () else trace.args
        if args == ():
            args = trace.args
        _, marked, n = self._bind(key, args)
        if n != trace.dim_length:
            raise NotSupportedEditRequest(request)
        if isinstance(request, Update):
            inner, w, discard = self._launch_runs(key, marked, n, request.constraint, trace.inner, "update")
        else:
            inner, w, discard = self._launch_runs(key, marked, n, None, trace.inner, "regenerate", request.selection)
        new = VmapTrace(self, inner, args, n)
        return new, w, Diff.unknown_change(new.get_retval()), Update(discard)


def vmap_combinator(*, in_axes=0):
    """``@genjax.vmap(in_axes=...)`` applied to a generative function (vmap.py:384-420)."""

    def decorator(f: StaticGenerativeFunction) -> Vmap:
        return Vmap(f, in_axes)

    return decorator


def repeat(*, n: int):
    """``@genjax.repeat(n=...)``: ``a -> [b]``, n independent runs on the same arguments (repeat.py:25-41)."""

    def decorator(f: StaticGenerativeFunction) -> Vmap:
        return Vmap(f, in_axes=None, axis_size=n)

    return decorator


StaticGenerativeFunction.vmap = lambda self, *, in_axes=0: Vmap(self, in_axes)
StaticGenerativeFunction.repeat = lambda self, *, n: Vmap(self, in_axes=None, axis_size=n)

"""``Scan``: a kernel ``(carry, x) -> (carry, y)`` unrolled over the leading axis of ``xs``.

API mirror of src/genjax/_src/generative_functions/combinators/scan.py
(``Scan:107``, ``ScanTrace:56``, ``simulate:199-238``, ``generate:240-297``,
``project:299-327``, ``edit_regenerate:417-507``, ``edit_update:509-602``,
``assess:634-660``, ``scan:672``, ``accumulate:791``, ``reduce:854``,
``iterate:916``, ``iterate_final:980``).  SURVEY.md section 8f-1.

Where the reference runs the kernel's GFI method under ``jax.lax.scan`` (and the
particle axis comes from an outer ``jax.vmap``), this class launches the
kernel's ONE fused model kernel once per time step over all particles: step t
reads the carry of step t-1 as per-particle arguments and its slice ``xs[t]``
as shared scalars / vectors, with the key chain of the reference
(``key_t = fold_in(key_{t-1}, t)``, scan.py:213, 268 -- applied lane-wise here).
Choices are addressed ``[t, addr]`` / ``[:, addr]``; the trace stores one fused
``StaticTrace`` per step and stacks on demand.

This is the API row, not the fast path: a bootstrap filter over a time series
belongs on ``inference.pf.ParticleFilter`` (one launch per step, no per-step
host work).  A ``Scan`` called inside an ``@gen`` body (``kernel.scan(n=T)(c, xs)
@ "addr"``) is UNROLLED into the caller's fused kernel (``capture_inline``): the
length is static, ``T x sites-per-step`` must fit the site table.  The same
unrolled form (``Scan.unrolled``) carries moves over all steps at once, e.g. HMC
over ``Selection.at["x"]`` of a scanned trace (inference/mcmc.py).
"""

from __future__ import annotations

from typing import Callable

import numpy as np
import torch

from ..core.choice_map import ChoiceMap, ChoiceMapNoValueAtAddress, Selection
from ..core.key import KeyBatch, fold_in_lanes
from ..runtime import cabi
from . import capture as cap
from .gfi import (
    Diff,
    EditRequest,
    GenerativeFunction,
    IndexRequest,
    NotSupportedEditRequest,
    Regenerate,
    Trace,
    Update,
)
from .static import Batched, StaticGenerativeFunction, _depends, _dev_tensor, _is_scalar_number, cap_norm

__all__ = ["Scan", "ScanTrace", "scan", "accumulate", "reduce", "iterate", "iterate_final"]


def _tree_map(fn: Callable, tree):
    leaves, shape = cap.flatten(tree)
    return cap.unflatten(shape, [fn(v) for v in leaves])


def _primal(v):
    return v.primal if isinstance(v, Diff) else v


class _Xs:
    """The scanned-over inputs: per leaf a host copy (scalar slices become launch scalars, no device sync per
    step) and a device copy (vector slices are passed as shared arguments)."""

    def __init__(self, xs, length: int | None, device):
        leaves, self.shape = cap.flatten(xs)
        self.items = []
        sizes = []
        for v in leaves:
            v = _primal(v)
            if isinstance(v, Batched):
                t = _dev_tensor(v.value, device)
                if t.ndim < 2:
                    raise ValueError("a per-particle scanned input needs axes [particle, time, ...]")
                sizes.append(int(t.shape[1]))
                self.items.append(("batched", t, None))
            else:
                t = v if isinstance(v, torch.Tensor) else torch.as_tensor(np.asarray(v))
                if t.ndim == 0:
                    raise ValueError("scanned inputs need a leading time axis")
                sizes.append(int(t.shape[0]))
                host = t.detach().cpu()
                self.items.append(("shared", _dev_tensor(t, device) if t.ndim > 1 else None, host))
        if len(set(sizes)) > 1:
            raise ValueError("scan got values with different leading axis sizes: " + ", ".join(str(s) for s in sizes) + ".")
        if sizes and length is not None and sizes[0] != length:
            raise ValueError(f"scan got `length` argument of {length} which disagrees with leading axis sizes [{sizes[0]}].")
        if not sizes and length is None:
            raise ValueError("scan needs `n=` when nothing is scanned over")
        self.length = sizes[0] if sizes else int(length)

    def at(self, t: int):
        out = []
        for kind, dev, host in self.items:
            if kind == "batched":
                out.append(Batched(dev[:, t].contiguous()))
            elif host.ndim == 1:
                x = host[t].item()
                out.append(x if host.dtype.is_floating_point else int(x))
            else:
                out.append(dev[t])
        return cap.unflatten(self.shape, out)


def _batch_size(key, trees, constraint: ChoiceMap | None) -> int | None:
    sizes = []
    if isinstance(key, KeyBatch):
        sizes.append(key.n)
    for tree in trees:
        for v in cap.flatten(tree)[0]:
            v = _primal(v)
            if isinstance(v, Batched):
                sizes.append(int(v.value.shape[0]))
    if constraint is not None:
        for _, v in constraint.leaves():
            if isinstance(v, Batched):
                sizes.append(int(v.value.shape[0]))
    if not sizes:
        return None
    if any(s != sizes[0] for s in sizes):
        raise ValueError(f"inconsistent particle-axis sizes {sizes}")
    return sizes[0]


def _carry_in(carry, n: int, device):
    """Every carry leaf as a per-particle [n, ...] argument, so that all steps share ONE compiled signature
    (step 0 would otherwise see launch scalars and later steps per-particle tensors)."""

    def one(v):
        v = _primal(v)
        if isinstance(v, Batched):
            t = _dev_tensor(v.value, device)
            if t.shape[0] != n:
                raise ValueError("carry does not match the particle-axis size")
            return Batched(t)
        if _is_scalar_number(v):
            dt = torch.int32 if isinstance(v, (bool, int, np.integer)) else torch.float32
            return Batched(torch.full((n,), v, dtype=dt, device=device))
        t = _dev_tensor(v, device)
        return Batched(t.unsqueeze(0).expand((n,) + tuple(t.shape)).contiguous())

    return _tree_map(one, carry)


def _carry_next(ret_carry, n: int, device):
    def one(v):
        if isinstance(v, torch.Tensor):
            return Batched(v)
        if v is None:
            return None
        dt = torch.int32 if isinstance(v, (bool, int, np.integer)) else torch.float32
        return Batched(torch.full((n,), v, dtype=dt, device=device))  # a constant carry leaf

    return _tree_map(one, ret_carry)


def _stack_time(per_step: list, batched: bool):
    """Leaf-wise stack of per-step pytrees with [n, ...] leaves -> [n, T, ...] (or [T, ...] for a scalar call)."""
    if not per_step:
        return None
    flat = [cap.flatten(y) for y in per_step]
    shape = flat[0][1]
    out = []
    for j in range(len(flat[0][0])):
        col = [f[0][j] for f in flat]
        if isinstance(col[0], torch.Tensor):
            st = torch.stack(col, dim=1)
            out.append(st if batched else st[0])
        else:
            out.append(col[0] if all(c == col[0] for c in col) else col)
    return cap.unflatten(shape, out)


def _stack_lists(per_step: list):
    """Leaf-wise lists over the steps of an unrolled scan: ``[(a_0, b_0), (a_1, b_1)] -> ([a_0, a_1], [b_0, b_1])``."""
    if not per_step:
        return None
    flat = [cap.flatten(y) for y in per_step]
    shape = flat[0][1]
    return cap.unflatten(shape, [cap.StackedList(f[0][j] for f in flat) for j in range(len(flat[0][0]))])


# -------------------------------------------------------------------- trace


class ScanTrace(Trace):
    """scan.py:56-99.  ``inner[t]`` is the fused trace of step t (particle axis leading)."""

    def __init__(self, gen_fn: "Scan", inner: list, args, carry_out, ys: list, score, n: int, batched: bool):
        self.gen_fn = gen_fn
        self.inner = inner
        self.args = args
        self.carry_out = carry_out  # pytree of Batched leaves
        self.ys = ys  # per-step pytrees
        self.score = score  # [n]
        self.n = n
        self.batched = batched
        self.scan_length = len(inner)

    def get_gen_fn(self):
        return self.gen_fn

    def get_args(self):
        return self.args

    def get_score(self):
        return self.score if self.batched else self.score[0]

    def get_retval(self):
        def view(v):
            v = v.value if isinstance(v, Batched) else v
            return v if self.batched or not isinstance(v, torch.Tensor) else v[0]

        ret = (_tree_map(view, self.carry_out), _stack_time(self.ys, self.batched))
        return self.gen_fn.post(self.args, ret) if self.gen_fn.post is not None else ret

    def get_choices(self) -> ChoiceMap:
        if not self.inner:
            return ChoiceMap.empty()
        chm = ChoiceMap.empty()
        for s in self.inner[0].cm.ir.sites:
            st = torch.stack([tr._site_value(s) for tr in self.inner], dim=1)
            chm = chm | ChoiceMap.entry(st if self.batched else st[0], *s.addr)
        return chm

    def get_inner_trace(self, t: int):
        return self.inner[t]

    def get_subtrace(self, *addr):
        return _ScanSubTrace(self, cap_norm(addr))


class _ScanSubTrace(Trace):
    """``tr.get_subtrace("y")`` of a scan: scores stacked over time (tests/core/generative/test_core.py:151-158)."""

    def __init__(self, parent: ScanTrace, addr: tuple):
        if not parent.inner:
            raise ChoiceMapNoValueAtAddress(addr)
        self.parent = parent
        self.addr = addr
        self.subs = [tr.get_subtrace(*addr) for tr in parent.inner]

    def get_score(self):
        st = torch.stack([s.get_score() for s in self.subs], dim=1)  # inner traces keep the particle axis
        return st if self.parent.batched else st[0]

    def get_choices(self) -> ChoiceMap:
        return self.parent.get_choices().get_submap(*self.addr)

    def get_retval(self):
        st = torch.stack([s.get_retval() for s in self.subs], dim=1)
        return st if self.parent.batched else st[0]

    def get_gen_fn(self):
        return self.subs[0].get_gen_fn()

    def get_args(self):
        raise NotImplementedError("per-site arguments are fused away")


# ------------------------------------------------------- generative function


class Scan(GenerativeFunction):
    """``Scan(kernel_gen_fn, length=n)`` / ``kernel.scan(n=n)``: type ``(c, [a]) -> (c, [b])``."""

    def __init__(self, kernel_gen_fn: StaticGenerativeFunction, length: int | None = None, post=None):
        if not isinstance(kernel_gen_fn, StaticGenerativeFunction):
            raise TypeError("Scan needs an @gen kernel of type (carry, x) -> (carry, y)")
        self.kernel_gen_fn = kernel_gen_fn
        self.length = length
        self.post = post  # (args, (carry, ys)) -> retval, for accumulate / reduce / iterate
        self.__name__ = f"scan({kernel_gen_fn.__name__})"

    def __repr__(self):
        return f"Scan({self.kernel_gen_fn!r}, length={self.length})"

    # -- helpers -----------------------------------------------------------
    @staticmethod
    def _unpack(args):
        args = tuple(args)
        if len(args) != 2:
            raise TypeError("a scanned generative function takes (carry, xs)")
        return args

    # -- nested use: ``kernel.scan(n=T)(carry, xs) @ "addr"`` inside an @gen body ------------------------------------
    def capture_inline(self, args):
        """The scan UNROLLED into the caller's fused kernel: step t's sites are recorded under ``(..., t, addr)`` (the
        reference addresses them ``[..., t, addr]`` / ``[..., :, addr]``, scan.py:81-99), the carry is threaded through
        as traced values.  The length must be static (``n=`` or the leading size of ``xs``) and ``T x sites per step``
        has to fit the kernel's site table (GJB_MAX_SITES); the per-step outputs come back as Python lists of length T."""
        c = cap.current_capture()
        if c is None:
            raise RuntimeError("a Scan can only be traced inside a @gen function body")
        carry, xs = self._unpack(args)
        leaves, shape = cap.flatten(xs)
        sizes = {int(v.shape[0]) for v in leaves if hasattr(v, "shape") and len(v.shape) >= 1}
        if len(sizes) > 1 or (sizes and self.length is not None and sizes != {self.length}):
            raise ValueError(f"scan got values with different leading axis sizes: {sorted(sizes)} (n={self.length})")
        if not sizes and self.length is None:
            raise ValueError("scan needs `n=` when nothing is scanned over")
        T = sizes.pop() if sizes else int(self.length)
        prefix, positions = c.prefix, c.scan_positions
        init, ys = carry, []
        try:
            c.scan_positions = positions + (len(prefix),)
            for t in range(T):
                c.prefix = prefix + (t,)
                x_t = cap.unflatten(shape, [v[t] for v in leaves])
                carry, y = self.kernel_gen_fn.capture_inline((carry, x_t))
                ys.append(y)
        finally:
            c.prefix, c.scan_positions = prefix, positions
        stacked = _stack_lists(ys)
        if self.post is None:
            return carry, stacked
        if self.post is _prepend_initial:  # accumulate / iterate: [init, c_1, ..., c_T]
            i_leaves, _ = cap.flatten(init)
            c_leaves, c_shape = cap.flatten(stacked, is_leaf=lambda v: isinstance(v, list))
            return cap.unflatten(c_shape, [cap.StackedList([i] + list(cs)) for i, cs in zip(i_leaves, c_leaves)])
        return self.post((init, xs), (carry, stacked))

    def unrolled(self, T: int) -> StaticGenerativeFunction:
        """The whole scan as ONE static model ``(carry, xs) -> carry`` (sites ``(t, addr)``): what a move over all steps
        at once -- HMC over ``Selection.at["x"]`` of a scanned trace, tests/inference/test_requests.py:237-255 -- runs on."""
        cache = self.__dict__.setdefault("_unrolled", {})
        fn = cache.get(T)
        if fn is None:
            inner = Scan(self.kernel_gen_fn, length=T)
            unpack = self._unpack

            def source(*args):
                carry, xs = unpack(args)
                return inner.capture_inline((carry, xs))[0]

            fn = StaticGenerativeFunction(source)
            fn.__name__ = f"{self.kernel_gen_fn.__name__}_unrolled{T}"
            cache[T] = fn
        return fn

    def _setup(self, key, args, constraint):
        device = cabi.require_cuda()
        carry, xs = self._unpack(args)
        xs_ = _Xs(xs, self.length, device)
        n = _batch_size(key, (carry, xs), constraint)
        batched = n is not None
        n = n if batched else 1
        return device, _carry_in(carry, n, device), xs_, n, batched

    def _finish(self, inner, args, carry, ys, score, weight, n, batched, device):
        if score is None:
            score = torch.zeros(n, dtype=torch.float32, device=device)
        tr = ScanTrace(self, inner, args, carry, ys, score, n, batched)
        if weight is None:
            return tr, None
        return tr, (weight if batched else weight[0])

    @staticmethod
    def _acc(tot, x):
        return x if tot is None else tot + x

    # -- GFI ---------------------------------------------------------------
    def simulate(self, key, args: tuple) -> ScanTrace:
        device, carry, xs, n, batched = self._setup(key, args, None)
        inner, ys, score, k = [], [], None, key
        for t in range(xs.length):
            k = fold_in_lanes(k, t)
            tr, _ = self.kernel_gen_fn._run(k, (carry, xs.at(t)), None, weight_mode="none", n=n, batched=True)
            ret_carry, y = tr.get_retval()
            carry = _carry_next(ret_carry, n, device)
            inner.append(tr)
            ys.append(y)
            score = self._acc(score, tr.score)
        return self._finish(inner, args, carry, ys, score, None, n, batched, device)[0]

    def generate(self, key, constraint: ChoiceMap, args: tuple):
        device, carry, xs, n, batched = self._setup(key, args, constraint)
        inner, ys, score, weight, k = [], [], None, None, key
        for t in range(xs.length):
            k = fold_in_lanes(k, t)
            tr, w = self.kernel_gen_fn._run(k, (carry, xs.at(t)), constraint.get_submap(t), weight_mode="generate", n=n,
                                            batched=True)
            ret_carry, y = tr.get_retval()
            carry = _carry_next(ret_carry, n, device)
            inner.append(tr)
            ys.append(y)
            score = self._acc(score, tr.score)
            weight = self._acc(weight, w)
        if weight is None:
            weight = torch.zeros(n, dtype=torch.float32, device=device)
        return self._finish(inner, args, carry, ys, score, weight, n, batched, device)

    def assess(self, sample: ChoiceMap, args: tuple):
        device, carry, xs, n, batched = self._setup(None, args, sample)
        ys, score = [], None
        for t in range(xs.length):
            tr, _ = self.kernel_gen_fn._run(None, (carry, xs.at(t)), sample.get_submap(t), weight_mode="none", n=n,
                                            batched=True)
            ret_carry, y = tr.get_retval()
            carry = _carry_next(ret_carry, n, device)
            ys.append(y)
            score = self._acc(score, tr.score)
        tr = self._finish([], args, carry, ys, score, None, n, batched, device)[0]
        return tr.get_score(), tr.get_retval()

    def project(self, key, trace: ScanTrace, selection: Selection):
        tot = None
        for tr in trace.inner:
            w = self.kernel_gen_fn.project(key, tr, selection)
            tot = self._acc(tot, w)
        if tot is None:
            tot = torch.zeros_like(trace.score)
        return tot if trace.batched else tot[0]

    def _edit_index(self, key, trace: ScanTrace, request: IndexRequest, argdiffs):
        """scan.py:329-415: the sub-request runs on step ``index`` with unchanged arguments; the next step is
        re-visited with the new carry, and -- like the reference -- the edit is only allowed when the carry LEAVING
        the kernel does not read the carry entering it (otherwise every later step would have to move)."""
        if not Diff.static_check_no_change(argdiffs if argdiffs not in (None, ()) else ()):
            raise AssertionError("IndexRequest needs unchanged arguments (scan.py:339)")
        idx = request.index
        T = trace.scan_length
        if not 0 <= idx < T:
            raise AssertionError(f"index {idx} is outside the scan of length {T}")
        device = cabi.require_cuda()
        n = trace.n
        old = trace.inner[idx]
        ir = old.cm.ir
        n_carry = len(cap.flatten(old.args[0])[0])
        ret_carry = cap.flatten(cap.unflatten(ir.ret_tree, list(ir.ret_leaves))[0])[0]
        assert not _depends(ret_carry, set(), set(range(n_carry))), \
            "IndexRequest on a scan needs a kernel whose outgoing carry does not read the incoming one (scan.py:368-372)"
        new_step, w, _, bwd = self.kernel_gen_fn.edit(key, old, request.request, Diff.no_change(old.args))
        inner = list(trace.inner)
        ys = list(trace.ys)
        inner[idx] = new_step
        carry, ys[idx] = new_step.get_retval()
        carry = _carry_next(carry, n, device)
        score = trace.score - old.score + new_step.score
        w = w.reshape(-1) if isinstance(w, torch.Tensor) else w
        if idx + 1 < T:
            nxt = trace.inner[idx + 1]
            nxt_args = (carry, nxt.args[1])
            new_next, w2, _, _ = self.kernel_gen_fn.edit(key, nxt, Update(ChoiceMap.empty()), Diff.unknown_change(nxt_args))
            inner[idx + 1] = new_next
            _, ys[idx + 1] = new_next.get_retval()
            score = score - nxt.score + new_next.score
            w = w + w2.reshape(-1)
            carry_out = trace.carry_out  # unchanged by the assertion above
        else:
            carry_out = carry
        new = ScanTrace(self, inner, trace.args, carry_out, ys, score, n, trace.batched)
        return new, (w if trace.batched else w[0]), Diff.unknown_change(new.get_retval()), IndexRequest(idx, bwd)

    def edit(self, key, trace: ScanTrace, request: EditRequest, argdiffs):
        if isinstance(request, IndexRequest):
            return self._edit_index(key, trace, request, argdiffs)
        if not isinstance(request, (Update, Regenerate)):
            if hasattr(request, "edit") and type(request).edit is not EditRequest.edit:
                return request.edit(key, trace, argdiffs)
            raise NotSupportedEditRequest(request)
        args = Diff.tree_primal(argdiffs) if argdiffs is not None and argdiffs != () else trace.args
        if args == ():
            args = trace.args
        constraint = request.constraint if isinstance(request, Update) else None
        device = cabi.require_cuda()
        carry, xs = self._unpack(args)
        xs = _Xs(xs, self.length, device)
        found = _batch_size(key, (carry,), constraint)
        if found is not None and found != trace.n:
            raise ValueError("edit arguments do not match the trace's particle-axis size")
        n, batched = trace.n, trace.batched
        carry = _carry_in(carry, n, device)
        if xs.length != trace.scan_length:
            raise NotSupportedEditRequest(request)  # changing the length of a scan is an `extend`, not an update
        inner, ys, score, weight, k = [], [], None, None, key
        discard = ChoiceMap.empty()
        for t in range(xs.length):
            k = fold_in_lanes(k, t)
            old = trace.inner[t]
            step_args = (carry, xs.at(t))
            if isinstance(request, Update):
                sub = Update(constraint.get_submap(t))
            else:
                sub = Regenerate(request.selection)
            tr, w, _, bwd = self.kernel_gen_fn.edit(k, old, sub, Diff.unknown_change(step_args))
            ret_carry, y = tr.get_retval()
            carry = _carry_next(ret_carry, n, device)
            inner.append(tr)
            ys.append(y)
            score = self._acc(score, tr.score)
            weight = self._acc(weight, w)
            if not bwd.constraint.static_is_empty():
                d = bwd.constraint if batched else bwd.constraint.map_leaves(
                    lambda v: v[0] if isinstance(v, torch.Tensor) and v.ndim >= 1 else v)  # scalar call: no particle axis
                discard = discard | d.extend(t)
        if weight is None:
            weight = torch.zeros(n, dtype=torch.float32, device=device)
        new_tr, w = self._finish(inner, args, carry, ys, score, weight, n, batched, device)
        return new_tr, w, Diff.unknown_change(new_tr.get_retval()), Update(discard)


# ------------------------------------------------------------- constructors


def scan(*, n: int | None = None):
    """``@genjax.scan(n=...)`` (scan.py:672-759)."""

    def decorator(f: StaticGenerativeFunction) -> Scan:
        return Scan(f, length=n)

    return decorator


def _with_source(f: StaticGenerativeFunction, source: Callable, name: str) -> StaticGenerativeFunction:
    out = StaticGenerativeFunction(source)
    out.__name__ = f"{f.__name__}_{name}"
    return out


def _prepend_initial(args, ret):
    """``[init, c_1, ..., c_T]`` (prepend_initial_acc, scan.py:762-788)."""
    init, (_, cs) = args[0], ret

    def cat(i, arr):
        if isinstance(i, Batched):  # per-particle init [n, ...] next to [n, T, ...]
            return torch.cat([i.value.to(arr.dtype).unsqueeze(1), arr], dim=1)
        it = torch.as_tensor(i, dtype=arr.dtype, device=arr.device)
        lead = tuple(arr.shape[: arr.ndim - it.ndim - 1])  # () for a scalar call, (n,) for a batched one
        return torch.cat([it.expand(lead + (1,) + tuple(it.shape)), arr], dim=len(lead))

    init_leaves, _ = cap.flatten(init)
    c_leaves, c_shape = cap.flatten(cs)
    return cap.unflatten(c_shape, [cat(i, c) for i, c in zip(init_leaves, c_leaves)])


def accumulate():
    """``(c, a) -> c`` to ``(c, [a]) -> [c]`` including the initial value (scan.py:791-851)."""

    def decorator(f: StaticGenerativeFunction) -> Scan:
        def both(c, x):
            r = f.source(c, x)
            return r, r

        return Scan(_with_source(f, both, "accumulate"), post=_prepend_initial)

    return decorator


def reduce():
    """``(c, a) -> c`` to ``(c, [a]) -> c`` (scan.py:854-913)."""

    def decorator(f: StaticGenerativeFunction) -> Scan:
        def carry_only(c, x):
            return f.source(c, x), None

        return Scan(_with_source(f, carry_only, "reduce"), post=lambda args, ret: ret[0])

    return decorator


class _Iterate(Scan):
    """``iterate`` / ``iterate_final`` take ``(init,)`` and scan over nothing (scan.py:916-1037)."""

    @staticmethod
    def _unpack(args):
        args = tuple(args)
        if len(args) != 1:
            raise TypeError("an iterated generative function takes (init,)")
        return args[0], None


def iterate(*, n: int):
    def decorator(f: StaticGenerativeFunction) -> Scan:
        def both(c, _):
            r = f.source(c)
            return r, r

        return _Iterate(_with_source(f, both, "iterate"), length=n, post=_prepend_initial)

    return decorator


def iterate_final(*, n: int):
    def decorator(f: StaticGenerativeFunction) -> Scan:
        def carry_only(c, _):
            return f.source(c), None

        return _Iterate(_with_source(f, carry_only, "iterate_final"), length=n, post=lambda args, ret: ret[0])

    return decorator


def _install_methods():
    """``gen_fn.scan(n=) / .accumulate() / .reduce() / .iterate(n=) / .iterate_final(n=)``
    (generative_function.py:776-1085)."""
    StaticGenerativeFunction.scan = lambda self, *, n=None: scan(n=n)(self)
    StaticGenerativeFunction.accumulate = lambda self: accumulate()(self)
    StaticGenerativeFunction.reduce = lambda self: reduce()(self)
    StaticGenerativeFunction.iterate = lambda self, *, n: iterate(n=n)(self)
    StaticGenerativeFunction.iterate_final = lambda self, *, n: iterate_final(n=n)(self)


_install_methods()

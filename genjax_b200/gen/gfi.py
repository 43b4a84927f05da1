"""Generative-function interface: abstract base, closure syntax, traces,
edit requests and argdiffs.

API mirror of src/genjax/_src/core/generative/generative_function.py
(``Trace:72``, ``GenerativeFunction:238``, ``importance:629-675``,
``update:611``, ``propose:677``, ``GenerativeFunctionClosure`` ``__matmul__:1568``,
``Update:1688``), concepts.py (``EditRequest:95``), requests.py
(``EmptyRequest:49``, ``Regenerate:64``) and core/compiler/interpreters/
incremental.py (``Diff:89``, ``NoChange``, ``UnknownChange``).
"""

from __future__ import annotations


from ..core.choice_map import ChoiceMap, Selection


# --------------------------------------------------------------------- Diff


class ChangeType:
    pass


class _NoChange(ChangeType):
    def __repr__(self):
        return "NoChange"


class _UnknownChange(ChangeType):
    def __repr__(self):
        return "UnknownChange"


NoChange = _NoChange()
UnknownChange = _UnknownChange()


class Diff:
    """A value tagged with change information (incremental.py:89-294)."""

    __slots__ = ("primal", "tangent")

    def __init__(self, primal, tangent: ChangeType):
        self.primal = primal
        self.tangent = tangent

    def get_primal(self):
        return self.primal

    def get_tangent(self):
        return self.tangent

    @staticmethod
    def _map(tree, fn):
        if isinstance(tree, Diff):
            return fn(tree)
        if isinstance(tree, tuple):
            return tuple(Diff._map(t, fn) for t in tree)
        if isinstance(tree, list):
            return [Diff._map(t, fn) for t in tree]
        if isinstance(tree, dict):
            return {k: Diff._map(v, fn) for k, v in tree.items()}
        return fn(tree)

    @staticmethod
    def no_change(tree):
        return Diff._map(tree, lambda v: Diff(v.primal if isinstance(v, Diff) else v, NoChange))

    @staticmethod
    def unknown_change(tree):
        return Diff._map(tree, lambda v: Diff(v.primal if isinstance(v, Diff) else v, UnknownChange))

    @staticmethod
    def tree_primal(tree):
        return Diff._map(tree, lambda v: v.primal if isinstance(v, Diff) else v)

    tree_diff_no_change = no_change
    tree_diff_unknown_change = unknown_change

    @staticmethod
    def tree_tangent(tree):
        return Diff._map(tree, lambda v: v.tangent if isinstance(v, Diff) else NoChange)

    @staticmethod
    def static_check_no_change(tree) -> bool:
        ok = True

        def chk(v):
            nonlocal ok
            if isinstance(v, Diff) and v.tangent is not NoChange:
                ok = False
            return v

        Diff._map(tree, chk)
        return ok

    @staticmethod
    def static_check_tree_diff(tree) -> bool:
        ok = True

        def chk(v):
            nonlocal ok
            if not isinstance(v, Diff):
                ok = False
            return v

        Diff._map(tree, chk)
        return ok

    def __repr__(self):
        return f"Diff({self.primal!r}, {self.tangent!r})"


# ----------------------------------------------------------------- requests


class NotSupportedEditRequest(Exception):
    """distribution.py:341-342, static.py:980-981."""

    def __init__(self, request):
        self.request = request
        super().__init__(request)


class EditRequest:
    """concepts.py:95-131: ``request.edit(key, trace, argdiffs)``."""

    def edit(self, key, tr: "Trace", argdiffs):
        return tr.get_gen_fn().edit(key, tr, self, argdiffs)

    def dimap(self, *, pre=lambda v: v, post=lambda v: v):
        return DiffAnnotate(self, argdiff_fn=pre, retdiff_fn=post)

    def map(self, post):
        return self.dimap(post=post)

    def contramap(self, pre):
        return self.dimap(pre=pre)


class PrimitiveEditRequest(EditRequest):
    pass


class EmptyRequest(EditRequest):
    """requests.py:49-61."""

    def edit(self, key, tr, argdiffs):
        if Diff.static_check_no_change(argdiffs):
            return tr, _zero_weight(tr), Diff.no_change(tr.get_retval()), EmptyRequest()
        return Update(ChoiceMap.empty()).edit(key, tr, argdiffs)


class Update(PrimitiveEditRequest):
    """generative_function.py:1688."""

    def __init__(self, constraint: ChoiceMap):
        self.constraint = constraint

    def __repr__(self):
        return f"Update({self.constraint!r})"


class Regenerate(PrimitiveEditRequest):
    """requests.py:64-66."""

    def __init__(self, selection: Selection):
        self.selection = selection

    def __repr__(self):
        return f"Regenerate({self.selection!r})"


class IndexRequest(EditRequest):
    """``IndexRequest(index, request)`` (generative_function.py ``IndexRequest``): apply ``request`` to ONE index of a
    vectorised trace (``Vmap`` lane, ``Scan`` step); handled by those combinators' ``edit``."""

    def __init__(self, index, request: EditRequest):
        self.index = int(index)
        self.request = request

    def __repr__(self):
        return f"IndexRequest({self.index}, {self.request!r})"


class StaticRequest(EditRequest):
    """static.py:130-131: per-address sub-requests."""

    def __init__(self, addressed: dict):
        self.addressed = dict(addressed)


class DiffAnnotate(EditRequest):
    """requests.py:70-95."""

    def __init__(self, request: EditRequest, argdiff_fn=lambda v: v, retdiff_fn=lambda v: v):
        self.request = request
        self.argdiff_fn = argdiff_fn
        self.retdiff_fn = retdiff_fn

    def edit(self, key, tr, argdiffs):
        new_tr, w, retdiff, bwd = self.request.edit(key, tr, self.argdiff_fn(argdiffs))
        return new_tr, w, self.retdiff_fn(retdiff), bwd


def _zero_weight(tr):
    import torch

    s = tr.get_score()
    return torch.zeros_like(s)


# -------------------------------------------------------------------- Trace


class Trace:
    """generative_function.py:72-230."""

    def get_args(self) -> tuple:
        raise NotImplementedError

    def get_retval(self):
        raise NotImplementedError

    def get_score(self):
        raise NotImplementedError

    def get_choices(self) -> ChoiceMap:
        raise NotImplementedError

    def get_sample(self) -> ChoiceMap:
        return self.get_choices()

    def get_gen_fn(self) -> "GenerativeFunction":
        raise NotImplementedError

    def edit(self, key, request: EditRequest, argdiffs=None):
        if argdiffs is None:
            argdiffs = Diff.no_change(self.get_args())
        return request.edit(key, self, argdiffs)

    def update(self, key, constraint: ChoiceMap, argdiffs=None):
        if argdiffs is None:
            argdiffs = Diff.no_change(self.get_args())
        return self.get_gen_fn().update(key, self, constraint, argdiffs)

    def project(self, key, selection: Selection):
        return self.get_gen_fn().project(key, self, selection)


# ------------------------------------------------------- GenerativeFunction


class GenerativeFunctionClosure:
    """``gen_fn(*args, **kwargs)`` (generative_function.py:1558-1690): ``closure @ "addr"`` traces it inside a
    ``@gen`` body; ``closure(key)`` runs it and returns the return value; the GFI methods take the REMAINING
    arguments.  Keyword arguments go through ``gen_fn.handle_kwargs()``, whose arguments are ``(args, kwargs)``."""

    def __init__(self, gen_fn, args: tuple, kwargs: dict):
        self.gen_fn = gen_fn
        self.args = args
        self.kwargs = kwargs

    def _packed_args(self):
        return (self.args, self.kwargs) if self.kwargs else self.args

    def __matmul__(self, addr):
        from .capture import trace_site

        return trace_site(addr, self.gen_fn, self._packed_args())

    def _target(self, args=(), kwargs=None):
        """(generative function to run, its argument tuple)."""
        full = self.args + tuple(args)
        kw = {**self.kwargs, **(kwargs or {})}
        if kw:
            return self.gen_fn.handle_kwargs(), (full, kw)
        return self.gen_fn, full

    def __call__(self, key, *args, **kwargs):
        fn, full = self._target(args, kwargs)
        return fn.simulate(key, full).get_retval()

    def handle_kwargs(self):
        return self

    # direct GFI use: genjax.normal(0., 1.).simulate(key, ())
    def simulate(self, key, args=()):
        fn, full = self._target(args)
        return fn.simulate(key, full)

    def generate(self, key, constraint, args=()):
        fn, full = self._target(args)
        return fn.generate(key, constraint, full)

    importance = generate

    def assess(self, sample, args=()):
        fn, full = self._target(args)
        return fn.assess(sample, full)

    def propose(self, key, args=()):
        fn, full = self._target(args)
        return fn.propose(key, full)

    def project(self, key, trace, selection):
        return self.gen_fn.project(key, trace, selection)

    def _full(self, args):
        return self._target(args)[1]


class GenerativeFunction:
    """Abstract GFI (generative_function.py:238-699)."""

    def __call__(self, *args, **kwargs) -> GenerativeFunctionClosure:
        return GenerativeFunctionClosure(self, args, kwargs)

    # -- abstract
    def simulate(self, key, args: tuple) -> Trace:
        raise NotImplementedError

    def assess(self, sample: ChoiceMap, args: tuple):
        raise NotImplementedError

    def generate(self, key, constraint: ChoiceMap, args: tuple):
        raise NotImplementedError

    def project(self, key, trace: Trace, selection: Selection):
        raise NotImplementedError

    def edit(self, key, trace: Trace, request: EditRequest, argdiffs):
        raise NotImplementedError

    # -- derived (generative_function.py:611-689)
    def importance(self, key, constraint: ChoiceMap, args: tuple):
        return self.generate(key, constraint, args)

    def update(self, key, trace: Trace, constraint: ChoiceMap, argdiffs=None):
        if argdiffs is None:
            argdiffs = Diff.no_change(trace.get_args())
        new_tr, w, retdiff, bwd = self.edit(key, trace, Update(constraint), argdiffs)
        assert isinstance(bwd, Update)
        return new_tr, w, retdiff, bwd.constraint

    def propose(self, key, args: tuple):
        tr = self.simulate(key, args)
        return tr.get_choices(), tr.get_score(), tr.get_retval()

    def marginal(self, *, selection: Selection | None = None, algorithm=None, reference_compat: bool = True):
        """``gen_fn.marginal(selection=..., algorithm=...)`` (generative_function.py ``marginal``; sp.py:208-273).
        ``reference_compat=False`` opts into the corrected ``random_weighted`` weight (see ``Marginal``)."""
        from ..inference.sp import Marginal

        return Marginal(self, Selection.all() if selection is None else selection, algorithm, reference_compat)

    def partial_apply(self, *bound):
        """``gen_fn.partial_apply(*args)`` (generative_function.py ``partial_apply``): the same generative function
        with its first arguments fixed; the choices keep their addresses."""
        raise NotImplementedError(f"partial_apply is not available for {type(self).__name__}")

    def handle_kwargs(self):
        """A version of this function whose arguments are ``(args, kwargs)`` (generative_function.py:1466-1480)."""
        return self

    def get_zero_trace(self, *args):
        raise NotImplementedError(f"get_zero_trace is not available for {type(self).__name__}")

"""ChoiceMap / Selection algebra (host-side, static addresses).

API mirror of src/genjax/_src/core/generative/choice_map.py for the subset the
hot path uses: ``Static`` + ``Choice`` maps (``ChoiceMap:847``, ``Choice:1397``,
``Static:1535``), the builder ``C[...]`` (``_ChoiceMapBuilder:752``) and the
``Selection`` algebra (``Selection:124``: all / none / leaf / static / and / or
/ complement).  In the reference these are pytrees whose leaves are traced
arrays; here they are plain host objects whose leaves are torch tensors (or
Python scalars) and they resolve, at capture time, to per-site flags
{sampled, constrained, selected} for the fused kernels.

Dynamic structure (``Indexed``, ``Switch``, ``Or`` with traced flags, masks)
is out of scope and raises ``NotImplementedError``.
"""

from __future__ import annotations

from typing import Any, Iterable

__all__ = [
    "ChoiceMap",
    "ChoiceMapBuilder",
    "ChoiceMapNoValueAtAddress",
    "Selection",
    "SelectionBuilder",
]


class ChoiceMapNoValueAtAddress(Exception):
    """Raised by ``chm[addr]`` when there is no value (choice_map.py:672, 1314-1315)."""

    def __init__(self, addr):
        self.addr = addr
        super().__init__(addr)


def _norm_addr(addr) -> tuple:
    if isinstance(addr, tuple):
        out = []
        for a in addr:
            out.extend(_norm_addr(a))
        return tuple(out)
    if addr is Ellipsis or isinstance(addr, slice):
        return (addr,)
    if isinstance(addr, (str, int)):
        return (addr,)
    raise TypeError(f"unsupported address component {addr!r}")


# ------------------------------------------------------------------ Selection


class Selection:
    """Set of addresses.  ``sel[addr]`` / ``addr in sel`` test membership;
    ``sel(addr)`` descends one level (choice_map.py:124-325)."""

    # kinds: all, none, static{name->Selection}, and, or, not
    def __init__(self, kind: str, payload: Any = None):
        self.kind = kind
        self.payload = payload

    # constructors -----------------------------------------------------
    @staticmethod
    def all() -> "Selection":
        return Selection("all")

    @staticmethod
    def none() -> "Selection":
        return Selection("none")

    @staticmethod
    def leaf() -> "Selection":
        return Selection("leaf")

    @staticmethod
    def at_addr(addr) -> "Selection":
        sel = Selection.all()
        for comp in reversed(_norm_addr(addr)):
            if comp is Ellipsis or isinstance(comp, slice):
                continue  # wildcard over an index axis: static models have none
            sel = Selection("static", {comp: sel})
        return sel

    class _At:
        def __getitem__(self, addr) -> "Selection":
            return Selection.at_addr(addr)

    at = _At()

    # algebra ----------------------------------------------------------
    def __or__(self, other: "Selection") -> "Selection":
        return Selection("or", (self, other))

    def __and__(self, other: "Selection") -> "Selection":
        return Selection("and", (self, other))

    def __invert__(self) -> "Selection":
        return Selection("not", self)

    def complement(self) -> "Selection":
        return ~self

    def extend(self, *addr) -> "Selection":
        sel = self
        for comp in reversed(_norm_addr(addr)):
            sel = Selection("static", {comp: sel})
        return sel

    # queries ----------------------------------------------------------
    def __call__(self, addr) -> "Selection":
        """Sub-selection under ``addr``."""
        sel = self
        for comp in _norm_addr(addr):
            sel = sel._descend(comp)
        return sel

    def _descend(self, comp) -> "Selection":
        k = self.kind
        if k in ("all", "none"):
            return self
        if k == "leaf":
            return Selection.none()
        if k == "static":
            return self.payload.get(comp, Selection.none())
        if k == "or":
            return self.payload[0]._descend(comp) | self.payload[1]._descend(comp)
        if k == "and":
            return self.payload[0]._descend(comp) & self.payload[1]._descend(comp)
        if k == "not":
            return ~self.payload._descend(comp)
        raise AssertionError(k)

    def check(self) -> bool:
        """Is the empty address ``()`` selected?"""
        k = self.kind
        if k in ("all", "leaf"):
            return True
        if k in ("none", "static"):
            return False
        if k == "or":
            return self.payload[0].check() or self.payload[1].check()
        if k == "and":
            return self.payload[0].check() and self.payload[1].check()
        if k == "not":
            return not self.payload.check()
        raise AssertionError(k)

    def __getitem__(self, addr) -> bool:
        return self(addr).check()

    def __contains__(self, addr) -> bool:
        return self(addr).check()

    def __repr__(self):
        if self.kind == "static":
            return "Selection{" + ", ".join(f"{k!r}: {v!r}" for k, v in self.payload.items()) + "}"
        if self.kind in ("or", "and"):
            op = " | " if self.kind == "or" else " & "
            return "(" + op.join(repr(p) for p in self.payload) + ")"
        if self.kind == "not":
            return f"~{self.payload!r}"
        return f"Selection.{self.kind}()"


class _SelectionBuilder:
    def __getitem__(self, addr) -> Selection:
        return Selection.at_addr(addr)


SelectionBuilder = _SelectionBuilder()


# ------------------------------------------------------------------ ChoiceMap


class ChoiceMap:
    """Immutable tree: either a leaf value (``Choice``) or a dict name -> ChoiceMap
    (``Static``).  Construct with ``ChoiceMap.d / kw / choice / empty`` or the
    builder ``C[...]``."""

    __slots__ = ("_value", "_children", "_has_value")

    def __init__(self, value=None, children=None, has_value=False):
        self._value = value
        self._children = dict(children) if children else {}
        self._has_value = has_value

    # constructors -----------------------------------------------------
    @staticmethod
    def empty() -> "ChoiceMap":
        return ChoiceMap()

    @staticmethod
    def choice(v) -> "ChoiceMap":
        if isinstance(v, ChoiceMap):
            return v
        return ChoiceMap(value=v, has_value=True)

    value = choice

    @staticmethod
    def d(d: dict) -> "ChoiceMap":
        out = ChoiceMap.empty()
        for addr, v in d.items():
            out = out | ChoiceMap.entry(v, *_norm_addr(addr))
        return out

    @staticmethod
    def kw(**kwargs) -> "ChoiceMap":
        return ChoiceMap.d(kwargs)

    @staticmethod
    def entry(v, *addr) -> "ChoiceMap":
        if isinstance(v, dict):
            v = ChoiceMap.d(v)
        chm = v if isinstance(v, ChoiceMap) else ChoiceMap.choice(v)
        return chm.extend(*addr)

    @staticmethod
    def builder():
        return ChoiceMapBuilder

    # structure --------------------------------------------------------
    def extend(self, *addr) -> "ChoiceMap":
        chm = self
        for comp in reversed(_norm_addr(addr)):
            if comp is Ellipsis or isinstance(comp, slice):
                continue  # vectorised leading axis: leaves already carry it
            if not isinstance(comp, str):
                raise NotImplementedError("indexed (dynamic) choice-map addresses are out of scope")
            if chm.static_is_empty():
                return ChoiceMap.empty()
            chm = ChoiceMap(children={comp: chm})
        return chm

    def static_is_empty(self) -> bool:
        return (not self._has_value) and all(c.static_is_empty() for c in self._children.values())

    def get_value(self):
        return self._value if self._has_value else None

    def has_value(self) -> bool:
        return self._has_value

    def get_submap(self, *addr) -> "ChoiceMap":
        chm = self
        for comp in _norm_addr(addr):
            if comp is Ellipsis or isinstance(comp, slice):
                continue
            chm = chm._children.get(comp, _EMPTY)
        return chm

    def __call__(self, *addr) -> "ChoiceMap":
        return self.get_submap(*addr)

    def __getitem__(self, addr):
        sub = self.get_submap(addr)
        if not sub._has_value:
            raise ChoiceMapNoValueAtAddress(addr)
        return sub._value

    def __contains__(self, addr) -> bool:
        return self.get_submap(addr)._has_value

    def keys(self) -> Iterable[str]:
        return self._children.keys()

    def leaves(self, prefix=()) -> Iterable[tuple[tuple, Any]]:
        """(address tuple, value) pairs in insertion order."""
        if self._has_value:
            yield prefix, self._value
        for k, c in self._children.items():
            yield from c.leaves(prefix + (k,))

    def to_dict(self) -> dict:
        return {(a[0] if len(a) == 1 else a): v for a, v in self.leaves()}

    # algebra ----------------------------------------------------------
    def merge(self, other: "ChoiceMap") -> "ChoiceMap":
        """``self | other``: self wins where both have a value (choice_map.py:1227)."""
        if other is None or other.static_is_empty():
            return self
        if self.static_is_empty():
            return other
        if self._has_value:
            return self
        if other._has_value:
            return other
        children = dict(self._children)
        for k, c in other._children.items():
            children[k] = children[k].merge(c) if k in children else c
        return ChoiceMap(children=children)

    def __or__(self, other):
        return self.merge(other)

    def __xor__(self, other: "ChoiceMap") -> "ChoiceMap":
        """Disjoint union (deprecated in the reference); raises on overlap."""
        for addr, _ in other.leaves():
            if addr in self or (addr == () and self._has_value):
                raise Exception(f"The two choice maps have an overlapping address {addr!r}.")
        return self.merge(other)

    def filter(self, selection: Selection) -> "ChoiceMap":
        """Keep the leaves whose address is in ``selection`` (choice_map.py:896)."""
        if self._has_value:
            return self if selection.check() else _EMPTY
        children = {}
        for k, c in self._children.items():
            f = c.filter(selection(k))
            if not f.static_is_empty():
                children[k] = f
        return ChoiceMap(children=children)

    def get_selection(self) -> Selection:
        if self._has_value:
            return Selection.all()
        if not self._children:
            return Selection.none()
        return Selection("static", {k: c.get_selection() for k, c in self._children.items()})

    def mask(self, flag):
        if isinstance(flag, bool):
            return self if flag else _EMPTY
        raise NotImplementedError("masked choice maps with traced flags are out of scope")

    class _AtSetter:
        def __init__(self, chm, addr):
            self.chm = chm
            self.addr = addr

        def set(self, v) -> "ChoiceMap":
            return ChoiceMap.entry(v, *self.addr) | self.chm

    class _At:
        def __init__(self, chm):
            self.chm = chm

        def __getitem__(self, addr):
            return ChoiceMap._AtSetter(self.chm, _norm_addr(addr))

    @property
    def at(self):
        return ChoiceMap._At(self)

    def map_leaves(self, fn) -> "ChoiceMap":
        if self._has_value:
            return ChoiceMap.choice(fn(self._value))
        return ChoiceMap(children={k: c.map_leaves(fn) for k, c in self._children.items()})

    def __repr__(self):
        if self._has_value:
            return f"Choice({self._value!r})"
        return "ChoiceMap{" + ", ".join(f"{k!r}: {c!r}" for k, c in self._children.items()) + "}"


_EMPTY = ChoiceMap()


class _Builder:
    """``C["x"].set(v)``, ``C["a", "b"].set(v)``, ``C.n()``, ``C.v(v)``, ``C.d({...})``, ``C.kw(...)``."""

    def __init__(self, addr=()):
        self.addr = addr

    def __getitem__(self, addr) -> "_Builder":
        return _Builder(self.addr + _norm_addr(addr))

    def set(self, v) -> ChoiceMap:
        return ChoiceMap.entry(v, *self.addr)

    def n(self) -> ChoiceMap:
        return ChoiceMap.empty()

    def v(self, v) -> ChoiceMap:
        return self.set(v)

    def d(self, d: dict) -> ChoiceMap:
        return self.set(ChoiceMap.d(d))

    def kw(self, **kwargs) -> ChoiceMap:
        return self.set(ChoiceMap.kw(**kwargs))


ChoiceMapBuilder = _Builder()

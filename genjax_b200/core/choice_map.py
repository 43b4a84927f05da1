"""ChoiceMap / Selection algebra (host-side).

API mirror of src/genjax/_src/core/generative/choice_map.py for the part the
hot path uses: ``Static`` + ``Choice`` maps (``ChoiceMap:847``, ``Choice:1397``,
``Static:1535``), vectorised leaves addressed through ``[:, "x"]`` / ``[i, "x"]``
(``Indexed:1448`` with full slices and static integer indices), the builder
``C[...]`` (``_ChoiceMapBuilder:752``) and the ``Selection`` algebra
(``Selection:124``: all / none / leaf / static incl. the ``...`` wildcard / and
/ or / complement, with the same constructor-time simplifications, so that
``==`` between selections behaves as in the reference's tests,
tests/core/test_choice_maps.py).  In the reference these are pytrees whose
leaves are traced arrays; here they are plain host objects whose leaves are
torch tensors (or Python scalars) and they resolve, at capture time, to per-site
flags {sampled, constrained, selected} for the fused kernels.

Dynamic structure: leaves may be ``Mask(value, flag)`` (core/mask.py) -- what the
traces of ``Switch`` / ``MaskCombinator`` models report and what a Mask-ed
constraint is written as (gen/switch.py, gen/static.py).  Array-valued index
addresses and ``chm.mask(traced flag)`` on whole maps raise
``NotImplementedError``.
"""

from __future__ import annotations

from typing import Any, Callable, Iterable

__all__ = [
    "ChoiceMap",
    "ChoiceMapBuilder",
    "ChoiceMapNoValueAtAddress",
    "Selection",
    "SelectionBuilder",
]

_FULL = slice(None, None, None)


class ChoiceMapNoValueAtAddress(Exception):
    """Raised by ``chm[addr]`` when there is no value (choice_map.py:672, 1314-1315)."""

    def __init__(self, addr):
        self.addr = addr
        super().__init__(addr)


def _is_index_array(a) -> bool:
    return hasattr(a, "shape") and hasattr(a, "dtype") and not isinstance(a, (str, bytes))


def _norm_addr(addr) -> tuple:
    """Flatten nested address tuples into one tuple of components (str / int / slice / Ellipsis)."""
    if isinstance(addr, tuple):
        out = []
        for a in addr:
            out.extend(_norm_addr(a))
        return tuple(out)
    if addr is Ellipsis or isinstance(addr, (slice, str)):
        return (addr,)
    if isinstance(addr, bool):
        raise TypeError(f"unsupported address component {addr!r}")
    if isinstance(addr, int):
        return (addr,)
    if _is_index_array(addr):
        if tuple(addr.shape) == ():
            return (int(addr),)  # a concrete 0-d index is a static index here (nothing is traced on the host)
        raise NotImplementedError("array-valued (dynamic) index addresses are out of scope")
    raise TypeError(f"unsupported address component {addr!r}")


def _validate_dynamic(addr: tuple, allow_partial_slice: bool) -> None:
    """Index components must be scalars first, then (for lookups only) one partial slice, then full slices
    (choice_map.py:695-749)."""
    dyn = [c for c in addr if isinstance(c, (slice, int))]
    k = 0
    while k < len(dyn) and isinstance(dyn[k], int):
        k += 1
    rest = dyn[k:]
    if rest and allow_partial_slice and rest[0] != _FULL:
        rest = rest[1:]
    if not all(isinstance(s, slice) and s == _FULL for s in rest):
        what = "an optional partial slice, and then only full slices" if allow_partial_slice else "full slices"
        raise ValueError(f"Address must consist of scalar components, followed by {what}. Found: {dyn}")


# ------------------------------------------------------------------ Selection


class Selection:
    """Set of addresses.  ``sel[addr]`` / ``addr in sel`` test membership;
    ``sel(addr)`` descends (choice_map.py:124-325)."""

    # kinds: all, none, leaf, static{component -> Selection} (component may be ``...``), and, or, not
    __slots__ = ("kind", "payload")

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
        comps = _norm_addr(addr)
        if comps == ():
            return Selection.leaf()
        return Selection.all().extend(*comps)

    # algebra (constructor-time simplifications of AndSel / OrSel / ComplementSel .build) -----------
    def __or__(self, other: "Selection") -> "Selection":
        if self.kind == "all" or other.kind == "none":
            return self
        if other.kind == "all" or self.kind == "none":
            return other
        if self == other:
            return self
        return Selection("or", (self, other))

    def __and__(self, other: "Selection") -> "Selection":
        if self.kind == "all" or other.kind == "none":
            return other
        if other.kind == "all" or self.kind == "none":
            return self
        if self == other:
            return self
        return Selection("and", (self, other))

    def __invert__(self) -> "Selection":
        if self.kind == "all":
            return Selection.none()
        if self.kind == "none":
            return Selection.all()
        if self.kind == "not":
            return self.payload
        return Selection("not", self)

    def complement(self) -> "Selection":
        return ~self

    def extend(self, *addr) -> "Selection":
        """Nest under the given components; ``...`` matches any one component (choice_map.py:283-312)."""
        sel = self
        for comp in reversed(_norm_addr(addr)):
            if isinstance(comp, slice):
                continue  # an index axis is transparent to selections (Indexed.filter passes them through)
            if sel.kind == "none":
                return sel
            sel = Selection("static", {comp: sel})
        return sel

    def filter(self, sample: "ChoiceMap") -> "ChoiceMap":
        return sample.filter(self)

    # queries ----------------------------------------------------------
    def __call__(self, addr) -> "Selection":
        """Sub-selection under ``addr``."""
        sel = self
        for comp in _norm_addr(addr):
            sel = sel.get_subselection(comp)
        return sel

    def get_subselection(self, comp) -> "Selection":
        if comp is Ellipsis or isinstance(comp, slice):
            raise TypeError("wildcards are only allowed when BUILDING a selection, not when querying one")
        k = self.kind
        if k in ("all", "none"):
            return self
        if k == "leaf":
            return Selection.none()
        if k == "static":
            if Ellipsis in self.payload:
                return self.payload[Ellipsis]
            return self.payload.get(comp, Selection.none())
        if k == "or":
            return self.payload[0].get_subselection(comp) | self.payload[1].get_subselection(comp)
        if k == "and":
            return self.payload[0].get_subselection(comp) & self.payload[1].get_subselection(comp)
        if k == "not":
            return ~self.payload.get_subselection(comp)
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

    def __eq__(self, other) -> bool:
        return isinstance(other, Selection) and self.kind == other.kind and self.payload == other.payload

    def __ne__(self, other) -> bool:
        return not self == other

    def __hash__(self):
        return hash(self.kind)

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
    """``S["x", "y"]``, ``S[..., "y"]``, ``S[()]``, ``S.all`` / ``S.none`` / ``S.leaf`` (choice_map.py:75-119)."""

    @property
    def all(self) -> Selection:
        return Selection.all()

    @property
    def none(self) -> Selection:
        return Selection.none()

    @property
    def leaf(self) -> Selection:
        return Selection.leaf()

    def __getitem__(self, addr) -> Selection:
        return Selection.at_addr(addr)


SelectionBuilder = _SelectionBuilder()
Selection.at = SelectionBuilder


# ------------------------------------------------------------------ ChoiceMap


def _leaf_equal(a, b) -> bool:
    if isinstance(a, ChoiceMap) or isinstance(b, ChoiceMap):
        return isinstance(a, ChoiceMap) and isinstance(b, ChoiceMap) and a == b
    ta, tb = hasattr(a, "shape"), hasattr(b, "shape")
    if ta or tb:
        try:
            import numpy as np

            xa = a.detach().cpu().numpy() if hasattr(a, "detach") else np.asarray(a)
            xb = b.detach().cpu().numpy() if hasattr(b, "detach") else np.asarray(b)
            return xa.shape == xb.shape and bool((xa == xb).all())
        except Exception:
            return False
    try:
        return bool(a == b)
    except Exception:
        return False


def _index_leaf(v, idx):
    """``v[idx]`` on a vectorised leaf (Choice.get_inner_map with an index component, choice_map.py:1437-1443)."""
    if isinstance(v, ChoiceMap):
        return v.get_submap(idx)
    if hasattr(v, "value") and type(v).__name__ == "Batched":  # genjax_b200.gen.static.Batched marker
        return type(v)(v.value[:, idx])  # the particle axis is invisible to addresses, as under jax.vmap
    if hasattr(v, "__getitem__") and not isinstance(v, (str, bytes)):
        return v[idx]
    raise TypeError(f"leaf {v!r} has no index axis")


class ChoiceMap:
    """Immutable tree: either a leaf value (``Choice``) or a dict component -> ChoiceMap
    (``Static``; integer components are static indices).  Construct with
    ``ChoiceMap.d / kw / from_mapping / choice / empty`` or the builder ``C[...]``."""

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
        if hasattr(v, "shape") and tuple(v.shape) == (0,):
            return ChoiceMap()  # an empty array is an empty choice map (Choice.build, choice_map.py:1411-1413)
        return ChoiceMap(value=v, has_value=True)

    value = choice

    @staticmethod
    def from_mapping(pairs: Iterable[tuple]) -> "ChoiceMap":
        """Address/value pairs; dict values nest; later pairs fill in, earlier ones win (choice_map.py:1043-1073)."""
        out = ChoiceMap.empty()
        for addr, v in pairs:
            out = out | ChoiceMap.entry(v, *_norm_addr(addr))
        return out

    @staticmethod
    def d(d: dict) -> "ChoiceMap":
        return ChoiceMap.from_mapping(d.items())

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
    def switch(idx, chms) -> "ChoiceMap":
        if isinstance(idx, int) and not isinstance(idx, bool):
            return list(chms)[idx]
        raise NotImplementedError("switch over a traced index is out of scope")

    # structure --------------------------------------------------------
    def extend(self, *addr) -> "ChoiceMap":
        chm = self
        for comp in reversed(_norm_addr(addr)):
            if comp is Ellipsis:
                raise TypeError("`...` is a Selection wildcard, not a choice-map address")
            if isinstance(comp, slice):
                if comp != _FULL and not chm.static_is_empty():
                    raise ValueError(f"Partial slices not supported: {comp}")
                continue  # full slice over a vectorised axis: the leaves already carry it (Indexed.build)
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

    def get_inner_map(self, comp) -> "ChoiceMap":
        """One address component down (``get_inner_map`` of Choice / Static / Indexed in the reference)."""
        if isinstance(comp, str):
            return self._children.get(comp, _EMPTY)
        if comp is Ellipsis:
            raise TypeError("`...` is a Selection wildcard, not a choice-map address")
        if isinstance(comp, slice):
            return self if comp == _FULL else self.map_leaves(lambda v: _index_leaf(v, comp))
        if isinstance(comp, int):
            # vectorised leaves are indexed; entries stored under this static index (C[comp, ...].set) win over them
            if self._has_value:
                return self.map_leaves(lambda v: _index_leaf(v, comp))
            rest = ChoiceMap(children={k: c for k, c in self._children.items() if not isinstance(k, int)})
            out = rest.map_leaves(lambda v: _index_leaf(v, comp))
            if comp in self._children:
                out = _prefer(self._children[comp], out)
            return out
        raise TypeError(f"unsupported address component {comp!r}")

    def get_submap(self, *addr) -> "ChoiceMap":
        comps = _norm_addr(addr)
        _validate_dynamic(comps, allow_partial_slice=True)
        chm = self
        for comp in comps:
            chm = chm.get_inner_map(comp)
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

    def keys(self) -> Iterable:
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
        """``self | other``: self wins where both have a value (Or.build, choice_map.py:1799-1833)."""
        if other is None or other.static_is_empty():
            return self
        if self.static_is_empty():
            return other
        if self._has_value and other._has_value:
            return self
        if self._has_value or other._has_value:
            raise Exception(f"Choice and non-Choice in Or: {self}, {other}")
        children = dict(self._children)
        for k, c in other._children.items():
            children[k] = children[k].merge(c) if k in children else c
        return ChoiceMap(children=children)

    def __or__(self, other):
        return self.merge(other)

    __add__ = __or__

    def __xor__(self, other: "ChoiceMap") -> "ChoiceMap":
        """Deprecated alias of ``|`` in the reference (choice_map.py:1261-1267)."""
        return self.merge(other)

    def __and__(self, other: "ChoiceMap") -> "ChoiceMap":
        """Addresses present in both, values from the right-hand side (choice_map.py:1272-1273)."""
        return other.filter(self.get_selection())

    def filter(self, selection) -> "ChoiceMap":
        """Keep the leaves whose address is in ``selection`` (choice_map.py:896); a bool acts as a mask."""
        if isinstance(selection, bool):
            return self if selection else _EMPTY
        if not isinstance(selection, Selection):
            raise NotImplementedError("masked choice maps with traced flags are out of scope")
        if self._has_value:
            return self if selection.check() else _EMPTY
        children = {}
        for k, c in self._children.items():
            # a static integer index is transparent to selections, like the reference's Indexed layer
            f = c.filter(selection if isinstance(k, int) else selection.get_subselection(k))
            if not f.static_is_empty():
                children[k] = f
        return ChoiceMap(children=children)

    def get_selection(self) -> Selection:
        """Selection of exactly the addresses holding a value (ChmSel, choice_map.py:624-660)."""
        if self._has_value:
            return Selection.leaf()
        if self.static_is_empty():
            return Selection.none()
        sel = Selection.none()
        for k, c in self._children.items():
            sub = c.get_selection()
            sel = sel | (sub if isinstance(k, int) else sub.extend(k))
        return sel

    def mask(self, flag):
        """``chm.mask(flag)`` (choice_map.py ``mask``): a concrete flag keeps or empties the map; a flag array wraps every
        leaf in ``Mask(value, flag)`` -- the form a Mask-ed constraint takes (distribution.py:129-142)."""
        if isinstance(flag, bool):
            return self.filter(flag)
        if hasattr(flag, "shape") and hasattr(flag, "dtype"):
            from .mask import Mask

            return self.map_leaves(lambda v: Mask.build(v, flag))
        return self.filter(_not_a_flag(flag))

    def simplify(self) -> "ChoiceMap":
        return self

    def invalid_subset(self, gen_fn, args) -> "ChoiceMap | None":
        """Choices that ``gen_fn(*args)`` can never visit, or None (choice_map.py:1344-1376)."""
        valid = Selection.none()
        for addr in gen_fn.get_site_addresses(args):
            valid = valid | Selection.leaf().extend(*addr)
        extras = self.filter(~valid)
        return None if extras.static_is_empty() else extras

    @property
    def at(self) -> "_Builder":
        return _Builder(self, ())

    def map_leaves(self, fn) -> "ChoiceMap":
        if self._has_value:
            return ChoiceMap.choice(fn(self._value))
        return ChoiceMap(children={k: c.map_leaves(fn) for k, c in self._children.items()})

    def __eq__(self, other) -> bool:
        if not isinstance(other, ChoiceMap):
            return NotImplemented
        if self._has_value or other._has_value:
            return self._has_value and other._has_value and _leaf_equal(self._value, other._value)
        mine = {k: c for k, c in self._children.items() if not c.static_is_empty()}
        theirs = {k: c for k, c in other._children.items() if not c.static_is_empty()}
        return mine.keys() == theirs.keys() and all(mine[k] == theirs[k] for k in mine)

    def __ne__(self, other) -> bool:
        r = self.__eq__(other)
        return r if r is NotImplemented else not r

    __hash__ = object.__hash__

    def __repr__(self):
        if self._has_value:
            return f"Choice({self._value!r})"
        return "ChoiceMap{" + ", ".join(f"{k!r}: {c!r}" for k, c in self._children.items()) + "}"


def _prefer(first: ChoiceMap, second: ChoiceMap) -> ChoiceMap:
    """``first | second`` where ``first`` also wins structurally (a value over a subtree and vice versa)."""
    if second.static_is_empty():
        return first
    if first.static_is_empty():
        return second
    if first._has_value or second._has_value:
        return first
    children = dict(second._children)
    for k, c in first._children.items():
        children[k] = _prefer(c, children[k]) if k in children else c
    return ChoiceMap(children=children)


def _not_a_flag(flag):
    raise NotImplementedError(f"masked choice maps with traced flags are out of scope (got {type(flag).__name__})")


_EMPTY = ChoiceMap()


class _Builder:
    """``C["x"].set(v)``, ``C["a", "b"].set(v)``, ``C[:, "x"].set(vec)``, ``C.n()``, ``C.v(v)``, ``C.d({...})``,
    ``C.kw(...)``, ``C.from_mapping(...)``; ``chm.at[addr].set(v)`` / ``.update(fn)`` edit an existing map
    (_ChoiceMapBuilder, choice_map.py:752-845)."""

    def __init__(self, base: ChoiceMap | None = None, addr=()):
        self.base = base
        self.addr = addr

    def __getitem__(self, addr) -> "_Builder":
        return _Builder(self.base, self.addr + _norm_addr(addr))

    def set(self, v) -> ChoiceMap:
        _validate_dynamic(self.addr, allow_partial_slice=False)
        new = ChoiceMap.entry(v, *self.addr)
        if self.base is None:
            return new
        return new | self.base  # the new entry wins over what the map held at that address (`chm + old`, :776)

    def update(self, f: Callable) -> ChoiceMap:
        """Replace what sits at the address by ``f(value)`` (or ``f(submap)`` if there is no value there)."""
        if self.base is None:
            return self.set(f(_EMPTY))
        sub = self.base.get_submap(self.addr)
        return self.set(f(sub.get_value() if sub.has_value() else sub))

    def n(self) -> ChoiceMap:
        return ChoiceMap.empty()

    def v(self, v) -> ChoiceMap:
        # a ChoiceMap passed to .v is stored AS A VALUE (choice_map.py:813-817: "not advisable", but tested)
        return self.set(ChoiceMap(value=v, has_value=True) if isinstance(v, ChoiceMap) else ChoiceMap.choice(v))

    def from_mapping(self, mapping) -> ChoiceMap:
        return self.set(ChoiceMap.from_mapping(mapping))

    def d(self, d: dict) -> ChoiceMap:
        return self.set(ChoiceMap.d(d))

    def kw(self, **kwargs) -> ChoiceMap:
        return self.set(ChoiceMap.kw(**kwargs))

    def switch(self, idx, chms) -> ChoiceMap:
        return self.set(ChoiceMap.switch(idx, chms))


ChoiceMapBuilder = _Builder()
ChoiceMap.builder = ChoiceMapBuilder

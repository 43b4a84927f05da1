"""Selection / ChoiceMap behaviour the reference's own suite asserts (/root/reference/tests/core/test_choice_maps.py),
restated over torch leaves for the address forms this package supports (static components, `...` wildcards in
selections, full slices / static integer indices over vectorised leaves).  Each test names the reference test it
follows.  Array-valued (dynamic) indices, Switch and traced masks are out of scope and must raise."""
import pytest
import torch
from hypothesis import assume, given, settings
from hypothesis import strategies as st

import genjax_b200 as gj
from genjax_b200 import ChoiceMap, ChoiceMapNoValueAtAddress, Selection
from genjax_b200 import ChoiceMapBuilder as C
from genjax_b200 import SelectionBuilder as S

# ------------------------------------------------------------------ selections (TestSelections)


def test_selection_prefix_semantics():  # test_selection
    sel = S["x"] | S["z", "y"]
    assert sel["x"] and sel["z", "y"] and sel["z", "y", "tail"]
    sel = S["x"]
    assert sel["x"] and sel["x", "y"] and sel["x", "y", "z"]
    sel = S["x", "y", "z"]
    assert sel["x", "y", "z"] and not sel["x"] and not sel["x", "y"]


def test_wildcard_selection():  # test_wildcard_selection
    sel = S["x"] | S[..., "y"]
    assert sel["x"] and sel["any_address", "y"] and sel["rando", "y", "tail"]
    assert not sel["any_address", "z"]


def test_selection_all_none():  # test_selection_all / test_selection_none
    a, n = Selection.all(), Selection.none()
    assert a == ~~a and a["x"] and a["y", "z"] and a[()]
    assert n == ~~n and not n["x"] and not n["y", "z"] and not n[()]
    assert Selection.none().extend("a", "b") == Selection.none()


def test_selection_builder_properties():  # test_selection_builder_properties
    assert S.all == Selection.all() and S.all["x"] and S.all[()]
    assert S.none == Selection.none() and not S.none["x"] and not S.none[()]
    leaf = S.leaf
    assert leaf == Selection.leaf()
    leaf = leaf.extend("a", "b")
    assert leaf["a", "b"] and not leaf["a"] and not leaf["a", "b", "c"]
    assert S[()] == Selection.leaf()
    assert () in S[()]


def test_selection_leaf_rejects_wildcard_queries():  # test_selection_leaf / test_ellipsis_not_allowed
    leaf = Selection.leaf().extend("x", "y")
    assert not leaf["x"] and leaf["x", "y"] and not leaf["x", "y", "z"]
    with pytest.raises(TypeError):
        leaf[..., "y"]
    with pytest.raises(TypeError):
        (S["a", "b", "c"] | S["x", "y", "z"])["a", ..., ...]


def test_selection_complement_and_or_simplify():  # test_selection_complement / _and / _or
    sel1, sel2 = S["x"] | S["y"], S["y"] | S["z"]
    comp = ~sel1
    assert not comp["x"] and not comp["y"] and comp["z"]
    assert ~~sel1 == sel1
    assert ~Selection.all() == Selection.none() and ~Selection.none() == Selection.all()

    both = sel1 & sel2
    assert not both["x"] and both["y"] and not both["z"]
    assert not both.check() and both.get_subselection("y").check()
    assert (Selection.all() & sel1) == sel1 and (sel1 & Selection.all()) == sel1
    assert (Selection.none() & sel1) == Selection.none() and (sel1 & Selection.none()) == Selection.none()
    assert sel1 & sel1 == sel1

    either = S["x"] | S["y"]
    assert either["x"] and either["y"] and not either["z"] and either.get_subselection("y").check()
    assert (Selection.all() | sel1) == Selection.all() and (sel1 | Selection.all()) == Selection.all()
    assert (Selection.none() | sel1) == sel1 and (sel1 | Selection.none()) == sel1
    assert sel1 | sel1 == sel1

    combined = (sel1 & sel2) | S["w"]  # test_selection_combination
    assert not combined["x"] and combined["y"] and not combined["z"] and combined["w"]


def test_selection_filter():  # test_selection_filter
    chm = ChoiceMap.kw(x=1, y=2, z=3)
    kept = (S["x"] | S["y"]).filter(chm)
    assert "x" in kept and "y" in kept and "z" not in kept
    assert kept["x"] == 1 and kept["y"] == 2
    assert Selection.none().filter(chm).static_is_empty()
    assert Selection.all().filter(chm) == chm
    nested = ChoiceMap.kw(a={"b": 1, "c": 2}, d=3)
    kept = (S["a", "b"] | S["d"]).filter(nested)
    assert "d" in kept and "b" in kept("a") and "c" not in kept("a")


def test_selection_contains_and_call():  # test_selection_contains / test_static_sel
    sel = S["x"] | S["y", "z"]
    assert "x" in sel and ("y", "z") in sel and "y" not in sel and "w" not in sel
    nested = S["c"].extend("a", "b")
    assert ("a", "b", "c") in nested and ("a", "b") not in nested
    assert not nested("a")("b").check() and nested("a")("b")("c").check()
    xy = Selection.at["x", "y"]
    assert not xy[()] and xy["x", "y"] and not xy["other_address"]
    assert xy("x") == Selection.at["y"] and xy("z") == Selection.none()
    inner = Selection.at["x"].extend("y")
    assert inner["y", "x"] and not inner["y"]


def test_selection_of_a_choice_map():  # test_chm_sel
    chm = C["x", "y"].set(3.0) | C["z"].set(5.0)
    sel = chm.get_selection()
    assert sel["x", "y"] and sel["z"] and not sel["w"] and sel("x")["y"]
    assert not sel["z", "below_a_leaf"]
    assert ChoiceMap.empty().get_selection() == Selection.none()


# ------------------------------------------------------------------ builder (TestChoiceMapBuilder)


def test_builder_set_and_membership():  # test_set / test_nested_set
    assert ChoiceMap.builder.set(1.0) == C[()].set(1.0)
    chm = C["a", "b"].set(1)
    assert chm["a", "b"] == 1 and ("a", "b") in chm and "a" not in chm and "b" in chm("a")
    chm = C["x"].set(C["y"].set(2))
    assert chm["x", "y"] == 2 and ("x", "y") in chm and "y" not in chm


def test_builder_update():  # test_update
    chm = C["x", "y"].set(2)
    assert chm.at["x"].update(lambda m: C["z"].set(m))["x", "z", "y"] == 2
    assert chm.at["x", "y"].update(lambda v: v * v)["x", "y"] == 4
    assert chm.at["q"].update(lambda m: C["z"].set(m))(("q", "z")).static_is_empty()
    assert chm.at["q"].update(lambda m: C["z"].set(2))["q", "z"] == 2


def test_builder_n_v_d_kw_from_mapping():  # test_empty / test_v_matches_set / test_from_mapping / test_d / test_kw
    assert C.n() == ChoiceMap.empty() and C["x", "y"].n() == ChoiceMap.empty()
    assert C["a", "b"].set(1) == C["a", "b"].v(1)
    inner = C["y"].v(2)
    assert C["x"].v(inner)("x").get_value() == inner
    chm = C["base"].from_mapping([("a", 1.0), (("b", "c"), 2.0), (("b", "d", "e"), {"f": 3.0})])
    assert chm["base", "a"] == 1 and chm["base", "b", "c"] == 2 and chm["base", "b", "d", "e", "f"] == 3
    assert ("b", "c") in chm("base")
    chm = C["top"].d({"x": 3, "y": {"z": 4, "w": C["bottom"].d({"v": 5})}})
    assert chm["top", "x"] == 3 and chm["top", "y", "z"] == 4 and chm["top", "y", "w", "bottom", "v"] == 5
    chm = C["root"].kw(a=1, b=C["nested"].kw(c=2, d={"deep": 3}))
    assert chm["root", "a"] == 1 and chm["root", "b", "nested", "c"] == 2 and chm["root", "b", "nested", "d", "deep"] == 3


def test_switch_needs_a_concrete_index():  # test_switch (concrete part); traced index is out of scope
    a, b, c = C["x"].set(1), C["y"].set(2), C["z"].set(3)
    assert C["root"].switch(1, [a, b, c])("root") == b
    assert ChoiceMap.switch(1, [a, b, c]) == b
    assert C["root"].switch(0, [C.n(), C.n()]).static_is_empty()
    with pytest.raises(NotImplementedError):
        ChoiceMap.switch(torch.tensor(1), [a, b, c])


# ------------------------------------------------------------------ choice maps (TestChoiceMap)


def test_choice_and_empty():  # test_empty / test_choice
    assert ChoiceMap.empty().static_is_empty()
    choice = ChoiceMap.choice(42.0)
    assert choice.get_value() == 42.0 and choice.has_value() and () in choice
    assert ChoiceMap.choice(torch.ones(0)).static_is_empty()


def test_kw_d_from_mapping():  # test_kv / test_d / test_from_mapping
    chm = ChoiceMap.kw(x=1, y=2)
    assert chm["x"] == 1 and chm["y"] == 2 and "x" in chm and "other_value" not in chm
    chm = ChoiceMap.d({"a": 1, "b": {"c": 2, "d": {"e": 3}}})
    assert chm["a"] == 1 and chm["b", "c"] == 2 and chm["b", "d", "e"] == 3 and ("b", "d", "e") in chm
    chm = ChoiceMap.from_mapping([("x", 1), (("y", "z"), 2), (("w", "v", "u"), 3)])
    assert chm["x"] == 1 and chm["y", "z"] == 2 and chm["w", "v", "u"] == 3 and ("w", "v", "u") in chm


def test_extend_through_at():  # test_extend_through_at
    base = ChoiceMap.kw(x=1, y={"z": 2})
    ext = base.at["y", "w"].set(3)
    assert ext["x"] == 1 and ext["y", "z"] == 2 and ext["y", "w"] == 3
    multi = base.at["y", "w"].set(3).at["a", "b", "c"].set(4)
    assert multi["y", "w"] == 3 and multi["a", "b", "c"] == 4 and multi["x"] == 1
    over = base.at["y", "z"].set(5)
    assert over["x"] == 1 and over["y", "z"] == 5
    nested = base.at["nested"].set(ChoiceMap.kw(a=6, b=7))
    assert nested["nested", "a"] == 6 and nested["nested", "b"] == 7 and nested["y", "z"] == 2
    assert base["y", "z"] == 2 and "nested" not in base.keys()  # the original is untouched


def test_mask_extend_merge():  # test_mask / test_extend / test_merge / test_static_is_empty
    chm = ChoiceMap.kw(x=1, y=2)
    assert chm.mask(True) == chm and chm.mask(False).static_is_empty()
    # a flag ARRAY wraps every leaf in Mask(value, flag) (choice_map.py ``mask``; the form of a Mask-ed constraint)
    from genjax_b200 import Mask

    masked = chm.mask(torch.tensor(True))
    assert isinstance(masked["x"], Mask) and masked["x"].value == 1 and bool(masked["x"].flag)
    with pytest.raises(NotImplementedError):
        chm.mask("not a flag")
    ext = ChoiceMap.choice(1).extend("a", "b")
    assert ext["a", "b"] == 1 and ext.get_value() is None and ext.get_submap("a", "b").get_value() == 1
    assert ChoiceMap.empty().extend("a", "b").static_is_empty()
    a, b = ChoiceMap.kw(x=1), ChoiceMap.kw(y=2)
    merged = a.merge(b)
    assert merged["x"] == 1 and merged["y"] == 2 and merged == a | b
    assert not ChoiceMap.kw(x=1).static_is_empty()


def test_or_xor_and():  # test_or_xor_access / test_xor / test_or / test_and
    left, right = ChoiceMap.kw(x=1, y=2), ChoiceMap.kw(z=3, w=4)
    for both in (left | right, left ^ right):
        assert both["x"] == 1 and both["y"] == 2 and both["z"] == 3 and both["w"] == 4
        with pytest.raises(ChoiceMapNoValueAtAddress):
            both["does_not_exist"]
    assert (ChoiceMap.empty() ^ ChoiceMap.empty()).static_is_empty()
    a = ChoiceMap.kw(x=1)
    assert (a ^ ChoiceMap.empty()) == a and (ChoiceMap.empty() ^ a) == a
    assert (a | ChoiceMap.empty()) == a and (ChoiceMap.empty() | a) == a
    assert (a | ChoiceMap.kw(y=2)).get_value() is None
    assert (ChoiceMap.choice(2.0) | ChoiceMap.choice(3.0)).get_value() == 2.0  # the left operand wins
    with pytest.raises(Exception, match="Choice and non-Choice in Or"):
        _ = C["x"].set(1.0) | C["x", "y"].set(2.0)

    c1, c2 = ChoiceMap.kw(x=1, y=2, z=3), ChoiceMap.kw(y=20, z=30, w=40)
    both = c1 & c2
    assert "x" not in both and "w" not in both and both["y"] == 20 and both["z"] == 30
    assert (c1 & ChoiceMap.empty()).static_is_empty() and (ChoiceMap.empty() & c1).static_is_empty()
    n1, n2 = ChoiceMap.kw(a={"b": 1, "c": 2}, d=3), ChoiceMap.kw(a={"b": 10, "d": 20}, d=30)
    both = n1 & n2
    assert both["a", "b"] == 10 and "c" not in both("a") and "d" not in both("a") and both["d"] == 30


def test_call_getitem_contains():  # test_call / test_getitem / test_contains / test_get_selection
    chm = ChoiceMap.kw(x={"y": 1})
    assert chm("x")("y") == ChoiceMap.choice(1)
    assert "x" not in chm and "y" in chm("x") and ("x", "y") in chm and "z" not in chm
    with pytest.raises(ChoiceMapNoValueAtAddress, match="y"):
        ChoiceMap.kw(x=1)["y"]
    sel = ChoiceMap.kw(x=1, y=2).get_selection()
    assert sel["x"] and sel["y"] and not sel["z"]


def test_simplify_is_identity_and_filters_push_down():  # test_simplify
    xyz = ChoiceMap.d({"x": 1, "y": 2, "z": 3})
    either = xyz.filter(S["x"]) | xyz.filter(S["y"])
    assert either.simplify() == ChoiceMap.d({"x": 1, "y": 2})
    with pytest.raises(ChoiceMapNoValueAtAddress, match="z"):
        either["z"]
    assert C["x"].set(None).simplify() == C["x"].set(None)


# ------------------------------------------------------------------ vectorised leaves and index components


def test_lookup_on_a_vector_leaf():  # test_lookup_dynamic
    chm = ChoiceMap.choice(torch.tensor([2.3, 4.4, 3.3]))
    assert chm.get_submap("x").static_is_empty()
    assert [float(chm[i]) for i in range(3)] == pytest.approx([2.3, 4.4, 3.3])
    assert ChoiceMap.empty().extend(slice(None, None, None)).static_is_empty()


def test_filter_through_a_full_slice():  # test_choicemap_filter_with_wildcard
    xs, ys = torch.tensor([1.0, 2.0, 3.0]), torch.tensor([4.0, 5.0, 6.0])
    chm = C[:].set({"x": xs, "y": ys})
    kept = chm.filter(S["x"])
    assert torch.equal(kept[:, "x"], xs)
    with pytest.raises(ChoiceMapNoValueAtAddress):
        kept[:, "y"]
    assert [float(kept[i, "x"]) for i in range(3)] == [1.0, 2.0, 3.0]


def test_static_index_component():  # test_choicemap_with_static_idx / test_choicemap_slice_validation (scalar part)
    chm = C[0].set({"x": 1.0, "y": 2.0})
    assert chm[0, "x"] == 1.0 and chm[0, "y"] == 2.0
    with pytest.raises(ChoiceMapNoValueAtAddress):
        chm[1, "x"]
    chm = C[0, "x", 1].set(10)
    assert chm[0, "x", 1] == 10
    assert C[torch.tensor(2), "y"].set(20)[2, "y"] == 20  # a concrete 0-d index is a static index
    sel = chm.get_selection()  # integer layers are transparent to selections
    assert sel["x"] and chm.filter(S["x"]) == chm and chm.filter(S["y"]).static_is_empty()


def test_slices():  # test_choicemap_slice
    for bad in (slice(None, 3), slice(0, 3), slice(0, 3, 1)):
        with pytest.raises(ValueError):
            C[bad, "x"].set(torch.tensor([1, 2]))
    with pytest.raises(ValueError):
        C[0, "x", 1:3].set(torch.tensor([1, 2]))
    vals = torch.arange(10)
    chm = C[:, "x"].set(vals)
    assert torch.equal(chm[:, "x"], vals)
    assert chm[1, "x"] == vals[1]
    assert chm[torch.tensor(5), "x"] == vals[5]
    assert torch.equal(chm[0:4, "x"], vals[0:4])
    with pytest.raises(ValueError):
        chm[0:4, 0:2, "x"]  # at most one partial slice in a lookup


def test_array_valued_indices_are_out_of_scope():  # test_access_dynamic: dynamic structure, SURVEY 8f-3
    with pytest.raises(NotImplementedError):
        C[torch.tensor([4, 8, 2]), "x"].set(torch.tensor([4.0, 8.0, 2.0]))


# ------------------------------------------------------------------ validation against a model


def test_invalid_subset():  # test_choicemap_validation / test_choicemap_nested_validation
    @gj.gen
    def model(x):
        y = gj.normal(x, 1.0) @ "y"
        z = gj.bernoulli(probs=0.5) @ "z"
        return y + z

    assert ChoiceMap.kw(y=1.0, z=1).invalid_subset(model, (0.0,)) is None
    only_x = ChoiceMap.kw(x=1.0)
    assert only_x.invalid_subset(model, (0.0,)) == only_x
    assert ChoiceMap.kw(y=1.0, z=1, extra=0.5).invalid_subset(model, (0.0,)) == ChoiceMap.kw(extra=0.5)

    @gj.gen
    def inner_model():
        a = gj.normal(0.0, 1.0) @ "a"
        b = gj.bernoulli(probs=0.5) @ "b"
        return a + b

    @gj.gen
    def outer_model():
        x = gj.normal(0.0, 1.0) @ "x"
        y = inner_model() @ "y"
        return x + y

    assert ChoiceMap.kw(x=1.0, y=ChoiceMap.kw(a=0.5, b=1)).invalid_subset(outer_model, ()) is None
    assert ChoiceMap.kw(x=1.0, y=ChoiceMap.kw(a=0.5)).invalid_subset(outer_model, ()) is None  # missing is fine
    extra_inner = ChoiceMap.kw(x=1.0, y=ChoiceMap.kw(a=0.5, b=1, c=2.0))
    assert extra_inner.invalid_subset(outer_model, ()) == ChoiceMap.kw(y=ChoiceMap.kw(c=2.0))
    extra_outer = ChoiceMap.kw(x=1.0, y=ChoiceMap.kw(a=0.5, b=1), z=3.0)
    assert extra_outer.invalid_subset(outer_model, ()) == ChoiceMap.kw(z=3.0)


# ------------------------------------------------------------------ path splitting (TestSubmap)

_maps = st.deferred(
    lambda: st.dictionaries(
        st.text(max_size=4), st.floats(allow_nan=False) | st.lists(st.floats(allow_nan=False), max_size=3) | _maps,
        min_size=1, max_size=3)
)


def _paths(mapping):
    out, stack = [], [((), mapping)]
    while stack:
        prefix, m = stack.pop()
        if isinstance(m, dict) and m:
            stack.extend(((*prefix, k), v) for k, v in m.items())
        else:
            out.append((prefix, m))
    return out


@settings(max_examples=60, deadline=None)
@given(_maps, st.data())
def test_get_submap_split_and_splat(mapping, data):  # test_get_submap_split_path / test_path_can_be_splat
    chm = ChoiceMap.d(mapping)
    path, value = data.draw(st.sampled_from(_paths(mapping)))
    assume(path)
    i = data.draw(st.integers(0, len(path)))
    assert chm.get_submap(path[:i])[path[i:]] == value
    assert chm.get_submap(path[:i], path[i:]) == chm.get_submap(path)
    assert chm.get_submap(path) == chm.get_submap(*path)

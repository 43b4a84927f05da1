"""Host-side logic that needs no GPU: ChoiceMap / Selection algebra (after
/root/reference/tests/core/test_choice_maps.py), the key tree, model capture
(site discovery, error conventions) and code generation."""
import numpy as np
import pytest

import genjax_b200 as gj
from genjax_b200 import C, S, ChoiceMap, Selection
from genjax_b200.gen import capture as cap
from genjax_b200.gen import codegen
from genjax_b200.gen.capture import ArgSpec
from genjax_b200.gen.expr import TracedControlFlow


# ---------------------------------------------------------------- selections
def test_selection_basic():
    new = S["x"] | S["z", "y"]
    assert new["x"] and new["z", "y"] and new["z", "y", "tail"]
    new = S["x", "y", "z"]
    assert new["x", "y", "z"] and not new["x"] and not new["x", "y"]


def test_selection_all_none_complement():
    assert Selection.all()["x"] and Selection.all()["y", "z"]
    assert not Selection.none()["x"]
    sel = S["x"] | S["y"]
    comp = ~sel
    assert not comp["x"] and not comp["y"] and comp["z"]
    both = S["x"] & S["x"]
    assert both["x"] and not both["y"]
    assert (S["x"] & S["y"])["x"] is False


def test_selection_extend_and_descend():
    sel = Selection.all().extend("a", "b")
    assert sel["a", "b"] and sel["a", "b", "c"] and not sel["a"]
    assert sel("a")["b"]
    assert "x" in S["x"]


# --------------------------------------------------------------- choice maps
def test_choicemap_builder_forms():
    chm = C["x"].set(1.0)
    assert chm["x"] == 1.0 and "x" in chm and "y" not in chm
    chm = C["a", "b"].set(2.0)
    assert chm["a", "b"] == 2.0 and chm("a")["b"] == 2.0
    assert C.kw(x=1, y=2)["y"] == 2
    assert C.d({"x": 1, "y": {"z": 3}})["y", "z"] == 3
    assert C.n().static_is_empty()
    assert C.v(5.0).get_value() == 5.0
    assert ChoiceMap.empty().static_is_empty()
    assert ChoiceMap.choice(3).has_value()


def test_choicemap_at_set_and_merge():
    chm = ChoiceMap.empty().at["x"].set(1.0).at["y", "z"].set(2.0)
    assert chm["x"] == 1.0 and chm["y", "z"] == 2.0
    merged = C["x"].set(1.0) | C["y"].set(2.0)
    assert merged["x"] == 1.0 and merged["y"] == 2.0
    # left-biased merge like the reference's `|` (choice_map.py:1227-1299)
    assert (C["x"].set(1.0) | C["x"].set(5.0))["x"] == 1.0
    assert C["x"].set(1.0).merge(C["y"].set(2.0))["y"] == 2.0


def test_choicemap_missing_value_raises():
    with pytest.raises(gj.ChoiceMapNoValueAtAddress):
        C["x"].set(1.0)["y"]


def test_choicemap_filter_and_selection_roundtrip():
    chm = C.d({"x": 1.0, "y": 2.0, "z": {"w": 3.0}})
    f = chm.filter(S["x"] | S["z"])
    assert "x" in f and ("z", "w") in f and "y" not in f
    g = chm.filter(~chm.filter(S["x"]).get_selection())
    assert "x" not in g and "y" in g
    assert sorted(a for a, _ in chm.leaves()) == [("x",), ("y",), ("z", "w")]


def test_target_filter_to_unconstrained():
    from genjax_b200.workloads import beta_bernoulli

    t = gj.Target(beta_bernoulli, (2.0, 2.0), C["v"].set(True))
    latents = t.filter_to_unconstrained(C.d({"p": 0.3, "v": True}))
    assert "p" in latents and "v" not in latents
    assert t["v"] is True
    with pytest.raises(TypeError):
        gj.Target(beta_bernoulli, (2.0, 2.0), {"v": True})


# ------------------------------------------------------------------- capture
def _specs(*kinds):
    return [ArgSpec(k, "f32", s) for k, s in kinds]


def test_capture_sites_in_program_order():
    from genjax_b200.workloads import lgssm_step

    ir = cap.capture(lgssm_step.source, "lgssm", _specs(("particle", ())), ("tuple", [("leaf", 0)]))
    assert [s.addr for s in ir.sites] == [("x",), ("y",)]
    assert [s.dist.name for s in ir.sites] == ["normal", "normal"]
    assert ir.width == 0 and len(ir.ret_leaves) == 1
    assert ir.ret_leaves[0].op == "site" and ir.ret_leaves[0].attr == 0


def test_capture_address_reuse_raises():
    @gj.gen
    def bad():
        gj.normal(0.0, 1.0) @ "x"
        gj.normal(0.0, 1.0) @ "x"

    with pytest.raises(gj.AddressReuse):
        cap.capture(bad.source, "bad", [], ("tuple", []))


def test_capture_traced_control_flow_raises():
    @gj.gen
    def bad(x):
        if x > 0:
            return gj.normal(0.0, 1.0) @ "a"
        return gj.normal(1.0, 1.0) @ "b"

    with pytest.raises(TracedControlFlow):
        cap.capture(bad.source, "bad", _specs(("particle", ())), ("tuple", [("leaf", 0)]))


def test_capture_nested_gen_inlines_under_prefix():
    @gj.gen
    def inner(m):
        return gj.normal(m, 1.0) @ "z"

    @gj.gen
    def outer(m):
        a = inner(m) @ "sub"
        return gj.normal(a, 2.0) @ "y"

    ir = cap.capture(outer.source, "outer", _specs(("scalar", ())), ("tuple", [("leaf", 0)]))
    assert [s.addr for s in ir.sites] == [("sub", "z"), ("y",)]


def test_exact_density_is_not_fusable():
    d = gj.exact_density(lambda key, a: a, lambda v, a: 0.0, "mine")

    @gj.gen
    def m():
        return d(1.0) @ "x"

    with pytest.raises(gj.NotFusable):
        cap.capture(m.source, "m", [], ("tuple", []))


def test_fingerprint_is_structural():
    from genjax_b200.workloads import lgssm_step

    a = cap.capture(lgssm_step.source, "a", _specs(("particle", ())), ("tuple", [("leaf", 0)]))
    b = cap.capture(lgssm_step.source, "b", _specs(("particle", ())), ("tuple", [("leaf", 0)]))
    c = cap.capture(lgssm_step.source, "c", _specs(("scalar", ())), ("tuple", [("leaf", 0)]))
    assert cap.ir_fingerprint(a) == cap.ir_fingerprint(b) != cap.ir_fingerprint(c)


# ------------------------------------------------------------------- codegen
def test_codegen_mappings():
    assert codegen.group_lanes(32) == 8 and codegen.group_lanes(8) == 2 and codegen.group_lanes(4) == 1
    assert codegen.group_lanes(12) == 0 and codegen.group_lanes(0) == 0 and codegen.group_lanes(256) == 0
    from genjax_b200.workloads import lgssm_step, lgssm_step_vec

    ir = cap.capture(lgssm_step.source, "lgssm", _specs(("particle", ())), ("tuple", [("leaf", 0)]))
    ir.digest = cap.ir_fingerprint(ir)
    src = codegen.generate(ir)
    assert "quad mapping" in src and "gjb_model_launch" in src and "gjb::Normal::" in src
    ir = cap.capture(lgssm_step_vec.source, "v", _specs(("particle", (32,)), ("shared", (32,)), ("shared", (32,))),
                     ("tuple", [("leaf", 0), ("leaf", 1), ("leaf", 2)]))
    ir.digest = cap.ir_fingerprint(ir)
    src = codegen.generate(ir)
    assert "group mapping, G=8" in src


def test_new_model_compiles_for_sm100a():
    """A model outside the prebuilt set goes capture -> codegen -> nvcc (sm_100a) -> loadable .so."""

    @gj.gen
    def m(x_prev, rate):
        s = gj.exponential(rate) @ "s"
        u = gj.uniform(0.0, 1.0) @ "u"
        b = gj.flip(gj.numpy.sigmoid(x_prev)) @ "b"
        y = gj.normal(x_prev * s + u, 1.0 + gj.numpy.exp(-s)) @ "y"
        return gj.numpy.where(b, y, -y)

    cm = m.prebuild([ArgSpec("particle", "f32", ()), ArgSpec("scalar", "f32", ())])
    assert cm.path.exists() and cm.info["mapping"] == "quad"
    assert [s["dist"] for s in cm.info["sites"]] == ["exponential", "uniform", "flip", "normal"]


def test_partial_apply_keeps_addresses():
    @gj.gen
    def m(a, b):
        x = gj.normal(a, b) @ "x"
        return gj.normal(x, 1.0) @ "y"

    p = m.partial_apply(0.5)
    ir = cap.capture(p.source, "p", _specs(("scalar", ())), ("tuple", [("leaf", 0)]))
    assert [s.addr for s in ir.sites] == [("x",), ("y",)]
    with pytest.raises(NotImplementedError):
        gj.normal.partial_apply(0.0)


def test_target_and_choice_map_are_pytrees_for_capture():
    """A custom proposal is a @gen function OF THE TARGET (custom_proposal.ipynb): Target / ChoiceMap arguments are
    flattened into traced leaves and rebuilt symbolically inside the body."""
    @gj.gen
    def model():
        x = gj.normal(0.0, 1.0) @ "x"
        gj.normal(x, 1.0) @ "y"

    target = gj.Target(model, (), C["y"].set(1.5))
    leaves, tree = cap.flatten((target,))
    assert leaves == [1.5] and tree[0] == "tuple" and tree[1][0][0] == "target"
    rebuilt = cap.unflatten(tree, ["SYM"])[0]
    assert isinstance(rebuilt, gj.Target) and rebuilt["y"] == "SYM" and rebuilt.p is model

    @gj.gen
    def proposal(t):
        return gj.normal(t["y"] / 2.0, 0.7) @ "x"

    ir = cap.capture(proposal.source, "proposal", [ArgSpec("scalar", "f32", ())], tree)
    assert [s.addr for s in ir.sites] == [("x",)] and ir.sites[0].args[0].op == "div"


def test_key_children_batch_and_scalar():
    from genjax_b200.core.key import KeyBatch, key_children

    kb = gj.split(gj.key(1), 8)
    a, b = key_children(kb)
    assert isinstance(a, KeyBatch) and a.n == b.n == 8 and a.words != b.words != kb.words and a.offset == kb.offset
    ka, kb2 = key_children(gj.key(1))
    assert ka == gj.split(gj.key(1))[0] and kb2 == gj.split(gj.key(1))[1]


def test_chain_kernels_are_generated_and_hmc_degrades_to_e_mode_without_a_gradient():
    """MH / HMC kernels are generated per (model, latent set).  A model whose log-density has no device gradient
    (lgamma of a latent) still gets its MH kernel; the HMC entry point then reports GJB_E_MODE instead of a wrong
    answer.  Argument validation happens before any GPU work, so this runs on the CPU."""
    from genjax_b200.gen.codegen_chain import ChainSpec
    from genjax_b200.gen.static import compile_ir

    @gj.gen
    def ok_model():
        x = gj.normal(0.0, 1.0) @ "x"
        gj.normal(x, 0.5) @ "y"

    ir = cap.capture(ok_model.source, "ok_model", [], ("tuple", []))
    cm = compile_ir(ir, chain=ChainSpec((0,), (None,)))
    assert cm.lib.gjb_model_mh_chain(None, None) == -1 and cm.lib.gjb_model_hmc_chain(None, None) == -1  # GJB_E_ARG

    @gj.gen
    def no_grad_model():
        a = gj.exponential(1.0) @ "a"
        gj.gamma(a + 1.0, 2.0) @ "g"  # log-density contains lgamma(a + 1): no digamma on the device

    ir = cap.capture(no_grad_model.source, "no_grad_model", [], ("tuple", []))
    cm = compile_ir(ir, chain=ChainSpec((0,), (None,)))
    assert cm.lib.gjb_model_mh_chain(None, None) == -1   # MH kernel exists (argument check fires)
    assert cm.lib.gjb_model_hmc_chain(None, None) == -3  # GJB_E_MODE: no gradient kernel for this model
    # a chain over an integer-valued site is refused at generation time
    @gj.gen
    def discrete():
        gj.flip(0.3) @ "b"

    from genjax_b200.gen.autodiff import NotDifferentiable

    ir = cap.capture(discrete.source, "discrete", [], ("tuple", []))
    with pytest.raises(NotDifferentiable):
        compile_ir(ir, chain=ChainSpec((0,), (None,)))


def test_gen_transfers_function_metadata():
    """test_static_gen_fn.py:38-79 (TestStaticGenFnMetadata)."""
    import genjax_b200 as gj

    def original_function(x: float, y: float) -> float:
        """This is a test function that adds two numbers."""
        return x + y

    wrapped = gj.gen(original_function)
    assert wrapped.__doc__ == original_function.__doc__ and wrapped.__name__ == original_function.__name__
    assert wrapped.__module__ == original_function.__module__ and wrapped.__qualname__ == original_function.__qualname__
    assert wrapped.__wrapped__ is original_function
    assert wrapped.__annotations__ == {"x": float, "y": float, "return": float}
    assert gj.gen(original_function).partial_apply(1.0).partial_apply(2.0).partial_args == (1.0, 2.0)


def test_zero_trace_and_site_addresses_need_no_device():
    import genjax_b200 as gj

    @gj.gen
    def model(x):
        y = gj.normal(x, 1.0) @ "y"
        z = gj.bernoulli(probs=0.7) @ ("z", "inner")
        return y + z

    zt = model.get_zero_trace(0.0)
    assert zt.get_args() == (0.0,) and zt.get_retval() == 0.0 and zt.get_score() == 0.0
    assert zt.get_choices()["y"] == 0.0 and zt.get_choices()["z", "inner"] == 0
    assert model.get_site_addresses((0.0,)) == [("y",), ("z", "inner")]
    with pytest.raises(RuntimeError):
        model.inline(0.0)  # only inside an @gen body


def test_distributions_reject_what_they_cannot_honour():
    """Unknown keywords and sample_shape=n fail at capture time instead of being dropped silently."""
    import genjax_b200 as gj

    def addresses(make):
        @gj.gen
        def m():
            return make() @ "v"

        return m.get_site_addresses(())

    for make in (lambda: gj.normal(0.0, 1.0, sample_shape=(2, 2)), lambda: gj.bernoulli(probs=0.3, sample_shape=2),
                 lambda: gj.categorical(logits=[0.0, 0.0], sample_shape=3)):
        with pytest.raises(NotImplementedError, match="sample_shape"):
            addresses(make)
    with pytest.raises(TypeError, match="unexpected keyword"):
        addresses(lambda: gj.uniform(0.0, 1.0, foo=1))
    for make in (lambda: gj.normal(loc=0.0, scale=1.0), lambda: gj.categorical(probs=[0.5, 0.5]),
                 lambda: gj.uniform(low=0.0, high=2.0), lambda: gj.categorical(logits=[0.0, 0.0], sample_shape=())):
        assert addresses(make) == [("v",)]


def test_weight_upper_bounds_from_the_model():
    """gen/bounds.py: the analytic supremum of a filter step's incremental weight."""
    import math

    import torch

    import genjax_b200 as gj
    from genjax_b200.inference.pf import ParticleFilter
    from genjax_b200.workloads import LG_R, hmm_step, lgssm_step, lgssm_step_vec

    half = 0.5 * math.log(2 * math.pi)
    b = ParticleFilter(lgssm_step, 8).weight_upper_bound(torch.zeros(8), gj.C["y"].set(torch.zeros(3)))
    assert b == pytest.approx(-(half + math.log(LG_R)), rel=1e-6)
    r = torch.tensor([0.5, 1.0, 2.0, 4.0])
    b = ParticleFilter(lgssm_step_vec, 8).weight_upper_bound(torch.zeros(8, 4), gj.C["y"].set(torch.zeros(3, 4)),
                                                             (torch.ones(4), r))
    assert b == pytest.approx(-(4 * half + float(torch.log(r).sum())), rel=1e-6)
    b = ParticleFilter(hmm_step, 8).weight_upper_bound(torch.zeros(8, dtype=torch.int32), gj.C["y"].set(torch.zeros(3, dtype=torch.int32)),
                                                       (torch.zeros(16, 16), torch.zeros(16, 16)))
    assert b == 0.0

    @gj.gen
    def hetero(x_prev):  # the observation noise depends on the particle: no particle-free bound
        x = gj.normal(x_prev, 1.0) @ "x"
        gj.normal(x, gj.numpy.exp(0.1 * x)) @ "y"
        return x

    assert ParticleFilter(hetero, 8).weight_upper_bound(torch.zeros(8), gj.C["y"].set(torch.zeros(3))) is None

    @gj.gen
    def tails(x_prev, s):  # long-tail observation sites: the mode of each sits at its loc
        x = gj.normal(x_prev, 1.0) @ "x"
        gj.cauchy(x, s) @ "y"
        gj.laplace(x, 2.0) @ "y2"
        gj.gumbel(x, s) @ "y3"
        gj.half_cauchy(x, s) @ "y4"
        return x

    obs = gj.C["y"].set(torch.zeros(3)).at["y2"].set(torch.zeros(3)).at["y3"].set(torch.zeros(3)).at["y4"].set(torch.zeros(3))
    b = ParticleFilter(tails, 8).weight_upper_bound(torch.zeros(8), obs, (0.5,))
    ls = math.log(0.5)
    assert b == pytest.approx((-math.log(math.pi) - ls) + (-math.log(2.0) - math.log(2.0)) + (-1.0 - ls) + (math.log(2 / math.pi) - ls), rel=1e-6)

    # the bound really bounds: oracle weights of the scalar model never exceed it
    from oracle import gfi as ogfi
    from oracle import rng

    def o_step(h, x_prev):
        x = h.normal("x", np.float32(0.9) * x_prev, np.float32(1.0))
        h.normal("y", x, np.float32(LG_R))
        return x

    _, w = ogfi.generate(o_step, rng.split(rng.key(0), 4096), {"y": np.float32(0.3)}, (np.zeros(4096, dtype=np.float32),))
    bound = ParticleFilter(lgssm_step, 8).weight_upper_bound(torch.zeros(8), gj.C["y"].set(torch.zeros(3)))
    assert w.max() <= bound + 1e-6 and w.max() > bound - 1e-3

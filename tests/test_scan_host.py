"""Host logic of the Scan combinator (genjax_b200/gen/scan.py) on CPU.

The kernel launch is replaced by tests/abi_emulator.py (the oracle's samplers / log-densities behind the real
``gjb_model_args`` structure), so these tests check the HOST plumbing -- per-step arguments, key chain, constraint
slicing, stacking, update / regenerate bookkeeping -- against the independent oracle restatement of
combinators/scan.py (oracle/gfi.py ``scan_*``) and against the outcomes the reference's own tests assert
(/root/reference/tests/generative_functions/test_scan_combinator.py).  They say nothing about the CUDA kernels; the
same scenarios run on a GPU in tests/test_zzz_unverified_gpu.py."""
import numpy as np
import pytest
import torch

import abi_emulator
import genjax_b200 as gj
from genjax_b200 import ChoiceMapBuilder as C
from oracle import dists as od
from oracle import gfi as ogfi
from oracle import rng

F32 = np.float32


@pytest.fixture(params=["ir", "host"])
def emu(monkeypatch, request):
    """"ir": the captured IR interpreted with the oracle; "host": the generated CUDA source compiled for the host
    (tests/host_kernels.py) -- the same scenarios on both."""
    return abi_emulator.install(monkeypatch, host_kernels=request.param == "host")


@gj.gen
def walk(x, std):
    nx = gj.normal(x, std) @ "x"
    y = gj.normal(2.0 * nx, 0.5) @ "y"
    return nx, nx + y


def o_walk(h, x, std):
    nx = h.normal("x", x, std)
    y = h.normal("y", (F32(2.0) * nx).astype(F32), F32(0.5))
    return nx, (nx + y).astype(F32)


STDS = np.array([2.0, 4.0, 3.0, 5.0, 1.0], dtype=F32)


def _np(t):
    return t.detach().cpu().numpy()


def test_emulator_matches_oracle_on_a_plain_model(emu):
    n = 37
    tr, w = walk.importance(gj.split(gj.key(5), n), C["y"].set(1.5), (gj.Batched(torch.linspace(-1, 1, n)), 0.7))
    otr, ow = ogfi.importance(o_walk, rng.split(rng.key(5), n), {"y": F32(1.5)}, (np.linspace(-1, 1, n, dtype=F32), F32(0.7)))
    np.testing.assert_allclose(_np(tr.get_choices()["x"]), otr.choices["x"], rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(_np(w), ow, rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(_np(tr.get_score()), otr.get_score(), rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("n", [None, 1, 6])
def test_scan_simulate_matches_oracle(emu, n):
    key = gj.key(314159) if n is None else gj.split(gj.key(314159), n)
    okey = rng.key(314159) if n is None else rng.split(rng.key(314159), n)
    model = walk.scan(n=5)
    tr = model.simulate(key, (0.25, torch.tensor(STDS)))
    _, ocarry, oys, oscore = ogfi.scan_simulate(o_walk, okey, F32(0.25), STDS)
    lead = () if n is None else (n,)
    chm = tr.get_choices()
    assert tuple(chm[:, "x"].shape) == lead + (5,)
    carry, ys = tr.get_retval()
    np.testing.assert_allclose(_np(carry).reshape(-1), np.broadcast_to(ocarry, (n or 1,)), rtol=1e-6, atol=1e-6)
    oys = np.stack([np.broadcast_to(y, (n or 1,)) for y in oys], axis=1)
    np.testing.assert_allclose(_np(ys).reshape(oys.shape), oys, rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(_np(tr.get_score()).reshape(-1), oscore, rtol=1e-5, atol=2e-5)
    # test_scan_length_inferred: the traced values are the scanned outputs; n= is optional when xs has a leading axis
    tr2 = walk.scan().simulate(key, (0.25, torch.tensor(STDS)))
    assert torch.equal(tr2.get_choices()[:, "x"], chm[:, "x"])
    # project on everything is the score (test_iterate_simple_normal)
    torch.testing.assert_close(tr.project(gj.key(1), gj.Selection.all()), tr.get_score(), rtol=1e-5, atol=2e-5)
    # get_subtrace stacks per-step scores (tests/core/generative/test_core.py:151-158)
    sub = tr.get_subtrace("y").get_score()
    assert tuple(sub.shape) == lead + (5,)
    torch.testing.assert_close(sub.sum(-1) + tr.get_subtrace("x").get_score().sum(-1), tr.get_score(), rtol=1e-5, atol=3e-5)


def test_scalar_lane_equals_lane_of_batch(emu):
    """A batched scan over split(key, n) and a scalar scan with split(key, n)[i] agree on lane i."""
    n = 5
    kb = gj.split(gj.key(9), n)
    model = walk.scan(n=5)
    batch = model.simulate(kb, (0.0, torch.tensor(STDS)))
    one = model.simulate(kb[3], (0.0, torch.tensor(STDS)))
    assert torch.equal(batch.get_choices()[:, "x"][3], one.get_choices()[:, "x"])
    assert one.get_score().shape == () and batch.get_score().shape == (n,)


def test_vmap_key_scan_shapes(emu):  # test_vmap_key_scan
    @gj.gen
    def model(x, _):
        y = gj.normal(x, 1.0) @ "y"
        return y, None

    results = model.scan().simulate(gj.split(gj.key(314159), 10), (1.0, torch.arange(5, dtype=torch.float32)))
    assert results.get_score().shape == (10,)
    assert results.get_choices()[:, "y"].shape == (10, 5)
    assert results.get_retval()[1] is None


def test_scan_importance_with_sliced_and_indexed_constraints(emu):
    n = 4
    ys = np.array([0.5, -1.0, 2.0, 0.0, 1.0], dtype=F32)
    model = walk.scan()
    tr, w = model.importance(gj.split(gj.key(2), n), C[:, "y"].set(torch.tensor(ys)), (0.1, torch.tensor(STDS)))
    _, _, _, oscore, ow = ogfi.scan_generate(o_walk, rng.split(rng.key(2), n), lambda t: {"y": ys[t]}, F32(0.1), STDS)
    np.testing.assert_allclose(_np(w), ow, rtol=1e-5, atol=2e-5)
    np.testing.assert_allclose(_np(tr.get_score()), oscore, rtol=1e-5, atol=2e-5)
    assert torch.equal(tr.get_choices()[:, "y"], torch.tensor(ys).expand(n, 5))
    # the index layer is optional (test_choicemap_scan: "index layer isn't required")
    tr2, w2 = model.importance(gj.split(gj.key(2), n), C["y"].set(torch.tensor(ys)), (0.1, torch.tensor(STDS)))
    assert torch.equal(w2, w)
    # one step constrained through a static index (test_iterate_simple_normal_importance)
    tr3, w3 = model.importance(gj.key(4), C[2, "x"].set(0.5), (0.1, torch.tensor(STDS)))
    xs = tr3.get_choices()[:, "x"]
    assert xs[2] == 0.5
    expect = od.normal_logpdf(F32(0.5), F32(xs[1].item()), STDS[2])
    assert w3.item() == pytest.approx(float(expect), rel=1e-5, abs=1e-5)
    # per-particle constraints: leaves [n, T] marked by vmap's in_axes=0
    per = torch.tensor(np.stack([ys + i for i in range(n)]))
    tr4, w4 = gj.vmap(model.importance, in_axes=(0, 0, None))(gj.split(gj.key(2), n), C[:, "y"].set(per), (0.1, torch.tensor(STDS)))
    _, _, _, _, ow4 = ogfi.scan_generate(o_walk, rng.split(rng.key(2), n), lambda t: {"y": _np(per)[:, t]}, F32(0.1), STDS)
    np.testing.assert_allclose(_np(w4), ow4, rtol=1e-5, atol=2e-5)
    assert torch.equal(tr4.get_choices()[:, "y"], per)


def test_scan_assess_equals_importance_with_everything_constrained(emu):
    n = 3
    model = walk.scan()
    tr = model.simulate(gj.split(gj.key(7), n), (0.0, torch.tensor(STDS)))
    chm = gj.vmap(lambda c: c, in_axes=0)(tr.get_choices())  # mark the [n, T] leaves as per-particle
    score, (carry, ys) = model.assess(chm, (0.0, torch.tensor(STDS)))
    torch.testing.assert_close(score, tr.get_score(), rtol=1e-5, atol=2e-5)
    torch.testing.assert_close(ys, tr.get_retval()[1])
    _, w = model.importance(gj.split(gj.key(8), n), chm, (0.0, torch.tensor(STDS)))
    torch.testing.assert_close(w, score, rtol=1e-5, atol=2e-5)
    with pytest.raises(gj.MissingAddress):
        model.assess(C[:, "x"].set(torch.zeros(5)), (0.0, torch.tensor(STDS)))


def test_scan_update_and_regenerate(emu):
    n = 4
    model = walk.scan()
    args = (0.3, torch.tensor(STDS))
    tr = model.simulate(gj.split(gj.key(11), n), args)
    otr, _, _, _ = ogfi.scan_simulate(o_walk, rng.split(rng.key(11), n), F32(0.3), STDS)

    # update one step's x: the change propagates through the carry to the next step's score (test_scan_update)
    new, w, _, bwd = model.update(gj.split(gj.key(12), n), tr, C[1, "x"].set(9.0), gj.Diff.no_change(args))
    onew, _, _, oscore, ow, odisc = ogfi.scan_update(o_walk, rng.split(rng.key(12), n), otr,
                                                     lambda t: {"x": F32(9.0)} if t == 1 else {}, F32(0.3), STDS)
    np.testing.assert_allclose(_np(w), ow, rtol=1e-4, atol=2e-4)
    np.testing.assert_allclose(_np(new.get_score()), oscore, rtol=1e-5, atol=2e-5)
    xs_new, xs_old = new.get_choices()[:, "x"], tr.get_choices()[:, "x"]
    assert (xs_new[:, 1] == 9.0).all() and torch.equal(xs_new[:, [0, 2, 3, 4]], xs_old[:, [0, 2, 3, 4]])
    torch.testing.assert_close(w, new.get_score() - tr.get_score(), rtol=1e-4, atol=2e-4)
    assert torch.equal(bwd[1, "x"], xs_old[:, 1]) and sorted(odisc) == [1]
    # and back again: the weights cancel (regenerate / update identities of test_requests.py:52-61)
    back, wb, _, _ = model.update(gj.split(gj.key(13), n), new, bwd, gj.Diff.no_change(args))
    torch.testing.assert_close(w + wb, torch.zeros(n), rtol=0, atol=5e-4)
    assert torch.equal(back.get_choices()[:, "x"], xs_old)

    # regenerate x everywhere: fresh samples from the step keys, y kept
    reg, wr, _, _ = model.edit(gj.split(gj.key(14), n), tr, gj.Regenerate(gj.S["x"]), gj.Diff.no_change(args))
    oreg, _, _, osc, owr = ogfi.scan_regenerate(o_walk, rng.split(rng.key(14), n), otr, {"x"}, F32(0.3), STDS)
    np.testing.assert_allclose(_np(wr), owr, rtol=1e-4, atol=3e-4)
    np.testing.assert_allclose(_np(reg.get_choices()[:, "x"]), np.stack([t.choices["x"] for t in oreg], 1), rtol=1e-6, atol=1e-6)
    assert torch.equal(reg.get_choices()[:, "y"], tr.get_choices()[:, "y"])


def test_scalar_call_update_has_no_particle_axis(emu):
    model = walk.scan()
    args = (0.3, torch.tensor(STDS))
    tr = model.simulate(gj.key(11), args)
    new, w, _, bwd = model.update(gj.key(12), tr, C[1, "x"].set(9.0), gj.Diff.no_change(args))
    assert w.shape == () and new.get_choices()[:, "x"].shape == (5,) and new.get_choices()[1, "x"] == 9.0
    assert bwd[1, "x"].shape == () and bwd[1, "x"] == tr.get_choices()[1, "x"]
    back, wb, _, _ = model.update(gj.key(13), new, bwd, gj.Diff.no_change(args))
    assert torch.equal(back.get_choices()[:, "x"], tr.get_choices()[:, "x"]) and (w + wb).abs().item() < 5e-4


def test_iterate_accumulate_reduce(emu):
    """test_iterate_simple_normal_importance / test_iterate / test_accumulate / test_reduce, with one random choice per
    step (kernels without any choice have no fused kernel to launch)."""

    @gj.gen
    def step(x):
        return gj.normal(x, 1.0) @ "z"

    it = step.iterate(n=10)
    tr, w = it.importance(gj.key(314159), C[3, "z"].set(0.5), (0.01,))
    zs = tr.get_choices()[:, "z"]
    assert zs[3] == 0.5
    assert w.item() == pytest.approx(float(od.normal_logpdf(F32(0.5), F32(zs[2].item()), F32(1.0))), rel=1e-5, abs=1e-5)
    out = tr.get_retval()
    assert out.shape == (11,) and out[0].item() == pytest.approx(0.01) and torch.equal(out[1:], zs)
    new, _, _, _ = it.update(gj.key(1), tr, C[3, "z"].set(1.0), gj.Diff.no_change((0.01,)))
    assert new.get_choices()[3, "z"] == 1.0
    assert step.iterate_final(n=10).simulate(gj.key(314159), (0.01,)).get_retval().shape == ()

    @gj.gen
    def add(acc, x):
        return acc + x + 0.0 * (gj.normal(0.0, 1.0) @ "eps")

    res = add.accumulate().simulate(gj.key(0), (0.0, torch.ones(4))).get_retval()
    assert torch.equal(res, torch.tensor([0.0, 1.0, 2.0, 3.0, 4.0]))
    assert add.reduce().simulate(gj.key(0), (0.0, torch.ones(10))).get_retval().item() == 10.0
    res = add.accumulate().simulate(gj.split(gj.key(0), 3), (0.0, torch.ones(4))).get_retval()
    assert torch.equal(res, torch.tensor([0.0, 1.0, 2.0, 3.0, 4.0]).expand(3, 5))


def test_scan_validation_and_zero_length(emu):  # test_scan_validation / test_zero_length_scan
    @gj.gen
    def foo(shift, d):
        x = gj.normal(d["loc"], d["scale"]) @ "x"
        return x + shift, None

    with pytest.raises(ValueError, match="scan got values with different leading axis sizes: 2, 1."):
        foo.scan().simulate(gj.key(0), (1.0, {"loc": torch.tensor([10.0, 12.0]), "scale": torch.tensor([1.0])}))
    with pytest.raises(ValueError):
        walk.scan(n=3).simulate(gj.key(0), (0.0, torch.tensor(STDS)))
    empty = walk.scan(n=0).simulate(gj.key(0), (2.0, torch.zeros(0)))
    assert empty.get_choices().static_is_empty() and empty.get_score().item() == 0.0
    walk.scan().importance(gj.key(1), empty.get_choices(), (2.0, torch.zeros(0)))
    with pytest.raises(TypeError):
        gj.Scan(gj.normal)  # only @gen kernels can be scanned


def test_numpy_api_composites_inside_a_model(emu):
    """gj.numpy composites (clip, logaddexp, mean, dot, logical ops) are traced into the model and agree with NumPy."""
    jnp = gj.numpy

    @gj.gen
    def model(x, w):
        a = gj.normal(jnp.clip(x, -1.0, 1.0), 1.0) @ "a"
        b = gj.normal(jnp.logaddexp(a, x), jnp.reciprocal(2.0 + jnp.square(x))) @ "b"
        inside = jnp.logical_and(a > -5.0, jnp.logical_not(a > 5.0))
        return jnp.where(inside, jnp.dot(w, w) * a + jnp.mean(w), jnp.negative(b)), jnp.divide(jnp.subtract(b, a), jnp.add(1.0, jnp.abs(x)))

    n = 64
    xs = torch.linspace(-3, 3, n)
    w = torch.tensor([0.5, -1.0, 2.0, 0.25])
    tr = model.simulate(gj.split(gj.key(3), n), (gj.Batched(xs), w))
    a, b = _np(tr.get_choices()["a"]), _np(tr.get_choices()["b"])
    x = xs.numpy()
    r0, r1 = tr.get_retval()
    np.testing.assert_allclose(_np(r0), float((w * w).sum()) * a + float(w.mean()), rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(_np(r1), (b - a) / (1 + np.abs(x)), rtol=1e-5, atol=1e-5)
    want = od.normal_logpdf(a, np.clip(x, -1, 1).astype(F32), F32(1.0)) + od.normal_logpdf(
        b, np.logaddexp(a, x).astype(F32), (1.0 / (2.0 + x * x)).astype(F32))
    np.testing.assert_allclose(_np(tr.get_score()), want, rtol=2e-5, atol=2e-5)
    assert float(jnp.logaddexp(0.0, 0.0)) == pytest.approx(np.log(2.0)) and float(jnp.clip(torch.tensor(3.0), 0.0, 1.0)) == 1.0

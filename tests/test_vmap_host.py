"""Host logic of the Vmap / repeat combinators (genjax_b200/gen/vmap_combinator.py) on CPU, through
tests/abi_emulator.py, following /root/reference/tests/generative_functions/test_vmap_combinator.py and
test_repeat_combinator.py.  Says nothing about the CUDA kernels (GPU twin: tests/test_zzz_unverified_gpu.py)."""
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
def kernel(x):
    z = gj.normal(x, 1.0) @ "z"
    return z


def o_kernel(h, x):
    return h.normal("z", x, F32(1.0))


def _lp(v, mu):
    return float(od.normal_logpdf(F32(v), F32(mu), F32(1.0)))


def test_vmap_simulate_score_project_assess(emu):  # test_vmap_combinator_simple_normal / _project / _assess
    model = gj.vmap(in_axes=(0,))(kernel)
    over = torch.arange(0, 50, dtype=torch.float32)
    tr = model.simulate(gj.key(314159), (over,))
    otr = ogfi.simulate(o_kernel, rng.split(rng.key(314159), 50), (np.arange(50, dtype=F32),))
    np.testing.assert_allclose(tr.get_choices()[:, "z"].numpy(), otr.choices["z"], rtol=1e-6, atol=1e-6)
    assert tr.get_score().shape == () and torch.equal(tr.get_score(), tr.inner.get_score().sum())
    assert tr.get_score().item() == pytest.approx(float(otr.get_score().sum()), rel=1e-5)
    assert torch.equal(tr.get_retval(), tr.get_choices()[:, "z"])
    assert tr.project(gj.key(1), gj.Selection.all()).item() == pytest.approx(tr.get_score().item(), rel=1e-6)
    assert tr.project(gj.key(1), gj.Selection.none()).item() == 0.0
    score, ret = model.assess(tr.get_choices(), (over,))
    assert score.item() == pytest.approx(tr.get_score().item(), rel=1e-6) and torch.equal(ret, tr.get_retval())
    assert kernel.vmap(in_axes=(0,)).simulate(gj.key(314159), (over,)).get_score() == tr.get_score()
    assert tr.get_subtrace("z").get_score().shape == (50,)  # tests/core/generative/test_core.py:141-149


def test_vmap_importance_vector_and_indexed_constraints(emu):
    """test_vmap_combinator_vector_choice_map_importance / _indexed_choice_map_importance."""
    model = kernel.vmap(in_axes=(0,))
    over = torch.arange(0, 3, dtype=torch.float32)
    _, w = model.importance(gj.key(314159), C[:, "z"].set(torch.tensor([3.0, 2.0, 3.0])), (over,))
    assert w.item() == pytest.approx(_lp(3.0, 0.0) + _lp(2.0, 1.0) + _lp(3.0, 2.0), rel=1e-6)
    tr, w = model.importance(gj.key(1), C[0, "z"].set(3.0), (over,))  # only index 0 is constrained
    assert w.item() == pytest.approx(_lp(3.0, 0.0), rel=1e-6)
    zs = tr.get_choices()[:, "z"]
    assert zs[0] == 3.0
    # the unconstrained lanes keep the values their lane keys give without any constraint
    free = model.simulate(gj.key(1), (over,)).get_choices()[:, "z"]
    assert torch.equal(zs[1:], free[1:])
    zv = [3.0, -1.0, 2.0]
    chm = C[0, "z"].set(zv[0]) | C[1, "z"].set(zv[1]) | C[2, "z"].set(zv[2])
    tr, w = model.importance(gj.key(2), chm, (over,))
    assert [tr.get_choices()[i, "z"].item() for i in range(3)] == zv
    assert w.item() == pytest.approx(sum(_lp(v, m) for v, m in zip(zv, [0.0, 1.0, 2.0])), rel=1e-6)
    tr, w = model.importance(gj.key(3), C[1, "z"].set(5.0), (over,))  # a run in the middle: three launches
    assert tr.get_choices()[1, "z"] == 5.0 and w.item() == pytest.approx(_lp(5.0, 1.0), rel=1e-6)
    assert tr.get_score().item() == pytest.approx(sum(_lp(tr.get_choices()[i, "z"].item(), i) for i in range(3)), rel=1e-5)


def test_vmap_update_and_regenerate(emu):
    model = kernel.vmap(in_axes=(0,))
    over = torch.arange(0, 4, dtype=torch.float32)
    tr = model.simulate(gj.key(5), (over,))
    old = tr.get_choices()[:, "z"]
    new, w, _, bwd = model.update(gj.key(6), tr, C[2, "z"].set(1.0), gj.Diff.no_change((over,)))
    assert new.get_choices()[2, "z"] == 1.0 and torch.equal(new.get_choices()[:, "z"][[0, 1, 3]], old[[0, 1, 3]])
    assert w.item() == pytest.approx(_lp(1.0, 2.0) - _lp(old[2].item(), 2.0), rel=1e-4, abs=1e-5)
    assert bwd[2, "z"] == old[2]
    allnew, w2, _, bwd2 = model.update(gj.key(7), tr, C[:, "z"].set(torch.zeros(4)), gj.Diff.no_change((over,)))
    assert torch.equal(bwd2[:, "z"], old) and (allnew.get_choices()[:, "z"] == 0).all()
    assert w2.item() == pytest.approx(allnew.get_score().item() - tr.get_score().item(), rel=1e-4, abs=1e-4)
    # a trace that was itself assembled from several runs can be updated again
    part, _ = model.importance(gj.key(9), C[1, "z"].set(5.0), (over,))
    again, w4, _, bwd4 = model.update(gj.key(10), part, C[3, "z"].set(-1.0), gj.Diff.no_change((over,)))
    assert again.get_choices()[1, "z"] == 5.0 and again.get_choices()[3, "z"] == -1.0
    assert bwd4[3, "z"] == part.get_choices()[3, "z"]
    assert w4.item() == pytest.approx(_lp(-1.0, 3.0) - _lp(part.get_choices()[3, "z"].item(), 3.0), rel=1e-4, abs=1e-5)
    reg, w3, _, _ = model.edit(gj.key(8), tr, gj.Regenerate(gj.S["z"]), gj.Diff.no_change((over,)))
    assert not torch.equal(reg.get_choices()[:, "z"], old)
    assert w3.item() == pytest.approx(reg.get_score().item() - tr.get_score().item(), rel=1e-4, abs=1e-4)


def test_repeat_and_in_axes_trees(emu):  # test_repeat_combinator.py / test_vmap_combinator_vmap_pytree / core test_repeat
    @gj.gen
    def model(x):
        return gj.normal(x, 1.0) @ "x"

    rep = model.repeat(n=3).simulate(gj.key(314159), (0.0,))
    vm = model.vmap().simulate(gj.key(314159), (torch.zeros(3),))
    assert torch.equal(rep.get_choices()[:, "x"], rep.get_retval()) and rep.get_retval().shape == (3,)
    assert torch.equal(vm.get_choices()[:, "x"], rep.get_choices()[:, "x"])  # same lane keys, same arguments
    assert gj.repeat(n=3)(model).simulate(gj.key(314159), (0.0,)).get_score() == rep.get_score()

    @gj.vmap(in_axes=(None, (0, None)))
    @gj.gen
    def foo(y, args):
        loc, (scale, _) = args
        x = gj.normal(loc, scale) @ "x"
        return x + y

    tr = foo.simulate(gj.key(0), (10.0, (torch.arange(3.0), (1.0, torch.arange(3)))))
    assert tr.get_retval().shape == (3,)
    torch.testing.assert_close(tr.get_retval(), tr.get_choices()[:, "x"] + 10.0)


def test_vmap_validation(emu):  # test_vmap_validation
    @gj.gen
    def foo(loc, scale):
        return gj.normal(loc, scale) @ "x"

    with pytest.raises(ValueError, match="vmap was requested to map its argument along axis 0, which implies that its "
                                         "rank should be at least 1, but is only 0"):
        foo.vmap(in_axes=(0, None)).simulate(gj.key(0), (10.0, torch.arange(3.0)))
    with pytest.raises(ValueError, match="vmap in_axes specification must be a tree prefix of the corresponding value"):
        foo.vmap(in_axes=(0, (0, None))).simulate(gj.key(0), (10.0, torch.arange(3.0)))
    with pytest.raises(IndexError):
        foo.vmap(in_axes=0).simulate(gj.key(0), (torch.arange(2.0), torch.arange(3.0)))
    # under an outer particle batch (a KeyBatch) the mapped axis is unrolled into the kernel: [particles, mapped] leaves
    tr = foo.vmap(in_axes=(0, None)).simulate(gj.split(gj.key(0), 4), (torch.arange(2.0), 1.0))
    assert tuple(tr.get_choices()[:, "x"].shape) == (4, 2) and tuple(tr.get_retval().shape) == (4, 2)


def test_closure_call_zero_length_and_repeat_importance(emu):
    """test_repeat_combinator_importance / test_repeat_matches_vmap (with one choice per call) / test_zero_length_vmap."""
    @gj.gen
    def model():
        return gj.normal(0.0, 1.0) @ "x"

    tr, w = model.repeat(n=10).importance(gj.key(314), C[1, "x"].set(3.0), ())
    assert w.item() == pytest.approx(_lp(tr.get_choices()[1, "x"].item(), 0.0), rel=1e-6) and tr.get_choices()[1, "x"] == 3.0

    @gj.gen
    def noisy_square(x):
        return x * x + 0.0 * (gj.normal(0.0, 1.0) @ "eps")

    rep = noisy_square.repeat(n=10)(2.0)(gj.key(314))
    assert rep.shape == (10,) and torch.equal(rep, torch.full((10,), 4.0))
    assert torch.equal(noisy_square.vmap()(torch.full((10,), 2.0))(gj.key(314)), rep)

    @gj.gen
    def step(state, sigma):
        new_x = gj.normal(state, sigma) @ "x"
        return new_x, new_x + 1

    empty = step.vmap(in_axes=(None, 0)).simulate(gj.key(20), (2.0, torch.zeros(0)))
    assert empty.get_choices().static_is_empty() and empty.get_score().item() == 0.0


def test_index_request_on_vmap_and_scan(emu):
    """TestVmapIndexRequest / TestScanIndexRequest (top-level form): the sub-request touches one index only."""
    @gj.gen
    def cell(mu):
        return gj.normal(mu, 1.0) @ "z"

    model = cell.vmap()
    over = torch.zeros(6)
    tr = model.simulate(gj.key(314159), (over,))
    old = tr.get_choices()[:, "z"]
    for idx in range(6):
        new, w, _, bwd = model.edit(gj.key(7), tr, gj.IndexRequest(idx, gj.Regenerate(gj.S["z"])), gj.Diff.no_change((over,)))
        z = new.get_choices()[:, "z"]
        keep = [i for i in range(6) if i != idx]
        assert z[idx] != old[idx] and torch.equal(z[keep], old[keep])
        assert w.item() == pytest.approx(_lp(z[idx].item(), 0.0) - _lp(old[idx].item(), 0.0), rel=1e-4, abs=1e-5)
        assert isinstance(bwd, gj.IndexRequest) and bwd.index == idx
        back, wb, _, _ = model.edit(gj.key(8), new, bwd, gj.Diff.no_change((over,)))
        assert torch.equal(back.get_choices()[:, "z"], old) and (w + wb).item() == pytest.approx(0.0, abs=1e-5)
        upd, wu, _, _ = model.edit(gj.key(9), tr, gj.IndexRequest(idx, gj.Update(C["z"].set(idx + 7.0))), gj.Diff.no_change((over,)))
        assert upd.get_choices()[idx, "z"] == idx + 7.0
        assert wu.item() == pytest.approx(_lp(idx + 7.0, 0.0) - _lp(old[idx].item(), 0.0), rel=1e-4, abs=1e-4)
    with pytest.raises(AssertionError):
        model.edit(gj.key(7), tr, gj.IndexRequest(11, gj.Regenerate(gj.S["z"])), gj.Diff.no_change((over,)))

    @gj.gen
    def kernel(carry, _):
        z = gj.normal(0.0, 1.0) @ "z"
        return z, None

    chain = kernel.scan(n=10)
    tr = chain.simulate(gj.key(1), (0.0, None))
    old = tr.get_choices()[:, "z"]
    for idx in (0, 4, 9):
        new, w, _, _ = chain.edit(gj.key(2), tr, gj.IndexRequest(idx, gj.Regenerate(gj.S["z"])), gj.Diff.no_change((0.0, None)))
        z = new.get_choices()[:, "z"]
        keep = [i for i in range(10) if i != idx]
        assert z[idx] != old[idx] and torch.equal(z[keep], old[keep])
        assert w.item() == pytest.approx(_lp(z[idx].item(), 0.0) - _lp(old[idx].item(), 0.0), rel=1e-4, abs=1e-5)
        assert new.get_score().item() == pytest.approx(sum(_lp(v.item(), 0.0) for v in z), rel=1e-5)
    with pytest.raises(AssertionError):
        chain.edit(gj.key(2), tr, gj.IndexRequest(11, gj.Regenerate(gj.S["z"])), gj.Diff.no_change((0.0, None)))

    @gj.gen
    def walk(carry, _):
        z = gj.normal(carry, 1.0) @ "z"
        return z + carry, None  # the outgoing carry reads the incoming one: every later step would move

    wtr = walk.scan(n=4).simulate(gj.key(3), (0.0, None))
    with pytest.raises(AssertionError):
        walk.scan(n=4).edit(gj.key(4), wtr, gj.IndexRequest(1, gj.Regenerate(gj.S["z"])), gj.Diff.no_change((0.0, None)))

"""Scan inside an @gen body (unrolled into the caller's fused kernel) and HMC over the choices of every step of a
scanned trace (reference: combinators/scan.py:81-99, tests/inference/test_requests.py:237-255)."""
import numpy as np
import pytest
import torch

from oracle import dists as od
from oracle import gfi as ogfi
from oracle import mcmc as omcmc
from oracle import rng as orng

pytestmark = pytest.mark.gpu
F32 = np.float32


def _gj():
    import genjax_b200 as gj

    return gj


def _np(t):
    return t.detach().cpu().numpy()


def _nested():
    gj = _gj()

    @gj.gen
    def step(c, x):
        z = gj.normal(c, 1.0) @ "z"
        y = gj.normal(z + x, 0.5) @ "y"
        return z, y

    @gj.gen
    def model(x0, xs):
        init = gj.normal(x0, 1.0) @ "init"
        final, ys = step.scan(n=4)(init, xs) @ "tracks"
        return final + ys[0]

    def o_model(h, x0, xs):
        c = h.normal("init", x0, F32(1.0))
        ys = []
        for t in range(4):
            z = h.normal(("tracks", t, "z"), c, F32(1.0))
            y = h.normal(("tracks", t, "y"), (z + xs[t]).astype(F32), F32(0.5))
            c = z
            ys.append(y)
        return (c + ys[0]).astype(F32)

    return model, o_model


def test_scan_nested_in_gen_matches_oracle(device):
    gj = _gj()
    model, o_model = _nested()
    n = 10_001
    xs = torch.tensor([0.5, -0.5, 1.0, 2.0])
    kb, okb = gj.split(gj.key(31), n), orng.split(orng.key(31), n)
    tr = model.simulate(kb, (0.25, xs))
    otr = ogfi.simulate(o_model, okb, (F32(0.25), xs.numpy()))
    chm = tr.get_choices()
    z = _np(chm["tracks", :, "z"])
    assert z.shape == (n, 4)
    for t in range(4):
        np.testing.assert_allclose(z[:, t], otr.choices[("tracks", t, "z")], rtol=3e-5, atol=3e-5)
        np.testing.assert_allclose(_np(chm["tracks", :, "y"])[:, t], otr.choices[("tracks", t, "y")], rtol=3e-5, atol=3e-5)
    np.testing.assert_allclose(_np(tr.get_score()), otr.get_score(), rtol=5e-5, atol=5e-5)
    np.testing.assert_allclose(_np(tr.get_retval()), otr.get_retval(), rtol=3e-5, atol=3e-5)

    # observations over the whole time axis, addressed like a scanned trace
    ys = torch.tensor([1.0, 0.0, 2.0, 2.5])
    tr2, w = model.importance(kb, gj.C["tracks", :, "y"].set(ys), (0.25, xs))
    otr2, ow = ogfi.generate(o_model, okb, {("tracks", t, "y"): F32(ys[t].item()) for t in range(4)}, (F32(0.25), xs.numpy()))
    np.testing.assert_allclose(_np(w), ow, rtol=5e-5, atol=5e-5)
    np.testing.assert_allclose(_np(tr2.get_choices()["tracks", :, "y"]), np.broadcast_to(ys.numpy(), (n, 4)))

    # a Selection names the address without the step index: regenerate every step's "z"
    tr3, w3, _, disc = tr2.edit(gj.split(gj.key(32), n), gj.Regenerate(gj.S["tracks", "z"]))
    otr3, ow3, _ = ogfi.regenerate(o_model, orng.split(orng.key(32), n), otr2, [("tracks", t, "z") for t in range(4)])
    np.testing.assert_allclose(_np(w3), ow3, rtol=1e-4, atol=1e-4)
    np.testing.assert_array_equal(_np(tr3.get_choices()["init"]), _np(tr2.get_choices()["init"]))
    assert (_np(tr3.get_choices()["tracks", :, "z"]) != _np(tr2.get_choices()["tracks", :, "z"])).all()


def test_iterate_and_accumulate_nested(device):
    """``f.iterate_final(n=)`` / ``f.accumulate()`` inside a body: return values of the unrolled forms."""
    gj = _gj()

    @gj.gen
    def inc(c):
        d = gj.uniform(0.0, 1.0) @ "d"
        return c + d

    @gj.gen
    def add(c, x):
        e = gj.normal(0.0, 0.1) @ "e"
        return c + x + e

    @gj.gen
    def model(xs):
        a = inc.iterate_final(n=3)(1.0) @ "it"
        acc = add.accumulate()(a, xs) @ "acc"
        return a, acc[0], acc[3]

    xs = torch.tensor([1.0, 2.0, 3.0])
    tr = model.simulate(gj.split(gj.key(1), 257), (xs,))
    a, first, last = (_np(v) for v in tr.get_retval())
    d = _np(tr.get_choices()["it", :, "d"])
    e = _np(tr.get_choices()["acc", :, "e"])
    np.testing.assert_allclose(a, 1.0 + d.sum(1), rtol=1e-6)
    np.testing.assert_array_equal(first, a)
    np.testing.assert_allclose(last, a + 6.0 + e.sum(1), rtol=1e-5, atol=1e-5)


def test_simple_scan_hmc(device):
    """tests/inference/test_requests.py:237-255: 50 bare HMC edits over ``Selection.at["x"]`` of a length-10 scan pull
    every x_t to its observation; one edit agrees with the oracle's restatement of hmc.py:156-211 on the joint target."""
    gj = _gj()
    from genjax_b200.inference.requests import HMC

    @gj.gen
    def kernel(z, scanned_in):
        z = gj.normal(z, 1.0) @ "x"
        _ = gj.normal(z, 0.01) @ "y"
        return z, None

    model = kernel.scan(n=10)
    n = 256
    vchm = gj.ChoiceMap.empty().at["y"].set(3.0 * torch.ones(10))
    tr, _ = model.importance(gj.split(gj.key(0), n), vchm, (0.0, None))
    request = HMC(gj.Selection.at["x"], 1e-2)

    x0 = _np(tr.get_choices()[:, "x"]).astype(F32)
    new_tr, w, _, _ = request.edit(gj.split(gj.key(5), n), tr, gj.Diff.no_change((0.0, None)))

    def lpg(q):
        q = q.astype(F32)
        prev = np.concatenate([np.zeros((q.shape[0], 1), dtype=F32), q[:, :-1]], axis=1)
        lp = np.zeros(q.shape[0], dtype=F32)
        for t in range(10):
            lp = (lp + od.normal_logpdf(q[:, t], prev[:, t], F32(1.0))).astype(F32)
            lp = (lp + od.normal_logpdf(F32(3.0), q[:, t], F32(0.01))).astype(F32)
        nxt = np.concatenate([q[:, 1:], q[:, -1:]], axis=1)
        g = -(q - prev) + (F32(3.0) - q) / F32(1e-4)
        g[:, :-1] += (nxt[:, :-1] - q[:, :-1])
        return lp, g.astype(F32)

    oq, olp, _, oalpha = omcmc.hmc_chain(lpg, x0, orng.split(orng.key(5), n), 1, 1e-2, 10, compat_stale_grad=True, accept=False)
    x1 = _np(new_tr.get_choices()[:, "x"])
    np.testing.assert_allclose(x1, oq, rtol=5e-4, atol=5e-4)
    sc = np.maximum(1.0, np.abs(olp))
    assert np.max(np.abs(_np(w) - oalpha) / sc) < 5e-3
    np.testing.assert_allclose(_np(new_tr.get_score()), olp, rtol=2e-3, atol=1.0)

    cur = tr
    key = gj.key(9)
    for i in range(50):
        cur, *_ = request.edit(gj.split(gj.fold_in(key, i), n), cur, gj.Diff.no_change((0.0, None)))
    x = _np(cur.get_choices()[:, "x"])
    assert x.shape == (n, 10)
    assert np.abs(x.mean(0) - 3.0).max() < 8e-3 * 3.0


def test_vmap_nested_in_gen_under_a_particle_batch(device):
    """``f.vmap(in_axes=...)(...) @ addr`` inside a body: the mapped axis is unrolled into the caller's kernel, which is
    also how a vmapped function runs under an outer particle batch (vmap.py:180-218).  Plain @gen callee, a masked
    callee over a vector of flags (test_mask_combinator.py:105-131, 160-176) and a masked distribution."""
    gj = _gj()

    @gj.gen
    def elem(m, s):
        z = gj.normal(m, s) @ "z"
        return z * 2.0

    @gj.gen
    def init():
        return gj.normal(0.0, 1.0) @ "x"

    masks = torch.tensor([True, False, True])

    @gj.gen
    def model(locs, x):
        zs = elem.vmap(in_axes=(0, None))(locs, 0.5) @ "elems"
        vm = init.mask().vmap(in_axes=(0,))(masks) @ "init"
        rats = gj.normal.mask().vmap(in_axes=(0, None, None))(masks, x, 1.0) @ "rats"
        return zs, vm, rats

    n = 5003
    locs = torch.tensor([-1.0, 0.0, 1.0, 2.0])
    tr = model.simulate(gj.split(gj.key(51), n), (locs, 0.25))
    zs, vm, rats = tr.get_retval()
    chm = tr.get_choices()
    z = _np(chm["elems", :, "z"])
    assert z.shape == (n, 4) and tuple(zs.shape) == (n, 4)
    np.testing.assert_allclose(_np(zs), 2.0 * z, rtol=1e-6)
    np.testing.assert_allclose(z.mean(0), locs.numpy(), atol=0.05)
    assert isinstance(vm, gj.Mask) and tuple(vm.value.shape) == (n, 3)
    np.testing.assert_array_equal(_np(vm.flag), np.broadcast_to(masks.numpy(), (n, 3)))
    xi = chm["init", :, "x"]
    assert isinstance(xi, gj.Mask) and tuple(xi.value.shape) == (n, 3)
    np.testing.assert_array_equal(_np(xi.value), _np(vm.value))
    r = chm["rats"]
    assert isinstance(r, gj.Mask) and tuple(r.value.shape) == (n, 3) and isinstance(rats, gj.Mask)
    mk = masks.numpy().astype(F32)
    want = (od.normal_logpdf(z, locs.numpy()[None, :], F32(0.5)).sum(1)
            + (mk[None, :] * od.normal_logpdf(_np(xi.value), F32(0.0), F32(1.0))).sum(1)
            + (mk[None, :] * od.normal_logpdf(_np(r.value), F32(0.25), F32(1.0))).sum(1))
    np.testing.assert_allclose(_np(tr.get_score()), want, rtol=1e-4, atol=1e-4)
    # the masked elements are drawn all the same (mask.py:158-165: the callee always runs)
    assert (np.abs(_np(xi.value)[:, 1]) > 0).all()

    # constraints over the mapped axis, per element and as a whole
    cons = gj.C["elems", :, "z"].set(torch.tensor([0.5, 0.5, 0.5, 0.5])) | gj.C["rats", 0].set(1.5)
    tr2, w = model.importance(gj.split(gj.key(52), n), cons, (locs, 0.25))
    want_w = float(od.normal_logpdf(F32(0.5), locs.numpy(), F32(0.5)).sum() + od.normal_logpdf(F32(1.5), F32(0.25), F32(1.0)))
    np.testing.assert_allclose(_np(w), want_w, rtol=1e-5)
    assert (_np(tr2.get_choices()["rats"].value)[:, 0] == 1.5).all()

    # top level: a vector of flags through mask().vmap() (test_mask_combinator.py:238-244)
    tr3 = init.mask().vmap().simulate(gj.key(1), (masks,))
    assert bool((tr3.get_retval().flag == masks.to(tr3.get_retval().flag.device)).all())


def test_top_level_vmap_under_an_outer_particle_batch(device):
    """``model.vmap(in_axes=...)`` called with a KeyBatch (``jax.vmap`` over split keys of a vmapped generative function,
    tests/generative_functions/test_vmap_combinator.py:208-228): two batch axes, the inner one unrolled into the kernel.
    simulate / importance / update (vmap.py:237-275) against the closed forms."""
    gj = _gj()

    @gj.gen
    def elem(m, s):
        z = gj.normal(m, s) @ "z"
        return z + 1.0

    vm = elem.vmap(in_axes=(0, None))
    n, k = 3001, 5
    locs = torch.linspace(-2.0, 2.0, k)
    kb = gj.split(gj.key(61), n)
    tr = vm.simulate(kb, (locs, 0.5))
    z = _np(tr.get_choices()[:, "z"])
    assert z.shape == (n, k)
    np.testing.assert_allclose(_np(tr.get_retval()), z + 1.0, rtol=1e-6)
    np.testing.assert_allclose(_np(tr.get_score()), od.normal_logpdf(z, locs.numpy()[None, :], F32(0.5)).sum(1), rtol=1e-4, atol=1e-4)
    assert abs(z.mean(0) - locs.numpy()).max() < 0.05

    obs = torch.tensor([0.1, 0.2, 0.3, 0.4, 0.5])
    tr2, w = vm.importance(kb, gj.C[:, "z"].set(obs), (locs, 0.5))
    want = float(od.normal_logpdf(obs.numpy(), locs.numpy(), F32(0.5)).sum())
    np.testing.assert_allclose(_np(w), want, rtol=1e-5)
    # update one element of every particle's vector (vmap.py:237-275 under the outer batch)
    tr3, w3, _, disc = tr.update(gj.split(gj.key(62), n), gj.C[2, "z"].set(0.0))
    z3 = _np(tr3.get_choices()[:, "z"])
    assert (z3[:, 2] == 0).all() and (np.delete(z3, 2, 1) == np.delete(z, 2, 1)).all()
    dw = od.normal_logpdf(F32(0.0), F32(locs[2].item()), F32(0.5)) - od.normal_logpdf(z[:, 2], F32(locs[2].item()), F32(0.5))
    np.testing.assert_allclose(_np(w3), dw, rtol=2e-4, atol=2e-4)
    np.testing.assert_array_equal(_np(disc[2, "z"]), z[:, 2])

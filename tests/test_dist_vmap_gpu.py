"""``dist.repeat(n=)`` / ``dist.vmap(in_axes=)`` for the scalar primitives (combinators/repeat.py:25-41,
vmap.py:180-218): one vector site of N independent draws, against the oracle's restatement (oracle/dists.py ``repeated``)
on the same Philox lanes, and through closed forms."""
import math

import numpy as np
import pytest
import torch

from oracle import dists as od
from oracle import gfi as ogfi
from oracle import rng as orng

pytestmark = pytest.mark.gpu
F32 = np.float32


def _gj():
    import genjax_b200 as gj

    return gj


def _np(t):
    return t.detach().cpu().numpy()


CASES = [
    ("uniform", (-1.0, 2.5)),
    ("exponential", (1.7,)),
    ("half_normal", (0.8,)),
    ("laplace", (0.3, 1.2)),
    ("cauchy", (0.0, 0.7)),
    ("log_normal", (0.1, 0.4)),
    ("gumbel", (1.0, 2.0)),
    ("logit_normal", (0.2, 0.9)),
    ("weibull", (1.5, 2.0)),
    ("half_cauchy", (0.5, 1.5)),
    ("kumaraswamy", (2.0, 3.0)),
    ("geometric", (0.3,)),
    ("normal", (0.5, 1.5)),
]


@pytest.mark.parametrize("name,args", CASES)
def test_repeat_matches_oracle(device, name, args):
    """``dist.repeat(n=6)`` inside an @gen body: values and score against the oracle; the score is the sum of the
    scalar primitive's log-densities (combinators/repeat.py:37-41)."""
    gj = _gj()
    dist = getattr(gj, name)
    D = 6

    @gj.gen
    def model():
        v = dist.repeat(n=D)(*args) @ "v"
        return v

    n = 4099
    tr = model.simulate(gj.split(gj.key(13), n), ())
    v = _np(tr.get_choices()["v"])
    assert v.shape == (n, D)
    sample, logpdf = od.repeated(name, D)
    words, idx = orng.lanes(orng.split(orng.key(13), n))
    ov = sample(words, idx, 1, *[F32(a) for a in args])
    close = np.isclose(v, ov, rtol=2e-4, atol=2e-5)
    assert close.mean() > 0.999, (name, (~close).sum())
    lp = logpdf(v, *[F32(a) for a in args])
    np.testing.assert_allclose(_np(tr.get_score()), lp, rtol=1e-4, atol=1e-4)
    # sum of the scalar primitive's own log-densities, element by element
    ref = sum(od.DISTS[name][1](v[:, k], *[F32(a) for a in args]).astype(np.float64) for k in range(D))
    np.testing.assert_allclose(_np(tr.get_score()), ref, rtol=1e-4, atol=1e-4)
    np.testing.assert_array_equal(_np(tr.get_retval()), v)
    # the draws are independent across elements
    if name not in ("cauchy", "half_cauchy"):
        c = np.corrcoef(v[:, 0], v[:, 1])[0, 1]
        assert abs(c) < 0.06


def test_vmap_in_axes_mapped_and_shared_arguments(device):
    """``exponential.vmap(in_axes=(0,))(rates)`` and ``laplace.vmap(in_axes=(0, None))(locs, scale)``: mapped arguments
    zip with the elements, shared ones broadcast (vmap.py:384; in_axes as jax.vmap's)."""
    gj = _gj()
    rates = torch.tensor([0.5, 1.0, 2.0, 4.0, 8.0])

    @gj.gen
    def model(locs, scale):
        e = gj.exponential.vmap(in_axes=(0,))(rates) @ "e"
        l = gj.laplace.vmap(in_axes=(0, None))(locs, scale) @ "l"
        return e, l

    n = 50_000
    locs = torch.tensor([-2.0, -1.0, 0.0, 1.0, 2.0])
    tr = model.simulate(gj.split(gj.key(2), n), (locs, 0.5))
    e, l = (_np(x) for x in tr.get_retval())
    np.testing.assert_allclose(e.mean(0), 1.0 / rates.numpy(), rtol=0.03)
    np.testing.assert_allclose(np.median(l, 0), locs.numpy(), atol=0.02)
    lp = od.repeated("exponential", 5)[1](e, rates.numpy()) + od.repeated("laplace", 5)[1](l, locs.numpy(), F32(0.5))
    np.testing.assert_allclose(_np(tr.get_score()), lp, rtol=1e-4, atol=1e-4)

    # importance with the vector site constrained: weight = its summed log-density
    obs = torch.tensor([0.1, 0.2, 0.3, 0.4, 0.5])
    tr2, w = model.importance(gj.split(gj.key(3), 64), gj.C["e"].set(obs), (locs, 0.5))
    want = float(np.sum(np.log(rates.numpy()) - rates.numpy() * obs.numpy()))
    np.testing.assert_allclose(_np(w), want, rtol=1e-5)

    with pytest.raises(ValueError):
        @gj.gen
        def bad():
            return gj.laplace.vmap(in_axes=(0,))(locs, 1.0) @ "x"

        bad.simulate(gj.key(0), ())


def test_repeat_rejects_what_it_cannot_map(device):
    gj = _gj()
    with pytest.raises(gj.NotFusable):
        gj.gamma.repeat(n=3)
    with pytest.raises(gj.NotFusable):
        gj.flip.vmap()


def test_repeat_under_switch_and_hmc_gradient(device):
    """A repeated site composes with the rest: under a Switch branch (zeros / no score where unselected) and as an
    HMC target (symbolic log-density of the vector site)."""
    gj = _gj()
    jnp = gj.numpy

    @gj.gen
    def a():
        return gj.uniform.repeat(n=4)(0.0, 2.0) @ "u"

    @gj.gen
    def b():
        return gj.exponential.repeat(n=4)(3.0) @ "u"

    @gj.gen
    def model():
        k = gj.flip(0.5) @ "k"
        return a.switch(b)(jnp.int32(k), (), ()) @ "s"

    n = 8192
    tr = model.simulate(gj.split(gj.key(4), n), ())
    k = _np(tr.get_choices()["k"]).astype(bool)
    u = _np(tr.get_choices()["s", "u"].unmask())
    lp = np.where(k, 4 * math.log(3.0) - 3.0 * u.sum(1), -4 * math.log(2.0)) + math.log(0.5)
    np.testing.assert_allclose(_np(tr.get_score()), lp, rtol=1e-4, atol=1e-4)
    assert (u[~k] <= 2.0).all() and abs(u[k].mean() - 1 / 3) < 0.02

    @gj.gen
    def target():
        z = gj.laplace.repeat(n=4)(1.0, 0.5) @ "z"
        return z

    from genjax_b200.inference.mcmc import HMC

    tr = target.simulate(gj.split(gj.key(6), 2048), ())
    tr2, w, _, _ = HMC(gj.S["z"], 0.05).edit(gj.split(gj.key(7), 2048), tr, ())
    z2 = _np(tr2.get_choices()["z"])
    want = (-np.abs(z2 - 1.0) / 0.5 - math.log(2 * 0.5)).sum(1)
    np.testing.assert_allclose(_np(tr2.get_score()), want, rtol=1e-4, atol=1e-4)

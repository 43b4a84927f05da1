"""mv_normal (full covariance; SURVEY kernel K10) on the fused path: Cholesky once per thread per launch, forward
substitution per particle; parity with the oracle's float32 restatement on the same Philox lanes."""
import numpy as np
import pytest
import torch

from oracle import dists as od
from oracle import rng as orng

pytestmark = pytest.mark.gpu
F32 = np.float32


@pytest.mark.parametrize("d", [3, 8])
def test_mv_normal_simulate_importance_assess(device, d):
    import genjax_b200 as gj

    @gj.gen
    def model(mu, cov):
        x = gj.mv_normal(mu, cov) @ "x"
        return x

    g = np.random.default_rng(d)
    A = g.standard_normal((d, d))
    cov = (A @ A.T + d * np.eye(d)).astype(F32)
    mu = g.standard_normal(d).astype(F32)
    args = (torch.from_numpy(mu), torch.from_numpy(cov))
    n = 20_000
    kb = gj.split(gj.key(3), n)
    tr = model.simulate(kb, args)
    x = tr.get_choices()["x"].cpu().numpy()
    words, idx = orng.lanes(orng.split(orng.key(3), n))
    ox = od.mv_normal_sample(words, idx, 1, mu, cov)
    np.testing.assert_allclose(x, ox, rtol=2e-5, atol=2e-5)
    np.testing.assert_allclose(tr.get_score().cpu().numpy(), od.mv_normal_logpdf(x, mu, cov), rtol=2e-5, atol=5e-5)
    # importance with the value constrained per particle: weight == logpdf; assess agrees
    v = torch.from_numpy((g.standard_normal((n, d)) * 2).astype(F32))
    chm = gj.vmap(lambda t: gj.C["x"].set(t), in_axes=0)(v)
    tr2, w = model.importance(kb, chm, args)
    want = od.mv_normal_logpdf(v.numpy(), mu, cov)
    np.testing.assert_allclose(w.cpu().numpy(), want, rtol=2e-5, atol=5e-5)
    score, _ = model.assess(chm, args)
    torch.testing.assert_close(score, tr2.get_score())
    assert abs(np.cov(x.T) - cov).max() < 0.05 * np.abs(cov).max() + 0.3

"""GPU tests of host-side features written after the round's 180 GPU-minutes were spent.  They compose entry points
the verified suite already exercises (gjb_model_launch through StaticGenerativeFunction._run), but have never run on
a device, so they carry the `unverified` marker and are skipped unless GJB_RUN_UNVERIFIED=1."""
import numpy as np
import pytest
import torch

from oracle import dists as od

pytestmark = [pytest.mark.gpu, pytest.mark.unverified]


def _gj():
    import genjax_b200 as gj

    return gj


def test_get_subtrace_scores_and_project(device):
    """tests/core/generative/test_core.py:27-37, 54-74, 77-113 (tupled addresses, project, nested get_subtrace)."""
    gj = _gj()

    @gj.gen
    def f():
        x = gj.normal(0.0, 1.0) @ "x"
        y = gj.normal(x, 2.0) @ "y"
        return x, y

    @gj.gen
    def g():
        x, y = f() @ "f"
        z = gj.normal(x + y, 1.0) @ ("z", "z0")
        return z

    n = 1000
    tr = f.simulate(gj.split(gj.key(0), n), ())
    xs, ys = tr.get_choices()["x"], tr.get_choices()["y"]
    sx, sy = tr.get_subtrace("x"), tr.get_subtrace("y")
    np.testing.assert_allclose(sx.get_score().cpu().numpy(), od.normal_logpdf(xs.cpu().numpy(), 0.0, 1.0), rtol=1e-5, atol=2e-5)
    np.testing.assert_allclose(sy.get_score().cpu().numpy(), od.normal_logpdf(ys.cpu().numpy(), xs.cpu().numpy(), 2.0), rtol=1e-5, atol=2e-5)
    assert torch.equal(sx.get_retval(), xs) and torch.equal(sx.get_choices().get_value(), xs)
    assert torch.equal(tr.project(gj.key(1), gj.S["x"]), sx.get_score())
    torch.testing.assert_close(tr.get_score(), sx.get_score() + sy.get_score(), rtol=1e-5, atol=2e-5)
    assert sx.get_gen_fn() is gj.normal

    tg = g.simulate(gj.split(gj.key(1), n), ())
    ftr = tg.get_subtrace("f")
    assert torch.equal(tg.get_subtrace("f", "x").get_score(), ftr.get_subtrace("x").get_score())
    torch.testing.assert_close(ftr.get_score(), ftr.get_subtrace("x").get_score() + ftr.get_subtrace("y").get_score())
    assert "x" in ftr.get_choices() and "y" in ftr.get_choices()
    zs = tg.get_subtrace("z", "z0")
    assert torch.equal(zs.get_score(), tg.project(gj.key(2), gj.Selection.at["z", "z0"]))
    with pytest.raises(gj.ChoiceMapNoValueAtAddress):
        tg.get_subtrace("nope")

    one = f.simulate(gj.key(3), ())  # scalar trace: 0-d views
    assert one.get_subtrace("x").get_score().shape == ()

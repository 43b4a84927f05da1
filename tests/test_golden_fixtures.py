"""Committed oracle fixtures (tests/golden/oracle_fixtures.npz, made by tests/golden/make_oracle_fixtures.py):
the oracle must keep reproducing them (integers exactly, floats to 1e-6), and the CUDA path is compared against the
frozen numbers without importing the oracle."""
import importlib.util
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
FIX = np.load(os.path.join(HERE, "golden", "oracle_fixtures.npz"))


def _build():
    spec = importlib.util.spec_from_file_location("make_oracle_fixtures", os.path.join(HERE, "golden", "make_oracle_fixtures.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.build()


def test_oracle_reproduces_committed_fixtures():
    now = _build()
    assert set(now) == set(FIX.files)
    for k in FIX.files:
        a, b = now[k], FIX[k]
        assert a.shape == b.shape and a.dtype == b.dtype, k
        if a.dtype.kind in "iub":
            assert np.array_equal(a, b), k
        else:
            np.testing.assert_allclose(a, b, rtol=1e-6, atol=1e-6, err_msg=k)


@pytest.mark.gpu
def test_cuda_matches_committed_fixtures(device):
    import torch

    import genjax_b200 as gj
    from genjax_b200.inference.pf import ParticleFilter
    from genjax_b200.runtime import smc_ops
    from genjax_b200.workloads import lgssm_step

    words = (0x12345678, 0x9ABCDEF0)
    w = smc_ops.philox_words(words, 1000, 3, 1, 64, device).cpu().numpy().view(np.uint32)
    assert np.array_equal(w, FIX["philox_words"])
    z = smc_ops.normal_fill(words, 1000, 2, 64, 8, device).cpu().numpy()
    np.testing.assert_allclose(z, FIX["normal_vec"], rtol=1e-5, atol=2e-6)
    # resampling of the frozen weights: bit-exact ancestors and integer mass
    lw = torch.from_numpy(FIX["resample_logw"]).to(device)
    ws = smc_ops.WeightWorkspace(lw.numel(), device)
    terms = ws.lse_terms(lw).cpu().numpy()
    assert terms[0] == FIX["lse_M_S"][0] and terms[1] == FIX["lse_M_S"][1]
    anc = torch.empty(lw.numel(), dtype=torch.int32, device=device)
    ws.systematic(lw, gj.key(5), anc)
    assert np.array_equal(anc.cpu().numpy(), FIX["resample_systematic"])
    # the 3-step filter: states / log-weights within fp32 tolerance, ancestors given the CUDA weights
    res = ParticleFilter(lgssm_step, 512, mode="graph").run(gj.key(99), torch.from_numpy(FIX["pf_x0"]),
                                               gj.C["y"].set(torch.from_numpy(FIX["pf_ys"])), record=True)
    np.testing.assert_allclose(res.history["log_weights"][0].cpu().numpy(), FIX["pf_logw"][0], rtol=1e-5, atol=2e-5)
    np.testing.assert_allclose(res.history["state"][0][0].cpu().numpy(), FIX["pf_states"][0], rtol=1e-5, atol=2e-6)
    assert res.log_increments[0].item() == pytest.approx(FIX["pf_logz_inc"][0], abs=1e-5)
    same = (res.ancestors[0].cpu().numpy() == FIX["pf_ancestors"][0]).mean()
    assert same > 0.99  # a 1-ulp weight difference can move a count boundary by one slot


# ------------------------------------------------------------------ combinators and the output-slot resampler

FIX2 = np.load(os.path.join(HERE, "golden", "combinator_fixtures.npz"))


def test_oracle_reproduces_combinator_fixtures():
    spec = importlib.util.spec_from_file_location("make_combinator_fixtures",
                                                  os.path.join(HERE, "golden", "make_combinator_fixtures.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    now = mod.build()
    assert set(now) == set(FIX2.files)
    for k in FIX2.files:
        a, b = now[k], FIX2[k]
        assert a.shape == b.shape and a.dtype == b.dtype, k
        if a.dtype.kind in "iub":
            assert np.array_equal(a, b), k
        else:
            np.testing.assert_allclose(a, b, rtol=1e-6, atol=1e-6, err_msg=k)
    # the push form (what the kernels compute today) gives the same ancestors as the frozen pull form
    from oracle import rng, smc

    key = rng.split(rng.key(11))[1]
    assert np.array_equal(smc.resample_systematic(FIX2["pull_logw"], key), FIX2["pull_ancestors_true_max"])


@pytest.mark.gpu
@pytest.mark.unverified
def test_cuda_combinators_match_committed_fixtures(device):
    import torch

    import genjax_b200 as gj

    @gj.gen
    def walk(x, std):
        nx = gj.normal(x, std) @ "x"
        y = gj.normal(2.0 * nx, 0.5) @ "y"
        return nx, nx + y

    @gj.gen
    def cell(x):
        return gj.normal(x, 1.0) @ "z"

    stds = torch.from_numpy(FIX2["scan_stds"]).to(device)
    tr = walk.scan(n=5).simulate(gj.split(gj.key(314159), 6), (0.25, stds))
    np.testing.assert_allclose(tr.get_choices()[:, "x"].cpu().numpy(), FIX2["scan_x"], rtol=1e-5, atol=2e-5)
    np.testing.assert_allclose(tr.get_choices()[:, "y"].cpu().numpy(), FIX2["scan_y"], rtol=1e-5, atol=4e-5)
    np.testing.assert_allclose(tr.get_score().cpu().numpy(), FIX2["scan_score"], rtol=1e-5, atol=5e-5)
    np.testing.assert_allclose(tr.get_retval()[0].cpu().numpy(), FIX2["scan_carry"], rtol=1e-5, atol=2e-5)
    yobs = torch.from_numpy(FIX2["scan_yobs"]).to(device)
    tr, w = walk.scan().importance(gj.split(gj.key(2), 6), gj.C[:, "y"].set(yobs), (0.1, stds))
    np.testing.assert_allclose(tr.get_choices()[:, "x"].cpu().numpy(), FIX2["scan_imp_x"], rtol=1e-5, atol=2e-5)
    np.testing.assert_allclose(w.cpu().numpy(), FIX2["scan_imp_weight"], rtol=1e-5, atol=5e-5)
    vt = cell.vmap().simulate(gj.key(314159), (torch.arange(50, dtype=torch.float32, device=device),))
    np.testing.assert_allclose(vt.get_choices()[:, "z"].cpu().numpy(), FIX2["vmap_z"], rtol=1e-5, atol=2e-6)
    np.testing.assert_allclose(vt.inner.get_score().cpu().numpy(), FIX2["vmap_score_lanes"], rtol=1e-5, atol=2e-5)

"""GPU parity of the fused model kernel and the particle-filter loop against
the oracle (teacher-forced per step) and the exact Kalman filter."""

import math

import numpy as np
import pytest
import torch

from oracle import dists as od
from oracle import gfi as ogfi
from oracle import rng as orng
from oracle import smc as osmc

pytestmark = pytest.mark.gpu

A_, Q_, C_, R_ = 0.9, 1.0, 1.0, 0.5


def _models():
    import genjax_b200 as gj

    @gj.gen
    def step(x_prev):
        x = gj.normal(A_ * x_prev, Q_) @ "x"
        gj.normal(C_ * x, R_) @ "y"
        return x

    @gj.gen
    def step_vec(x_prev, q, r):
        x = gj.mv_normal_diag(A_ * x_prev, q) @ "x"
        gj.mv_normal_diag(C_ * x, r) @ "y"
        return x

    return gj, step, step_vec


def o_step(h, x_prev):
    x = h.normal("x", np.float32(A_) * x_prev, np.float32(Q_))
    h.normal("y", np.float32(C_) * x, np.float32(R_))
    return x


def o_step_vec(h, x_prev, q, r):
    x = h.mv_normal_diag("x", np.float32(A_) * x_prev, q)
    h.mv_normal_diag("y", np.float32(C_) * x, r)
    return x


@pytest.mark.parametrize("n", [1, 5, 4096, 100_001])
def test_step_importance_matches_oracle(device, n):
    gj, step, _ = _models()
    key = gj.key(314159)
    keys = gj.split(key, n)
    g = np.random.default_rng(0)
    x_prev = g.standard_normal(n).astype(np.float32)
    tr, w = gj.vmap(step.importance, in_axes=(0, None, (0,)))(
        keys, gj.C["y"].set(0.7), (torch.from_numpy(x_prev).to(device),)
    )
    okeys = orng.split(orng.key(314159), n)
    otr, ow = ogfi.generate(o_step, okeys, {"y": np.float32(0.7)}, (x_prev,))
    np.testing.assert_allclose(tr.get_choices()["x"].cpu().numpy(), otr.choices["x"], rtol=1e-5, atol=2e-6)
    np.testing.assert_allclose(w.cpu().numpy(), ow, rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(tr.get_score().cpu().numpy(), otr.get_score(), rtol=1e-5, atol=1e-5)
    assert torch.equal(tr.get_retval(), tr.get_choices()["x"])


def test_assess_kat(device):
    """tests/generative_functions/test_static_gen_fn.py:317-318 of the reference."""
    import genjax_b200 as gj

    @gj.gen
    def model():
        y1 = gj.normal(0.0, 1.0) @ "y1"
        gj.normal(0.0, 1.0) @ "y2"
        return 0.0

    score, ret = model.assess(gj.C["y1"].set(1.0).at["y2"].set(-1.0), ())
    assert score.item() == pytest.approx(-2.837877, abs=1e-6)
    assert ret == 0.0


@pytest.mark.parametrize("d", [8, 32])
def test_step_vec_importance_matches_oracle(device, d):
    gj, _, step_vec = _models()
    n = 3001
    keys = gj.split(gj.key(7), n)
    g = np.random.default_rng(1)
    x_prev = g.standard_normal((n, d)).astype(np.float32)
    q = np.full(d, Q_, dtype=np.float32)
    r = (0.5 + 0.01 * np.arange(d)).astype(np.float32)
    y = g.standard_normal(d).astype(np.float32)
    tr, w = gj.vmap(step_vec.importance, in_axes=(0, None, (0, None, None)))(
        keys, gj.C["y"].set(torch.from_numpy(y)), (torch.from_numpy(x_prev).to(device), torch.from_numpy(q), torch.from_numpy(r))
    )
    otr, ow = ogfi.generate(o_step_vec, orng.split(orng.key(7), n), {"y": y}, (x_prev, q, r))
    np.testing.assert_allclose(tr.get_choices()["x"].cpu().numpy(), otr.choices["x"], rtol=1e-5, atol=2e-6)
    np.testing.assert_allclose(w.cpu().numpy(), ow, rtol=2e-5, atol=2e-4)
    np.testing.assert_allclose(tr.get_score().cpu().numpy(), otr.get_score(), rtol=2e-5, atol=2e-4)


@pytest.mark.parametrize("mode,use_graph", [("persistent", False), ("graph", False), ("graph", True)])
def test_particle_filter_teacher_forced_vs_oracle(device, mode, use_graph):
    """Per step: CUDA log-weights == oracle log-weights (fp32 tolerance) given the same
    inputs, and the CUDA ancestors are BIT-EXACT the oracle's resample of the CUDA weights."""
    gj, step, _ = _models()
    from genjax_b200.inference.pf import ParticleFilter

    n, T = 6000, 12
    ys = osmc.simulate_lgssm(0, T, 1, A_, Q_, C_, R_)[:, 0]
    g = np.random.default_rng(3)
    x0 = g.standard_normal(n).astype(np.float32)
    key = gj.key(99)
    pf = ParticleFilter(step, n, mode=mode)
    res = pf.run(key, torch.from_numpy(x0), gj.C["y"].set(torch.from_numpy(ys)), record=True, use_graph=use_graph)
    anc = res.ancestors.cpu().numpy()
    xs = res.history["state"][0].cpu().numpy()
    lws = res.history["log_weights"].cpu().numpy()
    okey = orng.key(99)
    x_in = x0
    for t in range(T):
        k_prop, k_res = osmc.pf_step_keys(okey, t)
        otr, ow = ogfi.generate(o_step, orng.split(k_prop, n), {"y": np.float32(ys[t])}, (x_in,))
        np.testing.assert_allclose(xs[t], otr.choices["x"], rtol=1e-5, atol=2e-6, err_msg=f"x step {t}")
        np.testing.assert_allclose(lws[t], ow, rtol=1e-5, atol=2e-5, err_msg=f"logw step {t}")
        exp_anc = osmc.resample_systematic(lws[t], k_res)
        assert np.array_equal(anc[t], exp_anc), f"ancestors step {t}"
        assert res.log_increments[t].item() == pytest.approx(osmc.log_mean_exp(lws[t]), abs=1e-9)
        x_in = xs[t][anc[t]]  # teacher forcing: continue from the CUDA state
    np.testing.assert_array_equal(res.state[0].cpu().numpy(), x_in)


def test_particle_filter_matches_kalman(device):
    gj, step, _ = _models()
    from genjax_b200.inference.pf import ParticleFilter

    n, T = 1 << 18, 50
    ys = osmc.simulate_lgssm(0, T, 1, A_, Q_, C_, R_)[:, 0]
    exact = osmc.kalman_logz(ys, A_, Q_, C_, R_)
    keys0 = gj.split(gj.key(1), n)
    x0 = gj.normal.sample(keys0, 0.0, 1.0)
    ests = []
    for seed in range(4):
        pf = ParticleFilter(step, n)
        res = pf.run(gj.key(seed), x0, gj.C["y"].set(torch.from_numpy(ys)))
        ests.append(res.log_marginal_likelihood.item())
    # sd(log Z-hat) of one run = 0.029 (8 seeds through tests/abi_emulator.py): the mean of 4 runs has sd 0.015
    assert np.mean(ests) == pytest.approx(exact, abs=0.05)
    assert np.std(ests) < 0.08


def test_particle_filter_vec_teacher_forced_vs_oracle(device):
    gj, _, step_vec = _models()
    from genjax_b200.inference.pf import ParticleFilter

    d, n, T = 8, 5000, 6
    ys = osmc.simulate_lgssm(2, T, d, A_, Q_, C_, R_)
    x0 = torch.randn(n, d, generator=torch.Generator().manual_seed(0))
    q = torch.full((d,), Q_)
    r = torch.full((d,), R_)
    res = ParticleFilter(step_vec, n, mode="graph").run(
        gj.key(5), x0, gj.C["y"].set(torch.from_numpy(ys)), shared_args=(q, r), record=True
    )
    anc = res.ancestors.cpu().numpy()
    xs = res.history["state"][0].cpu().numpy()
    lws = res.history["log_weights"].cpu().numpy()
    x_in = x0.numpy()
    okey = orng.key(5)
    for t in range(T):
        kp, kr = osmc.pf_step_keys(okey, t)
        otr, ow = ogfi.generate(o_step_vec, orng.split(kp, n), {"y": ys[t]}, (x_in, q.numpy(), r.numpy()))
        np.testing.assert_allclose(xs[t], otr.choices["x"], rtol=1e-5, atol=4e-6)
        np.testing.assert_allclose(lws[t], ow, rtol=2e-5, atol=2e-4)
        assert np.array_equal(anc[t], osmc.resample_systematic(lws[t], kr))
        x_in = xs[t][anc[t]]
    np.testing.assert_array_equal(res.state[0].cpu().numpy(), x_in)


def test_particle_filter_vec_matches_kalman(device):
    gj, _, step_vec = _models()
    from genjax_b200.inference.pf import ParticleFilter

    d, n, T = 8, 1 << 18, 20
    r_obs = 2.0  # weakly informative observations keep the 8-D bootstrap filter's variance small
    ys = osmc.simulate_lgssm(2, T, d, A_, Q_, C_, r_obs)
    exact = osmc.kalman_logz(ys, A_, Q_, C_, r_obs)
    x0 = torch.randn(n, d, generator=torch.Generator().manual_seed(0))
    q = torch.full((d,), Q_)
    r = torch.full((d,), r_obs)
    # one run of this 8-D bootstrap filter has sd(log Z-hat) = 0.24 (8 seeds through the oracle-backed emulation,
    # tests/abi_emulator.py; the Jensen bias is -0.03): the mean of 8 seeds has sd 0.084, tolerance = 4 sd
    ests = []
    pf = ParticleFilter(step_vec, n)
    for seed in range(5, 13):
        res = pf.run(gj.key(seed), x0, gj.C["y"].set(torch.from_numpy(ys)), shared_args=(q, r))
        ests.append(res.log_marginal_likelihood.item())
    assert np.mean(ests) == pytest.approx(exact, abs=0.34)
    assert np.std(ests, ddof=1) < 0.6


@pytest.mark.parametrize("n", [1, 7, 2049, 300_001, 3_000_000])
def test_persistent_filter_equals_three_kernel_filter(device, n):
    """The one-launch cooperative filter and the 3-launches-per-step filter are the same
    algorithm: bit-identical states, log-weights, ancestors and log-marginal increments
    (n = 3M exercises several 2048-particle tiles per CTA)."""
    gj, step, _ = _models()
    from genjax_b200.inference.pf import ParticleFilter

    T = 5
    ys = osmc.simulate_lgssm(1, T, 1, A_, Q_, C_, R_)[:, 0]
    x0 = torch.randn(n, generator=torch.Generator().manual_seed(n))
    out = {}
    for mode in ("persistent", "graph"):
        res = ParticleFilter(step, n, mode=mode).run(gj.key(17), x0, gj.C["y"].set(torch.from_numpy(ys)), record=True, use_graph=False)
        torch.cuda.synchronize()
        out[mode] = res
    a, b = out["persistent"], out["graph"]
    assert torch.equal(a.ancestors, b.ancestors)
    assert torch.equal(a.history["log_weights"], b.history["log_weights"])
    assert torch.equal(a.history["state"][0], b.history["state"][0])
    assert torch.equal(a.log_increments, b.log_increments)
    assert torch.equal(a.state[0], b.state[0])
    # and the last step's ancestors are the oracle's resample of those weights
    _, k_res = osmc.pf_step_keys(orng.key(17), T - 1)
    lw = a.history["log_weights"][T - 1].cpu().numpy()
    assert np.array_equal(a.ancestors[T - 1].cpu().numpy(), osmc.resample_systematic(lw, k_res))


def test_persistent_filter_non_record_matches_record(device):
    gj, step, _ = _models()
    from genjax_b200.inference.pf import ParticleFilter

    n, T = 50_000, 9
    ys = osmc.simulate_lgssm(4, T, 1, A_, Q_, C_, R_)[:, 0]
    x0 = torch.randn(n, generator=torch.Generator().manual_seed(1))
    r1 = ParticleFilter(step, n).run(gj.key(3), x0, gj.C["y"].set(torch.from_numpy(ys)), record=True)
    s1, z1 = r1.state[0].clone(), r1.log_increments.clone()
    r2 = ParticleFilter(step, n).run(gj.key(3), x0, gj.C["y"].set(torch.from_numpy(ys)), record=False)
    assert torch.equal(s1, r2.state[0]) and torch.equal(z1, r2.log_increments)


def o_hmm_step(h, z_prev, trans, obs):
    z = h.categorical("z", trans[z_prev])
    h.categorical("y", obs[z])
    return z


def _hmm_tables(K=16, sig_t=0.5, sig_o=0.5):
    """Banded circulant log-potentials a la discrete_hmm.py:42-52 (scaled_circulant)."""
    i = np.arange(K)
    d = np.minimum((i[:, None] - i[None, :]) % K, (i[None, :] - i[:, None]) % K).astype(np.float64)
    trans = -0.5 * (d / sig_t) ** 2
    obs = -0.5 * (d / sig_o) ** 2
    return trans.astype(np.float32), obs.astype(np.float32)


def test_hmm_filter_teacher_forced_and_exact(device):
    """16-state HMM bootstrap filter (BASELINE configs[3] model): categorical sites reading
    rows of shared logit tables staged in shared memory; integer particle state."""
    import genjax_b200 as gj
    from genjax_b200.inference.pf import ParticleFilter
    from genjax_b200.workloads import hmm_step

    K, n, T = 16, 40_000, 10
    trans, obs = _hmm_tables(K)
    g = np.random.default_rng(3)
    z = 0
    ys = np.empty(T, dtype=np.int32)
    pt = np.exp(trans - trans.max(1, keepdims=True)); pt /= pt.sum(1, keepdims=True)
    po = np.exp(obs - obs.max(1, keepdims=True)); po /= po.sum(1, keepdims=True)
    for t in range(T):
        z = g.choice(K, p=pt[z])
        ys[t] = g.choice(K, p=po[z])
    z0 = g.integers(0, K, n).astype(np.int32)
    res = ParticleFilter(hmm_step, n, mode="graph").run(
        gj.key(11), torch.from_numpy(z0), gj.C["y"].set(torch.from_numpy(ys)),
        shared_args=(torch.from_numpy(trans), torch.from_numpy(obs)), record=True)
    anc = res.ancestors.cpu().numpy()
    zs = res.history["state"][0].cpu().numpy()
    lws = res.history["log_weights"].cpu().numpy()
    okey = orng.key(11)
    z_in = z0
    mism = 0
    for t in range(T):
        kp, kr = osmc.pf_step_keys(okey, t)
        otr, ow = ogfi.generate(o_hmm_step, orng.split(kp, n), {"y": np.int32(ys[t])}, (z_in, trans, obs))
        same = zs[t] == otr.choices["z"]
        mism += int((~same).sum())  # a draw within 1 ulp of a CDF edge may land in the neighbouring state
        np.testing.assert_allclose(lws[t][same], ow[same], rtol=1e-5, atol=1e-5)
        assert np.array_equal(anc[t], osmc.resample_systematic(lws[t], kr))
        z_in = zs[t][anc[t]]
    assert mism <= 2e-5 * n * T
    # exact forward algorithm for the empirical initial distribution
    alpha = np.bincount(z0, minlength=K) / n
    ll = 0.0
    for t in range(T):
        alpha = (alpha @ pt) * po[:, ys[t]]
        ll += np.log(alpha.sum())
        alpha /= alpha.sum()
    assert res.log_marginal_likelihood.item() == pytest.approx(ll, abs=0.05)


def test_multi_gpu_global_resampling_equals_single_gpu(device):
    """R-rank filter with global resampling (peer-mapped ancestor writes / state gathers, push-poll
    exchanges) == single-GPU filter of R*n particles, bit for bit.  Needs >= 2 GPUs (gpurun --gpus 2)."""
    import os
    import subprocess
    import sys

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(root, "tests", "dist_pf_worker.py")]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0 and "DIST_PF_OK" in p.stdout, p.stdout[-3000:] + p.stderr[-3000:]


@pytest.mark.parametrize("use_graph", [False, True])
@pytest.mark.parametrize("obs_sd", [0.5, 0.0001])
def test_fused_mass_resample_equals_two_launches(device, use_graph, obs_sd):
    """gjb_mass_resample_systematic (one cooperative launch, masses kept in registers across a grid barrier)
    == gjb_weight_mass + gjb_resample_systematic, bit for bit; also inside a captured CUDA graph."""
    gj, step, _ = _models()
    from genjax_b200.inference.pf import ParticleFilter

    if obs_sd != 0.5:
        # a razor-sharp likelihood: a handful of particles own tens of thousands of offspring each, so the fused
        # kernel parks whole 4096-slot windows in the heavy list and the grid fills them after the second barrier
        @gj.gen
        def step(x_prev):  # noqa: F811
            x = gj.normal(A_ * x_prev, Q_) @ "x"
            gj.normal(C_ * x, obs_sd) @ "y"
            return x

    n, T = 200_003, 6
    ys = osmc.simulate_lgssm(3, T, 1, A_, Q_, C_, R_)[:, 0]
    x0 = torch.randn(n, generator=torch.Generator().manual_seed(2))
    outs = []
    for fuse in (True, False):
        pf = ParticleFilter(step, n, mode="graph")
        pf.fuse_mass_resample = fuse
        res = pf.run(gj.key(8), x0, gj.C["y"].set(torch.from_numpy(ys)), record=True, use_graph=use_graph)
        res = pf.run(gj.key(8), x0, gj.C["y"].set(torch.from_numpy(ys)), record=True, use_graph=use_graph)  # replay
        torch.cuda.synchronize()
        plan = next(iter(pf._plans.values()))
        assert plan.fuse_mass_resample == fuse
        outs.append((res.ancestors.clone(), res.log_increments.clone(), res.state[0].clone()))
    for a, b in zip(*outs):
        assert torch.equal(a, b)
    if obs_sd != 0.5:
        anc = outs[0][0][-1].cpu().numpy()
        counts = np.bincount(outs[0][0].cpu().numpy().reshape(-1) + np.repeat(np.arange(T) * n, n))
        assert counts.max() > 2 * 4096, counts.max()  # really degenerate: some parent owns whole 4096-slot windows
        lw = res.history["log_weights"][-1].cpu().numpy()
        _, k_res = osmc.pf_step_keys(orng.key(8), T - 1)
        assert np.array_equal(anc, osmc.resample_systematic(lw, k_res))


def test_particle_filter_multinomial_resampler(device):
    """``ParticleFilter(resampler="multinomial")``: the reference idiom's N independent categorical draws per step
    (mapping_tutorial.ipynb cell 37; inference/smc.py:102-109) -- ancestors bit-exact against the oracle's inverse-CDF
    multinomial over the same integer CDF, offspring j on lane j of split(k_res, N); CUDA-graph replay with a new key."""
    gj, step, _ = _models()
    from genjax_b200.inference.pf import ParticleFilter

    n, T = 20_000, 5
    ys = osmc.simulate_lgssm(4, T, 1, A_, Q_, C_, R_)[:, 0]
    x0 = np.random.default_rng(2).standard_normal(n).astype(np.float32)
    pf = ParticleFilter(step, n, resampler="multinomial")
    pf.run(gj.key(1), torch.from_numpy(x0), gj.C["y"].set(torch.from_numpy(ys)), record=True)  # captures the graph with another key
    res = pf.run(gj.key(33), torch.from_numpy(x0), gj.C["y"].set(torch.from_numpy(ys)), record=True)
    anc = res.ancestors.cpu().numpy()
    xs = res.history["state"][0].cpu().numpy()
    lws = res.history["log_weights"].cpu().numpy()
    x_in, okey = x0, orng.key(33)
    for t in range(T):
        kp, kr = osmc.pf_step_keys(okey, t)
        otr, ow = ogfi.generate(o_step, orng.split(kp, n), {"y": np.float32(ys[t])}, (x_in,))
        np.testing.assert_allclose(xs[t], otr.choices["x"], rtol=1e-5, atol=2e-6)
        np.testing.assert_allclose(lws[t], ow, rtol=1e-5, atol=2e-5)
        assert np.array_equal(anc[t], osmc.resample_multinomial(lws[t], orng.split(kr, n))), f"ancestors step {t}"
        assert res.log_increments[t].item() == pytest.approx(osmc.log_mean_exp(lws[t]), abs=1e-9)
        x_in = xs[t][anc[t]]
    np.testing.assert_array_equal(res.state[0].cpu().numpy(), x_in)
    # multinomial offspring counts are not sorted-unique like systematic ones: duplicates and gaps both occur
    assert len(np.unique(anc[-1])) < 0.7 * n
    with pytest.raises(ValueError):
        ParticleFilter(step, n, resampler="stratified")

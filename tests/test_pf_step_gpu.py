"""GPU parity of the single-launch filter step (ParticleFilter(mode="step"): gjb_model_pf_step / gjb_te_resample) and of
the stand-alone tile-exponent kernels (gjb_te_masses) with the oracle (oracle/smc.py te_*), through the C-ABI: bit-exact
within-tile CDFs, tile records and ancestors; proposals / log-weights within the fp32 tolerances of tests/test_pf_gpu.py;
log-marginal-likelihood against the exact Kalman filter and HMM forward algorithm at full size."""
import math

import numpy as np
import pytest
import torch

from oracle import rng as orng
from oracle import smc as osmc
from pf_step_common import check_against_oracle

pytestmark = pytest.mark.gpu
F32 = np.float32


def _wl():
    import genjax_b200 as gj
    from genjax_b200 import workloads as wl
    from genjax_b200.inference.pf import ParticleFilter

    return gj, wl, ParticleFilter


def o_step(h, x_prev):
    from genjax_b200.workloads import LG_A, LG_C, LG_Q, LG_R

    x = h.normal("x", F32(LG_A) * x_prev, F32(LG_Q))
    h.normal("y", F32(LG_C) * x, F32(LG_R))
    return x


@pytest.mark.parametrize("n,scale,kind", [(1, 1.0, "n"), (7, 1.0, "n"), (2048, 1.0, "n"), (2049, 2.0, "n"), (100_003, 0.3, "n"),
                                           (1 << 20, 2.0, "n"), (50_000, 30.0, "n"), (12_345, 300.0, "n"), (5000, 1.0, "one"),
                                           (5000, 1.0, "dead"), (7000, 1.0, "deadtile"), (4100, 1.0, "nan")])
def test_tile_exponent_kernels_bit_exact(device, n, scale, kind):
    from genjax_b200.runtime import smc_ops

    r = np.random.default_rng(n)
    lw = (r.standard_normal(n) * scale - 3.0).astype(F32)
    if kind == "one":
        lw[:] = -np.inf
        lw[n - 3] = 1.5
    elif kind == "dead":
        lw[:] = -np.inf
    elif kind == "deadtile":
        lw[2048:4096] = -np.inf
        lw[5] = np.nan
    elif kind == "nan":
        lw[::7] = np.nan
        lw[3] = np.inf
    ws = smc_ops.TeWorkspace(n, device).masses(torch.from_numpy(lw).to(device))
    tiles = ws.tiles
    q, e_p = osmc.te_tile_masses(lw)
    qp = np.zeros(tiles * 2048, dtype=np.uint64)
    qp[:n] = q
    want_cdf = np.cumsum(qp.reshape(tiles, 2048), axis=1, dtype=np.uint64).reshape(-1)
    assert np.array_equal(ws.cdf.cpu().numpy().view(np.uint64), want_cdf)
    recs = ws.recs.cpu().numpy()
    assert np.array_equal(recs[:, 0].view(np.uint64), want_cdf.reshape(tiles, 2048)[:, -1])
    live = recs[:, 0] != 0
    assert np.array_equal((recs[:, 1] & 0xFFFFFFFF).astype(np.uint32).view(np.int32)[live], e_p[live].astype(np.int32))
    key = orng.Key((0x1234 + n, 77), 5)

    class K:  # the product-side key type only needs words / index here
        words, index = key.words, key.index

    for out_lo, out_n in ((0, n), (n // 3, n - n // 3 - n // 5)):
        if out_n <= 0:
            continue
        anc = torch.full((out_n,), -7, dtype=torch.int32, device=device)
        ws.resample(K, anc, out_lo)
        want = osmc.resample_systematic_te(lw, key, out_lo, out_n)
        got = anc.cpu().numpy()
        assert np.array_equal(got, want), (np.flatnonzero(got != want)[:5], got[:8], want[:8])
        lse = ws.lse.cpu().numpy()
        lme = osmc.te_log_mean_exp(lw)
        if np.isfinite(lme):
            assert lse[2] == pytest.approx(lme, abs=1e-11)
            assert lse[1] == float(osmc.te_cdf(lw)[1])
        else:
            assert lse[1] == 0.0 and lse[2] == -np.inf


@pytest.mark.parametrize("n,T", [(7, 5), (2048, 5), (100_001, 4), (1 << 20, 3)])
@pytest.mark.parametrize("use_graph", [False, True])
def test_step_filter_teacher_forced_vs_oracle(device, n, T, use_graph):
    gj, wl, ParticleFilter = _wl()
    ys = osmc.simulate_lgssm(1, T, 1, wl.LG_A, wl.LG_Q, wl.LG_C, wl.LG_R)[:, 0]
    x0 = np.random.default_rng(n).standard_normal(n).astype(F32)
    obs = gj.C["y"].set(torch.from_numpy(ys))
    res = ParticleFilter(wl.lgssm_step, n, mode="step").run(gj.key(17), torch.from_numpy(x0), obs, record=True, use_graph=use_graph)
    check_against_oracle(res, x0, [{"y": F32(y)} for y in ys], o_step, 17, n, T)
    # non-record run (ping-pong buffers; weights and ancestors leave the kernels only at the end): identical
    res2 = ParticleFilter(wl.lgssm_step, n, mode="step").run(gj.key(17), torch.from_numpy(x0), obs, use_graph=use_graph)
    assert torch.equal(res2.log_increments, res.log_increments) and torch.equal(res2.state[0], res.state[0])
    # replaying the captured graph with another key / initial state gives that run's result
    if use_graph:
        pf = ParticleFilter(wl.lgssm_step, n, mode="step")
        a = pf.run(gj.key(3), torch.from_numpy(x0), obs)
        a_inc = a.log_increments.clone()
        b = pf.run(gj.key(17), torch.from_numpy(x0), obs)
        assert torch.equal(b.log_increments, res.log_increments) and not torch.equal(a_inc, b.log_increments)


@pytest.mark.parametrize("d,n", [(8, 5000), (32, 70_000)])
def test_step_filter_vector_model(device, d, n):
    gj, wl, ParticleFilter = _wl()
    T = 3
    r_sd = 0.5 * math.sqrt(d)
    q = np.full(d, wl.LG_Q, dtype=F32)
    r = np.full(d, r_sd, dtype=F32)
    ys = osmc.simulate_lgssm(2, T, d, wl.LG_A, wl.LG_Q, wl.LG_C, r_sd)
    x0 = np.random.default_rng(3).standard_normal((n, d)).astype(F32)

    def o_vec(h, x_prev, q_, r_):
        x = h.mv_normal_diag("x", F32(wl.LG_A) * x_prev, q_)
        h.mv_normal_diag("y", F32(wl.LG_C) * x, r_)
        return x

    res = ParticleFilter(wl.lgssm_step_vec, n, mode="step").run(
        gj.key(5), torch.from_numpy(x0), gj.C["y"].set(torch.from_numpy(ys)), (torch.from_numpy(q), torch.from_numpy(r)), record=True)
    # (a 32-D weight vector spans tens of nats: the two pipelines' 2^-36 quantisation steps sit at different references)
    check_against_oracle(res, x0, [{"y": y.astype(F32)} for y in ys], o_vec, 5, n, T, shared=(q, r), tol=(1e-5, 4e-6, 2e-5, 2e-4), lme_tol=5e-6)


def test_step_filter_integer_state_hmm(device):
    gj, wl, ParticleFilter = _wl()
    n, T, K = 25_000, 4, 16
    rg = np.random.default_rng(0)
    trans = rg.standard_normal((K, K)).astype(F32)
    obsl = rg.standard_normal((K, K)).astype(F32)
    ys = rg.integers(0, K, T).astype(np.int32)
    z0 = rg.integers(0, K, n).astype(np.int32)

    def o_hmm(h, z_prev, tl, ol):
        z = h.categorical("z", tl[z_prev])
        h.categorical("y", ol[z])
        return z

    res = ParticleFilter(wl.hmm_step, n, mode="step").run(
        gj.key(9), torch.from_numpy(z0), gj.C["y"].set(torch.from_numpy(ys)), (torch.from_numpy(trans), torch.from_numpy(obsl)), record=True)
    check_against_oracle(res, z0, [{"y": np.int32(y)} for y in ys], o_hmm, 9, n, T, shared=(trans, obsl))


def test_step_filter_at_the_bench_configuration(device):
    """BASELINE configs[1] at full size (N = 1 048 576, T = 100): steps 0, 49 and 99 teacher-forced against the oracle
    (bit-exact ancestors over 1 M particles), the log-marginal-likelihood within 4 sigma of the exact Kalman filter,
    and the step filter against the round-1 default (mode='graph': exact-max masses, 2 launches per step), which is a
    different realisation of the same estimator."""
    gj, wl, ParticleFilter = _wl()
    n, T = 1 << 20, 100
    ys = osmc.simulate_lgssm(0, T, 1, wl.LG_A, wl.LG_Q, wl.LG_C, wl.LG_R)[:, 0]
    exact = osmc.kalman_logz(ys, wl.LG_A, wl.LG_Q, wl.LG_C, wl.LG_R)
    x0 = np.random.default_rng(0).standard_normal(n).astype(F32)
    obs = gj.C["y"].set(torch.from_numpy(ys))
    res = ParticleFilter(wl.lgssm_step, n, mode="step").run(gj.key(314159), torch.from_numpy(x0), obs, record=True)
    check_against_oracle(res, x0, [{"y": F32(y)} for y in ys], o_step, 314159, n, T, steps=(0, 49, 99))
    # sd(log Z-hat) at N = 2^18 is 0.029 (tests/test_pf_gpu.py); at 2^20 half of that: 4 sigma = 0.06
    assert res.log_marginal_likelihood.item() == pytest.approx(exact, abs=0.06)
    old = ParticleFilter(wl.lgssm_step, n, mode="graph").run(gj.key(314159), torch.from_numpy(x0), obs)
    assert old.log_marginal_likelihood.item() == pytest.approx(exact, abs=0.06)
    # step 0 sees identical inputs in both pipelines: the estimate agrees to the masses' 2^-36 quantisation
    assert res.log_increments[0].item() == pytest.approx(old.log_increments[0].item(), abs=5e-7)


def test_step_filter_matches_kalman_over_seeds(device):
    gj, wl, ParticleFilter = _wl()
    n, T = 1 << 18, 50
    ys = osmc.simulate_lgssm(0, T, 1, wl.LG_A, wl.LG_Q, wl.LG_C, wl.LG_R)[:, 0]
    exact = osmc.kalman_logz(ys, wl.LG_A, wl.LG_Q, wl.LG_C, wl.LG_R)
    x0 = gj.normal.sample(gj.split(gj.key(1), n), 0.0, 1.0)
    pf = ParticleFilter(wl.lgssm_step, n, mode="step")
    ests = [pf.run(gj.key(seed), x0, gj.C["y"].set(torch.from_numpy(ys))).log_marginal_likelihood.item() for seed in range(4)]
    assert np.mean(ests) == pytest.approx(exact, abs=0.05)
    assert np.std(ests) < 0.08


def test_filter_result_diagnostics_checkpoint_and_html(device, tmp_path):
    """per-step ESS, the ancestry (genealogy) of the surviving lineages, a checkpoint of the
    run as CPU tensors, resuming a filter from it, and the jax-free render_html."""
    gj, wl, ParticleFilter = _wl()
    n, T = 6000, 6
    ys = osmc.simulate_lgssm(1, T, 1, wl.LG_A, wl.LG_Q, wl.LG_C, wl.LG_R)[:, 0]
    x0 = np.random.default_rng(7).standard_normal(n).astype(F32)
    obs = gj.C["y"].set(torch.from_numpy(ys))
    pf = ParticleFilter(wl.lgssm_step, n)
    res = pf.run(gj.key(4), torch.from_numpy(x0), obs, record=True)
    lws = res.history["log_weights"].cpu().numpy().astype(np.float64)
    w = np.exp(lws - lws.max(1, keepdims=True))
    np.testing.assert_allclose(res.ess.cpu().numpy(), w.sum(1) ** 2 / (w * w).sum(1), rtol=1e-5)
    assert (res.ess > 1).all() and (res.ess <= n).all()
    # genealogy: composing the recorded ancestors backwards
    anc = res.ancestors.cpu().numpy()
    cur = anc[T - 1].copy()
    gen = res.genealogy().cpu().numpy()
    assert np.array_equal(gen[T - 1], cur)
    for t in range(T - 2, -1, -1):
        cur = anc[t][cur]
        assert np.array_equal(gen[t], cur)
    assert len(np.unique(gen[0])) <= len(np.unique(gen[T - 1]))  # lineages coalesce backwards in time
    with pytest.raises(ValueError):
        pf.run(gj.key(4), torch.from_numpy(x0), obs).ess
    # checkpoint round trip + resume: filtering y[:3] then y[3:] from the saved state is a filter over all six steps
    path = tmp_path / "pf.pt"
    first = pf.run(gj.key(4), torch.from_numpy(x0), gj.C["y"].set(torch.from_numpy(ys[:3])), record=True)
    first.save(path)
    ck = type(first).load(path)
    assert ck["format"] == "genjax_b200.PFResult/1" and ck["ancestors"].shape == (3, n) and ck["ess"].shape == (3,)
    assert torch.equal(ck["state"][0], first.state[0].cpu())
    second = pf.run(gj.key(5), ck["state"][0], gj.C["y"].set(torch.from_numpy(ys[3:])))
    total = ck["log_marginal_likelihood"] + second.log_marginal_likelihood.item()
    exact = osmc.kalman_logz(ys, wl.LG_A, wl.LG_Q, wl.LG_C, wl.LG_R)
    assert total == pytest.approx(exact, abs=0.5)
    page = gj.render_html(res)
    assert page.startswith("<div") and "PFResult" in page and "ESS per step" in page
    assert "ChoiceMap" in gj.render_html(obs) and "float32" in gj.render_html({"x": torch.zeros(3)})


@pytest.mark.parametrize("n,T", [(7, 5), (6000, 6), (1 << 20, 12)])
def test_all_steps_in_one_cooperative_launch_equal_the_per_step_launches(device, n, T):
    """mode="steps" (gjb_model_pf_steps: the whole filter in ONE cooperative launch, one grid barrier per step) ==
    mode="step" (one launch per step), bit for bit: ancestors, states, weights, increments, final state; scalar and
    vector models; record and ping-pong buffers."""
    gj, wl, ParticleFilter = _wl()
    ys = osmc.simulate_lgssm(1, T, 1, wl.LG_A, wl.LG_Q, wl.LG_C, wl.LG_R)[:, 0]
    x0 = torch.from_numpy(np.random.default_rng(n).standard_normal(n).astype(F32))
    obs = gj.C["y"].set(torch.from_numpy(ys))
    for record in (True, False):
        a = ParticleFilter(wl.lgssm_step, n, mode="steps").run(gj.key(17), x0, obs, record=record)
        b = ParticleFilter(wl.lgssm_step, n, mode="step").run(gj.key(17), x0, obs, record=record)
        assert torch.equal(a.log_increments, b.log_increments) and torch.equal(a.state[0], b.state[0])
        if record:
            assert torch.equal(a.ancestors, b.ancestors) and torch.equal(a.history["log_weights"], b.history["log_weights"])
            assert torch.equal(a.history["state"][0], b.history["state"][0])
    if n >= 6000:
        d = 8
        q, r = torch.full((d,), wl.LG_Q), torch.full((d,), 0.5 * math.sqrt(d))
        ysv = torch.from_numpy(osmc.simulate_lgssm(2, T, d, wl.LG_A, wl.LG_Q, wl.LG_C, 0.5 * math.sqrt(d)))
        x0v = torch.from_numpy(np.random.default_rng(3).standard_normal((min(n, 70_000), d)).astype(F32))
        nv = x0v.shape[0]
        a = ParticleFilter(wl.lgssm_step_vec, nv, mode="steps").run(gj.key(5), x0v, gj.C["y"].set(ysv), (q, r), record=True)
        b = ParticleFilter(wl.lgssm_step_vec, nv, mode="step").run(gj.key(5), x0v, gj.C["y"].set(ysv), (q, r), record=True)
        assert torch.equal(a.ancestors, b.ancestors) and torch.equal(a.log_increments, b.log_increments)
        assert torch.equal(a.state[0], b.state[0])
    with pytest.raises(NotImplementedError):  # more windows than resident CTAs: the per-step form is the one to use
        ParticleFilter(wl.lgssm_step, 1 << 22, mode="steps").run(gj.key(1), torch.zeros(1 << 22), gj.C["y"].set(torch.from_numpy(ys)))

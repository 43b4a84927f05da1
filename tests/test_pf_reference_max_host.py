"""The reference-maximum filter step (ParticleFilter(reference_max="analytic")) on CPU: the
model kernel accumulates the exact integer masses relative to an analytic bound while the weights are in registers
(generated source run by tests/host_kernels.py, or the IR interpreter), the resampler takes the bound as its
reference -- against the oracle filter with the same reference (oracle/smc.py particle_filter(m_ref=...))."""
import numpy as np
import pytest
import torch

import abi_emulator
import genjax_b200 as gj
from genjax_b200.inference.pf import ParticleFilter
from genjax_b200.workloads import LG_A, LG_C, LG_Q, LG_R, hmm_step, lgssm_step, lgssm_step_vec
from oracle import gfi as ogfi
from oracle import rng as orng
from oracle import smc as osmc

F32 = np.float32


@pytest.fixture(params=["ir", "host"])
def emu(monkeypatch, request):
    return abi_emulator.install(monkeypatch, host_kernels=request.param == "host")


def o_step(h, x_prev):
    x = h.normal("x", F32(LG_A) * x_prev, F32(LG_Q))
    h.normal("y", F32(LG_C) * x, F32(LG_R))
    return x


@pytest.mark.parametrize("n", [7, 2048, 5000])
def test_analytic_reference_filter_matches_oracle(emu, n):
    T = 6
    ys = osmc.simulate_lgssm(1, T, 1, LG_A, LG_Q, LG_C, LG_R)[:, 0]
    x0 = np.random.default_rng(n).standard_normal(n).astype(F32)
    pf = ParticleFilter(lgssm_step, n, reference_max="analytic")
    bound = pf.weight_upper_bound(torch.from_numpy(x0), gj.C["y"].set(torch.from_numpy(ys)))
    res = pf.run(gj.key(17), torch.from_numpy(x0), gj.C["y"].set(torch.from_numpy(ys)), record=True)
    anc, lws, xs = res.ancestors.numpy(), res.history["log_weights"].numpy(), res.history["state"][0].numpy()
    x_in, okey = x0, orng.key(17)
    for t in range(T):
        kp, kr = osmc.pf_step_keys(okey, t)
        otr, ow = ogfi.generate(o_step, orng.split(kp, n), {"y": F32(ys[t])}, (x_in,))
        np.testing.assert_allclose(xs[t], otr.choices["x"], rtol=1e-5, atol=2e-6)
        np.testing.assert_allclose(lws[t], ow, rtol=1e-5, atol=2e-5)
        assert lws[t].max() <= bound + 1e-6
        # ancestors and masses: exact, given the kernel's own weights and the SAME reference
        assert np.array_equal(anc[t], osmc.resample_systematic_pull(lws[t], kr, M=F32(bound)))
        M, S, inc = res.lse_terms[t].numpy()
        assert M == float(F32(bound))
        assert S == float(int(osmc.det_exp_q((lws[t] - F32(bound)).astype(F32)).sum(dtype=np.uint64)))
        assert inc == pytest.approx(osmc.log_mean_exp_ref(lws[t], F32(bound)), abs=1e-12)
        assert inc == pytest.approx(osmc.log_mean_exp(lws[t]), abs=2e-7)  # same estimate as with the running max
        x_in = xs[t][anc[t]]
    np.testing.assert_array_equal(res.state[0].numpy(), x_in)
    # the oracle filter run end to end with the same reference agrees on the estimate
    ores = osmc.particle_filter(o_step, orng.key(17), x0, [{"y": F32(y)} for y in ys], m_ref=F32(bound))
    assert res.log_marginal_likelihood.item() == pytest.approx(ores["logz"], abs=2e-4)
    # non-record run (ping-pong buffers, tile-mass buffers alternate and clear each other): same result
    res2 = ParticleFilter(lgssm_step, n, reference_max="analytic").run(
        gj.key(17), torch.from_numpy(x0), gj.C["y"].set(torch.from_numpy(ys)))
    assert torch.equal(res2.log_increments, res.log_increments) and torch.equal(res2.state[0], res.state[0])


def test_analytic_reference_and_running_max_are_the_same_estimator(emu):
    n, T = 4096, 10
    ys = torch.from_numpy(osmc.simulate_lgssm(3, T, 1, LG_A, LG_Q, LG_C, LG_R)[:, 0])
    x0 = torch.randn(n, generator=torch.Generator().manual_seed(1))
    a = ParticleFilter(lgssm_step, n, reference_max="analytic").run(gj.key(5), x0, gj.C["y"].set(ys), record=True)
    b = ParticleFilter(lgssm_step, n, mode="graph").run(gj.key(5), x0, gj.C["y"].set(ys), record=True)
    # step 0 sees identical inputs: identical weights, estimate equal to fp64 rounding; ancestors may differ in a few
    # slots (masses are quantised relative to a different reference), after which the runs are different draws
    assert torch.equal(a.history["log_weights"][0], b.history["log_weights"][0])
    assert a.log_increments[0].item() == pytest.approx(b.log_increments[0].item(), abs=2e-7)
    assert (a.ancestors[0] != b.ancestors[0]).float().mean().item() < 0.01
    assert a.log_marginal_likelihood.item() == pytest.approx(b.log_marginal_likelihood.item(), abs=1.0)


def test_analytic_reference_needs_a_derivable_bound_and_graph_mode(emu):
    @gj.gen
    def hetero(x_prev):
        x = gj.normal(x_prev, 1.0) @ "x"
        gj.normal(x, gj.numpy.exp(0.1 * x)) @ "y"
        return x

    with pytest.raises(ValueError, match="no particle-free bound"):
        ParticleFilter(hetero, 64, reference_max="analytic").run(gj.key(0), torch.zeros(64), gj.C["y"].set(torch.zeros(3)))
    with pytest.raises(ValueError):
        ParticleFilter(lgssm_step, 64, mode="persistent", reference_max="analytic")
    # vector-site models run lane-group kernels, which have no mass instantiation yet
    with pytest.raises(NotImplementedError, match="scalar-site"):
        ParticleFilter(lgssm_step_vec, 64, reference_max="analytic").run(
            gj.key(0), torch.zeros(64, 4), gj.C["y"].set(torch.zeros(3, 4)), shared_args=(torch.ones(4), torch.ones(4)))
    # discrete model: bound 0, masses are the emission probabilities themselves
    n, T = 1000, 4
    g = np.random.default_rng(0)
    tl = torch.from_numpy(g.standard_normal((16, 16)).astype(F32))
    ol = torch.from_numpy(g.standard_normal((16, 16)).astype(F32))
    ys = torch.from_numpy(g.integers(0, 16, T).astype(np.int32))
    z0 = torch.from_numpy(g.integers(0, 16, n).astype(np.int32))
    res = ParticleFilter(hmm_step, n, reference_max="analytic").run(gj.key(2), z0, gj.C["y"].set(ys), shared_args=(tl, ol), record=True)
    assert (res.lse_terms[:, 0] == 0).all() and (res.lse_terms[:, 1] > 0).all()
    ref = ParticleFilter(hmm_step, n, mode="graph").run(gj.key(2), z0, gj.C["y"].set(ys), shared_args=(tl, ol), record=True)
    assert res.log_increments[0].item() == pytest.approx(ref.log_increments[0].item(), abs=2e-7)


@pytest.mark.parametrize("n", [7, 2048, 6000])
def test_single_pass_filter_matches_oracle(emu, n):
    """ParticleFilter(reference_max="analytic", single_pass=True): ONE launch per step (pull-resample the previous step
    into the CTA's own slots, gather, propose, score, accumulate masses) + one closing resampling launch.  Same
    ancestors, weights, estimate as the two-launch analytic filter and as the oracle with the same reference."""
    T = 5
    ys = osmc.simulate_lgssm(1, T, 1, LG_A, LG_Q, LG_C, LG_R)[:, 0]
    x0 = np.random.default_rng(n).standard_normal(n).astype(F32)
    obs = gj.C["y"].set(torch.from_numpy(ys))
    one = ParticleFilter(lgssm_step, n, reference_max="analytic", single_pass=True).run(gj.key(17), torch.from_numpy(x0), obs, record=True)
    two = ParticleFilter(lgssm_step, n, reference_max="analytic").run(gj.key(17), torch.from_numpy(x0), obs, record=True)
    assert torch.equal(one.ancestors, two.ancestors)
    assert torch.equal(one.history["log_weights"], two.history["log_weights"])
    assert torch.equal(one.history["state"][0], two.history["state"][0])
    assert torch.equal(one.lse_terms, two.lse_terms) and torch.equal(one.state[0], two.state[0])
    ores = osmc.particle_filter(o_step, orng.key(17), x0, [{"y": F32(y)} for y in ys],
                                m_ref=F32(ParticleFilter(lgssm_step, n).weight_upper_bound(torch.from_numpy(x0), obs)))
    assert one.log_marginal_likelihood.item() == pytest.approx(ores["logz"], abs=2e-4)
    # ping-pong buffers (no history)
    plain = ParticleFilter(lgssm_step, n, reference_max="analytic", single_pass=True).run(gj.key(17), torch.from_numpy(x0), obs)
    assert torch.equal(plain.log_increments, one.log_increments) and torch.equal(plain.state[0], one.state[0])


def test_single_pass_when_every_mass_underflows(emu):
    """An observation so far out that every weight sits more than the 25-nat range of the integer masses below the
    analytic bound (S == 0): both the two-launch and the single-pass filter leave that step unresampled (identity
    ancestors, log-increment -inf) and agree on every later step."""
    n, T = 3000, 4
    ys = np.array([0.2, 90.0, -0.4, 0.1], dtype=F32)  # y_1 = 90: weights near -16 000
    x0 = np.random.default_rng(3).standard_normal(n).astype(F32)
    obs = gj.C["y"].set(torch.from_numpy(ys))
    one = ParticleFilter(lgssm_step, n, reference_max="analytic", single_pass=True).run(gj.key(5), torch.from_numpy(x0), obs, record=True)
    two = ParticleFilter(lgssm_step, n, reference_max="analytic").run(gj.key(5), torch.from_numpy(x0), obs, record=True)
    assert one.lse_terms[1, 1].item() == 0.0 and one.log_increments[1].item() == -np.inf
    assert torch.equal(one.ancestors[1], torch.arange(n, dtype=torch.int32))
    assert torch.equal(one.ancestors, two.ancestors) and torch.equal(one.lse_terms, two.lse_terms)
    assert torch.equal(one.history["log_weights"], two.history["log_weights"]) and torch.equal(one.state[0], two.state[0])

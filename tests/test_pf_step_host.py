"""The single-launch filter step (ParticleFilter(mode="step"): gjb_model_pf_step + gjb_te_resample) on CPU: the
GENERATED pf_step_kernel and the core tile-exponent kernels run as written with real block semantics
(tests/simt_kernels.py) against the oracle filter over the same tile-exponent CDF (oracle/smc.py
particle_filter(masses="tile_exponent")).  Scalar (quad-mapped), vector (lane-group) and integer-state models."""
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


@pytest.fixture
def emu(monkeypatch):
    return abi_emulator.install(monkeypatch, host_kernels=True)


def o_step(h, x_prev):
    x = h.normal("x", F32(LG_A) * x_prev, F32(LG_Q))
    h.normal("y", F32(LG_C) * x, F32(LG_R))
    return x


from pf_step_common import check_against_oracle  # noqa: E402


@pytest.mark.parametrize("n", [7, 2048, 6000])
def test_step_filter_matches_oracle(emu, n):
    T = 5
    ys = osmc.simulate_lgssm(1, T, 1, LG_A, LG_Q, LG_C, LG_R)[:, 0]
    x0 = np.random.default_rng(n).standard_normal(n).astype(F32)
    pf = ParticleFilter(lgssm_step, n, mode="step")
    res = pf.run(gj.key(17), torch.from_numpy(x0), gj.C["y"].set(torch.from_numpy(ys)), record=True)
    check_against_oracle(res, x0, [{"y": F32(y)} for y in ys], o_step, 17, n, T)
    ores = osmc.particle_filter(o_step, orng.key(17), x0, [{"y": F32(y)} for y in ys], masses="tile_exponent")
    assert res.log_marginal_likelihood.item() == pytest.approx(ores["logz"], abs=2e-4)
    # non-record run (ping-pong state / CDF / tile-record buffers, weights written by the last step only)
    res2 = ParticleFilter(lgssm_step, n, mode="step").run(gj.key(17), torch.from_numpy(x0), gj.C["y"].set(torch.from_numpy(ys)))
    assert torch.equal(res2.log_increments, res.log_increments) and torch.equal(res2.state[0], res.state[0])


def test_step_filter_vector_model(emu):
    n, T, d = 3000, 3, 8
    q = np.full(d, LG_Q, dtype=F32)
    r = np.full(d, 0.5 * np.sqrt(d), dtype=F32)
    ys = osmc.simulate_lgssm(2, T, d, LG_A, LG_Q, LG_C, 0.5 * np.sqrt(d))
    x0 = np.random.default_rng(3).standard_normal((n, d)).astype(F32)

    def o_vec(h, x_prev, q_, r_):
        x = h.mv_normal_diag("x", F32(LG_A) * x_prev, q_)
        h.mv_normal_diag("y", F32(LG_C) * x, r_)
        return x

    pf = ParticleFilter(lgssm_step_vec, n, mode="step")
    res = pf.run(gj.key(5), torch.from_numpy(x0), gj.C["y"].set(torch.from_numpy(ys)), (torch.from_numpy(q), torch.from_numpy(r)), record=True)
    check_against_oracle(res, x0, [{"y": y.astype(F32)} for y in ys], o_vec, 5, n, T, shared=(q, r), tol=(1e-5, 2e-6, 2e-5, 2e-4))


def test_step_filter_integer_state(emu):
    from genjax_b200 import workloads as wl

    n, T, K = 2500, 4, 16
    rg = np.random.default_rng(0)
    trans = rg.standard_normal((K, K)).astype(F32)
    obsl = rg.standard_normal((K, K)).astype(F32)
    ys = rg.integers(0, K, T).astype(np.int32)
    z0 = rg.integers(0, K, n).astype(np.int32)

    def o_hmm(h, z_prev, tl, ol):
        z = h.categorical("z", tl[z_prev])
        h.categorical("y", ol[z])
        return z

    pf = ParticleFilter(hmm_step, n, mode="step")
    res = pf.run(gj.key(9), torch.from_numpy(z0), gj.C["y"].set(torch.from_numpy(ys)), (torch.from_numpy(trans), torch.from_numpy(obsl)), record=True)
    check_against_oracle(res, z0, [{"y": np.int32(y)} for y in ys], o_hmm, 9, n, T, shared=(trans, obsl))


@pytest.mark.parametrize("n", [7, 2048])
def test_all_steps_in_one_cooperative_launch_equal_the_per_step_launches(emu, n):
    """mode="steps" (gjb_model_pf_steps: every step of the filter in ONE cooperative launch, a grid barrier per step) ==
    mode="step", bit for bit; on the host a cooperative grid is one block, so up to one tile here (the device test runs
    1 M particles)."""
    T = 5
    ys = osmc.simulate_lgssm(1, T, 1, LG_A, LG_Q, LG_C, LG_R)[:, 0]
    x0 = torch.from_numpy(np.random.default_rng(n).standard_normal(n).astype(F32))
    obs = gj.C["y"].set(torch.from_numpy(ys))
    for record in (True, False):
        a = ParticleFilter(lgssm_step, n, mode="steps").run(gj.key(17), x0, obs, record=record)
        b = ParticleFilter(lgssm_step, n, mode="step").run(gj.key(17), x0, obs, record=record)
        assert torch.equal(a.log_increments, b.log_increments) and torch.equal(a.state[0], b.state[0])
        if record:
            assert torch.equal(a.ancestors, b.ancestors) and torch.equal(a.history["log_weights"], b.history["log_weights"])
            assert torch.equal(a.history["state"][0], b.history["state"][0])


@pytest.mark.parametrize("n", [2048, 6000])
def test_step_filter_table_form_matches_oracle(emu, monkeypatch, n):
    """The multi-GPU form of the step kernel on one device (GJB_STEP_TABLE=1: tile records go through the mailbox, the
    CTA that draws the last ticket builds the prefix / window table, the next launch consumes it -- te_finish_step +
    te_pull_table of csrc/gjb_step.cuh): same ancestors, states, weights and increments as the oracle and as the
    table-free form, bit for bit."""
    monkeypatch.setenv("GJB_STEP_TABLE", "1")
    T = 5
    ys = osmc.simulate_lgssm(1, T, 1, LG_A, LG_Q, LG_C, LG_R)[:, 0]
    x0 = np.random.default_rng(n).standard_normal(n).astype(F32)
    pf = ParticleFilter(lgssm_step, n, mode="step")
    res = pf.run(gj.key(17), torch.from_numpy(x0), gj.C["y"].set(torch.from_numpy(ys)), record=True)
    assert next(iter(pf._plans.values())).te_table
    check_against_oracle(res, x0, [{"y": F32(y)} for y in ys], o_step, 17, n, T)
    monkeypatch.setenv("GJB_STEP_TABLE", "0")
    ref = ParticleFilter(lgssm_step, n, mode="step").run(gj.key(17), torch.from_numpy(x0), gj.C["y"].set(torch.from_numpy(ys)), record=True)
    assert torch.equal(res.ancestors, ref.ancestors) and torch.equal(res.log_increments, ref.log_increments)
    assert torch.equal(res.state[0], ref.state[0])

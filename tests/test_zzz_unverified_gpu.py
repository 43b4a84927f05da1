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


# ------------------------------------------------------------------ Scan combinator (genjax_b200/gen/scan.py)

from oracle import gfi as ogfi  # noqa: E402
from oracle import rng  # noqa: E402

F32 = np.float32
STDS = np.array([2.0, 4.0, 3.0, 5.0, 1.0], dtype=F32)


def _walk(gj):
    @gj.gen
    def walk(x, std):
        nx = gj.normal(x, std) @ "x"
        y = gj.normal(2.0 * nx, 0.5) @ "y"
        return nx, nx + y

    return walk


def _o_walk(h, x, std):
    nx = h.normal("x", x, std)
    y = h.normal("y", (F32(2.0) * nx).astype(F32), F32(0.5))
    return nx, (nx + y).astype(F32)


def _np(t):
    return t.detach().cpu().numpy()


@pytest.mark.parametrize("n", [None, 1, 1000])
def test_scan_simulate_and_importance_match_oracle(device, n):
    """scan.py:199-297 against oracle/gfi.py scan_simulate / scan_generate (same scenarios as tests/test_scan_host.py,
    which runs them through the CPU emulation of the launch)."""
    gj = _gj()
    model = _walk(gj).scan(n=5)
    key = gj.key(314159) if n is None else gj.split(gj.key(314159), n)
    okey = rng.key(314159) if n is None else rng.split(rng.key(314159), n)
    stds = torch.tensor(STDS, device=device)
    tr = model.simulate(key, (0.25, stds))
    otr, ocarry, oys, oscore = ogfi.scan_simulate(_o_walk, okey, F32(0.25), STDS)
    lead = () if n is None else (n,)
    xs = tr.get_choices()[:, "x"]
    assert tuple(xs.shape) == lead + (5,)
    # chained steps: the rounding of step t feeds step t + 1 at scales up to 5, so the absolute tolerance is per chain
    np.testing.assert_allclose(_np(xs).reshape(-1, 5), np.stack([t.choices["x"] for t in otr], 1), rtol=1e-5, atol=2e-5)
    np.testing.assert_allclose(_np(tr.get_score()).reshape(-1), oscore, rtol=1e-5, atol=5e-5)
    carry, ys = tr.get_retval()
    np.testing.assert_allclose(_np(carry).reshape(-1), np.broadcast_to(ocarry, (n or 1,)), rtol=1e-5, atol=2e-5)
    assert tuple(ys.shape) == lead + (5,)
    sub = tr.get_subtrace("y").get_score()
    assert tuple(sub.shape) == lead + (5,)

    yobs = np.array([0.5, -1.0, 2.0, 0.0, 1.0], dtype=F32)
    tr2, w = model.importance(key, gj.C[:, "y"].set(torch.tensor(yobs, device=device)), (0.1, stds))
    _, _, _, osc2, ow = ogfi.scan_generate(_o_walk, okey, lambda t: {"y": yobs[t]}, F32(0.1), STDS)
    np.testing.assert_allclose(_np(w).reshape(-1), ow, rtol=1e-5, atol=5e-5)
    np.testing.assert_allclose(_np(tr2.get_score()).reshape(-1), osc2, rtol=1e-5, atol=5e-5)


def test_scan_lane_consistency_update_regenerate(device):
    gj = _gj()
    model = _walk(gj).scan()
    n = 64
    stds = torch.tensor(STDS, device=device)
    args = (0.3, stds)
    kb = gj.split(gj.key(11), n)
    tr = model.simulate(kb, args)
    one = model.simulate(kb[3], args)
    assert torch.equal(tr.get_choices()[:, "x"][3], one.get_choices()[:, "x"])
    otr, _, _, _ = ogfi.scan_simulate(_o_walk, rng.split(rng.key(11), n), F32(0.3), STDS)

    new, w, _, bwd = model.update(gj.split(gj.key(12), n), tr, gj.C[1, "x"].set(9.0), gj.Diff.no_change(args))
    _, _, _, oscore, ow, _ = ogfi.scan_update(_o_walk, rng.split(rng.key(12), n), otr,
                                              lambda t: {"x": F32(9.0)} if t == 1 else {}, F32(0.3), STDS)
    np.testing.assert_allclose(_np(w), ow, rtol=2e-4, atol=5e-4)
    np.testing.assert_allclose(_np(new.get_score()), oscore, rtol=1e-5, atol=5e-5)
    xs_new, xs_old = new.get_choices()[:, "x"], tr.get_choices()[:, "x"]
    assert (xs_new[:, 1] == 9.0).all() and torch.equal(xs_new[:, [0, 2, 3, 4]], xs_old[:, [0, 2, 3, 4]])
    assert torch.equal(bwd[1, "x"], xs_old[:, 1])
    back, wb, _, _ = model.update(gj.split(gj.key(13), n), new, bwd, gj.Diff.no_change(args))
    torch.testing.assert_close(w + wb, torch.zeros(n, device=device), rtol=0, atol=2e-3)
    assert torch.equal(back.get_choices()[:, "x"], xs_old)

    reg, wr, _, _ = model.edit(gj.split(gj.key(14), n), tr, gj.Regenerate(gj.S["x"]), gj.Diff.no_change(args))
    oreg, _, _, _, owr = ogfi.scan_regenerate(_o_walk, rng.split(rng.key(14), n), otr, {"x"}, F32(0.3), STDS)
    np.testing.assert_allclose(_np(wr), owr, rtol=2e-4, atol=5e-4)
    np.testing.assert_allclose(_np(reg.get_choices()[:, "x"]), np.stack([t.choices["x"] for t in oreg], 1), rtol=1e-5, atol=2e-5)
    assert torch.equal(reg.get_choices()[:, "y"], tr.get_choices()[:, "y"])

    chm = gj.vmap(lambda c: c, in_axes=0)(tr.get_choices())
    score, _ = model.assess(chm, args)
    torch.testing.assert_close(score, tr.get_score(), rtol=1e-5, atol=5e-5)


def test_iterate_and_accumulate(device):
    gj = _gj()

    @gj.gen
    def step(x):
        return gj.normal(x, 1.0) @ "z"

    it = step.iterate(n=10)
    tr, w = it.importance(gj.key(314159), gj.C[3, "z"].set(0.5), (0.01,))
    zs = tr.get_choices()[:, "z"]
    assert zs[3] == 0.5
    assert w.item() == pytest.approx(float(od.normal_logpdf(F32(0.5), F32(zs[2].item()), F32(1.0))), rel=1e-5, abs=2e-5)
    out = tr.get_retval()
    assert out.shape == (11,) and torch.equal(out[1:], zs)

    @gj.gen
    def add(acc, x):
        return acc + x + 0.0 * (gj.normal(0.0, 1.0) @ "eps")

    res = add.accumulate().simulate(gj.key(0), (0.0, torch.ones(4))).get_retval()
    assert torch.equal(res.cpu(), torch.tensor([0.0, 1.0, 2.0, 3.0, 4.0]))
    assert add.reduce().simulate(gj.key(0), (0.0, torch.ones(10))).get_retval().item() == 10.0


# ------------------------------------------------------------------ Vmap / repeat (genjax_b200/gen/vmap_combinator.py)


def test_vmap_combinator_and_repeat(device):
    """vmap.py:180-275, 363-375 and repeat.py:25-41 (scenarios of tests/test_vmap_host.py on the device)."""
    gj = _gj()

    @gj.gen
    def kernel(x):
        z = gj.normal(x, 1.0) @ "z"
        return z

    def o_kernel(h, x):
        return h.normal("z", x, F32(1.0))

    def lp(v, mu):
        return float(od.normal_logpdf(F32(v), F32(mu), F32(1.0)))

    model = gj.vmap(in_axes=(0,))(kernel)
    over = torch.arange(0, 50, dtype=torch.float32, device=device)
    tr = model.simulate(gj.key(314159), (over,))
    otr = ogfi.simulate(o_kernel, rng.split(rng.key(314159), 50), (np.arange(50, dtype=F32),))
    np.testing.assert_allclose(_np(tr.get_choices()[:, "z"]), otr.choices["z"], rtol=1e-5, atol=2e-6)
    assert tr.get_score().shape == () and tr.get_score().item() == pytest.approx(float(otr.get_score().sum()), rel=1e-5)
    assert tr.project(gj.key(1), gj.Selection.all()).item() == pytest.approx(tr.get_score().item(), rel=1e-5)
    score, ret = model.assess(tr.get_choices(), (over,))
    assert score.item() == pytest.approx(tr.get_score().item(), rel=1e-5) and torch.equal(ret, tr.get_retval())

    over3 = torch.arange(0, 3, dtype=torch.float32, device=device)
    _, w = model.importance(gj.key(314159), gj.C[:, "z"].set(torch.tensor([3.0, 2.0, 3.0], device=device)), (over3,))
    assert w.item() == pytest.approx(lp(3.0, 0.0) + lp(2.0, 1.0) + lp(3.0, 2.0), rel=1e-5)
    tr1, w1 = model.importance(gj.key(1), gj.C[1, "z"].set(5.0), (over3,))  # three launches: lanes [0,1), {1}, [2,3)
    assert tr1.get_choices()[1, "z"] == 5.0 and w1.item() == pytest.approx(lp(5.0, 1.0), rel=1e-5)
    free = model.simulate(gj.key(1), (over3,)).get_choices()[:, "z"]
    assert torch.equal(tr1.get_choices()[:, "z"][[0, 2]], free[[0, 2]])

    old = tr1.get_choices()[:, "z"]
    new, wu, _, bwd = model.update(gj.key(6), tr1, gj.C[2, "z"].set(1.0), gj.Diff.no_change((over3,)))
    assert new.get_choices()[2, "z"] == 1.0 and bwd[2, "z"] == old[2]
    assert wu.item() == pytest.approx(lp(1.0, 2.0) - lp(old[2].item(), 2.0), rel=1e-4, abs=2e-5)

    rep = kernel.repeat(n=3).simulate(gj.key(314159), (0.0,))
    vm = kernel.vmap().simulate(gj.key(314159), (torch.zeros(3, device=device),))
    assert rep.get_retval().shape == (3,) and torch.equal(vm.get_choices()[:, "z"], rep.get_choices()[:, "z"])


# ------------------------------------------------------------------ ParticleCollection checkpoints


def test_particle_collection_checkpoint_round_trip(device, tmp_path):
    gj = _gj()
    from genjax_b200.inference.smc import ImportanceK, ParticleCollection

    @gj.gen
    def model(mu):
        x = gj.normal(mu, 2.0) @ "x"
        gj.normal(x, 0.5) @ "y"
        return x

    target = gj.Target(model, (0.3,), gj.C["y"].set(1.0))
    pc = ImportanceK(target, k_particles=500).run_smc(gj.key(5))
    path = tmp_path / "pc.pt"
    pc.save(path)
    back = ParticleCollection.load(path, model, (0.3,))
    assert torch.equal(back.get_log_weights(), pc.get_log_weights())
    assert back.get_particles().get_choices() == pc.get_particles().get_choices()
    torch.testing.assert_close(back.get_particles().get_score(), pc.get_particles().get_score(), rtol=1e-6, atol=1e-6)
    assert back.get_log_marginal_likelihood_estimate() == pc.get_log_marginal_likelihood_estimate()
    sd = pc.state_dict()
    assert sd["diagnostics"]["ess"] == pytest.approx(pc.effective_sample_size().item())
    assert 1.0 <= sd["diagnostics"]["ess"] <= 500.0
    # the same key picks the same particle from the restored collection
    assert back.sample_particle(gj.key(9)).get_choices() == pc.sample_particle(gj.key(9)).get_choices()
    with pytest.raises(ValueError):
        ParticleCollection.load_state_dict(model, (0.3,), {"format": "something else"})


# ------------------------------------------------------------------ analytic weight bound (gen/bounds.py)


def test_filter_maxima_stay_below_the_analytic_bound(device):
    gj = _gj()
    from genjax_b200.inference.pf import ParticleFilter
    from genjax_b200.workloads import lgssm_step
    from oracle import smc as osmc

    n, T = 20_000, 12
    ys = torch.from_numpy(osmc.simulate_lgssm(0, T, 1, 0.9, 1.0, 1.0, 0.5)[:, 0])
    x0 = torch.randn(n, generator=torch.Generator().manual_seed(0))
    pf = ParticleFilter(lgssm_step, n, mode="graph")  # (lse_terms[:, 0] is the running max M in this mode)
    bound = pf.weight_upper_bound(x0, gj.C["y"].set(ys))
    res = pf.run(gj.key(1), x0, gj.C["y"].set(ys))
    m = res.lse_terms[:, 0].cpu().numpy()
    assert (m <= bound + 1e-6).all() and (m > bound - 0.01).all()  # 20 000 particles: some particle sits on the mode


# ------------------------------------------------------------------ reference-maximum filter step


@pytest.mark.parametrize("n", [7, 2048, 100_001])
def test_analytic_reference_filter_matches_oracle(device, n):
    """ParticleFilter(reference_max="analytic"): masses accumulated by model_kernel_static_mass relative to the
    analytic bound, resampler on those masses (same scenario as tests/test_pf_reference_max_host.py)."""
    gj = _gj()
    from genjax_b200.inference.pf import ParticleFilter
    from genjax_b200.workloads import LG_A, LG_C, LG_Q, LG_R, lgssm_step
    from oracle import smc as osmc

    def o_step(h, x_prev):
        x = h.normal("x", F32(LG_A) * x_prev, F32(LG_Q))
        h.normal("y", F32(LG_C) * x, F32(LG_R))
        return x

    T = 6
    ys = osmc.simulate_lgssm(1, T, 1, LG_A, LG_Q, LG_C, LG_R)[:, 0]
    x0 = np.random.default_rng(n).standard_normal(n).astype(F32)
    pf = ParticleFilter(lgssm_step, n, reference_max="analytic")
    bound = pf.weight_upper_bound(torch.from_numpy(x0), gj.C["y"].set(torch.from_numpy(ys)))
    for use_graph in (False, True):
        res = pf.run(gj.key(17), torch.from_numpy(x0), gj.C["y"].set(torch.from_numpy(ys)), record=True, use_graph=use_graph)
        anc, lws, xs = res.ancestors.cpu().numpy(), res.history["log_weights"].cpu().numpy(), res.history["state"][0].cpu().numpy()
        x_in, okey = x0, rng.key(17)
        for t in range(T):
            kp, kr = osmc.pf_step_keys(okey, t)
            otr, ow = ogfi.generate(o_step, rng.split(kp, n), {"y": F32(ys[t])}, (x_in,))
            np.testing.assert_allclose(xs[t], otr.choices["x"], rtol=1e-5, atol=2e-6)
            np.testing.assert_allclose(lws[t], ow, rtol=1e-5, atol=2e-5)
            assert lws[t].max() <= bound + 1e-6
            assert np.array_equal(anc[t], osmc.resample_systematic_pull(lws[t], kr, M=F32(bound)))
            M, S, inc = res.lse_terms[t].cpu().numpy()
            assert M == float(F32(bound))
            assert S == float(int(osmc.det_exp_q((lws[t] - F32(bound)).astype(F32)).sum(dtype=np.uint64)))
            assert inc == pytest.approx(osmc.log_mean_exp(lws[t]), abs=2e-7)
            x_in = xs[t][anc[t]]
        np.testing.assert_array_equal(res.state[0].cpu().numpy(), x_in)
    plain = ParticleFilter(lgssm_step, n, mode="graph").run(gj.key(17), torch.from_numpy(x0), gj.C["y"].set(torch.from_numpy(ys)), record=True)
    assert torch.equal(plain.history["log_weights"][0], res.history["log_weights"][0])
    assert plain.log_increments[0].item() == pytest.approx(res.log_increments[0].item(), abs=2e-7)


@pytest.mark.parametrize("n", [7, 2048, 100_001, 1 << 20])
def test_single_pass_filter_equals_the_two_launch_filter(device, n):
    """ParticleFilter(reference_max="analytic", single_pass=True): model_kernel_static_pull resolves the ancestors of
    its own 2048 slots from the previous step (output-slot resampling), gathers, proposes, scores and accumulates the
    masses in ONE launch per step.  Bit-identical ancestors, weights, states and estimate terms to the two-launch
    analytic filter (which the test above pins to the oracle); same scenario as tests/test_pf_reference_max_host.py."""
    import os

    if os.environ.get("GJB_EMULATE") and n > 10_000:  # both large sizes have passed there once; they take minutes
        pytest.skip("minutes under the SIMT host shim")
    gj = _gj()
    from genjax_b200.inference.pf import ParticleFilter
    from genjax_b200.workloads import LG_A, LG_C, LG_Q, LG_R, lgssm_step
    from oracle import smc as osmc

    T = 6
    ys = torch.from_numpy(osmc.simulate_lgssm(1, T, 1, LG_A, LG_Q, LG_C, LG_R)[:, 0])
    x0 = torch.from_numpy(np.random.default_rng(n % 1000).standard_normal(n).astype(F32))
    obs = gj.C["y"].set(ys)
    two = ParticleFilter(lgssm_step, n, reference_max="analytic").run(gj.key(17), x0, obs, record=True)
    for use_graph in (False, True):
        one = ParticleFilter(lgssm_step, n, reference_max="analytic", single_pass=True).run(gj.key(17), x0, obs, record=True, use_graph=use_graph)
        assert torch.equal(one.ancestors, two.ancestors)
        assert torch.equal(one.history["log_weights"], two.history["log_weights"])
        assert torch.equal(one.history["state"][0], two.history["state"][0])
        assert torch.equal(one.lse_terms, two.lse_terms) and torch.equal(one.state[0], two.state[0])
    plain = ParticleFilter(lgssm_step, n, reference_max="analytic", single_pass=True).run(gj.key(17), x0, obs)
    assert torch.equal(plain.log_increments, two.log_increments) and torch.equal(plain.state[0], two.state[0])
    with pytest.raises(NotImplementedError):  # the tile prefix lives in shared memory: up to 2048 tiles per device
        ParticleFilter(lgssm_step, (1 << 22) + 1, reference_max="analytic", single_pass=True).run(gj.key(0), torch.zeros((1 << 22) + 1), obs)


# ------------------------------------------------------------------ long-tail scalar wrappers

_LONG_TAIL = [("cauchy", (1.0, 2.0)), ("half_cauchy", (1.0, 2.0)), ("laplace", (-1.0, 0.5)), ("log_normal", (0.3, 0.8)),
              ("gumbel", (0.5, 1.5)), ("weibull", (1.7, 2.0)), ("kumaraswamy", (2.0, 3.0)), ("logit_normal", (0.3, 0.8)),
              ("geometric", (0.3,)), ("inverse_gamma", (3.0, 2.0)), ("chi2", (3.5,)), ("chi2", (0.8,)),
              ("student_t", (4.0, 1.0, 2.0)), ("poisson", (3.5,)), ("poisson", (42.0,))]


@pytest.mark.parametrize("name,args", _LONG_TAIL)
def test_long_tail_primitive_sample_and_logpdf_match_oracle(device, name, args):
    """tensorflow_probability/__init__.py:110, 174, 179, 214, 219, 309: dist.simulate over a KeyBatch == the oracle's
    inverse-CDF sampler on the same Philox lanes; score == oracle log-density (same bar as
    tests/test_gfi_gpu.py::test_primitive_sample_and_logpdf_match_oracle)."""
    gj = _gj()
    n = 50_001
    tr = getattr(gj, name).simulate(gj.split(gj.key(11), n), args)
    v = tr.get_retval().cpu().numpy()
    words, idx = rng.lanes(rng.split(rng.key(11), n))
    ov = od.DISTS[name][0](words, idx, 1, *[F32(a) for a in args])
    # tanf / expf of the device against a rounded float64 evaluation; the Cauchy tails amplify an ulp of the argument, an
    # ulp can flip a rejection of the gamma samplers or move a geometric draw across an integer
    assert (~np.isclose(v, ov, rtol=1e-4, atol=1e-5)).mean() < (2e-3 if name in ("inverse_gamma", "chi2", "geometric", "student_t", "poisson") else 1e-4), name
    np.testing.assert_allclose(tr.get_score().cpu().numpy(), od.DISTS[name][1](v, *[F32(a) for a in args]), rtol=2e-5,
                               atol=2e-4 if name == "poisson" else 2e-5)  # k log(rate) - lgamma(k + 1) cancels in float32
    kw = dict(zip(getattr(gj, name).kw_names, args))  # TFP's keyword spelling
    tr2 = getattr(gj, name).simulate(gj.split(gj.key(11), 64), ((), kw))
    assert torch.equal(tr2.get_retval(), tr.get_retval()[:64])


def test_long_tail_sites_inside_a_model(device):
    """All six as sites of one @gen model whose parameters depend on earlier sites: simulate, assess and importance
    agree with the oracle's log-densities; the constrained sites contribute the weight."""
    gj = _gj()

    @gj.gen
    def model(s):
        a = gj.cauchy(0.0, s) @ "a"
        b = gj.half_cauchy(a, 1.5) @ "b"
        c = gj.laplace(a, s) @ "c"
        d = gj.log_normal(0.1 * gj.numpy.tanh(c), 0.5) @ "d"  # bounded: the Cauchy tails of a, c stay out of exp
        e = gj.gumbel(c, d) @ "e"
        f = gj.weibull(1.0 + d, s) @ "f"
        return e + f

    n, s = 4096, 0.7
    tr = model.simulate(gj.split(gj.key(5), n), (s,))
    ch = {k: tr.get_choices()[k].cpu().numpy() for k in "abcdef"}
    want = (od.cauchy_logpdf(ch["a"], F32(0), F32(s)) + od.half_cauchy_logpdf(ch["b"], ch["a"], F32(1.5))
            + od.laplace_logpdf(ch["c"], ch["a"], F32(s)) + od.log_normal_logpdf(ch["d"], F32(0.1) * np.tanh(ch["c"]), F32(0.5))
            + od.gumbel_logpdf(ch["e"], ch["c"], ch["d"]) + od.weibull_logpdf(ch["f"], F32(1) + ch["d"], F32(s)))
    assert np.isfinite(want).all()
    np.testing.assert_allclose(tr.get_score().cpu().numpy(), want, rtol=1e-4, atol=1e-4)
    assert (ch["b"] >= ch["a"]).all() and (ch["d"] > 0).all() and (ch["f"] >= 0).all()
    obs = gj.C["e"].set(0.3).at["f"].set(1.1)
    tr2, w = model.importance(gj.split(gj.key(5), n), obs, (s,))
    c2 = {k: tr2.get_choices()[k].cpu().numpy() for k in "abcd"}
    for k in "abcd":  # the unconstrained sites draw what simulate drew on the same lanes
        np.testing.assert_array_equal(c2[k], ch[k])
    ww = od.gumbel_logpdf(F32(0.3), c2["c"], c2["d"]) + od.weibull_logpdf(F32(1.1), F32(1) + c2["d"], F32(s))
    assert not np.isnan(ww).any()  # -inf where the Cauchy tail of c pushes the observed e out of the Gumbel's reach
    np.testing.assert_allclose(w.cpu().numpy(), ww, rtol=1e-4, atol=1e-4)


def test_second_slice_sites_inside_a_model(device):
    """kumaraswamy, logit_normal, geometric, inverse_gamma, chi2 as sites of one model (quad-block and lane-stream
    samplers mixed): scores and importance weights equal the oracle's log-densities."""
    gj = _gj()

    @gj.gen
    def model(s):
        a = gj.kumaraswamy(2.0, s) @ "a"
        b = gj.logit_normal(a, 0.5) @ "b"
        k = gj.geometric(0.1 + 0.8 * b) @ "k"
        g = gj.inverse_gamma(2.0 + k, s) @ "g"
        c = gj.chi2(1.0 + a) @ "c"
        t = gj.student_t(2.0 + c, a, s) @ "t"
        p = gj.poisson(0.5 + 20.0 * b) @ "p"  # rates on both sides of the inversion / PTRS switch
        return g + c + t + p

    n, s = 4096, 1.5
    tr = model.simulate(gj.split(gj.key(6), n), (s,))
    ch = {k: tr.get_choices()[k].cpu().numpy() for k in "abkgctp"}
    assert ((ch["a"] > 0) & (ch["a"] < 1)).all() and ((ch["b"] > 0) & (ch["b"] < 1)).all()
    assert (ch["k"] >= 0).all() and (ch["k"] == np.floor(ch["k"])).all() and (ch["g"] > 0).all() and (ch["c"] > 0).all()
    want = (od.kumaraswamy_logpdf(ch["a"], F32(2), F32(s)) + od.logit_normal_logpdf(ch["b"], ch["a"], F32(0.5))
            + od.geometric_logpdf(ch["k"], F32(0.1) + F32(0.8) * ch["b"]) + od.inverse_gamma_logpdf(ch["g"], F32(2) + ch["k"], F32(s))
            + od.chi2_logpdf(ch["c"], F32(1) + ch["a"]) + od.student_t_logpdf(ch["t"], F32(2) + ch["c"], ch["a"], F32(s))
            + od.poisson_logpdf(ch["p"], F32(0.5) + F32(20) * ch["b"]))
    assert (ch["p"] >= 0).all() and (ch["p"] == np.floor(ch["p"])).all()
    assert np.isfinite(want).all()
    np.testing.assert_allclose(tr.get_score().cpu().numpy(), want, rtol=1e-4, atol=1e-4)
    obs = gj.C["g"].set(0.9).at["c"].set(2.5)
    tr2, w = model.importance(gj.split(gj.key(6), n), obs, (s,))
    c2 = {k: tr2.get_choices()[k].cpu().numpy() for k in "abk"}
    for k in "abk":
        np.testing.assert_array_equal(c2[k], ch[k])
    ww = od.inverse_gamma_logpdf(F32(0.9), F32(2) + c2["k"], F32(s)) + od.chi2_logpdf(F32(2.5), F32(1) + c2["a"])
    np.testing.assert_allclose(w.cpu().numpy(), ww, rtol=1e-4, atol=1e-4)

"""Shared checker of the single-launch filter step tests (CPU host run and -m gpu): teacher-forced comparison of a
recorded ParticleFilter(mode="step") run with the oracle over the same tile-exponent CDF."""
import numpy as np
import pytest

from oracle import gfi as ogfi
from oracle import rng as orng
from oracle import smc as osmc

F32 = np.float32


def check_against_oracle(res, x0, obs_list, o_model, key, n, T, shared=(), tol=(1e-5, 2e-6, 1e-5, 2e-5), steps=None, lme_tol=5e-7):
    anc, lws, xs = res.ancestors.cpu().numpy(), res.history["log_weights"].cpu().numpy(), res.history["state"][0].cpu().numpy()
    lse = res.lse_terms.cpu().numpy()
    x_in, okey = x0, orng.key(key)
    for t in range(T):
        if steps is None or t in steps:
            kp, kr = osmc.pf_step_keys(okey, t)
            otr, ow = ogfi.generate(o_model, orng.split(kp, n), obs_list[t], (x_in,) + tuple(shared))
            ox = otr.retval if not isinstance(otr.retval, tuple) else otr.retval[0]
            if xs.dtype == np.int32:
                np.testing.assert_array_equal(xs[t], ox)
            else:
                np.testing.assert_allclose(xs[t], ox, rtol=tol[0], atol=tol[1])
            np.testing.assert_allclose(lws[t], ow, rtol=tol[2], atol=tol[3])
            # ancestors and masses: exact, given the kernel's own weights
            assert np.array_equal(anc[t], osmc.resample_systematic_te(lws[t], kr)), f"ancestors differ at step {t}"
            E_ln2, S, inc = lse[t]
            _, S_o, e_o = osmc.te_cdf(lws[t])
            assert S == float(S_o) and E_ln2 == pytest.approx(e_o * np.log(2.0), abs=1e-12)
            assert inc == pytest.approx(osmc.te_log_mean_exp(lws[t]), abs=1e-11)
            assert inc == pytest.approx(osmc.log_mean_exp(lws[t]), abs=lme_tol)  # same estimate as the exact-max pipeline
        x_in = xs[t][anc[t]]  # teacher forcing: continue from the kernel's own state
    np.testing.assert_array_equal(res.state[0].cpu().numpy(), x_in)

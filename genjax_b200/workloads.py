"""The BASELINE.json configurations written against the public API
(SURVEY.md section 8d).  Used by bench.py, __graft_entry__ and the tests so
that every one of them exercises exactly the same captured models."""

from __future__ import annotations

from .gen.capture import ArgSpec
from .gen.distributions import beta, categorical, flip, mv_normal_diag, normal
from .gen.static import gen

# linear-Gaussian state-space model (config 2): std-devs q, r
LG_A, LG_Q, LG_C, LG_R = 0.9, 1.0, 1.0, 0.5


@gen
def lgssm_step(x_prev):
    """x_t ~ N(a x_{t-1}, q);  y_t ~ N(c x_t, r)   (scalar state, shape 2a)."""
    x = normal(LG_A * x_prev, LG_Q) @ "x"
    normal(LG_C * x, LG_R) @ "y"
    return x


@gen
def lgssm_step_vec(x_prev, q, r):
    """Diagonal d-dimensional LGSSM (shape 2b): mv_normal_diag sites."""
    x = mv_normal_diag(LG_A * x_prev, q) @ "x"
    mv_normal_diag(LG_C * x, r) @ "y"
    return x


@gen
def beta_bernoulli(alpha, beta_):
    """README quickstart model (config 1; reference README.md:88-93)."""
    p = beta(alpha, beta_) @ "p"
    v = flip(p) @ "v"
    return v


@gen
def hmm_step(z_prev, trans_logits, obs_logits):
    """16-state HMM kernel (config 4; template exact_testbed.py:61-68)."""
    z = categorical(logits=trans_logits[z_prev]) @ "z"
    categorical(logits=obs_logits[z]) @ "y"
    return z


def prebuild_all():
    """Compile every workload kernel for sm_100a (no GPU needed)."""
    out = {}
    for obs in (None, ("y",)):  # generic GFI variant + the bootstrap-filter variant (flags baked into pf_kernel)
        tag = "" if obs is None else "_pf"
        out["lgssm_step" + tag] = lgssm_step.prebuild([ArgSpec("particle", "f32", ())], pf_obs=obs)
        for d in (8, 32):
            out[f"lgssm_step_vec{d}" + tag] = lgssm_step_vec.prebuild(
                [ArgSpec("particle", "f32", (d,)), ArgSpec("shared", "f32", (d,)), ArgSpec("shared", "f32", (d,))], pf_obs=obs
            )
    out["hmm_step_pf"] = hmm_step.prebuild(
        [ArgSpec("particle", "i32", ()), ArgSpec("shared", "f32", (16, 16)), ArgSpec("shared", "f32", (16, 16))], pf_obs=("y",)
    )
    out["beta_bernoulli"] = beta_bernoulli.prebuild([ArgSpec("scalar", "f32", ()), ArgSpec("scalar", "f32", ())])
    out["hmm_step"] = hmm_step.prebuild(
        [ArgSpec("particle", "i32", ()), ArgSpec("shared", "f32", (16, 16)), ArgSpec("shared", "f32", (16, 16))]
    )
    return out

"""The BASELINE.json configurations written against the public API
(SURVEY.md section 8d).  Used by bench.py, __graft_entry__ and the tests so
that every one of them exercises exactly the same captured models."""

from __future__ import annotations

from .gen.capture import ArgSpec
from .gen import numpy_api as jnp
from .gen.distributions import beta, categorical, flip, gmm_diag, mv_normal_diag, normal
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


@gen
def gmm_target(logits, mu, sigma):
    """8-component 8-D Gaussian-mixture target (config 3): ONE site whose primitive is the mixture."""
    return gmm_diag(logits, mu, sigma) @ "x"


@gen
def eight_schools(sigma):
    """Hierarchical-normal (8-schools-style) model (config 5): all-Normal sites."""
    mu = normal(0.0, 5.0) @ "mu"
    log_tau = normal(0.0, 1.0) @ "log_tau"
    theta = normal.repeat(n=8)(mu, jnp.exp(log_tau)) @ "theta"
    mv_normal_diag(theta, sigma) @ "y"
    return theta


EIGHT_SCHOOLS_Y = [28.0, 8.0, -3.0, 7.0, -1.0, 1.0, 18.0, 12.0]
EIGHT_SCHOOLS_SIGMA = [15.0, 10.0, 16.0, 11.0, 9.0, 11.0, 10.0, 18.0]


def prebuild_chain_models():
    """Chain-kernel variants of configs 3 and 5 (no GPU needed)."""
    from .gen import capture as cap
    from .gen.codegen_chain import ChainSpec
    from .gen.static import compile_ir

    out = {}
    K, D = 8, 8
    specs = [ArgSpec("shared", "f32", (K,)), ArgSpec("shared", "f32", (K, D)), ArgSpec("shared", "f32", (K,))]
    tree = ("tuple", [("leaf", i) for i in range(3)])
    ir = cap.capture(gmm_target.source, "gmm_target", specs, tree)
    out["gmm_target_mh"] = compile_ir(ir, chain=ChainSpec((0,), (None,)))
    ir = cap.capture(eight_schools.source, "eight_schools", [ArgSpec("shared", "f32", (8,))], ("tuple", [("leaf", 0)]))
    out["eight_schools_hmc"] = compile_ir(ir, chain=ChainSpec((0, 1, 2), (None, None, None)))
    return out


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
    out.update(prebuild_chain_models())
    return out

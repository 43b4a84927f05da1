#!/usr/bin/env python
"""Reference-side CPU baseline: BASELINE.json configs[0..2] written against GenJAX's OWN API on jax[cpu].

This file is what a maintainer with ``jax`` / ``tensorflow_probability`` / ``genjax`` installed would run to put the
reference's number beside ours.  It imports nothing from this repository and nothing in this repository's product
path imports it.  In this image it CANNOT run (``jax``, ``jaxlib``, ``tensorflow_probability``, ``penzai`` are absent
and there is no network; ``pip install --no-index --find-links /opt/wheelhouse /root/reference`` stops at the missing
``poetry_dynamic_versioning`` build backend) -- ``bench.py --impl reference`` therefore times the oracle port and
says so (``cpu_baseline.kind = "port"``).  ``bench.py`` calls ``measure()`` below only when ``import jax, genjax``
succeeds (e.g. a future image with ``baseline/_ref`` populated).

The programs follow the reference's documented idioms:
  * configs[0]  README.md:81-119                     (beta-bernoulli ImportanceK, k = 50, 50 trials)
  * configs[1]  docs/cookbook/inactive/inference/importance_sampling.ipynb cell 16 and mapping_tutorial.ipynb
                cell 37 (vmapped ``step.importance`` + log-sum-exp + resample + gather per time step; the reference
                ships no resampler, so systematic resampling is written with ``jnp.cumsum`` / ``jnp.searchsorted``)
  * configs[2]  tests/inference/test_requests.py:168-193 (``Rejuvenate`` + the accept idiom), vmapped over chains

    python baseline/run_genjax_cpu.py --config 1 --particles 1048576 --T 100 --repeats 3
"""

from __future__ import annotations

import argparse
import json
import os
import time


def _imports():
    import jax  # noqa: F401  (ImportError here is the "not installable" signal bench.py looks for)
    import genjax  # noqa: F401

    return jax, genjax


# ------------------------------------------------------------------ configs[0]


def beta_bernoulli_quickstart():
    """README.md:81-119, verbatim structure; returns (estimate | v=True, estimate | v=False)."""
    jax, genjax = _imports()
    import jax.numpy as jnp
    from genjax import ChoiceMap, Target, beta, flip, gen
    from genjax.inference.smc import ImportanceK

    @gen
    def beta_bernoulli(a, b):
        p = beta(a, b) @ "p"
        v = flip(p) @ "v"
        return v

    @jax.jit
    def run_inference(obs):
        target = Target(beta_bernoulli, (2.0, 2.0), ChoiceMap.d({"v": obs}))
        alg = ImportanceK(target, k_particles=50)
        sub_keys = jax.random.split(jax.random.key(314159), 50)
        _, p_chm = jax.vmap(alg.random_weighted, in_axes=(0, None))(sub_keys, target)
        return jnp.mean(p_chm["p"])

    return float(run_inference(True)), float(run_inference(False))


# ------------------------------------------------------------------ configs[1]

LG_A, LG_Q, LG_C, LG_R = 0.9, 1.0, 1.0, 0.5


def make_lgssm_filter(n: int, d: int = 1):
    """Bootstrap filter over ``n`` particles: returns a jitted ``run(key, x0 [n(,d)], ys [T(,d)]) -> logZ``."""
    jax, genjax = _imports()
    import jax.numpy as jnp
    from genjax import ChoiceMapBuilder as C
    from genjax import gen, mv_normal_diag, normal
    from jax.scipy.special import logsumexp

    if d == 1:

        @gen
        def step(x_prev):
            x = normal(LG_A * x_prev, LG_Q) @ "x"
            _ = normal(LG_C * x, LG_R) @ "y"
            return x

    else:

        @gen
        def step(x_prev):
            x = mv_normal_diag(LG_A * x_prev, LG_Q * jnp.ones(d)) @ "x"
            _ = mv_normal_diag(LG_C * x, LG_R * jnp.ones(d)) @ "y"
            return x

    def one_step(carry, inp):
        x, logz = carry
        key_t, y = inp
        k_prop, k_res = jax.random.split(key_t)
        keys = jax.random.split(k_prop, n)
        tr, w = jax.vmap(step.importance, in_axes=(0, None, (0,)))(keys, C["y"].set(y), (x,))
        lse = logsumexp(w)
        # systematic resampling: offspring j takes the particle whose CDF interval holds (j + u) / n
        cdf = jnp.cumsum(jnp.exp(w - lse))
        u = (jnp.arange(n) + jax.random.uniform(k_res)) / n
        anc = jnp.clip(jnp.searchsorted(cdf, u), 0, n - 1)
        x_new = tr.get_retval()[anc]
        return (x_new, logz + lse - jnp.log(n)), None

    @jax.jit
    def run(key, x0, ys):
        keys = jax.random.split(key, ys.shape[0])
        (x, logz), _ = jax.lax.scan(one_step, (x0, jnp.float32(0.0)), (keys, ys))
        return logz, x

    return run


# ------------------------------------------------------------------ configs[2]


def make_gmm_mh(n_chains: int, n_steps: int, k: int = 8, d: int = 8, step_size: float = 0.5):
    """8-component 8-D diagonal Gaussian mixture, random-walk ``Rejuvenate`` + accept, vmapped over chains."""
    jax, genjax = _imports()
    import jax.numpy as jnp
    import jax.tree_util as jtu
    from genjax import ChoiceMapBuilder as C
    from genjax import ExactDensity, Pytree, StaticRequest, gen, mv_normal_diag
    from genjax.inference.requests import Rejuvenate
    from jax.scipy.special import logsumexp

    mu = jax.random.uniform(jax.random.key(1), (k, d), minval=-4.0, maxval=4.0)
    sigma = 0.7

    @Pytree.dataclass
    class GaussianMixture(ExactDensity):  # docs/cookbook/inactive/expressivity/custom_distribution.ipynb cell 9
        def sample(self, key, mu, sigma):
            k1, k2 = jax.random.split(key)
            comp = jax.random.randint(k1, (), 0, mu.shape[0])
            return mu[comp] + sigma * jax.random.normal(k2, (mu.shape[1],))

        def logpdf(self, x, mu, sigma):
            z = (x[None, :] - mu) / sigma
            comp = -0.5 * jnp.sum(z * z, axis=1) - mu.shape[1] * (jnp.log(sigma) + 0.5 * jnp.log(2 * jnp.pi))
            return logsumexp(comp) - jnp.log(mu.shape[0])

    gmm = GaussianMixture()

    @gen
    def model():
        return gmm(mu, sigma) @ "x"

    request = StaticRequest({"x": Rejuvenate(mv_normal_diag, lambda chm: (chm.get_value(), step_size * jnp.ones(d)))})

    def chain(key, x0):
        tr, _ = model.importance(key, C["x"].set(x0), ())

        def body(tr, key_t):
            k1, k2 = jax.random.split(key_t)
            new_tr, w, _, _ = request.edit(k1, tr, ())
            check = jnp.log(jax.random.uniform(k2)) < w
            tr = jtu.tree_map(lambda a, b: jnp.where(check, a, b), new_tr, tr)
            return tr, check

        tr, acc = jax.lax.scan(body, tr, jax.random.split(key, n_steps))
        return tr.get_choices()["x"], jnp.mean(acc)

    @jax.jit
    def run(key, x0):
        return jax.vmap(chain)(jax.random.split(key, n_chains), x0)

    return run


# ------------------------------------------------------------------ timing


def measure(config: int, particles: int, T: int, d: int = 1, repeats: int = 3) -> dict:
    """Units/s of one config on jax[cpu]: best of ``repeats`` after one compile+warm-up run."""
    jax, genjax = _imports()
    import jax.numpy as jnp
    import numpy as np

    jax.config.update("jax_platform_name", "cpu")
    g = np.random.default_rng(0)
    if config == 0:
        t0 = time.perf_counter()
        out = beta_bernoulli_quickstart()
        return {"config": 0, "estimates": out, "seconds": time.perf_counter() - t0, "readme": [0.6039314, 0.3679334]}
    if config == 1:
        run = make_lgssm_filter(particles, d)
        x0 = jnp.asarray(g.standard_normal(particles if d == 1 else (particles, d)), jnp.float32)
        ys = jnp.asarray(g.standard_normal(T if d == 1 else (T, d)), jnp.float32)
        args, units, unit = (jax.random.key(314159), x0, ys), particles * T, "particle-steps/s"
    elif config == 2:
        run = make_gmm_mh(particles, T)
        x0 = jnp.asarray(3.0 * g.standard_normal((particles, 8)), jnp.float32)
        args, units, unit = (jax.random.key(2), x0), particles * T, "chain-steps/s"
    else:
        raise SystemExit("configs 0, 1, 2 are written here; 3 and 4 are the same two programs at other sizes / HMC.edit")
    jax.block_until_ready(run(*args))
    best = float("inf")
    for _ in range(repeats):
        t0 = time.perf_counter()
        jax.block_until_ready(run(*args))
        best = min(best, time.perf_counter() - t0)
    return {"config": config, "value": units / best, "unit": unit, "seconds": best, "cores": os.cpu_count(),
            "jax": jax.__version__, "genjax": getattr(genjax, "__version__", "?")}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, default=1)
    ap.add_argument("--particles", type=int, default=1 << 20)
    ap.add_argument("--T", type=int, default=100)
    ap.add_argument("--dim", type=int, default=1)
    ap.add_argument("--repeats", type=int, default=3)
    a = ap.parse_args()
    try:
        print(json.dumps(measure(a.config, a.particles, a.T, a.dim, a.repeats)))
    except ImportError as e:
        print(json.dumps({"unavailable": f"{type(e).__name__}: {e}"}))


if __name__ == "__main__":
    main()

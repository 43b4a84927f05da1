"""Sequential Monte Carlo: particle collections, importance sampling, target
changes.

API mirror of src/genjax/_src/inference/smc.py: ``ParticleCollection:77``
(``get_log_marginal_likelihood_estimate:96-97``, ``sample_particle:102-109``),
``SMCAlgorithm:117`` (``random_weighted:162-179``, ``estimate_logpdf:181-198``),
``Importance:234``, ``ImportanceK:283`` (``run_smc:298-315``),
``ChangeTarget:360`` (``_reweight:378-384``).  The reference vmaps
``target.importance`` over K split keys; here that is ONE fused kernel launch
over the lanes of a ``KeyBatch``.
"""

from __future__ import annotations

import torch

from ..core.choice_map import ChoiceMap
from ..core.key import PRNGKey, key_children, split
from ..gen.static import Batched, StaticTrace, _rebatch
from ..runtime import smc_ops
from .sp import Algorithm, SampleDistribution, Target


class ParticleCollection:
    """Weighted particles (smc.py:77-109): a batched trace + log-weights."""

    def __init__(self, particles: StaticTrace, log_weights: torch.Tensor, is_valid=True):
        self.particles = particles
        self.log_weights = log_weights
        self.is_valid = is_valid
        self._ws = None

    def _workspace(self) -> smc_ops.WeightWorkspace:
        if self._ws is None:
            self._ws = smc_ops.WeightWorkspace(self.log_weights.numel(), self.log_weights.device)
        return self._ws

    def get_particles(self) -> StaticTrace:
        return self.particles

    def get_particle(self, idx) -> StaticTrace:
        if isinstance(idx, torch.Tensor) and idx.ndim == 0:
            idx = idx.reshape(1)
            tr = self.particles.take(idx)
            tr.batched = False
            return tr
        return self.particles.take(idx)

    def get_log_weights(self) -> torch.Tensor:
        return self.log_weights

    def __len__(self):
        return self.log_weights.numel()

    def __getitem__(self, idx):
        return self.get_particle(idx), self.log_weights[idx]

    def lse_terms(self) -> torch.Tensor:
        """Device float64 [M, S, log-mean-exp] from the exact integer mass."""
        return self._workspace().lse_terms(self.log_weights)

    def get_log_marginal_likelihood_estimate(self) -> torch.Tensor:
        """``logsumexp(lw) - log K`` (smc.py:96-97)."""
        return self.lse_terms()[2].to(torch.float32)

    def check_valid(self) -> bool:
        """False when every weight is zero / NaN (host sync)."""
        return bool(self.lse_terms()[1].item() > 0)

    def effective_sample_size(self) -> torch.Tensor:
        """(sum w)^2 / sum w^2 with w = exp(lw - M): one small kernel, fp64 accumulation in a fixed order."""
        return smc_ops.weight_ess(self.log_weights, self.lse_terms())

    # -- checkpointing (the on-disk side of a resumable run) ----
    def state_dict(self) -> dict:
        """Everything needed to rebuild this collection, as CPU tensors and plain Python values: the log-weights
        and, per address, the particles' choices.  The trace itself (score, return value, per-site log-densities)
        is NOT stored: ``load_state_dict`` recomputes it from the choices with one assess-mode launch of the model
        kernel, so a checkpoint stays valid across kernel / layout changes."""
        tr = self.particles
        choices = {}
        for s in tr.cm.ir.sites:
            v = tr.values[s.index]
            choices["/".join(str(a) for a in s.addr)] = {"value": v.detach().cpu(), "shared": bool(tr.bcast[s.index])}
        lse = self.lse_terms().detach().cpu()
        return {
            "format": "genjax_b200.ParticleCollection/1",
            "n": int(self.log_weights.numel()),
            "model": tr.get_gen_fn().__name__,
            "addresses": [list(s.addr) for s in tr.cm.ir.sites],
            "log_weights": self.log_weights.detach().cpu(),
            "choices": choices,
            "is_valid": bool(self.is_valid) if not isinstance(self.is_valid, torch.Tensor) else bool(self.is_valid.item()),
            "diagnostics": {"log_marginal_likelihood": float(lse[2]), "ess": float(self.effective_sample_size().item())},
        }

    @staticmethod
    def load_state_dict(gen_fn, args: tuple, state: dict) -> "ParticleCollection":
        """Rebuild a collection saved by ``state_dict`` for the model ``gen_fn`` called with ``args`` (per-particle
        arguments marked as in the original run, e.g. ``gj.Batched(x_prev)``)."""
        from ..runtime import cabi

        if state.get("format") != "genjax_b200.ParticleCollection/1":
            raise ValueError("not a genjax_b200 ParticleCollection checkpoint")
        device = cabi.require_cuda()
        n = int(state["n"])
        chm = ChoiceMap.empty()
        for addr in state["addresses"]:
            entry = state["choices"]["/".join(str(a) for a in addr)]
            v = entry["value"].to(device)
            chm = chm | ChoiceMap.entry(v if entry["shared"] else Batched(v), *addr)
        tr, _ = gen_fn._run(None, args, chm, weight_mode="none", n=n, batched=True)
        return ParticleCollection(tr, state["log_weights"].to(device), state["is_valid"])

    def save(self, path) -> None:
        torch.save(self.state_dict(), path)

    @staticmethod
    def load(path, gen_fn, args: tuple) -> "ParticleCollection":
        return ParticleCollection.load_state_dict(gen_fn, args, torch.load(path, map_location="cpu", weights_only=True))

    def sample_particle_index(self, key: PRNGKey) -> torch.Tensor:
        """One categorical draw over the normalised weights (smc.py:105-108)."""
        ws = self._workspace()
        ws.lse_terms(self.log_weights)
        n = self.log_weights.numel()
        cdf = torch.empty(n, dtype=torch.int64, device=self.log_weights.device)
        anc = torch.empty(1, dtype=torch.int32, device=self.log_weights.device)
        ws.multinomial(self.log_weights, key.words, key.index, anc, cdf)
        return anc

    def sample_particle(self, key: PRNGKey) -> StaticTrace:
        idx = self.sample_particle_index(key)
        tr = self.particles.take(idx.long())
        tr.batched = False
        return tr

    def resample(self, key: PRNGKey, method: str = "systematic") -> tuple["ParticleCollection", torch.Tensor]:
        """Resampled collection with uniform weights log-mean-exp(lw) and the ancestors."""
        ws = self._workspace()
        terms = ws.lse_terms(self.log_weights)
        n = self.log_weights.numel()
        anc = torch.empty(n, dtype=torch.int32, device=self.log_weights.device)
        if method == "systematic":
            ws.systematic(self.log_weights, key, anc)
        elif method == "multinomial":
            kb = split(key, n)
            cdf = torch.empty(n, dtype=torch.int64, device=self.log_weights.device)
            ws.multinomial(self.log_weights, kb.words, kb.offset, anc, cdf)
        else:
            raise ValueError(method)
        tr = self.particles.take(anc.long())
        lw = terms[2].to(torch.float32).expand(n).contiguous()
        return ParticleCollection(tr, lw, self.is_valid), anc


def _concat_traces(a: StaticTrace, b: StaticTrace) -> StaticTrace:
    """``tree_map(stack_to_first_dim, a, b)`` (smc.py:317-351): the particles of ``a`` followed by those of ``b``."""
    assert a.cm is b.cm, "traces of different models"
    ir = a.cm.ir

    def rows(tr, j, t):
        ev = tuple(ir.sites[j].value.shape)
        if tr.bcast[j] or t.ndim == len(ev):
            return t.reshape((1,) + ev).expand((tr.n,) + ev)
        return t.reshape((tr.n,) + ev)

    values = {j: torch.cat([rows(a, j, a.values[j]), rows(b, j, b.values[j])]).contiguous() for j in a.values}
    rets = []
    for ra, rb in zip(a.ret_leaves, b.ret_leaves):
        if isinstance(ra, torch.Tensor):
            rets.append(torch.cat([ra.reshape((a.n,) + tuple(ra.shape[1:])), rb.reshape((b.n,) + tuple(rb.shape[1:]))]))
        else:
            rets.append(ra)
    score = torch.cat([a.score.reshape(a.n), b.score.reshape(b.n)])
    return StaticTrace(a.gen_fn, a.cm, a.bound if a.bound is not None else b.bound, a.args, a.n + b.n, True, values, score,
                       rets, {j: False for j in values})


def _stack_chm(batch: ChoiceMap, n: int, one: ChoiceMap) -> ChoiceMap:
    """Batched choice map (n lanes) followed by one more lane holding ``one``."""
    out = ChoiceMap.empty()
    single = dict(one.leaves())
    for addr, v in batch.leaves():
        v = v.value if isinstance(v, Batched) else v
        v = torch.as_tensor(v)
        w = torch.as_tensor(single[addr]).to(v.device).to(v.dtype)
        ev = tuple(w.shape)
        out = out | ChoiceMap.entry(Batched(torch.cat([v.reshape((n,) + ev), w.reshape((1,) + ev)])), *addr)
    return out


class SMCAlgorithm(Algorithm):
    """smc.py:117-230."""

    def get_num_particles(self) -> int:
        raise NotImplementedError

    def get_final_target(self) -> Target:
        raise NotImplementedError

    def run_smc(self, key: PRNGKey) -> ParticleCollection:
        raise NotImplementedError

    def run_csmc(self, key: PRNGKey, retained: ChoiceMap) -> ParticleCollection:
        raise NotImplementedError

    # GenSP interface (smc.py:145-198)
    def random_weighted(self, key: PRNGKey, *args):
        (target,) = args
        kb = split(key)
        key, sub_key = kb[0], kb[1]
        algorithm = ChangeTarget(self, target)
        particle_collection = algorithm.run_smc(key)
        log_marginal = particle_collection.get_log_marginal_likelihood_estimate()
        particle = particle_collection.sample_particle(sub_key)
        log_density_estimate = particle.get_score() - log_marginal
        chm = target.filter_to_unconstrained(particle.get_choices())
        return log_density_estimate, chm

    #: smc.py:181-198 scores ``sample_particle(sub_key)`` of the conditional-SMC collection -- a particle drawn by weight,
    #: not necessarily the retained one.  That is the default here (results identical to the reference's); set
    #: ``reference_compat = False`` on an algorithm to score the retained particle (the last slot), whose choices are ``v``.
    reference_compat = True

    def estimate_logpdf(self, key: PRNGKey, v: ChoiceMap, *args):
        (target,) = args
        algorithm = ChangeTarget(self, target)
        if self.reference_compat:
            kb = split(key)
            key, sub_key = kb[0], kb[1]
            particle_collection = algorithm.run_csmc(key, v)
            particle = particle_collection.sample_particle(sub_key)
        else:
            particle_collection = algorithm.run_csmc(key, v)
            particle = particle_collection.get_particle(-1 % len(particle_collection))
        log_density_estimate = particle.get_score() - particle_collection.get_log_marginal_likelihood_estimate()
        return log_density_estimate

    def log_marginal_likelihood_estimate(self, key: PRNGKey, target: Target | None = None):
        algorithm = ChangeTarget(self, target) if target is not None else self
        return algorithm.run_smc(key).get_log_marginal_likelihood_estimate()

    def simulate(self, key, args):
        (target,) = args
        w, chm = self.random_weighted(key, target)
        return _SampleTrace(self, args, chm, w)

    # VI hooks (smc.py:204-230)
    def estimate_normalizing_constant(self, key: PRNGKey, target: Target):
        algorithm = ChangeTarget(self, target)
        key, sub_key = key_children(key)
        return algorithm.run_smc(sub_key).get_log_marginal_likelihood_estimate()

    def estimate_reciprocal_normalizing_constant(self, key: PRNGKey, target: Target, latent_choices: ChoiceMap, w):
        return ChangeTarget(self, target).run_csmc_for_normalizing_constant(key, latent_choices, w)


class _SampleTrace:
    def __init__(self, gen_fn, args, chm, score):
        self._gf, self._args, self._chm, self._score = gen_fn, args, chm, score

    def get_gen_fn(self):
        return self._gf

    def get_args(self):
        return self._args

    def get_choices(self):
        return self._chm

    get_sample = get_choices

    def get_retval(self):
        return self._chm

    def get_score(self):
        return self._score


def _as_batch(tr: StaticTrace, w: torch.Tensor):
    """A scalar-key trace as a 1-particle batch (``jnp.expand_dims(v, 0)``, smc.py:262)."""
    tr.batched = True
    return tr, w.reshape(1)


class Importance(SMCAlgorithm):
    """One-particle importance sampling (smc.py:234-279)."""

    def __init__(self, target: Target, q: SampleDistribution | None = None):
        self.target = target
        self.q = q

    def get_num_particles(self):
        return 1

    def get_final_target(self):
        return self.target

    def run_smc(self, key: PRNGKey):
        kb = split(key)
        key, sub_key = kb[0], kb[1]
        if self.q is not None:
            log_weight, choice = self.q.random_weighted(sub_key, self.target)
            tr, target_score = self.target.importance(key, choice)
            tr, w = _as_batch(tr, target_score - log_weight)
        else:
            tr, target_score = self.target.importance(key, ChoiceMap.empty())
            tr, w = _as_batch(tr, target_score)
        return ParticleCollection(tr, w, True)

    def run_csmc(self, key: PRNGKey, retained: ChoiceMap):
        kb = split(key)
        key, sub_key = kb[0], kb[1]
        q_score = self.q.estimate_logpdf(sub_key, retained, self.target) if self.q else 0.0
        tr, target_score = self.target.importance(key, retained)
        tr, w = _as_batch(tr, target_score - q_score)
        return ParticleCollection(tr, w, True)


class ImportanceK(SMCAlgorithm):
    """K-particle importance sampling (smc.py:283-351)."""

    def __init__(self, target: Target, q: SampleDistribution | None = None, k_particles: int = 2):
        self.target = target
        self.q = q
        self.k_particles = int(k_particles)

    def get_num_particles(self):
        return self.k_particles

    def get_final_target(self):
        return self.target

    def run_smc(self, key: PRNGKey):
        kb = split(key)
        key, sub_key = kb[0], kb[1]
        sub_keys = split(sub_key, self.get_num_particles())
        if self.q is not None:
            log_weights, choices = self.q.random_weighted(sub_keys, self.target)
            trs, target_scores = self.target.importance(sub_keys, choices)
            lw = target_scores - log_weights
        else:
            trs, lw = self.target.importance(sub_keys, ChoiceMap.empty())
        return ParticleCollection(trs, lw, True)

    def run_csmc(self, key: PRNGKey, retained: ChoiceMap):
        """Conditional importance sampling: K - 1 fresh particles plus the retained one, last (smc.py:317-351)."""
        kb = split(key)
        key, sub_key = kb[0], kb[1]
        k = self.get_num_particles()
        if k < 2:
            return Importance(self.target, self.q).run_csmc(key, retained)
        sub_keys = split(sub_key, k - 1)
        if self.q is not None:
            log_scores, choices = self.q.random_weighted(sub_keys, self.target)
            retained_score = self.q.estimate_logpdf(key, retained, self.target)
            stacked = _stack_chm(choices, k - 1, retained)
            stacked_scores = torch.cat([log_scores.reshape(k - 1), torch.as_tensor(retained_score).reshape(1).to(log_scores)])
            trs, target_scores = self.target.importance(split(key, k), stacked)
            return ParticleCollection(trs, target_scores - stacked_scores, True)
        ignored, ignored_scores = self.target.importance(sub_keys, ChoiceMap.empty())
        retained_tr, retained_score = self.target.importance(key, retained)
        retained_tr, retained_score = _as_batch(retained_tr, retained_score)
        trs = _concat_traces(ignored, retained_tr)
        return ParticleCollection(trs, torch.cat([ignored_scores.reshape(k - 1), retained_score.reshape(1)]), True)


class ChangeTarget(SMCAlgorithm):
    """Reweight a collection for a new target (smc.py:360-396)."""

    def __init__(self, prev: SMCAlgorithm, target: Target):
        self.prev = prev
        self.target = target

    def get_num_particles(self):
        return self.prev.get_num_particles()

    def get_final_target(self):
        return self.target

    def run_smc(self, key: PRNGKey) -> ParticleCollection:
        return self._reweight(self.prev.run_smc(key), key)

    def run_csmc(self, key: PRNGKey, retained: ChoiceMap) -> ParticleCollection:
        """smc.py:398-425: conditional SMC of the previous algorithm, then the same reweighting."""
        return self._reweight(self.prev.run_csmc(key, retained), key)

    def run_csmc_for_normalizing_constant(self, key: PRNGKey, latent_choices: ChoiceMap, w):
        """smc.py:432-465: the retained particle keeps the weight ``w`` it already has under the new target."""
        kb = split(key)
        key, sub_key = kb[0], kb[1]
        collection = self.prev.run_csmc(sub_key, latent_choices)
        k = self.get_num_particles()
        particles = collection.get_particles()
        lw = collection.get_log_weights()
        retained_score = particles.score[k - 1]
        retained_weight = lw[k - 1]
        w = torch.as_tensor(w).to(lw).reshape(())
        if k > 1:
            rew = self._reweight(collection, key).get_log_weights()  # lanes 0..k-2 are the rejected particles
            all_w = torch.cat([rew[: k - 1], (w - retained_score + retained_weight).reshape(1)]).contiguous()
        else:
            all_w = (w - retained_score + retained_weight).reshape(1).contiguous()
        total = ParticleCollection(particles, all_w, True).get_log_marginal_likelihood_estimate()  # logsumexp - log k
        return retained_score - total

    def _reweight(self, collection: ParticleCollection, key: PRNGKey) -> ParticleCollection:
        particles = collection.get_particles()
        # latents of every particle, batched along the particle axis
        latents = self.prev.get_final_target().filter_to_unconstrained(_rebatch(particles))
        sub_keys = split(key, self.get_num_particles())
        merged = self.target.constraint.merge(latents)
        # this_weight = new_weight - particle.get_score() + weight   (smc.py:383)
        new_tr, new_w = self.target.p._run(
            sub_keys, self.target.args, merged, weight_mode="generate", weight_in=collection.get_log_weights(),
            score_in=particles.score, n=particles.n, batched=True
        )
        return ParticleCollection(new_tr, new_w, True)

"""Targets and the GenSP algorithm interface.

API mirror of src/genjax/_src/inference/sp.py: ``Target:53`` (``importance``
:83-87 merges the target's constraint with the proposed choices,
``filter_to_unconstrained`` :89-91), ``SampleDistribution:101``,
``Algorithm:111``.
"""

from __future__ import annotations

from ..core.choice_map import ChoiceMap, Selection
from ..core.key import KeyBatch, key_children
from ..gen.gfi import GenerativeFunction


class Target:
    """An unnormalised target: generative function + args + constraint (sp.py:53-94)."""

    def __init__(self, p: GenerativeFunction, args: tuple, constraint: ChoiceMap):
        if not isinstance(p, GenerativeFunction):
            raise TypeError("Target.p must be a GenerativeFunction")
        if isinstance(p, SampleDistribution) and not getattr(p, "_allow_as_target", False):  # Marginal, Algorithm
            # sp.py:46-49,79: a Target cannot wrap a Marginal / Algorithm directly
            raise TypeError("Target.p cannot be a SampleDistribution (Marginal / Algorithm)")
        if not isinstance(constraint, ChoiceMap):
            raise TypeError("Target.constraint must be a ChoiceMap")
        self.p = p
        self.args = tuple(args)
        self.constraint = constraint

    def importance(self, key, constraint: ChoiceMap):
        merged = self.constraint.merge(constraint)
        return self.p.importance(key, merged, self.args)

    def filter_to_unconstrained(self, choice_map: ChoiceMap) -> ChoiceMap:
        selection = ~self.constraint.get_selection()
        return choice_map.filter(selection)

    def __getitem__(self, addr):
        return self.constraint[addr]

    def __repr__(self):
        return f"Target({self.p!r}, args={self.args!r}, constraint={self.constraint!r})"


class SampleDistribution(GenerativeFunction):
    """sp.py:101-108: distributions whose samples are choice maps."""

    def random_weighted(self, key, *args):
        raise NotImplementedError

    def estimate_logpdf(self, key, v, *args):
        raise NotImplementedError


class Algorithm(SampleDistribution):
    """sp.py:111-205: inference algorithms as sample distributions over a Target's latents."""

    def random_weighted(self, key, target: Target):
        raise NotImplementedError

    def estimate_logpdf(self, key, v: ChoiceMap, target: Target):
        raise NotImplementedError


class Marginal(SampleDistribution):
    """The marginal of a generative function over a selection of addresses (sp.py:208-252).

    ``random_weighted(key, *args)``: simulate, keep the selected choices, weight = projection of the trace on
    the COMPLEMENT of the selection, exactly as sp.py:227-228 (``reference_compat=False`` opts into the
    projection on the selection itself).  ``estimate_logpdf(key, v, *args)``: importance weight of ``v``.  With an ``algorithm`` the unselected choices are marginalised by (conditional) SMC -- available
    for scalar keys (the nested particle batch is not flattened into the outer one)."""

    def __init__(self, gen_fn: GenerativeFunction, selection: Selection | None = None, algorithm=None,
                 reference_compat: bool = True):
        """``reference_compat=True`` (default) is sp.py:227-228 to the letter: the weight is the projection of the
        trace on the COMPLEMENT of the selection, also when it is handed on to
        ``algorithm.estimate_reciprocal_normalizing_constant`` (sp.py:232-236).  That is 0 for a full selection,
        so a proposal wrapped as ``q = proposal.marginal()`` contributes no density to ``ImportanceK``'s
        ``target_scores - log_weights`` (smc.py:301-315) -- the reference's behaviour, reproduced here.
        ``reference_compat=False`` is the opt-in corrected estimator: it projects on the selection itself -- the
        quantity ``estimate_logpdf`` (sp.py:244-246) assigns to those choices -- which makes the p / q weights
        of a custom proposal the importance weights (closed-form check in tests/test_gfi_gpu.py)."""
        self.gen_fn = gen_fn
        self.selection = Selection.all() if selection is None else selection
        self.algorithm = algorithm
        self.reference_compat = bool(reference_compat)

    def random_weighted(self, key, *args):
        key, sub_key = key_children(key)
        tr = self.gen_fn.simulate(sub_key, args)
        choices = tr.get_choices()
        latent_choices = choices.filter(self.selection)
        key, sub_key = key_children(key)
        weight = tr.project(sub_key, ~self.selection if self.reference_compat else self.selection)
        if self.algorithm is None:
            return weight, latent_choices
        if isinstance(key, KeyBatch):
            raise NotImplementedError("Marginal with an inner algorithm under a batched key")
        target = Target(self.gen_fn, args, latent_choices)
        other_choices = choices.filter(~self.selection)
        Z = self.algorithm.estimate_reciprocal_normalizing_constant(key, target, other_choices, weight)
        return Z, latent_choices

    def estimate_logpdf(self, key, v: ChoiceMap, *args):
        if self.algorithm is None:
            _, weight = self.gen_fn.importance(key, v, args)
            return weight
        if isinstance(key, KeyBatch):
            raise NotImplementedError("Marginal with an inner algorithm under a batched key")
        target = Target(self.gen_fn, args, v)
        return self.algorithm.estimate_normalizing_constant(key, target)

    # a Marginal is itself a generative function whose only choice is the selected choice map
    def simulate(self, key, args):
        w, chm = self.random_weighted(key, *args)
        from .smc import _SampleTrace

        return _SampleTrace(self, args, chm, w)


def marginal(selection: Selection | None = None, algorithm=None, reference_compat: bool = True):
    """``@marginal(selection, algorithm)`` decorator (sp.py:260-273)."""

    def decorator(gen_fn: GenerativeFunction) -> Marginal:
        return Marginal(gen_fn, selection, algorithm, reference_compat)

    return decorator

"""Targets and the GenSP algorithm interface.

API mirror of src/genjax/_src/inference/sp.py: ``Target:53`` (``importance``
:83-87 merges the target's constraint with the proposed choices,
``filter_to_unconstrained`` :89-91), ``SampleDistribution:101``,
``Algorithm:111``.
"""

from __future__ import annotations

from ..core.choice_map import ChoiceMap
from ..gen.gfi import GenerativeFunction


class Target:
    """An unnormalised target: generative function + args + constraint (sp.py:53-94)."""

    def __init__(self, p: GenerativeFunction, args: tuple, constraint: ChoiceMap):
        if not isinstance(p, GenerativeFunction):
            raise TypeError("Target.p must be a GenerativeFunction")
        if isinstance(p, SampleDistribution) and not getattr(p, "_allow_as_target", False):
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

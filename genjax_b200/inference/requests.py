"""``genjax.inference.requests`` namespace (src/genjax/inference/requests.py)."""

from ..gen.gfi import DiffAnnotate, EmptyRequest, Regenerate, StaticRequest, Update
from .mcmc import HMC, Rejuvenate, SafeHMC

__all__ = ["HMC", "SafeHMC", "Rejuvenate", "Regenerate", "EmptyRequest", "DiffAnnotate", "StaticRequest", "Update"]

"""genjax_b200 -- a B200-native particle-parallel inference engine behind the
GenJAX API surface (``@gen`` / ``Target`` / ``ChoiceMap`` / GFI).

Only the hot path of genjax-community/genjax is rebuilt here (SURVEY.md
section 8): batched simulate / assess / importance / update over static
``@gen`` models, the SMC propose-weight-resample step, and batched MH / HMC
transitions -- each as hand-written sm_100a CUDA reached through the C-ABI in
``include/genjax_b200.h``.  There is no CPU fallback.
"""

from .core.choice_map import (
    ChoiceMap,
    ChoiceMapBuilder,
    ChoiceMapNoValueAtAddress,
    Selection,
    SelectionBuilder,
)
from .core.key import KeyBatch, PRNGKey, fold_in, key, split
from .gen import numpy_api as numpy
from .gen.capture import AddressReuse, MissingAddress
from .gen.distributions import (
    Distribution,
    ExactDensity,
    NotFusable,
    bernoulli,
    beta,
    categorical,
    cauchy,
    chi2,
    exact_density,
    exponential,
    flip,
    gamma,
    geometric,
    gmm_diag,
    gumbel,
    half_cauchy,
    half_normal,
    inverse_gamma,
    kumaraswamy,
    laplace,
    log_normal,
    logit_normal,
    mv_normal,
    mv_normal_diag,
    normal,
    poisson,
    register_primitive,
    student_t,
    uniform,
    weibull,
)
from .gen.gfi import (
    Diff,
    DiffAnnotate,
    EditRequest,
    EmptyRequest,
    GenerativeFunction,
    IndexRequest,
    NoChange,
    NotSupportedEditRequest,
    Regenerate,
    StaticRequest,
    Trace,
    UnknownChange,
    Update,
)
from .gen.static import Batched, StaticGenerativeFunction, StaticTrace, gen, vmap
from .gen.scan import Scan, ScanTrace, accumulate, iterate, iterate_final, reduce, scan
from .gen.vmap_combinator import Vmap, VmapTrace, repeat, vmap_combinator
from .gen.switch import MaskCombinator, Switch, mask, mix, or_else, switch
from .core.mask import Mask
from .inference.sp import Algorithm, Marginal, SampleDistribution, Target, marginal
from . import inference
from .core.render import render_html

C = ChoiceMapBuilder
S = SelectionBuilder

__version__ = "0.1.0"

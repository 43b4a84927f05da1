"""Properties of the output-slot ("pull") form of systematic resampling and of a bounded reference max -- the two
ingredients of the single-pass filter step planned in DESIGN.md section 10 (oracle only; no kernel uses them yet)."""
import math

import numpy as np
import pytest

from oracle import rng, smc

F32 = np.float32


@pytest.mark.parametrize("n,scale", [(1, 1.0), (7, 1.0), (2048, 1.0), (5000, 5.0), (50_000, 30.0)])
def test_pull_form_equals_push_form(n, scale):
    g = np.random.default_rng(n)
    logw = (scale * g.standard_normal(n)).astype(F32)
    key = rng.split(rng.key(3))[1]
    push = smc.resample_systematic(logw, key)
    pull = smc.resample_systematic_pull(logw, key)
    assert np.array_equal(push, pull)
    # any sub-range of output slots can be resolved on its own (a CTA / a rank owns a range of offspring)
    lo, m = n // 3, max(1, n // 4)
    assert np.array_equal(smc.resample_systematic_pull(logw, key, out_lo=lo, out_n=min(m, n - lo)), push[lo:lo + m])


def test_bounded_reference_max_is_a_valid_resampler():
    """With M = an upper bound of the weights instead of their max: offspring counts still follow the weights
    (|count_i - N w_i| < 1), the estimate of log mean exp agrees to fp64 rounding of the 2^-36 quantisation, and the
    result does not depend on how the particles are split into shards."""
    g = np.random.default_rng(0)
    n = 20_000
    z = g.standard_normal(n).astype(F32)
    lc = F32(0.5 * math.log(2 * math.pi) + math.log(0.5))
    logw = (F32(-0.5) * z * z - lc).astype(F32)  # Normal(0.5) observation density: bounded above by -lc
    bound = F32(-lc)
    assert bound >= logw.max()
    key = rng.split(rng.key(11))[1]
    anc = smc.resample_systematic_pull(logw, key, M=bound)
    counts = np.bincount(anc, minlength=n)
    w = np.exp(logw.astype(np.float64) - np.logaddexp.reduce(logw.astype(np.float64)))
    assert np.abs(counts - n * w).max() < 1.0 + 1e-3
    assert smc.log_mean_exp_ref(logw, bound) == pytest.approx(smc.log_mean_exp(logw), abs=1e-7)
    # shard independence: the integer masses relative to the SAME reference add up exactly
    parts = np.array_split(np.arange(n), 7)
    S = sum(int(smc.det_exp_q((logw[p] - bound).astype(F32)).sum(dtype=np.uint64)) for p in parts)
    assert S == int(smc.det_exp_q((logw - bound).astype(F32)).sum(dtype=np.uint64))
    # an absurdly loose bound underflows every mass: the caller must detect S == 0 and fall back to the true max
    assert smc.log_mean_exp_ref(logw, F32(1000.0)) == -math.inf

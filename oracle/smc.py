"""SMC restatement: log-sum-exp, resampling, ImportanceK, ChangeTarget, the
bootstrap particle-filter loop, and exact Kalman ground truth (oracle only).

Reference:
  * ``ParticleCollection.get_log_marginal_likelihood_estimate`` =
    ``logsumexp(lw) - log K``                     (inference/smc.py:96-97)
  * ``sample_particle``: ``idx ~ Categorical(lw - logsumexp(lw))`` (:102-109)
  * ``ImportanceK.run_smc``: ``key, sub = split(key); keys = split(sub, K)``;
    vmapped ``target.importance``; ``lw = target_scores - q_scores`` (:298-315)
  * ``ChangeTarget.run_smc``: ``prev.run_smc(key)`` then per particle
    ``w' = new_weight - particle.get_score() + w`` with keys ``split(key, K)``
    (the SAME key, :374,386)                                   (:370-396)
  * the extend / reweight / resample loop is a USER IDIOM, not library code
    (docs/cookbook/inactive/inference/importance_sampling.ipynb cell 16,
    mapping_tutorial.ipynb cell 37): per step a vmapped ``step.importance``
    with the observation constrained, ``log_w += w``, then N categorical
    draws over ``log_w - logsumexp(log_w)`` and a gather.  Drawing N offspring
    with N independent Gumbel-max passes is O(N^2); this build (and this
    oracle) draws them with ONE systematic (or multinomial) inverse-CDF pass
    over an exact integer CDF -- same target distribution, "parity unpinned"
    against the reference for the index stream itself.

The deterministic weight pipeline (bit-identical in gjb_resample.cuh):

    M   = max_i lw_i                                   (fp32, order-free)
    q_i = round(2^36 * exp(lw_i - M))  via det_exp_q   (uint64, <= 2^37)
    C_i = inclusive prefix sum of q                    (uint64, exact => any
          reduction tree / GPU count gives the same bits)
    S   = C_{N-1}
    cnt_i = clamp(ceil(C_i * (N / S) - u0), 0, N)      (fp64 rn mul, rn sub)
    ancestors[j] = i  for j in [cnt_{i-1}, cnt_i)      (systematic)
    logZ-hat    = M + log(S) - 36 log 2 - log N        (fp64 on the host)
"""

from __future__ import annotations

import math

import numpy as np

from . import gfi, rng

F32 = np.float32
U64 = np.uint64

Q_BITS = 36
_LOG2E = F32(1.4426950408889634)
_SQRT2 = F32(1.4142135623730951)
# Taylor coefficients of 2^g = exp(g ln 2), g in [-0.5, 0.5), degree 7
_EXP2_COEF = [F32((math.log(2.0) ** k) / math.factorial(k)) for k in range(8)]


def det_exp_q(x):
    """round(2^36 * exp(x)) for x <= 0 using only IEEE fp32 mul/add (no fma,
    no libm) so that NumPy and CUDA agree bit for bit.  NaN / -inf / tiny -> 0."""
    x = np.asarray(x, dtype=F32)
    with np.errstate(invalid="ignore", over="ignore"):
        t = (x * _LOG2E).astype(F32)
        ok = t >= F32(-62.0)  # False for NaN and -inf
        t = np.where(ok, t, F32(0.0)).astype(F32)
        t = np.minimum(t, F32(0.0))
        n = np.floor(t).astype(F32)
        g = ((t - n).astype(F32) - F32(0.5)).astype(F32)
        p = np.full(x.shape, _EXP2_COEF[7], dtype=F32)
        for k in range(6, -1, -1):
            p = ((p * g).astype(F32) + _EXP2_COEF[k]).astype(F32)
        p = (p * _SQRT2).astype(F32)
        m = (p * F32(2.0**Q_BITS)).astype(F32).astype(U64)
        sh = (-n).astype(np.int64).astype(U64)
        half = np.where(sh > 0, U64(1) << np.maximum(sh, U64(1)) - U64(1), U64(0)).astype(U64)
        q = (m + half) >> sh
    return np.where(ok, q, U64(0)).astype(U64)


def lse_terms(logw):
    """(M, S): fp32 max and exact integer mass of the weights relative to M."""
    logw = np.asarray(logw, dtype=F32)
    if logw.size == 0:
        return F32(-np.inf), 0
    with np.errstate(invalid="ignore"):
        M = np.fmax.reduce(logw)
        if not np.isfinite(M):
            M = F32(-np.inf) if not (M == np.inf) else M
        q = det_exp_q((logw - M).astype(F32))
    return F32(M), int(q.sum(dtype=U64))


def log_mean_exp(logw, n_total=None):
    """logsumexp(lw) - log N from the exact terms (fp64)."""
    M, S = lse_terms(logw)
    n = len(logw) if n_total is None else n_total
    if S == 0:
        return -math.inf
    return float(M) + math.log(S) - Q_BITS * math.log(2.0) - math.log(n)


def resample_u0(key: rng.Key):
    """The one uniform a systematic resample consumes: site 0, chunk 0, word 0."""
    w0, _, _, _ = rng.site_words(key.words, np.array([key.index], dtype=U64), 0, 0)
    return rng.u01(w0)[0]


def systematic_counts(logw, u0, n_out=None, M=None, S=None, c_offset=0):
    """Cumulative offspring counts cnt_i (int64) for local weights ``logw``.

    ``M``/``S``/``c_offset`` let a shard use the GLOBAL max, mass and the
    exclusive prefix of the shards before it (multi-GPU)."""
    logw = np.asarray(logw, dtype=F32)
    n = len(logw)
    n_out = n if n_out is None else n_out
    if M is None:
        M, S = lse_terms(logw)
    with np.errstate(invalid="ignore"):
        q = det_exp_q((logw - F32(M)).astype(F32))
    C = np.cumsum(q, dtype=U64) + U64(c_offset)
    if S == 0:
        return None, C
    scale = np.float64(n_out) / np.float64(S)
    pos = C.astype(np.float64) * scale - np.float64(u0)
    cnt = np.clip(np.ceil(pos), 0, n_out).astype(np.int64)
    cnt[C == U64(S)] = n_out  # the last unit of mass closes the range
    return cnt, C


def resample_systematic(logw, key: rng.Key):
    """ancestors int32[N]; identity when every weight is zero (is_valid False)."""
    n = len(logw)
    u0 = resample_u0(key)
    cnt, _ = systematic_counts(logw, u0)
    if cnt is None:
        return np.arange(n, dtype=np.int32)
    prev = np.concatenate([[0], cnt[:-1]])
    return np.repeat(np.arange(n, dtype=np.int32), (cnt - prev).astype(np.int64))


def resample_systematic_pull(logw, key: rng.Key, M=None, out_lo=0, out_n=None):
    """The same ancestors as ``resample_systematic``, resolved per OUTPUT slot: slot j takes the first particle i
    whose cumulative count exceeds j (``cnt_i > j``).  This is the formulation in which a CTA owns a range of
    offspring slots (and can go on to propagate them) instead of a range of parents; ``M`` may be any reference
    >= max(logw) known before the weights are (an analytic upper bound of the incremental weight) -- the masses,
    hence the ancestors, are then a deterministic function of (logw, M) and the max pass disappears."""
    logw = np.asarray(logw, dtype=F32)
    n = len(logw)
    out_n = n - out_lo if out_n is None else out_n
    u0 = resample_u0(key)
    if M is None:
        M, S = lse_terms(logw)
    else:
        with np.errstate(invalid="ignore"):
            S = int(det_exp_q((logw - F32(M)).astype(F32)).sum(dtype=U64))
    cnt, _ = systematic_counts(logw, u0, M=M, S=S)
    j = np.arange(out_lo, out_lo + out_n, dtype=np.int64)
    if cnt is None:
        return j.astype(np.int32)
    return np.searchsorted(cnt, j, side="right").astype(np.int32)


def log_mean_exp_ref(logw, M, n_total=None):
    """``log_mean_exp`` with the masses taken relative to a given reference ``M >= max(logw)``."""
    logw = np.asarray(logw, dtype=F32)
    with np.errstate(invalid="ignore"):
        S = int(det_exp_q((logw - F32(M)).astype(F32)).sum(dtype=U64))
    n = len(logw) if n_total is None else n_total
    if S == 0:
        return -math.inf
    return float(F32(M)) + math.log(S) - Q_BITS * math.log(2.0) - math.log(n)


# ------------------------------------------------- tile-exponent masses
#
# The single-launch filter step (csrc/gjb_step.cuh, gen/codegen.py ``pf_step_kernel``) cannot afford a pass for the
# global maximum before the masses are formed.  Each tile of TE_TILE consecutive particles (tile = global index //
# 2048, a constant of the ALGORITHM, not of the launch) therefore takes its masses relative to its own power-of-two
# reference 2^e_p, e_p = ceil(max_tile(lw * log2 e)); the consumer aligns the tiles with exact right shifts:
#
#     t_i = fl32(lw_i * log2e)  (non-finite -> mass 0);  r_i = rint(t_i);  g_i = t_i - r_i in [-1/2, 1/2]
#     q_i = rint(fl32(2^g_i) * 2^(36 - (e_p - r_i)))                     (fp32 fma polynomial, te_q)
#     c_i = inclusive prefix of q inside the tile (uint64),  T_p = the tile's mass
#     e   = max_p e_p (tiles with mass);  s_p = min(e - e_p, 63);  P_p = inclusive prefix of (T_p >> s_p);  S = P_last
#     C_i = P_{p-1} + (c_i >> s_p)        -- a monotone integer CDF: any CTA / GPU partition gives the same bits
#     cnt_i, ancestors as in systematic_counts;  log-mean-exp = e ln 2 + log S - 36 ln 2 - log N.
#
# A particle's effective mass floor(c_i / 2^s) - floor(c_{i-1} / 2^s) is within one unit (2^-36 of the heaviest
# particle's weight) of its exact share.  Same estimator; "parity unpinned" against the reference like every
# resampling index stream (module docstring).

TE_TILE = 2048
TE_E_NONE = -(2**31)
_TE_CLAMP = F32(2.0**20)
_TE_MAGIC = F32(12582912.0)  # 1.5 * 2^23: t + MAGIC rounds t to the nearest integer (ties to even), held in the low mantissa bits
# Taylor coefficients of 2^g, g in [-0.5, 0.5], degree 7 (|error| < 6e-9 relative), evaluated with fp32 FMAs (Horner)
_TE_COEF = _EXP2_COEF


def te_q(t, e):
    """Mass of a particle with scaled log-weight ``t`` (finite float32, |t| <= 2^20) relative to 2^e, e >= rint(t):
    rint(fl32(p * 2^(36 - (e - r)))) with r = rint(t), g = t - r, p = 2^g by fp32 FMA Horner -- operation for operation
    csrc/gjb_step.cuh te_q (one FADD for the rounding, no conversion-pipe instruction, one float -> uint64 convert)."""
    t = np.asarray(t, dtype=F32)
    tm = (t + _TE_MAGIC).astype(F32)
    r = tm.view(np.int32).astype(np.int64) - 0x4B400000
    g = (t - (tm - _TE_MAGIC).astype(F32)).astype(F32)
    p = np.full(t.shape, _TE_COEF[7], dtype=F32)
    for k in range(6, -1, -1):
        p = rng.fma32(p, g, _TE_COEF[k])
    ex = np.maximum(r + 163 - np.asarray(e, dtype=np.int64), 63)  # biased exponent of 2^(36 - (e - r)), at least 2^-64
    scale = (ex.astype(np.int32) << np.int32(23)).view(F32)
    return np.rint((p * scale).astype(F32)).astype(U64)


def te_tile_masses(logw):
    """(q uint64 [n], e_p int64 [tiles]) of tile-exponent masses; tiles are TE_TILE consecutive particles."""
    logw = np.asarray(logw, dtype=F32)
    n = len(logw)
    tiles = max(1, -(-n // TE_TILE))
    with np.errstate(invalid="ignore", over="ignore"):
        t = (logw * _LOG2E).astype(F32)
        ok = np.isfinite(t)
        t = np.clip(np.where(ok, t, F32(0.0)), -_TE_CLAMP, _TE_CLAMP).astype(F32)
        tp = np.full(tiles * TE_TILE, -np.inf, dtype=F32)
        tp[:n] = np.where(ok, t, F32(-np.inf))
        tmax = tp.reshape(tiles, TE_TILE).max(axis=1)
        live = np.isfinite(tmax)
        e_p = np.where(live, np.ceil(np.where(live, tmax, 0.0)), TE_E_NONE).astype(np.int64)
        e_i = e_p[np.arange(n) // TE_TILE]
        q = np.where(ok, te_q(t, np.where(ok, e_i, 0)), U64(0)).astype(U64)
    return q, e_p


def te_cdf(logw):
    """(C uint64 [n] global inclusive CDF, S int, e int or None): the tile-exponent CDF of the header comment."""
    q, e_p = te_tile_masses(logw)
    n = len(q)
    tiles = len(e_p)
    qp = np.zeros(tiles * TE_TILE, dtype=U64)
    qp[:n] = q
    c = np.cumsum(qp.reshape(tiles, TE_TILE), axis=1, dtype=U64)
    T = c[:, -1]
    livemass = T > 0
    if not livemass.any():
        return np.zeros(n, dtype=U64), 0, None
    e = int(e_p[livemass].max())
    s_p = np.where(livemass, np.minimum(e - np.where(livemass, e_p, e), 63), 63).astype(U64)
    Tp = T >> s_p
    P = np.cumsum(Tp, dtype=U64)
    base = np.concatenate([[U64(0)], P[:-1]]).astype(U64)
    C = (base[:, None] + (c >> s_p[:, None])).reshape(-1)[:n]
    return C, int(P[-1]), e


def te_log_mean_exp(logw, n_total=None):
    _, S, e = te_cdf(logw)
    n = len(logw) if n_total is None else n_total
    if S == 0:
        return -math.inf
    return e * math.log(2.0) + math.log(S) - Q_BITS * math.log(2.0) - math.log(n)


def te_counts(logw, u0, n_out=None):
    """Cumulative offspring counts under the tile-exponent CDF (None when no weight has mass)."""
    C, S, _ = te_cdf(logw)
    n_out = len(logw) if n_out is None else n_out
    if S == 0:
        return None
    scale = np.float64(n_out) / np.float64(S)
    pos = C.astype(np.float64) * scale - np.float64(u0)
    cnt = np.clip(np.ceil(pos), 0, n_out).astype(np.int64)
    cnt[C == U64(S)] = n_out
    return cnt


def resample_systematic_te(logw, key: rng.Key, out_lo=0, out_n=None):
    """Systematic ancestors of the offspring slots [out_lo, out_lo + out_n) under the tile-exponent CDF: slot j takes
    the first particle whose cumulative count exceeds j.  Identity when no weight has mass."""
    n = len(logw)
    out_n = n - out_lo if out_n is None else out_n
    j = np.arange(out_lo, out_lo + out_n, dtype=np.int64)
    cnt = te_counts(logw, resample_u0(key))
    if cnt is None:
        return j.astype(np.int32)
    return np.searchsorted(cnt, j, side="right").astype(np.int32)


def _mulhi64(a, b):
    """floor(a*b / 2^64) for uint64 arrays / scalars (32-bit limbs)."""
    a = np.asarray(a, dtype=U64)
    b = np.asarray(b, dtype=U64)
    m = U64(0xFFFFFFFF)
    s = U64(32)
    a0, a1 = a & m, a >> s
    b0, b1 = b & m, b >> s
    with np.errstate(over="ignore"):
        t = a0 * b0
        w1 = a1 * b0 + (t >> s)
        w2 = a0 * b1 + (w1 & m)
        return a1 * b1 + (w1 >> s) + (w2 >> s)


def resample_multinomial(logw, key: rng.KeyBatch | rng.Key):
    """N iid categorical draws: offspring j takes r_j = (w0 << 32 | w1) of its
    own Philox lane (site 0, chunk 1), target = mulhi64(r_j, S), ancestor =
    first i with C_i > target.  (Same law as the reference idiom's
    ``jax.random.categorical`` per offspring.)"""
    logw = np.asarray(logw, dtype=F32)
    n = len(logw)
    M, S = lse_terms(logw)
    if S == 0:
        return np.arange(n, dtype=np.int32)
    q = det_exp_q((logw - M).astype(F32))
    C = np.cumsum(q, dtype=U64)
    idx = np.arange(n, dtype=U64) + U64(key.index if isinstance(key, rng.Key) else key.offset)
    w0, w1, _, _ = rng.site_words(key.words, idx, 0, 1)
    r = (w0.astype(U64) << U64(32)) | w1.astype(U64)
    tgt = _mulhi64(r, U64(S))
    return np.searchsorted(C, tgt, side="right").astype(np.int32)


# ------------------------------------------------------------ ImportanceK


class OParticles:
    def __init__(self, trace: gfi.OTrace, log_weights):
        self.trace = trace
        self.log_weights = np.asarray(log_weights, dtype=F32)

    def log_marginal_likelihood_estimate(self):
        return log_mean_exp(self.log_weights)


def importance_k(model, args, constraint, key: rng.Key, k):
    """ImportanceK(target, k_particles=k).run_smc(key) without a custom q."""
    kb = rng.split(key)
    sub_key = kb[1]
    sub_keys = rng.split(sub_key, k)
    tr, w = gfi.generate(model, sub_keys, constraint, args)
    return OParticles(tr, w)


def change_target(prev: OParticles, key: rng.Key, model, args, new_constraint, old_constraint_addrs):
    """ChangeTarget._reweight (smc.py:378-384) applied to ``prev``."""
    n = len(prev.log_weights)
    latents = {a: v for a, v in prev.trace.choices.items() if a not in old_constraint_addrs}
    chm = dict(latents)
    chm.update(new_constraint)
    sub_keys = rng.split(key, n)
    tr, w = gfi.generate(model, sub_keys, chm, args)
    new_w = ((w - prev.trace.get_score()).astype(F32) + prev.log_weights).astype(F32)
    return OParticles(tr, new_w)


def sample_particle_index(log_weights, key: rng.Key):
    """Categorical(lw - LSE) draw: one multinomial offspring (lane of ``key``)."""
    lw = np.asarray(log_weights, dtype=F32)
    M, S = lse_terms(lw)
    q = det_exp_q((lw - M).astype(F32))
    C = np.cumsum(q, dtype=U64)
    w0, w1, _, _ = rng.site_words(key.words, np.array([key.index], dtype=U64), 0, 1)
    r = (w0.astype(U64) << U64(32)) | w1.astype(U64)
    tgt = _mulhi64(r, U64(S))
    return int(np.searchsorted(C, tgt, side="right")[0])


# ------------------------------------------------------- bootstrap filter


def pf_step_keys(key: rng.Key, t: int):
    """(propose key, resample key) of filter step t: split(fold_in(key, t))."""
    kb = rng.split(rng.fold_in(key, t))
    return kb[0], kb[1]


def particle_filter(step_model, key: rng.Key, state0, observations, shared_args=(), resampler="systematic",
                    record=False, m_ref=None, masses="exact_max"):
    """Bootstrap PF.  ``step_model(h, *state, *shared_args)`` returns the new
    state (array or tuple of arrays); ``observations`` is a list of dicts
    addr -> value constrained at each step.

    Per step t:  keys = split(k_prop, N); (tr, w) = importance;
    logZ += log mean exp(w); ancestors = resample(w, k_res); state = retval[anc].
    Returns dict(state, logz (float64 total), logz_inc list, history).
    """
    state = tuple(np.asarray(s) for s in (state0 if isinstance(state0, tuple) else (state0,)))
    n = state[0].shape[0]
    logz = 0.0
    incs = []
    hist = []
    for t, obs in enumerate(observations):
        k_prop, k_res = pf_step_keys(key, t)
        keys = rng.split(k_prop, n)
        tr, w = gfi.generate(step_model, keys, obs, state + tuple(shared_args))
        # m_ref: a reference maximum known before the weights are (an analytic bound of the incremental weight);
        # the integer masses, hence ancestors and estimate, are then taken relative to it instead of max(w)
        # masses="tile_exponent": the single-launch step's CDF (tile-exponent masses, see above)
        if masses == "tile_exponent":
            inc = te_log_mean_exp(w)
        else:
            inc = log_mean_exp(w) if m_ref is None else log_mean_exp_ref(w, m_ref)
        incs.append(inc)
        logz += inc
        if resampler == "systematic" and masses == "tile_exponent":
            anc = resample_systematic_te(w, k_res)
        elif resampler == "systematic":
            anc = resample_systematic(w, k_res) if m_ref is None else resample_systematic_pull(w, k_res, M=F32(m_ref))
        else:
            anc = resample_multinomial(w, rng.split(k_res, n))
        rv = tr.retval if isinstance(tr.retval, tuple) else (tr.retval,)
        new_state = tuple(np.asarray(r)[anc] for r in rv)
        if record:
            hist.append(dict(pre_state=rv, logw=w, ancestors=anc, choices=tr.choices))
        state = new_state
    return dict(state=state, logz=logz, logz_inc=incs, history=hist)


# ------------------------------------------------------------ exact truth


def kalman_logz(ys, a, q, c, r, m0=0.0, p0=1.0):
    """Exact log p(y_1:T) of x_0~N(m0,p0), x_t~N(a x_{t-1}, q^2), y_t~N(c x_t, r^2)
    (std-devs q, r), independently per dimension; ys: [T] or [T, d]. float64."""
    ys = np.asarray(ys, dtype=np.float64)
    if ys.ndim == 1:
        ys = ys[:, None]
    T, d = ys.shape
    m = np.full(d, m0, dtype=np.float64)
    p = np.full(d, p0, dtype=np.float64)
    ll = 0.0
    for t in range(T):
        m = a * m
        p = a * a * p + q * q
        s = c * c * p + r * r
        resid = ys[t] - c * m
        ll += float(np.sum(-0.5 * (np.log(2 * np.pi * s) + resid * resid / s)))
        k = p * c / s
        m = m + k * resid
        p = (1 - k * c) * p
    return ll


def simulate_lgssm(seed, T, d, a, q, c, r):
    """Synthetic observations y_1:T for the linear-Gaussian SSM (NumPy PCG64)."""
    g = np.random.default_rng(seed)
    x = g.standard_normal(d)
    ys = np.empty((T, d), dtype=np.float32)
    for t in range(T):
        x = a * x + q * g.standard_normal(d)
        ys[t] = c * x + r * g.standard_normal(d)
    return ys

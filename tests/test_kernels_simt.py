"""The block-level CUDA kernels, executed on the CPU with real block semantics (tests/simt_kernels.py: one OS thread
per CUDA thread, real barriers / shuffles / atomics) and compared with the oracle bit for bit -- the same contracts
tests/test_core_gpu.py checks on the device.  A CPU regression guard for kernel edits; the device run stays the
authority on the device's own behaviour."""
import ctypes as C

import numpy as np
import pytest

import simt_kernels
from genjax_b200.runtime import cabi
from oracle import rng, smc

F32 = np.float32
NEG_INF = 0x007FFFFF


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


@pytest.fixture(scope="module")
def core():
    return simt_kernels.core()


def _terms(core, lw):
    n = lw.size
    wmax = np.array([NEG_INF], dtype=np.uint32)
    core.s_weight_max(_p(lw), C.c_int64(n), _p(wmax), C.c_int(max(1, min(4, (n + 255) // 256))))
    tm = np.zeros((n + 2047) // 2048, dtype=np.uint64)
    core.s_weight_mass(_p(lw), C.c_int64(n), _p(wmax), None, _p(tm))
    return wmax, tm


def _resample_args(lw, wmax, tm, key, anc, lse=None, **kw):
    R = cabi.ResampleArgs()
    R.logw, R.n, R.wmax, R.tile_mass = lw.ctypes.data, lw.size, wmax.ctypes.data, tm.ctypes.data
    R.n_total, R.out_lo, R.out_n, R.anc_base = kw.get("n_total", lw.size), kw.get("out_lo", 0), anc.size, kw.get("anc_base", 0)
    R.key0, R.key1, R.key_index = key.words[0], key.words[1], key.index
    R.ancestors = anc.ctypes.data
    if lse is not None:
        R.lse_out = lse.ctypes.data
    return R


@pytest.mark.parametrize("n,scale", [(1, 1.0), (5, 1.0), (2048, 1.0), (2049, 1.0), (4097, 3.0), (12_000, 0.1), (9000, 30.0)])
def test_mass_and_systematic_resampling_kernels(core, n, scale):
    """weight_max_kernel, weight_mass_kernel, lse_finalize_kernel, resample_systematic_kernel: exact integer mass, bit-exact
    ancestors (balanced direct write-out, max-scan windows and the single-owner fill of degenerate weights)."""
    lw = (scale * np.random.default_rng(n).standard_normal(n)).astype(F32)
    wmax, tm = _terms(core, lw)
    M, S = smc.lse_terms(lw)
    assert int(tm.sum()) == S
    out = np.zeros(3)
    core.s_lse_finalize(_p(tm), C.c_int(tm.size), _p(wmax), None, C.c_int64(n), _p(out))
    assert out[0] == float(M) and out[1] == float(S) and out[2] == pytest.approx(smc.log_mean_exp(lw), abs=1e-12)
    key = rng.split(rng.key(n))[1]
    anc, lse = np.full(n, -1, dtype=np.int32), np.zeros(3)
    core.s_resample_systematic(C.byref(_resample_args(lw, wmax, tm, key, anc, lse)))
    assert np.array_equal(anc, smc.resample_systematic(lw, key))
    assert lse[1] == float(S)
    # an output window resolved on its own (what a shard / a rank does)
    lo, m = n // 3, max(1, n // 2)
    m = min(m, n - lo)
    part = np.full(m, -1, dtype=np.int32)
    core.s_resample_systematic(C.byref(_resample_args(lw, wmax, tm, key, part, out_lo=lo)))
    assert np.array_equal(part, smc.resample_systematic(lw, key)[lo:lo + m])


def test_degenerate_and_invalid_weights(core):
    n = 6000
    lw = np.full(n, -np.inf, dtype=F32)
    lw[4321] = 0.0  # one particle owns every offspring: whole windows filled by a single owner
    wmax, tm = _terms(core, lw)
    anc = np.full(n, -1, dtype=np.int32)
    key = rng.key(3)
    core.s_resample_systematic(C.byref(_resample_args(lw, wmax, tm, key, anc)))
    assert (anc == 4321).all()
    lw[:] = -np.inf  # every weight zero: identity ancestors, S == 0
    wmax, tm = _terms(core, lw)
    assert int(tm.sum()) == 0
    core.s_resample_systematic(C.byref(_resample_args(lw, wmax, tm, key, anc)))
    assert np.array_equal(anc, np.arange(n, dtype=np.int32))


@pytest.mark.parametrize("n", [1, 700, 2048])
def test_fused_mass_resample_kernel_single_tile(core, n):
    """mass_resample_kernel (cooperative on the device) as a grid of one block: same ancestors as the two launches."""
    lw = (2.0 * np.random.default_rng(n + 1).standard_normal(n)).astype(F32)
    wmax, tm = _terms(core, lw)
    key = rng.split(rng.key(n + 1))[1]
    anc, lse = np.full(n, -1, dtype=np.int32), np.zeros(3)
    scratch = np.zeros(1, dtype=np.uint64)
    heavy = np.zeros(cabi.GJB_HEAVY_WS_WORDS, dtype=np.uint32)
    R = _resample_args(lw, wmax, scratch, key, anc, lse)
    R.heavy_ws = heavy.ctypes.data
    assert core.s_mass_resample_one_block(C.byref(R)) == 0
    assert np.array_equal(anc, smc.resample_systematic(lw, key))
    assert lse[2] == pytest.approx(smc.log_mean_exp(lw), abs=1e-12)


def test_multinomial_and_gather_kernels(core):
    n = 5000
    lw = (2.0 * np.random.default_rng(7).standard_normal(n)).astype(F32)
    wmax, tm = _terms(core, lw)
    kb = rng.split(rng.key(6), n)
    cdf, anc = np.zeros(n, dtype=np.uint64), np.full(n, -1, dtype=np.int32)
    core.s_multinomial(_p(lw), C.c_int64(n), _p(wmax), _p(tm), _p(cdf), C.c_uint32(kb.words[0]), C.c_uint32(kb.words[1]),
                       C.c_uint64(kb.offset), C.c_int64(n), _p(anc))
    assert np.array_equal(anc, smc.resample_multinomial(lw, kb))
    src = np.random.default_rng(8).integers(0, 1 << 31, (n, 3)).astype(np.uint32)
    dst = np.zeros((n, 3), dtype=np.uint32)
    core.s_gather_rows(_p(src), _p(anc), _p(dst), C.c_int64(n), C.c_int(3), C.c_int(4))
    assert np.array_equal(dst, src[anc])


@pytest.mark.parametrize("n,scale,kind", [(1, 1.0, "n"), (7, 1.0, "n"), (2048, 1.0, "n"), (2049, 2.0, "n"), (6000, 0.3, "n"),
                                           (9000, 30.0, "n"), (12_345, 300.0, "n"), (5000, 1.0, "one"), (5000, 1.0, "dead"),
                                           (7000, 1.0, "deadtile"), (4100, 1.0, "nan")])
def test_tile_exponent_masses_and_pull_resampling(core, n, scale, kind):
    """te_mass_kernel (te_publish) and te_resample_kernel (te_pull): within-tile CDFs, tile records, ancestors and the
    log-mean-exp terms of the tile-exponent pipeline, bit for bit against oracle/smc.py -- balanced and degenerate
    weights, dead particles, a dead tile, NaN / +inf weights, partial last tile, output windows."""
    r = np.random.default_rng(n)
    lw = (r.standard_normal(n) * scale - 3.0).astype(F32)
    if kind == "one":
        lw[:] = -np.inf
        lw[n - 3] = 1.5
    elif kind == "dead":
        lw[:] = -np.inf
    elif kind == "deadtile":
        lw[2048:4096] = -np.inf
        lw[5] = np.nan
    elif kind == "nan":
        lw[::7] = np.nan
        lw[3] = np.inf
    tiles = (n + 2047) // 2048
    cdf = np.zeros(tiles * 2048, dtype=np.uint64)
    recs = np.zeros(tiles, dtype=[("mass", np.uint64), ("e", np.int32), ("pad", np.int32)])
    core.s_te_masses(_p(lw), C.c_int64(n), _p(cdf), _p(recs))
    q, e_p = smc.te_tile_masses(lw)
    qp = np.zeros(tiles * 2048, dtype=np.uint64)
    qp[:n] = q
    want_cdf = np.cumsum(qp.reshape(tiles, 2048), axis=1, dtype=np.uint64).reshape(-1)
    assert np.array_equal(cdf, want_cdf)
    assert np.array_equal(recs["mass"], want_cdf.reshape(tiles, 2048)[:, -1])
    live = recs["mass"] > 0
    assert np.array_equal(recs["e"][live], e_p[live].astype(np.int32))
    key = rng.Key((0x1234 + n, 77), 5)
    kd = np.array([key.words[0], key.words[1], key.index & 0xFFFFFFFF, key.index >> 32], dtype=np.uint32)
    for out_lo, out_n in ((0, n), (n // 3, n - n // 3 - n // 5)):
        if out_n <= 0:
            continue
        anc = np.full(out_n, -7, dtype=np.int32)
        lse = np.zeros(3, dtype=np.float64)
        A = cabi.TeResampleArgs()
        A.cdf, A.recs, A.n_tiles_total, A.n_total = cdf.ctypes.data, recs.ctypes.data, tiles, n
        A.out_lo, A.out_n, A.key_dev, A.ancestors, A.lse_out = out_lo, out_n, kd.ctypes.data, anc.ctypes.data, lse.ctypes.data
        core.s_te_resample(C.byref(A))
        want = smc.resample_systematic_te(lw, key, out_lo, out_n)
        assert np.array_equal(anc, want), (np.flatnonzero(anc != want)[:5], anc[:8], want[:8])
        lme = smc.te_log_mean_exp(lw)
        if np.isfinite(lme):
            assert lse[2] == pytest.approx(lme, abs=1e-12, rel=1e-13)
            assert lse[1] == float(smc.te_cdf(lw)[1])
        else:
            assert lse[1] == 0.0 and lse[2] == -np.inf


def test_device_key_table_equals_the_host_key_tree(core):
    """pf_key_table_kernel (threefry2x32-20 on the device) == core/key.py pf_key_table == oracle/smc.py pf_step_keys:
    proposal words, resample key, multinomial lane words of every step."""
    import genjax_b200 as gj
    from genjax_b200.core.key import pf_key_table

    for seed, T in ((314159, 100), (7, 1), (2**40 + 5, 333)):
        k = gj.fold_in(gj.key(seed), 3)
        w0, w1 = k.collapsed()
        out = np.zeros((T, 8), dtype=np.uint32)
        core.s_pf_key_table(C.c_uint32(w0), C.c_uint32(w1), C.c_int(T), _p(out))
        assert np.array_equal(out, pf_key_table(k, T))
        okey = rng.fold_in(rng.key(seed), 3)
        for t in (0, T - 1):
            kp, kr = smc.pf_step_keys(okey, t)
            assert tuple(out[t, 2:4]) == kr.words and out[t, 4] == kr.index
            assert tuple(out[t, 0:2]) == rng.split(kp, 4).words and tuple(out[t, 6:8]) == rng.split(kr, 4).words

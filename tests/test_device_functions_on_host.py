"""The CUDA SOURCE of the exact-arithmetic device functions, run on the host.

genjax_b200/csrc/gjb_rng.cuh and gjb_resample.cuh are compiled by g++ against tests/host_shim/cuda_runtime.h
(intrinsics restated as plain IEEE operations, -ffp-contract=off) and their scalar functions are compared with the
oracle bit for bit: Philox4x32-10 and the lane / quad counter layout, the uniform conversion, the ordered float
encoding of the running max, det_exp_q (the 2^36 fixed-point exponential every integer mass comes from),
offspring_cnt (the fp64 count every ancestor comes from) and the resampling uniform.  This is a CPU regression guard on
the kernel source text -- the device execution itself is checked by the -m gpu tests."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import rng, smc

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
F32 = np.float32


@pytest.fixture(scope="module")
def lib(tmp_path_factory):
    out = tmp_path_factory.mktemp("devfn") / "libdevfn.so"
    cmd = ["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-shared", f"-I{HERE}/host_shim",
           f"-I{ROOT}/genjax_b200/csrc", f"-I{ROOT}/include", "-o", str(out), f"{HERE}/host_shim/device_functions.cpp"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    lib = C.CDLL(str(out))
    lib.h_resample_u0.restype = C.c_double
    lib.h_resample_u0.argtypes = [C.c_uint32, C.c_uint32, C.c_uint64]
    return lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def test_philox_and_counter_layout(lib):
    out = np.zeros(4, dtype=np.uint32)
    for ctr, key, want in (((0, 0, 0, 0), (0, 0), (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)),
                           ((0xFFFFFFFF,) * 4, (0xFFFFFFFF,) * 2, (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)),
                           ((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0),
                            (0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1))):  # Random123 known answers
        lib.h_philox(_p(np.array(ctr, dtype=np.uint32)), C.c_uint32(key[0]), C.c_uint32(key[1]), _p(out))
        assert tuple(int(x) for x in out) == want
    words = (0x12345678, 0x9ABCDEF0)
    for idx in (0, 5, 1 << 33, (1 << 40) + 12345):
        for site, chunk in ((1, 0), (3, 2), (7, 0xFFFF)):
            lib.h_lane_words(C.c_uint32(words[0]), C.c_uint32(words[1]), C.c_uint64(idx), C.c_uint32(site), C.c_uint32(chunk), _p(out))
            want = rng.site_words(words, np.array([idx], dtype=np.uint64), site, chunk)
            assert [int(x) for x in out] == [int(w[0]) for w in want]
    idx = np.arange(64, dtype=np.uint64) + np.uint64(1000)
    want = rng.quad_slot_words(words, idx, 2)  # particle idx takes word (idx & 3) of the block of quad idx >> 2
    for i, w in zip(idx, want):
        lib.h_quad_words(C.c_uint32(words[0]), C.c_uint32(words[1]), C.c_uint64(int(i) >> 2), C.c_uint32(2), C.c_uint32(0), _p(out))
        assert int(out[int(i) & 3]) == int(w)


def test_uniform_conversion_and_ordered_float_encoding(lib):
    g = np.random.default_rng(0)
    bits = np.concatenate([g.integers(0, 1 << 32, 4096, dtype=np.uint64).astype(np.uint32),
                           np.array([0, 1, 511, 512, 0xFFFFFFFF, 0xFFFFFE00], dtype=np.uint32)])
    out = np.zeros(bits.size, dtype=F32)
    lib.h_u01(_p(bits), C.c_int(bits.size), _p(out))
    assert np.array_equal(out, rng.u01(bits)) and out.min() > 0.0 and out.max() < 1.0
    f = np.concatenate([g.standard_normal(1000).astype(F32) * F32(50), np.array([0.0, -0.0, np.inf, -np.inf, 1e-45, -1e-45], dtype=F32)])
    enc = np.zeros(f.size, dtype=np.uint32)
    lib.h_fenc(_p(f), C.c_int(f.size), _p(enc))
    keep = ~((f == 0) & np.signbit(f))  # -0.0 == 0.0 as floats but encodes just below it
    order = np.argsort(f[keep], kind="stable")
    assert np.all(np.diff(enc[keep][order].astype(np.int64)) >= 0)  # the encoding preserves the order: atomicMax works
    assert int(enc[f == -np.inf][0]) == 0x007FFFFF  # GJB_WMAX_NEG_INF
    dec = np.zeros(f.size, dtype=F32)
    lib.h_fdec(_p(enc), C.c_int(f.size), _p(dec))
    assert np.array_equal(dec.view(np.uint32), f.view(np.uint32))
    z = np.zeros(2 * 64, dtype=F32)
    b0, b1 = bits[:64].copy(), bits[64:128].copy()
    lib.h_box_muller(_p(b0), _p(b1), C.c_int(64), _p(z))
    want = rng.box_muller(b0, b1)
    np.testing.assert_allclose(z.reshape(64, 2), np.stack(want, 1), rtol=2e-5, atol=2e-6)  # libm sin/cos, not sincospi


def test_det_exp_q_bit_exact(lib):
    g = np.random.default_rng(1)
    x = np.concatenate([-np.abs(g.standard_normal(20000) * 8).astype(F32), -np.linspace(0, 45, 4001).astype(F32),
                        np.array([0.0, -0.0, -1e-30, -43.0, -44.0, -100.0, -np.inf, np.nan], dtype=F32)])
    out = np.zeros(x.size, dtype=np.uint64)
    lib.h_det_exp_q(_p(x), C.c_int(x.size), _p(out))
    assert np.array_equal(out, smc.det_exp_q(x))
    assert abs(int(out[x == 0][0]) - (1 << 36)) <= 1 << 13  # exp(0) to 2^-23 relative (fp32 polynomial), same on both sides
    assert out[np.isnan(x)][0] == 0 and out[np.isinf(x)][0] == 0


def test_offspring_counts_and_resample_uniform_bit_exact(lib):
    g = np.random.default_rng(2)
    for n, scale in ((5000, 1.0), (100_003, 6.0)):
        logw = (scale * g.standard_normal(n)).astype(F32)
        key = rng.split(rng.key(n))[1]
        u0 = smc.resample_u0(key)
        assert lib.h_resample_u0(C.c_uint32(key.words[0]), C.c_uint32(key.words[1]), C.c_uint64(key.index)) == float(u0)
        M, S = smc.lse_terms(logw)
        cnt, Cq = smc.systematic_counts(logw, u0)
        got = np.zeros(n, dtype=np.int32)
        lib.h_offspring_cnt(_p(np.ascontiguousarray(Cq)), C.c_int(n), C.c_uint64(S), C.c_int32(n), C.c_double(float(u0)), _p(got))
        assert np.array_equal(got, cnt.astype(np.int32))


def test_distribution_log_densities_and_samplers(lib):
    """gjb_dist.cuh against oracle/dists.py: same float32 operation order, host libm on both sides."""
    from oracle import dists as od

    g = np.random.default_rng(3)
    n = 4000
    v = g.standard_normal(n).astype(F32)
    a = (g.standard_normal(n) * 2).astype(F32)
    b = (0.2 + g.random(n) * 3).astype(F32)
    pos = (0.05 + g.random(n) * 4).astype(F32)
    unit = (0.01 + 0.98 * g.random(n)).astype(F32)
    bits = g.integers(0, 2, n).astype(F32)
    cases = [
        (0, v, a, b, od.normal_logpdf(v, a, b)), (8, v, a, b, od.normal_logpdf(v, a, b)),
        (1, unit, np.zeros(n, F32), b + 1, od.uniform_logpdf(unit, F32(0.0), b + 1)),
        (2, pos, b, b, od.exponential_logpdf(pos, b)), (3, pos, b, b, od.half_normal_logpdf(pos, b)),
        (4, pos, b + 0.5, pos, od.gamma_logpdf(pos, b + 0.5, pos)), (5, unit, b + 0.5, pos, od.beta_logpdf(unit, b + 0.5, pos)),
        (6, bits, unit, unit, od.flip_logpdf(bits.astype(bool), unit)), (7, bits, a, a, od.bernoulli_logpdf(bits.astype(bool), a)),
        (9, v, a, b, od.cauchy_logpdf(v, a, b)), (10, a + pos, a, b, od.half_cauchy_logpdf(a + pos, a, b)),
        (10, a - pos, a, b, od.half_cauchy_logpdf(a - pos, a, b)), (11, v, a, b, od.laplace_logpdf(v, a, b)),
        (12, pos, a, b, od.log_normal_logpdf(pos, a, b)), (12, -pos, a, b, od.log_normal_logpdf(-pos, a, b)),
        (13, v, a, b, od.gumbel_logpdf(v, a, b)), (14, pos, b + 0.3, b, od.weibull_logpdf(pos, b + 0.3, b)),
        (15, unit, b + 0.3, pos, od.kumaraswamy_logpdf(unit, b + 0.3, pos)), (15, unit + 1, b, b, od.kumaraswamy_logpdf(unit + 1, b, b)),
        (16, unit, a, b, od.logit_normal_logpdf(unit, a, b)), (16, -unit, a, b, od.logit_normal_logpdf(-unit, a, b)),
        (17, np.floor(pos * 3), unit, unit, od.geometric_logpdf(np.floor(pos * 3), unit)),
        (18, pos, b + 0.5, pos[::-1], od.inverse_gamma_logpdf(pos, b + 0.5, pos[::-1])), (19, pos, b + 0.5, b, od.chi2_logpdf(pos, b + 0.5)),
    ]
    for which, x, p, q, want in cases:
        out = np.zeros(n, dtype=F32)
        x, p, q = (np.ascontiguousarray(t, dtype=F32) for t in (x, p, q))
        lib.h_logpdf(C.c_int(which), _p(x), _p(p), _p(q), C.c_int(n), _p(out))
        np.testing.assert_allclose(out, want, rtol=3e-6, atol=3e-6, err_msg=f"logpdf case {which}")
    # the inverse-CDF samplers of the long-tail wrappers on the oracle's own draws of lanes 77.. at site 2
    words, idx = (0x12345678, 0x9ABCDEF0), np.arange(n, dtype=np.uint64) + np.uint64(77)
    u_q, z_q = rng.quad_u01(words, idx, 2), rng.quad_normal(words, idx, 2)
    for which, name, p, q in ((9, "cauchy", a, b), (10, "half_cauchy", a, b), (11, "laplace", a, b), (12, "log_normal", a / 4, b / 3),
                              (13, "gumbel", a, b), (14, "weibull", b + 0.3, b), (15, "kumaraswamy", b + 0.3, pos),
                              (16, "logit_normal", a, b), (17, "geometric", unit, None)):
        out = np.zeros(n, dtype=F32)
        draw = np.ascontiguousarray(z_q if name in ("log_normal", "logit_normal") else u_q, dtype=F32)
        p = np.ascontiguousarray(p, dtype=F32)
        params = (p,) if q is None else (p, np.ascontiguousarray(q, dtype=F32))
        lib.h_sample(C.c_int(which), _p(draw), _p(p), _p(params[-1]), C.c_int(n), _p(out))
        want = od.DISTS[name][0](words, idx, 2, *params)
        if name == "geometric":  # floor of a quotient: an ulp can move a draw sitting on an integer boundary
            assert np.mean(out != want) < 0.002 and np.abs(out - want).max() <= 1
            continue
        np.testing.assert_allclose(out, want, rtol=2e-5, atol=2e-6, err_msg=name)  # libm tanf / expf vs rounded float64
    logits = np.array([0.1, -0.4, 1.3, 0.0, -2.0], dtype=F32)
    u = np.ascontiguousarray(g.random(n), dtype=F32)
    draws, lp = np.zeros(n, dtype=np.int32), np.zeros(n, dtype=F32)
    lib.h_categorical(_p(logits), C.c_int(5), _p(u), C.c_int(n), _p(draws), _p(lp))
    np.testing.assert_allclose(lp, od.categorical_logpdf(draws, logits), rtol=3e-6, atol=3e-6)
    freq = np.bincount(draws, minlength=5) / n
    np.testing.assert_allclose(freq, np.exp(logits) / np.exp(logits).sum(), atol=0.03)
    words, idx = (0x12345678, 0x9ABCDEF0), np.arange(n, dtype=np.uint64) + np.uint64(77)
    ga, be = np.zeros(n, dtype=F32), np.zeros(n, dtype=F32)
    rates = np.ascontiguousarray(np.where(np.arange(n) % 3 == 0, 0.2 + pos, 10 + 30 * pos), dtype=F32)  # both sampler branches
    lib.h_poisson(C.c_uint32(words[0]), C.c_uint32(words[1]), C.c_uint64(77), C.c_int(n), C.c_uint32(4), _p(rates), _p(ga), _p(be))
    want_k = od.poisson_sample(words, idx, 4, rates)
    assert np.mean(ga != want_k) < 0.002  # an ulp of expf / logf can move a draw sitting on an acceptance boundary
    np.testing.assert_allclose(be, od.poisson_logpdf(ga, rates), rtol=1e-5, atol=2e-4)  # k log(rate) - lgamma(k + 1): terms near 600, one ulp is 6e-5
    for aa, bb in ((2.5, 1.5), (0.4, 2.0)):
        lib.h_gamma_beta(C.c_uint32(words[0]), C.c_uint32(words[1]), C.c_uint64(77), C.c_int(n), C.c_uint32(3), C.c_float(aa),
                         C.c_float(bb), _p(ga), _p(be))
        want_g, want_b = od.gamma_sample(words, idx, 3, F32(aa), F32(bb)), od.beta_sample(words, idx, 3, F32(aa), F32(bb))
        # a one-ulp libm difference can flip a rejection: allow a handful of lanes to take a different attempt
        assert np.mean(~np.isclose(ga, want_g, rtol=2e-5, atol=1e-6)) < 0.002
        assert np.mean(~np.isclose(be, want_b, rtol=2e-5, atol=1e-6)) < 0.002
        lib.h_inverse_gamma_chi2(C.c_uint32(words[0]), C.c_uint32(words[1]), C.c_uint64(77), C.c_int(n), C.c_uint32(3), C.c_float(aa),
                                 C.c_float(bb), _p(ga), _p(be))
        assert np.mean(~np.isclose(ga, od.inverse_gamma_sample(words, idx, 3, F32(aa), F32(bb)), rtol=2e-5, atol=1e-6)) < 0.002
        assert np.mean(~np.isclose(be, od.chi2_sample(words, idx, 3, F32(2 * aa)), rtol=2e-5, atol=1e-6)) < 0.002
        lib.h_student_t(C.c_uint32(words[0]), C.c_uint32(words[1]), C.c_uint64(77), C.c_int(n), C.c_uint32(3), C.c_float(2 * aa),
                        C.c_float(0.5), C.c_float(bb), _p(ga), _p(be))
        assert np.mean(~np.isclose(ga, od.student_t_sample(words, idx, 3, F32(2 * aa), F32(0.5), F32(bb)), rtol=5e-5, atol=5e-6)) < 0.002
        np.testing.assert_allclose(be, od.student_t_logpdf(ga, F32(2 * aa), F32(0.5), F32(bb)), rtol=3e-6, atol=3e-6)

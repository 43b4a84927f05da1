"""Pins the oracle's counter-based RNGs against published known-answer vectors
(tests/golden/reference_kats.json) and checks the bits->variate maps."""
import json
import os

import numpy as np

from oracle import rng

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_kats.json")))


def _h(xs):
    return [int(x, 16) for x in xs]


def test_threefry2x32_random123_kat():
    for v in GOLD["threefry2x32_20"]["vectors"]:
        out = rng.threefry2x32(np.array(_h(v["key"]), dtype=np.uint32), np.array(_h(v["ctr"]), dtype=np.uint32))
        assert [int(x) for x in out] == _h(v["out"])


def test_threefry_host_key_tree_matches_oracle():
    # the product's host key tree (genjax_b200/core/key.py) and the oracle's are two implementations
    from genjax_b200.core import key as gk

    for v in GOLD["threefry2x32_20"]["vectors"]:
        k, c = _h(v["key"]), _h(v["ctr"])
        assert list(gk.threefry2x32(k[0], k[1], c[0], c[1])) == _h(v["out"])
        a, b = gk.threefry2x32_np(k[0], k[1], c[0], c[1])
        assert [int(a), int(b)] == _h(v["out"])
    ok, pk = rng.key(314159), gk.key(314159)
    assert ok.words == pk.words == (0, 314159)
    for d in (0, 1, 7, 2**31 + 5):
        assert rng.fold_in(ok, d).words == gk.fold_in(pk, d).words
    so, sp = rng.split(ok, 5), gk.split(pk, 5)
    assert so.words == sp.words and so.n == sp.n
    # lane collapse (a split lane used as a parent key)
    assert rng.split(so[3], 2).words == gk.split(sp[3], 2).words
    assert rng.fold_in(so[4], 9).words == gk.fold_in(sp[4], 9).words


def test_pf_key_table_matches_oracle_step_keys():
    from genjax_b200.core import key as gk
    from oracle import smc

    tab = gk.pf_key_table(gk.key(99), 6)
    for t in range(6):
        k_prop, k_res = smc.pf_step_keys(rng.key(99), t)
        lanes = rng.split(k_prop, 4)
        assert (int(tab[t, 0]), int(tab[t, 1])) == lanes.words
        assert (int(tab[t, 2]), int(tab[t, 3])) == k_res.words
        assert int(tab[t, 4]) | (int(tab[t, 5]) << 32) == k_res.index


def test_philox4x32_10_random123_kat():
    for v in GOLD["philox4x32_10"]["vectors"]:
        c, k = _h(v["ctr"]), _h(v["key"])
        out = rng.philox4x32_10(c[0], c[1], c[2], c[3], k[0], k[1])
        assert [int(x) for x in out] == _h(v["out"])


def test_u01_open_interval_and_exact():
    bits = np.array([0, 1, 511, 512, 0xFFFFFFFF, 0x80000000], dtype=np.uint32)
    u = rng.u01(bits)
    assert u.dtype == np.float32
    assert np.all(u > 0) and np.all(u < 1)
    assert u[0] == np.float32(2.0**-24) and u[4] == np.float32(1 - 2.0**-24)


def test_normal_stream_moments_and_layout():
    words = (0x12345678, 0x9ABCDEF0)
    idx = np.arange(200_000, dtype=np.uint64)
    z = rng.normal_vec(words, idx, site=3, d=6)
    assert abs(z.mean()) < 5e-3 and abs(z.std() - 1) < 5e-3
    # kurtosis of a normal is 3
    assert abs(((z - z.mean()) ** 4).mean() / z.var() ** 2 - 3) < 0.05
    # lane i of a batch == the scalar lane (vmap-over-split-keys property)
    z7 = rng.normal_vec(words, np.array([7], dtype=np.uint64), site=3, d=6)
    assert np.array_equal(z7[0], z[7])
    # distinct sites / chunks are distinct streams
    z2 = rng.normal_vec(words, idx[:1000], site=4, d=6)
    assert abs(np.corrcoef(z[:1000, 0], z2[:, 0])[0, 1]) < 0.1

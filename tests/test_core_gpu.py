"""GPU parity of the model-independent kernels (libgjb_core.so) against the
oracle: RNG bit-exact, exact integer LSE mass, bit-exact resample indices."""

import math

import numpy as np
import pytest
import torch

from oracle import rng as orng
from oracle import smc as osmc

pytestmark = pytest.mark.gpu


def _ops():
    from genjax_b200.runtime import smc_ops

    return smc_ops


def test_philox_words_bit_exact(device):
    ops = _ops()
    words = (0xA4093822, 0x299F31D0)
    n, off = 5000, (1 << 32) - 100  # crosses the 32-bit lane boundary
    got = ops.philox_words(words, off, 3, 7, n, device).cpu().numpy().view(np.uint32)
    idx = np.arange(n, dtype=np.uint64) + np.uint64(off)
    exp = np.stack(orng.site_words(words, idx, 3, 7), axis=1)
    assert np.array_equal(got, exp)


def test_philox_random123_kat(device):
    ops = _ops()
    got = ops.philox_words((0, 0), 0, 0, 0, 1, device).cpu().numpy().view(np.uint32)[0]
    assert [hex(int(x)) for x in got] == ["0x6627e8d5", "0xe169c58d", "0xbc57ac4c", "0x9b00dbd8"]


@pytest.mark.parametrize("d", [1, 3, 8, 32])
def test_normal_fill_matches_oracle(device, d):
    ops = _ops()
    words = (123, 456)
    n = 4096
    got = ops.normal_fill(words, 17, 2, n, d, device).cpu().numpy()
    exp = orng.normal_vec(words, np.arange(n, dtype=np.uint64) + np.uint64(17), 2, d)
    np.testing.assert_allclose(got, exp, rtol=2e-6, atol=2e-6)
    assert abs(got.mean()) < 5 / math.sqrt(n * d) and abs(got.std() - 1) < 0.05


@pytest.mark.parametrize("n", [1, 7, 2048, 2049, 100_003, 1 << 20])
def test_lse_terms_exact(device, n):
    ops = _ops()
    g = np.random.default_rng(n)
    lw = (g.standard_normal(n) * 3 - 5).astype(np.float32)
    if n > 10:
        lw[3] = -np.inf
        lw[5] = np.nan
    ws = ops.WeightWorkspace(n, device)
    terms = ws.lse_terms(torch.from_numpy(lw).to(device)).cpu().numpy()
    M, S = osmc.lse_terms(lw)
    assert terms[0] == float(M)
    assert int(terms[1]) == S  # exact integer mass
    assert terms[2] == pytest.approx(osmc.log_mean_exp(lw), rel=1e-12, abs=1e-12)
    finite = lw[np.isfinite(lw)].astype(np.float64)
    ref = np.log(np.sum(np.exp(finite - finite.max()))) + finite.max() - math.log(n)
    assert terms[2] == pytest.approx(ref, abs=2e-6)


@pytest.mark.parametrize("n,scale", [(1, 1.0), (5, 1.0), (2048, 1.0), (4097, 3.0), (100_003, 0.1), (1 << 20, 2.0), (50_000, 30.0)])
def test_systematic_ancestors_bit_exact(device, n, scale):
    ops = _ops()
    g = np.random.default_rng(n + 1)
    lw = (g.standard_normal(n) * scale).astype(np.float32)
    key = orng.Key((0xDEADBEEF, 0x12345678), 9)
    exp = osmc.resample_systematic(lw, key)
    from genjax_b200.core.key import PRNGKey

    ws = ops.WeightWorkspace(n, device)
    lw_d = torch.from_numpy(lw).to(device)
    ws.lse_terms(lw_d)
    anc = torch.full((n,), -1, dtype=torch.int32, device=device)
    ws.systematic(lw_d, PRNGKey(key.words, key.index), anc)
    got = anc.cpu().numpy()
    assert np.array_equal(got, exp)
    assert np.all(np.diff(got) >= 0)


def test_systematic_degenerate_weights(device):
    ops = _ops()
    from genjax_b200.core.key import PRNGKey

    n = 10_000
    lw = np.full(n, -np.inf, dtype=np.float32)
    lw[1234] = 0.0
    ws = ops.WeightWorkspace(n, device)
    lw_d = torch.from_numpy(lw).to(device)
    ws.lse_terms(lw_d)
    anc = torch.empty(n, dtype=torch.int32, device=device)
    ws.systematic(lw_d, PRNGKey((1, 2), 0), anc)
    assert torch.all(anc == 1234)
    # all-zero mass: identity ancestors
    lw[:] = -np.inf
    lw_d = torch.from_numpy(lw).to(device)
    terms = ws.lse_terms(lw_d)
    ws.systematic(lw_d, PRNGKey((1, 2), 0), anc)
    assert terms[1].item() == 0
    assert torch.equal(anc.cpu(), torch.arange(n, dtype=torch.int32))


def test_systematic_sharded_equals_single(device):
    """Two shards using the global max / mass / offsets reproduce the 1-GPU ancestors."""
    ops = _ops()
    from genjax_b200.core.key import PRNGKey

    n = 30_000
    g = np.random.default_rng(5)
    lw = (g.standard_normal(n) * 2).astype(np.float32)
    key = orng.Key((7, 8), 3)
    exp = osmc.resample_systematic(lw, key)
    M, S = osmc.lse_terms(lw)
    half = n // 2
    m_glob = torch.tensor([float(M)], dtype=torch.float32, device=device)
    s_tot = torch.tensor([S], dtype=torch.int64, device=device)
    out = np.full(n, -1, dtype=np.int32)
    offs = [0, int(osmc.det_exp_q((lw[:half] - M).astype(np.float32)).sum(dtype=np.uint64))]
    for r, (lo, hi) in enumerate([(0, half), (half, n)]):
        part = torch.from_numpy(lw[lo:hi]).to(device)
        ws = ops.WeightWorkspace(hi - lo, device)
        ws.mass_pass(part, m_global=m_glob)
        c_off = torch.tensor([offs[r]], dtype=torch.int64, device=device)
        anc = torch.full((n,), -1, dtype=torch.int32, device=device)
        # each shard writes the offspring it OWNS (whole window here), global ancestor ids
        ws.systematic(part, PRNGKey(key.words, key.index), anc, n_total=n, out_lo=0, anc_base=lo,
                      m_global=m_glob, c_offset=c_off, s_total=s_tot)
        a = anc.cpu().numpy()
        mask = a >= 0
        assert np.all(out[mask] == -1)
        out[mask] = a[mask]
    assert np.array_equal(out, exp)


@pytest.mark.parametrize("n", [1, 100, 5000, 70_000])
def test_multinomial_ancestors_bit_exact(device, n):
    ops = _ops()
    g = np.random.default_rng(n + 2)
    lw = (g.standard_normal(n) * 1.5).astype(np.float32)
    kb = orng.KeyBatch((11, 22), n, 5)
    exp = osmc.resample_multinomial(lw, kb)
    ws = ops.WeightWorkspace(n, device)
    lw_d = torch.from_numpy(lw).to(device)
    ws.lse_terms(lw_d)
    anc = torch.empty(n, dtype=torch.int32, device=device)
    cdf = torch.empty(n, dtype=torch.int64, device=device)
    ws.multinomial(lw_d, kb.words, kb.offset, anc, cdf)
    assert np.array_equal(anc.cpu().numpy(), exp)


@pytest.mark.parametrize("shape", [(1000,), (1000, 3), (513, 32), (100, 8)])
def test_gather_rows(device, shape):
    ops = _ops()
    g = torch.Generator().manual_seed(0)
    src = torch.randn(shape, generator=g).to(device)
    anc = torch.randint(0, shape[0], (777,), generator=g, dtype=torch.int32).to(device)
    out = ops.gather_rows(src, anc)
    assert torch.equal(out, src[anc.long()])


def test_device_key_table_equals_the_host_key_tree(device):
    """gjb_pf_key_table: the filter's per-step keys derived on the device == core/key.py pf_key_table (host threefry)."""
    import genjax_b200 as gj
    from genjax_b200.core.key import pf_key_table
    from genjax_b200.runtime import cabi

    for seed, T in ((314159, 100), (7, 1), (2**40 + 5, 1000)):
        k = gj.fold_in(gj.key(seed), 3)
        w0, w1 = k.collapsed()
        out = torch.zeros((T, 8), dtype=torch.int32, device=device)
        cabi.check(cabi.core().gjb_pf_key_table(w0, w1, T, out.data_ptr(), cabi.stream_ptr(device)), "gjb_pf_key_table")
        assert np.array_equal(out.cpu().numpy().view(np.uint32), pf_key_table(k, T))

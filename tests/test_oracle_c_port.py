"""The C / OpenMP restatement of the oracle's bootstrap filter (oracle/c/pf_port.c, used as bench.py's CPU arm)
equals the NumPy oracle: same Philox lanes, same float32 operation order, same integer CDF."""
import numpy as np
import pytest

from genjax_b200.core.key import key as pkey, pf_key_table
from oracle import cport, rng
from oracle import smc as osmc

F32 = np.float32


@pytest.fixture(scope="module")
def built():
    if cport.lib() is None:
        pytest.skip("gcc not available")
    return cport


@pytest.mark.parametrize("n,d", [(1, 1), (7, 1), (4099, 1), (50_000, 1), (3001, 8)])
def test_c_port_matches_numpy_oracle(built, n, d):
    T = 5
    a, q, c, r = 0.9, 1.0, 1.0, 0.5
    g = np.random.default_rng(n)
    ys = g.standard_normal((T, d) if d > 1 else T).astype(F32)
    x0 = g.standard_normal((n, d) if d > 1 else n).astype(F32)
    if d == 1:
        def step(h, x_prev):
            x = h.normal("x", F32(a) * x_prev, F32(q))
            h.normal("y", F32(c) * x, F32(r))
            return x
        shared, obs = (), [{"y": F32(y)} for y in ys]
    else:
        qv, rv = np.full(d, q, F32), np.full(d, r, F32)

        def step(h, x_prev, qq, rr):
            x = h.mv_normal_diag("x", F32(a) * x_prev, qq)
            h.mv_normal_diag("y", F32(c) * x, rr)
            return x
        shared, obs = (qv, rv), [{"y": y} for y in ys]
    ref = osmc.particle_filter(step, rng.key(42), x0, obs, shared_args=shared, record=True)
    got = built.pf_lgssm(x0, ys, a, q, c, r, pf_key_table(pkey(42), T))
    np.testing.assert_allclose(got["logw_last"], ref["history"][-1]["logw"], rtol=2e-6, atol=2e-6)
    assert np.array_equal(got["ancestors_last"], ref["history"][-1]["ancestors"])
    np.testing.assert_allclose(got["state"].reshape(x0.shape), ref["state"][0], rtol=2e-6, atol=2e-6)
    np.testing.assert_allclose(got["logz_inc"], ref["logz_inc"], rtol=0, atol=1e-9)


def test_c_port_matches_kalman(built):
    T, n = 30, 200_000
    ys = osmc.simulate_lgssm(0, T, 1, 0.9, 1.0, 1.0, 0.5)[:, 0]
    x0 = np.random.default_rng(1).standard_normal(n).astype(F32)
    got = built.pf_lgssm(x0, ys, 0.9, 1.0, 1.0, 0.5, pf_key_table(pkey(7), T))
    assert got["logz_inc"].sum() == pytest.approx(osmc.kalman_logz(ys, 0.9, 1.0, 1.0, 0.5), abs=0.1)

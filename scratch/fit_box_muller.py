"""Fits the fp32 polynomial coefficients of the bit-reproducible Box-Muller (csrc/gjb_rng.cuh box_muller, oracle/rng.py):
least-squares on Chebyshev nodes in float64 (close to minimax for these smooth kernels), coefficients rounded to fp32.
Prints C hex-float literals and the worst errors over a dense grid."""
import numpy as np
from numpy.polynomial import chebyshev as Ch, polynomial as P

def fit(fn, lo, hi, deg, weight=None, n=4001):
    k = np.arange(n)
    x = 0.5 * (lo + hi) + 0.5 * (hi - lo) * np.cos(np.pi * (k + 0.5) / n)
    y = fn(x)
    w = np.ones_like(x) if weight is None else weight(x)
    V = np.vander(x, deg + 1, increasing=True)
    c, *_ = np.linalg.lstsq(V * w[:, None], y * w, rcond=None)
    return c

# -2 log1p(f) = -2 f + f^2 R2(f),  f in [sqrt(1/2) - 1, sqrt(2) - 1]
lo, hi = np.sqrt(0.5) - 1, np.sqrt(2.0) - 1
def R2(f):
    f = np.where(np.abs(f) < 1e-9, 1e-9, f)
    return (-2 * np.log1p(f) + 2 * f) / (f * f)
c_log = fit(R2, lo, hi, 8)
# sin(2 pi r) = r * S(r^2), cos(2 pi r) = 1 + r^2 C(r^2),  r in [-1/8, 1/8]
def S(z):
    r = np.sqrt(np.maximum(z, 1e-30)); return np.sin(2 * np.pi * r) / r
def Cc(z):
    z = np.maximum(z, 1e-30); r = np.sqrt(z); return (np.cos(2 * np.pi * r) - 1) / z
c_sin = fit(S, 0.0, 1 / 64, 3)
c_cos = fit(Cc, 0.0, 1 / 64, 3)
def hexf(v): return float(np.float32(v)).hex()
print("log R2:", [hexf(v) for v in c_log])
print("sin S :", [hexf(v) for v in c_sin])
print("cos C :", [hexf(v) for v in c_cos])
# error check in float64 with fp32-rounded coefficients
f = np.linspace(lo, hi, 2_000_001)
cl = np.float32(c_log).astype(np.float64)
approx = -2 * f + f * f * P.polyval(f, cl)
print("max abs err of -2log1p:", np.abs(approx + 2 * np.log1p(f)).max(), "max rel:", np.nanmax(np.abs((approx + 2 * np.log1p(f)) / (2 * np.log1p(f) + 1e-300))[np.abs(f) > 1e-6]))
r = np.linspace(-1 / 8, 1 / 8, 2_000_001)
z = r * r
print("sin err:", np.abs(r * P.polyval(z, np.float32(c_sin).astype(np.float64)) - np.sin(2 * np.pi * r)).max())
print("cos err:", np.abs(1 + z * P.polyval(z, np.float32(c_cos).astype(np.float64)) - np.cos(2 * np.pi * r)).max())

"""Counter-based RNG restatement (oracle; test infrastructure only).

What the reference does: every random choice draws from JAX's threefry2x32
PRNG -- ``jax.random.key/split/fold_in`` at static.py:260-263 (per-site key
``fold_in(key, counter)``, counter from 1), smc.py:154,171,299-300,386,
vmap.py:186,201, hmc.py:125,167,180 -- and TFP samplers turn the bits into
variates (tensorflow_probability/__init__.py:52-62).  jax and tfp are not
vendored under /root/reference and are not installable here, so:

  * threefry2x32 and Philox4x32-10 are restated from the published algorithms
    (Salmon et al., SC'11, "Parallel random numbers: as easy as 1, 2, 3") and
    PINNED against the Random123 known-answer vectors (tests/test_oracle_rng.py);
  * the key tree (split / fold_in layout) and the bits->variate maps are this
    build's own, fully specified below; sampled VALUES are "parity unpinned"
    against the reference (nothing in its tests pins one), and bit-level parity
    is defined between this oracle and the CUDA kernels.

Stream layout shared with genjax_b200/csrc/gjb_rng.cuh:

    words = philox4x32_10(ctr=(idx_lo, idx_hi, chunk, site), key=(k0, k1))

``(k0, k1)`` is the batch key, ``idx`` the GLOBAL particle/chain index (so a
result does not depend on how particles are sharded over GPUs), ``site`` the
1-based visit counter of the random choice inside the model (static.py:260-263
starts its counter at 1 too) and ``chunk`` numbers successive 4-word blocks a
site consumes (vector sites: chunk = dim // 4; rejection samplers: attempt).

SCALAR sites driven by one uniform or one normal use QUAD streams: the four
consecutive particles of global quad ``idx >> 2`` share the block
``philox(ctr=(quad_lo, quad_hi, 0, site))`` and particle ``idx`` takes slot
``idx & 3`` -- word ``slot`` for uniform-driven samplers, component ``slot``
of ``(BM(w0,w1), BM(w2,w3))`` for normal-driven ones (one Philox block and
two Box-Muller pairs per four particles instead of per particle).
"""

from __future__ import annotations

import numpy as np

U32 = np.uint32
U64 = np.uint64
_M32 = 0xFFFFFFFF

# --------------------------------------------------------------------------
# threefry2x32 (20 rounds) -- used host-side for the key tree
# --------------------------------------------------------------------------

_ROT = ((13, 15, 26, 6), (17, 29, 16, 24))


def _rotl32(x, r):
    x = x.astype(U32)
    return ((x << U32(r)) | (x >> U32(32 - r))).astype(U32)


def threefry2x32(key, ctr):
    """threefry2x32-20.  key: (2,) uint32; ctr: (..., 2) uint32 -> (..., 2)."""
    key = np.asarray(key, dtype=U32)
    ctr = np.asarray(ctr, dtype=U32)
    ks0, ks1 = key[0], key[1]
    ks2 = U32(0x1BD11BDA) ^ ks0 ^ ks1
    ks = (ks0, ks1, ks2)
    with np.errstate(over="ignore"):
        x0 = (ctr[..., 0] + ks0).astype(U32)
        x1 = (ctr[..., 1] + ks1).astype(U32)
        for g in range(5):
            rots = _ROT[g % 2]
            for r in rots:
                x0 = (x0 + x1).astype(U32)
                x1 = _rotl32(x1, r)
                x1 = x1 ^ x0
            x0 = (x0 + ks[(g + 1) % 3]).astype(U32)
            x1 = (x1 + ks[(g + 2) % 3] + U32(g + 1)).astype(U32)
    return np.stack([x0, x1], axis=-1)


# --------------------------------------------------------------------------
# Philox4x32-10 -- the per-particle device stream
# --------------------------------------------------------------------------

_PH_M0 = U64(0xD2511F53)
_PH_M1 = U64(0xCD9E8D57)
_PH_W0 = 0x9E3779B9
_PH_W1 = 0xBB67AE85


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Philox4x32-10.  All inputs broadcastable uint32 arrays -> 4 uint32 arrays."""
    c0, c1, c2, c3 = np.broadcast_arrays(
        *(np.asarray(c, dtype=U32) for c in (c0, c1, c2, c3))
    )
    k0 = int(k0) & _M32
    k1 = int(k1) & _M32
    for _ in range(10):
        p0 = c0.astype(U64) * _PH_M0
        p1 = c2.astype(U64) * _PH_M1
        hi0 = (p0 >> U64(32)).astype(U32)
        lo0 = (p0 & U64(_M32)).astype(U32)
        hi1 = (p1 >> U64(32)).astype(U32)
        lo1 = (p1 & U64(_M32)).astype(U32)
        c0, c1, c2, c3 = (hi1 ^ c1 ^ U32(k0), lo1, hi0 ^ c3 ^ U32(k1), lo0)
        k0 = (k0 + _PH_W0) & _M32
        k1 = (k1 + _PH_W1) & _M32
    return c0, c1, c2, c3


# --------------------------------------------------------------------------
# key tree (host side; mirrors genjax_b200/core/key.py)
# --------------------------------------------------------------------------


class Key:
    """A PRNG key = 2 threefry words + a 64-bit lane index.

    ``split(key, n)[i]`` is ``Key(child_words, index=i)``: the batch shares one
    pair of words and lane ``i`` is addressed through the Philox counter, so a
    batched GFI call over ``split(key, n)`` and a scalar call with
    ``split(key, n)[i]`` produce the same numbers for lane ``i`` (what
    ``jax.vmap`` over ``jax.random.split`` guarantees in the reference).
    """

    __slots__ = ("words", "index")

    def __init__(self, words, index=0):
        self.words = (int(words[0]) & _M32, int(words[1]) & _M32)
        self.index = int(index)

    def __repr__(self):
        return f"Key({self.words[0]:#010x},{self.words[1]:#010x};{self.index})"


def key(seed: int) -> Key:
    """``jax.random.key(seed)``: key data = [seed >> 32, seed & 0xffffffff]."""
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    return Key((seed >> 32, seed & _M32), 0)


def _collapse(k: Key):
    """Fold the lane index into the words (identity for index 0)."""
    if k.index == 0:
        return k.words
    out = threefry2x32(
        np.array(k.words, dtype=U32),
        np.array([(k.index >> 32) & _M32 ^ 0x5851F42D, k.index & _M32], dtype=U32),
    )
    return (int(out[0]), int(out[1]))


def fold_in(k: Key, data: int) -> Key:
    """``jax.random.fold_in``: hash the key with one 32-bit datum."""
    w = _collapse(k)
    out = threefry2x32(np.array(w, dtype=U32), np.array([0, int(data) & _M32], dtype=U32))
    return Key((int(out[0]), int(out[1])), 0)


class KeyBatch:
    """Lazy result of ``split(key, n)``: n lanes over one pair of words."""

    def __init__(self, words, n, offset=0):
        self.words = (int(words[0]) & _M32, int(words[1]) & _M32)
        self.n = int(n)
        self.offset = int(offset)

    def __len__(self):
        return self.n

    def __getitem__(self, i):
        if isinstance(i, slice):
            start, stop, step = i.indices(self.n)
            assert step == 1
            return KeyBatch(self.words, max(0, stop - start), self.offset + start)
        if i < 0:
            i += self.n
        if not 0 <= i < self.n:
            raise IndexError(i)
        return Key(self.words, self.offset + i)

    def __iter__(self):
        return (self[i] for i in range(self.n))


def split(k: Key, n: int = 2) -> KeyBatch:
    """``jax.random.split``: child words = threefry(key, [0x73706c74, 0])."""
    w = _collapse(k)
    out = threefry2x32(np.array(w, dtype=U32), np.array([0x73706C74, 0], dtype=U32))
    return KeyBatch((int(out[0]), int(out[1])), n, 0)


def fold_in_lanes(k, data: int):
    """``fold_in`` applied lane-wise (``jax.vmap(lambda key: fold_in(key, data))``): the words are hashed with
    ``data``, every lane keeps its index (mirrors genjax_b200/core/key.py ``fold_in_lanes``; the chained key of the
    Scan combinator, scan.py:213, 268)."""
    out = threefry2x32(np.array(k.words, dtype=U32), np.array([0x666F6C64, int(data) & _M32], dtype=U32))
    w = (int(out[0]), int(out[1]))
    if isinstance(k, KeyBatch):
        return KeyBatch(w, k.n, k.offset)
    return Key(w, k.index)


def lanes(k) -> tuple[tuple[int, int], np.ndarray]:
    """(words, uint64 lane indices) for a Key (1 lane) or KeyBatch (n lanes)."""
    if isinstance(k, KeyBatch):
        return k.words, (np.arange(k.n, dtype=np.uint64) + np.uint64(k.offset))
    return k.words, np.array([k.index], dtype=np.uint64)


# --------------------------------------------------------------------------
# bits -> variates (float32, operation order shared with gjb_rng.cuh)
# --------------------------------------------------------------------------

F32 = np.float32
_TWO_NEG23 = F32(2.0**-23)



def fma32(a, b, c):
    """Correctly rounded float32 fused multiply-add ``round32(a * b + c)`` (CUDA ``__fmaf_rn``, C ``fmaf``).

    NumPy has no fma.  The product of two float32 is exact in float64; the float64 sum ``p + c`` may round, and
    rounding that once more to float32 could double-round.  The sum is therefore re-rounded TO ODD (Boldo &
    Melquiond, "Emulation of FMA and correctly rounded sums", 2008): with the exact error of the float64 addition
    from TwoSum, an inexact sum whose last mantissa bit is even is moved to its neighbour on the error's side,
    which is the odd one; rounding a 53-bit round-to-odd value to 24 bits is then the correctly rounded result.
    Checked against glibc ``fmaf`` in tests/test_oracle_rng.py."""
    a = np.asarray(a, dtype=F32).astype(np.float64)
    b = np.asarray(b, dtype=F32).astype(np.float64)
    c = np.asarray(c, dtype=F32).astype(np.float64)
    with np.errstate(invalid="ignore", over="ignore"):
        p = a * b
        s = p + c
        bb = s - p
        err = (p - (s - bb)) + (c - bb)
        even = (s.view(np.int64) & np.int64(1)) == 0
        fix = np.isfinite(s) & (err != 0.0) & even
        s = np.where(fix, np.nextafter(s, np.where(err > 0.0, np.inf, -np.inf)), s)
    return s.astype(F32)


def site_words(words, idx, site, chunk=0):
    """The 4 Philox words of (lane idx, site, chunk)."""
    idx = np.asarray(idx, dtype=np.uint64)
    lo = (idx & np.uint64(_M32)).astype(U32)
    hi = (idx >> np.uint64(32)).astype(U32)
    chunk = np.broadcast_to(np.asarray(chunk, dtype=U32), lo.shape) if np.ndim(chunk) == 0 else np.asarray(chunk, dtype=U32)
    return philox4x32_10(lo, hi, chunk, U32(site), words[0], words[1])


def u01(bits):
    """uint32 -> float32 in (0,1): ((bits >> 9) + 0.5) * 2^-23 (exact in fp32)."""
    return ((bits >> U32(9)).astype(F32) + F32(0.5)) * _TWO_NEG23


# Box-Muller on fp32 polynomials with correctly rounded FMAs (scratch/fit_box_muller.py fitted the coefficients): every
# operation below is one IEEE fp32 operation of csrc/gjb_rng.cuh box_muller, so sampled normals are BIT-EXACT between
# this oracle and the device (round 1 used libm log / sincospi on the device: rtol 1e-5 only).
_BM_LOG = [F32(float.fromhex(h)) for h in (
    "0x1.fffffep-1", "-0x1.55554ep-1", "0x1.000206p-1", "-0x1.99a3ecp-2", "0x1.548882p-2", "-0x1.22973ap-2",
    "0x1.0c524cp-2", "-0x1.0696e4p-2", "0x1.4237fep-3")]          # -2 log1p(f) = -2 f + f^2 R(f)
_BM_SIN = [F32(float.fromhex(h)) for h in ("0x1.921fb6p+2", "-0x1.4abbbap+5", "0x1.465ec4p+6", "-0x1.2d9b7cp+6")]  # sin(2 pi r) = r S(r^2)
_BM_COS = [F32(float.fromhex(h)) for h in ("-0x1.3bd3ccp+4", "0x1.03c1eap+6", "-0x1.55cb9ap+6", "0x1.db6578p+5")]  # cos(2 pi r) = 1 + r^2 C(r^2)
_BM_NEG_2LN2 = F32(float.fromhex("-0x1.62e43p+0"))
_MAGIC = F32(12582912.0)  # 1.5 * 2^23: adding it rounds to the nearest integer, which lands in the low mantissa bits


def box_muller(b0, b1):
    """Two N(0,1) float32 from two uint32 words: r = sqrt(-2 ln u1), (z0, z1) = r (cos, sin)(2 pi u2)."""
    u1 = np.asarray(u01(b0), dtype=F32)
    u2 = np.asarray(u01(b1), dtype=F32)
    # -2 ln u1: u1 = 2^e m, m in [sqrt(1/2), sqrt(2)), f = m - 1 (exact)
    tb = u1.view(np.int32) - np.int32(0x3F3504F3)
    e = tb >> np.int32(23)
    m = ((tb & np.int32(0x007FFFFF)) + np.int32(0x3F3504F3)).view(F32)
    f = (m - F32(1.0)).astype(F32)
    ef = e.astype(F32)
    f2 = (f * f).astype(F32)
    R = np.full(f.shape, _BM_LOG[8], dtype=F32)
    for k in range(7, -1, -1):
        R = fma32(R, f, _BM_LOG[k])
    L = fma32(ef, _BM_NEG_2LN2, fma32(f2, R, (F32(-2.0) * f).astype(F32)))
    r = np.sqrt(L).astype(F32)
    # angle 2 pi u2 = j pi/2 + 2 pi rr, j = rint(4 u2), rr in [-1/8, 1/8] (exact)
    tm = fma32(u2, F32(4.0), _MAGIC)
    q = tm.view(np.int32) & np.int32(3)
    jf = (tm - _MAGIC).astype(F32)
    rr = fma32(jf, F32(-0.25), u2)
    z = (rr * rr).astype(F32)
    ps = fma32(z, _BM_SIN[3], _BM_SIN[2])
    ps = fma32(z, ps, _BM_SIN[1])
    ps = fma32(z, ps, _BM_SIN[0])
    sn = (rr * ps).astype(F32)
    pc = fma32(z, _BM_COS[3], _BM_COS[2])
    pc = fma32(z, pc, _BM_COS[1])
    pc = fma32(z, pc, _BM_COS[0])
    cs = fma32(z, pc, F32(1.0))
    swap = (q & 1) != 0
    a = np.where(swap, sn, cs)
    b = np.where(swap, cs, sn)
    a = np.where(((q + 1) & 2) != 0, -a, a).astype(F32)
    b = np.where((q & 2) != 0, -b, b).astype(F32)
    return (r * a).astype(F32), (r * b).astype(F32)


def normal4(words, idx, site, chunk=0):
    """Four N(0,1) variates of one Philox block: (z0,z1)=BM(w0,w1), (z2,z3)=BM(w2,w3)."""
    w0, w1, w2, w3 = site_words(words, idx, site, chunk)
    z0, z1 = box_muller(w0, w1)
    z2, z3 = box_muller(w2, w3)
    return z0, z1, z2, z3


def quad_slot_words(words, idx, site):
    """The word slot ``idx & 3`` of the quad block of each lane (uint32 [n])."""
    idx = np.asarray(idx, dtype=np.uint64)
    w = site_words(words, idx >> np.uint64(2), site, 0)
    sub = (idx & np.uint64(3)).astype(np.int64)
    return np.choose(sub, w)


def quad_u01(words, idx, site):
    """One uniform per lane from the quad stream of a scalar site."""
    return u01(quad_slot_words(words, idx, site))


def quad_normal(words, idx, site):
    """One N(0,1) per lane from the quad stream of a scalar site."""
    idx = np.asarray(idx, dtype=np.uint64)
    z = normal4(words, idx >> np.uint64(2), site, 0)
    sub = (idx & np.uint64(3)).astype(np.int64)
    return np.choose(sub, z).astype(F32)


def normal_vec(words, idx, site, d):
    """[n, d] standard normals: dim j comes from chunk j // 4, slot j % 4."""
    idx = np.asarray(idx, dtype=np.uint64)
    out = np.empty((idx.shape[0], d), dtype=F32)
    for c in range((d + 3) // 4):
        z = normal4(words, idx, site, c)
        for s in range(4):
            j = 4 * c + s
            if j < d:
                out[:, j] = z[s]
    return out

"""PRNG keys for the particle-parallel engine.

Mirrors the ``jax.random.key / split / fold_in`` surface the reference uses
(static.py:260-263, inference/smc.py:154,171,299-300,386, vmap.py:186,201,
hmc.py:125,167,180) without materialising one key per particle: a key is two
threefry words plus a 64-bit lane index, and ``split(key, n)`` is a lazy
``KeyBatch`` whose lane ``i`` is addressed through the Philox counter inside
the fused kernels (csrc/gjb_rng.cuh).  A batched GFI call over
``split(key, n)`` and a scalar call with ``split(key, n)[i]`` give the same
numbers for lane ``i`` -- the property ``jax.vmap`` over split keys has in the
reference.

The key tree (threefry2x32-20 over the words) is host-side integer work; the
exact JAX counter layout cannot be verified in this image (no jax), so sampled
values are distribution-equal, not bit-equal, to the reference's.
"""

from __future__ import annotations

_M32 = 0xFFFFFFFF
_ROT = ((13, 15, 26, 6), (17, 29, 16, 24))


def _rotl(x, r):
    return ((x << r) | (x >> (32 - r))) & _M32


def threefry2x32(k0: int, k1: int, c0: int, c1: int) -> tuple[int, int]:
    """threefry2x32-20 (Salmon et al. 2011) on Python ints."""
    ks = (k0 & _M32, k1 & _M32, (0x1BD11BDA ^ k0 ^ k1) & _M32)
    x0 = (c0 + ks[0]) & _M32
    x1 = (c1 + ks[1]) & _M32
    for g in range(5):
        for r in _ROT[g % 2]:
            x0 = (x0 + x1) & _M32
            x1 = _rotl(x1, r) ^ x0
        x0 = (x0 + ks[(g + 1) % 3]) & _M32
        x1 = (x1 + ks[(g + 2) % 3] + g + 1) & _M32
    return x0, x1


class PRNGKey:
    """Two key words + lane index (see module docstring)."""

    __slots__ = ("words", "index")

    def __init__(self, words, index: int = 0):
        self.words = (int(words[0]) & _M32, int(words[1]) & _M32)
        self.index = int(index)

    def collapsed(self) -> tuple[int, int]:
        if self.index == 0:
            return self.words
        return threefry2x32(
            self.words[0], self.words[1], ((self.index >> 32) & _M32) ^ 0x5851F42D, self.index & _M32
        )

    def __repr__(self):
        return f"PRNGKey({self.words[0]:#010x}, {self.words[1]:#010x}; lane {self.index})"

    def __eq__(self, other):
        return isinstance(other, PRNGKey) and self.words == other.words and self.index == other.index

    def __hash__(self):
        return hash((self.words, self.index))


class KeyBatch:
    """Lazy ``split(key, n)``: n lanes sharing one pair of words."""

    __slots__ = ("words", "n", "offset")

    def __init__(self, words, n: int, offset: int = 0):
        self.words = (int(words[0]) & _M32, int(words[1]) & _M32)
        self.n = int(n)
        self.offset = int(offset)

    def __len__(self):
        return self.n

    @property
    def shape(self):
        return (self.n,)

    def __getitem__(self, i):
        if isinstance(i, slice):
            start, stop, step = i.indices(self.n)
            if step != 1:
                raise IndexError("KeyBatch slices must be contiguous")
            return KeyBatch(self.words, max(0, stop - start), self.offset + start)
        i = int(i)
        if i < 0:
            i += self.n
        if not 0 <= i < self.n:
            raise IndexError(i)
        return PRNGKey(self.words, self.offset + i)

    def __iter__(self):
        return (self[i] for i in range(self.n))

    def __repr__(self):
        return f"KeyBatch({self.words[0]:#010x}, {self.words[1]:#010x}; lanes [{self.offset}, {self.offset + self.n}))"


def key(seed: int) -> PRNGKey:
    """``jax.random.key(seed)``."""
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    return PRNGKey((seed >> 32, seed & _M32), 0)


PRNGKey.from_seed = staticmethod(key)


def fold_in(k: PRNGKey, data: int) -> PRNGKey:
    """``jax.random.fold_in(key, data)``."""
    w = k.collapsed()
    return PRNGKey(threefry2x32(w[0], w[1], 0, int(data) & _M32), 0)


def split(k: PRNGKey, num: int = 2) -> KeyBatch:
    """``jax.random.split(key, num)`` (lazy)."""
    w = k.collapsed()
    return KeyBatch(threefry2x32(w[0], w[1], 0x73706C74, 0), num, 0)


def key_children(k, num: int = 2) -> list:
    """``split`` that also works lane-wise on a ``KeyBatch``: child j of a batch is the batch whose words are
    hashed with j and whose lanes are unchanged (what ``jax.vmap(jax.random.split)`` over the batch gives, one
    child batch per column).  For a scalar key this is ``list(split(k, num))``."""
    if isinstance(k, KeyBatch):
        return [KeyBatch(threefry2x32(k.words[0], k.words[1], 0x73706C74, j + 1), k.n, k.offset) for j in range(num)]
    return list(split(k, num))


def fold_in_lanes(k, data: int):
    """``fold_in`` applied lane-wise -- ``jax.vmap(lambda key: jax.random.fold_in(key, data))(keys)``: the words
    are hashed with ``data`` and every lane keeps its index, for a ``KeyBatch`` and for a single lane alike, so a
    batched call and a scalar call on lane i of the batch stay in step (the chained key of ``Scan``, scan.py:213)."""
    w = threefry2x32(k.words[0], k.words[1], 0x666F6C64, int(data) & _M32)
    if isinstance(k, KeyBatch):
        return KeyBatch(w, k.n, k.offset)
    if isinstance(k, PRNGKey):
        return PRNGKey(w, k.index)
    raise TypeError(f"expected a PRNGKey or KeyBatch, got {type(k).__name__}")


def lanes_of(k) -> tuple[tuple[int, int], int, int]:
    """(words, first lane, n lanes) of a PRNGKey (1 lane) or KeyBatch."""
    if isinstance(k, KeyBatch):
        return k.words, k.offset, k.n
    if isinstance(k, PRNGKey):
        return k.words, k.index, 1
    raise TypeError(f"expected a PRNGKey or KeyBatch, got {type(k).__name__}")


# ---------------------------------------------------------------- vectorised


def threefry2x32_np(k0, k1, c0, c1):
    """NumPy threefry2x32-20 over broadcastable uint32 arrays (host-side key tables)."""
    import numpy as np

    U = np.uint32
    k0, k1, c0, c1 = (np.asarray(a, dtype=U) for a in np.broadcast_arrays(k0, k1, c0, c1))
    ks = (k0, k1, (U(0x1BD11BDA) ^ k0 ^ k1).astype(U))
    with np.errstate(over="ignore"):
        x0 = (c0 + ks[0]).astype(U)
        x1 = (c1 + ks[1]).astype(U)
        for g in range(5):
            for r in _ROT[g % 2]:
                x0 = (x0 + x1).astype(U)
                x1 = ((x1 << U(r)) | (x1 >> U(32 - r))).astype(U) ^ x0
            x0 = (x0 + ks[(g + 1) % 3]).astype(U)
            x1 = (x1 + ks[(g + 2) % 3] + U(g + 1)).astype(U)
    return x0, x1


def pf_key_table(k: PRNGKey, n_steps: int):
    """uint32 [T, 8] rows {prop_k0, prop_k1, res_k0, res_k1, res_idx_lo, res_idx_hi, mn_k0, mn_k1}.

    Step t uses ``k_prop, k_res = split(fold_in(key, t))``; the proposal lanes
    are ``split(k_prop, N)``, the systematic resampler draws its one uniform from
    ``k_res`` and the multinomial resampler gives offspring j lane j of
    ``split(k_res, N)`` (words mn_k0, mn_k1).
    """
    import numpy as np

    w = k.collapsed()
    t = np.arange(n_steps, dtype=np.uint32)
    f0, f1 = threefry2x32_np(w[0], w[1], np.uint32(0), t)  # fold_in(key, t)
    s0, s1 = threefry2x32_np(f0, f1, np.uint32(0x73706C74), np.uint32(0))  # split(.) words; lanes 0, 1
    p0, p1 = threefry2x32_np(s0, s1, np.uint32(0x73706C74), np.uint32(0))  # split(k_prop = lane 0, N)
    tab = np.zeros((n_steps, 8), dtype=np.uint32)
    tab[:, 0], tab[:, 1], tab[:, 2], tab[:, 3] = p0, p1, s0, s1
    tab[:, 4] = 1  # k_res = lane 1 of the split
    c0, c1 = threefry2x32_np(s0, s1, np.uint32(0x5851F42D), np.uint32(1))  # k_res collapsed (lane 1 folded into the words)
    tab[:, 6], tab[:, 7] = threefry2x32_np(c0, c1, np.uint32(0x73706C74), np.uint32(0))  # split(k_res, N) words
    return tab

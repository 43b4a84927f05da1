"""Thin Python drivers over the model-independent C-ABI entry points
(libgjb_core.so): log-sum-exp terms, resampling, ancestor gather.

These replace ``jax.scipy.special.logsumexp`` (inference/smc.py:25,97,107),
the categorical draw in ``ParticleCollection.sample_particle``
(inference/smc.py:102-109) and ``tree_map(lambda v: v[idx])``
(inference/smc.py:90-91) of the reference.  Everything is enqueue-only on the
current CUDA stream; nothing here reads results back to the host.
"""

from __future__ import annotations

import torch

from ..core.key import PRNGKey
from . import cabi

TILE = 2048


class WeightWorkspace:
    """Per-collection device scratch: running max, tile masses, LSE triple."""

    def __init__(self, n: int, device):
        self.n = int(n)
        self.device = device
        self.tiles = max(1, (self.n + TILE - 1) // TILE)
        self.wmax = torch.empty(1, dtype=torch.int32, device=device)  # ordered-uint encoded float
        self.tile_mass = torch.empty(self.tiles, dtype=torch.int64, device=device)
        self.lse = torch.empty(3, dtype=torch.float64, device=device)
        self.heavy = torch.zeros(cabi.GJB_HEAVY_WS_WORDS, dtype=torch.int32, device=device)
        self.reset_max()

    def reset_max(self):
        cabi.check(cabi.core().gjb_wmax_reset(self.wmax.data_ptr(), cabi.stream_ptr(self.device)), "gjb_wmax_reset")

    def max_pass(self, logw: torch.Tensor):
        cabi.check(
            cabi.core().gjb_weight_max(cabi.ptr(logw), logw.numel(), self.wmax.data_ptr(), cabi.stream_ptr(self.device)),
            "gjb_weight_max",
        )

    def mass_pass(self, logw: torch.Tensor, m_global: torch.Tensor | None = None):
        cabi.check(
            cabi.core().gjb_weight_mass(
                cabi.ptr(logw), logw.numel(), self.wmax.data_ptr(), cabi.ptr(m_global), self.tile_mass.data_ptr(),
                cabi.stream_ptr(self.device),
            ),
            "gjb_weight_mass",
        )

    def finalize(self, n_total: int | None = None, m_global: torch.Tensor | None = None, out: torch.Tensor | None = None):
        out = self.lse if out is None else out
        cabi.check(
            cabi.core().gjb_lse_finalize(
                self.tile_mass.data_ptr(), self.n, self.wmax.data_ptr(), cabi.ptr(m_global),
                self.n if n_total is None else int(n_total), out.data_ptr(), cabi.stream_ptr(self.device),
            ),
            "gjb_lse_finalize",
        )
        return out

    def lse_terms(self, logw: torch.Tensor, have_max: bool = False) -> torch.Tensor:
        """[M, S, log-mean-exp] (float64, device) of ``logw``."""
        if not have_max:
            self.reset_max()
            self.max_pass(logw)
        self.mass_pass(logw)
        return self.finalize()

    def systematic_args(self, logw: torch.Tensor, key: PRNGKey | None, ancestors: torch.Tensor, *, n_total=None,
                        out_lo=0, anc_base=0, m_global=None, c_offset=None, s_total=None, key_dev=None, lse_out=None,
                        wmax=None, wmax_next=None) -> cabi.ResampleArgs:
        R = cabi.ResampleArgs()
        R.logw = cabi.ptr(logw)
        R.n = logw.numel()
        R.wmax = (self.wmax if wmax is None else wmax).data_ptr()
        R.m_global = cabi.ptr(m_global)
        R.tile_mass = self.tile_mass.data_ptr()
        R.c_offset = cabi.ptr(c_offset)
        R.s_total = cabi.ptr(s_total)
        R.n_total = self.n if n_total is None else int(n_total)
        R.out_lo = int(out_lo)
        R.out_n = ancestors.numel()
        R.anc_base = int(anc_base)
        if key is not None:
            R.key0, R.key1 = key.words
            R.key_index = key.index
        R.key_dev = cabi.ptr(key_dev)
        R.ancestors = ancestors.data_ptr()
        R.lse_out = cabi.ptr(lse_out)
        R.wmax_next = cabi.ptr(wmax_next)
        R.heavy_ws = self.heavy.data_ptr()
        return R

    def systematic(self, logw: torch.Tensor, key: PRNGKey | None, ancestors: torch.Tensor, **kw):
        """Ancestors for offspring [out_lo, out_lo + len(ancestors)); needs mass_pass done."""
        import ctypes as C

        R = self.systematic_args(logw, key, ancestors, **kw)
        cabi.check(cabi.core().gjb_resample_systematic(C.byref(R), cabi.stream_ptr(self.device)),
                   "gjb_resample_systematic")
        return ancestors

    def multinomial(self, logw: torch.Tensor, words, idx_offset: int, ancestors: torch.Tensor, cdf: torch.Tensor):
        cabi.check(
            cabi.core().gjb_resample_multinomial(
                cabi.ptr(logw), logw.numel(), self.wmax.data_ptr(), self.tile_mass.data_ptr(), cdf.data_ptr(),
                words[0], words[1], int(idx_offset), ancestors.numel(), ancestors.data_ptr(),
                cabi.stream_ptr(self.device),
            ),
            "gjb_resample_multinomial",
        )
        return ancestors


class TeWorkspace:
    """Tile-exponent CDF of n log-weights (include/genjax_b200.h section 1c): ``masses(logw)`` then
    ``resample(key, ancestors)`` -- the stand-alone form of what the single-launch filter step does in flight."""

    def __init__(self, n: int, device):
        self.n = int(n)
        self.device = device
        self.tiles = (self.n + cabi.TE_TILE - 1) // cabi.TE_TILE
        if self.tiles > cabi.TE_MAX_TILES:
            raise cabi.GjbError(f"tile-exponent resampling spans at most {cabi.TE_MAX_TILES} tiles")
        self.cdf = torch.empty(self.tiles * cabi.TE_TILE, dtype=torch.int64, device=device)
        self.recs = torch.empty((self.tiles, 2), dtype=torch.int64, device=device)
        self.lse = torch.empty(3, dtype=torch.float64, device=device)
        self._key = torch.empty(4, dtype=torch.int32, device=device)

    def masses(self, logw: torch.Tensor):
        cabi.check(cabi.core().gjb_te_masses(cabi.ptr(logw), logw.numel(), self.cdf.data_ptr(), self.recs.data_ptr(),
                                             cabi.stream_ptr(self.device)), "gjb_te_masses")
        return self

    def resample(self, key: PRNGKey, ancestors: torch.Tensor, out_lo: int = 0) -> torch.Tensor:
        """ancestors[j - out_lo] for offspring j in [out_lo, out_lo + len(ancestors)); ``self.lse`` = {E ln 2, S, log-mean-exp}."""
        import ctypes as C

        import numpy as np

        kd = np.array([key.words[0], key.words[1], key.index & 0xFFFFFFFF, key.index >> 32], dtype=np.uint32).view(np.int32)
        self._key.copy_(torch.from_numpy(kd))
        R = cabi.TeResampleArgs()
        R.cdf, R.recs = self.cdf.data_ptr(), self.recs.data_ptr()
        R.n_tiles_total, R.n_total = self.tiles, self.n
        R.out_lo, R.out_n = int(out_lo), ancestors.numel()
        R.key_dev, R.ancestors, R.lse_out = self._key.data_ptr(), ancestors.data_ptr(), self.lse.data_ptr()
        cabi.check(cabi.core().gjb_te_resample(C.byref(R), cabi.stream_ptr(self.device)), "gjb_te_resample")
        return ancestors


def gather_rows(src: torch.Tensor, ancestors: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
    """out[j] = src[ancestors[j]] over the leading axis (4-byte element types)."""
    if src.element_size() != 4:
        raise cabi.GjbError("gather_rows handles 4-byte element types")
    n_out = ancestors.numel()
    row = 1
    for d in src.shape[1:]:
        row *= d
    if out is None:
        out = torch.empty((n_out,) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device)
    cabi.check(
        cabi.core().gjb_gather_rows(cabi.ptr(src), cabi.ptr(ancestors), cabi.ptr(out), n_out, row * 4,
                                    cabi.stream_ptr(src.device)),
        "gjb_gather_rows",
    )
    return out


def accept_mask(u: torch.Tensor, w: torch.Tensor) -> torch.Tensor:
    """int32 [n]: ``log(u) < w`` (the MH accept test of the reference's tests, test_requests.py:136-137)."""
    n = u.numel()
    mask = torch.empty(n, dtype=torch.int32, device=u.device)
    cabi.check(cabi.core().gjb_accept_mask(cabi.ptr(u.contiguous()), cabi.ptr(w.contiguous()), n, cabi.ptr(mask),
                                           cabi.stream_ptr(u.device)), "gjb_accept_mask")
    return mask


def select_rows(mask: torch.Tensor, a: torch.Tensor, b: torch.Tensor, n: int, event_shape: tuple) -> torch.Tensor:
    """out[i] = a[i] if mask[i] else b[i] over the leading axis; ``a`` / ``b`` of event shape only are broadcast."""
    row = 1
    for d in event_shape:
        row *= d
    a_b = a.numel() == row and n != 1
    b_b = b.numel() == row and n != 1
    if a.dtype != b.dtype or a.element_size() != 4:
        raise cabi.GjbError("select_rows handles equal 4-byte element types")
    out = torch.empty((n,) + tuple(event_shape), dtype=a.dtype, device=mask.device)
    cabi.check(cabi.core().gjb_select_rows(cabi.ptr(mask), cabi.ptr(a.contiguous()), cabi.ptr(b.contiguous()), cabi.ptr(out), n,
                                           row * 4, int(a_b), int(b_b), cabi.stream_ptr(mask.device)), "gjb_select_rows")
    return out


def weight_ess(logw: torch.Tensor, lse3: torch.Tensor) -> torch.Tensor:
    """Effective sample size of ``logw`` given its lse terms (float64 0-d)."""
    out = torch.empty(1, dtype=torch.float64, device=logw.device)
    cabi.check(cabi.core().gjb_weight_ess(cabi.ptr(logw), logw.numel(), cabi.ptr(lse3), cabi.ptr(out),
                                          cabi.stream_ptr(logw.device)), "gjb_weight_ess")
    return out[0]


def philox_words(words, idx_offset, site, chunk, n, device) -> torch.Tensor:
    out = torch.empty((n, 4), dtype=torch.int32, device=device)
    cabi.check(
        cabi.core().gjb_philox_fill(words[0], words[1], int(idx_offset), int(site), int(chunk), n, out.data_ptr(),
                                    cabi.stream_ptr(device)),
        "gjb_philox_fill",
    )
    return out


def normal_fill(words, idx_offset, site, n, d, device) -> torch.Tensor:
    out = torch.empty((n, d), dtype=torch.float32, device=device)
    cabi.check(
        cabi.core().gjb_normal_fill(words[0], words[1], int(idx_offset), int(site), n, d, out.data_ptr(),
                                    cabi.stream_ptr(device)),
        "gjb_normal_fill",
    )
    return out

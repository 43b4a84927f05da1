"""ctypes bindings of the C-ABI in ``include/genjax_b200.h``.

This is the ONLY way the Python host reaches the CUDA kernels; there is no
CPU or PyTorch-op fallback -- a missing library or a non-CUDA tensor raises.
"""

from __future__ import annotations

import ctypes as C
from pathlib import Path

import torch

from . import build

ABI_VERSION = 16  # GJB_ABI_VERSION of include/genjax_b200.h this binding mirrors (tests/test_abi.py keeps them equal)
GJB_MAX_SITES = 32
GJB_MAX_ARGS = 16
GJB_MAX_RETS = 16
GJB_HEAVY_WS_WORDS = 4 + 3 * 1024
TE_TILE = 2048
TE_MAX_TILES = 4096
TE_LL_WORDS = 4
TE_MAILBOX_WORDS = 2 * TE_MAX_TILES * TE_LL_WORDS
STEP_PDL = 1
STEP_FLAGWAIT = 2
STEP_LIGHT = 4
MASS_MAX_PARTICLES = 1 << 27

SITE_SAMPLE = 1
SITE_WEIGHT = 2
SITE_BCAST = 4
CHAIN_HAVE_LOGP = 1
CHAIN_NO_ACCEPT = 2

_p = C.c_void_p
_i64 = C.c_int64
_u64 = C.c_uint64
_u32 = C.c_uint32
_i32 = C.c_int32


class GjbError(RuntimeError):
    pass


class ModelArgs(C.Structure):
    """``gjb_model_args`` (include/genjax_b200.h)."""

    _fields_ = [
        ("n", _i64),
        ("idx_offset", _u64),
        ("key0", _u32),
        ("key1", _u32),
        ("key_dev", _p),
        ("gather", _p),
        ("peer_args", _p),
        ("link", _p),
        ("wait_off", _u64),
        ("push_off", _u64),
        ("args", _p * GJB_MAX_ARGS),
        ("scalars", C.c_float * GJB_MAX_ARGS),
        ("site_in", _p * GJB_MAX_SITES),
        ("site_out", _p * GJB_MAX_SITES),
        ("ret_out", _p * GJB_MAX_RETS),
        ("site_flags", _u32 * GJB_MAX_SITES),
        ("score_in", _p),
        ("weight_in", _p),
        ("score_out", _p),
        ("weight_out", _p),
        ("wmax", _p),
        ("m_ref", _p),
        ("tile_mass", _p),
        ("tile_mass_clear", _p),
        ("tile_mass_clear_n", _i64),
        ("pull_logw", _p),
        ("pull_tile_mass", _p),
        ("pull_m_ref", _p),
        ("pull_key", _p),
        ("pull_ancestors", _p),
        ("pull_lse", _p),
        ("pull_n_total", _i64),
    ]


GJB_MAX_RANKS = 16
XCHG_MAX, XCHG_MASS, XCHG_BARRIER = 0, 1, 2


class Peers(C.Structure):
    """``gjb_peers`` (include/genjax_b200.h)."""

    _fields_ = [("world", _i32), ("rank", _i32), ("n_per_rank", _i64), ("base", _p * GJB_MAX_RANKS),
                ("div_mul", _u32), ("div_shr", _u32)]


GJB_PAD_SLOTS = 4
GJB_PAD_WORDS = GJB_PAD_SLOTS * GJB_MAX_RANKS * 2


class Link(C.Structure):
    """``gjb_link`` (include/genjax_b200.h)."""

    _fields_ = [("rank", _i32), ("world", _i32), ("pads", _p * GJB_MAX_RANKS), ("epoch", _p), ("counter", _p)]


class XchgArgs(C.Structure):
    """``gjb_xchg_args`` (include/genjax_b200.h)."""

    _fields_ = [
        ("rank", _i32),
        ("world", _i32),
        ("mode", _i32),
        ("n_tiles", _i32),
        ("pads", _p * GJB_MAX_RANKS),
        ("epoch", _p),
        ("tag_offset", _u64),
        ("wmax", _p),
        ("tile_mass", _p),
        ("m_global", _p),
        ("c_offset", _p),
        ("s_total", _p),
    ]


class ResampleArgs(C.Structure):
    """``gjb_resample_args`` (include/genjax_b200.h)."""

    _fields_ = [
        ("logw", _p),
        ("n", _i64),
        ("wmax", _p),
        ("m_global", _p),
        ("tile_mass", _p),
        ("c_offset", _p),
        ("s_total", _p),
        ("n_total", _i64),
        ("out_lo", _i64),
        ("out_n", _i64),
        ("anc_base", _i64),
        ("key0", _u32),
        ("key1", _u32),
        ("key_index", _u64),
        ("key_dev", _p),
        ("ancestors", _p),
        ("lse_out", _p),
        ("wmax_next", _p),
        ("heavy_ws", _p),
    ]


class PfArgs(C.Structure):
    """``gjb_pf_args`` (include/genjax_b200.h)."""

    _fields_ = [
        ("n", _i64),
        ("n_total", _i64),
        ("idx_offset", _u64),
        ("T", _i32),
        ("record", _i32),
        ("n_state", _i32),
        ("reserved", _i32),
        ("keys", _p),
        ("state0", _p * GJB_MAX_ARGS),
        ("state_buf", _p * GJB_MAX_ARGS),
        ("state_stride", _i64 * GJB_MAX_ARGS),
        ("shared", _p * GJB_MAX_ARGS),
        ("scalars", C.c_float * GJB_MAX_ARGS),
        ("obs", _p * GJB_MAX_SITES),
        ("obs_stride", _i64 * GJB_MAX_SITES),
        ("site_flags", _u32 * GJB_MAX_SITES),
        ("logw", _p),
        ("ancestors", _p),
        ("lse", _p),
        ("wmax", _p),
        ("cta_mass", _p),
        ("barrier", _p),
    ]


class TileRec(C.Structure):
    """``gjb_tile_rec`` (include/genjax_b200.h section 1c)."""

    _fields_ = [("mass", _u64), ("e", _i32), ("reserved", _i32)]


class StepTable(C.Structure):
    """``gjb_step_table`` (include/genjax_b200.h section 1c)."""

    _fields_ = [("S", _u64), ("E", _i32), ("n_tiles_total", _i32), ("tag", _u32), ("reserved", _u32), ("pre", _u64 * TE_MAX_TILES),
                ("win", (_i32 * 2) * TE_MAX_TILES), ("shf", C.c_uint8 * TE_MAX_TILES)]


class StepLink(C.Structure):
    """``gjb_step_link`` (include/genjax_b200.h section 1c)."""

    _fields_ = [("rank", _i32), ("world", _i32), ("tiles_per_rank", _i32), ("reserved", _i32),
                ("mailbox", _p * GJB_MAX_RANKS), ("epoch", _p), ("ticket", _p)]


class TeTableArgs(C.Structure):
    """``gjb_te_table_args`` (include/genjax_b200.h section 1c)."""

    _fields_ = [("link", _p), ("step", _i32), ("flags", _u32), ("slot_offset", _i64), ("n_local", _i64), ("n_total", _i64),
                ("reskey", _p), ("table_out", _p), ("lse_out", _p)]


class TeResampleArgs(C.Structure):
    """``gjb_te_resample_args`` (include/genjax_b200.h section 1c)."""

    _fields_ = [
        ("cdf", _p),
        ("recs", _p),
        ("table", _p),
        ("cdf_peers", _p),
        ("n_tiles_total", _i32),
        ("reserved", _i32),
        ("n_total", _i64),
        ("out_lo", _i64),
        ("out_n", _i64),
        ("key_dev", _p),
        ("ancestors", _p),
        ("lse_out", _p),
    ]


class StepArgs(C.Structure):
    """``gjb_step_args`` (include/genjax_b200.h section 2)."""

    _fields_ = [
        ("n", _i64),
        ("n_total", _i64),
        ("idx_offset", _u64),
        ("slot_offset", _i64),
        ("step", _i32),
        ("flags", _u32),
        ("key_dev", _p),
        ("args", _p * GJB_MAX_ARGS),
        ("scalars", C.c_float * GJB_MAX_ARGS),
        ("peer_args", _p),
        ("site_in", _p * GJB_MAX_SITES),
        ("state_out", _p * GJB_MAX_RETS),
        ("weight_out", _p),
        ("prev_cdf", _p),
        ("cdf_peers", _p),
        ("table_in", _p),
        ("prev_recs", _p),
        ("n_tiles_total", _i32),
        ("reserved", _i32),
        ("prev_lse", _p),
        ("prev_key", _p),
        ("ancestors_out", _p),
        ("cdf_out", _p),
        ("recs_out", _p),
        ("link", _p),
        ("table_out", _p),
        ("lse_out", _p),
    ]


class StepsArgs(C.Structure):
    """``gjb_steps_args`` (include/genjax_b200.h section 2)."""

    _fields_ = [
        ("n", _i64),
        ("idx_offset", _u64),
        ("T", _i32),
        ("record", _i32),
        ("keys", _p),
        ("state0", _p * GJB_MAX_RETS),
        ("state_buf", _p * GJB_MAX_RETS),
        ("state_stride", _i64 * GJB_MAX_RETS),
        ("shared", _p * GJB_MAX_ARGS),
        ("scalars", C.c_float * GJB_MAX_ARGS),
        ("obs", _p * GJB_MAX_SITES),
        ("obs_stride", _i64 * GJB_MAX_SITES),
        ("logw", _p),
        ("ancestors", _p),
        ("cdf", _p),
        ("recs", _p),
        ("lse", _p),
    ]


class ChainArgs(C.Structure):
    """``gjb_chain_args`` (include/genjax_b200.h)."""

    _fields_ = [
        ("n", _i64),
        ("idx_offset", _u64),
        ("key0", _u32),
        ("key1", _u32),
        ("args", _p * GJB_MAX_ARGS),
        ("scalars", C.c_float * GJB_MAX_ARGS),
        ("site_in", _p * GJB_MAX_SITES),
        ("site_flags", _u32 * GJB_MAX_SITES),
        ("state", _p),
        ("logp", _p),
        ("accept_count", _p),
        ("alpha_out", _p),
        ("state_width", _i32),
        ("flags", _u32),
        ("n_steps", _i32),
        ("step0", _i32),
        ("step_size", C.c_float),
        ("n_leapfrog", _i32),
        ("compat_stale_grad", _i32),
    ]


CORE_PROTOTYPES = {
    "gjb_abi_version": (C.c_int, []),
    "gjb_resample_workspace_bytes": (_i64, [_i64]),
    "gjb_wmax_reset": (C.c_int, [_p, _p]),
    "gjb_weight_max": (C.c_int, [_p, _i64, _p, _p]),
    "gjb_weight_mass": (C.c_int, [_p, _i64, _p, _p, _p, _p]),
    "gjb_lse_finalize": (C.c_int, [_p, _i64, _p, _p, _i64, _p, _p]),
    "gjb_resample_systematic": (C.c_int, [C.POINTER(ResampleArgs), _p]),
    "gjb_mass_resample_fits": (C.c_int, [_i64]),
    "gjb_mass_resample_systematic": (C.c_int, [C.POINTER(ResampleArgs), _p]),
    "gjb_resample_multinomial": (C.c_int, [_p, _i64, _p, _p, _p, _u32, _u32, _u64, _i64, _p, _p]),
    "gjb_resample_multinomial_keydev": (C.c_int, [_p, _i64, _p, _p, _p, _p, _u64, _i64, _p, _p]),
    "gjb_accept_mask": (C.c_int, [_p, _p, _i64, _p, _p]),
    "gjb_select_rows": (C.c_int, [_p, _p, _p, _p, _i64, _i32, _i32, _i32, _p]),
    "gjb_weight_ess": (C.c_int, [_p, _i64, _p, _p, _p]),
    "gjb_gather_rows": (C.c_int, [_p, _p, _p, _i64, _i32, _p]),
    "gjb_exchange": (C.c_int, [C.POINTER(XchgArgs), _p]),
    "gjb_peers_set_divisor": (C.c_int, [C.POINTER(Peers)]),
    "gjb_epoch_bump": (C.c_int, [_p, _p]),
    "gjb_weight_mass_linked": (C.c_int, [_p, _i64, _p, _p, _u64, _u64, _p]),
    "gjb_weight_mass_prefix_linked": (C.c_int, [_p, _i64, _p, _p, _p, _u64, _u64, _p]),
    "gjb_resample_systematic_pull": (C.c_int, [C.POINTER(ResampleArgs), C.POINTER(Peers), C.POINTER(Peers), _p, _u64, _u64, _p]),
    "gjb_resample_systematic_linked": (C.c_int, [C.POINTER(ResampleArgs), C.POINTER(Peers), _p, _u64, _u64, _u64, _p]),
    "gjb_resample_systematic_peers": (C.c_int, [C.POINTER(ResampleArgs), C.POINTER(Peers), _p]),
    "gjb_gather_rows_peers": (C.c_int, [C.POINTER(Peers), _p, _p, _i64, _i32, _p]),
    "gjb_te_masses": (C.c_int, [_p, _i64, _p, _p, _p]),
    "gjb_te_resample": (C.c_int, [C.POINTER(TeResampleArgs), _p]),
    "gjb_te_table": (C.c_int, [C.POINTER(TeTableArgs), _p]),
    "gjb_pf_key_table": (C.c_int, [_u32, _u32, _i32, _p, _p]),
    "gjb_philox_fill": (C.c_int, [_u32, _u32, _u64, _u32, _u32, _i64, _p, _p]),
    "gjb_normal_fill": (C.c_int, [_u32, _u32, _u64, _u32, _i64, _i32, _p, _p]),
}

MODEL_PROTOTYPES = {
    "gjb_model_info": (C.c_char_p, []),
    "gjb_model_launch": (C.c_int, [C.POINTER(ModelArgs), _p]),
    "gjb_model_pf_grid": (C.c_int, [_i64]),
    "gjb_model_pf_run": (C.c_int, [C.POINTER(PfArgs), _p]),
    "gjb_model_pf_step": (C.c_int, [C.POINTER(StepArgs), _p]),
    "gjb_model_pf_steps_fits": (C.c_int, [_i64]),
    "gjb_model_pf_steps": (C.c_int, [C.POINTER(StepsArgs), _p]),
    "gjb_model_mh_chain": (C.c_int, [C.POINTER(ChainArgs), _p]),
    "gjb_model_hmc_chain": (C.c_int, [C.POINTER(ChainArgs), _p]),
}

_core = None


def _bind(lib, protos):
    for name, (res, argtypes) in protos.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = argtypes
    return lib


def core_library_path() -> Path:
    return build.LIB / "libgjb_core.so"


def core():
    """The loaded core library (built on demand); raises if unavailable."""
    global _core
    if _core is None:
        path = build.build_core()
        _core = _bind(C.CDLL(str(path)), CORE_PROTOTYPES)
        if _core.gjb_abi_version() != ABI_VERSION:
            raise GjbError(f"libgjb_core.so reports ABI {_core.gjb_abi_version()}, this binding is for {ABI_VERSION}")
    return _core


def load_model_library(path: Path):
    return _bind(C.CDLL(str(path)), MODEL_PROTOTYPES)


def check(code: int, what: str) -> None:
    if code == 0:
        return
    if code < 0:
        names = {-1: "GJB_E_ARG", -2: "GJB_E_RANGE", -3: "GJB_E_MODE"}
        raise GjbError(f"{what}: {names.get(code, code)}")
    raise GjbError(f"{what}: CUDA error {code}")


def ptr(t: torch.Tensor | None) -> int | None:
    """Device pointer of a CUDA tensor (None stays null)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise GjbError("genjax_b200 kernels need CUDA tensors (there is no CPU fallback)")
    if not t.is_contiguous():
        raise GjbError("tensor must be contiguous")
    return t.data_ptr()


def stream_ptr(device=None) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def require_cuda() -> torch.device:
    if not torch.cuda.is_available():
        raise GjbError(
            "genjax_b200 needs a CUDA device: the hot path is hand-written sm_100a CUDA and has no CPU fallback"
        )
    return torch.device("cuda", torch.cuda.current_device())

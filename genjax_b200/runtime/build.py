"""nvcc driver: compiles the core library and per-model translation units
IN-TREE (``genjax_b200/_lib/*.so``) for sm_100a.

Built files are git-ignored but travel to the GPU box with the snapshot; a
model that was not pre-built is compiled on first use (nvcc is in the image).
"""

from __future__ import annotations

import hashlib
import os
import subprocess
import threading
from pathlib import Path

PKG = Path(__file__).resolve().parent.parent
CSRC = PKG / "csrc"
LIB = PKG / "_lib"
INCLUDE = PKG.parent / "include"

NVCC_FLAGS = [
    "-gencode",
    "arch=compute_100a,code=sm_100a",
    "-O3",
    "-std=c++17",
    "-lineinfo",
    # no implicit mul+add contraction: every fp32 operation of the generated model bodies and of the distribution library is
    # then ONE IEEE operation, exactly what the NumPy oracle executes (explicit FMAs are written as __fmaf_rn and restated
    # with oracle/rng.py fma32), so sampled values are bit-exact against the oracle instead of within a tolerance
    "-fmad=false",
    "--shared",
    "-Xcompiler",
    "-fPIC",
    "-Xptxas",
    "-v",
]

_lock = threading.Lock()
_model_locks: dict[str, threading.Lock] = {}


class BuildError(RuntimeError):
    pass


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise BuildError("nvcc not found")


def _header_digest(extra: bool = True) -> str:
    h = hashlib.sha256()
    for p in sorted(list(CSRC.glob("*.cuh")) + [INCLUDE / "genjax_b200.h"]):
        h.update(p.name.encode())
        h.update(p.read_bytes())
    h.update(" ".join(NVCC_FLAGS + (_extra_flags() if extra else [])).encode())
    return h.hexdigest()


def _extra_flags() -> list[str]:
    """Extra nvcc flags from $GJB_NVCC_EXTRA (diagnostic builds, e.g. -DGJB_TRACE); part of every digest."""
    return os.environ.get("GJB_NVCC_EXTRA", "").split()


def _compile(src: Path, out: Path, log: Path, extra: bool = True) -> None:
    cmd = [_nvcc(), *NVCC_FLAGS, *(_extra_flags() if extra else []), f"-I{CSRC}", f"-I{INCLUDE}", "-o", str(out), str(src)]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    log.write_text(" ".join(cmd) + "\n" + proc.stdout + proc.stderr)
    if proc.returncode != 0:
        raise BuildError(f"nvcc failed for {src}:\n{proc.stderr[-4000:]}")


def build_core(force: bool = False) -> Path:
    """Compile csrc/gjb_core.cu -> _lib/libgjb_core.so (skipped when up to date)."""
    with _lock:
        LIB.mkdir(exist_ok=True)
        src = CSRC / "gjb_core.cu"
        out = LIB / "libgjb_core.so"
        stamp = LIB / "libgjb_core.hash"
        # (the diagnostic flags of $GJB_NVCC_EXTRA are for the generated model kernels: the core library keeps one build,
        # so a traced multi-process run does not recompile it in every rank)
        digest = hashlib.sha256(src.read_bytes() + _header_digest(extra=False).encode()).hexdigest()
        if not force and out.exists() and stamp.exists() and stamp.read_text() == digest:
            return out
        _compile(src, out, LIB / "libgjb_core.log", extra=False)
        stamp.write_text(digest)
        return out


def model_digest(source: str) -> str:
    return hashlib.sha256((source + _header_digest()).encode()).hexdigest()[:20]


def build_model(source: str, force: bool = False) -> Path:
    """Compile one generated model translation unit -> _lib/model_<digest>.so."""
    digest = model_digest(source)
    out = LIB / f"model_{digest}.so"
    with _lock:
        lock = _model_locks.setdefault(digest, threading.Lock())
    with lock:  # per model: distinct models compile concurrently
        if out.exists() and not force:
            return out
        LIB.mkdir(exist_ok=True)
        src = LIB / f"model_{digest}.cu"
        src.write_text(source)
        tmp = LIB / f"model_{digest}.tmp{os.getpid()}.so"
        _compile(src, tmp, LIB / f"model_{digest}.log")
        os.replace(tmp, out)
        return out

"""ctypes loader of the C / OpenMP restatement of the oracle's bootstrap filter (oracle/c/pf_port.c).
Test infrastructure and CPU baseline of bench.py only -- see oracle/__init__.py for the rules."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "liboracle_pf.so")
FAST_LIB = os.path.join(HERE, "_build", "liboracle_pf_fast.so")
_lib = None
_fast = None


def build(force: bool = False) -> str | None:
    """Compile oracle/c/pf_port.c with gcc (OpenMP when available); returns the library path or None."""
    src = os.path.join(HERE, "c", "pf_port.c")
    fsrc = os.path.join(HERE, "c", "pf_port_fast.c")
    if (not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(src)
            and os.path.exists(FAST_LIB) and os.path.getmtime(FAST_LIB) >= os.path.getmtime(fsrc)):
        return LIB
    try:
        subprocess.run(["make", "-C", os.path.join(HERE, "c"), "-B"], check=True, capture_output=True)
    except Exception:
        return None
    return LIB if os.path.exists(LIB) else None


def lib():
    global _lib
    if _lib is None:
        path = build()
        if path is None:
            return None
        _lib = C.CDLL(path)
        _lib.pf_lgssm.restype = C.c_int
        _lib.pf_lgssm.argtypes = [C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_float,
                                  C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.pf_port_threads.restype = C.c_int
    return _lib


_SIG = [C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_float,
        C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]


def fast_lib():
    """The performance build (oracle/c/pf_port_fast.c: quad-wise Philox, float libm, -O3 -march=native), or None."""
    global _fast
    if _fast is None:
        if build() is None or not os.path.exists(FAST_LIB):
            return None
        _fast = C.CDLL(FAST_LIB)
        _fast.pf_lgssm_fast.restype = C.c_int
        _fast.pf_lgssm_fast.argtypes = _SIG
        _fast.pf_fast_threads.restype = C.c_int
    return _fast


def threads() -> int:
    lb = lib()
    return int(lb.pf_port_threads()) if lb is not None else 0


def set_threads(n: int) -> None:
    lb = lib()
    if lb is not None:
        lb.pf_port_set_threads(int(n))


def pf_lgssm(x0, ys, a, q, c, r, key_table, fast: bool = False, threads_: int | None = None):
    """Runs the filter; returns dict(state, logz_inc, logw_last, ancestors_last).  x0: [n] or [n, d] float32;
    ys: [T] or [T, d]; q, r: [d]; key_table: uint32 [T, 8] (genjax_b200.core.key.pf_key_table).
    ``fast=True``: the performance build (same algorithm and streams; proposals agree to float32 rounding)."""
    lb = fast_lib() if fast else lib()
    if lb is None:
        raise RuntimeError("oracle C port is not built (gcc missing?)")
    if threads_:
        (lb.pf_fast_set_threads if fast else lb.pf_port_set_threads)(int(threads_))
    x = np.array(x0, dtype=np.float32, copy=True, order="C")
    n = x.shape[0]
    d = 1 if x.ndim == 1 else x.shape[1]
    ys = np.ascontiguousarray(ys, dtype=np.float32)
    T = ys.shape[0]
    q = np.ascontiguousarray(np.broadcast_to(np.asarray(q, dtype=np.float32), (d,)))
    r = np.ascontiguousarray(np.broadcast_to(np.asarray(r, dtype=np.float32), (d,)))
    keys = np.ascontiguousarray(key_table, dtype=np.uint32)
    inc = np.empty(T, dtype=np.float64)
    lw = np.empty(n, dtype=np.float32)
    anc = np.empty(n, dtype=np.int32)
    rc = (lb.pf_lgssm_fast if fast else lb.pf_lgssm)(n, T, d, x.ctypes.data, ys.ctypes.data, float(a), q.ctypes.data, float(c), r.ctypes.data,
                     keys.ctypes.data, inc.ctypes.data, lw.ctypes.data, anc.ctypes.data)
    if rc != 0:
        raise MemoryError("pf_lgssm failed")
    return dict(state=x, logz_inc=inc, logw_last=lw, ancestors_last=anc)

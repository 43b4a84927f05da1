"""TEST INFRASTRUCTURE ONLY -- the GENERATED CUDA source of a model, compiled for the host and executed thread by thread.

`codegen.generate(ir, ...)` output is cut before its `extern "C"` launch section (the `<<< >>>` syntax is nvcc-only),
compiled by g++ against tests/host_shim/cuda_runtime.h (intrinsics restated as plain IEEE operations, barriers and
shuffles as no-ops) and driven by a small appended function that calls the kernel body once per thread of ONE block
of one CTA-thread each.  This runs the very per-particle code the GPU runs -- site order, flag handling, hoisted
constants, quad RNG streams, 128-bit load/store guards, MH / HMC loops with the generated log-density, gradient and
proposal functions -- on a CPU, for the quad-mapped (scalar-site) kernels and the chain kernels.  What it cannot run:
lane-group kernels (vector sites: warp shuffles), the cooperative filter kernel and everything in libgjb_core (block
scans); those stay with the interpreter in tests/abi_emulator.py.  It says nothing about the device's own arithmetic
(libm vs CUDA math, FMA contraction): that is what the -m gpu tests are for."""

from __future__ import annotations

import ctypes as C
import hashlib
import os
import subprocess
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
_CACHE: dict = {}


def _deps_digest() -> str:
    """Headers the host build includes: a cached library is stale when any of them changed."""
    h = hashlib.sha256()
    import glob

    for f in sorted(glob.glob(os.path.join(ROOT, "genjax_b200", "csrc", "*.cuh")) + glob.glob(os.path.join(HERE, "host_shim", "*.h"))
                    + [os.path.join(ROOT, "include", "genjax_b200.h")]):
        h.update(open(f, "rb").read())
    return h.hexdigest() + os.environ.get("GJB_SIMT_FLAGS", "")

_DRIVER = r'''
extern "C" int host_model_launch(const gjb_model_args* a) {
  gridDim.x = 1; gridDim.y = 1; gridDim.z = 1; blockDim.x = 1; blockIdx.x = 0;
  for (unsigned t = 0; t < (unsigned)kThreads; ++t) { threadIdx.x = t; model_kernel(*a); }
  return 0;
}
'''
_MASS_DRIVER = r'''
extern "C" int host_model_launch_mass(const gjb_model_args* a) {
  gridDim.x = 1; gridDim.y = 1; gridDim.z = 1; blockDim.x = 1; blockIdx.x = 0;
  for (unsigned t = 0; t < (unsigned)kThreads; ++t) { threadIdx.x = t; model_kernel_static_mass(*a); }
  return 0;
}
'''
_CHAIN_DRIVER = r'''
extern "C" int host_%(kind)s_chain(const gjb_chain_args* a) {
  gridDim.x = 1; gridDim.y = 1; gridDim.z = 1; blockDim.x = 1; blockIdx.x = 0; threadIdx.x = 0;
  %(kind)s_chain_kernel(*a);  // a grid-stride loop over the chains: one "thread" walks all of them
  return 0;
}
'''


def is_host_runnable(source: str) -> bool:
    return "(quad mapping)" in source.split("\n", 1)[0]


def build(source: str):
    """ctypes library of the host build of one generated model source (cached by content)."""
    digest = hashlib.sha256((source + _deps_digest()).encode()).hexdigest()[:20]
    lib = _CACHE.get(digest)
    if lib is not None:
        return lib
    cut = source.index('extern "C" {')
    body = source[:cut].replace("extern __shared__ __align__(16) unsigned char dyn_smem[];", "static unsigned char dyn_smem[1 << 16];")
    text = body + _DRIVER
    if "model_kernel_static_mass(" in body:
        text += _MASS_DRIVER
    for kind in ("mh", "hmc"):
        if f"{kind}_chain_kernel(" in body:
            text += _CHAIN_DRIVER % {"kind": kind}
    d = os.path.join(tempfile.gettempdir(), "gjb_host_kernels")
    os.makedirs(d, exist_ok=True)
    cpp, so = os.path.join(d, f"model_{digest}.cpp"), os.path.join(d, f"model_{digest}.so")
    if not os.path.exists(so):
        with open(cpp, "w") as f:
            f.write(text)
        cmd = ["g++", "-std=c++17", "-O1", "-ffp-contract=off", "-fPIC", "-shared", "-w", f"-I{HERE}/host_shim",
               f"-I{ROOT}/genjax_b200/csrc", f"-I{ROOT}/include", "-o", so + ".tmp", cpp]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("host build of a generated kernel failed:\n" + r.stderr[-4000:])
        os.replace(so + ".tmp", so)
    lib = C.CDLL(so)
    _CACHE[digest] = lib
    return lib

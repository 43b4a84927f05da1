"""TEST INFRASTRUCTURE ONLY -- block-level CUDA kernels executed on the CPU with real block semantics.

The kernels of genjax_b200/csrc/gjb_core.cu (and generated model sources) are compiled by g++ against
tests/host_shim_simt/cuda_runtime.h: every CUDA thread of a block is an OS thread, __syncthreads is a barrier, warp
shuffles exchange values between the lanes of a warp, atomics are atomic, blocks run one after the other.  The
`<<< >>>` launch section of the source is cut off and replaced by small drivers (simt::launch).  This runs tile scans,
the max-scan write-out, block reductions and lane-group kernels AS WRITTEN; it says nothing about the device's
arithmetic, memory model or timing, and cooperative kernels only run as grids of one block."""

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

    for f in sorted(glob.glob(os.path.join(ROOT, "genjax_b200", "csrc", "*.cuh")) + glob.glob(os.path.join(HERE, "host_shim_simt", "*.h"))
                    + [os.path.join(ROOT, "include", "genjax_b200.h")]):
        h.update(open(f, "rb").read())
    return h.hexdigest() + os.environ.get("GJB_SIMT_FLAGS", "")

CORE_DRIVERS = r'''
extern "C" {
int s_weight_max(const float* logw, int64_t n, uint32_t* wmax, int grid) {
  simt::launch(grid, 256, [=] { gjb::weight_max_kernel(logw, n, wmax); });
  return 0;
}
int s_weight_mass(const float* logw, int64_t n, const uint32_t* wmax, const float* m_global, uint64_t* tile_mass) {
  const int tiles = (int)((n + gjb::kTile - 1) / gjb::kTile);
  simt::launch(tiles, gjb::kThreads, [=] { gjb::weight_mass_kernel(logw, n, wmax, m_global, tile_mass); });
  return 0;
}
int s_resample_systematic(const gjb_resample_args* a) {
  const int tiles = (int)((a->n + gjb::kTile - 1) / gjb::kTile);
  gjb_peers none; memset(&none, 0, sizeof(none)); none.world = 1;
  const gjb_resample_args R = *a;
  simt::launch(tiles, gjb::kThreads, [=] { gjb::resample_systematic_kernel(R, none, nullptr, 0, 0, 0); });
  return 0;
}
int s_mass_resample_one_block(const gjb_resample_args* a) {  // cooperative kernel: single-tile grids only
  if (a->n > gjb::kTile) return -1;
  const gjb_resample_args R = *a;
  simt::launch(1, gjb::kThreads, [=] { gjb::mass_resample_kernel(R); });
  return 0;
}
int s_lse_finalize(const uint64_t* tile_mass, int n_tiles, const uint32_t* wmax, const float* m_global, int64_t n_total, double* out) {
  simt::launch(1, 256, [=] { gjb::lse_finalize_kernel(tile_mass, n_tiles, wmax, m_global, n_total, out); });
  return 0;
}
int s_multinomial(const float* logw, int64_t n, const uint32_t* wmax, const uint64_t* tile_mass, uint64_t* cdf, uint32_t k0,
                  uint32_t k1, uint64_t idx_offset, int64_t n_out, int32_t* anc) {
  const int tiles = (int)((n + gjb::kTile - 1) / gjb::kTile);
  simt::launch(tiles, gjb::kThreads, [=] { gjb::cdf_kernel(logw, n, wmax, tile_mass, cdf); });
  simt::launch((int)((n_out + 255) / 256), 256, [=] { gjb::multinomial_search_kernel(cdf, n, k0, k1, idx_offset, n_out, anc); });
  return 0;
}
int s_philox_fill(uint32_t k0, uint32_t k1, uint64_t off, uint32_t site, uint32_t chunk, int64_t n, uint4* out) {
  simt::launch((int)((n + 255) / 256 < 4 ? (n + 255) / 256 : 4), 256, [=] { gjb::philox_fill_kernel(k0, k1, off, site, chunk, n, out); });
  return 0;
}
int s_normal_fill(uint32_t k0, uint32_t k1, uint64_t off, uint32_t site, int64_t n, int d, float* out) {
  simt::launch(4, 256, [=] { gjb::normal_fill_kernel(k0, k1, off, site, n, d, out); });
  return 0;
}
int s_te_masses(const float* logw, int64_t n, uint64_t* cdf, gjb_tile_rec* recs) {
  const int tiles = (int)((n + gjb::kTeTile - 1) / gjb::kTeTile);
  simt::launch(tiles, gjb::kThreads, [=] { gjb::te_mass_kernel(logw, n, cdf, recs); });
  return 0;
}
int s_te_resample(const gjb_te_resample_args* a) {
  const gjb_te_resample_args A = *a;
  const int ctas = (int)((A.out_n + gjb::kTeTile - 1) / gjb::kTeTile);
  simt::launch(ctas, gjb::kThreads, [=] { gjb::te_resample_kernel(A); });
  return 0;
}
int s_te_table(const gjb_te_table_args* a) {
  const gjb_te_table_args A = *a;
  simt::launch(1, gjb::kTabThreads, [=] { gjb::te_table_kernel(A); });
  return 0;
}
int s_pf_key_table(uint32_t k0, uint32_t k1, int T, uint32_t* out) {
  simt::launch(2, 128, [=] { gjb::pf_key_table_kernel(k0, k1, T, out); });
  return 0;
}
int s_gather_rows(const uint32_t* src, const int32_t* anc, uint32_t* dst, int64_t n_out, int w, int grid) {
  simt::launch(grid, 256, [=] { gjb::gather_rows_kernel<uint32_t>(src, anc, dst, n_out, w); });
  return 0;
}
}
'''


def _compile(text: str, tag: str):
    digest = hashlib.sha256((text + _deps_digest()).encode()).hexdigest()[:20]
    lib = _CACHE.get(digest)
    if lib is not None:
        return lib
    d = os.path.join(tempfile.gettempdir(), "gjb_simt_kernels")
    os.makedirs(d, exist_ok=True)
    cpp, so = os.path.join(d, f"{tag}_{digest}.cpp"), os.path.join(d, f"{tag}_{digest}.so")
    if not os.path.exists(so):
        with open(cpp, "w") as f:
            f.write(text)
        cmd = ["g++", "-std=c++17", "-O1", "-ffp-contract=off", "-fPIC", "-shared", "-w", "-pthread", *os.environ.get("GJB_SIMT_FLAGS", "").split(),
               f"-I{HERE}/host_shim_simt", f"-I{ROOT}/genjax_b200/csrc", f"-I{ROOT}/include", "-o", so + ".tmp", cpp]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("SIMT host build failed:\n" + r.stderr[-4000:])
        os.replace(so + ".tmp", so)
    lib = C.CDLL(so)
    _CACHE[digest] = lib
    return lib


def core():
    """libgjb_core's kernels (everything above its `extern "C"` launch section) with SIMT drivers."""
    src = open(os.path.join(ROOT, "genjax_b200", "csrc", "gjb_core.cu")).read()
    return _compile(src[: src.index('extern "C" {')] + CORE_DRIVERS, "core")


MODEL_DRIVER = r'''
extern "C" int s_model_launch(const gjb_model_args* a, int grid) {
  const gjb_model_args A = *a;
  simt::launch(grid, kThreads, [=] { model_kernel(A); });
  return 0;
}
'''
PULL_DRIVER = r'''
extern "C" int s_model_launch_pull(const gjb_model_args* a) {
  const gjb_model_args A = *a;
  const int tiles = (int)((A.n + gjb::kTile - 1) / gjb::kTile);
  simt::launch(tiles, kThreads, [=] { model_kernel_static_pull(A); });
  return 0;
}
'''
STEP_DRIVER = r'''
extern "C" int s_pf_step(const gjb_step_args* a) {
  const gjb_step_args A = *a;
  const int tiles = (int)((A.n + gjb::kTeTile - 1) / gjb::kTeTile);
  if (A.link) simt::launch(tiles, kThreads, [=] { pf_step_kernel_t<true>(A); });
  else simt::launch(tiles, kThreads, [=] { pf_step_kernel_t<false>(A); });
  return 0;
}
'''
STEPS_DRIVER = r'''
extern "C" int s_pf_steps(const gjb_steps_args* q) {  // cooperative: a grid of ONE block only (n <= 2048)
  const gjb_steps_args Q = *q;
  if (Q.n > gjb::kTeTile) return -2;
  simt::launch(1, kThreads, [=] { pf_steps_kernel(Q); });
  return 0;
}
'''
PF_DRIVER = r'''
extern "C" int s_pf_run(const gjb_pf_args* q) {  // the persistent cooperative filter as a grid of ONE block
  const gjb_pf_args Q = *q;
  simt::launch(1, 32, [=] { pf_init_kernel(Q.wmax, Q.barrier); });
  simt::launch(1, kThreads, [=] { pf_kernel(Q); });
  return 0;
}
'''
CHAIN_DRIVER = r'''
extern "C" int s_%(kind)s_chain(const gjb_chain_args* a) {
  const gjb_chain_args A = *a;
  simt::launch(1, 128, [=] { %(kind)s_chain_kernel(A); });
  return 0;
}
'''


def model(source: str):
    """A generated model source (quad- or lane-group-mapped) with a SIMT driver for its generic model_kernel."""
    body = source[: source.index('extern "C" {')].replace(
        "extern __shared__ __align__(16) unsigned char dyn_smem[];", "static unsigned char dyn_smem[1 << 16];")
    text = body + MODEL_DRIVER
    if "model_kernel_static_pull(" in body:
        text += PULL_DRIVER
    if "pf_kernel(" in body:
        text += PF_DRIVER
    if "pf_step_kernel_t(" in body:
        text += STEP_DRIVER
    if "pf_steps_kernel(" in body:
        text += STEPS_DRIVER
    for kind in ("mh", "hmc"):
        if f"{kind}_chain_kernel(" in body:
            text += CHAIN_DRIVER % {"kind": kind}
    return _compile(text, "model")

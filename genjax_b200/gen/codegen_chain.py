"""Batched MCMC kernels for a captured model: ``gjb_model_mh_chain`` and
``gjb_model_hmc_chain`` (include/genjax_b200.h).

One thread owns one chain; the chain state (the selected latent sites, Dtot
floats), its log-density and the gradient live in registers for ``n_steps``
transitions per launch -- only the state row is read and written in HBM.

Reference semantics reproduced per transition:
  * MH  : ``Rejuvenate(proposal, mapping).edit`` on every selected address --
          ``proposed ~ proposal(*mapping(cur))``, ``w = logp(new) - logp(old)
          + bwd - fwd`` (inference/requests/rejuvenate.py:70-94) -- followed by
          the user-side accept ``log(uniform) < w`` and ``where(check, new,
          old)`` (tests/inference/test_requests.py:136-137, 190-191);
  * HMC : ``HMC(selection, eps, L).edit`` -- momenta ~ N(0,1), L leapfrog
          steps, ``alpha = logp_L - logp_0 + logN(-p_L) - logN(p_0)``
          (inference/requests/hmc.py:156-211) -- then the same accept.
          ``compat_stale_grad=1`` reproduces hmc.py:186, which carries the
          gradient of the INITIAL position into the first half-kick of every
          leapfrog step; 0 is the textbook integrator.

The three device functions the hand-written loops call are generated from the
model IR with gen/autodiff.py: ``logp_fn`` (total log-density at a state),
``logp_grad_fn`` (value and gradient) and ``prop_fn`` (proposal arguments at a
state).
"""

from __future__ import annotations

from dataclasses import dataclass

from . import autodiff as AD
from . import expr as E
from .capture import ModelIR
from .expr import Expr, F32, I32


@dataclass(frozen=True)
class ChainSpec:
    """Which sites form the chain state and how MH proposes them.

    latent   : site indices (program order) that make up the state
    proposals: per latent site, ``None`` (random walk: normal(cur, step_size))
               or a callable ``mapping(cur_expr) -> (loc, scale)`` (Rejuvenate's
               ``argument_mapping`` applied to the current value; the proposal
               family is normal / mv_normal_diag)."""

    latent: tuple
    proposals: tuple = ()

    def key(self):
        return ("chain", self.latent, tuple(None if p is None else id(p) for p in self.proposals))


def _layout(ir: ModelIR, spec: ChainSpec):
    """offset and width of every latent site inside the state row."""
    offs, off = {}, 0
    for j in spec.latent:
        s = ir.sites[j]
        if s.value.dtype != F32:
            raise AD.NotDifferentiable(f"chain state must be float-valued; site {s.addr} is integer")
        w = s.value.shape[0] if s.value.ndim else 1
        offs[j] = (off, w)
        off += w
    return offs, off


class _Fn:
    """One generated device function: an emitter whose site / argument nodes are
    bound to the chain context."""

    def __init__(self, ir: ModelIR, spec: ChainSpec, tag: str, qname: str = "q"):
        from .codegen import _Emitter

        self.ir, self.spec, self.tag = ir, spec, tag
        self.em = _Emitter(ir, group=False, const_prefix=f"cv_{tag}")
        self.offs, self.dtot = _layout(ir, spec)
        em = self.em
        em.hoist_vectors = False  # one thread per chain: keep per-thread constants scalar
        for s in ir.sites:
            j = s.index
            if j in self.offs:
                off, w = self.offs[j]
                em.names[s.value._id] = f"({qname} + {off})" if s.value.ndim else f"{qname}[{off}]"
            else:
                em.names[s.value._id] = f"X.o{j}"
        for i, (a, ae) in enumerate(zip(ir.args, ir.arg_exprs)):
            if a.kind == "particle":
                em.names[ae._id] = f"X.a{i}"

    def bind(self, e: Expr, name: str):
        self.em.names[e._id] = name

    def uni_struct(self) -> list[str]:
        t = self.tag
        out = [f"struct Uni_{t} {{", "  float sc[NA];"]
        out.extend(self.em.uni_decl)
        out.append("};")
        out.append(f"__device__ __forceinline__ void make_uni_{t}(Uni_{t}& U, const float* __restrict__ scalars) {{")
        out.append("  for (int i = 0; i < NA; ++i) U.sc[i] = scalars[i];")
        out.extend(self.em.uni_init)
        out.append("}")
        return out


def _ctx_struct(ir: ModelIR, spec: ChainSpec) -> list[str]:
    out = ["struct Ctx {  // per-chain constants: per-chain arguments and the values of the unselected sites"]
    for i, a in enumerate(ir.args):
        if a.kind == "particle":
            ct = "int" if a.dtype == I32 else "float"
            out.append(f"  {ct} a{i}{'[%d]' % a.shape[0] if a.shape else ''};")
    for s in ir.sites:
        if s.index not in spec.latent:
            ct = "int" if s.value.dtype == I32 else "float"
            out.append(f"  {ct} o{s.index}{'[%d]' % s.value.shape[0] if s.value.ndim else ''};")
    out.append("  int unused_;")
    out.append("};")
    return out


def _load_ctx(ir: ModelIR, spec: ChainSpec) -> list[str]:
    out = ["  Ctx X; X.unused_ = 0;"]
    for i, a in enumerate(ir.args):
        if a.kind != "particle":
            continue
        cast = "const int*" if a.dtype == I32 else "const float*"
        if a.shape == ():
            out.append(f"  X.a{i} = __ldg(reinterpret_cast<{cast}>(A.args[{i}]) + c);")
        else:
            D = a.shape[0]
            out.append(f"  for (int k = 0; k < {D}; ++k) X.a{i}[k] = __ldg(reinterpret_cast<const float*>(A.args[{i}]) + c * {D} + k);")
    for s in ir.sites:
        j = s.index
        if j in spec.latent:
            continue
        cast = "const int*" if s.value.dtype == I32 else "const float*"
        if s.value.ndim == 0:
            out.append(f"  X.o{j} = __ldg(reinterpret_cast<{cast}>(A.site_in[{j}]) + ((A.site_flags[{j}] & GJB_SITE_BCAST) ? 0 : c));")
        else:
            D = s.value.shape[0]
            out.append(f"  for (int k = 0; k < {D}; ++k) X.o{j}[k] = __ldg(reinterpret_cast<const float*>(A.site_in[{j}]) + ((A.site_flags[{j}] & GJB_SITE_BCAST) ? 0 : c * {D}) + k);")
    return out


def _copy_out(fn: _Fn, e: Expr, dst: str, off: int, w: int) -> list[str]:
    """statements storing value ``e`` (scalar or width-w vector) into dst[off:off+w]"""
    em = fn.em
    em.emit_expr(e)
    r = em.ref(e)
    if e.ndim == 0:
        return [f"      for (int k = 0; k < {w}; ++k) {dst}[{off} + k] = {r};"] if w > 1 else [f"      {dst}[{off}] = {r};"]
    return [f"      for (int k = 0; k < {w}; ++k) {dst}[{off} + k] = {r}[k];"]


def generate_chain(ir: ModelIR, spec: ChainSpec) -> tuple[str, str]:
    """(namespace-scope device code, extern "C" launchers) for the chain kernels."""
    offs, dtot = _layout(ir, spec)
    if dtot == 0:
        raise ValueError("no latent sites selected for the chain")
    if dtot > 64:
        raise NotImplementedError("chain state wider than 64 floats")
    logp = AD.model_logp(ir)
    lat_vals = [ir.sites[j].value for j in spec.latent]

    ns: list[str] = []
    ns.append(f"constexpr int kD = {dtot};  // chain state width")
    ns.extend(_ctx_struct(ir, spec))

    # ---- logp_fn
    f1 = _Fn(ir, spec, "lp")
    f1.em.emit_expr(logp)
    body1 = list(f1.em.lines) + [f"      return {f1.em.ref(logp)};"]
    ns.extend(f1.uni_struct())
    ns.append("__device__ __forceinline__ float logp_fn(const Uni_lp& U, const Ctx& X, const float* q) {")
    ns.extend(body1)
    ns.append("}")

    # ---- logp_grad_fn
    have_grad = True
    try:
        grads = AD.grad(logp, lat_vals)
        f2 = _Fn(ir, spec, "lg")
        f2.em.emit_expr(logp)
        stores = []
        for j, g in zip(spec.latent, grads):
            off, w = offs[j]
            stores.extend(_copy_out(f2, g, "g", off, w))
        ns.extend(f2.uni_struct())
        ns.append("__device__ __forceinline__ float logp_grad_fn(const Uni_lg& U, const Ctx& X, const float* q, float* g) {")
        ns.extend(f2.em.lines)
        ns.extend(stores)
        ns.append(f"      return {f2.em.ref(logp)};")
        ns.append("}")
    except AD.NotDifferentiable as err:
        have_grad = False
        ns.append(f"// no gradient kernel: {err}")

    # ---- prop_fn: proposal arguments (loc, scale) at a state
    f3 = _Fn(ir, spec, "pr")
    stores = []
    custom = False
    for n_, j in enumerate(spec.latent):
        off, w = offs[j]
        mapping = spec.proposals[n_] if n_ < len(spec.proposals) else None
        cur = ir.sites[j].value
        if mapping is None:
            loc, scale = cur, Expr("chain_step", (), F32, ())
        else:
            custom = True
            loc, scale = (E.lift(x) for x in mapping(cur))
        for e_ in E.topo([scale, loc]):
            if e_.op == "chain_step":
                f3.bind(e_, "step_size")
        stores.extend(_copy_out(f3, loc, "loc", off, w))
        stores.extend(_copy_out(f3, scale, "scale", off, w))
    ns.extend(f3.uni_struct())
    ns.append("__device__ __forceinline__ void prop_fn(const Uni_pr& U, const Ctx& X, const float* q, float step_size, float* loc, float* scale) {")
    ns.extend(f3.em.lines)
    ns.extend(stores)
    ns.append("}")
    consts = f1.em.consts + (f2.em.consts if have_grad else []) + f3.em.consts
    load = load_ctx_code(ir, spec)
    from .codegen import _shared_decls

    _, stage = _shared_decls(ir)
    stage_code = "\n".join([ln.replace("ARGS", "A.args") for ln in stage] + (["  __syncthreads();"] if stage else []))
    ns_final = consts + ns
    ns_final.append(_MH_KERNEL.replace("CHAIN_LOAD_CTX", load).replace("CHAIN_STAGE", stage_code))
    if have_grad:
        ns_final.append(_HMC_KERNEL.replace("CHAIN_LOAD_CTX", load).replace("CHAIN_STAGE", stage_code))
    ext = _EXTERN_MH
    ext += _EXTERN_HMC if have_grad else _EXTERN_HMC_STUB
    return "\n".join(ns_final), ext


_COMMON = r"""
// one Philox stream per chain: ctr = (chain_lo, chain_hi, chunk, transition + 1)
__device__ __forceinline__ void chain_normals(const gjb::Lane& l, uint32_t t1, float* z) {
#pragma unroll
  for (int c = 0; c < (kD + 3) / 4; ++c) {
    const float4 n4 = gjb::normal4(l, t1, (uint32_t)c);
    const float zz[4] = {n4.x, n4.y, n4.z, n4.w};
#pragma unroll
    for (int s = 0; s < 4; ++s) if (4 * c + s < kD) z[4 * c + s] = zz[s];
  }
}
__device__ __forceinline__ float chain_uniform(const gjb::Lane& l, uint32_t t1) {
  return gjb::u01(l.words(t1, 0xFFFFu).x);
}
__device__ __forceinline__ float std_normal_logpdf_sum(const float* p, float sign) {
  float s = 0.0f;
#pragma unroll
  for (int k = 0; k < kD; ++k) s += gjb::Normal::logpdf(sign * p[k], 0.0f, 1.0f);
  return s;
}
"""

_MH_KERNEL = _COMMON + r"""
__global__ void __launch_bounds__(128) mh_chain_kernel(const __grid_constant__ gjb_chain_args A) {
CHAIN_STAGE
  Uni_lp ULP; make_uni_lp(ULP, A.scalars);
  Uni_pr UPR; make_uni_pr(UPR, A.scalars);
  for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < A.n; c += (int64_t)gridDim.x * blockDim.x) {
    CHAIN_LOAD_CTX
    float q[kD], prop[kD], loc[kD], scale[kD], z[kD];
#pragma unroll
    for (int k = 0; k < kD; ++k) q[k] = A.state[c * kD + k];
    float lp = (A.flags & GJB_CHAIN_HAVE_LOGP) ? A.logp[c] : logp_fn(ULP, X, q);
    int acc = A.accept_count ? A.accept_count[c] : 0;
    float alpha = 0.0f;
    const gjb::Lane lane = gjb::make_lane(A.key0, A.key1, A.idx_offset + (uint64_t)c);
    for (int s = 0; s < A.n_steps; ++s) {
      const uint32_t t1 = (uint32_t)(A.step0 + s) + 1u;
      chain_normals(lane, t1, z);
      prop_fn(UPR, X, q, A.step_size, loc, scale);
      float fwd = 0.0f;
#pragma unroll
      for (int k = 0; k < kD; ++k) { prop[k] = loc[k] + scale[k] * z[k]; fwd += gjb::Normal::logpdf(prop[k], loc[k], scale[k]); }
      // backward proposal arguments: at the PROPOSED state (Metropolis-Hastings), or -- compat -- at the OLD state, which is
      // what rejuvenate.py:84-86 does (`argument_mapping(bwd_chm)` with bwd_chm the discarded, i.e. old, choices)
      if (!A.compat_stale_grad) prop_fn(UPR, X, prop, A.step_size, loc, scale);
      float bwd = 0.0f;
#pragma unroll
      for (int k = 0; k < kD; ++k) bwd += gjb::Normal::logpdf(q[k], loc[k], scale[k]);
      const float lp_new = logp_fn(ULP, X, prop);
      alpha = ((lp_new - lp) + bwd) - fwd;  // rejuvenate.py:88: w + bwd_score - fwd_score
      const bool ok = (A.flags & GJB_CHAIN_NO_ACCEPT) ? true : (logf(chain_uniform(lane, t1)) < alpha);
      if (ok) {
#pragma unroll
        for (int k = 0; k < kD; ++k) q[k] = prop[k];
        lp = lp_new;
        ++acc;
      }
    }
#pragma unroll
    for (int k = 0; k < kD; ++k) A.state[c * kD + k] = q[k];
    A.logp[c] = lp;
    if (A.accept_count) A.accept_count[c] = acc;
    if (A.alpha_out) A.alpha_out[c] = alpha;
  }
}
"""

_HMC_KERNEL = r"""
__global__ void __launch_bounds__(128) hmc_chain_kernel(const __grid_constant__ gjb_chain_args A) {
CHAIN_STAGE
  Uni_lg ULG; make_uni_lg(ULG, A.scalars);
  const float eps = A.step_size;
  for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < A.n; c += (int64_t)gridDim.x * blockDim.x) {
    CHAIN_LOAD_CTX
    float q0[kD], q[kD], g0[kD], gc[kD], g[kD], p[kD];
#pragma unroll
    for (int k = 0; k < kD; ++k) q0[k] = A.state[c * kD + k];
    float lp0 = logp_grad_fn(ULG, X, q0, g0);
    int acc = A.accept_count ? A.accept_count[c] : 0;
    float alpha = 0.0f;
    const gjb::Lane lane = gjb::make_lane(A.key0, A.key1, A.idx_offset + (uint64_t)c);
    for (int s = 0; s < A.n_steps; ++s) {
      const uint32_t t1 = (uint32_t)(A.step0 + s) + 1u;
      chain_normals(lane, t1, p);                           // sample_momenta, hmc.py:120-130
      const float k0 = std_normal_logpdf_sum(p, 1.0f);      // assess_momenta(p_0)
#pragma unroll
      for (int k = 0; k < kD; ++k) { q[k] = q0[k]; gc[k] = g0[k]; g[k] = g0[k]; }  // (g = g0: n_leapfrog == 0 must not carry an uninitialised gradient)
      float lp = lp0;
      for (int l = 0; l < A.n_leapfrog; ++l) {              // kernel, hmc.py:170-186
#pragma unroll
        for (int k = 0; k < kD; ++k) { p[k] = p[k] + (eps * 0.5f) * gc[k]; q[k] = q[k] + eps * p[k]; }
        lp = logp_grad_fn(ULG, X, q, g);
#pragma unroll
        for (int k = 0; k < kD; ++k) {
          p[k] = p[k] + (eps * 0.5f) * g[k];
          if (!A.compat_stale_grad) gc[k] = g[k];           // hmc.py:186 carries the OLD gradient when compat
        }
      }
      const float k1 = std_normal_logpdf_sum(p, -1.0f);     // assess_momenta(-p_L)
      alpha = ((lp - lp0) + k1) - k0;                       // hmc.py:196-203
      const bool ok = (A.flags & GJB_CHAIN_NO_ACCEPT) ? true : (logf(chain_uniform(lane, t1)) < alpha);
      if (ok) {
#pragma unroll
        for (int k = 0; k < kD; ++k) { q0[k] = q[k]; g0[k] = g[k]; }
        lp0 = lp;
        ++acc;
      }
    }
#pragma unroll
    for (int k = 0; k < kD; ++k) A.state[c * kD + k] = q0[k];
    A.logp[c] = lp0;
    if (A.accept_count) A.accept_count[c] = acc;
    if (A.alpha_out) A.alpha_out[c] = alpha;
  }
}
"""

_CHECK = r"""
  if (!a || a->n < 0 || !a->state || !a->logp || a->n_steps < 0) return GJB_E_ARG;
  if (a->state_width != kD) return GJB_E_ARG;
  if (a->n == 0) return 0;
"""

_EXTERN_MH = r"""
int gjb_model_mh_chain(const gjb_chain_args* a, void* stream) {""" + _CHECK + r"""
  // chains keep their state in registers for the whole launch: spread them over ALL SMs before filling one -- 8192 chains
  // (one GPU's share of configs[4]) in 128-thread blocks would occupy 64 of the 148 SMs
  int tpb = 128;
  while (tpb > 32 && (a->n + tpb - 1) / tpb < 2 * 148) tpb >>= 1;
  int64_t blocks = (a->n + tpb - 1) / tpb;
  const int64_t cap = gjb::resident_blocks((const void*)mh_chain_kernel, tpb, 16);
  if (blocks > cap) blocks = cap;
  mh_chain_kernel<<<(int)blocks, tpb, 0, (cudaStream_t)stream>>>(*a);
  return (int)cudaGetLastError();
}
"""

_EXTERN_HMC = r"""
int gjb_model_hmc_chain(const gjb_chain_args* a, void* stream) {""" + _CHECK + r"""
  if (a->n_leapfrog < 0) return GJB_E_ARG;
  // chains keep their state in registers for the whole launch: spread them over ALL SMs before filling one -- 8192 chains
  // (one GPU's share of configs[4]) in 128-thread blocks would occupy 64 of the 148 SMs
  int tpb = 128;
  while (tpb > 32 && (a->n + tpb - 1) / tpb < 2 * 148) tpb >>= 1;
  int64_t blocks = (a->n + tpb - 1) / tpb;
  const int64_t cap = gjb::resident_blocks((const void*)hmc_chain_kernel, tpb, 16);
  if (blocks > cap) blocks = cap;
  hmc_chain_kernel<<<(int)blocks, tpb, 0, (cudaStream_t)stream>>>(*a);
  return (int)cudaGetLastError();
}
"""

_EXTERN_HMC_STUB = r"""
int gjb_model_hmc_chain(const gjb_chain_args* a, void* stream) { (void)a; (void)stream; return GJB_E_MODE; }
"""


def load_ctx_code(ir: ModelIR, spec: ChainSpec) -> str:
    return "\n".join("  " + ln for ln in _load_ctx(ir, spec))

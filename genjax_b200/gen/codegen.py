"""ModelIR -> one CUDA translation unit (sm_100a) exporting the per-model
C-ABI of include/genjax_b200.h (``gjb_model_info`` / ``gjb_model_launch`` /
``gjb_model_mh_chain`` / ``gjb_model_hmc_chain``).

The generated ``model_kernel`` is the fused, batched form of the reference's
static-language handlers (static.py:254-278, 298-321, 341-380, 407-466,
616-673): sites are visited in program order; per-site launch flags choose
sample-vs-read and whether the site's logpdf joins the importance weight, so
ONE kernel serves simulate / assess / generate / update / regenerate.

Two thread mappings:
  * "quad"  (all-scalar models or odd event widths): a thread owns 4
    consecutive particles, 128-bit loads/stores of every [N] array;
  * "group" (event width D, D % 4 == 0, D/4 a power of two <= 32): G = D/4
    lanes own one particle, each lane one float4 of every width-D value;
    logpdf partials are reduced with warp shuffles.
"""

from __future__ import annotations

import json

from . import expr as E
from .capture import ModelIR
from .expr import Expr, F32, I32

_UN_FN = {
    "neg": "gjb::f_neg", "exp": "gjb::f_exp", "log": "gjb::f_log", "sqrt": "gjb::f_sqrt", "abs": "gjb::f_abs",
    "tanh": "gjb::f_tanh", "sigmoid": "gjb::f_sigmoid", "log1p": "gjb::f_log1p", "expm1": "gjb::f_expm1",
    "square": "gjb::f_square", "floor": "gjb::f_floor", "sin": "gjb::f_sin", "cos": "gjb::f_cos",
    "softplus": "gjb::f_softplus", "lgamma": "gjb::f_lgamma", "reciprocal": "gjb::f_reciprocal",
}
_BIN_OP = {"add": "+", "sub": "-", "mul": "*", "div": "/", "lt": "<", "le": "<=", "gt": ">", "ge": ">=",
           "eq": "==", "ne": "!=", "and": "&&", "or": "||"}
_BIN_FN = {"pow": "gjb::f_pow", "min": "gjb::f_min", "max": "gjb::f_max"}


def group_lanes(width: int) -> int:
    """Lanes per particle for the group mapping, 0 if the model must use quads."""
    if width > 0 and width % 4 == 0:
        g = width // 4
        if g <= 32 and (g & (g - 1)) == 0:
            return g
    return 0


class _Emitter:
    """Emits the per-particle body.  ``group`` selects lane-distributed V4
    vectors (True) or per-thread float arrays (False)."""

    def __init__(self, ir: ModelIR, group: bool):
        self.ir = ir
        self.group = group
        self.D = ir.width
        self.lines: list[str] = []
        self.names: dict[int, str] = {}
        self.consts: list[str] = []  # namespace-scope __constant__ arrays

    def w(self, s: str):
        self.lines.append("      " + s)

    # ---- type helpers
    def ctype(self, e: Expr) -> str:
        if e.ndim == 0:
            return "int" if e.dtype == I32 else "float"
        if self.group and e.shape[0] == self.D and e.op not in ("row", "constvec") and not self._is_shared_vec(e):
            return "gjb::V4"
        return "array"

    @staticmethod
    def _is_shared_vec(e: Expr) -> bool:
        return e.op == "arg" and e.attr["kind"] == "shared"

    def is_ptr_vec(self, e: Expr) -> bool:
        """Vector values that live behind a pointer (indexable as name[k])."""
        return e.ndim == 1 and (e.op in ("row", "constvec") or self._is_shared_vec(e) or not self.group
                                or e.shape[0] != self.D)

    def ref(self, e: Expr) -> str:
        return self.names[e._id]

    def elem_ref(self, e: Expr, k: str) -> str:
        """name of element k of a (possibly scalar, broadcast) operand inside a loop."""
        if e.ndim == 0:
            return self.ref(e)
        return f"{self.ref(e)}[{k}]"

    def as_v4(self, e: Expr) -> str:
        """operand as a lane V4 (group mode)."""
        if e.ndim == 0:
            return f"gjb::v4_splat((float){self.ref(e)})"
        if self.ctype(e) == "gjb::V4":
            return self.ref(e)
        # pointer vector of width D: this lane's 4 elements
        return f"gjb::v4_load({self.ref(e)} + 4 * lane)" if self._aligned_ptr(e) else (
            f"gjb::V4{{{{{self.ref(e)}[4*lane], {self.ref(e)}[4*lane+1], {self.ref(e)}[4*lane+2], {self.ref(e)}[4*lane+3]}}}}")

    def _aligned_ptr(self, e: Expr) -> bool:
        return self._is_shared_vec(e)

    # ---- expression emission
    def emit_expr(self, e: Expr):
        if e._id in self.names:
            return
        for i in e.ins:
            self.emit_expr(i)
        n = f"e{len(self.names)}"
        op = e.op
        if op == "const":
            self.names[e._id] = E.fmt_float(e.attr) if e.dtype == F32 else str(int(e.attr))
            return
        if op == "arg":
            self.names[e._id] = self.arg_name(e)
            return
        if op == "site":
            self.names[e._id] = f"s{e.attr}"
            return
        if op == "constvec":
            vals = ", ".join(E.fmt_float(v) if e.dtype == F32 else f"(float){int(v)}" for v in e.attr)
            cname = f"cv{len(self.consts)}"
            self.consts.append(f"__device__ const float {cname}[{len(e.attr)}] = {{{vals}}};")
            self.names[e._id] = cname
            return
        self.names[e._id] = n
        t = self.ctype(e)
        if op == "row":
            mat, idx = e.ins
            K = mat.shape[1]
            self.w(f"const float* {n} = {self.ref(mat)} + (int)({self.ref(idx)}) * {K};")
            return
        if op == "gather1":
            vec, idx = e.ins
            self.w(f"const float {n}_f = {self.ref(vec)}[(int)({self.ref(idx)})];")
            if e.dtype == I32:
                self.w(f"const int {n} = __float_as_int({n}_f);")
            else:
                self.w(f"const float {n} = {n}_f;")
            return
        if op == "elem":
            (vec,) = e.ins
            j = int(e.attr)
            if self.ctype(vec) == "gjb::V4":
                self.w(f"const float {n} = __shfl_sync(0xffffffffu, {self.ref(vec)}.v[{j % 4}], (threadIdx.x & 31 & ~(G - 1)) + {j // 4});")
            else:
                self.w(f"const float {n} = {self.ref(vec)}[{j}];")
            return
        if op == "sum":
            (x,) = e.ins
            if self.ctype(x) == "gjb::V4":
                self.w(f"const float {n} = gjb::group_sum<G>(gjb::v4_hsum({self.ref(x)}));")
            else:
                K = x.shape[0]
                self.w(f"float {n} = 0.0f;")
                self.w(f"for (int k = 0; k < {K}; ++k) {n} += {self.ref(x)}[k];")
            return
        if op == "cast":
            (x,) = e.ins
            if t in ("int", "float"):
                self.w(f"const {t} {n} = ({t}){self.ref(x)};")
                return
            raise NotImplementedError("vector casts")
        # elementwise ----------------------------------------------------
        if t in ("int", "float"):
            self.w(f"const {t} {n} = {self.scalar_rhs(e, None)};")
        elif t == "gjb::V4":
            self.w(f"const gjb::V4 {n} = {self.v4_rhs(e)};")
        else:
            K = e.shape[0]
            self.w(f"float {n}[{K}];")
            self.w(f"for (int k = 0; k < {K}; ++k) {n}[k] = {self.scalar_rhs(e, 'k')};")

    def scalar_rhs(self, e: Expr, k) -> str:
        r = (lambda x: self.elem_ref(x, k)) if k is not None else self.ref
        op = e.op
        if op in _UN_FN:
            return f"{_UN_FN[op]}({r(e.ins[0])})"
        if op == "logical_not":
            return f"(!({r(e.ins[0])}) ? 1 : 0)"
        if op in _BIN_OP:
            a, b = r(e.ins[0]), r(e.ins[1])
            if e.dtype == I32 and op in ("lt", "le", "gt", "ge", "eq", "ne", "and", "or"):
                return f"(({a} {_BIN_OP[op]} {b}) ? 1 : 0)"
            return f"({a} {_BIN_OP[op]} {b})"
        if op in _BIN_FN:
            return f"{_BIN_FN[op]}({r(e.ins[0])}, {r(e.ins[1])})"
        if op == "where":
            return f"(({r(e.ins[0])}) ? ({r(e.ins[1])}) : ({r(e.ins[2])}))"
        raise NotImplementedError(f"op {op}")

    def v4_rhs(self, e: Expr) -> str:
        op = e.op
        if op in _UN_FN:
            return f"{_UN_FN[op]}({self.as_v4(e.ins[0])})"
        if op in _BIN_OP and op in ("add", "sub", "mul", "div"):
            return f"({self.as_v4(e.ins[0])} {_BIN_OP[op]} {self.as_v4(e.ins[1])})"
        if op in _BIN_FN:
            return f"{_BIN_FN[op]}({self.as_v4(e.ins[0])}, {self.as_v4(e.ins[1])})"
        raise NotImplementedError(f"vector op {op} in a lane-group kernel")

    def arg_name(self, e: Expr) -> str:
        i = e.attr["index"]
        kind = e.attr["kind"]
        if kind == "scalar":
            return f"((int)A.scalars[{i}])" if e.dtype == I32 else f"A.scalars[{i}]"
        if kind == "shared":
            if e.ndim == 0:
                return f"(__float_as_int(sh{i}[0]))" if e.dtype == I32 else f"sh{i}[0]"
            return f"sh{i}"
        return f"a{i}"  # particle arg: local loaded before the body

    # ---- sites
    def emit_site(self, s):
        d = s.dist
        j = s.index
        for a in s.args:
            self.emit_expr(a)
        fl = f"fl{j}"
        need = f"(need_score || ({fl} & GJB_SITE_WEIGHT))"
        self.w(f"// site {j} {'/'.join(map(str, s.addr))!r}: {d.name}")
        if not d.vector:
            vt = "int" if s.value.dtype == I32 else "float"
            if d.name == "categorical":
                lg = s.args[0]
                a = [(self.ref(lg), lg.shape[0])]
            else:
                a = [self.ref(x) for x in s.args]
            self.w(f"{vt} s{j};")
            self.w(f"if ({fl} & GJB_SITE_SAMPLE) s{j} = {d.emit_sample(j, a, self)}; else s{j} = in_s{j};")
            self.w(f"if {need} {{ const float lp = {d.emit_logpdf(f's{j}', a, self)}; score += lp; if ({fl} & GJB_SITE_WEIGHT) weight += lp; }}")
            return
        if d.name != "mv_normal_diag":
            raise NotImplementedError(d.name)
        loc, scale = s.args
        D = s.value.shape[0]
        if self.group:
            self.w(f"gjb::V4 s{j};")
            self.w(f"if ({fl} & GJB_SITE_SAMPLE) s{j} = gjb::mvn_diag_sample(rng, {j + 1}u, (uint32_t)lane, {self.as_v4(loc)}, {self.as_v4(scale)}); else s{j} = in_s{j};")
            self.w(f"if {need} {{ const float lp = gjb::mvn_diag_logpdf4(s{j}, {self.as_v4(loc)}, {self.as_v4(scale)}); vscore += lp; if ({fl} & GJB_SITE_WEIGHT) vweight += lp; }}")
        else:
            self.w(f"float s{j}[{D}];")
            self.w(f"if ({fl} & GJB_SITE_SAMPLE) {{")
            self.w(f"  for (int c = 0; c < {(D + 3) // 4}; ++c) {{ const float4 z = gjb::normal4(rng, {j + 1}u, (uint32_t)c); const float zz[4] = {{z.x, z.y, z.z, z.w}};")
            self.w(f"    for (int t = 0; t < 4; ++t) {{ const int k = 4 * c + t; if (k < {D}) s{j}[k] = {self.elem_ref(loc, 'k')} + {self.elem_ref(scale, 'k')} * zz[t]; }} }}")
            self.w(f"}} else {{ for (int k = 0; k < {D}; ++k) s{j}[k] = in_s{j}[k]; }}")
            self.w(f"if {need} {{ float lp = 0.0f; for (int k = 0; k < {D}; ++k) lp += gjb::Normal::logpdf(s{j}[k], {self.elem_ref(loc, 'k')}, {self.elem_ref(scale, 'k')}); score += lp; if ({fl} & GJB_SITE_WEIGHT) weight += lp; }}")


def _info_json(ir: ModelIR, mapping: str, G: int) -> str:
    info = {
        "name": ir.name,
        "digest": ir.digest,
        "mapping": mapping,
        "lanes_per_particle": G,
        "width": ir.width,
        "args": [{"kind": a.kind, "dtype": a.dtype, "shape": list(a.shape)} for a in ir.args],
        "sites": [
            {"addr": list(s.addr), "dist": s.dist.name, "dtype": s.value.dtype, "shape": list(s.value.shape)}
            for s in ir.sites
        ],
        "n_rets": len(ir.ret_leaves),
    }
    return json.dumps(info)


def _shared_decls(ir: ModelIR) -> tuple[list[str], list[str]]:
    decl, stage = [], []
    for i, a in enumerate(ir.args):
        if a.kind == "shared":
            n = 1
            for d in a.shape:
                n *= d
            n = max(n, 1)
            decl.append(f"  __shared__ __align__(16) float sh{i}[{(n + 3) // 4 * 4}];")
            stage.append(f"  gjb::stage_shared(sh{i}, A.args[{i}], {n});")
    return decl, stage


def generate(ir: ModelIR) -> str:
    G = group_lanes(ir.width)
    if G:
        # categorical tables etc. are allowed only as pointer vectors; everything of width D is lane-distributed
        return _generate_group(ir, G)
    return _generate_quad(ir)


# ============================================================== quad mapping


def _generate_quad(ir: ModelIR) -> str:
    em = _Emitter(ir, group=False)
    ns = len(ir.sites)
    out: list[str] = []
    decl, stage = _shared_decls(ir)

    # body ----------------------------------------------------------------
    for s in ir.sites:
        em.emit_site(s)
    ret_names = []
    for r in ir.ret_leaves:
        if isinstance(r, Expr):
            em.emit_expr(r)
            ret_names.append(em.ref(r))
        else:
            ret_names.append(E.fmt_float(float(r)))
    body = "\n".join(em.lines)

    P = ["struct P {"]
    pre: list[str] = []   # loads before the body (per quad)
    bind: list[str] = []  # per-u local bindings
    post: list[str] = []  # stores after the body
    save: list[str] = []  # per-u saves into P
    for i, a in enumerate(ir.args):
        if a.kind != "particle":
            continue
        ct = "int" if a.dtype == I32 else "float"
        if a.shape == ():
            P.append(f"  {ct} a{i};")
            conv = "(int)w[u]" if a.dtype == I32 else "gjb::as_f(w[u])"
            pre.append(f"    gjb::load4(A.args[{i}], i0, nv, g, A.gather != nullptr, false, w);")
            pre.append(f"    for (int u = 0; u < 4; ++u) p[u].a{i} = {conv};")
            bind.append(f"      const {ct} a{i} = p[u].a{i};")
        else:
            D = a.shape[0]
            P.append(f"  float a{i}[{D}];")
            pre.append(f"    for (int u = 0; u < 4; ++u) {{ const int64_t row = A.gather ? (int64_t)g[u] : i0 + u;")
            pre.append(f"      for (int k = 0; k < {D}; ++k) p[u].a{i}[k] = (u < nv) ? __ldg(reinterpret_cast<const float*>(A.args[{i}]) + row * {D} + k) : 0.0f; }}")
            bind.append(f"      const float* a{i} = p[u].a{i};")
    for s in ir.sites:
        j = s.index
        ct = "int" if s.value.dtype == I32 else "float"
        if s.value.ndim == 0:
            P.append(f"  {ct} s{j};")
            conv = "(int)w[u]" if s.value.dtype == I32 else "gjb::as_f(w[u])"
            pre.append(f"    if (!(fl{j} & GJB_SITE_SAMPLE)) {{ gjb::load4(A.site_in[{j}], i0, nv, g0, false, (fl{j} & GJB_SITE_BCAST) != 0, w);")
            pre.append(f"      for (int u = 0; u < 4; ++u) p[u].s{j} = {conv}; }}")
            bind.append(f"      const {ct} in_s{j} = p[u].s{j};")
            save.append(f"      p[u].s{j} = s{j};")
            post.append(f"    if (A.site_out[{j}]) {{ for (int u = 0; u < 4; ++u) w[u] = gjb::as_u(p[u].s{j}); gjb::store4(A.site_out[{j}], i0, nv, w); }}")
        else:
            D = s.value.shape[0]
            P.append(f"  float s{j}[{D}];")
            pre.append(f"    if (!(fl{j} & GJB_SITE_SAMPLE)) {{ for (int u = 0; u < 4; ++u) for (int k = 0; k < {D}; ++k)")
            pre.append(f"      p[u].s{j}[k] = (u < nv) ? __ldg(reinterpret_cast<const float*>(A.site_in[{j}]) + ((fl{j} & GJB_SITE_BCAST) ? 0 : (i0 + u) * {D}) + k) : 0.0f; }}")
            bind.append(f"      const float* in_s{j} = p[u].s{j};")
            save.append(f"      for (int k = 0; k < {D}; ++k) p[u].s{j}[k] = s{j}[k];")
            post.append(f"    if (A.site_out[{j}]) {{ for (int u = 0; u < nv; ++u) for (int k = 0; k < {D}; ++k) reinterpret_cast<float*>(A.site_out[{j}])[(i0 + u) * {D} + k] = p[u].s{j}[k]; }}")
    for k, r in enumerate(ir.ret_leaves):
        is_vec = isinstance(r, Expr) and r.ndim == 1
        is_int = isinstance(r, Expr) and r.dtype == I32
        if is_vec:
            D = r.shape[0]
            P.append(f"  float r{k}[{D}];")
            save.append(f"      for (int k = 0; k < {D}; ++k) p[u].r{k}[k] = {ret_names[k]}[k];")
            post.append(f"    if (A.ret_out[{k}]) {{ for (int u = 0; u < nv; ++u) for (int k = 0; k < {D}; ++k) reinterpret_cast<float*>(A.ret_out[{k}])[(i0 + u) * {D} + k] = p[u].r{k}[k]; }}")
        else:
            P.append(f"  {'int' if is_int else 'float'} r{k};")
            save.append(f"      p[u].r{k} = {ret_names[k]};")
            post.append(f"    if (A.ret_out[{k}]) {{ for (int u = 0; u < 4; ++u) w[u] = gjb::as_u(p[u].r{k}); gjb::store4(A.ret_out[{k}], i0, nv, w); }}")
    P.append("  float score, weight;")
    P.append("};")

    out.append(f"// generated by genjax_b200.gen.codegen -- model '{ir.name}' [{ir.digest}] (quad mapping)")
    out.append('#include "gjb_model.cuh"')
    out.append("namespace {")
    out.extend(em.consts)
    out.extend(P)
    out.append("constexpr int kThreads = 256;")
    out.append("__global__ void __launch_bounds__(kThreads) model_kernel(const __grid_constant__ gjb_model_args A) {")
    out.extend(decl)
    out.extend(stage)
    if stage:
        out.append("  __syncthreads();")
    for j in range(ns):
        out.append(f"  const uint32_t fl{j} = A.site_flags[{j}];")
    out.append("  const bool need_score = A.score_out != nullptr;")
    out.append("  const uint32_t key0 = A.key_dev ? __ldg(A.key_dev) : A.key0, key1 = A.key_dev ? __ldg(A.key_dev + 1) : A.key1;")
    out.append("  float run_max = -INFINITY;")
    out.append("  const int32_t g0[4] = {0, 0, 0, 0};")
    out.append("  const int64_t nq = (A.n + 3) >> 2;")
    out.append("  for (int64_t q = blockIdx.x * (int64_t)kThreads + threadIdx.x; q < nq; q += (int64_t)gridDim.x * kThreads) {")
    out.append("    const int64_t i0 = q << 2;")
    out.append("    const int nv = (A.n - i0) < 4 ? (int)(A.n - i0) : 4;")
    out.append("    P p[4];")
    out.append("    int32_t g[4];")
    out.append("    gjb::load4_idx(A.gather, i0, nv, g);")
    out.append("    uint32_t w[4];")
    out.append("    (void)g0; (void)w;")
    out.extend(pre)
    out.append("#pragma unroll")
    out.append("    for (int u = 0; u < 4; ++u) {")
    out.append("      const gjb::Lane rng = gjb::make_lane(key0, key1, A.idx_offset + (uint64_t)(i0 + u));")
    out.append("      (void)rng;")
    out.append("      float score = 0.0f, weight = 0.0f;")
    out.extend(bind)
    out.append(body)
    out.extend(save)
    out.append("      p[u].score = score; p[u].weight = weight;")
    out.append("    }")
    out.extend(post)
    out.append("    if (A.score_out) { for (int u = 0; u < 4; ++u) w[u] = gjb::as_u(p[u].score); gjb::store4(A.score_out, i0, nv, w); }")
    out.append("    if (A.weight_out || A.wmax) {")
    out.append("      uint32_t wi[4] = {0u, 0u, 0u, 0u}, si[4] = {0u, 0u, 0u, 0u};")
    out.append("      if (A.weight_in) gjb::load4(A.weight_in, i0, nv, g0, false, false, wi);")
    out.append("      if (A.score_in) gjb::load4(A.score_in, i0, nv, g0, false, false, si);")
    out.append("      for (int u = 0; u < 4; ++u) {")
    out.append("        float t = p[u].weight;")
    out.append("        if (A.weight_in) t = gjb::as_f(wi[u]) + t;")
    out.append("        if (A.score_in) t = t - gjb::as_f(si[u]);")
    out.append("        w[u] = gjb::as_u(t);")
    out.append("        if (u < nv) run_max = fmaxf(run_max, t);")
    out.append("      }")
    out.append("      if (A.weight_out) gjb::store4(A.weight_out, i0, nv, w);")
    out.append("    }")
    out.append("  }")
    out.append("  if (A.wmax) gjb::block_wmax(run_max, A.wmax);")
    out.append("}")
    out.append("}  // namespace")
    out.append(_extern_c(ir, "quad", 1, work_per_thread=4))
    return "\n".join(out) + "\n"


# ============================================================= group mapping


def _generate_group(ir: ModelIR, G: int) -> str:
    em = _Emitter(ir, group=True)
    D = ir.width
    ns = len(ir.sites)
    decl, stage = _shared_decls(ir)
    for s in ir.sites:
        em.emit_site(s)
    ret_names = []
    for r in ir.ret_leaves:
        if isinstance(r, Expr):
            em.emit_expr(r)
            ret_names.append(em.ref(r) if r.ndim == 0 else em.as_v4(r))
        else:
            ret_names.append(E.fmt_float(float(r)))
    body = "\n".join(em.lines)

    pre: list[str] = []
    post: list[str] = []
    for i, a in enumerate(ir.args):
        if a.kind != "particle":
            continue
        if a.shape == ():
            ct = "int" if a.dtype == I32 else "float"
            cast = "const int*" if a.dtype == I32 else "const float*"
            pre.append(f"      const {ct} a{i} = valid ? __ldg(reinterpret_cast<{cast}>(A.args[{i}]) + row) : 0;")
        else:
            pre.append(f"      const gjb::V4 a{i} = valid ? gjb::v4_ldg(reinterpret_cast<const float*>(A.args[{i}]) + row * {D} + 4 * lane) : gjb::v4_splat(0.0f);")
    for s in ir.sites:
        j = s.index
        if s.value.ndim == 0:
            ct = "int" if s.value.dtype == I32 else "float"
            cast = "const int*" if s.value.dtype == I32 else "const float*"
            pre.append(f"      {ct} in_s{j} = 0;")
            pre.append(f"      if (!(fl{j} & GJB_SITE_SAMPLE) && valid) in_s{j} = __ldg(reinterpret_cast<{cast}>(A.site_in[{j}]) + ((fl{j} & GJB_SITE_BCAST) ? 0 : i));")
            post.append(f"      if (A.site_out[{j}] && valid && lane == 0) reinterpret_cast<{ct}*>(A.site_out[{j}])[i] = s{j};")
        else:
            pre.append(f"      gjb::V4 in_s{j} = gjb::v4_splat(0.0f);")
            pre.append(f"      if (!(fl{j} & GJB_SITE_SAMPLE) && valid) in_s{j} = gjb::v4_ldg(reinterpret_cast<const float*>(A.site_in[{j}]) + ((fl{j} & GJB_SITE_BCAST) ? 0 : i * {D}) + 4 * lane);")
            post.append(f"      if (A.site_out[{j}] && valid) *reinterpret_cast<float4*>(reinterpret_cast<float*>(A.site_out[{j}]) + i * {D} + 4 * lane) = gjb::v4_to(s{j});")
    for k, r in enumerate(ir.ret_leaves):
        if isinstance(r, Expr) and r.ndim == 1:
            post.append(f"      if (A.ret_out[{k}] && valid) *reinterpret_cast<float4*>(reinterpret_cast<float*>(A.ret_out[{k}]) + i * {D} + 4 * lane) = gjb::v4_to({ret_names[k]});")
        else:
            ct = "int" if isinstance(r, Expr) and r.dtype == I32 else "float"
            post.append(f"      if (A.ret_out[{k}] && valid && lane == 0) reinterpret_cast<{ct}*>(A.ret_out[{k}])[i] = {ret_names[k]};")

    out: list[str] = []
    out.append(f"// generated by genjax_b200.gen.codegen -- model '{ir.name}' [{ir.digest}] (group mapping, G={G})")
    out.append('#include "gjb_model.cuh"')
    out.append("namespace {")
    out.extend(em.consts)
    out.append("constexpr int kThreads = 256;")
    out.append(f"constexpr int G = {G};")
    out.append("constexpr int kPPB = kThreads / G;  // particles per block iteration")
    out.append("__global__ void __launch_bounds__(kThreads) model_kernel(const __grid_constant__ gjb_model_args A) {")
    out.extend(decl)
    out.extend(stage)
    if stage:
        out.append("  __syncthreads();")
    for j in range(ns):
        out.append(f"  const uint32_t fl{j} = A.site_flags[{j}];")
    out.append("  const bool need_score = A.score_out != nullptr;")
    out.append("  const uint32_t key0 = A.key_dev ? __ldg(A.key_dev) : A.key0, key1 = A.key_dev ? __ldg(A.key_dev + 1) : A.key1;")
    out.append("  float run_max = -INFINITY;")
    out.append("  const int lane = threadIdx.x & (G - 1);")
    out.append("  const int sub = threadIdx.x / G;")
    out.append("  for (int64_t base = (int64_t)blockIdx.x * kPPB; base < A.n; base += (int64_t)gridDim.x * kPPB) {")
    out.append("    {")
    out.append("      const int64_t i = base + sub;")
    out.append("      const bool valid = i < A.n;")
    out.append("      const int64_t row = valid ? (A.gather ? (int64_t)__ldg(A.gather + i) : i) : 0;")
    out.append("      (void)row;")
    out.append("      const gjb::Lane rng = gjb::make_lane(key0, key1, A.idx_offset + (uint64_t)i);")
    out.append("      float score = 0.0f, weight = 0.0f, vscore = 0.0f, vweight = 0.0f;")
    out.extend(pre)
    out.append(body)
    out.append("      if (need_score) score += gjb::group_sum<G>(vscore);")
    out.append("      weight += gjb::group_sum<G>(vweight);")
    out.extend(post)
    out.append("      if (A.score_out && valid && lane == 0) A.score_out[i] = score;")
    out.append("      if ((A.weight_out || A.wmax) && valid) {")
    out.append("        float t = weight;")
    out.append("        if (A.weight_in) t = __ldg(A.weight_in + i) + t;")
    out.append("        if (A.score_in) t = t - __ldg(A.score_in + i);")
    out.append("        if (A.weight_out && lane == 0) A.weight_out[i] = t;")
    out.append("        run_max = fmaxf(run_max, t);")
    out.append("      }")
    out.append("    }")
    out.append("  }")
    out.append("  if (A.wmax) gjb::block_wmax(run_max, A.wmax);")
    out.append("}")
    out.append("}  // namespace")
    out.append(_extern_c(ir, "group", G, work_per_thread=0))
    return "\n".join(out) + "\n"


def _extern_c(ir: ModelIR, mapping: str, G: int, work_per_thread: int) -> str:
    info = _info_json(ir, mapping, G).replace("\\", "\\\\").replace('"', '\\"')
    if mapping == "quad":
        work = "(a->n + 3) / 4"
        per_block = "kThreads"
    else:
        work = "a->n"
        per_block = "kPPB"
    return f"""
extern "C" {{
const char* gjb_model_info(void) {{ return "{info}"; }}

int gjb_model_launch(const gjb_model_args* a, void* stream) {{
  if (!a || a->n < 0) return GJB_E_ARG;
  if (a->n == 0) return 0;
  const int64_t work = {work};
  int64_t blocks = (work + {per_block} - 1) / {per_block};
  const int64_t cap = 148 * 8;  // persistent-style grid: a multiple of the 148 SMs
  if (blocks > cap) blocks = cap;
  model_kernel<<<(int)blocks, kThreads, 0, (cudaStream_t)stream>>>(*a);
  return (int)cudaGetLastError();
}}

int gjb_model_mh_chain(const gjb_chain_args* a, void* stream) {{ (void)a; (void)stream; return GJB_E_MODE; }}
int gjb_model_hmc_chain(const gjb_chain_args* a, void* stream) {{ (void)a; (void)stream; return GJB_E_MODE; }}
}}
"""

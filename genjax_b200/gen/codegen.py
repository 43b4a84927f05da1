"""ModelIR -> one CUDA translation unit (sm_100a) exporting the per-model
C-ABI of include/genjax_b200.h (``gjb_model_info`` / ``gjb_model_launch`` /
``gjb_model_pf_run`` / ``gjb_model_mh_chain`` / ``gjb_model_hmc_chain``).

The generated per-particle body is the fused, batched form of the reference's
static-language handlers (static.py:254-278, 298-321, 341-380, 407-466,
616-673): sites are visited in program order; per-site launch flags choose
sample-vs-read and whether the site's logpdf joins the importance weight, so
ONE body serves simulate / assess / generate / update / regenerate.  It is
instantiated twice:

  * ``model_kernel``  one GFI call over n particles (grid-stride, one wave);
  * ``pf_kernel``     the whole T-step bootstrap particle filter as ONE
    persistent cooperative launch: per step (A) ancestor gather + propose +
    logpdf + running max, (B) exact integer weight mass per CTA, (C) CDF scan
    + systematic offspring ranges, separated by grid barriers.

Two thread mappings:
  * "quad"  (all-scalar models or odd event widths): a thread owns the 4
    consecutive particles of a global quad, 128-bit loads/stores of every [N]
    array, ONE Philox block per (quad, sampled site);
  * "group" (event width D, D % 4 == 0, D/4 a power of two <= 32): G = D/4
    lanes own one particle, each lane one float4 of every width-D value;
    logpdf partials are reduced with warp shuffles.

Particle-invariant sub-expressions (functions of constants, scalar and shared
arguments only: 1/scale, log scale, table rows ...) are hoisted out of the
particle loop into a per-thread ``Uni`` struct computed once per launch.
"""

from __future__ import annotations

import json
import os

import numpy as np

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
    vectors (True) or per-thread float arrays (False).  Values that do not
    depend on particle data are emitted into the ``Uni`` struct instead."""

    def __init__(self, ir: ModelIR, group: bool, const_prefix: str = "cv"):
        self.ir = ir
        self.group = group
        self.const_prefix = const_prefix
        self.D = ir.width
        self.lines: list[str] = []  # per-particle body
        self.uni_decl: list[str] = []  # members of struct Uni
        self.uni_init: list[str] = []  # statements of make_uni()
        self.names: dict[int, str] = {}
        self.consts: list[str] = []  # namespace-scope constant arrays
        self._uniform: dict[int, bool] = {}
        self._v4cache: dict[int, str] = {}
        self.n_tmp = 0
        self.hoist_vectors = True  # False: width-D temporaries are recomputed per use instead of living in Uni

    def w(self, s: str):
        self.lines.append("      " + s)

    def fresh(self, prefix="e") -> str:
        self.n_tmp += 1
        return f"{prefix}{self.n_tmp}"

    # ---- classification
    def is_uniform(self, e: Expr) -> bool:
        u = self._uniform.get(e._id)
        if u is None:
            if e.op in ("const", "constvec"):
                u = True
            elif e.op == "arg":
                u = e.attr["kind"] != "particle"
            elif e.op in ("site", "chain_step"):
                u = False
            else:
                u = all(self.is_uniform(i) for i in e.ins)
                if u and e.ndim == 1 and e.op != "row" and not self.hoist_vectors:
                    u = False
            self._uniform[e._id] = u
        return u

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
        # pointer vector of width D: this lane's 4 elements, loaded once per launch when particle-invariant
        r = self.ref(e)
        load = (f"gjb::v4_load({r} + 4 * lane)" if self._is_shared_vec(e) else
                f"gjb::V4{{{{{r}[4*lane], {r}[4*lane+1], {r}[4*lane+2], {r}[4*lane+3]}}}}")
        if not self.is_uniform(e):
            return load
        c = self._v4cache.get(e._id)
        if c is None:
            c = self.fresh("v")
            self.uni_decl.append(f"  gjb::V4 {c};")
            self.uni_init.append(f"  U.{c} = {load.replace(r, self._uni_local(r))};")
            c = f"U.{c}"
            self._v4cache[e._id] = c
        return c

    @staticmethod
    def _uni_local(name: str) -> str:
        return name

    # ---- expression emission
    def emit_expr(self, e: Expr):
        if e._id in self.names:
            return
        for i in e.ins:
            self.emit_expr(i)
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
            cname = f"{self.const_prefix}{len(self.consts)}"
            self.consts.append(f"__device__ const float {cname}[{len(e.attr)}] = {{{vals}}};")
            self.names[e._id] = cname
            return
        uni = self.is_uniform(e)
        n = self.fresh("h" if uni else "e")
        t = self.ctype(e)

        def out(decl_type: str, rhs: str, array: int = 0):
            """define value `n` (in Uni or in the body)"""
            if uni:
                self.uni_decl.append(f"  {decl_type} {n}{f'[{array}]' if array else ''};")
                self.names[e._id] = f"U.{n}"
            else:
                self.names[e._id] = n
            return f"U.{n}" if uni else n

        emit = self.uni_init.append if uni else self.w
        pad = "  " if uni else ""
        if op == "row":
            mat, idx = e.ins
            K = mat.shape[1]
            nm = out("const float*", "")
            emit(f"{pad}{'' if uni else 'const float* '}{nm} = {self.ref(mat)} + (int)({self.ref(idx)}) * {K};")
            return
        if op == "gather1":
            vec, idx = e.ins
            ct = "int" if e.dtype == I32 else "float"
            nm = out(ct, "")
            val = f"{self.ref(vec)}[(int)({self.ref(idx)})]"
            if e.dtype == I32:
                val = f"__float_as_int({val})"
            emit(f"{pad}{'' if uni else 'const ' + ct + ' '}{nm} = {val};")
            return
        if op == "elem":
            (vec,) = e.ins
            j = int(e.attr)
            nm = out("float", "")
            if self.ctype(vec) == "gjb::V4":
                val = f"__shfl_sync(0xffffffffu, {self.ref(vec)}.v[{j % 4}], (threadIdx.x & 31 & ~(G - 1)) + {j // 4})"
            else:
                val = f"{self.ref(vec)}[{j}]"
            emit(f"{pad}{'' if uni else 'const float '}{nm} = {val};")
            return
        if op == "sum":
            (x,) = e.ins
            nm = out("float", "")
            if self.ctype(x) == "gjb::V4":
                emit(f"{pad}{'' if uni else 'const float '}{nm} = gjb::group_sum<G>(gjb::v4_hsum({self.ref(x)}));")
            else:
                K = x.shape[0]
                emit(f"{pad}{'' if uni else 'float '}{nm} = 0.0f;")
                emit(f"{pad}for (int k = 0; k < {K}; ++k) {nm} += {self.ref(x)}[k];")
            return
        if op == "cast":
            (x,) = e.ins
            if t in ("int", "float"):
                nm = out(t, "")
                emit(f"{pad}{'' if uni else 'const ' + t + ' '}{nm} = ({t}){self.ref(x)};")
                return
            raise NotImplementedError("vector casts")
        # elementwise ----------------------------------------------------
        if t in ("int", "float"):
            nm = out(t, "")
            emit(f"{pad}{'' if uni else 'const ' + t + ' '}{nm} = {self.scalar_rhs(e, None)};")
        elif t == "gjb::V4":
            rhs = self.v4_rhs(e)
            nm = out("gjb::V4", "")
            emit(f"{pad}{'' if uni else 'const gjb::V4 '}{nm} = {rhs};")
        else:
            K = e.shape[0]
            nm = out("float", "", array=K)
            if not uni:
                emit(f"float {nm}[{K}];")
            emit(f"{pad}for (int k = 0; k < {K}; ++k) {nm}[k] = {self.scalar_rhs(e, 'k')};")

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
            return f"((int)U.sc[{i}])" if e.dtype == I32 else f"U.sc[{i}]"
        if kind == "shared":
            if e.ndim == 0:
                return f"(__float_as_int(sh{i}[0]))" if e.dtype == I32 else f"sh{i}[0]"
            return f"sh{i}"
        return f"a{i}"  # particle arg: local loaded before the body

    # ---- particle-invariant pieces of the Normal family ----------------
    def normal_consts(self, j: int, scale: Expr, vec_width: int = 0) -> tuple[str, str]:
        """(inv, lc) names: inv = 1/scale, lc = sum_k (0.5 log 2pi + log scale_k); hoisted when scale is uniform."""
        uni = self.is_uniform(scale)
        emit = self.uni_init.append if uni else self.w
        pad = "  " if uni else ""
        pre = "U." if uni else ""
        inv, lc = f"inv{j}", f"lc{j}"
        if vec_width == 0:
            if uni:
                self.uni_decl.append(f"  float {inv}, {lc};")
            emit(f"{pad}{'' if uni else 'const float '}{pre}{inv} = 1.0f / {self.ref(scale)};")
            emit(f"{pad}{'' if uni else 'const float '}{pre}{lc} = gjb::kHalfLog2Pi + logf({self.ref(scale)});")
        elif self.group:
            if uni:
                self.uni_decl.append(f"  gjb::V4 {inv}; float {lc};")
            sv = self.as_v4(scale)
            emit(f"{pad}{'' if uni else 'const gjb::V4 '}{pre}{inv} = gjb::f_reciprocal({sv});")
            emit(f"{pad}{'' if uni else 'const float '}{pre}{lc} = gjb::v4_hsum(gjb::kHalfLog2Pi + gjb::f_log({sv}));")
        else:
            D = vec_width
            if uni:
                self.uni_decl.append(f"  float {inv}[{D}]; float {lc};")
            else:
                emit(f"float {inv}[{D}]; float {lc};")
            emit(f"{pad}{pre}{lc} = 0.0f;")
            emit(f"{pad}for (int k = 0; k < {D}; ++k) {{ {pre}{inv}[k] = 1.0f / {self.elem_ref(scale, 'k')}; "
                 f"{pre}{lc} += gjb::kHalfLog2Pi + logf({self.elem_ref(scale, 'k')}); }}")
        return pre + inv, pre + lc

    # ---- sites
    def emit_site(self, s):
        d = s.dist
        j = s.index
        for a in s.args:
            self.emit_expr(a)
        fl = f"FL({j})"
        samp, wt = f"({fl} & GJB_SITE_SAMPLE)", f"({fl} & GJB_SITE_WEIGHT)"
        if s.cmask is not None:
            # Mask-ed constraint (distribution.py:129-142): per particle, bit 0 = take the supplied value, else draw one;
            # bit 1 = keep this site out of the weight
            self.emit_expr(s.cmask)
            cm = self.ref(s.cmask)
            samp = f"(({fl} & GJB_SITE_SAMPLE) || !({cm} & 1))"
            wt = f"(({fl} & GJB_SITE_WEIGHT) && !({cm} & 2))"
        gates = []
        for pred in (s.live, s.scored):
            if pred is not None:
                self.emit_expr(pred)
                gates.append(f"({self.ref(pred)} != 0)")
        live = gates[0] if s.live is not None else None
        need = f"(need_score || {wt})"
        if gates:
            need = f"({need} && {' && '.join(gates)})"
        self.w(f"// site {j} {'/'.join(map(str, s.addr))!r}: {d.name}"
               + (" [dynamic: exists / is scored only where its Switch branch / Mask flag holds]" if gates else ""))
        if not d.vector:
            vt = "int" if s.value.dtype == I32 else "float"
            if d.name == "categorical":
                lg = s.args[0]
                a = [(self.ref(lg), lg.shape[0])]
            else:
                a = [self.ref(x) for x in s.args]
            self.w(f"{vt} s{j};")
            self.w(f"if {samp} s{j} = {d.emit_sample(j, a, self)}; else s{j} = in_s{j};")
            if live:
                self.w(f"if (!{live}) s{j} = 0;")
            if d.name == "normal":
                inv, lc = self.normal_consts(j, s.args[1])
                lp = f"gjb::Normal::logpdf_r(s{j}, {a[0]}, {inv}, {lc})"
            else:
                lp = d.emit_logpdf(f"s{j}", a, self)
            self.w(f"if {need} {{ const float lp = {lp}; score += lp; if {wt} weight += lp; }}")
            return
        if getattr(d, "base", None) is not None:
            # dist.repeat / dist.vmap: N independent draws of a scalar primitive in one vector site (per-thread arrays)
            D = s.value.shape[0]
            b = d.base
            ak = ", ".join(self.elem_ref(x, "k") for x in s.args)
            draw = "zz[t]" if b.rng_kind == "normal" else "gjb::u01(ww[t])"
            self.w(f"float s{j}[{D}];")
            self.w(f"if {samp} {{")
            self.w(f"  for (int c = 0; c < {(D + 3) // 4}; ++c) {{")
            if b.rng_kind == "normal":
                self.w(f"    const float4 z = gjb::normal4(rng, {j + 1}u, (uint32_t)c); const float zz[4] = {{z.x, z.y, z.z, z.w}};")
            else:
                self.w(f"    const uint4 wd = rng.words({j + 1}u, (uint32_t)c); const uint32_t ww[4] = {{wd.x, wd.y, wd.z, wd.w}};")
            self.w(f"    for (int t = 0; t < 4; ++t) {{ const int k = 4 * c + t; if (k < {D}) s{j}[k] = gjb::{b.cuda}::sample({draw}, {ak}); }} }}")
            self.w(f"}} else {{ for (int k = 0; k < {D}; ++k) s{j}[k] = in_s{j}[k]; }}")
            if live:
                self.w(f"if (!{live}) {{ for (int k = 0; k < {D}; ++k) s{j}[k] = 0.0f; }}")
            self.w(f"if {need} {{ float lp = 0.0f; for (int k = 0; k < {D}; ++k) lp += gjb::{b.cuda}::logpdf(s{j}[k], {ak}); "
                   f"score += lp; if {wt} weight += lp; }}")
            return
        if d.name == "gmm_diag":
            logits, mu, sigma = s.args
            K, D = mu.shape
            a = f"{self.ref(logits)}, {self.ref(mu)}, {self.ref(sigma)}"
            self.w(f"float s{j}[{D}];")
            self.w(f"if {samp} gjb::GmmDiag::sample<{K}, {D}>(rng, {j + 1}u, {a}, s{j}); else {{ for (int k = 0; k < {D}; ++k) s{j}[k] = in_s{j}[k]; }}")
            if live:
                self.w(f"if (!{live}) {{ for (int k = 0; k < {D}; ++k) s{j}[k] = 0.0f; }}")
            self.w(f"if {need} {{ const float lp = gjb::GmmDiag::logpdf<{K}, {D}>(s{j}, {a}); score += lp; if {wt} weight += lp; }}")
            return
        if d.name == "mv_normal":
            loc, cov = s.args
            D = s.value.shape[0]
            # Cholesky factor of the shared covariance: particle-invariant, once per thread per launch
            self.uni_decl.append(f"  float L{j}[{D * D}]; float ld{j};")
            self.uni_init.append(f"  U.ld{j} = gjb::MvNormal::cholesky<{D}>({self.ref(cov)}, U.L{j});")
            lc = self.ref(loc)
            self.w(f"float s{j}[{D}];")
            self.w(f"if {samp} gjb::MvNormal::sample<{D}>(rng, {j + 1}u, {lc}, U.L{j}, s{j}); else {{ for (int k = 0; k < {D}; ++k) s{j}[k] = in_s{j}[k]; }}")
            if live:
                self.w(f"if (!{live}) {{ for (int k = 0; k < {D}; ++k) s{j}[k] = 0.0f; }}")
            self.w(f"if {need} {{ const float lp = gjb::MvNormal::logpdf<{D}>(s{j}, {lc}, U.L{j}, U.ld{j}); score += lp; if {wt} weight += lp; }}")
            return
        if d.name != "mv_normal_diag":
            raise NotImplementedError(d.name)
        loc, scale = s.args
        D = s.value.shape[0]
        inv, lc = self.normal_consts(j, scale, vec_width=D)
        if self.group:
            self.w(f"gjb::V4 s{j};")
            self.w(f"if {samp} s{j} = gjb::mvn_diag_sample(rng, {j + 1}u, (uint32_t)lane, {self.as_v4(loc)}, {self.as_v4(scale)}); else s{j} = in_s{j};")
            if live:
                self.w(f"if (!{live}) s{j} = gjb::v4_splat(0.0f);")
            self.w(f"if {need} {{ const float lp = gjb::mvn_diag_logpdf4_r(s{j}, {self.as_v4(loc)}, {inv}, {lc}); vscore += lp; if {wt} vweight += lp; }}")
        else:
            self.w(f"float s{j}[{D}];")
            self.w(f"if {samp} {{")
            self.w(f"  for (int c = 0; c < {(D + 3) // 4}; ++c) {{ const float4 z = gjb::normal4(rng, {j + 1}u, (uint32_t)c); const float zz[4] = {{z.x, z.y, z.z, z.w}};")
            self.w(f"    for (int t = 0; t < 4; ++t) {{ const int k = 4 * c + t; if (k < {D}) s{j}[k] = {self.elem_ref(loc, 'k')} + {self.elem_ref(scale, 'k')} * zz[t]; }} }}")
            self.w(f"}} else {{ for (int k = 0; k < {D}; ++k) s{j}[k] = in_s{j}[k]; }}")
            if live:
                self.w(f"if (!{live}) {{ for (int k = 0; k < {D}; ++k) s{j}[k] = 0.0f; }}")
            self.w(f"if {need} {{ float lp = -{lc}; for (int k = 0; k < {D}; ++k) {{ const float z = s{j}[k] * {inv}[k] - {self.elem_ref(loc, 'k')} * {inv}[k]; lp -= 0.5f * (z * z); }} score += lp; if {wt} weight += lp; }}")


def _info_json(ir: ModelIR, mapping: str, G: int, pf_step: bool = False) -> str:
    info = {
        "pf_step": bool(pf_step),
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
    """namespace-scope __shared__ blocks for the shared (un-batched) arguments + their staging calls."""
    decl, stage = [], []
    for i, a in enumerate(ir.args):
        if a.kind == "shared":
            n = 1
            for d in a.shape:
                n *= d
            n = max(n, 1)
            decl.append(f"__shared__ __align__(16) float sh{i}[{(n + 3) // 4 * 4}];")
            stage.append(f"  gjb::stage_shared(sh{i}, ARGS[{i}], {n});")
    return decl, stage


def generate(ir: ModelIR, pf_obs: tuple | None = None, chain=None) -> str:
    """CUDA source for ``ir``.  ``pf_obs`` (site indices observed at every
    filter step) bakes the per-site flags of the persistent filter kernel in at
    compile time, so the paths a bootstrap filter never takes (reading proposed
    sites, scoring unweighted ones) are removed from ``pf_kernel``."""
    G = group_lanes(ir.width)
    if any(s.dist.vector and s.dist.name != "mv_normal_diag" for s in ir.sites):
        G = 0  # lane-group kernels know mv_normal_diag only; other vector primitives run on per-thread arrays
    gen = _Generator(ir, G, pf_obs, chain)
    return gen.source()


class _Generator:
    def __init__(self, ir: ModelIR, G: int, pf_obs: tuple | None = None, chain=None):
        self.ir = ir
        self.G = G
        self.chain = chain
        self.pf_obs = None if pf_obs is None else tuple(sorted(int(j) for j in pf_obs))
        self.group = G > 0
        self.em = _Emitter(ir, group=self.group)
        self.ns = len(ir.sites)
        self.na = max(len(ir.args), 1)
        # what the model kernels write through ret_out: the return leaves, then the validity flags of dynamic sites
        self.all_rets = list(ir.ret_leaves) + list(ir.flag_leaves)
        self.nr = max(len(self.all_rets), 1)
        em = self.em
        for s in ir.sites:
            em.emit_site(s)
        self.ret_names = []
        for r in self.all_rets:
            if isinstance(r, Expr):
                em.emit_expr(r)
                if self.group:
                    self.ret_names.append(em.ref(r) if r.ndim == 0 else em.as_v4(r))
                else:
                    self.ret_names.append(em.ref(r))
            else:
                self.ret_names.append(E.fmt_float(float(r)))
        self.body = "\n".join(em.lines)
        self.needs_lane = any(s.dist.rng_kind == "lane" for s in ir.sites)

    # ------------------------------------------------------------ pieces
    def header(self) -> list[str]:
        ir, em = self.ir, self.em
        decl, stage = _shared_decls(ir)
        self.stage = stage
        out = [f"// generated by genjax_b200.gen.codegen -- model '{ir.name}' [{ir.digest}] "
               f"({'group mapping, G=%d' % self.G if self.group else 'quad mapping'})",
               *(["#define GJB_NO_PREFETCH  // parent rows are loaded tile by tile (prefetching the next tile spills at 64 registers: measured slower)"]
                 if os.environ.get("GJB_STEP_PREFETCH", "0") == "0" else []),
               '#include "gjb_model.cuh"', '#include "gjb_resample.cuh"', "namespace {"]
        out.extend(em.consts)
        out.append("constexpr int kThreads = 256;")
        out.append(f"constexpr int kPfMinBlocks = {3 if self.group else 4};  // CTAs per SM the filter kernel is compiled for")
        out.append("constexpr int kPfCacheTiles = 2;  // tiles per CTA whose masses stay in shared memory between phases")
        out.append(f"constexpr int NS = {max(self.ns, 1)}, NA = {self.na}, NR = {self.nr};")
        out.append(f"constexpr int kNState = {len(self.ir.ret_leaves)};  // return leaves proper (NR also counts the validity flags of dynamic sites)")
        if self.group:
            out.append(f"constexpr int G = {self.G};")
            out.append("constexpr int kPPB = kThreads / G;  // particles per block iteration")
        out.extend(decl)
        if self.pf_obs is not None:
            vals = ", ".join("(GJB_SITE_WEIGHT | GJB_SITE_BCAST)" if j in self.pf_obs else "GJB_SITE_SAMPLE"
                             for j in range(self.ns)) or "0u"
        else:
            vals = ", ".join("0u" for _ in range(max(self.ns, 1)))
        out.append(f"__device__ constexpr uint32_t kPfFl[NS] = {{{vals}}};  // flags baked into pf_kernel")
        out.append(f"constexpr uint32_t kPfFl_host[NS] = {{{vals}}};")
        out.append(f"constexpr bool kPfStatic = {'true' if self.pf_obs is not None else 'false'};")
        out.append("#define FL(j) (kSt ? kPfFl[j] : fl[j])")
        out.append("struct Io {  // what one launch (or one filter step) reads and writes")
        out.append("  const void* args[NA]; const void* site_in[NS]; void* site_out[NS]; void* ret_out[NR];")
        out.append("  const int32_t* gather; const float* score_in; const float* weight_in; float* score_out; float* weight_out;")
        out.append("  const gjb_peers* peers;  // nullable device array [NA]: gathered rows may live on peer ranks")
        out.append("  const float* m_ref; unsigned long long* tile_mass;  // reference-maximum step (kMass instantiation only)")
        out.append("  float* te_w;  // nullable (generic pointer): the weights also go here (single-launch step: shared memory)")
        out.append("};")
        out.append("struct Uni {  // particle-invariant values, computed once per thread per launch")
        out.append("  float sc[NA];")
        out.extend(em.uni_decl)
        out.append("};")
        out.append("__device__ __forceinline__ void make_uni(Uni& U, const float* __restrict__ scalars) {")
        if self.group:
            out.append("  const int lane = threadIdx.x & (G - 1); (void)lane;")
        out.append("  for (int i = 0; i < NA; ++i) U.sc[i] = scalars[i];")
        out.extend(em.uni_init)
        out.append("}")
        return out

    def stage_lines(self, args_expr: str) -> list[str]:
        if not self.stage:
            return []
        return [s.replace("ARGS", args_expr) for s in self.stage] + ["  __syncthreads();"]

    # ------------------------------------------------------------- quads
    def run_quads(self) -> list[str]:
        ir = self.ir
        P = ["struct P {"]
        pre: list[str] = []
        bind: list[str] = []
        post: list[str] = []
        save: list[str] = []
        for i, a in enumerate(ir.args):
            if a.kind != "particle":
                continue
            ct = "int" if a.dtype == I32 else "float"
            if a.shape == ():
                P.append(f"  {ct} a{i};")
                conv = "(int)w[u]" if a.dtype == I32 else "gjb::as_f(w[u])"
                pre.append(f"    gjb::load4<kCg>(io.args[{i}], i0, lo, hi, g, io.gather != nullptr, false, w, (io.gather && io.peers) ? io.peers + {i} : nullptr);")
                pre.append(f"    for (int u = 0; u < 4; ++u) p[u].a{i} = {conv};")
                bind.append(f"      const {ct} a{i} = p[u].a{i};")
            else:
                D = a.shape[0]
                P.append(f"  float a{i}[{D}];")
                pre.append(f"    for (int u = 0; u < 4; ++u) {{ int64_t row = io.gather ? (int64_t)g[u] : i0 + u;")
                pre.append(f"      const float* rb = reinterpret_cast<const float*>(gjb::arg_base(io.args[{i}], (io.gather && io.peers) ? io.peers + {i} : nullptr, row));")
                pre.append(f"      for (int k = 0; k < {D}; ++k) p[u].a{i}[k] = (u >= lo && u < hi) ? gjb::ldf<kCg>(rb + row * {D} + k) : 0.0f; }}")
                bind.append(f"      const float* a{i} = p[u].a{i};")
        for s in ir.sites:
            j = s.index
            ct = "int" if s.value.dtype == I32 else "float"
            if s.value.ndim == 0:
                P.append(f"  {ct} s{j};")
                conv = "(int)w[u]" if s.value.dtype == I32 else "gjb::as_f(w[u])"
                pre.append(f"    if (!(FL({j}) & GJB_SITE_SAMPLE)) {{ gjb::load4<false>(io.site_in[{j}], i0, lo, hi, g0, false, (FL({j}) & GJB_SITE_BCAST) != 0, w);")
                pre.append(f"      for (int u = 0; u < 4; ++u) p[u].s{j} = {conv}; }}")
                bind.append(f"      const {ct} in_s{j} = p[u].s{j};")
                save.append(f"      p[u].s{j} = s{j};")
                post.append(f"    if (io.site_out[{j}]) {{ for (int u = 0; u < 4; ++u) w[u] = gjb::as_u(p[u].s{j}); gjb::store4(io.site_out[{j}], i0, lo, hi, w); }}")
            else:
                D = s.value.shape[0]
                P.append(f"  float s{j}[{D}];")
                pre.append(f"    if (!(FL({j}) & GJB_SITE_SAMPLE)) {{ for (int u = 0; u < 4; ++u) for (int k = 0; k < {D}; ++k)")
                pre.append(f"      p[u].s{j}[k] = (u >= lo && u < hi) ? __ldg(reinterpret_cast<const float*>(io.site_in[{j}]) + ((FL({j}) & GJB_SITE_BCAST) ? 0 : (i0 + u) * {D}) + k) : 0.0f; }}")
                bind.append(f"      const float* in_s{j} = p[u].s{j};")
                save.append(f"      for (int k = 0; k < {D}; ++k) p[u].s{j}[k] = s{j}[k];")
                post.append(f"    if (io.site_out[{j}]) {{ for (int u = lo; u < hi; ++u) for (int k = 0; k < {D}; ++k) reinterpret_cast<float*>(io.site_out[{j}])[(i0 + u) * {D} + k] = p[u].s{j}[k]; }}")
        for k, r in enumerate(self.all_rets):
            is_vec = isinstance(r, Expr) and r.ndim == 1
            is_int = isinstance(r, Expr) and r.dtype == I32
            if is_vec:
                D = r.shape[0]
                P.append(f"  float r{k}[{D}];")
                save.append(f"      for (int k = 0; k < {D}; ++k) p[u].r{k}[k] = {self.ret_names[k]}[k];")
                post.append(f"    if (io.ret_out[{k}]) {{ for (int u = lo; u < hi; ++u) for (int k = 0; k < {D}; ++k) reinterpret_cast<float*>(io.ret_out[{k}])[(i0 + u) * {D} + k] = p[u].r{k}[k]; }}")
            else:
                P.append(f"  {'int' if is_int else 'float'} r{k};")
                save.append(f"      p[u].r{k} = {self.ret_names[k]};")
                post.append(f"    if (io.ret_out[{k}]) {{ for (int u = 0; u < 4; ++u) w[u] = gjb::as_u(p[u].r{k}); gjb::store4(io.ret_out[{k}], i0, lo, hi, w); }}")
        P.append("  float score, weight;")
        P.append("};")

        # RNG of one quad: the Philox block (and its Box-Muller normals) of every scalar site that may be sampled from the
        # quad stream.  Emitted as its own function so that the single-launch filter step can run it BEFORE it resolves its
        # ancestors (pure ALU work that overlaps the tile-record loads), and as a struct the body reads.
        qrng_decl: list[str] = ["struct QRng {  // quad-stream random words / normals of one quad, per scalar site"]
        qrng_fill: list[str] = []
        rngpre: list[str] = ["    QRng R;",
                             "    if (kPre) R = *pre_rng; else quad_rng<kSt>(fl, key0, key1, quad0 + (uint64_t)ql, R);"]
        n_q = 0
        for s in ir.sites:
            j = s.index
            kind = s.dist.rng_kind
            if s.dist.vector or kind == "lane":
                continue
            n_q += 1
            qrng_decl.append(f"  uint4 W{j};")
            qrng_fill.append(f"  R.W{j} = make_uint4(0u, 0u, 0u, 0u);")
            rngpre.append(f"    const uint4& W{j} = R.W{j}; (void)W{j};")
            if kind == "normal":
                qrng_decl.append(f"  float4 Z{j};")
                qrng_fill.append(f"  R.Z{j} = make_float4(0.f, 0.f, 0.f, 0.f);")
                rngpre.append(f"    const float4& Z{j} = R.Z{j}; (void)Z{j};")
            draw_if = "true" if s.cmask is not None else f"FL({j}) & GJB_SITE_SAMPLE"
            qrng_fill.append(f"  if ({draw_if}) {{ R.W{j} = gjb::quad_words(key0, key1, quad, {j + 1}u);"
                             + (f" R.Z{j} = gjb::normal4_of(R.W{j});" if kind == "normal" else "") + " }")
        if n_q == 0:
            qrng_decl.append("  int unused;")
        qrng_decl.append("};")
        qrng_fn = ["template <bool kSt>",
                   "__device__ __forceinline__ void quad_rng(const uint32_t (&fl)[NS], uint32_t key0, uint32_t key1, uint64_t quad, QRng& R) {",
                   "  (void)fl; (void)key0; (void)key1; (void)quad; (void)R;"] + qrng_fill + ["}"]

        out = list(P)
        out.extend(qrng_decl)
        out.extend(qrng_fn)
        out.append("// quads [ql_begin, ql_end) step ql_stride of the launch; local particle i0 = 4*ql - (idx_offset & 3)")
        out.append("// kSm: io.gather points to shared memory and the weights also go to io.te_w; kPre: the quad's RNG was drawn by the caller")
        out.append("template <bool kCg, bool kSt, bool kMass = false, bool kSm = false, bool kPre = false>")
        out.append("__device__ __forceinline__ void run_quads(const Io& io, const Uni& U, const uint32_t (&fl)[NS], int64_t n,")
        out.append("    uint64_t idx_offset, uint32_t key0, uint32_t key1, int64_t ql_begin, int64_t ql_end, int64_t ql_stride, float& run_max,")
        out.append("    const QRng* pre_rng = nullptr) {")
        out.append("  const bool need_score = !kSt && io.score_out != nullptr;")
        out.append("  const float mref = kMass ? __ldg(io.m_ref) : 0.0f;  // reference maximum known before the launch")
        out.append("  const int shift = kSt ? 0 : (int)(idx_offset & 3);")
        out.append("  const uint64_t quad0 = idx_offset >> 2;")
        out.append("  const int32_t g0[4] = {0, 0, 0, 0};")
        out.append("  for (int64_t ql = ql_begin; ql < ql_end; ql += ql_stride) {")
        out.append("    const int64_t i0 = (ql << 2) - shift;")
        out.append("    const int lo = i0 < 0 ? (int)(-i0) : 0;")
        out.append("    const int hi = (n - i0) < 4 ? (int)(n - i0) : 4;")
        out.append("    P p[4];")
        out.append("    int32_t g[4];")
        out.append("    if (kSm) gjb::load4_idx_gen(io.gather, i0, lo, hi, g); else gjb::load4_idx<kCg>(io.gather, i0, lo, hi, g);")
        out.append("    uint32_t w[4];")
        out.append("    (void)g0; (void)w; (void)quad0;")
        out.extend(pre)
        out.extend(rngpre)
        out.append("#pragma unroll")
        out.append("    for (int u = 0; u < 4; ++u) {")
        out.append("      const int sub = u; (void)sub;")
        if self.needs_lane:
            out.append("      const gjb::Lane rng = gjb::make_lane(key0, key1, idx_offset + (uint64_t)(i0 + u));")
        out.append("      float score = 0.0f, weight = 0.0f;")
        out.extend(bind)
        out.append(self.body)
        out.extend(save)
        out.append("      p[u].score = score; p[u].weight = weight;")
        out.append("    }")
        out.extend(post)
        out.append("    if (!kSt && io.score_out) { for (int u = 0; u < 4; ++u) w[u] = gjb::as_u(p[u].score); gjb::store4(io.score_out, i0, lo, hi, w); }")
        out.append("    {")
        out.append("      uint32_t wi[4] = {0u, 0u, 0u, 0u}, si[4] = {0u, 0u, 0u, 0u};")
        out.append("      if (!kSt && io.weight_in) gjb::load4<false>(io.weight_in, i0, lo, hi, g0, false, false, wi);")
        out.append("      if (!kSt && io.score_in) gjb::load4<false>(io.score_in, i0, lo, hi, g0, false, false, si);")
        out.append("      for (int u = 0; u < 4; ++u) {")
        out.append("        float t = p[u].weight;")
        out.append("        if (!kSt && io.weight_in) t = gjb::as_f(wi[u]) + t;")
        out.append("        if (!kSt && io.score_in) t = t - gjb::as_f(si[u]);")
        out.append("        w[u] = gjb::as_u(t);")
        out.append("        if (u >= lo && u < hi) run_max = fmaxf(run_max, t);")
        out.append("      }")
        out.append("      if (io.weight_out) gjb::store4(io.weight_out, i0, lo, hi, w);")
        out.append("      if (kSm) gjb::store4(io.te_w, i0, lo, hi, w);")
        out.append("      if (kMass) {  // exact integer mass of the quad, added to its tile (a quad never straddles a tile)")
        out.append("        unsigned long long qs = 0ull;")
        out.append("        for (int u = 0; u < 4; ++u) if (u >= lo && u < hi) qs += gjb::det_exp_q(__fadd_rn(gjb::as_f(w[u]), -mref));")
        out.append("        if (qs) atomicAdd(io.tile_mass + ((i0 + lo) / gjb::kTile), qs);")
        out.append("      }")
        out.append("    }")
        out.append("  }")
        out.append("}")
        return out

    # ------------------------------------------------------------ groups
    def run_groups(self) -> list[str]:
        ir = self.ir
        D = ir.width
        pre: list[str] = []
        post: list[str] = []
        for i, a in enumerate(ir.args):
            if a.kind != "particle":
                continue
            if a.shape == ():
                ct = "int" if a.dtype == I32 else "float"
                cast = "const int*" if a.dtype == I32 else "const float*"
                pre.append(f"      int64_t row{i} = row; const void* rb{i} = gjb::arg_base(io.args[{i}], (io.gather && io.peers) ? io.peers + {i} : nullptr, row{i});")
                pre.append(f"      const {ct} a{i} = valid ? gjb::ldx<kCg>(reinterpret_cast<{cast}>(rb{i}) + row{i}) : 0;")
            else:
                pre.append(f"      int64_t row{i} = row; const void* rb{i} = gjb::arg_base(io.args[{i}], (io.gather && io.peers) ? io.peers + {i} : nullptr, row{i});")
                pre.append(f"      const gjb::V4 a{i} = valid ? gjb::v4_ld<kCg>(reinterpret_cast<const float*>(rb{i}) + row{i} * {D} + 4 * lane) : gjb::v4_splat(0.0f);")
        for s in ir.sites:
            j = s.index
            if s.value.ndim == 0:
                ct = "int" if s.value.dtype == I32 else "float"
                cast = "const int*" if s.value.dtype == I32 else "const float*"
                pre.append(f"      {ct} in_s{j} = 0;")
                pre.append(f"      if (!(FL({j}) & GJB_SITE_SAMPLE) && valid) in_s{j} = __ldg(reinterpret_cast<{cast}>(io.site_in[{j}]) + ((FL({j}) & GJB_SITE_BCAST) ? 0 : i));")
                post.append(f"      if (io.site_out[{j}] && valid && lane == 0) reinterpret_cast<{ct}*>(io.site_out[{j}])[i] = s{j};")
                kind = s.dist.rng_kind
                if kind != "lane":
                    pre.append(f"      uint4 W{j} = make_uint4(0u, 0u, 0u, 0u); (void)W{j};")
                    if kind == "normal":
                        pre.append(f"      float4 Z{j} = make_float4(0.f, 0.f, 0.f, 0.f); (void)Z{j};")
                    draw_if = "true" if s.cmask is not None else f"FL({j}) & GJB_SITE_SAMPLE"
                    pre.append(f"      if ({draw_if}) {{ W{j} = gjb::quad_words(key0, key1, (idx_offset + (uint64_t)i) >> 2, {j + 1}u);"
                               + (f" Z{j} = gjb::normal4_of(W{j});" if kind == "normal" else "") + " }")
            else:
                pre.append(f"      gjb::V4 in_s{j} = gjb::v4_splat(0.0f);")
                pre.append(f"      if (!(FL({j}) & GJB_SITE_SAMPLE) && valid) in_s{j} = gjb::v4_ldg(reinterpret_cast<const float*>(io.site_in[{j}]) + ((FL({j}) & GJB_SITE_BCAST) ? 0 : i * {D}) + 4 * lane);")
                post.append(f"      if (io.site_out[{j}] && valid) *reinterpret_cast<float4*>(reinterpret_cast<float*>(io.site_out[{j}]) + i * {D} + 4 * lane) = gjb::v4_to(s{j});")
        for k, r in enumerate(self.all_rets):
            if isinstance(r, Expr) and r.ndim == 1:
                post.append(f"      if (io.ret_out[{k}] && valid) *reinterpret_cast<float4*>(reinterpret_cast<float*>(io.ret_out[{k}]) + i * {D} + 4 * lane) = gjb::v4_to({self.ret_names[k]});")
            else:
                ct = "int" if isinstance(r, Expr) and r.dtype == I32 else "float"
                post.append(f"      if (io.ret_out[{k}] && valid && lane == 0) reinterpret_cast<{ct}*>(io.ret_out[{k}])[i] = {self.ret_names[k]};")

        out: list[str] = []
        out.append("// particles [p_begin, p_end): block-iteration base steps by p_stride; G lanes per particle")
        out.append("template <bool kCg, bool kSt, bool kSm = false>  // kSm: io.gather points to shared memory")
        out.append("__device__ __forceinline__ void run_groups(const Io& io, const Uni& U, const uint32_t (&fl)[NS], int64_t n,")
        out.append("    uint64_t idx_offset, uint32_t key0, uint32_t key1, int64_t p_begin, int64_t p_end, int64_t p_stride, float& run_max) {")
        out.append("  const bool need_score = !kSt && io.score_out != nullptr;")
        out.append("  const int lane = threadIdx.x & (G - 1);")
        out.append("  const int sub_p = threadIdx.x / G;")
        out.append("  for (int64_t base = p_begin; base < p_end; base += p_stride) {")
        out.append("    {")
        out.append("      const int64_t i = base + sub_p;")
        out.append("      const bool valid = i < p_end && i < n;")
        out.append("      const int64_t row = valid ? (io.gather ? (int64_t)(kSm ? io.gather[i] : gjb::ldx<kCg>(io.gather + i)) : i) : 0;")
        out.append("      (void)row;")
        out.append("      const int sub = (int)((idx_offset + (uint64_t)i) & 3); (void)sub;")
        out.append("      const gjb::Lane rng = gjb::make_lane(key0, key1, idx_offset + (uint64_t)i); (void)rng;")
        out.append("      float score = 0.0f, weight = 0.0f, vscore = 0.0f, vweight = 0.0f;")
        out.extend(pre)
        out.append(self.body)
        out.append("      if (need_score) score += gjb::group_sum<G>(vscore);")
        out.append("      weight += gjb::group_sum<G>(vweight);")
        out.extend(post)
        out.append("      if (!kSt && io.score_out && valid && lane == 0) io.score_out[i] = score;")
        out.append("      if (valid) {")
        out.append("        float t = weight;")
        out.append("        if (!kSt && io.weight_in) t = __ldg(io.weight_in + i) + t;")
        out.append("        if (!kSt && io.score_in) t = t - __ldg(io.score_in + i);")
        out.append("        if (io.weight_out && lane == 0) io.weight_out[i] = t;")
        out.append("        if (kSm && lane == 0) io.te_w[i] = t;")
        out.append("        run_max = fmaxf(run_max, t);")
        out.append("      }")
        out.append("    }")
        out.append("  }")
        out.append("}")
        return out

    # ------------------------------------------------------------ kernels
    def model_kernel(self, static: bool = False, mass: bool = False) -> list[str]:
        name = ("model_kernel_static_mass" if mass else "model_kernel_static") if static else "model_kernel"
        # lane-group (vector) models stream 16 bytes per lane per row: 4 CTAs per SM (64 registers) keep ~33 KB of loads in
        # flight per SM, what HBM3e needs; at the compiler's own 71 registers (3 CTAs) the scoring pass sat at 0.53 of peak
        lb = "kThreads, 4" if self.group else "kThreads"
        out = [f"__global__ void __launch_bounds__({lb}) {name}(const __grid_constant__ gjb_model_args A) {{"]
        if mass:
            out.append("  // the tile masses of the NEXT step's buffer are zeroed here (nobody reads them during this launch)")
            out.append("  for (int64_t i = blockIdx.x * (int64_t)kThreads + threadIdx.x; i < A.tile_mass_clear_n; i += (int64_t)gridDim.x * kThreads)")
            out.append("    if (A.tile_mass_clear) A.tile_mass_clear[i] = 0ull;")
        out.append("  if (A.link && A.wait_off) {  // multi-GPU: the peers' ancestor writes of the previous step have landed")
        out.append("    __shared__ uint64_t link_vals[GJB_MAX_RANKS];")
        out.append("    gjb::link_wait(A.link, A.wait_off, link_vals);")
        out.append("  }")
        out.extend(self.stage_lines("A.args"))
        out.append("  Uni U; make_uni(U, A.scalars);")
        out.append("  uint32_t fl[NS];")
        out.append(f"  for (int j = 0; j < {self.ns}; ++j) fl[j] = A.site_flags[j];")
        out.append("  Io io;")
        out.append("  for (int i = 0; i < NA; ++i) io.args[i] = A.args[i];")
        out.append("  for (int j = 0; j < NS; ++j) { io.site_in[j] = A.site_in[j]; io.site_out[j] = A.site_out[j]; }")
        out.append("  for (int k = 0; k < NR; ++k) io.ret_out[k] = A.ret_out[k];")
        out.append("  io.gather = A.gather; io.score_in = A.score_in; io.weight_in = A.weight_in; io.score_out = A.score_out; io.weight_out = A.weight_out;")
        out.append("  io.peers = A.peer_args;")
        out.append("  io.m_ref = A.m_ref; io.tile_mass = A.tile_mass;" if mass else "  io.m_ref = nullptr; io.tile_mass = nullptr;")
        out.append("  io.te_w = nullptr;")
        out.append("  const uint32_t key0 = A.key_dev ? __ldg(A.key_dev) : A.key0, key1 = A.key_dev ? __ldg(A.key_dev + 1) : A.key1;")
        out.append("  float run_max = -INFINITY;")
        if self.group:
            out.append(f"  run_groups<false, {'true' if static else 'false'}>(io, U, fl, A.n, A.idx_offset, key0, key1, (int64_t)blockIdx.x * kPPB, A.n, (int64_t)gridDim.x * kPPB, run_max);")
        else:
            out.append("  const int64_t nq = (A.n + (int64_t)(A.idx_offset & 3) + 3) >> 2;")
            targs = "false, true, true" if mass else f"false, {'true' if static else 'false'}"
            out.append(f"  run_quads<{targs}>(io, U, fl, A.n, A.idx_offset, key0, key1, blockIdx.x * (int64_t)kThreads + threadIdx.x, nq, (int64_t)gridDim.x * kThreads, run_max);")
        out.append("  if (A.wmax) gjb::block_wmax(run_max, A.wmax);")
        out.append("  if (A.link && A.push_off) {  // multi-GPU: the CTA that finishes last publishes this rank's max")
        out.append("    if (gjb::link_last_block(A.link)) gjb::link_push(A.link, A.push_off, (uint64_t)__ldcg(A.wmax));")
        out.append("  }")
        out.append("}")
        return out

    def pull_kernel(self) -> list[str]:
        """Single-pass filter step: pull-resample the previous step into this CTA's own slots, then gather + propose +
        logpdf + masses for exactly those slots (include/genjax_b200.h, gjb_model_args.pull_*)."""
        out = ["__global__ void __launch_bounds__(kThreads) model_kernel_static_pull(const __grid_constant__ gjb_model_args A) {"]
        out.append("  __shared__ gjb::TileSmem tsm;")
        out.append("  __shared__ __align__(16) int32_t heads[gjb::kWin];")
        out.append("  __shared__ uint64_t pre[gjb::kPullMaxTiles];")
        out.append("  for (int64_t i = blockIdx.x * (int64_t)kThreads + threadIdx.x; i < A.tile_mass_clear_n; i += (int64_t)gridDim.x * kThreads)")
        out.append("    if (A.tile_mass_clear) A.tile_mass_clear[i] = 0ull;")
        out.extend(self.stage_lines("A.args"))
        out.append("  Uni U; make_uni(U, A.scalars);")
        out.append("  uint32_t fl[NS];")
        out.append(f"  for (int j = 0; j < {self.ns}; ++j) fl[j] = A.site_flags[j];")
        out.append("  Io io;")
        out.append("  for (int i = 0; i < NA; ++i) io.args[i] = A.args[i];")
        out.append("  for (int j = 0; j < NS; ++j) { io.site_in[j] = A.site_in[j]; io.site_out[j] = A.site_out[j]; }")
        out.append("  for (int k = 0; k < NR; ++k) io.ret_out[k] = A.ret_out[k];")
        out.append("  io.gather = nullptr; io.score_in = nullptr; io.weight_in = nullptr; io.score_out = nullptr; io.weight_out = A.weight_out;")
        out.append("  io.peers = nullptr; io.m_ref = A.m_ref; io.tile_mass = A.tile_mass; io.te_w = nullptr;")
        out.append("  const int64_t w_lo = (int64_t)blockIdx.x * gjb::kTile;")
        out.append("  const int64_t w_n = (A.n - w_lo) < gjb::kTile ? (A.n - w_lo) : gjb::kTile;")
        out.append("  if (A.pull_logw) {  // ancestors of MY slots from the previous step's weights and tile masses")
        out.append("    const float Mp = __ldg(A.pull_m_ref);")
        out.append("    const double u0 = gjb::resample_u0(__ldg(A.pull_key), __ldg(A.pull_key + 1), (uint64_t)__ldg(A.pull_key + 2) | ((uint64_t)__ldg(A.pull_key + 3) << 32));")
        out.append("    const uint64_t S = gjb::pull_ancestors<false>(A.pull_logw, A.n, A.pull_tile_mass, (int)gridDim.x, Mp, A.pull_n_total, u0, w_lo, w_n, A.pull_ancestors + w_lo, tsm, heads, pre);")
        out.append("    if (blockIdx.x == 0 && threadIdx.x == 0 && A.pull_lse) {")
        out.append("      A.pull_lse[0] = (double)Mp; A.pull_lse[1] = (double)S;")
        out.append("      A.pull_lse[2] = S ? (double)Mp + log((double)S) - gjb::kQLog - log((double)A.pull_n_total) : -INFINITY;")
        out.append("    }")
        out.append("    __syncthreads();  // the CTA's own ancestor stores are visible to all of its threads (read back through L2)")
        out.append("    io.gather = A.pull_ancestors;")
        out.append("  }")
        out.append("  const uint32_t key0 = A.key_dev ? __ldg(A.key_dev) : A.key0, key1 = A.key_dev ? __ldg(A.key_dev + 1) : A.key1;")
        out.append("  float run_max = -INFINITY;")
        out.append("  run_quads<true, true, true>(io, U, fl, A.n, A.idx_offset, key0, key1, (w_lo >> 2) + threadIdx.x, (w_lo + w_n + 3) >> 2, kThreads, run_max);")
        out.append("}")
        return out

    def step_kernel(self) -> list[str]:
        """The single-launch filter step (include/genjax_b200.h ``gjb_model_pf_step``): one CTA per 2048 offspring
        slots -- te_pull (ancestors of the CTA's slots from the previous step's tile-exponent CDF) -> gather +
        propose + logpdf through the generated body -> te_publish (within-tile CDF + tile record of the new
        weights).  Block-level synchronisation only; csrc/gjb_step.cuh."""
        ir = self.ir
        out = ["// kDist: the launch carries a gjb_step_link (several GPUs, or the table forms on one device): peer loads, the prefix",
               "// tables and the cross-rank hand-off at the end of the CTA.  kDist = false is the single-device table-free step: none of",
               "// that code is in its kernel (two instantiations, chosen by the launcher), so it costs the hot path no registers.",
               "template <bool kDist>",
               "__global__ void __launch_bounds__(kThreads, 4) pf_step_kernel_t(const __grid_constant__ gjb_step_args A) {"]
        out.append("  __shared__ gjb::TeSmem sm;")
        out.extend(self.stage_lines("A.args"))
        out.append("  Uni U; make_uni(U, A.scalars);")
        out.append("  uint32_t fl[NS];")
        out.append(f"  for (int j = 0; j < {self.ns}; ++j) fl[j] = kPfFl[j];")
        out.append("  const int tid = threadIdx.x;")
        out.append("  const int64_t w_loc = (int64_t)blockIdx.x * gjb::kTeTile;  // first (local) slot of this CTA's window")
        out.append("  const int w_n = (A.n - w_loc) < gjb::kTeTile ? (int)(A.n - w_loc) : gjb::kTeTile;")
        out.append("  Io io;")
        out.append("  for (int i = 0; i < NA; ++i) io.args[i] = A.args[i];")
        out.append("  for (int j = 0; j < NS; ++j) { io.site_in[j] = A.site_in[j]; io.site_out[j] = nullptr; }")
        out.append("  for (int k = 0; k < NR; ++k) io.ret_out[k] = nullptr;")
        for k, r in enumerate(ir.ret_leaves):
            if isinstance(r, Expr) and r.op == "site" and r.attr not in (self.pf_obs or ()):
                out.append(f"  io.site_out[{r.attr}] = A.state_out[{k}];  // a sampled site that is the next state: written once")
            else:
                out.append(f"  io.ret_out[{k}] = A.state_out[{k}];")
        out.append("  io.gather = nullptr; io.score_in = nullptr; io.weight_in = nullptr; io.score_out = nullptr; io.weight_out = A.weight_out;")
        out.append("  io.peers = nullptr; io.m_ref = nullptr; io.tile_mass = nullptr;")
        wbuf = "reinterpret_cast<float*>(sm.pre)" if self.group else "reinterpret_cast<float*>(sm.heads)"
        out.append(f"  float* const wbuf = {wbuf};  // this window's new weights (block-shared)")
        out.append("  io.te_w = wbuf - w_loc;")
        hoist = (not self.group) and os.environ.get("GJB_STEP_HOIST", "1") != "0"
        out.append("  gjb::pdl_launch_dependents();  // the next launch of the stream may queue up behind this one right away")
        out.append("  GJB_TP(14); GJB_TP(0);")
        out.append("  const uint32_t key0 = __ldg(A.key_dev), key1 = __ldg(A.key_dev + 1);")
        if not self.group:
            out.append("  const int64_t q0 = (w_loc >> 2) + tid * 2;  // the thread's own 8 slots = 2 global quads")
            out.append("  const int64_t qw = (w_loc + w_n + 3) >> 2;")
        if hoist:
            out.append("  // their random numbers do not depend on the previous launch, so they are drawn BEFORE this launch waits for it")
            out.append("  // (programmatic dependent launch: this overlaps the previous kernel's tail)")
            out.append("  QRng R0, R1;")
            out.append("  quad_rng<true>(fl, key0, key1, (A.idx_offset >> 2) + (uint64_t)q0, R0);")
            out.append("  quad_rng<true>(fl, key0, key1, (A.idx_offset >> 2) + (uint64_t)q0 + 1, R1);")
        out.append("  GJB_TP(1);")
        out.append("  // constants of the link (not written by any launch): read before the wait, off the critical path")
        out.append("  int lk_world = 1, lk_tpr = 0, lk_rank = 0; const uint64_t* lk_box = nullptr;")
        out.append("  if (kDist && A.link && A.step > 0 && (A.flags & GJB_STEP_LIGHT)) {")
        out.append("    lk_world = A.link->world; lk_tpr = A.link->tiles_per_rank; lk_rank = A.link->rank;")
        out.append("    lk_box = gjb::te_mail_slot(A.link->mailbox[lk_rank], A.step - 1, 0);")
        out.append("  }")
        out.append("  // table form: the table kernel runs BESIDE the step kernels and flags its table with the step's tag; the CTAs spin on")
        out.append("  // that flag (te_pull_table) instead of waiting for a kernel boundary -- every record behind the table was mailed after")
        out.append("  // its tile's data was fenced into L2, so the flag covers the bulk data too.  Table-free form: wait for the previous launch.")
        out.append("  // (measured on B200: the flag hand-off is SLOWER than the kernel boundary, 30.7 vs 22.5 us per step on one device, so it is")
        out.append("  // opt-in: GJB_STEP_FLAGWAIT)")
        out.append("  const bool flagged = kDist && A.table_in && A.link && (A.flags & GJB_STEP_FLAGWAIT);")
        out.append("  if (!flagged) gjb::pdl_wait();")
        out.append("  GJB_TP(12);")
        out.append("  if (A.prev_cdf) {  // ancestors of MY slots: output-slot systematic resampling of the previous step")
        out.append("    const double u0 = gjb::resample_u0(__ldg(A.prev_key), __ldg(A.prev_key + 1), (uint64_t)__ldg(A.prev_key + 2) | ((uint64_t)__ldg(A.prev_key + 3) << 32));")
        out.append("    int32_t anc[gjb::kTeItems];")
        out.append("    int E;")
        out.append("    const int64_t w_glob = A.slot_offset + w_loc;")
        out.append("    if (kDist && A.table_in) {  // the previous launch's last CTA left the prefix table: no prefix work here")
        out.append("      // (peer memory is read through L2 (ld.global.cg); on one device the read-only path is used: griddepcontrol.wait")
        out.append("      // above makes the previous launch's writes visible to it, measured 1.4 us per step faster than .cg)")
        out.append("      if (A.cdf_peers && A.link && (A.flags & GJB_STEP_LIGHT))  // rank-level table: the tile prefix of the parents' rank(s) is formed here")
        out.append("        gjb::te_pull_light<true>(A.table_in, lk_box, lk_world, lk_tpr, lk_rank, A.cdf_peers, A.n_total, u0, w_glob, w_n, sm, anc, &E);")
        out.append("      else if (A.cdf_peers) gjb::te_pull_table<true>(A.table_in, (int)blockIdx.x, A.prev_cdf, A.cdf_peers, A.n_total, u0, w_glob, w_n, sm, anc, &E,")
        out.append("                                                flagged ? gjb::te_tag(A.link, A.step - 1) : 0u);")
        out.append("      else gjb::te_pull_table<false>(A.table_in, (int)blockIdx.x, A.prev_cdf, nullptr, A.n_total, u0, w_glob, w_n, sm, anc, &E,")
        out.append("                                     flagged ? gjb::te_tag(A.link, A.step - 1) : 0u);")
        out.append("    } else {           // single device, table-free: every CTA forms the tile prefix from the plain records")
        out.append("      uint64_t S;")
        out.append("      if (A.n_tiles_total <= 2 * kThreads) {")
        out.append("        const gjb::TeRecs2 recs2 = gjb::te_load_recs2<false>(A.prev_recs, A.n_tiles_total);")
        out.append("        S = gjb::te_pull<false, true>(A.prev_recs, A.n_tiles_total, A.prev_cdf, nullptr, A.n_total, u0, w_glob, w_n, sm, anc, &E, &recs2);")
        out.append("      } else {")
        out.append("        S = gjb::te_pull<false, false>(A.prev_recs, A.n_tiles_total, A.prev_cdf, nullptr, A.n_total, u0, w_glob, w_n, sm, anc, &E);")
        out.append("      }")
        out.append("      if (blockIdx.x == 0 && tid == 0 && A.prev_lse) gjb::te_write_lse(A.prev_lse, E, S, A.n_total);")
        out.append("    }")
        out.append("    const int4 a0 = make_int4(anc[0], anc[1], anc[2], anc[3]), a1 = make_int4(anc[4], anc[5], anc[6], anc[7]);")
        out.append("    *reinterpret_cast<int4*>(sm.heads + tid * gjb::kTeItems) = a0;  // (a thread's own slots: no hazard with the scan)")
        out.append("    *reinterpret_cast<int4*>(sm.heads + tid * gjb::kTeItems + 4) = a1;")
        out.append("    if (A.ancestors_out) {")
        out.append("      int32_t* o = A.ancestors_out + w_loc + tid * gjb::kTeItems;")
        out.append("      if (tid * gjb::kTeItems + gjb::kTeItems <= w_n && (reinterpret_cast<uintptr_t>(o) & 15) == 0) { reinterpret_cast<int4*>(o)[0] = a0; reinterpret_cast<int4*>(o)[1] = a1; }")
        out.append("      else for (int k = 0; k < gjb::kTeItems; ++k) if (tid * gjb::kTeItems + k < w_n) o[k] = anc[k];")
        out.append("    }")
        out.append("    io.gather = sm.heads - w_loc;")
        out.append("    io.peers = kDist ? A.peer_args : nullptr;")
        out.append("  }")
        out.append("  float run_max = -INFINITY;")
        if self.group:
            out.append("  __syncthreads();  // every group reads ancestors other threads resolved")
            out.append("  if (kDist && A.cdf_peers) run_groups<true, true, true>(io, U, fl, A.n, A.idx_offset, key0, key1, w_loc, w_loc + w_n, kPPB, run_max);")
            out.append("  else run_groups<false, true, true>(io, U, fl, A.n, A.idx_offset, key0, key1, w_loc, w_loc + w_n, kPPB, run_max);")
            out.append("  __syncthreads();  // the window's weights are complete")
        else:
            if hoist:
                out.append("  if (kDist && A.cdf_peers) {")
                out.append("    if (q0 < qw) run_quads<true, true, false, true, true>(io, U, fl, A.n, A.idx_offset, key0, key1, q0, q0 + 1, 1, run_max, &R0);")
                out.append("    if (q0 + 1 < qw) run_quads<true, true, false, true, true>(io, U, fl, A.n, A.idx_offset, key0, key1, q0 + 1, q0 + 2, 1, run_max, &R1);")
                out.append("  } else {")
                out.append("    if (q0 < qw) run_quads<false, true, false, true, true>(io, U, fl, A.n, A.idx_offset, key0, key1, q0, q0 + 1, 1, run_max, &R0);")
                out.append("    if (q0 + 1 < qw) run_quads<false, true, false, true, true>(io, U, fl, A.n, A.idx_offset, key0, key1, q0 + 1, q0 + 2, 1, run_max, &R1);")
                out.append("  }")
            else:
                out.append("  const int64_t qe = q0 + 2 < qw ? q0 + 2 : qw;")
                out.append("  if (kDist && A.cdf_peers) run_quads<true, true, false, true>(io, U, fl, A.n, A.idx_offset, key0, key1, q0, qe, 1, run_max);")
                out.append("  else run_quads<false, true, false, true>(io, U, fl, A.n, A.idx_offset, key0, key1, q0, qe, 1, run_max);")
        out.append("  GJB_TP(8);")
        out.append("  float lw[gjb::kTeItems];")
        out.append("#pragma unroll")
        out.append("  for (int k = 0; k < gjb::kTeItems; ++k) lw[k] = (tid * gjb::kTeItems + k < w_n) ? wbuf[tid * gjb::kTeItems + k] : -INFINITY;")
        if self.group:
            out.append("  __syncthreads();  // wbuf aliases sm.pre: everyone has its weights before te_publish reuses the scratch")
        out.append("  gjb::te_publish(lw, A.cdf_out + w_loc, A.recs_out ? A.recs_out + blockIdx.x : nullptr, sm);")
        out.append("  GJB_TP(11);")
        out.append("  if (kDist && A.link) gjb::te_finish_step(A.link, A.step, A.slot_offset, A.n, A.n_total, A.key_dev + 2, A.table_out, A.lse_out, sm, (A.flags & GJB_STEP_LIGHT) != 0);")
        out.append("  GJB_TP(15);")
        out.append("}")
        return out

    def steps_kernel(self) -> list[str]:
        """ALL T filter steps in ONE cooperative launch (include/genjax_b200.h ``gjb_model_pf_steps``; one device, every
        window's CTA co-resident): the body of ``pf_step_kernel`` (table-free form) in a loop over t, with ONE grid-wide
        barrier per step where the per-step form has a kernel boundary.  The random numbers of step t + 1 are drawn BEFORE
        that barrier, so they overlap the wait for the slowest CTA of step t."""
        ir = self.ir
        out = ["__global__ void __launch_bounds__(kThreads, 4) pf_steps_kernel(const __grid_constant__ gjb_steps_args Q) {"]
        out.append("  __shared__ gjb::TeSmem sm;")
        out.extend(self.stage_lines("Q.shared"))
        out.append("  Uni U; make_uni(U, Q.scalars);")
        out.append("  uint32_t fl[NS];")
        out.append(f"  for (int j = 0; j < {self.ns}; ++j) fl[j] = kPfFl[j];")
        out.append("  const int tid = threadIdx.x;")
        out.append("  const int64_t n = Q.n;")
        out.append("  const int64_t w_loc = (int64_t)blockIdx.x * gjb::kTeTile;")
        out.append("  const int w_n = (n - w_loc) < gjb::kTeTile ? (int)(n - w_loc) : gjb::kTeTile;")
        out.append("  const int n_tiles = (int)gridDim.x;")
        out.append("  const int64_t cdf_stride = (int64_t)n_tiles * gjb::kTeTile;")
        wbuf = "reinterpret_cast<float*>(sm.pre)" if self.group else "reinterpret_cast<float*>(sm.heads)"
        out.append(f"  float* const wbuf = {wbuf};")
        if not self.group:
            out.append("  const int64_t q0 = (w_loc >> 2) + tid * 2;")
            out.append("  const int64_t qw = (w_loc + w_n + 3) >> 2;")
        out.append("  for (int t = 0; t < Q.T; ++t) {")
        out.append("    const int slot = Q.record ? t : (t & 1);")
        out.append("    const int pslot = Q.record ? t - 1 : ((t - 1) & 1);")
        out.append("    const uint32_t* kt = Q.keys + 8 * t;")
        out.append("    const uint32_t key0 = __ldg(kt), key1 = __ldg(kt + 1);")
        if not self.group:
            out.append("    QRng R0, R1;  // this step's random numbers: independent of the previous step, drawn before its barrier")
            out.append("    quad_rng<true>(fl, key0, key1, (Q.idx_offset >> 2) + (uint64_t)q0, R0);")
            out.append("    quad_rng<true>(fl, key0, key1, (Q.idx_offset >> 2) + (uint64_t)q0 + 1, R1);")
        out.append("    if (t > 0) cooperative_groups::this_grid().sync();  // step t - 1: every CDF row, tile record and state row is written")
        out.append("    Io io;")
        out.append("    for (int i = 0; i < NA; ++i) io.args[i] = Q.shared[i];")
        out.append("    for (int j = 0; j < NS; ++j) { io.site_in[j] = nullptr; io.site_out[j] = nullptr; }")
        out.append("    for (int k = 0; k < NR; ++k) io.ret_out[k] = nullptr;")
        out.append("    io.gather = nullptr; io.score_in = nullptr; io.weight_in = nullptr; io.score_out = nullptr; io.peers = nullptr;")
        out.append("    io.m_ref = nullptr; io.tile_mass = nullptr;")
        for i in range(len(ir.ret_leaves)):
            out.append(f"    io.args[{i}] = t == 0 ? Q.state0[{i}] : (const void*)((const char*)Q.state_buf[{i}] + (int64_t)pslot * Q.state_stride[{i}]);")
        for sdef in ir.sites:
            j = sdef.index
            out.append(f"    if (Q.obs[{j}]) io.site_in[{j}] = (const char*)Q.obs[{j}] + (int64_t)t * Q.obs_stride[{j}];")
        for k, r in enumerate(ir.ret_leaves):
            dst = f"(void*)((char*)Q.state_buf[{k}] + (int64_t)slot * Q.state_stride[{k}])"
            if isinstance(r, Expr) and r.op == "site" and r.attr not in (self.pf_obs or ()):
                out.append(f"    io.site_out[{r.attr}] = {dst};")
            else:
                out.append(f"    io.ret_out[{k}] = {dst};")
        out.append("    io.weight_out = Q.record ? Q.logw + (int64_t)t * n : (t == Q.T - 1 ? Q.logw : nullptr);")
        out.append("    io.te_w = wbuf - w_loc;")
        out.append("    if (t > 0) {  // ancestors of MY slots: output-slot systematic resampling of step t - 1")
        out.append("      const uint32_t* kp = Q.keys + 8 * (t - 1) + 2;")
        out.append("      const double u0 = gjb::resample_u0(__ldg(kp), __ldg(kp + 1), (uint64_t)__ldg(kp + 2) | ((uint64_t)__ldg(kp + 3) << 32));")
        out.append("      const gjb_tile_rec* recs = Q.recs + (int64_t)((t - 1) & 1) * n_tiles;")
        out.append("      const uint64_t* cdf = Q.cdf + (int64_t)((t - 1) & 1) * cdf_stride;")
        out.append("      int32_t anc[gjb::kTeItems];")
        out.append("      int E;")
        out.append("      uint64_t S;")
        out.append("      if (n_tiles <= 2 * kThreads) {")
        out.append("        const gjb::TeRecs2 recs2 = gjb::te_load_recs2<true>(recs, n_tiles);")
        out.append("        S = gjb::te_pull<true, true>(recs, n_tiles, cdf, nullptr, n, u0, w_loc, w_n, sm, anc, &E, &recs2);")
        out.append("      } else {")
        out.append("        S = gjb::te_pull<true, false>(recs, n_tiles, cdf, nullptr, n, u0, w_loc, w_n, sm, anc, &E);")
        out.append("      }")
        out.append("      if (blockIdx.x == 0 && tid == 0) gjb::te_write_lse(Q.lse + 3 * (t - 1), E, S, n);")
        out.append("      const int4 a0 = make_int4(anc[0], anc[1], anc[2], anc[3]), a1 = make_int4(anc[4], anc[5], anc[6], anc[7]);")
        out.append("      *reinterpret_cast<int4*>(sm.heads + tid * gjb::kTeItems) = a0;")
        out.append("      *reinterpret_cast<int4*>(sm.heads + tid * gjb::kTeItems + 4) = a1;")
        out.append("      if (Q.record) {")
        out.append("        int32_t* o = Q.ancestors + (int64_t)(t - 1) * n + w_loc + tid * gjb::kTeItems;")
        out.append("        if (tid * gjb::kTeItems + gjb::kTeItems <= w_n && (reinterpret_cast<uintptr_t>(o) & 15) == 0) { reinterpret_cast<int4*>(o)[0] = a0; reinterpret_cast<int4*>(o)[1] = a1; }")
        out.append("        else for (int k = 0; k < gjb::kTeItems; ++k) if (tid * gjb::kTeItems + k < w_n) o[k] = anc[k];")
        out.append("      }")
        out.append("      io.gather = sm.heads - w_loc;")
        out.append("    }")
        out.append("    float run_max = -INFINITY;")
        if self.group:
            out.append("    __syncthreads();")
            out.append("    run_groups<true, true, true>(io, U, fl, n, Q.idx_offset, key0, key1, w_loc, w_loc + w_n, kPPB, run_max);")
            out.append("    __syncthreads();")
        else:
            out.append("    if (q0 < qw) run_quads<true, true, false, true, true>(io, U, fl, n, Q.idx_offset, key0, key1, q0, q0 + 1, 1, run_max, &R0);")
            out.append("    if (q0 + 1 < qw) run_quads<true, true, false, true, true>(io, U, fl, n, Q.idx_offset, key0, key1, q0 + 1, q0 + 2, 1, run_max, &R1);")
        out.append("    float lw[gjb::kTeItems];")
        out.append("#pragma unroll")
        out.append("    for (int k = 0; k < gjb::kTeItems; ++k) lw[k] = (tid * gjb::kTeItems + k < w_n) ? wbuf[tid * gjb::kTeItems + k] : -INFINITY;")
        if self.group:
            out.append("    __syncthreads();")
        out.append("    gjb::te_publish(lw, Q.cdf + (int64_t)(t & 1) * cdf_stride + w_loc, Q.recs + (int64_t)(t & 1) * n_tiles + blockIdx.x, sm);")
        out.append("    __syncthreads();  // the shared scratch is reused by the next step")
        out.append("  }")
        out.append("}")
        return out

    def pf_kernel(self) -> list[str]:
        """The persistent particle-filter kernel (None when the model's return
        leaves cannot feed back as its leading particle arguments)."""
        ir = self.ir
        out = ["__global__ void __launch_bounds__(kThreads, kPfMinBlocks) pf_kernel(const __grid_constant__ gjb_pf_args Q) {"]
        out.extend(self.stage_lines("Q.shared"))
        out.append("  Uni U; make_uni(U, Q.scalars);")
        out.append("  uint32_t fl[NS];")
        out.append(f"  for (int j = 0; j < {self.ns}; ++j) fl[j] = Q.site_flags[j];")
        out.append("  constexpr bool kSt = kPfStatic;")
        out.append("  __shared__ gjb::TileSmem tsm;")
        out.append("  extern __shared__ __align__(16) unsigned char dyn_smem[];")
        out.append("  uint64_t* qcache = reinterpret_cast<uint64_t*>(dyn_smem);  // masses of this CTA's tiles, phase B -> phase C")
        out.append("  const int tid = threadIdx.x;")
        out.append("  const int64_t n = Q.n;")
        out.append("  const int64_t G_ = gridDim.x;")
        out.append("  const int64_t chunk = (((n + G_ - 1) / G_) + 7) & ~(int64_t)7;  // particles per CTA, multiple of 8")
        out.append("  const int64_t c0 = (int64_t)blockIdx.x * chunk < n ? (int64_t)blockIdx.x * chunk : n;")
        out.append("  const int64_t c1 = c0 + chunk < n ? c0 + chunk : n;")
        out.append("  const bool cache_q = (chunk + gjb::kTile - 1) / gjb::kTile <= kPfCacheTiles;")
        out.append("  for (int t = 0; t < Q.T; ++t) {")
        out.append("    const int slot = Q.record ? t : (t & 1);")
        out.append("    const int pslot = Q.record ? t - 1 : ((t - 1) & 1);")
        out.append("    Io io;")
        out.append("    for (int i = 0; i < NA; ++i) io.args[i] = nullptr;")
        out.append("    for (int j = 0; j < NS; ++j) { io.site_in[j] = nullptr; io.site_out[j] = nullptr; }")
        out.append("    for (int k = 0; k < NR; ++k) io.ret_out[k] = nullptr;")
        out.append("    io.score_in = nullptr; io.weight_in = nullptr; io.score_out = nullptr; io.peers = nullptr; io.m_ref = nullptr; io.tile_mass = nullptr; io.te_w = nullptr;")
        for i in range(len(ir.ret_leaves)):
            out.append(f"    io.args[{i}] = t == 0 ? Q.state0[{i}] : (const void*)((const char*)Q.state_buf[{i}] + (int64_t)pslot * Q.state_stride[{i}]);")
        out.append("    io.gather = t == 0 ? nullptr : Q.ancestors + (int64_t)pslot * n;")
        for s in ir.sites:
            j = s.index
            out.append(f"    if (Q.obs[{j}]) io.site_in[{j}] = (const char*)Q.obs[{j}] + (int64_t)t * Q.obs_stride[{j}];")
        for k, r in enumerate(ir.ret_leaves):
            dst = f"(void*)((char*)Q.state_buf[{k}] + (int64_t)slot * Q.state_stride[{k}])"
            if isinstance(r, Expr) and r.op == "site":
                # a sampled site that is returned: write it once, as the next state
                out.append(f"    if (FL({r.attr}) & GJB_SITE_SAMPLE) io.site_out[{r.attr}] = {dst}; else io.ret_out[{k}] = {dst};")
            else:
                out.append(f"    io.ret_out[{k}] = {dst};")
        out.append("    float* lw = Q.logw + (Q.record ? (int64_t)t * n : 0);")
        out.append("    io.weight_out = lw;")
        out.append("    const uint32_t* kt = Q.keys + 8 * t;")
        out.append("    const uint32_t key0 = __ldg(kt), key1 = __ldg(kt + 1);")
        out.append("    // ---- phase A: gather + propose + logpdf over this CTA's particles, running max")
        out.append("    float run_max = -INFINITY;")
        if self.group:
            out.append("    run_groups<true, kPfStatic>(io, U, fl, n, Q.idx_offset, key0, key1, c0, c1, kPPB, run_max);")
        else:
            out.append("    run_quads<true, kPfStatic>(io, U, fl, n, Q.idx_offset, key0, key1, (c0 >> 2) + tid, (c1 + 3) >> 2, kThreads, run_max);")
        out.append("    uint32_t* wm = Q.wmax + (t & 1);")
        out.append("    gjb::block_wmax(run_max, wm);")
        out.append("    gjb::grid_barrier(Q.barrier, (uint32_t)G_);")
        out.append("    // ---- phase B: exact integer mass of this CTA's weights relative to the global max")
        out.append("    const float M = gjb::fdec(__ldcg(wm));")
        out.append("    uint64_t mass = 0;")
        out.append("    for (int64_t tb = c0; tb < c1; tb += gjb::kTile)")
        out.append("      mass += gjb::tile_mass_of<true>(lw, c1, tb, M, tsm.red, cache_q ? qcache + (tb - c0) : nullptr);")
        out.append("    if (tid == 0) __stcg(reinterpret_cast<unsigned long long*>(Q.cta_mass) + blockIdx.x, (unsigned long long)mass);")
        out.append("    gjb::grid_barrier(Q.barrier, (uint32_t)G_);")
        out.append("    // ---- phase C: CDF offset of this CTA, systematic offspring ranges -> ancestors")
        out.append("    uint64_t pre = 0, tot = 0;")
        out.append("    for (int b = tid; b < (int)G_; b += kThreads) {")
        out.append("      const uint64_t v = __ldcg(reinterpret_cast<const unsigned long long*>(Q.cta_mass) + b);")
        out.append("      tot += v; if (b < (int)blockIdx.x) pre += v;")
        out.append("    }")
        out.append("    pre = gjb::block_sum_u64(pre, tsm.red);")
        out.append("    tot = gjb::block_sum_u64(tot, tsm.red);")
        out.append("    const uint64_t S = tot;")
        out.append("    int32_t* anc = Q.ancestors + (int64_t)slot * n;")
        out.append("    if (blockIdx.x == 0 && tid == 0) {")
        out.append("      double* o = Q.lse + 3 * t;")
        out.append("      o[0] = (double)M; o[1] = (double)S;")
        out.append("      o[2] = S ? (double)M + log((double)S) - gjb::kQLog - log((double)Q.n_total) : -INFINITY;")
        out.append("      Q.wmax[(t + 1) & 1] = GJB_WMAX_NEG_INF;")
        out.append("    }")
        out.append("    if (S == 0) {")
        out.append("      for (int64_t i = c0 + tid; i < c1; i += kThreads) anc[i] = (int32_t)i;")
        out.append("    } else {")
        out.append("      const double u0 = gjb::resample_u0(__ldg(kt + 2), __ldg(kt + 3), (uint64_t)__ldg(kt + 4) | ((uint64_t)__ldg(kt + 5) << 32));")
        out.append("      uint64_t off = pre;")
        out.append("      for (int64_t tb = c0; tb < c1; tb += gjb::kTile)")
        out.append("        off += gjb::resample_tile<true>(lw, c1, tb, M, off, S, Q.n_total, u0, 0, n, 0, anc, tsm, reinterpret_cast<int32_t*>(qcache + (cache_q ? (tb - c0) : 0)), cache_q ? qcache + (tb - c0) : nullptr);")
        out.append("    }")
        out.append("    gjb::grid_barrier(Q.barrier, (uint32_t)G_);")
        out.append("  }")
        out.append("}")
        out.append("__global__ void pf_init_kernel(uint32_t* wmax, uint32_t* barrier) {")
        out.append("  if (threadIdx.x == 0) { wmax[0] = GJB_WMAX_NEG_INF; wmax[1] = GJB_WMAX_NEG_INF; barrier[0] = 0u; barrier[1] = 0u; }")
        out.append("}")
        return out

    def pf_supported(self) -> bool:
        ir = self.ir
        n_state = len(ir.ret_leaves)
        if n_state == 0 or n_state > len(ir.args):
            return False
        for r, a in zip(ir.ret_leaves, ir.args):
            if not isinstance(r, Expr) or a.kind != "particle":
                return False
            if tuple(r.shape) != tuple(a.shape) or r.dtype != a.dtype:
                return False
        return all(a.kind != "particle" for a in ir.args[n_state:])

    def source(self) -> str:
        out = self.header()
        out.extend(self.run_groups() if self.group else self.run_quads())
        out.extend(self.model_kernel())
        if self.pf_obs is not None:
            out.extend(self.model_kernel(static=True))
            if not self.group:
                out.extend(self.model_kernel(static=True, mass=True))
                out.extend(self.pull_kernel())
        pf = self.pf_supported()
        if pf:
            out.extend(self.pf_kernel())
        # (static shared memory: TeSmem is 45 KB of the 48 KB a kernel may declare; large staged argument blocks do not fit)
        staged = sum(max(1, int(np.prod(a.shape))) for a in self.ir.args if a.kind == "shared")
        self.has_step = pf and self.pf_obs is not None and staged <= 640
        if self.has_step:
            out.extend(self.step_kernel())
            out.extend(self.steps_kernel())
        chain_ext = None
        if self.chain is not None:
            from . import codegen_chain

            chain_ns, chain_ext = codegen_chain.generate_chain(self.ir, self.chain)
            out.append(chain_ns)
        out.append("}  // namespace")
        out.append(self.extern_c(pf, chain_ext))
        return "\n".join(out) + "\n"

    def extern_c(self, pf: bool, chain_ext: str | None = None) -> str:
        ir = self.ir
        mapping = "group" if self.group else "quad"
        info = _info_json(ir, mapping, max(self.G, 1), getattr(self, "has_step", False)).replace("\\", "\\\\").replace('"', '\\"')
        if self.group:
            work = "a->n"
            per_block = "kPPB"
        else:
            work = "(a->n + (int64_t)(a->idx_offset & 3) + 3) / 4"
            per_block = "kThreads"
        if pf:
            pf_code = f"""
static int pf_smem_for(int64_t n, int grid) {{
  const int64_t chunk = (((n + grid - 1) / grid) + 7) & ~(int64_t)7;
  const int64_t tiles = (chunk + gjb::kTile - 1) / gjb::kTile;
  return tiles <= kPfCacheTiles ? (int)(tiles * gjb::kTile * 8) : gjb::kTile * 8;  // >= one window of heads
}}

static int pf_grid_for(int64_t n) {{
  // sized with the largest dynamic shared memory the kernel may ask for
  const int resident = gjb::resident_blocks((const void*)pf_kernel, kThreads, 4, kPfCacheTiles * gjb::kTile * 8);
  int64_t want = (n + 255) / 256;  // at least 256 particles per CTA
  if (want < 1) want = 1;
  return (int)(want < resident ? want : resident);
}}

int gjb_model_pf_grid(int64_t n) {{ return n < 0 ? GJB_E_ARG : pf_grid_for(n); }}

int gjb_model_pf_run(const gjb_pf_args* a, void* stream) {{
  if (!a || a->n <= 0 || a->T <= 0 || !a->keys || !a->logw || !a->ancestors || !a->lse || !a->wmax || !a->cta_mass || !a->barrier)
    return GJB_E_ARG;
  if (a->n_state != {len(ir.ret_leaves)}) return GJB_E_ARG;
  if ((a->idx_offset & 3) != 0) return GJB_E_ARG;  // quads must align with the global quad streams
  if (a->n_total >= GJB_MASS_MAX_PARTICLES || a->n >= GJB_MASS_MAX_PARTICLES) return GJB_E_RANGE;
  for (int i = 0; i < a->n_state; ++i) if (!a->state0[i] || !a->state_buf[i]) return GJB_E_ARG;
  if (kPfStatic) for (int j = 0; j < {self.ns}; ++j) if (a->site_flags[j] != kPfFl_host[j]) return GJB_E_MODE;  // flags are baked in
  const int grid = pf_grid_for(a->n);
  const int smem = pf_smem_for(a->n, grid);
  static bool attr_set = false;
  if (!attr_set) {{
    cudaFuncSetAttribute((const void*)pf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kPfCacheTiles * gjb::kTile * 8);
    attr_set = true;
  }}
  pf_init_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(a->wmax, a->barrier);
  void* params[1] = {{(void*)a}};
  const cudaError_t e = cudaLaunchCooperativeKernel((const void*)pf_kernel, dim3(grid), dim3(kThreads), params, smem, (cudaStream_t)stream);
  if (e != cudaSuccess) return (int)e;
  return (int)cudaGetLastError();
}}
"""
        else:
            pf_code = """
int gjb_model_pf_grid(int64_t n) { (void)n; return GJB_E_MODE; }
int gjb_model_pf_run(const gjb_pf_args* a, void* stream) { (void)a; (void)stream; return GJB_E_MODE; }
"""
        if self.pf_obs is not None:
            static_dispatch = f"""// a launch whose flags are exactly the baked-in filter flags (and that asks for no score / weight
  // accumulation, and whose lanes start on a quad boundary) takes the specialised instantiation
  bool is_static = !a->score_in && !a->weight_in && !a->score_out && (a->idx_offset & 3) == 0;
  for (int j = 0; j < {self.ns}; ++j) is_static = is_static && a->site_flags[j] == kPfFl_host[j];
  if (a->tile_mass || a->m_ref) {{{{  // reference-maximum step: masses accumulated while the weights are in registers
    if (!is_static || !a->tile_mass || !a->m_ref || !a->weight_out || {'true' if self.group else 'false'}) return GJB_E_MODE;
    if (a->pull_ancestors) {{{{  // single-pass step: one CTA per 2048 offspring slots
      const int64_t tiles = (a->n + gjb::kTile - 1) / gjb::kTile;
      if (tiles > gjb::kPullMaxTiles || a->idx_offset != 0 || a->gather) return GJB_E_RANGE;
      if (a->pull_logw && (!a->pull_tile_mass || !a->pull_m_ref || !a->pull_key || a->pull_n_total <= 0 || a->pull_n_total >= GJB_MASS_MAX_PARTICLES)) return GJB_E_ARG;
      {'return GJB_E_MODE;' if self.group else 'model_kernel_static_pull<<<(int)tiles, kThreads, 0, (cudaStream_t)stream>>>(*a);'}
      return (int)cudaGetLastError();
    }}}}
    {'return GJB_E_MODE;' if self.group else 'model_kernel_static_mass<<<(int)blocks, kThreads, 0, (cudaStream_t)stream>>>(*a);'}
    return (int)cudaGetLastError();
  }}}}
  if (is_static) {{{{
    model_kernel_static<<<(int)blocks, kThreads, 0, (cudaStream_t)stream>>>(*a);
    return (int)cudaGetLastError();
  }}}}"""
        else:
            static_dispatch = "if (a->tile_mass || a->m_ref) return GJB_E_MODE;  // needs the filter-flag instantiation"
        if getattr(self, "has_step", False):
            step_code = """
static int pf_steps_capacity() {  // CTAs of pf_steps_kernel that are co-resident on this device
  return gjb::resident_blocks((const void*)pf_steps_kernel, kThreads, 8);
}
int gjb_model_pf_steps_fits(int64_t n) {
  if (n <= 0) return 0;
  return (n + gjb::kTeTile - 1) / gjb::kTeTile <= pf_steps_capacity() ? 1 : 0;
}
int gjb_model_pf_steps(const gjb_steps_args* a, void* stream) {
  if (!a || a->n <= 0 || a->T <= 0 || !a->keys || !a->cdf || !a->recs || !a->lse) return GJB_E_ARG;
  if ((a->idx_offset & 3) != 0) return GJB_E_ARG;
  if (a->n > (1LL << 26) || (a->n + gjb::kTeTile - 1) / gjb::kTeTile > gjb::kTeMaxTiles) return GJB_E_RANGE;
  if ((reinterpret_cast<uintptr_t>(a->cdf) & 15) || (reinterpret_cast<uintptr_t>(a->recs) & 15)) return GJB_E_ARG;
  if (a->record ? (!a->logw || !a->ancestors) : !a->logw) return GJB_E_ARG;
  for (int k = 0; k < kNState; ++k) if (!a->state0[k] || !a->state_buf[k]) return GJB_E_ARG;
  const int64_t tiles = (a->n + gjb::kTeTile - 1) / gjb::kTeTile;
  if (tiles > pf_steps_capacity()) return GJB_E_RANGE;  // every window's CTA must be resident (one grid barrier per step)
  void* params[1] = {(void*)a};
  const cudaError_t e = cudaLaunchCooperativeKernel((const void*)pf_steps_kernel, dim3((unsigned)tiles), dim3(kThreads), params, 0, (cudaStream_t)stream);
  if (e != cudaSuccess) return (int)e;
  return (int)cudaGetLastError();
}
#ifdef GJB_TRACE
int gjb_model_trace_read(unsigned long long* dst, int n) {  // scratch/trace_step.py
  cudaDeviceSynchronize();
  return (int)cudaMemcpyFromSymbol(dst, gjb_trace_buf, sizeof(unsigned long long) * (size_t)(n < 1024 * 16 ? n : 1024 * 16));
}
#endif
int gjb_model_pf_step(const gjb_step_args* a, void* stream) {
  if (!a || a->n <= 0 || a->n_total < a->n || !a->key_dev || !a->cdf_out) return GJB_E_ARG;
  if (!a->link && !a->recs_out) return GJB_E_ARG;  // table-free form needs the plain records (with link: table_out, or NULL = mail only)
  if (!a->link && (a->cdf_peers || a->peer_args || a->slot_offset != 0 || a->n_total != a->n)) return GJB_E_MODE;  // several devices need the table form
  if ((a->idx_offset & 3) != 0 || a->slot_offset < 0 || (a->slot_offset % gjb::kTeTile) != 0) return GJB_E_ARG;
  if (a->step < 0 || a->step >= 65535) return GJB_E_RANGE;  // 16-bit step field of the record tags
  if (reinterpret_cast<uintptr_t>(a->cdf_out) & 15) return GJB_E_ARG;
  for (int k = 0; k < kNState; ++k) if (!a->state_out[k]) return GJB_E_ARG;
  // S <= n_total * (2^36 + 1) must stay below 2^63 (signed conversion in offspring_cnt)
  if (a->n_total > (1LL << 26)) return GJB_E_RANGE;
  if ((a->n_total + gjb::kTeTile - 1) / gjb::kTeTile > gjb::kTeMaxTiles) return GJB_E_RANGE;
  if (a->prev_cdf) {
    if ((!a->table_in && !a->prev_recs) || !a->prev_key) return GJB_E_ARG;
    if (!a->table_in && (a->n_tiles_total <= 0 || a->n_tiles_total > gjb::kTeMaxTiles || (int64_t)a->n_tiles_total * gjb::kTeTile < a->n_total)) return GJB_E_ARG;
    if ((reinterpret_cast<uintptr_t>(a->prev_cdf) & 15) || (reinterpret_cast<uintptr_t>(a->prev_recs) & 15)) return GJB_E_ARG;
  }
  if (!a->link && (a->table_in || a->table_out || a->cdf_peers || a->peer_args)) return GJB_E_ARG;  // those need the link
  const int64_t tiles = (a->n + gjb::kTeTile - 1) / gjb::kTeTile;
  if (a->flags & GJB_STEP_PDL) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)tiles); cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = 0; cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return (int)cudaLaunchKernelEx(&cfg, a->link ? pf_step_kernel_t<true> : pf_step_kernel_t<false>, *a);
  }
  if (a->link) pf_step_kernel_t<true><<<(int)tiles, kThreads, 0, (cudaStream_t)stream>>>(*a);
  else pf_step_kernel_t<false><<<(int)tiles, kThreads, 0, (cudaStream_t)stream>>>(*a);
  return (int)cudaGetLastError();
}
"""
        else:
            step_code = """
int gjb_model_pf_step(const gjb_step_args* a, void* stream) { (void)a; (void)stream; return GJB_E_MODE; }
int gjb_model_pf_steps_fits(int64_t n) { (void)n; return 0; }
int gjb_model_pf_steps(const gjb_steps_args* a, void* stream) { (void)a; (void)stream; return GJB_E_MODE; }
"""
        chain_code = chain_ext if chain_ext is not None else """
int gjb_model_mh_chain(const gjb_chain_args* a, void* stream) { (void)a; (void)stream; return GJB_E_MODE; }
int gjb_model_hmc_chain(const gjb_chain_args* a, void* stream) { (void)a; (void)stream; return GJB_E_MODE; }
"""
        return f"""
extern "C" {{
const char* gjb_model_info(void) {{ return "{info}"; }}

int gjb_model_launch(const gjb_model_args* a, void* stream) {{
  if (!a || a->n < 0) return GJB_E_ARG;
  if (a->n == 0) return 0;
  const int64_t work = {work};
  int64_t blocks = (work + {per_block} - 1) / {per_block};
  const int64_t cap = gjb::resident_blocks((const void*)model_kernel, kThreads, 8);  // one full wave, grid-stride inside
  if (blocks > cap) blocks = cap;
  {static_dispatch}
  model_kernel<<<(int)blocks, kThreads, 0, (cudaStream_t)stream>>>(*a);
  return (int)cudaGetLastError();
}}
{pf_code}
{step_code}
{chain_code}
}}
"""

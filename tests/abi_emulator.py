"""TEST INFRASTRUCTURE ONLY -- a CPU emulation of the C-ABI entry points the single-device host code reaches:
``gjb_model_launch``, ``gjb_model_mh_chain`` / ``gjb_model_hmc_chain`` and the LSE / resampling / gather calls of
libgjb_core.

Purpose: exercise the HOST side (argument binding, per-site flags, pointer tables, choice-map / trace plumbing, the
``Scan`` combinator, ``get_subtrace`` ...) in the CPU suite, where no CUDA device exists.  The emulator reads the very
``gjb_model_args`` structure the host fills (through the raw pointers, like the kernel does), interprets the captured
model IR with the oracle's samplers / log-densities (oracle/dists.py, oracle/rng.py) and writes the outputs back
through the pointers.

What a green run proves: the host code fills the ABI consistently with the documented contract
(include/genjax_b200.h).  What it does NOT prove: anything about the CUDA kernels -- those are only ever checked on a
real B200 by the ``-m gpu`` tests.  Nothing under genjax_b200/ imports this file; it is installed by monkeypatching
inside tests (``install(monkeypatch)``), and the product keeps failing loudly without CUDA.
"""

from __future__ import annotations

import ctypes as C
import json
import os

import numpy as np
import torch

from oracle import dists as od

F32, I32 = np.float32, np.int32

_UN = {
    "neg": np.negative, "exp": np.exp, "log": np.log, "sqrt": np.sqrt, "abs": np.abs, "tanh": np.tanh,
    "sigmoid": lambda x: F32(1) / (F32(1) + np.exp(-x)), "log1p": np.log1p, "expm1": np.expm1, "square": np.square,
    "floor": np.floor, "sin": np.sin, "cos": np.cos, "softplus": lambda x: np.logaddexp(F32(0), x),
    "reciprocal": lambda x: F32(1) / x, "logical_not": lambda x: (x == 0).astype(I32),
}
_BIN = {
    "add": np.add, "sub": np.subtract, "mul": np.multiply, "div": np.divide, "pow": np.power, "min": np.minimum,
    "max": np.maximum, "lt": np.less, "le": np.less_equal, "gt": np.greater, "ge": np.greater_equal,
    "eq": np.equal, "ne": np.not_equal, "and": lambda a, b: (a != 0) & (b != 0), "or": lambda a, b: (a != 0) | (b != 0),
}


def _typed(v, dtype):
    return np.asarray(v).astype(F32 if dtype == "f32" else I32)


def evaluate(e, env, cache):
    """Vectorised float32 / int32 interpreter of the captured DAG; values carry a leading particle axis or broadcast."""
    if e._id in cache:
        return cache[e._id]
    ins = [evaluate(i, env, cache) for i in e.ins]
    op = e.op
    if e.shape != () and op not in ("row", "gather1", "elem", "sum"):
        # a scalar-shaped operand that carries the particle axis broadcasts against the event axis of the result
        ins = [v[..., None] if (i.shape == () and np.ndim(v) >= 1) else v for i, v in zip(e.ins, ins)]
    if op == "const":
        v = F32(e.attr) if e.dtype == "f32" else I32(e.attr)
    elif op == "constvec":
        v = _typed(e.attr, e.dtype)
    elif op == "site":
        v = env[("site", e.attr)]
    elif op == "arg":
        v = env[("arg", e.attr["index"])]
    elif op == "chain_step":
        v = env["step_size"]
    elif op == "lgamma":
        from scipy.special import gammaln

        v = gammaln(ins[0]).astype(F32)
    elif op in ("row", "gather1"):
        v = ins[0][np.asarray(ins[1]).astype(np.int64)]
    elif op == "elem":
        v = ins[0][..., int(e.attr)]
    elif op == "sum":
        v = np.sum(ins[0], axis=-1, dtype=F32)
    elif op == "cast":
        v = _typed(ins[0], e.dtype)
    elif op == "where":
        v = np.where(np.asarray(ins[0]) != 0, ins[1], ins[2])
    elif op in _UN:
        with np.errstate(all="ignore"):
            v = _UN[op](ins[0])
    elif op in _BIN:
        with np.errstate(all="ignore"):
            v = _BIN[op](ins[0], ins[1])
        if op in ("lt", "le", "gt", "ge", "eq", "ne", "and", "or"):
            v = np.asarray(v).astype(I32)
    else:
        raise NotImplementedError(f"emulator: op {op}")
    if isinstance(v, np.ndarray) and v.dtype == np.float64:
        v = v.astype(F32)
    cache[e._id] = v
    return v


def _accumulate_tile_mass(A, n, t):
    """Reference-maximum step (gjb_model_args.m_ref / tile_mass): exact integer masses of the weights relative to
    *m_ref, added per 2048-particle tile; the next step's buffer is zeroed."""
    if not (A.tile_mass or A.m_ref):
        return
    from oracle import smc as osmc

    if not (A.tile_mass and A.m_ref) or int(A.idx_offset) & 3:
        raise ValueError("emulator: m_ref and tile_mass come together, on a quad-aligned launch")
    if A.tile_mass_clear and A.tile_mass_clear_n > 0:
        _arr(A.tile_mass_clear, int(A.tile_mass_clear_n), C.c_uint64, np.uint64)[:] = 0
    mref = _view(A.m_ref, 1, F32)[0]
    with np.errstate(invalid="ignore"):
        q = osmc.det_exp_q((np.asarray(t, dtype=F32) - mref).astype(F32))
    tiles = (n + 2047) // 2048
    tm = _arr(A.tile_mass, tiles, C.c_uint64, np.uint64)
    for b in range(tiles):
        tm[b] += q[b * 2048:(b + 1) * 2048].sum(dtype=np.uint64)


def _view(ptr, count, dtype):
    if not ptr or count == 0:
        return None
    ct = C.c_float if dtype == F32 else C.c_int32
    return np.ctypeslib.as_array((ct * count).from_address(ptr))


def _prod(shape):
    out = 1
    for d in shape:
        out *= d
    return out


class EmulatedModelLib:
    """Stands in for ``ctypes.CDLL(model_<digest>.so)``: ``gjb_model_info`` and ``gjb_model_launch`` only."""

    def __init__(self, ir, chain=None):
        self.ir = ir
        self.chain = chain
        self.launches = 0

    def gjb_model_info(self):
        sites = [{"addr": list(s.addr), "dist": s.dist.name} for s in self.ir.sites]
        return json.dumps({"name": self.ir.name, "sites": sites, "emulated": True, "pf_step": True}).encode()

    def gjb_model_pf_steps_fits(self, n):
        return 1 if n <= 2048 else 0  # the SIMT host run of a cooperative kernel is a grid of ONE block

    def gjb_model_pf_steps(self, a_ref, stream):
        import simt_kernels
        from genjax_b200.gen import codegen

        source = getattr(self, "source", None) or codegen.generate(self.ir, getattr(self, "pf_obs", None), self.chain)
        lib = simt_kernels.model(source)
        self.launches += 1
        return lib.s_pf_steps(a_ref)

    def gjb_model_pf_step(self, a_ref, stream):
        """The single-launch filter step is a block-level kernel: always the GENERATED source with real block semantics
        (tests/simt_kernels.py), also when the other launches of this library interpret the IR."""
        import simt_kernels
        from genjax_b200.gen import codegen

        A = a_ref._obj
        if A.peer_args or A.cdf_peers:
            raise NotImplementedError("emulator: multi-GPU links are a GPU-only path")
        if int(A.n) > 60_000:
            raise NotImplementedError("emulator: the SIMT host run of pf_step_kernel is kept to small filters")
        source = getattr(self, "source", None) or codegen.generate(self.ir, getattr(self, "pf_obs", None), self.chain)
        lib = simt_kernels.model(source)
        if not hasattr(lib, "s_pf_step"):
            return -3
        self.launches += 1
        return lib.s_pf_step(a_ref)

    def gjb_model_launch(self, a_ref, stream):
        A = a_ref._obj
        ir = self.ir
        n = int(A.n)
        if n < 0:
            return -1
        if n == 0:
            return 0
        self.launches += 1
        if A.peer_args or A.link:
            raise NotImplementedError("emulator: multi-GPU links are a GPU-only path")
        gather_ptr = A.gather
        if A.pull_ancestors and A.pull_logw:  # single-pass step: output-slot resampling of the previous step first
            from oracle import rng as orng
            from oracle import smc as osmc

            lwp = _view(A.pull_logw, n, F32)
            mp = _view(A.pull_m_ref, 1, F32)[0]
            kd = _arr(A.pull_key, 4, C.c_uint32, np.uint32)
            key = orng.Key((int(kd[0]), int(kd[1])), int(kd[2]) | (int(kd[3]) << 32))
            anc = osmc.resample_systematic_pull(lwp, key, M=mp)
            _view(A.pull_ancestors, n, I32)[:] = anc
            if A.pull_lse:
                tiles = (n + 2047) // 2048
                S = int(_arr(A.pull_tile_mass, tiles, C.c_uint64, np.uint64).sum(dtype=np.uint64))
                out = _arr(A.pull_lse, 3, C.c_double, np.float64)
                out[0], out[1] = float(mp), float(S)
                out[2] = float(mp) + np.log(float(S)) - 36 * np.log(2.0) - np.log(float(A.pull_n_total)) if S else -np.inf
            gather_ptr = A.pull_ancestors
        words = (int(A.key0), int(A.key1))
        if A.key_dev:  # filter steps read their key words from the device key table
            kd = _arr(A.key_dev, 2, C.c_uint32, np.uint32)
            words = (int(kd[0]), int(kd[1]))
        idx = np.uint64(A.idx_offset) + np.arange(n, dtype=np.uint64)
        gather = _view(gather_ptr, n, I32)
        env = {}
        for i, spec in enumerate(ir.args):
            dt = F32 if spec.dtype == "f32" else I32
            if spec.kind == "scalar":
                env[("arg", i)] = dt(A.scalars[i])
            elif spec.kind == "shared":
                env[("arg", i)] = _view(A.args[i], max(_prod(spec.shape), 1), dt).reshape(spec.shape).copy()
            else:
                row = _prod(spec.shape)
                if gather is not None:  # rows of a LARGER source array, picked by ancestor index
                    rows = int(gather.max()) + 1
                    src = _view(A.args[i], rows * row, dt).reshape((rows,) + tuple(spec.shape))
                    env[("arg", i)] = src[gather].copy()
                else:
                    env[("arg", i)] = _view(A.args[i], n * row, dt).reshape((n,) + tuple(spec.shape)).copy()
        cache = {}
        score = np.zeros(n, dtype=F32)
        weight = np.zeros(n, dtype=F32)
        need_score = bool(A.score_out)
        for s in ir.sites:
            j = s.index
            fl = int(A.site_flags[j])
            ev = tuple(s.value.shape)
            dt = F32 if s.value.dtype == "f32" else I32
            args = [evaluate(e, env, cache) for e in s.args]
            if getattr(s.dist, "base", None) is not None:  # dist.repeat / dist.vmap: a vector site of iid scalar draws
                sample, logpdf = od.repeated(s.dist.base.name, ev[0])
            elif s.dist.name in od.DISTS:
                sample, logpdf = od.DISTS[s.dist.name]
            else:  # no oracle sampler: score through the symbolic log-density (gen/autodiff.py), refuse to sample
                sample, logpdf = None, None
            # dynamic structure (SiteSpec.live / scored / cmask): per-particle predicates
            cmw = None if getattr(s, "cmask", None) is None else np.broadcast_to(np.asarray(evaluate(s.cmask, env, cache)), (n,)).astype(np.int64)
            live = None if getattr(s, "live", None) is None else np.broadcast_to(np.asarray(evaluate(s.live, env, cache)), (n,)) != 0
            scored = None if getattr(s, "scored", None) is None else np.broadcast_to(np.asarray(evaluate(s.scored, env, cache)), (n,)) != 0

            def _draw():
                if sample is None:
                    raise NotImplementedError(f"emulator: no oracle sampler for {s.dist.name}")
                with np.errstate(all="ignore"):
                    d = np.asarray(sample(words, idx, j + 1, *args))
                return np.broadcast_to(d.astype(dt), (n,) + ev).copy()

            if fl & 1:
                v = _draw()
            else:
                if not A.site_in[j]:
                    return -1
                if fl & 4:
                    v = _view(A.site_in[j], max(_prod(ev), 1), dt).reshape(ev).copy()
                    v = np.broadcast_to(v, (n,) + ev)
                else:
                    v = _view(A.site_in[j], n * _prod(ev), dt).reshape((n,) + ev).copy()
                if cmw is not None:
                    take = (cmw & 1).astype(bool).reshape((n,) + (1,) * len(ev))
                    v = np.where(take, v, _draw()).astype(dt)
            if live is not None:
                v = np.where(live.reshape((n,) + (1,) * len(ev)), v, 0).astype(dt)
            gate = None
            for g_ in (live, scored):
                if g_ is not None:
                    gate = g_ if gate is None else (gate & g_)
            if gate is not None or cmw is not None:  # the general (masked) accumulation
                if need_score or (fl & 2):
                    vv = v.astype(bool) if getattr(s.dist, "bool_valued", False) else v
                    with np.errstate(all="ignore"):
                        if logpdf is None:
                            from genjax_b200.gen import autodiff as AD

                            env[("site", j)] = v
                            lp = np.broadcast_to(np.asarray(evaluate(AD.logpdf_expr(s.dist, s.value, s.args), env, {}), dtype=F32), (n,))
                        else:
                            lp = np.broadcast_to(np.asarray(logpdf(vv, *args), dtype=F32), (n,))
                    on = np.ones(n, dtype=bool) if gate is None else gate
                    score = np.where(on, (score + lp).astype(F32), score)
                    if fl & 2:
                        won = on if cmw is None else (on & ((cmw & 2) == 0))
                        weight = np.where(won, (weight + lp).astype(F32), weight)
            elif need_score or (fl & 2):
                vv = v.astype(bool) if getattr(s.dist, "bool_valued", False) else v
                if logpdf is None:
                    from genjax_b200.gen import autodiff as AD

                    env[("site", j)] = v
                    lp = np.broadcast_to(np.asarray(evaluate(AD.logpdf_expr(s.dist, s.value, s.args), env, {}), dtype=F32), (n,))
                else:
                    lp = np.broadcast_to(np.asarray(logpdf(vv, *args), dtype=F32), (n,))
                score = (score + lp).astype(F32)
                if fl & 2:
                    weight = (weight + lp).astype(F32)
            env[("site", j)] = v
            out = _view(A.site_out[j], n * _prod(ev), dt)
            if out is not None:
                out[:] = np.ascontiguousarray(v).reshape(-1)
        for k, r in enumerate(list(ir.ret_leaves) + list(getattr(ir, "flag_leaves", []))):
            if not A.ret_out[k]:
                continue
            val = evaluate(r, env, cache)
            dt = F32 if r.dtype == "f32" else I32
            full = np.broadcast_to(np.asarray(val).astype(dt), (n,) + tuple(r.shape))
            _view(A.ret_out[k], n * _prod(r.shape), dt)[:] = np.ascontiguousarray(full).reshape(-1)
        if A.score_out:
            _view(A.score_out, n, F32)[:] = score
        if A.weight_out:
            t = weight
            if A.weight_in:
                t = (_view(A.weight_in, n, F32) + t).astype(F32)
            if A.score_in:
                t = (t - _view(A.score_in, n, F32)).astype(F32)
            _view(A.weight_out, n, F32)[:] = t
            _accumulate_tile_mass(A, n, t)
            if A.wmax:  # running maximum of the weights in the order-preserving integer encoding (atomicMax)
                w = _arr(A.wmax, 1, C.c_uint32, np.uint32)
                with np.errstate(invalid="ignore"):
                    m = np.fmax.reduce(t)
                if not np.isnan(m):
                    w[0] = max(int(w[0]), _enc(m))
        elif A.wmax:
            raise NotImplementedError("emulator: wmax without weight_out")
        return 0


# ------------------------------------------------------------------ chain entry points (gjb_model_mh_chain / _hmc_chain)


def _chain_common(lib, A):
    """(n, layout, env builder) of a chain launch: per-chain arguments and the unselected sites are constants, the
    selected sites are columns of the state row (gen/codegen_chain.py Ctx / _layout)."""
    ir, spec = lib.ir, lib.chain
    n = int(A.n)
    offs, off = {}, 0
    for j in spec.latent:
        sv = ir.sites[j].value
        w = sv.shape[0] if sv.ndim else 1
        offs[j] = (off, w)
        off += w
    base = {"step_size": F32(A.step_size)}
    for i, a in enumerate(ir.args):
        dt = F32 if a.dtype == "f32" else I32
        if a.kind == "scalar":
            base[("arg", i)] = dt(A.scalars[i])
        elif a.kind == "shared":
            base[("arg", i)] = _view(A.args[i], max(_prod(a.shape), 1), dt).reshape(a.shape).copy()
        else:
            base[("arg", i)] = _view(A.args[i], n * _prod(a.shape), dt).reshape((n,) + tuple(a.shape)).copy()
    for s in ir.sites:
        j = s.index
        if j in spec.latent:
            continue
        ev = tuple(s.value.shape)
        dt = F32 if s.value.dtype == "f32" else I32
        if int(A.site_flags[j]) & 4:
            base[("site", j)] = np.broadcast_to(_view(A.site_in[j], max(_prod(ev), 1), dt).reshape(ev), (n,) + ev).copy()
        else:
            base[("site", j)] = _view(A.site_in[j], n * _prod(ev), dt).reshape((n,) + ev).copy()

    def env_of(q):
        env = dict(base)
        for j, (o, w) in offs.items():
            env[("site", j)] = q[:, o] if ir.sites[j].value.ndim == 0 else q[:, o:o + w]
        return env

    return n, offs, off, env_of


def _emulated_chain(lib, a_ref, kind):
    from genjax_b200.gen import autodiff as AD
    from genjax_b200.gen import expr as E
    from oracle import mcmc as omcmc
    from oracle import rng as orng

    A = a_ref._obj
    ir, spec = lib.ir, lib.chain
    if spec is None:
        return -3
    n, offs, dtot, env_of = _chain_common(lib, A)
    if int(A.state_width) != dtot:
        return -1
    logp_e = AD.model_logp(ir)
    lat_vals = [ir.sites[j].value for j in spec.latent]

    def full(v):
        return np.broadcast_to(np.asarray(v, dtype=F32), (n,)).astype(F32)

    def logp(q):
        return full(evaluate(logp_e, env_of(np.asarray(q, dtype=F32)), {}))

    state = _view(A.state, n * dtot, F32).reshape(n, dtot)
    key = orng.KeyBatch((int(A.key0), int(A.key1)), n, int(A.idx_offset))
    accept = not (int(A.flags) & 2)
    if kind == "mh":
        props = []
        for k, j in enumerate(spec.latent):
            mapping = spec.proposals[k] if k < len(spec.proposals) else None
            cur = ir.sites[j].value
            props.append((cur, E.Expr("chain_step", (), "f32", ())) if mapping is None else tuple(E.lift(x) for x in mapping(cur)))

        def proposal(q):
            env, cache = env_of(np.asarray(q, dtype=F32)), {}
            loc, scale = np.empty((n, dtot), dtype=F32), np.empty((n, dtot), dtype=F32)
            for (le, se), j in zip(props, spec.latent):
                o, w = offs[j]
                for dst, ex in ((loc, le), (scale, se)):
                    val = np.asarray(evaluate(ex, env, cache), dtype=F32)
                    if val.ndim >= 1 and val.shape[0] == n and ex.shape == ():
                        val = val[:, None]
                    dst[:, o:o + w] = np.broadcast_to(val, (n, w))
            return loc, scale

        q, lp, acc, alpha = omcmc.mh_chain(logp, state.copy(), key, int(A.n_steps), float(A.step_size), proposal, accept,
                                           int(A.step0), bwd_at_old=bool(A.compat_stale_grad))
    else:
        try:
            grads = AD.grad(logp_e, lat_vals)
        except AD.NotDifferentiable:
            return -3

        def logp_grad(q):
            env, cache = env_of(np.asarray(q, dtype=F32)), {}
            lp = full(evaluate(logp_e, env, cache))
            g = np.empty((n, dtot), dtype=F32)
            for ge, j in zip(grads, spec.latent):
                o, w = offs[j]
                val = np.asarray(evaluate(ge, env, cache), dtype=F32)
                if val.ndim >= 1 and val.shape[0] == n and ge.shape == ():
                    val = val[:, None]
                g[:, o:o + w] = np.broadcast_to(val, (n, w))
            return lp, g

        q, lp, acc, alpha = omcmc.hmc_chain(logp_grad, state.copy(), key, int(A.n_steps), float(A.step_size),
                                            int(A.n_leapfrog), bool(A.compat_stale_grad), accept, int(A.step0))
    state[:] = q
    _view(A.logp, n, F32)[:] = lp
    _view(A.accept_count, n, I32)[:] += acc.astype(I32)
    if A.alpha_out:
        _view(A.alpha_out, n, F32)[:] = alpha
    return 0


def _emulated_pf_run(lib, q_ref):
    """``gjb_model_pf_run``: the T-step bootstrap filter of the persistent kernel, as the same three phases per step
    (propose + weight + max, exact mass, systematic resampling) run through the emulated single-step entry points."""
    from genjax_b200.runtime import cabi

    Q = q_ref._obj
    ir = lib.ir
    n, T, rec = int(Q.n), int(Q.T), bool(Q.record)
    if n <= 0 or T <= 0 or int(Q.idx_offset) & 3:
        return -1
    keys = _arr(Q.keys, T * 8, C.c_uint32, np.uint32).reshape(T, 8)
    core = EmulatedCore()
    wmax = np.zeros(1, dtype=np.uint32)
    tiles = max(1, (n + _TILE - 1) // _TILE)
    tile_mass = np.zeros(tiles, dtype=np.uint64)
    n_state = int(Q.n_state)
    for t in range(T):
        slot, pslot = (t, t - 1) if rec else (t & 1, (t - 1) & 1)
        A = cabi.ModelArgs()
        A.n, A.idx_offset = n, int(Q.idx_offset)
        A.key0, A.key1 = int(keys[t, 0]), int(keys[t, 1])
        for i, a in enumerate(ir.args):
            if i < n_state:
                A.args[i] = Q.state0[i] if t == 0 else Q.state_buf[i] + pslot * int(Q.state_stride[i])
            elif a.kind == "scalar":
                A.scalars[i] = Q.scalars[i]
            else:
                A.args[i] = Q.shared[i]
        if t > 0:
            A.gather = Q.ancestors + pslot * n * 4
        for s_ in ir.sites:
            j = s_.index
            A.site_flags[j] = Q.site_flags[j]
            if Q.obs[j]:
                A.site_in[j] = Q.obs[j] + t * int(Q.obs_stride[j])
        for k, r in enumerate(ir.ret_leaves):
            buf = Q.state_buf[k] + slot * int(Q.state_stride[k])
            if r.op == "site" and (int(Q.site_flags[r.attr]) & 1):
                A.site_out[r.attr] = buf
            else:
                A.ret_out[k] = buf
        lw = Q.logw + (t * n * 4 if rec else 0)
        A.weight_out = lw
        wmax[0] = 0x007FFFFF
        A.wmax = wmax.ctypes.data
        rc = lib.gjb_model_launch(C.byref(A), 0)
        if rc:
            return rc
        core.gjb_weight_mass(lw, n, wmax.ctypes.data, None, tile_mass.ctypes.data, 0)
        R = cabi.ResampleArgs()
        R.logw, R.n, R.wmax, R.tile_mass = lw, n, wmax.ctypes.data, tile_mass.ctypes.data
        R.n_total, R.out_lo, R.out_n, R.anc_base = int(Q.n_total), 0, n, 0
        R.key0, R.key1 = int(keys[t, 2]), int(keys[t, 3])
        R.key_index = int(keys[t, 4]) | (int(keys[t, 5]) << 32)
        R.ancestors = Q.ancestors + slot * n * 4
        R.lse_out = Q.lse + t * 24
        core.gjb_resample_systematic(C.byref(R), 0)
    return 0


def _pf_run(lib, q_ref):
    """Small filters run the REAL persistent kernel as a grid of one block under the SIMT shim (when the generated
    source is at hand); larger ones use the per-step emulation above."""
    source = getattr(lib, "source", None)
    Q = q_ref._obj
    if source is not None and "pf_kernel(" in source and int(Q.n) <= 20000:
        import simt_kernels

        return simt_kernels.model(source).s_pf_run(q_ref)
    return _emulated_pf_run(lib, q_ref)


EmulatedModelLib.gjb_model_pf_grid = lambda self, n: -1 if n < 0 else 1
EmulatedModelLib.gjb_model_pf_run = lambda self, q_ref, stream: _pf_run(self, q_ref)
EmulatedModelLib.gjb_model_mh_chain = lambda self, a_ref, stream: _emulated_chain(self, a_ref, "mh")
EmulatedModelLib.gjb_model_hmc_chain = lambda self, a_ref, stream: _emulated_chain(self, a_ref, "hmc")


# ------------------------------------------------------------------ libgjb_core.so (single-device entry points)

_TILE = 2048


def _enc(f) -> int:
    u = int(np.asarray(f, dtype=F32).view(np.uint32))
    return (u ^ 0xFFFFFFFF) & 0xFFFFFFFF if u >> 31 else u ^ 0x80000000


def _dec(u: int):
    u = int(u)
    b = u ^ 0x80000000 if u >> 31 else (u ^ 0xFFFFFFFF) & 0xFFFFFFFF
    return np.asarray(b, dtype=np.uint32).view(F32)[()]


def _arr(ptr, count, ct, dtype):
    if not ptr or count == 0:
        return None
    return np.ctypeslib.as_array((ct * count).from_address(ptr)).view(dtype)


class EmulatedCore:
    """The model-independent entry points the single-device SMC drivers use (runtime/smc_ops.py), restated with
    oracle/smc.py.  Multi-GPU, fused cooperative and filter entry points are GPU-only and absent on purpose."""

    def gjb_abi_version(self):
        return 15

    def gjb_mass_resample_fits(self, n):
        return 0

    def gjb_pf_key_table(self, key0, key1, T, out, stream):
        from genjax_b200.core.key import PRNGKey, pf_key_table

        _arr(out, 8 * T, C.c_uint32, np.uint32)[:] = pf_key_table(PRNGKey((key0, key1), 0), T).reshape(-1)
        return 0

    def gjb_epoch_bump(self, epoch, stream):
        _arr(epoch, 1, C.c_uint64, np.uint64)[0] += np.uint64(1)
        return 0

    # the tile-exponent kernels (include/genjax_b200.h section 1c) always run as written (tests/simt_kernels.py)
    def gjb_te_masses(self, logw, n, cdf, recs, stream):
        import simt_kernels

        return simt_kernels.core().s_te_masses(C.c_void_p(logw), C.c_int64(n), C.c_void_p(cdf), C.c_void_p(recs))

    def gjb_te_table(self, a_ref, stream):
        import simt_kernels

        return simt_kernels.core().s_te_table(a_ref)

    def gjb_te_resample(self, a_ref, stream):
        import simt_kernels

        return simt_kernels.core().s_te_resample(a_ref)

    def gjb_wmax_reset(self, wmax, stream):
        _arr(wmax, 1, C.c_uint32, np.uint32)[0] = 0x007FFFFF
        return 0

    def gjb_weight_max(self, logw, n, wmax, stream):
        w = _arr(wmax, 1, C.c_uint32, np.uint32)
        if n > 0:
            x = _view(logw, n, F32)
            with np.errstate(invalid="ignore"):
                m = np.fmax.reduce(x)
            if not np.isnan(m):
                w[0] = max(int(w[0]), _enc(m))
        return 0

    @staticmethod
    def _max(wmax, m_global):
        if m_global:
            return _view(m_global, 1, F32)[0]
        return _dec(_arr(wmax, 1, C.c_uint32, np.uint32)[0])

    def gjb_weight_mass(self, logw, n, wmax, m_global, tile_mass, stream):
        from oracle import smc as osmc

        M = self._max(wmax, m_global)
        x = _view(logw, n, F32)
        with np.errstate(invalid="ignore"):
            q = osmc.det_exp_q((x - M).astype(F32))
        tiles = max(1, (n + _TILE - 1) // _TILE)
        out = _arr(tile_mass, tiles, C.c_uint64, np.uint64)
        for b in range(tiles):
            out[b] = q[b * _TILE:(b + 1) * _TILE].sum(dtype=np.uint64)
        return 0

    def gjb_lse_finalize(self, tile_mass, n, wmax, m_global, n_total, lse_out, stream):
        import math

        tiles = max(1, (n + _TILE - 1) // _TILE)
        S = int(_arr(tile_mass, tiles, C.c_uint64, np.uint64).sum(dtype=np.uint64))
        M = float(self._max(wmax, m_global))
        out = _arr(lse_out, 3, C.c_double, np.float64)
        out[0], out[1] = M, float(S)
        out[2] = M + math.log(S) - 36 * math.log(2.0) - math.log(n_total) if S else -math.inf
        return 0

    def gjb_resample_systematic(self, r_ref, stream):
        from oracle import rng as orng
        from oracle import smc as osmc

        R = r_ref._obj
        c_off = int(_arr(R.c_offset, 1, C.c_uint64, np.uint64)[0]) if R.c_offset else 0
        s_tot = int(_arr(R.s_total, 1, C.c_uint64, np.uint64)[0]) if R.s_total else None
        key0, key1, key_index = int(R.key0), int(R.key1), int(R.key_index)
        if R.key_dev:
            kd = _arr(R.key_dev, 4, C.c_uint32, np.uint32)
            key0, key1, key_index = int(kd[0]), int(kd[1]), int(kd[2]) | (int(kd[3]) << 32)
        n = int(R.n)
        logw = _view(R.logw, n, F32)
        M = self._max(R.wmax, R.m_global)
        tiles = max(1, (n + _TILE - 1) // _TILE)
        S = int(_arr(R.tile_mass, tiles, C.c_uint64, np.uint64).sum(dtype=np.uint64)) if s_tot is None else s_tot
        anc = _view(R.ancestors, int(R.out_n), I32)
        lo = int(R.out_lo)
        if S == 0:
            anc[:] = np.arange(lo, lo + int(R.out_n), dtype=I32) - lo + int(R.anc_base)
        else:
            u0 = osmc.resample_u0(orng.Key((key0, key1), key_index))
            cnt, _ = osmc.systematic_counts(logw, u0, n_out=int(R.n_total), M=M, S=S, c_offset=c_off)
            scale = np.float64(int(R.n_total)) / np.float64(S)
            start = 0 if c_off == 0 else int(np.clip(np.ceil(np.float64(c_off) * scale - np.float64(u0)), 0, int(R.n_total)))
            prev = np.concatenate([[start], cnt[:-1]])
            # a shard writes only the offspring of ITS parents that fall inside the window; other slots are untouched
            for i in np.nonzero(cnt > prev)[0]:
                a, b = max(int(prev[i]), lo), min(int(cnt[i]), lo + int(R.out_n))
                if b > a:
                    anc[a - lo:b - lo] = i + int(R.anc_base)
        if R.lse_out:
            self.gjb_lse_finalize(R.tile_mass, n, R.wmax, R.m_global, int(R.n_total), R.lse_out, stream)
        if R.wmax_next:
            _arr(R.wmax_next, 1, C.c_uint32, np.uint32)[0] = 0x007FFFFF
        return 0

    def gjb_resample_multinomial(self, logw, n, wmax, tile_mass, cdf, key0, key1, idx_offset, n_out, ancestors, stream):
        from oracle import rng as orng
        from oracle import smc as osmc

        x = _view(logw, n, F32)
        M = self._max(wmax, None)
        with np.errstate(invalid="ignore"):
            q = osmc.det_exp_q((x - M).astype(F32))
        Cq = np.cumsum(q, dtype=np.uint64)
        S = int(Cq[-1]) if n else 0
        anc = _view(ancestors, n_out, I32)
        if S == 0:
            anc[:] = np.arange(n_out, dtype=I32)
            return 0
        idx = np.arange(n_out, dtype=np.uint64) + np.uint64(idx_offset)
        w0, w1, _, _ = orng.site_words((key0, key1), idx, 0, 1)
        r = (w0.astype(np.uint64) << np.uint64(32)) | w1.astype(np.uint64)
        anc[:] = np.searchsorted(Cq, osmc._mulhi64(r, np.uint64(S)), side="right").astype(I32)
        return 0

    def gjb_resample_multinomial_keydev(self, logw, n, wmax, tile_mass, cdf, key_dev, idx_offset, n_out, ancestors, stream):
        kd = _arr(key_dev, 2, C.c_uint32, np.uint32)
        return self.gjb_resample_multinomial(logw, n, wmax, tile_mass, cdf, int(kd[0]), int(kd[1]), idx_offset, n_out, ancestors, stream)

    def gjb_philox_fill(self, key0, key1, idx_offset, site, chunk, n, out4, stream):
        from oracle import rng as orng

        idx = np.arange(n, dtype=np.uint64) + np.uint64(idx_offset)
        w = np.stack(orng.site_words((key0, key1), idx, site, chunk), axis=1).astype(np.uint32)
        _arr(out4, 4 * n, C.c_uint32, np.uint32)[:] = w.reshape(-1)
        return 0

    def gjb_normal_fill(self, key0, key1, idx_offset, site, n, d, out, stream):
        from oracle import rng as orng

        idx = np.arange(n, dtype=np.uint64) + np.uint64(idx_offset)
        _view(out, n * d, F32)[:] = orng.normal_vec((key0, key1), idx, site, d).reshape(-1)
        return 0

    def gjb_accept_mask(self, u, w, n, mask, stream):
        with np.errstate(divide="ignore", invalid="ignore"):
            _view(mask, n, I32)[:] = (np.log(_view(u, n, F32)).astype(F32) < _view(w, n, F32)).astype(I32)
        return 0

    def gjb_select_rows(self, mask, a, b, out, n, row_bytes, a_bcast, b_bcast, stream):
        words = row_bytes // 4
        m = _view(mask, n, I32).astype(bool)[:, None]
        av = _view(a, words if a_bcast else n * words, I32).reshape(-1, words)
        bv = _view(b, words if b_bcast else n * words, I32).reshape(-1, words)
        _view(out, n * words, I32).reshape(n, words)[:] = np.where(m, av, bv)
        return 0

    def gjb_weight_ess(self, logw, n, lse3, out, stream):
        M = F32(_arr(lse3, 3, C.c_double, np.float64)[0])
        with np.errstate(invalid="ignore"):
            w = np.exp((_view(logw, n, F32) - M).astype(F32)).astype(F32).astype(np.float64)
        s2 = float((w * w).sum())
        _arr(out, 1, C.c_double, np.float64)[0] = float(w.sum()) ** 2 / s2 if s2 > 0 else 0.0
        return 0

    def gjb_gather_rows(self, src, ancestors, dst, n_out, row_bytes, stream):
        words = row_bytes // 4
        anc = _view(ancestors, n_out, I32)
        rows = int(anc.max()) + 1 if n_out else 0
        s = _view(src, rows * words, I32).reshape(rows, words)
        _view(dst, n_out * words, I32).reshape(n_out, words)[:] = s[anc]
        return 0


class HostKernelModelLib(EmulatedModelLib):
    """Same stand-in, but the model launch and the chain launches execute the GENERATED CUDA source compiled for the
    host (tests/host_kernels.py) instead of interpreting the IR; the filter run reuses the per-step launch."""

    def __init__(self, ir, chain, hostlib):
        super().__init__(ir, chain)
        self.h = hostlib

    def gjb_model_launch(self, a_ref, stream):
        A = a_ref._obj
        n = int(A.n)
        if n < 0:
            return -1
        if n == 0:
            return 0
        if A.peer_args or A.link:
            raise NotImplementedError("host kernels: multi-GPU links are a GPU-only path")
        self.launches += 1
        wmax = A.wmax
        A.wmax = None  # the block-level max reduction needs a real thread block; redone below from the weights
        try:
            if A.pull_ancestors:  # single-pass step: a block-level kernel, run with real block semantics
                import simt_kernels

                if not (A.tile_mass and A.m_ref and A.weight_out) or A.gather or int(A.idx_offset):
                    return -3
                rc = simt_kernels.model(self.source).s_model_launch_pull(a_ref)
            elif A.tile_mass or A.m_ref:  # the launcher's dispatch: the filter-flag instantiation with masses
                if not hasattr(self.h, "host_model_launch_mass") or not (A.tile_mass and A.m_ref and A.weight_out):
                    return -3
                rc = self.h.host_model_launch_mass(a_ref)
            else:
                rc = self.h.host_model_launch(a_ref)
        finally:
            A.wmax = wmax
        if rc == 0 and wmax:
            if not A.weight_out:
                raise NotImplementedError("host kernels: wmax without weight_out")
            t = _view(A.weight_out, n, F32)
            w = _arr(wmax, 1, C.c_uint32, np.uint32)
            with np.errstate(invalid="ignore"):
                m = np.fmax.reduce(t)
            if not np.isnan(m):
                w[0] = max(int(w[0]), _enc(m))
        return rc

    def gjb_model_mh_chain(self, a_ref, stream):
        return self.h.host_mh_chain(a_ref) if hasattr(self.h, "host_mh_chain") else -3

    def gjb_model_hmc_chain(self, a_ref, stream):
        return self.h.host_hmc_chain(a_ref) if hasattr(self.h, "host_hmc_chain") else -3


class SimtKernelModelLib(EmulatedModelLib):
    """The generated source of a lane-group (vector-site) model run with real block semantics (tests/simt_kernels.py):
    slower than the thread-at-a-time host build, so only used where that one cannot run."""

    def __init__(self, ir, chain, source):
        super().__init__(ir, chain)
        import simt_kernels

        self.s = simt_kernels.model(source)
        from genjax_b200.gen import codegen

        self.ppb = 256 // max(codegen.group_lanes(ir.width), 1)

    def gjb_model_launch(self, a_ref, stream):
        A = a_ref._obj
        n = int(A.n)
        if n < 0:
            return -1
        if n == 0:
            return 0
        if A.peer_args or A.link:
            raise NotImplementedError("SIMT kernels: multi-GPU links are a GPU-only path")
        if A.tile_mass or A.m_ref:
            return -3
        self.launches += 1
        return self.s.s_model_launch(a_ref, C.c_int(max(1, min(4, (n + self.ppb - 1) // self.ppb))))

    def gjb_model_mh_chain(self, a_ref, stream):
        return self.s.s_mh_chain(a_ref) if hasattr(self.s, "s_mh_chain") else -3

    def gjb_model_hmc_chain(self, a_ref, stream):
        return self.s.s_hmc_chain(a_ref) if hasattr(self.s, "s_hmc_chain") else -3


class SimtCore(EmulatedCore):
    """libgjb_core's single-device entry points on their REAL kernels, run with block semantics on the CPU
    (tests/simt_kernels.py) for inputs small enough to stay quick; larger ones fall back to the oracle restatement."""

    LIMIT = 40_000

    def __init__(self):
        import simt_kernels

        self.k = simt_kernels.core()

    def gjb_weight_max(self, logw, n, wmax, stream):
        if n > self.LIMIT or n <= 0:
            return super().gjb_weight_max(logw, n, wmax, stream)
        return self.k.s_weight_max(C.c_void_p(logw), C.c_int64(n), C.c_void_p(wmax), C.c_int(max(1, min(4, (n + 255) // 256))))

    def gjb_weight_mass(self, logw, n, wmax, m_global, tile_mass, stream):
        if n > self.LIMIT or n <= 0:
            return super().gjb_weight_mass(logw, n, wmax, m_global, tile_mass, stream)
        return self.k.s_weight_mass(C.c_void_p(logw), C.c_int64(n), C.c_void_p(wmax), C.c_void_p(m_global), C.c_void_p(tile_mass))

    def gjb_lse_finalize(self, tile_mass, n, wmax, m_global, n_total, lse_out, stream):
        tiles = max(1, (n + _TILE - 1) // _TILE)
        return self.k.s_lse_finalize(C.c_void_p(tile_mass), C.c_int(tiles), C.c_void_p(wmax), C.c_void_p(m_global),
                                     C.c_int64(n_total), C.c_void_p(lse_out))

    def gjb_resample_systematic(self, r_ref, stream):
        R = r_ref._obj
        if int(R.n) > self.LIMIT or int(R.n) <= 0 or int(R.out_n) <= 0:
            return super().gjb_resample_systematic(r_ref, stream)
        return self.k.s_resample_systematic(r_ref)

    def gjb_resample_multinomial(self, logw, n, wmax, tile_mass, cdf, key0, key1, idx_offset, n_out, ancestors, stream):
        if n > self.LIMIT or n <= 0 or n_out <= 0:
            return super().gjb_resample_multinomial(logw, n, wmax, tile_mass, cdf, key0, key1, idx_offset, n_out, ancestors, stream)
        return self.k.s_multinomial(C.c_void_p(logw), C.c_int64(n), C.c_void_p(wmax), C.c_void_p(tile_mass), C.c_void_p(cdf),
                                    C.c_uint32(key0), C.c_uint32(key1), C.c_uint64(idx_offset), C.c_int64(n_out), C.c_void_p(ancestors))

    def gjb_philox_fill(self, key0, key1, idx_offset, site, chunk, n, out4, stream):
        if n > self.LIMIT or n <= 0:
            return super().gjb_philox_fill(key0, key1, idx_offset, site, chunk, n, out4, stream)
        return self.k.s_philox_fill(C.c_uint32(key0), C.c_uint32(key1), C.c_uint64(idx_offset), C.c_uint32(site), C.c_uint32(chunk),
                                    C.c_int64(n), C.c_void_p(out4))

    def gjb_normal_fill(self, key0, key1, idx_offset, site, n, d, out, stream):
        if n * d > 4 * self.LIMIT or n <= 0:
            return super().gjb_normal_fill(key0, key1, idx_offset, site, n, d, out, stream)
        return self.k.s_normal_fill(C.c_uint32(key0), C.c_uint32(key1), C.c_uint64(idx_offset), C.c_uint32(site), C.c_int64(n),
                                    C.c_int(d), C.c_void_p(out))

    def gjb_gather_rows(self, src, ancestors, dst, n_out, row_bytes, stream):
        if n_out > self.LIMIT or n_out <= 0:
            return super().gjb_gather_rows(src, ancestors, dst, n_out, row_bytes, stream)
        return self.k.s_gather_rows(C.c_void_p(src), C.c_void_p(ancestors), C.c_void_p(dst), C.c_int64(n_out), C.c_int(row_bytes // 4),
                                    C.c_int(max(1, min(4, (n_out * (row_bytes // 4) + 255) // 256))))


class _PrebuildPool:
    """Background nvcc jobs of the prebuild pass (joined at interpreter exit)."""

    def __init__(self):
        self._ex = None
        self._futs = []

    def submit(self, fn, *a):
        import atexit
        from concurrent.futures import ThreadPoolExecutor

        if self._ex is None:
            self._ex = ThreadPoolExecutor(max_workers=int(os.environ.get("GJB_PREBUILD_JOBS", "4")))
            atexit.register(self.join)
        self._futs.append(self._ex.submit(fn, *a))

    def join(self):
        for f in self._futs:
            try:
                f.result()
            except Exception as e:  # a model that does not compile must fail the GPU test, not the prebuild pass
                print(f"[prebuild] {type(e).__name__}: {str(e)[-400:]}")
        self._futs.clear()


_PREBUILD_POOL = _PrebuildPool()


class _EmulatedCompiledModel:
    def __init__(self, ir, chain=None, pf_obs=None, host_kernels=False):
        self.ir = ir
        self.lib = None
        if os.environ.get("GJB_PREBUILD") == "1":
            # scripts/prebuild_test_models.py: also cross-compile the sm_100a library of every model the GPU tests
            # will ask for, so the GPU box loads genjax_b200/_lib/model_*.so instead of spending its minutes in nvcc
            from genjax_b200.gen import codegen as _cg
            from genjax_b200.runtime import build as _gb

            _PREBUILD_POOL.submit(_gb.build_model, _cg.generate(ir, pf_obs, chain))
        if host_kernels:
            import host_kernels as hk
            from genjax_b200.gen import codegen

            source = codegen.generate(ir, pf_obs, chain)
            if hk.is_host_runnable(source):
                self.lib = HostKernelModelLib(ir, chain, hk.build(source))
            else:
                self.lib = SimtKernelModelLib(ir, chain, source)
            self.lib.source = source
        if self.lib is None:
            self.lib = EmulatedModelLib(ir, chain)
        self.lib.pf_obs = pf_obs
        self.path = None
        self.info = json.loads(self.lib.gjb_model_info().decode())


def install(monkeypatch, host_kernels=None):
    """Route the host through the emulator for the duration of one test.  ``host_kernels=True`` (or
    GJB_EMULATE_KERNELS=host) executes the generated CUDA source compiled for the host where that is possible
    (tests/host_kernels.py) instead of interpreting the IR."""
    import os

    if host_kernels is None:
        host_kernels = os.environ.get("GJB_EMULATE_KERNELS") == "host"
    from genjax_b200.gen import capture as cap
    from genjax_b200.gen import static
    from genjax_b200.runtime import cabi

    cpu = torch.device("cpu")
    monkeypatch.setattr(cabi, "require_cuda", lambda: cpu)
    monkeypatch.setattr(cabi, "stream_ptr", lambda device=None: 0)
    monkeypatch.setattr(cabi, "ptr", lambda t: None if t is None else t.data_ptr())

    def compile_ir(ir, pf_obs=None, chain=None):
        ir.digest = cap.ir_fingerprint(ir)
        return _EmulatedCompiledModel(ir, chain, pf_obs, host_kernels)

    monkeypatch.setattr(static, "compile_ir", compile_ir)

    # module-level @gen functions (genjax_b200.workloads) keep their compiled models in a per-object cache that other
    # tests may have filled with REAL libraries: under emulation every generative function uses a cache of its own
    cache_name = "_emu_cache_host" if host_kernels else "_emu_cache_ir"

    def with_private_cache(method):
        def wrapped(self, *args, **kwargs):
            saved = self._cache
            self._cache = self.__dict__.setdefault(cache_name, {})
            try:
                return method(self, *args, **kwargs)
            finally:
                self._cache = saved

        return wrapped

    for name in ("compiled_for", "prebuild"):
        monkeypatch.setattr(static.StaticGenerativeFunction, name, with_private_cache(getattr(static.StaticGenerativeFunction, name)))
    from genjax_b200.inference import mcmc

    monkeypatch.setattr(mcmc, "compile_ir", compile_ir)
    core = SimtCore() if host_kernels else EmulatedCore()
    monkeypatch.setattr(cabi, "core", lambda: core)
    # CUDA graphs are a device feature: under emulation the filter enqueues its launches eagerly on every run
    from genjax_b200.inference import pf

    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
    eager = pf._Plan.execute
    monkeypatch.setattr(pf._Plan, "execute", lambda self, key, state0, shared, obs, use_graph: eager(self, key, state0, shared, obs, False))
    return cpu

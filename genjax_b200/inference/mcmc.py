"""Batched MCMC on fused chain kernels: Metropolis-Hastings with
``Rejuvenate``-style proposals and Hamiltonian Monte Carlo.

API mirror of the reference's edit requests
  * ``Rejuvenate(proposal, argument_mapping)``  (inference/requests/rejuvenate.py:45-94)
  * ``HMC(selection, eps, L)`` / ``SafeHMC``     (inference/requests/hmc.py:139-223)
whose ``.edit(key, trace, argdiffs)`` returns ``(new_trace, weight, retdiff,
bwd_request)`` WITHOUT accepting, plus the accept idiom the reference leaves to
user code (tests/inference/test_requests.py:136-137, 190-191):

    check = log(uniform(key)) < w ;  tr = where(check, new_tr, tr)

``mh_chain`` / ``hmc_chain`` run ``n_steps`` propose-weight-accept transitions
per launch with the chain state in registers (``gjb_model_mh_chain`` /
``gjb_model_hmc_chain``, gen/codegen_chain.py): the reference's Python loop of
``request.edit`` + accept, one chain per lane of the batched trace.
"""

from __future__ import annotations

import ctypes as C

import torch

from ..core.choice_map import ChoiceMap, Selection, _norm_addr
from ..core.key import lanes_of
from ..gen.codegen_chain import ChainSpec
from ..gen.gfi import Diff, EditRequest, Update
from ..gen.static import StaticTrace, _rebatch, compile_ir
from ..runtime import cabi


def _chain_model(trace: StaticTrace, latent: tuple, proposals: tuple):
    """The model variant carrying chain kernels for this (state, proposal) choice."""
    gf = trace.gen_fn
    cache = gf.__dict__.setdefault("_chain_cache", {})
    key = (id(trace.cm), latent, tuple(None if p is None else id(p) for p in proposals))
    cm = cache.get(key)
    if cm is None:
        ir = trace.cm.ir
        cm = compile_ir(ir, chain=ChainSpec(latent, proposals))
        cache[key] = cm
    return cm


def _selected_sites(trace: StaticTrace, selection: Selection) -> tuple:
    return tuple(s.index for s in trace.cm.ir.sites if selection(s.sel_addr).check())


class ChainResult:
    def __init__(self, trace: StaticTrace, accept_count: torch.Tensor, alpha: torch.Tensor, n_steps: int):
        self.trace = trace
        self.accept_count = accept_count
        self.alpha = alpha
        self.n_steps = n_steps

    @property
    def accept_rate(self) -> torch.Tensor:
        return self.accept_count.float().mean() / max(self.n_steps, 1)


def _run_chain(kind: str, key, trace: StaticTrace, latent: tuple, proposals: tuple, *, n_steps: int, step_size: float,
               n_leapfrog: int = 0, compat_stale_grad: bool = False, accept: bool = True, step0: int = 0,
               rebuild_trace: bool = True) -> ChainResult:
    if not latent:
        raise ValueError("the selection matches no random choice of the model")
    device = cabi.require_cuda()
    ir = trace.cm.ir
    cm = _chain_model(trace, latent, proposals)
    n = trace.n
    words, lane0, _ = lanes_of(key)

    # state row = selected sites in program order
    cols = []
    for j in latent:
        v = trace.values[j]
        if trace.bcast[j]:
            v = v.reshape((1,) + tuple(ir.sites[j].value.shape)).expand((n,) + tuple(ir.sites[j].value.shape))
        cols.append(v.reshape(n, -1).to(torch.float32))
    state = torch.cat(cols, dim=1).contiguous()
    logp = trace.score.to(torch.float32).clone().contiguous()
    acc = torch.zeros(n, dtype=torch.int32, device=device)
    alpha = torch.empty(n, dtype=torch.float32, device=device)

    A = cabi.ChainArgs()
    A.n = n
    A.idx_offset = lane0
    A.key0, A.key1 = words
    keep = []
    bound = trace.bound
    if bound is None:
        raise ValueError("this trace was produced by take(); re-create it with generate/update before running chains")
    for i, (spec, pl) in enumerate(zip(bound.specs, bound.payload)):
        if spec.kind == "scalar":
            A.scalars[i] = pl
        else:
            A.args[i] = pl.data_ptr()
            keep.append(pl)
    for s in ir.sites:
        j = s.index
        if j in latent:
            continue
        t = trace.values[j]
        A.site_in[j] = t.data_ptr()
        A.site_flags[j] = cabi.SITE_BCAST if trace.bcast[j] else 0
    A.state = state.data_ptr()
    A.logp = logp.data_ptr()
    A.accept_count = acc.data_ptr()
    A.alpha_out = alpha.data_ptr()
    A.state_width = state.shape[1]
    # the kernel re-evaluates logp(state) with its own expression order on entry (no GJB_CHAIN_HAVE_LOGP),
    # so alpha never mixes two floating-point formulas of the same density
    A.flags = 0 if accept else cabi.CHAIN_NO_ACCEPT
    A.n_steps = int(n_steps)
    A.step0 = int(step0)
    A.step_size = float(step_size)
    A.n_leapfrog = int(n_leapfrog)
    A.compat_stale_grad = int(bool(compat_stale_grad))
    fn = cm.lib.gjb_model_mh_chain if kind == "mh" else cm.lib.gjb_model_hmc_chain
    cabi.check(fn(C.byref(A), cabi.stream_ptr(device)), f"gjb_model_{kind}_chain")

    if not rebuild_trace:
        return ChainResult(_StateOnly(state, logp, latent, ir), acc, alpha, n_steps)
    # new trace: every site constrained (moved latents from the state row, the rest unchanged);
    # one assess-mode launch of the model kernel recomputes score and return value
    chm = _rebatch(trace)
    off = 0
    for j in latent:
        s = ir.sites[j]
        w = s.value.shape[0] if s.value.ndim else 1
        v = state[:, off:off + w].contiguous()
        if s.value.ndim == 0:
            v = v.reshape(n)
        from ..gen.static import Batched

        chm = ChoiceMap.entry(Batched(v), *s.addr) | chm
        off += w
    new_tr, _ = trace.gen_fn._run(None, trace.args, chm, weight_mode="none", n=n, batched=trace.batched)
    return ChainResult(new_tr, acc, alpha, n_steps)


class _StateOnly:
    """Final chain state without rebuilding a trace (benchmark path)."""

    def __init__(self, state, logp, latent, ir):
        self.state, self.logp, self.latent, self.ir = state, logp, latent, ir


# ---------------------------------------------------------------- chain drivers


def mh_chain(key, trace: StaticTrace, selection: Selection, *, step_size: float = 1.0, n_steps: int = 1,
             proposals: dict | None = None, step0: int = 0, rebuild_trace: bool = True) -> ChainResult:
    """``n_steps`` Metropolis-Hastings transitions per chain: a normal random walk
    of scale ``step_size`` on every selected site (``Rejuvenate(normal, lambda chm:
    (chm.get_value(), step_size))``), or ``proposals[addr] = mapping`` with
    ``mapping(cur) -> (loc, scale)`` for a custom normal proposal."""
    latent = _selected_sites(trace, selection)
    ir = trace.cm.ir
    props = []
    proposals = {_norm_addr(k): v for k, v in (proposals or {}).items()}
    for j in latent:
        props.append(proposals.get(ir.sites[j].addr))
    return _run_chain("mh", key, trace, latent, tuple(props), n_steps=n_steps, step_size=step_size, step0=step0,
                      rebuild_trace=rebuild_trace)


def hmc_chain(key, trace: StaticTrace, selection: Selection, *, eps: float, L: int = 10, n_iters: int = 1,
              compat_stale_grad: bool = False, step0: int = 0, rebuild_trace: bool = True) -> ChainResult:
    """``n_iters`` x (``HMC(selection, eps, L).edit`` + accept) per chain.
    ``compat_stale_grad=True`` integrates exactly like the reference (hmc.py:186);
    the default is the textbook leapfrog."""
    latent = _selected_sites(trace, selection)
    return _run_chain("hmc", key, trace, latent, (), n_steps=n_iters, step_size=eps, n_leapfrog=L,
                      compat_stale_grad=compat_stale_grad, step0=step0, rebuild_trace=rebuild_trace)


def mh_accept(key, new_trace: StaticTrace, old_trace: StaticTrace, weight: torch.Tensor) -> tuple:
    """The accept idiom of tests/inference/test_requests.py:136-137:
    ``check = log(uniform(key)) < w; tr = where(check, new, old)`` per chain."""
    from ..gen.distributions import uniform
    from ..gen.static import Batched

    from ..runtime import smc_ops

    u = uniform.sample(key, 0.0, 1.0) if old_trace.batched else uniform.sample(key, 0.0, 1.0).reshape(1)
    mask = smc_ops.accept_mask(u, weight.reshape(-1).to(torch.float32))  # log(u) < w, one small kernel
    check = mask.bool()
    ir = old_trace.cm.ir
    chm = ChoiceMap.empty()
    n = old_trace.n
    for s in ir.sites:
        j = s.index
        a, b = new_trace.values[j], old_trace.values[j]
        if new_trace.bcast[j] and old_trace.bcast[j]:
            chm = chm | ChoiceMap.entry(a, *s.addr)
            continue
        ev = tuple(s.value.shape)
        chm = chm | ChoiceMap.entry(Batched(smc_ops.select_rows(mask, a, b, n, ev)), *s.addr)  # where(check, new, old)
    tr, _ = old_trace.gen_fn._run(None, old_trace.args, chm, weight_mode="none", n=n, batched=old_trace.batched)
    return tr, check


# --------------------------------------------------------------- edit requests


class Rejuvenate(EditRequest):
    """``Rejuvenate(proposal, argument_mapping)`` (rejuvenate.py:45-94), used inside
    ``StaticRequest({"addr": Rejuvenate(...)})``: propose the addressed choice from
    ``proposal(*argument_mapping(current choice map))`` and weight by
    ``w + bwd_score - fwd_score``.  Normal-family proposals are fused."""

    def __init__(self, proposal, argument_mapping, reference_compat: bool = True):
        """``reference_compat=True`` (default) scores the backward move exactly as rejuvenate.py:84-86: the proposal's
        arguments are ``argument_mapping(bwd_chm)`` with ``bwd_chm`` the discarded (OLD) choices, so the backward score is
        ``log q(old; mapping(old))``.  ``reference_compat=False`` is Metropolis-Hastings' ``log q(old; mapping(new))``
        (what ``mh_chain`` uses).  Identical for proposals that ignore the current choice."""
        from ..gen.distributions import mv_normal_diag, normal

        self.reference_compat = bool(reference_compat)
        if proposal not in (normal, mv_normal_diag) and getattr(proposal, "name", None) != "mv_normal_diag":
            raise NotImplementedError("fused Rejuvenate supports normal / mv_normal_diag proposals")
        self.proposal = proposal
        self.argument_mapping = argument_mapping

    def _mapping(self):
        am = self.argument_mapping

        def mapping(cur):
            return am(ChoiceMap.choice(cur))

        # one compiled variant per Rejuvenate object
        if not hasattr(self, "_mapped"):
            self._mapped = mapping
        return self._mapped

    def moved_sites(self, trace: StaticTrace, addr: tuple) -> tuple:
        return _selected_sites(trace, Selection.all().extend(*addr))

    def edit_at(self, key, trace: StaticTrace, addr: tuple, argdiffs):
        if not Diff.static_check_no_change(argdiffs if argdiffs not in (None, ()) else ()):
            raise NotImplementedError("Rejuvenate with changed arguments")
        sel = Selection.all().extend(*addr)
        latent = _selected_sites(trace, sel)
        res = _run_chain("mh", key, trace, latent, (self._mapping(),) * len(latent), n_steps=1, step_size=1.0,
                         accept=False, compat_stale_grad=self.reference_compat)
        old = ChoiceMap.empty()
        for j in latent:
            s = trace.cm.ir.sites[j]
            old = old | ChoiceMap.entry(trace._site_value(s), *s.addr)
        w = res.alpha if trace.batched else res.alpha[0]
        return res.trace, w, Diff.unknown_change(res.trace.get_retval()), Update(old)

    def edit(self, key, tr, argdiffs):
        raise NotImplementedError("address a Rejuvenate request with StaticRequest({addr: Rejuvenate(...)})")


class HMC(EditRequest):
    """``HMC(selection, eps, L=10)`` (hmc.py:139-211): one HMC move, no accept; the
    returned weight is alpha.  Integrates exactly like the reference, including the
    carried-gradient behaviour of hmc.py:186."""

    def __init__(self, selection: Selection, eps, L: int = 10):
        self.selection = selection
        self.eps = float(eps)
        self.L = int(L)

    def moved_sites(self, trace: StaticTrace, addr: tuple = ()) -> tuple:
        return _selected_sites(trace, self.selection.extend(*addr))

    def edit_at(self, key, tr: StaticTrace, addr: tuple, argdiffs):
        """The move on the callee at ``addr`` (``StaticRequest({addr: HMC(...)})``): the selection is relative to it."""
        assert Diff.static_check_no_change(argdiffs if argdiffs not in (None, ()) else ()), \
            "HMC needs unchanged arguments (hmc.py:163)"
        latent = self.moved_sites(tr, addr)
        res = _run_chain("hmc", key, tr, latent, (), n_steps=1, step_size=self.eps, n_leapfrog=self.L,
                         compat_stale_grad=True, accept=False)
        old = ChoiceMap.empty()
        for j in latent:
            s = tr.cm.ir.sites[j]
            old = old | ChoiceMap.entry(tr._site_value(s), *s.addr)
        w = res.alpha if tr.batched else res.alpha[0]
        # the return value changes exactly when it reads a moved choice (the retdiff HMC's inner Update reports)
        from ..gen.static import _depends

        ret = res.trace.get_retval()
        changed = _depends(tr.cm.ir.ret_leaves, set(latent), set())
        return res.trace, w, (Diff.unknown_change(ret) if changed else Diff.no_change(ret)), Update(old)

    def edit(self, key, tr, argdiffs):
        from ..gen.scan import ScanTrace

        if isinstance(tr, ScanTrace):
            new_tr, w, retdiff, _ = self._edit_scan(key, tr, argdiffs)
        else:
            new_tr, w, retdiff, _ = self.edit_at(key, tr, (), argdiffs)
        # hmc.py:205-210: the backward request of an HMC move is the same move (``edit_at``, the form a StaticRequest
        # composes, keeps handing back the discarded choices)
        return new_tr, w, retdiff, HMC(self.selection, self.eps, self.L)

    def _edit_scan(self, key, tr, argdiffs):
        """HMC over the selected choices of EVERY step of a scanned trace at once (tests/inference/test_requests.py:
        237-255): the scan is unrolled into one static model (``Scan.unrolled``), the move runs in its fused chain kernel
        (one launch: all steps' latents in registers, the gradient through the whole chain of carries), and the scanned
        trace is rebuilt by an update with the moved values."""
        from ..gen.static import Batched

        assert Diff.static_check_no_change(argdiffs if argdiffs not in (None, ()) else ()), \
            "HMC needs unchanged arguments (hmc.py:163)"
        scan = tr.gen_fn
        T = tr.scan_length
        U = scan.unrolled(T)
        chm = ChoiceMap.empty()
        for t, inner in enumerate(tr.inner):
            for s in inner.cm.ir.sites:
                v = inner.values[s.index]
                chm = chm | ChoiceMap.entry(v if inner.bcast[s.index] else Batched(v), t, *s.addr)
        utr, _ = U._run(None, tuple(tr.args), chm, weight_mode="none", n=tr.n, batched=True)
        new_u, w, _, _ = self.edit_at(key, utr, (), ())
        moved = ChoiceMap.empty()
        for j in self.moved_sites(utr):
            s = utr.cm.ir.sites[j]
            moved = moved | ChoiceMap.entry(Batched(new_u.values[j]), *s.addr)
        new_tr, _, retdiff, bwd = scan.edit(key, tr, Update(moved), Diff.no_change(tr.args))
        return new_tr, (w if tr.batched else w[0]), retdiff, bwd


def SafeHMC(selection: Selection, eps, L: int = 10):
    """hmc.py:214-223: ``HMC(...).map(retdiff_assertion)`` -- an HMC move that asserts the return value of the
    generative function it is addressed at does not change (it must not read a moved choice)."""

    def retdiff_assertion(retdiff):
        assert Diff.static_check_no_change(retdiff), "SafeHMC: the return value depends on a moved choice"
        return retdiff

    return HMC(selection, eps, L).map(retdiff_assertion)

"""Particle filter with GLOBAL resampling over the GPUs of one box
(SURVEY.md section 8e): one process per GPU, particles block partitioned.

Per step every rank runs the same three kernels as the single-GPU filter on
its own block, with three tiny cross-rank exchanges between them:

    model kernel   gather x[anc] (rows may live on a peer: NVLink loads through
                   peer-mapped pointers) + propose + logpdf + local max
    exchange MAX   -> global max M
    mass kernel    exact integer mass of the local weights relative to M
    exchange MASS  -> mass of the ranks before this one (CDF offset), total S
    resample       local CDF scan on top of the offset; the ancestors of this
                   rank's parents' offspring are WRITTEN INTO THE OWNING RANK's
                   ancestor buffer (NVLink stores)
    exchange BARRIER

The exchanges are single-CTA kernels pushing 16 bytes into every peer's pad
(``gjb_exchange``): no NCCL call sits on the data path; ``torch.distributed``
is used only for the symmetric-memory rendezvous.  RNG lanes, the integer CDF
and the systematic positions are functions of GLOBAL particle indices, so an
R-rank run reproduces the single-GPU run of R*n particles bit for bit.
"""

from __future__ import annotations

import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

from ..core.choice_map import ChoiceMap
from ..core.key import PRNGKey, pf_key_table
from ..gen.capture import ArgSpec
from ..gen.expr import Expr
from ..gen.static import StaticGenerativeFunction, _dev_tensor
from ..runtime import cabi, smc_ops
from .pf import PFResult

_PAD_WORDS = cabi.GJB_PAD_WORDS  # uint64 [GJB_PAD_SLOTS][GJB_MAX_RANKS][2]


def shard_bounds(n_total: int, world: int, rank: int) -> tuple[int, int]:
    """Global particle ids [lo, hi) owned by ``rank`` (equal blocks, multiple of 4)."""
    if n_total % world:
        raise ValueError("the global particle count must be divisible by the number of ranks")
    n = n_total // world
    if n % 4:
        raise ValueError("particles per rank must be a multiple of 4 (quad RNG streams, 128-bit rows)")
    return rank * n, (rank + 1) * n


class SymmArena:
    """One symmetric-memory allocation per rank carved into named buffers; the
    same offsets are valid on every rank, ``peer_ptr`` gives a buffer's address
    in any rank's copy."""

    def __init__(self, nbytes: int, device, group):
        import torch.distributed._symmetric_memory as symm_mem

        self.nbytes = (int(nbytes) + 255) // 256 * 256
        self.buf = symm_mem.empty(self.nbytes, dtype=torch.uint8, device=device)
        self.buf.zero_()
        self.handle = symm_mem.rendezvous(self.buf, group)
        self.world = self.handle.world_size
        self.rank = self.handle.rank
        self.ptrs = [int(p) for p in self.handle.buffer_ptrs]
        self.off = 0

    def take(self, shape, dtype) -> tuple[torch.Tensor, int]:
        n = int(np.prod(shape)) * torch.empty((), dtype=dtype).element_size()
        off = (self.off + 255) // 256 * 256
        if off + n > self.nbytes:
            raise MemoryError("symmetric arena exhausted")
        t = self.buf[off:off + n].view(dtype).view(shape)
        self.off = off + n
        return t, off

    def peers(self, off: int, n_per_rank: int) -> cabi.Peers:
        P = cabi.Peers()
        P.world, P.rank, P.n_per_rank = self.world, self.rank, int(n_per_rank)
        for r in range(self.world):
            P.base[r] = self.ptrs[r] + off
        cabi.check(cabi.core().gjb_peers_set_divisor(C.byref(P)), "gjb_peers_set_divisor")
        return P


class DistributedParticleFilter:
    """``DistributedParticleFilter(step, n_per_rank)``: the bootstrap filter of
    ``ParticleFilter`` over ``world * n_per_rank`` particles with global
    systematic resampling.  ``state0`` / results are this rank's block."""

    def __init__(self, step: StaticGenerativeFunction, n_per_rank: int, group=None, fused: bool = True,
                 mode: str = "step"):
        """mode "step" (default): ONE launch per filter step on every rank (``gjb_model_pf_step``, the single-device
        default): each CTA resolves the ancestors of its own 2048 offspring slots from the previous step's
        tile-exponent CDF -- parent rows and parent states are read over NVLink when they live on a peer -- proposes,
        scores, and mails its tile record to every rank; the CTA that finishes last on a rank waits for the records
        of ALL ranks (the one cross-rank hand-off of the step) and builds the prefix table the next launch consumes.
        n_per_rank must be a multiple of 2048.
        mode "pull" (round 1): every rank resolves the ancestors of its OWN slots, reading peers'
        log-weight tiles over NVLink -- 2 fused hand-offs per step, no cross-rank stores, no barrier;
        mode "push": owners write ancestors into the peers' buffers -- 3 hand-offs per step, fused into
        the kernels (``fused=True``) or as 3 extra single-CTA launches (``fused=False``)."""
        if mode not in ("step", "pull", "push"):
            raise ValueError(mode)
        self.mode = mode
        self.fused = bool(fused) or mode == "pull"
        if not dist.is_initialized():
            raise RuntimeError("init torch.distributed first (one process per GPU)")
        self.step = step
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        if self.world > cabi.GJB_MAX_RANKS:
            raise ValueError("too many ranks")
        self.n = int(n_per_rank)
        self.n_total = self.n * self.world
        shard_bounds(self.n_total, self.world, self.rank)
        self._plans: dict = {}

    def run(self, key: PRNGKey, state0, observations: ChoiceMap, shared_args: tuple = (), *, record: bool = False,
            use_graph: bool = True) -> PFResult:
        device = cabi.require_cuda()
        state0 = state0 if isinstance(state0, (tuple, list)) else (state0,)
        state0 = tuple(_dev_tensor(s, device) for s in state0)
        obs, T = {}, None
        for addr, v in observations.leaves():
            t = _dev_tensor(v, device)
            T = t.shape[0] if T is None else T
            obs[addr] = t
        shared = tuple(s if isinstance(s, (int, float)) else _dev_tensor(s, device) for s in shared_args)
        sig = (tuple((tuple(s.shape[1:]), s.dtype) for s in state0),
               tuple((tuple(s.shape), s.dtype) if isinstance(s, torch.Tensor) else ("scalar",) for s in shared),
               tuple((a, tuple(v.shape[1:])) for a, v in obs.items()), T, record)
        plan = self._plans.get(sig)
        if plan is None:
            plan = (_DistStepPlan if self.mode == "step" else _DistPlan)(self, state0, shared, obs, T, record, device)
            self._plans[sig] = plan
        return plan.execute(key, state0, shared, obs, use_graph)


class _DistPlan:
    def __init__(self, pf: DistributedParticleFilter, state0, shared, obs, T, record, device):
        self.pf, self.device, self.T, self.record = pf, device, T, record
        n, world, rank = pf.n, pf.world, pf.rank
        specs = [ArgSpec("particle", "i32" if s.dtype == torch.int32 else "f32", tuple(s.shape[1:])) for s in state0]
        for s in shared:
            if isinstance(s, torch.Tensor):
                specs.append(ArgSpec("shared", "i32" if s.dtype == torch.int32 else "f32", tuple(s.shape)))
            else:
                specs.append(ArgSpec("scalar", "i32" if isinstance(s, int) else "f32", ()))
        self.cm = pf.step.prebuild(specs, pf_obs=tuple(obs.keys()))  # filter variant: site flags baked in
        ir = self.ir = self.cm.ir
        if len(ir.ret_leaves) != len(state0):
            raise ValueError("the step must return one leaf per state leaf")
        if 3 * T + 3 >= (1 << 16):
            raise ValueError("at most 21844 filter steps per run (16-bit hand-off offsets)")
        slots = T if record else 2
        self.slots = slots
        # ---- symmetric buffers: state ping-pong, ancestors, exchange pad
        row_elems = [int(np.prod(s.shape[1:])) if s.ndim > 1 else 1 for s in state0]
        need = (sum(slots * n * r * 4 + 512 for r in row_elems) + 2 * (slots * n * 4 + 512) + _PAD_WORDS * 8 + 1024
                + 2 * 8 * ((n + smc_ops.TILE - 1) // smc_ops.TILE + 1) + 1024)
        self.arena = SymmArena(need, device, pf.group)
        self.bufs, self.buf_off = [], []
        for s in state0:
            t, off = self.arena.take((slots,) + tuple(s.shape), s.dtype)
            self.bufs.append(t)
            self.buf_off.append(off)
        self.anc, self.anc_off = self.arena.take((slots, n), torch.int32)
        self.pad, self.pad_off = self.arena.take((_PAD_WORDS,), torch.int64)
        self.pull = pf.mode == "pull"
        n_tiles = max(1, (n + smc_ops.TILE - 1) // smc_ops.TILE)
        if self.pull:  # peers read these: log-weights and inclusive tile prefixes, double buffered across steps
            self.logw_sym, self.logw_off = self.arena.take((slots, n), torch.float32)
            self.tpre, self.tpre_off = self.arena.take((2, n_tiles), torch.int64)
        # ---- local buffers
        self.state_in = tuple(torch.empty_like(s) for s in state0)
        self.shared = tuple(torch.empty_like(s) if isinstance(s, torch.Tensor) else s for s in shared)
        self.obs = {a: torch.empty_like(v) for a, v in obs.items()}
        self.keys = torch.empty((T, 8), dtype=torch.int32, device=device)
        self.logw = torch.empty((T if record else 1, n), dtype=torch.float32, device=device)
        self.lse = torch.empty((T, 3), dtype=torch.float64, device=device)
        self.ws = smc_ops.WeightWorkspace(n, device)
        self.wmax2 = torch.empty(2, dtype=torch.int32, device=device)
        self.m_global = torch.empty(1, dtype=torch.float32, device=device)
        self.c_offset = torch.empty(1, dtype=torch.int64, device=device)
        self.s_total = torch.empty(1, dtype=torch.int64, device=device)
        self.epoch = torch.zeros(1, dtype=torch.int64, device=device)
        self.counter = torch.zeros(1, dtype=torch.int32, device=device)
        L = cabi.Link()
        L.rank, L.world = rank, world
        for r in range(world):
            L.pads[r] = self.arena.ptrs[r] + self.pad_off
        L.epoch = self.epoch.data_ptr()
        L.counter = self.counter.data_ptr()
        self.link = torch.from_numpy(np.frombuffer(bytes(L), dtype=np.uint8).copy()).to(device)
        self.final = tuple(torch.empty_like(s) for s in state0)
        # device tables of peer pointers: per slot, gjb_peers[GJB_MAX_ARGS]
        self.peer_tabs = []
        for sl in range(slots):
            arr = (cabi.Peers * cabi.GJB_MAX_ARGS)()
            for i, (s, off) in enumerate(zip(state0, self.buf_off)):
                slot_bytes = n * row_elems[i] * 4
                arr[i] = self.arena.peers(off + sl * slot_bytes, n)
            host = np.frombuffer(bytes(arr), dtype=np.uint8).copy()
            self.peer_tabs.append(torch.from_numpy(host).to(device))
        self.row_elems = row_elems
        self.obs_sites = {ir.site_index(a): a for a in obs}
        self.graph = None
        dist.barrier(pf.group)  # every rank has zeroed its arena before anyone pushes into it
        self._build()

    def _build(self):
        pf, ir, T, n = self.pf, self.ir, self.T, self.pf.n
        rank, world = pf.rank, pf.world
        self.steps = []
        for t in range(T):
            slot = t if self.record else (t & 1)
            pslot = (t - 1) if self.record else ((t - 1) & 1)
            A = cabi.ModelArgs()
            A.n = n
            A.idx_offset = rank * n
            A.key_dev = self.keys[t].data_ptr()
            for i in range(len(self.state_in)):
                A.args[i] = self.state_in[i].data_ptr() if t == 0 else self.bufs[i][pslot].data_ptr()
            if t > 0:
                A.gather = self.anc[pslot].data_ptr()
                A.peer_args = self.peer_tabs[pslot].data_ptr()
            for k, s in enumerate(self.shared):
                i = len(self.state_in) + k
                if isinstance(s, torch.Tensor):
                    A.args[i] = s.data_ptr()
                else:
                    A.scalars[i] = float(s)
            for site in ir.sites:
                j = site.index
                if j in self.obs_sites:
                    A.site_in[j] = self.obs[self.obs_sites[j]][t].data_ptr()
                    A.site_flags[j] = cabi.SITE_WEIGHT | cabi.SITE_BCAST
                else:
                    A.site_flags[j] = cabi.SITE_SAMPLE
            for k, r in enumerate(ir.ret_leaves):
                if not isinstance(r, Expr):
                    raise ValueError("the step's return value must depend on its choices / arguments")
                out = self.bufs[k][slot]
                if r.op == "site" and r.attr not in self.obs_sites:
                    A.site_out[r.attr] = out.data_ptr()
                else:
                    A.ret_out[k] = out.data_ptr()
            lw = self.logw_sym[slot] if self.pull else self.logw[t if self.record else 0]
            A.weight_out = lw.data_ptr()
            wm = self.wmax2[t & 1:]
            A.wmax = wm.data_ptr()
            if self.pull:
                A.link = self.link.data_ptr()
                A.wait_off = 0  # no barrier: peers only READ this rank's buffers
                A.push_off = 2 * t + 1
            elif pf.fused:
                A.link = self.link.data_ptr()
                A.wait_off = 3 * t if t > 0 else 0  # BARRIER of step t-1 has offset 3(t-1)+3
                A.push_off = 3 * t + 1

            def xchg(mode, off):
                X = cabi.XchgArgs()
                X.rank, X.world, X.mode = rank, world, mode
                X.n_tiles = self.ws.tiles
                for r in range(world):
                    X.pads[r] = self.arena.ptrs[r] + self.pad_off
                X.epoch = self.epoch.data_ptr()
                X.tag_offset = off
                X.wmax = wm.data_ptr()
                X.tile_mass = self.ws.tile_mass.data_ptr()
                X.m_global = self.m_global.data_ptr()
                X.c_offset = self.c_offset.data_ptr()
                X.s_total = self.s_total.data_ptr()
                return X

            R = self.ws.systematic_args(
                lw, None, self.anc[slot], n_total=pf.n_total, out_lo=0, anc_base=rank * n, m_global=self.m_global,
                c_offset=self.c_offset, s_total=self.s_total, key_dev=self.keys[t][2:], lse_out=self.lse[t], wmax=wm,
                wmax_next=self.wmax2[(t + 1) & 1:],
            )
            R.out_n = pf.n_total
            anc_peers = self.arena.peers(self.anc_off + slot * n * 4, n)
            lw_peers = pre_peers = None
            if self.pull:
                R.out_lo, R.out_n = rank * n, n  # my own slots only
                lw_peers = self.arena.peers(self.logw_off + slot * n * 4, n)
                pre_peers = self.arena.peers(self.tpre_off + (t & 1) * self.tpre.shape[1] * 8, n)
            self.steps.append(dict(A=A, lw=lw, wm=wm, R=R, anc_peers=anc_peers, lw_peers=lw_peers, pre_peers=pre_peers,
                                   x_max=xchg(cabi.XCHG_MAX, 3 * t + 1), x_mass=xchg(cabi.XCHG_MASS, 3 * t + 2),
                                   x_bar=xchg(cabi.XCHG_BARRIER, 3 * t + 3)))
        self.x_final = xchg(cabi.XCHG_BARRIER, 3 * T + 1)
        self.x_end = xchg(cabi.XCHG_BARRIER, 3 * T + 2)
        last = (T - 1) if self.record else ((T - 1) & 1)
        self.final_peers = [self.arena.peers(off + last * n * self.row_elems[i] * 4, n) for i, off in enumerate(self.buf_off)]
        self.last = last

    def _enqueue(self):
        core = cabi.core()
        stream = cabi.stream_ptr(self.device)
        lib = self.cm.lib
        cabi.check(core.gjb_wmax_reset(self.wmax2.data_ptr(), stream), "gjb_wmax_reset")
        for t, st in enumerate(self.steps):
            cabi.check(lib.gjb_model_launch(C.byref(st["A"]), stream), "gjb_model_launch")
            if self.pull:
                cabi.check(core.gjb_weight_mass_prefix_linked(st["lw"].data_ptr(), st["lw"].numel(), self.ws.tile_mass.data_ptr(),
                                                              self.tpre[t & 1].data_ptr(), self.link.data_ptr(), 2 * t + 1,
                                                              2 * t + 2, stream), "gjb_weight_mass_prefix_linked")
                cabi.check(core.gjb_resample_systematic_pull(C.byref(st["R"]), C.byref(st["lw_peers"]), C.byref(st["pre_peers"]),
                                                             self.link.data_ptr(), 2 * t + 1, 2 * t + 2, stream),
                           "gjb_resample_systematic_pull")
                continue
            if self.pf.fused:
                cabi.check(core.gjb_weight_mass_linked(st["lw"].data_ptr(), st["lw"].numel(), self.ws.tile_mass.data_ptr(),
                                                       self.link.data_ptr(), 3 * t + 1, 3 * t + 2, stream),
                           "gjb_weight_mass_linked")
                cabi.check(core.gjb_resample_systematic_linked(C.byref(st["R"]), C.byref(st["anc_peers"]),
                                                               self.link.data_ptr(), 3 * t + 1, 3 * t + 2, 3 * t + 3, stream),
                           "gjb_resample_systematic_linked")
                continue
            cabi.check(core.gjb_exchange(C.byref(st["x_max"]), stream), "gjb_exchange(max)")
            cabi.check(core.gjb_weight_mass(st["lw"].data_ptr(), st["lw"].numel(), st["wm"].data_ptr(),
                                            self.m_global.data_ptr(), self.ws.tile_mass.data_ptr(), stream), "gjb_weight_mass")
            cabi.check(core.gjb_exchange(C.byref(st["x_mass"]), stream), "gjb_exchange(mass)")
            cabi.check(core.gjb_resample_systematic_peers(C.byref(st["R"]), C.byref(st["anc_peers"]), stream),
                       "gjb_resample_systematic_peers")
            cabi.check(core.gjb_exchange(C.byref(st["x_bar"]), stream), "gjb_exchange(barrier)")
        if self.pf.fused and not self.pull:  # all ranks finished their last resample before anyone gathers the final state
            cabi.check(core.gjb_exchange(C.byref(self.x_final), stream), "gjb_exchange(final barrier)")
        for k in range(len(self.bufs)):
            cabi.check(core.gjb_gather_rows_peers(C.byref(self.final_peers[k]), self.anc[self.last].data_ptr(),
                                                  self.final[k].data_ptr(), self.pf.n, self.row_elems[k] * 4, stream),
                       "gjb_gather_rows_peers")
        if self.pull:  # nobody overwrites state / log-weights of this run while a peer still gathers from them
            cabi.check(core.gjb_exchange(C.byref(self.x_end), stream), "gjb_exchange(end of run)")
        cabi.check(core.gjb_epoch_bump(self.epoch.data_ptr(), stream), "gjb_epoch_bump")

    def launches_per_run(self) -> int:
        return (3 + 3 * self.T if self.pf.fused else 2 + 6 * self.T) + len(self.bufs)  # pull: 3/step too, 2 hand-offs

    def execute(self, key, state0, shared, obs, use_graph):
        # the per-step key table is derived on the device from the run key's two words (3 us; the NumPy threefry of
        # core/key.py pf_key_table costs 0.4-1.2 ms of host time per run, a quarter of a 1 M-particle, 100-step run)
        w0, w1 = key.collapsed()
        cabi.check(cabi.core().gjb_pf_key_table(w0, w1, self.T, self.keys.data_ptr(), cabi.stream_ptr(self.device)), "gjb_pf_key_table")
        for dst, src in zip(self.state_in, state0):
            if dst.data_ptr() != src.data_ptr():
                dst.copy_(src, non_blocking=True)
        for dst, src in zip(self.shared, shared):
            if isinstance(dst, torch.Tensor) and dst.data_ptr() != src.data_ptr():
                dst.copy_(src, non_blocking=True)
        for a in self.obs:
            if self.obs[a].data_ptr() != obs[a].data_ptr():
                self.obs[a].copy_(obs[a], non_blocking=True)
        if use_graph:
            if self.graph is None:
                self._enqueue()
                torch.cuda.current_stream(self.device).synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._enqueue()
                self.graph = g
            self.graph.replay()
        else:
            self._enqueue()
        inc = self.lse[:, 2]
        hist = {"state": tuple(self.bufs), "log_weights": self.logw_sym if self.pull else self.logw} if self.record else None
        return PFResult(state=self.final, log_marginal_likelihood=inc.sum(), log_increments=inc, lse_terms=self.lse,
                        ancestors=self.anc if self.record else None, history=hist)


class _DistStepPlan:
    """Launch arguments of the global-resampling filter on the single-launch step kernel (mode="step")."""

    def __init__(self, pf: DistributedParticleFilter, state0, shared, obs, T, record, device):
        import os

        self.pf, self.device, self.T, self.record = pf, device, T, record
        self.stepmode = True
        n, world, rank = pf.n, pf.world, pf.rank
        if n % cabi.TE_TILE:
            raise ValueError(f"mode='step': particles per rank must be a multiple of {cabi.TE_TILE}")
        tiles = n // cabi.TE_TILE
        if tiles * world > cabi.TE_MAX_TILES or pf.n_total > (1 << 26):
            raise NotImplementedError(f"mode='step' resamples over at most {cabi.TE_MAX_TILES} tiles of {cabi.TE_TILE} particles")
        if T >= 65535:
            raise ValueError("at most 65534 filter steps per run (16-bit step field of the record tags)")
        specs = [ArgSpec("particle", "i32" if s.dtype == torch.int32 else "f32", tuple(s.shape[1:])) for s in state0]
        for s in shared:
            if isinstance(s, torch.Tensor):
                specs.append(ArgSpec("shared", "i32" if s.dtype == torch.int32 else "f32", tuple(s.shape)))
            else:
                specs.append(ArgSpec("scalar", "i32" if isinstance(s, int) else "f32", ()))
        self.cm = pf.step.prebuild(specs, pf_obs=tuple(obs.keys()))
        ir = self.ir = self.cm.ir
        if not self.cm.info.get("pf_step", False):
            raise NotImplementedError("mode='step': this model's return value does not feed back as its state")
        if len(ir.ret_leaves) != len(state0):
            raise ValueError("the step must return one leaf per state leaf")
        slots = T if record else 2
        self.slots, self.tiles = slots, tiles
        row_elems = [int(np.prod(s.shape[1:])) if s.ndim > 1 else 1 for s in state0]
        self.row_elems = row_elems
        # ---- symmetric buffers: what peers read (state ping-pong, CDF rows) or write (tile-record mailbox, barrier pad)
        need = (sum(slots * n * r * 4 + 512 for r in row_elems) + 2 * n * 8 + 512 + cabi.TE_MAILBOX_WORDS * 8 + 512
                + _PAD_WORDS * 8 + 1024)
        self.arena = SymmArena(need, device, pf.group)
        self.bufs, self.buf_off = [], []
        for s in state0:
            t, off = self.arena.take((slots,) + tuple(s.shape), s.dtype)
            self.bufs.append(t)
            self.buf_off.append(off)
        self.cdf, self.cdf_off = self.arena.take((2, n), torch.int64)
        self.mailbox, self.mail_off = self.arena.take((cabi.TE_MAILBOX_WORDS,), torch.int64)
        self.pad, self.pad_off = self.arena.take((_PAD_WORDS,), torch.int64)
        # ---- local buffers
        self.state_in = tuple(torch.empty_like(s) for s in state0)
        self.shared = tuple(torch.empty_like(s) if isinstance(s, torch.Tensor) else s for s in shared)
        self.obs = {a: torch.empty_like(v) for a, v in obs.items()}
        self.keys = torch.empty((T, 8), dtype=torch.int32, device=device)
        self.logw = torch.empty((T if record else 1, n), dtype=torch.float32, device=device)
        self.anc = torch.empty((slots, n), dtype=torch.int32, device=device)
        self.lse = torch.empty((T, 3), dtype=torch.float64, device=device)
        self.tables = torch.zeros((2, C.sizeof(cabi.StepTable)), dtype=torch.uint8, device=device)
        self.epoch = torch.zeros(1, dtype=torch.int64, device=device)
        self.ticket = torch.zeros(1, dtype=torch.int32, device=device)
        self.final = tuple(torch.empty_like(s) for s in state0)
        L = cabi.StepLink()
        L.rank, L.world, L.tiles_per_rank = rank, world, tiles
        for r in range(world):
            L.mailbox[r] = self.arena.ptrs[r] + self.mail_off
        L.epoch, L.ticket = self.epoch.data_ptr(), self.ticket.data_ptr()
        self.link = torch.from_numpy(np.frombuffer(bytes(L), dtype=np.uint8).copy()).to(device)

        def dev_bytes(obj):
            return torch.from_numpy(np.frombuffer(bytes(obj), dtype=np.uint8).copy()).to(device)

        # device tables of peer pointers: per state slot gjb_peers[GJB_MAX_ARGS]; per CDF parity one gjb_peers
        self.peer_tabs = []
        for sl in range(slots):
            arr = (cabi.Peers * cabi.GJB_MAX_ARGS)()
            for i, off in enumerate(self.buf_off):
                arr[i] = self.arena.peers(off + sl * n * row_elems[i] * 4, n)
            self.peer_tabs.append(dev_bytes(arr))
        self.cdf_peers = [dev_bytes(self.arena.peers(self.cdf_off + par * n * 8, n)) for par in range(2)]
        self.obs_sites = {ir.site_index(a): a for a in obs}
        self.graph = None
        pdl = os.environ.get("GJB_PDL", "1") != "0"
        self.table_kernel = os.environ.get("GJB_STEP_TABLE_KERNEL", "0") == "1"
        # rank-level table (GJB_STEP_LIGHT=1, opt-in): the last CTA reduces the records to S, E and the ranks' prefix; the
        # consumers form the tile prefix of their parents' rank(s) themselves.  Measured on 2 / 4 / 8 B200s
        # (profiles/r2_call26_*): 39.8 / 42.9 / 51.7 us per step against 29.0 / 36.0 / 53.0 for the per-tile table built by
        # the last CTA -- the longer dependent-load chain in every consumer CTA costs more than the shorter serial tail
        # saves until 8 ranks, so the per-tile table stays the default.
        self.light = (not self.table_kernel) and os.environ.get("GJB_STEP_LIGHT", "0") == "1"
        self.table_full = torch.zeros(C.sizeof(cabi.StepTable), dtype=torch.uint8, device=device)
        dist.barrier(pf.group)  # every rank has zeroed its arena before anyone mails into it
        # ---- per-step arguments
        self.sargs = []
        self.targs = []
        for t in range(T):
            A = cabi.StepArgs()
            A.n, A.n_total, A.idx_offset, A.slot_offset = n, pf.n_total, rank * n, rank * n
            A.step = t
            A.flags = (cabi.STEP_PDL if (pdl and t > 0) else 0) | (cabi.STEP_FLAGWAIT if os.environ.get("GJB_STEP_FLAGWAIT") == "1" else 0)
            if self.light:
                A.flags |= cabi.STEP_LIGHT
            A.key_dev = self.keys[t].data_ptr()
            slot = t if record else (t & 1)
            pslot = (t - 1) if record else ((t - 1) & 1)
            for i in range(len(self.state_in)):
                A.args[i] = self.state_in[i].data_ptr() if t == 0 else self.bufs[i][pslot].data_ptr()
                A.state_out[i] = self.bufs[i][slot].data_ptr()
            for k, s in enumerate(self.shared):
                i = len(self.state_in) + k
                if isinstance(s, torch.Tensor):
                    A.args[i] = s.data_ptr()
                else:
                    A.scalars[i] = float(s)
            for j, addr in self.obs_sites.items():
                A.site_in[j] = self.obs[addr][t].data_ptr()
            if record:
                A.weight_out = self.logw[t].data_ptr()
            elif t == T - 1:
                A.weight_out = self.logw[0].data_ptr()
            if t > 0:
                A.prev_cdf = self.cdf[(t - 1) & 1].data_ptr()
                A.cdf_peers = self.cdf_peers[(t - 1) & 1].data_ptr()
                A.peer_args = self.peer_tabs[pslot].data_ptr()
                A.table_in = self.tables[(t - 1) & 1].data_ptr()
                A.prev_key = self.keys[t - 1][2:].data_ptr()
                if record:
                    A.ancestors_out = self.anc[t - 1].data_ptr()
            A.cdf_out = self.cdf[t & 1].data_ptr()
            A.link = self.link.data_ptr()
            if not self.table_kernel:  # default: the step kernel's last CTA builds the table (measured faster, see pf.py)
                A.table_out = self.tables[t & 1].data_ptr()
                A.lse_out = self.lse[t].data_ptr()
            self.sargs.append(A)
            B = cabi.TeTableArgs()
            B.link, B.step, B.flags = self.link.data_ptr(), t, (cabi.STEP_PDL if pdl else 0)
            B.slot_offset, B.n_local, B.n_total = rank * n, n, pf.n_total
            B.reskey = self.keys[t][2:].data_ptr()
            B.table_out = self.tables[t & 1].data_ptr()
            B.lse_out = self.lse[t].data_ptr()
            self.targs.append(B)
        last = (T - 1) if record else ((T - 1) & 1)
        self.last = last
        R = cabi.TeResampleArgs()
        R.cdf = self.cdf[(T - 1) & 1].data_ptr()
        R.table = self.tables[(T - 1) & 1].data_ptr()
        if self.light:
            # the closing resampling reads a per-tile table: built once per run by the table kernel from the last step's records
            self.targs[T - 1].table_out = self.table_full.data_ptr()
            R.table = self.table_full.data_ptr()
        R.cdf_peers = self.cdf_peers[(T - 1) & 1].data_ptr()
        R.n_tiles_total, R.n_total, R.out_lo, R.out_n = tiles * world, pf.n_total, rank * n, n
        R.key_dev = self.keys[T - 1][2:].data_ptr()
        R.ancestors = self.anc[last].data_ptr()
        self.close = R
        self.final_peers = [self.arena.peers(off + last * n * row_elems[i] * 4, n) for i, off in enumerate(self.buf_off)]
        X = cabi.XchgArgs()
        X.rank, X.world, X.mode = rank, world, cabi.XCHG_BARRIER
        for r in range(world):
            X.pads[r] = self.arena.ptrs[r] + self.pad_off
        X.epoch = self.epoch.data_ptr()
        X.tag_offset = 1
        self.x_end = X

    def _enqueue(self):
        core = cabi.core()
        stream = cabi.stream_ptr(self.device)
        lib = self.cm.lib
        for t in range(self.T):
            cabi.check(lib.gjb_model_pf_step(C.byref(self.sargs[t]), stream), "gjb_model_pf_step")
            if self.table_kernel:
                cabi.check(core.gjb_te_table(C.byref(self.targs[t]), stream), "gjb_te_table")
        if self.light:
            cabi.check(core.gjb_te_table(C.byref(self.targs[self.T - 1]), stream), "gjb_te_table(closing)")
        cabi.check(core.gjb_te_resample(C.byref(self.close), stream), "gjb_te_resample")
        for k in range(len(self.bufs)):
            cabi.check(core.gjb_gather_rows_peers(C.byref(self.final_peers[k]), self.anc[self.last].data_ptr(),
                                                  self.final[k].data_ptr(), self.pf.n, self.row_elems[k] * 4, stream),
                       "gjb_gather_rows_peers")
        # nobody overwrites state / CDF rows of this run while a peer still reads them
        cabi.check(core.gjb_exchange(C.byref(self.x_end), stream), "gjb_exchange(end of run)")
        cabi.check(core.gjb_epoch_bump(self.epoch.data_ptr(), stream), "gjb_epoch_bump")

    def launches_per_run(self) -> int:
        return (2 if self.table_kernel else 1) * self.T + 3 + len(self.bufs) + (1 if self.light else 0)  # step kernel (+ table kernel) per step, closing resample, gathers, barrier, epoch

    def execute(self, key, state0, shared, obs, use_graph):
        # the per-step key table is derived on the device from the run key's two words (3 us; the NumPy threefry of
        # core/key.py pf_key_table costs 0.4-1.2 ms of host time per run, a quarter of a 1 M-particle, 100-step run)
        w0, w1 = key.collapsed()
        cabi.check(cabi.core().gjb_pf_key_table(w0, w1, self.T, self.keys.data_ptr(), cabi.stream_ptr(self.device)), "gjb_pf_key_table")
        for dst, src in zip(self.state_in, state0):
            if dst.data_ptr() != src.data_ptr():
                dst.copy_(src, non_blocking=True)
        for dst, src in zip(self.shared, shared):
            if isinstance(dst, torch.Tensor) and dst.data_ptr() != src.data_ptr():
                dst.copy_(src, non_blocking=True)
        for a in self.obs:
            if self.obs[a].data_ptr() != obs[a].data_ptr():
                self.obs[a].copy_(obs[a], non_blocking=True)
        if use_graph:
            if self.graph is None:
                self._enqueue()
                torch.cuda.current_stream(self.device).synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._enqueue()
                self.graph = g
            self.graph.replay()
        else:
            self._enqueue()
        inc = self.lse[:, 2]
        hist = {"state": tuple(self.bufs), "log_weights": self.logw} if self.record else None
        return PFResult(state=self.final, log_marginal_likelihood=inc.sum(), log_increments=inc, lse_terms=self.lse,
                        ancestors=self.anc if self.record else None, history=hist)

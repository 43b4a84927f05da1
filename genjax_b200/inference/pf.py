"""Bootstrap / guided particle filter: the extend - reweight - resample loop.

The reference ships NO particle-filter class: the loop is a user idiom
(docs/cookbook/inactive/inference/importance_sampling.ipynb cell 16,
mapping_tutorial.ipynb cell 37; reweight formula inference/smc.py:383):

    for t:  trs, w = vmap(step.importance, (0, None, 0))(keys_t, C["y"].set(y_t), (x_prev,))
            log_w += w;  idx ~ categorical(log_w - logsumexp(log_w)) per offspring
            x_prev = trs.get_retval()[idx]

``ParticleFilter`` is that idiom on the GPU:

  * ``mode="auto"`` (default) -> ``"step"``: ONE launch per filter step
    (``gjb_model_pf_step``, gen/codegen.py ``pf_step_kernel``, csrc/gjb_step.cuh):
    every CTA resolves the ancestors of its own 2048 offspring slots from the
    previous step's tile-exponent integer CDF (output-slot systematic
    resampling), gathers, proposes, scores, and publishes the CDF row + tile
    record of its new weights; T launches + a closing ``gjb_te_resample`` in one
    CUDA graph, programmatic dependent launch between the steps;
  * ``mode="steps"``: the same step body for all T steps in ONE cooperative
    launch (``gjb_model_pf_steps``), a grid barrier per step;
  * ``mode="graph"``: round 1's two launches per step over the exact-max CDF
    (``gjb_model_launch`` + ``gjb_mass_resample_systematic``), also the form
    behind ``resampler="multinomial"`` and ``reference_max="analytic"``;
  * ``mode="persistent"``: round 1's cooperative kernel, three grid barriers per step.

"step" / "steps" agree bit for bit with each other, "graph" / "persistent" with each
other; the two families are different realisations of the same estimator (their
integer CDFs quantise the masses against different references).
"""

from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np
import torch

from ..core.choice_map import ChoiceMap
from ..core.key import PRNGKey, pf_key_table
from ..gen.capture import ArgSpec
from ..gen.expr import Expr, I32
from ..gen.static import StaticGenerativeFunction, _dev_tensor
from ..runtime import cabi, smc_ops


@dataclass
class PFResult:
    state: tuple  # final (resampled) per-particle state leaves
    log_marginal_likelihood: torch.Tensor  # float64 0-d, sum of the increments
    log_increments: torch.Tensor  # float64 [T]
    lse_terms: torch.Tensor  # float64 [T, 3] = (M or E ln 2, S, increment)
    ancestors: torch.Tensor | None  # int32 [T, N] when record=True
    history: dict | None  # per-step pre-resampling state / log-weights when record=True
    _ess: torch.Tensor | None = None

    @property
    def ess(self) -> torch.Tensor:
        """Effective sample size (sum w)^2 / sum w^2 of every step's pre-resampling weights, float64 [T] (SURVEY section 5:
        the per-step diagnostic of an SMC run).  Needs ``record=True`` (the weights of every step); one small kernel per step."""
        if self._ess is None:
            if self.history is None:
                raise ValueError("per-step ESS needs the log-weights of every step: run with record=True")
            lw = self.history["log_weights"]
            T = lw.shape[0]
            ess = torch.empty(T, dtype=torch.float64, device=lw.device)
            ws = smc_ops.WeightWorkspace(lw.shape[1], lw.device)
            for t in range(T):
                ess[t] = smc_ops.weight_ess(lw[t].contiguous(), ws.lse_terms(lw[t].contiguous()))
            self._ess = ess
        return self._ess

    def genealogy(self) -> torch.Tensor:
        """int32 [T, N]: row t holds, for every FINAL particle, the index of its ancestor among step t's particles (the
        ancestry history of the surviving lineages, composed from the recorded per-step ancestors)."""
        if self.ancestors is None:
            raise ValueError("the genealogy needs the ancestors of every step: run with record=True")
        T, n = self.ancestors.shape
        out = torch.empty((T, n), dtype=torch.int32, device=self.ancestors.device)
        cur = self.ancestors[T - 1]
        out[T - 1] = cur
        for t in range(T - 2, -1, -1):
            cur = smc_ops.gather_rows(self.ancestors[t].contiguous(), cur.contiguous())
            out[t] = cur
        return out

    def state_dict(self) -> dict:
        """Checkpoint of a filter run as CPU tensors: final state, log-marginal-likelihood terms and, for a
        recorded run, the ancestry and effective-sample-size history.  ``ParticleFilter.run(key, state, obs[t0:])`` from
        ``state`` resumes the filter; the estimate of the whole run is the sum of the parts."""
        d = {"format": "genjax_b200.PFResult/1", "state": [s.detach().cpu() for s in self.state],
             "log_increments": self.log_increments.detach().cpu(), "lse_terms": self.lse_terms.detach().cpu(),
             "log_marginal_likelihood": float(self.log_marginal_likelihood.item())}
        if self.ancestors is not None:
            d["ancestors"] = self.ancestors.detach().cpu()
            d["genealogy"] = self.genealogy().detach().cpu()
            d["ess"] = self.ess.detach().cpu()
        return d

    def save(self, path) -> None:
        torch.save(self.state_dict(), path)

    @staticmethod
    def load(path) -> dict:
        d = torch.load(path, map_location="cpu", weights_only=True)
        if d.get("format") != "genjax_b200.PFResult/1":
            raise ValueError("not a genjax_b200 PFResult checkpoint")
        return d


class ParticleFilter:
    """``ParticleFilter(step, n_particles)`` with ``step(*state, *shared)`` a
    static ``@gen`` kernel whose return value is the next state.

    observations: ChoiceMap whose leaves carry a leading time axis [T, ...];
    every step constrains those addresses to the t-th slice (broadcast over
    particles).  Every other site is proposed from the model (bootstrap)."""

    def __init__(self, step: StaticGenerativeFunction, n_particles: int, *, n_state: int = 1, resampler: str = "systematic",
                 idx_offset: int = 0, n_total: int | None = None, mode: str = "auto", reference_max: str = "running",
                 single_pass: bool = False):
        """``reference_max``: what the exact integer weight masses are taken relative to.  "running" (default): the
        maximum of the step's weights, found by a running max in the model kernel, masses in the resampling launch.
        "analytic" (graph mode, scalar-site models): an upper bound of the incremental weight derived from the model
        (``gen/bounds.py``); the masses are then accumulated in the model kernel itself and the step needs neither
        the max nor the mass pass.  Same estimator, ancestors differ in the last bits of the
        masses; a bound so loose that every mass underflows shows up as ``lse_terms[:, 1] == 0``."""
        if resampler not in ("systematic", "multinomial"):
            raise ValueError(f"resampler must be 'systematic' or 'multinomial', not {resampler!r}")
        if mode not in ("auto", "persistent", "graph", "step", "steps"):
            raise ValueError(mode)
        if mode == "steps" and reference_max != "running":
            raise ValueError("mode='steps' forms its masses per tile (tile-exponent CDF); it takes no reference_max")
        if resampler == "multinomial":
            # the reference idiom itself: N independent categorical draws over the normalised weights per step
            # (mapping_tutorial.ipynb cell 37; inference/smc.py:102-109), as inverse-CDF draws over the exact integer CDF
            if mode not in ("auto", "graph") or reference_max != "running" or single_pass or idx_offset or n_total not in (None, n_particles):
                raise NotImplementedError("resampler='multinomial' runs in mode='graph' on one device (model launch + mass + "
                                          "CDF + search per step)")
            mode = "graph"
        self.resampler = resampler
        if mode == "auto":
            # the single-launch step kernel (tile-exponent CDF) wherever it applies -- the fastest verified form (B200, 1 M
            # particles, d = 1: 17-18 us per step against 21.8 for the two-launch exact-max form); a plan falls back to
            # "graph" when the model or the particle count is outside what the step kernel handles (see _Plan)
            mode = "graph" if (reference_max != "running" or single_pass) else "auto"
        if mode == "step" and reference_max != "running":
            raise ValueError("mode='step' forms its masses per tile (tile-exponent CDF); it takes no reference_max")
        if reference_max not in ("running", "analytic"):
            raise ValueError(reference_max)
        if reference_max == "analytic" and mode != "graph":
            raise ValueError("reference_max='analytic' needs mode='graph'")
        if single_pass and (reference_max != "analytic" or idx_offset != 0):
            raise ValueError("single_pass=True needs reference_max='analytic' on one device")
        # single_pass: ONE launch per step -- every CTA first resolves the ancestors of its own 2048 offspring slots
        # from the previous step's weights (output-slot resampling), then gathers, proposes, scores and accumulates
        # the masses of exactly those slots; a final resampling launch closes the run
        self.single_pass = bool(single_pass)
        self.reference_max = reference_max
        self.mode = mode
        self.fuse_mass_resample = True  # graph mode: gjb_mass_resample_systematic when the particle count fits
        self.step = step
        self.n = int(n_particles)
        self.n_state = n_state
        self.idx_offset = int(idx_offset)
        self.n_total = int(n_total) if n_total is not None else self.n
        self._plans: dict = {}

    def weight_upper_bound(self, state0, observations: ChoiceMap, shared_args: tuple = ()):
        """Analytic supremum over particles of ONE step's incremental log-weight (the sum of the observed sites'
        log-densities), or None when it cannot be derived from the model (``gen/bounds.py``).  Host-only: the step
        is captured symbolically, nothing is launched.  A reference maximum of this kind is what the single-pass
        filter step of DESIGN.md section 10 needs; today it is a diagnostic (``lse_terms[:, 0] <= bound``)."""
        from ..gen import bounds
        from ..gen import capture as cap_

        state0 = state0 if isinstance(state0, (tuple, list)) else (state0,)
        specs, values = [], {}
        for i, s in enumerate(state0):
            t = torch.as_tensor(s)
            specs.append(ArgSpec("particle", "i32" if t.dtype in (torch.int32, torch.int64) else "f32", tuple(t.shape[1:])))
        for k, s in enumerate(shared_args):
            i = len(state0) + k
            if isinstance(s, (int, float)):
                specs.append(ArgSpec("scalar", "i32" if isinstance(s, int) else "f32", ()))
                values[i] = s
            else:
                t = torch.as_tensor(s).detach().cpu()
                specs.append(ArgSpec("shared", "i32" if t.dtype in (torch.int32, torch.int64) else "f32", tuple(t.shape)))
                values[i] = t.numpy()
        tree = ("tuple", [("leaf", i) for i in range(len(specs))])
        ir = cap_.capture(self.step.source, self.step.__name__, specs, tree)
        sites = [ir.site_index(addr) for addr, _ in observations.leaves()]
        b = bounds.log_weight_upper_bound(ir, sites)
        return None if b is None else bounds.evaluate_invariant(b, values)

    # ------------------------------------------------------------------ plan
    def _plan(self, state0: tuple, shared: tuple, obs: dict, T: int, record: bool, device):
        sig = (tuple((tuple(s.shape[1:]), s.dtype) for s in state0),
               tuple((tuple(s.shape), s.dtype) if isinstance(s, torch.Tensor) else ("scalar", type(s).__name__) for s in shared),
               tuple((a, tuple(v.shape[1:])) for a, v in obs.items()), T, record)
        plan = self._plans.get(sig)
        if plan is None:
            plan = _Plan(self, state0, shared, obs, T, record, device)
            self._plans[sig] = plan
        return plan

    def run(self, key: PRNGKey, state0, observations: ChoiceMap, shared_args: tuple = (), *, record: bool = False,
            use_graph: bool = True) -> PFResult:
        device = cabi.require_cuda()
        state0 = state0 if isinstance(state0, (tuple, list)) else (state0,)
        state0 = tuple(_dev_tensor(s, device) for s in state0)
        for s in state0:
            if s.shape[0] != self.n:
                raise ValueError("initial state must have n_particles rows")
        obs = {}
        T = None
        for addr, v in observations.leaves():
            t = _dev_tensor(v, device)
            T = t.shape[0] if T is None else T
            if t.shape[0] != T:
                raise ValueError("observation leaves disagree on the number of steps")
            obs[addr] = t
        if T is None:
            raise ValueError("no observations")
        shared = tuple(s if isinstance(s, (int, float)) else _dev_tensor(s, device) for s in shared_args)
        plan = self._plan(state0, shared, obs, T, record, device)
        return plan.execute(key, state0, shared, obs, use_graph)


class _Plan:
    """Pre-built launch arguments (and CUDA graph) for one (model, N, T) shape."""

    def __init__(self, pf: ParticleFilter, state0, shared, obs, T, record, device):
        self.pf = pf
        self.device = device
        self.T = T
        self.record = record
        n = pf.n
        step = pf.step
        specs = [ArgSpec("particle", "i32" if s.dtype == torch.int32 else "f32", tuple(s.shape[1:])) for s in state0]
        for s in shared:
            if isinstance(s, torch.Tensor):
                specs.append(ArgSpec("shared", "i32" if s.dtype == torch.int32 else "f32", tuple(s.shape)))
            else:
                specs.append(ArgSpec("scalar", "i32" if isinstance(s, int) else "f32", ()))
        self.cm = step.prebuild(specs, pf_obs=tuple(obs.keys()))  # filter variant: site flags baked in
        ir = self.cm.ir
        self.ir = ir
        # static buffers (graph replays read/write these)
        self.state_in = tuple(torch.empty_like(s) for s in state0)
        self.shared = tuple(torch.empty_like(s) if isinstance(s, torch.Tensor) else s for s in shared)
        self.obs = {a: torch.empty_like(v) for a, v in obs.items()}
        self.keys = torch.empty((T, 8), dtype=torch.int32, device=device)
        self.logw = torch.empty(n, dtype=torch.float32, device=device)
        self.anc = torch.empty((T if record else 2, n), dtype=torch.int32, device=device)
        self.lse = torch.empty((T, 3), dtype=torch.float64, device=device)
        self.ws = smc_ops.WeightWorkspace(n, device)
        self.wmax2 = torch.empty(2, dtype=torch.int32, device=device)
        # the next state = the model's return leaves (ping-pong buffers)
        rets = ir.ret_leaves
        if len(rets) != len(state0):
            raise ValueError(f"step returns {len(rets)} leaves but the state has {len(state0)}")
        self.bufs = []
        for r, s in zip(rets, state0):
            if not isinstance(r, Expr):
                raise ValueError("the step's return value must depend on its choices / arguments")
            want = torch.int32 if r.dtype == I32 else torch.float32
            if want != s.dtype or tuple(r.shape) != tuple(s.shape[1:]):
                raise ValueError("the step's return value must have the dtype and shape of the state it replaces")
            depth = T if record else 2
            self.bufs.append(torch.empty((depth,) + tuple(s.shape), dtype=s.dtype, device=device))
        self.logw_hist = torch.empty((T, n), dtype=torch.float32, device=device) if record else None
        self.final = tuple(torch.empty_like(s) for s in state0)
        self.obs_sites = {}
        for addr in obs:
            self.obs_sites[ir.site_index(addr)] = addr
        self.graph = None
        self.persistent = pf.mode == "persistent"
        tiles = (n + cabi.TE_TILE - 1) // cabi.TE_TILE
        step_ok = (tiles <= cabi.TE_MAX_TILES and pf.n_total == n and n <= (1 << 26) and pf.idx_offset % 4 == 0
                   and T < 65535 and self.cm.info.get("pf_step", False))
        # "steps": the same step body for ALL T steps in ONE cooperative launch (one grid barrier per step instead of a kernel
        # boundary); needs a co-resident CTA per 2048-slot window (1 M particles: 512 of 592 slots on a B200)
        self.stepsmode = pf.mode == "steps"
        if self.stepsmode and not (step_ok and self.cm.lib.gjb_model_pf_steps_fits(n)):
            raise NotImplementedError("mode='steps' needs a step-capable model on one device and a resident CTA per 2048 particles")
        self.stepmode = pf.mode in ("step", "steps") or (pf.mode == "auto" and step_ok)
        self.mode = ("steps" if self.stepsmode else "step") if self.stepmode else ("graph" if pf.mode == "auto" else pf.mode)
        if self.stepmode:
            if tiles > cabi.TE_MAX_TILES or pf.n_total > (1 << 26):
                raise NotImplementedError(f"mode='step' resamples over at most {cabi.TE_MAX_TILES} tiles of {cabi.TE_TILE} "
                                          "particles; use mode='graph' beyond that")
            if not self.cm.info.get("pf_step", False):
                raise NotImplementedError("mode='step': this model's return value does not feed back as its state")
            self.te_tiles = tiles
            import os

            # table form: the last CTA of each launch builds the tile-prefix table for the next (the form several devices
            # need); table-free form: every CTA forms the prefix from the plain tile records.  The redundant prefix work
            # of the table-free form grows with the SQUARE of the tile count, the last CTA's serial tail linearly.
            # Measured on one B200 (profiles/r2_call37/38_table_form_cost_1gpu.txt, us per step, table-free / table):
            # 1024 tiles 31.3 / 41.8, 1536 tiles 44.9 / 57.1, 2048 tiles 100.4 / 75.5, 4096 tiles 389.6 / 147.5 -- so the
            # table form is the default from 2048 tiles (4 194 304 particles) up.  Same ancestors either way.
            env = os.environ.get("GJB_STEP_TABLE")
            self.te_table = (env == "1") if env is not None else (tiles >= 2048 and not self.stepsmode)
            # ... and who builds the table: the step kernel's last CTA (default; B200: 22.5 us per step on one device, 28 us
            # on two) or gjb_te_table, a small kernel resident beside the step kernel (24.5 / 32 us: its launch and the
            # dependent launch behind it cost more than the last CTA's serial tail)
            self.te_table_kernel = os.environ.get("GJB_STEP_TABLE_KERNEL", "0") == "1"
            self.te_recs = torch.empty((2, tiles, 2), dtype=torch.int64, device=device)
            self.te_cdf = torch.empty((2, tiles * cabi.TE_TILE), dtype=torch.int64, device=device)
            # what the last CTA of each launch leaves for the next (one table per step parity), the tile-record
            # mailbox (single device: only this rank mails into it), run epoch and CTA ticket
            self.te_tables = torch.zeros((2, C.sizeof(cabi.StepTable)), dtype=torch.uint8, device=device)
            self.te_mailbox = torch.zeros(cabi.TE_MAILBOX_WORDS, dtype=torch.int64, device=device)
            self.te_epoch = torch.zeros(1, dtype=torch.int64, device=device)
            self.te_ticket = torch.zeros(1, dtype=torch.int32, device=device)
            L = cabi.StepLink()
            L.rank, L.world, L.tiles_per_rank = 0, 1, tiles
            L.mailbox[0] = self.te_mailbox.data_ptr()
            L.epoch, L.ticket = self.te_epoch.data_ptr(), self.te_ticket.data_ptr()
            self.te_link = torch.from_numpy(np.frombuffer(bytes(L), dtype=np.uint8).copy()).to(device)
        self.analytic = pf.reference_max == "analytic"
        if self.analytic:
            from ..gen import bounds

            from ..gen import codegen

            if codegen.group_lanes(ir.width):
                raise NotImplementedError("reference_max='analytic' exists for scalar-site (quad-mapped) models only: "
                                          "the lane-group kernels of vector-site models have no mass instantiation yet")
            self.bound_expr = bounds.log_weight_upper_bound(ir, sorted(self.obs_sites))
            if self.bound_expr is None:
                raise ValueError("reference_max='analytic': no particle-free bound of the observed sites' log-density "
                                 "can be derived for this model (gen/bounds.py)")
            self.m_ref = torch.empty(1, dtype=torch.float32, device=device)
            self.tm2 = torch.zeros((2, self.ws.tiles), dtype=torch.int64, device=device)
            self._m_ref_value = None
            self.single_pass = pf.single_pass
            if self.single_pass:
                if self.ws.tiles > 2048:
                    raise NotImplementedError("single_pass handles up to 2048 tiles (4 194 304 particles) per device")
                self.tm3 = torch.zeros((3, self.ws.tiles), dtype=torch.int64, device=device)  # accumulate / read / clear
                self.logw2 = torch.empty((2, n), dtype=torch.float32, device=device)
        if self.persistent:
            self._build_pf_args()
        elif self.stepmode:
            self._build_step_args()
            if self.stepsmode:
                self._build_steps_args()
        else:
            self._build_args()

    def _build_pf_args(self):
        """One ``gjb_pf_args`` for the persistent kernel."""
        pf, ir, T, n = self.pf, self.ir, self.T, self.pf.n
        lib = self.cm.lib
        grid = lib.gjb_model_pf_grid(n)
        cabi.check(min(grid, 0), "gjb_model_pf_grid")
        self.pf_grid = grid
        if self.record:
            self.logw = self.logw_hist
        self.cta_mass = torch.empty(grid, dtype=torch.int64, device=self.device)
        self.barrier = torch.zeros(2, dtype=torch.int32, device=self.device)
        Q = cabi.PfArgs()
        Q.n, Q.n_total, Q.idx_offset = n, pf.n_total, pf.idx_offset
        Q.T, Q.record, Q.n_state = T, int(self.record), len(self.state_in)
        Q.keys = self.keys.data_ptr()
        for i, s in enumerate(self.state_in):
            Q.state0[i] = s.data_ptr()
            Q.state_buf[i] = self.bufs[i].data_ptr()
            Q.state_stride[i] = self.bufs[i][0].numel() * 4
        for k, s in enumerate(self.shared):
            i = len(self.state_in) + k
            if isinstance(s, torch.Tensor):
                Q.shared[i] = s.data_ptr()
            else:
                Q.scalars[i] = float(s)
        for site in ir.sites:
            j = site.index
            if j in self.obs_sites:
                o = self.obs[self.obs_sites[j]]
                Q.obs[j] = o.data_ptr()
                Q.obs_stride[j] = (o[0].numel() if o.ndim > 1 else 1) * 4
                Q.site_flags[j] = cabi.SITE_WEIGHT | cabi.SITE_BCAST
            else:
                Q.site_flags[j] = cabi.SITE_SAMPLE
        Q.logw = self.logw.data_ptr()
        Q.ancestors = self.anc.data_ptr()
        Q.lse = self.lse.data_ptr()
        Q.wmax = self.wmax2.data_ptr()
        Q.cta_mass = self.cta_mass.data_ptr()
        Q.barrier = self.barrier.data_ptr()
        self.pf_args = Q

    def _build_step_args(self):
        """One ``gjb_step_args`` per step (ONE launch each) + the closing ``gjb_te_resample_args``."""
        import os

        pf, ir, T, n = self.pf, self.ir, self.T, self.pf.n
        if pf.idx_offset % 4:
            raise ValueError("mode='step' needs idx_offset % 4 == 0 (quad RNG streams)")
        if T >= 65535:
            raise ValueError("mode='step' tags its tile records with a 16-bit step number: at most 65534 steps per run")
        pdl = os.environ.get("GJB_PDL", "1") != "0"
        self.sargs = []
        self.targs = []
        for t in range(T):
            A = cabi.StepArgs()
            A.n, A.n_total, A.idx_offset, A.slot_offset = n, n, pf.idx_offset, 0
            A.step = t
            # programmatic dependent launch behind the previous step kernel: this launch draws its random numbers while
            # that one drains (the first step follows copies / other kernels and is launched normally)
            A.flags = (cabi.STEP_PDL if (pdl and t > 0) else 0) | (cabi.STEP_FLAGWAIT if os.environ.get("GJB_STEP_FLAGWAIT") == "1" else 0)
            A.key_dev = self.keys[t].data_ptr()
            slot = t if self.record else (t & 1)
            prev_slot = (t - 1) if self.record else ((t - 1) & 1)
            for i in range(len(self.state_in)):
                A.args[i] = self.state_in[i].data_ptr() if t == 0 else self.bufs[i][prev_slot].data_ptr()
                A.state_out[i] = self.bufs[i][slot].data_ptr()
            for k, s in enumerate(self.shared):
                i = len(self.state_in) + k
                if isinstance(s, torch.Tensor):
                    A.args[i] = s.data_ptr()
                else:
                    A.scalars[i] = float(s)
            for j, addr in self.obs_sites.items():
                A.site_in[j] = self.obs[addr][t].data_ptr()
            # the log-weights only leave the kernel when somebody reads them: every step when recording, else the last
            if self.record:
                A.weight_out = self.logw_hist[t].data_ptr()
            elif t == T - 1:
                A.weight_out = self.logw.data_ptr()
            if t > 0:
                A.prev_cdf = self.te_cdf[(t - 1) & 1].data_ptr()
                A.prev_key = self.keys[t - 1][2:].data_ptr()
                if self.te_table:
                    A.table_in = self.te_tables[(t - 1) & 1].data_ptr()
                else:
                    A.prev_recs = self.te_recs[(t - 1) & 1].data_ptr()
                    A.n_tiles_total = self.te_tiles
                    A.prev_lse = self.lse[t - 1].data_ptr()
                if self.record:
                    A.ancestors_out = self.anc[t - 1].data_ptr()
            A.cdf_out = self.te_cdf[t & 1].data_ptr()
            if self.te_table:
                A.link = self.te_link.data_ptr()
                if not self.te_table_kernel:  # the step kernel's last CTA builds the table
                    A.table_out = self.te_tables[t & 1].data_ptr()
                    A.lse_out = self.lse[t].data_ptr()
                B = cabi.TeTableArgs()
                B.link, B.step, B.flags = self.te_link.data_ptr(), t, (cabi.STEP_PDL if pdl else 0)
                B.slot_offset, B.n_local, B.n_total = 0, n, n
                B.reskey = self.keys[t][2:].data_ptr()
                B.table_out = self.te_tables[t & 1].data_ptr()
                B.lse_out = self.lse[t].data_ptr()
                self.targs.append(B)
            else:
                A.recs_out = self.te_recs[t & 1].data_ptr()
            self.sargs.append(A)
        last = (T - 1) if self.record else ((T - 1) & 1)
        R = cabi.TeResampleArgs()
        R.cdf = self.te_cdf[(T - 1) & 1].data_ptr()
        if self.te_table:
            R.table = self.te_tables[(T - 1) & 1].data_ptr()
        else:
            R.recs = self.te_recs[(T - 1) & 1].data_ptr()
            R.lse_out = self.lse[T - 1].data_ptr()
        R.n_tiles_total, R.n_total, R.out_lo, R.out_n = self.te_tiles, n, 0, n
        R.key_dev = self.keys[T - 1][2:].data_ptr()
        R.ancestors = self.anc[last].data_ptr()
        self.te_close = R

    def _build_steps_args(self):
        """One ``gjb_steps_args`` for the cooperative all-steps launch (the closing resampling is ``self.te_close``)."""
        pf, ir, T, n = self.pf, self.ir, self.T, self.pf.n
        if self.te_table:
            raise NotImplementedError("mode='steps' is the table-free form")
        Q = cabi.StepsArgs()
        Q.n, Q.idx_offset, Q.T, Q.record = n, pf.idx_offset, T, int(self.record)
        Q.keys = self.keys.data_ptr()
        for i, s in enumerate(self.state_in):
            Q.state0[i] = s.data_ptr()
            Q.state_buf[i] = self.bufs[i].data_ptr()
            Q.state_stride[i] = self.bufs[i][0].numel() * 4
        for k, s in enumerate(self.shared):
            i = len(self.state_in) + k
            if isinstance(s, torch.Tensor):
                Q.shared[i] = s.data_ptr()
            else:
                Q.scalars[i] = float(s)
        for j, addr in self.obs_sites.items():
            o = self.obs[addr]
            Q.obs[j] = o.data_ptr()
            Q.obs_stride[j] = (o[0].numel() if o.ndim > 1 else 1) * 4
        Q.logw = (self.logw_hist if self.record else self.logw).data_ptr()
        if self.record:
            Q.ancestors = self.anc.data_ptr()
        Q.cdf = self.te_cdf.data_ptr()
        Q.recs = self.te_recs.data_ptr()
        Q.lse = self.lse.data_ptr()
        self.steps_args = Q

    def _build_args(self):
        pf, ir, T, n = self.pf, self.ir, self.T, self.pf.n
        self.margs = []
        self.rargs = []
        for t in range(T):
            A = cabi.ModelArgs()
            A.n = n
            A.idx_offset = pf.idx_offset
            A.key_dev = self.keys[t].data_ptr()
            slot = t if self.record else (t & 1)
            prev_slot = (t - 1) if self.record else ((t - 1) & 1)
            for i in range(len(self.state_in)):
                A.args[i] = self.state_in[i].data_ptr() if t == 0 else self.bufs[i][prev_slot].data_ptr()
            if t > 0:
                A.gather = self.anc[prev_slot].data_ptr()
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
            # outputs: only what the next step needs
            for k, r in enumerate(ir.ret_leaves):
                out = self.bufs[k][slot]
                if r.op == "site" and not (r.attr in self.obs_sites):
                    A.site_out[r.attr] = out.data_ptr()
                else:
                    A.ret_out[k] = out.data_ptr()
            lw = self.logw_hist[t] if self.record else self.logw
            A.weight_out = lw.data_ptr()
            if self.analytic and self.single_pass:
                lw = self.logw_hist[t] if self.record else self.logw2[t & 1]
                A.weight_out = lw.data_ptr()
                A.gather = None  # the kernel gathers through the ancestors it resolves itself
                A.m_ref = self.m_ref.data_ptr()
                A.tile_mass = self.tm3[t % 3].data_ptr()
                A.tile_mass_clear = self.tm3[(t + 1) % 3].data_ptr()
                A.tile_mass_clear_n = self.ws.tiles
                A.pull_n_total = pf.n_total
                if t == 0:
                    A.pull_ancestors = self.anc[slot].data_ptr()  # selects the single-pass kernel; nothing to resample yet
                else:
                    lw_prev = self.logw_hist[t - 1] if self.record else self.logw2[(t - 1) & 1]
                    A.pull_logw = lw_prev.data_ptr()
                    A.pull_tile_mass = self.tm3[(t - 1) % 3].data_ptr()
                    A.pull_m_ref = self.m_ref.data_ptr()
                    A.pull_key = self.keys[t - 1][2:].data_ptr()
                    A.pull_ancestors = self.anc[prev_slot].data_ptr()
                    A.pull_lse = self.lse[t - 1].data_ptr()
                self.margs.append(A)
                if t == T - 1:  # the run ends with a plain resampling of the last step's weights
                    R = self.ws.systematic_args(
                        lw, None, self.anc[slot], n_total=pf.n_total, out_lo=pf.idx_offset, anc_base=pf.idx_offset,
                        key_dev=self.keys[t][2:], lse_out=self.lse[t], m_global=self.m_ref,
                    )
                    R.tile_mass = self.tm3[t % 3].data_ptr()
                    self.rargs.append((lw, R))
                continue
            if self.analytic:
                # masses relative to the analytic bound, accumulated by the model kernel into this step's tile buffer;
                # the other buffer (next step's) is zeroed by the same launch
                A.m_ref = self.m_ref.data_ptr()
                A.tile_mass = self.tm2[t & 1].data_ptr()
                A.tile_mass_clear = self.tm2[(t + 1) & 1].data_ptr()
                A.tile_mass_clear_n = self.ws.tiles
                self.margs.append(A)
                R = self.ws.systematic_args(
                    lw, None, self.anc[slot], n_total=pf.n_total, out_lo=pf.idx_offset, anc_base=pf.idx_offset,
                    key_dev=self.keys[t][2:], lse_out=self.lse[t], m_global=self.m_ref,
                )
                R.tile_mass = self.tm2[t & 1].data_ptr()
                self.rargs.append((lw, R))
                continue
            A.wmax = self.wmax2[t & 1 :].data_ptr()
            self.margs.append(A)
            if pf.resampler == "multinomial":
                self.rargs.append((lw, None))
                continue
            R = self.ws.systematic_args(
                lw, None, self.anc[slot], n_total=pf.n_total, out_lo=pf.idx_offset, anc_base=pf.idx_offset,
                key_dev=self.keys[t][2:], lse_out=self.lse[t], wmax=self.wmax2[t & 1 :],
                wmax_next=self.wmax2[(t + 1) & 1 :],
            )
            self.rargs.append((lw, R))

    def _enqueue(self):
        """Enqueue one filter run.  With ``GJB_NVTX=1`` the run and its phases are bracketed by NVTX ranges (``pf.run`` >
        ``pf.steps`` / ``pf.close``) for Nsight Systems timelines; ranges are host-side markers and cost nothing on the device."""
        import os

        if os.environ.get("GJB_NVTX") == "1":
            torch.cuda.nvtx.range_push(f"pf.run[{self.mode}] n={self.pf.n} T={self.T}")
            try:
                return self._enqueue_impl()
            finally:
                torch.cuda.nvtx.range_pop()
        return self._enqueue_impl()

    def _enqueue_impl(self):
        core = cabi.core()
        stream = cabi.stream_ptr(self.device)
        lib = self.cm.lib
        nvtx = __import__("os").environ.get("GJB_NVTX") == "1"
        if self.persistent:
            cabi.check(lib.gjb_model_pf_run(C.byref(self.pf_args), stream), "gjb_model_pf_run")
            last = (self.T - 1) if self.record else ((self.T - 1) & 1)
            for k in range(len(self.bufs)):
                smc_ops.gather_rows(self.bufs[k][last], self.anc[last], self.final[k])
            return
        if self.stepmode:  # ONE launch per step, the closing resampling of the last step, the final gather(s)
            if nvtx:
                torch.cuda.nvtx.range_push("pf.steps: resample(t-1) + gather + propose + logpdf + masses(t)")
            for t in range(0 if self.stepsmode else self.T):  # (mode "steps": one cooperative launch below instead)
                cabi.check(lib.gjb_model_pf_step(C.byref(self.sargs[t]), stream), "gjb_model_pf_step")
                if self.te_table and self.te_table_kernel:
                    cabi.check(core.gjb_te_table(C.byref(self.targs[t]), stream), "gjb_te_table")
            if self.stepsmode:
                cabi.check(lib.gjb_model_pf_steps(C.byref(self.steps_args), stream), "gjb_model_pf_steps")
            if nvtx:
                torch.cuda.nvtx.range_pop()
                torch.cuda.nvtx.range_push("pf.close: resample(T-1) + final gather")
            cabi.check(core.gjb_te_resample(C.byref(self.te_close), stream), "gjb_te_resample")
            last = (self.T - 1) if self.record else ((self.T - 1) & 1)
            for k in range(len(self.bufs)):
                smc_ops.gather_rows(self.bufs[k][last], self.anc[last], self.final[k])
            cabi.check(core.gjb_epoch_bump(self.te_epoch.data_ptr(), stream), "gjb_epoch_bump")  # new record tags next run
            if nvtx:
                torch.cuda.nvtx.range_pop()
            return
        if self.analytic and self.single_pass:  # 1 launch per step + one closing resampling launch
            self.tm3.zero_()
            for t in range(self.T):
                cabi.check(lib.gjb_model_launch(C.byref(self.margs[t]), stream), "gjb_model_launch")
            cabi.check(core.gjb_resample_systematic(C.byref(self.rargs[-1][1]), stream), "gjb_resample_systematic")
            last = (self.T - 1) if self.record else ((self.T - 1) & 1)
            for k in range(len(self.bufs)):
                smc_ops.gather_rows(self.bufs[k][last], self.anc[last], self.final[k])
            return
        if self.analytic:  # 2 launches per step: model kernel (weights + masses), resampler on the given masses
            self.tm2.zero_()
            for t in range(self.T):
                cabi.check(lib.gjb_model_launch(C.byref(self.margs[t]), stream), "gjb_model_launch")
                cabi.check(core.gjb_resample_systematic(C.byref(self.rargs[t][1]), stream), "gjb_resample_systematic")
            last = (self.T - 1) if self.record else ((self.T - 1) & 1)
            for k in range(len(self.bufs)):
                smc_ops.gather_rows(self.bufs[k][last], self.anc[last], self.final[k])
            return
        if self.pf.resampler == "multinomial":
            # per step: model launch (weights + running max) | exact integer tile masses | log-mean-exp | CDF + one
            # inverse-CDF search per offspring, offspring j on lane j of split(k_res, N)
            if getattr(self, "mn_cdf", None) is None:
                self.mn_cdf = torch.empty(self.pf.n, dtype=torch.int64, device=self.device)
            n = self.pf.n
            for t in range(self.T):
                wm = self.wmax2[t & 1 :]
                cabi.check(core.gjb_wmax_reset(wm.data_ptr(), stream), "gjb_wmax_reset")
                cabi.check(lib.gjb_model_launch(C.byref(self.margs[t]), stream), "gjb_model_launch")
                lw = self.rargs[t][0]
                cabi.check(core.gjb_weight_mass(lw.data_ptr(), n, wm.data_ptr(), None, self.ws.tile_mass.data_ptr(), stream), "gjb_weight_mass")
                cabi.check(core.gjb_lse_finalize(self.ws.tile_mass.data_ptr(), n, wm.data_ptr(), None, n, self.lse[t].data_ptr(), stream),
                           "gjb_lse_finalize")
                slot = t if self.record else (t & 1)
                cabi.check(core.gjb_resample_multinomial_keydev(lw.data_ptr(), n, wm.data_ptr(), self.ws.tile_mass.data_ptr(),
                                                                self.mn_cdf.data_ptr(), self.keys[t][6:].data_ptr(), 0, n,
                                                                self.anc[slot].data_ptr(), stream), "gjb_resample_multinomial_keydev")
            last = (self.T - 1) if self.record else ((self.T - 1) & 1)
            for k in range(len(self.bufs)):
                smc_ops.gather_rows(self.bufs[k][last], self.anc[last], self.final[k])
            return
        cabi.check(core.gjb_wmax_reset(self.wmax2.data_ptr(), stream), "gjb_wmax_reset")
        fused = getattr(self, "fuse_mass_resample", None)
        if fused is None:
            # mass + resample as one cooperative launch when every tile's CTA is resident (2 launches per step)
            fused = self.fuse_mass_resample = bool(core.gjb_mass_resample_fits(self.pf.n)) and self.pf.fuse_mass_resample
        for t in range(self.T):
            cabi.check(lib.gjb_model_launch(C.byref(self.margs[t]), stream), "gjb_model_launch")
            lw, R = self.rargs[t]
            if fused:
                cabi.check(core.gjb_mass_resample_systematic(C.byref(R), stream), "gjb_mass_resample_systematic")
                continue
            cabi.check(
                core.gjb_weight_mass(lw.data_ptr(), lw.numel(), self.wmax2[t & 1 :].data_ptr(), None,
                                     self.ws.tile_mass.data_ptr(), stream),
                "gjb_weight_mass",
            )
            cabi.check(core.gjb_resample_systematic(C.byref(R), stream), "gjb_resample_systematic")
        last = (self.T - 1) if self.record else ((self.T - 1) & 1)
        for k in range(len(self.bufs)):
            smc_ops.gather_rows(self.bufs[k][last], self.anc[last], self.final[k])

    def launches_per_run(self) -> int:
        if self.persistent:
            return 2 + len(self.bufs)  # init + persistent filter kernel + final gather(s)
        if self.stepmode:
            return (1 if self.stepsmode else self.T) + 2 + len(self.bufs)  # step kernel(s), closing resample, final gather(s), epoch bump
        if self.analytic and self.single_pass:
            return self.T + 1 + len(self.bufs)  # (+ one memset node)
        if self.analytic:
            return 2 * self.T + len(self.bufs)  # (+ one memset node)
        if self.pf.resampler == "multinomial":
            return 6 * self.T + len(self.bufs)
        return 1 + (2 if getattr(self, "fuse_mass_resample", False) else 3) * self.T + len(self.bufs)

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
        if self.analytic:
            from ..gen import bounds

            values = {}
            for k, sh in enumerate(shared):
                values[len(self.state_in) + k] = sh.detach().cpu().numpy() if isinstance(sh, torch.Tensor) else sh
            m = bounds.evaluate_invariant(self.bound_expr, values)
            if m != self._m_ref_value:
                self.m_ref.copy_(torch.tensor([m], dtype=torch.float32), non_blocking=True)
                self._m_ref_value = m
        if use_graph and not self.persistent:
            if self.graph is None:
                # warm-up launch outside capture (module load), then capture once
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
        hist = None
        if self.record:
            hist = {"state": tuple(b for b in self.bufs), "log_weights": self.logw_hist}
        return PFResult(
            state=self.final,
            log_marginal_likelihood=inc.sum(),
            log_increments=inc,
            lse_terms=self.lse,
            ancestors=self.anc if self.record else None,
            history=hist,
        )

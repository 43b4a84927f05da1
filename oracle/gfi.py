"""Static-language GFI restatement, vectorised over particles (oracle only).

Follows src/genjax/_src/generative_functions/static.py:
  * sites are visited in program order; site i (1-based, counting EVERY
    ``trace_p`` site whether constrained or not) draws from the key
    ``fold_in(key, i)`` -- SimulateHandler :254-278, GenerateHandler :341-380,
    UpdateHandler :407-466 -- here ``site=i`` in the Philox counter;
  * simulate: score = sum of site scores (StaticTrace.get_score :102-105);
  * assess:   score accumulates ``logpdf(v)`` of every site (:310-321), a
    missing address raises ``MissingAddress`` (:317-318);
  * generate/importance: unconstrained site -> sample, w += 0; constrained
    site -> w += logpdf (:368-380 with distribution.py:117-147);
  * update:   every site is re-scored under the new args, w += new - old
    score; constrained sites take the new value and the old one goes to the
    discard (distribution.py:179-244);
  * regenerate: selected site -> fresh sample, w += new - old score
    (distribution.py:258-277); unselected -> old value re-scored (:279-298).

An oracle model is a plain function ``model(h, *args)`` calling
``h.site(addr, dist_name, *dist_args)`` (or the sugar ``h.normal(addr, ...)``)
and returning its retval; all values are NumPy arrays with a leading particle
axis (or scalars that broadcast).
"""

from __future__ import annotations

import numpy as np

from . import dists, rng

F32 = np.float32


class MissingAddress(Exception):
    pass


class AddressReuse(Exception):
    pass


class OTrace:
    def __init__(self, model, args, retval, choices, scores, n, flags=None, site_rec=None):
        self.model = model
        self.args = args
        self.retval = retval
        self.choices = choices  # addr -> values [n, ...]
        self.scores = scores  # addr -> float32 [n]
        self.n = n
        self.flags = flags or {}  # addr -> bool [n]: validity of the sites under a Switch branch / Mask
        self.site_rec = site_rec or {}  # (addr, branch tags) -> (values, flag, score) of every such site

    def get_score(self):
        tot = np.zeros(self.n, dtype=F32)
        for s in self.scores.values():
            tot = (tot + s).astype(F32)
        return tot

    def get_choices(self):
        return dict(self.choices)

    def get_retval(self):
        return self.retval

    def take(self, idx):
        """tree_map(lambda v: v[idx]) -- ParticleCollection.get_particle (smc.py:90-91)."""
        idx = np.asarray(idx)

        def g(v):
            v = np.asarray(v)
            return v[idx] if v.ndim >= 1 and v.shape[0] == self.n else v

        n = int(idx.shape[0]) if idx.ndim else 1
        args = tuple(g(a) for a in self.args)
        rv = g(self.retval) if self.retval is not None else None
        return OTrace(
            self.model,
            args,
            rv,
            {k: g(v) for k, v in self.choices.items()},
            {k: g(v) for k, v in self.scores.items()},
            n,
            {k: g(v) for k, v in self.flags.items()},
            {k: tuple(g(x) for x in v) for k, v in self.site_rec.items()},
        )


class _Handler:
    def __init__(self, words, idx):
        self.words = words
        self.idx = idx
        self.n = len(idx)
        self.counter = 0
        self.choices = {}
        self.scores = {}
        self.weight = np.zeros(self.n, dtype=F32)
        # dynamic structure (combinators/switch.py, mask.py): where the current call exists / is scored
        self.live = None  # bool [n]: the enclosing Switch branches are the selected ones
        self.scored = None  # bool [n]: the enclosing MaskCombinator flags hold
        self.branch = ()  # ((switch id, branch), ...): which sites may share an address
        self.flags = {}  # addr -> bool [n] validity of a dynamic site
        self.site_rec = {}  # (addr, branch tags) -> (values, flag, score) per dynamic site
        self._excl = {}  # addr -> branch tags recorded there
        self._switches = 0

    def _gate(self):
        g = None
        for m in (self.live, self.scored):
            if m is not None:
                g = m if g is None else (g & m)
        return g

    def _w_add(self, lp, on=None):
        """weight += lp where the site is valid (and ``on`` holds)."""
        g = self._gate()
        if on is not None:
            g = on if g is None else (g & on)
        new = (self.weight + lp).astype(F32)
        self.weight = new if g is None else np.where(g, new, self.weight)

    def _record(self, addr, v, score):
        score = np.broadcast_to(np.asarray(score, dtype=F32), (self.n,)).copy()
        g = self._gate()
        if g is None and not self.branch:
            if addr in self.choices:
                raise AddressReuse(addr)
            self.choices[addr] = v
            self.scores[addr] = score
            return v
        # a site under a Switch branch / Mask: zeros outside the selected branch (switch.py:171-180: the sub-traces of
        # the unselected branches are zero-filled), score only where valid (mask.py:84: score = check * inner score)
        v = np.asarray(v)
        if self.live is not None:
            v = np.where(self.live.reshape((self.n,) + (1,) * (v.ndim - 1)), np.broadcast_to(v, (self.n,) + v.shape[1:]), 0).astype(v.dtype)
        score = np.where(g, score, F32(0)) if g is not None else score
        for other in self._excl.get(addr, ()):
            if not any(i == j and b != c for i, b in self.branch for j, c in other):
                raise AddressReuse(addr)
        self._excl.setdefault(addr, []).append(self.branch)
        flag = g if g is not None else np.ones(self.n, dtype=bool)
        self.site_rec[(addr, self.branch)] = (v, flag, score)
        if addr in self.choices:  # the same address in another branch of one Switch: the valid side wins (Mask.__or__)
            old_f = self.flags[addr]
            self.choices[addr] = np.where(old_f.reshape((self.n,) + (1,) * (v.ndim - 1)), self.choices[addr], v)
            self.scores[addr] = (self.scores[addr] + score).astype(F32)
            self.flags[addr] = old_f | flag
        else:
            self.choices[addr] = v
            self.scores[addr] = score
            self.flags[addr] = flag
        return v

    # ---- combinators
    def switch(self, idx, branches, branch_args):
        """``Switch(*branches)(idx, *branch_args)`` (switch.py:160-181): every branch is visited (so the site counter
        runs over all of them), the selected one's sites exist; returns the selected return value."""
        assert len(branches) == len(branch_args)
        k = np.clip(np.broadcast_to(np.asarray(idx).astype(np.int64), (self.n,)), 0, len(branches) - 1)
        self._switches += 1
        sid = self._switches
        rets = []
        for i, (f, a) in enumerate(zip(branches, branch_args)):
            old_live, old_branch = self.live, self.branch
            sel = k == i
            self.live = sel if old_live is None else (old_live & sel)
            self.branch = old_branch + ((sid, i),)
            try:
                rets.append(f(self, *a))
            finally:
                self.live, self.branch = old_live, old_branch
        return _tree_choose(k, rets, self.n)

    def mask(self, check, f, *args):
        """``MaskCombinator(f)(check, *args)`` (mask.py:158-165) -> (retval, flag)."""
        c = np.broadcast_to(np.asarray(check), (self.n,)) != 0
        old = self.scored
        self.scored = c if old is None else (old & c)
        try:
            ret = f(self, *args)
        finally:
            self.scored = old
        return ret, c

    def _logpdf(self, dist, v, args):
        lp = dists.DISTS[dist][1](v, *args)
        return np.broadcast_to(np.asarray(lp, dtype=F32), (self.n,))

    def _sample(self, dist, args):
        return dists.DISTS[dist][0](self.words, self.idx, self.counter, *args)

    def site(self, addr, dist, *args):
        self.counter += 1
        return self.handle(addr, dist, args)

    def __getattr__(self, name):
        if name in dists.DISTS:
            return lambda addr, *args: self.site(addr, name, *args)
        raise AttributeError(name)


def _tree_choose(k, rets, n):
    if isinstance(rets[0], (tuple, list)):
        return type(rets[0])(_tree_choose(k, [r[j] for r in rets], n) for j in range(len(rets[0])))
    if rets[0] is None:
        return None
    arrs = [np.asarray(r) for r in rets]
    dt = np.result_type(*[a.dtype for a in arrs])
    if dt == np.float64:
        dt = F32
    shape = np.broadcast_shapes(*[a.shape[1:] if (a.ndim and a.shape[0] == n) else a.shape for a in arrs])
    full = [np.broadcast_to(a.astype(dt), (n,) + shape) if not (a.ndim and a.shape[0] == n) else a.astype(dt) for a in arrs]
    out = full[-1]
    for i in range(len(full) - 2, -1, -1):
        out = np.where((k == i).reshape((n,) + (1,) * len(shape)), full[i], out)
    return out


def _is_masked(v):
    return isinstance(v, tuple) and len(v) == 3 and v[0] == "mask"


class _Simulate(_Handler):
    def handle(self, addr, dist, args):
        with np.errstate(all="ignore"):
            v = self._sample(dist, args)
            lp = self._logpdf(dist, v, args)
        return self._record(addr, v, lp)


class _Assess(_Handler):
    def __init__(self, chm, n):
        super().__init__((0, 0), np.zeros(n, dtype=np.uint64))
        self.chm = chm

    def handle(self, addr, dist, args):
        if addr not in self.chm:
            raise MissingAddress(addr)
        v = self.chm[addr]
        if _is_masked(v):  # assess scores the wrapped value whatever the flag says (distribution.py:404-417)
            v = v[1]
        with np.errstate(all="ignore"):
            lp = self._logpdf(dist, v, args)
        v = self._record(addr, v, lp)
        self._w_add(lp)
        return v


class _Generate(_Handler):
    def __init__(self, words, idx, chm):
        super().__init__(words, idx)
        self.chm = chm

    def handle(self, addr, dist, args):
        with np.errstate(all="ignore"):
            if addr in self.chm and _is_masked(self.chm[addr]):
                # Mask-ed constraint (distribution.py:129-142): constrained where the flag holds, drawn elsewhere
                _, cv, flag = self.chm[addr]
                flag = np.broadcast_to(np.asarray(flag), (self.n,)) != 0
                sv = np.asarray(self._sample(dist, args))
                cv = np.broadcast_to(np.asarray(cv).astype(sv.dtype), sv.shape)
                v = np.where(flag.reshape((self.n,) + (1,) * (sv.ndim - 1)), cv, sv)
                lp = self._logpdf(dist, v, args)
                self._w_add(lp, on=flag)
            elif addr in self.chm:
                v = self.chm[addr]
                lp = self._logpdf(dist, v, args)
                self._w_add(lp)
            else:
                v = self._sample(dist, args)
                lp = self._logpdf(dist, v, args)
        return self._record(addr, v, lp)


class _Update(_Handler):
    def __init__(self, words, idx, prev: OTrace, chm):
        super().__init__(words, idx)
        self.prev = prev
        self.chm = chm
        self.discard = {}

    def handle(self, addr, dist, args):
        rec = self.prev.site_rec.get((addr, self.branch))
        if rec is not None:
            return self._handle_dynamic(addr, dist, args, rec, addr in self.chm, False)
        old_v = self.prev.choices[addr]
        if addr in self.chm and _is_masked(self.chm[addr]):
            # distribution.py:190-226: the new value where the flag holds, the old one elsewhere; discard masked alike
            _, cv, flag = self.chm[addr]
            flag = np.broadcast_to(np.asarray(flag), (self.n,)) != 0
            ov = np.asarray(old_v)
            v = np.where(flag.reshape((self.n,) + (1,) * (ov.ndim - 1)), np.broadcast_to(np.asarray(cv).astype(ov.dtype), ov.shape), ov)
            self.discard[addr] = ("mask", old_v, flag)
        elif addr in self.chm:
            v = self.chm[addr]
            self.discard[addr] = old_v
        else:
            v = old_v
        lp = self._logpdf(dist, v, args)
        self.weight = (self.weight + (lp - self.prev.scores[addr]).astype(F32)).astype(F32)
        self._record(addr, v, lp)
        return v

    def _handle_dynamic(self, addr, dist, args, rec, constrained, selected):
        """A site under a Switch branch / Mask.  A branch that was not selected before holds no value: where it comes
        alive the site is drawn afresh (switch.py:226-246 generates the new branch with ``simulate``); the weight is
        the new valid score minus the old valid score (switch.py:299-300; mask.py:213-256: t->t the move's weight, t->f
        minus the old score, f->t the new score, f->f zero)."""
        old_v, old_flag, old_score = rec
        old_v = np.asarray(old_v)
        shp = (self.n,) + (1,) * (old_v.ndim - 1)
        with np.errstate(all="ignore"):
            if constrained:
                c = self.chm[addr]
                if _is_masked(c):
                    flag = np.broadcast_to(np.asarray(c[2]), (self.n,)) != 0
                    v = np.where(flag.reshape(shp), np.broadcast_to(np.asarray(c[1]).astype(old_v.dtype), old_v.shape), old_v)
                    self.discard[addr] = ("mask", old_v, flag)
                else:
                    v = c
                    self.discard[addr] = old_v
            elif selected:
                v = self._sample(dist, args)
                self.discard[addr] = old_v
            else:
                v = old_v
                if self.live is not None:
                    fresh = np.asarray(self._sample(dist, args))
                    v = np.where(old_flag.reshape(shp), old_v, np.broadcast_to(fresh.astype(old_v.dtype), old_v.shape))
            lp = self._logpdf(dist, v, args)
        g = self._gate()
        glp = np.where(g, lp, F32(0)) if g is not None else lp
        self.weight = (self.weight + (glp - old_score).astype(F32)).astype(F32)
        return self._record(addr, v, lp)


class _Regenerate(_Update):
    def __init__(self, words, idx, prev: OTrace, selected):
        super().__init__(words, idx, prev, {})
        self.selected = set(selected)

    def handle(self, addr, dist, args):
        rec = self.prev.site_rec.get((addr, self.branch))
        if rec is not None:
            return self._handle_dynamic(addr, dist, args, rec, False, addr in self.selected)
        old_v = self.prev.choices[addr]
        if addr in self.selected:
            v = self._sample(dist, args)
            self.discard[addr] = old_v
        else:
            v = old_v
        lp = self._logpdf(dist, v, args)
        self.weight = (self.weight + (lp - self.prev.scores[addr]).astype(F32)).astype(F32)
        self._record(addr, v, lp)
        return v


def _n_of(key, args, chm=None):
    if isinstance(key, rng.KeyBatch):
        return key.n
    return 1


def simulate(model, key, args):
    words, idx = rng.lanes(key)
    h = _Simulate(words, idx)
    rv = model(h, *args)
    return OTrace(model, args, rv, h.choices, h.scores, h.n, h.flags, h.site_rec)


def assess(model, chm, args, n=1):
    h = _Assess(chm, n)
    rv = model(h, *args)
    return h.weight, rv


def generate(model, key, chm, args):
    words, idx = rng.lanes(key)
    h = _Generate(words, idx, chm)
    rv = model(h, *args)
    return OTrace(model, args, rv, h.choices, h.scores, h.n, h.flags, h.site_rec), h.weight


importance = generate


def update(model, key, trace: OTrace, chm, args=None):
    words, idx = rng.lanes(key)
    h = _Update(words, idx, trace, chm)
    args = trace.args if args is None else args
    rv = model(h, *args)
    return OTrace(model, args, rv, h.choices, h.scores, h.n, h.flags, h.site_rec), h.weight, h.discard


def regenerate(model, key, trace: OTrace, selected, args=None):
    words, idx = rng.lanes(key)
    h = _Regenerate(words, idx, trace, selected)
    args = trace.args if args is None else args
    rv = model(h, *args)
    return OTrace(model, args, rv, h.choices, h.scores, h.n, h.flags, h.site_rec), h.weight, h.discard


# --------------------------------------------------------------------------
# Scan combinator (src/genjax/_src/generative_functions/combinators/scan.py)
# --------------------------------------------------------------------------
#
# A scanned kernel is an oracle model ``kernel(h, carry, x) -> (carry, y)``.  The key chain is the reference's
# (``key_t = fold_in(key_{t-1}, t)``, scan.py:213, 268), applied lane-wise (rng.fold_in_lanes); scores and weights
# are summed over the steps in step order (``jnp.sum(scores)``, :236, 296); constraints are addressed per step
# (``constraint.get_submap(idx)``, :270): here ``chm_at(t)`` returns the step's ``{addr: values}`` dict.


def _x_at(xs, t):
    if xs is None:
        return None
    if isinstance(xs, (tuple, list)):
        return type(xs)(_x_at(x, t) for x in xs)
    if isinstance(xs, dict):
        return {k: _x_at(v, t) for k, v in xs.items()}
    return np.asarray(xs)[t]


def _scan_len(xs, length):
    if xs is None:
        return int(length)
    leaves = []

    def go(x):
        if isinstance(x, (tuple, list)):
            for y in x:
                go(y)
        elif isinstance(x, dict):
            for y in x.values():
                go(y)
        elif x is not None:
            leaves.append(np.asarray(x).shape[0])

    go(xs)
    assert len(set(leaves)) == 1 and (length is None or leaves[0] == length)
    return leaves[0]


def scan_simulate(kernel, key, carry, xs, length=None):
    """scan.py:199-238 -> (per-step traces, carry_out, [y_t], score)."""
    T = _scan_len(xs, length)
    traces, ys, score, k = [], [], None, key
    for t in range(T):
        k = rng.fold_in_lanes(k, t)
        tr = simulate(kernel, k, (carry, _x_at(xs, t)))
        carry, y = tr.get_retval()
        traces.append(tr)
        ys.append(y)
        score = tr.get_score() if score is None else (score + tr.get_score()).astype(F32)
    return traces, carry, ys, score


def scan_generate(kernel, key, chm_at, carry, xs, length=None):
    """scan.py:240-297 -> (per-step traces, carry_out, [y_t], score, weight)."""
    T = _scan_len(xs, length)
    traces, ys, score, weight, k = [], [], None, None, key
    for t in range(T):
        k = rng.fold_in_lanes(k, t)
        tr, w = generate(kernel, k, chm_at(t), (carry, _x_at(xs, t)))
        carry, y = tr.get_retval()
        traces.append(tr)
        ys.append(y)
        score = tr.get_score() if score is None else (score + tr.get_score()).astype(F32)
        weight = w if weight is None else (weight + w).astype(F32)
    return traces, carry, ys, score, weight


def scan_assess(kernel, chm_at, carry, xs, length=None, n=1):
    """scan.py:634-660 -> (score, carry_out, [y_t])."""
    T = _scan_len(xs, length)
    ys, score = [], None
    for t in range(T):
        s, (carry, y) = assess(kernel, chm_at(t), (carry, _x_at(xs, t)), n=n)
        ys.append(y)
        score = s if score is None else (score + s).astype(F32)
    return score, carry, ys


def scan_update(kernel, key, traces, chm_at, carry, xs, length=None):
    """scan.py:509-602: every step is re-visited with ``Update(constraint(t))`` and the new carry
    -> (per-step traces, carry_out, [y_t], score, weight, {t: discard})."""
    T = _scan_len(xs, length)
    new, ys, score, weight, k, discard = [], [], None, None, key, {}
    for t in range(T):
        k = rng.fold_in_lanes(k, t)
        tr, w, d = update(kernel, k, traces[t], chm_at(t), (carry, _x_at(xs, t)))
        carry, y = tr.get_retval()
        new.append(tr)
        ys.append(y)
        score = tr.get_score() if score is None else (score + tr.get_score()).astype(F32)
        weight = w if weight is None else (weight + w).astype(F32)
        if d:
            discard[t] = d
    return new, carry, ys, score, weight, discard


def scan_regenerate(kernel, key, traces, selected, carry, xs, length=None):
    """scan.py:417-507: ``Regenerate(selection)`` at every step."""
    T = _scan_len(xs, length)
    new, ys, score, weight, k = [], [], None, None, key
    for t in range(T):
        k = rng.fold_in_lanes(k, t)
        tr, w, _ = regenerate(kernel, k, traces[t], selected, (carry, _x_at(xs, t)))
        carry, y = tr.get_retval()
        new.append(tr)
        ys.append(y)
        score = tr.get_score() if score is None else (score + tr.get_score()).astype(F32)
        weight = w if weight is None else (weight + w).astype(F32)
    return new, carry, ys, score, weight

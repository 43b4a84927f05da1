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
    def __init__(self, model, args, retval, choices, scores, n):
        self.model = model
        self.args = args
        self.retval = retval
        self.choices = choices  # addr -> values [n, ...]
        self.scores = scores  # addr -> float32 [n]
        self.n = n

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

    def _record(self, addr, v, score):
        if addr in self.choices:
            raise AddressReuse(addr)
        self.choices[addr] = v
        self.scores[addr] = np.broadcast_to(np.asarray(score, dtype=F32), (self.n,)).copy()

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


class _Simulate(_Handler):
    def handle(self, addr, dist, args):
        v = self._sample(dist, args)
        self._record(addr, v, self._logpdf(dist, v, args))
        return v


class _Assess(_Handler):
    def __init__(self, chm, n):
        super().__init__((0, 0), np.zeros(n, dtype=np.uint64))
        self.chm = chm

    def handle(self, addr, dist, args):
        if addr not in self.chm:
            raise MissingAddress(addr)
        v = self.chm[addr]
        lp = self._logpdf(dist, v, args)
        self._record(addr, v, lp)
        self.weight = (self.weight + lp).astype(F32)
        return v


class _Generate(_Handler):
    def __init__(self, words, idx, chm):
        super().__init__(words, idx)
        self.chm = chm

    def handle(self, addr, dist, args):
        if addr in self.chm:
            v = self.chm[addr]
            lp = self._logpdf(dist, v, args)
            self.weight = (self.weight + lp).astype(F32)
        else:
            v = self._sample(dist, args)
            lp = self._logpdf(dist, v, args)
        self._record(addr, v, lp)
        return v


class _Update(_Handler):
    def __init__(self, words, idx, prev: OTrace, chm):
        super().__init__(words, idx)
        self.prev = prev
        self.chm = chm
        self.discard = {}

    def handle(self, addr, dist, args):
        old_v = self.prev.choices[addr]
        if addr in self.chm:
            v = self.chm[addr]
            self.discard[addr] = old_v
        else:
            v = old_v
        lp = self._logpdf(dist, v, args)
        self.weight = (self.weight + (lp - self.prev.scores[addr]).astype(F32)).astype(F32)
        self._record(addr, v, lp)
        return v


class _Regenerate(_Handler):
    def __init__(self, words, idx, prev: OTrace, selected):
        super().__init__(words, idx)
        self.prev = prev
        self.selected = set(selected)
        self.discard = {}

    def handle(self, addr, dist, args):
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
    return OTrace(model, args, rv, h.choices, h.scores, h.n)


def assess(model, chm, args, n=1):
    h = _Assess(chm, n)
    rv = model(h, *args)
    return h.weight, rv


def generate(model, key, chm, args):
    words, idx = rng.lanes(key)
    h = _Generate(words, idx, chm)
    rv = model(h, *args)
    return OTrace(model, args, rv, h.choices, h.scores, h.n), h.weight


importance = generate


def update(model, key, trace: OTrace, chm, args=None):
    words, idx = rng.lanes(key)
    h = _Update(words, idx, trace, chm)
    args = trace.args if args is None else args
    rv = model(h, *args)
    return OTrace(model, args, rv, h.choices, h.scores, h.n), h.weight, h.discard


def regenerate(model, key, trace: OTrace, selected, args=None):
    words, idx = rng.lanes(key)
    h = _Regenerate(words, idx, trace, selected)
    args = trace.args if args is None else args
    rv = model(h, *args)
    return OTrace(model, args, rv, h.choices, h.scores, h.n), h.weight, h.discard


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

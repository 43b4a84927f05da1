"""The N>1 path on CPU (gloo, world_size 2): the sharding arithmetic the multi-GPU
filter relies on -- block bounds, one max reduction, one all-gather of integer rank
masses giving every rank its CDF offset, owner-side offspring ranges -- reproduces the
unsharded resampler bit for bit, and every rank derives the same per-step keys."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import rng as orng
from oracle import smc as osmc

F32 = np.float32


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_total, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from genjax_b200.core.key import key as pkey, pf_key_table
        from genjax_b200.inference.pf_dist import shard_bounds

        lo, hi = shard_bounds(n_total, world, rank)
        g = np.random.default_rng(123)
        lw_all = (g.standard_normal(n_total) * 2.5).astype(F32)
        lw = lw_all[lo:hi]
        # exchange MAX
        m = torch.tensor([float(np.max(lw))], dtype=torch.float32)
        dist.all_reduce(m, op=dist.ReduceOp.MAX)
        M = F32(m.item())
        # exchange MASS: integer masses relative to the GLOBAL max
        q = osmc.det_exp_q((lw - M).astype(F32))
        mine = torch.tensor([int(q.sum(dtype=np.uint64))], dtype=torch.int64)
        masses = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(masses, mine)
        masses = [int(x.item()) for x in masses]
        c_offset, S = sum(masses[:rank]), sum(masses)
        # owner-side offspring ranges over the global index space
        k_res = osmc.pf_step_keys(orng.key(7), 3)[1]
        u0 = osmc.resample_u0(k_res)
        cnt, _ = osmc.systematic_counts(lw, u0, n_out=n_total, M=M, S=S, c_offset=c_offset)
        prev = np.concatenate([[0 if rank == 0 else None], cnt[:-1]]) if rank == 0 else None
        first = torch.tensor([int(cnt[-1])], dtype=torch.int64)  # my last cumulative count = next rank's start
        ends = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(ends, first)
        start = 0 if rank == 0 else int(ends[rank - 1].item())
        prev = np.concatenate([[start], cnt[:-1]])
        anc_global = np.repeat(np.arange(lo, hi, dtype=np.int32), (cnt - prev).astype(np.int64))
        # every rank's pieces concatenate to the unsharded ancestors
        pieces = [None] * world
        dist.all_gather_object(pieces, anc_global)
        full = np.concatenate(pieces)
        ref = osmc.resample_systematic(lw_all, k_res)
        ok = np.array_equal(full, ref)
        # identical key tables on every rank (no collective needed for keys)
        tab = pf_key_table(pkey(99), 5)
        tabs = [None] * world
        dist.all_gather_object(tabs, tab)
        ok = ok and all(np.array_equal(t, tab) for t in tabs)
        # the global log-mean-exp from the exchanged terms equals the unsharded one
        lme = float(M) + np.log(S) - osmc.Q_BITS * np.log(2.0) - np.log(n_total)
        ok = ok and abs(lme - osmc.log_mean_exp(lw_all)) < 1e-12
        out[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [8, 4096, 100_000])
def test_sharded_resampling_world2_gloo(n_total):
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), n_total, out), nprocs=world, join=True)
    assert dict(out) == {0: True, 1: True}


def test_shard_bounds_validation():
    from genjax_b200.inference.pf_dist import shard_bounds

    assert shard_bounds(1 << 20, 8, 3) == (3 << 17, 4 << 17)
    with pytest.raises(ValueError):
        shard_bounds(10, 4, 0)
    with pytest.raises(ValueError):
        shard_bounds(12, 2, 0)  # 6 per rank: not a multiple of 4

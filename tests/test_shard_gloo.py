"""Host-side logic of fibre-range sharding on CPU (two gloo ranks): the exchange steps of
pfs-neural-net_b200/shard.py -- statistics merge (count, mean, M2) with cached shard sizes, plain and packed
sums, the replicated() scope -- and the fibre-range partition of a general edge list.
The sharded Block itself is compared with the single-GPU Block in tests/test_gpu_wide_shard.py."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pfs_neural_net_b200 import shard


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(3)
    full = torch.randn(1000, 6, generator=g) * 3 + 40          # |mean| >> std: the merge must not cancel
    bounds = [0, 380, 1000]
    mine = full[bounds[rank]:bounds[rank + 1]]
    ok = True
    # outside the scope everything is the identity
    t = torch.ones(3)
    ok &= shard.allreduce_sum(t) is t and not shard.active() and shard.world_size() == 1
    with shard.fibre_sharded():
        ok &= shard.active() and shard.world_size() == world
        mean = mine.mean(0)
        m2 = ((mine - mean) ** 2).sum(0)
        n, gmean, gm2 = shard.allreduce_moments(float(mine.shape[0]), mean, m2)
        ok &= n == 1000.0
        ok &= torch.allclose(gmean, full.mean(0), rtol=1e-6)
        ok &= torch.allclose(gm2, ((full - full.mean(0)) ** 2).sum(0), rtol=1e-5)
        s = shard.allreduce_sum(mine.sum(0))
        ok &= torch.allclose(s, full.sum(0), rtol=1e-6)
        with shard.replicated():
            ok &= not shard.active()
            r = torch.full((2,), 5.0)
            ok &= torch.equal(shard.allreduce_sum(r), torch.full((2,), 5.0))     # class rows: no exchange
        calls, nbytes = shard.traffic()
        ok &= calls == 3 and nbytes > 0                       # shard sizes (once), the moment gather, the sum
        # a second merge of the same shard sizes costs ONE collective and no host synchronisation
        n2, gmean2, gm22 = shard.allreduce_moments(float(mine.shape[0]), mean, m2)
        ok &= shard.traffic()[0] == 4 and n2 == 1000.0 and torch.equal(gmean2, gmean) and torch.equal(gm22, gm2)
        ok &= shard.shard_counts(mine.shape[0]) == (380, 620) and shard.total_count(mine.shape[0]) == 1000
        # several small tensors, one collective
        a, b, c = mine.sum(0), mine[:, :2].sum(0).reshape(1, 2), (mine ** 2).sum(0)
        pa, pb, pc = shard.allreduce_packed([a, b, c])
        ok &= shard.traffic()[0] == 5 and pb.shape == (1, 2)
        ok &= torch.allclose(pa, full.sum(0), rtol=1e-6) and torch.allclose(pb[0], full[:, :2].sum(0), rtol=1e-6)
        ok &= torch.allclose(pc, (full ** 2).sum(0), rtol=1e-6)
    ok &= not shard.active()
    out[rank] = bool(ok)
    dist.destroy_process_group()


def test_shard_exchanges_two_ranks():
    world = 2
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
        assert dict(out) == {0: True, 1: True}


def test_sharding_requires_process_group():
    import pytest
    with pytest.raises(RuntimeError):
        with shard.fibre_sharded():
            pass


def test_partition_fibres_covers_the_edge_list_once():
    g = torch.Generator().manual_seed(5)
    S, T, world = 37, 6, 4
    keep = torch.rand(S * T, generator=g) < 0.3
    e = torch.nonzero(keep).flatten()
    e = e[torch.randperm(e.numel(), generator=g)]
    ei = torch.stack([e // T, e % T])
    seen = []
    for r in range(world):
        local, sl, pos = shard.partition_fibres(ei, S, world, r)
        assert local.shape == (2, pos.numel())
        assert torch.equal(local[0] + sl.start, ei[0][pos]) and torch.equal(local[1], ei[1][pos])
        assert (local[0] >= 0).all() and (local[0] < sl.stop - sl.start).all()
        assert torch.equal(pos, torch.sort(pos).values)          # the edge list's own order is kept
        seen.append(pos)
    allpos = torch.cat(seen)
    assert torch.equal(torch.sort(allpos).values, torch.arange(ei.shape[1]))

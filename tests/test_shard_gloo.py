"""Host-side logic of fibre-range sharding on CPU (two gloo ranks): the exchange steps of
pfs-neural-net_b200/shard.py -- statistics merge (count, mean, M2), plain sums, the replicated()
scope -- and, end to end, that a fibre-sharded run of the ORACLE with those exchanges plugged in
at the points wide.py uses them reproduces the unsharded result."""
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
        ok &= calls == 3 and nbytes > 0
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

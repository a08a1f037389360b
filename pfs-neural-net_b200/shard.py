"""Fibre-range sharding of ONE large graph over the ranks of a process group (BASELINE config C4;
SURVEY.md section 8e "Fibre-range edge sharding").

Rank r owns a contiguous range of fibres and all their edges (the dense layout is fibre-major, so
that is a contiguous slab of x_e); x_t, u and the weights are replicated.  Every fibre-side
statistic is then local, and the layer needs one exchange per reduction over "all edges / all
fibres": the BatchNorm statistics of the edge and source models, the class aggregate of the target
model (the tensor the north star names), the class-table gradients and the fibre-local parts of the
weight gradients.  They are all small fp32 tensors, all-reduced (sum) over NCCL/NVLink on the
compute stream; with no group active every function here is the identity.

Usage:   with shard.fibre_sharded(group):  y = block((edge_index_local, x_s_local, x_t, x_e_local, u))
"""
import contextlib

import torch
import torch.distributed as dist

_state = {"group": None, "active": False, "replicated": 0, "bytes": 0, "calls": 0}


def active():
    return _state["active"] and _state["replicated"] == 0


@contextlib.contextmanager
def fibre_sharded(group=None):
    """Run the enclosed module calls (forward AND the backward they record) as one fibre shard."""
    if not dist.is_initialized():
        raise RuntimeError("fibre_sharded() needs torch.distributed to be initialised")
    prev = (_state["group"], _state["active"])
    _state["group"], _state["active"] = group, True
    try:
        yield
    finally:
        _state["group"], _state["active"] = prev


@contextlib.contextmanager
def replicated():
    """Rows that every shard holds in full (classes, the global row): reductions stay local."""
    _state["replicated"] += 1
    try:
        yield
    finally:
        _state["replicated"] -= 1


def world_size():
    return dist.get_world_size(_state["group"]) if _state["active"] else 1


def allreduce_sum(t):
    """Sum of a (small, fp32) tensor over the shards; identity when not sharded."""
    if not active():
        return t
    t = t.contiguous()
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=_state["group"])
    _state["bytes"] += t.numel() * t.element_size()
    _state["calls"] += 1
    return t


def allreduce_moments(n, mean, m2):
    """Combine per-shard (count, mean, sum of squared deviations) into the global ones (Chan et al.);
    returns (n_total as float, mean, m2)."""
    if not active():
        return n, mean, m2
    # sum of n, n*mean, and m2 + n*mean^2 would cancel; exchange the three moments and merge exactly:
    # M2 = sum_r [ m2_r + n_r (mean_r - mean)^2 ]
    cnt = torch.full((1,), float(n), dtype=torch.float64, device=mean.device)
    s1 = mean.double() * float(n)
    pack = torch.cat([cnt, s1])
    dist.all_reduce(pack, op=dist.ReduceOp.SUM, group=_state["group"])
    n_tot = float(pack[0].item())
    g_mean = pack[1:] / n_tot
    dev = m2.double() + float(n) * (mean.double() - g_mean) ** 2
    dist.all_reduce(dev, op=dist.ReduceOp.SUM, group=_state["group"])
    _state["bytes"] += (pack.numel() + dev.numel()) * 8
    _state["calls"] += 2
    return n_tot, g_mean.float(), dev.float()


def traffic():
    """(collective calls, bytes) since the last reset -- bench.py reports them."""
    return _state["calls"], _state["bytes"]


def reset_traffic():
    _state["calls"] = _state["bytes"] = 0

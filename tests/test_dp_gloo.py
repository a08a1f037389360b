"""Host-side logic of the data-parallel path on CPU: two gloo ranks, flat gradient bucket all-reduce
with a missing gradient on one rank, graph sharding, buffer broadcast."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pfs_neural_net_b200 import dp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(4, 3), torch.nn.BatchNorm1d(3), torch.nn.Linear(3, 2))
    dp.broadcast_parameters(model)
    x = torch.full((5, 4), float(rank + 1))
    model(x).sum().backward()
    if rank == 1:
        model[2].bias.grad = None                      # a tensor without gradient on one rank only
    local = [None if p.grad is None else p.grad.clone() for p in model.parameters()]
    bucket = dp.GradBucket(model.parameters())
    bucket.all_reduce()
    gathered = [None] * world
    dist.all_gather_object(gathered, local)
    ok = True
    for i, p in enumerate(model.parameters()):
        expect = sum((g[i] if g[i] is not None else torch.zeros_like(p)) for g in gathered)
        ok &= torch.allclose(p.grad, expect, atol=1e-6)
    model[1].running_mean.fill_(float(rank))
    dp.broadcast_buffers(model, src=0)
    ok &= float(model[1].running_mean[0]) == 0.0
    out[rank] = bool(ok)
    dist.destroy_process_group()


def test_grad_bucket_allreduce_two_ranks():
    world = 2
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
        assert dict(out) == {0: True, 1: True}


def test_shard_graphs_partitions_every_graph_once():
    for world in (1, 2, 4, 8):
        seen = sorted(g for r in range(world) for g in dp.shard_graphs(256, r, world))
        assert seen == list(range(256))
        assert {len(dp.shard_graphs(256, r, world)) for r in range(world)} == {256 // world}


def test_single_process_bucket_is_a_noop():
    lin = torch.nn.Linear(3, 2)
    lin(torch.ones(1, 3)).sum().backward()
    g = lin.weight.grad.clone()
    dp.GradBucket(lin.parameters()).all_reduce()
    assert torch.equal(lin.weight.grad, g)


def _block_worker(rank, world, port, out):
    """DP over graphs with the Block's real parameter set: every rank runs its share of the graphs (the CPU oracle does the
    arithmetic on the module's own parameters, so `.grad` lands where the product path puts it), one flat-bucket
    all-reduce; the result must equal the single-process sum over all graphs."""
    from oracle import block_oracle as bo
    from pfs_neural_net_b200 import gnn
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    F, S, T, G = 4, 9, 5, 6
    torch.manual_seed(100 + rank)                        # different initial weights per rank ...
    blk = gnn.Block(F)
    dp.broadcast_parameters(blk)                         # ... until rank 0's are broadcast
    ei = bo.complete_bipartite(S, T)
    g = torch.Generator().manual_seed(5)
    ins = [torch.randn(G, S, F, generator=g), torch.randn(G, T, F, generator=g), torch.randn(G, S * T, F, generator=g),
           torch.randn(G, 1, F, generator=g)]
    ups = [torch.randn(t.shape, generator=g) for t in ins]

    def run(graphs):
        for p in blk.parameters():
            p.grad = None
        state = dict(blk.named_parameters())
        state.update(dict(blk.named_buffers()))
        for i in graphs:
            outs = bo.block(state, "", ei, ins[0][i], ins[1][i], ins[2][i], ins[3][i], training=True, buffers={})
            torch.autograd.backward(list(outs), [u[i] for u in ups])
        return {k: (None if p.grad is None else p.grad.clone()) for k, p in blk.named_parameters()}

    full = run(range(G))                                 # every rank can compute the whole batch: the expected sum
    run(dp.shard_graphs(G, rank, world))
    bucket = dp.GradBucket(blk.parameters())
    bucket.all_reduce()
    ok = True
    for k, p in blk.named_parameters():
        scale = max(full[k].abs().max().item(), 1e-6)
        ok &= (p.grad - full[k]).abs().max().item() <= 1e-4 * max(scale, full[k[:-4] + "weight"].abs().max().item()
                                                                  if k.endswith("bias") and ".norm." not in k else scale)
    out[rank] = bool(ok)
    dist.destroy_process_group()


def test_block_gradients_over_two_ranks_equal_the_single_process_sum():
    world = 2
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_block_worker, args=(world, _free_port(), out), nprocs=world, join=True)
        assert dict(out) == {0: True, 1: True}

"""Size-independent properties at BASELINE.json's full sizes, where the CPU oracle cannot follow:

* C4 shard (12 500 fibres x 512 classes = 6.4 M edges, Fdim 128, bf16, tensor-core path), one Block forward+backward:
  bit-exact run-to-run determinism (every reduction has a fixed order, no atomics) and **fibre-permutation
  equivariance** -- relabelling the fibres permutes x_s' and the edge rows of x_e' the same way and leaves x_t', u'
  and every parameter gradient unchanged (up to the reassociation of the sums over fibres).
* C5 (10 % Bernoulli edge list of 100 000 x 512, shuffled, int64 edge_index), Fdim 128 bf16: the CSR/CSC path gives the
  same result for two different orderings of the same edge set.
"""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _dev(min_gb):
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    if torch.cuda.get_device_properties(0).total_memory < min_gb * 2 ** 30:
        pytest.skip("needs %d GB of device memory" % min_gb)
    return torch.device("cuda:0")


def _block(F, dev):
    from pfs_neural_net_b200 import gnn
    torch.manual_seed(0)
    blk = gnn.Block(F)
    g = torch.Generator().manual_seed(1)
    for m in blk.modules():
        if isinstance(m, torch.nn.BatchNorm1d):
            m.weight.data = 0.5 + torch.rand(m.weight.shape, generator=g)
            m.bias.data = 2 * torch.rand(m.bias.shape, generator=g) - 1
    return blk.to(torch.bfloat16).to(dev).train()


def _run(blk, ei, ins, ups):
    for p in blk.parameters():
        p.grad = None
    xs = [t.detach().clone().requires_grad_(True) for t in ins]
    _, o_s, o_t, o_e, o_u = blk((ei, *xs))
    torch.autograd.backward([o_s, o_t, o_e, o_u], ups)
    torch.cuda.synchronize()
    grads = {k: p.grad.clone() for k, p in blk.named_parameters()}
    return (o_s, o_t, o_e, o_u), [x.grad for x in xs], grads


def _close(a, b, tol, what):
    a, b = a.float(), b.float()
    err = (a - b).abs().max().item() / max(b.abs().max().item(), 1e-30)
    assert err < tol, (what, err)


def test_c4_shard_determinism_and_fibre_permutation_equivariance():
    dev = _dev(60)
    S, T, F = 12500, 512, 128
    E = S * T
    blk = _block(F, dev)
    ei = torch.cartesian_prod(torch.arange(S), torch.arange(T)).T.contiguous().to(dev)
    gen = torch.Generator(device=dev).manual_seed(3)
    bf = torch.bfloat16
    ins = [torch.randn(S, F, generator=gen, device=dev).to(bf), torch.randn(T, F, generator=gen, device=dev).to(bf),
           torch.randn(E, F, generator=gen, device=dev).to(bf), torch.randn(1, F, generator=gen, device=dev).to(bf)]
    ups = [torch.randn(t.shape, generator=gen, device=dev).to(bf) for t in ins]
    state = {k: v.clone() for k, v in blk.state_dict().items()}
    out1, gin1, gp1 = _run(blk, ei, ins, ups)
    blk.load_state_dict(state)
    out2, gin2, gp2 = _run(blk, ei, ins, ups)
    for a, b in zip(out1 + tuple(gin1), out2 + tuple(gin2)):
        assert torch.equal(a, b)                                  # bit-exact
    for k in gp1:
        assert torch.equal(gp1[k], gp2[k]), k
    # relabel the fibres
    perm = torch.randperm(S, generator=torch.Generator().manual_seed(4)).to(dev)
    eperm = (perm[:, None] * T + torch.arange(T, device=dev)[None, :]).reshape(-1)
    ins_p = [ins[0][perm].contiguous(), ins[1], ins[2][eperm].contiguous(), ins[3]]
    ups_p = [ups[0][perm].contiguous(), ups[1], ups[2][eperm].contiguous(), ups[3]]
    del out2, gin2, gp2
    blk.load_state_dict(state)
    out3, gin3, gp3 = _run(blk, ei, ins_p, ups_p)
    tol = 2e-2                                                    # bf16 outputs, sums over fibres reassociated
    _close(out3[0], out1[0][perm], tol, "x_s")
    _close(out3[2], out1[2][eperm], tol, "x_e")
    _close(out3[1], out1[1], tol, "x_t")
    _close(out3[3], out1[3], tol, "u")
    _close(gin3[0], gin1[0][perm], 6e-2, "grad x_s")
    _close(gin3[2], gin1[2][eperm], 6e-2, "grad x_e")
    _close(gin3[1], gin1[1], 6e-2, "grad x_t")
    for k in gp1:
        scale = max(gp1[k].float().abs().max().item(),
                    gp1[k[:-4] + "weight"].float().abs().max().item() if k.endswith("bias") else 0.0)
        # parameter gradients: sums over 6.4 M edges / 12 500 fibres of bf16 factors whose last bit depends on the
        # summation order upstream (the kurtosis columns of the fibre MLP amplify it most): Frobenius-norm relative
        d = (gp3[k].float() - gp1[k].float()).norm().item()
        # (+ one bf16 ulp per entry; biases behind a BatchNorm are cancellations of large terms: looser)
        rel = 2e-1 if k.endswith("bias") else 5e-2
        assert d < rel * gp1[k].float().norm().item() + 2 ** -7 * scale * gp1[k].numel() ** 0.5, (k, d)


def test_c5_edge_order_invariance():
    dev = _dev(60)
    S, T, F = 100000, 512, 128
    blk = _block(F, dev)
    g = torch.Generator().manual_seed(7)
    e = torch.nonzero(torch.rand(S * T, generator=g) < 0.1).flatten()
    E = e.numel()
    orders = [e[torch.randperm(E, generator=g)], e]               # shuffled, and sorted (fibre-major)
    gen = torch.Generator(device=dev).manual_seed(5)
    bf = torch.bfloat16
    x_s, x_t, u = (torch.randn(n, F, generator=gen, device=dev).to(bf) for n in (S, T, 1))
    xe_by_edge = torch.randn(S * T, F, generator=torch.Generator().manual_seed(6)).to(bf)[e]   # keyed by (fibre, class)
    up_by_edge = torch.randn(E, F, generator=torch.Generator().manual_seed(8)).to(bf)
    ups_n = [torch.randn(t.shape, generator=gen, device=dev).to(bf) for t in (x_s, x_t, u)]
    state = {k: v.clone() for k, v in blk.state_dict().items()}
    results = []
    for order in orders:
        pos = torch.searchsorted(e, order)                        # index of every edge in the sorted list
        ei = torch.stack([order // T, order % T]).contiguous().to(dev)
        ins = [x_s, x_t, xe_by_edge[pos].to(dev), u]
        ups = [ups_n[0], ups_n[1], up_by_edge[pos].to(dev), ups_n[2]]
        blk.load_state_dict(state)
        out, gin, gp = _run(blk, ei, ins, ups)
        inv = torch.empty_like(pos)
        inv[pos] = torch.arange(E)
        inv = inv.to(dev)
        results.append((out[0], out[1], out[2][inv], out[3], gin[0], gin[1], gin[2][inv], gp))
        del out, gin
    a, b = results
    # gradients pass through the cubic of the moment backward: a last-bit change of a bf16 message moves them by a few ulps
    for i, what in enumerate(("x_s", "x_t", "x_e", "u", "grad x_s", "grad x_t", "grad x_e")):
        _close(a[i], b[i], 2e-2 if i < 4 else 6e-2, what)
    for k in a[7]:
        scale = max(b[7][k].float().abs().max().item(),
                    b[7][k[:-4] + "weight"].float().abs().max().item() if k.endswith("bias") else 0.0)
        d = (a[7][k].float() - b[7][k].float()).norm().item()
        rel = 2e-1 if k.endswith("bias") else 5e-2
        assert d < rel * b[7][k].float().norm().item() + 2 ** -7 * scale * b[7][k].numel() ** 0.5, (k, d)

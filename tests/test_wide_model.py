"""tests/wide_model.py (the Block with the bf16 roundings of the wide path, test infrastructure) is the oracle when
no tensor is rounded: outputs and every gradient equal oracle/block_oracle.py in fp64."""
import pytest
import torch

from oracle import block_oracle as bo
from tests import wide_model as wm


def _case(kind, F, S, T, seed):
    g = torch.Generator().manual_seed(seed)
    ei = bo.complete_bipartite(S, T)
    if kind == "csr":
        keep = torch.rand(S * T, generator=g) < 0.6
        keep[:T] = False          # an empty fibre
        keep[3::T] = False        # an empty class
        ei = ei[:, keep]
        ei = ei[:, torch.randperm(ei.shape[1], generator=g)]
    E = ei.shape[1]
    ins = [torch.randn(n, F, generator=g, dtype=torch.float64) for n in (S, T, E, 1)]
    ups = [torch.randn(n, F, generator=g, dtype=torch.float64) for n in (S, T, E, 1)]
    return ei, ins, ups


@pytest.mark.parametrize("kind,F,S,T", [("dense", 8, 30, 6), ("csr", 16, 25, 8)])
def test_unrounded_model_is_the_oracle(kind, F, S, T):
    ei, ins, ups = _case(kind, F, S, T, seed=5)
    sd0 = bo.cast_state(bo.random_block_state(F, seed=2), torch.float64)
    res = []
    for which in ("oracle", "model"):
        sd = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v) for k, v in sd0.items()}
        xs = [t.clone().requires_grad_(True) for t in ins]
        if which == "oracle":
            outs = bo.block(sd, "", ei, *xs, training=True, buffers={})
        else:
            outs = wm.block_rounded(sd, ei, *xs, rounding=())
        torch.autograd.backward(list(outs), ups)
        res.append((outs, xs, sd))
    (o_a, x_a, s_a), (o_b, x_b, s_b) = res
    for a, b in zip(o_a, o_b):
        assert torch.allclose(a, b, rtol=1e-10, atol=1e-12)
    for a, b in zip(x_a, x_b):
        assert torch.allclose(a.grad, b.grad, rtol=1e-9, atol=1e-11)
    for k, v in s_a.items():
        if v.is_floating_point() and v.requires_grad:
            assert torch.allclose(v.grad, s_b[k].grad, rtol=1e-9, atol=1e-10), k


def test_rounding_moves_gradients_more_than_outputs():
    """The property the bf16 tolerance policy rests on: rounding ONE hidden activation tensor
    to bf16 in the forward leaves the outputs within 1e-2 of fp64 but moves gradients by several 1e-2."""
    F, S, T = 32, 96, 64
    ei = bo.complete_bipartite(S, T)
    g = torch.Generator().manual_seed(7)
    sd0 = {k: (v.bfloat16().double() if v.is_floating_point() else v) for k, v in bo.random_block_state(F, seed=3).items()}
    ins = [torch.randn(n, F, generator=g).bfloat16().double() for n in (S, T, S * T, 1)]
    ups = [torch.randn(n, F, generator=g).bfloat16().double() for n in (S, T, S * T, 1)]
    out = {}
    for name, R in (("exact", ()), ("a1", ("a1",))):
        sd = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v) for k, v in sd0.items()}
        xs = [t.clone().requires_grad_(True) for t in ins]
        outs = wm.block_rounded(sd, ei, *xs, rounding=R, rms_eps=float(torch.finfo(torch.bfloat16).eps))
        torch.autograd.backward(list(outs), ups)
        out[name] = (outs, xs)
    err = lambda a, b: ((a - b).abs().max() / b.abs().max()).item()
    fwd = max(err(a, b) for a, b in zip(out["a1"][0], out["exact"][0]))
    gxe = err(out["a1"][1][2].grad, out["exact"][1][2].grad)
    assert fwd < 1e-2 and gxe > 2e-2, (fwd, gxe)

"""Checks the hand-derived forward/backward algebra of tests/kernel_model.py (what the CUDA kernels
implement) against autograd over the oracle restatement, in fp64, per module."""
import pytest
import torch

from oracle import block_oracle as bo
from tests import kernel_model as km
from tests.util import nerr, upstream


def sub(sd, prefix):
    return {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}


def setup(seed, F=10, S=19, T=12, kind="dense", u_zero=False):
    g = torch.Generator().manual_seed(seed)
    sd = bo.random_block_state(F, seed=seed, dtype=torch.float64)
    for k in sd:
        if k.endswith("running_mean"):
            sd[k] = torch.randn(sd[k].shape, generator=g, dtype=torch.float64) * 0.3
        if k.endswith("running_var"):
            sd[k] = 0.5 + torch.rand(sd[k].shape, generator=g, dtype=torch.float64)
    ei = bo.complete_bipartite(S, T)
    if kind == "sparse":
        ei = ei[:, torch.rand(S * T, generator=g) < 0.4]
        ei = ei[:, torch.randperm(ei.shape[1], generator=g)]
    E = ei.shape[1]
    x_s = torch.randn(S, F, generator=g, dtype=torch.float64)
    x_t = torch.randn(T, F, generator=g, dtype=torch.float64)
    x_e = torch.randn(E, F, generator=g, dtype=torch.float64)
    u = torch.zeros(1, F, dtype=torch.float64) if u_zero else torch.randn(1, F, generator=g, dtype=torch.float64)
    return sd, ei, x_s, x_t, x_e, u


def autograd_ref(fn, sd, prefix, ins, training, normed):
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items()
              if k.startswith(prefix) and v.is_floating_point() and "running" not in k}
    full = dict(sd)
    full.update(params)
    ins = [t.clone().requires_grad_(True) for t in ins]
    buffers = {}
    out = fn(full, ins, buffers)
    (out * upstream(out)).sum().backward()
    return out.detach(), [t.grad if t.grad is not None else torch.zeros_like(t) for t in ins], \
        {k[len(prefix):]: p.grad for k, p in params.items() if p.grad is not None}, buffers


def check_grads(mine, ref):
    assert set(mine) == set(ref), (set(mine) ^ set(ref))
    scale = max(v.abs().max().item() for v in ref.values())
    for k in ref:
        assert (mine[k] - ref[k]).abs().max().item() <= 1e-9 * max(scale, 1.0), k


@pytest.mark.parametrize("training,normed,kind", [(True, True, "dense"), (False, True, "dense"),
                                                  (True, False, "dense"), (True, True, "sparse"),
                                                  (False, True, "sparse")])
def test_edge(training, normed, kind):
    sd, ei, x_s, x_t, x_e, u = setup(1, kind=kind)
    P = "edge_model."
    out_r, gin_r, gp_r, buf = autograd_ref(
        lambda full, ins, b: bo.edge_model(full, P, ins[0], ins[1], ei, ins[2], ins[3], training, normed, b),
        sd, P, [x_s, x_t, x_e, u], training, normed)
    p = sub(sd, P)
    rm, rv = p.get("norm.running_mean"), p.get("norm.running_var")
    out, saved = km.edge_fwd(p, x_s, x_t, ei[0], ei[1], x_e, u, training, normed, rm, rv)
    assert nerr(out, out_r) < 1e-11
    g = upstream(out)
    dx_s, dx_t, dx_e, du, grads = km.edge_bwd(p, x_s, x_t, ei[0], ei[1], x_e, u, out, saved, g, training, normed, rm, rv)
    for a, b in zip((dx_s, dx_t, dx_e, du), gin_r):
        assert nerr(a, b) < 1e-9
    check_grads(grads, gp_r)
    if training and normed:
        assert nerr(saved["running_mean"], buf[P + "norm.running_mean"]) < 1e-11
        assert nerr(saved["running_var"], buf[P + "norm.running_var"]) < 1e-11
        assert int(buf[P + "norm.num_batches_tracked"]) == 2


@pytest.mark.parametrize("training,normed,kind", [(True, True, "dense"), (False, True, "dense"),
                                                  (True, False, "dense"), (True, True, "sparse")])
def test_source(training, normed, kind):
    sd, ei, x_s, x_t, x_e, u = setup(2, kind=kind)
    P = "s_model."
    out_r, gin_r, gp_r, buf = autograd_ref(
        lambda full, ins, b: bo.s_model(full, P, ins[0], ins[1], ei, ins[2], ins[3], training, normed, b),
        sd, P, [x_s, x_t, x_e, u], training, normed)
    p = sub(sd, P)
    rm, rv = p.get("norm.running_mean"), p.get("norm.running_var")
    out, saved = km.source_fwd(p, x_s, x_t, ei[0], ei[1], x_e, u, training, normed, rm, rv)
    assert nerr(out, out_r) < 1e-11
    g = upstream(out)
    grads_in = km.source_bwd(p, x_s, x_t, ei[0], ei[1], x_e, u, saved, g, training, normed, rm, rv)
    for a, b in zip(grads_in[:4], gin_r):
        assert nerr(a, b) < 1e-8
    check_grads(grads_in[4], gp_r)
    if training and normed:
        assert nerr(saved["running_mean"], buf[P + "norm.running_mean"]) < 1e-11
        assert nerr(saved["running_var"], buf[P + "norm.running_var"]) < 1e-11


@pytest.mark.parametrize("training,normed,kind", [(True, True, "dense"), (False, True, "dense"),
                                                  (True, False, "dense"), (True, True, "sparse")])
def test_target(training, normed, kind):
    sd, ei, x_s, x_t, x_e, u = setup(3, kind=kind)
    P = "t_model."
    out_r, gin_r, gp_r, buf = autograd_ref(
        lambda full, ins, b: bo.t_model(full, P, ins[0], ins[1], ei, ins[2], ins[3], training, normed, b),
        sd, P, [x_s, x_t, x_e, u], training, normed)
    p = sub(sd, P)
    rm, rv = p.get("norm.running_mean"), p.get("norm.running_var")
    out, saved = km.target_fwd(p, x_s, x_t, ei[0], ei[1], x_e, u, training, normed, rm, rv)
    assert nerr(out, out_r) < 1e-11
    g = upstream(out)
    grads_in = km.target_bwd(p, x_s, x_t, ei[0], ei[1], x_e, u, saved, g, training, normed, rm, rv)
    for a, b in zip(grads_in[:4], gin_r):
        assert nerr(a, b) < 1e-9
    check_grads(grads_in[4], gp_r)


@pytest.mark.parametrize("normed", [True, False])
def test_global(normed):
    sd, ei, x_s, x_t, x_e, u = setup(4)
    P = "global_model."
    out_r, gin_r, gp_r, _ = autograd_ref(
        lambda full, ins, b: bo.global_model(full, P, ins[0], ins[1], ins[2], normed),
        sd, P, [x_s, x_t, u], True, normed)
    p = sub(sd, P)
    out, saved = km.global_fwd(p, x_s, x_t, u, normed)
    assert nerr(out, out_r) < 1e-11
    dxs, dxt, du, grads = km.global_bwd(p, x_s, x_t, u, saved, upstream(out), normed)
    for a, b in zip((dxs, dxt, du), gin_r):
        assert nerr(a, b) < 1e-9
    check_grads(grads, gp_r)

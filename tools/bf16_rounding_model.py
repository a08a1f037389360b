#!/usr/bin/env python
"""Where the bf16 error of a Block comes from: the reference Block in fp64 (tests/wide_model.py) with chosen tensors
rounded to bf16 in the forward and / or in the backward.  CPU only.

    python tools/bf16_rounding_model.py > profiles/r02_bf16_rounding_model.txt

Columns: norm-wise error max|a-b| / max|b| against the unrounded fp64 run, for the outputs, the input gradients and
five weight gradients."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import block_oracle as bo  # noqa: E402
from tests import wide_model as wm  # noqa: E402

GRAD_POINTS = ("h1", "z", "xe2", "hs", "m", "h3", "ht", "h3t", "asum", "agg", "xs2")
KEYS = ["edge_model.0.weight", "s_model.node_mlp_1.0.weight", "s_model.node_mlp_2.0.weight", "t_model.node_mlp_2.0.weight",
        "t_model.node_mlp_1.2.weight"]


def err(a, b):
    return ((a - b).abs().max() / b.abs().max()).item()


def run(sd0, ei, ins, ups, R, RG):
    sd = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v) for k, v in sd0.items()}
    xs = [t.clone().requires_grad_(True) for t in ins]
    o = wm.block_rounded(sd, ei, *xs, rounding=R, grad_rounding=RG, rms_eps=float(torch.finfo(torch.bfloat16).eps))
    torch.autograd.backward(list(o), ups)
    return o, xs, sd


def main():
    print(__doc__)
    for (F, S, T, seed) in [(32, 96, 64, 3), (128, 96, 64, 1)]:
        sd0 = {k: (v.bfloat16().double() if v.is_floating_point() else v) for k, v in bo.random_block_state(F, seed=seed).items()}
        ei = bo.complete_bipartite(S, T)
        g = torch.Generator().manual_seed(7)
        r = lambda *s: torch.randn(*s, generator=g).bfloat16().double()
        ins = [r(S, F), r(T, F), r(S * T, F), r(1, F)]
        ups = [r(S, F), r(T, F), r(S * T, F), r(1, F)]
        ref = run(sd0, ei, ins, ups, (), ())
        print("\nF=%d S=%d T=%d (complete graph, train mode)" % (F, S, T))
        print("%-52s %8s %8s %8s | %8s %8s %8s | %s" % ("rounded tensors", "x_s", "x_t", "x_e", "g_x_s", "g_x_t", "g_x_e",
                                                          " ".join(k.replace("node_mlp_", "mlp")[:18].rjust(18) for k in KEYS)))
        rows = [("forward: all bf16 storage points; backward: all", wm.ALL_ROUNDING, GRAD_POINTS),
                ("forward: all bf16 storage points; backward: exact", wm.ALL_ROUNDING, ()),
                ("forward: exact; backward: all", (), GRAD_POINTS),
                ("forward: default wide.PREC set; backward: all", wm.WIDE_ROUNDING, GRAD_POINTS),
                ("forward: default wide.PREC set; backward: exact", wm.WIDE_ROUNDING, ())]
        rows += [("forward: only %s" % n, (n,), ()) for n in ("a1", "xe2", "a_s", "m", "hcat", "a3", "xs2", "a_t", "asum", "agg", "a3t")]
        for name, R, RG in rows:
            o, xi, st = run(sd0, ei, ins, ups, R, RG)
            print("%-52s %8.1e %8.1e %8.1e | %8.1e %8.1e %8.1e | %s" % (
                name, err(o[0], ref[0][0]), err(o[1], ref[0][1]), err(o[2], ref[0][2]),
                err(xi[0].grad, ref[1][0].grad), err(xi[1].grad, ref[1][1].grad), err(xi[2].grad, ref[1][2].grad),
                " ".join(("%.1e" % err(st[k].grad, ref[2][k].grad)).rjust(18) for k in KEYS)))


if __name__ == "__main__":
    main()

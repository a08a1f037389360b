#!/usr/bin/env python
"""Per-tensor error table of the bf16 wide path against the oracle (fp64; fp32 for the big case), next to the
reference's own path executed in bf16 -- for several settings of wide.PREC (which intermediates stay fp32).

    python tools/wide_error_table.py [--cases smoke,c4mid,...] [--prec "none;z32;z32,m32;z32,m32,at32"]

Norm-wise error max|a-b| / max|b| per tensor (north-star bf16 bound 1e-2)."""
import argparse
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import block_oracle as bo  # noqa: E402

CASES = {
    # name: (kind, F, S, T, oracle dtype, with bf16 reference)
    "smoke": ("dense", 32, 96, 64, torch.float64, True),
    "t1": ("dense", 32, 200, 64, torch.float64, True),
    "t3": ("csr", 32, 150, 64, torch.float64, True),
    "t4": ("dense", 128, 96, 64, torch.float64, True),
    "t5": ("dense", 32, 40, 128, torch.float64, True),
    "t6": ("dense", 128, 12, 256, torch.float64, True),
    "c4mid": ("dense", 128, 1024, 512, torch.float32, False),
}
RMS_EPS_BF16 = float(torch.finfo(torch.bfloat16).eps)


def graph(kind, S, T, seed):
    g = torch.Generator().manual_seed(seed)
    ei = bo.complete_bipartite(S, T)
    if kind == "dense":
        return ei
    keep = torch.rand(S * T, generator=g) < 0.6
    keep[:T] = False
    keep[3::T] = False
    ei = ei[:, keep]
    return ei[:, torch.randperm(ei.shape[1], generator=g)]


def err(a, b, scale=None):
    a, b = a.detach().double().cpu(), b.detach().double()
    den = b.abs().max().item() if scale is None else scale
    return (a - b).abs().max().item() / max(den, 1e-30)


def run_oracle(state, ei, ins, ups, dtype):
    sd = bo.cast_state(state, dtype)
    sd = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v) for k, v in sd.items()}
    xs = [t.detach().clone().to(dtype).requires_grad_(True) for t in ins]
    outs = bo.block(sd, "", ei, *xs, training=True, buffers={}, rms_eps=RMS_EPS_BF16)
    torch.autograd.backward(list(outs), [u.to(dtype) for u in ups])
    return outs, xs, sd


def collect(outs, xs, named_grads, ref_outs, ref_xs, ref_sd):
    rep = {}
    for n, a, b in zip(("x_s", "x_t", "x_e", "u"), outs, ref_outs):
        rep[n] = err(a, b)
    for n, a, b in zip(("g_x_s", "g_x_t", "g_x_e", "g_u"), xs, ref_xs):
        rep[n] = err(a.grad, b.grad)
    for k, g in named_grads.items():
        ref = ref_sd[k].grad
        if ref is None or g is None:
            continue
        scale = None
        if k.endswith("bias") and ".norm." not in k:
            scale = max(ref.abs().max().item(), ref_sd[k[:-4] + "weight"].grad.abs().max().item())
        rep["grad " + k] = err(g, ref, scale)
    return rep


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", default="smoke,t1,t4,t6")
    ap.add_argument("--prec", default="none;z32;z32,m32;z32,m32,at32")
    ap.add_argument("--top", type=int, default=12)
    args = ap.parse_args()
    from pfs_neural_net_b200 import gnn, wide as pw
    dev = torch.device("cuda:0")
    for cname in args.cases.split(","):
        kind, F, S, T, odt, with_ref = CASES[cname]
        ei = graph(kind, S, T, seed=F + S)
        E = ei.shape[1]
        state = {k: (v.bfloat16() if v.is_floating_point() else v) for k, v in bo.random_block_state(F, seed=1 if cname != "smoke" else 3).items()}
        g = torch.Generator().manual_seed(7)
        r = lambda *s: torch.randn(*s, generator=g).bfloat16()
        ins = [r(S, F), r(T, F), r(E, F), r(1, F)]
        ups = [torch.randn(s, generator=g).bfloat16() for s in ((S, F), (T, F), (E, F), (1, F))]
        t0 = time.time()
        ref_outs, ref_xs, ref_sd = run_oracle(state, ei, ins, ups, odt)
        t_or = time.time() - t0
        cols = {}
        if with_ref:
            q_outs, q_xs, q_sd = run_oracle(state, ei, ins, ups, torch.bfloat16)
            cols["ref-in-bf16"] = collect(q_outs, q_xs, {k: v.grad for k, v in q_sd.items() if v.is_floating_point() and v.requires_grad},
                                          ref_outs, ref_xs, ref_sd)
        for prec in args.prec.split(";"):
            pw.PREC.clear()
            pw.PREC.update(x for x in prec.split(",") if x and x != "none")
            blk = gnn.Block(F).to(torch.bfloat16)
            blk.load_state_dict(state, strict=True)
            blk = blk.to(dev).train()
            x = [t.detach().clone().to(dev).requires_grad_(True) for t in ins]
            _, o_s, o_t, o_e, o_u = blk((ei.to(dev), *x))
            torch.autograd.backward([o_s, o_t, o_e, o_u], [u.to(dev) for u in ups])
            torch.cuda.synchronize()
            cols[prec] = collect((o_s, o_t, o_e, o_u), x, {k: p.grad for k, p in blk.named_parameters()}, ref_outs, ref_xs, ref_sd)
        names = list(cols)
        print("\n== %s: %s F=%d S=%d T=%d E=%d (oracle %s, %.1f s)" % (cname, kind, F, S, T, E, str(odt).replace("torch.", ""), t_or))
        print("%-38s " % "tensor" + " ".join("%14s" % n[:14] for n in names))
        keys = list(next(iter(cols.values())))
        fwd = keys[:8]
        rest = sorted(keys[8:], key=lambda k: -max(c.get(k, 0) for n, c in cols.items() if n != "ref-in-bf16"))[:args.top]
        for k in fwd + rest:
            print("%-38s " % k[:38] + " ".join("%14.2e" % cols[n].get(k, float("nan")) for n in names))
        for n in names:
            over = [k for k in keys if not cols[n].get(k, 0) < 1e-2]
            print("  %-14s tensors over 1e-2: %d of %d %s" % (n[:14], len(over), len(keys), over[:6]))


if __name__ == "__main__":
    main()

"""Wide-feature path (Fdim >= 32, bf16): the update modules of reference src/gnn.py:73-223 composed
from the tensor-core / HBM-bound primitives of libpfs_b200.so (wide_ops.py), one graph per call.

Same algebra as the narrow fp32 kernels (tests/kernel_model.py is the executable
statement of it): first-layer split into gathered node tables, closed-form double BatchNorm, target
aggregation commuted with its Linear, per-fibre cubic for the moment backward.  What differs is the
execution: at these widths every Linear is a dense contraction, so it runs as a tcgen05 GEMM
(`gemm_nt` forward / input gradients, `gemm_tn` weight gradients), activations are stored in bf16
(a1, messages, hidden activations are SAVED for the backward instead of recomputed), statistics and
all reductions are fp32 and deterministic.  PyTorch is used for memory and for O(Fdim) coefficient
vectors (BatchNorm / RMSNorm scalars per feature); everything that scales with edges or nodes is a
library kernel.

Fibre-range sharding (BASELINE config C4): when `shard` (see shard.py) is active every reduction
over "all edges / all fibres" below goes through `shard.allreduce_*` -- the BatchNorm statistics,
the class-side aggregates and the class-table gradients -- and nothing else changes.
"""
import torch

from . import _abi
from . import wide_ops as wo
from . import shard as _shard

BF16, F32 = torch.bfloat16, torch.float32
# Where fp32 is kept between the tensor-core GEMMs (north-star bf16 tolerance 1e-2):
#   z32  -- the EdgeModel pre-BatchNorm output z [E,F]: statistics AND the normalised x_e' are computed from fp32 z
#   m32  -- the SModel messages m [E,2F]: the moment statistics (E[m^2] - mean^2, third / fourth central moments over
#           std^3 / std^4 amplify a 2^-9 rounding of m) and their backward read fp32 m
#   at32 -- the TModel hidden activations feeding the class sums (fp32 copy next to the bf16 one kept for the mask)
#   node32 -- node-level GEMM operands (O(S + T) rows: the SModel statistics and hidden layer, the TModel class sums,
#           aggregate and hidden layer) enter their bf16 GEMMs as [hi | lo] pairs against [W | W] (wide_ops.split), i.e.
#           to ~2^-17 instead of 2^-9; at T = 512 the node-level GEMMs are ~1 % of the step's FLOPs
# PFS_WIDE_PREC=<comma list> overrides the default set (A/B runs of tools/wide_error_table.py); "none" = all bf16.
import os as _os
_PREC_DEFAULT = "z32,m32,at32,node32"
PREC = set(filter(None, _os.environ.get("PFS_WIDE_PREC", _PREC_DEFAULT).replace("none", "").split(",")))
BN_EPS = 1e-5
BN_MOMENTUM = 0.1
SLOPE = 0.1


def supported(F, dtype):
    """The wide path serves bf16 tensors with Fdim a multiple of 8 (TMA alignment), at least 16."""
    return dtype == BF16 and F >= 16 and F % 8 == 0


class WideTopology:
    """What the wide kernels need from `edge_index`: segments of rows per fibre / per class and, for a
    general edge list, the int32 src / tgt of every edge (dense canonical order needs no arrays)."""

    def __init__(self, topo):
        self.S, self.T, self.E, self.device = topo.S, topo.T, topo.E, topo.device
        if topo.canonical:
            self.src = self.tgt = None
            self.fibres = wo.Segments(0, self.S, self.S, self.T)
            self.classes = wo.Segments(1, self.T, self.S, self.T)
            self.class_count = torch.full((self.T,), float(self.S), dtype=F32, device=self.device)
        else:
            a = topo.csr()
            ei = topo.edge_index
            self.src = ei[0].to(torch.int32).contiguous()
            self.tgt = ei[1].to(torch.int32).contiguous()
            self.csc_eid = a["csr_eid"][a["csc_q"].long()].contiguous()      # one-time index composition
            self.fibres = wo.Segments(2, self.S, ptr=a["csr_rowptr"], lst=a["csr_eid"])
            self.classes = wo.Segments(2, self.T, ptr=a["csc_colptr"], lst=self.csc_eid)
            cp = a["csc_colptr"]
            self.class_count = (cp[1:] - cp[:-1]).to(F32)
        self._class_count_total = {}
        self.div = self.T      # dense: src = e // T, tgt = e % T
        # canonical order with whole 128-edge tiles inside one fibre: the class rows of a tile are contiguous, so
        # x_t[tgt] is a second GEMM operand and P_s[src] a per-tile bias row -- no gathered tables at all
        self.tiled_dense = bool(topo.canonical and self.T % 128 == 0)


    def class_count_total(self):
        """Edges per class over all fibre shards: a property of the partition, exchanged once per process group."""
        if not _shard.active():
            return self.class_count
        key = _shard.group_key()
        hit = self._class_count_total.get(key)
        if hit is None:
            hit = _shard.allreduce_sum(self.class_count.clone())
            self._class_count_total[key] = hit
        return hit


# ------------------------------------------------------------------------------------------------ helpers
def _f32(t):
    return wo.cast(t.contiguous(), F32)


def _bf(t):
    return wo.cast(t.contiguous(), BF16)


def _like(g32, ref):
    """fp32 gradient -> dtype of the tensor it belongs to."""
    return wo.cast(g32.contiguous(), ref.dtype) if ref.dtype != F32 else g32


def _colsum(x):
    return wo.colstats(1, x)[0]


def _bn_train_stats(y, rows_local):
    """(n, mean, var) of the columns of y over ALL rows (all shards)."""
    st = wo.colstats(0, y)
    n, mean, m2 = _shard.allreduce_moments(float(rows_local), st[0], st[1])
    return n, mean, m2 / n


def _update_running(rm, rv, nbt, mean, var_unbiased, steps=1):
    if rm is None or rv is None:
        return
    with torch.no_grad():
        rm.mul_(1 - BN_MOMENTUM).add_((BN_MOMENTUM * mean).to(rm.dtype))
        rv.mul_(1 - BN_MOMENTUM).add_((BN_MOMENTUM * var_unbiased).to(rv.dtype))
        if nbt is not None:
            nbt.add_(steps)


def _single_bn_fwd(y, rows, training, gamma, beta, rm, rv, nbt):
    """BatchNorm1d over the rows of y (fp32 [rows, F]) -> bf16 output and what the backward needs."""
    g, b = gamma.float(), beta.float()
    if training:
        n, mu, var = _bn_train_stats(y, rows)
        if n <= 1:
            raise ValueError("Expected more than 1 value per channel when training")
        r = torch.rsqrt(var + BN_EPS)
        _update_running(rm, rv, nbt, mu, var * (n / (n - 1)))
    else:
        mu, r = rm.float(), torch.rsqrt(rv.float() + BN_EPS)
        n = float(rows)
    a = g * r
    out = wo.rowmap(0, y, a.contiguous(), (b - a * mu).contiguous())
    return out, mu.contiguous(), r.contiguous(), n


def _single_bn_bwd(gout, y, mu, r, n, training, gamma):
    """dy (bf16) and the affine gradients of a BatchNorm1d whose input y (fp32) was saved."""
    g = gamma.float()
    st = _shard.allreduce_sum(wo.colstats(1, gout, v=y, p0=mu, p1=r))          # [2,F]: one exchange
    sg, sgx = st[0], st[1]
    a = (g * r).contiguous()
    if training:
        dy = wo.rowmap(1, gout, a, (sg / n).contiguous(), v=y, p0=mu, p1=r, c2=(sgx / n).contiguous())
    else:
        dy = wo.rowmap(0, gout, a, torch.zeros_like(a))
    return dy, sgx, sg


# ------------------------------------------------------------------------------------------------ EdgeModel
class WideEdgeFunction(torch.autograd.Function):
    """EdgeModel (reference src/gnn.py:73-101) on the wide path."""

    @staticmethod
    def forward(ctx, topo, training, normed, x_s, x_t, x_e, u, w1, b1, w2, b2, gamma, beta, rm, rv, nbt):
        wt = topo.wide()
        x_s, x_t, x_e, u, w1, w2 = (t.contiguous() for t in (x_s, x_t, x_e, u, w1, w2))
        F = x_e.shape[1]
        E = x_e.shape[0]
        uvec = wo.gemm_nt(u, w1[:, 3 * F:], bias=_f32(b1), want="f32")                          # W1_u.u + b1
        Ps = wo.gemm_nt(x_s, w1[:, :F], want="f32")                                             # [S,4F] fp32
        if wt.tiled_dense:
            # [x_t[tgt] | x_e] . [W1_t | W1_e]^T + P_s[src] (tile-constant) + (W1_u.u + b1)
            a1 = wo.gemm_nt(x_e, w1[:, F:3 * F], A2=x_t, a2_mod=wt.T, bias=uvec[0].contiguous(), bias_rows=Ps,
                            bias_rows_div=wt.T, act=True)
        else:
            Pt = wo.gemm_nt(x_t, w1[:, F:2 * F], bias=uvec[0].contiguous(), want="f32")         # [T,4F] fp32
            a1 = wo.gemm_nt(x_e, w1[:, 2 * F:3 * F], tab0=Ps, idx0=wt.src, div0=wt.div, tab1=Pt, idx1=wt.tgt,
                            mod1=wt.div, act=True)                                              # [E,4F] bf16
        z32 = "z32" in PREC and normed
        z = wo.gemm_nt(a1, w2, bias=_f32(b2), want="f32" if z32 else "bf16")                    # [E,F]
        saved_small = {}
        if normed:
            g, b = gamma.float(), beta.float()
            if training:
                n, mu, var = _bn_train_stats(z, E)
                if n <= 1:
                    raise ValueError("Expected more than 1 value per channel when training")
                r1 = torch.rsqrt(var + BN_EPS)
                var2 = g * g * var * r1 * r1
                r2 = torch.rsqrt(var2 + BN_EPS)
                A = g * g * r1 * r2
                shift = b - A * mu
                unb = n / (n - 1)
                # two momentum updates: the norm is applied twice (src/gnn.py:82,101)
                _update_running(rm, rv, None, mu, var * unb)
                _update_running(rm, rv, nbt, b, var2 * unb, steps=2)
                saved_small = dict(var=var, r1=r1, r2=r2, n=n)
            else:
                rmf, rvf = rm.float(), rv.float()
                a = g * torch.rsqrt(rvf + BN_EPS)
                A = a * a
                shift = a * (b - rmf - a * rmf) + b
                saved_small = dict(a=a)
            saved_small.update(A=A.contiguous(), shift=shift.contiguous())
            # bf16 z is normalised in place (it is not needed again); fp32 z goes to a fresh bf16 x_e'
            out = wo.rowmap(0, z, saved_small["A"], saved_small["shift"], out=None if z32 else z)
        else:
            out = z
        ctx.topo, ctx.training, ctx.normed, ctx.small = topo, training, normed, saved_small
        ctx.buffers = (rm, rv)
        ctx.save_for_backward(x_s, x_t, x_e, u, w1, b1, w2, b2, gamma, beta, a1, out)
        return out

    @staticmethod
    def backward(ctx, g):
        x_s, x_t, x_e, u, w1, b1, w2, b2, gamma, beta, a1, xe2 = ctx.saved_tensors
        wt = ctx.topo.wide()
        sm = ctx.small
        F, E = x_e.shape[1], x_e.shape[0]
        H = 4 * F
        g = g.contiguous()
        g_gamma = g_beta = None
        if not ctx.normed:
            dz = g
        else:
            gm, bt = gamma.float(), beta.float()
            A = sm["A"]
            if ctx.training:
                r1, r2, var, n = sm["r1"], sm["r2"], sm["var"], sm["n"]
                g2 = gm * gm * r2
                inv = torch.where(g2 != 0, 1.0 / g2, torch.zeros_like(g2)).contiguous()        # xhat1 = (x_e' - beta) inv
                st = _shard.allreduce_sum(wo.colstats(1, g, v=xe2, p0=bt.contiguous(), p1=inv))
                sg, sgx = st[0], st[1]
                gbar, mgx = sg / n, sgx / n
                s = gm * r2
                q = var * r1 * r1
                kappa = s * s + 1 - s * s * q
                dz = wo.rowmap(1, g, A, gbar.contiguous(), v=xe2, p0=bt.contiguous(), p1=inv, c2=(mgx * kappa).contiguous())
                g_gamma = n * mgx * s * (2 - s * s * q)
                g_beta = sg
            else:
                rm, rv = (t.float() for t in ctx.buffers)
                a, shift = sm["a"], sm["shift"]
                c = torch.rsqrt(rv + BN_EPS)
                invA = torch.where(A != 0, 1.0 / A, torch.zeros_like(A)).contiguous()
                st = _shard.allreduce_sum(wo.colstats(1, g, v=xe2, p0=shift, p1=invA))          # sum g, sum g z
                sg, sgz = st[0], st[1]
                dz = wo.rowmap(0, g, A, torch.zeros_like(A))
                g_gamma = c * (2 * a * (sgz - rm * sg) + (bt - rm) * sg)
                g_beta = (a + 1) * sg
        w2t = wo.transpose(w2)                                   # [4F, F]
        w1t = wo.transpose(w1)                                   # [4F(in), 4F(out)]: row blocks = input slices
        dh1 = wo.gemm_nt(dz, w2t, mask=a1)                       # (dz W2) . lrelu'(h1)   [E,4F] bf16
        g_w2 = wo.gemm_tn(dz, a1)
        g_b2 = _colsum(dz)
        dPs = wo.segsum(wt.fibres, dh1, want="bf16")             # [S,4F]
        dPt32 = _shard.allreduce_sum(wo.segsum(wt.classes, dh1, want="f32"))   # [T,4F] (summed over fibre shards)
        dPt = _bf(dPt32)
        tot = _colsum(dPt32)                                     # [4F]
        tot16 = _bf(tot[None])
        g_w1 = torch.empty(H, H, dtype=F32, device=g.device)
        wo.gemm_tn(dPs, x_s, out=g_w1[:, :F])
        wo.gemm_tn(dPt, x_t, out=g_w1[:, F:2 * F])
        wo.gemm_tn(dh1, x_e, out=g_w1[:, 2 * F:3 * F])
        wo.gemm_tn(tot16, u, out=g_w1[:, 3 * F:])
        g_x_s = wo.gemm_nt(dPs, w1t[:F])
        g_x_t = wo.gemm_nt(dPt, w1t[F:2 * F])
        g_x_e = wo.gemm_nt(dh1, w1t[2 * F:3 * F])
        g_u = wo.gemm_nt(tot16, w1t[3 * F:])
        # fibre-local parameter gradients are partial under sharding; class-side ones are already global
        # (one packed exchange for all of them)
        if _shard.active():
            w1s, w1e, g_w2, g_b2 = _shard.allreduce_packed([g_w1[:, :F], g_w1[:, 2 * F:3 * F], g_w2, g_b2])
            g_w1[:, :F] = w1s
            g_w1[:, 2 * F:3 * F] = w1e
        return (None, None, None, g_x_s, g_x_t, g_x_e, g_u, _like(g_w1, w1), _like(tot, b1), _like(g_w2, w2),
                _like(g_b2, b2), None if g_gamma is None else _like(g_gamma, gamma),
                None if g_beta is None else _like(g_beta, beta), None, None, None)


# ------------------------------------------------------------------------------------------------ SModel
class WideSourceFunction(torch.autograd.Function):
    """SModel (reference src/gnn.py:104-154) on the wide path."""

    @staticmethod
    def forward(ctx, topo, training, normed, x_s, x_t, x_e, u, w1, b1, w2, b2, w3, b3, w4, b4, gamma, beta, rm, rv, nbt):
        wt = topo.wide()
        x_s, x_t, x_e, u, w1, w2, w3, w4 = (t.contiguous() for t in (x_s, x_t, x_e, u, w1, w2, w3, w4))
        S, F = x_s.shape
        if wt.tiled_dense:
            a_s = wo.gemm_nt(x_e, w1, A2=x_t, a2_mod=wt.T, bias=_f32(b1), act=True)             # [x_t[tgt] | x_e] . W1^T + b1
        else:
            Qt = wo.gemm_nt(x_t, w1[:, :F], bias=_f32(b1), want="f32")                          # [T,2F] fp32
            a_s = wo.gemm_nt(x_e, w1[:, F:], tab1=Qt, idx1=wt.tgt, mod1=wt.div, act=True)       # [E,2F]
        m = wo.gemm_nt(a_s, w2, bias=_f32(b2), want="f32" if "m32" in PREC else "bf16")         # messages [E,2F]
        moments = wo.moments_fwd(wt.fibres, m)                                                  # [S,5,2F] fp32
        b3eff = wo.gemm_nt(u, w3[:, 9 * F:], bias=_f32(b3), want="f32")                         # [1,10F]
        if "node32" in PREC:
            hc = wo.source_hcat(x_s, moments, with_lo=True)                                     # [S,17F] = [hcat | lo(stats)]
            hcat = hc[:, :9 * F]
            a3c = wo.split(wo.gemm_nt(hc, torch.cat([w3[:, :9 * F], w3[:, F:9 * F]], 1), bias=b3eff[0].contiguous(),
                                      act=True, want="f32"))                                    # [S,20F] = [a3 | lo(a3)]
            a3 = a3c[:, :10 * F]
            y = wo.gemm_nt(a3c, torch.cat([w4, w4], 1), bias=_f32(b4), want="f32")              # [S,F] fp32
            del hc, a3c
        else:
            hcat = wo.source_hcat(x_s, moments)                                                 # [S,9F] bf16
            a3 = wo.gemm_nt(hcat, w3[:, :9 * F], bias=b3eff[0].contiguous(), act=True)          # [S,10F] bf16
            y = wo.gemm_nt(a3, w4, bias=_f32(b4), want="f32")                                   # [S,F] fp32
        bn = None
        if normed:
            out, mu, r, n = _single_bn_fwd(y, S, training, gamma, beta, rm, rv, nbt)
            bn = (mu, r, n)
        else:
            out = _bf(y)
        ctx.topo, ctx.training, ctx.normed, ctx.bn = topo, training, normed, bn
        ctx.buffers = (rm, rv)
        ctx.save_for_backward(x_s, x_t, x_e, u, w1, b1, w2, b2, w3, b3, w4, b4, gamma, beta, a_s, m, moments, hcat, a3, y)
        return out

    @staticmethod
    def backward(ctx, g):
        (x_s, x_t, x_e, u, w1, b1, w2, b2, w3, b3, w4, b4, gamma, beta, a_s, m, moments, hcat, a3, y) = ctx.saved_tensors
        wt = ctx.topo.wide()
        S, F = x_s.shape
        M2, J, K9 = 2 * F, 10 * F, 9 * F
        g = g.contiguous()
        g_gamma = g_beta = None
        if ctx.normed:
            mu, r, n = ctx.bn
            dy, g_gamma, g_beta = _single_bn_bwd(g, y, mu, r, n, ctx.training, gamma)
        else:
            dy = g
        w4t, w3t = wo.transpose(w4), wo.transpose(w3)            # [10F, F], [10F(in), 10F(out)]
        g_w4 = wo.gemm_tn(dy, a3)
        g_b4 = _colsum(dy)
        dh3 = wo.gemm_nt(dy, w4t, mask=a3)                       # [S,10F] bf16
        tot3 = _colsum(dh3)
        tot3_16 = _bf(tot3[None])
        g_w3 = torch.empty(J, J, dtype=F32, device=g.device)
        wo.gemm_tn(dh3, hcat, out=g_w3[:, :K9])
        wo.gemm_tn(tot3_16, u, out=g_w3[:, K9:])
        g_u = wo.gemm_nt(tot3_16, w3t[K9:], want="f32")
        dh = wo.gemm_nt(dh3, w3t[:K9], want="f32")               # gradient of hcat [S,9F] fp32
        g_x_s, coef = wo.source_coef(wt.fibres, dh, moments, F)  # dx_s bf16, cubic coefficients [S,4,2F]
        dm = wo.source_dm_seg(wt.fibres, m, moments, coef) if 2 * F <= 2048 else wo.source_dm(m, moments, coef, wt.src, wt.div)
        w2t, w1t = wo.transpose(w2), wo.transpose(w1)            # [2F,2F], [2F(in), 2F(out)]
        g_w2 = wo.gemm_tn(dm, a_s)
        g_b2 = _colsum(dm)
        dhs = wo.gemm_nt(dm, w2t, mask=a_s)                      # [E,2F] bf16
        dQt32 = _shard.allreduce_sum(wo.segsum(wt.classes, dhs, want="f32"))
        dQt = _bf(dQt32)
        g_b1 = _colsum(dQt32)
        g_w1 = torch.empty(M2, M2, dtype=F32, device=g.device)
        wo.gemm_tn(dQt, x_t, out=g_w1[:, :F])
        wo.gemm_tn(dhs, x_e, out=g_w1[:, F:])
        g_x_t = wo.gemm_nt(dQt, w1t[:F])
        g_x_e = wo.gemm_nt(dhs, w1t[F:])
        # fibre-local partial sums under sharding
        if _shard.active():
            w1e, g_w2, g_b2, g_w3, g_w4, g_b4, tot3, g_u = _shard.allreduce_packed(
                [g_w1[:, F:], g_w2, g_b2, g_w3, g_w4, g_b4, tot3, g_u])
            g_w1[:, F:] = w1e
        return (None, None, None, g_x_s, g_x_t, g_x_e, _like(g_u, u), _like(g_w1, w1), _like(g_b1, b1), _like(g_w2, w2),
                _like(g_b2, b2), _like(g_w3, w3), _like(tot3, b3), _like(g_w4, w4), _like(g_b4, b4),
                None if g_gamma is None else _like(g_gamma, gamma), None if g_beta is None else _like(g_beta, beta),
                None, None, None)


# ------------------------------------------------------------------------------------------------ TModel
class WideTargetFunction(torch.autograd.Function):
    """TModel (reference src/gnn.py:157-192) on the wide path; the class aggregate is the tensor that is
    all-reduced under fibre-range sharding (north star)."""

    @staticmethod
    def forward(ctx, topo, training, normed, x_s, x_t, x_e, u, w1, b1, w2, b2, w3, b3, w4, b4, gamma, beta, rm, rv, nbt):
        wt = topo.wide()
        x_s, x_t, x_e, u, w1, w2, w3, w4 = (t.contiguous() for t in (x_s, x_t, x_e, u, w1, w2, w3, w4))
        T, F = x_t.shape
        Rs = wo.gemm_nt(x_s, w1[:, :F], bias=_f32(b1), want="f32")                              # [S,2F] fp32
        want = "both" if "at32" in PREC else "bf16"
        if wt.tiled_dense:
            a_t = wo.gemm_nt(x_e, w1[:, F:], bias_rows=Rs, bias_rows_div=wt.T, act=True, want=want)   # R_s[src] is tile-constant
        else:
            a_t = wo.gemm_nt(x_e, w1[:, F:], tab0=Rs, idx0=wt.src, div0=wt.div, act=True, want=want)  # [E,2F] bf16
        a_t, a_sum_in = a_t if want == "both" else (a_t, a_t)
        asum32 = _shard.allreduce_sum(wo.segsum(wt.classes, a_sum_in, want="f32"))              # [T,2F]
        del a_sum_in
        cnt = wt.class_count_total()                             # edges per class over ALL shards (cached per topology)
        b3eff = wo.gemm_nt(u, w3[:, 3 * F:], bias=_f32(b3), want="f32")
        if "node32" in PREC:
            asum_c = wo.split(asum32)                                                           # [T,4F] = [asum | lo]
            asum = asum_c[:, :2 * F]
            agg32 = wo.gemm_nt(asum_c, torch.cat([w2, w2], 1), bias=_f32(b2), bias_rowscale=cnt, want="f32")
            hc = torch.empty(T, 5 * F, dtype=BF16, device=x_t.device)                           # [x_t | agg | lo(agg)]
            hc[:, :F].copy_(x_t)
            wo.split(agg32, out=hc[:, F:])
            hcat = hc[:, :3 * F]
            a3c = wo.split(wo.gemm_nt(hc, torch.cat([w3[:, :3 * F], w3[:, F:3 * F]], 1), bias=b3eff[0].contiguous(),
                                      act=True, want="f32"))                                    # [T,8F] = [a3 | lo(a3)]
            a3 = a3c[:, :4 * F]
            y = wo.gemm_nt(a3c, torch.cat([w4, w4], 1), bias=_f32(b4), want="f32")              # [T,F] fp32
            del asum_c, hc, a3c
        else:
            asum = _bf(asum32)
            hcat = torch.empty(T, 3 * F, dtype=BF16, device=x_t.device)
            hcat[:, :F].copy_(x_t)
            wo.gemm_nt(asum, w2, bias=_f32(b2), bias_rowscale=cnt, out_bf16=hcat[:, F:], want="none")   # agg = W2 asum + cnt b2
            a3 = wo.gemm_nt(hcat, w3[:, :3 * F], bias=b3eff[0].contiguous(), act=True)          # [T,4F] bf16
            y = wo.gemm_nt(a3, w4, bias=_f32(b4), want="f32")                                   # [T,F] fp32
        bn = None
        if normed:
            with _shard.replicated():                     # class rows are replicated on every shard
                out, mu, r, n = _single_bn_fwd(y, T, training, gamma, beta, rm, rv, nbt)
            bn = (mu, r, n)
        else:
            out = _bf(y)
        ctx.topo, ctx.training, ctx.normed, ctx.bn = topo, training, normed, bn
        ctx.buffers = (rm, rv)
        ctx.save_for_backward(x_s, x_t, x_e, u, w1, b1, w2, b2, w3, b3, w4, b4, gamma, beta, a_t, asum, cnt, hcat, a3, y)
        return out

    @staticmethod
    def backward(ctx, g):
        (x_s, x_t, x_e, u, w1, b1, w2, b2, w3, b3, w4, b4, gamma, beta, a_t, asum, cnt, hcat, a3, y) = ctx.saved_tensors
        wt = ctx.topo.wide()
        T, F = x_t.shape
        M2, H = 2 * F, 4 * F
        g = g.contiguous()
        g_gamma = g_beta = None
        with _shard.replicated():                         # everything on class rows is computed redundantly
            if ctx.normed:
                mu, r, n = ctx.bn
                dy, g_gamma, g_beta = _single_bn_bwd(g, y, mu, r, n, ctx.training, gamma)
            else:
                dy = g
        w4t, w3t, w2t = wo.transpose(w4), wo.transpose(w3), wo.transpose(w2)
        g_w4 = wo.gemm_tn(dy, a3)
        g_b4 = _colsum(dy)
        dh3 = wo.gemm_nt(dy, w4t, mask=a3)                       # [T,4F]
        tot3 = _colsum(dh3)
        tot3_16 = _bf(tot3[None])
        g_w3 = torch.empty(H, H, dtype=F32, device=g.device)
        wo.gemm_tn(dh3, hcat, out=g_w3[:, :3 * F])
        wo.gemm_tn(tot3_16, u, out=g_w3[:, 3 * F:])
        g_u = wo.gemm_nt(tot3_16, w3t[3 * F:], want="f32")
        dh = wo.gemm_nt(dh3, w3t[:3 * F])                        # [T,3F] bf16: [dx_t | dagg]
        g_x_t = dh[:, :F].contiguous()
        dagg = dh[:, F:]
        g_w2 = wo.gemm_tn(dagg, asum)
        g_b2 = wo.colstats(1, dagg, roww=cnt)[0]
        dasum = wo.gemm_nt(dagg, w2t, want="f32")                # [T,2F] table gathered by tgt
        dht = wo.gather_mask(dasum, wt.tgt, wt.div, a_t)         # [E,2F] bf16
        w1t = wo.transpose(w1)
        dRs = wo.segsum(wt.fibres, dht, want="bf16")             # [S,2F]
        g_b1 = _colsum(dRs)
        g_w1 = torch.empty(M2, M2, dtype=F32, device=g.device)
        wo.gemm_tn(dRs, x_s, out=g_w1[:, :F])
        wo.gemm_tn(dht, x_e, out=g_w1[:, F:])
        g_b1, g_w1 = _shard.allreduce_packed([g_b1, g_w1])
        g_x_s = wo.gemm_nt(dRs, w1t[:F])
        g_x_e = wo.gemm_nt(dht, w1t[F:])
        return (None, None, None, g_x_s, g_x_t, g_x_e, _like(g_u, u), _like(g_w1, w1), _like(g_b1, b1), _like(g_w2, w2),
                _like(g_b2, b2), _like(g_w3, w3), _like(tot3, b3), _like(g_w4, w4), _like(g_b4, b4),
                None if g_gamma is None else _like(g_gamma, gamma), None if g_beta is None else _like(g_beta, beta),
                None, None, None)


# ------------------------------------------------------------------------------------------------ GlobalModel
def _rms_fwd(x, w, eps):
    r = torch.rsqrt((x * x).mean(-1, keepdim=True) + eps)
    return x * r * w, r


def _rms_bwd(g, x, r, w):
    gw = g * w
    return r * gw - x * r ** 3 * (gw * x).mean(-1, keepdim=True), (g * x * r).sum(0)


class WideGlobalFunction(torch.autograd.Function):
    """GlobalModel (reference src/gnn.py:195-223): mean pools as column sums over the node rows, the
    [1,3F] MLP as two M=1 GEMMs, the two RMSNorms as O(F) vector math."""

    @staticmethod
    def forward(ctx, normed, x_s, x_t, u, w1, b1, w2, b2, rms_w, eps):
        x_s, x_t, u, w1, w2 = (t.contiguous() for t in (x_s, x_t, u, w1, w2))
        S, F = x_s.shape
        T = x_t.shape[0]
        n_s = float(_shard.total_count(S))                       # fibres over all shards (host int, cached)
        mean_s = _shard.allreduce_sum(_colsum(x_s)) / n_s
        mean_t = _colsum(x_t) / T
        hcat = torch.cat([u.float().reshape(1, F), mean_s[None], mean_t[None]], 1)
        hcat16 = _bf(hcat)
        a = wo.gemm_nt(hcat16, w1, bias=_f32(b1), act=True)      # [1,3F] bf16
        y = wo.gemm_nt(a, w2, bias=_f32(b2), want="f32")         # [1,F]
        if normed:
            w = rms_w.float()
            o1, r1 = _rms_fwd(y, w, eps)
            o2, r2 = _rms_fwd(o1, w, eps)
            ctx.rms = (o1, r1, r2)
            out = o2
        else:
            out = y
        ctx.normed, ctx.S_total = normed, n_s
        ctx.save_for_backward(x_s, x_t, u, w1, b1, w2, b2, rms_w, hcat16, a, y)
        return _bf(out).reshape(u.shape)

    @staticmethod
    def backward(ctx, g):
        x_s, x_t, u, w1, b1, w2, b2, rms_w, hcat16, a, y = ctx.saved_tensors
        S, F = x_s.shape
        T = x_t.shape[0]
        g32 = g.float().reshape(1, F)
        g_rms = None
        if ctx.normed:
            w = rms_w.float()
            o1, r1, r2 = ctx.rms
            d1, gw2 = _rms_bwd(g32, o1, r2, w)
            dy, gw1 = _rms_bwd(d1, y, r1, w)
            g_rms = _like(gw1 + gw2, rms_w)
        else:
            dy = g32
        dy16 = _bf(dy)
        g_w2 = wo.gemm_tn(dy16, a)
        dh = wo.gemm_nt(dy16, wo.transpose(w2), mask=a)          # [1,3F] bf16
        g_w1 = wo.gemm_tn(dh, hcat16)
        dcat = wo.gemm_nt(dh, wo.transpose(w1), want="f32")      # [1,3F]
        g_u = _bf(dcat[:, :F]).reshape(u.shape)
        g_x_s = _bf((dcat[:, F:2 * F] / ctx.S_total).expand(S, F))
        g_x_t = _bf((dcat[:, 2 * F:] / T).expand(T, F))
        return (None, g_x_s, g_x_t, g_u, _like(g_w1, w1), _like(dh.float()[0], b1), _like(g_w2, w2), _like(dy[0], b2),
                g_rms, None)


# ------------------------------------------------------------------------------------------------ time head
class WideTimeHeadFunction(torch.autograd.Function):
    """GNN.edge_prediction (reference src/gnn.py:307-312) on the wide path: the first layer is a tcgen05 GEMM with
    the LeakyReLU in its epilogue, the F -> 1 layer, softplus and scale are one streaming kernel."""

    @staticmethod
    def forward(ctx, scale, x_e, w1, b1, w2, b2):
        x_e, w1 = x_e.contiguous(), w1.contiguous()
        a = wo.gemm_nt(x_e, w1, bias=_f32(b1), act=True)                     # [E,F] bf16
        w2f, b2f = _f32(w2.reshape(-1)), _f32(b2.reshape(-1))
        pred, time, _, _ = wo.head_fwd(a, w2f, b2f, scale)
        ctx.scale = float(scale)
        ctx.save_for_backward(x_e, w1, b1, w2, b2, a, w2f, pred)
        return wo.cast(time, x_e.dtype)

    @staticmethod
    def backward(ctx, g):
        x_e, w1, b1, w2, b2, a, w2f, pred = ctx.saved_tensors
        E = x_e.shape[0]
        gp, da = wo.head_bwd(a, w2f, pred, _f32(g.reshape(-1)), ctx.scale)
        g_w2 = wo.colstats(1, a, roww=gp[:E])[0]                              # sum_e gp[e] a[e]
        g_b2 = _colsum(gp.view(-1, 2)).sum().reshape(1)
        g_w1 = wo.gemm_tn(da, x_e)
        g_b1 = _colsum(da)
        g_x_e = wo.gemm_nt(da, wo.transpose(w1))
        g_w1, g_b1, g_w2, g_b2 = _shard.allreduce_packed([g_w1, g_b1, g_w2, g_b2])
        return None, g_x_e, _like(g_w1, w1), _like(g_b1, b1), _like(g_w2.reshape(w2.shape), w2), _like(g_b2.reshape(b2.shape), b2)


def integer_times(tgt, x_e, w1, b1, w2, b2, scale, class_hours):
    """(time, visits, time_int), fp32 [E] each: the integer definition of DESIGN.md section 8 on the wide path;
    `tgt` = int32 class of every edge."""
    with torch.no_grad():
        a = wo.gemm_nt(x_e.contiguous(), w1.contiguous(), bias=_f32(b1), act=True)
        _, time, visits, time_int = wo.head_fwd(a, _f32(w2.reshape(-1)), _f32(b2.reshape(-1)), scale,
                                                class_hours=class_hours.float().contiguous(), tgt=tgt.contiguous(),
                                                T=int(class_hours.shape[0]))
    return time, visits, time_int

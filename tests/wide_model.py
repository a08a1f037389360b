"""TEST INFRASTRUCTURE ONLY -- the reference Block (src/gnn.py:73-259) with the bf16 roundings of the wide path.

`block_rounded(..., rounding=WIDE_ROUNDING)` is the oracle's Block in fp64 except that the tensors the tensor-core
path stores in bf16 between its GEMMs (the hidden activations a1 / a_s of the per-edge MLPs and the module outputs)
are rounded to bf16 in the forward, with a straight-through gradient.  With `rounding=()` it IS the oracle
(tests/test_wide_model.py holds it to oracle/block_oracle.py at 1e-10).

Why it exists: the gradient of this network is discontinuous in its forward values (LeakyReLU derivative masks; the
third / fourth standardised moments scale like std^-3, std^-4), so ANY forward executed in bf16 -- the reference's own
path run in torch.bfloat16 included -- has gradients that differ from the fp64 gradient by 3e-2 .. 3e-1 norm-wise while
its forward outputs agree to < 1e-2 (tools/bf16_rounding_model.py, profiles/r02_bf16_rounding_model.txt: rounding a1
ALONE already moves grad x_e by 7e-2).  Against this model the mask decisions coincide with the kernels', so the
backward kernels can be held to the plain 1e-2 of the north star (tests/test_gpu_wide_parity.py).
"""
import torch
import torch.nn.functional as Fn

# forward tensors the wide path keeps in bf16 with the default wide.PREC ("z32,m32,at32,node32")
WIDE_ROUNDING = ("a1", "xe2", "a_s", "xs2", "xt2", "u_in", "ga", "u2")
# all bf16 storage points of the all-bf16 variant (PFS_WIDE_PREC=none)
ALL_ROUNDING = WIDE_ROUNDING + ("z", "m", "hcat", "a3", "asum", "agg", "a3t")


class _RoundGrad(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        return g.bfloat16().to(g.dtype)


def _lrelu(x):
    return torch.where(x > 0, x, 0.1 * x)


def _bn(y, g, b, eps=1e-5):
    mu = y.mean(0)
    var = ((y - mu) ** 2).mean(0)
    return (y - mu) / torch.sqrt(var + eps) * g + b


def _rms(w, x, eps):
    return x * torch.rsqrt((x * x).mean(-1, keepdim=True) + eps) * w


def block_rounded(sd, edge_index, x_s, x_t, x_e, u, rounding=WIDE_ROUNDING, grad_rounding=(), rms_eps=None):
    """(x_s', x_t', x_e', u') of one train-mode, normed Block; `sd` holds fp64 tensors keyed like a Block state_dict."""
    R, RG = set(rounding), set(grad_rounding)

    def rnd(name, x):
        if name in RG:
            x = _RoundGrad.apply(x)
        if name in R:
            x = x + (x.bfloat16().to(x.dtype) - x).detach()      # straight-through rounding
        return x

    src, tgt = edge_index[0], edge_index[1]
    S, T, E = x_s.shape[0], x_t.shape[0], x_e.shape[0]
    dt = x_e.dtype
    if rms_eps is None:
        rms_eps = float(torch.finfo(dt).eps)
    p = lambda k: sd[k]
    # EdgeModel (src/gnn.py:98-101): MLP, then the BatchNorm twice
    h = torch.cat([x_s[src], x_t[tgt], x_e, u.expand(E, -1)], 1) @ p("edge_model.0.weight").T + p("edge_model.0.bias")
    a1 = rnd("a1", _lrelu(rnd("h1", h)))
    z = rnd("z", a1 @ p("edge_model.2.weight").T + p("edge_model.2.bias"))
    g, b = p("edge_model.norm.weight"), p("edge_model.norm.bias")
    xe2 = rnd("xe2", _bn(_bn(z, g, b), g, b))
    # SModel (src/gnn.py:135-154)
    q = lambda k: p("s_model." + k)
    a_s = rnd("a_s", _lrelu(rnd("hs", torch.cat([x_t[tgt], xe2], 1) @ q("node_mlp_1.0.weight").T + q("node_mlp_1.0.bias"))))
    m = rnd("m", a_s @ q("node_mlp_1.2.weight").T + q("node_mlp_1.2.bias"))
    cnt = torch.zeros(S, dtype=dt).index_add(0, src, torch.ones(E, dtype=dt)).clamp(min=1)[:, None]
    ssum = lambda v: torch.zeros(S, v.shape[1], dtype=dt).index_add(0, src, v)
    mean = ssum(m) / cnt
    var = Fn.leaky_relu(ssum(m * m) / cnt - mean ** 2)
    std = torch.sqrt(var + 1e-6)
    skew = ssum((m - mean[src]) ** 3) / cnt / std ** 3
    kurt = ssum((m - mean[src]) ** 4) / cnt / std ** 4
    mean, var, skew, kurt = (torch.nan_to_num(t, nan=0.0) for t in (mean, var, skew, kurt))
    std = torch.sqrt(var + 1e-6)
    hcat = torch.cat([x_s] + [rnd("hcat", t) for t in (mean, std, skew, kurt)] + [u.expand(S, -1)], 1)
    a3 = rnd("a3", _lrelu(rnd("h3", hcat @ q("node_mlp_2.0.weight").T + q("node_mlp_2.0.bias"))))
    xs2 = rnd("xs2", _bn(a3 @ q("node_mlp_2.2.weight").T + q("node_mlp_2.2.bias"), q("norm.weight"), q("norm.bias")))
    # TModel (src/gnn.py:187-192); the sum of W2 a + b2 over a class commutes with the Linear
    q = lambda k: p("t_model." + k)
    a_t = rnd("a_t", _lrelu(rnd("ht", torch.cat([xs2[src], xe2], 1) @ q("node_mlp_1.0.weight").T + q("node_mlp_1.0.bias"))))
    asum = rnd("asum", torch.zeros(T, a_t.shape[1], dtype=dt).index_add(0, tgt, a_t))
    cntt = torch.zeros(T, dtype=dt).index_add(0, tgt, torch.ones(E, dtype=dt))[:, None]
    agg = rnd("agg", asum @ q("node_mlp_1.2.weight").T + cntt * q("node_mlp_1.2.bias"))
    a3t = rnd("a3t", _lrelu(rnd("h3t", torch.cat([x_t, agg, u.expand(T, -1)], 1) @ q("node_mlp_2.0.weight").T + q("node_mlp_2.0.bias"))))
    xt2 = rnd("xt2", _bn(a3t @ q("node_mlp_2.2.weight").T + q("node_mlp_2.2.bias"), q("norm.weight"), q("norm.bias")))
    # GlobalModel (src/gnn.py:220-223): mean pools, MLP, RMSNorm twice
    q = lambda k: p("global_model." + k)
    gh = rnd("u_in", torch.cat([u, xs2.mean(0, keepdim=True), xt2.mean(0, keepdim=True)], -1))
    ga = rnd("ga", _lrelu(gh @ q("0.weight").T + q("0.bias")))
    y = ga @ q("2.weight").T + q("2.bias")
    w = q("norm.weight")
    u2 = rnd("u2", _rms(w, _rms(w, y, rms_eps), rms_eps))
    return xs2, xt2, xe2, u2

"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference message-passing layer.

This file is the parity oracle for the hot path.  Only `tests/`,
`__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of `bench.py` may import
it; the product package (`pfs-neural-net_b200/`) never does.

It restates, as pure functions over a `state_dict`-style mapping, what the reference's
`nn.Module`s in `/root/reference/src/gnn.py` compute.  It is written for any floating dtype so
the same code gives the fp64 ground truth and the fp32 "what the reference would print" result.
Gradients come from torch autograd over these functions (exactly how the reference gets its own).

PARITY PIN: the reference has no tests or golden vectors of its own, so this
restatement is pinned against outputs of the *unmodified* reference modules run in the build
container (`oracle/ref_loader.py` + `oracle/make_golden.py` -> `tests/golden/*.pt`; checked by
`tests/test_oracle_golden.py`, and live against the imported reference when /root/reference
exists).  Third-party arithmetic: `torch_scatter.scatter` (PyPI torch-scatter, unpinned,
reference README.md:58-60) is restated in `segment_sum` / `segment_mean` from its published
algorithm (zeros().scatter_add_() and a count clamped to >= 1).
"""
import math

import torch
import torch.nn.functional as Fn

LRELU_MLP = 0.1      # reference src/gnn.py:69  (LeakyReLU(0.1) inside every MLP)
LRELU_VAR = 0.01     # reference src/gnn.py:141 (F.leaky_relu default slope on the variance)
BN_EPS = 1e-5        # torch.nn.BatchNorm1d default, reference src/gnn.py:82,118,170
BN_MOMENTUM = 0.1
STD_EPS = 1e-6       # reference src/gnn.py:142,149


# --------------------------------------------------------------------------------------
# third-party scatter (reference call sites src/gnn.py:140,141,143,144,190; src/train.py:48,61)
# --------------------------------------------------------------------------------------
def segment_sum(values, index, num_segments):
    """torch_scatter.scatter(values, index, dim=0, dim_size=num_segments, reduce='sum')."""
    shape = (num_segments,) + tuple(values.shape[1:])
    out = torch.zeros(shape, dtype=values.dtype, device=values.device)
    idx = index.reshape((-1,) + (1,) * (values.dim() - 1)).expand_as(values)
    return out.scatter_add(0, idx, values)


def segment_mean(values, index, num_segments):
    """reduce='mean': sum / clamp(count, min=1) -- empty segments give 0."""
    total = segment_sum(values, index, num_segments)
    count = segment_sum(torch.ones(index.shape, dtype=values.dtype, device=values.device), index, num_segments)
    count = count.clamp(min=1).reshape((-1,) + (1,) * (values.dim() - 1))
    return total / count


# --------------------------------------------------------------------------------------
# building blocks
# --------------------------------------------------------------------------------------
def mlp(sd, prefix, x):
    """reference src/gnn.py:65-71: Linear -> LeakyReLU(0.1) -> Linear, children '0' and '2'."""
    h = Fn.linear(x, sd[prefix + "0.weight"], sd[prefix + "0.bias"])
    h = Fn.leaky_relu(h, LRELU_MLP)
    return Fn.linear(h, sd[prefix + "2.weight"], sd[prefix + "2.bias"])


def batch_norm(sd, prefix, x, training, buffers=None):
    """torch.nn.BatchNorm1d over the rows of x [N, C].

    training: batch statistics (biased variance for the output, unbiased for running_var),
    running buffers updated with momentum 0.1 and num_batches_tracked += 1.  `buffers`, when
    given, is a dict that receives the updated buffer values (the inputs in `sd` are not
    mutated); eval mode normalises with the running buffers.
    """
    w, b = sd[prefix + "weight"], sd[prefix + "bias"]
    rm = sd[prefix + "running_mean"] if buffers is None or prefix + "running_mean" not in buffers \
        else buffers[prefix + "running_mean"]
    rv = sd[prefix + "running_var"] if buffers is None or prefix + "running_var" not in buffers \
        else buffers[prefix + "running_var"]
    if training:
        n = x.shape[0]
        if n <= 1:
            raise ValueError("Expected more than 1 value per channel when training")
        mean = x.mean(dim=0)
        var = ((x - mean) ** 2).mean(dim=0)
        if buffers is not None:
            nbt_key = prefix + "num_batches_tracked"
            nbt = buffers.get(nbt_key, sd.get(nbt_key, torch.zeros((), dtype=torch.long)))
            buffers[prefix + "running_mean"] = (1 - BN_MOMENTUM) * rm.to(x.dtype) + BN_MOMENTUM * mean.detach()
            buffers[prefix + "running_var"] = (1 - BN_MOMENTUM) * rv.to(x.dtype) \
                + BN_MOMENTUM * var.detach() * (n / (n - 1))
            buffers[nbt_key] = nbt + 1
    else:
        mean, var = rm.to(x.dtype), rv.to(x.dtype)
    return (x - mean) / torch.sqrt(var + BN_EPS) * w + b


def rms_norm(weight, x, eps=None):
    """torch.nn.RMSNorm(F) with eps=None -> finfo(x.dtype).eps (reference src/gnn.py:203).  `eps` overrides
    it so that an fp64 run can stand in for the reference executed in another dtype (bf16 parity)."""
    if eps is None:
        eps = torch.finfo(x.dtype).eps
    return x * torch.rsqrt(x.pow(2).mean(dim=-1, keepdim=True) + eps) * weight


# --------------------------------------------------------------------------------------
# the four update modules (reference src/gnn.py:73-223) and Block (src/gnn.py:226-259)
# --------------------------------------------------------------------------------------
def edge_model(sd, prefix, x_s, x_t, edge_index, x_e, u, training=True, normed=True, buffers=None):
    """reference src/gnn.py:73-101.  NOTE the norm is applied TWICE when normed: `EdgeModel`
    subclasses the Sequential `MLP`, so `self.norm` is also its 4th child;
    both applications share one BatchNorm1d (two running-stat updates per forward)."""
    src, tgt = edge_index[0], edge_index[1]
    h = torch.cat([x_s[src], x_t[tgt], x_e, u.expand(x_e.shape[0], -1)], dim=-1)
    z = mlp(sd, prefix, h)
    if not normed:
        return z
    z = batch_norm(sd, prefix + "norm.", z, training, buffers)
    return batch_norm(sd, prefix + "norm.", z, training, buffers)


def source_moments(msg, src, num_src):
    """reference src/gnn.py:140-151: per-fibre mean / std / skew / kurtosis of the messages."""
    mean = segment_mean(msg, src, num_src)
    var = Fn.leaky_relu(segment_mean(msg ** 2, src, num_src) - mean ** 2, LRELU_VAR)
    std = torch.sqrt(var + STD_EPS)
    skew = segment_mean((msg - mean[src]) ** 3, src, num_src) / std ** 3
    kurt = segment_mean((msg - mean[src]) ** 4, src, num_src) / std ** 4
    mean = torch.nan_to_num(mean, nan=0.0)
    var = torch.nan_to_num(var, nan=0.0)
    std = torch.sqrt(var + STD_EPS)
    skew = torch.nan_to_num(skew, nan=0.0)
    kurt = torch.nan_to_num(kurt, nan=0.0)
    return mean, std, skew, kurt


def s_model(sd, prefix, x_s, x_t, edge_index, x_e, u, training=True, normed=True, buffers=None):
    """reference src/gnn.py:104-154."""
    src, tgt = edge_index[0], edge_index[1]
    msg = mlp(sd, prefix + "node_mlp_1.", torch.cat([x_t[tgt], x_e], dim=1))
    mean, std, skew, kurt = source_moments(msg, src, x_s.shape[0])
    h = torch.cat([x_s, mean, std, skew, kurt, u.expand(x_s.shape[0], -1)], dim=-1)
    y = mlp(sd, prefix + "node_mlp_2.", h)
    return batch_norm(sd, prefix + "norm.", y, training, buffers) if normed else y


def t_model(sd, prefix, x_s, x_t, edge_index, x_e, u, training=True, normed=True, buffers=None):
    """reference src/gnn.py:157-192."""
    src, tgt = edge_index[0], edge_index[1]
    msg = mlp(sd, prefix + "node_mlp_1.", torch.cat([x_s[src], x_e], dim=1))
    agg = segment_sum(msg, tgt, x_t.shape[0])
    h = torch.cat([x_t, agg, u.expand(x_t.shape[0], -1)], dim=-1)
    y = mlp(sd, prefix + "node_mlp_2.", h)
    return batch_norm(sd, prefix + "norm.", y, training, buffers) if normed else y


def global_model(sd, prefix, x_s, x_t, u, normed=True, rms_eps=None):
    """reference src/gnn.py:195-223; RMSNorm applied twice for the same Sequential reason."""
    h = torch.cat([u, x_s.mean(dim=0, keepdim=True), x_t.mean(dim=0, keepdim=True)], dim=-1)
    y = mlp(sd, prefix, h)
    if not normed:
        return y
    w = sd[prefix + "norm.weight"]
    return rms_norm(w, rms_norm(w, y, rms_eps), rms_eps)


def block(sd, prefix, edge_index, x_s, x_t, x_e, u, training=True, normed=True, buffers=None,
          e_model=True, s_model_on=True, t_model_on=True, u_model=True, rms_eps=None):
    """reference src/gnn.py:243-259: edge -> source -> target -> global, each stage consuming
    the previous stage's outputs."""
    if e_model:
        x_e = edge_model(sd, prefix + "edge_model.", x_s, x_t, edge_index, x_e, u, training, normed, buffers)
    if s_model_on:
        x_s = s_model(sd, prefix + "s_model.", x_s, x_t, edge_index, x_e, u, training, normed, buffers)
    if t_model_on:
        x_t = t_model(sd, prefix + "t_model.", x_s, x_t, edge_index, x_e, u, training, normed, buffers)
    if u_model:
        u = global_model(sd, prefix + "global_model.", x_s, x_t, u, normed, rms_eps)
    return x_s, x_t, x_e, u


def gnn_forward(sd, num_blocks, edge_index, x_s, x_t, x_e, u, training=True, normed=True, buffers=None):
    """reference src/gnn.py:280-305: encoders, `num_blocks` Blocks; returns (x_s, x_t, x_e, u)."""
    x_s = mlp(sd, "encoder_s.", x_s)
    x_t = mlp(sd, "encoder_t.", x_t)
    for b in range(num_blocks):
        x_s, x_t, x_e, u = block(sd, "mpb.%d." % b, edge_index, x_s, x_t, x_e, u, training, normed, buffers)
    return x_s, x_t, x_e, u


def edge_prediction(sd, x_e, scale=1.0):
    """reference src/gnn.py:307-312.  `GNN.round` (src/gnn.py:321-325) tests the bound method
    `self.train`, which is always truthy, so it is the identity in train AND eval mode."""
    return Fn.softplus(mlp(sd, "decoder_e.", x_e)) * scale


def node_prediction(sd, x_s, scale=1.0):
    """reference src/gnn.py:314-319 (round is the identity, see edge_prediction)."""
    return torch.softmax(mlp(sd, "decoder_s.", x_s), dim=-1) * scale


def integer_times(time, class_hours, tgt):
    """The integer definition adopted for "rounded integer times" (SURVEY.md section 8a row A8):
    visits = round-half-even(time / T_i[tgt]) and time_int = visits * T_i, which is what the
    reference plots (src/train.py:257) and what `softfloor` relaxes (src/train.py:21-27,43-49)."""
    per_visit = class_hours[tgt]
    visits = torch.round(time.reshape(-1) / per_visit)
    return visits, visits * per_visit


# --------------------------------------------------------------------------------------
# helpers shared by tests and bench
# --------------------------------------------------------------------------------------
def cast_state(sd, dtype):
    return {k: (v.to(dtype) if v.is_floating_point() else v.clone()) for k, v in sd.items()}


def complete_bipartite(num_src, num_tgt):
    """canonical dense order of reference src/train.py:94: e = k * T + i."""
    k = torch.arange(num_src).repeat_interleave(num_tgt)
    i = torch.arange(num_tgt).repeat(num_src)
    return torch.stack([k, i])


def block_param_shapes(F):
    """state_dict entries of one reference Block, in registration order."""
    def mlp_shapes(p, d1, d2, d3):
        return [(p + "0.weight", (d2, d1)), (p + "0.bias", (d2,)), (p + "2.weight", (d3, d2)), (p + "2.bias", (d3,))]

    def bn(p):
        return [(p + "weight", (F,)), (p + "bias", (F,)), (p + "running_mean", (F,)), (p + "running_var", (F,)),
                (p + "num_batches_tracked", ())]
    out = []
    out += mlp_shapes("edge_model.", 4 * F, 4 * F, F) + bn("edge_model.norm.")
    out += mlp_shapes("s_model.node_mlp_1.", 2 * F, 2 * F, 2 * F) + mlp_shapes("s_model.node_mlp_2.", 10 * F, 10 * F, F)
    out += bn("s_model.norm.")
    out += mlp_shapes("t_model.node_mlp_1.", 2 * F, 2 * F, 2 * F) + mlp_shapes("t_model.node_mlp_2.", 4 * F, 4 * F, F)
    out += bn("t_model.norm.")
    out += mlp_shapes("global_model.", 3 * F, 3 * F, F) + [("global_model.norm.weight", (F,))]
    return out


def random_block_state(F, seed=0, dtype=torch.float32, affine_jitter=True):
    """Random Block weights with torch's default Linear init distribution (U(-1/sqrt(fan_in), ..))
    and, per SURVEY.md section 4.2, NON-trivial norm affine parameters so double-norm bugs show."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for name, shape in block_param_shapes(F):
        if name.endswith("num_batches_tracked"):
            sd[name] = torch.zeros((), dtype=torch.long)
        elif name.endswith("running_mean"):
            sd[name] = torch.zeros(shape, dtype=dtype)
        elif name.endswith("running_var"):
            sd[name] = torch.ones(shape, dtype=dtype)
        elif ".norm." in name and name.endswith("weight"):
            sd[name] = (0.5 + torch.rand(shape, generator=g, dtype=torch.float64)).to(dtype) if affine_jitter \
                else torch.ones(shape, dtype=dtype)
        elif ".norm." in name and name.endswith("bias"):
            sd[name] = (2 * torch.rand(shape, generator=g, dtype=torch.float64) - 1).to(dtype) if affine_jitter \
                else torch.zeros(shape, dtype=dtype)
        else:
            fan_in = shape[1] if len(shape) == 2 else None
            if fan_in is None:
                # bias: fan_in of the matching weight, registered just before it
                fan_in = sd[name.replace("bias", "weight")].shape[1]
            bound = 1.0 / math.sqrt(fan_in)
            sd[name] = ((2 * torch.rand(shape, generator=g, dtype=torch.float64) - 1) * bound).to(dtype)
    return sd

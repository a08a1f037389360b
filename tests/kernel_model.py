"""TEST INFRASTRUCTURE ONLY -- the algebra the CUDA kernels implement, written out in torch.

Each function mirrors one C-ABI op of include/pfs_b200.h (same decomposition, same saved tensors,
hand-derived backward -- no autograd), so that (a) the hand-derived formulas are checked on the
CPU, in fp64, against autograd over oracle/block_oracle.py (tests/test_kernel_model.py), and
(b) a failing GPU parity test can be localised phase by phase.  Nothing in the product imports it.

Decomposition:
  * first-layer split: W1.[x_s[src] | x_t[tgt] | x_e | u] = P_s[src] + P_t[tgt] + W1_e.x_e with per-node
    tables P_s = x_s.W1_s^T and P_t = x_t.W1_t^T + W1_u.u + b1 (reference src/gnn.py:100);
  * the edge BatchNorm applied twice (src/gnn.py:82,101) in closed form: one statistics pass;
  * scatter-sum of the target messages commuted with their last Linear (src/gnn.py:189-190).
"""
import torch

EPS_BN = 1e-5
MOM = 0.1
SLOPE = 0.1
SLOPE_VAR = 0.01
STD_EPS = 1e-6


def lrelu(x, s=SLOPE):
    return torch.where(x > 0, x, s * x)


def dlrelu(x, s=SLOPE):
    return torch.where(x > 0, torch.ones_like(x), torch.full_like(x, s))


def seg_sum(v, idx, n):
    out = torch.zeros((n,) + tuple(v.shape[1:]), dtype=v.dtype)
    return out.index_add(0, idx, v)


def bn_train_fwd(y, gamma, beta):
    n = y.shape[0]
    mu = y.mean(0)
    var = ((y - mu) ** 2).mean(0)
    r = 1 / torch.sqrt(var + EPS_BN)
    out = (y - mu) * r * gamma + beta
    return out, mu, var, r


def bn_train_bwd(g, y, mu, r, gamma):
    xh = (y - mu) * r
    gbar = g.mean(0)
    gx = (g * xh).mean(0)
    dy = gamma * r * (g - gbar - xh * gx)
    return dy, (g * xh).sum(0), g.sum(0)


# ----------------------------------------------------------------------------- edge model
def edge_fwd(p, x_s, x_t, src, tgt, x_e, u, training=True, normed=True, rm=None, rv=None):
    F = x_e.shape[1]
    W1, b1, W2, b2 = p["0.weight"], p["0.bias"], p["2.weight"], p["2.bias"]
    Ps = x_s @ W1[:, :F].T
    Pt = x_t @ W1[:, F:2 * F].T + (u @ W1[:, 3 * F:].T + b1)
    h1 = Ps[src] + Pt[tgt] + x_e @ W1[:, 2 * F:3 * F].T
    z = lrelu(h1) @ W2.T + b2
    saved = {}
    if not normed:
        return z, saved
    gam, bet = p["norm.weight"], p["norm.bias"]
    if training:
        n = z.shape[0]
        mu = z.mean(0)
        var = ((z - mu) ** 2).mean(0)
        r1 = 1 / torch.sqrt(var + EPS_BN)
        var2 = gam * gam * var * r1 * r1
        r2 = 1 / torch.sqrt(var2 + EPS_BN)
        A = gam * gam * r1 * r2
        shift = bet - A * mu
        saved.update(mu=mu, var=var, r1=r1, r2=r2)
        unb = n / (n - 1)
        rm1 = (1 - MOM) * rm + MOM * mu
        rv1 = (1 - MOM) * rv + MOM * var * unb
        saved["running_mean"] = (1 - MOM) * rm1 + MOM * bet      # mean of the first BN output is beta
        saved["running_var"] = (1 - MOM) * rv1 + MOM * var2 * unb
    else:
        a = gam / torch.sqrt(rv + EPS_BN)
        A = a * a
        shift = a * (bet - rm - a * rm) + bet
        saved.update(a=a)
    saved.update(A=A, shift=shift)
    return A * z + shift, saved


def edge_bwd(p, x_s, x_t, src, tgt, x_e, u, xe2, saved, g, training=True, normed=True, rm=None, rv=None):
    F = x_e.shape[1]
    S, T, E = x_s.shape[0], x_t.shape[0], x_e.shape[0]
    W1, b1, W2, b2 = p["0.weight"], p["0.bias"], p["2.weight"], p["2.bias"]
    grads = {}
    if not normed:
        dz = g
    elif training:
        gam, bet = p["norm.weight"], p["norm.bias"]
        r1, r2, var = saved["r1"], saved["r2"], saved["var"]
        xh1 = (xe2 - bet) / (gam * gam * r2)          # normalised first-BN input recovered from the output
        gbar, mgx = g.mean(0), (g * xh1).mean(0)
        s = gam * r2
        q = var * r1 * r1
        kappa = s * s + 1 - s * s * q
        dz = gam * gam * r1 * r2 * (g - gbar - xh1 * mgx * kappa)
        grads["norm.weight"] = E * mgx * s * (2 - s * s * q)
        grads["norm.bias"] = E * gbar
    else:
        gam, bet = p["norm.weight"], p["norm.bias"]
        a, A, shift = saved["a"], saved["A"], saved["shift"]
        c = 1 / torch.sqrt(rv + EPS_BN)
        z = (xe2 - shift) / A
        y1 = a * (z - rm) + bet
        dz = A * g
        grads["norm.weight"] = (g * c * ((y1 - rm) + a * (z - rm))).sum(0)
        grads["norm.bias"] = (g * (a + 1)).sum(0)
    # recompute the hidden layer
    Ps = x_s @ W1[:, :F].T
    Pt = x_t @ W1[:, F:2 * F].T + (u @ W1[:, 3 * F:].T + b1)
    h1 = Ps[src] + Pt[tgt] + x_e @ W1[:, 2 * F:3 * F].T
    a1 = lrelu(h1)
    dh1 = (dz @ W2) * dlrelu(h1)
    grads["2.weight"] = dz.T @ a1
    grads["2.bias"] = dz.sum(0)
    dPs, dPt = seg_sum(dh1, src, S), seg_sum(dh1, tgt, T)
    tot = dPt.sum(0)
    dW1 = torch.zeros_like(W1)
    dW1[:, :F] = dPs.T @ x_s
    dW1[:, F:2 * F] = dPt.T @ x_t
    dW1[:, 2 * F:3 * F] = dh1.T @ x_e
    dW1[:, 3 * F:] = torch.outer(tot, u[0])
    grads["0.weight"] = dW1
    grads["0.bias"] = tot
    dx_s = dPs @ W1[:, :F]
    dx_t = dPt @ W1[:, F:2 * F]
    dx_e = dh1 @ W1[:, 2 * F:3 * F]
    du = (tot @ W1[:, 3 * F:])[None]
    return dx_s, dx_t, dx_e, du, grads


# ----------------------------------------------------------------------------- source model
def source_fwd(p, x_s, x_t, src, tgt, xe2, u, training=True, normed=True, rm=None, rv=None):
    F = xe2.shape[1]
    S = x_s.shape[0]
    W1, b1, W2, b2 = (p["node_mlp_1." + k] for k in ("0.weight", "0.bias", "2.weight", "2.bias"))
    W3, b3, W4, b4 = (p["node_mlp_2." + k] for k in ("0.weight", "0.bias", "2.weight", "2.bias"))
    Qt = x_t @ W1[:, :F].T + b1
    hs = Qt[tgt] + xe2 @ W1[:, F:].T
    m = lrelu(hs) @ W2.T + b2
    cnt = seg_sum(torch.ones(src.shape[0], dtype=m.dtype), src, S).clamp(min=1)[:, None]
    mean = seg_sum(m, src, S) / cnt
    ex2 = seg_sum(m * m, src, S) / cnt
    d = m - mean[src]
    c2, c3, c4 = (seg_sum(d ** k, src, S) / cnt for k in (2, 3, 4))
    var_raw = ex2 - mean * mean
    var = lrelu(var_raw, SLOPE_VAR)
    std = torch.sqrt(var + STD_EPS)
    skew, kurt = c3 / std ** 3, c4 / std ** 4
    stats = dict(cnt=cnt, mean=mean, ex2=ex2, c2=c2, c3=c3, c4=c4)      # what the edge kernel writes
    # finite values assumed here (nan_to_num is the identity); the kernels clamp like torch does
    hcat = torch.cat([x_s, mean, std, skew, kurt], dim=1)
    b3eff = b3 + u @ W3[:, 9 * F:].T
    h3 = hcat @ W3[:, :9 * F].T + b3eff
    y = lrelu(h3) @ W4.T + b4
    saved = dict(stats=stats, y=y)
    if not normed:
        return y, saved
    gam, bet = p["norm.weight"], p["norm.bias"]
    if training:
        out, mu, var_b, r = bn_train_fwd(y, gam, bet)
        n = y.shape[0]
        saved.update(mu=mu, r=r, running_mean=(1 - MOM) * rm + MOM * mu,
                     running_var=(1 - MOM) * rv + MOM * var_b * n / (n - 1))
    else:
        out = (y - rm) / torch.sqrt(rv + EPS_BN) * gam + bet
    return out, saved


def source_bwd(p, x_s, x_t, src, tgt, xe2, u, saved, g, training=True, normed=True, rm=None, rv=None):
    F = xe2.shape[1]
    S, T = x_s.shape[0], x_t.shape[0]
    W1, b1, W2, b2 = (p["node_mlp_1." + k] for k in ("0.weight", "0.bias", "2.weight", "2.bias"))
    W3, b3, W4, b4 = (p["node_mlp_2." + k] for k in ("0.weight", "0.bias", "2.weight", "2.bias"))
    grads = {}
    y = saved["y"]
    if not normed:
        dy = g
    elif training:
        dy, grads["norm.weight"], grads["norm.bias"] = bn_train_bwd(g, y, saved["mu"], saved["r"], p["norm.weight"])
    else:
        c = 1 / torch.sqrt(rv + EPS_BN)
        dy = g * c * p["norm.weight"]
        grads["norm.weight"] = (g * (y - rm) * c).sum(0)
        grads["norm.bias"] = g.sum(0)
    st = saved["stats"]
    cnt, mean, ex2, c2, c3, c4 = (st[k] for k in ("cnt", "mean", "ex2", "c2", "c3", "c4"))
    var_raw = ex2 - mean * mean
    var = lrelu(var_raw, SLOPE_VAR)
    std = torch.sqrt(var + STD_EPS)
    skew, kurt = c3 / std ** 3, c4 / std ** 4
    hcat = torch.cat([x_s, mean, std, skew, kurt], dim=1)
    b3eff = b3 + u @ W3[:, 9 * F:].T
    h3 = hcat @ W3[:, :9 * F].T + b3eff
    a3 = lrelu(h3)
    grads["node_mlp_2.2.weight"] = dy.T @ a3
    grads["node_mlp_2.2.bias"] = dy.sum(0)
    dh3 = (dy @ W4) * dlrelu(h3)
    tot3 = dh3.sum(0)
    dW3 = torch.zeros_like(W3)
    dW3[:, :9 * F] = dh3.T @ hcat
    dW3[:, 9 * F:] = torch.outer(tot3, u[0])
    grads["node_mlp_2.0.weight"] = dW3
    grads["node_mlp_2.0.bias"] = tot3
    du = (tot3 @ W3[:, 9 * F:])[None]
    dh = dh3 @ W3[:, :9 * F]
    dx_s = dh[:, :F]
    d_mean, d_std, d_skew, d_kurt = (dh[:, F + 2 * F * i:F + 2 * F * (i + 1)] for i in range(4))
    # moments backward -> per-fibre polynomial coefficients
    d_c3 = d_skew / std ** 3
    d_c4 = d_kurt / std ** 4
    d_std_tot = d_std - 3 * c3 / std ** 4 * d_skew - 4 * c4 / std ** 5 * d_kurt
    d_var_raw = d_std_tot / (2 * std) * dlrelu(var_raw, SLOPE_VAR)
    d_mu = d_mean - 2 * mean * d_var_raw - 3 * c2 * d_c3 - 4 * c3 * d_c4
    A0, A1, A2, A3 = d_mu / cnt, 2 * d_var_raw / cnt, 3 * d_c3 / cnt, 4 * d_c4 / cnt
    # per-edge pass (recompute the messages)
    Qt = x_t @ W1[:, :F].T + b1
    hs = Qt[tgt] + xe2 @ W1[:, F:].T
    as_ = lrelu(hs)
    m = as_ @ W2.T + b2
    d = m - mean[src]
    dm = A0[src] + A1[src] * m + A2[src] * d * d + A3[src] * d * d * d
    grads["node_mlp_1.2.weight"] = dm.T @ as_
    grads["node_mlp_1.2.bias"] = dm.sum(0)
    dhs = (dm @ W2) * dlrelu(hs)
    dQt = seg_sum(dhs, tgt, T)
    dW1 = torch.zeros_like(W1)
    dW1[:, :F] = dQt.T @ x_t
    dW1[:, F:] = dhs.T @ xe2
    grads["node_mlp_1.0.weight"] = dW1
    grads["node_mlp_1.0.bias"] = dQt.sum(0)
    dx_t = dQt @ W1[:, :F]
    dxe2 = dhs @ W1[:, F:]
    return dx_s, dx_t, dxe2, du, grads


# ----------------------------------------------------------------------------- target model
def target_fwd(p, xs2, x_t, src, tgt, xe2, u, training=True, normed=True, rm=None, rv=None):
    F = xe2.shape[1]
    T = x_t.shape[0]
    W1, b1, W2, b2 = (p["node_mlp_1." + k] for k in ("0.weight", "0.bias", "2.weight", "2.bias"))
    W3, b3, W4, b4 = (p["node_mlp_2." + k] for k in ("0.weight", "0.bias", "2.weight", "2.bias"))
    Rs = xs2 @ W1[:, :F].T + b1
    ht = Rs[src] + xe2 @ W1[:, F:].T
    asum = seg_sum(lrelu(ht), tgt, T)
    cnt = seg_sum(torch.ones(tgt.shape[0], dtype=xe2.dtype), tgt, T)[:, None]
    agg = asum @ W2.T + cnt * b2
    b3eff = b3 + u @ W3[:, 3 * F:].T
    h3 = torch.cat([x_t, agg], 1) @ W3[:, :3 * F].T + b3eff
    y = lrelu(h3) @ W4.T + b4
    saved = dict(asum=asum, cnt=cnt, y=y)
    if not normed:
        return y, saved
    gam, bet = p["norm.weight"], p["norm.bias"]
    if training:
        out, mu, var_b, r = bn_train_fwd(y, gam, bet)
        n = y.shape[0]
        saved.update(mu=mu, r=r, running_mean=(1 - MOM) * rm + MOM * mu,
                     running_var=(1 - MOM) * rv + MOM * var_b * n / (n - 1))
    else:
        out = (y - rm) / torch.sqrt(rv + EPS_BN) * gam + bet
    return out, saved


def target_bwd(p, xs2, x_t, src, tgt, xe2, u, saved, g, training=True, normed=True, rm=None, rv=None):
    F = xe2.shape[1]
    S = xs2.shape[0]
    W1, b1, W2, b2 = (p["node_mlp_1." + k] for k in ("0.weight", "0.bias", "2.weight", "2.bias"))
    W3, b3, W4, b4 = (p["node_mlp_2." + k] for k in ("0.weight", "0.bias", "2.weight", "2.bias"))
    grads = {}
    y, asum, cnt = saved["y"], saved["asum"], saved["cnt"]
    if not normed:
        dy = g
    elif training:
        dy, grads["norm.weight"], grads["norm.bias"] = bn_train_bwd(g, y, saved["mu"], saved["r"], p["norm.weight"])
    else:
        c = 1 / torch.sqrt(rv + EPS_BN)
        dy = g * c * p["norm.weight"]
        grads["norm.weight"] = (g * (y - rm) * c).sum(0)
        grads["norm.bias"] = g.sum(0)
    agg = asum @ W2.T + cnt * b2
    hcat = torch.cat([x_t, agg], 1)
    b3eff = b3 + u @ W3[:, 3 * F:].T
    h3 = hcat @ W3[:, :3 * F].T + b3eff
    a3 = lrelu(h3)
    grads["node_mlp_2.2.weight"] = dy.T @ a3
    grads["node_mlp_2.2.bias"] = dy.sum(0)
    dh3 = (dy @ W4) * dlrelu(h3)
    tot3 = dh3.sum(0)
    dW3 = torch.zeros_like(W3)
    dW3[:, :3 * F] = dh3.T @ hcat
    dW3[:, 3 * F:] = torch.outer(tot3, u[0])
    grads["node_mlp_2.0.weight"] = dW3
    grads["node_mlp_2.0.bias"] = tot3
    du = (tot3 @ W3[:, 3 * F:])[None]
    dh = dh3 @ W3[:, :3 * F]
    dx_t, dagg = dh[:, :F], dh[:, F:]
    grads["node_mlp_1.2.weight"] = dagg.T @ asum
    grads["node_mlp_1.2.bias"] = (cnt * dagg).sum(0)
    dasum = dagg @ W2                                  # [T, 2F] table gathered by tgt
    Rs = xs2 @ W1[:, :F].T + b1
    ht = Rs[src] + xe2 @ W1[:, F:].T
    dht = dasum[tgt] * dlrelu(ht)
    dRs = seg_sum(dht, src, S)
    dW1 = torch.zeros_like(W1)
    dW1[:, :F] = dRs.T @ xs2
    dW1[:, F:] = dht.T @ xe2
    grads["node_mlp_1.0.weight"] = dW1
    grads["node_mlp_1.0.bias"] = dRs.sum(0)
    dxs2 = dRs @ W1[:, :F]
    dxe2 = dht @ W1[:, F:]
    return dxs2, dx_t, dxe2, du, grads


# ----------------------------------------------------------------------------- global model
def rms_fwd(x, w):
    eps = torch.finfo(x.dtype).eps
    r = torch.rsqrt((x * x).mean(-1, keepdim=True) + eps)
    return x * r * w, r


def rms_bwd(g, x, r, w):
    gw = g * w
    dx = r * gw - x * r ** 3 * (gw * x).mean(-1, keepdim=True)
    return dx, (g * x * r).sum(0)


def global_fwd(p, xs2, xt2, u, normed=True):
    W1, b1, W2, b2 = p["0.weight"], p["0.bias"], p["2.weight"], p["2.bias"]
    hcat = torch.cat([u, xs2.mean(0, keepdim=True), xt2.mean(0, keepdim=True)], 1)
    h = hcat @ W1.T + b1
    y = lrelu(h) @ W2.T + b2
    if not normed:
        return y, {}
    w = p["norm.weight"]
    o1, r1 = rms_fwd(y, w)
    o2, r2 = rms_fwd(o1, w)
    return o2, dict(y=y, o1=o1, r1=r1, r2=r2)


def global_bwd(p, xs2, xt2, u, saved, g, normed=True):
    F = u.shape[1]
    W1, b1, W2, b2 = p["0.weight"], p["0.bias"], p["2.weight"], p["2.bias"]
    grads = {}
    hcat = torch.cat([u, xs2.mean(0, keepdim=True), xt2.mean(0, keepdim=True)], 1)
    h = hcat @ W1.T + b1
    a = lrelu(h)
    if normed:
        w = p["norm.weight"]
        d1, gw2 = rms_bwd(g, saved["o1"], saved["r2"], w)
        dy, gw1 = rms_bwd(d1, saved["y"], saved["r1"], w)
        grads["norm.weight"] = gw1 + gw2
    else:
        dy = g
    grads["2.weight"] = dy.T @ a
    grads["2.bias"] = dy.sum(0)
    dh = (dy @ W2) * dlrelu(h)
    grads["0.weight"] = dh.T @ hcat
    grads["0.bias"] = dh.sum(0)
    dcat = dh @ W1
    du = dcat[:, :F]
    dxs2 = dcat[:, F:2 * F].expand(xs2.shape[0], -1) / xs2.shape[0]
    dxt2 = dcat[:, 2 * F:].expand(xt2.shape[0], -1) / xt2.shape[0]
    return dxs2, dxt2, du, grads

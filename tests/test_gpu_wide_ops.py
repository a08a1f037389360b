"""GPU tests of the wide-feature primitives (tcgen05 bf16 GEMMs, segmented reductions, statistics)
against plain torch fp32/fp64 references of the same op on the same bf16-rounded inputs."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return torch.device("cuda:0")


def _rel(a, b):
    a, b = a.double(), b.double()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


@pytest.mark.parametrize("M,N,K", [(128, 128, 128), (300, 512, 128), (1000, 128, 512), (77, 256, 64), (5, 32, 32),
                                   (4096, 1280, 1152), (1, 512, 128), (513, 72, 40)])
def test_gemm_nt_plain(M, N, K):
    from pfs_neural_net_b200 import wide_ops as wo
    dev = _dev()
    g = torch.Generator(device="cpu").manual_seed(M * 7 + N + K)
    A = torch.randn(M, K, generator=g).to(dev).bfloat16()
    B = torch.randn(N, K, generator=g).to(dev).bfloat16()
    ref = A.double() @ B.double().T
    c16, c32 = wo.gemm_nt(A, B, want="both")
    torch.cuda.synchronize()
    assert _rel(c32, ref) < 1e-5
    assert _rel(c16, ref) < 6e-3
    assert torch.equal(c16, c32.bfloat16())


def test_gemm_nt_epilogue():
    from pfs_neural_net_b200 import wide_ops as wo
    dev = _dev()
    g = torch.Generator(device="cpu").manual_seed(3)
    S, T, K, N = 37, 24, 64, 256
    E = S * T
    A = torch.randn(E, K, generator=g).to(dev).bfloat16()
    B = torch.randn(N, K, generator=g).to(dev).bfloat16()
    bias = torch.randn(N, generator=g).to(dev)
    rs = torch.rand(E, generator=g).to(dev)
    t0 = torch.randn(S, N, generator=g).to(dev)
    t1 = torch.randn(T, N, generator=g).to(dev)
    mask = torch.randn(E, N, generator=g).to(dev).bfloat16()
    mask[0, :8] = 0
    src = torch.arange(E, device=dev) // T
    tgt = torch.arange(E, device=dev) % T
    base = A.double() @ B.double().T + bias.double() * rs.double()[:, None] + t0.double()[src] + t1.double()[tgt]
    act = torch.where(base > 0, base, 0.1 * base)
    ref = act * torch.where(mask.double() > 0, 1.0, 0.1)
    # dense (div / mod) addressing
    out = wo.gemm_nt(A, B, bias=bias, bias_rowscale=rs, tab0=t0, div0=T, tab1=t1, mod1=T, mask=mask, act=True, want="f32")
    assert _rel(out, ref) < 1e-5
    # explicit index arrays, shuffled rows
    perm = torch.randperm(E, generator=g).to(dev)
    out2 = wo.gemm_nt(A[perm].contiguous(), B, bias=bias, bias_rowscale=rs[perm].contiguous(), tab0=t0,
                      idx0=src[perm].int().contiguous(), tab1=t1, idx1=tgt[perm].int().contiguous(),
                      mask=mask[perm].contiguous(), act=True, want="f32")
    assert _rel(out2, ref[perm]) < 1e-5
    # strided operands: column slices of wider matrices, strided bf16 output
    W = torch.randn(N, 3 * K, generator=g).to(dev).bfloat16()
    buf = torch.zeros(E, 2 * N, dtype=torch.bfloat16, device=dev)
    wo.gemm_nt(A, W[:, K:2 * K], out_bf16=buf[:, N:], want="none")
    ref3 = A.double() @ W[:, K:2 * K].double().T
    assert _rel(buf[:, N:], ref3) < 6e-3
    assert buf[:, :N].abs().max().item() == 0


@pytest.mark.parametrize("E,J,K", [(256, 128, 128), (1000, 512, 128), (5000, 128, 512), (70, 256, 64), (3, 32, 32),
                                   (20000, 1280, 1152), (1, 512, 128), (12345, 40, 72)])
def test_gemm_tn(E, J, K):
    from pfs_neural_net_b200 import wide_ops as wo
    dev = _dev()
    g = torch.Generator(device="cpu").manual_seed(E + J + K)
    D = torch.randn(E, J, generator=g).to(dev).bfloat16()
    X = torch.randn(E, K, generator=g).to(dev).bfloat16()
    ref = D.double().T @ X.double()
    out = wo.gemm_tn(D, X)
    torch.cuda.synchronize()
    assert _rel(out, ref) < 2e-5
    out2 = wo.gemm_tn(D, X)
    assert torch.equal(out, out2)            # deterministic
    # into a column slice, accumulating
    big = torch.ones(J, 2 * K + 8, dtype=torch.float32, device=dev)
    wo.gemm_tn(D, X, out=big[:, 8:8 + K], accumulate=True)
    assert _rel(big[:, 8:8 + K], ref + 1) < 2e-5
    assert (big[:, :8] == 1).all() and (big[:, 8 + K:] == 1).all()


def test_colstats_rowmap():
    from pfs_neural_net_b200 import wide_ops as wo
    dev = _dev()
    g = torch.Generator(device="cpu").manual_seed(5)
    R, C = 10007, 96
    x = (torch.randn(R, C, generator=g) * 3 + 50).to(dev).bfloat16()
    st = wo.colstats(0, x)
    xd = x.double()
    assert _rel(st[0], xd.mean(0)) < 1e-6
    assert _rel(st[1], ((xd - xd.mean(0)) ** 2).sum(0)) < 1e-4
    y = torch.randn(R, C, generator=g).to(dev)      # fp32 input
    st = wo.colstats(0, y)
    assert _rel(st[0], y.double().mean(0)) < 1e-5
    assert _rel(st[1], ((y.double() - y.double().mean(0)) ** 2).sum(0)) < 1e-5
    gg = torch.randn(R, C, generator=g).to(dev).bfloat16()
    p0, p1 = torch.randn(C, generator=g).to(dev), torch.rand(C, generator=g).to(dev) + 0.5
    w = torch.rand(R, generator=g).to(dev)
    st = wo.colstats(1, gg, v=y, p0=p0, p1=p1, roww=w)
    assert _rel(st[0], (w.double()[:, None] * gg.double()).sum(0)) < 1e-5
    assert _rel(st[1], (w.double()[:, None] * gg.double() * (y.double() - p0.double()) * p1.double()).sum(0)) < 1e-5
    a, b, c2 = torch.randn(C, generator=g).to(dev), torch.randn(C, generator=g).to(dev), torch.randn(C, generator=g).to(dev)
    o = wo.rowmap(0, x, a, b)
    assert _rel(o, a.double() * xd + b.double()) < 6e-3
    o = wo.rowmap(1, gg, a, b, v=y, p0=p0, p1=p1, c2=c2)
    assert _rel(o, a.double() * (gg.double() - b.double() - (y.double() - p0.double()) * p1.double() * c2.double())) < 6e-3


def test_segsum_and_moments():
    from pfs_neural_net_b200 import wide_ops as wo
    dev = _dev()
    g = torch.Generator(device="cpu").manual_seed(9)
    S, T, C = 301, 40, 64
    E = S * T
    x = torch.randn(E, C, generator=g).to(dev).bfloat16()
    xd = x.double().view(S, T, C)
    fib = wo.Segments(0, S, S, T)
    cls = wo.Segments(1, T, S, T)
    assert _rel(wo.segsum(fib, x), xd.sum(1)) < 1e-5
    o32, o16 = wo.segsum(cls, x, want="both")
    assert _rel(o32, xd.sum(0)) < 1e-5
    assert torch.equal(o16, o32.bfloat16())
    # listed segments: random subset, ragged, with empty segments
    keep = torch.rand(E, generator=g) < 0.3
    keep[:T] = False                          # fibre 0 is empty
    rows = torch.nonzero(keep).flatten()
    seg_of = rows // T
    order = torch.argsort(seg_of, stable=True)
    lst = rows[order].int().to(dev)
    counts = torch.bincount(seg_of, minlength=S)
    ptr = torch.zeros(S + 1, dtype=torch.int32)
    ptr[1:] = counts.cumsum(0)
    ptr = ptr.to(dev)
    seg = wo.Segments(2, S, ptr=ptr, lst=lst)
    ref = torch.zeros(S, C, dtype=torch.float64, device=dev).index_add(0, seg_of.to(dev), x.double()[rows.to(dev)])
    assert _rel(wo.segsum(seg, x), ref) < 1e-5
    # moments (dense fibres)
    mo = wo.moments_fwd(fib, x)
    mean = xd.mean(1)
    d = xd - mean[:, None]
    for i, r in enumerate((mean, (xd ** 2).mean(1), (d ** 2).mean(1), (d ** 3).mean(1), (d ** 4).mean(1))):
        assert _rel(mo[:, i], r) < 1e-5, i
    mo2 = wo.moments_fwd(seg, x)
    assert mo2[0].abs().max().item() == 0     # empty fibre: all zeros, no NaN
    k = int(torch.nonzero(counts > 3)[0])
    xs = x.double()[lst[ptr[k]:ptr[k + 1]].long()]
    assert _rel(mo2[k, 0], xs.mean(0)) < 1e-5
    assert _rel(mo2[k, 3], ((xs - xs.mean(0)) ** 3).mean(0)) < 1e-4


def test_cast_transpose_gather():
    from pfs_neural_net_b200 import wide_ops as wo
    dev = _dev()
    g = torch.Generator(device="cpu").manual_seed(11)
    w = torch.randn(70, 200, generator=g).to(dev).bfloat16()
    assert torch.equal(wo.transpose(w), w.T.contiguous())
    assert torch.equal(wo.transpose(w[:, 40:104]), w[:, 40:104].T.contiguous())
    assert torch.equal(wo.cast(w, torch.float32), w.float())
    f = torch.randn(1000, generator=g).to(dev)
    assert torch.equal(wo.cast(f, torch.bfloat16), f.bfloat16())
    tab = torch.randn(12, 64, generator=g).to(dev)
    act = torch.randn(120, 64, generator=g).to(dev).bfloat16()
    o = wo.gather_mask(tab, None, 12, act)
    ref = tab[torch.arange(120, device=dev) % 12] * torch.where(act.float() > 0, 1.0, 0.1)
    assert _rel(o, ref) < 6e-3

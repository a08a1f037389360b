"""Python wrappers of the wide-feature primitives of libpfs_b200.so (include/pfs_b200.h, "Wide-feature
path"): one ctypes call per function, on the current CUDA stream, PyTorch owning the memory.
bf16 tensors in, bf16 / fp32 tensors out; no CPU fallback (CPU tensors raise)."""
import ctypes as ct

import torch

from . import _abi

BF16, F32 = torch.bfloat16, torch.float32
_DT = {torch.bfloat16: 0, torch.float32: 1}


def _stream(dev):
    return torch.cuda.current_stream(dev).cuda_stream


def _chk(t, dtype=None, name="tensor"):
    if t is None:
        return
    if not t.is_cuda:
        raise _abi.PfsError("pfs_b200 wide kernels need CUDA tensors, %s is on %s (no CPU fallback)" % (name, t.device))
    if dtype is not None and t.dtype != dtype:
        raise _abi.PfsError("%s must be %s, got %s" % (name, dtype, t.dtype))


def _rows2d(t, name):
    """2-D view with unit inner stride (column slices of row-major matrices are fine)."""
    if t.dim() != 2 or (t.shape[1] > 1 and t.stride(1) != 1):
        raise _abi.PfsError("%s must be a 2-D tensor with contiguous rows, got shape %s strides %s"
                            % (name, tuple(t.shape), t.stride()))
    return t


_ws_cache = {}


def _workspace(dev, nbytes):
    key = str(dev)
    ws = _ws_cache.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(int(nbytes * 1.25) + 4096, dtype=torch.uint8, device=dev)
        _ws_cache[key] = ws
    return ws


class Segments:
    """Row segments for the segmented reductions (fibres or classes; dense or listed)."""

    def __init__(self, mode, nseg, S=0, T=0, ptr=None, lst=None):
        self.mode, self.nseg, self.S, self.T, self.ptr, self.list = mode, int(nseg), int(S), int(T), ptr, lst
        s = _abi.WideSegments()
        s.mode, s.nseg, s.S, s.T = mode, self.nseg, self.S, self.T
        s.ptr, s.list = _abi.ptr(ptr), _abi.ptr(lst)
        self.struct = s


def gemm_nt(A, B, bias=None, bias_rowscale=None, tab0=None, idx0=None, div0=0, tab1=None, idx1=None, mod1=0,
            mask=None, act=False, out_bf16=None, out_f32=None, want="bf16", A2=None, a2_mod=0, bias_rows=None,
            bias_rows_div=0):
    """epilogue(A[M,K] @ B[N,K]^T) on tcgen05; returns the bf16 and/or fp32 result (want in {'bf16','f32','both'}).
    Dense-layout operands: A2 [a2_mod, K2] contracted first ([A2[m % a2_mod] | A[m]] @ B^T with B [N, K2 + K]) and
    bias_rows [*, N] added per 128-row tile (row m // bias_rows_div)."""
    A, B = _rows2d(A, "A"), _rows2d(B, "B")
    _chk(A, BF16, "A"), _chk(B, BF16, "B")
    M, K = A.shape
    N = B.shape[0]
    K2 = 0
    if A2 is not None:
        A2 = _rows2d(A2, "A2")
        _chk(A2, BF16, "A2")
        K2 = A2.shape[1]
        if A2.shape[0] != a2_mod:
            raise _abi.PfsError("gemm_nt: A2 has %d rows, a2_mod is %d" % (A2.shape[0], a2_mod))
    _chk(bias_rows, F32, "bias_rows")
    if B.shape[1] != K + K2:
        raise _abi.PfsError("gemm_nt: A %s (+ A2 %d columns) vs B %s" % (tuple(A.shape), K2, tuple(B.shape)))
    dev = A.device
    for t, n in ((bias, "bias"), (bias_rowscale, "bias_rowscale"), (tab0, "tab0"), (tab1, "tab1")):
        _chk(t, F32, n)
    for t, n in ((idx0, "idx0"), (idx1, "idx1")):
        _chk(t, torch.int32, n)
    _chk(mask, BF16, "mask")
    if want in ("bf16", "both") and out_bf16 is None:
        out_bf16 = torch.empty(M, N, dtype=BF16, device=dev)
    if want in ("f32", "both") and out_f32 is None:
        out_f32 = torch.empty(M, N, dtype=F32, device=dev)
    a = _abi.WideGemmArgs()
    a.A, a.lda, a.B, a.ldb, a.M, a.N, a.K = A.data_ptr(), A.stride(0), B.data_ptr(), B.stride(0), M, N, K
    a.bias, a.bias_rowscale = _abi.ptr(bias), _abi.ptr(bias_rowscale)
    a.tab0, a.idx0, a.div0 = _abi.ptr(tab0), _abi.ptr(idx0), int(div0)
    a.tab1, a.idx1, a.mod1 = _abi.ptr(tab1), _abi.ptr(idx1), int(mod1)
    if mask is not None:
        mask = _rows2d(mask, "mask")
        a.mask, a.ldmask = mask.data_ptr(), mask.stride(0)
    a.act = int(bool(act))
    if A2 is not None:
        a.A2, a.lda2, a.K2, a.a2_mod = A2.data_ptr(), A2.stride(0), K2, int(a2_mod)
    if bias_rows is not None:
        a.bias_rows, a.bias_rows_div = bias_rows.contiguous().data_ptr(), int(bias_rows_div)
    if out_bf16 is not None:
        out_bf16 = _rows2d(out_bf16, "out_bf16")
        a.out_bf16, a.ldc = out_bf16.data_ptr(), out_bf16.stride(0)
    if out_f32 is not None:
        out_f32 = _rows2d(out_f32, "out_f32")
        a.out_f32, a.ldf = out_f32.data_ptr(), out_f32.stride(0)
    with torch.cuda.device(dev):
        a.stream = _stream(dev)
        _abi.check(_abi.load_library().pfs_wide_gemm_nt(ct.byref(a)), "pfs_wide_gemm_nt")
    if want == "both":
        return out_bf16, out_f32
    return out_bf16 if want == "bf16" else out_f32


def gemm_tn(D, X, out=None, accumulate=False):
    """out[J,Kx] (+)= D[E,J]^T @ X[E,Kx] (fp32), contraction over rows, deterministic split reduction."""
    D, X = _rows2d(D, "D"), _rows2d(X, "X")
    _chk(D, BF16, "D"), _chk(X, BF16, "X")
    E, J = D.shape
    Kx = X.shape[1]
    if X.shape[0] != E:
        raise _abi.PfsError("gemm_tn: D %s vs X %s" % (tuple(D.shape), tuple(X.shape)))
    dev = D.device
    if out is None:
        out = torch.empty(J, Kx, dtype=F32, device=dev)
    out = _rows2d(out, "out")
    _chk(out, F32, "out")
    lib = _abi.load_library()
    ws = _workspace(dev, lib.pfs_wide_gemm_tn_workspace(E, J, Kx))
    with torch.cuda.device(dev):
        _abi.check(lib.pfs_wide_gemm_tn(D.data_ptr(), D.stride(0), X.data_ptr(), X.stride(0), E, J, Kx, out.data_ptr(),
                                        out.stride(0), int(bool(accumulate)), ws.data_ptr(), ws.numel(), _stream(dev)),
                   "pfs_wide_gemm_tn")
    return out


def colstats(kind, g, v=None, p0=None, p1=None, roww=None):
    """kind 0: (mean, M2) of the columns of g; kind 1: (sum w g, sum w g (v - p0) p1).  Returns fp32 [2, C]."""
    g = _rows2d(g, "g")
    R, C = g.shape
    dev = g.device
    _chk(g, None, "g")
    out = torch.empty(2, C, dtype=F32, device=dev)
    lib = _abi.load_library()
    ws = _workspace(dev, lib.pfs_wide_colstats_workspace(R, C))
    if v is not None:
        v = _rows2d(v, "v")
        _chk(v, None, "v")
    with torch.cuda.device(dev):
        _abi.check(lib.pfs_wide_colstats(kind, g.data_ptr(), _DT[g.dtype], g.stride(0), _abi.ptr(v),
                                         _DT[v.dtype] if v is not None else 0, v.stride(0) if v is not None else 0,
                                         _abi.ptr(p0), _abi.ptr(p1), _abi.ptr(roww), R, C, out.data_ptr(), ws.data_ptr(),
                                         ws.numel(), _stream(dev)), "pfs_wide_colstats")
    return out


def rowmap(kind, x, a, b, v=None, p0=None, p1=None, c2=None, out=None):
    """kind 0: a x + b; kind 1: a (x - b - (v - p0) p1 c2); per-column fp32 coefficients, bf16 output."""
    x = _rows2d(x, "x")
    R, C = x.shape
    dev = x.device
    _chk(x, None, "x")
    if out is None:
        out = torch.empty(R, C, dtype=BF16, device=dev)
    out = _rows2d(out, "out")
    if v is not None:
        v = _rows2d(v, "v")
    with torch.cuda.device(dev):
        _abi.check(_abi.load_library().pfs_wide_rowmap(
            kind, x.data_ptr(), _DT[x.dtype], x.stride(0), _abi.ptr(v), _DT[v.dtype] if v is not None else 0,
            v.stride(0) if v is not None else 0, a.data_ptr(), b.data_ptr(), _abi.ptr(p0), _abi.ptr(p1), _abi.ptr(c2),
            R, C, out.data_ptr(), out.stride(0), _stream(dev)), "pfs_wide_rowmap")
    return out


def segsum(seg, x, want="f32"):
    """Segmented row sums of x (bf16 or fp32) -> fp32 and/or bf16 [nseg, C]."""
    x = _rows2d(x, "x")
    _chk(x, None, "x")
    C = x.shape[1]
    dev = x.device
    o32 = torch.empty(seg.nseg, C, dtype=F32, device=dev) if want in ("f32", "both") else None
    o16 = torch.empty(seg.nseg, C, dtype=BF16, device=dev) if want in ("bf16", "both") else None
    lib = _abi.load_library()
    ws = _workspace(dev, lib.pfs_wide_segsum_workspace(ct.byref(seg.struct), C))
    with torch.cuda.device(dev):
        _abi.check(lib.pfs_wide_segsum(ct.byref(seg.struct), x.data_ptr(), _DT[x.dtype], x.stride(0), C, _abi.ptr(o32), _abi.ptr(o16),
                                       ws.data_ptr(), ws.numel(), _stream(dev)), "pfs_wide_segsum")
    if want == "both":
        return o32, o16
    return o32 if want == "f32" else o16


def moments_fwd(seg, m):
    _chk(m, None, "m")
    m = m.contiguous()
    C = m.shape[1]
    out = torch.empty(seg.nseg, 5, C, dtype=F32, device=m.device)
    with torch.cuda.device(m.device):
        _abi.check(_abi.load_library().pfs_wide_moments_fwd(ct.byref(seg.struct), m.data_ptr(), _DT[m.dtype], C, out.data_ptr(),
                                                            _stream(m.device)), "pfs_wide_moments_fwd")
    return out


def source_hcat(x_s, moments, with_lo=False):
    """[x_s | mean | std | skew | kurt] bf16 [S, 9F]; with_lo: [S, 17F], the statistics' bf16 remainders appended."""
    _chk(x_s, BF16, "x_s")
    x_s = x_s.contiguous()
    S, F = x_s.shape
    out = torch.empty(S, (17 if with_lo else 9) * F, dtype=BF16, device=x_s.device)
    with torch.cuda.device(x_s.device):
        _abi.check(_abi.load_library().pfs_wide_source_hcat(x_s.data_ptr(), moments.data_ptr(), S, F, out.data_ptr(),
                                                            out.stride(0), int(with_lo), _stream(x_s.device)),
                   "pfs_wide_source_hcat")
    return out


def split(x, out=None):
    """fp32 [R, C] -> bf16 [R, 2C] = [hi | lo] with x = hi + lo to ~2^-17 (node-level GEMM operands)."""
    x = _rows2d(x, "x")
    _chk(x, F32, "x")
    R, C = x.shape
    if out is None:
        out = torch.empty(R, 2 * C, dtype=BF16, device=x.device)
    out = _rows2d(out, "out")
    with torch.cuda.device(x.device):
        _abi.check(_abi.load_library().pfs_wide_split(x.data_ptr(), x.stride(0), R, C, out.data_ptr(), out.stride(0),
                                                      _stream(x.device)), "pfs_wide_split")
    return out


def source_coef(seg, dh, moments, F):
    _chk(dh, F32, "dh")
    dh = dh.contiguous()
    S = dh.shape[0]
    dx_s = torch.empty(S, F, dtype=BF16, device=dh.device)
    coef = torch.empty(S, 4, 2 * F, dtype=F32, device=dh.device)
    with torch.cuda.device(dh.device):
        _abi.check(_abi.load_library().pfs_wide_source_coef(ct.byref(seg.struct), dh.data_ptr(), moments.data_ptr(), S, F,
                                                            dx_s.data_ptr(), coef.data_ptr(), _stream(dh.device)),
                   "pfs_wide_source_coef")
    return dx_s, coef


def source_dm(m, moments, coef, src, T):
    _chk(m, None, "m")
    E, C = m.shape
    out = torch.empty(E, C, dtype=BF16, device=m.device)
    with torch.cuda.device(m.device):
        _abi.check(_abi.load_library().pfs_wide_source_dm(m.data_ptr(), _DT[m.dtype], moments.data_ptr(), coef.data_ptr(), _abi.ptr(src),
                                                          int(T), E, C, out.data_ptr(), _stream(m.device)),
                   "pfs_wide_source_dm")
    return out


def source_dm_seg(seg, m, moments, coef):
    """source_dm with one CTA per fibre segment."""
    _chk(m, None, "m")
    E, C = m.shape
    out = torch.empty(E, C, dtype=BF16, device=m.device)
    with torch.cuda.device(m.device):
        _abi.check(_abi.load_library().pfs_wide_source_dm_seg(ct.byref(seg.struct), m.data_ptr(), _DT[m.dtype], moments.data_ptr(),
                                                              coef.data_ptr(), C, out.data_ptr(), _stream(m.device)),
                   "pfs_wide_source_dm_seg")
    return out


def gather_mask(tab, idx, mod, act):
    _chk(tab, F32, "tab"), _chk(act, BF16, "act")
    E, C = act.shape
    out = torch.empty(E, C, dtype=BF16, device=act.device)
    with torch.cuda.device(act.device):
        _abi.check(_abi.load_library().pfs_wide_gather_mask(tab.data_ptr(), _abi.ptr(idx), int(mod), act.data_ptr(), E, C,
                                                            out.data_ptr(), _stream(act.device)), "pfs_wide_gather_mask")
    return out


def cast(t, dtype):
    """bf16 <-> fp32 copy through the library's cast kernel."""
    if t.dtype == dtype:
        return t
    _chk(t, None, "t")
    t = t.contiguous()
    out = torch.empty(t.shape, dtype=dtype, device=t.device)
    if t.numel() == 0:
        return out
    with torch.cuda.device(t.device):
        _abi.check(_abi.load_library().pfs_wide_cast(t.data_ptr(), _DT[t.dtype], out.data_ptr(), _DT[dtype], t.numel(),
                                                     _stream(t.device)), "pfs_wide_cast")
    return out


def transpose(w):
    """bf16 [R, C] (rows may be strided) -> contiguous [C, R]."""
    w = _rows2d(w, "w")
    _chk(w, BF16, "w")
    R, C = w.shape
    out = torch.empty(C, R, dtype=BF16, device=w.device)
    with torch.cuda.device(w.device):
        _abi.check(_abi.load_library().pfs_wide_transpose(w.data_ptr(), R, C, w.stride(0), out.data_ptr(), _stream(w.device)),
                   "pfs_wide_transpose")
    return out


def head_fwd(a, w2, b2, scale, class_hours=None, tgt=None, T=0):
    """Time head tail: pred = a . w2 + b2, time = softplus(pred) * scale (+ integer visits / times)."""
    _chk(a, BF16, "a"), _chk(w2, F32, "w2"), _chk(b2, F32, "b2")
    a = a.contiguous()
    E, F = a.shape
    dev = a.device
    pred = torch.empty(E, dtype=F32, device=dev)
    time = torch.empty(E, dtype=F32, device=dev)
    visits = time_int = None
    if class_hours is not None:
        _chk(class_hours, F32, "class_hours")
        visits, time_int = torch.empty(E, dtype=F32, device=dev), torch.empty(E, dtype=F32, device=dev)
    with torch.cuda.device(dev):
        _abi.check(_abi.load_library().pfs_wide_head_fwd(
            a.data_ptr(), w2.data_ptr(), b2.data_ptr(), float(scale), E, F, _abi.ptr(class_hours), _abi.ptr(tgt), int(T),
            pred.data_ptr(), time.data_ptr(), _abi.ptr(visits), _abi.ptr(time_int), _stream(dev)), "pfs_wide_head_fwd")
    return pred, time, visits, time_int


def head_bwd(a, w2, pred, g_time, scale):
    """(gp [E] fp32 padded to an even length with zeros, da [E,F] bf16) of the time head."""
    _chk(a, BF16, "a"), _chk(g_time, F32, "g_time")
    E, F = a.shape
    dev = a.device
    gp = torch.zeros(E + (E & 1), dtype=F32, device=dev)
    da = torch.empty(E, F, dtype=BF16, device=dev)
    with torch.cuda.device(dev):
        _abi.check(_abi.load_library().pfs_wide_head_bwd(a.data_ptr(), w2.data_ptr(), pred.data_ptr(), g_time.data_ptr(),
                                                         float(scale), E, F, gp.data_ptr(), da.data_ptr(), _stream(dev)),
                   "pfs_wide_head_bwd")
    return gp, da

"""TEST INFRASTRUCTURE ONLY -- stand-in for the third-party `torch_scatter` package.

The reference imports `scatter`/`scatter_mean` (reference src/gnn.py:4, src/train.py:4) from
PyPI `torch-scatter` (version unpinned, README.md:58-60; not vendored under /root/reference and
not installable offline).  Upstream torch-scatter 2.1.x implements `reduce='sum'` as
`zeros(dim_size).scatter_add_(dim, index, src)` and `reduce='mean'` as that sum divided by a
second scatter_add_ of ones, clamped to >= 1.  That published algorithm is restated here; it is
the *definition* the parity report is anchored on.

Only the call shapes the reference uses are supported: a 1-D `index` along `dim` 0 (or the only
dim of a 1-D `src`) -- reference src/gnn.py:140-144,190 and src/train.py:48,61.
"""
import torch


def _norm_dim(src, dim):
    return dim + src.dim() if dim < 0 else dim


def scatter_sum(src, index, dim=-1, out=None, dim_size=None):
    dim = _norm_dim(src, dim)
    if index.dim() != 1 or dim != 0:
        if not (src.dim() == 1 and index.dim() == 1):
            raise ValueError("shim supports a 1-D index along dim 0 only")
    if dim_size is None:
        dim_size = int(index.max()) + 1 if index.numel() else 0
    shape = list(src.shape)
    shape[0] = dim_size
    if out is None:
        out = torch.zeros(shape, dtype=src.dtype, device=src.device)
    idx = index.reshape([-1] + [1] * (src.dim() - 1)).expand_as(src)
    return out.scatter_add_(0, idx, src)


def scatter_mean(src, index, dim=-1, out=None, dim_size=None):
    total = scatter_sum(src, index, dim, out, dim_size)
    count = scatter_sum(torch.ones(index.shape, dtype=src.dtype, device=src.device), index, 0,
                        None, total.shape[0])
    count = count.clamp_(min=1).reshape([-1] + [1] * (src.dim() - 1))
    return total.div_(count)


def scatter(src, index, dim=-1, out=None, dim_size=None, reduce='sum'):
    if reduce in ('sum', 'add'):
        return scatter_sum(src, index, dim, out, dim_size)
    if reduce == 'mean':
        return scatter_mean(src, index, dim, out, dim_size)
    raise ValueError(f"reduce={reduce!r} is not used by the reference and not restated here")

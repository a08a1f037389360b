"""On-disk graphs.

* `load_graph(path)`: reads the reference's `graphs/graph-*.pt` -- a pickled `gnn.BipartiteData`
  (torch_geometric `Data` with a `GlobalStorage`, reference src/graph.py:83) -- WITHOUT torch_geometric:
  unknown classes in the pickle are mapped to attribute bags and the tensors come out of the
  `_store._mapping` dict.  Also reads the PyG-free format `save_graph` writes.
* `make_graph(class_info, nfibers, fdim)`: what reference src/graph.py:14-67 builds (a complete
  fibre x class graph with zero edge / fibre / global features), but in the canonical dense order
  e = k*T + i of src/train.py:94 instead of the class-permuted order an unstable argsort leaves in
  `graphs/graph-0.pt`, so the kernels take the index-free dense path.
* `save_graph(path, graph, with_index=True)`: plain dict of tensors, optionally with the int32
  CSR / CSC arrays of the edge list precomputed on the host (`csr_arrays`), so loading a general
  fibre-target visibility graph does not need a device sort.
"""
import pickle

import torch

from .gnn import BipartiteData

FORMAT = "pfs_b200.graph.v1"


class _Bag:
    """Stand-in for a class the unpickler cannot import (torch_geometric / the reference's gnn module)."""

    def __init__(self, *args, **kwargs):
        self._args, self._kwargs = args, kwargs

    def __setstate__(self, state):
        self.__dict__.update(state if isinstance(state, dict) else {"_state": state})


# What a graph file may reference: the tensor-rebuild machinery of torch.save and plain containers.  Every other
# global named by the pickle (torch_geometric / the reference's gnn classes, and anything a crafted file might ask
# for) becomes an inert attribute bag: loading a graph never imports or calls code the file chooses.
_ALLOWED = {
    ("collections", "OrderedDict"), ("builtins", "dict"), ("builtins", "list"), ("builtins", "tuple"), ("builtins", "set"),
    ("builtins", "int"), ("builtins", "float"), ("builtins", "bool"), ("builtins", "str"), ("builtins", "complex"),
    ("torch", "Size"), ("torch", "device"), ("torch", "dtype"),
    ("torch._utils", "_rebuild_tensor_v2"), ("torch._utils", "_rebuild_tensor"), ("torch._utils", "_rebuild_parameter"),
    ("torch._utils", "_rebuild_device_tensor_from_numpy"), ("torch._tensor", "_rebuild_from_type_v2"),
    ("torch.serialization", "_get_layout"),
}
_ALLOWED_TORCH_ATTRS = ("Storage", "Tensor")     # torch.FloatStorage, torch.LongTensor, ...: legacy type tags


class _Unpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if (module, name) in _ALLOWED:
            return super().find_class(module, name)
        if module == "torch" and (name.endswith(_ALLOWED_TORCH_ATTRS) or hasattr(torch, name) and isinstance(
                getattr(torch, name), (torch.dtype, torch.layout, torch.memory_format))):
            return super().find_class(module, name)
        return type(name, (_Bag,), {"__module__": module})


class _PickleModule:
    """`pickle_module` for torch.load: the stock pickle with a tolerant class lookup."""
    __name__ = "pfs_b200_pickle"
    Unpickler = _Unpickler
    load = staticmethod(lambda f, **kw: _Unpickler(f, **kw).load())
    loads = staticmethod(pickle.loads)
    dump, dumps = staticmethod(pickle.dump), staticmethod(pickle.dumps)
    HIGHEST_PROTOCOL = pickle.HIGHEST_PROTOCOL


def _tensors_of(obj):
    if isinstance(obj, dict):
        if obj.get("format") == FORMAT:
            return {k: v for k, v in obj.items() if torch.is_tensor(v)}
        return {k: v for k, v in obj.items() if torch.is_tensor(v)}
    store = getattr(obj, "_store", None)
    mapping = getattr(store, "_mapping", None) if store is not None else None
    if mapping is None:
        mapping = {k: v for k, v in vars(obj).items() if torch.is_tensor(v)}
    return {k: v for k, v in mapping.items() if torch.is_tensor(v)}


def load_graph(path, device=None):
    """BipartiteData from a reference `graph-*.pt` or a `save_graph` file; tensors saved from another device (the
    shipped file was written from `mps`) are mapped to the CPU first."""
    obj = torch.load(path, map_location="cpu", pickle_module=_PickleModule, weights_only=False)
    t = _tensors_of(obj)
    missing = [k for k in ("edge_index", "x_s", "x_t", "x_e", "x_u") if k not in t]
    if missing:
        raise ValueError("%s: not a bipartite graph file (missing %s)" % (path, ", ".join(missing)))
    g = BipartiteData.__new__(BipartiteData)
    for k, v in t.items():
        setattr(g, k, v if device is None else v.to(device))
    g.num_nodes = len(g.x_t)
    return g


def csr_arrays(edge_index, nfibers, nclasses):
    """Host-side int32 CSR (fibre-sorted, stable) and CSC (class-sorted positions) arrays of an edge list: the same
    arrays `pfs_build_topology` produces on the device (include/pfs_b200.h, struct pfs_topology)."""
    src, tgt = edge_index[0].cpu(), edge_index[1].cpu()
    order = torch.sort(src, stable=True).indices
    csr_src, csr_tgt = src[order], tgt[order]
    rowptr = torch.zeros(nfibers + 1, dtype=torch.int64)
    rowptr[1:] = torch.bincount(src, minlength=nfibers).cumsum(0)
    cscq = torch.sort(csr_tgt, stable=True).indices
    colptr = torch.zeros(nclasses + 1, dtype=torch.int64)
    colptr[1:] = torch.bincount(tgt, minlength=nclasses).cumsum(0)
    i32 = torch.int32
    return dict(csr_rowptr=rowptr.to(i32), csr_eid=order.to(i32), csr_src=csr_src.to(i32), csr_tgt=csr_tgt.to(i32),
                csc_colptr=colptr.to(i32), csc_q=cscq.to(i32))


def is_canonical(edge_index, nclasses):
    e = torch.arange(edge_index.shape[1])
    return bool((edge_index[0].cpu() == e // nclasses).all() and (edge_index[1].cpu() == e % nclasses).all())


def make_graph(class_info, nfibers, fdim, pad_class_features=True):
    """Complete fibre x class graph of reference src/graph.py:14-67 in canonical dense order (src/train.py:94)."""
    class_info = torch.as_tensor(class_info, dtype=torch.float32)
    T = class_info.shape[0]
    x_t = class_info
    if pad_class_features and x_t.shape[1] < fdim:            # reference src/graph.py:76: pad the class table to Fdim
        x_t = torch.cat([x_t, torch.zeros(T, fdim - x_t.shape[1])], 1)
    k = torch.arange(nfibers).repeat_interleave(T)
    i = torch.arange(T).repeat(nfibers)
    g = BipartiteData.__new__(BipartiteData)
    g.edge_index = torch.stack([k, i])
    g.x_s = torch.zeros(nfibers, fdim)
    g.x_t, g.num_nodes = x_t, T
    g.x_e = torch.zeros(nfibers * T, fdim)
    g.x_u = torch.zeros(1, fdim)
    return g


def save_graph(path, graph, with_index=True):
    d = {"format": FORMAT}
    for k in ("edge_index", "x_s", "x_t", "x_e", "x_u"):
        d[k] = getattr(graph, k).detach().cpu()
    S, T = d["x_s"].shape[0], d["x_t"].shape[0]
    d["canonical"] = torch.tensor(is_canonical(d["edge_index"], T))
    if with_index and not bool(d["canonical"]):
        d.update(csr_arrays(d["edge_index"], S, T))
    torch.save(d, path)
    return path

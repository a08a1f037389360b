"""The training loss of the reference (`softfloor` + `loss_function`, reference src/train.py:21-80) on
the B200 kernels `pfs_loss_fwd` / `pfs_loss_bwd`: the step right after the
message-passing path, consuming `GNN.edge_prediction`'s times.

`loss_function(gnn, graph, class_info, ...)` keeps the reference's argument meaning and return values
(`(loss, totutils)`, or the 7-tuple of `finaloutput=True`); the module-level globals the reference
reads (`gnn`, NFIBERS, NCLASSES, NFIELDS, TOTAL_TIME, wutils, wvar; src/config.py:16-28) are explicit
arguments with the reference's values as defaults.  The uniform noise of `softfloor` is drawn with
`torch.rand_like` exactly where the reference draws it (so a seeded run consumes the generator the
same way) or can be passed in.  fp32, one graph, dense canonical edge order; no CPU fallback.
"""
import ctypes as ct

import torch

from . import _abi

NOISELEVEL = 0.3          # softfloor default (reference src/train.py:21)


class LossFunction(torch.autograd.Function):
    """loss(time): forward = pfs_loss_fwd, backward = pfs_loss_bwd (gradient w.r.t. the edge times only)."""

    @staticmethod
    def forward(ctx, time, noise, hours, counts, S, T, consts):
        for t, n in ((time, "time"), (noise, "noise"), (hours, "hours"), (counts, "counts")):
            if not t.is_cuda:
                raise _abi.PfsError("pfs_b200 loss needs CUDA tensors, %s is on %s (no CPU fallback)" % (n, t.device))
            if t.dtype != torch.float32:
                raise _abi.PfsError("pfs_b200 loss is fp32, %s is %s" % (n, t.dtype))
        time, noise, hours, counts = (t.contiguous() for t in (time, noise, hours, counts))
        E = S * T
        if time.numel() != E or noise.numel() != E or hours.numel() != T or counts.numel() != T:
            raise _abi.PfsError("loss: expected %d edge times / noise values and %d class rows" % (E, T))
        dev = time.device
        f32 = dict(dtype=torch.float32, device=dev)
        lib = _abi.load_library()
        a = _abi.LossArgs()
        a.S, a.T = S, T
        out = dict(galaxies=torch.empty(E, **f32), time2=torch.empty(E, **f32), fibre_time=torch.empty(S, **f32),
                   n_prime=torch.empty(T, **f32), class_mean=torch.empty(T, **f32), class_coef=torch.empty(T, **f32),
                   scalars=torch.zeros(8, **f32))
        ws = torch.empty(lib.pfs_loss_workspace_bytes(S, T), dtype=torch.uint8, device=dev)
        for k, v in dict(time=time, noise=noise, hours=hours, counts=counts, **out).items():
            setattr(a, k, v.data_ptr())
        sharp_dev = consts.get("sharpness_dev")
        for k, v in consts.items():
            if k != "sharpness_dev":
                setattr(a, k, float(v))
        a.sharpness_dev = _abi.ptr(sharp_dev)
        a.workspace, a.workspace_bytes = ws.data_ptr(), ws.numel()
        with torch.cuda.device(dev):
            a.stream = torch.cuda.current_stream(dev).cuda_stream
            _abi.check(lib.pfs_loss_fwd(ct.byref(a)), "pfs_loss_fwd")
        ctx.S, ctx.T, ctx.consts = S, T, dict(consts)
        ctx.save_for_backward(time, noise, hours, counts, out["time2"], out["fibre_time"], out["class_mean"], out["class_coef"])
        sc = out["scalars"]
        ctx.mark_non_differentiable(sc, out["n_prime"], out["fibre_time"], out["time2"])
        return sc[0].clone(), sc, out["n_prime"], out["fibre_time"], out["time2"]

    @staticmethod
    def backward(ctx, g_loss, *_):
        time, noise, hours, counts, time2, fibre_time, class_mean, class_coef = ctx.saved_tensors
        dev = time.device
        g_time = torch.empty_like(time)
        gl = g_loss.detach().to(torch.float32).reshape(1).contiguous()
        a = _abi.LossArgs()
        a.S, a.T = ctx.S, ctx.T
        for k, v in dict(time=time, noise=noise, hours=hours, counts=counts, time2=time2, fibre_time=fibre_time,
                         class_mean=class_mean, class_coef=class_coef, g_loss=gl, g_time=g_time).items():
            setattr(a, k, v.data_ptr())
        for k, v in ctx.consts.items():
            if k != "sharpness_dev":
                setattr(a, k, float(v))
        a.sharpness_dev = _abi.ptr(ctx.consts.get("sharpness_dev"))
        with torch.cuda.device(dev):
            a.stream = torch.cuda.current_stream(dev).cuda_stream
            _abi.check(_abi.load_library().pfs_loss_bwd(ct.byref(a)), "pfs_loss_bwd")
        return g_time, None, None, None, None, None, None


def loss_from_times(time, class_info, nfibers, nclasses, nfields=10, total_time=42, wutils=2000.0, wvar=1.0, pclass=0.1,
                    pfiber=1.0, sharpness=0.5, noise=None):
    """Loss terms from the edge times [E] (or [E,1]); returns (loss, scalars, n_prime, fibre_time, time2) where
    scalars = [loss, totutils, class_penalty, fibre_penalty, variance, #minima, 0, 0]."""
    time = time.reshape(-1)
    if time.dtype != torch.float32:
        time = time.float()                       # bf16 head output: dtype plumbing, the loss itself is fp32
    if noise is None:
        noise = torch.rand_like(time)             # the draw of softfloor, reference src/train.py:22
    hours = class_info[:, 0].to(torch.float32)
    counts = (class_info[:, 1] / nfields).to(torch.float32)
    consts = dict(total_time=total_time, wutils=wutils, wvar=wvar, pclass=pclass, pfiber=pfiber, noiselevel=NOISELEVEL)
    if torch.is_tensor(sharpness):                # device scalar: the value is read by the kernels (CUDA-graph replays)
        consts.update(sharpness=0.0, sharpness_dev=sharpness.to(torch.float32).reshape(1))
    else:
        consts.update(sharpness=sharpness)
    return LossFunction.apply(time, noise.reshape(-1).to(torch.float32), hours, counts, int(nfibers), int(nclasses), consts)


def loss_function(gnn, graph, class_info, pclass=0.1, pfiber=1.0, sharpness=0.5, finaloutput=False, *, nfibers=None,
                  nclasses=None, nfields=10, total_time=42, wutils=2000.0, wvar=1.0, noise=None):
    """reference src/train.py:29-80 with its globals as arguments (`gnn` first; NFIBERS / NCLASSES default to the
    graph's node counts)."""
    from .topology import get_topology
    nclasses = int(nclasses if nclasses is not None else graph.x_t.shape[0])
    nfibers = int(nfibers if nfibers is not None else graph.x_s.shape[0])
    topo = get_topology(graph.edge_index, nfibers, nclasses)
    if not topo.canonical:
        raise _abi.PfsError("the loss needs the canonical dense edge order e = k*T + i (reference src/train.py:67 "
                            "reshapes the times to [NFIBERS, NCLASSES])")
    time = gnn.edge_prediction(graph.x_e, scale=total_time / nclasses).squeeze(-1)
    loss, sc, n_prime, fibre_time, time2 = loss_from_times(time, class_info, nfibers, nclasses, nfields, total_time, wutils,
                                                            wvar, pclass, pfiber, sharpness, noise)
    if not finaloutput:
        return loss, sc[1]
    counts = class_info[:, 1] / nfields
    comp = (n_prime / counts).detach().cpu().numpy()
    return loss, sc[1], comp, n_prime, fibre_time.detach().cpu().numpy(), time2, sc[4]

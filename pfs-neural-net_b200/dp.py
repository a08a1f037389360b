"""Data parallelism over graph batches (BASELINE config 3).

Every graph is an independent unit with its own BatchNorm statistics, so graphs are sharded over
ranks (one process per GPU) with NO collective on the data path; the only exchange is ONE all-reduce
(sum) of a flat fp32 gradient bucket per step over NCCL / NVLink.  The reference has no distributed
code at all; this module is new work.

  * parameters whose gradient is None on a rank (the reference loss leaves the last Block's node
    models without gradient) are zero-filled so every rank reduces the same
    bucket layout;
  * BatchNorm running buffers are per rank (the reference is single-process); `broadcast_buffers`
    copies rank 0's at checkpoint time.
"""
import torch
import torch.distributed as dist


def shard_graphs(num_graphs, rank, world_size):
    """Graph ids owned by `rank`: g = rank (mod world_size)."""
    return list(range(rank, num_graphs, world_size))


class GradBucket:
    """Flat fp32 bucket over `params` (registration order): pack -> all_reduce(sum) -> the parameters' `.grad`
    become VIEWS of the reduced bucket (no copy back).

    pack is one multi-tensor copy (plus a zero fill of the slices whose gradient is None on this rank, so that
    every rank reduces the same layout); nothing is launched per parameter after the collective.  The whole
    sequence is CUDA-graph capturable (static bucket, no host synchronisation)."""

    def __init__(self, params, process_group=None):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("no trainable parameters")
        dev = self.params[0].device
        self.numel = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(self.numel, dtype=torch.float32, device=dev)
        self.views = []
        off = 0
        for p in self.params:
            self.views.append(self.flat[off:off + p.numel()].view_as(p))
            off += p.numel()
        self.group = process_group

    def pack(self):
        have = [(v, p.grad) for v, p in zip(self.views, self.params)
                if p.grad is not None and p.grad.data_ptr() != v.data_ptr()]
        miss = [v for v, p in zip(self.views, self.params) if p.grad is None]
        if have:
            torch._foreach_copy_([v for v, _ in have], [g for _, g in have])
        if miss:
            torch._foreach_zero_(miss)

    def unpack(self):
        """Point every parameter's gradient at its slice of the bucket (fp32 parameters: a view, no copy)."""
        for v, p in zip(self.views, self.params):
            if p.dtype == torch.float32:
                p.grad = v
            elif p.grad is None:
                p.grad = v.to(p.dtype)
            else:
                p.grad.copy_(v)

    def world(self):
        if not (dist.is_available() and dist.is_initialized()):
            return 1
        return dist.get_world_size(self.group)

    def all_reduce(self, average=False):
        """Sum (or average) the gradients of all ranks; a no-op without an initialised process group or with a
        single rank.  Afterwards `p.grad` is a view of the bucket."""
        if self.world() == 1:
            return
        self.pack()
        dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
        if average:
            self.flat.div_(self.world())
        self.unpack()


def broadcast_buffers(module, src=0, process_group=None):
    """Make rank `src`'s BatchNorm running buffers the checkpointed ones (DDP-style)."""
    if not (dist.is_available() and dist.is_initialized()):
        return
    for b in module.buffers():
        dist.broadcast(b, src=src, group=process_group)


def broadcast_parameters(module, src=0, process_group=None):
    if not (dist.is_available() and dist.is_initialized()):
        return
    for p in module.parameters():
        dist.broadcast(p.data, src=src, group=process_group)

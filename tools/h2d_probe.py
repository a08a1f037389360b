#!/usr/bin/env python
"""Bare pinned host->device bandwidth per GPU with 1..N ranks copying at the same time (torchrun, one rank per GPU):
separates the fabric limit of the box from bench.py's copy pattern.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/h2d_probe.py [--bind]

Prints one JSON line per concurrency level k = 1, 2, 4, ... N: ranks 0..k-1 copy, the others wait."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    bind = "--bind" in sys.argv
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = None
    if bind:
        import bench
        numa = bench.numa_bind(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    nbytes = 512 << 20
    host = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    host.fill_(1)
    dst = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    aff = sorted(os.sched_getaffinity(0))
    k = 1
    while k <= world:
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
        gbs = 0.0
        if rank < k:
            dst.copy_(host, non_blocking=True)
            torch.cuda.synchronize(dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(8):
                dst.copy_(host, non_blocking=True)
            e1.record()
            torch.cuda.synchronize(dev)
            gbs = 8 * nbytes / (e0.elapsed_time(e1) * 1e-3) / 1e9
        t = torch.tensor([gbs], device=dev)
        allv = [torch.zeros(1, device=dev) for _ in range(world)]
        if world > 1:
            dist.all_gather(allv, t)
        else:
            allv = [t]
        if rank == 0:
            v = [float(x) for x in allv][:k]
            print(json.dumps({"concurrent_ranks": k, "bind": bind, "h2d_gbs_per_gpu": [round(x, 1) for x in v],
                              "min": round(min(v), 1), "sum": round(sum(v), 1), "rank0_numa": numa,
                              "rank0_cpus": "%d..%d (%d)" % (aff[0], aff[-1], len(aff))}))
        k *= 2
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

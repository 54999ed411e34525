"""Pinned host-to-device copy bandwidth per rank, alone and with every rank copying at once (run under torchrun):
explains the end-to-end leg of bench.py at N > 1.  python -m torch.distributed.run --nproc-per-node N tools/probe_h2d.py"""
import json
import os

import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
nbytes = 1 << 30
h = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
d = torch.empty(nbytes, dtype=torch.uint8, device=dev)


def copies(n):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        d.copy_(h, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    return n * nbytes / (e0.elapsed_time(e1) * 1e6)


copies(2)
alone = []
for r in range(world):           # one rank at a time
    if world > 1:
        dist.barrier()
    if r == rank:
        alone.append(copies(5))
if world > 1:
    dist.barrier()
together = copies(10)
out = torch.tensor([alone[0], together], device=dev, dtype=torch.float64)
if world > 1:
    g = [torch.zeros_like(out) for _ in range(world)]
    dist.all_gather(g, out)
else:
    g = [out]
if rank == 0:
    print(json.dumps({"ranks": world, "pinned_h2d_GBps_alone": [round(float(x[0]), 2) for x in g],
                      "pinned_h2d_GBps_all_ranks_at_once": [round(float(x[1]), 2) for x in g]}))
if world > 1:
    dist.destroy_process_group()

"""tools/h2d_aggregate.py (under torchrun, one rank per GPU): what the box's host -> device path gives when EVERY rank copies
its 3.24 GB of pinned node planes at the same time, with no kernels running — the floor of the end-to-end step at N GPUs."""
import json
import os

import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 3 * 513 ** 3
h = torch.empty(n, dtype=torch.float64).pin_memory()
d = torch.empty(n, dtype=torch.float64, device="cuda")
d.copy_(h, non_blocking=True)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 3
e0.record()
for _ in range(reps):
    d.copy_(h, non_blocking=True)
e1.record()
torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1) / reps], device="cuda")
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
if rank == 0:
    t = float(ms.item())
    print(json.dumps({"n_gpus": world, "bytes_per_rank": n * 8, "ms_max_over_ranks": t, "GBs_per_rank": n * 8 / t / 1e6,
                      "GBs_aggregate": world * n * 8 / t / 1e6}))
if world > 1:
    dist.destroy_process_group()

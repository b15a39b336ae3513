"""Concurrent host->device bandwidth of N ranks (development probe): does the end-to-end path scale past the host?

  python -m torch.distributed.run --nproc-per-node N tools/probe/h2d_scaling.py [--affinity]
"""
import os
import sys
import time

import torch
import torch.distributed as dist


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    if "--affinity" in sys.argv:
        import pynvml
        pynvml.nvmlInit()
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl")
    n = 256 << 20
    h = torch.empty(n, dtype=torch.uint8).pin_memory()
    h.fill_(1)
    d = torch.empty(n, dtype=torch.uint8, device="cuda")
    for _ in range(3):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    reps = 20
    for _ in range(reps):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    gbs = n * reps / dt / 1e9
    t = torch.tensor([gbs], device="cuda")
    if world > 1:
        lst = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(lst, t)
        if rank == 0:
            v = [round(float(x), 1) for x in lst]
            print("affinity" if "--affinity" in sys.argv else "default", "per-rank GB/s", v, "sum", round(sum(v), 1), "cpus", sorted(os.sched_getaffinity(0))[:4], "...", flush=True)
        dist.destroy_process_group()
    else:
        print("single", round(gbs, 1), flush=True)


if __name__ == "__main__":
    main()

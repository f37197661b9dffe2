"""Host <-> device copy bandwidth with N ranks copying AT THE SAME TIME (what bounds the host-buffer step, bench.py `e2e`,
when several GPUs of one box share the host): pinned buffers of the host-step's sizes (the [B,2] i32 action array and
the [B] x 72-byte record array at 65,536 envs, and a large buffer), each rank on its own GPU.

    python profiles/prof_pcie.py                                              # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 profiles/prof_pcie.py

Rank 0 prints per-rank and aggregate GB/s (barrier before every timed loop so the copies overlap across ranks)."""
import os
import time

import torch
import torch.distributed as dist

rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)


def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def gather(x):
    t = torch.tensor([x], dtype=torch.float64, device=dev)
    if world > 1:
        out = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(out, t)
        return [float(o.item()) for o in out]
    return [x]


for nbytes in (65536 * 8, 65536 * 72, 64 * 1024 * 1024):
    d = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    h = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    for name, fn in (("d2h", lambda: h.copy_(d, non_blocking=True)), ("h2d", lambda: d.copy_(h, non_blocking=True))):
        for _ in range(5):
            fn()
        barrier()
        t0 = time.perf_counter()
        reps = 50
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / reps
        rates = gather(nbytes / dt / 1e9)
        if rank == 0:
            print("%s %9d B x %d ranks: %.1f us/copy on rank 0, per-rank GB/s min %.1f max %.1f, aggregate %.1f GB/s"
                  % (name, nbytes, world, dt * 1e6, min(rates), max(rates), sum(rates)), flush=True)
# launch + synchronize latency with all ranks busy
x = torch.zeros(1, device=dev)
for _ in range(10):
    x.add_(1); torch.cuda.synchronize()
barrier()
t0 = time.perf_counter()
for _ in range(200):
    x.add_(1); torch.cuda.synchronize()
lat = gather((time.perf_counter() - t0) / 200 * 1e6)
if rank == 0:
    print("launch + synchronize: %.1f us (max over ranks %.1f us); host cores available to this process: %d"
          % (lat[0], max(lat), len(os.sched_getaffinity(0))), flush=True)
if world > 1:
    dist.destroy_process_group()

"""Host-to-device copy bandwidth of the box, per GPU: every rank alone, then all ranks at once
(torchrun, one process per GPU). Explains the end-to-end scaling of bench.py: a match needs
208 MB of input every 5.4 ms = 38 GB/s per GPU.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29555 tools/h2d_bandwidth.py
"""
import json
import os

import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
MB = 256
host = torch.empty(MB << 20, dtype=torch.uint8).pin_memory()
dev = torch.empty(MB << 20, dtype=torch.uint8, device="cuda")


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def measure(active, iters=20):
    barrier()
    gbs = 0.0
    if active:
        for _ in range(3):
            dev.copy_(host, non_blocking=True)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(iters):
            dev.copy_(host, non_blocking=True)
        b.record()
        torch.cuda.synchronize()
        gbs = iters * (MB << 20) / (a.elapsed_time(b) * 1e-3) / 1e9
    barrier()
    return gbs


alone = []
for r in range(world):
    g = measure(rank == r)
    t = torch.tensor([g], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    alone.append(float(t.item()))
g = measure(True)
t = torch.tensor([g], dtype=torch.float64, device="cuda")
lo = t.clone()
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
if rank == 0:
    print(json.dumps({"bench": "pinned host-to-device copy, 256 MB", "n_gpus": world,
                      "alone_gb_s_per_gpu": [round(x, 1) for x in alone],
                      "concurrent_gb_s_total": round(float(t.item()), 1),
                      "concurrent_gb_s_min_per_gpu": round(float(lo.item()), 1),
                      "needed_gb_s_per_gpu_for_bench": 38.5}), flush=True)
if world > 1:
    dist.destroy_process_group()

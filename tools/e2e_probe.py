"""Where does the end-to-end path lose host-link bandwidth when many GPUs run at once? (torchrun, one rank per GPU)

Runs bench.py's e2e loop only (bicos_b200_match_host_begin/_end, 2 frames in flight, pinned host stacks) and reports
the aggregate input rate in GB/s. BICOS_B200_HOST_PROBE (temporary switch in cabi.cu) takes pieces out of the
pipeline: "nod2h" = no result downloads, "nocompute,nod2h" = band uploads only.
"""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import libbicos_b200 as lb  # noqa: E402
from libbicos_b200 import synth  # noqa: E402

N, ROWS, COLS, FRAMES = 33, 1536, 2048, 4
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
if os.environ.get("PROBE_PLAIN"):
    # plain copies: 256 MB host-to-device on one stream, optionally 32 MB device-to-host on another at the same time
    MB = 1 << 20
    hsrc = torch.empty(256 * MB, dtype=torch.uint8).pin_memory()
    ddst = torch.empty(256 * MB, dtype=torch.uint8, device="cuda")
    dsrc = torch.empty(32 * MB, dtype=torch.uint8, device="cuda")
    hdst = torch.empty(32 * MB, dtype=torch.uint8).pin_memory()
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    both = os.environ["PROBE_PLAIN"] == "both"
    for rep in range(2):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(s1):
            a.record()
            for _ in range(8):
                ddst.copy_(hsrc, non_blocking=True)
            b.record()
        if both:
            with torch.cuda.stream(s2):
                for _ in range(8):
                    hdst.copy_(dsrc, non_blocking=True)
        torch.cuda.synchronize()
    g = torch.tensor([8 * 256 * MB / (a.elapsed_time(b) * 1e-3) / 1e9], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(g, op=dist.ReduceOp.SUM)
    if rank == 0:
        print(json.dumps({"probe": "plain 256 MB host-to-device copies" + (" + 32 MB device-to-host at the same time" if both else ""),
                          "n_gpus": world, "input_gb_s_total": round(float(g.item()), 1)}), flush=True)
    if world > 1:
        dist.destroy_process_group()
    sys.exit(0)
cfg = lb.Config(nxcorr_threshold=0.96, min_variance=2.0, subpixel_step=0.1, consistency=True, max_lr_diff=1)
host = []
for f in range(FRAMES):
    l, r, _ = synth.make_stacks(N, ROWS, COLS, np.uint8, frame=rank * FRAMES + f, xp=torch, device="cuda")
    host.append((l.cpu().pin_memory().numpy(), r.cpu().pin_memory().numpy()))
outs = [(torch.empty((ROWS, COLS), dtype=torch.float32).pin_memory().numpy(), torch.empty((ROWS, COLS), dtype=torch.float32).pin_memory().numpy())
        for _ in range(FRAMES)]
hh = [lb.Handle(local), lb.Handle(local)]


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def step():
    for f, ((l, r), out) in enumerate(zip(host, outs)):
        hh[f % 2].match_host_end()
        hh[f % 2].match_host_begin(l, r, cfg, out=out)
    for x in hh:
        x.match_host_end()


for _ in range(2):
    step()
barrier()
t0 = time.perf_counter()
steps = 6
for _ in range(steps):
    step()
torch.cuda.synchronize()
ms = (time.perf_counter() - t0) * 1e3 / steps
t = torch.tensor([ms], dtype=torch.float64, device="cuda")
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
ms = float(t.item())
if rank == 0:
    gbs = world * FRAMES * 2 * N * ROWS * COLS / (ms * 1e-3) / 1e9
    print(json.dumps({"probe": os.environ.get("BICOS_B200_HOST_PROBE", "full pipeline"), "n_gpus": world, "ms_per_step": ms,
                      "input_gb_s_total": round(gbs, 1), "mpx_per_s": round(world * FRAMES * ROWS * COLS / ms / 1e3)}), flush=True)
if world > 1:
    dist.destroy_process_group()

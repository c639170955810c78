"""BASELINE configs[4]: a batch of 256 independent 2x33 1920x1200 stereo stacks, frame-sharded
over the GPUs of one node (one process per GPU, torchrun; also runs on 1 GPU). No data-path
communication: every rank matches its own frames; one barrier on each side of the timed region.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29544 tools/bench_batch.py [--frames 256] [--distinct 8]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import libbicos_b200 as lb  # noqa: E402
from libbicos_b200 import sharding, synth  # noqa: E402
from tools.bench_configs import CONFIGS  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=256)
    ap.add_argument("--distinct", type=int, default=8, help="distinct synthetic frames per GPU, cycled")
    ap.add_argument("--config", default="C5")
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n, dt, rows, cols, kw, note = CONFIGS[args.config]
    cfg = lb.Config(**kw)
    h = lb.Handle(local)
    mine = sharding.frame_indices(rank, world, args.frames)
    distinct = min(args.distinct, len(mine))
    frames = [synth.make_stacks(n, rows, cols, dt, frame=mine[f], xp=torch, device="cuda")[:2] for f in range(distinct)]
    outs = [h.match(l, r, cfg) for l, r in frames]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # throughput mode: the rank's frames in batches of `distinct` through bicos_b200_match_batch
    def run(count):
        done = 0
        while done < count:
            k = min(distinct, count - done)
            h.match_batch(frames[:k], cfg, outs=outs[:k])
            done += k

    run(min(distinct, len(mine)))
    barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    run(len(mine))
    b.record()
    barrier()
    t = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    if rank == 0:
        px = rows * cols
        print(json.dumps({"bench": "frame-sharded batch", "config": args.config, "note": note, "n_gpus": world,
                          "frames": args.frames, "frames_per_gpu": len(mine), "distinct_frames_per_gpu": distinct,
                          "batch_ms": ms, "frames_per_s": args.frames / ms * 1e3, "mpx_per_s": args.frames * px / ms / 1e3,
                          "ms_per_frame_per_gpu": ms / len(mine)}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

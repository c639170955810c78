"""Single match, row-sharded over N GPUs (one process per GPU, torchrun): strong scaling.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29533 tools/bench_rowshard.py [--config metric] [--iters 20]

Every rank holds only its row block of both stacks (generated locally: the synthetic scene is
counter-based). Two ways to assemble the result on rank 0 are timed:
  peer    PeerAssembly: refine kernels store their rows into rank 0's images over NVLink
  gather  match into local tensors, then torch.distributed gather (NCCL) of disparity + corrmap
Device time per match = max over ranks of CUDA-event time around [barrier .. own kernels done];
the rank-0 result is checked against a single-GPU match of the whole image.
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
    ap.add_argument("--config", default="metric")
    ap.add_argument("--iters", type=int, default=20)
    args = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    n, dt, rows, cols, kw, note = CONFIGS[args.config]
    cfg = lb.Config(**kw)
    h = lb.Handle(local)
    lo, hi = sharding.row_range(rank, world, rows)
    l, r, _ = synth.make_stacks(n, rows, cols, dt, row0=lo, rows=hi - lo, xp=torch, device="cuda")

    def timed(fn, finish):
        for _ in range(3):
            fn()
            finish()
        ts = []
        for _ in range(args.iters):
            dist.barrier()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            res = finish(record=b)
            t = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ts.append(float(t.item()))
        return float(np.median(ts)), float(np.min(ts)), res

    # ---- peer-memory assembly ------------------------------------------------------------------
    pa = sharding.PeerAssembly(h, rows, cols, cfg, local)

    def peer_finish(record=None):
        if record is not None:
            record.record()
        return pa.finish()

    peer_med, peer_min, (pd, pc) = timed(lambda: pa.match(l, r), peer_finish)

    # ---- NCCL gather ---------------------------------------------------------------------------
    out = h.match(l, r, cfg)

    def gather_fn():
        h.match(l, r, cfg, out=out)

    def gather_finish(record=None):
        d = sharding.gather_rows(out[0], rows)
        c = sharding.gather_rows(out[1], rows) if out[1] is not None else None
        if record is not None:
            record.record()
        torch.cuda.synchronize()
        return d, c

    gat_med, gat_min, (gd, gc) = timed(gather_fn, gather_finish)

    if rank == 0:
        fl, fr, _ = synth.make_stacks(n, rows, cols, dt, xp=torch, device="cuda")
        want = h.match(fl, fr, cfg)
        single_med = None
        ts = []
        for _ in range(5):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            h.match(fl, fr, cfg, out=want)
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        single_med = float(np.median(ts))

        def same(x, y):
            return bool(torch.equal(torch.nan_to_num(x.double(), nan=-7.0), torch.nan_to_num(y.double(), nan=-7.0)))

        ok = same(pd, want[0]) and same(gd, want[0])
        if want[1] is not None:
            ok = ok and same(pc, want[1]) and same(gc, want[1])
        px = rows * cols
        print(json.dumps({
            "bench": "row-sharded single match", "config": args.config, "note": note, "n_gpus": world,
            "rows_per_gpu": hi - lo, "single_gpu_ms": single_med,
            "peer_ms": peer_med, "peer_ms_min": peer_min, "peer_mpx_per_s": px / peer_med / 1e3,
            "peer_speedup_vs_1gpu": single_med / peer_med,
            "gather_ms": gat_med, "gather_ms_min": gat_min, "gather_mpx_per_s": px / gat_med / 1e3,
            "matches_single_gpu_result": ok,
        }), flush=True)
    pa.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

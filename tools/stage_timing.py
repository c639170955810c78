"""Per-stage CUDA-event timings at a given size (development aid; bench.py is the contract)."""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import libbicos_b200 as lb
from libbicos_b200 import synth


def timeit(fn, iters=10, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts)), float(np.min(ts))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=33)
    ap.add_argument("--rows", type=int, default=1536)
    ap.add_argument("--cols", type=int, default=2048)
    ap.add_argument("--u16", action="store_true")
    ap.add_argument("--full", action="store_true")
    args = ap.parse_args()
    dt = np.uint16 if args.u16 else np.uint8
    tdt = torch.uint16 if args.u16 else torch.uint8
    l, r, _ = synth.make_stacks(args.n, args.rows, args.cols, dt, xp=torch, device="cuda")
    h = lb.Handle(0)
    P = args.rows * args.cols
    k = lb.descriptor_words(args.n, args.full)
    eb = 2 if args.u16 else 1
    print(f"n={args.n} {args.cols}x{args.rows} K={k} dtype={tdt}")

    med, mn = timeit(lambda: h.transform(l, args.full))
    tb = args.n * P * eb + P * 4 * k
    print(f"transform (one stack): {med:.3f} ms (min {mn:.3f})  {tb / mn / 1e6:.0f} GB/s")
    d0, _ = h.transform(l, args.full)
    d1, _ = h.transform(r, args.full)
    for flags, name in ((1, "NODUPES"), (2, "CONSISTENCY"), (3, "BOTH"), (0, "PLAIN")):
        med, mn = timeit(lambda: h.search(d0, d1, k, args.cols, flags), iters=5, warmup=2)
        popc = args.cols * P * k
        print(f"search {name:12s}: {med:.3f} ms (min {mn:.3f})  {popc / mn / 1e9:.2f} T popc32/s  {args.cols * P / mn / 1e9:.3f} T pairs/s")
    for kw, name in (
        (dict(nxcorr_threshold=0.96, min_variance=2.0), "agree"),
        (dict(nxcorr_threshold=0.96, min_variance=2.0, subpixel_step=0.1), "agree_subpixel 0.1"),
        (dict(nxcorr_threshold=0.96, min_variance=2.0, subpixel_step=0.1, double=True), "agree_subpixel 0.1 f64"),
    ):
        cfg = lb.Config(**kw)
        keys = h.search(d0, d1, k, args.cols, cfg.flags)
        med, mn = timeit(lambda: h.refine(l, r, cfg, keys, want_raw=False), iters=5, warmup=2)
        print(f"refine {name:24s}: {med:.3f} ms (min {mn:.3f})")
    for kw, name in (
        (dict(nxcorr_threshold=0.96, min_variance=2.0), "config1 (NoDuplicates, integer)"),
        (dict(nxcorr_threshold=0.96, min_variance=2.0, subpixel_step=0.1, consistency=True, max_lr_diff=1),
         "config2 (Consistency, subpixel 0.1)"),
    ):
        cfg = lb.Config(**kw)
        out = h.match(l, r, cfg)
        med, mn = timeit(lambda: h.match(l, r, cfg, out=out), iters=10, warmup=3)
        valid = (~torch.isnan(out[0]) & (out[0] != -32768)).float().mean().item()
        print(f"match {name:38s}: {med:.3f} ms (min {mn:.3f})  {P / med / 1e3:.1f} Mpx/s  valid {valid:.3f}")


if __name__ == "__main__":
    main()

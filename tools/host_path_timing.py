"""Host-resident entry points, wall clock: pinned vs pageable inputs, sync vs two frames in flight,
and pybicos.match (what a Python user of the reference calls). Development aid."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import libbicos_b200 as lb
from libbicos_b200 import pybicos, synth

N, ROWS, COLS = 33, 1536, 2048
cfg = lb.Config(nxcorr_threshold=0.96, min_variance=2.0, subpixel_step=0.1, consistency=True, max_lr_diff=1)
l, r, _ = synth.make_stacks(N, ROWS, COLS, np.uint8, xp=torch, device="cuda")
lp, rp = l.cpu().pin_memory().numpy(), r.cpu().pin_memory().numpy()
lg, rg = np.array(lp), np.array(rp)  # pageable copies
outp = (torch.empty((ROWS, COLS), dtype=torch.float32).pin_memory().numpy(), torch.empty((ROWS, COLS), dtype=torch.float32).pin_memory().numpy())
h = lb.Handle(0)


def wall(fn, iters=8, warmup=2):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        t0 = time.perf_counter()
        fn()
        ts.append((time.perf_counter() - t0) * 1e3)
    return np.median(ts), np.min(ts)


print("match_host, pinned in/out      : %.2f ms (min %.2f)" % wall(lambda: h.match_host(lp, rp, cfg, out=outp)))
print("match_host, pageable in/out    : %.2f ms (min %.2f)" % wall(lambda: h.match_host(lg, rg, cfg)))
pc = pybicos.Config()
pc.nxcorr_threshold = 0.96
pc.min_variance = 2.0
pc.subpixel_step = 0.1
pc.set_consistency(1, False)
ll, rl = list(lg), list(rg)
print("pybicos.match (lists of arrays): %.2f ms (min %.2f)" % wall(lambda: pybicos.match(ll, rl, pc), iters=5))

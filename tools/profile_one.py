"""One BICOS::match of the bench workload (for ncu): 2 transform launches, 1 search, 1 refine."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import libbicos_b200 as lb
from libbicos_b200 import synth

variant = sys.argv[1] if len(sys.argv) > 1 else "config2"
kw = dict(nxcorr_threshold=0.96, min_variance=2.0)
if variant == "config2":
    kw.update(subpixel_step=0.1, consistency=True, max_lr_diff=1)
if variant == "c4":  # 256-bit descriptors (n = 64), a row band of the 4096-column configuration
    kw = dict(nxcorr_threshold=0.96, min_variance=2.0)
    l, r, _ = synth.make_stacks(64, 3000, 4096, np.uint8, rows=256, xp=torch, device="cuda")
else:
    l, r, _ = synth.make_stacks(33, 1536, 2048, np.uint8, xp=torch, device="cuda")
h = lb.Handle(0)
h.set_overlap(False)  # whole-frame kernels one after the other: what a per-kernel profile wants (ncu serialises anyway)
cfg = lb.Config(**kw)
out = h.match(l, r, cfg)
torch.cuda.synchronize()
for _ in range(0 if variant == "c4" else 2):
    h.match(l, r, cfg, out=out)
torch.cuda.synchronize()
print("ok", float(torch.nan_to_num(out[0]).sum()))

"""Every BASELINE.json config on one B200. Prints one JSON line per config: device-resident ms
per match (CUDA events, median of `--iters` after 3 warm-ups), Mpx/s and per-stage times.

`python bench.py --table` runs the same table with the reference's own CUDA backend timed beside
every config (bench.py is the one measurement script that may load the baselines under oracle/):
the UNMODIFIED reference src/impl/cuda.cu compiled for sm_100a (`make -C oracle refcuda`), on the
same synthetic stacks, plus how many disparities differ between the two (the reference's CUDA
build contracts the interpolation into FMAs and leaves unevaluated corrmap cells uninitialised,
so only the integer part is expected to agree everywhere).
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import libbicos_b200 as lb  # noqa: E402
from libbicos_b200 import synth  # noqa: E402

C1 = dict(nxcorr_threshold=0.96, min_variance=2.0)
C2 = dict(nxcorr_threshold=0.96, min_variance=2.0, subpixel_step=0.1, consistency=True, max_lr_diff=1)
CONFIGS = {
    # name: (n, dtype, rows, cols, config, note)
    "C1": (33, np.uint8, 1024, 1280, C1, "configs[0]: LIMITED, NoDuplicates, integer"),
    "C2": (33, np.uint8, 1024, 1280, C2, "configs[1]: + subpixel 0.1, Consistency{1}, float"),
    "metric": (33, np.uint8, 1536, 2048, C2, "BASELINE metric size, configs[1] Config"),
    "C3": (16, np.uint16, 2048, 2448, dict(nxcorr_threshold=0.96, min_variance=2.0, mode_full=True, double=True),
           "configs[2] at n=16 (largest FULL stack the reference accepts: 227 bits), double"),
    "C3n20": (20, np.uint16, 2048, 2448, dict(nxcorr_threshold=0.96, min_variance=2.0, mode_full=True, double=True,
                                              wide_descriptors=True),
              "configs[2] as named: n=20 FULL = 363 bits -> 12-word descriptors (extension: the reference throws above 256 bits)"),
    "C4": (64, np.uint8, 3000, 4096, C1, "configs[3]: n=64 -> 250 bits -> 256-bit descriptors"),
    "C5": (33, np.uint8, 1200, 1920, C1, "configs[4]: one frame of the batch (frames are independent)"),
}


def median_ms(fn, iters, warmup=3):
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


def run_table(configs, iters=10, refcuda=None):
    """`refcuda`: an object with time(l, r, ...) / match(l, r, ...) of the reference CUDA backend, or None."""
    h = lb.Handle(0)
    for name in configs:
        n, dt, rows, cols, kw, note = CONFIGS[name]
        l, r, _ = synth.make_stacks(n, rows, cols, dt, xp=torch, device="cuda")
        cfg = lb.Config(**kw)
        out = h.match(l, r, cfg)
        h.set_profiling(True)
        med, mn = median_ms(lambda: h.match(l, r, cfg, out=out), iters)
        stage_ms, cnt = h.stage_times()
        h.set_profiling(False)
        px = rows * cols
        disp = out[0]
        valid = (~torch.isnan(disp) & (disp != -32768)).float().mean().item()
        line = {
            "config": name, "note": note, "n": n, "dtype": np.dtype(dt).name, "rows": rows, "cols": cols,
            "K": lb.descriptor_words(n, cfg.mode_full, cfg.wide_descriptors), "cfg": kw, "ms_per_match": med, "ms_min": mn,
            "mpx_per_s": px / med / 1e3, "valid_frac": valid,
            "stage_ms": {k: v / max(cnt, 1) for k, v in zip(("transform_x2", "search", "refine"), stage_ms)},
        }
        if refcuda is not None and kw.get("wide_descriptors"):
            line["reference_cuda"] = {"error": "input stacks too large: the reference rejects more than 256 descriptor bits"}
        elif refcuda is not None:
            ln, rn = l.cpu().numpy(), r.cpu().numpy()
            if dt == np.uint16:
                ln, rn = ln.view(np.uint16), rn.view(np.uint16)
            try:
                ref_ms, ref_min = refcuda.time(ln, rn, warmup=3, iters=max(7, iters), **kw)
                rd, rc = refcuda.match(ln, rn, **kw)
                got = disp.cpu().numpy()
                ref_invalid = np.isnan(rd) if rd.dtype.kind == "f" else rd == -32768
                got_invalid = np.isnan(got) | (got == -32768)
                both = ~ref_invalid & ~got_invalid
                line["reference_cuda"] = {
                    "ms_per_match": ref_ms, "ms_min": ref_min, "mpx_per_s": px / ref_ms / 1e3,
                    "speedup": ref_ms / med, "speedup_vs_ref_min": ref_min / mn,
                    "valid_mask_mismatch": int((ref_invalid != got_invalid).sum()),
                    "disparity_mismatch_gt_1e-3": int((np.abs(rd.astype(np.float64)[both] - got[both]) > 1e-3).sum()),
                    "pixels": int(px),
                }
            except Exception as e:  # the baseline must never take the product's numbers down with it
                line["reference_cuda"] = {"error": str(e)}
        print(json.dumps(line), flush=True)
        del l, r, out
        torch.cuda.empty_cache()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="C1,C2,metric,C3,C3n20,C4,C5")
    ap.add_argument("--iters", type=int, default=10)
    args = ap.parse_args()
    run_table(args.configs.split(","), args.iters)


if __name__ == "__main__":
    main()

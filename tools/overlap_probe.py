"""Does the FP32-bound refine of one frame run beside the tensor-bound search of the next?

    python tools/overlap_probe.py [--frames 4] [--rounds 10] [--bands 1]

Stage-level C ABI calls (bicos_b200_transform / _search / _refine) on two CUDA streams: stream A carries
transform + search of unit u, stream B the refine of unit u - 1 (a unit = one frame, or one row band of a frame).
Key buffers are double-buffered. Prints ms per frame for (a) bicos_b200_match back to back, (b) the staged
calls on one stream, (c) the two-stream pipeline, and checks that (c) reproduces (a) bit for bit.
"""
import argparse
import ctypes
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import libbicos_b200 as lb  # noqa: E402
from libbicos_b200 import capi, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=4)
    ap.add_argument("--rounds", type=int, default=10)
    ap.add_argument("--bands", type=int, default=1)
    ap.add_argument("--rows", type=int, default=1536)
    ap.add_argument("--cols", type=int, default=2048)
    ap.add_argument("--n", type=int, default=33)
    ap.add_argument("--priority", type=int, default=1, help="1: search stream gets the higher priority")
    args = ap.parse_args()
    n, rows, cols = args.n, args.rows, args.cols
    cfg = lb.Config(nxcorr_threshold=0.96, min_variance=2.0, subpixel_step=0.1, consistency=True, max_lr_diff=1)
    ccfg = cfg.to_c()
    L = lb.lib()
    h = lb.Handle(0)
    frames = [synth.make_stacks(n, rows, cols, np.uint8, frame=f, xp=torch, device="cuda")[:2] for f in range(args.frames)]
    K = lb.descriptor_words(n)
    pitch_words = (cols * K + 3) // 4 * 4
    flags = cfg.flags | capi.FLAG_TOP_BIT_FREE

    ref = [h.match(l, r, cfg) for l, r in frames]
    torch.cuda.synchronize()

    def timeit(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(args.rounds):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / (args.rounds * args.frames)

    outs = [(torch.empty_like(d), torch.empty_like(c)) for d, c in ref]

    def serial():
        for (l, r), o in zip(frames, outs):
            h.match(l, r, cfg, out=o)

    t_match = timeit(serial)

    # ---- staged: units = (frame, band) ----
    cuts = [rows * b // args.bands for b in range(args.bands + 1)]
    units = [(f, cuts[b], cuts[b + 1]) for f in range(args.frames) for b in range(args.bands)]
    band_max = max(c1 - c0 for _, c0, c1 in units)
    desc = [torch.empty((band_max, pitch_words), dtype=torch.int32, device="cuda") for _ in range(2)]
    keys = [[torch.empty((band_max, cols), dtype=torch.int32, device="cuda") for _ in range(4)] for _ in range(2)]
    sa = torch.cuda.Stream(priority=-1 if args.priority else 0)
    sb = torch.cuda.Stream(priority=0)
    ev_search = [torch.cuda.Event() for _ in range(2)]
    ev_refine = [torch.cuda.Event() for _ in range(2)]

    def planes(stack, r0):
        eb = stack.element_size()
        return capi._ptr_array([stack.data_ptr() + (t * stack.stride(0) + r0 * stack.stride(1)) * eb for t in range(n)])

    def enqueue_search(u, slot, stream):
        f, r0, r1 = units[u]
        l, r = frames[f]
        sp = ctypes.c_void_p(stream.cuda_stream)
        pitch = l.stride(1)
        for stack, d in ((l, desc[0]), (r, desc[1])):
            capi._check(L.bicos_b200_transform(h._h, planes(stack, r0), n, r1 - r0, cols, pitch, 0, 0, d.data_ptr(), pitch_words, sp))
        k = keys[slot]
        capi._check(L.bicos_b200_search(h._h, desc[0].data_ptr(), desc[1].data_ptr(), K, r1 - r0, cols, pitch_words, flags,
                                        k[0].data_ptr(), None, k[2].data_ptr(), None, sp))

    def enqueue_refine(u, slot, stream):
        f, r0, r1 = units[u]
        l, r = frames[f]
        d, c = outs[f]
        k = keys[slot]
        capi._check(L.bicos_b200_refine(h._h, planes(l, r0), planes(r, r0), n, r1 - r0, cols, l.stride(1), 0, ctypes.byref(ccfg),
                                        k[0].data_ptr(), None, k[2].data_ptr(), None, None,
                                        d.data_ptr() + r0 * d.stride(0) * 4, d.stride(0) * 4,
                                        c.data_ptr() + r0 * c.stride(0) * 4, c.stride(0) * 4, ctypes.c_void_p(stream.cuda_stream)))

    def staged_serial():
        st = torch.cuda.current_stream()
        for u in range(len(units)):
            enqueue_search(u, 0, st)
            enqueue_refine(u, 0, st)

    trace = {}

    def pipelined(record=False):
        cur = torch.cuda.current_stream()
        start = torch.cuda.Event()
        start.record(cur)
        sa.wait_event(start)
        sb.wait_event(start)
        for u in range(len(units)):
            slot = u & 1
            if u >= 2:
                sa.wait_event(ev_refine[slot])  # the refine that read this key buffer
            if record:
                trace[u] = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
                trace[u][0].record(sa)
            enqueue_search(u, slot, sa)
            if record:
                trace[u][1].record(sa)
            ev_search[slot].record(sa)
            sb.wait_event(ev_search[slot])
            if record:
                trace[u][2].record(sb)
            enqueue_refine(u, slot, sb)
            if record:
                trace[u][3].record(sb)
            ev_refine[slot].record(sb)
        done = torch.cuda.Event()
        done.record(sb)
        cur.wait_event(done)
        if record:
            trace["start"] = torch.cuda.Event(enable_timing=True)
        return start

    t_staged = timeit(staged_serial)
    for d, c in outs:
        d.zero_()
        c.zero_()
    t_pipe = timeit(pipelined)
    torch.cuda.synchronize()
    # timeline of one pass: [transform+search begin, end] on stream A, [refine begin, end] on stream B, ms from the first event
    t0 = torch.cuda.Event(enable_timing=True)
    t0.record()
    pipelined(record=True)
    torch.cuda.synchronize()
    timeline = [[round(t0.elapsed_time(e), 3) for e in trace[u]] for u in range(len(units))]
    same = all(torch.equal(torch.nan_to_num(a, nan=-7.0), torch.nan_to_num(b, nan=-7.0))
               for (a, c1), (b, c2) in zip(outs, ref)) and all(
        torch.equal(torch.nan_to_num(c1, nan=-7.0), torch.nan_to_num(c2, nan=-7.0)) for (a, c1), (b, c2) in zip(outs, ref))
    px = rows * cols
    print(json.dumps({"probe": "search/refine overlap", "frames": args.frames, "bands": args.bands, "priority": args.priority,
                      "kernel": lb.last_search_kernel(),
                      "ms_per_frame": {"match_serial": t_match, "staged_serial": t_staged, "two_streams": t_pipe},
                      "mpx_per_s": {"match_serial": px / t_match / 1e3, "two_streams": px / t_pipe / 1e3},
                      "pipelined_equals_serial": bool(same), "carveout": os.environ.get("BICOS_B200_REFINE_CARVEOUT", "default"),
                      "timeline_ms[search0,search1,refine0,refine1]": timeline[: 8]}))


if __name__ == "__main__":
    main()

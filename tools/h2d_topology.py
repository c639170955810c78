"""Which GPUs of the box share a host-to-device path? One process, pinned host buffers, concurrent copies.

    python tools/h2d_topology.py [--mb 256] [--iters 10]

Prints one JSON object: the NVML common-ancestor matrix and PCI bus ids, every GPU's bandwidth alone, the aggregate
of every pair (0, k), and of the subsets a 2- and 4-rank job could take (first ordinals, evenly spread, topology-aware
pick of libbicos_b200.topology.pick_devices), with plain pinned and with write-combined pinned host memory.
bench.py's end-to-end number at N < visible GPUs depends on which subset the ranks use.
"""
import argparse
import ctypes
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from libbicos_b200 import topology  # noqa: E402

cudart = ctypes.CDLL("libcudart.so.12") if os.path.exists("/usr/local/cuda/lib64/libcudart.so.12") else None
if cudart is None:
    try:
        cudart = ctypes.CDLL("libcudart.so")
    except OSError:
        cudart = None


def host_alloc(nbytes, write_combined):
    """Pinned portable host memory as a torch uint8 tensor (cudaHostAlloc through ctypes; torch has no WC flag)."""
    if cudart is None:
        return torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    ptr = ctypes.c_void_p()
    flags = 0x01 | (0x04 if write_combined else 0)  # cudaHostAllocPortable | cudaHostAllocWriteCombined
    rc = cudart.cudaHostAlloc(ctypes.byref(ptr), ctypes.c_size_t(nbytes), ctypes.c_uint(flags))
    if rc != 0:
        raise RuntimeError(f"cudaHostAlloc failed: {rc}")
    buf = (ctypes.c_ubyte * nbytes).from_address(ptr.value)
    return torch.frombuffer(buf, dtype=torch.uint8)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mb", type=int, default=256)
    ap.add_argument("--iters", type=int, default=10)
    args = ap.parse_args()
    n = torch.cuda.device_count()
    nbytes = args.mb << 20
    dev = []
    for d in range(n):
        torch.cuda.set_device(d)
        dev.append((torch.empty(nbytes, dtype=torch.uint8, device=f"cuda:{d}"), torch.cuda.Stream(device=d)))
    hosts = {False: [host_alloc(nbytes, False) for _ in range(n)], True: [host_alloc(nbytes, True) for _ in range(n)]}

    def run(subset, wc=False):
        """aggregate GB/s of concurrent copies to the GPUs of `subset`, and the slowest GPU's own rate"""
        evs = {}
        for rep in range(2):  # first repetition warms up
            for d in subset:
                torch.cuda.set_device(d)
                buf, st = dev[d]
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                with torch.cuda.stream(st):
                    a.record()
                    for _ in range(args.iters if rep else 2):
                        buf.copy_(hosts[wc][d], non_blocking=True)
                    b.record()
                evs[d] = (a, b)
            for d in subset:
                torch.cuda.synchronize(d)
        rates = [args.iters * nbytes / (evs[d][0].elapsed_time(evs[d][1]) * 1e-3) / 1e9 for d in subset]
        return round(sum(rates), 1), round(min(rates), 1)

    out = {"bench": f"pinned host-to-device copies, {args.mb} MB, one process", "visible_gpus": n,
           "nvml": topology.describe(), "alone_gb_s": [run([d])[0] for d in range(n)]}
    out["pairs_with_gpu0_gb_s_total"] = {str(k): run([0, k])[0] for k in range(1, n)}
    subsets = {}
    for world in (2, 4, 8):
        if world > n:
            continue
        cand = {"first": list(range(world)), "spread": [i * n // world for i in range(world)],
                "picked": topology.pick_devices(world)[0]}
        for name, sub in cand.items():
            tot, lo = run(sub)
            tot_wc, lo_wc = run(sub, wc=True)
            subsets[f"n{world}_{name}"] = {"gpus": sub, "total": tot, "min_per_gpu": lo, "total_write_combined": tot_wc,
                                          "min_per_gpu_write_combined": lo_wc}
    out["subsets_gb_s"] = subsets
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()

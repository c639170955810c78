"""Host-side multi-GPU logic on CPU: world_size-2 gloo processes, the oracle standing in for the
CUDA matcher (the oracle is allowed here: this is a test)."""

import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

from libbicos_b200 import sharding, synth  # noqa: E402

KW = dict(nxcorr_threshold=0.9, subpixel_step=0.25, min_variance=1.0, consistency=True, max_lr_diff=1)
N, ROWS, COLS = 9, 37, 96  # odd row count: uneven blocks


def test_row_ranges_cover_exactly():
    for rows in (1, 7, 37, 1536):
        for world in (1, 2, 3, 8):
            blocks = [sharding.row_range(r, world, rows) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == rows
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1
    assert sum(len(sharding.frame_indices(r, 8, 256)) for r in range(8)) == 256


def _worker(rank, world, port, tmp):
    sys.path.insert(0, ROOT)
    import oracle

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        def match_fn(s0, s1):
            d, c = oracle.port.match(s0.numpy(), s1.numpy(), **KW)
            return torch.from_numpy(d), torch.from_numpy(c)

        # row-sharded single match: every rank generates only its own rows
        lo, hi = sharding.row_range(rank, world, ROWS)
        l, r, _ = synth.make_stacks(N, ROWS, COLS, np.uint8, seed=21, row0=lo, rows=hi - lo)
        disp, corr = sharding.match_row_sharded(match_fn, torch.from_numpy(l), torch.from_numpy(r), ROWS)
        if rank == 0:
            np.save(os.path.join(tmp, "disp.npy"), disp.numpy())
            np.save(os.path.join(tmp, "corr.npy"), corr.numpy())
        else:
            assert disp is None and corr is None

        # frame-sharded batch: no communication, every frame done exactly once
        def load(f):
            a, b, _ = synth.make_stacks(N, 12, COLS, np.uint8, seed=21, frame=f)
            return torch.from_numpy(a), torch.from_numpy(b)

        done = sharding.match_frames_sharded(match_fn, load, 5)
        np.save(os.path.join(tmp, f"frames_{rank}.npy"), np.array([f for f, _, _ in done]))
        for f, d, _ in done:
            np.save(os.path.join(tmp, f"frame_{f}.npy"), d.numpy())
    finally:
        dist.destroy_process_group()


def test_row_and_frame_sharding_world2(tmp_path, oracles):
    world = 2
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    l, r, _ = synth.make_stacks(N, ROWS, COLS, np.uint8, seed=21)
    want_d, want_c = oracles.port.match(l, r, **KW)
    assert np.array_equal(np.load(tmp_path / "disp.npy"), want_d, equal_nan=True)
    assert np.array_equal(np.load(tmp_path / "corr.npy"), want_c, equal_nan=True)
    frames = sorted(int(f) for k in range(world) for f in np.load(tmp_path / f"frames_{k}.npy"))
    assert frames == [0, 1, 2, 3, 4]
    for f in frames:
        a, b, _ = synth.make_stacks(N, 12, COLS, np.uint8, seed=21, frame=f)
        assert np.array_equal(np.load(tmp_path / f"frame_{f}.npy"), oracles.port.match(a, b, **KW)[0], equal_nan=True)

"""Row- and frame-sharding of BICOS::match across the GPUs of one node (one process per GPU).

Every stage of the path reads only its own image row (reference descriptor_transform.hpp:131-134,
bicos.hpp:91-94, agree.hpp:83,154-156), so
  * a single match is row-sharded: rank g owns rows [g*H/G, (g+1)*H/G) of all 2n input images,
    runs the whole path on them, and the only exchange is the gather of the output rows
    (disparity + corrmap) to rank 0 -- the one real collective on this path;
  * a batch of stereo stacks is frame-sharded with no data-path communication at all.

torch.distributed is plumbing (NCCL on GPUs, gloo in the CPU tests); the matching itself is the
`match_fn` passed in, by default Handle.match from libbicos_b200.capi. On GPUs the gather can be
replaced by PeerAssembly: the kernels store their rows straight into rank 0's images over NVLink.
"""

from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple


def row_range(rank: int, world: int, rows: int) -> Tuple[int, int]:
    """Contiguous row block of `rank`: blocks differ by at most one row and cover [0, rows)."""
    if not 0 <= rank < world:
        raise ValueError("rank out of range")
    return rows * rank // world, rows * (rank + 1) // world


def frame_indices(rank: int, world: int, frames: int) -> List[int]:
    """Frames of a batch handled by `rank` (contiguous blocks, like rows)."""
    lo, hi = row_range(rank, world, frames)
    return list(range(lo, hi))


def gather_rows(local, rows: int, dst: int = 0, group=None):
    """Gather per-rank row blocks [my_rows, W] (split by row_range) into [rows, W] on `dst`.

    Blocks are padded to the largest block so that one fixed-size gather suffices. Returns the
    assembled tensor on `dst` and None elsewhere.
    """
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    lo, hi = row_range(rank, world, rows)
    if local.shape[0] != hi - lo:
        raise ValueError(f"rank {rank} holds {local.shape[0]} rows, expected {hi - lo}")
    block = max(row_range(r, world, rows)[1] - row_range(r, world, rows)[0] for r in range(world))
    padded = local
    if local.shape[0] != block:
        padded = torch.zeros((block,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        padded[: local.shape[0]] = local
    padded = padded.contiguous()
    if rank == dst:
        parts = [torch.empty_like(padded) for _ in range(world)]
        dist.gather(padded, parts, dst=dst, group=group)
        out = torch.empty((rows,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        for r, part in enumerate(parts):
            a, b = row_range(r, world, rows)
            out[a:b] = part[: b - a]
        return out
    dist.gather(padded, None, dst=dst, group=group)
    return None


def match_row_sharded(match_fn: Callable, stack0_rows, stack1_rows, rows: int, dst: int = 0, group=None):
    """One match, row-sharded. Each rank passes ONLY its row block of both stacks ([n, my_rows, W]).

    `match_fn(stack0_rows, stack1_rows) -> (disparity, corrmap or None)` runs the path on the
    local rows. Returns (disparity, corrmap) assembled on `dst`, (None, None) on the other ranks.
    """
    disp, corr = match_fn(stack0_rows, stack1_rows)
    full_disp = gather_rows(disp, rows, dst, group)
    full_corr = gather_rows(corr, rows, dst, group) if corr is not None else None
    return full_disp, full_corr


class PeerAssembly:
    """Output assembly over NVLink peer memory for row-sharded matches (one process per GPU).

    Rank `dst` owns the full-size disparity / corrmap images; every other rank maps them once
    (CUDA IPC, C ABI bicos_b200_shared_*) and its refine kernel stores its rows straight into
    them, so a match needs no gather afterwards: only a stream synchronisation and a barrier.
    Build it once per image size / output types and reuse it for every match.
    """

    def __init__(self, handle, rows: int, cols: int, cfg, device: int, dst: int = 0, group=None):
        import torch
        import torch.distributed as dist

        from .capi import TYPE_16S, TYPE_64F, SharedImage, lib
        import ctypes

        self.handle, self.rows, self.cols, self.cfg, self.dst, self.group = handle, rows, cols, cfg, dst, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        ccfg = cfg.to_c()
        dt = lib().bicos_b200_disparity_type(ctypes.byref(ccfg))
        ct = lib().bicos_b200_corrmap_type(ctypes.byref(ccfg))
        disp_dtype = torch.int16 if dt == TYPE_16S else torch.float32
        corr_dtype = None if not ct else (torch.float64 if ct == TYPE_64F else torch.float32)
        handles = [None, None]
        if self.rank == dst:
            self.disp = SharedImage.create(device, rows, cols, disp_dtype)
            self.corr = SharedImage.create(device, rows, cols, corr_dtype) if corr_dtype is not None else None
            handles = [self.disp.handle, self.corr.handle if self.corr else None]
        dist.broadcast_object_list(handles, src=dst, group=group)
        if self.rank != dst:
            self.disp = SharedImage.open(device, handles[0], rows, cols, disp_dtype)
            self.corr = SharedImage.open(device, handles[1], rows, cols, corr_dtype) if handles[1] else None
        self.lo, self.hi = row_range(self.rank, self.world, rows)

    def match(self, stack0_rows, stack1_rows):
        """Enqueue this rank's rows ([n, my_rows, W]) of one match; results land on rank `dst`."""
        if stack0_rows.shape[1] != self.hi - self.lo:
            raise ValueError(f"rank {self.rank} holds {stack0_rows.shape[1]} rows, expected {self.hi - self.lo}")
        self.handle.match_raw(stack0_rows, stack1_rows, self.cfg, self.disp.row_ptr(self.lo), self.disp.pitch,
                              self.corr.row_ptr(self.lo) if self.corr else None, self.corr.pitch if self.corr else 0)

    def finish(self):
        """Wait for every rank's rows. Returns (disparity, corrmap) tensors aliasing the assembled
        images on rank `dst`, (None, None) elsewhere."""
        import torch
        import torch.distributed as dist

        torch.cuda.synchronize()
        dist.barrier(group=self.group)
        if self.rank != self.dst:
            return None, None
        return self.disp.tensor(), (self.corr.tensor() if self.corr else None)

    def close(self):
        import torch.distributed as dist

        dist.barrier(group=self.group)  # nobody may still be storing into the owner's memory
        if self.rank != self.dst:
            self.disp.close()
            if self.corr:
                self.corr.close()
        dist.barrier(group=self.group)  # mappings are gone before the owner frees
        if self.rank == self.dst:
            self.disp.close()
            if self.corr:
                self.corr.close()


def match_frames_sharded(match_fn: Callable, load_frame: Callable[[int], Sequence], frames: int,
                         rank: Optional[int] = None, world: Optional[int] = None):
    """A batch of `frames` independent stereo stacks, frame-sharded: no communication.

    `load_frame(f) -> (stack0, stack1)` produces frame f on this rank's device. Returns the list
    of (frame index, disparity, corrmap) this rank produced.
    """
    if rank is None or world is None:
        import torch.distributed as dist

        rank, world = dist.get_rank(), dist.get_world_size()
    results = []
    for f in frame_indices(rank, world, frames):
        s0, s1 = load_frame(f)
        disp, corr = match_fn(s0, s1)
        results.append((f, disp, corr))
    return results

"""Which GPUs should the ranks of a job use when it has fewer ranks than the node has GPUs?

Host-resident callers (bicos_b200_match_host, pybicos) are bound by the host-to-device link, and GPUs that hang off
the same PCIe switch or host bridge share it: round 1 measured 120 GB/s for four ranks on the first four ordinals of an
8 x B200 node against 178 GB/s for eight. `pick_devices` spreads the ranks over the node's PCIe tree as NVML reports it
(nvmlDeviceGetTopologyCommonAncestor), and evenly over the ordinals where NVML shows a flat tree (virtual machines).
Nothing is hard-coded to one machine; every rank computes the same list.
"""

from __future__ import annotations

from typing import List, Tuple


def _nvml_matrix():
    """(levels, bus_ids) in CUDA ordinal order, or None: levels[i][j] = NVML common-ancestor level of GPUs i and j
    (0 same board, 10 one PCIe switch, 20 several switches, 30 one host bridge, 40 one NUMA node, 50 across the system)."""
    import pynvml
    import torch

    pynvml.nvmlInit()
    n = torch.cuda.device_count()
    handles = []
    for i in range(n):
        uuid = str(torch.cuda.get_device_properties(i).uuid)
        uuid = uuid if uuid.startswith("GPU-") else "GPU-" + uuid
        try:
            handles.append(pynvml.nvmlDeviceGetHandleByUUID(uuid.encode()))
        except TypeError:
            handles.append(pynvml.nvmlDeviceGetHandleByUUID(uuid))
    levels = [[0] * n for _ in range(n)]
    for i in range(n):
        for j in range(n):
            if i != j:
                levels[i][j] = int(pynvml.nvmlDeviceGetTopologyCommonAncestor(handles[i], handles[j]))
    bus = []
    for h in handles:
        b = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus.append(b.decode() if isinstance(b, bytes) else str(b))
    return levels, bus


def describe() -> dict:
    try:
        levels, bus = _nvml_matrix()
        return {"common_ancestor_level": levels, "pci_bus_id": bus}
    except Exception as e:  # noqa: BLE001 - a description, never a requirement
        return {"unavailable": f"{type(e).__name__}: {e}"}


def spread(world: int, visible: int) -> List[int]:
    return [i * visible // world for i in range(world)]


def choose(world: int, levels) -> Tuple[List[int], str]:
    """Greedy farthest-point selection on the common-ancestor levels: start at GPU 0, then always the GPU whose
    closest already chosen GPU is farthest (ties: the larger sum of levels, then the lower ordinal). Pure function
    of the matrix, so that every rank agrees."""
    visible = len(levels)
    flat = len({levels[i][j] for i in range(visible) for j in range(visible) if i != j}) <= 1
    if flat:
        return spread(world, visible), "evenly spread over the ordinals (NVML reports a flat PCIe tree)"
    chosen = [0]
    while len(chosen) < world:
        best = max((d for d in range(visible) if d not in chosen),
                   key=lambda d: (min(levels[d][c] for c in chosen), sum(levels[d][c] for c in chosen), -d))
        chosen.append(best)
    return sorted(chosen), "spread over the PCIe tree (NVML common-ancestor levels)"


def pick_devices(world: int) -> Tuple[List[int], str]:
    """CUDA ordinals for ranks 0 .. world-1 of a single-node job, and how they were chosen."""
    import torch

    visible = torch.cuda.device_count()
    if world >= visible:
        return list(range(world)), "one rank per visible GPU"
    try:
        levels, _ = _nvml_matrix()
        return choose(world, levels)
    except Exception as e:  # noqa: BLE001
        return spread(world, visible), f"evenly spread over the ordinals (NVML unavailable: {type(e).__name__})"

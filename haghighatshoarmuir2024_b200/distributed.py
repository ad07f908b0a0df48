"""Multi-GPU plumbing: clips are independent, so the batch is sharded by clip across
ranks (one process per GPU) and the only exchange is a final sum of DoA histograms
(the Monte-Carlo loop of paper_plots/target_snn_localization.py:447-467).

Works with any torch.distributed backend: NCCL over NVLink for CUDA tensors on the
GPU box, gloo for the CPU tests.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous clip range [lo, hi) of `rank`; sizes differ by at most one."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def doa_histogram(doa: torch.Tensor, num_doa: int, group_id: Optional[torch.Tensor] = None,
                  num_groups: int = 1) -> torch.Tensor:
    """int64 [num_groups, num_doa] counts of DoA indices (group = e.g. SNR x band cell)."""
    doa = doa.to(torch.int64).reshape(-1)
    if group_id is None:
        flat = doa
    else:
        flat = group_id.to(torch.int64).reshape(-1) * num_doa + doa
    hist = torch.bincount(flat, minlength=num_groups * num_doa)
    return hist.reshape(num_groups, num_doa)


def reduce_histograms(hist: torch.Tensor, group=None) -> torch.Tensor:
    """Sum per-rank histograms in place over all ranks (one small all_reduce)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(hist, op=dist.ReduceOp.SUM, group=group)
    return hist


def max_over_ranks(value: float, device: torch.device, group=None) -> float:
    """Device-timed milliseconds -> max over ranks (the bench's timing rule)."""
    t = torch.tensor([value], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())


def _parse_cpulist(text: str):
    cpus = []
    for part in text.strip().split(","):
        if not part:
            continue
        if "-" in part:
            lo, hi = part.split("-")
            cpus.extend(range(int(lo), int(hi) + 1))
        else:
            cpus.append(int(part))
    return cpus


def bind_to_gpu_numa(local_rank: int, world_local: int = 1) -> dict:
    """Pin this process (and therefore the pinned host buffers it allocates afterwards: first touch) to the CPUs of
    the NUMA node its GPU hangs off, taking the `local_rank`-th of `world_local` equal shares of that node's CPUs so
    that the ranks of one box do not sit on the same cores.  The host-buffer path (micloc_snn_run_host) streams
    tens of GB/s per GPU over PCIe: staging memory on the far socket halves that.  Returns what was done;
    never raises (a box without sysfs topology just stays unbound)."""
    info = {"bound": False}
    try:
        import os
        bus = torch.cuda.get_device_properties(local_rank).pci_bus_id if hasattr(
            torch.cuda.get_device_properties(local_rank), "pci_bus_id") else None
        dom = getattr(torch.cuda.get_device_properties(local_rank), "pci_domain_id", 0)
        dev = getattr(torch.cuda.get_device_properties(local_rank), "pci_device_id", 0)
        if bus is None:
            return info
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev:02x}.0/numa_node"
        node = int(open(path).read().strip())
        if node < 0:
            node = 0
        cpus = _parse_cpulist(open(f"/sys/devices/system/node/node{node}/cpulist").read())
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if not allowed:
            return info
        share = max(len(allowed) // max(world_local, 1), 1)
        mine = allowed[(local_rank % max(world_local, 1)) * share:][:share] or allowed
        os.sched_setaffinity(0, mine)
        info = {"bound": True, "numa_node": node, "cpus": len(mine), "first_cpu": mine[0]}
    except Exception as e:                       # noqa: BLE001 -- topology files are optional
        info["why"] = repr(e)[:80]
    return info

"""Multi-GPU sharding of the hot path: independent CPIs round-robin over ranks.

The path has no exchange step between CPIs (BASELINE config 4: "stream of independent CPIs sharded
round-robin"), so the data path needs no collective; the only communication is the final gather of
the finished maps to rank 0 (NCCL over NVLink on GPUs, gloo in the CPU tests).  One process per GPU.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def cpi_indices(n_cpis: int, rank: int, world: int) -> list[int]:
    """CPI c is processed by rank c mod world."""
    return list(range(rank, n_cpis, world))


def gather_maps(local_maps: torch.Tensor, n_cpis: int, rank: int, world: int, dst: int = 0):
    """local_maps: [k_r, nDop, nDel] complex64, the maps of cpi_indices(n_cpis, rank, world) in order.
    Returns on ``dst`` a tensor [n_cpis, nDop, nDel] in CPI order, None elsewhere."""
    if world == 1:
        return local_maps
    k_max = (n_cpis + world - 1) // world
    shape = tuple(local_maps.shape[1:])
    buf = torch.zeros((k_max,) + shape, dtype=local_maps.dtype, device=local_maps.device)
    buf[: local_maps.shape[0]] = local_maps
    flat = torch.view_as_real(buf).contiguous()   # gloo has no complex support; NCCL does not care
    gathered = [torch.empty_like(flat) for _ in range(world)] if rank == dst else None
    dist.gather(flat, gathered, dst=dst)
    if rank != dst:
        return None
    out = torch.empty((n_cpis,) + shape, dtype=local_maps.dtype, device=local_maps.device)
    for r in range(world):
        idx = cpi_indices(n_cpis, r, world)
        out[idx] = torch.view_as_complex(gathered[r])[: len(idx)]
    return out

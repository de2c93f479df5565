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


# ------------------------------------------------------------------------------------------
# ONE large CPI split over the ranks (BASELINE config 5)
# ------------------------------------------------------------------------------------------
def block_range(n: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous block [start, start+count) of n items for `rank` (sizes differ by at most one)."""
    base, rem = divmod(n, world)
    start = rank * base + min(rank, rem)
    return start, base + (1 if rank < rem else 0)


def _all_gather_rows(local: torch.Tensor, counts: list[int], world: int) -> torch.Tensor:
    """Concatenate row blocks of different heights from every rank (padded all_gather)."""
    k_max = max(counts)
    shape = tuple(local.shape[1:])
    buf = torch.zeros((k_max,) + shape, dtype=local.dtype, device=local.device)
    buf[: local.shape[0]] = local
    flat = torch.view_as_real(buf).contiguous()
    parts = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(parts, flat)
    return torch.cat([torch.view_as_complex(parts[r])[: counts[r]] for r in range(world)], dim=0)


def caf_single_cpi_sharded(amb, d_x_local: torch.Tensor, d_y_local: torch.Tensor, rank: int, world: int,
                           stream=None, dst: int = 0):
    """Cross-ambiguity function of ONE CPI whose IQ is spread over the ranks.

    The range stage is independent per batch, so rank r correlates the contiguous block of batches
    ``block_range(nDopplerBins, r, world)`` -- a contiguous 1/world slice of the IQ, which is all it
    holds (``d_x_local``, ``d_y_local``: complex64 CUDA tensors of that many batches * nCorr samples).
    The Doppler stage needs every batch of a delay column, hence ONE exchange: an all-gather of the
    range matrix (nDop x nDel complex64, <= 17 MB).  Each rank then transforms its tile of delay columns
    (a column's chirp-z transform is indivisible; tiling the OUTPUT Doppler bins instead would repeat the
    whole transform on every rank) and the tiles are gathered to ``dst`` and assembled into the Map.
    Returns the [nDop, nDel] complex64 map on ``dst``, None elsewhere.
    ``amb`` needs ``geometry``, ``range_device`` and ``doppler_device`` (blah2_b200.process.Ambiguity).
    """
    g = amb.geometry
    n_dop, n_del, n_corr = g.n_doppler_bins, g.n_delay_bins, g.n_corr
    b0, nb = block_range(n_dop, rank, world)
    assert d_x_local.numel() == nb * n_corr and d_y_local.numel() == nb * n_corr, "local IQ must hold exactly the rank's batches"
    dev = d_x_local.device
    # The torch operations around the two kernels (copies, collectives, transposes) run on torch's CURRENT stream;
    # the kernels must be ordered with them, so without an explicit stream they go onto that one -- never onto the
    # handle's private non-blocking stream.  An explicit stream must be current while this runs.
    if stream is None and dev.type == "cuda":
        stream = torch.cuda.current_stream(dev)
    s_ptr = stream.cuda_stream if stream is not None and hasattr(stream, "cuda_stream") else stream
    R_local = torch.empty((nb, n_del), dtype=torch.complex64, device=dev)
    amb.range_device(d_x_local, d_y_local, b0, nb, R_local, s_ptr)
    if world > 1:
        counts = [block_range(n_dop, r, world)[1] for r in range(world)]
        R_full = _all_gather_rows(R_local, counts, world)
    else:
        R_full = R_local
    c0, nc = block_range(n_del, rank, world)
    tile = torch.empty((n_dop, nc), dtype=torch.complex64, device=dev)
    amb.doppler_device(R_full.contiguous(), c0, nc, tile, s_ptr)
    if world == 1:
        return tile
    ccounts = [block_range(n_del, r, world)[1] for r in range(world)]
    k_max = max(ccounts)
    buf = torch.zeros((k_max, n_dop), dtype=torch.complex64, device=dev)
    buf[:nc] = tile.transpose(0, 1)
    flat = torch.view_as_real(buf).contiguous()
    gathered = [torch.empty_like(flat) for _ in range(world)] if rank == dst else None
    dist.gather(flat, gathered, dst=dst)
    if rank != dst:
        return None
    cols = torch.cat([torch.view_as_complex(gathered[r])[: ccounts[r]] for r in range(world)], dim=0)  # [nDel, nDop]
    return cols.transpose(0, 1).contiguous()

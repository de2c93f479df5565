"""Multi-GPU sharding of the hot path: independent CPIs round-robin over ranks.

The path has no exchange step between CPIs (BASELINE config 4: "stream of independent CPIs sharded
round-robin"), so the data path needs no collective; the only communication is the final gather of
the finished maps to rank 0 (NCCL over NVLink on GPUs, gloo in the CPU tests).  One process per GPU.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch
import torch.distributed as dist


class Comm:
    """The C-ABI communicator (include/b200dd.h b200dd_comm_*: NCCL issued from C on its own stream).

    One per process / GPU.  The 128-byte NCCL id is created on rank 0 and handed to the other ranks through
    torch.distributed's default process group (any backend) -- plumbing; the data path below never touches torch.
    """

    def __init__(self, rank: int, world: int, device: int):
        from . import capi
        self._capi = capi
        self._lib = capi.load()
        ident = np.zeros(128, dtype=np.uint8)
        if rank == 0:
            capi.check(self._lib.b200dd_comm_get_unique_id(capi.ptr(ident)))
        if world > 1:
            t = torch.from_numpy(ident)
            if dist.get_backend() == "nccl":
                t = t.cuda(device)
            dist.broadcast(t, src=0)
            ident = t.cpu().numpy().copy()
        h = C.c_void_p()
        capi.check(self._lib.b200dd_comm_create(int(rank), int(world), capi.ptr(ident), int(device), C.byref(h)))
        self._h, self.rank, self.world = h, rank, world

    @staticmethod
    def _sp(stream):
        if stream is None:
            return None
        return C.c_void_p(stream.cuda_stream if hasattr(stream, "cuda_stream") else int(stream))

    def gather_async(self, send: torch.Tensor, recv, dst: int = 0, after=None):
        """Every rank's `send` (same byte size) lands on rank dst at recv[r]; returns at once."""
        nbytes = send.numel() * send.element_size()
        self._capi.check(self._lib.b200dd_comm_gather_async(self._h, self._capi.ptr(send),
                                                            self._capi.ptr(recv) if recv is not None else None,
                                                            nbytes, int(dst), self._sp(after)))

    def gatherv_async(self, send: torch.Tensor, recv, sizes, offsets, dst: int = 0, after=None):
        nbytes = send.numel() * send.element_size()
        sz = (C.c_size_t * self.world)(*[int(v) for v in sizes])
        of = (C.c_size_t * self.world)(*[int(v) for v in offsets])
        self._capi.check(self._lib.b200dd_comm_gatherv_async(self._h, self._capi.ptr(send), nbytes,
                                                             self._capi.ptr(recv) if recv is not None else None,
                                                             sz, of, int(dst), self._sp(after)))

    def allgatherv_async(self, send: torch.Tensor, recv: torch.Tensor, sizes, offsets, after=None):
        sz = (C.c_size_t * self.world)(*[int(v) for v in sizes])
        of = (C.c_size_t * self.world)(*[int(v) for v in offsets])
        self._capi.check(self._lib.b200dd_comm_allgatherv_async(self._h, self._capi.ptr(send), self._capi.ptr(recv), sz, of,
                                                                self._sp(after)))

    def allreduce_f64_async(self, buf: torch.Tensor, after=None):
        count = buf.numel() * (2 if buf.is_complex() else 1)
        self._capi.check(self._lib.b200dd_comm_allreduce_f64_async(self._h, self._capi.ptr(buf), count, self._sp(after)))

    def sendrecv_async(self, send, send_peer: int, recv, recv_peer: int, after=None):
        """One send (to send_peer) and / or one receive (from recv_peer) in one group; a peer of -1 skips that half."""
        sb = send.numel() * send.element_size() if (send is not None and send_peer >= 0) else 0
        rb = recv.numel() * recv.element_size() if (recv is not None and recv_peer >= 0) else 0
        self._capi.check(self._lib.b200dd_comm_sendrecv_async(self._h, self._capi.ptr(send) if sb else None, sb, int(send_peer),
                                                              self._capi.ptr(recv) if rb else None, rb, int(recv_peer),
                                                              self._sp(after)))

    def join(self, stream):
        """`stream` waits for everything enqueued on the communicator so far."""
        self._capi.check(self._lib.b200dd_comm_join(self._h, self._sp(stream)))

    def sync(self):
        self._capi.check(self._lib.b200dd_comm_sync(self._h))

    def close(self):
        if getattr(self, "_h", None):
            self._lib.b200dd_comm_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


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

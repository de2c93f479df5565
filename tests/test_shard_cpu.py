"""CPU, world_size 2, gloo: the multi-GPU sharding logic (CPI round-robin + final map gather)."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from blah2_b200.shard import cpi_indices, gather_maps


def test_round_robin_assignment_covers_every_cpi_once():
    for n, w in [(64, 8), (7, 2), (3, 4), (1, 2)]:
        seen = sorted(i for r in range(w) for i in cpi_indices(n, r, w))
        assert seen == list(range(n))
        assert max(len(cpi_indices(n, r, w)) for r in range(w)) - min(len(cpi_indices(n, r, w)) for r in range(w)) <= 1


def _worker(rank, world, port, n_cpis, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        idx = cpi_indices(n_cpis, rank, world)
        # a stand-in "map" whose content identifies the CPI it came from
        local = torch.stack([torch.full((5, 3), complex(c, -c), dtype=torch.complex64) for c in idx]) if idx else \
            torch.zeros((0, 5, 3), dtype=torch.complex64)
        out = gather_maps(local, n_cpis, rank, world)
        if rank == 0:
            q.put(out.numpy())
        else:
            assert out is None
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_cpis", [4, 7])
def test_gather_maps_world2_gloo(n_cpis):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + n_cpis
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_cpis, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert out.shape == (n_cpis, 5, 3)
    for c in range(n_cpis):
        assert np.all(out[c] == complex(c, -c))


# ---- one CPI split over ranks: orchestration checked with a CPU stand-in for the two CUDA stages ------
class _CpuAmbiguity:
    """Same interface as blah2_b200.process.Ambiguity.range_device / doppler_device, computed with the
    ORACLE on CPU tensors -- test infrastructure for the sharding logic only."""

    def __init__(self, geom):
        from types import SimpleNamespace
        from oracle import blah2_oracle as O
        self.O = O
        self.g = O.ambiguity_geometry(*geom)
        self.geometry = SimpleNamespace(n_doppler_bins=self.g.nDopplerBins, n_delay_bins=self.g.nDelayBins,
                                        n_corr=self.g.nCorr)

    def range_device(self, x, y, b0, nb, R, stream=None):
        import copy
        g = copy.copy(self.g)
        g.nDopplerBins = nb
        R.copy_(torch.from_numpy(self.O.range_matrix(x.numpy().astype(np.complex128), y.numpy().astype(np.complex128),
                                                     g).astype(np.complex64)))

    def doppler_device(self, R, c0, nc, tile, stream=None):
        full = self.O.doppler_transform(R.numpy().astype(np.complex128), self.g)
        tile.copy_(torch.from_numpy(full[:, c0:c0 + nc].astype(np.complex64)))


_GEOM = (-3, 20, -50, 50, 10000, 4000, True)


def _worker_cpi(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from blah2_b200.scene import random_iq
        from blah2_b200.shard import block_range, caf_single_cpi_sharded
        amb = _CpuAmbiguity(_GEOM)
        x, y = random_iq(_GEOM[5], 9)
        b0, nb = block_range(amb.g.nDopplerBins, rank, world)
        nC = amb.g.nCorr
        xl = torch.from_numpy(x[b0 * nC:(b0 + nb) * nC].astype(np.complex64))   # each rank holds only its slice
        yl = torch.from_numpy(y[b0 * nC:(b0 + nb) * nC].astype(np.complex64))
        m = caf_single_cpi_sharded(amb, xl, yl, rank, world)
        if rank == 0:
            q.put(m.numpy())
        else:
            assert m is None
    finally:
        dist.destroy_process_group()


def test_single_cpi_split_over_two_ranks_gloo():
    from blah2_b200.scene import random_iq
    from blah2_b200.shard import block_range
    from oracle import blah2_oracle as O
    assert [block_range(41, r, 8) for r in range(8)][:3] == [(0, 6), (6, 5), (11, 5)]
    assert sum(block_range(41, r, 8)[1] for r in range(8)) == 41
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker_cpi, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = q.get(timeout=180)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    x, y = random_iq(_GEOM[5], 9)
    ref, _, _ = O.ambiguity_process(x.astype(np.complex64), y.astype(np.complex64), O.ambiguity_geometry(*_GEOM))
    assert out.shape == ref.shape
    assert np.max(np.abs(out - ref)) / np.max(np.abs(ref)) < 1e-5


# ---- SingleCpiPlan (clutter filter + CAF of one CPI over the ranks) with CPU stand-ins for the CUDA stages ----------
class _CpuChunkFilter:
    """Same interface as blah2_b200.process.WienerHopfChunk, computed on CPU tensors from the DEFINITIONS
    (WienerHopf.cpp:76-108 sums restricted to the chunk, :125-160 filter) -- test infrastructure only."""

    def __init__(self, delayMin, delayMax, nSamples, c0, nc):
        self.nBins = delayMax - delayMin
        self.sh = -delayMin
        self.c0, self.nc = c0, nc
        self.xl, self.xr, self.yr = max(self.nBins - 1 - self.sh, 0), self.nBins - 1 + self.sh, self.nBins - 1

    def halos(self):
        return self.xl, self.xr, self.yr

    def _xs(self, x_loc, i):           # xs[i] = x[i + sh], i a global index array
        return x_loc[i + self.sh - (self.c0 - self.xl)]

    def corr_device(self, x_loc, y_loc, ab, stream=None):
        x, y = x_loc.numpy().astype(np.complex128), y_loc.numpy().astype(np.complex128)
        n = np.arange(self.c0, self.c0 + self.nc)
        a = np.array([np.sum(self._xs(x, n) * np.conj(self._xs(x, n + k))) for k in range(self.nBins)])
        b = np.array([np.sum(y[n + k - self.c0] * np.conj(self._xs(x, n))) for k in range(self.nBins)])
        ab.copy_(torch.from_numpy(np.concatenate([a, b])))

    def filter_device(self, ab, x_loc, y_loc, y_out, stream=None):
        import scipy.linalg as sla
        v = ab.numpy()
        a, b = v[:self.nBins], v[self.nBins:]
        # A(i,j) = a[j-i] (j >= i), conj(a[i-j]) (i > j)   (WienerHopf.cpp:85-97)
        A = sla.toeplitz(np.conj(a), a)
        w = np.linalg.solve(A, b)
        x, y = x_loc.numpy().astype(np.complex128), y_loc.numpy().astype(np.complex128)
        out = np.empty(self.nc, dtype=np.complex128)
        for j in range(self.nc):
            i = self.c0 + j
            k = np.arange(0, min(self.nBins, i + 1))
            out[j] = y[j] - np.sum(w[k] * self._xs(x, i - k))
        y_out.copy_(torch.from_numpy(out.astype(np.complex64)))


class _CpuAmbiguityPlus(_CpuAmbiguity):
    def place_tile(self, tile, c0, nc, m, stream=None):
        m[:, c0:c0 + nc] = tile.reshape(m.shape[0], nc)


_GEOM2 = (-3, 20, -50, 50, 10000, 4000, True)
_CLUT2 = (-2, 9)


def _worker_plan(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from blah2_b200.scene import make_scene
        from blah2_b200.shard import SingleCpiPlan, TorchComm
        amb = _CpuAmbiguityPlus(_GEOM2)
        sc = make_scene(_GEOM2[5], _GEOM2[4], seed=5, n_clutter=6)
        plan = SingleCpiPlan(TorchComm(rank, world), amb, _GEOM2[5], "cpu", clutter=_CLUT2, whc_factory=_CpuChunkFilter)
        plan.x_own.copy_(torch.from_numpy(sc.x[plan.s0:plan.s0 + plan.ns].astype(np.complex64)))   # each rank holds only its slice
        plan.y_own.copy_(torch.from_numpy(sc.y[plan.s0:plan.s0 + plan.ns].astype(np.complex64)))
        m = plan.run(None)
        if rank == 0:
            q.put(m.numpy().copy())
        else:
            assert m is None
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_single_cpi_plan_with_clutter_filter_gloo(world):
    """Halo exchange (circular right halos, one-sided left halo), all-reduce of the partial correlations, replicated
    solve, per-chunk filter, range stage on the rank's batches, all-gather, column tiles, gather + placement: the map
    on rank 0 equals the oracle's WienerHopf + Ambiguity on the whole CPI."""
    from blah2_b200.scene import make_scene
    from oracle import blah2_oracle as O
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 33500 + (os.getpid() % 2000) + world
    procs = [ctx.Process(target=_worker_plan, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = q.get(timeout=240)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    sc = make_scene(_GEOM2[5], _GEOM2[4], seed=5, n_clutter=6)
    x, y = sc.x.astype(np.complex64).astype(np.complex128), sc.y.astype(np.complex64).astype(np.complex128)
    ok, yf = O.wienerhopf_process(x, y, *_CLUT2)
    assert ok
    ref, _, _ = O.ambiguity_process(x, yf.astype(np.complex64), O.ambiguity_geometry(*_GEOM2))
    assert out.shape == ref.shape
    assert np.max(np.abs(out - ref)) / np.max(np.abs(ref)) < 1e-4   # complex64 hand-offs of y' between the stages

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

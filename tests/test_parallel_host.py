"""CPU tier: host-side logic of the element-slab partitioning (world_size 2 and 4, gloo)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from dg_multigrid_solver_b200.parallel import SlabPartition, exchange_halo, gather_rows, scatter_rows
    Ni, b, Nj = 5, 4, 8 * world
    part = SlabPartition(Nj, world, rank)
    gl, gh = int(part.has_lo), int(part.has_hi)
    rows = part.rows + gl + gh
    row = Ni * b
    # owned rows carry their global row index, ghost rows start at -1
    v = torch.full((rows * row,), -1.0, dtype=torch.float64)
    for jl in range(part.rows):
        v[(gl + jl) * row:(gl + jl + 1) * row] = float(part.j0 + jl)
    exchange_halo(v, Ni, b, gl, gh, rank, world)
    ok = True
    if gl:
        ok &= bool((v[:row] == part.j0 - 1).all())
    if gh:
        ok &= bool((v[-row:] == part.j1).all())
    # one-directional exchange (the pipeline of the exact lexicographic sweep)
    w = torch.full((rows * row,), -1.0, dtype=torch.float64)
    for jl in range(part.rows):
        w[(gl + jl) * row:(gl + jl + 1) * row] = float(part.j0 + jl)
    exchange_halo(w, Ni, b, gl, gh, rank, world, upward=False, downward=True)
    if gl:
        ok &= bool((w[:row] == -1).all())
    if gh:
        ok &= bool((w[-row:] == part.j1).all())
    owned = v[gl * row:(gl + part.rows) * row].clone()
    full = gather_rows(owned, world, rank)
    if rank == 0:
        ref = torch.arange(Nj, dtype=torch.float64).repeat_interleave(row)
        ok &= bool(torch.equal(full, ref))
        full = full * 2
    back = scatter_rows(full, owned, world, rank)
    ok &= bool(torch.equal(back, owned * 2))
    s = torch.tensor([float(owned.sum())], dtype=torch.float64)
    dist.all_reduce(s)
    ok &= abs(float(s) - row * Nj * (Nj - 1) / 2) < 1e-9
    out[rank] = ok
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_halo_exchange_gather_scatter(world):
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert all(out[r] for r in range(world)), dict(out)


def test_partition_and_level_split():
    from dg_multigrid_solver_b200.parallel import SlabPartition, distributed_levels, slab_nodes
    p = SlabPartition(2048, 8, 3)
    assert (p.j0, p.j1, p.rows, p.has_lo, p.has_hi) == (768, 1024, 256, True, True)
    assert not SlabPartition(2048, 8, 0).has_lo and not SlabPartition(2048, 8, 7).has_hi
    with pytest.raises(ValueError):
        SlabPartition(10, 4, 0)
    f = [2, 4, 8, 16, 32, 64, 128, 256, 512]
    assert distributed_levels(2048, 1, f) == [2, 4, 8, 16, 32, 64, 128, 256]
    assert distributed_levels(2048, 8, f) == [2, 4, 8, 16, 32]
    assert distributed_levels(2048, 8, f, min_rows=64) == [2, 4]
    xn = np.arange(9 * 5, dtype=np.float64).reshape(9, 5)       # 4 element rows, Pg = 2
    lx, ly, lo, hi = slab_nodes(xn, xn, 2, SlabPartition(4, 2, 1), halo=1)
    assert (lo, hi) == (1, 0) and lx.shape == (7, 5) and lx[0, 0] == xn[2, 0]

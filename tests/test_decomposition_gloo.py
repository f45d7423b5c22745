"""Host-side logic of the slab decomposition (dune-vof_b200/decomposition.py) on CPU: world_size 2 over gloo."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import dune_vof_b200.decomposition as D


def test_split_last_axis_covers_everything():
    for n in (7, 64, 513):
        for world in (1, 2, 3, 8):
            parts = [D.split_last_axis(n, world, r) for r in range(world)]
            assert parts[0][0] == 0
            for (lo, cnt), (lo2, _) in zip(parts, parts[1:]):
                assert lo + cnt == lo2
            assert parts[-1][0] + parts[-1][1] == n
            assert max(c for _, c in parts) - min(c for _, c in parts) <= 1


def test_halo_width_table():
    # Young's 2 (flags + reconstructions of the first halo layer are recomputed), height function 4, Swartz 3
    assert D.HALO_WIDTH == {0: 2, 1: 4, 2: 3}


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, halo, owned, layer, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # global field: value = global layer index * 1000 + position in layer
        lo = rank * owned
        t = torch.full((owned + 2 * halo, layer), -1.0, dtype=torch.float64)
        for k in range(owned):
            t[halo + k] = (lo + k) * 1000.0 + torch.arange(layer, dtype=torch.float64)
        ex = D.HaloExchanger(owned, halo, rank, world)
        ex.exchange(t)
        ok = True
        for k in range(-halo, owned + halo):
            g = lo + k
            row = t[halo + k]
            if 0 <= g < owned * world:
                ok = ok and bool(torch.equal(row, g * 1000.0 + torch.arange(layer, dtype=torch.float64)))
            else:
                ok = ok and bool(torch.all(row == -1.0))       # outside the global domain: untouched
        # the time-step estimate: MIN all-reduce (evolution.hh:75)
        dt = torch.tensor([0.5 + rank], dtype=torch.float64)
        dist.all_reduce(dt, op=dist.ReduceOp.MIN)
        ok = ok and float(dt) == 0.5
        out[rank] = 1 if ok else 0
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("halo", [2, 4])
def test_halo_exchange_world2(halo):
    world = 2
    out = mp.get_context("spawn").Array("i", [0] * world)
    port = _free_port()
    procs = [mp.get_context("spawn").Process(target=_worker, args=(r, world, port, halo, 6, 10, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert list(out) == [1] * world


def test_exchanger_rejects_thin_slab():
    with pytest.raises(ValueError):
        D.HaloExchanger(owned_layers=1, halo=2, rank=0, world=2)


def test_split_by_cost_balances_a_band():
    """a band of work in the middle third of the layers: the cost-weighted split gives every slab about the same cost,
    covers every layer once and respects the minimal thickness (the halo width)"""
    n, world = 1024, 8
    costs = np.full(n, 1.0)
    costs[200:520] += 400.0                      # the sphere's band: z in [0.2, 0.5] of a 1024^3 grid
    counts = D.split_by_cost(costs, world, min_layers=2)
    assert len(counts) == world and sum(counts) == n and min(counts) >= 2
    edges = np.concatenate([[0], np.cumsum(counts)])
    per_slab = [costs[edges[r]:edges[r + 1]].sum() for r in range(world)]
    assert max(per_slab) <= 1.15 * costs.sum() / world
    # uniform cost -> the even split
    assert D.split_by_cost(np.ones(64), 8, 2) == [8] * 8
    # all work in one layer: nothing better than isolating it, but every slab keeps its minimal thickness
    spike = np.zeros(40)
    spike[17] = 5.0
    c = D.split_by_cost(spike, 4, 3)
    assert sum(c) == 40 and min(c) >= 3
    with pytest.raises(ValueError):
        D.split_by_cost(np.ones(5), 4, 2)

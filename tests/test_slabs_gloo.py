"""k-slab partition + node-plane halo exchange, world_size 2 and 3 over gloo on CPU
(the NCCL path on GPUs uses the same plan and the same torch.distributed calls)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import curvigrid_b200 as cg
from curvigrid_b200 import slabs


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, ncell_slab, nhalo, nj, ni, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        plan = slabs.make_plan(rank, world, ncell_slab, nhalo)
        lo, hi = plan.required
        olo, ohi = plan.owned
        # global truth: node plane m holds the value 1000*m + j*10 + i (+ coord offset)
        k = torch.arange(lo, hi, dtype=torch.float64)[:, None, None]
        j = torch.arange(nj, dtype=torch.float64)[None, :, None]
        i = torch.arange(ni, dtype=torch.float64)[None, None, :]
        bufs = []
        for c in range(3):
            truth = 1000.0 * k + 10.0 * j + i + 0.25 * c
            b = truth.clone()
            b[: olo - lo] = -1.0          # planes owned by other ranks are unknown before the exchange
            b[ohi - lo:] = -1.0
            bufs.append(b)
        slabs.exchange_node_planes(plan, bufs)
        ok = all(torch.equal(b, 1000.0 * k + 10.0 * j + i + 0.25 * c) for c, b in enumerate(bufs))
        out[rank] = int(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,ncell_slab,nhalo", [(2, 12, 5), (2, 7, 2), (3, 20, 5), (4, 9, 2)])
def test_exchange_fills_required_planes(world, ncell_slab, nhalo):
    port = _free_port()
    out = mp.Manager().dict()
    mp.spawn(_worker, args=(world, port, ncell_slab, nhalo, 4, 5, out), nprocs=world, join=True)
    assert [out[r] for r in range(world)] == [1] * world


@pytest.mark.parametrize("world,ncell,nhalo", [(1, 16, 5), (2, 16, 5), (4, 64, 5), (8, 1536, 5), (3, 7, 2), (8, 10, 5)])
def test_plan_properties(world, ncell, nhalo):
    plans = [slabs.make_plan(r, world, ncell, nhalo) for r in range(world)]
    # cell planes: a partition of cell.full
    assert plans[0].cells[0] == 0 and plans[-1].cells[1] == ncell + 2 * nhalo
    for a, b in zip(plans[:-1], plans[1:]):
        assert a.cells[1] == b.cells[0]
    # owned node planes: a partition of the interior node planes 0..ncell
    covered = []
    for p in plans:
        covered += list(range(*p.owned))
        assert p.required[0] <= p.owned[0] and p.owned[1] <= p.required[1] or p.owned[0] == p.owned[1]
    assert covered == list(range(ncell + 1))
    # sends and receives pair up and cover every non-owned required plane exactly once
    for p in plans:
        need = set(range(*p.required)) - set(range(*p.owned))
        got = []
        for peer, lo, hi in p.recvs:
            assert (p.rank, lo, hi) in plans[peer].sends
            got += list(range(lo, hi))
        assert sorted(got) == sorted(need)
    # same rule as the C library's cg_grid_required_node_range (checked on the GPU box in test_gpu_slabs.py)
    for p in plans:
        k0, k1 = p.cells
        lo, hi = slabs.required_node_range(k0, k1, nhalo, ncell)
        assert 0 <= lo < hi <= ncell + 1
        assert lo <= max(min(k0 - 1 - nhalo, ncell), 0) and hi - 1 >= min(max(k1 + 2 - nhalo, 0), ncell)


def test_slab_grid_kwargs():
    plan = slabs.make_plan(1, 2, 16, 5)
    kw = slabs.slab_grid_kwargs(plan, (8, 6), 5)
    assert kw["global_celldims"] == (8, 6, 16)
    assert kw["window"][0] == (0, 0, plan.cells[0]) and kw["window"][1] == (18, 16, plan.cells[1] - plan.cells[0])
    assert kw["node_origin"] == (0, 0, plan.required[0])


@pytest.mark.parametrize("world", [1, 2, 3, 4, 7])
@pytest.mark.parametrize("ncell,nhalo", [(40, 5), (13, 2), (9, 0), (64, 3)])
def test_overlap_schedule_covers_slab_and_needs_owned_planes_only(world, ncell, nhalo):
    """slabs.overlap_schedule: the early work touches owned node planes only, early + late cover every padded plane
    and every cell plane exactly once (pure index algebra; the GPU side is tests/test_gpu_slabs.py)."""
    for rank in range(world):
        plan = slabs.make_plan(rank, world, ncell, nhalo)
        k0, k1 = plan.cells
        nw = k1 - k0
        if nw <= 0:
            continue
        s = slabs.overlap_schedule(plan)
        cover = np.zeros(nw, dtype=int)
        for lo, hi in s["eval_early"] + s["eval_late"]:
            cover[lo:hi] += 1
        assert (cover == 1).all(), (rank, world, s)
        pcover = np.zeros(nw + 4, dtype=int)
        for lo, hi in s["pad_early"] + s["pad_late"]:
            pcover[lo:hi] += 1
        assert (pcover == 1).all(), (rank, world, s)
        recv = set()
        for _, lo, hi in plan.recvs:
            recv.update(range(lo, hi))
        early_pad = set()
        for lo, hi in s["pad_early"]:
            early_pad.update(range(lo, hi))
        for e in early_pad:  # dependencies of a padded plane: clamp(m) and, when extrapolated, its inner neighbour
            m = k0 + e - 1 - nhalo
            c = min(max(m, 0), ncell)
            deps = {c} if m == c else {c, c + 1 if m < c else c - 1}
            assert not (deps & recv), (rank, world, e, deps)
        for lo, hi in s["eval_early"]:  # a cell plane reads the padded planes k .. k+4
            for k in range(lo, hi):
                assert set(range(k, k + 5)) <= early_pad, (rank, world, k)
        if world == 1:
            assert not s["eval_late"] and not s["pad_late"]

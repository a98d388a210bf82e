"""Slab windows on one GPU: every rank's slab, evaluated as its own handle from only the node
planes the plan says it holds, must reproduce the corresponding planes of the single-grid
result (same kernels, same padded values; 1e-13 because the compiler contracts FMAs differently
in a marching kernel's prologue and loop body, a 1-ulp effect)."""
import ctypes

import numpy as np
import pytest
import torch

from util import assert_metrics_close, grid_outputs, mild_nodes

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("profile", ["full", "compact"])
def test_slabs_reproduce_single_grid(cg, world, profile):
    h = 2
    nodes = mild_nodes((9, 8, 14), seed=31)          # [i,j,k], 13 cells along k
    ni, nj, nk = (s - 1 for s in nodes[0].shape)
    whole = cg.DiscreteGrid(*nodes, h, profile=profile)
    if profile == "full":
        ref = grid_outputs(cg, whole)
    else:
        ref = {"J": cg.cell_metrics(whole).cpu().numpy(),
               "act": [a.cpu().numpy() for a in cg.face_metrics(whole)]}
    plane = (ni + 2 * h) * (nj + 2 * h)
    for rank in range(world):
        plan = cg.slabs.make_plan(rank, world, nk, h)
        lo, hi = plan.required
        local = [torch.from_numpy(np.ascontiguousarray(a.transpose(2, 1, 0)[lo:hi])).cuda() for a in nodes]
        kw = cg.slabs.slab_grid_kwargs(plan, (ni, nj), h)
        g = cg.DiscreteGrid(*local, h, profile=profile, **kw)
        # the C library agrees with the Python plan about the planes it needs
        a, b = ctypes.c_int64(), ctypes.c_int64()
        cg._lib.check(cg._lib.lib().cg_grid_required_node_range(g._h, 2, ctypes.byref(a), ctypes.byref(b)))
        assert (a.value, b.value) == plan.required
        k0, k1 = plan.cells
        if profile == "full":
            got = grid_outputs(cg, g)
            want = {key: ref[key][..., k0 * plane:k1 * plane, :]
                    for key in ("cell_fwd", "cell_inv", "face_fwd", "face_inv", "face_cons")}
            assert_metrics_close(got, want, rtol=1e-13, label=f"slab {rank}/{world}")
        else:
            assert np.allclose(cg.cell_metrics(g).cpu().numpy(), ref["J"][k0:k1], rtol=1e-13, atol=0)
            for ax in range(3):
                a, b = cg.face_metrics(g)[ax].cpu().numpy(), ref["act"][ax][k0:k1]
                assert np.abs(a - b).max() <= 1e-13 * np.abs(b).max()
        res = cg.gcl(g)
        assert max(res) < 1e-13


@pytest.mark.parametrize("profile", ["full", "compact"])
def test_plane_ranged_pad_and_eval_match_one_evaluation(cg, profile):
    """cg_grid_pad_planes / cg_grid_eval_planes (the pieces slabs.step_overlapped launches around the exchange):
    padding and evaluating a slab in the early / late pieces of its overlap schedule gives the arrays of one
    cg_grid_eval (1e-13: chunk boundaries differ, see above)."""
    h = 3
    nodes = mild_nodes((35, 9, 30), seed=5)
    ni, nj, nk = (s - 1 for s in nodes[0].shape)
    world, rank = 3, 1
    plan = cg.slabs.make_plan(rank, world, nk, h)
    lo, hi = plan.required
    local = [torch.from_numpy(np.ascontiguousarray(a.transpose(2, 1, 0)[lo:hi])).cuda() for a in nodes]
    kw = cg.slabs.slab_grid_kwargs(plan, (ni, nj), h)
    ref_g = cg.DiscreteGrid(*local, h, profile=profile, **kw)
    g = cg.DiscreteGrid(*local, h, profile=profile, cache_mode="lazy", **kw)
    sched = cg.slabs.overlap_schedule(plan)
    assert sched["eval_early"] and len(sched["eval_late"]) == 2
    lib, st = cg._lib.lib(), g._stream()
    g._bind_metric_outputs(3)
    i64 = ctypes.c_int64
    for a, b in sched["pad_early"]:
        cg._lib.check(lib.cg_grid_pad_planes(g._h, i64(a), i64(b), st))
    for a, b in sched["eval_early"]:
        cg._lib.check(lib.cg_grid_eval_planes(g._h, 3, i64(a), i64(b), 0, st))
    assert cg.cell_metrics_valid(g) is False and cg.face_metrics_valid(g) is False
    for a, b in sched["pad_late"]:
        cg._lib.check(lib.cg_grid_pad_planes(g._h, i64(a), i64(b), st))
    for a, b in sched["eval_late"]:
        cg._lib.check(lib.cg_grid_eval_planes(g._h, 3, i64(a), i64(b), 0, st))
    cg._lib.check(lib.cg_grid_eval_planes(g._h, 3, i64(0), i64(0), 1, st))
    assert cg.cell_metrics_valid(g) and cg.face_metrics_valid(g)
    if profile == "full":
        assert_metrics_close(grid_outputs(cg, g), grid_outputs(cg, ref_g), rtol=1e-13, label="ranged")
    else:
        assert np.allclose(cg.cell_metrics(g).cpu().numpy(), cg.cell_metrics(ref_g).cpu().numpy(), rtol=1e-13, atol=0)
        for ax in range(3):
            a, b = cg.face_metrics(g)[ax].cpu().numpy(), cg.face_metrics(ref_g)[ax].cpu().numpy()
            assert np.abs(a - b).max() <= 1e-13 * np.abs(b).max()
    # argument checks
    with pytest.raises(cg.CurvigridArgumentError):
        cg._lib.check(lib.cg_grid_eval_planes(g._h, 3, i64(2), i64(1), 0, st))
    with pytest.raises(cg.CurvigridArgumentError):
        cg._lib.check(lib.cg_grid_pad_planes(g._h, i64(0), i64(10 ** 6), st))


@pytest.mark.parametrize("nchunks", [1, 3, 8, 1000])
@pytest.mark.parametrize("profile", ["full", "compact"])
def test_update_from_host_pipelined(cg, nchunks, profile):
    """cg_grid_update_from_host: pipelined H2D + pad + evaluation in pieces gives the arrays of set_nodes + one
    evaluation, for new node values (update!(grid, x, y, z), discrete_grid.jl:605-647)."""
    h = 3
    nodes0 = mild_nodes((20, 9, 26), seed=1)
    nodes1 = mild_nodes((20, 9, 26), seed=2)
    dev0 = [torch.from_numpy(np.ascontiguousarray(a.transpose(2, 1, 0))).cuda() for a in nodes0]
    g = cg.DiscreteGrid(*dev0, h, profile=profile, cache_mode="lazy")
    cg.refresh_metrics(g)
    pinned = [torch.from_numpy(np.ascontiguousarray(a.transpose(2, 1, 0))).pin_memory() for a in nodes1]
    g.update_from_host(*pinned, nchunks=nchunks)
    assert cg.cell_metrics_valid(g) and cg.face_metrics_valid(g)
    ref_g = cg.DiscreteGrid(*nodes1, h, profile=profile)
    if profile == "full":
        assert_metrics_close(grid_outputs(cg, g), grid_outputs(cg, ref_g), rtol=5e-13, label="pipelined")
    else:
        assert np.allclose(cg.cell_metrics(g).cpu().numpy(), cg.cell_metrics(ref_g).cpu().numpy(), rtol=5e-13, atol=0)
        for ax in range(3):
            a, b = cg.face_metrics(g)[ax].cpu().numpy(), cg.face_metrics(ref_g)[ax].cpu().numpy()
            assert np.abs(a - b).max() <= 5e-13 * np.abs(b).max()
    assert max(cg.gcl(g)) < 1e-12


@pytest.mark.parametrize("profile", ["full", "compact"])
def test_update_from_host_accumulates_gcl_plane_by_plane(cg, profile):
    """CG_WHAT_GCL: the residuals reduced behind the pieces of the pipelined host update are the residuals of one
    reduction over the finished arrays, bit for bit (a max does not depend on the order), a poisoned conserved entry still
    comes back as NaN from a fresh reduction, and any later evaluation drops the accumulated values."""
    h = 3
    nodes = mild_nodes((20, 9, 26), seed=4)
    pinned = [torch.from_numpy(np.ascontiguousarray(a.transpose(2, 1, 0))).pin_memory() for a in nodes]
    g = cg.DiscreteGrid(*nodes, h, profile=profile, cache_mode="lazy", conserved_metric_scheme="endpoint")
    cg.refresh_metrics(g)
    want = cg.gcl(g)                       # one reduction over the whole grid
    launches0 = cg._lib.lib().cg_launch_count()
    g.update_from_host(*pinned, nchunks=5, gcl=True)
    n_update = cg._lib.lib().cg_launch_count() - launches0
    got = cg.gcl(g)                        # reads the accumulated maxima: no further launch
    assert cg._lib.lib().cg_launch_count() - launches0 == n_update
    assert got == want and max(got) > 0.0
    cg.refresh_metrics(g)                  # another evaluation: the accumulated values are stale, gcl reduces again
    launches1 = cg._lib.lib().cg_launch_count()
    assert cg.gcl(g) == want
    assert cg._lib.lib().cg_launch_count() == launches1 + 1

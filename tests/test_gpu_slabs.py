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

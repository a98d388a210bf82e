"""The FUSED cell+face evaluation (`refresh_metrics`, eager constructors, `update!` + refresh) runs the
compile-time configurations of the marching kernels (MODE_FULL / MODE_COMPACT, csrc/cg_kernels_tiled.cuh), the
staged 2-D kernel and the staged MappedGrid kernel; `cell_metrics` + `face_metrics` on a lazy grid run MODE_CELL
and MODE_FACE.  This module checks the fused paths against the oracle and against the separate refreshes, over
chunk boundaries of the marching axis, ragged warp tiles, FP32 and the COMPACT profile.

Tolerance: FP64 relative error <= 1e-12 (BASELINE.json north_star); FP32 as in test_gpu_parity.py.
"""
import ctypes

import numpy as np
import pytest

from util import RTOL_F64, assert_fp32_close, assert_metrics_close, domain_indices, grid_outputs, mild_nodes

pytestmark = pytest.mark.gpu
SCHEMES = ("endpoint", "gradient", "curvature")


def _fused(cg, nodes, nhalo, scheme, kernel=1, **kw):
    g = cg.DiscreteGrid(*nodes, nhalo, conserved_metric_scheme=scheme, cache_mode="lazy", **kw)
    g.set_kernel(kernel)
    cg.refresh_metrics(g)  # ONE evaluation with CG_WHAT_CELL | CG_WHAT_FACE
    assert cg.cell_metrics_valid(g) and cg.face_metrics_valid(g)
    return g


@pytest.mark.parametrize("scheme", SCHEMES)
@pytest.mark.parametrize("shape,nhalo", [((8, 7, 6), 2), ((37, 9, 8), 2), ((5, 9, 4), 5), ((12, 9), 3), ((33, 29), 5), ((9,), 2)])
def test_fused_refresh_matches_oracle(cg, oracle, shape, nhalo, scheme):
    nodes = mild_nodes(shape, seed=11 + len(shape))
    ref = oracle.discrete_eval(nodes, nhalo, scheme)
    got = grid_outputs(cg, _fused(cg, nodes, nhalo, scheme))
    assert got["nfull"] == ref["nfull"]
    assert_metrics_close(got, ref, rtol=RTOL_F64, label=f"fused {shape} h={nhalo} {scheme}")


@pytest.mark.parametrize("scheme", SCHEMES)
@pytest.mark.parametrize("shape", [(12, 6, 70), (6, 70, 12), (40, 5, 131)])
def test_fused_marching_over_chunk_boundaries(cg, scheme, shape):
    """Marching axes longer than one chunk (64 steps): the carried state of a chunk (line-stencil pipeline, lagged
    conserved metrics, node ring) is re-primed at every chunk start.  Reference: the gather kernels (thread per
    cell, no carried state), themselves checked against the oracle in test_gpu_parity.py.

    Tolerance 4e-12: on a mesh this long the coordinates reach ~100 cell sizes, the potentials (d_b q1) q2 are ~100 x
    the metric they difference to, and BOTH kernel families (and the oracle) sit at 1.5e-12 .. 1.9e-12 of each other
    (tools/check_vs_oracle.py); a chunk-boundary bug shows up as an O(1) error."""
    nodes = mild_nodes(shape, seed=31)
    ref = grid_outputs(cg, _fused(cg, nodes, 2, scheme, kernel=0))
    got = grid_outputs(cg, _fused(cg, nodes, 2, scheme, kernel=1))
    assert_metrics_close(got, ref, rtol=4e-12, label=f"chunks {shape} {scheme}")
    sep = cg.DiscreteGrid(*nodes, 2, conserved_metric_scheme=scheme, cache_mode="lazy")
    assert_metrics_close(grid_outputs(cg, sep), ref, rtol=4e-12, label=f"chunks separate {shape} {scheme}")


def test_fused_equals_separate_refreshes(cg):
    """MODE_FULL and MODE_CELL + MODE_FACE are the same source compiled with different switches: they agree to
    round-off (the compiler contracts FMAs differently, so not bit for bit)."""
    nodes = mild_nodes((37, 11, 9), seed=5)
    a = grid_outputs(cg, _fused(cg, nodes, 3, "curvature"))
    b = grid_outputs(cg, cg.DiscreteGrid(*nodes, 3, cache_mode="lazy"))
    assert_metrics_close(a, b, rtol=1e-13, label="fused vs separate")


@pytest.mark.parametrize("shape,nhalo", [((35, 9, 70), 2), ((12, 9), 3)])
def test_fused_compact_profile(cg, shape, nhalo):
    nodes = mild_nodes(shape, seed=7)
    N = len(shape)
    fo = grid_outputs(cg, _fused(cg, nodes, nhalo, "curvature"))
    comp = _fused(cg, nodes, nhalo, "curvature", profile="compact")
    J = cg.cell_metrics(comp).reshape(-1).cpu().numpy()
    assert np.abs(J - fo["cell_fwd"][:, N * N]).max() <= 1e-13 * np.abs(J).max()
    act = cg.face_metrics(comp)
    for a in range(N):
        got = act[a].reshape(-1, N).cpu().numpy()
        want = np.stack([fo["face_cons"][a][:, a + N * c] for c in range(N)], axis=1)
        assert np.abs(got - want).max() <= 1e-12 * np.abs(want).max()
    assert max(cg.gcl(comp)) < 1e-13


@pytest.mark.parametrize("shape,nhalo", [((8, 7, 6), 2), ((12, 9), 3)])
def test_fused_fp32(cg, oracle, shape, nhalo):
    nodes32 = [a.astype(np.float32) for a in mild_nodes(shape, seed=8)]
    ref = oracle.discrete_eval([a.astype(np.float64) for a in nodes32], nhalo, "curvature")
    got = grid_outputs(cg, _fused(cg, nodes32, nhalo, "curvature", T=np.float32))
    dom = domain_indices(ref["nfull"], nhalo)
    assert_fp32_close(got, ref, nodes32, dom, label="fused fp32")   # eps32 x conditioning (util.fp32_tolerances)


def test_nodes_changed_then_fused_refresh(cg, oracle):
    """The bench's step: the caller rewrites its device node buffers in place, cg_grid_nodes_changed re-pads them and
    ONE fused evaluation refreshes both caches."""
    import torch
    nodes = mild_nodes((9, 8, 7), seed=2)
    as_dev = lambda a: torch.from_numpy(np.ascontiguousarray(a.transpose(2, 1, 0))).cuda()  # memory order [k, j, i]
    dev = [as_dev(a) for a in nodes]
    g = cg.DiscreteGrid(*dev, 2, conserved_metric_scheme="gradient", cache_mode="lazy")
    cg.refresh_metrics(g)
    assert_metrics_close(grid_outputs(cg, g), oracle.discrete_eval(nodes, 2, "gradient"), rtol=RTOL_F64, label="initial")
    moved = [np.asfortranarray(a + 0.01 * np.sin(1.3 * a)) for a in nodes]
    for d, m in zip(dev, moved):
        d.copy_(as_dev(m))
    cg._lib.check(cg._lib.lib().cg_grid_nodes_changed(g._h, g._stream()))
    cg.refresh_metrics(g)
    assert_metrics_close(grid_outputs(cg, g), oracle.discrete_eval(moved, 2, "gradient"), rtol=RTOL_F64, label="moved nodes")
    assert max(cg.gcl(g)) < 5e-13


def test_bound_outputs_must_be_16_byte_aligned(cg):
    """The kernels write with 16-B vector stores and TMA bulk copies (include/curvigrid_cuda.h)."""
    import torch
    nodes = mild_nodes((6, 5, 4), seed=1)
    g = cg.DiscreteGrid(*nodes, 2, cache_mode="lazy")
    L = cg._lib
    ncell = int(np.prod(g.nfull))
    buf = torch.zeros(ncell * 10 + 2, dtype=torch.float64, device="cuda")
    rc = L.lib().cg_grid_bind_output(g._h, L.CG_ARR_CELL_FORWARD, 0, ctypes.c_void_p(buf.data_ptr() + 8),
                                     ctypes.c_size_t(ncell * 80))
    assert rc == L.CG_ERR_ARG and b"aligned" in L.lib().cg_last_error()
    rc = L.lib().cg_grid_bind_output(g._h, L.CG_ARR_CELL_FORWARD, 0, ctypes.c_void_p(buf.data_ptr()),
                                     ctypes.c_size_t(ncell * 80))
    assert rc == L.CG_OK

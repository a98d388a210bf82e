"""Parity of the CUDA path (through the C ABI) with the oracle — the `-m gpu` gate.

Tolerance: FP64 relative error <= 1e-12 (BASELINE.json north_star), measured per
cell against the largest entry of the same matrix (J against itself), over ALL of
cell.full unless stated.  GCL residuals use the reference's own bounds
(test_2d.jl:183, test_3d.jl:190-196, test_3d_continuous.jl:60-62).
"""
import ctypes

import numpy as np
import pytest

from util import RTOL_F64, assert_fp32_close, assert_metrics_close, assert_three_way, domain_indices, grid_outputs, metric_err, mild_nodes

pytestmark = pytest.mark.gpu

SCHEMES = ("endpoint", "gradient", "curvature")
KERNELS = (0, 1)  # 0 gather reference kernels, 1 tiled kernels


def _make(cg, nodes, nhalo, scheme, kernel, **kw):
    g = cg.DiscreteGrid(*nodes, nhalo, conserved_metric_scheme=scheme, cache_mode="lazy", **kw)
    g.set_kernel(kernel)
    return g


@pytest.mark.parametrize("kernel", KERNELS)
@pytest.mark.parametrize("scheme", SCHEMES)
@pytest.mark.parametrize("shape,nhalo", [((9,), 2), ((12, 9), 3), ((8, 7, 6), 2), ((5, 9, 4), 5)])
def test_discrete_matches_oracle(cg, oracle, shape, nhalo, scheme, kernel):
    nodes = mild_nodes(shape, seed=len(shape))
    ref = oracle.discrete_eval(nodes, nhalo, scheme)
    g = _make(cg, nodes, nhalo, scheme, kernel)
    got = grid_outputs(cg, g)
    assert got["nfull"] == ref["nfull"]
    assert_metrics_close(got, ref, rtol=RTOL_F64, label=f"{shape} h={nhalo} {scheme} k{kernel}")


@pytest.mark.parametrize("kernel", (17, 18, 20))
@pytest.mark.parametrize("scheme", SCHEMES)
def test_each_marching_kernel_alone(cg, oracle, scheme, kernel):
    """Development masks: one marching kernel (xi+cell / eta / zeta) + gather for the other axes."""
    nodes = mild_nodes((37, 9, 8), seed=21)   # > 29 cells along xi: two warp tiles
    ref = oracle.discrete_eval(nodes, 2, scheme)
    g = _make(cg, nodes, 2, scheme, kernel)
    assert_metrics_close(grid_outputs(cg, g), ref, rtol=RTOL_F64, label=f"mask {kernel - 16} {scheme}")


@pytest.mark.parametrize("kernel", KERNELS)
def test_halo_coords_included_discards_caller_halo(cg, oracle, kernel):
    """discrete_grid.jl:55-61, 556-566: the caller's halo nodes are stripped and re-created
    by linear extrapolation."""
    h = 2
    inner = mild_nodes((7, 6, 5), seed=7)
    padded = []
    rng = np.random.default_rng(3)
    for a in inner:
        big = rng.standard_normal(tuple(s + 2 * h for s in a.shape))  # garbage halo
        big[h:-h, h:-h, h:-h] = a
        padded.append(np.asfortranarray(big))
    ref = oracle.discrete_eval(inner, h, "curvature")
    g = _make(cg, padded, h, "curvature", kernel, halo_coords_included=True)
    assert_metrics_close(grid_outputs(cg, g), ref, label="halo_coords_included")


@pytest.mark.parametrize("kernel", KERNELS)
@pytest.mark.parametrize("shape,nhalo", [((12, 9), 3), ((8, 7, 6), 2)])
def test_fp32_variant(cg, oracle, shape, nhalo, kernel):
    nodes32 = [a.astype(np.float32) for a in mild_nodes(shape, seed=11, noise=0.0)]
    ref = oracle.discrete_eval([a.astype(np.float64) for a in nodes32], nhalo, "curvature")  # same (rounded) inputs
    g = _make(cg, nodes32, nhalo, "curvature", kernel, T=np.float32)
    got = grid_outputs(cg, g)
    dom = domain_indices(ref["nfull"], nhalo)
    assert_fp32_close(got, ref, nodes32, dom, label="fp32")   # eps32 x conditioning (util.fp32_tolerances)


@pytest.mark.parametrize("kernel", KERNELS)
@pytest.mark.parametrize("shape,nhalo", [((12, 9), 3), ((8, 7, 6), 2)])
def test_compact_profile_is_a_view_of_full(cg, shape, nhalo, kernel):
    nodes = mild_nodes(shape, seed=5)
    N = len(shape)
    full = _make(cg, nodes, nhalo, "curvature", kernel)
    comp = _make(cg, nodes, nhalo, "curvature", kernel, profile="compact")
    fo = grid_outputs(cg, full)
    J = cg.cell_metrics(comp).reshape(-1).cpu().numpy()
    assert np.abs(J - fo["cell_fwd"][:, N * N]).max() <= 1e-13 * np.abs(J).max()
    act = cg.face_metrics(comp)
    for a in range(N):
        row = fo["face_cons"][a].reshape(-1, N, N)[:, :, a]  # [cell][c] of entry (r=a, c): index a + N*c
        got = act[a].reshape(-1, N).cpu().numpy()
        want = np.stack([fo["face_cons"][a][:, a + N * c] for c in range(N)], axis=1)
        assert np.abs(got - want).max() <= 1e-13 * np.abs(want).max()
    assert max(cg.gcl(comp)) < 1e-13


@pytest.mark.parametrize("shape,nhalo", [((9,), 2), ((12, 9), 3), ((8, 7, 6), 2)])
def test_coordinates_match_oracle(cg, oracle, shape, nhalo):
    """generation.jl:187-309: node / centroid / high-side face coordinates."""
    nodes = mild_nodes(shape, seed=2)
    N = len(shape)
    g = _make(cg, nodes, nhalo, "curvature", 1)
    ref_n = oracle.discrete_coords(nodes, nhalo, 0)
    ref_c = oracle.discrete_coords(nodes, nhalo, 1)
    for q in range(N):
        assert np.abs(g.node_coordinates[q].reshape(-1).cpu().numpy() - ref_n[q]).max() < 1e-12
        assert np.abs(g.centroid_coordinates[q].reshape(-1).cpu().numpy() - ref_c[q]).max() < 1e-12
    for a in range(N):
        ref_f = oracle.discrete_coords(nodes, nhalo, 2 + a)
        for q in range(N):
            assert np.abs(g.face_coordinates[a][q].reshape(-1).cpu().numpy() - ref_f[q]).max() < 1e-12


# ---------------------------------------------------------------------------
# the reference's own known-answer / property tests, on the GPU
# ---------------------------------------------------------------------------
@pytest.mark.parametrize("kernel", KERNELS)
def test_rectilinear_3d_known_answers(cg, kernel):
    """test_3d.jl:40-135 (all of cell.full, like the halo-included variant :112-135)."""
    wl = cg.workloads
    x, y, z, dx, dy, dz = wl.rectilinear_nodes_3d(0.0, 2.0, 1, 3, -1, 2, 10, 20, 30)
    g = _make(cg, [x, y, z], 5, "curvature", kernel)
    o = grid_outputs(cg, g)
    J = dx * dy * dz
    fwd = np.array([dx, 0, 0, 0, dy, 0, 0, 0, dz, J])
    inv = np.array([1 / dx, 0, 0, 0, 1 / dy, 0, 0, 0, 1 / dz, J])
    cons = np.array([J / dx, 0, 0, 0, J / dy, 0, 0, 0, J / dz])
    assert np.abs(o["cell_fwd"] - fwd).max() < 1e-14
    assert np.abs(o["cell_inv"] - inv).max() < 1e-12
    assert np.abs(o["face_fwd"] - fwd).max() < 1e-13
    assert np.abs(o["face_inv"] - inv).max() < 1e-12
    assert np.abs(o["face_cons"] - cons).max() < 1e-14
    assert max(cg.gcl(g)) < np.finfo(np.float64).eps


@pytest.mark.parametrize("kernel", KERNELS)
def test_rectilinear_2d_known_answers(cg, kernel):
    """test_2d.jl:18-63, 124-151"""
    wl = cg.workloads
    x, y, dx, dy = wl.rectilinear_nodes_2d(0, 2, 1, 3, 40, 80)
    g = _make(cg, [x, y], 5, "curvature", kernel)
    o = grid_outputs(cg, g)
    J = dx * dy
    assert np.abs(o["cell_fwd"] - np.array([dx, 0, 0, dy, J])).max() < 1e-14
    assert np.abs(o["cell_inv"] - np.array([1 / dx, 0, 0, 1 / dy, J])).max() < 1e-12
    assert np.abs(o["face_fwd"] - np.array([dx, 0, 0, dy, J])).max() < 1e-13
    assert np.abs(o["face_cons"] - np.array([dy, 0, 0, dx])).max() < 1e-14


@pytest.mark.parametrize("kernel", KERNELS)
@pytest.mark.parametrize("scheme", SCHEMES)
def test_wavy_gcl_bounds(cg, scheme, kernel):
    """test_2d.jl:153-238 (< 5e-15) and test_3d.jl:149-245 (< 5e-13), incl. halo-included input."""
    wl = cg.workloads
    g2 = _make(cg, wl.wavy_nodes_2d(41, 41), 5, scheme, kernel)
    assert max(cg.gcl(g2)) < 5e-15
    g3 = _make(cg, wl.wavy_nodes_3d(20, 20, 20), 5, scheme, kernel)
    assert max(cg.gcl(g3)) < 5e-13
    g3h = _make(cg, wl.wavy_nodes_3d(20, 20, 20), 5, scheme, kernel, halo_coords_included=True)
    assert max(cg.gcl(g3h)) < 5e-13


# ---------------------------------------------------------------------------
# larger sizes: size-independent properties + sampled oracle cells + kernel cross-check
# ---------------------------------------------------------------------------
def test_3d_128_sampled_oracle_and_kernel_crosscheck(cg, oracle):
    wl = cg.workloads
    nodes = wl.wavy_nodes_3d(129, 129, 129)
    h = 5
    g1 = _make(cg, nodes, h, "curvature", 1)
    o1 = grid_outputs(cg, g1)
    assert max(cg.gcl(g1)) < 5e-13
    rng = np.random.default_rng(0)
    ntot = int(np.prod(o1["nfull"]))
    sample = np.sort(rng.choice(ntot, size=96, replace=False))
    got = {k: (o1[k][..., sample, :]) for k in ("cell_fwd", "cell_inv", "face_fwd", "face_inv", "face_cons")}
    # cell.domain within 1e-12 of the long-double evaluation of the reference's algorithm (util.assert_three_way)
    rep = assert_three_way(oracle, got, nodes, h, "curvature", sample, o1["nfull"], label="128^3 sampled")
    assert rep["face_cons"]["domain_gpu_vs_truth"] <= 1e-12
    g0 = _make(cg, nodes, h, "curvature", 0)
    o0 = grid_outputs(cg, g0)
    dom = domain_indices(o1["nfull"], h)
    # two independently rounded evaluations, each within 1e-12 of the exact value
    assert_metrics_close(o1, o0, rtol=4 * RTOL_F64, where=dom, label="tiled vs gather")


def test_2d_1024_gradient_properties(cg, oracle):
    wl = cg.workloads
    nodes = wl.wavy_nodes_2d(1025, 1025)
    g = _make(cg, nodes, 5, "gradient", 1)
    o = grid_outputs(cg, g)
    assert max(cg.gcl(g)) < 5e-15
    rng = np.random.default_rng(1)
    sample = np.sort(rng.choice(int(np.prod(o["nfull"])), size=256, replace=False))
    got = {k: (o[k][..., sample, :]) for k in ("cell_fwd", "cell_inv", "face_fwd", "face_inv", "face_cons")}
    assert_three_way(oracle, got, nodes, 5, "gradient", sample, o["nfull"], label="1024^2 sampled")


# ---------------------------------------------------------------------------
# cache lifecycle and error behaviour (test_unified_grid_types.jl:84-137, 139-198)
# ---------------------------------------------------------------------------
def test_cache_lifecycle(cg):
    nodes = mild_nodes((6, 5, 4), seed=4)
    g = cg.DiscreteGrid(*nodes, 2)  # eager
    assert cg.cell_metrics_valid(g) and cg.face_metrics_valid(g)
    cg.update(g, 0.5, {})
    assert not cg.cell_metrics_valid(g) and not cg.face_metrics_valid(g)
    cg.cell_metrics(g)
    assert cg.cell_metrics_valid(g) and not cg.face_metrics_valid(g)
    cg.face_metrics(g)
    assert cg.face_metrics_valid(g)
    cg.invalidate_face_metrics(g)
    assert cg.cell_metrics_valid(g) and not cg.face_metrics_valid(g)
    cg.refresh_face_metrics(g)
    assert cg.face_metrics_valid(g)
    lazy = cg.DiscreteGrid(*nodes, 2, cache_mode="lazy")
    assert not cg.cell_metrics_valid(lazy)
    off = cg.DiscreteGrid(*nodes, 2, compute_metrics=False, cache_mode="off")
    with pytest.raises(ValueError):
        cg.cell_metrics(off)


def test_error_behaviour(cg):
    nodes = mild_nodes((6, 5), seed=4)
    with pytest.raises(TypeError):
        cg.DiscreteGrid(*nodes, 2, conserved_metric_scheme="nope")
    with pytest.raises(ValueError):
        cg.DiscreteGrid(*nodes, 2, cache_mode="sometimes")
    with pytest.raises(ValueError):
        cg.DiscreteGrid(nodes[0], nodes[1][:-1], 2)
    with pytest.raises(ValueError):
        cg.DiscreteGrid(*nodes, 2, interpolation="cubic")
    L = cg._lib
    h = ctypes.c_void_p()
    rc = L.lib().cg_grid_create(0, 4, 0, L.i64x3((4, 4, 4)), 2, 2, 0, None, None, None, ctypes.byref(h))
    assert rc == L.CG_ERR_ARG and b"dimension" in L.lib().cg_last_error()
    rc = L.lib().cg_grid_create(0, 2, 0, L.i64x3((4, 4)), 2, 9, 0, None, None, None, ctypes.byref(h))
    assert rc == L.CG_ERR_ARG


def test_host_buffer_entry_point(cg, oracle):
    """cg_discrete_metrics_host: host pointers in, host pointers out, one call."""
    nodes = mild_nodes((7, 6, 5), seed=9)
    h = 2
    ref = oracle.discrete_eval(nodes, h, "gradient")
    nc = int(np.prod(ref["nfull"]))
    out = {"cell_fwd": np.zeros((nc, 10)), "cell_inv": np.zeros((nc, 10)), "face_fwd": np.zeros((3, nc, 10)),
           "face_inv": np.zeros((3, nc, 10)), "face_cons": np.zeros((3, nc, 9))}
    L = cg._lib
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    flat = [np.asfortranarray(a) for a in nodes]
    rc = L.lib().cg_discrete_metrics_host(0, 3, L.CG_F64, p(flat[0]), p(flat[1]), p(flat[2]),
                                          L.i64x3([s - 1 for s in nodes[0].shape]), h, 0, 1, p(out["cell_fwd"]),
                                          p(out["cell_inv"]), p(out["face_fwd"]), p(out["face_inv"]),
                                          p(out["face_cons"]))
    L.check(rc)
    assert_metrics_close(out, ref, label="host entry point")


# ---------------------------------------------------------------------------
# BASELINE.json full sizes (configs 2 and 3): size-independent properties + sampled oracle cells.  The arrays stay on
# the device (122 GB at 512^3); sampled cells are gathered there.
# ---------------------------------------------------------------------------
def _sampled(cg, grid, sample):
    import torch
    idx = torch.from_numpy(sample).to(f"cuda:{grid.device}")
    N = grid.dim
    K = N * N + 1
    cm, fm = cg.cell_metrics(grid), cg.face_metrics(grid)
    pick = lambda t, k: t.reshape(-1, k)[idx].cpu().numpy()
    return {"cell_fwd": pick(cm.forward, K), "cell_inv": pick(cm.inverse, K),
            "face_fwd": np.stack([pick(fm[a].forward, K) for a in range(N)]),
            "face_inv": np.stack([pick(fm[a].inverse, K) for a in range(N)]),
            "face_cons": np.stack([pick(fm[a].conserved, N * N) for a in range(N)])}


def test_3d_512_full_size_properties(cg, oracle):
    """BASELINE config 3 (3-D DiscreteGrid 512^3 wavy, FULL cell+face, Curvature): GCL at round-off over the whole
    domain, 64 random cells of cell.full against the oracle, and the exact scaling law of the metrics under
    x -> 2x (a power of two commutes with every floating-point operation: J -> 8 J, conserved rows -> 4 x, bitwise)."""
    import torch
    nodes = cg.workloads.wavy_nodes_3d(513, 513, 513)
    g = cg.DiscreteGrid(*nodes, 5, cache_mode="lazy")
    cg.refresh_metrics(g)
    assert max(cg.gcl(g)) < 5e-13
    rng = np.random.default_rng(7)
    nfull = tuple(n + 10 for n in (512, 512, 512))
    sample = np.sort(rng.choice(int(np.prod(nfull)), size=64, replace=False))
    got = _sampled(cg, g, sample)
    # North-star tolerance at the BASELINE size: on cell.domain every entry is within 1e-12 of the EXACT value of the
    # reference's algorithm on these inputs (long-double oracle).  The reference's own Float64 arithmetic (FP64 oracle)
    # is 7e-12 from that value here — eps (|x|/dx)^2, |x|/dx = 256 — so a direct 1e-12 comparison of two FP64
    # evaluations is not meaningful at this size; the kernels take first differences between neighbouring nodes and
    # stay at eps |x|/dx (6e-14 measured, tools/roundoff_gpu.py).
    rep = assert_three_way(oracle, got, nodes, 5, "curvature", sample, nfull, label="512^3 sampled")
    assert rep["face_cons"]["domain_gpu_vs_truth"] < rep["face_cons"]["domain_oracle64_vs_truth"]
    # a second sample inside cell.domain only, near the corners of the box where |x| / dx is largest
    far = np.concatenate([rng.integers(nfull[0] - 40, nfull[0] - 6, (24, 3)), rng.integers(6, 40, (24, 3))])
    s2 = np.sort(far[:, 0] + nfull[0] * (far[:, 1] + nfull[1] * far[:, 2])).astype(np.int64)
    assert_three_way(oracle, _sampled(cg, g, s2), nodes, 5, "curvature", s2, nfull, label="512^3 corners")
    del g
    torch.cuda.empty_cache()
    g1 = cg.DiscreteGrid(*nodes, 5, cache_mode="lazy", profile="compact")
    g2 = cg.DiscreteGrid(*[2.0 * a for a in nodes], 5, cache_mode="lazy", profile="compact")
    assert torch.equal(cg.cell_metrics(g2), 8.0 * cg.cell_metrics(g1))
    for a in range(3):
        assert torch.equal(cg.face_metrics(g2)[a], 4.0 * cg.face_metrics(g1)[a])


def test_2d_4096_full_size_properties(cg, oracle):
    """BASELINE config 2 (2-D DiscreteGrid 4096^2, GradientCorrected): GCL bound of test_2d.jl:183-238 on the whole
    domain, 256 random cells against the oracle, and the bitwise scaling law under x -> 2x."""
    import torch
    nodes = cg.workloads.wavy_nodes_2d(4097, 4097)
    g = cg.DiscreteGrid(*nodes, 5, conserved_metric_scheme="gradient", cache_mode="lazy")
    cg.refresh_metrics(g)
    assert max(cg.gcl(g)) < 5e-15
    rng = np.random.default_rng(8)
    sample = np.sort(rng.choice(4106 * 4106, size=256, replace=False))
    got = _sampled(cg, g, sample)
    assert_three_way(oracle, got, nodes, 5, "gradient", sample, (4106, 4106), label="4096^2 sampled")
    g2 = cg.DiscreteGrid(*[2.0 * a for a in nodes], 5, conserved_metric_scheme="gradient", cache_mode="lazy")
    f1, f2 = cg.face_metrics(g), cg.face_metrics(g2)
    for a in range(2):
        assert torch.equal(f2[a].conserved, 2.0 * f1[a].conserved)      # J inv(F): one power of the scale in 2-D
        assert torch.equal(f2[a].forward[..., :4], 2.0 * f1[a].forward[..., :4])
        assert torch.equal(f2[a].forward[..., 4], 4.0 * f1[a].forward[..., 4])

"""Legacy MEG6 pipeline (CurvilinearGrid3D :meg6 / :meg6_symmetric) on the GPU against the
array-pass oracle.  Both sides keep the reference's operation order and are built without FMA
contraction, so agreement is expected to the last bit almost everywhere; the gate is 1e-13 relative
(per array, against the array's max) plus the reference's own known answers."""
import numpy as np
import pytest

from util import mild_nodes

pytestmark = pytest.mark.gpu


def _dom(nfull, h=5):
    idx = np.meshgrid(*[np.arange(h, n - h) for n in nfull], indexing="ij")
    return (idx[0] + nfull[0] * (idx[1] + nfull[1] * idx[2])).ravel()


@pytest.mark.parametrize("symmetric", [False, True])
@pytest.mark.parametrize("halo_included", [False, True])
def test_legacy_matches_oracle(cg, oracle, symmetric, halo_included):
    shape = (13, 12, 14) if not halo_included else (19, 18, 20)
    nodes = mild_nodes(shape, seed=41)
    ref = oracle.legacy_meg6_3d(nodes, 5, halo_coords_included=halo_included, symmetric=symmetric)
    mesh = cg.legacy.CurvilinearGrid3D(*nodes, "meg6_symmetric" if symmetric else "meg6",
                                       halo_coords_included=halo_included)
    assert mesh.celldims_full == ref["nfull"]
    cell, edge, cent = mesh._cell.cpu().numpy(), mesh._edge.cpu().numpy(), mesh._cent.cpu().numpy()
    sel = slice(None) if halo_included else _dom(ref["nfull"])
    assert np.array_equal(cent, ref["centroid"])
    for name, got, want in (("cell", cell, ref["cell"]), ("edge", edge, ref["edge"])):
        g, w = got[..., sel], want[..., sel]
        if halo_included and name == "edge":
            continue  # last planes are not written by interpolate_to_edge!; compared below on the valid range
        scale = np.abs(w).max(axis=-1, keepdims=True)
        scale[scale == 0] = 1.0
        err = (np.abs(g - w) / scale).max()
        assert err <= 1e-13, (name, err)
        assert (g == w).mean() > 0.99, (name, (g == w).mean())
    # edges: everything the reference writes (incl. the lo-1 face plane) matches
    w, g = ref["edge"], edge
    scale = np.abs(w).max(axis=-1, keepdims=True)
    scale[scale == 0] = 1.0
    assert (np.abs(g - w) / scale).max() <= 1e-13


def test_legacy_rectilinear_known_answers(cg):
    """test_3d.jl:137-147 / rectilinear_compare_to_standard.jl: forward = (dx,dy,dz), J, hatted = J/dx..."""
    wl = cg.workloads
    x, y, z, dx, dy, dz = wl.rectilinear_nodes_3d(0.0, 2.0, 1.0, 3.0, -1.0, 2.0, 8, 9, 10)
    mesh = cg.legacy.CurvilinearGrid3D(x, y, z, "meg6")
    m = mesh.cell_center_metrics
    h = 5
    sl = (slice(h, -h),) * 3
    J = dx * dy * dz
    assert abs(m["J"][sl].cpu().numpy() - J).max() < 1e-14
    assert abs(m["x₁"]["ξ"][sl].cpu().numpy() - dx).max() < 1e-14
    assert abs(m["x₂"]["η"][sl].cpu().numpy() - dy).max() < 1e-14
    assert abs(m["x₃"]["ζ"][sl].cpu().numpy() - dz).max() < 1e-14
    assert abs(m["x₁"]["η"][sl].cpu().numpy()).max() == 0.0
    assert abs(m["ξ"]["x₁"][sl].cpu().numpy() - 1 / dx).max() < 1e-12
    assert abs(m["ξ̂"]["x₁"][sl].cpu().numpy() - J / dx).max() < 1e-14
    e = mesh.edge_metrics["i₊½"]
    assert abs(e["ξ̂"]["x₁"][sl].cpu().numpy() - J / dx).max() < 1e-14
    assert abs(e["J"][sl].cpu().numpy() - J).max() < 1e-14


def test_legacy_errors_and_static(cg):
    nodes = mild_nodes((12, 12, 12), seed=1)
    with pytest.raises(ValueError):
        cg.legacy.CurvilinearGrid3D(*nodes, "meg9")
    with pytest.raises(ValueError):
        cg.legacy.CurvilinearGrid3D(nodes[0], nodes[1][:-1], nodes[2], "meg6")
    with pytest.raises(cg.CurvigridArgumentError):
        cg.legacy.CurvilinearGrid3D(*mild_nodes((5, 12, 12), seed=1), "meg6")   # < 6 interior cells
    mesh = cg.legacy.CurvilinearGrid3D(*nodes, "meg6", is_static=True)
    with pytest.warns(UserWarning):
        cg.legacy.update(mesh, False, False)
    assert mesh.launches > 0


@pytest.mark.parametrize("symmetric", [False, True])
def test_legacy_spherical_basis_grid(cg, oracle, symmetric):
    """SphericalBasisCurvilinearGrid3D (3d_spherical_basis.jl:191-315): the MEG6 pipeline on the stored (r, θ, ϕ)
    coordinates + Cartesian node coordinates."""
    terms = cg.workloads.spherical_rtp_3d_terms((12, 11, 13))
    nodes = cg.workloads.nodes_from_terms(terms, 0.0, (12, 11, 13), 0)
    ref = oracle.legacy_meg6_3d(nodes, 5, halo_coords_included=False, symmetric=symmetric)
    mesh = cg.legacy.SphericalBasisCurvilinearGrid3D(*nodes, "meg6_symmetric" if symmetric else "meg6")
    sel = _dom(ref["nfull"])
    assert np.array_equal(mesh._cent.cpu().numpy(), ref["centroid"])
    for got, want in ((mesh._cell.cpu().numpy(), ref["cell"]), (mesh._edge.cpu().numpy(), ref["edge"])):
        g, w = got[..., sel], want[..., sel]
        scale = np.abs(w).max(axis=-1, keepdims=True)
        scale[scale == 0] = 1.0
        assert (np.abs(g - w) / scale).max() <= 1e-13
    r, th, ph = (np.asarray(a).transpose(2, 1, 0) for a in nodes)
    c = mesh.cartesian_node_coordinates
    np.testing.assert_allclose(c["x"].cpu().numpy(), r * np.sin(th) * np.cos(ph), rtol=1e-14, atol=1e-15)
    np.testing.assert_allclose(c["y"].cpu().numpy(), r * np.sin(th) * np.sin(ph), rtol=1e-14, atol=1e-15)
    np.testing.assert_allclose(c["z"].cpu().numpy(), r * np.cos(th), rtol=1e-14, atol=1e-15)
    assert set(mesh.centroid_coordinates) == {"r", "θ", "ϕ"}
    # the centroid radius is the eight-node mean
    assert abs(float(mesh.centroid_coordinates["r"][5, 5, 5]) - r[0:2, 0:2, 0:2].mean()) < 1e-14
    # moving mesh: new nodes -> update refreshes metrics and Cartesian coordinates
    cg.legacy.set_node_coordinates(mesh, nodes[0] * 1.5, nodes[1], nodes[2])
    cg.legacy.update(mesh, True, False)
    np.testing.assert_allclose(mesh.cartesian_node_coordinates["z"].cpu().numpy(), 1.5 * r * np.cos(th), rtol=1e-14, atol=1e-15)


@pytest.mark.parametrize("halo_included", [False, True])
def test_legacy_2d_matches_oracle(cg, oracle, halo_included):
    """CurvilinearGrid2D :meg6 (2d.jl:616-746) against the array-pass oracle."""
    n = (23, 19) if halo_included else (14, 12)
    x, y = cg.workloads.wavy_nodes_2d(*n)
    ref = oracle.legacy_meg6_2d((x, y), 5, halo_coords_included=halo_included)
    mesh = cg.legacy.CurvilinearGrid2D(x, y, "meg6", halo_coords_included=halo_included)
    assert mesh.celldims_full == ref["nfull"]
    assert np.array_equal(mesh._cent.cpu().numpy(), ref["centroid"])
    for name, got, want in (("cell", mesh._cell.cpu().numpy(), ref["cell"]), ("edge", mesh._edge.cpu().numpy(), ref["edge"])):
        scale = np.abs(want).max(axis=-1, keepdims=True)
        scale[scale == 0] = 1.0
        err = (np.abs(got - want) / scale).max()
        assert err <= 1e-13, (name, err)
        assert (got == want).mean() > 0.99, (name, (got == want).mean())
    sym = cg.legacy.CurvilinearGrid2D(x, y, "meg6_symmetric", halo_coords_included=halo_included)
    assert np.array_equal(sym._cell.cpu().numpy(), mesh._cell.cpu().numpy())   # conservative_metrics.jl:417-421


def test_legacy_2d_known_answers_and_errors(cg):
    """test_discretizations.jl:148-191 style: x = i, y = 2 j -> x_xi = 1, y_eta = 2, xi_x = 1, eta_y = 1/2, J = 2."""
    ni, nj = 12, 10
    x = np.arange(ni, dtype=float)[:, None] * np.ones((1, nj))
    y = 2.0 * np.arange(nj, dtype=float)[None, :] * np.ones((ni, 1))
    mesh = cg.legacy.CurvilinearGrid2D(x, y, "meg6")
    m, h = mesh.cell_center_metrics, 5
    sl = (slice(h, -h),) * 2
    assert abs(m["x₁"]["ξ"][sl].cpu().numpy() - 1).max() < 1e-13
    assert abs(m["x₂"]["η"][sl].cpu().numpy() - 2).max() < 1e-13
    assert abs(m["x₁"]["η"][sl].cpu().numpy()).max() == 0.0
    assert abs(m["J"][sl].cpu().numpy() - 2).max() < 1e-13
    assert abs(m["ξ"]["x₁"][sl].cpu().numpy() - 1).max() < 1e-13
    assert abs(m["η"]["x₂"][sl].cpu().numpy() - 0.5).max() < 1e-13
    assert abs(m["ξ̂"]["x₁"][sl].cpu().numpy() - 2).max() < 1e-13
    assert abs(m["η̂"]["x₂"][sl].cpu().numpy() - 1).max() < 1e-13
    e = mesh.edge_metrics["j₊½"]
    assert abs(e["η̂"]["x₂"][sl].cpu().numpy() - 1).max() < 1e-13
    assert abs(e["J"][sl].cpu().numpy() - 2).max() < 1e-13
    with pytest.raises(ValueError):
        cg.legacy.CurvilinearGrid2D(x, y[:-1], "meg6")
    with pytest.raises(cg.CurvigridArgumentError):
        cg.legacy.CurvilinearGrid2D(x[:5], y[:5], "meg6")   # < 6 interior cells
    static = cg.legacy.CurvilinearGrid2D(x, y, "meg6", is_static=True)
    with pytest.raises(RuntimeError):
        cg.legacy.update2d(static, False, False)


@pytest.mark.parametrize("axis", ["x", "y"])
def test_legacy_axisymmetric_2d_snaps_after_centroids(cg, oracle, axis):
    """AxisymmetricGrid2D update! (2d.jl:644-671): centroids from the nodes as given, then the axis row / column of
    node.domain is snapped, then the metrics; a second update sees the snapped nodes."""
    x, y = cg.workloads.wavy_nodes_2d(15, 13)
    x, y = x + 0.01, y + 0.02                      # first row / column slightly off the axis
    mesh = cg.legacy.AxisymmetricGrid2D(x, y, True, axis, "meg6")
    ref = oracle.legacy_meg6_2d((x, y), 5)
    assert np.array_equal(mesh._cell.cpu().numpy(), cg.legacy.CurvilinearGrid2D(x, y, "meg6")._cell.cpu().numpy())
    scale = np.abs(ref["cell"]).max(axis=-1, keepdims=True)
    scale[scale == 0] = 1.0
    assert (np.abs(mesh._cell.cpu().numpy() - ref["cell"]) / scale).max() <= 1e-13
    n = mesh.node_coordinates
    xs, ys = x.copy(), y.copy()
    if axis == "x":
        assert float(n["y"][0].abs().max()) == 0.0 and float(n["x"][0].abs().max()) > 0.0
        ys[:, 0] = 0.0
    else:
        assert float(n["x"][:, 0].abs().max()) == 0.0
        xs[0, :] = 0.0
    cg.legacy.update2d(mesh, True, False)          # now the metrics of the snapped mesh
    ref2 = oracle.legacy_meg6_2d((xs, ys), 5)
    scale = np.abs(ref2["cell"]).max(axis=-1, keepdims=True)
    scale[scale == 0] = 1.0
    assert (np.abs(mesh._cell.cpu().numpy() - ref2["cell"]) / scale).max() <= 1e-13
    with pytest.raises(ValueError):
        cg.legacy.AxisymmetricGrid2D(x, y, True, "z", "meg6")


@pytest.mark.parametrize("halo_included", [False, True])
def test_legacy_1d_matches_oracle(cg, oracle, halo_included):
    """CurvilinearGrid1D(x, :meg6) update! (1d.jl:388-408) against the oracle's array-pass restatement."""
    n = 41
    s = np.linspace(0.0, 1.0, n + (10 if halo_included else 0))
    x = s + 0.04 * np.sin(2 * np.pi * s) + 0.3
    mesh = cg.legacy.CurvilinearGrid1D(x, "meg6", halo_coords_included=halo_included)
    ref = oracle.legacy_meg6_1d(x, 5, halo_included)
    for got, want in ((mesh._cell.cpu().numpy(), ref["cell"]), (mesh._edge[0].cpu().numpy(), ref["edge"]),
                      (mesh._cent.cpu().numpy(), ref["centroid"])):
        scale = max(1e-300, np.abs(want).max())
        assert np.abs(got - want).max() <= 1e-13 * scale
    # known answers on a uniform mesh: x_xi = dx, xi_x = 1 / dx, J = dx, hatted = 1 (conservative_metrics.jl:239-245)
    u = cg.legacy.CurvilinearGrid1D(np.linspace(2.0, 4.0, 21), "meg6")
    d = u.cell_center_metrics
    dom = slice(5, -5)
    assert np.allclose(d["x₁"]["ξ"][dom].cpu().numpy(), 0.1, rtol=0, atol=1e-13)
    assert np.allclose(d["ξ̂"]["x₁"][dom].cpu().numpy(), 1.0, rtol=0, atol=1e-13)
    assert np.allclose(u.edge_metrics["i₊½"]["J"][6:-6].cpu().numpy(), 0.1, rtol=0, atol=1e-13)
    with pytest.raises(cg.CurvigridArgumentError):
        cg.legacy.CurvilinearGrid1D(np.linspace(0, 1, 5), "meg6")    # < 6 interior cells


def test_legacy_check_valid_metrics(cg):
    """_check_valid_metrics (3d.jl:600-677, 1d.jl:457-482): a folded mesh (negative Jacobian) or a non-finite metric is
    reported with the reference's messages; the check is one fused device reduction (cg_legacy_check_valid)."""
    x, y, z = cg.workloads.wavy_nodes_3d(15, 14, 13)
    ok = cg.legacy.CurvilinearGrid3D(x, y, z, "meg6")           # passes the check inside update
    assert float(ok.cell_center_metrics["J"][5:-5, 5:-5, 5:-5].min()) > 0
    with pytest.raises(RuntimeError, match="Invalid centroid metrics found"):
        cg.legacy.CurvilinearGrid3D(-x, y, z, "meg6")           # mirrored mesh: J < 0 everywhere
    bad = np.array(x)
    bad[7, 7, 7] = np.nan
    with pytest.raises(RuntimeError, match="Invalid (edge|centroid) metrics found"):
        cg.legacy.CurvilinearGrid3D(bad, y, z, "meg6")
    b1 = np.linspace(0.0, 1.0, 30)
    b1[12] = np.inf
    with pytest.raises(RuntimeError, match="Invalid (edge|centroid) metrics found"):
        cg.legacy.CurvilinearGrid1D(b1, "meg6")


def test_legacy_1d_decreasing_mesh_is_accepted(cg):
    m = cg.legacy.CurvilinearGrid1D(np.linspace(1.0, 0.0, 30), "meg6")
    assert float(m.cell_center_metrics["J"][5:-5].min()) > 0 and float(m.cell_center_metrics["x₁"]["ξ"][5:-5].max()) < 0


@pytest.mark.parametrize("symmetric", [False, True])
@pytest.mark.parametrize("halo_included", [False, True])
@pytest.mark.parametrize("shape", [(131, 58, 63), (122, 62, 115)])
def test_legacy_fused_equals_pass_by_pass_bitwise(cg, monkeypatch, shape, symmetric, halo_included):
    """The fused launches (first + second derivative + MEG reconstruction / edge interpolation through shared memory,
    shell-only clearing) reproduce the pass-by-pass pipeline (CG_LEGACY_FUSED=0) bit for bit, on grids that need
    several tiles per axis and end in short tiles (57 = 54 + 3 and 121 = 118 + 3 cells: the one-sided stencils of the last
    three cells reach back into the tile before; 114 = 2 x 54 + 6), with the TMA brick load (even row length) and
    without it, from poisoned output buffers."""
    import torch
    nodes = mild_nodes(shape, seed=7)
    scheme = "meg6_symmetric" if symmetric else "meg6"
    mesh = cg.legacy.CurvilinearGrid3D(*nodes, scheme, halo_coords_included=halo_included)
    assert mesh.launches <= 40   # 20 with the tile kernels, 36 with CG_LEGACY_MARCH=1 (two end strips per marched pass)
    fused = [t.clone() for t in (mesh._cell, mesh._edge, mesh._cent)]
    for t in (mesh._cell, mesh._edge, mesh._cent, mesh._scratch):
        t.fill_(float("nan"))   # nothing may survive from the caller's buffers
    cg.legacy.update(mesh, True, halo_included)
    for a, b in zip(fused, (mesh._cell, mesh._edge, mesh._cent)):
        assert torch.equal(a, b)
    monkeypatch.setenv("CG_LEGACY_FUSED", "0")
    cg.legacy.update(mesh, True, halo_included)
    assert mesh.launches > 200
    for name, a, b in zip(("cell", "edge", "centroid"), fused, (mesh._cell, mesh._edge, mesh._cent)):
        assert torch.equal(a, b), name


@pytest.mark.parametrize("halo_included", [False, True])
def test_legacy_fused_2d_1d_equal_pass_by_pass_bitwise(cg, monkeypatch, halo_included):
    import torch
    x, y = cg.workloads.wavy_nodes_2d(150, 131)
    m2 = cg.legacy.CurvilinearGrid2D(x, y, "meg6", halo_coords_included=halo_included)
    s = np.linspace(0.0, 1.0, 400)
    m1 = cg.legacy.CurvilinearGrid1D(s + 0.04 * np.sin(2 * np.pi * s) + 0.3, "meg6", halo_coords_included=halo_included)
    fused = [t.clone() for m in (m2, m1) for t in (m._cell, m._edge, m._cent)]
    monkeypatch.setenv("CG_LEGACY_FUSED", "0")
    cg.legacy.update2d(m2, True, halo_included)
    cg.legacy.update1d(m1, True, halo_included)
    for a, b in zip(fused, [t for m in (m2, m1) for t in (m._cell, m._edge, m._cent)]):
        assert torch.equal(a, b)


def test_legacy_fused_matches_oracle_multi_tile(cg, oracle):
    """Oracle parity of the fused path on a grid with several tiles per axis (the small cases above fit one tile)."""
    nodes = mild_nodes((125, 64, 70), seed=5)
    ref = oracle.legacy_meg6_3d(nodes, 5, halo_coords_included=False, symmetric=False)
    mesh = cg.legacy.CurvilinearGrid3D(*nodes, "meg6")
    sel = _dom(ref["nfull"])
    for name, got, want in (("cell", mesh._cell.cpu().numpy(), ref["cell"]), ("edge", mesh._edge.cpu().numpy(), ref["edge"])):
        scale = np.abs(want).max(axis=-1, keepdims=True)
        scale[scale == 0] = 1.0
        assert (np.abs(got - want) / scale).max() <= 1e-13, name
        assert (got[..., sel] == want[..., sel]).mean() > 0.99, name

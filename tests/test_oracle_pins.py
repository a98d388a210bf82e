"""Pins of the oracle: the reference's own known-answer and property tests re-encoded
(SURVEY.md §4, §8c).  CPU only.  The reference stores no golden vectors and no Julia runtime
exists here, so these properties ARE the pin ("parity partially pinned")."""
import numpy as np
import pytest

import curvigrid_b200 as cg
from util import domain_indices

wl = cg.workloads
EPS = np.finfo(np.float64).eps
SCHEMES = ("endpoint", "gradient", "curvature")


def _approx(a, b):  # Julia's isapprox default: rtol = sqrt(eps)
    return np.all(np.abs(a - b) <= np.sqrt(EPS) * np.maximum(np.abs(a), np.abs(b)) + 1e-300) or np.allclose(a, b, atol=1e-12)


@pytest.mark.parametrize("halo_included", [False, True])
def test_rectilinear_3d_known_answers(oracle, halo_included):
    """test_3d.jl:40-135: every entry on domain / i+1/2 domains; with halo-included input on cell.full."""
    h = 5
    ni, nj, nk = (8, 10, 12) if not halo_included else (18, 20, 22)  # the halo variant strips 2h nodes
    x, y, z, dx, dy, dz = wl.rectilinear_nodes_3d(0.0, 2.0, 1, 3, -1, 2, ni, nj, nk)
    r = oracle.discrete_eval([x, y, z], h, "curvature", halo_coords_included=halo_included)
    J = dx * dy * dz
    fwd = np.array([dx, 0, 0, 0, dy, 0, 0, 0, dz, J])
    inv = np.array([1 / dx, 0, 0, 0, 1 / dy, 0, 0, 0, 1 / dz, J])
    cons = np.array([J / dx, 0, 0, 0, J / dy, 0, 0, 0, J / dz])
    # halo-included input: asserted on cell.full (test_3d.jl:124-135); otherwise on cell.domain
    assert np.abs(r["cell_fwd"] - fwd).max() < 1e-13
    assert np.abs(r["cell_inv"] - inv).max() < 1e-11
    assert np.abs(r["face_fwd"] - fwd).max() < 1e-13
    assert np.abs(r["face_inv"] - inv).max() < 1e-11
    assert np.abs(r["face_cons"] - cons).max() < 1e-13
    assert oracle.gcl_max(r["face_cons"], r["nfull"], h).max() < EPS


def test_rectilinear_2d_known_answers(oracle):
    """test_2d.jl:18-63 and the halo-included variant :124-151 (cell.full 40x80, domain 6:35 x 6:75)."""
    x, y, dx, dy = wl.rectilinear_nodes_2d(0, 2, 1, 3, 40, 80)
    r = oracle.discrete_eval([x, y], 5, "curvature", halo_coords_included=True)
    assert r["nfull"] == (40, 80)
    J = dx * dy
    assert np.abs(r["cell_fwd"] - np.array([dx, 0, 0, dy, J])).max() < 1e-14
    assert np.abs(r["cell_inv"] - np.array([1 / dx, 0, 0, 1 / dy, J])).max() < 1e-12
    assert np.abs(r["face_fwd"] - np.array([dx, 0, 0, dy, J])).max() < 1e-13
    assert np.abs(r["face_inv"] - np.array([1 / dx, 0, 0, 1 / dy, J])).max() < 1e-12
    assert np.abs(r["face_cons"] - np.array([dy, 0, 0, dx])).max() < 1e-14
    # coord / centroid of the halo-included grid (test_2d.jl:139-142)
    nodes = oracle.discrete_coords([x, y], 5, 0, halo_coords_included=True)
    cents = oracle.discrete_coords([x, y], 5, 1, halo_coords_included=True)
    ilo = 5
    node_lin = (ilo + 1) + 41 * (ilo + 1)
    assert np.allclose([nodes[0][node_lin], nodes[1][node_lin]], [0.3, 1.15], atol=1e-14)
    cell_lin = ilo + 40 * ilo
    assert np.allclose([cents[0][cell_lin], cents[1][cell_lin]], [0.275, 1.1375], atol=1e-14)


@pytest.mark.parametrize("scheme", ("gradient", "curvature"))
@pytest.mark.parametrize("halo_included", [False, True])
def test_wavy_2d_gcl(oracle, scheme, halo_included):
    """test_2d.jl:153-238: GCL < 5e-15 on the 41x41 wavy mesh (the reference asserts it for the default
    CurvatureCorrected scheme; Gradient has the same conserved 2-D metrics; EndpointAverage is not
    GCL-exact in 2-D on a general mesh and the reference does not claim it)."""
    x, y = wl.wavy_nodes_2d(41, 41)
    r = oracle.discrete_eval([x, y], 5, scheme, what=2, halo_coords_included=halo_included)
    assert oracle.gcl_max(r["face_cons"], r["nfull"], 5).max() < 5e-15


@pytest.mark.parametrize("scheme,n", [("endpoint", 20), ("gradient", 20), ("curvature", 14)])
def test_wavy_3d_gcl(oracle, scheme, n):
    """test_3d.jl:149-245: GCL < 5e-13 on the wavy mesh (curvature at 14^3 nodes to bound CPU time;
    the 20^3 curvature case runs on the GPU in test_gpu_parity.py)."""
    x, y, z = wl.wavy_nodes_3d(n, n, n)
    r = oracle.discrete_eval([x, y, z], 5, scheme, what=2)
    assert oracle.gcl_max(r["face_cons"], r["nfull"], 5).max() < 5e-13


def test_mapped_gcl(oracle):
    """test_2d_continuous.jl:68-92 (wavy 2-D, < 1e-14, also after update! with new t),
    test_3d_continuous.jl:9-62 (wavy 3-D, < 1e-14), test_3d.jl:247-282 (sphere sector, < 5e-13)."""
    for t in (0.0, 0.7):
        r = oracle.mapped_eval(wl.wavy_2d_terms((101, 101), ct=0.1), t, (101, 101), 5, "curvature", what=2)
        assert oracle.gcl_max(r["face_cons"], r["nfull"], 5).max() < 1e-14
    r = oracle.mapped_eval(wl.wavy_3d_terms((12, 12, 12)), 0.0, (12, 12, 12), 3, "curvature", what=2)
    assert oracle.gcl_max(r["face_cons"], r["nfull"], 3).max() < 1e-14
    r = oracle.mapped_eval(wl.sphere_sector_3d_terms((10, 10, 10)), 0.0, (10, 10, 10), 2, "curvature", what=2)
    assert oracle.gcl_max(r["face_cons"], r["nfull"], 2).max() < 5e-13


def test_schemes_differ_in_3d_but_conserved_gradient_equals_curvature_in_2d(oracle):
    """test_unified_grid_types.jl:139-198 (three schemes give different face values); SURVEY A.2
    (2-D: Ghat of Gradient and Curvature coincide, G differs)."""
    x, y, z = wl.wavy_nodes_3d(8, 8, 8)
    rs = {s: oracle.discrete_eval([x, y, z], 2, s, what=2) for s in SCHEMES}
    assert np.abs(rs["endpoint"]["face_cons"] - rs["gradient"]["face_cons"]).max() > 1e-6
    assert np.abs(rs["gradient"]["face_cons"] - rs["curvature"]["face_cons"]).max() > 1e-8
    x2, y2 = wl.wavy_nodes_2d(17, 17)
    r2 = {s: oracle.discrete_eval([x2, y2], 2, s, what=2) for s in SCHEMES}
    assert np.abs(r2["gradient"]["face_cons"] - r2["curvature"]["face_cons"]).max() < 1e-15
    assert np.abs(r2["gradient"]["face_inv"] - r2["curvature"]["face_inv"]).max() > 1e-10


def _active_rows(r, N):
    return [np.stack([r["face_cons"][a][:, a + N * c] for c in range(N)], axis=1) for a in range(N)]


@pytest.mark.parametrize("dim", [2, 3])
def test_mapped_vs_discrete_convergence(oracle, dim):
    """test_mapped_discrete_face_geometry.jl:269-337: a DiscreteGrid built from the MappedGrid's
    node.full coordinates (halo_coords_included) has GCL < 1e-12 and its face metric vectors converge
    to the mapped ones with observed order > 1."""
    h = 2
    res = [(8,) * dim, (12,) * dim, (16,) * dim] if dim == 2 else [(6,) * 3, (9,) * 3, (12,) * 3]
    errs = []
    for cd in res:
        terms = wl.generic_3d_terms(cd) if dim == 3 else wl.wavy_2d_terms(cd, L=2.0)
        m = oracle.mapped_eval(terms, 0.0, cd, h, "curvature", what=2)
        nodes = wl.nodes_from_terms(terms, 0.0, cd, h, include_halo=True)
        d = oracle.discrete_eval(nodes, h, "curvature", what=2, halo_coords_included=True)
        assert oracle.gcl_max(m["face_cons"], m["nfull"], h).max() < 1e-12
        assert oracle.gcl_max(d["face_cons"], d["nfull"], h).max() < 1e-12
        dom = domain_indices(m["nfull"], h)
        am, ad = _active_rows(m, dim), _active_rows(d, dim)
        # relative to the face-metric magnitude so different resolutions compare (the reference compares
        # physical metric vectors, which carry the same factor on both grids)
        e = max(np.abs(ad[a][dom] - am[a][dom]).max() / np.abs(am[a][dom]).max() for a in range(dim))
        errs.append(e)
    if max(errs) < 1e-12:
        return
    assert errs[1] < errs[0] and errs[2] < errs[1] and errs[2] < 0.35 * errs[0]
    orders = [np.log(errs[i] / errs[i + 1]) / np.log(res[i + 1][0] / res[i][0]) for i in range(2)]
    assert min(orders) > 1.0


def test_cell_inverse_stores_det_of_forward(oracle):
    """metrics.jl:1875-1876: inverse[I] = Metric(G, J) with J = det(F), not det(G)."""
    x, y, z = wl.wavy_nodes_3d(6, 6, 6)
    r = oracle.discrete_eval([x, y, z], 2, "curvature", what=1)
    assert np.array_equal(r["cell_fwd"][:, 9], r["cell_inv"][:, 9])
    F = r["cell_fwd"][:, :9].reshape(-1, 3, 3).transpose(0, 2, 1)
    assert np.allclose(np.linalg.det(F), r["cell_fwd"][:, 9], rtol=1e-12)


def test_golden_vectors_unchanged(oracle):
    """tests/golden/*.npz were written by tests/golden/make_golden.py from this oracle."""
    import os
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    for name in ("discrete3d", "discrete2d", "discrete1d"):
        z = np.load(os.path.join(here, f"{name}.npz"))
        nodes = [z[f"node{q}"] for q in range(sum(1 for k in z.files if k.startswith("node")))]
        for sch in SCHEMES:
            r = oracle.discrete_eval(nodes, int(z["nhalo"]), sch)
            for k in ("cell_fwd", "cell_inv", "face_fwd", "face_inv", "face_cons"):
                assert np.allclose(r[k], z[f"{sch}_{k}"], rtol=1e-13, atol=1e-13), (name, sch, k)
    for name in ("mapped2d_readme", "mapped3d_wavy"):
        z = np.load(os.path.join(here, f"{name}.npz"))
        for sch in SCHEMES:
            r = oracle.mapped_eval(z["terms"], float(z["t"]), tuple(int(c) for c in z["celldims"]), int(z["nhalo"]), sch)
            for k in ("cell_fwd", "cell_inv", "face_fwd", "face_inv", "face_cons"):
                assert np.allclose(r[k], z[f"{sch}_{k}"], rtol=1e-13, atol=1e-13), (name, sch, k)


# ---------------------------------------------------------------------------
# legacy MEG6 pipeline
# ---------------------------------------------------------------------------
def test_legacy_stencil_floating_point_error(oracle):
    """test_discretizations.jl:193-221: Kahan sum + tol_diff zero the derivative of 1e4 + O(100 eps) noise."""
    import ctypes
    lib = oracle.lib()
    lib.cgo_legacy_central6.restype = ctypes.c_double
    rng = np.random.default_rng(0)
    for _ in range(20):
        v = 1e4 * np.ones(6) + 100 * EPS * rng.random(6)
        naive = -1 / 60 * v[0] + 3 / 20 * v[1] - 3 / 4 * v[2] + 3 / 4 * v[3] - 3 / 20 * v[4] + 1 / 60 * v[5]
        d = lib.cgo_legacy_central6(v.ctypes.data_as(ctypes.POINTER(ctypes.c_double)))
        assert d == 0.0
    assert abs(naive) >= 0.0


def test_legacy_stencils_exact_on_polynomials(oracle):
    """test_discretizations.jl:41-146 (derivatives of linear fields) + order of the 6-point one-sided stencils:
    exact for polynomials up to degree 5."""
    import ctypes
    lib = oracle.lib()
    lib.cgo_legacy_central6.restype = ctypes.c_double
    dp = lambda a: a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))
    xs = np.arange(6, dtype=np.float64)
    for deg in range(1, 6):
        f = 0.3 * xs ** deg + 1.7 * xs + 2.0
        df = 0.3 * deg * xs ** (deg - 1) + 1.7
        lo, hi = np.zeros(3), np.zeros(3)
        lib.cgo_legacy_mixed(dp(f), dp(lo), dp(hi))
        assert np.allclose(lo, df[:3], rtol=1e-11, atol=1e-11), (deg, lo, df[:3])
        assert np.allclose(hi, df[3:], rtol=1e-11, atol=1e-11), (deg, hi, df[3:])
    xs7 = np.arange(-3, 4, dtype=np.float64)
    for deg in range(1, 7):
        f = 0.2 * xs7 ** deg + 0.9 * xs7
        w = np.ascontiguousarray(np.delete(f, 3))
        d = lib.cgo_legacy_central6(dp(w))
        assert abs(d - (0.9 if deg > 1 else 1.1)) < 1e-11


def test_legacy_rectilinear_and_2d_style_known_answers(oracle):
    """test_3d.jl:137-147 and test_discretizations.jl:148-191 (x=i, y=2j -> x_xi=1, y_eta=2, xi_x=1, eta_y=1/2),
    here on the 3-D pipeline with z = 3k."""
    n = (17, 18, 19)
    i, j, k = np.meshgrid(*[np.arange(1, m + 1, dtype=np.float64) for m in n], indexing="ij")
    r = oracle.legacy_meg6_3d([i, 2 * j, 3 * k], 5, halo_coords_included=True)
    nf = r["nfull"]
    c = r["cell"]
    idx = np.meshgrid(*[np.arange(5, m - 5) for m in nf], indexing="ij")
    dom = (idx[0] + nf[0] * (idx[1] + nf[1] * idx[2])).ravel()
    assert np.allclose(c[1][dom], 1.0) and np.allclose(c[5][dom], 2.0) and np.allclose(c[9][dom], 3.0)
    assert np.allclose(c[10][dom], 1.0) and np.allclose(c[14][dom], 0.5) and np.allclose(c[18][dom], 1 / 3)
    assert np.allclose(c[0][dom], 6.0)
    assert np.abs(c[[2, 3, 4, 6, 7, 8]][:, dom]).max() == 0.0
    # halo-included: metrics are evaluated on cell.full with one-sided stencils at the array edges
    assert np.allclose(c[1], 1.0) and np.allclose(c[0], 6.0)


@pytest.mark.parametrize("symmetric", [False, True])
def test_legacy_wavy_gcl_small(oracle, symmetric):
    """The legacy conservative metrics are GCL-consistent to ~1e-14 on the wavy mesh
    (test_2d_axisymmetric.jl / test_3d_spherical.jl assert eps-level GCL for :meg6_symmetric)."""
    x, y, z = wl.wavy_nodes_3d(16, 16, 16)
    r = oracle.legacy_meg6_3d([x, y, z], 5, symmetric=symmetric)
    nf = r["nfull"]
    E = r["edge"].reshape(3, 19, nf[2], nf[1], nf[0])
    h = 5
    worst = 0.0
    for c in range(3):
        tot = 0.0
        for ax in range(3):
            A = E[ax, 10 + 3 * ax + c]
            hi = [slice(h, nf[2] - h), slice(h, nf[1] - h), slice(h, nf[0] - h)]
            lo = list(hi)
            d = 2 - ax
            lo[d] = slice(h - 1, (nf[2], nf[1], nf[0])[d] - h - 1)
            tot = tot + A[tuple(hi)] - A[tuple(lo)]
        worst = max(worst, np.abs(tot).max())
    assert worst < 5e-14


def test_legacy_2d_known_answers(oracle):
    """CurvilinearGrid2D legacy pipeline: x = i, y = 2 j (test_discretizations.jl:148-191) and a rotated / sheared
    affine map (every stencil is exact on linear fields): forward metrics, J, inverse and hatted metrics, edges."""
    ni, nj, h = 12, 11, 5
    i = np.arange(ni, dtype=float)[:, None] * np.ones((1, nj))
    j = np.arange(nj, dtype=float)[None, :] * np.ones((ni, 1))
    A = np.array([[1.3, 0.4], [-0.2, 0.9]])
    x, y = A[0, 0] * i + A[0, 1] * j + 0.7, A[1, 0] * i + A[1, 1] * j - 1.1
    r = oracle.legacy_meg6_2d((x, y), h)
    nf = r["nfull"]
    dom = (np.arange(h, nf[0] - h)[:, None] + nf[0] * np.arange(h, nf[1] - h)[None, :]).ravel()
    c = r["cell"][:, dom]
    det = np.linalg.det(A)
    Ainv = np.linalg.inv(A)
    assert np.abs(c[0] - det).max() < 1e-13
    for q in range(2):
        for a in range(2):
            assert np.abs(c[1 + 2 * q + a] - A[q, a]).max() < 1e-13
    for row in range(2):
        for k in range(2):
            assert np.abs(c[5 + 2 * row + k] - Ainv[row, k]).max() < 1e-13
            assert np.abs(c[9 + 2 * row + k] - Ainv[row, k] * det).max() < 1e-13
    for ax in range(2):
        e = r["edge"][ax][:, dom]
        # the one-sided 6-point stencils next to the domain edge amplify round-off of the (constant) input: 1e-12
        assert np.abs(e[0] - det).max() < 1e-12
        for m in range(4):
            assert np.abs(e[1 + m] - Ainv.ravel()[m]).max() < 1e-12
            assert np.abs(e[5 + m] - Ainv.ravel()[m] * det).max() < 1e-12
    # centroids are the four-node mean (2d.jl:734-746)
    cx = r["centroid"][0][dom].reshape(nf[1] - 2 * h, nf[0] - 2 * h)
    assert abs(cx[0, 0] - 0.25 * (x[0, 0] + x[1, 0] + x[1, 1] + x[0, 1])) < 1e-15

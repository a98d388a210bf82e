"""MappedGrid (separable term-list mapping) on the GPU against the oracle's nested-AD
evaluation of the same analytic closures.  Tolerance 1e-12 relative (per-matrix scale)."""
import numpy as np
import pytest

from util import RTOL_F64, assert_fp32_close, assert_metrics_close, domain_indices, grid_outputs

pytestmark = pytest.mark.gpu
SCHEMES = ("endpoint", "gradient", "curvature")


def _cases(wl):
    return {
        "readme2d": (wl.readme_wavy_2d_terms(), (16, 12), 3),            # BASELINE config 1 mapping
        "wavy2d": (wl.wavy_2d_terms((21, 17), ct=0.3), (21, 17), 2),
        "cyl2d": (wl.cylindrical_2d_terms((12, 10)), (12, 10), 2),
        "wavy3d": (wl.wavy_3d_terms((9, 8, 7), time_amp=0.1), (9, 8, 7), 2),
        "sphere3d": (wl.sphere_sector_3d_terms((8, 7, 6)), (8, 7, 6), 2),
        "rtp3d": (wl.spherical_rtp_3d_terms((8, 8, 8)), (8, 8, 8), 2),
        "affine1d": ([wl.term(0, 1.0, {0: (wl.AFFINE, 0.3, 0.25)}), wl.term(0, 0.02, {0: (wl.SIN, 0.0, 0.7)})], (11,), 2),
    }


@pytest.mark.parametrize("scheme", SCHEMES)
@pytest.mark.parametrize("name", ["readme2d", "wavy2d", "cyl2d", "wavy3d", "sphere3d", "rtp3d", "affine1d"])
def test_mapped_matches_oracle(cg, oracle, name, scheme):
    terms, celldims, h = _cases(cg.workloads)[name]
    t = 0.37
    ref = oracle.mapped_eval(terms, t, celldims, h, scheme)
    g = cg.MappedGrid(terms, {}, celldims, h, conserved_metric_scheme=scheme, t=t)
    got = grid_outputs(cg, g)
    assert got["nfull"] == ref["nfull"]
    assert_metrics_close(got, ref, rtol=RTOL_F64, label=f"{name} {scheme}")


@pytest.mark.parametrize("name", ["readme2d", "wavy3d", "affine1d"])
def test_mapped_coordinates(cg, oracle, name):
    terms, celldims, h = _cases(cg.workloads)[name]
    N = len(celldims)
    g = cg.MappedGrid(terms, {}, celldims, h, t=0.2)
    for which, arrs in ((0, g.node_coordinates), (1, g.centroid_coordinates)):
        ref = oracle.mapped_coords(terms, 0.2, celldims, h, which)
        for q in range(N):
            assert np.abs(arrs[q].reshape(-1).cpu().numpy() - ref[q]).max() < 1e-12
    for a in range(N):
        ref = oracle.mapped_coords(terms, 0.2, celldims, h, 2 + a)
        for q in range(N):
            assert np.abs(g.face_coordinates[a][q].reshape(-1).cpu().numpy() - ref[q]).max() < 1e-12


def test_update_reevaluates_time_dependent_mapping(cg, oracle):
    """update!(grid, t, params) then lazy refresh (mapped_grid.jl:551-593; BASELINE config 4)."""
    wl = cg.workloads
    celldims, h = (8, 7, 6), 2
    terms = wl.wavy_3d_terms(celldims, time_amp=0.1)
    g = cg.MappedGrid(terms, {}, celldims, h)
    for t in (0.0, 0.5, 1.3):
        cg.update(g, t)
        assert not cg.cell_metrics_valid(g) and not cg.face_metrics_valid(g)
        got = grid_outputs(cg, g)
        ref = oracle.mapped_eval(terms, t, celldims, h, "curvature")
        assert_metrics_close(got, ref, label=f"t={t}")
        assert cg.cell_metrics_valid(g) and cg.face_metrics_valid(g)


def test_local_grid_equals_slices_of_global(cg):
    """test_unified_grid_types.jl:413-488 — the pin for partitioned analytic grids."""
    wl = cg.workloads
    celldims, h = (12, 10, 9), 2
    terms = wl.wavy_3d_terms(celldims)
    G = cg.MappedGrid(terms, {}, celldims, h)
    dom = ((3, 9), (2, 7), (4, 9))
    Lg = cg.local_grid(G, dom)
    go, lo = grid_outputs(cg, G), grid_outputs(cg, Lg)
    nfg, nfl = go["nfull"], lo["nfull"]
    # local full cell (i,j,k) <-> global full cell (i+lo_i, ...)
    idx = np.meshgrid(*[np.arange(n) for n in nfl], indexing="ij")
    lin_l = (idx[0] + nfl[0] * (idx[1] + nfl[1] * idx[2])).ravel()
    lin_g = ((idx[0] + dom[0][0]) + nfg[0] * ((idx[1] + dom[1][0]) + nfg[1] * (idx[2] + dom[2][0]))).ravel()
    sub = {k: go[k][..., lin_g, :] for k in ("cell_fwd", "cell_inv", "face_fwd", "face_inv", "face_cons")}
    loc = {k: lo[k][..., lin_l, :] for k in sub}
    assert_metrics_close(loc, sub, rtol=1e-12, label="local_grid")


@pytest.mark.parametrize("scheme", SCHEMES)
def test_mapped_gcl_bounds(cg, scheme):
    """test_3d_continuous.jl:55-62 (wavy 41^3, < 1e-14) and test_3d.jl:247-282 (sphere sector, < 5e-13)."""
    wl = cg.workloads
    g = cg.MappedGrid(wl.wavy_3d_terms((41, 41, 41)), {}, (41, 41, 41), 5, conserved_metric_scheme=scheme)
    assert max(cg.gcl(g)) < 1e-14
    s = cg.MappedGrid(wl.sphere_sector_3d_terms((20, 20, 20)), {}, (20, 20, 20), 5, conserved_metric_scheme=scheme)
    assert max(cg.gcl(s)) < 5e-13
    w2 = cg.MappedGrid(wl.wavy_2d_terms((401, 401)), {}, (401, 401), 5, conserved_metric_scheme=scheme)
    assert max(cg.gcl(w2)) < 1e-14


def test_mapped_fp32_and_compact(cg, oracle):
    wl = cg.workloads
    celldims, h = (9, 8, 7), 2
    terms = wl.wavy_3d_terms(celldims)
    ref = oracle.mapped_eval(terms, 0.0, celldims, h, "curvature")
    g32 = cg.MappedGrid(terms, {}, celldims, h, T=np.float32)
    dom = domain_indices(ref["nfull"], h)
    # analytic mapping in FP32: eps32 x conditioning (coordinates up to 6, spacing ~1.3 on this 9 x 8 x 7 mesh)
    assert_fp32_close(grid_outputs(cg, g32), ref, wl.nodes_from_terms(terms, 0.0, celldims, h), dom, label="mapped fp32", dx=1.3)
    comp = cg.MappedGrid(terms, {}, celldims, h, profile="compact")
    J = cg.cell_metrics(comp).reshape(-1).cpu().numpy()
    assert np.abs(J - ref["cell_fwd"][:, 9]).max() < 1e-12 * np.abs(J).max()
    for a in range(3):
        got = cg.face_metrics(comp)[a].reshape(-1, 3).cpu().numpy()
        want = np.stack([ref["face_cons"][a][:, a + 3 * c] for c in range(3)], axis=1)
        assert np.abs(got - want).max() < 1e-12 * np.abs(want).max()


@pytest.mark.parametrize("name", ["wavy3d", "sphere3d", "rtp3d"])
def test_mapped_ad_thomas_lombard(cg, oracle, name):
    """ADThomasLombardMetric (metrics.jl:607-668, pair tables :2336-3080, fill :3082-3188; test_3d_continuous.jl:167-300)
    through its own kernel (k_mapped_adtl3d_staged) against the oracle's INDEPENDENT restatement of the pair-table
    assembly (outer reconstruction first, array difference second: oracle/metric_oracle.hpp AdtlPairTable): GCL < 1e-12,
    active rows equal to the CurvatureCorrected rows at the reference's 1e-12, inactive rows zero, face forward metric =
    Jacobian at the face centre."""
    terms, celldims, h = _cases(cg.workloads)[name]
    t = 0.37
    ref = oracle.mapped_eval(terms, t, celldims, h, "adtl")
    g = cg.MappedGrid(terms, {}, celldims, h, conserved_metric_scheme=cg.ADThomasLombardMetric, t=t)
    got = grid_outputs(cg, g)
    assert_metrics_close(got, ref, rtol=RTOL_F64, label=f"{name} adtl")
    assert max(cg.gcl(g)) < 1e-12                                  # test_3d_continuous.jl:296
    cur = grid_outputs(cg, cg.MappedGrid(terms, {}, celldims, h, conserved_metric_scheme="curvature", t=t))
    ocur = oracle.mapped_eval(terms, t, celldims, h, "curvature", what=2)
    rows = np.arange(9) % 3                                        # column-major: row index of entry k
    for ax in range(3):
        a, c = got["face_cons"][ax], cur["face_cons"][ax]
        scale = np.abs(c).max()
        assert np.abs(a[..., rows == ax] - c[..., rows == ax]).max() <= 1e-12 * scale   # :297
        assert not a[..., rows != ax].any()
        # the two restatements of the oracle (pair tables / closure graph) agree as well
        assert np.abs(ref["face_cons"][ax][..., rows == ax] - ocur["face_cons"][ax][..., rows == ax]).max() <= 1e-12 * scale
    # the face forward metric differs from the reconstructed one of the Curvature scheme on a curved mesh
    assert np.abs(got["face_fwd"] - cur["face_fwd"]).max() > 0
    # a face-only refresh and the COMPACT profile go through the CurvatureCorrected evaluation + overwrite
    g2 = cg.MappedGrid(terms, {}, celldims, h, conserved_metric_scheme=cg.ADThomasLombardMetric, t=t, cache_mode="lazy")
    cg.refresh_cell_metrics(g2)
    cg.refresh_face_metrics(g2)
    assert_metrics_close(grid_outputs(cg, g2), ref, rtol=RTOL_F64, label=f"{name} adtl separate refreshes")
    gc = cg.MappedGrid(terms, {}, celldims, h, conserved_metric_scheme=cg.ADThomasLombardMetric, t=t, profile="compact")
    for ax in range(3):
        want = np.stack([ref["face_cons"][ax][:, ax + 3 * c] for c in range(3)], axis=1)
        gotc = cg.face_metrics(gc)[ax].reshape(-1, 3).cpu().numpy()
        assert np.abs(gotc - want).max() <= 1e-12 * np.abs(want).max()


def test_ad_thomas_lombard_dimension_and_grid_kind(cg, oracle):
    wl = cg.workloads
    terms, celldims, h = _cases(wl)["wavy2d"]
    with pytest.warns(UserWarning):                                # metrics.jl:1669-1673
        g = cg.MappedGrid(terms, {}, celldims, h, conserved_metric_scheme="adtl", t=0.1)
    ref = oracle.mapped_eval(terms, 0.1, celldims, h, "curvature")
    assert_metrics_close(grid_outputs(cg, g), ref, rtol=RTOL_F64, label="adtl 2d = curvature")
    with pytest.raises(cg.CurvigridArgumentError):
        cg.DiscreteGrid(*wl.wavy_nodes_3d(9, 9, 9), 2, conserved_metric_scheme="adtl")

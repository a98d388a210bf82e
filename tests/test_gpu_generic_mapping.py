"""MappedGrid with an ARBITRARY analytic mapping given as CUDA source (cg_grid_set_mapping_source: NVRTC-compiled jet
arithmetic + generic kernels) against the oracle's nested-AD evaluation of the same closures.  The reference's MappedGrid
takes any Julia closures (mapped_grid.jl:421-540); the non-separable test mappings are FuncMap ids 1-3 of
oracle/metric_oracle.hpp.  Tolerance 1e-12 relative (per-matrix scale), as for every other FP64 path."""
import numpy as np
import pytest

from util import RTOL_F64, assert_metrics_close, domain_indices, grid_outputs, metric_err

pytestmark = pytest.mark.gpu
SCHEMES = ("endpoint", "gradient", "curvature")

SRC = {
    1: """
template <class S> __device__ void cg_map(S xi, S eta, S zeta, double t, const double* p, S& x, S& y, S& z) {
  x = xi + p[0] * sin(p[1] * (eta + 0.5 * zeta) + 0.3 * t);
  y = eta * (1.0 + p[2] * xi) / (1.0 + p[3] * zeta * zeta);
  z = zeta + p[0] * exp(-p[4] * (xi - eta) * (xi - eta)) * cos(t) + sqrt(1.0 + 0.01 * xi * xi);
}""",
    2: """
template <class S> __device__ void cg_map(S xi, S eta, S zeta, double t, const double* p, S& x, S& y, S& z) {
  x = xi + p[0] * sin(p[1] * xi * eta + t);
  y = eta + p[2] * sqrt(1.0 + p[3] * xi * xi + 0.5 * eta * eta);
}""",
    3: """
template <class S> __device__ void cg_map(S xi, S eta, S zeta, double t, const double* p, S& x, S& y, S& z) {
  x = xi + p[0] * sin(p[1] * xi) * exp(-p[2] * xi) + p[3] * t;
}""",
}
CASES = {1: ((9, 8, 7), [0.3, 0.4, 0.01, 0.001, 0.02]), 2: ((14, 11), [0.2, 0.05, 0.3, 0.1]), 3: ((23,), [0.1, 0.6, 0.02, 0.5])}


@pytest.mark.parametrize("scheme", SCHEMES)
@pytest.mark.parametrize("map_id", [1, 2, 3])
def test_generic_mapping_matches_oracle(cg, oracle, map_id, scheme):
    celldims, prm = CASES[map_id]
    h, t = 3, 0.7
    ref = oracle.funcmap_eval(len(celldims), map_id, prm, t, celldims, h, scheme)
    g = cg.MappedGrid(SRC[map_id], prm, celldims, h, conserved_metric_scheme=scheme, t=t)
    got = grid_outputs(cg, g)
    assert got["nfull"] == ref["nfull"]
    assert_metrics_close(got, ref, rtol=RTOL_F64, label=f"generic map {map_id} {scheme}")


@pytest.mark.parametrize("map_id", [1, 2, 3])
def test_generic_mapping_coordinates(cg, oracle, map_id):
    celldims, prm = CASES[map_id]
    N, h, t = len(celldims), 3, 0.2
    g = cg.MappedGrid(SRC[map_id], prm, celldims, h, t=t)
    for which, arrs in ((0, g.node_coordinates), (1, g.centroid_coordinates)):
        ref = oracle.funcmap_coords(N, map_id, prm, t, celldims, h, which)
        for q in range(N):
            assert np.abs(arrs[q].reshape(-1).cpu().numpy() - ref[q]).max() < 1e-12
    for a in range(N):
        ref = oracle.funcmap_coords(N, map_id, prm, t, celldims, h, 2 + a)
        for q in range(N):
            assert np.abs(g.face_coordinates[a][q].reshape(-1).cpu().numpy() - ref[q]).max() < 1e-12


@pytest.mark.parametrize("name", ["readme2d", "wavy3d", "rtp3d", "affine1d"])
def test_generic_path_equals_table_path_on_separable_mappings(cg, oracle, name):
    """Any term list also runs through the generic kernels (workloads.terms_to_cuda_source): the two MappedGrid paths and
    the oracle agree on the reference's own test mappings."""
    wl = cg.workloads
    terms, celldims, h = {
        "readme2d": (wl.readme_wavy_2d_terms(), (16, 12), 3),
        "wavy3d": (wl.wavy_3d_terms((9, 8, 7), time_amp=0.1), (9, 8, 7), 2),
        "rtp3d": (wl.spherical_rtp_3d_terms((8, 8, 8)), (8, 8, 8), 2),
        "affine1d": ([wl.term(0, 1.0, {0: (wl.AFFINE, 0.3, 0.25)}), wl.term(0, 0.02, {0: (wl.SIN, 0.0, 0.7)})], (11,), 2),
    }[name]
    t = 0.37
    gen = grid_outputs(cg, cg.MappedGrid(wl.terms_to_cuda_source(terms), None, celldims, h, t=t))
    tab = grid_outputs(cg, cg.MappedGrid(terms, {}, celldims, h, t=t))
    ref = oracle.mapped_eval(terms, t, celldims, h, "curvature")
    assert_metrics_close(gen, ref, rtol=RTOL_F64, label=f"{name} generic vs oracle")
    assert_metrics_close(gen, tab, rtol=RTOL_F64, label=f"{name} generic vs tables")


def test_generic_mapping_update_params_window_compact_fp32(cg, oracle):
    celldims, prm = CASES[1]
    h = 3
    g = cg.MappedGrid(SRC[1], prm, celldims, h, cache_mode="lazy")
    # update!(grid, t, params): new time and new parameters, lazily re-evaluated (mapped_grid.jl:551-593)
    prm2 = [0.25, 0.3, 0.02, 0.002, 0.03]
    cg.update(g, 1.1, prm2)
    assert not cg.cell_metrics_valid(g) and not cg.face_metrics_valid(g)
    ref = oracle.funcmap_eval(3, 1, prm2, 1.1, celldims, h, "curvature")
    assert_metrics_close(grid_outputs(cg, g), ref, rtol=RTOL_F64, label="after update!")
    # GCL residual of the conserved metrics at round-off (the acceptance functional of the reference, gcl.jl:20-117)
    assert max(cg.gcl(g)) < 1e-12
    # local_grid: a window of the same mapping reproduces the slices of the global grid (mapped_grid.jl:344-379)
    loc = cg.local_grid(g, ((2, 7), (1, 6), (0, 7)))
    full = ref
    nf = ref["nfull"]
    idx = np.arange(int(np.prod(nf))).reshape(nf[::-1]).transpose(2, 1, 0)   # [i, j, k] -> linear (i fastest)
    sel = idx[2:7 + 2 * h, 1:6 + 2 * h, 0:7 + 2 * h].transpose(2, 1, 0).ravel()
    got = grid_outputs(cg, loc)
    for k in ("cell_fwd", "cell_inv"):
        assert metric_err(got[k], full[k][sel], True) <= RTOL_F64
    assert metric_err(got["face_cons"], full["face_cons"][:, sel], False) <= RTOL_F64
    # COMPACT profile: J and the active rows
    gc = cg.MappedGrid(SRC[1], prm2, celldims, h, profile="compact", t=1.1)
    J = cg.cell_metrics(gc).reshape(-1).cpu().numpy()
    assert np.abs(J - ref["cell_fwd"][:, 9]).max() <= 1e-12 * np.abs(J).max()
    act = cg.face_metrics(gc)
    for a in range(3):
        want = np.stack([ref["face_cons"][a][:, a + 3 * c] for c in range(3)], axis=1)
        assert np.abs(act[a].reshape(-1, 3).cpu().numpy() - want).max() <= 1e-12 * np.abs(want).max()
    # FP32 storage: FP64 arithmetic, cast on store
    g32 = cg.MappedGrid(SRC[1], prm2, celldims, h, T=np.float32, t=1.1)
    got32 = grid_outputs(cg, g32)
    assert metric_err(got32["cell_fwd"], ref["cell_fwd"], True) < 1e-6
    assert metric_err(got32["face_cons"], ref["face_cons"], False) < 1e-6


def test_generic_mapping_compile_error_is_an_argument_error(cg):
    bad = "template <class S> __device__ void cg_map(S xi, S eta, S zeta, double t, const double* p, S& x, S& y, S& z) { x = nope(xi); }"
    with pytest.raises(cg.CurvigridArgumentError, match="nope"):
        cg.MappedGrid(bad, None, (8, 8), 2)
    with pytest.raises(cg.CurvigridArgumentError):   # ADThomasLombardMetric needs a term list
        cg.MappedGrid(SRC[1], CASES[1][1], (8, 8, 8), 2, conserved_metric_scheme=cg.ADThomasLombardMetric)

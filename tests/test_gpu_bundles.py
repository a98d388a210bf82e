"""Device-side solver bundles: face_flux_geometry for every face and cellvolumes, against the
single-index mirrors of face_geometry.jl:393-416 / generation.jl:615-726 and a numpy restatement on the conserved
arrays (which the parity tests pin to the oracle)."""
import math

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _scale_face(cs, q):
    # conservation_face_metric_component_scale, face_geometry.jl:302-332 (numpy, whole arrays)
    N = len(q)
    if cs == "spherical":
        r = np.abs(q[0])
        st = np.sin(q[1]) if N > 1 else None
        return ([4 * math.pi * r * r] if N == 1 else [2 * math.pi * r * r * st, 2 * math.pi * r * st] if N == 2
                else [r * r * st, r * st, r])
    if cs == "cylindrical":
        return [2 * math.pi * np.abs(q[0])] * N
    if cs == "axi_x":
        return [2 * math.pi * np.abs(q[1])] * N
    if cs == "axi_y":
        return [2 * math.pi * np.abs(q[0])] * N
    return [np.ones_like(q[0])] * N


_CS = {"unit": ("CurvilinearCS", "CartesianBasis"), "spherical": ("SphericalCS", "SphericalBasis"),
       "cylindrical": ("CylindricalCS", "CartesianBasis"), "axi_x": ("AxisymmetricCS{:x}", "CartesianBasis"),
       "axi_y": ("AxisymmetricCS{:y}", "CartesianBasis")}


def _grid(cg, dim, cs, profile):
    if dim == 3:
        nodes = cg.workloads.nodes_from_terms(cg.workloads.spherical_rtp_3d_terms((9, 8, 7)), 0.0, (9, 8, 7), 0) \
            if cs == "spherical" else cg.workloads.wavy_nodes_3d(10, 9, 8)
    else:
        x, y = cg.workloads.wavy_nodes_2d(14, 12)
        nodes = (x + 1.5, y + 0.7)  # positive radii / polar angles for the scaled systems
    csn, basis = _CS[cs]
    return cg.DiscreteGrid(*nodes, 3, coordinate_system=csn, basis=basis, cache_mode="lazy", profile=profile)


@pytest.mark.parametrize("dim,cs", [(3, "unit"), (3, "spherical"), (2, "unit"), (2, "cylindrical"), (2, "axi_x"),
                                    (2, "axi_y"), (2, "spherical")])
@pytest.mark.parametrize("profile", ["full", "compact"])
def test_face_flux_geometry_arrays(dim, cs, profile):
    import curvigrid_b200 as cg

    g = _grid(cg, dim, cs, profile)
    N = g.dim
    fm = cg.face_metrics(g)
    for axis in range(N):
        b = cg.face_flux_geometry_arrays(g, axis)
        if profile == "compact":
            row = [fm[axis][..., c].cpu().numpy() for c in range(N)]
        else:
            cons = fm[axis].conserved.cpu().numpy()
            row = [cons[..., axis + N * c] for c in range(N)]
        q = [t.cpu().numpy() for t in g.face_coordinates[axis]]
        sc = _scale_face(cs, q)
        mv = np.stack([sc[c] * row[c] for c in range(N)])
        area = np.sqrt((mv * mv).sum(0))
        np.testing.assert_allclose(b.metric_vector.cpu().numpy(), mv, rtol=1e-13, atol=0)
        np.testing.assert_allclose(b.area.cpu().numpy(), area, rtol=1e-13, atol=0)
        np.testing.assert_allclose(b.normal.cpu().numpy(), mv / area, rtol=1e-12, atol=1e-15)
        if profile == "full":  # the single-index mirror of the reference accessor, both sides of a cell
            idx = tuple(4 + d for d in range(N))
            loc = "ijk"[axis]
            hi, lo = cg.face_flux_geometry(g, idx, loc + "hi"), cg.face_flux_geometry(g, idx, loc + "lo")
            t = tuple(reversed(idx))
            tl = list(idx); tl[axis] -= 1; tl = tuple(reversed(tl))
            np.testing.assert_allclose(b.metric_vector.cpu().numpy()[(slice(None),) + t], hi.metric_vector, rtol=1e-13)
            np.testing.assert_allclose(-b.metric_vector.cpu().numpy()[(slice(None),) + tl], lo.metric_vector, rtol=1e-13)
            assert abs(float(b.area[t]) - hi.area) <= 1e-13 * hi.area


@pytest.mark.parametrize("dim,cs", [(3, "unit"), (3, "spherical"), (2, "cylindrical"), (2, "axi_x"), (2, "spherical")])
@pytest.mark.parametrize("profile", ["full", "compact"])
def test_cellvolumes(dim, cs, profile):
    import curvigrid_b200 as cg

    g = _grid(cg, dim, cs, profile)
    N = g.dim
    vol = cg.cellvolumes(g, include_halo=True).cpu().numpy()
    cm = cg.cell_metrics(g)
    J = cm.cpu().numpy() if profile == "compact" else cm.forward.cpu().numpy()[..., -1]
    q = [t.cpu().numpy() for t in g.centroid_coordinates]
    if cs == "spherical":
        sc = (q[0] ** 2 * np.sin(q[1])) if N == 3 else 2 * math.pi * q[0] ** 2 * np.sin(q[1])
    elif cs == "cylindrical":
        sc = 2 * math.pi * np.abs(q[0])
    elif cs == "axi_x":
        sc = 2 * math.pi * np.abs(q[1])
    else:
        sc = np.ones_like(J)
    np.testing.assert_allclose(vol, sc * np.abs(J), rtol=1e-13)
    h = g.nhalo
    dom = cg.cellvolumes(g).cpu().numpy()
    assert dom.shape == tuple(n - 2 * h for n in vol.shape)
    np.testing.assert_array_equal(dom, vol[tuple(slice(h, n - h) for n in vol.shape)])
    if profile == "full":
        idx = tuple(4 + d for d in range(N))
        assert abs(cg.cellvolume(g, idx) - vol[tuple(reversed(idx))]) <= 1e-13 * abs(vol[tuple(reversed(idx))])


def test_bundle_errors():
    import curvigrid_b200 as cg

    g = _grid(cg, 2, "unit", "full")
    with pytest.raises(ValueError):
        cg.face_flux_geometry_arrays(g, 2)

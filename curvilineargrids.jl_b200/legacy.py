"""Host-side mirror of the reference's LEGACY curvilinear grid for the metric path.

    CurvilinearGrid3D(x, y, z, :meg6 | :meg6_symmetric; halo_coords_included)   3d.jl:51-115
    update!(mesh, force, include_halo_region)                                   3d.jl:472-493
    mesh.cell_center_metrics / mesh.edge_metrics (SoA, metric_soa.jl:2-170)
    mesh.centroid_coordinates

(paths under /root/reference/src/grids/curvilinear_mapped_grids/ and src/grids/).  The device
pipeline is cg_legacy_meg6_update_3d in libcurvigrid_cuda.so (csrc/cg_legacy.cu).
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

from . import _lib as L

_SCHEMES = {"meg6": False, "meg6_symmetric": True}
_AX = ("ξ", "η", "ζ")
_X = ("x₁", "x₂", "x₃")


class CurvilinearGrid3D:
    nhalo = 5  # MEG6 (DiscretizationSchemes.jl:53-60)

    def __init__(self, x, y, z, discretization_scheme="meg6", *, device=None, halo_coords_included=False,
                 is_static=False, init_metrics=True):
        import torch
        if not torch.cuda.is_available():
            raise RuntimeError("curvigrid: a CUDA device is required (no CPU fallback)")
        key = str(discretization_scheme).lower().lstrip(":")
        if key not in _SCHEMES:
            # GridTypes.jl:595-626 knows :meg2/:meg4 too; their one-sided edge kernels only exist for
            # SixthOrder in the reference (derivative_kernels.jl:184-292), so only MEG6 is buildable
            raise ValueError(f"Unknown or unsupported discretization scheme {discretization_scheme!r}; "
                             "use :meg6 or :meg6_symmetric")
        self.symmetric = _SCHEMES[key]
        self.discretization_scheme_name = key
        self.device = int(torch.cuda.current_device() if device is None else device)
        self.halo_coords_included = bool(halo_coords_included)
        self.is_static = bool(is_static)
        shp = tuple(np.shape(x))
        if tuple(np.shape(y)) != shp or tuple(np.shape(z)) != shp:
            raise ValueError("Size mismatch for the given x, y, z coordinate arrays, they must all be the same")
        self.nnodes = shp
        h = self.nhalo
        self.celldims_full = tuple(s - 1 + (0 if halo_coords_included else 2 * h) for s in shp)
        dev = f"cuda:{self.device}"
        self._nodes = [torch.from_numpy(np.ascontiguousarray(np.asarray(a, dtype=np.float64).transpose(2, 1, 0))).to(dev)
                       for a in (x, y, z)]
        nc = int(np.prod(self.celldims_full))
        self._cell = torch.zeros((28, nc), dtype=torch.float64, device=dev)
        self._edge = torch.zeros((3, 19, nc), dtype=torch.float64, device=dev)
        self._cent = torch.zeros((3, nc), dtype=torch.float64, device=dev)
        self._scratch = torch.empty((7, nc), dtype=torch.float64, device=dev)
        self.launches = 0
        if init_metrics:
            update(self, True, self.halo_coords_included)

    # --- SoA views in the reference's naming (arrays shaped [nk, nj, ni]) ------------------------
    def _v(self, t):
        return t.view(tuple(reversed(self.celldims_full)))

    @property
    def centroid_coordinates(self):
        return {"x": self._v(self._cent[0]), "y": self._v(self._cent[1]), "z": self._v(self._cent[2])}

    @property
    def cell_center_metrics(self):
        c = self._cell
        out = {"J": self._v(c[0])}
        for q in range(3):
            out[_X[q]] = {_AX[a]: self._v(c[1 + 3 * q + a]) for a in range(3)}
        for r in range(3):
            out[_AX[r]] = {_X[k]: self._v(c[10 + 3 * r + k]) for k in range(3)}
            out[_AX[r] + "̂"] = {_X[k]: self._v(c[19 + 3 * r + k]) for k in range(3)}
        return out

    @property
    def edge_metrics(self):
        names = ("i₊½", "j₊½", "k₊½")
        out = {}
        for ax in range(3):
            e = self._edge[ax]
            d = {"J": self._v(e[0])}
            for r in range(3):
                d[_AX[r]] = {_X[k]: self._v(e[1 + 3 * r + k]) for k in range(3)}
                d[_AX[r] + "̂"] = {_X[k]: self._v(e[10 + 3 * r + k]) for k in range(3)}
            out[names[ax]] = d
        return out


def _check_valid_metrics(mesh, dim, cell_rows, positive_row, edge_rows, edge_rows_per_axis):
    """_check_valid_metrics (1d.jl:457-482, 2d.jl:748-788, 3d.jl:600-677): ONE fused device reduction
    (cg_legacy_check_valid) instead of the reference's ~60 all(isfinite.(...)) passes; same error messages."""
    import torch
    nf = (ctypes.c_int64 * 3)(*(list(mesh.celldims_full) + [1] * (3 - dim)))
    cr = (ctypes.c_int * max(1, len(cell_rows)))(*cell_rows)
    er = (ctypes.c_int * max(1, len(edge_rows)))(*edge_rows)
    c_ok, e_ok = ctypes.c_int(), ctypes.c_int()
    st = ctypes.c_void_p(torch.cuda.current_stream(mesh.device).cuda_stream)
    rc = L.lib().cg_legacy_check_valid(mesh.device, dim, nf, mesh.nhalo, ctypes.c_void_p(mesh._cell.data_ptr()), cr,
                                       len(cell_rows), positive_row, ctypes.c_void_p(mesh._edge.data_ptr()),
                                       edge_rows_per_axis, er, len(edge_rows), ctypes.byref(c_ok), ctypes.byref(e_ok), st)
    if rc != L.CG_OK:
        raise L.CurvigridError(rc, L.lib().cg_legacy_last_error().decode())
    if not e_ok.value:
        raise RuntimeError("Invalid edge metrics found")
    if not c_ok.value:
        raise RuntimeError("Invalid centroid metrics found")


def update(mesh, force=False, include_halo_region=False):
    """update!(mesh::AbstractCurvilinearGrid3D, force, include_halo_region) — 3d.jl:472-493."""
    import torch
    if mesh.is_static and not force:
        import warnings
        warnings.warn("Attempting to update grid metrics when grid.is_static = true!")
        return None
    if bool(include_halo_region) != mesh.halo_coords_included:
        # the reference evaluates on cell.full only when the node arrays carry real halo coordinates
        raise ValueError("include_halo_region must match halo_coords_included")
    nn = (ctypes.c_int64 * 3)(*mesh.nnodes)
    nl = ctypes.c_uint64()
    p = lambda t: ctypes.c_void_p(t.data_ptr())
    st = ctypes.c_void_p(torch.cuda.current_stream(mesh.device).cuda_stream)
    rc = L.lib().cg_legacy_meg6_update_3d(
        mesh.device, p(mesh._nodes[0]), p(mesh._nodes[1]), p(mesh._nodes[2]), nn, mesh.nhalo,
        int(mesh.halo_coords_included), int(mesh.symmetric), p(mesh._cell), p(mesh._edge), p(mesh._cent),
        p(mesh._scratch), ctypes.c_size_t(mesh._scratch.numel() * 8), st, ctypes.byref(nl))
    if rc != L.CG_OK:
        msg = L.lib().cg_legacy_last_error().decode()
        raise (L.CurvigridArgumentError if rc == L.CG_ERR_ARG else L.CurvigridError)(rc, msg)
    mesh.launches = int(nl.value)
    # 3d.jl:600-677: J finite and > 0, the 9 forward and 9 inverse metrics finite on cell.domain (rows 0..18 of the
    # cell SoA); the 9 hatted edge metrics of each axis finite on expand(domain, axis, -1) (rows 10..18)
    if os.environ.get("CG_LEGACY_FUSED", "1") != "0":
        # the fused launches tested their own values while writing them (cg_legacy_validity): two flag words to read
        c_ok, e_ok = ctypes.c_int(), ctypes.c_int()
        rc = L.lib().cg_legacy_validity(mesh.device, ctypes.byref(c_ok), ctypes.byref(e_ok), st)
        if rc != L.CG_OK:
            raise L.CurvigridError(rc, L.lib().cg_legacy_last_error().decode())
        if not e_ok.value:
            raise RuntimeError("Invalid edge metrics found")
        if not c_ok.value:
            raise RuntimeError("Invalid centroid metrics found")
        return None
    _check_valid_metrics(mesh, 3, list(range(19)), 0, list(range(10, 19)), 19)
    return None


def set_node_coordinates(mesh, x, y, z):
    """Mutate mesh.node_coordinates in place (the moving-mesh use of the legacy path), then call update."""
    import torch
    for t, a in zip(mesh._nodes, (x, y, z)):
        t.copy_(torch.from_numpy(np.ascontiguousarray(np.asarray(a, dtype=np.float64).transpose(2, 1, 0))))


class SphericalBasisCurvilinearGrid3D(CurvilinearGrid3D):
    """SphericalBasisCurvilinearGrid3D(r, θ, ϕ, :meg6 | :meg6_symmetric) — 3d_spherical_basis.jl:51-215.

    The reference stores (r, θ, ϕ) as the node coordinates and runs the SAME MEG6 pipeline on them
    (`update!`, 3d_spherical_basis.jl:191-215: centroid kernel :269-315 = the eight-node mean, then
    `update_metrics!`): the metrics are taken in the stored coordinates.  The only extra is
    `cartesian_node_coordinates` (x = r sinθ cosϕ, y = r sinθ sinϕ, z = r cosθ, :217-252), refreshed by
    `update`.  Here: the device pipeline of CurvilinearGrid3D on the (r, θ, ϕ) arrays plus that conversion as
    device tensor arithmetic (a read-side convenience, not a metric kernel)."""

    def __init__(self, r, theta, phi, discretization_scheme="meg6", **kw):
        super().__init__(r, theta, phi, discretization_scheme, **kw)
        self._refresh_cartesian()

    def _refresh_cartesian(self):
        import torch
        r, th, ph = self._nodes
        st = torch.sin(th)
        self._xyz = (r * st * torch.cos(ph), r * st * torch.sin(ph), r * torch.cos(th))

    @property
    def node_coordinates(self):
        return {"r": self._nodes[0], "θ": self._nodes[1], "ϕ": self._nodes[2]}

    @property
    def centroid_coordinates(self):
        return {"r": self._v(self._cent[0]), "θ": self._v(self._cent[1]), "ϕ": self._v(self._cent[2])}

    @property
    def cartesian_node_coordinates(self):
        return {"x": self._xyz[0], "y": self._xyz[1], "z": self._xyz[2]}


_update_3d = update


def update(mesh, force=False, include_halo_region=False):  # noqa: F811
    """update!(mesh, force, include_halo_region) — 3d.jl:472-493 / 3d_spherical_basis.jl:191-215."""
    out = _update_3d(mesh, force, include_halo_region)
    if isinstance(mesh, SphericalBasisCurvilinearGrid3D) and hasattr(mesh, "_xyz") and (not mesh.is_static or force):
        mesh._refresh_cartesian()
    return out


class CurvilinearGrid2D:
    """CurvilinearGrid2D(x, y, :meg6 | :meg6_symmetric; halo_coords_included) — 2d.jl:120-200, `update!` :616-637.

    Device pipeline: cg_legacy_meg6_update_2d (centroids, MEG6 forward metrics, inverse = inv [x_ξ x_η; y_ξ y_η],
    hatted = inverse * J, edge interpolation).  In 2-D the symmetric scheme is the same computation
    (conservative_metrics.jl:417-421).  SoA views in the reference's naming, arrays shaped [nj, ni]."""
    nhalo = 5

    def __init__(self, x, y, discretization_scheme="meg6", *, device=None, halo_coords_included=False, is_static=False,
                 init_metrics=True):
        import torch
        if not torch.cuda.is_available():
            raise RuntimeError("curvigrid: a CUDA device is required (no CPU fallback)")
        key = str(discretization_scheme).lower().lstrip(":")
        if key not in _SCHEMES:
            raise ValueError(f"Unknown or unsupported discretization scheme {discretization_scheme!r}; "
                             "use :meg6 or :meg6_symmetric")
        self.symmetric = _SCHEMES[key]
        self.discretization_scheme_name = key
        self.device = int(torch.cuda.current_device() if device is None else device)
        self.halo_coords_included = bool(halo_coords_included)
        self.is_static = bool(is_static)
        shp = tuple(np.shape(x))
        if tuple(np.shape(y)) != shp or len(shp) != 2:
            raise ValueError("Size mismatch for the given x, y coordinate arrays, they must all be the same")
        self.nnodes = shp
        h = self.nhalo
        self.celldims_full = tuple(s - 1 + (0 if halo_coords_included else 2 * h) for s in shp)
        dev = f"cuda:{self.device}"
        self._nodes = [torch.from_numpy(np.ascontiguousarray(np.asarray(a, dtype=np.float64).T)).to(dev) for a in (x, y)]
        nc = int(np.prod(self.celldims_full))
        self._cell = torch.zeros((13, nc), dtype=torch.float64, device=dev)
        self._edge = torch.zeros((2, 9, nc), dtype=torch.float64, device=dev)
        self._cent = torch.zeros((2, nc), dtype=torch.float64, device=dev)
        self._scratch = torch.empty((4, nc), dtype=torch.float64, device=dev)
        self.launches = 0
        if init_metrics:
            update2d(self, True, self.halo_coords_included)

    def _v(self, t):
        return t.view(tuple(reversed(self.celldims_full)))

    @property
    def centroid_coordinates(self):
        return {"x": self._v(self._cent[0]), "y": self._v(self._cent[1])}

    @property
    def cell_center_metrics(self):
        c = self._cell
        out = {"J": self._v(c[0])}
        for q in range(2):
            out[_X[q]] = {_AX[a]: self._v(c[1 + 2 * q + a]) for a in range(2)}
        for r in range(2):
            out[_AX[r]] = {_X[k]: self._v(c[5 + 2 * r + k]) for k in range(2)}
            out[_AX[r] + "̂"] = {_X[k]: self._v(c[9 + 2 * r + k]) for k in range(2)}
        return out

    @property
    def edge_metrics(self):
        names = ("i₊½", "j₊½")
        out = {}
        for ax in range(2):
            e = self._edge[ax]
            d = {"J": self._v(e[0])}
            for r in range(2):
                d[_AX[r]] = {_X[k]: self._v(e[1 + 2 * r + k]) for k in range(2)}
                d[_AX[r] + "̂"] = {_X[k]: self._v(e[5 + 2 * r + k]) for k in range(2)}
            out[names[ax]] = d
        return out


def update2d(mesh, force=False, include_halo_region=False):
    """update!(mesh::AbstractCurvilinearGrid2D, force, include_halo_region) — 2d.jl:616-637 (a static grid without
    `force` is an error there, unlike the 3-D warning)."""
    import torch
    if mesh.is_static and not force:
        raise RuntimeError("Attempting to update grid metrics when grid.is_static = true!")
    if bool(include_halo_region) != mesh.halo_coords_included:
        raise ValueError("include_halo_region must match halo_coords_included")
    nn = (ctypes.c_int64 * 2)(*mesh.nnodes)
    nl = ctypes.c_uint64()
    p = lambda t: ctypes.c_void_p(t.data_ptr())
    st = ctypes.c_void_p(torch.cuda.current_stream(mesh.device).cuda_stream)
    rc = L.lib().cg_legacy_meg6_update_2d(
        mesh.device, p(mesh._nodes[0]), p(mesh._nodes[1]), nn, mesh.nhalo, int(mesh.halo_coords_included),
        p(mesh._cell), p(mesh._edge), p(mesh._cent), p(mesh._scratch), ctypes.c_size_t(mesh._scratch.numel() * 8), st,
        ctypes.byref(nl))
    if rc != L.CG_OK:
        msg = L.lib().cg_legacy_last_error().decode()
        raise (L.CurvigridArgumentError if rc == L.CG_ERR_ARG else L.CurvigridError)(rc, msg)
    # _check_valid_metrics (2d.jl:748-753): only `@assert all(isfinite.(J[domain]))` is live in the reference
    try:
        _check_valid_metrics(mesh, 2, [0], -1, [], 9)
    except RuntimeError as e:
        raise AssertionError("non-finite cell-centre Jacobian on cell.domain") from e
    mesh.launches = int(nl.value)
    return None


class AxisymmetricGrid2D(CurvilinearGrid2D):
    """AxisymmetricGrid2D(x, y, snap_to_axis, rotational_axis, :meg6 | :meg6_symmetric) — 2d.jl:511-600; `update!`
    :644-671: centroids, THEN (if `snap_to_axis`) the first node row / column of node.domain is put on the axis
    (`_snap_nodes_to_axis`, :814-827: y = 0 for rotational_axis = :x, x = 0 for :y), then the metrics from the
    centroids.  The snap therefore acts on the node coordinates a later update reads."""

    def __init__(self, x, y, snap_to_axis=True, rotational_axis="x", discretization_scheme="meg6", **kw):
        axis = str(rotational_axis).lstrip(":")
        if axis not in ("x", "y"):
            raise ValueError("rotational_axis must be :x or :y")
        self.snap_to_axis = bool(snap_to_axis)
        self.rotational_axis = axis
        super().__init__(x, y, discretization_scheme, **kw)

    def _snap_nodes_to_axis(self):
        o = 0 if not self.halo_coords_included else self.nhalo  # node.domain inside the stored node arrays [j, i]
        n_i, n_j = self.nnodes
        if self.rotational_axis == "x":
            self._nodes[1][o, o:n_i - o] = 0.0   # y[domain[:, 1]]
        else:
            self._nodes[0][o:n_j - o, o] = 0.0   # x[domain[1, :]]

    @property
    def node_coordinates(self):
        return {"x": self._nodes[0], "y": self._nodes[1]}


_update_2d = update2d


def update2d(mesh, force=False, include_halo_region=False):  # noqa: F811
    """update!(mesh, force, include_halo_region) for CurvilinearGrid2D (2d.jl:616-637) and AxisymmetricGrid2D (:644-671)."""
    out = _update_2d(mesh, force, include_halo_region)
    if isinstance(mesh, AxisymmetricGrid2D) and mesh.snap_to_axis:
        mesh._snap_nodes_to_axis()
    return out


class CurvilinearGrid1D:
    """CurvilinearGrid1D(x, :meg6; halo_coords_included) — 1d.jl:73-138; `update!` :388-408: centroids 0.5 (x_i + x_i+1),
    x_xi by the MEG6 cell-centre derivative (forward_metrics.jl:11-27), xi_x = 1 / x_xi, J = |x_xi|, hatted = xi_x J
    (conservative_metrics.jl:239-245), J / xi_x / hatted interpolated to the i+1/2 edges, `_check_valid_metrics`."""
    nhalo = 5

    def __init__(self, x, discretization_scheme="meg6", *, device=None, halo_coords_included=False, is_static=False,
                 init_metrics=True):
        import torch
        if not torch.cuda.is_available():
            raise RuntimeError("curvigrid: a CUDA device is required (no CPU fallback)")
        key = str(discretization_scheme).lower().lstrip(":")
        if key not in _SCHEMES:
            raise ValueError(f"Unknown or unsupported discretization scheme {discretization_scheme!r}; use :meg6")
        self.discretization_scheme_name = key
        self.device = int(torch.cuda.current_device() if device is None else device)
        self.halo_coords_included = bool(halo_coords_included)
        self.is_static = bool(is_static)
        x = np.ascontiguousarray(np.asarray(x, dtype=np.float64).ravel())
        self.nnodes = (x.size,)
        self.celldims_full = (x.size - 1 + (0 if halo_coords_included else 2 * self.nhalo),)
        dev = f"cuda:{self.device}"
        nc = self.celldims_full[0]
        self._nodes = [torch.from_numpy(x).to(dev)]
        self._cell = torch.zeros((4, nc), dtype=torch.float64, device=dev)
        self._edge = torch.zeros((1, 3, nc), dtype=torch.float64, device=dev)
        self._cent = torch.zeros((nc,), dtype=torch.float64, device=dev)
        self._scratch = torch.empty((3, nc), dtype=torch.float64, device=dev)
        self.launches = 0
        if init_metrics:
            update1d(self, True, self.halo_coords_included)

    @property
    def centroid_coordinates(self):
        return {"x": self._cent}

    @property
    def cell_center_metrics(self):
        c = self._cell
        return {"J": c[0], "x₁": {"ξ": c[1]}, "ξ": {"x₁": c[2]}, "ξ̂": {"x₁": c[3]}}

    @property
    def edge_metrics(self):
        e = self._edge[0]
        return {"i₊½": {"J": e[0], "ξ": {"x₁": e[1]}, "ξ̂": {"x₁": e[2]}}}


def update1d(mesh, force=False, include_halo_region=False):
    """update!(mesh::AbstractCurvilinearGrid1D, force, include_halo_region) — 1d.jl:388-408 (a static grid without
    `force` is silently left alone there)."""
    import torch
    if mesh.is_static and not force:
        return None
    if bool(include_halo_region) != mesh.halo_coords_included:
        raise ValueError("include_halo_region must match halo_coords_included")
    nn = (ctypes.c_int64 * 1)(*mesh.nnodes)
    nl = ctypes.c_uint64()
    p = lambda t: ctypes.c_void_p(t.data_ptr())
    st = ctypes.c_void_p(torch.cuda.current_stream(mesh.device).cuda_stream)
    rc = L.lib().cg_legacy_meg6_update_1d(mesh.device, p(mesh._nodes[0]), nn, mesh.nhalo, int(mesh.halo_coords_included),
                                          p(mesh._cell), p(mesh._edge), p(mesh._cent), p(mesh._scratch),
                                          ctypes.c_size_t(mesh._scratch.numel() * 8), st, ctypes.byref(nl))
    if rc != L.CG_OK:
        msg = L.lib().cg_legacy_last_error().decode()
        raise (L.CurvigridArgumentError if rc == L.CG_ERR_ARG else L.CurvigridError)(rc, msg)
    mesh.launches = int(nl.value)
    # 1d.jl:457-482: J finite and > 0, xi.x1 and x1.xi finite on cell.domain; the hatted edge metric on expand(domain, 1, -1)
    _check_valid_metrics(mesh, 1, [0, 1, 2], 0, [2], 3)
    return None

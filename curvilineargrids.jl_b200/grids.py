"""Host-side mirror of the reference's unified-grid interface for the metric path.

Same names, argument meaning and error behaviour as the Julia API the hot path
sits behind (paths under /root/reference/src/grids/unified_grids/):

    DiscreteGrid(x[,y[,z]], nhalo; ...)     discrete_grid.jl:428-590
    MappedGrid(mapping, params, celldims, nhalo; ...)   mapped_grid.jl:421-540
    cell_metrics / face_metrics / refresh_* / invalidate_*   metric_cache_api.jl:16-127
    update!(grid, t, params)  -> update(grid, t, params)      mapped_grid.jl:551-593
    local_grid(grid, global_cell_domain)                      mapped_grid.jl:344-379
    gcl(grid)                                                 ../gcl.jl:119-204
    face_flux_geometry(grid, idx, loc), cellvolume(grid, idx) face_geometry.jl:393-416, generation.jl:615-726

All numerical work happens in libcurvigrid_cuda.so; torch supplies device
memory (outputs are torch CUDA tensors bound into the library) and streams.
Array shapes are reversed w.r.t. Julia so that the memory is identical:
Julia `Array{Metric{3}}(ni,nj,nk)` <-> torch `[nk, nj, ni, 10]`.
"""
from __future__ import annotations

import ctypes
import math
from collections import namedtuple

import numpy as np

from . import _lib as L

EndpointAverageReconstruction = "endpoint"
GradientCorrectedReconstruction = "gradient"
CurvatureCorrectedReconstruction = "curvature"
# ADThomasLombardMetric (metrics.jl:135): 3-D MappedGrid; face forward metric = Jacobian at the face centre, conserved =
# active row only (equal to the CurvatureCorrected row); in 1-D / 2-D it simplifies to CurvatureCorrected (:1669-1673)
ADThomasLombardMetric = "adtl"
_SCHEMES = {"endpoint": 0, "gradient": 1, "curvature": 2, "adtl": 3}

CellMetrics = namedtuple("CellMetrics", "forward inverse")
FaceMetrics = namedtuple("FaceMetrics", "forward inverse conserved")
FaceFluxGeometry = namedtuple("FaceFluxGeometry", "coordinate metric_vector area normal")

_CACHE_MODES = ("eager", "lazy", "off")


def _torch():
    import torch
    return torch


def _scheme_id(s):
    if isinstance(s, str) and s in _SCHEMES:
        return _SCHEMES[s]
    # the reference raises a TypeError for a non-trait scheme (test_unified_grid_types.jl:192-197)
    raise TypeError(f"conserved_metric_scheme must be one of {list(_SCHEMES)}, got {s!r}")


def _validate_cache_mode(mode):
    # unified_grids.jl:54-63
    if mode not in _CACHE_MODES:
        raise ValueError(f"Unsupported cache_mode={mode!r}. Expected one of: 'eager', 'lazy', 'off'.")
    return mode


class _UnifiedGrid:
    """Common storage/caches of MappedGrid and DiscreteGrid (unified_grids.jl:34-70)."""

    def __init__(self, dim, celldims, nhalo, *, device, T, scheme, profile, cache_mode, compute_metrics,
                 coordinate_system, basis, window, global_offset):
        torch = _torch()
        if not torch.cuda.is_available():
            raise RuntimeError("curvigrid: a CUDA device is required (no CPU fallback)")
        if nhalo < 0:
            raise AssertionError("nhalo >= 0")
        _validate_cache_mode(cache_mode)
        self.dim = dim
        self.celldims = tuple(int(c) for c in celldims)
        self.nhalo = int(nhalo)
        # default: the current torch device (one process per GPU under torchrun)
        self.device = int(torch.cuda.current_device() if device is None else device)
        self.T = np.dtype(T)
        if self.T not in (np.dtype(np.float64), np.dtype(np.float32)):
            raise ValueError("T must be float64 or float32")
        self.tdtype = torch.float64 if self.T == np.float64 else torch.float32
        self.scheme = scheme
        self.profile = profile
        self.coordinate_system = coordinate_system
        self.basis = basis
        self.nfull = tuple(c + 2 * self.nhalo for c in self.celldims)
        if window is None:
            self.window_origin = (0,) * dim
            self.window_cells = self.nfull
        else:
            self.window_origin, self.window_cells = tuple(window[0]), tuple(window[1])
        self.global_offset = tuple(global_offset) if global_offset is not None else (0,) * dim
        # cache policy: discrete_grid.jl:292-316 / mapped_grid.jl:241-260
        self._metrics_disabled = (not compute_metrics) and cache_mode == "off"
        self.cache_mode = cache_mode if compute_metrics else ("lazy" if cache_mode == "eager" else cache_mode)
        self.state = {"t": 0.0, "params": {}}
        self._h = ctypes.c_void_p()
        L.check(L.lib().cg_grid_create(
            self.device, dim, L.CG_F64 if self.T == np.float64 else L.CG_F32, L.i64x3(self.celldims), self.nhalo,
            _scheme_id(scheme), L.CG_PROFILE_COMPACT if profile == "compact" else L.CG_PROFILE_FULL,
            L.i64x3(self.window_origin, 0), L.i64x3(self.window_cells), L.i64x3(self.global_offset, 0),
            ctypes.byref(self._h)))
        self._bound = {}
        self._coords_enabled = False

    # ---- plumbing ---------------------------------------------------------
    def __del__(self):
        try:
            if getattr(self, "_h", None) and self._h.value:
                L.lib().cg_grid_destroy(self._h)
                self._h = ctypes.c_void_p()
        except Exception:
            pass

    def close(self):
        self.__del__()
        self._bound = {}

    def _stream(self):
        return ctypes.c_void_p(_torch().cuda.current_stream(self.device).cuda_stream)

    def _cell_shape(self):
        return tuple(reversed(self.window_cells))

    def _node_shape(self):
        return tuple(reversed([w + 1 for w in self.window_cells]))

    def _bind(self, array, axis, shape):
        key = (array, axis)
        if key not in self._bound:
            torch = _torch()
            t = torch.empty(shape, dtype=self.tdtype, device=f"cuda:{self.device}")
            L.check(L.lib().cg_grid_bind_output(self._h, array, axis, ctypes.c_void_p(t.data_ptr()),
                                                ctypes.c_size_t(t.numel() * t.element_size())))
            self._bound[key] = t
        return self._bound[key]

    def _require_metric_storage(self, fn):
        if self._metrics_disabled:
            raise ValueError(f"`{fn}` requires metric storage, but this grid was built with "
                             "compute_metrics=false, cache_mode=:off.")

    def _bind_metric_outputs(self, what):
        N = self.dim
        cs = self._cell_shape()
        if self.profile == "compact":
            if what & L.CG_WHAT_CELL:
                self._bind(L.CG_ARR_COMPACT_J, 0, cs)
            if what & L.CG_WHAT_FACE:
                for a in range(N):
                    self._bind(L.CG_ARR_COMPACT_ACTIVE, a, cs + (N,))
            return
        if what & L.CG_WHAT_CELL:
            self._bind(L.CG_ARR_CELL_FORWARD, 0, cs + (N * N + 1,))
            self._bind(L.CG_ARR_CELL_INVERSE, 0, cs + (N * N + 1,))
        if what & L.CG_WHAT_FACE:
            for a in range(N):
                self._bind(L.CG_ARR_FACE_FORWARD, a, cs + (N * N + 1,))
                self._bind(L.CG_ARR_FACE_INVERSE, a, cs + (N * N + 1,))
                self._bind(L.CG_ARR_FACE_CONSERVED, a, cs + (N * N,))

    def _bind_coord_outputs(self):
        N = self.dim
        for q in range(N):
            self._bind(L.CG_ARR_NODE_COORD, q, self._node_shape())
            self._bind(L.CG_ARR_CENTROID_COORD, q, self._cell_shape())
            for a in range(N):
                self._bind(L.CG_ARR_FACE_COORD, 3 * a + q, self._cell_shape())
        self._coords_enabled = True

    def _eval(self, what):
        if what & (L.CG_WHAT_CELL | L.CG_WHAT_FACE):
            self._bind_metric_outputs(what)
        if what & L.CG_WHAT_COORDS:
            self._bind_coord_outputs()
        L.check(L.lib().cg_grid_eval(self._h, what, self._stream()))

    def _valid(self):
        c, f = ctypes.c_int(), ctypes.c_int()
        L.check(L.lib().cg_grid_valid(self._h, ctypes.byref(c), ctypes.byref(f)))
        return bool(c.value), bool(f.value)

    def set_kernel(self, which):
        """0 = gather reference kernels, 1 = tiled kernels (default)."""
        L.check(L.lib().cg_grid_set_option(self._h, L.CG_OPT_KERNEL, ctypes.c_int64(which)))

    def last_eval_ms(self):
        ms = ctypes.c_float()
        L.check(L.lib().cg_grid_last_eval_ms(self._h, ctypes.byref(ms)))
        return ms.value

    def synchronize(self):
        L.check(L.lib().cg_grid_sync(self._h, self._stream()))

    # ---- payload views ------------------------------------------------------
    def _cell_payload(self):
        if self.profile == "compact":
            return self._bound[(L.CG_ARR_COMPACT_J, 0)]
        return CellMetrics(self._bound[(L.CG_ARR_CELL_FORWARD, 0)], self._bound[(L.CG_ARR_CELL_INVERSE, 0)])

    def _face_payload(self):
        if self.profile == "compact":
            return tuple(self._bound[(L.CG_ARR_COMPACT_ACTIVE, a)] for a in range(self.dim))
        return tuple(FaceMetrics(self._bound[(L.CG_ARR_FACE_FORWARD, a)], self._bound[(L.CG_ARR_FACE_INVERSE, a)],
                                 self._bound[(L.CG_ARR_FACE_CONSERVED, a)]) for a in range(self.dim))

    # ---- coordinates --------------------------------------------------------
    def _coords_valid(self):
        v = ctypes.c_int()
        L.check(L.lib().cg_grid_coords_valid(self._h, ctypes.byref(v)))
        return bool(v.value)

    def _ensure_coords(self):
        # the library re-evaluates stored coordinates whenever the nodes / mapping state change (update!,
        # set_nodes, update_from_host); ask it instead of trusting a Python-side flag
        if not self._coords_enabled or not self._coords_valid():
            self._eval(L.CG_WHAT_COORDS)

    @property
    def node_coordinates(self):
        self._ensure_coords()
        return tuple(self._bound[(L.CG_ARR_NODE_COORD, q)] for q in range(self.dim))

    @property
    def centroid_coordinates(self):
        self._ensure_coords()
        return tuple(self._bound[(L.CG_ARR_CENTROID_COORD, q)] for q in range(self.dim))

    @property
    def face_coordinates(self):
        self._ensure_coords()
        return tuple(tuple(self._bound[(L.CG_ARR_FACE_COORD, 3 * a + q)] for q in range(self.dim))
                     for a in range(self.dim))

    @property
    def iterators(self):
        """0-based half-open ranges of generation.jl:3-33 (`full` / `domain`)."""
        h = self.nhalo
        cell_full = tuple((0, n) for n in self.nfull)
        cell_dom = tuple((h, n - h) for n in self.nfull)
        node_full = tuple((0, n + 1) for n in self.nfull)
        node_dom = tuple((h, n + 1 - h) for n in self.nfull)
        return {"cell": {"full": cell_full, "domain": cell_dom}, "node": {"full": node_full, "domain": node_dom},
                "nhalo": h}


# ---------------------------------------------------------------------------
class DiscreteGrid(_UnifiedGrid):
    """DiscreteGrid(x[,y[,z]], nhalo; ...) — discrete_grid.jl:428-590.

    Coordinates are numpy arrays indexed [i,j,k] (any memory order), torch CUDA
    tensors in memory order (shape [nk,nj,ni]), or pinned torch CPU tensors in
    memory order.
    """

    def __init__(self, *args, device=None, T=None, compute_metrics=True, halo_coords_included=False,
                 coordinate_system="CurvilinearCS", basis="CartesianBasis", interpolation="linear",
                 cache_mode="eager", conserved_metric_scheme=CurvatureCorrectedReconstruction, profile="full",
                 window=None, global_celldims=None, node_origin=None):
        *coords, nhalo = args
        dim = len(coords)
        if dim < 1 or dim > 3:
            raise ValueError("DiscreteGrid takes 1 to 3 coordinate arrays")
        if interpolation != "linear":
            raise ValueError("`DiscreteGrid` currently supports only linear interpolation (`interpolation=:linear`).")
        torch = _torch()
        shapes = []
        for c in coords:
            if isinstance(c, torch.Tensor):
                shapes.append(tuple(reversed(c.shape)))
            else:
                shapes.append(tuple(np.shape(c)))
        if any(s != shapes[0] for s in shapes):
            raise ValueError("x, y, and z arrays must have matching dimensions.")
        if T is None:
            c0 = coords[0]
            T = {torch.float32: np.float32}.get(c0.dtype, np.float64) if isinstance(c0, torch.Tensor) else (
                np.float32 if np.asarray(c0).dtype == np.float32 else np.float64)
        nn = shapes[0]
        self.halo_coords_included = bool(halo_coords_included)
        if global_celldims is not None:
            # slab piece of a larger grid: `coords` cover the interior node planes starting at
            # `node_origin`; `window` = (origin, cells) of the stored part of the global cell.full
            celldims = tuple(global_celldims)
            if window is None or node_origin is None:
                raise ValueError("global_celldims requires window and node_origin")
        else:
            celldims = tuple(n - 1 - (2 * nhalo if halo_coords_included else 0) for n in nn)
        if any(c < 1 for c in celldims):
            raise ValueError("DiscreteGrid needs at least 2 interior nodes per axis")
        super().__init__(dim, celldims, nhalo, device=device, T=T, scheme=conserved_metric_scheme, profile=profile,
                         cache_mode=cache_mode, compute_metrics=compute_metrics,
                         coordinate_system=coordinate_system, basis=basis, window=window, global_offset=None)
        self._node_origin = node_origin
        self._set_nodes(coords)
        self.state = {"t": 0.0, "params": {}}
        if not self._metrics_disabled and self.cache_mode == "eager":
            self._eval(L.CG_WHAT_CELL | L.CG_WHAT_FACE)

    def _set_nodes(self, coords):
        torch = _torch()
        ptrs, keep = [], []
        nn = None
        for c in coords:
            if isinstance(c, torch.Tensor):
                t = c.to(self.tdtype).contiguous()
                keep.append(t)
                ptrs.append(ctypes.c_void_p(t.data_ptr()))
                nn = tuple(reversed(t.shape))
            else:
                a = np.asfortranarray(np.asarray(c, dtype=self.T))
                keep.append(a)
                ptrs.append(ctypes.c_void_p(a.ctypes.data))
                nn = a.shape
        ptrs += [None] * (3 - len(ptrs))
        if self._node_origin is not None:
            origin, count = L.i64x3(self._node_origin, 0), L.i64x3(nn)
        else:
            origin, count = None, None
        L.check(L.lib().cg_grid_set_nodes(self._h, ptrs[0], ptrs[1], ptrs[2], int(self.halo_coords_included),
                                          origin, count, self._stream()))
        self._node_inputs = keep  # keep device inputs alive: the library reads them in place
        self._node_input_shape = tuple(reversed(nn))  # memory order [k, j, i]
        if any(isinstance(k, np.ndarray) or not k.is_cuda for k in keep):
            self.synchronize()  # host memory must outlive the async copy

    def update_from_host(self, *coords, what=3, nchunks=8, comm=None, plan=None, gcl=False):
        """update!(grid, x, y, z) from HOST arrays + refresh of both caches, pipelined (cg_grid_update_from_host):
        the H2D copy of a piece of node planes overlaps the metric kernels of the piece before.  `coords`: pinned
        CPU torch tensors (or numpy arrays) in memory order [k, j, i] with the extents of the constructor's arrays.
        With `comm` (slabs.NcclComm) and `plan` the grid is a k-slab: only the owned planes are read from the host
        arrays and the neighbour planes arrive through NCCL meanwhile (cg_grid_update_from_host_dist).  `gcl=True`
        (CG_WHAT_GCL): the GCL residuals are reduced plane by plane behind the kernels; the next gcl(grid) only reads them."""
        torch = _torch()
        if self.dim != 3:
            raise ValueError("update_from_host: 3-D DiscreteGrid only")
        ptrs = []
        for c in coords:
            t = c if isinstance(c, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(c, dtype=self.T))
            if t.is_cuda or not t.is_contiguous() or t.dtype != self.tdtype:
                raise ValueError("update_from_host needs contiguous CPU arrays of the grid's dtype")
            if tuple(t.shape) != self._node_input_shape:
                raise ValueError(f"update_from_host: arrays must have the extents of the constructor's arrays "
                                 f"{self._node_input_shape}, got {tuple(t.shape)}")
            ptrs.append(t)
        if len(ptrs) != 3:
            raise ValueError("update_from_host takes three coordinate arrays")
        self._bind_metric_outputs(what)
        if gcl:
            what = what | L.CG_WHAT_GCL
        if comm is not None and comm.nranks > 1:
            L.check(L.lib().cg_grid_update_from_host_dist(
                self._h, *[ctypes.c_void_p(t.data_ptr()) for t in ptrs], what, int(nchunks), comm._h, comm.rank,
                comm.nranks, comm.bounds(plan), self._stream()))
        else:
            L.check(L.lib().cg_grid_update_from_host(self._h, *[ctypes.c_void_p(t.data_ptr()) for t in ptrs], what,
                                                     int(nchunks), self._stream()))
        self._host_keepalive = ptrs  # the copies are asynchronous

    def set_nodes(self, *coords):
        """Moving sampled mesh: new node arrays (the reference re-creates the grid;
        legacy `update!(mesh, force, include_halo)` is the analogue, 3d.jl:472-493)."""
        self._set_nodes(coords)


class MappedGrid(_UnifiedGrid):
    """MappedGrid(mapping_terms, params, celldims, nhalo; ...) — mapped_grid.jl:421-540.

    The mapping is a separable term list (workloads.py), a callable `params -> term list`, or — for mappings that
    are not separable — a `str` holding the CUDA C++ source of

        template <class S> __device__ void cg_map(S xi, S eta, S zeta, double t, const double* p, S& x, S& y, S& z);

    which the library compiles with NVRTC (cg_grid_set_mapping_source; `params` is then a sequence of floats, the `p`
    array).  Arbitrary Julia closures cannot cross a C ABI.
    """

    def __init__(self, mapping, params, celldims, nhalo, *, device=None, T=np.float64, compute_metrics=True,
                 coordinate_system="CurvilinearCS", basis="CartesianBasis", cache_mode="eager",
                 conserved_metric_scheme=CurvatureCorrectedReconstruction, profile="full",
                 global_cell_indices=None, t=0.0):
        dim = len(celldims)
        goff = None
        if global_cell_indices is not None:
            # generation.jl:10-31: must describe the local non-halo cell domain
            sizes = tuple(hi - lo for lo, hi in global_cell_indices)
            if sizes != tuple(celldims):
                raise ValueError(f"global_cell_indices must describe the local non-halo cell domain with size "
                                 f"{tuple(celldims)}; got {sizes}")
            goff = tuple(lo for lo, _ in global_cell_indices)
        if conserved_metric_scheme == ADThomasLombardMetric and dim < 3:
            import warnings
            warnings.warn("ADThomasLombardMetric simplifies to the direct metric form in 1-D/2-D.")  # metrics.jl:1670
        super().__init__(dim, celldims, nhalo, device=device, T=T, scheme=conserved_metric_scheme, profile=profile,
                         cache_mode=cache_mode, compute_metrics=compute_metrics,
                         coordinate_system=coordinate_system, basis=basis, window=None, global_offset=goff)
        self.mapping = mapping
        self.state = {"t": float(t), "params": params}
        self._push_mapping()
        L.check(L.lib().cg_grid_update(self._h, ctypes.c_double(float(t)), self._stream()))
        if not self._metrics_disabled and self.cache_mode == "eager":
            self._eval(L.CG_WHAT_CELL | L.CG_WHAT_FACE)

    def _push_mapping(self):
        if isinstance(self.mapping, str):  # generic mapping: compiled once, parameters on every update!
            if not getattr(self, "_source_loaded", False):
                L.check(L.lib().cg_grid_set_mapping_source(self._h, self.mapping.encode()))
                self._source_loaded = True
            prm = self.state["params"]
            arr = np.ascontiguousarray(np.asarray([] if prm is None else prm, dtype=np.float64).ravel())
            L.check(L.lib().cg_grid_set_mapping_params(self._h, arr.ctypes.data_as(ctypes.POINTER(ctypes.c_double)),
                                                       int(arr.size)))
            return
        terms = self.mapping(self.state["params"]) if callable(self.mapping) else self.mapping
        tab = np.ascontiguousarray(np.asarray(terms, dtype=np.float64).reshape(-1, 15))
        L.check(L.lib().cg_grid_set_mapping_terms(self._h, tab.shape[0],
                                                  tab.ctypes.data_as(ctypes.POINTER(ctypes.c_double))))


# ---------------------------------------------------------------------------
# metric cache API (metric_cache_api.jl:16-127)
# ---------------------------------------------------------------------------
def invalidate_cell_metrics(grid):
    if grid._metrics_disabled:
        return None
    L.check(L.lib().cg_grid_invalidate(grid._h, L.CG_WHAT_CELL))


def invalidate_face_metrics(grid):
    if grid._metrics_disabled:
        return None
    L.check(L.lib().cg_grid_invalidate(grid._h, L.CG_WHAT_FACE))


def refresh_cell_metrics(grid, include_halo_region=False):
    grid._require_metric_storage("refresh_cell_metrics!")
    grid._eval(L.CG_WHAT_CELL)
    return grid._cell_payload()


def refresh_face_metrics(grid, include_halo_region=False):
    grid._require_metric_storage("refresh_face_metrics!")
    grid._eval(L.CG_WHAT_FACE)
    return grid._face_payload()


def refresh_metrics(grid):
    """Both caches in ONE fused launch (the reference needs 1 + N kernel launches)."""
    grid._require_metric_storage("refresh_metrics!")
    grid._eval(L.CG_WHAT_CELL | L.CG_WHAT_FACE)
    return grid._cell_payload(), grid._face_payload()


def cell_metrics(grid, refresh=False):
    grid._require_metric_storage("cell_metrics")
    if refresh or not grid._valid()[0]:
        return refresh_cell_metrics(grid)
    return grid._cell_payload()


def face_metrics(grid, refresh=False):
    grid._require_metric_storage("face_metrics")
    if refresh or not grid._valid()[1]:
        return refresh_face_metrics(grid)
    return grid._face_payload()


def cell_metrics_valid(grid):
    return (not grid._metrics_disabled) and grid._valid()[0]


def face_metrics_valid(grid):
    return (not grid._metrics_disabled) and grid._valid()[1]


def update(grid, t=0.0, params=None):
    """update!(grid, t, params): refresh coordinates, store state, invalidate both
    caches; metrics are recomputed on the next access (mapped_grid.jl:551-593)."""
    if params is None:
        params = grid.state.get("params", {})
    grid.state = {"t": float(t), "params": params}
    if isinstance(grid, MappedGrid):
        grid._push_mapping()
    L.check(L.lib().cg_grid_update(grid._h, ctypes.c_double(float(t)), grid._stream()))
    return None


def local_grid(grid, global_cell_domain):
    """local_grid(grid::MappedGrid, global_cell_domain) — mapped_grid.jl:344-379.
    `global_cell_domain`: per-axis 0-based half-open (lo, hi) ranges of interior cells."""
    if not isinstance(grid, MappedGrid):
        raise TypeError("local_grid is defined for MappedGrid")
    for (lo, hi), n in zip(global_cell_domain, grid.celldims):
        if lo < 0 or hi > n or hi <= lo:
            raise ValueError("Requested local grid is outside the parent grid's interior cell domain.")
    celldims = tuple(hi - lo for lo, hi in global_cell_domain)
    goff = tuple(g + lo for g, (lo, _) in zip(grid.global_offset, global_cell_domain))
    return MappedGrid(grid.mapping, grid.state["params"], celldims, grid.nhalo, device=grid.device, T=grid.T,
                      coordinate_system=grid.coordinate_system, basis=grid.basis, cache_mode=grid.cache_mode,
                      conserved_metric_scheme=grid.scheme, profile=grid.profile,
                      global_cell_indices=tuple((o, o + c) for o, c in zip(goff, celldims)), t=grid.state["t"])


def kernel_names_3d():
    """The launches of one 3-D DiscreteGrid cell+face evaluation (bench.py `roofline.kernel`)."""
    return "k_face_xi + k_face_zeta<SWAP> (eta) + k_face_zeta (one cell+face evaluation = 3 launches)"


def fp64_instr_per_cell():
    """FP64 thread instructions per cell of the three marching kernels, FULL profile, Curvature (SASS of the loop
    bodies: xi 736 per warp step and 29 cells, eta / zeta 697 per 31)."""
    return 32.0 * (736.0 / 29.0 + 2.0 * 697.0 / 31.0)


def gcl(grid, low_plane=None):
    """max |I_c| over cell.domain for each physical column (../gcl.jl:20-117 reduced on device).

    `low_plane` (k-slab windows): the previous slab's last conserved plane of the slab axis (`seam_plane` of that
    grid); with it the window's first plane — whose low face lives on the neighbour — is checked as well."""
    face_metrics(grid)
    out = (ctypes.c_double * 3)()
    if low_plane is None:
        L.check(L.lib().cg_grid_gcl_max(grid._h, out, grid._stream()))
    else:
        L.check(L.lib().cg_grid_gcl_max_seam(grid._h, ctypes.c_void_p(low_plane.data_ptr()), out, grid._stream()))
    return tuple(out[c] for c in range(grid.dim))


def seam_plane(grid):
    """The last plane (along the slab axis) of the conserved metrics of the slab axis, as a device tensor view: what
    the next slab's `gcl(grid, low_plane=...)` needs."""
    face_metrics(grid)
    ax = grid.dim - 1
    if grid.profile == "compact":
        return grid._bound[(L.CG_ARR_COMPACT_ACTIVE, ax)][-1]
    return grid._bound[(L.CG_ARR_FACE_CONSERVED, ax)][-1]


# ---------------------------------------------------------------------------
# read-side consumers (stay host code in the reference too; single-index reads)
# ---------------------------------------------------------------------------
def _cell_scale(cs, basis, q):
    # face_geometry.jl:277-300
    N = len(q)
    if cs == "CylindricalCS" and basis == "CartesianBasis":
        return 2 * math.pi * abs(q[0])
    if cs == "AxisymmetricCS{:x}" and basis == "CartesianBasis":
        return 2 * math.pi * abs(q[1])
    if cs == "AxisymmetricCS{:y}" and basis == "CartesianBasis":
        return 2 * math.pi * abs(q[0])
    if cs == "SphericalCS" and basis == "SphericalBasis":
        if N == 1:
            return 4 * math.pi * q[0] ** 2
        if N == 2:
            return 2 * math.pi * q[0] ** 2 * math.sin(q[1])
        return q[0] ** 2 * math.sin(q[1])
    return 1.0


def _face_component_scale(cs, basis, q):
    # face_geometry.jl:302-332
    N = len(q)
    if cs == "SphericalCS" and basis == "SphericalBasis":
        r = abs(q[0])
        if N == 1:
            return [4 * math.pi * r * r]
        s = math.sin(q[1])
        if N == 2:
            return [2 * math.pi * r * r * s, 2 * math.pi * r * s]
        return [r * r * s, r * s, r]
    return [_cell_scale(cs, basis, q)] * N


_LOCS = {"ilo": (0, "lo"), "ihi": (0, "hi"), "jlo": (1, "lo"), "jhi": (1, "hi"), "klo": (2, "lo"), "khi": (2, "hi")}


def face_flux_geometry(grid, idx, loc):
    """face_geometry.jl:393-416; idx = 0-based full cell index (i[,j[,k]])."""
    axis, side = _LOCS[loc]
    if axis >= grid.dim:
        raise ValueError(f"Invalid face location {loc} for a {grid.dim}-D grid")
    N = grid.dim
    fidx = list(idx)
    if side == "lo":
        fidx[axis] -= 1
    tidx = tuple(reversed(fidx))
    q = [float(grid.face_coordinates[axis][c][tidx]) for c in range(N)]
    cons = face_metrics(grid)[axis].conserved[tidx].cpu().numpy().reshape(N, N, order="F")
    row = cons[axis, :]
    scale = _face_component_scale(grid.coordinate_system, grid.basis, q)
    sgn = 1.0 if side == "hi" else -1.0
    mv = np.array([sgn * scale[c] * row[c] for c in range(N)])
    area = float(np.linalg.norm(mv))
    normal = mv / area if area > 0 else np.zeros(N)
    return FaceFluxGeometry(np.array(q), mv, area, normal)


def cellvolume(grid, idx):
    """generation.jl:615-726: scale(q_centroid) * |cell.forward[I].J|."""
    tidx = tuple(reversed(tuple(idx)))
    q = [float(grid.centroid_coordinates[c][tidx]) for c in range(grid.dim)]
    J = float(cell_metrics(grid).forward[tidx][-1])
    return _cell_scale(grid.coordinate_system, grid.basis, q) * abs(J)


# ---------------------------------------------------------------------------
# solver-facing bundles on the device
# ---------------------------------------------------------------------------
def _coordsys_id(cs, basis):
    if cs == "CylindricalCS" and basis == "CartesianBasis":
        return L.CG_CS_CYLINDRICAL
    if cs == "AxisymmetricCS{:x}" and basis == "CartesianBasis":
        return L.CG_CS_AXISYMMETRIC_X
    if cs == "AxisymmetricCS{:y}" and basis == "CartesianBasis":
        return L.CG_CS_AXISYMMETRIC_Y
    if cs == "SphericalCS" and basis == "SphericalBasis":
        return L.CG_CS_SPHERICAL_BASIS
    return L.CG_CS_UNIT


def _refresh_coords_if_scaled(grid, csys):
    if csys != L.CG_CS_UNIT:
        grid._eval(L.CG_WHAT_COORDS)


def face_flux_geometry_arrays(grid, axis):
    """`face_flux_geometry(grid, I, hi side of axis)` (face_geometry.jl:393-416) for every cell of the window in one
    launch.  Returns a FaceFluxGeometry of device tensors: mapped_coordinate = the face coordinates of `axis`,
    metric_vector [N, *cells], area [*cells], normal [N, *cells]; the low face of cell I is minus the entry I - e_axis."""
    torch = _torch()
    if not 0 <= axis < grid.dim:
        raise ValueError(f"Invalid face axis {axis} for a {grid.dim}-D grid")
    face_metrics(grid)
    csys = _coordsys_id(grid.coordinate_system, grid.basis)
    _refresh_coords_if_scaled(grid, csys)
    shape = grid._cell_shape()
    dt, dev = grid.tdtype, f"cuda:{grid.device}"
    mv = torch.empty((grid.dim,) + tuple(shape), dtype=dt, device=dev)
    nrm = torch.empty_like(mv)
    area = torch.empty(tuple(shape), dtype=dt, device=dev)
    L.check(L.lib().cg_grid_face_flux_geometry(grid._h, axis, csys, ctypes.c_void_p(mv.data_ptr()),
                                               ctypes.c_void_p(area.data_ptr()), ctypes.c_void_p(nrm.data_ptr()),
                                               grid._stream()))
    q = grid.face_coordinates[axis] if csys != L.CG_CS_UNIT or grid._coords_enabled else None
    return FaceFluxGeometry(q, mv, area, nrm)


def cellvolumes(grid, include_halo=False):
    """cellvolumes(grid; include_halo) (generation.jl:942-990): scale(q_centroid) * |det F| per cell, device tensor."""
    torch = _torch()
    cell_metrics(grid)
    csys = _coordsys_id(grid.coordinate_system, grid.basis)
    _refresh_coords_if_scaled(grid, csys)
    vol = torch.empty(tuple(grid._cell_shape()), dtype=grid.tdtype, device=f"cuda:{grid.device}")
    L.check(L.lib().cg_grid_cell_volumes(grid._h, csys, ctypes.c_void_p(vol.data_ptr()), grid._stream()))
    if include_halo or tuple(grid.window_cells) != tuple(grid.nfull):
        return vol
    h = grid.nhalo
    return vol[tuple(slice(h, n - h) for n in vol.shape)]

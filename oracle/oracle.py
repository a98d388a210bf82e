"""ORACLE — TEST INFRASTRUCTURE ONLY (ctypes front-end of liboracle.so).

Importers allowed: tests/, __graft_entry__.smoke(), bench.py (cpu_baseline and
--impl reference legs).  The product package never imports this module.

parity: PARTIALLY PINNED (see oracle/metric_oracle.hpp header).
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = {}
# base real type of the restatement: "f64" = liboracle.so (the reference's Float64 arithmetic), "ld" = long double
# (64-bit significand), "quad" = __float128.  The extended builds take and return float64 arrays; tests use them as
# truth for the last-digits question "which FP64 side owns the difference".
_LIBNAME = {"f64": "liboracle.so", "ld": "liboracle_ld.so", "quad": "liboracle_q.so"}

SCHEMES = {"endpoint": 0, "gradient": 1, "curvature": 2, "adtl": 3}


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def lib(precision="f64"):
    if precision not in _LIB:
        path = os.path.join(_HERE, _LIBNAME[precision])
        if not os.path.exists(path):
            build()
        L = ctypes.CDLL(path)
        L.cgo_num_threads.restype = ctypes.c_int
        L.cgo_real_digits.restype = ctypes.c_int
        _LIB[precision] = L
    return _LIB[precision]


def num_threads():
    return lib().cgo_num_threads()


def set_num_threads(n):
    lib().cgo_set_num_threads(int(n))


def _dp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double)) if a is not None else None


def _ip(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)) if a is not None else None


def _scheme(s):
    return SCHEMES[s] if isinstance(s, str) else int(s)


def strip_halo(a, nhalo):
    """discrete_grid.jl:55-61"""
    if nhalo <= 0:
        return a
    sl = tuple(slice(nhalo, s - nhalo) for s in a.shape)
    return a[sl]


def _alloc(dim, nout, what):
    KM, KC = dim * dim + 1, dim * dim
    out = {}
    out["cell_fwd"] = np.zeros((nout, KM)) if what & 1 else None
    out["cell_inv"] = np.zeros((nout, KM)) if what & 1 else None
    out["face_fwd"] = np.zeros((dim, nout, KM)) if what & 2 else None
    out["face_inv"] = np.zeros((dim, nout, KM)) if what & 2 else None
    out["face_cons"] = np.zeros((dim, nout, KC)) if what & 2 else None
    return out


def discrete_eval(nodes, nhalo, scheme="curvature", sample=None, what=3, halo_coords_included=False, precision="f64"):
    """DiscreteGrid(x[,y[,z]], nhalo; halo_coords_included, conserved_metric_scheme).

    Returns dict of packed arrays: cell_fwd/cell_inv [ncell, N*N+1],
    face_fwd/face_inv [N, ncell, N*N+1], face_cons [N, ncell, N*N], plus `nfull`.
    ncell = prod(nfull) (Fortran order) or len(sample)."""
    dim = len(nodes)
    nodes = [np.asarray(a, dtype=np.float64) for a in nodes]
    if halo_coords_included:
        nodes = [strip_halo(a, nhalo) for a in nodes]
    flat = [np.ascontiguousarray(a.ravel(order="F")) for a in nodes]
    nn = np.array(list(nodes[0].shape) + [1] * (3 - dim), dtype=np.int64)
    nfull = tuple(int(s - 1 + 2 * nhalo) for s in nodes[0].shape)
    ntot = int(np.prod(nfull))
    samp = None if sample is None else np.ascontiguousarray(sample, dtype=np.int64)
    nout = ntot if samp is None else len(samp)
    out = _alloc(dim, nout, what)
    xyz = flat + [None] * (3 - dim)
    rc = lib(precision).cgo_discrete_eval(
        ctypes.c_int(dim), _dp(xyz[0]), _dp(xyz[1]), _dp(xyz[2]), _ip(nn), ctypes.c_int(nhalo),
        ctypes.c_int(_scheme(scheme)), ctypes.c_int64(0 if samp is None else len(samp)), _ip(samp),
        ctypes.c_int(what), _dp(out["cell_fwd"]), _dp(out["cell_inv"]), _dp(out["face_fwd"]),
        _dp(out["face_inv"]), _dp(out["face_cons"]))
    if rc != 0:
        raise ValueError("cgo_discrete_eval: unsupported dim/scheme")
    out["nfull"] = nfull
    return out


def mapped_eval(terms, t, celldims, nhalo, scheme="curvature", sample=None, what=3, global_offset=None,
                precision="f64"):
    """MappedGrid(x[,y[,z]], params, celldims, nhalo) with a term-list mapping."""
    dim = len(celldims)
    tab = np.ascontiguousarray(np.asarray(terms, dtype=np.float64).reshape(-1, 15))
    nc = np.array(list(celldims) + [1] * (3 - dim), dtype=np.int64)
    go = np.zeros(3, dtype=np.int64)
    if global_offset is not None:
        go[:dim] = global_offset
    nfull = tuple(int(c + 2 * nhalo) for c in celldims)
    ntot = int(np.prod(nfull))
    samp = None if sample is None else np.ascontiguousarray(sample, dtype=np.int64)
    nout = ntot if samp is None else len(samp)
    out = _alloc(dim, nout, what)
    rc = lib(precision).cgo_mapped_eval(
        ctypes.c_int(dim), ctypes.c_int(tab.shape[0]), _dp(tab), ctypes.c_double(t), _ip(nc),
        ctypes.c_int(nhalo), _ip(go), ctypes.c_int(_scheme(scheme)),
        ctypes.c_int64(0 if samp is None else len(samp)), _ip(samp), ctypes.c_int(what),
        _dp(out["cell_fwd"]), _dp(out["cell_inv"]), _dp(out["face_fwd"]), _dp(out["face_inv"]),
        _dp(out["face_cons"]))
    if rc != 0:
        raise ValueError("cgo_mapped_eval: unsupported dim/scheme")
    out["nfull"] = nfull
    return out


def discrete_coords(nodes, nhalo, which, halo_coords_included=False):
    """which: 0 node, 1 centroid, 2+axis face centre.  Returns [dim, npoints] (Fortran order points)."""
    dim = len(nodes)
    nodes = [np.asarray(a, dtype=np.float64) for a in nodes]
    if halo_coords_included:
        nodes = [strip_halo(a, nhalo) for a in nodes]
    flat = [np.ascontiguousarray(a.ravel(order="F")) for a in nodes] + [None] * (3 - dim)
    nn = np.array(list(nodes[0].shape) + [1] * (3 - dim), dtype=np.int64)
    nfull = [s - 1 + 2 * nhalo for s in nodes[0].shape]
    npts = int(np.prod([n + (1 if which == 0 else 0) for n in nfull]))
    out = np.zeros((dim, npts))
    lib().cgo_discrete_coords(ctypes.c_int(dim), _dp(flat[0]), _dp(flat[1]), _dp(flat[2]), _ip(nn),
                              ctypes.c_int(nhalo), ctypes.c_int(which), _dp(out))
    return out


def mapped_coords(terms, t, celldims, nhalo, which, global_offset=None):
    dim = len(celldims)
    tab = np.ascontiguousarray(np.asarray(terms, dtype=np.float64).reshape(-1, 15))
    nc = np.array(list(celldims) + [1] * (3 - dim), dtype=np.int64)
    go = np.zeros(3, dtype=np.int64)
    if global_offset is not None:
        go[:dim] = global_offset
    npts = int(np.prod([c + 2 * nhalo + (1 if which == 0 else 0) for c in celldims]))
    out = np.zeros((dim, npts))
    lib().cgo_mapped_coords(ctypes.c_int(dim), ctypes.c_int(tab.shape[0]), _dp(tab), ctypes.c_double(t),
                            _ip(nc), ctypes.c_int(nhalo), _ip(go), ctypes.c_int(which), _dp(out))
    return out


def funcmap_eval(dim, map_id, params, t, celldims, nhalo, scheme="curvature", sample=None, what=3, global_offset=None,
                 precision="f64"):
    """MappedGrid with one of the NON-separable analytic test mappings of metric_oracle.hpp (FuncMap ids 1-3)."""
    prm = np.ascontiguousarray(np.asarray(params, dtype=np.float64))
    nc = np.array(list(celldims) + [1] * (3 - dim), dtype=np.int64)
    go = np.zeros(3, dtype=np.int64)
    if global_offset is not None:
        go[:dim] = global_offset
    nfull = tuple(int(c + 2 * nhalo) for c in celldims)
    samp = None if sample is None else np.ascontiguousarray(sample, dtype=np.int64)
    nout = int(np.prod(nfull)) if samp is None else len(samp)
    out = _alloc(dim, nout, what)
    rc = lib(precision).cgo_funcmap_eval(
        ctypes.c_int(dim), ctypes.c_int(map_id), _dp(prm), ctypes.c_double(t), _ip(nc), ctypes.c_int(nhalo), _ip(go),
        ctypes.c_int(_scheme(scheme)), ctypes.c_int64(0 if samp is None else len(samp)), _ip(samp), ctypes.c_int(what),
        _dp(out["cell_fwd"]), _dp(out["cell_inv"]), _dp(out["face_fwd"]), _dp(out["face_inv"]), _dp(out["face_cons"]))
    if rc != 0:
        raise ValueError("cgo_funcmap_eval: unsupported dim/scheme")
    out["nfull"] = nfull
    return out


def funcmap_coords(dim, map_id, params, t, celldims, nhalo, which, global_offset=None):
    prm = np.ascontiguousarray(np.asarray(params, dtype=np.float64))
    nc = np.array(list(celldims) + [1] * (3 - dim), dtype=np.int64)
    go = np.zeros(3, dtype=np.int64)
    if global_offset is not None:
        go[:dim] = global_offset
    npts = int(np.prod([c + 2 * nhalo + (1 if which == 0 else 0) for c in celldims]))
    out = np.zeros((dim, npts))
    lib().cgo_funcmap_coords(ctypes.c_int(dim), ctypes.c_int(map_id), _dp(prm), ctypes.c_double(t), _ip(nc),
                             ctypes.c_int(nhalo), _ip(go), ctypes.c_int(which), _dp(out))
    return out


def gcl_max(face_cons, nfull, nhalo):
    """gcl.jl:20-117: max |I_c| over cell.domain per physical column c."""
    dim = len(nfull)
    cons = np.ascontiguousarray(face_cons, dtype=np.float64)
    nf = np.array(list(nfull) + [1] * (3 - dim), dtype=np.int64)
    out = np.zeros(3)
    lib().cgo_gcl_max(ctypes.c_int(dim), _dp(cons), _ip(nf), ctypes.c_int(nhalo), _dp(out))
    return out[:dim]


def legacy_meg6_3d(nodes, nhalo=5, halo_coords_included=False, symmetric=False):
    """Legacy CurvilinearGrid3D(x,y,z,:meg6[_symmetric]) update! (oracle/legacy_meg6.hpp).
    Returns dict: cell [28, ncell] (J, fwd 9, inv 9, hat 9), edge [3, 19, ncell] (J, inv 9, hat 9),
    centroid [3, ncell], nfull (cell.full dims); arrays are Fortran-ordered over cells."""
    nodes = [np.asarray(a, dtype=np.float64) for a in nodes]
    flat = [np.ascontiguousarray(a.ravel(order="F")) for a in nodes]
    nn = np.array(nodes[0].shape, dtype=np.int64)
    nfull = tuple(int(s - 1 + (0 if halo_coords_included else 2 * nhalo)) for s in nodes[0].shape)
    nc = int(np.prod(nfull))
    cell = np.zeros((28, nc))
    edge = np.zeros((3, 19, nc))
    cent = np.zeros((3, nc))
    lib().cgo_legacy_meg6_3d(_dp(flat[0]), _dp(flat[1]), _dp(flat[2]), _ip(nn), ctypes.c_int(nhalo),
                             ctypes.c_int(int(halo_coords_included)), ctypes.c_int(int(symmetric)), _dp(cell),
                             _dp(edge), _dp(cent))
    return {"cell": cell, "edge": edge, "centroid": cent, "nfull": nfull}


def legacy_meg6_2d(nodes, nhalo=5, halo_coords_included=False):
    """Legacy CurvilinearGrid2D(x,y,:meg6[_symmetric]) update! (oracle/legacy_meg6.hpp; the symmetric scheme is the
    same computation in 2-D).  Returns dict: cell [13, ncell] (J, fwd 4, inv 4, hat 4), edge [2, 9, ncell]
    (J, inv 4, hat 4), centroid [2, ncell], nfull; arrays are Fortran-ordered over cells."""
    nodes = [np.asarray(a, dtype=np.float64) for a in nodes]
    flat = [np.ascontiguousarray(a.ravel(order="F")) for a in nodes]
    nn = np.array(nodes[0].shape, dtype=np.int64)
    nfull = tuple(int(s - 1 + (0 if halo_coords_included else 2 * nhalo)) for s in nodes[0].shape)
    nc = int(np.prod(nfull))
    cell = np.zeros((13, nc))
    edge = np.zeros((2, 9, nc))
    cent = np.zeros((2, nc))
    lib().cgo_legacy_meg6_2d(_dp(flat[0]), _dp(flat[1]), _ip(nn), ctypes.c_int(nhalo),
                             ctypes.c_int(int(halo_coords_included)), _dp(cell), _dp(edge), _dp(cent))
    return {"cell": cell, "edge": edge, "centroid": cent, "nfull": nfull}


def legacy_meg6_1d(x, nhalo=5, halo_coords_included=False):
    """Legacy CurvilinearGrid1D(x, :meg6) update! (oracle/legacy_meg6.hpp).  Returns dict: cell [4, ncell] (J, x_xi, xi_x,
    hatted), edge [3, ncell] (J, xi_x, hatted at i+1/2), centroid [ncell], nfull."""
    x = np.ascontiguousarray(np.asarray(x, dtype=np.float64).ravel())
    nn = np.array([x.size], dtype=np.int64)
    nc = int(x.size - 1 + (0 if halo_coords_included else 2 * nhalo))
    cell, edge, cent = np.zeros((4, nc)), np.zeros((3, nc)), np.zeros(nc)
    lib().cgo_legacy_meg6_1d(_dp(x), _ip(nn), ctypes.c_int(nhalo), ctypes.c_int(int(halo_coords_included)), _dp(cell),
                             _dp(edge), _dp(cent))
    return {"cell": cell, "edge": edge, "centroid": cent, "nfull": (nc,)}

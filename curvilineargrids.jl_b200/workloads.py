"""Synthetic inputs for tests and bench: the reference's own test meshes/mappings.

Every generator restates an input the reference builds in its tests or README
(no RNG; BASELINE.json `configs`).  Arrays are returned
Fortran-ordered (column-major, like Julia) so a raveled buffer matches the
reference's memory layout.

Mappings are *separable term lists* (the form the CUDA MappedGrid path takes
through the C ABI, see include/curvigrid_cuda.h):

    q_c(t, xi, eta, zeta) = sum_m coef_m(t) * f_m1(xi) * f_m2(eta) * f_m3(zeta)
    coef(t) = c0 + c1*t + c2*sin(w*t + ph)
    f in {1, a + b*s, sin(a + b*s), cos(a + b*s)}

Each term is a row of 15 doubles:
    [coord, c0, c1, c2, w, ph, kind1, a1, b1, kind2, a2, b2, kind3, a3, b3]
"""
from __future__ import annotations

import math

import numpy as np

ONE, AFFINE, SIN, COS = 0, 1, 2, 3


# --------------------------------------------------------------------------
# term-list helpers
# --------------------------------------------------------------------------
def term(coord, c0, factors, c1=0.0, c2=0.0, w=0.0, ph=0.0):
    """factors: dict axis -> (kind, a, b); missing axes are the constant 1."""
    row = [float(coord), c0, c1, c2, w, ph]
    for ax in range(3):
        kind, a, b = factors.get(ax, (ONE, 0.0, 0.0))
        row += [float(kind), float(a), float(b)]
    return row


def eval_terms(terms, t, *coords):
    """numpy evaluation of a term list at broadcastable computational coords."""
    terms = np.asarray(terms, dtype=np.float64).reshape(-1, 15)
    dim = len(coords)
    out = [0.0] * dim
    for r in terms:
        c = int(r[0])
        v = r[1] + r[2] * t + r[3] * math.sin(r[4] * t + r[5])
        for ax in range(dim):
            kind, a, b = int(r[6 + 3 * ax]), r[7 + 3 * ax], r[8 + 3 * ax]
            if kind == AFFINE:
                v = v * (a + b * coords[ax])
            elif kind == SIN:
                v = v * np.sin(a + b * coords[ax])
            elif kind == COS:
                v = v * np.cos(a + b * coords[ax])
        out[c] = out[c] + v
    return out


def nodes_from_terms(terms, t, celldims, nhalo, include_halo=False):
    """Node arrays of a MappedGrid: node I (1-based, full) sits at xi = I - nhalo
    (/root/reference/src/grids/unified_grids/generation.jl:209-221)."""
    dim = len(celldims)
    axes = []
    for d in range(dim):
        if include_halo:
            axes.append(np.arange(1 - nhalo, celldims[d] + 1 + nhalo + 1, dtype=np.float64))
        else:
            axes.append(np.arange(1, celldims[d] + 2, dtype=np.float64))
    mesh = np.meshgrid(*axes, indexing="ij")
    out = eval_terms(terms, t, *mesh)
    return [np.asfortranarray(np.broadcast_to(o, mesh[0].shape)) for o in out]


# --------------------------------------------------------------------------
# DiscreteGrid node generators
# --------------------------------------------------------------------------
def rectilinear_nodes_2d(x0, x1, y0, y1, ni, nj):
    """/root/reference/test/unit/test_2d.jl:1-16"""
    xn = np.linspace(x0, x1, ni + 1)
    yn = np.linspace(y0, y1, nj + 1)
    x, y = np.meshgrid(xn, yn, indexing="ij")
    return np.asfortranarray(x), np.asfortranarray(y), (x1 - x0) / ni, (y1 - y0) / nj


def rectilinear_nodes_3d(x0, x1, y0, y1, z0, z1, ni, nj, nk):
    """/root/reference/test/unit/test_3d.jl:17-38"""
    xn = np.linspace(x0, x1, ni + 1)
    yn = np.linspace(y0, y1, nj + 1)
    zn = np.linspace(z0, z1, nk + 1)
    x, y, z = np.meshgrid(xn, yn, zn, indexing="ij")
    return (np.asfortranarray(x), np.asfortranarray(y), np.asfortranarray(z),
            (x1 - x0) / ni, (y1 - y0) / nj, (z1 - z0) / nk)


def wavy_nodes_2d(nx, ny, a0=0.1):
    """/root/reference/test/unit/test_2d.jl:154-181 (nx, ny = node counts)."""
    i = np.arange(nx, dtype=np.float64)[:, None]
    j = np.arange(ny, dtype=np.float64)[None, :]
    x1d = i / (nx - 1)
    y1d = j / (ny - 1)
    w = a0 * np.sin(2 * np.pi * x1d) * np.sin(2 * np.pi * y1d)
    return np.asfortranarray(x1d + w), np.asfortranarray(y1d + w)


def wavy_nodes_3d(ni, nj, nk, L=4.0):
    """/root/reference/test/unit/test_3d.jl:153-178 (ni, nj, nk = node counts)."""
    dx, dy, dz = L / ni, L / nj, L / nk
    i = np.arange(ni, dtype=np.float64)[:, None, None]
    j = np.arange(nj, dtype=np.float64)[None, :, None]
    k = np.arange(nk, dtype=np.float64)[None, None, :]
    sp = lambda a: np.sin(np.pi * a)
    x = -L / 2 + dx * (i + sp(j * dy) * sp(k * dz))
    y = -L / 2 + dy * (j + sp(k * dz) * sp(i * dx))
    z = -L / 2 + dz * (k + sp(i * dx) * sp(j * dy))
    shp = (ni, nj, nk)
    return tuple(np.asfortranarray(np.broadcast_to(a, shp)) for a in (x, y, z))


# --------------------------------------------------------------------------
# MappedGrid term lists
# --------------------------------------------------------------------------
def readme_wavy_2d_terms():
    """/root/reference/README.md:40-43:
    x = xi + 0.15 sin(0.3 xi) cos(0.2 eta); y = eta + 0.10 cos(0.2 xi) sin(0.3 eta)"""
    return [
        term(0, 1.0, {0: (AFFINE, 0.0, 1.0)}),
        term(0, 0.15, {0: (SIN, 0.0, 0.3), 1: (COS, 0.0, 0.2)}),
        term(1, 1.0, {1: (AFFINE, 0.0, 1.0)}),
        term(1, 0.10, {0: (COS, 0.0, 0.2), 1: (SIN, 0.0, 0.3)}),
    ]


def wavy_2d_terms(celldims=(401, 401), L=12.0, ct=0.0):
    """/root/reference/test/unit/test_2d_continuous.jl:21-66; `ct` adds the
    `p.ct*t` drift of test_unified_grid_backends.jl:63-64."""
    ni, nj = celldims
    dx, dy = L / ni, L / nj
    Ax, Ay, n = 0.25 / dx, 0.5 / dy, 0.5
    xmin = ymin = -L / 2
    return [
        term(0, 1.0, {0: (AFFINE, xmin - dx, dx)}),
        term(0, dx * Ax, {1: (SIN, -math.pi * n * dy, math.pi * n * dy)}),
        term(0, 0.0, {}, c1=ct),
        term(1, 1.0, {1: (AFFINE, ymin - dy, dy)}),
        term(1, dy * Ay, {0: (SIN, -math.pi * n * dx, math.pi * n * dx)}),
    ]


def cylindrical_2d_terms(celldims=(41, 41)):
    """/root/reference/test/unit/test_2d_continuous.jl:36-66"""
    ni, nj = celldims
    rmin, rmax = 1.0, 4.0
    tmin, tmax = math.radians(5), math.radians(25)
    dr, dth = (rmax - rmin) / ni, (tmax - tmin) / nj
    return [
        term(0, 1.0, {0: (AFFINE, rmin - dr, dr), 1: (COS, tmin - dth, dth)}),
        term(1, 1.0, {0: (AFFINE, rmin - dr, dr), 1: (SIN, tmin - dth, dth)}),
    ]


def wavy_3d_terms(celldims=(41, 41, 41), L=12.0, time_amp=0.0):
    """/root/reference/test/unit/test_3d_continuous.jl:10-53.  `time_amp` > 0
    multiplies the wave amplitudes by (1 + time_amp*sin(t)) (BASELINE config 4;
    SURVEY.md §8d)."""
    ni, nj, nk = celldims
    d = [L / ni, L / nj, L / nk]
    A = [0.2 / d[0], 0.4 / d[1], 0.6 / d[2]]
    n = 0.5
    mn = -L / 2
    out = []
    for c in range(3):
        a1, a2 = (c + 1) % 3, (c + 2) % 3
        out.append(term(c, 1.0, {c: (AFFINE, mn - d[c], d[c])}))
        amp = d[c] * A[c]
        out.append(term(c, amp, {a1: (SIN, -math.pi * n * d[a1], math.pi * n * d[a1]),
                                 a2: (SIN, -math.pi * n * d[a2], math.pi * n * d[a2])},
                        c2=amp * time_amp, w=1.0))
    return out


def sphere_sector_3d_terms(celldims=(20, 20, 20)):
    """/root/reference/test/unit/test_3d.jl:247-257 (Cartesian x,y,z of an r-theta-phi sector)."""
    ni, nj, nk = celldims
    r0, r1 = 1.0, 3.0
    t0, t1 = math.radians(35), math.radians(180 - 35)
    p0, p1 = math.radians(45), math.radians(360 - 45)
    dr, dt, dp = (r1 - r0) / ni, (t1 - t0) / nj, (p1 - p0) / nk
    R = (AFFINE, r0 - dr, dr)
    return [
        term(0, 1.0, {0: R, 1: (SIN, t0 - dt, dt), 2: (COS, p0 - dp, dp)}),
        term(1, 1.0, {0: R, 1: (SIN, t0 - dt, dt), 2: (SIN, p0 - dp, dp)}),
        term(2, 1.0, {0: R, 1: (COS, t0 - dt, dt)}),
    ]


def spherical_rtp_3d_terms(celldims):
    """/root/reference/test/unit/test_mapped_discrete_face_geometry.jl:196-226:
    perturbed (r, theta, phi) node coordinates of the SphericalCS/SphericalBasis case
    (BASELINE config 5)."""
    ni, nj, nk = celldims
    r0, th0, ph0 = 1.2, 0.45, 0.25
    dr, dth, dph = 0.4 / ni, 0.35 / nj, 0.45 / nk
    ar, ath, aph = 0.012, 0.008, 0.006
    pi = math.pi
    s2i = (SIN, -2 * pi / ni, 2 * pi / ni)
    s2j = (SIN, -2 * pi / nj, 2 * pi / nj)
    s1k = (SIN, -pi / nk, pi / nk)
    return [
        term(0, 1.0, {0: (AFFINE, r0 - dr, dr)}),
        term(0, ar, {1: s2j, 2: s1k}),
        term(1, 1.0, {1: (AFFINE, th0 - dth, dth)}),
        term(1, ath, {0: s2i}),
        term(2, 1.0, {2: (AFFINE, ph0 - dph, dph)}),
        term(2, aph, {0: s2i, 1: s2j}),
    ]


def generic_3d_terms(celldims):
    """/root/reference/test/unit/test_mapped_discrete_face_geometry.jl:228-256"""
    ni, nj, nk = celldims
    q0 = [0.6, 0.8, 1.0]
    dq = [0.7 / ni, 0.6 / nj, 0.5 / nk]
    amp = [0.014, 0.011, 0.009]
    n = [ni, nj, nk]
    pi = math.pi
    s2 = lambda ax: (SIN, -2 * pi / n[ax], 2 * pi / n[ax])
    out = []
    for c in range(3):
        a1, a2 = (c + 1) % 3, (c + 2) % 3
        out.append(term(c, 1.0, {c: (AFFINE, q0[c] - dq[c], dq[c])}))
        out.append(term(c, amp[c], {a1: s2(a1), a2: s2(a2)}))
    return out


def affine_terms(dim, origin, spacing):
    """Rectilinear MappedGrid: q_c = origin_c + (xi_c - 1) * spacing_c."""
    return [term(c, 1.0, {c: (AFFINE, origin[c] - spacing[c], spacing[c])}) for c in range(dim)]


def terms_to_cuda_source(terms):
    """CUDA C++ source of `cg_map` (MappedGrid generic path, cg_grid_set_mapping_source) for a term list: lets any
    separable mapping run through the NVRTC-compiled kernels as well (cross-checks of the two MappedGrid paths)."""
    rows = np.asarray(terms, dtype=np.float64).reshape(-1, 15)
    names = ("xi", "eta", "zeta")
    expr = {0: ["0.0 * xi"], 1: ["0.0 * eta"], 2: ["0.0 * zeta"]}
    for r in rows:
        coef = f"({float(r[1])!r} + {float(r[2])!r} * t + {float(r[3])!r} * sin({float(r[4])!r} * t + {float(r[5])!r}))"
        fac = []
        for d in range(3):
            kind, a, b = int(r[6 + 3 * d]), float(r[7 + 3 * d]), float(r[8 + 3 * d])
            arg = f"({a!r} + {b!r} * {names[d]})"
            if kind == AFFINE:
                fac.append(arg)
            elif kind == SIN:
                fac.append(f"sin{arg}")
            elif kind == COS:
                fac.append(f"cos{arg}")
        expr[int(r[0])].append(" * ".join([coef] + fac))
    body = "\n".join(f"  {q} = {' + '.join(expr[c])};" for c, q in enumerate("xyz"))
    return ("template <class S> __device__ void cg_map(S xi, S eta, S zeta, double t, const double* p, S& x, S& y, S& z) {\n"
            + body + "\n}\n")


# The three NON-separable test mappings (oracle/metric_oracle.hpp FuncMap ids 1-3; tools/julia_goldens.jl states them as
# Julia closures) as CUDA source for MappedGrid(mapping_source, params, ...): cg_grid_set_mapping_source.
_MAP_SIG = "template <class S> __device__ void cg_map(S xi, S eta, S zeta, double t, const double* p, S& x, S& y, S& z)"
GENERIC_TEST_SOURCES = {
    1: _MAP_SIG + """ {
  x = xi + p[0] * sin(p[1] * (eta + 0.5 * zeta) + 0.3 * t);
  y = eta * (1.0 + p[2] * xi) / (1.0 + p[3] * zeta * zeta);
  z = zeta + p[0] * exp(-p[4] * (xi - eta) * (xi - eta)) * cos(t) + sqrt(1.0 + 0.01 * xi * xi);
}""",
    2: _MAP_SIG + """ {
  x = xi + p[0] * sin(p[1] * xi * eta + t);
  y = eta + p[2] * sqrt(1.0 + p[3] * xi * xi + 0.5 * eta * eta);
}""",
    3: _MAP_SIG + """ {
  x = xi + p[0] * sin(p[1] * xi) * exp(-p[2] * xi) + p[3] * t;
}""",
}

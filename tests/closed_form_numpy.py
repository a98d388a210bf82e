"""Vectorised numpy statement of the closed forms the CUDA DiscreteGrid kernels use.

Test infrastructure: it lets the CPU-only suite check the *algorithm* of the
kernels against the oracle (literal nested-AD restatement of the
reference) without a GPU.  The CUDA code does not call this.

3-D conserved metrics use the edge/line formulation:
  * pairs (P, a) with a != face normal f collapse to differences of *edge
    scalars* Psi (potential reconstructed first to the a-face, then along f to
    the edge shared by two a-faces) -> discrete Stokes form, GCL by construction;
  * pairs with a == f collapse to a 4-cell line stencil of per-cell scalars
    (V+, V-, U) along f.
"""
from __future__ import annotations

import numpy as np

LAM = {"endpoint": (0.0, 0.0), "gradient": (0.5, 0.0), "curvature": (0.5, 1.0 / 12.0),
       0: (0.0, 0.0), 1: (0.5, 0.0), 2: (0.5, 1.0 / 12.0)}


def pad_nodes(A, lo, hi):
    """Extend an interior node array by `lo[d]` / `hi[d]` nodes per axis with the
    Interpolations.jl `Line()` rule: value at clamp(I) plus
    sum_d (I_d - c_d) * one-sided difference in the boundary cell (no cross terms)."""
    A = np.asarray(A, dtype=np.float64)
    dim = A.ndim
    n = A.shape
    idx = [np.arange(-lo[d], n[d] + hi[d]) for d in range(dim)]
    cl = [np.clip(ix, 0, n[d] - 1) for d, ix in enumerate(idx)]
    mesh_c = np.meshgrid(*cl, indexing="ij")
    base = A[tuple(mesh_c)]
    out = base.copy()
    for d in range(dim):
        off = idx[d] - cl[d]  # signed distance outside
        nb = np.where(off > 0, cl[d] - 1, np.where(off < 0, cl[d] + 1, cl[d]))
        mesh_n = list(mesh_c)
        shape = [1] * dim
        shape[d] = -1
        mesh_n[d] = np.broadcast_to(nb.reshape(shape), base.shape)
        g = np.where((off > 0).reshape(shape), base - A[tuple(mesh_n)],
                     np.where((off < 0).reshape(shape), A[tuple(mesh_n)] - base, 0.0))
        out = out + off.reshape(shape) * g
    return out


def _avg(A, ax):
    s0 = [slice(None)] * A.ndim
    s1 = [slice(None)] * A.ndim
    s0[ax] = slice(0, -1)
    s1[ax] = slice(1, None)
    return 0.5 * (A[tuple(s1)] + A[tuple(s0)])


def _dif(A, ax):
    s0 = [slice(None)] * A.ndim
    s1 = [slice(None)] * A.ndim
    s0[ax] = slice(0, -1)
    s1[ax] = slice(1, None)
    return A[tuple(s1)] - A[tuple(s0)]


def _sh(A, ax, lo, n):
    s = [slice(None)] * A.ndim
    s[ax] = slice(lo, lo + n)
    return A[tuple(s)]


def _inv3(M):
    """M[..., r, c] -> inverse, det (adjugate form)."""
    a = M
    c00 = a[..., 1, 1] * a[..., 2, 2] - a[..., 1, 2] * a[..., 2, 1]
    c01 = a[..., 1, 2] * a[..., 2, 0] - a[..., 1, 0] * a[..., 2, 2]
    c02 = a[..., 1, 0] * a[..., 2, 1] - a[..., 1, 1] * a[..., 2, 0]
    det = a[..., 0, 0] * c00 + a[..., 0, 1] * c01 + a[..., 0, 2] * c02
    inv = np.empty_like(a)
    inv[..., 0, 0] = c00
    inv[..., 1, 0] = c01
    inv[..., 2, 0] = c02
    inv[..., 0, 1] = a[..., 0, 2] * a[..., 2, 1] - a[..., 0, 1] * a[..., 2, 2]
    inv[..., 1, 1] = a[..., 0, 0] * a[..., 2, 2] - a[..., 0, 2] * a[..., 2, 0]
    inv[..., 2, 1] = a[..., 0, 1] * a[..., 2, 0] - a[..., 0, 0] * a[..., 2, 1]
    inv[..., 0, 2] = a[..., 0, 1] * a[..., 1, 2] - a[..., 0, 2] * a[..., 1, 1]
    inv[..., 1, 2] = a[..., 0, 2] * a[..., 1, 0] - a[..., 0, 0] * a[..., 1, 2]
    inv[..., 2, 2] = a[..., 0, 0] * a[..., 1, 1] - a[..., 0, 1] * a[..., 1, 0]
    return inv / det[..., None, None], det


def _pack_metric(M, J):
    """M[..., r, c], J[...] -> [ncell(F-order), N*N+1] in Julia element layout."""
    N = M.shape[-1]
    cells = M.shape[:-2]
    out = np.empty(cells + (N * N + 1,))
    for c in range(N):
        for r in range(N):
            out[..., r + N * c] = M[..., r, c]
    out[..., N * N] = J
    return out.reshape((-1, N * N + 1), order="F") if False else _f_flat(out)


def _f_flat(A):
    """[cells..., K] -> [ncell (Fortran order over cells), K]"""
    K = A.shape[-1]
    nd = A.ndim - 1
    return np.ascontiguousarray(np.transpose(A, list(range(nd - 1, -1, -1)) + [nd]).reshape(-1, K))


def _pack_cons(M):
    N = M.shape[-1]
    out = np.empty(M.shape[:-2] + (N * N,))
    for c in range(N):
        for r in range(N):
            out[..., r + N * c] = M[..., r, c]
    return _f_flat(out)


def clamp_shift(nfull, nhalo, dim):
    """delta (per axis, per full cell index) from the cell centre to its clamped
    point: Interpolations.gradient evaluates at clamp(xi) (discrete_grid.jl:113-142)."""
    out = []
    for d in range(dim):
        n_nodes = nfull[d] - 2 * nhalo + 1
        xi = np.arange(nfull[d]) + 1 - nhalo + 0.5
        out.append(np.clip(xi, 1.0, float(n_nodes)) - xi)
    return out


# ===========================================================================
# 3-D
# ===========================================================================
def discrete_3d(nodes, nhalo, scheme="curvature"):
    lam1, lam2 = LAM[scheme]
    endpoint = lam1 == 0.0
    x, y, z = [np.asarray(a, dtype=np.float64) for a in nodes]
    h = nhalo
    nfull = tuple(s - 1 + 2 * h for s in x.shape)
    # ext cells = full cells shifted by +1: ext cell e <-> full cell e-1; ext cells 0..nfull+2
    Q = [pad_nodes(a, (h + 1,) * 3, (h + 2,) * 3) for a in (x, y, z)]

    # a-face forms  FA[a][q] = dict(0,u,v,uv) with u=(a+1)%3, v=(a+2)%3
    FA = []
    for a in range(3):
        u, v = (a + 1) % 3, (a + 2) % 3
        per_q = []
        for q in range(3):
            N = Q[q]
            # first differences between neighbouring nodes, THEN averages (round-off rule of the kernels, cg_math.cuh)
            per_q.append({"0": _avg(_avg(N, u), v), "u": _avg(_dif(N, u), v),
                          "v": _avg(_dif(N, v), u), "uv": _dif(_dif(N, u), v)})
        FA.append(per_q)

    # cell Taylor forms T[q][frozenset(axes)]
    T = []
    for q in range(3):
        a = 0
        u, v = 1, 2
        F = FA[a][q]
        T.append({(): _avg(F["0"], a), (u,): _avg(F["u"], a), (v,): _avg(F["v"], a), (u, v): _avg(F["uv"], a),
                  (a,): _avg(_avg(_dif(Q[q], a), u), v), (a, u): _dif(F["u"], a), (a, v): _dif(F["v"], a),
                  (a, u, v): _dif(F["uv"], a)})

    def tc(q, *axes):
        return T[q][tuple(sorted(axes))]

    def full(A):  # ext-cell array -> full cells
        return A[1:1 + nfull[0], 1:1 + nfull[1], 1:1 + nfull[2]]

    out = {"nfull": nfull}

    # ---- cell metrics --------------------------------------------------
    Fc = np.empty(tc(0).shape + (3, 3))
    for q in range(3):
        for d in range(3):
            Fc[..., q, d] = tc(q, d)
    Gc, Jc = _inv3(Fc)
    # forward = gradient at the clamped point
    dsh = clamp_shift(nfull, h, 3)
    D = np.meshgrid(*dsh, indexing="ij")
    Ff = np.empty((nfull) + (3, 3))
    for q in range(3):
        for d in range(3):
            e, f = [a for a in range(3) if a != d]
            Ff[..., q, d] = (full(tc(q, d)) + full(tc(q, d, e)) * D[e] + full(tc(q, d, f)) * D[f]
                             + full(tc(q, d, e, f)) * D[e] * D[f])
    _, Jf = _inv3(Ff)
    out["cell_fwd"] = _pack_metric(Ff, Jf)
    out["cell_inv"] = _pack_metric(full(Gc), Jf)

    # ---- face G (inverse Jacobian reconstructed) -----------------------
    face_fwd, face_inv = [], []
    for f in range(3):
        Fp = np.zeros_like(Fc)  # dF/d(delta_f): column d != f gets q_{fd}
        for q in range(3):
            for d in range(3):
                if d != f:
                    Fp[..., q, d] = tc(q, f, d)
        GFp = Gc @ Fp
        G1 = -GFp @ Gc
        G2 = -2.0 * GFp @ G1
        L = Gc + lam1 * G1 + lam2 * G2
        R = Gc - lam1 * G1 + lam2 * G2
        Gface = 0.5 * (_sh(L, f, 0, L.shape[f] - 1) + _sh(R, f, 1, R.shape[f] - 1))
        Gface = full(Gface)
        Fface, detG = _inv3(Gface)
        _, Jface = _inv3(Fface)
        face_fwd.append(_pack_metric(Fface, Jface))
        face_inv.append(_pack_metric(Gface, Jface))
    out["face_fwd"] = np.stack(face_fwd)
    out["face_inv"] = np.stack(face_inv)

    # ---- conserved Ghat -------------------------------------------------
    def face_form(a, q, *axes):
        u, v = (a + 1) % 3, (a + 2) % 3
        key = "".join("u" if ax == u else "v" for ax in sorted(axes, key=lambda ax: (ax - a) % 3)) or "0"
        return FA[a][q][key]

    def edge_scalar(a, f, b, q1, q2):
        """Psi^{a->f}[(d_b q1) q2] on the edges at (c + a/2 + f/2), indexed by ext cell c
        (valid c: 0..n-2 along a and f)."""
        t = 3 - a - f
        if endpoint:
            p0 = tc(q1, b) * tc(q2)
            s = 0.5 * (_sh(p0, a, 0, p0.shape[a] - 1) + _sh(p0, a, 1, p0.shape[a] - 1))
            return 0.5 * (_sh(s, f, 0, s.shape[f] - 1) + _sh(s, f, 1, s.shape[f] - 1))
        kap = 0.5 * (0.125 - lam2)
        # face polynomial along f on the a-face at ext node plane (c+1) along a
        F1 = lambda *ax: _sh(face_form(a, q1, *ax), a, 1, Q[0].shape[a] - 2)
        F2 = lambda *ax: _sh(face_form(a, q2, *ax), a, 1, Q[0].shape[a] - 2)
        if b == f:
            e0 = F1(f) * F2()
            e1 = F1(f) * F2(f)
            e2 = 0.0 * e0
            pa0 = 2 * tc(q1, a, f) * tc(q2, a)
            pa1 = 2 * tc(q1, a, f) * tc(q2, a, f)
            pa2 = 0.0 * pa0
        else:
            assert b == t
            e0 = F1(t) * F2()
            e1 = F1(t) * F2(f) + F1(f, t) * F2()
            e2 = F1(f, t) * F2(f)
            pa0 = 2 * tc(q1, a, t) * tc(q2, a)
            pa1 = 2 * (tc(q1, a, t) * tc(q2, a, f) + tc(q1, a, f, t) * tc(q2, a))
            pa2 = 2 * tc(q1, a, f, t) * tc(q2, a, f)
        na = pa0.shape[a] - 1
        E0 = e0 - kap * (_sh(pa0, a, 0, na) + _sh(pa0, a, 1, na))
        E1 = e1 - kap * (_sh(pa1, a, 0, na) + _sh(pa1, a, 1, na))
        E2 = e2 - kap * (_sh(pa2, a, 0, na) + _sh(pa2, a, 1, na))
        Wp = 0.5 * E0 + 0.5 * lam1 * E1 + lam2 * E2
        Wm = 0.5 * E0 - 0.5 * lam1 * E1 + lam2 * E2
        nf = Wp.shape[f] - 1
        return _sh(Wp, f, 0, nf) + _sh(Wm, f, 1, nf)

    def edge_diff(a, f, b, q1, q2):
        """Psi(I + f/2 + a/2) - Psi(I + f/2 - a/2) for full cells I."""
        psi = edge_scalar(a, f, b, q1, q2)  # indexed by ext cell c; edge at c+a/2+f/2
        # full cell I = ext cell I+1: hi edge c = I+1, lo edge c = I (along a); along f: c = I+1; along t: I+1
        sl_hi = [slice(1, 1 + nfull[d]) for d in range(3)]
        sl_lo = list(sl_hi)
        sl_lo[a] = slice(0, nfull[a])
        return psi[tuple(sl_hi)] - psi[tuple(sl_lo)]

    def line_term(f, b, q1, q2):
        p0 = tc(q1, b) * tc(q2)
        p1 = tc(q1, b) * tc(q2, f) + tc(q1, b, f) * tc(q2)
        p2 = tc(q1, b, f) * tc(q2, f)
        Vp = 0.5 * p0 + lam1 * p1 + (2 * lam2 + lam1 * lam1) * p2
        Vm = 0.5 * p0 - lam1 * p1 + (2 * lam2 + lam1 * lam1) * p2
        U = 0.5 * p0 + (2 * lam2 - lam1 * lam1) * p2
        n = nfull[f]
        # full cell I = ext I+1
        C = 0.5 * (_sh(Vp - 2 * U, f, 1, n) - _sh(Vp, f, 0, n) + _sh(2 * U - Vm, f, 2, n) + _sh(Vm, f, 3, n))
        sl = [slice(1, 1 + nfull[d]) for d in range(3)]
        sl[f] = slice(None)
        return C[tuple(sl)]

    cons = []
    for f in range(3):
        s, t = (f + 1) % 3, (f + 2) % 3
        Gh = np.empty(nfull + (3, 3))
        for c in range(3):
            q1, q2 = (c + 1) % 3, (c + 2) % 3
            Gh[..., f, c] = edge_diff(t, f, s, q1, q2) - edge_diff(s, f, t, q1, q2)
            Gh[..., s, c] = line_term(f, t, q1, q2) - edge_diff(t, f, f, q1, q2)
            Gh[..., t, c] = edge_diff(s, f, f, q1, q2) - line_term(f, s, q1, q2)
        cons.append(_pack_cons(Gh))
    out["face_cons"] = np.stack(cons)
    return out


# ===========================================================================
# 2-D  
# ===========================================================================
def _inv2(M):
    det = M[..., 0, 0] * M[..., 1, 1] - M[..., 0, 1] * M[..., 1, 0]
    inv = np.empty_like(M)
    inv[..., 0, 0] = M[..., 1, 1]
    inv[..., 0, 1] = -M[..., 0, 1]
    inv[..., 1, 0] = -M[..., 1, 0]
    inv[..., 1, 1] = M[..., 0, 0]
    return inv / det[..., None, None], det


def discrete_2d(nodes, nhalo, scheme="curvature"):
    lam1, lam2 = LAM[scheme]
    x, y = [np.asarray(a, dtype=np.float64) for a in nodes]
    h = nhalo
    nfull = tuple(s - 1 + 2 * h for s in x.shape)
    Q = [pad_nodes(a, (h,) * 2, (h + 1,) * 2) for a in (x, y)]  # ext cells = full + 1 on the high side
    # cell forms: q0, q_xi, q_eta, q_xieta on ext cells
    T = []
    for q in range(2):
        N = Q[q]
        T.append({(): _avg(_avg(N, 0), 1), (0,): _avg(_dif(N, 0), 1), (1,): _dif(_avg(N, 0), 1),
                  (0, 1): _dif(_dif(N, 0), 1)})
    Fc = np.empty(T[0][()].shape + (2, 2))
    for q in range(2):
        for d in range(2):
            Fc[..., q, d] = T[q][(d,)]
    Gc, Jc = _inv2(Fc)
    out = {"nfull": nfull}
    full = lambda A: A[:nfull[0], :nfull[1]]
    dsh = clamp_shift(nfull, h, 2)
    D = np.meshgrid(*dsh, indexing="ij")
    Ff = np.empty(nfull + (2, 2))
    for q in range(2):
        for d in range(2):
            e = 1 - d
            Ff[..., q, d] = full(T[q][(d,)]) + full(T[q][(0, 1)]) * D[e]
    _, Jf = _inv2(Ff)
    out["cell_fwd"] = _pack_metric(Ff, Jf)
    out["cell_inv"] = _pack_metric(full(Gc), Jf)

    face_fwd, face_inv, face_cons = [], [], []
    for f in range(2):
        o = 1 - f
        Fp = np.zeros_like(Fc)
        for q in range(2):
            Fp[..., q, o] = T[q][(0, 1)]
        GFp = Gc @ Fp
        G1 = -GFp @ Gc
        G2 = -2.0 * GFp @ G1
        L = Gc + lam1 * G1 + lam2 * G2
        R = Gc - lam1 * G1 + lam2 * G2
        n = L.shape[f] - 1
        Gface = full(0.5 * (_sh(L, f, 0, n) + _sh(R, f, 1, n)))
        Fface, _ = _inv2(Gface)
        _, Jface = _inv2(Fface)
        face_fwd.append(_pack_metric(Fface, Jface))
        face_inv.append(_pack_metric(Gface, Jface))
        # Ghat = J*inv(F) = [[y_eta, -x_eta], [-y_xi, x_xi]]: every entry is linear along f,
        # so the curvature term vanishes and only lam1 matters.
        H = np.empty_like(Fc)
        H[..., 0, 0] = Fc[..., 1, 1]
        H[..., 0, 1] = -Fc[..., 0, 1]
        H[..., 1, 0] = -Fc[..., 1, 0]
        H[..., 1, 1] = Fc[..., 0, 0]
        Hp = np.zeros_like(Fc)  # d/d(delta_f) of H: only entries built from the o-column of F vary
        if f == 0:  # d/dxi: y_eta' = y_xieta, x_eta' = x_xieta
            Hp[..., 0, 0] = T[1][(0, 1)]
            Hp[..., 0, 1] = -T[0][(0, 1)]
        else:  # d/deta: y_xi' = y_xieta, x_xi' = x_xieta
            Hp[..., 1, 0] = -T[1][(0, 1)]
            Hp[..., 1, 1] = T[0][(0, 1)]
        L = H + lam1 * Hp
        R = H - lam1 * Hp
        Hface = full(0.5 * (_sh(L, f, 0, n) + _sh(R, f, 1, n)))
        face_cons.append(_pack_cons(Hface))
    out["face_fwd"] = np.stack(face_fwd)
    out["face_inv"] = np.stack(face_inv)
    out["face_cons"] = np.stack(face_cons)
    return out

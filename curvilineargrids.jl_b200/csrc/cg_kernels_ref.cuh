// Reference-quality ("gather") CUDA kernels: one thread per cell, every
// quantity re-derived from the padded node arrays.  They are the first correct
// device path and stay in the library as the on-device cross-check for the
// tiled kernels (cg_grid_set_option(CG_OPT_KERNEL, 0)).
//
// What they replace in the reference (all under
// /root/reference/src/grids/unified_grids/):
//   k_pad_nodes            discrete_grid.jl:55-71  (strip halo + Line() extrapolation)
//   k_coords               generation.jl:187-309   (node / centroid / face coordinates)
//   k_metrics{1,2,3}d_ref  metrics.jl:1795-2053    (cell + face fill kernels)
//   k_gcl                  ../gcl.jl:20-117        (GCL residual, max-reduced)
#pragma once
#include <cstdint>

#include "cg_math.cuh"

namespace cg {

struct Dims {
  int dim;
  int nhalo;
  int64_t nw[3];    // window size in full cells (what is stored locally)
  int64_t ne[3];    // ext node array dims = nw + 4 on active axes, 1 otherwise
  int64_t se[3];    // ext node strides
  int64_t w0[3];    // window origin in global full-cell indices
  int64_t gn[3];    // global interior cell counts
  int64_t r0, r1;   // 3-D marching kernels: window cell planes [r0, r1) of axis 2 to evaluate (default 0, nw[2])
};

template <class T>
struct NodeSrc {
  const T* p[3];
  int64_t s0[3];  // global interior node index of element 0 along each axis
  int64_t sn[3];  // nodes available along each axis
  int64_t st[3];  // strides (elements)
};

// ---------------------------------------------------------------------------
// Line() padding: ext node e <-> global interior node m = w0 + e - 1 - nhalo.
// value = A[c] + sum_d (m_d - c_d) * one-sided difference in the boundary cell,
// c = clamp(m) (no cross terms).
// ---------------------------------------------------------------------------
template <class T, int DIM>
__global__ void __launch_bounds__(256) k_pad_nodes(Dims D, NodeSrc<T> S, T* ex, T* ey, T* ez, int e2_lo) {
  // one block per ext row (j, k); threads stride over i.  DIM is a template parameter so that every
  // per-axis quantity lives in a register (a runtime dimension count put them in local memory).
  const int64_t e1 = DIM > 1 ? (int64_t)(blockIdx.x % (unsigned)D.ne[1]) : 0;
  const int64_t e2 = DIM > 2 ? (int64_t)(blockIdx.x / (unsigned)D.ne[1]) + e2_lo : 0;  // e2_lo: first padded plane
  int64_t base_jk = 0, off1 = 0, off2 = 0, nb1 = 0, nb2 = 0;
  if (DIM > 1) {
    const int64_t m = D.w0[1] + e1 - 1 - D.nhalo;
    const int64_t c = m < 0 ? 0 : (m > D.gn[1] ? D.gn[1] : m);
    off1 = m - c;
    nb1 = off1 > 0 ? -S.st[1] : S.st[1];
    base_jk += (c - S.s0[1]) * S.st[1];
  }
  if (DIM > 2) {
    const int64_t m = D.w0[2] + e2 - 1 - D.nhalo;
    const int64_t c = m < 0 ? 0 : (m > D.gn[2] ? D.gn[2] : m);
    off2 = m - c;
    nb2 = off2 > 0 ? -S.st[2] : S.st[2];
    base_jk += (c - S.s0[2]) * S.st[2];
  }
  const int64_t row = e1 * D.se[1] + e2 * D.se[2];
  const T* src[3] = {S.p[0], S.p[1], S.p[2]};
  T* out[3] = {ex, ey, ez};
  for (int64_t i = threadIdx.x; i < D.ne[0]; i += blockDim.x) {
    const int64_t m = D.w0[0] + i - 1 - D.nhalo;
    const int64_t c = m < 0 ? 0 : (m > D.gn[0] ? D.gn[0] : m);
    const int64_t off0 = m - c;
    const int64_t nb0 = off0 > 0 ? -S.st[0] : S.st[0];
    const int64_t base = base_jk + (c - S.s0[0]) * S.st[0];
#pragma unroll
    for (int q = 0; q < DIM; ++q) {
      const T* A = src[q];
      const T v0 = A[base];
      T v = v0;
      if (off0 != 0) {
        const T nb = A[base + nb0];
        v += T(off0) * (off0 > 0 ? (v0 - nb) : (nb - v0));
      }
      if (DIM > 1 && off1 != 0) {
        const T nb = A[base + nb1];
        v += T(off1) * (off1 > 0 ? (v0 - nb) : (nb - v0));
      }
      if (DIM > 2 && off2 != 0) {
        const T nb = A[base + nb2];
        v += T(off2) * (off2 > 0 ? (v0 - nb) : (nb - v0));
      }
      out[q][row + i] = v;
    }
  }
}

// ---------------------------------------------------------------------------
// Coordinates (generation.jl:187-309): node.full copy-out of the ext arrays,
// centroids (mean of the 2^N corners) and high-side face centres (mean of the
// 2^(N-1) face corners).  Output arrays are SoA like the reference.
// ---------------------------------------------------------------------------
template <class T>
struct CoordOut {
  T* node[3];
  T* centroid[3];
  T* face[3][3];  // [axis][coord]
};

template <class T>
__global__ void k_coords(Dims D, const T* ex, const T* ey, const T* ez, CoordOut<T> O) {
  const T* E[3] = {ex, ey, ez};
  int64_t nn[3] = {1, 1, 1};
  for (int d = 0; d < D.dim; ++d) nn[d] = D.nw[d] + 1;
  const int64_t ntot = nn[0] * nn[1] * nn[2];
  for (int64_t lin = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; lin < ntot;
       lin += (int64_t)gridDim.x * blockDim.x) {
    int64_t n[3], r = lin;
    n[0] = r % nn[0];
    r /= nn[0];
    n[1] = r % nn[1];
    n[2] = r / nn[1];
    int64_t eb = 0;
    bool is_cell = true;
    int64_t clin = 0, cst = 1;
    for (int d = 0; d < D.dim; ++d) {
      eb += (n[d] + 1) * D.se[d];
      if (n[d] >= D.nw[d]) is_cell = false;
      clin += n[d] * cst;
      cst *= D.nw[d];
    }
    for (int q = 0; q < D.dim; ++q) {
      if (O.node[q]) O.node[q][lin] = E[q][eb];
    }
    if (!is_cell) continue;
    const int nc = 1 << D.dim;
    for (int q = 0; q < D.dim; ++q) {
      T sum = 0, fs[3] = {0, 0, 0};
      for (int c = 0; c < nc; ++c) {
        int64_t o = eb;
        for (int d = 0; d < D.dim; ++d)
          if ((c >> d) & 1) o += D.se[d];
        T v = E[q][o];
        sum += v;
        for (int a = 0; a < D.dim; ++a)
          if ((c >> a) & 1) fs[a] += v;
      }
      if (O.centroid[q]) O.centroid[q][clin] = sum / T(nc);
      for (int a = 0; a < D.dim; ++a)
        if (O.face[a][q]) O.face[a][q][clin] = fs[a] / T(nc / 2);
    }
  }
}

// ---------------------------------------------------------------------------
// output descriptors (Julia AoS layouts)
// ---------------------------------------------------------------------------
template <class T>
struct MetricOut {
  T* cell_fwd;      // [cell][N*N+1]
  T* cell_inv;
  T* face_fwd[3];   // [cell][N*N+1]
  T* face_inv[3];
  T* face_cons[3];  // [cell][N*N]
  // compact ("solver") profile: J of the cell and the active conserved row per axis
  T* cjac;          // [cell]
  T* cactive[3];    // [cell][N]
};

// shift from the cell centre to the clamped point (Interpolations.gradient
// evaluates at clamp(xi): discrete_grid.jl:113-142).  gi = global full index.
template <class T>
CG_HD T clamp_delta(int64_t gi, int nhalo, int64_t gn) {
  // xi = gi + 1 - nhalo + 1/2, interior nodes span [1, gn+1]
  T xi = T(gi + 1 - nhalo) + T(0.5);
  T lo = T(1), hi = T(gn + 1);
  T c = xi < lo ? lo : (xi > hi ? hi : xi);
  return c - xi;
}

// =============================== 1-D ========================================
// Ghat = inv(F)*det(F) == 1 (metrics.jl:317-323); G = 1/x_xi reconstructed.
template <class T>
__global__ void k_metrics1d_ref(Dims D, const T* ex, MetricOut<T> O, int scheme, int what) {
  const T lam1 = lam1_of<T>(scheme), lam2 = lam2_of<T>(scheme);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < D.nw[0]; i += (int64_t)gridDim.x * blockDim.x) {
    const T* n = ex + (i + 1);
    T F0 = n[1] - n[0], F1 = n[2] - n[1];
    if (what & 1) {
      if (O.cell_fwd) { O.cell_fwd[2 * i] = F0; O.cell_fwd[2 * i + 1] = F0; }
      if (O.cell_inv) { O.cell_inv[2 * i] = T(1) / F0; O.cell_inv[2 * i + 1] = F0; }
      if (O.cjac) O.cjac[i] = F0;
    }
    if (what & 2) {
      // x is linear inside a cell: G' = G'' = 0
      T G = T(0.5) * (T(1) / F0 + T(1) / F1);
      (void)lam1; (void)lam2;
      T Ff = T(1) / G;
      T gh = T(0.5) * ((T(1) / F0) * F0 + (T(1) / F1) * F1);
      if (O.face_fwd[0]) { O.face_fwd[0][2 * i] = Ff; O.face_fwd[0][2 * i + 1] = Ff; }
      if (O.face_inv[0]) { O.face_inv[0][2 * i] = G; O.face_inv[0][2 * i + 1] = Ff; }
      if (O.face_cons[0]) O.face_cons[0][i] = gh;
      if (O.cactive[0]) O.cactive[0][i] = gh;
    }
  }
}

// =============================== 2-D ========================================
template <class T>
struct Cell2 {
  T F[2][2];   // F[q][d] = d x_q / d xi_d at the centre
  T tw[2];     // x_xieta, y_xieta
};
template <class T>
CG_HD Cell2<T> cell2_load(const T* ex, const T* ey, int64_t b, int64_t s1) {
  Cell2<T> c;
  const T* E[2] = {ex, ey};
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    Form4<T> f = face_form(E[q][b], E[q][b + 1], E[q][b + s1], E[q][b + s1 + 1]);
    c.F[q][0] = f.cu;
    c.F[q][1] = f.cv;
    c.tw[q] = f.cuv;
  }
  return c;
}

template <class T>
__global__ void k_metrics2d_ref(Dims D, const T* ex, const T* ey, MetricOut<T> O, int scheme, int what) {
  const T lam1 = lam1_of<T>(scheme), lam2 = lam2_of<T>(scheme);
  const int64_t ntot = D.nw[0] * D.nw[1];
  const int64_t s1 = D.se[1];
  for (int64_t lin = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; lin < ntot;
       lin += (int64_t)gridDim.x * blockDim.x) {
    int64_t i = lin % D.nw[0], j = lin / D.nw[0];
    int64_t b = (i + 1) + (j + 1) * s1;
    Cell2<T> c0 = cell2_load(ex, ey, b, s1);
    T G0[2][2];
    inv2(c0.F, G0);
    if (what & 1) {
      T dl[2] = {clamp_delta<T>(D.w0[0] + i, D.nhalo, D.gn[0]), clamp_delta<T>(D.w0[1] + j, D.nhalo, D.gn[1])};
      T Ff[2][2];
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        Ff[q][0] = c0.F[q][0] + c0.tw[q] * dl[1];
        Ff[q][1] = c0.F[q][1] + c0.tw[q] * dl[0];
      }
      T J = Ff[0][0] * Ff[1][1] - Ff[0][1] * Ff[1][0];
      if (O.cell_fwd) {
        T* o = O.cell_fwd + 5 * lin;
        o[0] = Ff[0][0]; o[1] = Ff[1][0]; o[2] = Ff[0][1]; o[3] = Ff[1][1]; o[4] = J;
      }
      if (O.cell_inv) {
        T* o = O.cell_inv + 5 * lin;
        o[0] = G0[0][0]; o[1] = G0[1][0]; o[2] = G0[0][1]; o[3] = G0[1][1]; o[4] = J;
      }
      if (O.cjac) O.cjac[lin] = J;
    }
    if (what & 2) {
#pragma unroll
      for (int f = 0; f < 2; ++f) {
        const int o_ax = 1 - f;
        Cell2<T> c1 = cell2_load(ex, ey, b + (f == 0 ? 1 : s1), s1);
        // conserved: Ghat = [[y_eta, -x_eta], [-y_xi, x_xi]], linear along f
        T H0[2][2] = {{c0.F[1][1], -c0.F[0][1]}, {-c0.F[1][0], c0.F[0][0]}};
        T H1[2][2] = {{c1.F[1][1], -c1.F[0][1]}, {-c1.F[1][0], c1.F[0][0]}};
        T Hp0[2][2] = {{0, 0}, {0, 0}}, Hp1[2][2] = {{0, 0}, {0, 0}};
        if (f == 0) {
          Hp0[0][0] = c0.tw[1]; Hp0[0][1] = -c0.tw[0];
          Hp1[0][0] = c1.tw[1]; Hp1[0][1] = -c1.tw[0];
        } else {
          Hp0[1][0] = -c0.tw[1]; Hp0[1][1] = c0.tw[0];
          Hp1[1][0] = -c1.tw[1]; Hp1[1][1] = c1.tw[0];
        }
        T Gh[2][2];
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
          for (int cc = 0; cc < 2; ++cc)
            Gh[r][cc] = T(0.5) * ((H0[r][cc] + lam1 * Hp0[r][cc]) + (H1[r][cc] - lam1 * Hp1[r][cc]));
        if (O.face_cons[f]) {
          T* o = O.face_cons[f] + 4 * lin;
          o[0] = Gh[0][0]; o[1] = Gh[1][0]; o[2] = Gh[0][1]; o[3] = Gh[1][1];
        }
        if (O.cactive[f]) {
          T* o = O.cactive[f] + 2 * lin;
          o[0] = Gh[f][0]; o[1] = Gh[f][1];
        }
        if (O.face_fwd[f] || O.face_inv[f]) {
          T G1[2][2];
          inv2(c1.F, G1);
          T Fp0[2][2] = {{0, 0}, {0, 0}}, Fp1[2][2] = {{0, 0}, {0, 0}};
          Fp0[0][o_ax] = c0.tw[0]; Fp0[1][o_ax] = c0.tw[1];
          Fp1[0][o_ax] = c1.tw[0]; Fp1[1][o_ax] = c1.tw[1];
          T L0[2][2], R0[2][2], L1[2][2], R1[2][2];
          ginv_states2(G0, Fp0, lam1, lam2, L0, R0);
          ginv_states2(G1, Fp1, lam1, lam2, L1, R1);
          T Gf[2][2], Ff[2][2];
#pragma unroll
          for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int cc = 0; cc < 2; ++cc) Gf[r][cc] = T(0.5) * (L0[r][cc] + R1[r][cc]);
          inv2(Gf, Ff);
          T J = Ff[0][0] * Ff[1][1] - Ff[0][1] * Ff[1][0];
          if (O.face_fwd[f]) {
            T* o = O.face_fwd[f] + 5 * lin;
            o[0] = Ff[0][0]; o[1] = Ff[1][0]; o[2] = Ff[0][1]; o[3] = Ff[1][1]; o[4] = J;
          }
          if (O.face_inv[f]) {
            T* o = O.face_inv[f] + 5 * lin;
            o[0] = Gf[0][0]; o[1] = Gf[1][0]; o[2] = Gf[0][1]; o[3] = Gf[1][1]; o[4] = J;
          }
        }
      }
    }
  }
}

// =============================== 3-D ========================================
template <class T>
struct Ext3 {
  const T* q[3];
  int64_t st[3];
};

// all 8 Taylor coefficients of one coordinate in the cell whose low-corner node is b;
// index = axis bitmask (bit0 xi, bit1 eta, bit2 zeta)
template <class T>
CG_HD void cell_taylor(const T* A, const int64_t (&st)[3], int64_t b, T (&c)[8]) {
  Form4<T> lo = face_form(A[b], A[b + st[0]], A[b + st[1]], A[b + st[0] + st[1]]);
  const int64_t bz = b + st[2];
  Form4<T> hi = face_form(A[bz], A[bz + st[0]], A[bz + st[1]], A[bz + st[0] + st[1]]);
  c[0] = T(0.5) * (lo.c0 + hi.c0);
  c[1] = T(0.5) * (lo.cu + hi.cu);
  c[2] = T(0.5) * (lo.cv + hi.cv);
  c[3] = T(0.5) * (lo.cuv + hi.cuv);
  // first zeta-difference: the four zeta-edges of the cell (not a difference of plane sums, see cg_math.cuh)
  c[4] = T(0.25) * (((A[bz] - A[b]) + (A[bz + st[0]] - A[b + st[0]])) +
                    ((A[bz + st[1]] - A[b + st[1]]) + (A[bz + st[0] + st[1]] - A[b + st[0] + st[1]])));
  c[5] = hi.cu - lo.cu;
  c[6] = hi.cv - lo.cv;
  c[7] = hi.cuv - lo.cuv;
}

// face form of coordinate A on the plane of nodes starting at b with transverse axes (u,v)
template <class T>
CG_HD Form4<T> plane_form(const T* A, const int64_t (&st)[3], int64_t b, int u, int v) {
  return face_form(A[b], A[b + st[u]], A[b + st[v]], A[b + st[u] + st[v]]);
}

// E^a polynomial (along f) on the a-face on the HIGH side of the cell whose
// low-corner node is c, for the potential (d_b q1) q2   
template <class T>
CG_HD Poly2<T> eface_poly(const Ext3<T>& X, int q1, int q2, int a, int f, int b, int64_t c, T kap) {
  const int t = 3 - a - f;
  const int64_t sa = X.st[a];
  Form4<T> A0 = plane_form(X.q[q1], X.st, c, f, t), A1 = plane_form(X.q[q1], X.st, c + sa, f, t),
           A2 = plane_form(X.q[q1], X.st, c + 2 * sa, f, t);
  Form4<T> B0 = plane_form(X.q[q2], X.st, c, f, t), B1 = plane_form(X.q[q2], X.st, c + sa, f, t),
           B2 = plane_form(X.q[q2], X.st, c + 2 * sa, f, t);
  Form4<T> dA0 = form_dif(A1, A0), dA1 = form_dif(A2, A1), dB0 = form_dif(B1, B0), dB1 = form_dif(B2, B1);
  Poly2<T> e;
  if (b == f) {
    e.e0 = A1.cu * B1.c0 - (T(2) * kap) * (dA0.cu * dB0.c0 + dA1.cu * dB1.c0);
    e.e1 = A1.cu * B1.cu - (T(2) * kap) * (dA0.cu * dB0.cu + dA1.cu * dB1.cu);
    e.e2 = T(0);
  } else {
    e.e0 = A1.cv * B1.c0 - (T(2) * kap) * (dA0.cv * dB0.c0 + dA1.cv * dB1.c0);
    e.e1 = A1.cv * B1.cu + A1.cuv * B1.c0 -
           (T(2) * kap) * (dA0.cv * dB0.cu + dA0.cuv * dB0.c0 + dA1.cv * dB1.cu + dA1.cuv * dB1.c0);
    e.e2 = A1.cuv * B1.cu - (T(2) * kap) * (dA0.cuv * dB0.cu + dA1.cuv * dB1.cu);
  }
  return e;
}

// cell-centre value of the potential (d_b q1) q2 (Endpoint scheme)
template <class T>
CG_HD T potential_centre(const Ext3<T>& X, int q1, int q2, int b, int64_t c) {
  T t1[8], t2[8];
  cell_taylor(X.q[q1], X.st, c, t1);
  cell_taylor(X.q[q2], X.st, c, t2);
  return t1[1 << b] * t2[0];
}

// edge scalar Psi^{a->f}[(d_b q1) q2] at the edge (c + a/2 + f/2)
template <class T>
CG_HD T edge_scalar(const Ext3<T>& X, int scheme, int q1, int q2, int a, int f, int b, int64_t c) {
  if (scheme == ENDPOINT) {
    return T(0.25) * (potential_centre(X, q1, q2, b, c) + potential_centre(X, q1, q2, b, c + X.st[a]) +
                      potential_centre(X, q1, q2, b, c + X.st[f]) +
                      potential_centre(X, q1, q2, b, c + X.st[a] + X.st[f]));
  }
  const T lam1 = lam1_of<T>(scheme), lam2 = lam2_of<T>(scheme);
  const T kap = T(0.5) * (T(0.125) - lam2);
  Poly2<T> e0 = eface_poly(X, q1, q2, a, f, b, c, kap);
  Poly2<T> e1 = eface_poly(X, q1, q2, a, f, b, c + X.st[f], kap);
  return w_plus(e0, lam1, lam2) + w_minus(e1, lam1, lam2);
}
template <class T>
CG_HD T edge_diff(const Ext3<T>& X, int scheme, int q1, int q2, int a, int f, int b, int64_t c) {
  return edge_scalar(X, scheme, q1, q2, a, f, b, c) - edge_scalar(X, scheme, q1, q2, a, f, b, c - X.st[a]);
}

// line scalars of (d_b q1) q2 along f in the cell with low-corner node c
template <class T>
CG_HD LineS<T> line_at(const Ext3<T>& X, int q1, int q2, int f, int b, int64_t c, T lam1, T lam2) {
  const int t = 3 - f - b;
  Form4<T> Alo = plane_form(X.q[q1], X.st, c, b, t), Ahi = plane_form(X.q[q1], X.st, c + X.st[f], b, t);
  Form4<T> Blo = plane_form(X.q[q2], X.st, c, b, t), Bhi = plane_form(X.q[q2], X.st, c + X.st[f], b, t);
  T q1b = T(0.5) * (Alo.cu + Ahi.cu), q1bf = Ahi.cu - Alo.cu;
  T q20 = T(0.5) * (Blo.c0 + Bhi.c0), q2f = Bhi.c0 - Blo.c0;
  return line_scalars(q1b * q20, q1b * q2f + q1bf * q20, q1bf * q2f, lam1, lam2);
}
template <class T>
CG_HD T line_term(const Ext3<T>& X, int scheme, int q1, int q2, int f, int b, int64_t c) {
  const T lam1 = lam1_of<T>(scheme), lam2 = lam2_of<T>(scheme);
  const int64_t sf = X.st[f];
  LineS<T> m1 = line_at(X, q1, q2, f, b, c - sf, lam1, lam2), z0 = line_at(X, q1, q2, f, b, c, lam1, lam2),
           p1 = line_at(X, q1, q2, f, b, c + sf, lam1, lam2), p2 = line_at(X, q1, q2, f, b, c + 2 * sf, lam1, lam2);
  return T(0.5) * ((z0.vp - T(2) * z0.u) - m1.vp + (T(2) * p1.u - p1.vm) + p2.vm);
}

template <class T>
CG_HD void store_metric3(T* o, const T (&M)[3][3], T J) {
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int r = 0; r < 3; ++r) o[r + 3 * c] = M[r][c];
  o[9] = J;
}

template <class T>
__global__ void k_metrics3d_ref(Dims D, const T* ex, const T* ey, const T* ez, MetricOut<T> O, int scheme,
                                int what, int axis_mask = 15 /* bit f: face axis f, bit 3: cell */) {
  Ext3<T> X;
  X.q[0] = ex; X.q[1] = ey; X.q[2] = ez;
  X.st[0] = D.se[0]; X.st[1] = D.se[1]; X.st[2] = D.se[2];
  const T lam1 = lam1_of<T>(scheme), lam2 = lam2_of<T>(scheme);
  const int64_t ntot = D.nw[0] * D.nw[1] * D.nw[2];
  for (int64_t lin = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; lin < ntot;
       lin += (int64_t)gridDim.x * blockDim.x) {
    int64_t i = lin % D.nw[0], r = lin / D.nw[0];
    int64_t j = r % D.nw[1], k = r / D.nw[1];
    const int64_t c = (i + 1) * X.st[0] + (j + 1) * X.st[1] + (k + 1) * X.st[2];
    T Tq[3][8];
#pragma unroll
    for (int q = 0; q < 3; ++q) cell_taylor(X.q[q], X.st, c, Tq[q]);
    T F[3][3], G[3][3];
#pragma unroll
    for (int q = 0; q < 3; ++q)
#pragma unroll
      for (int d = 0; d < 3; ++d) F[q][d] = Tq[q][1 << d];
    inv3(F, G);
    if ((what & 1) && (axis_mask & 8)) {
      T dl[3] = {clamp_delta<T>(D.w0[0] + i, D.nhalo, D.gn[0]), clamp_delta<T>(D.w0[1] + j, D.nhalo, D.gn[1]),
                 clamp_delta<T>(D.w0[2] + k, D.nhalo, D.gn[2])};
      T Ff[3][3];
#pragma unroll
      for (int q = 0; q < 3; ++q)
#pragma unroll
        for (int d = 0; d < 3; ++d) {
          const int e = (d + 1) % 3, g = (d + 2) % 3;
          Ff[q][d] = Tq[q][1 << d] + Tq[q][(1 << d) | (1 << e)] * dl[e] + Tq[q][(1 << d) | (1 << g)] * dl[g] +
                     Tq[q][7] * dl[e] * dl[g];
        }
      T J = det3(Ff);
      if (O.cell_fwd) store_metric3(O.cell_fwd + 10 * lin, Ff, J);
      if (O.cell_inv) store_metric3(O.cell_inv + 10 * lin, G, J);
      if (O.cjac) O.cjac[lin] = J;
    }
    if (!(what & 2)) continue;
    for (int f = 0; f < 3; ++f) {
      if (!((axis_mask >> f) & 1)) continue;
      const int s = (f + 1) % 3, t = (f + 2) % 3;
      // ---- reconstructed inverse Jacobian, forward = inv, J = det(forward) ----
      if (O.face_fwd[f] || O.face_inv[f]) {
        T Tn[3][8];
#pragma unroll
        for (int q = 0; q < 3; ++q) cell_taylor(X.q[q], X.st, c + X.st[f], Tn[q]);
        T F1[3][3], G1[3][3], Fp0[3][3], Fp1[3][3];
        for (int q = 0; q < 3; ++q)
          for (int d = 0; d < 3; ++d) {
            F1[q][d] = Tn[q][1 << d];
            Fp0[q][d] = d == f ? T(0) : Tq[q][(1 << d) | (1 << f)];
            Fp1[q][d] = d == f ? T(0) : Tn[q][(1 << d) | (1 << f)];
          }
        inv3(F1, G1);
        T L0[3][3], R0[3][3], L1[3][3], R1[3][3], Gf[3][3], Ff[3][3];
        ginv_states3(G, Fp0, lam1, lam2, L0, R0);
        ginv_states3(G1, Fp1, lam1, lam2, L1, R1);
        for (int a = 0; a < 3; ++a)
          for (int b = 0; b < 3; ++b) Gf[a][b] = T(0.5) * (L0[a][b] + R1[a][b]);
        inv3(Gf, Ff);
        T J = det3(Ff);
        if (O.face_fwd[f]) store_metric3(O.face_fwd[f] + 10 * lin, Ff, J);
        if (O.face_inv[f]) store_metric3(O.face_inv[f] + 10 * lin, Gf, J);
      }
      // ---- conserved metrics ----
      T Gh[3][3];
      const bool full_rows = O.face_cons[f] != nullptr;
      for (int col = 0; col < 3; ++col) {
        const int q1 = (col + 1) % 3, q2 = (col + 2) % 3;
        Gh[f][col] = edge_diff(X, scheme, q1, q2, t, f, s, c) - edge_diff(X, scheme, q1, q2, s, f, t, c);
        if (full_rows) {
          Gh[s][col] = line_term(X, scheme, q1, q2, f, t, c) - edge_diff(X, scheme, q1, q2, t, f, f, c);
          Gh[t][col] = edge_diff(X, scheme, q1, q2, s, f, f, c) - line_term(X, scheme, q1, q2, f, s, c);
        }
      }
      if (full_rows) {
        T* o = O.face_cons[f] + 9 * lin;
        for (int cc = 0; cc < 3; ++cc)
          for (int rr = 0; rr < 3; ++rr) o[rr + 3 * cc] = Gh[rr][cc];
      }
      if (O.cactive[f]) {
        T* o = O.cactive[f] + 3 * lin;
        o[0] = Gh[f][0]; o[1] = Gh[f][1]; o[2] = Gh[f][2];
      }
    }
  }
}

// ---------------------------------------------------------------------------
// GCL residual (gcl.jl:20-117) over cell.domain of the window, max-reduced.
// Works on the FULL profile (cons[axis][cell][N*N], active entry [axis + N*c])
// or on the COMPACT profile (active[axis][cell][N], entry [c]).
// Non-negative doubles order like their bit patterns, so atomicMax on the bits.
// ---------------------------------------------------------------------------
template <class T, int N, bool COMPACT>
__global__ void __launch_bounds__(256) k_gcl(Dims D, const T* __restrict__ c0, const T* __restrict__ c1,
                                             const T* __restrict__ c2, int64_t lo0, int64_t lo1, int64_t lo2,
                                             int64_t hi0, int64_t hi1, int64_t hi2, const T* __restrict__ low,
                                             unsigned long long* out_bits) {
  // `low` (may be NULL): conserved metrics of the slab axis (the last one) on the cell plane just BELOW the window,
  // [nw0 * nw1][K] — the previous rank's last plane; with it the window-local plane 0 is evaluated too (its low
  // face along the slab axis lives on the neighbour).  N and the profile are template parameters: every per-axis
  // quantity stays in a register (a run-time dimension count put them in local memory: 10.3 ms at 512^3).
  const T* C[3] = {c0, c1, c2};
  constexpr int K = COMPACT ? N : N * N;
  const int64_t lo[3] = {lo0, lo1, lo2}, hi[3] = {hi0, hi1, hi2};
  const int64_t ext0 = hi[0] - lo[0], ext1 = N > 1 ? hi[1] - lo[1] : 1, ext2 = N > 2 ? hi[2] - lo[2] : 1;
  const int64_t cs[3] = {1, D.nw[0], D.nw[0] * D.nw[1]};
  const int64_t nrows = ext1 * ext2;
  double mx[N];
  bool bad[N];  // a non-finite residual: maximum(abs.(I)) of the reference propagates NaN
#pragma unroll
  for (int c = 0; c < N; ++c) {
    mx[c] = 0.0;
    bad[c] = false;
  }
  // a block walks along rows (i fastest: coalesced), rows are distributed over blockIdx.y
  for (int64_t row = blockIdx.y; row < nrows; row += gridDim.y) {
    const int64_t j = N > 1 ? lo[1] + row % ext1 : 0, k = N > 2 ? lo[2] + row / ext1 : 0;
    for (int64_t ii = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; ii < ext0; ii += (int64_t)gridDim.x * blockDim.x) {
      const int64_t idx[3] = {lo[0] + ii, j, k};
      const int64_t cl = idx[0] + idx[1] * cs[1] + idx[2] * cs[2];
      T acc[N];
#pragma unroll
      for (int c = 0; c < N; ++c) acc[c] = T(0);
#pragma unroll
      for (int ax = 0; ax < N; ++ax) {
        const bool seam = ax == N - 1 && idx[ax] == 0;
        const T* hi_p = C[ax] + cl * K;
        const T* lo_p = seam ? low + (cl - idx[ax] * cs[ax]) * K : C[ax] + (cl - cs[ax]) * K;
#pragma unroll
        for (int c = 0; c < N; ++c) {
          const int e = COMPACT ? c : (ax + N * c);
          acc[c] += hi_p[e] - lo_p[e];  // same order as the reference: per column, axes in order
        }
      }
#pragma unroll
      for (int c = 0; c < N; ++c) {
        const double a = fabs((double)acc[c]);
        if (a > mx[c]) mx[c] = a;
        if (!(a <= 1.7976931348623157e308)) bad[c] = true;  // NaN or Inf
      }
    }
  }
#pragma unroll
  for (int c = 0; c < N; ++c) {
    double v = mx[c];
    for (int o = 16; o > 0; o >>= 1) {
      double w = __shfl_down_sync(0xffffffffu, v, o);
      v = w > v ? w : v;
    }
    const unsigned anybad = __ballot_sync(0xffffffffu, bad[c]);
    // the bit pattern of a quiet NaN is larger than that of any finite value or Inf: atomicMax keeps it
    if ((threadIdx.x & 31) == 0) {
      if (anybad) atomicMax(out_bits + c, 0x7ff8000000000000ull);
      else if (v > 0.0) atomicMax(out_bits + c, (unsigned long long)__double_as_longlong(v));
    }
  }
}


// ---------------------------------------------------------------------------
// Solver-facing bundles: face_flux_geometry (face_geometry.jl:393-416) and cellvolume
// (generation.jl:615-726, 942-990) for every cell of the window, SoA.
//   metric vector of the face I+1/2 on `axis` = component_scale(q_face) .* conserved[axis][I][axis, :]
//   (conservation_face_metric_component_scale, face_geometry.jl:302-332), area = |mv|, normal = mv / area;
//   the low face of cell I is the negative of the entry I - e_axis (face_geometry.jl:369-373, 410).
//   volume = conservation_cell_metric_scale(q_centroid) * |det F_cell| (face_geometry.jl:277-300).
// csys: 0 unit scale, 1 CylindricalCS, 2 AxisymmetricCS{:x}, 3 AxisymmetricCS{:y}, 4 SphericalCS + SphericalBasis
// ---------------------------------------------------------------------------
template <class T>
CG_HD T cell_scale_of(int csys, int N, const T (&q)[3]) {
  const T pi = T(3.14159265358979323846);
  switch (csys) {
    case 1: return T(2) * pi * fabs(q[0]);
    case 2: return N == 2 ? T(2) * pi * fabs(q[1]) : T(1);
    case 3: return N == 2 ? T(2) * pi * fabs(q[0]) : T(1);
    case 4:
      if (N == 1) return T(4) * pi * q[0] * q[0];
      if (N == 2) return T(2) * pi * q[0] * q[0] * sin(q[1]);
      return q[0] * q[0] * sin(q[1]);
    default: return T(1);
  }
}
template <class T>
CG_HD void face_scale_of(int csys, int N, const T (&q)[3], T (&sc)[3]) {
  if (csys == 4) {
    const T pi = T(3.14159265358979323846);
    const T r = fabs(q[0]);
    if (N == 1) {
      sc[0] = T(4) * pi * r * r;
    } else if (N == 2) {
      const T st = sin(q[1]);
      sc[0] = T(2) * pi * r * r * st;
      sc[1] = T(2) * pi * r * st;
    } else {
      const T st = sin(q[1]);
      sc[0] = r * r * st;
      sc[1] = r * st;
      sc[2] = r;
    }
    return;
  }
  const T s = cell_scale_of(csys, N, q);
  sc[0] = sc[1] = sc[2] = s;
}
// cons: FULL profile conserved array of `axis` ([cell][N*N], column-major) or the COMPACT active row ([cell][N]);
// fq[c]: face coordinate c of `axis` (may be null when csys == 0); outputs [c][cell] / [cell], any may be null
template <class T>
__global__ void __launch_bounds__(256) k_face_flux_bundle(int N, int axis, int csys, int compact, int64_t ncell,
                                                          const T* __restrict__ cons, const T* fq0, const T* fq1,
                                                          const T* fq2, T* mv, T* area, T* normal) {
  const T* FQ[3] = {fq0, fq1, fq2};
  for (int64_t cl = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; cl < ncell; cl += (int64_t)gridDim.x * blockDim.x) {
    T q[3] = {0, 0, 0}, sc[3] = {1, 1, 1}, v[3] = {0, 0, 0};
    if (csys != 0)
      for (int c = 0; c < N; ++c) q[c] = FQ[c][cl];
    face_scale_of(csys, N, q, sc);
    T a2 = 0;
    for (int c = 0; c < N; ++c) {
      const T row = compact ? cons[cl * N + c] : cons[cl * (N * N) + axis + N * c];
      v[c] = sc[c] * row;
      a2 += v[c] * v[c];
    }
    const T a = sqrt(a2);
    const T ia = a > T(0) ? T(1) / a : T(0);
    for (int c = 0; c < N; ++c) {
      if (mv) mv[c * ncell + cl] = v[c];
      if (normal) normal[c * ncell + cl] = v[c] * ia;
    }
    if (area) area[cl] = a;
  }
}
// J: cell_metrics.forward ([cell][N*N+1], J last) or the COMPACT det F array ([cell])
template <class T>
__global__ void __launch_bounds__(256) k_cell_volumes(int N, int csys, int compact, int64_t ncell,
                                                      const T* __restrict__ Jsrc, const T* cq0, const T* cq1,
                                                      const T* cq2, T* vol) {
  const T* CQ[3] = {cq0, cq1, cq2};
  for (int64_t cl = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; cl < ncell; cl += (int64_t)gridDim.x * blockDim.x) {
    T q[3] = {0, 0, 0};
    if (csys != 0)
      for (int c = 0; c < N; ++c) q[c] = CQ[c][cl];
    const T J = compact ? Jsrc[cl] : Jsrc[cl * (N * N + 1) + N * N];
    vol[cl] = cell_scale_of(csys, N, q) * fabs(J);
  }
}

}  // namespace cg

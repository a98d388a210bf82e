// MappedGrid device path for separable term-list mappings.
//
// The reference evaluates MappedGrid metrics by nested forward-mode AD of the
// user's mapping closures, point by point
// (/root/reference/src/grids/unified_grids/metrics.jl:183-341, 670-902,
// 1536-1651, kernels 1795-2053).  A Julia closure cannot cross a C ABI; the
// mapping arrives as a *separable term list*
//     q_c(t,xi) = sum_m coef_m(t) * f_m1(xi_1) f_m2(xi_2) f_m3(xi_3)
// (every analytic mapping in the reference's tests and README has this form).
// Every operator of the conserved-metric closure graph — the face
// reconstruction E (metrics.jl:141-178), the cell-centred difference
// D_a = E_a(c) - E_a(c - e_a) (:1536-1651) and the outer face reconstruction
// with its first/second derivatives (:784-850) — is linear and acts along ONE
// axis, so on a separable term it acts on a single 1-D factor.  The kernels
// therefore
//   1. build small 1-D operator tables per (term pair, axis) on the device, in
//      FP64, from the analytic derivatives of the factors (k_mapped_tables), and
//   2. assemble all 27 conserved entries of a cell as sums of triple products of
//      table values, and the reconstructed inverse Jacobian from analytic
//      F, dF, d2F (k_mapped_metrics{1,2,3}d).
// This is exact (same values as AD to round-off), costs O(N) preprocessing, and
// leaves the O(N^3) kernel bound by its 856 B/cell of output.
#pragma once
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <vector>

#include "cg_kernels_ref.cuh"
#include "cg_kernels_tiled.cuh"

namespace cg {

constexpr int M_MAXT = 24;  // terms (one extra slot is used for the unit term)
constexpr int M_MAXP = 96;  // term pairs over all columns
enum MOp { OP_V0 = 0, OP_D0 = 1, OP_E0 = 2, OP_FD0 = 3, OP_V1 = 4, OP_E1 = 5, OP_COUNT = 6 };

struct MTerm {
  int coord;
  int kind[3];
  double coef;
  double a[3], b[3];
};
struct MTerms {
  int n;
  MTerm t[M_MAXT];
};
struct MPair {
  int m1, m2, col;
};
struct MPairs {
  int n;
  int cstart[4];  // 3-D: pairs of column c are p[cstart[c] .. cstart[c+1]) (the list is built column by column)
  MPair p[M_MAXP];
  double coef[M_MAXP + 1];  // coef(m1) * coef(m2) at the current time (one entry of slack for the prefetch)
};

// n-th derivative of a factor (kind 0: 1, 1: a+b*s, 2: sin(a+b*s), 3: cos(a+b*s))
__host__ __device__ inline double fac_der(int kind, double a, double b, double s, int n) {
  if (kind == 0) return n == 0 ? 1.0 : 0.0;
  if (kind == 1) return n == 0 ? a + b * s : (n == 1 ? b : 0.0);
  double sn, cs;
#if defined(__CUDA_ARCH__)
  sincos(a + b * s, &sn, &cs);
#else
  sn = std::sin(a + b * s);
  cs = std::cos(a + b * s);
#endif
  double bn = 1.0;
  for (int i = 0; i < n; ++i) bn *= b;
  const int r = n & 3;
  double v;
  if (kind == 2) v = r == 0 ? sn : (r == 1 ? cs : (r == 2 ? -sn : -cs));
  else v = r == 0 ? cs : (r == 1 ? -sn : (r == 2 ? -cs : sn));
  return v * bn;
}

// centre of full cell `idx` along axis d (generation.jl:247-260: xi = Iglobal - nhalo + 1/2)
__host__ __device__ inline double cell_centre(const Dims& D, int d, int64_t idx) {
  return double(D.w0[d] + idx + 1 - D.nhalo) + 0.5;
}

// ---------------------------------------------------------------------------
// 1-D operator tables.  For pair (m1,m2), axis d, variant v (derivative order of
// the q1 factor), h(s) = f_{m1,d}^{(v)}(s) * f_{m2,d}(s):
//   V  = h(s_I)
//   E  = 1/2 [ (h + l1 h' + l2 h'')(s_I) + (h - l1 h' + l2 h'')(s_I + 1) ]      face reconstruction
//   D  = E(s_I) - E(s_I - 1)                                                       cell-centred difference
//   FD = face reconstruction of the function u(s) = E[h](s) - E[h](s-1), with u', u''
// tab[((p*3 + d)*L + I)*OP_COUNT + op]: the six operator values of one position are contiguous (48 B, 16-B
// aligned), so a cell reads them with three 16-byte loads per axis
// ---------------------------------------------------------------------------
template <class T>
__global__ void k_mapped_tables(Dims D, MTerms TM, MPairs PR, double lam1, double lam2, T* tab, int64_t L, T* tabx = nullptr, int64_t Lx = 0) {
  const int64_t ntot = (int64_t)PR.n * 3 * L;
  for (int64_t lin = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; lin < ntot;
       lin += (int64_t)gridDim.x * blockDim.x) {
    const int64_t I = lin % L;
    const int d = (int)((lin / L) % 3);
    const int p = (int)(lin / (3 * L));
    if (d >= D.dim || I >= D.nw[d]) continue;
    const MTerm& t1 = TM.t[PR.p[p].m1];
    const MTerm& t2 = TM.t[PR.p[p].m2];
    const double s0 = cell_centre(D, d, I);
    for (int v = 0; v < 2; ++v) {
      // h^(n) at positions s0-1, s0, s0+1, s0+2 for n = 0..4
      double hn[4][5];
      for (int pos = 0; pos < 4; ++pos) {
        const double s = s0 + double(pos - 1);
        double f[6], g[5];
        for (int n = 0; n < 6; ++n) f[n] = fac_der(t1.kind[d], t1.a[d], t1.b[d], s, n + v);
        for (int n = 0; n < 5; ++n) g[n] = fac_der(t2.kind[d], t2.a[d], t2.b[d], s, n);
        const double binom[5][5] = {{1, 0, 0, 0, 0}, {1, 1, 0, 0, 0}, {1, 2, 1, 0, 0}, {1, 3, 3, 1, 0}, {1, 4, 6, 4, 1}};
        for (int n = 0; n < 5; ++n) {
          double acc = 0.0;
          for (int k = 0; k <= n; ++k) acc += binom[n][k] * f[k] * g[n - k];
          hn[pos][n] = acc;
        }
      }
      // E_n(pos) for pos in {s0-1, s0, s0+1}, n = 0..2
      double En[3][3];
      for (int pos = 0; pos < 3; ++pos)
        for (int n = 0; n < 3; ++n)
          En[pos][n] = 0.5 * ((hn[pos][n] + lam1 * hn[pos][n + 1] + lam2 * hn[pos][n + 2]) +
                              (hn[pos + 1][n] - lam1 * hn[pos + 1][n + 1] + lam2 * hn[pos + 1][n + 2]));
      T* base = tab + ((int64_t)(p * 3 + d) * L + I) * OP_COUNT;
      if (v == 0) {
        base[OP_V0] = T(hn[1][0]);
        base[OP_E0] = T(En[1][0]);
        base[OP_D0] = T(En[1][0] - En[0][0]);
        double u0[3], u1[3];  // u, u', u'' at s0 and s0+1
        for (int n = 0; n < 3; ++n) {
          u0[n] = En[1][n] - En[0][n];
          u1[n] = En[2][n] - En[1][n];
        }
        base[OP_FD0] = T(0.5 * ((u0[0] + lam1 * u0[1] + lam2 * u0[2]) + (u1[0] - lam1 * u1[1] + lam2 * u1[2])));
      } else {
        base[OP_V1] = T(hn[1][0]);
        base[OP_E1] = T(En[1][0]);
      }
    }
    // xi axis, transposed copy tabx[(p * OP_COUNT + op) * Lx + I] (Lx = L rounded up to 16 values, so every row starts
    // on a 128-byte line): a warp whose lanes are consecutive cells reads one operator value of a pair as one
    // contiguous, aligned 256-byte segment (k_mapped_cons_uni)
    if (tabx && d == 0) {
      const T* base = tab + ((int64_t)(p * 3) * L + I) * OP_COUNT;
      for (int o = 0; o < OP_COUNT; ++o) tabx[((int64_t)p * OP_COUNT + o) * Lx + I] = base[o];
    }
  }
}

// factor derivative tables for the analytic Jacobian: ftab[((m*3 + d)*L1 + I)*4 + order], I in [0, nw_d]
// (the four derivative orders of one position are contiguous: two 16-byte loads)
template <class T>
__global__ void k_mapped_factors(Dims D, MTerms TM, T* ftab, int64_t L1, double shift = 0.0) {
  const int64_t ntot = (int64_t)TM.n * 3 * L1;
  for (int64_t lin = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; lin < ntot;
       lin += (int64_t)gridDim.x * blockDim.x) {
    const int64_t I = lin % L1;
    const int d = (int)((lin / L1) % 3);
    const int m = (int)(lin / (3 * L1));
    T* base = ftab + ((int64_t)(m * 3 + d) * L1 + I) * 4;
    if (d >= D.dim) {
      for (int o = 0; o < 4; ++o) base[o] = o == 0 ? T(1) : T(0);
      continue;
    }
    if (I > D.nw[d]) continue;
    const MTerm& t = TM.t[m];
    const double s = cell_centre(D, d, I) + shift;  // shift = 1/2: the high-side face of the cell (ADThomasLombardMetric)
    for (int o = 0; o < 4; ++o) base[o] = T(fac_der(t.kind[d], t.a[d], t.b[d], s, o));
  }
}

// analytic F, dF/d(xi_f), d2F/d(xi_f)^2 at the centre of cell (i0,i1,i2); N = dim
template <class T, int N>
__device__ __forceinline__ void mapped_jac(const MTerms& TM, const T* ftab, int64_t L1, const int64_t (&ix)[3], int f,
                                           T (&F)[N][N], T (&Fp)[N][N], T (&Fpp)[N][N]) {
#pragma unroll
  for (int q = 0; q < N; ++q)
#pragma unroll
    for (int d = 0; d < N; ++d) F[q][d] = Fp[q][d] = Fpp[q][d] = T(0);
  for (int m = 0; m < TM.n; ++m) {
    const int q = TM.t[m].coord;
    T a[3][4];
#pragma unroll
    for (int d = 0; d < 3; ++d)
#pragma unroll
      for (int o = 0; o < 4; ++o) a[d][o] = d < N ? ftab[((int64_t)(m * 3 + d) * L1 + ix[d]) * 4 + o] : (o == 0 ? T(1) : T(0));
    const T c = T(TM.t[m].coef);
#pragma unroll
    for (int d = 0; d < N; ++d) {
      T v0 = c, v1 = c, v2 = c;
#pragma unroll
      for (int ax = 0; ax < N; ++ax) {
        const int o = (ax == d) ? 1 : 0;
        v0 *= a[ax][o];
        v1 *= a[ax][o + (ax == f ? 1 : 0)];
        v2 *= a[ax][o + (ax == f ? 2 : 0)];
      }
#pragma unroll
      for (int qq = 0; qq < N; ++qq)
        if (qq == q) {
          F[qq][d] += v0;
          Fp[qq][d] += v1;
          Fpp[qq][d] += v2;
        }
    }
  }
}

// L / R states of G = inv(F) along one axis with F'' != 0:
//   G' = -G F' G,  G'' = 2 G F' G F' G - G F'' G
template <class T>
__device__ __forceinline__ void mapped_states3(const T (&F)[3][3], const T (&Fp)[3][3], const T (&Fpp)[3][3], T lam1,
                                               T lam2, T (&G)[3][3], T (&Lo)[3][3], T (&Ro)[3][3]) {
  inv3(F, G);
  T A1[3][3], G1[3][3], G2[3][3], A2[3][3], H[3][3];
  mul3(G, Fp, A1);
  mul3(A1, G, G1);
  mul3(A1, G1, G2);
  mul3(G, Fpp, A2);
  mul3(A2, G, H);
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      T s = G[r][c] + lam2 * (T(2) * G2[r][c] - H[r][c]);
      T d = lam1 * G1[r][c];
      Lo[r][c] = s - d;
      Ro[r][c] = s + d;
    }
}
template <class T>
__device__ __forceinline__ void mapped_states2(const T (&F)[2][2], const T (&Fp)[2][2], const T (&Fpp)[2][2], T lam1,
                                               T lam2, T (&G)[2][2], T (&Lo)[2][2], T (&Ro)[2][2]) {
  inv2(F, G);
  T A1[2][2], G1[2][2], G2[2][2], A2[2][2], H[2][2];
  mul2(G, Fp, A1);
  mul2(A1, G, G1);
  mul2(A1, G1, G2);
  mul2(G, Fpp, A2);
  mul2(A2, G, H);
#pragma unroll
  for (int r = 0; r < 2; ++r)
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      T s = G[r][c] + lam2 * (T(2) * G2[r][c] - H[r][c]);
      T d = lam1 * G1[r][c];
      Lo[r][c] = s - d;
      Ro[r][c] = s + d;
    }
}

// =============================== 3-D ========================================
template <class T>
__global__ void __launch_bounds__(128) k_mapped_metrics3d(Dims D, MTerms TM, MPairs PR, const T* __restrict__ tab,
                                                          const T* __restrict__ ftab, int64_t L, MetricOut<T> O,
                                                          int scheme, int what) {
  const T lam1 = lam1_of<T>(scheme), lam2 = lam2_of<T>(scheme);
  const int64_t L1 = L + 1;
  const int64_t ntot = D.nw[0] * D.nw[1] * D.nw[2];
  for (int64_t lin = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; lin < ntot;
       lin += (int64_t)gridDim.x * blockDim.x) {
    int64_t ix[3];
    ix[0] = lin % D.nw[0];
    int64_t r = lin / D.nw[0];
    ix[1] = r % D.nw[1];
    ix[2] = r / D.nw[1];
    T F[3][3], Fp[3][3], Fpp[3][3], G[3][3], Lc[3][3], Rc[3][3];
    if (what & 1) {
      mapped_jac<T, 3>(TM, ftab, L1, ix, 0, F, Fp, Fpp);
      T J = inv3(F, G);
      if (O.cell_fwd) store_metric3(O.cell_fwd + 10 * lin, F, J);
      if (O.cell_inv) store_metric3(O.cell_inv + 10 * lin, G, J);
      if (O.cjac) O.cjac[lin] = J;
    }
    if (!(what & 2)) continue;
    const bool full = O.face_cons[0] != nullptr;
    // ---- conserved metrics: sums over term pairs of triple products of 1-D table values ----
    T Gh[3][3][3];  // [face][row][col]
#pragma unroll
    for (int f = 0; f < 3; ++f)
#pragma unroll
      for (int rr = 0; rr < 3; ++rr)
#pragma unroll
        for (int cc = 0; cc < 3; ++cc) Gh[f][rr][cc] = T(0);
    for (int p = 0; p < PR.n; ++p) {
      T tv[3][OP_COUNT];
#pragma unroll
      for (int d = 0; d < 3; ++d)
#pragma unroll
        for (int o = 0; o < OP_COUNT; ++o) tv[d][o] = tab[((int64_t)(p * 3 + d) * L + ix[d]) * OP_COUNT + o];
      const T C = T(TM.t[PR.p[p].m1].coef * TM.t[PR.p[p].m2].coef);
      const int col = PR.p[p].col;
#pragma unroll
      for (int f = 0; f < 3; ++f) {
        const int s = (f + 1) % 3, t = (f + 2) % 3;
        T act = C * tv[f][OP_E0] * (tv[s][OP_V1] * tv[t][OP_D0] - tv[s][OP_D0] * tv[t][OP_V1]);
        T rs = C * tv[s][OP_V0] * (tv[f][OP_FD0] * tv[t][OP_V1] - tv[f][OP_E1] * tv[t][OP_D0]);
        T rt = C * tv[t][OP_V0] * (tv[f][OP_E1] * tv[s][OP_D0] - tv[f][OP_FD0] * tv[s][OP_V1]);
#pragma unroll
        for (int cc = 0; cc < 3; ++cc)
          if (cc == col) {
            Gh[f][f][cc] += act;
            Gh[f][s][cc] += rs;
            Gh[f][t][cc] += rt;
          }
      }
    }
#pragma unroll
    for (int f = 0; f < 3; ++f) {
      if (O.cactive[f]) {
        T* o = O.cactive[f] + 3 * lin;
        o[0] = Gh[f][f][0]; o[1] = Gh[f][f][1]; o[2] = Gh[f][f][2];
      }
      if (!full) continue;
      T* o = O.face_cons[f] + 9 * lin;
#pragma unroll
      for (int cc = 0; cc < 3; ++cc)
#pragma unroll
        for (int rr = 0; rr < 3; ++rr) o[rr + 3 * cc] = Gh[f][rr][cc];
      // ---- reconstructed inverse Jacobian ----
      T Gn[3][3], Ln[3][3], Rn[3][3], Gf[3][3], Ff[3][3];
      mapped_jac<T, 3>(TM, ftab, L1, ix, f, F, Fp, Fpp);
      mapped_states3(F, Fp, Fpp, lam1, lam2, G, Lc, Rc);
      int64_t in[3] = {ix[0], ix[1], ix[2]};
      in[f] += 1;
      mapped_jac<T, 3>(TM, ftab, L1, in, f, F, Fp, Fpp);
      mapped_states3(F, Fp, Fpp, lam1, lam2, Gn, Ln, Rn);
#pragma unroll
      for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = 0; b < 3; ++b) Gf[a][b] = T(0.5) * (Lc[a][b] + Rn[a][b]);
      inv3(Gf, Ff);
      T J = det3(Ff);
      store_metric3(O.face_fwd[f] + 10 * lin, Ff, J);
      store_metric3(O.face_inv[f] + 10 * lin, Gf, J);
    }
  }
}

// ---------------------------------------------------------------------------
// FULL-profile fast path of the 3-D kernel.  Same arithmetic as k_mapped_metrics3d, organised like the
// DiscreteGrid marching kernels: a warp owns 31 consecutive cells of a j-row (lane 31 is the +xi neighbour)
// and walks along k;
//   * every AoS output goes through the warp's staging buffers and the TMA unit (cg_kernels_tiled.cuh);
//   * the L / R states of the reconstructed inverse Jacobian are computed once per cell and axis where the
//     neighbour is in reach: the +xi neighbour's R state comes by warp shuffle, the +zeta neighbour's state is
//     the next step's own state (carried); only the +eta neighbour is recomputed;
//   * one pass over the mapping terms per cell position accumulates F and the f-derivatives of every axis
//     needed there; the terms are visited coordinate by coordinate (MCoordTerms), no run-time selects.
// ---------------------------------------------------------------------------
#ifndef CG_MAPPED_UNROLL_P
#define CG_MAPPED_UNROLL_P 2
#endif
#ifndef CG_MAPPED_UNROLL_T
#define CG_MAPPED_UNROLL_T 2
#endif
constexpr int MAPPED_UNROLL_P = CG_MAPPED_UNROLL_P, MAPPED_UNROLL_T = CG_MAPPED_UNROLL_T;
// vector loads of N consecutive table values: 16-byte pieces when N * sizeof(T) is a multiple of 16 (FP64 rows, FP32
// factor rows), 8-byte pieces otherwise (FP32 operator rows: 24 B); the address is aligned to the piece
template <class T, int N>
__device__ __forceinline__ void ldg_vec(const T* __restrict__ p, T (&v)[N]) {
  if constexpr (sizeof(T) == 8) {
    static_assert(N % 2 == 0, "table rows are 16-byte multiples");
#pragma unroll
    for (int k = 0; k < N / 2; ++k) {
      const double2 w = __ldg(reinterpret_cast<const double2*>(p) + k);
      v[2 * k] = w.x; v[2 * k + 1] = w.y;
    }
  } else if constexpr (N % 4 == 0) {
#pragma unroll
    for (int k = 0; k < N / 4; ++k) {
      const float4 w = __ldg(reinterpret_cast<const float4*>(p) + k);
      v[4 * k] = w.x; v[4 * k + 1] = w.y; v[4 * k + 2] = w.z; v[4 * k + 3] = w.w;
    }
  } else {
    static_assert(N % 2 == 0, "table rows are 8-byte multiples");
#pragma unroll
    for (int k = 0; k < N / 2; ++k) {
      const float2 w = __ldg(reinterpret_cast<const float2*>(p) + k);
      v[2 * k] = w.x; v[2 * k + 1] = w.y;
    }
  }
}

struct MCoordTerms {
  int cnt[3];
  int idx[3][M_MAXT];
  int flat[M_MAXT + 1];     // the same terms as one sequence ordered by coordinate (software pipelining of the loads)
  double coef[M_MAXT + 1];  // their time-dependent coefficients in that order
};
inline MCoordTerms coord_terms(const MTerms& TM) {
  MCoordTerms C;
  for (int q = 0; q < 3; ++q) C.cnt[q] = 0;
  for (int m = 0; m < TM.n; ++m) {
    const int q = TM.t[m].coord;
    if (q >= 0 && q < 3) C.idx[q][C.cnt[q]++] = m;
  }
  int s = 0;
  for (int q = 0; q < 3; ++q)
    for (int mm = 0; mm < C.cnt[q]; ++mm) {
      C.flat[s] = C.idx[q][mm];
      C.coef[s++] = TM.t[C.idx[q][mm]].coef;
    }
  C.flat[s] = s > 0 ? C.flat[s - 1] : 0;  // sentinel: the prefetch of "one past the end" stays in range
  C.coef[s] = 0.0;
  return C;
}
// F and, for every axis f with (AXES >> f) & 1, dF/d(xi_f) and d2F/d(xi_f)^2 at the centre of cell ix
template <class T, int AXES>
__device__ __forceinline__ void mapped_jac_multi(const MTerms& TM, const MCoordTerms& CT, const T* __restrict__ ftab,
                                                 int64_t L1, const int64_t (&ix)[3], T (&F)[3][3], T (&Fp)[3][3][3],
                                                 T (&Fpp)[3][3][3]) {
  // software pipeline over the flat term sequence: the table rows of the NEXT term are requested before the current
  // term is accumulated (8 warps per SM cannot hide an L1 round trip per term otherwise)
  T a[3][4], an[3][4];
  T c = T(CT.coef[0]), cn;
#pragma unroll
  for (int d = 0; d < 3; ++d) ldg_vec<T, 4>(ftab + ((int64_t)(CT.flat[0] * 3 + d) * L1 + ix[d]) * 4, a[d]);
  int s = 0;
#pragma unroll
  for (int q = 0; q < 3; ++q) {
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      F[q][d] = T(0);
#pragma unroll
      for (int f = 0; f < 3; ++f) Fp[f][q][d] = Fpp[f][q][d] = T(0);
    }
#pragma unroll(MAPPED_UNROLL_T)
    for (int mm = 0; mm < CT.cnt[q]; ++mm) {
      ++s;
      const int mnext = CT.flat[s];  // sentinel past the end
      cn = T(CT.coef[s]);
#pragma unroll
      for (int d = 0; d < 3; ++d) ldg_vec<T, 4>(ftab + ((int64_t)(mnext * 3 + d) * L1 + ix[d]) * 4, an[d]);
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        // orders: o_ax = (ax == d) for F; + (ax == f) for dF/dxi_f; + 2 (ax == f) for the second derivative
        const int e = (d + 1) % 3, g = (d + 2) % 3;
        const T ce = c * a[e][0], cg_ = c * a[g][0];
        F[q][d] += ce * a[g][0] * a[d][1];
#pragma unroll
        for (int f = 0; f < 3; ++f) {
          if (!((AXES >> f) & 1)) continue;
          if (f == d) {
            Fp[f][q][d] += ce * a[g][0] * a[d][2];
            Fpp[f][q][d] += ce * a[g][0] * a[d][3];
          } else if (f == e) {
            Fp[f][q][d] += cg_ * a[e][1] * a[d][1];
            Fpp[f][q][d] += cg_ * a[e][2] * a[d][1];
          } else {
            Fp[f][q][d] += ce * a[g][1] * a[d][1];
            Fpp[f][q][d] += ce * a[g][2] * a[d][1];
          }
        }
      }
      c = cn;
#pragma unroll
      for (int d = 0; d < 3; ++d)
#pragma unroll
        for (int o = 0; o < 4; ++o) a[d][o] = an[d][o];
    }
  }
}
// L / R states from G = inv(F) already known
template <class T>
__device__ __forceinline__ void mapped_states3_g(const T (&G)[3][3], const T (&Fp)[3][3], const T (&Fpp)[3][3], T lam1,
                                                 T lam2, T (&Lo)[3][3], T (&Ro)[3][3]) {
  T A1[3][3], G1[3][3], G2[3][3], A2[3][3], H[3][3];
  mul3(G, Fp, A1);
  mul3(A1, G, G1);
  mul3(A1, G1, G2);
  mul3(G, Fpp, A2);
  mul3(A2, G, H);
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      T s = G[r][c] + lam2 * (T(2) * G2[r][c] - H[r][c]);
      T d = lam1 * G1[r][c];
      Lo[r][c] = s - d;
      Ro[r][c] = s + d;
    }
}

constexpr int MAPPED_OUT = 31;
#ifndef CG_MWY
#define CG_MWY 1
#endif
// one warp per block: row, plane and every staging / TMA address are uniform (see ZWY in cg_kernels_tiled.cuh)
constexpr int MWY = CG_MWY;
// Staging in four phases per step (cell arrays, then one face axis at a time, all in the same three buffers): 7.5 KB
// per warp instead of 27.5 KB, which leaves ~190 KB of L1 for the operator-table rows (they missed L1 on every step
// with 220 KB of shared memory per SM: 44.6 -> 41.2 ms with two phases).  The wait before a phase's first
// staging store follows the arithmetic of that phase, so the TMA unit has had time to read the previous one.
template <class T> constexpr int mapped_stage_vals = 2 * stage_len<T, 10> + stage_len<T, 9>;
#ifndef CG_MAPPED_MINB
#define CG_MAPPED_MINB 8
#endif
template <class T>
__global__ void __launch_bounds__(32 * MWY, CG_MAPPED_MINB / MWY) k_mapped_metrics3d_staged(Dims D, MTerms TM, MCoordTerms CT, MPairs PR,
                                                                  const T* __restrict__ tab, const T* __restrict__ ftab,
                                                                  int64_t L, MetricOut<T> O, int scheme, int kchunk) {
  extern __shared__ __align__(128) unsigned char cg_dyn_smem[];
  const T lam1 = lam1_of<T>(scheme), lam2 = lam2_of<T>(scheme);
  const int64_t L1 = L + 1;
  const int lane = threadIdx.x;
  const int64_t nw0 = D.nw[0], nw1 = D.nw[1], nw2 = D.nw[2];
  const int64_t j = MWY == 1 ? (int64_t)blockIdx.y : (int64_t)blockIdx.y * blockDim.y + threadIdx.y;
  if (j >= nw1) return;  // whole warp
  const int64_t i0 = (int64_t)blockIdx.x * MAPPED_OUT;
  const int nvalid = (int)(nw0 - i0 < MAPPED_OUT ? nw0 - i0 : MAPPED_OUT);
  const bool out_ok = lane < nvalid;
  // lane 31 (and lanes past the row end) evaluate the cell nw0, i.e. the +xi neighbour of the last cell: the
  // factor tables hold nw + 1 entries per axis
  const int64_t i = i0 + lane < nw0 ? i0 + lane : nw0;
  const int64_t k0 = (int64_t)blockIdx.z * kchunk;
  const int64_t k1 = k0 + kchunk < nw2 ? k0 + kchunk : nw2;
  T* const sg = reinterpret_cast<T*>(cg_dyn_smem) + (MWY == 1 ? 0 : threadIdx.y) * mapped_stage_vals<T>;

  // carried along zeta: F, G = inv F, det F and the zeta-axis L state of the cell of the next step
  T Fc[3][3], Gc[3][3], Jc, Lz[3][3], Fp0[3][3], Fpp0[3][3], Fp1[3][3], Fpp1[3][3];
  auto cell_state = [&](int64_t m, T (&F)[3][3], T (&G)[3][3], T& J, T (&L2)[3][3], T (&R2)[3][3], T (&P0)[3][3],
                        T (&PP0)[3][3], T (&P1)[3][3], T (&PP1)[3][3]) {
    const int64_t ix[3] = {i, j, m};
    T Fp[3][3][3], Fpp[3][3][3];
    mapped_jac_multi<T, 7>(TM, CT, ftab, L1, ix, F, Fp, Fpp);
    J = inv3(F, G);
    mapped_states3_g(G, Fp[2], Fpp[2], lam1, lam2, L2, R2);
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        P0[r][c] = Fp[0][r][c]; PP0[r][c] = Fpp[0][r][c];
        P1[r][c] = Fp[1][r][c]; PP1[r][c] = Fpp[1][r][c];
      }
  };
  {
    T R2[3][3];
    cell_state(k0, Fc, Gc, Jc, Lz, R2, Fp0, Fpp0, Fp1, Fpp1);
  }
  for (int64_t m = k0; m < k1; ++m) {
    const int64_t ix[3] = {i, j, m};
    const int64_t cl0 = i0 + nw0 * (j + nw1 * m);
    bulk_wait_read();  // the TMA unit has finished reading last step's staging buffers (lanes > 0 own no groups)
    __syncwarp();
    T* const b10 = sg;                          // two Metric buffers and one ConservedMetric buffer, four phases
    T* const b9 = sg + 2 * stage_len<T, 10>;
    {
      Stage<T, 10> Scf(O.cell_fwd, cl0, b10), Sci(O.cell_inv, cl0, b10 + stage_len<T, 10>);
      put10(Scf.slot(lane), Fc, Jc);  // every lane stages (no branch); slots >= nvalid are never sent
      put10(Sci.slot(lane), Gc, Jc);
      fence_async_smem();
      __syncwarp();
      Scf.issue(nvalid, lane);
      Sci.issue(nvalid, lane);
      if (lane == 0) bulk_commit();
    }
    // ---- conserved metrics: sums over term pairs of triple products of 1-D table values ----
    T Gh[3][3][3];  // [face][row][col]
#pragma unroll
    for (int f = 0; f < 3; ++f)
#pragma unroll
      for (int rr = 0; rr < 3; ++rr)
#pragma unroll
        for (int cc = 0; cc < 3; ++cc) Gh[f][rr][cc] = T(0);
    // the pair list is ordered by column: three loops with a compile-time column (no predicated accumulation),
    // software-pipelined over the flat list: the rows of pair p + 1 are requested before pair p is accumulated
    {
      const int np = PR.n;
      int64_t ixc[3];
#pragma unroll
      for (int d = 0; d < 3; ++d) ixc[d] = ix[d] < D.nw[d] ? ix[d] : D.nw[d] - 1;
      T tv[3][OP_COUNT], tn[3][OP_COUNT];
      T C = T(PR.coef[0]), Cn;
#pragma unroll
      for (int d = 0; d < 3; ++d) ldg_vec<T, OP_COUNT>(tab + ((int64_t)d * L + ixc[d]) * OP_COUNT, tv[d]);
#pragma unroll
      for (int cc = 0; cc < 3; ++cc) {
#pragma unroll(MAPPED_UNROLL_P)
        for (int p = PR.cstart[cc]; p < PR.cstart[cc + 1]; ++p) {
          const int pn = p + 1 < np ? p + 1 : p;  // the last prefetch re-reads the last pair
          Cn = T(PR.coef[p + 1]);
#pragma unroll
          for (int d = 0; d < 3; ++d) ldg_vec<T, OP_COUNT>(tab + ((int64_t)(pn * 3 + d) * L + ixc[d]) * OP_COUNT, tn[d]);
#pragma unroll
          for (int f = 0; f < 3; ++f) {
            const int s = (f + 1) % 3, t = (f + 2) % 3;
            Gh[f][f][cc] += C * tv[f][OP_E0] * (tv[s][OP_V1] * tv[t][OP_D0] - tv[s][OP_D0] * tv[t][OP_V1]);
            Gh[f][s][cc] += C * tv[s][OP_V0] * (tv[f][OP_FD0] * tv[t][OP_V1] - tv[f][OP_E1] * tv[t][OP_D0]);
            Gh[f][t][cc] += C * tv[t][OP_V0] * (tv[f][OP_E1] * tv[s][OP_D0] - tv[f][OP_FD0] * tv[s][OP_V1]);
          }
          C = Cn;
#pragma unroll
          for (int d = 0; d < 3; ++d)
#pragma unroll
            for (int o = 0; o < OP_COUNT; ++o) tv[d][o] = tn[d][o];
        }
      }
    }
    // ---- reconstructed inverse Jacobian on the three faces ----
#pragma unroll
    for (int f = 0; f < 3; ++f) {
      T Gf[3][3], Ff[3][3];
      if (f == 0) {  // +xi neighbour: its R state by shuffle
        T L0[3][3], R0[3][3];
        mapped_states3_g(Gc, Fp0, Fpp0, lam1, lam2, L0, R0);
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
          for (int b = 0; b < 3; ++b) Gf[a][b] = T(0.5) * (L0[a][b] + __shfl_down_sync(FULLMASK, R0[a][b], 1));
      } else if (f == 1) {  // +eta neighbour: recomputed
        T L1s[3][3], R1s[3][3], Fn[3][3], Gn[3][3], Ln[3][3], Rn[3][3], Fpn[3][3][3], Fppn[3][3][3];
        mapped_states3_g(Gc, Fp1, Fpp1, lam1, lam2, L1s, R1s);
        const int64_t in[3] = {i, j + 1, m};
        mapped_jac_multi<T, 2>(TM, CT, ftab, L1, in, Fn, Fpn, Fppn);
        inv3(Fn, Gn);
        mapped_states3_g(Gn, Fpn[1], Fppn[1], lam1, lam2, Ln, Rn);
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
          for (int b = 0; b < 3; ++b) Gf[a][b] = T(0.5) * (L1s[a][b] + Rn[a][b]);
      } else {  // +zeta neighbour: the cell of the next step; its F, G, L state are carried
        T Ln[3][3], Rn[3][3];
        cell_state(m + 1, Fc, Gc, Jc, Ln, Rn, Fp0, Fpp0, Fp1, Fpp1);
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
          for (int b = 0; b < 3; ++b) {
            Gf[a][b] = T(0.5) * (Lz[a][b] + Rn[a][b]);
            Lz[a][b] = Ln[a][b];
          }
      }
      T J = rcp(inv3(Gf, Ff));  // det(F_face) = 1 / det(G_face)
      // one phase per face axis: wait until the TMA unit has read the previous phase, stage, hand over
      bulk_wait_read();
      __syncwarp();
      Stage<T, 10> Sff(O.face_fwd[f], cl0, b10), Sfi(O.face_inv[f], cl0, b10 + stage_len<T, 10>);
      Stage<T, 9> Sfc(O.face_cons[f], cl0, b9);
      put9(Sfc.slot(lane), Gh[f]);
      put10(Sff.slot(lane), Ff, J);
      put10(Sfi.slot(lane), Gf, J);
      fence_async_smem();
      __syncwarp();
      Sff.issue(nvalid, lane);
      Sfi.issue(nvalid, lane);
      Sfc.issue(nvalid, lane);
      if (lane == 0) bulk_commit();
    }
  }
  if (lane == 0) bulk_wait_read();  // shared memory must outlive the TMA reads
}

template <class T>
inline bool launch_mapped_metrics3d_staged(const Dims& D, const MTerms& TM, const MPairs& PR, const T* tab, const T* ftab,
                                           int64_t L, const MetricOut<T>& O, int scheme, cudaStream_t st) {
  const unsigned gx = (unsigned)((D.nw[0] + MAPPED_OUT - 1) / MAPPED_OUT), gy = (unsigned)((D.nw[1] + MWY - 1) / MWY);
  // 8 warps per SM: chunks along k so that the last wave is (nearly) full (7.5 waves -> 15 at 512^3)
#ifdef CG_MAPPED_NCHUNK  // experiment: fixed number of chunks along k
  const int kchunk = (int)((D.nw[2] + CG_MAPPED_NCHUNK - 1) / CG_MAPPED_NCHUNK);
#else
  const int kchunk = pick_chunk((int64_t)gx * gy, D.nw[2], 148LL * CG_MAPPED_MINB / MWY, 32, 0.4);
#endif
  const int smem = MWY * mapped_stage_vals<T> * (int)sizeof(T);
  if (cudaFuncSetAttribute(k_mapped_metrics3d_staged<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return false;
  k_mapped_metrics3d_staged<T><<<dim3(gx, gy, (unsigned)((D.nw[2] + kchunk - 1) / kchunk)), dim3(32, MWY), smem, st>>>(
      D, TM, coord_terms(TM), PR, tab, ftab, L, O, scheme, kchunk);
  return true;
}

// ---------------------------------------------------------------------------
// Round 2: the FULL-profile kernel with the WARP-UNIFORM part of every product factored out.
//
// A warp owns 31 cells of one j-row and walks along k, so of the three 1-D table values that every product
// multiplies, only the xi one differs between lanes; the eta value (index j) and the zeta value (index m) are the
// same for all 32 lanes.  k_mapped_metrics3d_staged made every lane load all three rows and multiply them
// (12 pairs x 9 sixteen-byte loads and ~42 FP64 operations per pair and cell; 6 terms x 6 loads and ~48 operations
// per term for the Jacobian derivatives; the kernel sat at 8 warps/SM waiting for those loads: 33 % FP64 pipe,
// 2.4 long-scoreboard stalls per issue).  Here, once per warp step,
//   * lane p multiplies the eta and zeta rows of pair p into the 15 coefficients U_p that the nine conserved entries
//     need (pair_uniform: 24 operations on one lane instead of ~27 on every lane), and
//   * lane s multiplies the eta and zeta factor-derivative rows of term s into the 7 products c f_eta^(a) f_zeta^(b)
//     that F, dF/dxi_f and d2F/dxi_f^2 need, for the three cell positions of a step (term_uniform),
// into a small shared-memory table; every lane then reads the coefficients as 16-byte broadcasts and does ONE
// multiply-add per (xi value, coefficient): 15 per pair, 9-12 per term and position.  Per-lane global loads are the
// xi rows only; they are the same for every step of the march AND for every j, so the blocks are ordered j-fastest:
// the warps resident on an SM work on the same xi tile and share its rows in L1 (24 KB instead of 24 KB per warp).
// The zeta rows of the next step are requested one phase before they are multiplied.
// FP64 operations per cell: ~1840 -> ~1300 (the L / R states of inv F are now half of them).
// ---------------------------------------------------------------------------
template <class T, int N>
__device__ __forceinline__ void lds_vec(const T* p, T (&v)[N]) {
  if constexpr (sizeof(T) == 8) {
    static_assert(N % 2 == 0, "16-byte pieces");
#pragma unroll
    for (int k = 0; k < N / 2; ++k) {
      const double2 w = reinterpret_cast<const double2*>(p)[k];
      v[2 * k] = w.x; v[2 * k + 1] = w.y;
    }
  } else {
    static_assert(N % 4 == 0, "16-byte pieces");
#pragma unroll
    for (int k = 0; k < N / 4; ++k) {
      const float4 w = reinterpret_cast<const float4*>(p)[k];
      v[4 * k] = w.x; v[4 * k + 1] = w.y; v[4 * k + 2] = w.z; v[4 * k + 3] = w.w;
    }
  }
}
template <class T, int N>
__device__ __forceinline__ void sts_vec(T* p, const T (&v)[N]) {
  if constexpr (sizeof(T) == 8) {
#pragma unroll
    for (int k = 0; k < N / 2; ++k) reinterpret_cast<double2*>(p)[k] = make_double2(v[2 * k], v[2 * k + 1]);
  } else {
#pragma unroll
    for (int k = 0; k < N / 4; ++k) reinterpret_cast<float4*>(p)[k] = make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
  }
}
constexpr int UNI_PAIR = 16;  // coefficients per pair (15 + 1 of padding)
constexpr int UNI_TERM = 8;   // products per term and cell position (7 + 1 of padding)
// staging: two Metric buffers (four phases per step) and one ConservedMetric buffer per face axis (the three
// conserved arrays leave together with the cell arrays, so the 27 accumulators are dead before the Jacobian phases)
template <class T> constexpr int uni_stage_vals = 2 * stage_len<T, 10> + 3 * stage_len<T, 9>;
#ifndef CG_UNI_MINB
#define CG_UNI_MINB 8
#endif
#ifndef CG_UNI_UNROLL_T
#define CG_UNI_UNROLL_T 2
#endif
#ifndef CG_UNI_UNROLL_P
#define CG_UNI_UNROLL_P 2
#endif
constexpr int UNI_UNROLL_T = CG_UNI_UNROLL_T, UNI_UNROLL_P = CG_UNI_UNROLL_P;
template <class T, int SCH>
__global__ void __launch_bounds__(32, CG_UNI_MINB) k_mapped_metrics3d_uni(Dims D, MCoordTerms CT, MPairs PR,
                                                                           const T* __restrict__ tab,
                                                                           const T* __restrict__ ftab, int64_t L,
                                                                           MetricOut<T> O, int nt, int kchunk) {
  extern __shared__ __align__(128) unsigned char cg_dyn_smem[];
  const T lam1 = lam1_of<T>(SCH), lam2 = lam2_of<T>(SCH);
  const int64_t L1 = L + 1;
  const int lane = threadIdx.x;
  const int64_t nw0 = D.nw[0], nw1 = D.nw[1], nw2 = D.nw[2];
  const int64_t j = (int64_t)blockIdx.x;  // j-fastest block order: resident warps share the xi rows in L1
  const int64_t i0 = (int64_t)blockIdx.y * MAPPED_OUT;
  const int nvalid = (int)(nw0 - i0 < MAPPED_OUT ? nw0 - i0 : MAPPED_OUT);
  // lane 31 (and lanes past the row end) evaluate the cell nw0, i.e. the +xi neighbour of the last cell: the
  // factor tables hold nw + 1 entries per axis; the operator tables hold nw (their values are not sent)
  const int64_t i = i0 + lane < nw0 ? i0 + lane : nw0;
  const int64_t ic = i < nw0 ? i : nw0 - 1;
  const int64_t k0 = (int64_t)blockIdx.z * kchunk;
  const int64_t k1 = k0 + kchunk < nw2 ? k0 + kchunk : nw2;
  const int np = PR.n;
  T* const sg = reinterpret_cast<T*>(cg_dyn_smem);
  T* const b10 = sg;
  T* const b9 = sg + 2 * stage_len<T, 10>;
  T* const pairU = sg + uni_stage_vals<T>;          // [np][UNI_PAIR]
  T* const termU = pairU + (int64_t)np * UNI_PAIR;  // [3 positions][nt][UNI_TERM], flat (coordinate-ordered) term index
  const T* const xrow = tab + ic * OP_COUNT;                   // + (p * 3) * L * OP_COUNT : xi row of pair p
  const T* const yrow = tab + (L + j) * OP_COUNT;              // + (p * 3) * L * OP_COUNT : eta row of pair p
  const T* const zrow = tab + 2 * L * OP_COUNT;                // + ((p * 3) * L + m) * OP_COUNT
  const T* const fx = ftab + i * 4;                            // + (m * 3) * L1 * 4 : xi factor row of term m
  const int64_t pstride = 3 * L * OP_COUNT, tstride = 3 * L1 * 4;

  // ---- warp-uniform coefficient tables -------------------------------------------------------------------------
  // pair p: U[0..14] such that, with x = the pair's xi row,
  //   Gh[xi][xi]    += x.E0 U0               Gh[xi][eta]   += x.FD0 U1 + x.E1 U3    Gh[xi][zeta]  += x.FD0 U2 + x.E1 U4
  //   Gh[eta][eta]  += x.D0 U5 + x.V1 U9     Gh[eta][zeta] += x.D0 U6 + x.V1 U10    Gh[eta][xi]   += x.V0 U13
  //   Gh[zeta][zeta]+= x.D0 U7 + x.V1 U11    Gh[zeta][eta] += x.D0 U8 + x.V1 U12    Gh[zeta][xi]  += x.V0 U14
  // ([face][row]; the column is the pair's).  zpre = the zeta row of the pair `lane`, loaded earlier.
  auto pair_uniform = [&](int64_t mz, const T (&zpre)[OP_COUNT]) {
    for (int p = lane; p < np; p += 32) {
      T y[OP_COUNT], z[OP_COUNT];
      ldg_vec<T, OP_COUNT>(yrow + p * pstride, y);
      if (p == lane) {
#pragma unroll
        for (int o = 0; o < OP_COUNT; ++o) z[o] = zpre[o];
      } else {
        ldg_vec<T, OP_COUNT>(zrow + p * pstride + mz * OP_COUNT, z);
      }
      const T C = T(PR.coef[p]);
#pragma unroll
      for (int o = 0; o < OP_COUNT; ++o) z[o] *= C;
      T u[UNI_PAIR];
      u[0] = y[OP_V1] * z[OP_D0] - y[OP_D0] * z[OP_V1];
      u[1] = y[OP_V0] * z[OP_V1];
      u[2] = -(z[OP_V0] * y[OP_V1]);
      u[3] = -(y[OP_V0] * z[OP_D0]);
      u[4] = z[OP_V0] * y[OP_D0];
      u[5] = y[OP_E0] * z[OP_V1];
      u[6] = -(z[OP_V0] * y[OP_E1]);
      u[7] = -(z[OP_E0] * y[OP_V1]);
      u[8] = y[OP_V0] * z[OP_E1];
      u[9] = -(y[OP_E0] * z[OP_D0]);
      u[10] = z[OP_V0] * y[OP_FD0];
      u[11] = z[OP_E0] * y[OP_D0];
      u[12] = -(y[OP_V0] * z[OP_FD0]);
      u[13] = y[OP_E1] * z[OP_D0] - y[OP_FD0] * z[OP_V1];
      u[14] = z[OP_FD0] * y[OP_V1] - z[OP_E1] * y[OP_D0];
      u[15] = T(0);
      sts_vec<T, UNI_PAIR>(pairU + p * UNI_PAIR, u);
    }
  };
  // term s (flat index), products P(a, b) = coef f_eta^(a) f_zeta^(b):
  //   position 0 = (j, mA)     : [P00 P10 P01 P20 P11 P30 P21 -]   dF/dxi, d2F/dxi2, dF/deta, d2F/deta2 of the cell
  //   position 1 = (j + 1, mA) : same products                      F, dF/deta, d2F/deta2 of the +eta neighbour
  //   position 2 = (j, mC)     : [P00 P10 P01 P11 P02 P12 P03 -]   F, dF/dzeta, d2F/dzeta2 of the +zeta neighbour
  // zA / zC: the zeta factor rows (4 derivative orders) of the term `lane` at mA / mC, loaded earlier.
  auto term_uniform = [&](int64_t mA, int64_t mC, const T (&zA_pre)[4], const T (&zC_pre)[4]) {
    for (int s = lane; s < nt; s += 32) {
      const int m = CT.flat[s];
      const T* const base = ftab + (int64_t)m * tstride;
      T ya[4], yb[4], zA[4], zC[4];
      ldg_vec<T, 4>(base + (L1 + j) * 4, ya);
      ldg_vec<T, 4>(base + (L1 + j + 1) * 4, yb);
      if (s == lane) {
#pragma unroll
        for (int o = 0; o < 4; ++o) { zA[o] = zA_pre[o]; zC[o] = zC_pre[o]; }
      } else {
        ldg_vec<T, 4>(base + (2 * L1 + mA) * 4, zA);
        ldg_vec<T, 4>(base + (2 * L1 + mC) * 4, zC);
      }
      const T c = T(CT.coef[s]);
      T u[UNI_TERM];
      {
        const T z0 = c * zA[0], z1 = c * zA[1];
        u[0] = ya[0] * z0; u[1] = ya[1] * z0; u[2] = ya[0] * z1; u[3] = ya[2] * z0;
        u[4] = ya[1] * z1; u[5] = ya[3] * z0; u[6] = ya[2] * z1; u[7] = T(0);
        sts_vec<T, UNI_TERM>(termU + (0 * nt + s) * UNI_TERM, u);
        u[0] = yb[0] * z0; u[1] = yb[1] * z0; u[2] = yb[0] * z1; u[3] = yb[2] * z0;
        u[4] = yb[1] * z1; u[5] = yb[3] * z0; u[6] = yb[2] * z1;
        sts_vec<T, UNI_TERM>(termU + (1 * nt + s) * UNI_TERM, u);
      }
      {
        const T y0 = c * ya[0], y1 = c * ya[1];
        u[0] = y0 * zC[0]; u[1] = y1 * zC[0]; u[2] = y0 * zC[1]; u[3] = y1 * zC[1];
        u[4] = y0 * zC[2]; u[5] = y1 * zC[2]; u[6] = y0 * zC[3];
        sts_vec<T, UNI_TERM>(termU + (2 * nt + s) * UNI_TERM, u);
      }
    }
  };
  // the zeta rows a lane multiplies in the next uniform phase (first pass of the loops above)
  auto load_zpair = [&](int64_t mz, T (&z)[OP_COUNT]) {
    const int p = lane < np ? lane : np - 1;
    ldg_vec<T, OP_COUNT>(zrow + p * pstride + mz * OP_COUNT, z);
  };
  auto load_zterm = [&](int64_t mz, T (&z)[4]) {
    const int s = lane < nt ? lane : nt - 1;
    ldg_vec<T, 4>(ftab + (int64_t)CT.flat[s] * tstride + (2 * L1 + mz) * 4, z);
  };

  // ---- per-lane accumulation over the terms: one multiply-add per (xi factor value, uniform product) -------------
  // F, dF/dxi_2, d2F/dxi_2^2 at the +zeta position (termU position 2)
  auto jac_zeta = [&](T (&F)[3][3], T (&Fp)[3][3], T (&Fpp)[3][3]) {
    int s = 0;
    T a[4], an[4], u[UNI_TERM], un[UNI_TERM];
    ldg_vec<T, 4>(fx + (int64_t)CT.flat[0] * tstride, a);
    lds_vec<T, UNI_TERM>(termU + (2 * nt) * UNI_TERM, u);
#pragma unroll
    for (int q = 0; q < 3; ++q) {
#pragma unroll
      for (int d = 0; d < 3; ++d) F[q][d] = Fp[q][d] = Fpp[q][d] = T(0);
#pragma unroll(UNI_UNROLL_T)
      for (int mm = 0; mm < CT.cnt[q]; ++mm, ++s) {
        // the rows of the NEXT term are requested before this one is accumulated (flat[] and termU carry a sentinel)
        ldg_vec<T, 4>(fx + (int64_t)CT.flat[s + 1] * tstride, an);
        lds_vec<T, UNI_TERM>(termU + (2 * nt + s + 1) * UNI_TERM, un);
        F[q][0] += a[1] * u[0]; F[q][1] += a[0] * u[1]; F[q][2] += a[0] * u[2];
        Fp[q][0] += a[1] * u[2]; Fp[q][1] += a[0] * u[3]; Fp[q][2] += a[0] * u[4];
        Fpp[q][0] += a[1] * u[4]; Fpp[q][1] += a[0] * u[5]; Fpp[q][2] += a[0] * u[6];
#pragma unroll
        for (int o = 0; o < 4; ++o) a[o] = an[o];
#pragma unroll
        for (int o = 0; o < UNI_TERM; ++o) u[o] = un[o];
      }
    }
  };
  // dF/dxi, d2F/dxi^2 of the cell (termU position 0)
  auto jac_xi = [&](T (&Fp)[3][3], T (&Fpp)[3][3]) {
    int s = 0;
    T a[4], an[4], u[4], un[4];
    ldg_vec<T, 4>(fx + (int64_t)CT.flat[0] * tstride, a);
    lds_vec<T, 4>(termU, u);
#pragma unroll
    for (int q = 0; q < 3; ++q) {
#pragma unroll
      for (int d = 0; d < 3; ++d) Fp[q][d] = Fpp[q][d] = T(0);
#pragma unroll(UNI_UNROLL_T)
      for (int mm = 0; mm < CT.cnt[q]; ++mm, ++s) {
        ldg_vec<T, 4>(fx + (int64_t)CT.flat[s + 1] * tstride, an);
        lds_vec<T, 4>(termU + (0 * nt + s + 1) * UNI_TERM, un);
        Fp[q][0] += a[2] * u[0]; Fp[q][1] += a[1] * u[1]; Fp[q][2] += a[1] * u[2];
        Fpp[q][0] += a[3] * u[0]; Fpp[q][1] += a[2] * u[1]; Fpp[q][2] += a[2] * u[2];
#pragma unroll
        for (int o = 0; o < 4; ++o) { a[o] = an[o]; u[o] = un[o]; }
      }
    }
  };
  // dF/deta, d2F/deta^2 of the cell (position 0) and F, dF/deta, d2F/deta^2 of the +eta neighbour (position 1)
  auto jac_eta = [&](T (&Fp)[3][3], T (&Fpp)[3][3], T (&Fn)[3][3], T (&Fpn)[3][3], T (&Fppn)[3][3]) {
    int s = 0;
    T a[4], an[4], u[UNI_TERM], un[UNI_TERM], v[UNI_TERM], vn[UNI_TERM];
    ldg_vec<T, 4>(fx + (int64_t)CT.flat[0] * tstride, a);
    lds_vec<T, UNI_TERM>(termU, u);
    lds_vec<T, UNI_TERM>(termU + nt * UNI_TERM, v);
#pragma unroll
    for (int q = 0; q < 3; ++q) {
#pragma unroll
      for (int d = 0; d < 3; ++d) Fp[q][d] = Fpp[q][d] = Fn[q][d] = Fpn[q][d] = Fppn[q][d] = T(0);
#pragma unroll(UNI_UNROLL_T)
      for (int mm = 0; mm < CT.cnt[q]; ++mm, ++s) {
        ldg_vec<T, 4>(fx + (int64_t)CT.flat[s + 1] * tstride, an);
        lds_vec<T, UNI_TERM>(termU + (0 * nt + s + 1) * UNI_TERM, un);
        lds_vec<T, UNI_TERM>(termU + (1 * nt + s + 1) * UNI_TERM, vn);
        Fp[q][0] += a[1] * u[1]; Fp[q][1] += a[0] * u[3]; Fp[q][2] += a[0] * u[4];
        Fpp[q][0] += a[1] * u[3]; Fpp[q][1] += a[0] * u[5]; Fpp[q][2] += a[0] * u[6];
        Fn[q][0] += a[1] * v[0]; Fn[q][1] += a[0] * v[1]; Fn[q][2] += a[0] * v[2];
        Fpn[q][0] += a[1] * v[1]; Fpn[q][1] += a[0] * v[3]; Fpn[q][2] += a[0] * v[4];
        Fppn[q][0] += a[1] * v[3]; Fppn[q][1] += a[0] * v[5]; Fppn[q][2] += a[0] * v[6];
#pragma unroll
        for (int o = 0; o < 4; ++o) a[o] = an[o];
#pragma unroll
        for (int o = 0; o < UNI_TERM; ++o) { u[o] = un[o]; v[o] = vn[o]; }
      }
    }
  };

  // carried along zeta: F, G = inv F, det F and the zeta-axis L state of the cell of the current step
  T Fc[3][3], Gc[3][3], Jc, Lz[3][3];
  T zt_cur[4];  // zeta factor row of the term `lane` at the plane after the current one
  {
    // prologue: position 2 at (j, k0) gives the first cell; then the tables of step k0
    T zp[OP_COUNT], za[4], zc[4];
    load_zterm(k0, za);
    term_uniform(k0, k0, za, za);
    __syncwarp();
    T Fp[3][3], Fpp[3][3], R2[3][3];
    jac_zeta(Fc, Fp, Fpp);
    Jc = inv3(Fc, Gc);
    mapped_states3_g(Gc, Fp, Fpp, lam1, lam2, Lz, R2);
    __syncwarp();
    load_zpair(k0, zp);
    load_zterm(k0 + 1, zc);
    pair_uniform(k0, zp);
    term_uniform(k0, k0 + 1, za, zc);
#pragma unroll
    for (int o = 0; o < 4; ++o) zt_cur[o] = zc[o];
    __syncwarp();
  }
  for (int64_t m = k0; m < k1; ++m) {
    const int64_t cl0 = i0 + nw0 * (j + nw1 * m);
    const bool more = m + 1 < k1;
    // the zeta rows of the next step's pairs: requested now, multiplied after the conserved sums
    T zp[OP_COUNT];
    load_zpair(more ? m + 1 : m, zp);
    // ---- conserved metrics: 15 multiply-adds per pair -------------------------------------------------------------
    T Gh[3][3][3];  // [face][row][col]
#pragma unroll
    for (int f = 0; f < 3; ++f)
#pragma unroll
      for (int rr = 0; rr < 3; ++rr)
#pragma unroll
        for (int cc = 0; cc < 3; ++cc) Gh[f][rr][cc] = T(0);
    {
      T x[OP_COUNT], xn[OP_COUNT], u[UNI_PAIR], un[UNI_PAIR];
      ldg_vec<T, OP_COUNT>(xrow, x);
      lds_vec<T, UNI_PAIR>(pairU, u);
#pragma unroll
      for (int cc = 0; cc < 3; ++cc) {
#pragma unroll(UNI_UNROLL_P)
        for (int p = PR.cstart[cc]; p < PR.cstart[cc + 1]; ++p) {
          const int pn = p + 1 < np ? p + 1 : p;  // the last prefetch re-reads the last pair
          ldg_vec<T, OP_COUNT>(xrow + pn * pstride, xn);
          lds_vec<T, UNI_PAIR>(pairU + pn * UNI_PAIR, un);
          Gh[0][0][cc] += x[OP_E0] * u[0];
          Gh[0][1][cc] += x[OP_FD0] * u[1] + x[OP_E1] * u[3];
          Gh[0][2][cc] += x[OP_FD0] * u[2] + x[OP_E1] * u[4];
          Gh[1][1][cc] += x[OP_D0] * u[5] + x[OP_V1] * u[9];
          Gh[1][2][cc] += x[OP_D0] * u[6] + x[OP_V1] * u[10];
          Gh[1][0][cc] += x[OP_V0] * u[13];
          Gh[2][2][cc] += x[OP_D0] * u[7] + x[OP_V1] * u[11];
          Gh[2][1][cc] += x[OP_D0] * u[8] + x[OP_V1] * u[12];
          Gh[2][0][cc] += x[OP_V0] * u[14];
#pragma unroll
          for (int o = 0; o < OP_COUNT; ++o) x[o] = xn[o];
#pragma unroll
          for (int o = 0; o < UNI_PAIR; ++o) u[o] = un[o];
        }
      }
    }
    // ---- phase 0: cell arrays + the three conserved arrays ------------------------------------------------------------
    bulk_wait_read();  // the TMA unit has finished reading last step's staging buffers (lanes > 0 own no groups)
    __syncwarp();      // ... and every lane has read pairU
    {
      Stage<T, 10> Scf(O.cell_fwd, cl0, b10), Sci(O.cell_inv, cl0, b10 + stage_len<T, 10>);
      Stage<T, 9> Sc0(O.face_cons[0], cl0, b9), Sc1(O.face_cons[1], cl0, b9 + stage_len<T, 9>),
          Sc2(O.face_cons[2], cl0, b9 + 2 * stage_len<T, 9>);
      put10(Scf.slot(lane), Fc, Jc);  // every lane stages (no branch); slots >= nvalid are never sent
      put10(Sci.slot(lane), Gc, Jc);
      put9(Sc0.slot(lane), Gh[0]);
      put9(Sc1.slot(lane), Gh[1]);
      put9(Sc2.slot(lane), Gh[2]);
      if (more) pair_uniform(m + 1, zp);  // next step's pair coefficients (pairU is free: barrier above)
      fence_async_smem();
      __syncwarp();
      Scf.issue(nvalid, lane);
      Sci.issue(nvalid, lane);
      Sc0.issue(nvalid, lane);
      Sc1.issue(nvalid, lane);
      Sc2.issue(nvalid, lane);
      if (lane == 0) bulk_commit();
    }
    // the zeta factor rows the next step's term products need (plane m + 2)
    T zt_next[4];
    load_zterm(m + 2 <= nw2 ? m + 2 : nw2, zt_next);
    // ---- reconstructed inverse Jacobian on the three faces ----------------------------------------------------------
    auto face_out = [&](int f, const T (&Gf)[3][3]) {
      T Ff[3][3];
      const T J = rcp(inv3(Gf, Ff));  // det(F_face) = 1 / det(G_face)
      bulk_wait_read();               // the previous phase has left the two Metric buffers
      __syncwarp();
      Stage<T, 10> Sff(O.face_fwd[f], cl0, b10), Sfi(O.face_inv[f], cl0, b10 + stage_len<T, 10>);
      put10(Sff.slot(lane), Ff, J);
      put10(Sfi.slot(lane), Gf, J);
      fence_async_smem();
      __syncwarp();
      Sff.issue(nvalid, lane);
      Sfi.issue(nvalid, lane);
      if (lane == 0) bulk_commit();
    };
    {  // xi face: the +xi neighbour's R state by shuffle
      T Fp[3][3], Fpp[3][3], L0[3][3], R0[3][3], Gf[3][3];
      jac_xi(Fp, Fpp);
      mapped_states3_g(Gc, Fp, Fpp, lam1, lam2, L0, R0);
#pragma unroll
      for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = 0; b < 3; ++b) Gf[a][b] = T(0.5) * (L0[a][b] + __shfl_down_sync(FULLMASK, R0[a][b], 1));
      face_out(0, Gf);
    }
    {  // eta face: the +eta neighbour is recomputed (9 multiply-adds per term)
      T Fp[3][3], Fpp[3][3], Fn[3][3], Fpn[3][3], Fppn[3][3], L1s[3][3], R1s[3][3], Gn[3][3], Ln[3][3], Rn[3][3], Gf[3][3];
      jac_eta(Fp, Fpp, Fn, Fpn, Fppn);
      mapped_states3_g(Gc, Fp, Fpp, lam1, lam2, L1s, R1s);
      inv3(Fn, Gn);
      mapped_states3_g(Gn, Fpn, Fppn, lam1, lam2, Ln, Rn);
#pragma unroll
      for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = 0; b < 3; ++b) Gf[a][b] = T(0.5) * (L1s[a][b] + Rn[a][b]);
      face_out(1, Gf);
    }
    {  // zeta face: the +zeta neighbour is the cell of the next step; its F, G, J and L state are carried
      T Fp[3][3], Fpp[3][3], Ln[3][3], Rn[3][3], Gf[3][3];
      jac_zeta(Fc, Fp, Fpp);
      Jc = inv3(Fc, Gc);
      mapped_states3_g(Gc, Fp, Fpp, lam1, lam2, Ln, Rn);
#pragma unroll
      for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = 0; b < 3; ++b) {
          Gf[a][b] = T(0.5) * (Lz[a][b] + Rn[a][b]);
          Lz[a][b] = Ln[a][b];
        }
      face_out(2, Gf);
    }
    // ---- next step's term products ------------------------------------------------------------------------------------
    __syncwarp();  // every lane has read termU
    if (more) term_uniform(m + 1, m + 2 <= nw2 ? m + 2 : nw2, zt_cur, zt_next);
#pragma unroll
    for (int o = 0; o < 4; ++o) zt_cur[o] = zt_next[o];
    __syncwarp();
  }
  if (lane == 0) bulk_wait_read();  // shared memory must outlive the TMA reads
}

template <class T>
inline bool launch_mapped_metrics3d_uni(const Dims& D, const MTerms& TM, const MPairs& PR, const T* tab, const T* ftab,
                                        int64_t L, const MetricOut<T>& O, int scheme, cudaStream_t st) {
  const unsigned gx = (unsigned)((D.nw[0] + MAPPED_OUT - 1) / MAPPED_OUT);
  if (gx > 65535u || D.nw[1] <= 0) return false;
  const int smem = (int)((uni_stage_vals<T> + (int64_t)PR.n * UNI_PAIR + (3 * (int64_t)TM.n + 1) * UNI_TERM) * sizeof(T));
  int per_sm = (int)(200 * 1024 / (smem + 1024));
  if (per_sm > CG_UNI_MINB) per_sm = CG_UNI_MINB;
  if (per_sm < 1) return false;
  const int kchunk = pick_chunk((int64_t)gx * D.nw[1], D.nw[2], 148LL * per_sm, 32, 1.0);
  const dim3 grid((unsigned)D.nw[1], gx, (unsigned)((D.nw[2] + kchunk - 1) / kchunk));
  const MCoordTerms CT = coord_terms(TM);
  auto go = [&](auto kern) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return false;
    kern<<<grid, 32, smem, st>>>(D, CT, PR, tab, ftab, L, O, TM.n, kchunk);
    return true;
  };
  if (scheme == ENDPOINT) return go(k_mapped_metrics3d_uni<T, ENDPOINT>);
  if (scheme == GRADIENT) return go(k_mapped_metrics3d_uni<T, GRADIENT>);
  return go(k_mapped_metrics3d_uni<T, CURVATURE>);
}


// ---------------------------------------------------------------------------
// ADThomasLombardMetric as its OWN kernel (metrics.jl:607-668 closures, pair tables :2336-3080, fill :3082-3188), FULL
// profile.  What the scheme stores per face is less than the reconstruction schemes and needs no neighbour cell:
//   forward = the mapping's Jacobian AT THE FACE CENTRE (:3162-3168), inverse = inv, J = det,
//   conserved = the active row only (:3174-3182), rows of the other two axes are zero.
// The active row is the same sum over term pairs of triple products of 1-D operator-table values as in
// k_mapped_metrics3d_staged (outer reconstruction along the face axis and the two cell-centred differences act on
// different axes, so their order does not matter: the reference's pair tables reconstruct first, its closure graph
// differences first, and its test asserts the two to agree at 1e-12, test_3d_continuous.jl:296-297).  No inverse-
// Jacobian reconstruction, no inactive rows: about a third of the arithmetic of the CurvatureCorrected kernel, no
// carried state, 16 one-warp blocks per SM; the kernel is bound by its 856 B of output per cell.
// fhalf = the factor tables at centre + 1/2 (the high-side face position along an axis).
// ---------------------------------------------------------------------------
#ifndef CG_ADTL_MINB
#define CG_ADTL_MINB 16
#endif
template <class T>
__global__ void __launch_bounds__(32, CG_ADTL_MINB) k_mapped_adtl3d_staged(Dims D, MTerms TM, MPairs PR,
                                                                            const T* __restrict__ tab,
                                                                            const T* __restrict__ ftab,
                                                                            const T* __restrict__ fhalf, int64_t L,
                                                                            MetricOut<T> O, int kchunk) {
  extern __shared__ __align__(128) unsigned char cg_dyn_smem[];
  const int64_t L1 = L + 1;
  const int lane = threadIdx.x;
  const int64_t nw0 = D.nw[0], nw1 = D.nw[1], nw2 = D.nw[2];
  const int64_t j = (int64_t)blockIdx.y;
  if (j >= nw1) return;
  const int64_t i0 = (int64_t)blockIdx.x * 32;
  const int nvalid = (int)(nw0 - i0 < 32 ? nw0 - i0 : 32);
  const int64_t i = i0 + lane < nw0 ? i0 + lane : nw0 - 1;  // lanes past the row end repeat the last cell (never sent)
  const int64_t k0 = (int64_t)blockIdx.z * kchunk;
  const int64_t k1 = k0 + kchunk < nw2 ? k0 + kchunk : nw2;
  T* const b10 = reinterpret_cast<T*>(cg_dyn_smem);
  T* const b9 = b10 + 2 * stage_len<T, 10>;
  // F[q][d] = sum over the terms of coordinate q of coef * prod_ax f_ax^(ax == d) with the factor values of axis `fa`
  // taken at the face position (fhalf) when fa >= 0
  auto jacobian = [&](const int64_t (&ix)[3], int fa, T (&F)[3][3]) {
#pragma unroll
    for (int q = 0; q < 3; ++q)
#pragma unroll
      for (int d = 0; d < 3; ++d) F[q][d] = T(0);
    for (int m = 0; m < TM.n; ++m) {
      T a[3][2];
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        const T* tb = (d == fa ? fhalf : ftab) + ((int64_t)(m * 3 + d) * L1 + ix[d]) * 4;
        if constexpr (sizeof(T) == 8) {
          const double2 w = __ldg(reinterpret_cast<const double2*>(tb));
          a[d][0] = w.x;
          a[d][1] = w.y;
        } else {
          a[d][0] = __ldg(tb);
          a[d][1] = __ldg(tb + 1);
        }
      }
      const T c = T(TM.t[m].coef);
      const int q = TM.t[m].coord;
      const T v0 = c * a[0][1] * a[1][0] * a[2][0], v1 = c * a[0][0] * a[1][1] * a[2][0], v2 = c * a[0][0] * a[1][0] * a[2][1];
#pragma unroll
      for (int qq = 0; qq < 3; ++qq)
        if (qq == q) {
          F[qq][0] += v0;
          F[qq][1] += v1;
          F[qq][2] += v2;
        }
    }
  };
  for (int64_t m = k0; m < k1; ++m) {
    const int64_t ix[3] = {i, j, m};
    const int64_t cl0 = i0 + nw0 * (j + nw1 * m);
    T F[3][3], G[3][3];
    // ---- cell arrays ----
    jacobian(ix, -1, F);
    T J = inv3(F, G);
    bulk_wait_read();  // the TMA unit has finished reading the previous phase (lanes > 0 own no groups)
    __syncwarp();
    {
      Stage<T, 10> Scf(O.cell_fwd, cl0, b10), Sci(O.cell_inv, cl0, b10 + stage_len<T, 10>);
      put10(Scf.slot(lane), F, J);
      put10(Sci.slot(lane), G, J);
      fence_async_smem();
      __syncwarp();
      Scf.issue(nvalid, lane);
      Sci.issue(nvalid, lane);
      if (lane == 0) bulk_commit();
    }
    // ---- active rows: sums over the term pairs of triple products of 1-D table values ----
    T act[3][3];  // [face][col]
#pragma unroll
    for (int f = 0; f < 3; ++f)
#pragma unroll
      for (int cc = 0; cc < 3; ++cc) act[f][cc] = T(0);
#pragma unroll
    for (int cc = 0; cc < 3; ++cc) {
      for (int p = PR.cstart[cc]; p < PR.cstart[cc + 1]; ++p) {
        T tv[3][OP_COUNT];
#pragma unroll
        for (int d = 0; d < 3; ++d) ldg_vec<T, OP_COUNT>(tab + ((int64_t)(p * 3 + d) * L + ix[d]) * OP_COUNT, tv[d]);
        const T C = T(PR.coef[p]);
#pragma unroll
        for (int f = 0; f < 3; ++f) {
          const int s = (f + 1) % 3, t = (f + 2) % 3;
          act[f][cc] += C * tv[f][OP_E0] * (tv[s][OP_V1] * tv[t][OP_D0] - tv[s][OP_D0] * tv[t][OP_V1]);
        }
      }
    }
    // ---- one phase per face axis ----
#pragma unroll
    for (int f = 0; f < 3; ++f) {
      jacobian(ix, f, F);
      J = inv3(F, G);
      T Gh[3][3];
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int cc = 0; cc < 3; ++cc) Gh[r][cc] = r == f ? act[f][cc] : T(0);
      bulk_wait_read();
      __syncwarp();
      Stage<T, 10> Sff(O.face_fwd[f], cl0, b10), Sfi(O.face_inv[f], cl0, b10 + stage_len<T, 10>);
      Stage<T, 9> Sfc(O.face_cons[f], cl0, b9);
      put9(Sfc.slot(lane), Gh);
      put10(Sff.slot(lane), F, J);
      put10(Sfi.slot(lane), G, J);
      fence_async_smem();
      __syncwarp();
      Sff.issue(nvalid, lane);
      Sfi.issue(nvalid, lane);
      Sfc.issue(nvalid, lane);
      if (lane == 0) bulk_commit();
    }
  }
  if (lane == 0) bulk_wait_read();  // shared memory must outlive the TMA reads
}

template <class T>
inline bool launch_mapped_adtl3d_staged(const Dims& D, const MTerms& TM, const MPairs& PR, const T* tab, const T* ftab,
                                        const T* fhalf, int64_t L, const MetricOut<T>& O, cudaStream_t st) {
  const unsigned gx = (unsigned)((D.nw[0] + 31) / 32), gy = (unsigned)D.nw[1];
  const int kchunk = pick_chunk((int64_t)gx * gy, D.nw[2], 148LL * CG_ADTL_MINB, 32, 0.2);
  const int smem = mapped_stage_vals<T> * (int)sizeof(T);
  if (cudaFuncSetAttribute(k_mapped_adtl3d_staged<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return false;
  k_mapped_adtl3d_staged<T><<<dim3(gx, gy, (unsigned)((D.nw[2] + kchunk - 1) / kchunk)), 32, smem, st>>>(
      D, TM, PR, tab, ftab, fhalf, L, O, kchunk);
  return true;
}

// ---------------------------------------------------------------------------
// ADThomasLombardMetric (metrics.jl:607-668, fill :3082-3188) on top of the CurvatureCorrected evaluation: the face
// forward metric is the mapping's Jacobian AT THE FACE CENTRE, inverse = inv(F), J = det(F) (:3162-3168), and the
// conserved matrix keeps only its active row (:3174-3182; the active row equals the CurvatureCorrected one, asserted
// by the reference at 1e-12, test_3d_continuous.jl:296-297).  fhalf = the factor tables at centre + 1/2.
// ---------------------------------------------------------------------------
template <class T>
__global__ void __launch_bounds__(128) k_mapped_adtl_faces(Dims D, MTerms TM, const T* __restrict__ ftab,
                                                           const T* __restrict__ fhalf, int64_t L1, MetricOut<T> O) {
  const int64_t nc = D.nw[0] * D.nw[1] * D.nw[2];
  for (int64_t cl = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; cl < nc; cl += (int64_t)gridDim.x * blockDim.x) {
    const int64_t ix[3] = {cl % D.nw[0], (cl / D.nw[0]) % D.nw[1], cl / (D.nw[0] * D.nw[1])};
#pragma unroll
    for (int f = 0; f < 3; ++f) {
      T F[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}}, G[3][3];
      for (int m = 0; m < TM.n; ++m) {
        const int q = TM.t[m].coord;
        T a[3][2];
#pragma unroll
        for (int d = 0; d < 3; ++d) {
          const T* tb = (d == f ? fhalf : ftab) + ((int64_t)(m * 3 + d) * L1 + ix[d]) * 4;
          a[d][0] = tb[0];
          a[d][1] = tb[1];
        }
        const T c = T(TM.t[m].coef);
#pragma unroll
        for (int d = 0; d < 3; ++d) {
          const T v = c * a[d][1] * a[(d + 1) % 3][0] * a[(d + 2) % 3][0];
#pragma unroll
          for (int qq = 0; qq < 3; ++qq)
            if (qq == q) F[qq][d] += v;
        }
      }
      const T J = inv3(F, G);
      if (O.face_fwd[f]) {
        T* o = O.face_fwd[f] + 10 * cl;
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
          for (int r = 0; r < 3; ++r) o[r + 3 * c] = F[r][c];
        o[9] = J;
      }
      if (O.face_inv[f]) {
        T* o = O.face_inv[f] + 10 * cl;
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
          for (int r = 0; r < 3; ++r) o[r + 3 * c] = G[r][c];
        o[9] = J;
      }
      if (O.face_cons[f]) {
        T* o = O.face_cons[f] + 9 * cl;
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
          for (int r = 0; r < 3; ++r)
            if (r != f) o[r + 3 * c] = T(0);
      }
    }
  }
}

// =============================== 2-D ========================================
// Ghat = J inv(F) = [[y_eta, -x_eta], [-y_xi, x_xi]] reconstructed entry-wise (metrics.jl:276-282,
// 577-582); pairs are (term, unit term), so V1/E1 hold the differentiated factor.
template <class T>
__global__ void k_mapped_metrics2d(Dims D, MTerms TM, MPairs PR, const T* __restrict__ tab,
                                   const T* __restrict__ ftab, int64_t L, MetricOut<T> O, int scheme, int what) {
  const T lam1 = lam1_of<T>(scheme), lam2 = lam2_of<T>(scheme);
  const int64_t L1 = L + 1;
  const int64_t ntot = D.nw[0] * D.nw[1];
  for (int64_t lin = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; lin < ntot;
       lin += (int64_t)gridDim.x * blockDim.x) {
    int64_t ix[3] = {lin % D.nw[0], lin / D.nw[0], 0};
    T F[2][2], Fp[2][2], Fpp[2][2], G[2][2], Lc[2][2], Rc[2][2];
    if (what & 1) {
      mapped_jac<T, 2>(TM, ftab, L1, ix, 0, F, Fp, Fpp);
      T J = inv2(F, G);
      if (O.cell_fwd) { T* o = O.cell_fwd + 5 * lin; o[0] = F[0][0]; o[1] = F[1][0]; o[2] = F[0][1]; o[3] = F[1][1]; o[4] = J; }
      if (O.cell_inv) { T* o = O.cell_inv + 5 * lin; o[0] = G[0][0]; o[1] = G[1][0]; o[2] = G[0][1]; o[3] = G[1][1]; o[4] = J; }
      if (O.cjac) O.cjac[lin] = J;
    }
    if (!(what & 2)) continue;
    for (int f = 0; f < 2; ++f) {
      T Gh[2][2] = {{0, 0}, {0, 0}};
      for (int p = 0; p < PR.n; ++p) {
        const int q = TM.t[PR.p[p].m1].coord;
        const T C = T(TM.t[PR.p[p].m1].coef);
        const T* t0 = tab + ((int64_t)(p * 3 + 0) * L + ix[0]) * OP_COUNT;
        const T* t1 = tab + ((int64_t)(p * 3 + 1) * L + ix[1]) * OP_COUNT;
        // d q / d xi  and  d q / d eta reconstructed to the f-face
        T qxi = C * (f == 0 ? t0[OP_E1] : t0[OP_V1]) * (f == 1 ? t1[OP_E0] : t1[OP_V0]);
        T qeta = C * (f == 0 ? t0[OP_E0] : t0[OP_V0]) * (f == 1 ? t1[OP_E1] : t1[OP_V1]);
        if (q == 0) {
          Gh[1][1] += qxi;
          Gh[0][1] -= qeta;
        } else {
          Gh[0][0] += qeta;
          Gh[1][0] -= qxi;
        }
      }
      if (O.face_cons[f]) { T* o = O.face_cons[f] + 4 * lin; o[0] = Gh[0][0]; o[1] = Gh[1][0]; o[2] = Gh[0][1]; o[3] = Gh[1][1]; }
      if (O.cactive[f]) { T* o = O.cactive[f] + 2 * lin; o[0] = Gh[f][0]; o[1] = Gh[f][1]; }
      if (!O.face_fwd[f] && !O.face_inv[f]) continue;
      T Gn[2][2], Ln[2][2], Rn[2][2], Gf[2][2], Ff[2][2];
      mapped_jac<T, 2>(TM, ftab, L1, ix, f, F, Fp, Fpp);
      mapped_states2(F, Fp, Fpp, lam1, lam2, G, Lc, Rc);
      int64_t in[3] = {ix[0], ix[1], 0};
      in[f] += 1;
      mapped_jac<T, 2>(TM, ftab, L1, in, f, F, Fp, Fpp);
      mapped_states2(F, Fp, Fpp, lam1, lam2, Gn, Ln, Rn);
      for (int a = 0; a < 2; ++a)
        for (int b = 0; b < 2; ++b) Gf[a][b] = T(0.5) * (Lc[a][b] + Rn[a][b]);
      inv2(Gf, Ff);
      T J = Ff[0][0] * Ff[1][1] - Ff[0][1] * Ff[1][0];
      if (O.face_fwd[f]) { T* o = O.face_fwd[f] + 5 * lin; o[0] = Ff[0][0]; o[1] = Ff[1][0]; o[2] = Ff[0][1]; o[3] = Ff[1][1]; o[4] = J; }
      if (O.face_inv[f]) { T* o = O.face_inv[f] + 5 * lin; o[0] = Gf[0][0]; o[1] = Gf[1][0]; o[2] = Gf[0][1]; o[3] = Gf[1][1]; o[4] = J; }
    }
  }
}

// =============================== 1-D ========================================
template <class T>
__global__ void k_mapped_metrics1d(Dims D, MTerms TM, const T* __restrict__ ftab, int64_t L, MetricOut<T> O,
                                   int scheme, int what) {
  const T lam1 = lam1_of<T>(scheme), lam2 = lam2_of<T>(scheme);
  const int64_t L1 = L + 1;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < D.nw[0]; i += (int64_t)gridDim.x * blockDim.x) {
    T Fv[2] = {0, 0}, F1[2] = {0, 0}, F2[2] = {0, 0};
    for (int m = 0; m < TM.n; ++m)
      for (int c = 0; c < 2; ++c) {
        const T* a = ftab + ((int64_t)(m * 3) * L1 + i + c) * 4;
        const T cf = T(TM.t[m].coef);
        Fv[c] += cf * a[1];
        F1[c] += cf * a[2];
        F2[c] += cf * a[3];
      }
    if (what & 1) {
      if (O.cell_fwd) { O.cell_fwd[2 * i] = Fv[0]; O.cell_fwd[2 * i + 1] = Fv[0]; }
      if (O.cell_inv) { O.cell_inv[2 * i] = T(1) / Fv[0]; O.cell_inv[2 * i + 1] = Fv[0]; }
      if (O.cjac) O.cjac[i] = Fv[0];
    }
    if (what & 2) {
      T st[2][3];  // g, g', g'' with g = 1/F
      for (int c = 0; c < 2; ++c) {
        T g = T(1) / Fv[c];
        st[c][0] = g;
        st[c][1] = -F1[c] * g * g;
        st[c][2] = T(2) * F1[c] * F1[c] * g * g * g - F2[c] * g * g;
      }
      T Gf = T(0.5) * ((st[0][0] + lam1 * st[0][1] + lam2 * st[0][2]) + (st[1][0] - lam1 * st[1][1] + lam2 * st[1][2]));
      T Ff = T(1) / Gf;
      if (O.face_fwd[0]) { O.face_fwd[0][2 * i] = Ff; O.face_fwd[0][2 * i + 1] = Ff; }
      if (O.face_inv[0]) { O.face_inv[0][2 * i] = Gf; O.face_inv[0][2 * i + 1] = Ff; }
      // inv(F)*det(F) == 1 identically (metrics.jl:317-323)
      if (O.face_cons[0]) O.face_cons[0][i] = T(1);
      if (O.cactive[0]) O.cactive[0][i] = T(1);
    }
  }
}

// coordinates of a MappedGrid: the mapping evaluated at node / centroid / face-centre positions
// (generation.jl:187-309).  The mapping is separable, and all evaluation points are node positions (integer xi) or
// cell centres (xi + 1/2) per axis — the high-side face centre on axis a is the centre moved half a cell along a,
// i.e. the NODE position I_a + 1 there.  k_mapped_ctab tabulates every factor at the nodes and centres of its axis
// once per update! (O(N) sin / cos); the O(N^3) kernel is products of table values and 120 B of coalesced stores
// per cell (the first version evaluated ~90 sincos per node: 21 ms at 512^3).
//   ctab[((m*3 + d)*L1 + I)*2 + w],  w = 0: f_md(node I),  w = 1: f_md(centre I),  I in [0, nw_d]
template <class T>
__global__ void k_mapped_ctab(Dims D, MTerms TM, T* ctab, int64_t L1) {
  const int64_t ntot = (int64_t)TM.n * 3 * L1;
  for (int64_t lin = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; lin < ntot;
       lin += (int64_t)gridDim.x * blockDim.x) {
    const int64_t I = lin % L1;
    const int d = (int)((lin / L1) % 3);
    const int m = (int)(lin / (3 * L1));
    T* base = ctab + ((int64_t)(m * 3 + d) * L1 + I) * 2;
    if (d >= D.dim) {
      base[0] = base[1] = T(1);
      continue;
    }
    if (I > D.nw[d]) continue;
    const MTerm& t = TM.t[m];
    const double s = double(D.w0[d] + I + 1 - D.nhalo);  // node I of the window (generation.jl:209-221)
    base[0] = T(fac_der(t.kind[d], t.a[d], t.b[d], s, 0));
    base[1] = T(fac_der(t.kind[d], t.a[d], t.b[d], s + 0.5, 0));
  }
}

// one thread per node; blockIdx.y / .z are the node row and plane (no divisions), DIM is a template parameter so
// that every per-axis quantity stays in a register
template <class T, int DIM>
__global__ void __launch_bounds__(128) k_mapped_coords(Dims D, MTerms TM, const T* __restrict__ ctab, int L1,
                                                       CoordOut<T> O) {
  const int n0 = blockIdx.x * blockDim.x + threadIdx.x;
  const int n1 = DIM > 1 ? (int)blockIdx.y : 0, n2 = DIM > 2 ? (int)blockIdx.z : 0;
  const int nw0 = (int)D.nw[0], nw1 = DIM > 1 ? (int)D.nw[1] : 1, nw2 = DIM > 2 ? (int)D.nw[2] : 1;
  if (n0 > nw0) return;
  const int n[3] = {n0, n1, n2};
  const bool is_cell = n0 < nw0 && (DIM < 2 || n1 < nw1) && (DIM < 3 || n2 < nw2);
  const int64_t lin = n0 + (int64_t)(nw0 + 1) * (n1 + (int64_t)(DIM > 1 ? nw1 + 1 : 1) * n2);
  const int64_t clin = n0 + (int64_t)nw0 * (n1 + (int64_t)nw1 * n2);
  T qn[DIM], qc[DIM], qf[DIM][DIM];
#pragma unroll
  for (int c = 0; c < DIM; ++c) {
    qn[c] = qc[c] = T(0);
#pragma unroll
    for (int a = 0; a < DIM; ++a) qf[a][c] = T(0);
  }
  for (int m = 0; m < TM.n; ++m) {
    T fn[DIM], fc[DIM], fh[DIM];  // factor at node I, centre I, node I + 1 (the face position) per axis
#pragma unroll
    for (int d = 0; d < DIM; ++d) {
      const T* b = ctab + ((m * 3 + d) * L1 + n[d]) * 2;
      if constexpr (sizeof(T) == 8) {
        const double2 v = *reinterpret_cast<const double2*>(b);  // (node I, centre I): 16-B aligned pair
        fn[d] = v.x;
        fc[d] = v.y;
      } else {
        fn[d] = b[0];
        fc[d] = b[1];
      }
      fh[d] = b[2];  // node I + 1 (one entry of slack follows the last node)
    }
    const T c = T(TM.t[m].coef);
    const int q = TM.t[m].coord;
    T vn = c, vc = c;
#pragma unroll
    for (int d = 0; d < DIM; ++d) {
      vn *= fn[d];
      vc *= fc[d];
    }
    T vf[DIM];
#pragma unroll
    for (int a = 0; a < DIM; ++a) {
      vf[a] = c;
#pragma unroll
      for (int d = 0; d < DIM; ++d) vf[a] *= (d == a ? fh[d] : fc[d]);
    }
#pragma unroll
    for (int qq = 0; qq < DIM; ++qq)
      if (qq == q) {
        qn[qq] += vn;
        qc[qq] += vc;
#pragma unroll
        for (int a = 0; a < DIM; ++a) qf[a][qq] += vf[a];
      }
  }
#pragma unroll
  for (int c = 0; c < DIM; ++c) {
    if (O.node[c]) O.node[c][lin] = qn[c];
    if (is_cell) {
      if (O.centroid[c]) O.centroid[c][clin] = qc[c];
#pragma unroll
      for (int a = 0; a < DIM; ++a)
        if (O.face[a][c]) O.face[a][c][clin] = qf[a][c];
    }
  }
}
template <class T>
inline void launch_mapped_coords(const Dims& D, const MTerms& TM, const T* ctab, int64_t L1, const CoordOut<T>& O,
                                 cudaStream_t st) {
  const dim3 grid((unsigned)((D.nw[0] + 1 + 127) / 128), D.dim > 1 ? (unsigned)(D.nw[1] + 1) : 1u,
                  D.dim > 2 ? (unsigned)(D.nw[2] + 1) : 1u);
  if (D.dim == 3) k_mapped_coords<T, 3><<<grid, 128, 0, st>>>(D, TM, ctab, (int)L1, O);
  else if (D.dim == 2) k_mapped_coords<T, 2><<<grid, 128, 0, st>>>(D, TM, ctab, (int)L1, O);
  else k_mapped_coords<T, 1><<<grid, 128, 0, st>>>(D, TM, ctab, (int)L1, O);
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
struct MappedTables {
  bool dirty = true;
  int launches = 0;
  void* tab = nullptr;
  void* tabx = nullptr;   // xi rows of `tab`, transposed (one array per operator)
  void* ftab = nullptr;
  void* fhalf = nullptr;  // factor tables at centre + 1/2 (ADThomasLombardMetric only)
  void* ctab = nullptr;   // factor values at nodes and centres (coordinates)
  size_t tab_bytes = 0, ftab_bytes = 0, fhalf_bytes = 0, ctab_bytes = 0;
  bool ctab_dirty = true;
  int64_t L = 0;
  MTerms terms;
  MPairs pairs;
  void free_all() {
    if (tab) cudaFree(tab);
    if (tabx) cudaFree(tabx);
    if (ftab) cudaFree(ftab);
    if (fhalf) cudaFree(fhalf);
    if (ctab) cudaFree(ctab);
    tab = tabx = ftab = fhalf = ctab = nullptr;
    tab_bytes = ftab_bytes = fhalf_bytes = ctab_bytes = 0;
  }
};

inline int64_t mapped_lx(int64_t L) { return (L + 15) & ~int64_t(15); }  // row stride of the transposed xi table
// returns 0 ok, 1 too many terms, 2 too many pairs, 3 cuda error
template <class T>
inline int mapped_prepare(MappedTables& M, const std::vector<double>& table, double t, const Dims& D, int scheme,
                          cudaStream_t st, bool want_half = false) {
  if (!M.dirty && M.tab) return 0;
  const int nt = (int)(table.size() / 15);
  if (nt + 1 > M_MAXT) return 1;
  M.terms.n = nt;
  for (int m = 0; m < nt; ++m) {
    const double* r = &table[15 * m];
    MTerm& T1 = M.terms.t[m];
    T1.coord = (int)r[0];
    T1.coef = r[1] + r[2] * t + r[3] * std::sin(r[4] * t + r[5]);
    for (int d = 0; d < 3; ++d) {
      T1.kind[d] = d < D.dim ? (int)r[6 + 3 * d] : 0;
      T1.a[d] = r[7 + 3 * d];
      T1.b[d] = r[8 + 3 * d];
    }
  }
  // unit term (all factors 1) used as q2 in the 2-D pair list
  MTerm& U = M.terms.t[nt];
  U.coord = 0;
  U.coef = 1.0;
  for (int d = 0; d < 3; ++d) { U.kind[d] = 0; U.a[d] = U.b[d] = 0.0; }
  M.pairs.n = 0;
  for (int c = 0; c < 4; ++c) M.pairs.cstart[c] = 0;
  if (D.dim == 3) {
    for (int col = 0; col < 3; ++col) {
      const int q1 = (col + 1) % 3, q2 = (col + 2) % 3;
      M.pairs.cstart[col] = M.pairs.n;
      for (int m1 = 0; m1 < nt; ++m1)
        for (int m2 = 0; m2 < nt; ++m2)
          if (M.terms.t[m1].coord == q1 && M.terms.t[m2].coord == q2) {
            if (M.pairs.n >= M_MAXP) return 2;
            M.pairs.p[M.pairs.n++] = MPair{m1, m2, col};
          }
    }
    M.pairs.cstart[3] = M.pairs.n;
    for (int p = 0; p < M.pairs.n; ++p) M.pairs.coef[p] = M.terms.t[M.pairs.p[p].m1].coef * M.terms.t[M.pairs.p[p].m2].coef;
    M.pairs.coef[M.pairs.n] = 0.0;
  } else if (D.dim == 2) {
    for (int m1 = 0; m1 < nt; ++m1) {
      if (M.pairs.n >= M_MAXP) return 2;
      M.pairs.p[M.pairs.n++] = MPair{m1, nt, 0};
    }
  }
  int64_t L = 1;
  for (int d = 0; d < D.dim; ++d) L = D.nw[d] > L ? D.nw[d] : L;
  M.L = L;
  const size_t tb = (size_t)(M.pairs.n > 0 ? M.pairs.n : 1) * 3 * OP_COUNT * L * sizeof(T);
  const size_t fb = (size_t)nt * 3 * 4 * (L + 1) * sizeof(T);
  if (tb > M.tab_bytes) {
    if (M.tab) cudaFree(M.tab);
    if (M.tabx) cudaFree(M.tabx);
    M.tab = M.tabx = nullptr;
    M.tab_bytes = 0;
    if (cudaMalloc(&M.tab, tb) != cudaSuccess) return 3;
    if (cudaMalloc(&M.tabx, (tb / 3 / L) * mapped_lx(L)) != cudaSuccess) return 3;
    M.tab_bytes = tb;
  }
  if (fb > M.ftab_bytes) {
    if (M.ftab) cudaFree(M.ftab);
    if (cudaMalloc(&M.ftab, fb) != cudaSuccess) return 3;
    M.ftab_bytes = fb;
  }
  const double lam1 = scheme == ENDPOINT ? 0.0 : 0.5, lam2 = scheme == CURVATURE ? 1.0 / 12.0 : 0.0;
  MTerms TMd = M.terms;
  TMd.n = nt;  // device loops over real terms only; the unit term is addressed by index nt
  if (M.pairs.n > 0) {
    const int64_t n = (int64_t)M.pairs.n * 3 * L;
    k_mapped_tables<T><<<(unsigned)((n + 127) / 128), 128, 0, st>>>(D, TMd, M.pairs, lam1, lam2, (T*)M.tab, L, D.dim == 3 ? (T*)M.tabx : nullptr, mapped_lx(L));
    M.launches++;
  }
  {
    const int64_t n = (int64_t)nt * 3 * (L + 1);
    k_mapped_factors<T><<<(unsigned)((n + 127) / 128), 128, 0, st>>>(D, TMd, (T*)M.ftab, L + 1);
    M.launches++;
    if (want_half) {
      if (fb > M.fhalf_bytes) {
        if (M.fhalf) cudaFree(M.fhalf);
        if (cudaMalloc(&M.fhalf, fb) != cudaSuccess) return 3;
        M.fhalf_bytes = fb;
      }
      k_mapped_factors<T><<<(unsigned)((n + 127) / 128), 128, 0, st>>>(D, TMd, (T*)M.fhalf, L + 1, 0.5);
      M.launches++;
    }
  }
  if (cudaGetLastError() != cudaSuccess) return 3;
  M.dirty = false;
  M.ctab_dirty = true;
  return 0;
}

// the coordinate tables of the current terms (after mapped_prepare); returns 0 ok, 3 cuda error
template <class T>
inline int mapped_prepare_coords(MappedTables& M, const Dims& D, cudaStream_t st) {
  if (!M.ctab_dirty && M.ctab) return 0;
  const int nt = M.terms.n;
  const int64_t L1 = M.L + 1;
  const size_t cb = (size_t)nt * 3 * 2 * (L1 + 1) * sizeof(T);  // one entry of slack: the kernel reads node I + 1
  if (cb > M.ctab_bytes) {
    if (M.ctab) cudaFree(M.ctab);
    if (cudaMalloc(&M.ctab, cb) != cudaSuccess) return 3;
    M.ctab_bytes = cb;
  }
  const int64_t n = (int64_t)nt * 3 * L1;
  k_mapped_ctab<T><<<(unsigned)((n + 127) / 128), 128, 0, st>>>(D, M.terms, (T*)M.ctab, L1);
  M.launches++;
  if (cudaGetLastError() != cudaSuccess) return 3;
  M.ctab_dirty = false;
  return 0;
}

}  // namespace cg

// Marching kernels for the 3-D DiscreteGrid metric fill (the fast path) and the
// coalesced 2-D kernel.
//
// Replaces /root/reference/src/grids/unified_grids/metrics.jl:1850-1877 (cell
// fill) and :1962-2053 (three face fills) with three warp-autonomous kernels, one
// per face axis (the xi kernel also writes the cell metrics):
//   * a warp owns 29 (xi kernel) or 31 (eta / zeta kernels) consecutive cells of one row; the remaining lanes are
//     the -1 / +1,+2 neighbours the stencils need; it marches along k (the eta kernel along j);
//   * xi-neighbours come from warp shuffles, the neighbours along the marching axis are carried in registers from
//     the previous step or looked ahead one plane, the third axis is recomputed from the node rows;
//   * nodes arrive through a per-warp cp.async ring in shared memory, outputs leave through per-warp staging
//     buffers and TMA bulk stores; no block barriers: warps never wait for each other.
// Math: DESIGN.md §3 (edge scalars Psi + 4-cell line stencils), same closed forms
// as tests/closed_form_numpy.py and the gather kernels in cg_kernels_ref.cuh.
#pragma once
#include <cuda_runtime.h>

#include <cstdlib>

#include "cg_kernels_ref.cuh"

namespace cg {

#define CG_D __device__ __forceinline__
#ifndef CG_MINB
#define CG_MINB 2
#endif
#ifndef CG_ZETA_MINB
#define CG_ZETA_MINB 2
#endif
#ifndef CG_COMPACT_MINB
#define CG_COMPACT_MINB 3
#endif
#ifndef CG_XI_COMPACT_MINB
#define CG_XI_COMPACT_MINB 2  // blocks (of 4 warps) per SM of the COMPACT xi kernel: at 3 (168 registers) it spills since the
                              // forms take first differences between neighbouring nodes (6.2 -> 5.0 ms at 512^3)
#endif
#ifndef CG_WY
#define CG_WY 4
#endif
constexpr int WY = CG_WY;       // warps (rows) per block of the marching kernels
constexpr int NT = 32 * WY;     // threads per block
#ifndef CG_ZWY
#define CG_ZWY 1
#endif
// The eta / zeta kernels run ONE warp per block: row index, chunk and every staging / TMA address are then functions
// of blockIdx alone, i.e. they live in uniform registers (UBLKCP straight from UR operands, no R2UR loops).
constexpr int ZWY = CG_ZWY;
#ifndef CG_PICK_CHUNK
#define CG_PICK_CHUNK 0  // 1 = wave-aware chunk lengths (pick_chunk; mixed results: 384^3 +2 %, 512^3 and 300^3 -1..2 %); 0 = the ~30-waves policy
#endif
#ifndef CG_XWY
#define CG_XWY 4
#endif
constexpr int XWY = CG_XWY;  // warps per block of the xi kernel
constexpr int XNT = 32 * XWY;
#ifndef CG_JFAST
#define CG_JFAST 0
#endif
// one-warp blocks: 1 = blockIdx.x counts rows (fastest), blockIdx.y the xi segments, so the blocks resident on one
// SM are neighbouring rows and share node rows in L1
constexpr int JFAST = CG_JFAST;
#ifndef CG_Z_UNROLL
#define CG_Z_UNROLL 1
#endif
constexpr int Z_UNROLL = CG_Z_UNROLL;  // unroll factor of the eta / zeta marching loop
#ifndef CG_Z_SLOTS
#define CG_Z_SLOTS 3
#endif
#ifndef CG_Z_DBUF
#define CG_Z_DBUF 0
#endif
// eta / zeta kernels: node ring slots (3: the plane needed in the next step is fetched now; 4: one more step of
// slack) and double-buffered output staging (the TMA unit may still read last step's chunks while this step stages)
constexpr int Z_SLOTS = CG_Z_SLOTS;
constexpr int Z_DBUF = CG_Z_DBUF;
constexpr int ZNT = 32 * ZWY;
constexpr int WARP_OUT = 29;  // output cells per warp along xi (kernels that need lanes -1, +1, +2)
constexpr int ZETA_OUT = 31;  // the zeta / eta(SWAP) kernel only needs lane -1
constexpr unsigned FULLMASK = 0xffffffffu;
// MODE: the two hot configurations are compiled with every output switch as a compile-time constant, so the
// loop body is straight-line code the scheduler can interleave freely and unused work disappears:
//   MODE_FULL    : FULL profile, cell + face outputs (all nine AoS arrays)
//   MODE_COMPACT : COMPACT profile, cell + face outputs (J of the cell, active conserved row per axis)
//   MODE_FACE    : FULL profile, face outputs only  (refresh_face_metrics!; xi kernel only — the eta / zeta kernels
//                  never write cell arrays and use MODE_FULL)
//   MODE_CELL    : FULL profile, cell outputs only  (refresh_cell_metrics!; xi kernel only)
//   MODE_GENERIC : anything else, switches read at run time
enum { MODE_GENERIC = 0, MODE_FULL = 1, MODE_COMPACT = 2, MODE_FACE = 3, MODE_CELL = 4 };

template <class T>
struct PF {
  Form4<T> q[3];
};
#ifndef CG_IX
#define CG_IX uint32_t
#endif
// node index type of the marching kernels: 32-bit (launchers fall back to the gather kernels when the padded
// node array has >= 2^31 elements); one IMAD.WIDE per address instead of 64-bit add chains
using Ix = CG_IX;
template <class T, class IX = Ix>
struct NodeP {
  using ix = IX;
  const T* q[3];
  IX s1, s2;
};

template <class T> CG_D T ldn(const T* p) { return __ldg(p); }
// RAW bilinear form of 4 face nodes n[bu][bv] (no normalisation: pure sums and differences)
//   c0 = 4 * (true c0), cu = 2 * (true cu), cv = 2 * (true cv), cuv = true cuv
// Every consumer below is bilinear in form coefficients, so the powers of two fold into its constants.
// First differences are taken between neighbouring nodes (cg_math.cuh: round-off rule), never between node sums.
template <class T>
CG_D Form4<T> raw_form(T n00, T n10, T n01, T n11) {
  T s0 = n10 + n00, d0 = n10 - n00, s1 = n11 + n01, d1 = n11 - n01;
  return {s0 + s1, d0 + d1, (n01 - n00) + (n11 - n10), d1 - d0};
}
// zeta-plane forms (fixed k), transverse (xi, eta)
template <class T, class IX>
CG_D PF<T> zplane(const NodeP<T, IX>& N, IX b) {
  PF<T> f;
#pragma unroll
  for (int q = 0; q < 3; ++q) {
    const T* A = N.q[q] + b;
    f.q[q] = raw_form(ldn(A), ldn(A + 1), ldn(A + N.s1), ldn(A + N.s1 + 1));
  }
  return f;
}
// sum / difference of a node column across two consecutive zeta planes, per coordinate
template <class T>
struct ZPair {
  T s[3], d[3], lo[3];  // sum, difference (exact) and lower node of the column
};
template <class T, class IX>
CG_D ZPair<T> zpair(const NodeP<T, IX>& N, IX b) {
  ZPair<T> p;
#pragma unroll
  for (int q = 0; q < 3; ++q) {
    const T lo = ldn(N.q[q] + b), hi = ldn(N.q[q] + b + N.s2);
    p.s[q] = hi + lo;
    p.d[q] = hi - lo;
    p.lo[q] = lo;
  }
  return p;
}
// raw form of the plane through two node columns a, b (u = from a to b, v = zeta): an eta-plane from two
// xi-adjacent columns, a xi-plane from two eta-adjacent columns; 5 operations per coordinate.  The u-difference
// (b.hi - a.hi) + (b.lo - a.lo) is formed as 2 (b.lo - a.lo) + (b.d - a.d): neighbour differences only.
template <class T>
CG_D PF<T> pair_form(const ZPair<T>& a, const ZPair<T>& b) {
  PF<T> f;
#pragma unroll
  for (int q = 0; q < 3; ++q) {
    const T cuv = b.d[q] - a.d[q];
    f.q[q] = {a.s[q] + b.s[q], fma(T(2), b.lo[q] - a.lo[q], cuv), a.d[q] + b.d[q], cuv};
  }
  return f;
}
// ---------------------------------------------------------------------------
// Per-warp ring of node planes in shared memory.  The marching kernels run at 8 warps/SM, too few to hide a
// global-load miss, and prefetch.global.L1 does not shorten it (measured).  The plane that is first needed in the
// NEXT step is therefore copied with cp.async, asynchronously, while the current step computes; all node reads of
// a step are LDS.  Layout [slot][row][coordinate][column]: lanes read consecutive columns (conflict-free).
// ---------------------------------------------------------------------------
CG_D void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
CG_D void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// A plane row is fetched with 16-byte cp.async by the first NL lanes.  The padded node arrays have a row pitch that is a multiple of 16 bytes (cg_api.cu: dims()), so the
// row segment starting at the column i0 rounded DOWN to a 16-byte boundary is aligned; `ph` = i0 mod VEC is the
// warp-uniform column phase every read adds.  Columns past the end of a row belong to lanes without output.
template <class T, int ROWS, int SLOTS>
struct RingV {
  static constexpr int VEC = 16 / (int)sizeof(T);
  static constexpr int NL = (34 + 2 * (VEC - 1)) / VEC;  // lanes that copy
  static constexpr int COLS = VEC * NL;
  static constexpr int VALS = SLOTS * ROWS * 3 * COLS;
  T* base;  // already shifted by the phase
  CG_D T* row(int plane, int r, int q) const {
    const int slot = SLOTS == 4 ? (plane & 3) : plane % SLOTS;
    return base + ((slot * ROWS + r) * 3 + q) * COLS;
  }
  // g0 = index of (aligned first column + VEC * lane, ring row 0, that plane); base_unshifted = base - ph
  template <class IX>
  CG_D void fetch(const NodeP<T, IX>& N, int plane, IX g0, int lane, int ph) const {
    if (lane < NL) {
#pragma unroll
      for (int r = 0; r < ROWS; ++r)
#pragma unroll
        for (int q = 0; q < 3; ++q) {
          const uint32_t d = (uint32_t)__cvta_generic_to_shared(row(plane, r, q) - ph + VEC * lane);
          asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(d), "l"(N.q[q] + g0 + (IX)r * N.s1) : "memory");
        }
    }
  }
  CG_D PF<T> zplane(int plane, int r, int lane) const {
    PF<T> f;
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      const T* a = row(plane, r, q) + lane;
      const T* b = row(plane, r + 1, q) + lane;
      f.q[q] = raw_form(a[0], a[1], b[0], b[1]);
    }
    return f;
  }
  CG_D ZPair<T> zpair(int plane, int r, int c) const {
    ZPair<T> p;
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      const T lo = row(plane, r, q)[c], hi = row(plane + 1, r, q)[c];
      p.s[q] = hi + lo;
      p.d[q] = hi - lo;
      p.lo[q] = lo;
    }
    return p;
  }
};

template <class T>
CG_D PF<T> swap_uv(const PF<T>& p) {
  PF<T> o;
#pragma unroll
  for (int q = 0; q < 3; ++q) o.q[q] = {p.q[q].c0, p.q[q].cv, p.q[q].cu, p.q[q].cuv};
  return o;
}

// cell Taylor coefficients (true values) from the two RAW zeta-planes of the cell (bitmask: 1 xi, 2 eta, 4 zeta);
// dz[q] = sum of the cell's four zeta-edge differences (the first zeta-difference is NOT taken between the plane sums:
// round-off rule, cg_math.cuh)
template <class T>
CG_D void taylor_from_z(const PF<T>& lo, const PF<T>& hi, const T (&dz)[3], T (&c)[3][8]) {
#pragma unroll
  for (int q = 0; q < 3; ++q) {
    c[q][0] = T(0.125) * (lo.q[q].c0 + hi.q[q].c0);
    c[q][1] = T(0.25) * (lo.q[q].cu + hi.q[q].cu);
    c[q][2] = T(0.25) * (lo.q[q].cv + hi.q[q].cv);
    c[q][3] = T(0.5) * (lo.q[q].cuv + hi.q[q].cuv);
    c[q][4] = T(0.25) * dz[q];
    c[q][5] = T(0.5) * (hi.q[q].cu - lo.q[q].cu);
    c[q][6] = T(0.5) * (hi.q[q].cv - lo.q[q].cv);
    c[q][7] = hi.q[q].cuv - lo.q[q].cuv;
  }
}
// Half states W+-_f of the six E^a polynomials on the a-face between cells c and
// c+a.  P0/P1/P2: RAW forms on the low face of c, the shared face, the high face of
// c+a, Form4 ordered (u = f, v = t).  Raw products carry 8 (cu c0, cv c0), 4 (cu cu, cv cu, cuv c0) or
// 2 (cuv cu) times the true value; the factors are folded into the reconstruction weights.  Index = bsel*3 + col, bsel 0: b = f, 1: b = t;
// col selects (q1,q2) = (col+1, col+2) mod 3.       
template <class T, int SCH>
CG_D void eface6(const PF<T>& P0, const PF<T>& P1, const PF<T>& P2, T (&wp)[6], T (&wm)[6]) {
  if (SCH == ENDPOINT) {
    // E = 1/2 (P_c + P_{c+a}) with cell-centre values; W+ = W- = E/2
    T q0c[3], q0n[3], qfc[3], qfn[3], qtc[3], qtn[3];  // raw sums: 8 x and 4 x the cell-centre values
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      q0c[q] = P0.q[q].c0 + P1.q[q].c0;
      q0n[q] = P1.q[q].c0 + P2.q[q].c0;
      qfc[q] = P0.q[q].cu + P1.q[q].cu;
      qfn[q] = P1.q[q].cu + P2.q[q].cu;
      qtc[q] = P0.q[q].cv + P1.q[q].cv;
      qtn[q] = P1.q[q].cv + P2.q[q].cv;
    }
#pragma unroll
    for (int col = 0; col < 3; ++col) {
      const int q1 = (col + 1) % 3, q2 = (col + 2) % 3;
      wp[col] = wm[col] = T(0.25 / 32) * (qfc[q1] * q0c[q2] + qfn[q1] * q0n[q2]);
      wp[3 + col] = wm[3 + col] = T(0.25 / 32) * (qtc[q1] * q0c[q2] + qtn[q1] * q0n[q2]);
    }
    return;
  }
  const T lam1 = T(0.5), lam2 = SCH == CURVATURE ? T(1.0 / 12.0) : T(0);
  const T k2 = T(0.125) - lam2;  // 2*kappa
  Form4<T> dA[3], dB[3];
#pragma unroll
  for (int q = 0; q < 3; ++q) {
    dA[q] = form_dif(P1.q[q], P0.q[q]);
    dB[q] = form_dif(P2.q[q], P1.q[q]);
  }
#pragma unroll
  for (int col = 0; col < 3; ++col) {
    const int q1 = (col + 1) % 3, q2 = (col + 2) % 3;
    const Form4<T>&A1 = P1.q[q1], &B1 = P1.q[q2];
    const Form4<T>&a0 = dA[q1], &b0 = dA[q2], &a1 = dB[q1], &b1 = dB[q2];
    // b = f
    T e0 = A1.cu * B1.c0 - k2 * (a0.cu * b0.c0 + a1.cu * b1.c0);
    T e1 = A1.cu * B1.cu - k2 * (a0.cu * b0.cu + a1.cu * b1.cu);
    T h0 = T(0.0625) * e0, h1 = (T(0.125) * lam1) * e1;  // raw e0 = 8 x, raw e1 = 4 x
    wp[col] = h0 + h1;
    wm[col] = h0 - h1;
    // b = t
    e0 = A1.cv * B1.c0 - k2 * (a0.cv * b0.c0 + a1.cv * b1.c0);
    e1 = A1.cv * B1.cu + A1.cuv * B1.c0 - k2 * (a0.cv * b0.cu + a0.cuv * b0.c0 + a1.cv * b1.cu + a1.cuv * b1.c0);
    T e2 = A1.cuv * B1.cu - k2 * (a0.cuv * b0.cu + a1.cuv * b1.cu);
    h0 = T(0.0625) * e0 + (T(0.5) * lam2) * e2;  // raw e0 = 8 x, e1 = 4 x, e2 = 2 x
    h1 = (T(0.125) * lam1) * e1;
    wp[3 + col] = h0 + h1;
    wm[3 + col] = h0 - h1;
  }
}

// W+-( E on the a-face between rows 1|2 ) - W+-( E on the a-face between rows 0|1 ) for four consecutive
// a-planes P0..P3: the second a-derivative of the middle cell cancels, so the difference needs two face
// products and two cell products instead of two and four (same index convention as eface6).
template <class T, int SCH>
CG_D void eface6_diff(const PF<T>& P0, const PF<T>& P1, const PF<T>& P2, const PF<T>& P3, T (&wp)[6], T (&wm)[6]) {
  if (SCH == ENDPOINT) {
    // E = 1/2 (P_c + P_{c+a}) with cell-centre values: difference = 1/2 (P_{c+a} - P_{c-a}); W+ = W- = E/2
    T q0a[3], q0b[3], qfa[3], qfb[3], qta[3], qtb[3];  // raw sums: 8 x and 4 x the cell-centre values
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      q0a[q] = P0.q[q].c0 + P1.q[q].c0;
      q0b[q] = P2.q[q].c0 + P3.q[q].c0;
      qfa[q] = P0.q[q].cu + P1.q[q].cu;
      qfb[q] = P2.q[q].cu + P3.q[q].cu;
      qta[q] = P0.q[q].cv + P1.q[q].cv;
      qtb[q] = P2.q[q].cv + P3.q[q].cv;
    }
#pragma unroll
    for (int col = 0; col < 3; ++col) {
      const int q1 = (col + 1) % 3, q2 = (col + 2) % 3;
      wp[col] = wm[col] = T(0.25 / 32) * (qfb[q1] * q0b[q2] - qfa[q1] * q0a[q2]);
      wp[3 + col] = wm[3 + col] = T(0.25 / 32) * (qtb[q1] * q0b[q2] - qta[q1] * q0a[q2]);
    }
    return;
  }
  const T lam1 = T(0.5), lam2 = SCH == CURVATURE ? T(1.0 / 12.0) : T(0);
  const T k2 = T(0.125) - lam2;
  Form4<T> dL[3], dH[3];
#pragma unroll
  for (int q = 0; q < 3; ++q) {
    dL[q] = form_dif(P1.q[q], P0.q[q]);
    dH[q] = form_dif(P3.q[q], P2.q[q]);
  }
#pragma unroll
  for (int col = 0; col < 3; ++col) {
    const int q1 = (col + 1) % 3, q2 = (col + 2) % 3;
    const Form4<T>&A1 = P1.q[q1], &B1 = P1.q[q2], &A2 = P2.q[q1], &B2 = P2.q[q2];
    const Form4<T>&al = dL[q1], &bl = dL[q2], &ah = dH[q1], &bh = dH[q2];
    T e0 = (A2.cu * B2.c0 - A1.cu * B1.c0) - k2 * (ah.cu * bh.c0 - al.cu * bl.c0);
    T e1 = (A2.cu * B2.cu - A1.cu * B1.cu) - k2 * (ah.cu * bh.cu - al.cu * bl.cu);
    T h0 = T(0.0625) * e0, h1 = (T(0.125) * lam1) * e1;  // raw e0 = 8 x, raw e1 = 4 x
    wp[col] = h0 + h1;
    wm[col] = h0 - h1;
    e0 = (A2.cv * B2.c0 - A1.cv * B1.c0) - k2 * (ah.cv * bh.c0 - al.cv * bl.c0);
    e1 = ((A2.cv * B2.cu + A2.cuv * B2.c0) - (A1.cv * B1.cu + A1.cuv * B1.c0)) -
         k2 * ((ah.cv * bh.cu + ah.cuv * bh.c0) - (al.cv * bl.cu + al.cuv * bl.c0));
    T e2 = (A2.cuv * B2.cu - A1.cuv * B1.cu) - k2 * (ah.cuv * bh.cu - al.cuv * bl.cu);
    h0 = T(0.0625) * e0 + (T(0.5) * lam2) * e2;  // raw e0 = 8 x, e1 = 4 x, e2 = 2 x
    h1 = (T(0.125) * lam1) * e1;
    wp[3 + col] = h0 + h1;
    wm[3 + col] = h0 - h1;
  }
}

template <class T> CG_D T shfl_dn(T v, int d) { return __shfl_down_sync(FULLMASK, v, d); }
template <class T> CG_D T shfl_up(T v, int d) { return __shfl_up_sync(FULLMASK, v, d); }

// line scalars of the potential (d_b q1) q2 along axis f from the cell's Taylor coefficients
template <class T>
CG_D LineS<T> line_from_taylor(const T (&c)[3][8], int q1, int q2, int f, int b, T lam1, T lam2) {
  const int mb = 1 << b, mf = 1 << f;
  T p0 = c[q1][mb] * c[q2][0];
  T p1 = c[q1][mb] * c[q2][mf] + c[q1][mb | mf] * c[q2][0];
  T p2 = c[q1][mb | mf] * c[q2][mf];
  return line_scalars(p0, p1, p2, lam1, lam2);
}

// RAW Taylor sums of the cell from its two RAW eta-planes: craw[q][S] = 2^(3 - |S|) * (true coefficient), pure
// sums / differences.  Every consumer in the eta / zeta kernels is a product of two coefficients (line scalars) or
// homogeneous in F (inverse Jacobian states), so the powers of two fold into compile-time weights.
// de[q] = sum of the cell's four eta-edge differences (round-off rule, cg_math.cuh).
template <class T>
CG_D void taylor_raw_e(const PF<T>& lo, const PF<T>& hi, const T (&de)[3], T (&c)[3][8]) {
#pragma unroll
  for (int q = 0; q < 3; ++q) {
    c[q][0] = lo.q[q].c0 + hi.q[q].c0;
    c[q][1] = lo.q[q].cu + hi.q[q].cu;
    c[q][2] = de[q];
    c[q][3] = hi.q[q].cu - lo.q[q].cu;
    c[q][4] = lo.q[q].cv + hi.q[q].cv;
    c[q][5] = lo.q[q].cuv + hi.q[q].cuv;
    c[q][6] = hi.q[q].cv - lo.q[q].cv;
    c[q][7] = hi.q[q].cuv - lo.q[q].cuv;
  }
}
// line scalars from RAW Taylor sums: p0 carries 32 x, p1 16 x, p2 8 x the true value
template <class T>
CG_D LineS<T> line_from_taylor_raw(const T (&c)[3][8], int q1, int q2, int f, int b, T lam1, T lam2) {
  const int mb = 1 << b, mf = 1 << f;
  const T p0 = c[q1][mb] * c[q2][0];
  const T p1 = c[q1][mb] * c[q2][mf] + c[q1][mb | mf] * c[q2][0];
  const T p2 = c[q1][mb | mf] * c[q2][mf];
  const T a = (T(2) * lam2 + lam1 * lam1) * T(0.125), bq = (T(2) * lam2 - lam1 * lam1) * T(0.125), l1 = lam1 * T(0.0625);
  LineS<T> s;
  const T base = T(0.5 / 32) * p0 + a * p2;
  s.vp = base + l1 * p1;
  s.vm = base - l1 * p1;
  s.u = T(0.5 / 32) * p0 + bq * p2;
  return s;
}

// ---------------------------------------------------------------------------
// Output staging.  The metric arrays are AoS (80-B / 72-B elements, layout of the Julia arrays): a lane
// storing its own element touches one 128-B line per lane and instruction, which saturates the L1TEX
// wavefront queue (measured: round 1, profiles/).  The cells a warp writes in one step are consecutive
// along xi, i.e. ONE contiguous chunk per array.  Every lane puts its element into the warp's staging
// buffer in shared memory and lane 0 hands the chunk to the TMA unit (cp.async.bulk shared -> global):
// no global-store instructions, full-sector writes.  Chunks whose ends are not 16-B aligned (72-B
// elements, FP32) keep the same 16-B phase in shared memory; the <= 3 head / tail values go out as
// plain stores.
// ---------------------------------------------------------------------------
CG_D void bulk_store(void* gdst, const void* ssrc, uint32_t bytes) {
#ifdef CG_NO_STORE  // experiment: time the kernels without their bulk stores
  if (bytes != 0xffffffffu) return;
#endif
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst),
               "r"((uint32_t)__cvta_generic_to_shared(ssrc)), "r"(bytes)
               : "memory");
}
CG_D void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// all bulk stores of this thread have finished READING shared memory (the staging buffers may be rewritten)
CG_D void bulk_wait_read() {
#ifndef CG_NO_WAIT  // experiment: results are wrong without the wait, only the timing is of interest
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
#endif
}
// ... all but the most recent group (double-buffered staging: the group of two steps ago used this buffer)
CG_D void bulk_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
CG_D void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

template <class T, int E> constexpr int stage_len = 32 * E + 16 / (int)sizeof(T);  // values per array

// staging slot of one array for one warp step: where the chunk goes in global memory and in shared memory
template <class T, int E>
struct Stage {
  T* g;   // global address of the chunk's first value
  T* s;   // shared address of the chunk's first value (same 16-B phase as g)
  CG_D Stage(T* arr, int64_t cl0, T* sbuf) {
#ifdef CG_STORE_WRAP  // experiment: every store lands in the first 16 MB of its array (L2-resident): SM-side store cost
    cl0 &= (int64_t)0x1ffff;
#endif
    g = arr + (int64_t)E * cl0;
    constexpr int A = 16 / (int)sizeof(T);
    if constexpr ((E * sizeof(T)) % 16 == 0) s = sbuf;  // array bases are 16-B aligned (cudaMalloc / checked on bind)
    else s = sbuf + (int)((reinterpret_cast<uintptr_t>(g) / sizeof(T)) & (A - 1));
  }
  CG_D T* slot(int rel) const { return s + rel * E; }
  // hand the chunk to the TMA unit.  All lanes of the warp call this AFTER the valid lanes have written
  // their slots, executed fence_async_smem() and passed a __syncwarp().
  CG_D void issue(int nvalid, int lane) const {
    constexpr int A = 16 / (int)sizeof(T);
    if constexpr ((E * sizeof(T)) % 16 == 0) {  // whole elements are 16-B multiples: no head / tail
      if (lane == 0) bulk_store(g, s, (uint32_t)(nvalid * E * sizeof(T)));
      return;
    }
    const int nvals = nvalid * E;
    const int head = (int)((A - (reinterpret_cast<uintptr_t>(g) / sizeof(T))) & (A - 1));
    const int tail = (nvals - head) & (A - 1);
    const int body = nvals - head - tail;
    if (lane == 0 && body > 0) bulk_store(g + head, s + head, (uint32_t)(body * sizeof(T)));
    if (lane < head) g[lane] = s[lane];
    if (lane < tail) g[head + body + lane] = s[head + body + lane];
  }
  CG_D void flush(int nvalid, int lane) const {
    fence_async_smem();
    __syncwarp();
    issue(nvalid, lane);
  }
};

template <class T>
CG_D void put10(T* o, const T (&M)[3][3], T J) {  // o: staging slot (16-B aligned for FP64)
  if constexpr (sizeof(T) == 8) {
    double2* p = reinterpret_cast<double2*>(o);
    p[0] = make_double2(M[0][0], M[1][0]);
    p[1] = make_double2(M[2][0], M[0][1]);
    p[2] = make_double2(M[1][1], M[2][1]);
    p[3] = make_double2(M[0][2], M[1][2]);
    p[4] = make_double2(M[2][2], J);
  } else {
    o[0] = M[0][0]; o[1] = M[1][0]; o[2] = M[2][0];
    o[3] = M[0][1]; o[4] = M[1][1]; o[5] = M[2][1];
    o[6] = M[0][2]; o[7] = M[1][2]; o[8] = M[2][2];
    o[9] = J;
  }
}
template <class T>
CG_D void put9(T* o, const T (&M)[3][3]) {
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int r = 0; r < 3; ++r) o[r + 3 * c] = M[r][c];
}

// ===========================================================================
// Kernel A: cell metrics + xi-face metrics.  f = xi, s = eta, t = zeta.
// ===========================================================================
// dynamic shared memory: per warp, staging buffers of the five AoS arrays this kernel writes
// per-warp shared memory: [node ring | staging buffers]; the COMPACT configuration stages nothing
template <class T, int MODE> constexpr int xi_warp_vals =
    RingV<T, 4, 4>::VALS + (MODE == MODE_COMPACT ? 0 : 4 * stage_len<T, 10> + stage_len<T, 9>);
template <class T, int SCH, int MODE>
__global__ void __launch_bounds__(XNT, (MODE == MODE_COMPACT ? CG_XI_COMPACT_MINB : CG_MINB) * 4 / XWY) k_face_xi(Dims D, const T* ex, const T* ey, const T* ez,
                                                         MetricOut<T> O, int what, int kchunk) {
  extern __shared__ __align__(128) unsigned char cg_dyn_smem[];
  const int lane = threadIdx.x;
  const int64_t nw0 = D.nw[0], nw1 = D.nw[1];
  const unsigned bx = (XWY == 1 && JFAST) ? blockIdx.y : blockIdx.x, by = (XWY == 1 && JFAST) ? blockIdx.x : blockIdx.y;
  const int64_t j = XWY == 1 ? (int64_t)by : (int64_t)by * blockDim.y + threadIdx.y;
  if (j >= nw1) return;  // whole warp
  const int64_t i0 = (int64_t)bx * WARP_OUT;  // first output cell of the warp (lane 1)
  const int64_t i = i0 + lane - 1;
  const int64_t ic = i > nw0 + 1 ? nw0 + 1 : i;
  const bool out_ok = lane >= 1 && lane <= WARP_OUT && i < nw0;
  const int nvalid = (int)(nw0 - i0 < WARP_OUT ? nw0 - i0 : WARP_OUT);
  const int rel = lane - 1;
  // staging slot of this lane: every lane writes (no branch); the slots 29 .. 31 are never sent
  const int srel = lane == 0 ? 31 : lane - 1;
  const int64_t k0 = D.r0 + (int64_t)blockIdx.z * kchunk;  // planes [D.r0, D.r1) of the window
  const int64_t k1 = k0 + kchunk < D.r1 ? k0 + kchunk : D.r1;
  NodeP<T> N;
  N.q[0] = ex; N.q[1] = ey; N.q[2] = ez;
  N.s1 = (Ix)D.se[1]; N.s2 = (Ix)D.se[2];
  const T lam1 = lam1_of<T>(SCH), lam2 = lam2_of<T>(SCH);
  constexpr bool ALL = MODE == MODE_FULL, CMP = MODE == MODE_COMPACT, FACE = MODE == MODE_FACE, CELL = MODE == MODE_CELL;
  constexpr bool GEN = MODE == MODE_GENERIC;
  const bool full = ALL || FACE || (GEN && O.face_cons[0] != nullptr);
  const bool do_cell = ALL || CMP || CELL || (GEN && (what & 1));
  const bool do_cons = ALL || CMP || FACE || (GEN && (what & 2));
  const bool do_gf = ALL || FACE || (GEN && (what & 2) && full);
  const bool st_cf = ALL || CELL || (GEN && do_cell && O.cell_fwd), st_ci = ALL || CELL || (GEN && do_cell && O.cell_inv);
  const bool st_cj = CMP || (GEN && O.cjac), st_ca = CMP || (GEN && O.cactive[0]);
  const Ix nbase = (Ix)(ic + 1) + (Ix)(j + 1) * N.s1 + N.s2;  // node (ic, j, plane 0)
  const int64_t cs1 = nw0, cs2 = nw0 * nw1;
  T* const wsm = reinterpret_cast<T*>(cg_dyn_smem) + (XWY == 1 ? 0 : threadIdx.y) * xi_warp_vals<T, MODE>;
  using RingT = RingV<T, 4, 4>;
  T* const sg = wsm + RingT::VALS;
  T* const sg_cf = sg, * const sg_ci = sg + stage_len<T, 10>, * const sg_ff = sg + 2 * stage_len<T, 10>,
         * const sg_fi = sg + 3 * stage_len<T, 10>, * const sg_fc = sg + 4 * stage_len<T, 10>;
  // node ring: ring rows 0..3 = node rows j-1 .. j+2, ring column c (after the phase shift) = node column i0 - 1 + c,
  // fetched in 16-byte pieces (RingV)
  const int ph = (int)(i0 & (RingT::VEC - 1));
  const RingT R{wsm + ph};
  const Ix rg0 = (Ix)(i0 - ph + RingT::VEC * lane) + (Ix)j * N.s1 + N.s2;  // plane 0, row j-1
#pragma unroll
  for (int pp = 0; pp < 3; ++pp) R.fetch(N, (int)(k0 + pp), rg0 + (Ix)(k0 + pp) * N.s2, lane, ph);
  cp_async_commit();

  const T dl0 = clamp_delta<T>(D.w0[0] + i, D.nhalo, D.gn[0]), dl1 = clamp_delta<T>(D.w0[1] + j, D.nhalo, D.gn[1]);

  T psi_bot[6];
  if (do_cons) {
    T wp[6], wm[6];
    const Ix b = nbase + (Ix)(k0 - 1) * N.s2;
    eface6<T, SCH>(zplane(N, b), zplane(N, b + N.s2), zplane(N, b + 2 * N.s2), wp, wm);
#pragma unroll
    for (int e = 0; e < 6; ++e) psi_bot[e] = wp[e] + shfl_dn(wm[e], 1);
  }
  PF<T> z0 = zplane(N, nbase + (Ix)k0 * N.s2), z1 = zplane(N, nbase + (Ix)(k0 + 1) * N.s2);
  for (int64_t m = k0; m < k1; ++m) {
    const Ix b = nbase + (Ix)m * N.s2;
    const int64_t cl0 = i0 + j * cs1 + m * cs2;  // first output cell of the warp in this plane
    const int64_t cl = cl0 + rel;
    // the TMA unit has finished reading last step's staging buffers; the node planes m .. m+2 have landed
    bulk_wait_read();  // lanes other than 0 own no bulk groups: their wait returns at once
    cp_async_wait_all();
    __syncwarp();
    if (m + 1 < k1) {  // plane m+3 (first needed in the next step) into the slot plane m-1 has left
      R.fetch(N, (int)(m + 3), rg0 + (Ix)(m + 3) * N.s2, lane, ph);
      cp_async_commit();
    }
    const int pm = (int)m;
    PF<T> z2 = R.zplane(pm + 2, 1, lane);  // looked-ahead plane; z0, z1 are carried
    Stage<T, 10> Scf(O.cell_fwd, cl0, sg_cf), Sci(O.cell_inv, cl0, sg_ci), Sff(O.face_fwd[0], cl0, sg_ff),
        Sfi(O.face_inv[0], cl0, sg_fi);
    Stage<T, 9> Sfc(O.face_cons[0], cl0, sg_fc);
    // eta-planes j, j+1 of the cell (u = xi, v = zeta) from the zeta-pairs of its four node columns
    const PF<T> e1 = pair_form(R.zpair(pm, 1, lane), R.zpair(pm, 1, lane + 1)),
                e2 = pair_form(R.zpair(pm, 2, lane), R.zpair(pm, 2, lane + 1));
    T c[3][8];
    {
      T dz[3];
#pragma unroll
      for (int q = 0; q < 3; ++q) dz[q] = e1.q[q].cv + e2.q[q].cv;
      taylor_from_z(z0, z1, dz, c);
    }
    if (do_cell || do_gf) {
      T F[3][3], G[3][3];
#pragma unroll
      for (int q = 0; q < 3; ++q)
#pragma unroll
        for (int d = 0; d < 3; ++d) F[q][d] = c[q][1 << d];
      const T detF = inv3(F, G);
      if (do_cell) {
        // cell.forward is evaluated at the clamped point (discrete_grid.jl:113-142): only halo cells differ
        // from the centre value, so the correction is skipped by warps that hold interior cells only
        T Ff[3][3], J = detF;
#pragma unroll
        for (int q = 0; q < 3; ++q)
#pragma unroll
          for (int d = 0; d < 3; ++d) Ff[q][d] = F[q][d];
        const T dl2 = clamp_delta<T>(D.w0[2] + m, D.nhalo, D.gn[2]);
        if (__any_sync(FULLMASK, dl0 != T(0)) || dl1 != T(0) || dl2 != T(0)) {
          const T dl[3] = {dl0, dl1, dl2};
#pragma unroll
          for (int q = 0; q < 3; ++q)
#pragma unroll
            for (int d = 0; d < 3; ++d) {
              const int e = (d + 1) % 3, g = (d + 2) % 3;
              Ff[q][d] = c[q][1 << d] + c[q][(1 << d) | (1 << e)] * dl[e] + c[q][(1 << d) | (1 << g)] * dl[g] +
                         c[q][7] * dl[e] * dl[g];
            }
          J = det3(Ff);
        }
        if (st_cf) put10(Scf.slot(srel), Ff, J);
        if (st_ci) put10(Sci.slot(srel), G, J);
        if (st_cj && out_ok) O.cjac[cl] = J;
      }
      if (do_gf) {
        // ---- reconstructed inverse Jacobian on the xi face ----
        T Fa[3], Fb[3], L[3][3], R[3][3], Gf[3][3], Ff[3][3];
#pragma unroll
        for (int q = 0; q < 3; ++q) {
          Fa[q] = c[q][3];
          Fb[q] = c[q][5];
        }
        ginv_states3_ax<0>(G, Fa, Fb, lam1, lam2, L, R);
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
          for (int cc = 0; cc < 3; ++cc) Gf[r][cc] = T(0.5) * (L[r][cc] + shfl_dn(R[r][cc], 1));
        T Jf = rcp(inv3(Gf, Ff));  // det(F_face) = 1 / det(G_face)
        put10(Sff.slot(srel), Ff, Jf);
        put10(Sfi.slot(srel), Gf, Jf);
      }
    }
    if (do_cons) {
      // ---- zeta-direction edge scalars: top face (m + 1/2), looked ahead one plane ----
      T psi_top[6];
      {
        T wp[6], wm[6];
        eface6<T, SCH>(z0, z1, z2, wp, wm);
#pragma unroll
        for (int e = 0; e < 6; ++e) psi_top[e] = wp[e] + shfl_dn(wm[e], 1);
      }
      // ---- eta-direction edge scalars on both eta faces of the cell ----
      T dpsi_eta[6];
      {
        // eta-planes j-1 .. j+2 (u = xi, v = zeta) from the zeta-pairs of the node columns i, i+1
        const PF<T> e0 = pair_form(R.zpair(pm, 0, lane), R.zpair(pm, 0, lane + 1)),
                    e3 = pair_form(R.zpair(pm, 3, lane), R.zpair(pm, 3, lane + 1));
        T wp[6], wm[6];
        eface6_diff<T, SCH>(e0, e1, e2, e3, wp, wm);  // (face j + 1/2) - (face j - 1/2)
#pragma unroll
        for (int e = 0; e < 6; ++e) dpsi_eta[e] = wp[e] + shfl_dn(wm[e], 1);
      }
      T Gh[3][3];
#pragma unroll
      for (int col = 0; col < 3; ++col) {
        Gh[0][col] = (psi_top[3 + col] - psi_bot[3 + col]) - dpsi_eta[3 + col];
        Gh[1][col] = -(psi_top[col] - psi_bot[col]);
        Gh[2][col] = dpsi_eta[col];
      }
#pragma unroll
      for (int e = 0; e < 6; ++e) psi_bot[e] = psi_top[e];
      if (st_ca && out_ok) {
        T* o = O.cactive[0] + 3 * cl;
        o[0] = Gh[0][0]; o[1] = Gh[0][1]; o[2] = Gh[0][2];
      }
      if (full) {
        // ---- xi line stencils (neighbours by shuffle): rows eta (+, b = zeta) and zeta (-, b = eta) ----
#pragma unroll
        for (int col = 0; col < 3; ++col) {
          const int q1 = (col + 1) % 3, q2 = (col + 2) % 3;
#pragma unroll
          for (int bb = 1; bb <= 2; ++bb) {
            LineS<T> s = line_from_taylor(c, q1, q2, 0, bb, lam1, lam2);
            T v = (s.vp - T(2) * s.u) - shfl_up(s.vp, 1) + shfl_dn(T(2) * s.u - s.vm, 1) + shfl_dn(s.vm, 2);
            if (bb == 2) Gh[1][col] += T(0.5) * v;
            else Gh[2][col] -= T(0.5) * v;
          }
        }
        put9(Sfc.slot(srel), Gh);
      }
    }
    z0 = z1;
    z1 = z2;
    // ---- one flush point: staged arrays -> TMA ----
    fence_async_smem();
    __syncwarp();
    if (st_cf) Scf.issue(nvalid, lane);
    if (st_ci) Sci.issue(nvalid, lane);
    if (do_gf) {
      Sff.issue(nvalid, lane);
      Sfi.issue(nvalid, lane);
    }
    if (do_cons && full) Sfc.issue(nvalid, lane);
    if (lane == 0) bulk_commit();
  }
  if (lane == 0) bulk_wait_read();  // shared memory must outlive the TMA reads
}

// ===========================================================================
// Kernel C: zeta-face metrics.  f = zeta, s = xi, t = eta.  Marches with the
// "front" cell p = m+1; the face (m + 1/2) is written at array plane m.
//   active row zeta : +dPsi^{eta->zeta}[b=xi] across eta - dPsi^{xi->zeta}[b=eta] across xi
//   row xi          : +line_zeta[b=eta]                   - dPsi^{eta->zeta}[b=zeta] across eta
//   row eta         : +dPsi^{xi->zeta}[b=zeta] across xi  - line_zeta[b=xi]
// ===========================================================================
template <class T>
struct CellZ {
  T wpe[6];  // W+_zeta [E^eta(j+1/2)] - W+_zeta [E^eta(j-1/2)]   (difference across eta already taken)
  T wme[6];
  T wpx[6];  // W+_zeta [E^xi(i+1/2)]
  T wmx[6];
  T L[3][3], R[3][3];
  LineS<T> ln[6];  // zeta-line scalars of the cell (index = bb*3 + col; bb 0: b = xi, 1: b = eta)
};

// node sources of the zeta kernel: global memory (prologue) or the warp's ring (marching loop).
// zp(r, c): zeta-pair across planes p, p+1 of the column (row j + r, col i + c); zpl(d): raw zeta-plane p + d
template <class T>
struct GlobalSrc {
  const NodeP<T>& N;
  Ix b;  // node (ic, j, p)
  CG_D ZPair<T> zp(int r, int c) const { return zpair(N, b + (Ix)r * N.s1 + (Ix)c); }
  CG_D PF<T> zpl(int d) const { return zplane(N, b + (Ix)d * N.s2); }
};
template <class T, class RING>
struct RingSrc {
  const RING& R;
  int p, lane;
  CG_D ZPair<T> zp(int r, int c) const { return R.zpair(p, r + 1, lane + c); }
  CG_D PF<T> zpl(int d) const { return R.zplane(p + d, 1, lane); }
};

// the six zeta-line scalars of the cell whose lower plane is the source's plane (index = bb*3 + col; bb 0: b = xi, 1: b = eta)
template <class T, class SRC>
CG_D void zlines(const SRC& S, T lam1, T lam2, LineS<T> (&ln)[6]) {
  T c[3][8];
  {
    const ZPair<T> p00 = S.zp(0, 0), p01 = S.zp(0, 1), p10 = S.zp(1, 0), p11 = S.zp(1, 1);
    T dz[3];
#pragma unroll
    for (int q = 0; q < 3; ++q) dz[q] = (p00.d[q] + p01.d[q]) + (p10.d[q] + p11.d[q]);
    taylor_from_z(S.zpl(0), S.zpl(1), dz, c);
  }
#pragma unroll
  for (int col = 0; col < 3; ++col) {
    const int q1 = (col + 1) % 3, q2 = (col + 2) % 3;
    ln[col] = line_from_taylor(c, q1, q2, 2, 0, lam1, lam2);
    ln[3 + col] = line_from_taylor(c, q1, q2, 2, 1, lam1, lam2);
  }
}

template <class T, int SCH, bool GF, class SRC>
CG_D void cellz_compute(const SRC& S, bool full, T lam1, T lam2, CellZ<T>& o) {
  // All plane forms of the cell at the source's plane p come from the zeta-pairs (sum, difference across
  // planes p, p+1) of the node columns rows j-1 .. j+2 x cols i, i+1 and rows j, j+1 x col i+2.
  const ZPair<T> p00 = S.zp(0, 0), p01 = S.zp(0, 1), p10 = S.zp(1, 0), p11 = S.zp(1, 1);
  const PF<T> E1 = pair_form(p00, p01), E2 = pair_form(p10, p11);  // eta-planes j, j+1 (u = xi, v = zeta)
  {
    // eta faces j-1/2 and j+1/2: eta-planes j-1 .. j+2 -> need (u,v) = (zeta, xi)
    const PF<T> E0 = pair_form(S.zp(-1, 0), S.zp(-1, 1)), E3 = pair_form(S.zp(2, 0), S.zp(2, 1));
    eface6_diff<T, SCH>(swap_uv(E0), swap_uv(E1), swap_uv(E2), swap_uv(E3), o.wpe, o.wme);
  }
  T de[3];  // sum of the cell's four eta-edge differences
  {
    // xi face i+1/2: xi-planes i, i+1, i+2, forms (eta,zeta) -> need (u,v) = (zeta, eta)
    const PF<T> X0 = pair_form(p00, p10), X1 = pair_form(p01, p11), X2 = pair_form(S.zp(0, 2), S.zp(1, 2));
    eface6<T, SCH>(swap_uv(X0), swap_uv(X1), swap_uv(X2), o.wpx, o.wmx);
#pragma unroll
    for (int q = 0; q < 3; ++q) de[q] = X0.q[q].cu + X1.q[q].cu;
  }
  if (!full) return;
  T c[3][8];
  taylor_raw_e(E1, E2, de, c);  // one (raw) Taylor expansion of the cell serves the line scalars and the L / R states
#pragma unroll
  for (int col = 0; col < 3; ++col) {
    const int q1 = (col + 1) % 3, q2 = (col + 2) % 3;
    o.ln[col] = line_from_taylor_raw(c, q1, q2, 2, 0, lam1, lam2);
    o.ln[3 + col] = line_from_taylor_raw(c, q1, q2, 2, 1, lam1, lam2);
  }
  if (GF) {
    // raw F = 4 F, raw dF/d(zeta) = 2 dF/d(zeta): inv gives G / 4, and with (2 lam1, 4 lam2) the states come out as
    // L / 4, R / 4 (G' ~ G F' G, G'' ~ G F' G F' G); the face average folds the 4 back in
    T F[3][3], G[3][3], Fa[3], Fb[3];
#pragma unroll
    for (int q = 0; q < 3; ++q) {
#pragma unroll
      for (int d = 0; d < 3; ++d) F[q][d] = c[q][1 << d];
      Fa[q] = c[q][5];  // column xi of dF/d(zeta)
      Fb[q] = c[q][6];  // column eta
    }
    inv3(F, G);
    ginv_states3_ax<2>(G, Fa, Fb, T(2) * lam1, T(4) * lam2, o.L, o.R);
  }
}

// SWAP = false: zeta faces, marching along k.  SWAP = true: ETA faces with the same code, marching along j:
// the strides of eta and zeta are exchanged, i.e. the kernel works in the left-handed labelling
// (xi, eta' = zeta, zeta' = eta).  In that labelling every Thomas-Lombard entry changes sign and the rows /
// Jacobian columns eta <-> zeta are exchanged; the stores undo both.
template <class T> constexpr int zeta_stage_vals = 2 * stage_len<T, 10> + stage_len<T, 9>;
template <class T, int MODE> constexpr int zeta_warp_vals =
    RingV<T, 4, Z_SLOTS>::VALS + (MODE == MODE_COMPACT ? 0 : (1 + Z_DBUF) * zeta_stage_vals<T>);
template <class T, int SCH, bool SWAP, int MODE>
__global__ void __launch_bounds__(ZNT, (MODE == MODE_COMPACT ? CG_COMPACT_MINB : CG_ZETA_MINB) * 4 / ZWY) k_face_zeta(Dims D, const T* ex, const T* ey, const T* ez,
                                                                MetricOut<T> O, int kchunk) {
  constexpr bool GF = true;
  constexpr int AX = SWAP ? 1 : 2;  // face axis written
  extern __shared__ __align__(128) unsigned char cg_dyn_smem[];
  // state carried from step to step (registers):
  //   wpe, wpx : W+ of the cell below the face (eta-difference / xi-face polynomials)
  //   F0..F2   : line stencil pipeline.  The front cell m+1 contributes V- to the face m-1, (2U - V-) to m,
  //              (V+ - 2U) to m+1 and -V+ to m+2, so the conserved metrics of the face m-1 are completed (and
  //              stored) in step m; F0, F1, -F2 are the partial sums of the faces m-1, m, m+1 at step entry
  //   GhP      : conserved metrics of the face m-1 without the V- term (kernel labelling)
  //   Lpk      : L state of the reconstructed inverse Jacobian of the cell below the face
  T wpe[6], wpx[6], F0[6], F1[6], F2[6], GhP[3][3], Lpk[9];
  T* const wsm = reinterpret_cast<T*>(cg_dyn_smem) + (ZWY == 1 ? 0 : threadIdx.y) * zeta_warp_vals<T, MODE>;
  using RingT = RingV<T, 4, Z_SLOTS>;
  T* const stg = wsm + RingT::VALS;
  const int lane = threadIdx.x;
  const int64_t nw0 = D.nw[0];
  // "row" axis (blockIdx.y) and marching axis; the plane range [D.r0, D.r1) of axis 2 restricts whichever is zeta
  const int64_t row0 = SWAP ? D.r0 : 0, nw1 = SWAP ? D.r1 : D.nw[1];
  const int64_t mar0 = SWAP ? 0 : D.r0, nw2 = SWAP ? D.nw[1] : D.r1;
  const unsigned bx = (ZWY == 1 && JFAST) ? blockIdx.y : blockIdx.x, by = (ZWY == 1 && JFAST) ? blockIdx.x : blockIdx.y;
  const int64_t j = row0 + (ZWY == 1 ? (int64_t)by : (int64_t)by * blockDim.y + threadIdx.y);
  if (j >= nw1) return;
  const int64_t i0 = (int64_t)bx * ZETA_OUT;  // first output cell of the warp (lane 1)
  const int64_t i = i0 + lane - 1;
  const int64_t ic = i > nw0 ? nw0 : i;  // this kernel reads node columns ic .. ic+2
  const bool out_ok = lane >= 1 && lane <= ZETA_OUT && i < nw0;
  const int nvalid = (int)(nw0 - i0 < ZETA_OUT ? nw0 - i0 : ZETA_OUT);
  const int rel = lane - 1;
  // staging slot of this lane: every lane writes (no branch); slot 31 is never sent (31 output cells per warp)
  const int srel = lane == 0 ? 31 : lane - 1;
  const int64_t k0 = mar0 + (int64_t)blockIdx.z * kchunk;
  const int64_t k1 = k0 + kchunk < nw2 ? k0 + kchunk : nw2;
  NodeP<T> N;
  N.q[0] = ex; N.q[1] = ey; N.q[2] = ez;
  N.s1 = (Ix)D.se[SWAP ? 2 : 1]; N.s2 = (Ix)D.se[SWAP ? 1 : 2];
  const T lam1 = lam1_of<T>(SCH), lam2 = lam2_of<T>(SCH);
  constexpr bool ALL = MODE == MODE_FULL, CMP = MODE == MODE_COMPACT;
  const bool full = ALL || (!CMP && O.face_cons[AX] != nullptr);
  const Ix nbase = (Ix)(ic + 1) + (Ix)(j + 1) * N.s1 + N.s2;
  // cell strides of the row axis and the marching axis in the output arrays
  const int64_t cs1 = SWAP ? D.nw[0] * D.nw[1] : D.nw[0], cs2 = SWAP ? D.nw[0] : D.nw[0] * D.nw[1];

  // node ring: ring rows 0..3 = node rows j-1 .. j+2, ring column c (after the phase shift) = node column i0 - 1 + c;
  // planes k0+1 .. k0+Z_SLOTS-1 are primed here, plane m+Z_SLOTS follows at the top of step m into the slot the plane m
  // has left (no barrier inside a step)
  const int ph = (int)(i0 & (RingT::VEC - 1));
  const RingT R{wsm + ph};
  const Ix rg0 = (Ix)(i0 - ph + RingT::VEC * lane) + (Ix)j * N.s1 + N.s2;
#pragma unroll
  for (int pp = 1; pp < Z_SLOTS; ++pp) R.fetch(N, (int)(k0 + pp), rg0 + (Ix)(k0 + pp) * N.s2, lane, ph);
  cp_async_commit();
  // carried state of the first step
  {
    const GlobalSrc<T> S0{N, nbase + (Ix)k0 * N.s2};
    if (full) {
      LineS<T> lm[6];  // cell k0 - 1
      zlines(GlobalSrc<T>{N, nbase + (Ix)(k0 - 1) * N.s2}, lam1, lam2, lm);
#pragma unroll
      for (int e = 0; e < 6; ++e) F2[e] = lm[e].vp;
    }
    CellZ<T> c0;  // cell k0
    cellz_compute<T, SCH, GF>(S0, full, lam1, lam2, c0);
#pragma unroll
    for (int e = 0; e < 6; ++e) {
      wpe[e] = c0.wpe[e];
      wpx[e] = c0.wpx[e];
    }
    if (full) {
#pragma unroll
      for (int e = 0; e < 6; ++e) {  // the face k0-1 belongs to the previous chunk: F0 is not needed
        F0[e] = T(0);
        F1[e] = (c0.ln[e].vp - T(2) * c0.ln[e].u) - F2[e];
        F2[e] = c0.ln[e].vp;
      }
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int cc = 0; cc < 3; ++cc) {
          GhP[r][cc] = T(0);
          if (GF) Lpk[3 * r + cc] = c0.L[r][cc];
        }
    }
  }
  // completes the conserved metrics of the face below (GhP + V- term), stages and hands them to the TMA unit
  auto finish_cons = [&](const T (&vm)[6], int64_t cl0_prev, T* sg_fc) {
    T Go[3][3];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int cc = 0; cc < 3; ++cc) Go[r][cc] = GhP[r][cc];
#pragma unroll
    for (int e = 0; e < 6; ++e) {
      const T v = T(0.5) * (F0[e] + vm[e]);
      if (e >= 3) Go[0][e - 3] += v;  // row xi: + line[b = eta]
      else Go[1][e] -= v;             // row eta: - line[b = xi]
    }
    if (SWAP) {
#pragma unroll
      for (int col = 0; col < 3; ++col) {
        T r0 = -Go[0][col], r1 = -Go[2][col], r2 = -Go[1][col];
        Go[0][col] = r0; Go[1][col] = r1; Go[2][col] = r2;
      }
    }
    Stage<T, 9> Sc(O.face_cons[AX], cl0_prev, sg_fc);
    put9(Sc.slot(srel), Go);
    return Sc;
  };
#pragma unroll(Z_UNROLL)
  for (int64_t m = k0; m < k1; ++m) {
    const int64_t cl0 = i0 + j * cs1 + m * cs2;  // first output cell of the warp in this step
    const int64_t cl = cl0 + rel;
    // the TMA unit has finished reading the staging buffers this step writes (lanes other than 0 own no bulk groups:
    // their wait returns at once); the node planes m+1, m+2 have landed
    if (Z_DBUF) bulk_wait_read1();
    else bulk_wait_read();
    cp_async_wait_all();
    __syncwarp();
    if (m + 1 < k1) {  // the next plane the ring lacks, into the slot the plane m has left
      R.fetch(N, (int)(m + Z_SLOTS), rg0 + (Ix)(m + Z_SLOTS) * N.s2, lane, ph);
      cp_async_commit();
    }
    T* const sgm = stg + (Z_DBUF ? (int)((m - k0) & 1) * zeta_stage_vals<T> : 0);
    T* const sg_ff = sgm, * const sg_fi = sgm + stage_len<T, 10>, * const sg_fc = sgm + 2 * stage_len<T, 10>;
    const RingSrc<T, RingT> S{R, (int)(m + 1), lane};
    CellZ<T> cf;  // front cell m + 1
    cellz_compute<T, SCH, GF>(S, full, lam1, lam2, cf);
    T Gh[3][3];  // rows 0 xi, 1 eta, 2 zeta
#pragma unroll
    for (int col = 0; col < 3; ++col) {
      // a = eta (= t, sigma +1, a' = xi): bsel 1 is b = xi, bsel 0 is b = zeta; difference across eta is in wpe/wme
      T de_a = wpe[3 + col] + cf.wme[3 + col];
      T de_f = wpe[col] + cf.wme[col];
      // a = xi (= s, sigma -1, a' = eta): bsel 1 is b = eta, bsel 0 is b = zeta
      T px_a = wpx[3 + col] + cf.wmx[3 + col];
      T px_f = wpx[col] + cf.wmx[col];
      T dx_a = px_a - shfl_up(px_a, 1);
      T dx_f = px_f - shfl_up(px_f, 1);
      Gh[2][col] = de_a - dx_a;
      Gh[0][col] = -de_f;
      Gh[1][col] = dx_f;
    }
#pragma unroll
    for (int e = 0; e < 6; ++e) {
      wpe[e] = cf.wpe[e];
      wpx[e] = cf.wpx[e];
    }
    if ((CMP || (!ALL && O.cactive[AX])) && out_ok) {
      T* o = O.cactive[AX] + 3 * cl;
      const T sgn = SWAP ? T(-1) : T(1);
      o[0] = sgn * Gh[2][0]; o[1] = sgn * Gh[2][1]; o[2] = sgn * Gh[2][2];
    }
    if (!full) continue;
    Stage<T, 10> Sf(O.face_fwd[AX], cl0, sg_ff), Si(O.face_inv[AX], cl0, sg_fi);
    // zeta line stencil: the front cell's V- completes the face m-1 (stored now, one step late)
    T vm[6];
#pragma unroll
    for (int e = 0; e < 6; ++e) vm[e] = cf.ln[e].vm;
    const Stage<T, 9> Sc = finish_cons(vm, cl0 - cs2, sg_fc);
#pragma unroll
    for (int e = 0; e < 6; ++e) {
      F0[e] = F1[e] + (T(2) * cf.ln[e].u - cf.ln[e].vm);
      F1[e] = (cf.ln[e].vp - T(2) * cf.ln[e].u) - F2[e];
      F2[e] = cf.ln[e].vp;
    }
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int cc = 0; cc < 3; ++cc) GhP[r][cc] = Gh[r][cc];
    if (GF) {
      T Gf[3][3], Ff[3][3];
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int cc = 0; cc < 3; ++cc) {
          Gf[r][cc] = T(2) * (Lpk[3 * r + cc] + cf.R[r][cc]);  // states carry 1/4 (cellz_compute)
          Lpk[3 * r + cc] = cf.L[r][cc];
        }
      T Jf = rcp(inv3(Gf, Ff));  // det(F_face) = 1 / det(G_face)
      if (SWAP) {  // back to the true labelling: columns eta <-> zeta of F, rows eta <-> zeta of G (det changes sign)
        Jf = -Jf;
#pragma unroll
        for (int q = 0; q < 3; ++q) {
          T t1 = Ff[q][1]; Ff[q][1] = Ff[q][2]; Ff[q][2] = t1;
          T t2 = Gf[1][q]; Gf[1][q] = Gf[2][q]; Gf[2][q] = t2;
        }
      }
      put10(Sf.slot(srel), Ff, Jf);
      put10(Si.slot(srel), Gf, Jf);
    }
    // ---- one flush point: staged arrays -> TMA ----
    fence_async_smem();
    __syncwarp();
    if (GF) {
      Sf.issue(nvalid, lane);
      Si.issue(nvalid, lane);
    }
    if (m > k0) Sc.issue(nvalid, lane);
    if (lane == 0) bulk_commit();
  }
  if (full) {
    // the last face of the chunk needs V- of the cell k1 + 1
    if (lane == 0) bulk_wait_read();
    __syncwarp();
    const GlobalSrc<T> S1{N, nbase + (Ix)(k1 + 1) * N.s2};
    LineS<T> l2[6];
    zlines(S1, lam1, lam2, l2);
    T vm[6];
#pragma unroll
    for (int e = 0; e < 6; ++e) vm[e] = l2[e].vm;
    const Stage<T, 9> Sc = finish_cons(vm, i0 + j * cs1 + (k1 - 1) * cs2, stg + 2 * stage_len<T, 10>);
    fence_async_smem();
    __syncwarp();
    Sc.issue(nvalid, lane);
    if (lane == 0) bulk_commit();
  }
  if (lane == 0) bulk_wait_read();  // shared memory must outlive the TMA reads
}

// ===========================================================================
// Kernel XZ (round 2): cell metrics + xi-face metrics + zeta-face metrics in ONE k-marching kernel (FULL profile).
//
// The xi kernel and the zeta kernel both march along k over the same node rows and both rebuild, per cell, the
// eta- and xi-plane forms, the Taylor expansion and inv(F).  Fused, that work is done once (about 1300 FP64
// instructions per warp step against 766 + 718), and — more important — the combined kernel writes 624 B per
// cell for that arithmetic, which moves it from the FP64 issue limit to the HBM limit.  The eta faces stay in
// k_face_zeta<SWAP> (marching along j).
//
// Step m of a chunk [k0, k1) works on the FRONT cell p = m + 1 (its node planes p, p+1 and the looked-ahead p+2
// are in the ring) and writes
//     the cell arrays and the xi-face arrays of cell p            (skipped in the last step: p = k1 is the next chunk's)
//     the zeta-face forward / inverse metrics of the face m+1/2   (skipped in the first step m = k0 - 1)
//     the conserved metrics of the zeta face (m-1)+1/2            (completed by the front cell's V-, one step late)
// so a chunk runs k1 - k0 + 1 steps, m = k0-1 .. k1-1; every chunk is alike (no special first / last chunk).
// Outputs leave in two staging phases per step through the same buffers (cell + xi, then zeta).
// ===========================================================================
// Shared memory per warp: node ring (3 slots) | staging (two Metric buffers + one ConservedMetric buffer, used in
// three phases per step: cell arrays, xi-face arrays, zeta-face arrays) | parking of carried state.  The kernel
// carries 54 values per lane from step to step; with the 72 live values of an eface6_diff that is more than 255
// registers hold (the first version spilled 53 values per step into a 28 KB L1: 26.8 ms).  The 36 values that are
// touched once per step (partial conserved metrics of the face below, L state of the cell below, the line-stencil
// pipeline) are therefore parked in per-lane shared-memory slots; the ring keeps 3 slots (the plane of the NEXT
// step is fetched once the last read of the front cell's lower plane is done) so that 8 one-warp blocks still fit
// an SM (27 KB each).  Measured at 512^3: 21.2 ms (three TMA round trips per step; parking in an L2-resident
// global scratch with six staging buffers instead: 24.5 ms) against 19.1 ms for k_face_xi + k_face_zeta.
template <class T> constexpr int xz_stage_vals = 2 * stage_len<T, 10> + stage_len<T, 9>;
constexpr int XZ_PARK = 36;  // GhP 0..8 | Lpk 9..17 | F0 18..23 | F1 24..29 | F2 30..35
template <class T> constexpr int xz_warp_vals = RingV<T, 4, 3>::VALS + xz_stage_vals<T> + XZ_PARK * 32;
// parked values: per-lane shared-memory slots ([value][lane]: conflict-free, no synchronisation, a lane only touches
// its own)
template <class T> CG_D T park_ld(const T* p) { return *p; }
template <class T> CG_D void park_st(T* p, T v) { *p = v; }
#ifndef CG_XZ_MINB
#define CG_XZ_MINB 8
#endif
template <class T, int SCH>
__global__ void __launch_bounds__(32, CG_XZ_MINB) k_face_xz(Dims D, const T* ex, const T* ey, const T* ez, MetricOut<T> O,
                                                            int kchunk) {
  extern __shared__ __align__(128) unsigned char cg_dyn_smem[];
  // carried from step to step in registers:
  //   psi_bot : Psi^{zeta->xi} on the zeta face below the front cell (xi-face conserved metrics)
  //   wpe, wpx : W+ of the cell below the zeta face (as in k_face_zeta)
  // parked in shared memory: GhP (conserved metrics of the face below without the V- term), Lpk (L state below),
  // F0, F1, F2 (line-stencil pipeline of k_face_zeta)
  T psi_bot[6], wpe[6], wpx[6];
  T* const wsm = reinterpret_cast<T*>(cg_dyn_smem);
  using RingT = RingV<T, 4, 3>;
  T* const stg = wsm + RingT::VALS;
  T* const sg_a = stg, * const sg_b = stg + stage_len<T, 10>, * const sg_9 = stg + 2 * stage_len<T, 10>;
  T* const sg_c = sg_a, * const sg_d = sg_b, * const sg_9b = sg_9;  // one set of buffers, three phases
  const int lane = threadIdx.x;
  T* const park = stg + xz_stage_vals<T> + lane;  // value v of this lane: park[32 * v]
  const int64_t nw0 = D.nw[0], nw1 = D.nw[1];
  const int64_t j = (int64_t)blockIdx.y;
  if (j >= nw1) return;
  const int64_t i0 = (int64_t)blockIdx.x * WARP_OUT;  // first output cell of the warp (lane 1)
  const int64_t i = i0 + lane - 1;
  const int64_t ic = i > nw0 + 1 ? nw0 + 1 : i;
  const int nvalid = (int)(nw0 - i0 < WARP_OUT ? nw0 - i0 : WARP_OUT);
  const int srel = lane == 0 ? 31 : lane - 1;  // staging slot: every lane writes; slots 29 .. 31 are never sent
  const int64_t k0 = D.r0 + (int64_t)blockIdx.z * kchunk;
  const int64_t k1 = k0 + kchunk < D.r1 ? k0 + kchunk : D.r1;
  NodeP<T> N;
  N.q[0] = ex; N.q[1] = ey; N.q[2] = ez;
  N.s1 = (Ix)D.se[1]; N.s2 = (Ix)D.se[2];
  const T lam1 = lam1_of<T>(SCH), lam2 = lam2_of<T>(SCH);
  const Ix nbase = (Ix)(ic + 1) + (Ix)(j + 1) * N.s1 + N.s2;  // node (ic, j, plane 0)
  const int64_t cs1 = nw0, cs2 = nw0 * nw1;
  const int ph = (int)(i0 & (RingT::VEC - 1));
  const RingT R{wsm + ph};
  const Ix rg0 = (Ix)(i0 - ph + RingT::VEC * lane) + (Ix)j * N.s1 + N.s2;  // plane 0, row j-1
  // the first step (front cell k0) needs the planes k0, k0+1, k0+2
#pragma unroll
  for (int pp = 0; pp < 3; ++pp) R.fetch(N, (int)(k0 + pp), rg0 + (Ix)(k0 + pp) * N.s2, lane, ph);
  cp_async_commit();
  const T dl0 = clamp_delta<T>(D.w0[0] + i, D.nhalo, D.gn[0]), dl1 = clamp_delta<T>(D.w0[1] + j, D.nhalo, D.gn[1]);
  // zeta coordinate of the front cell's centre (clamp_delta in floating point: no 64-bit integers in the loop)
  T xi_z = T(D.w0[2] + k0 + 1 - D.nhalo) + T(0.5);
  const T hi_z = T(D.gn[2] + 1);
  // carried state at the entry of the first step (m = k0 - 1, front cell k0): V+ of the cell k0 - 1 and the edge
  // scalars on the zeta face below the cell k0; the zeta-face state of "cell k0 - 1" is not used (its face belongs
  // to the previous chunk)
  {
    LineS<T> lm[6];
    zlines(GlobalSrc<T>{N, nbase + (Ix)(k0 - 1) * N.s2}, lam1, lam2, lm);
#pragma unroll
    for (int v = 0; v < 30; ++v) park_st(park + 32 * v, T(0));
#pragma unroll
    for (int e = 0; e < 6; ++e) {
      park_st(park + 32 * (30 + e), lm[e].vp);  // F2
      wpe[e] = wpx[e] = T(0);
    }
    T wp[6], wm[6];
    const Ix b = nbase + (Ix)(k0 - 1) * N.s2;
    eface6<T, SCH>(zplane(N, b), zplane(N, b + N.s2), zplane(N, b + 2 * N.s2), wp, wm);
#pragma unroll
    for (int e = 0; e < 6; ++e) psi_bot[e] = wp[e] + shfl_dn(wm[e], 1);
  }
  for (int64_t m = k0 - 1; m < k1; ++m) {
    const int64_t p = m + 1;              // front cell
    const bool do_x = p < k1;             // cell + xi-face outputs of the cell p
    const bool do_z = m >= k0;            // zeta face m + 1/2 (and the conserved metrics of the face below it)
    const int64_t cl0p = i0 + j * cs1 + p * cs2;  // first output cell of the warp in the plane p
    const int64_t cl0m = cl0p - cs2;
    // the node planes p .. p+2 have landed (the wait for last step's staging buffers comes right before the first
    // staging store of this step: by then the TMA unit has long finished reading them)
    cp_async_wait_all();
    __syncwarp();
    const int pp = (int)p;
    T Gh_x[3][3], Gh_z[3][3];  // conserved metrics of the xi face of the cell p / of the zeta face m (rows 0 xi, 1 eta, 2 zeta)
    // ---- (1) zeta faces p-1/2 (carried) and p+1/2 of the cell, for the xi face: zeta-planes p, p+1, p+2 ---------
    if (do_x) {
      T wp[6], wm[6];
      eface6<T, SCH>(R.zplane(pp, 1, lane), R.zplane(pp + 1, 1, lane), R.zplane(pp + 2, 1, lane), wp, wm);
#pragma unroll
      for (int e = 0; e < 6; ++e) {
        const T top = wp[e] + shfl_dn(wm[e], 1);
        const T d = top - psi_bot[e];
        psi_bot[e] = top;
        if (e >= 3) Gh_x[0][e - 3] = d;    // active row: + (psi_top - psi_bot)[b = eta]
        else Gh_x[1][e] = -d;              // row eta:    - (psi_top - psi_bot)[b = xi]
      }
    }
    // ---- (2) eta faces j-1/2 and j+1/2 of the cell: eta-planes j-1 .. j+2, once with (u, v) = (xi, zeta) for the xi
    //          face and once with (zeta, xi) for the zeta face; the Taylor sums of the cell from the inner two --------
    T c[3][8];  // RAW Taylor sums of the front cell
    {
      const ZPair<T> p00 = R.zpair(pp, 1, lane), p01 = R.zpair(pp, 1, lane + 1), p10 = R.zpair(pp, 2, lane),
                     p11 = R.zpair(pp, 2, lane + 1);
      const PF<T> E1 = pair_form(p00, p01), E2 = pair_form(p10, p11);  // eta-planes j, j+1 (u = xi, v = zeta)
      T de[3];  // sum of the four eta-edge differences: 2 (lo) + (d) of the two column pairs (round-off rule)
#pragma unroll
      for (int q = 0; q < 3; ++q)
        de[q] = fma(T(2), (p10.lo[q] - p00.lo[q]) + (p11.lo[q] - p01.lo[q]), (p10.d[q] - p00.d[q]) + (p11.d[q] - p01.d[q]));
      const PF<T> E0 = pair_form(R.zpair(pp, 0, lane), R.zpair(pp, 0, lane + 1)),
                  E3 = pair_form(R.zpair(pp, 3, lane), R.zpair(pp, 3, lane + 1));
      T wp[6], wm[6];
      if (do_x) {
        eface6_diff<T, SCH>(E0, E1, E2, E3, wp, wm);  // (face j + 1/2) - (face j - 1/2), f = xi
#pragma unroll
        for (int col = 0; col < 3; ++col) {
          Gh_x[0][col] -= wp[3 + col] + shfl_dn(wm[3 + col], 1);
          Gh_x[2][col] = wp[col] + shfl_dn(wm[col], 1);
        }
      }
      eface6_diff<T, SCH>(swap_uv(E0), swap_uv(E1), swap_uv(E2), swap_uv(E3), wp, wm);  // f = zeta
#pragma unroll
      for (int col = 0; col < 3; ++col) {
        // a = eta (sigma +1, a' = xi): bsel 1 is b = xi, bsel 0 is b = zeta
        Gh_z[2][col] = wpe[3 + col] + wm[3 + col];
        Gh_z[0][col] = -(wpe[col] + wm[col]);
      }
#pragma unroll
      for (int e = 0; e < 6; ++e) wpe[e] = wp[e];
      taylor_raw_e(E1, E2, de, c);  // last use of the inner eta-planes
    }
    // ---- (3) xi face i+1/2 for the zeta face: xi-planes i, i+1, i+2 with (u, v) = (zeta, eta) --------------------
    {
      const ZPair<T> a0 = R.zpair(pp, 1, lane), a1 = R.zpair(pp, 1, lane + 1), a2 = R.zpair(pp, 1, lane + 2),
                     b0 = R.zpair(pp, 2, lane), b1 = R.zpair(pp, 2, lane + 1), b2 = R.zpair(pp, 2, lane + 2);
      T wp[6], wm[6];
      eface6<T, SCH>(swap_uv(pair_form(a0, b0)), swap_uv(pair_form(a1, b1)), swap_uv(pair_form(a2, b2)), wp, wm);
#pragma unroll
      for (int col = 0; col < 3; ++col) {
        // a = xi (sigma -1, a' = eta): bsel 1 is b = eta, bsel 0 is b = zeta
        const T px_a = wpx[3 + col] + wm[3 + col];
        const T px_f = wpx[col] + wm[col];
        Gh_z[2][col] -= px_a - shfl_up(px_a, 1);
        Gh_z[1][col] = px_f - shfl_up(px_f, 1);
      }
#pragma unroll
      for (int e = 0; e < 6; ++e) wpx[e] = wp[e];
    }
    // the front cell's lower plane has been read for the last time: the plane the NEXT step lacks goes into its slot
    __syncwarp();
    if (m + 1 < k1) {
      R.fetch(N, (int)(p + 3), rg0 + (Ix)(p + 3) * N.s2, lane, ph);
      cp_async_commit();
    }
    // ---- (4) inverse Jacobian of the cell: raw F = 4 F, inv gives G / 4 -------------------------------------------
    T G[3][3], detF;
    {
      T F[3][3];
#pragma unroll
      for (int q = 0; q < 3; ++q)
#pragma unroll
        for (int d = 0; d < 3; ++d) F[q][d] = c[q][1 << d];
      detF = T(1.0 / 64.0) * inv3(F, G);
    }
    if (do_x) {
      // ================= phase A1: cell arrays of the cell p =======================================================
      {
        Stage<T, 10> Scf(O.cell_fwd, cl0p, sg_a), Sci(O.cell_inv, cl0p, sg_b);
        // cell.forward is evaluated at the clamped point (discrete_grid.jl:113-142): only halo cells differ from the
        // centre value, so the correction is skipped by warps that hold interior cells only
        T Ff[3][3], Gt[3][3], J = detF;
#pragma unroll
        for (int q = 0; q < 3; ++q)
#pragma unroll
          for (int d = 0; d < 3; ++d) {
            Ff[q][d] = T(0.25) * c[q][1 << d];
            Gt[q][d] = T(4) * G[q][d];
          }
        const T dl2 = (xi_z < T(1) ? T(1) : (xi_z > hi_z ? hi_z : xi_z)) - xi_z;
        if (__any_sync(FULLMASK, dl0 != T(0)) || dl1 != T(0) || dl2 != T(0)) {
          const T dl[3] = {dl0, dl1, dl2};
#pragma unroll
          for (int q = 0; q < 3; ++q)
#pragma unroll
            for (int d = 0; d < 3; ++d) {
              const int e = (d + 1) % 3, g = (d + 2) % 3;
              Ff[q][d] = T(0.25) * c[q][1 << d] + T(0.5) * (c[q][(1 << d) | (1 << e)] * dl[e] + c[q][(1 << d) | (1 << g)] * dl[g]) +
                         c[q][7] * dl[e] * dl[g];
            }
          J = det3(Ff);
        }
        bulk_wait_read();  // last step's groups have left the staging buffers (issued most of a step ago)
        __syncwarp();
        put10(Scf.slot(srel), Ff, J);
        put10(Sci.slot(srel), Gt, J);
        fence_async_smem();
        __syncwarp();
        Scf.issue(nvalid, lane);
        Sci.issue(nvalid, lane);
        if (lane == 0) bulk_commit();
      }
      // ================= phase A2: xi-face arrays of the cell p ====================================================
      T Gf[3][3], Ff[3][3], Jf;
      {
        // reconstructed inverse Jacobian on the xi face (states carry 1/4, see cellz_compute)
        T Fa[3], Fb[3], L[3][3], Rr[3][3];
#pragma unroll
        for (int q = 0; q < 3; ++q) {
          Fa[q] = c[q][3];  // column eta of dF/d(xi)
          Fb[q] = c[q][5];  // column zeta
        }
        ginv_states3_ax<0>(G, Fa, Fb, T(2) * lam1, T(4) * lam2, L, Rr);
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
          for (int cc = 0; cc < 3; ++cc) Gf[r][cc] = T(2) * (L[r][cc] + shfl_dn(Rr[r][cc], 1));
        Jf = rcp(inv3(Gf, Ff));  // det(F_face) = 1 / det(G_face)
      }
      // xi line stencils (neighbours by shuffle): rows eta (+, b = zeta) and zeta (-, b = eta)
#pragma unroll
      for (int col = 0; col < 3; ++col) {
        const int q1 = (col + 1) % 3, q2 = (col + 2) % 3;
#pragma unroll
        for (int bb = 1; bb <= 2; ++bb) {
          const LineS<T> sc = line_from_taylor_raw(c, q1, q2, 0, bb, lam1, lam2);
          const T v = (sc.vp - T(2) * sc.u) - shfl_up(sc.vp, 1) + shfl_dn(T(2) * sc.u - sc.vm, 1) + shfl_dn(sc.vm, 2);
          if (bb == 2) Gh_x[1][col] += T(0.5) * v;
          else Gh_x[2][col] -= T(0.5) * v;
        }
      }
      bulk_wait_read();  // the cell arrays have left the two Metric buffers
      __syncwarp();
      Stage<T, 10> Sff(O.face_fwd[0], cl0p, sg_c), Sfi(O.face_inv[0], cl0p, sg_d);
      Stage<T, 9> Sfc(O.face_cons[0], cl0p, sg_9);
      put10(Sff.slot(srel), Ff, Jf);
      put10(Sfi.slot(srel), Gf, Jf);
      put9(Sfc.slot(srel), Gh_x);
      fence_async_smem();
      __syncwarp();
      Sff.issue(nvalid, lane);
      Sfi.issue(nvalid, lane);
      Sfc.issue(nvalid, lane);
      if (lane == 0) bulk_commit();
    }
    // ================= phase B: zeta face m + 1/2 ================================================================
    // zeta-line scalars of the front cell and its zeta-axis L / R states
    LineS<T> ln[6];
#pragma unroll
    for (int col = 0; col < 3; ++col) {
      const int q1 = (col + 1) % 3, q2 = (col + 2) % 3;
      ln[col] = line_from_taylor_raw(c, q1, q2, 2, 0, lam1, lam2);
      ln[3 + col] = line_from_taylor_raw(c, q1, q2, 2, 1, lam1, lam2);
    }
    T Lz[3][3], Rz[3][3];
    {
      T Fa[3], Fb[3];
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        Fa[q] = c[q][5];  // column xi of dF/d(zeta)
        Fb[q] = c[q][6];  // column eta
      }
      ginv_states3_ax<2>(G, Fa, Fb, T(2) * lam1, T(4) * lam2, Lz, Rz);
    }
    if (do_z) {
      T Go[3][3], Gf[3][3], Ff[3][3];
      // the front cell's V- completes the conserved metrics of the face m - 1 (stored one step late)
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int cc = 0; cc < 3; ++cc) Go[r][cc] = park_ld(park + 32 * (3 * r + cc));
#pragma unroll
      for (int e = 0; e < 6; ++e) {
        const T v = T(0.5) * (park_ld(park + 32 * (18 + e)) + ln[e].vm);  // F0
        if (e >= 3) Go[0][e - 3] += v;  // row xi: + line[b = eta]
        else Go[1][e] -= v;             // row eta: - line[b = xi]
      }
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int cc = 0; cc < 3; ++cc) Gf[r][cc] = T(2) * (park_ld(park + 32 * (9 + 3 * r + cc)) + Rz[r][cc]);
      const T Jf = rcp(inv3(Gf, Ff));
      bulk_wait_read();  // the xi-face arrays (or, in the last step, the previous step's arrays) have left the buffers
      __syncwarp();
      Stage<T, 10> Sf(O.face_fwd[2], cl0m, sg_a), Si(O.face_inv[2], cl0m, sg_b);
      Stage<T, 9> Sc(O.face_cons[2], cl0m - cs2, sg_9b);
      put9(Sc.slot(srel), Go);
      put10(Sf.slot(srel), Ff, Jf);
      put10(Si.slot(srel), Gf, Jf);
      fence_async_smem();
      __syncwarp();
      Sf.issue(nvalid, lane);
      Si.issue(nvalid, lane);
      if (m > k0) Sc.issue(nvalid, lane);
      if (lane == 0) bulk_commit();
    }
    // ---- carried state of the next step ----
    xi_z += T(1);
#pragma unroll
    for (int e = 0; e < 6; ++e) {
      const T f1 = park_ld(park + 32 * (24 + e)), f2 = park_ld(park + 32 * (30 + e));
      park_st(park + 32 * (18 + e), f1 + (T(2) * ln[e].u - ln[e].vm));        // F0
      park_st(park + 32 * (24 + e), (ln[e].vp - T(2) * ln[e].u) - f2);        // F1
      park_st(park + 32 * (30 + e), ln[e].vp);                                // F2
    }
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int cc = 0; cc < 3; ++cc) {
        park_st(park + 32 * (3 * r + cc), Gh_z[r][cc]);
        park_st(park + 32 * (9 + 3 * r + cc), Lz[r][cc]);
      }
  }
  {
    // the last zeta face of the chunk needs V- of the cell k1 + 1
    if (lane == 0) bulk_wait_read();
    __syncwarp();
    const GlobalSrc<T> S1{N, nbase + (Ix)(k1 + 1) * N.s2};
    LineS<T> l2[6];
    zlines(S1, lam1, lam2, l2);
    T Go[3][3];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int cc = 0; cc < 3; ++cc) Go[r][cc] = park_ld(park + 32 * (3 * r + cc));
#pragma unroll
    for (int e = 0; e < 6; ++e) {
      const T v = T(0.5) * (park_ld(park + 32 * (18 + e)) + l2[e].vm);
      if (e >= 3) Go[0][e - 3] += v;
      else Go[1][e] -= v;
    }
    Stage<T, 9> Sc(O.face_cons[2], i0 + j * cs1 + (k1 - 1) * cs2, sg_9b);
    put9(Sc.slot(srel), Go);
    fence_async_smem();
    __syncwarp();
    Sc.issue(nvalid, lane);
    if (lane == 0) bulk_commit();
  }
  if (lane == 0) bulk_wait_read();  // shared memory must outlive the TMA reads
}

// Chunk length along the marching axis for `tiles` warp tiles of `per_block` warps each, `slots` resident blocks on the
// GPU: among 1 .. extent / min_chunk chunks per tile, the count with the best (wave quantisation) x (prologue
// amortisation) efficiency; `prologue` = cost of a chunk's set-up in steps.
inline int pick_chunk(int64_t tiles, int64_t extent, int64_t slots, int min_chunk, double prologue) {
  int best_n = 1;
  double best = -1.0;
  const int64_t max_n = extent / min_chunk > 0 ? extent / min_chunk : 1;
  for (int64_t n = 1; n <= max_n && n <= 64; ++n) {
    const double waves = (double)(tiles * n) / (double)slots;
    const double full = (double)(int64_t)waves + ((double)(int64_t)waves < waves ? 1.0 : 0.0);
    const double len = (double)((extent + n - 1) / n);
    // quantisation of the last wave x amortisation of the chunk prologue x half a wave of ragged tail
    const double eff = waves / full * len / (len + prologue) * waves / (waves + 0.5);
    if (eff > best + 1e-9) {
      best = eff;
      best_n = (int)n;
    }
  }
  int c = (int)((extent + best_n - 1) / best_n);
  return c < 1 ? 1 : c;
}

// ---------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------
template <class T>
inline bool tiled3d_supported(const Dims& D, const MetricOut<T>&) {
  // gridDim.y carries rows (j, or k planes in the eta kernel), gridDim.z the chunks: both are limited to 65535; a
  // window that skewed falls back to the gather kernel
  const bool grid_ok = D.nw[1] <= 65535 && D.nw[2] <= 65535;
  return D.dim == 3 && grid_ok && (sizeof(Ix) == 8 || D.ne[0] * D.ne[1] * D.ne[2] < (int64_t(1) << 31));
}
// ===========================================================================
// 2-D: one thread per PAIR of xi-adjacent cells (2p, 2p+1) of a row.  The pair's 80 B of every
// `Metric` array and 64 B of every `ConservedMetric` array are contiguous and 16-B aligned, so all
// stores are 16-B vector stores, and the 4x3 node patch is loaded once for both cells.
// Replaces metrics.jl:1822-1848 (cell) and :1906-1960 (faces) for DiscreteGrid.
// ===========================================================================
template <class T>
CG_D void face2d(const Cell2<T>& c0, const Cell2<T>& c1, const T (&G0)[2][2], const T (&G1)[2][2], int f, T lam1,
                 T lam2, bool want_g, T (&Gh)[2][2], T (&Gf)[2][2], T (&Ff)[2][2], T& J) {
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
#pragma unroll
  for (int r = 0; r < 2; ++r)
#pragma unroll
    for (int cc = 0; cc < 2; ++cc)
      Gh[r][cc] = T(0.5) * ((H0[r][cc] + lam1 * Hp0[r][cc]) + (H1[r][cc] - lam1 * Hp1[r][cc]));
  if (!want_g) return;
  const int o_ax = 1 - f;
  T Fp0[2][2] = {{0, 0}, {0, 0}}, Fp1[2][2] = {{0, 0}, {0, 0}};
  Fp0[0][o_ax] = c0.tw[0]; Fp0[1][o_ax] = c0.tw[1];
  Fp1[0][o_ax] = c1.tw[0]; Fp1[1][o_ax] = c1.tw[1];
  T L0[2][2], R0[2][2], L1[2][2], R1[2][2];
  ginv_states2(G0, Fp0, lam1, lam2, L0, R0);
  ginv_states2(G1, Fp1, lam1, lam2, L1, R1);
#pragma unroll
  for (int r = 0; r < 2; ++r)
#pragma unroll
    for (int cc = 0; cc < 2; ++cc) Gf[r][cc] = T(0.5) * (L0[r][cc] + R1[r][cc]);
  inv2(Gf, Ff);
  J = Ff[0][0] * Ff[1][1] - Ff[0][1] * Ff[1][0];
}

// store the Metric elements of the pair (10 values, 80 B, 16-B aligned for FP64)
template <class T>
CG_D void store_pair5(T* o, const T (&A)[2][2], T JA, const T (&B)[2][2], T JB, bool second) {
  if (sizeof(T) == 8 && second) {
    double2* p = reinterpret_cast<double2*>(o);
    p[0] = make_double2(A[0][0], A[1][0]);
    p[1] = make_double2(A[0][1], A[1][1]);
    p[2] = make_double2(JA, B[0][0]);
    p[3] = make_double2(B[1][0], B[0][1]);
    p[4] = make_double2(B[1][1], JB);
  } else {
    o[0] = A[0][0]; o[1] = A[1][0]; o[2] = A[0][1]; o[3] = A[1][1]; o[4] = JA;
    if (second) { o[5] = B[0][0]; o[6] = B[1][0]; o[7] = B[0][1]; o[8] = B[1][1]; o[9] = JB; }
  }
}
template <class T>
CG_D void store_pair4(T* o, const T (&A)[2][2], const T (&B)[2][2], bool second) {
  if (sizeof(T) == 8) {
    double2* p = reinterpret_cast<double2*>(o);
    p[0] = make_double2(A[0][0], A[1][0]);
    p[1] = make_double2(A[0][1], A[1][1]);
    if (second) {
      p[2] = make_double2(B[0][0], B[1][0]);
      p[3] = make_double2(B[0][1], B[1][1]);
    }
  } else {
    o[0] = A[0][0]; o[1] = A[1][0]; o[2] = A[0][1]; o[3] = A[1][1];
    if (second) { o[4] = B[0][0]; o[5] = B[1][0]; o[6] = B[0][1]; o[7] = B[1][1]; }
  }
}

template <class T>
__global__ void __launch_bounds__(128) k_metrics2d_pair(Dims D, const T* ex, const T* ey, MetricOut<T> O, int scheme,
                                                        int what) {
  const T lam1 = lam1_of<T>(scheme), lam2 = lam2_of<T>(scheme);
  const int64_t nw0 = D.nw[0], nw1 = D.nw[1];
  const int64_t npair = (nw0 + 1) / 2;
  const int64_t s1 = D.se[1];
  const int64_t ntot = npair * nw1;
  for (int64_t lin = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; lin < ntot;
       lin += (int64_t)gridDim.x * blockDim.x) {
    const int64_t p = lin % npair, j = lin / npair;
    const int64_t i = 2 * p;
    const bool second = i + 1 < nw0;
    const int64_t b = (i + 1) + (j + 1) * s1;
    const int64_t cl = i + j * nw0;
    Cell2<T> c00 = cell2_load(ex, ey, b, s1), c10 = cell2_load(ex, ey, b + 1, s1);
    T G00[2][2], G10[2][2];
    inv2(c00.F, G00);
    inv2(c10.F, G10);
    if (what & 1) {
      T dl1 = clamp_delta<T>(D.w0[1] + j, D.nhalo, D.gn[1]);
      T dlA = clamp_delta<T>(D.w0[0] + i, D.nhalo, D.gn[0]), dlB = clamp_delta<T>(D.w0[0] + i + 1, D.nhalo, D.gn[0]);
      T FA[2][2], FB[2][2];
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        FA[q][0] = c00.F[q][0] + c00.tw[q] * dl1;
        FA[q][1] = c00.F[q][1] + c00.tw[q] * dlA;
        FB[q][0] = c10.F[q][0] + c10.tw[q] * dl1;
        FB[q][1] = c10.F[q][1] + c10.tw[q] * dlB;
      }
      T JA = FA[0][0] * FA[1][1] - FA[0][1] * FA[1][0], JB = FB[0][0] * FB[1][1] - FB[0][1] * FB[1][0];
      if (O.cell_fwd) store_pair5(O.cell_fwd + 5 * cl, FA, JA, FB, JB, second);
      if (O.cell_inv) store_pair5(O.cell_inv + 5 * cl, G00, JA, G10, JB, second);
      if (O.cjac) {
        O.cjac[cl] = JA;
        if (second) O.cjac[cl + 1] = JB;
      }
    }
    if (!(what & 2)) continue;
    const bool want_g = O.face_fwd[0] != nullptr;
    T GhA[2][2], GhB[2][2], GfA[2][2], GfB[2][2], FfA[2][2], FfB[2][2], JA = 0, JB = 0;
    {  // xi faces: (c00,c10) and (c10,c20)
      Cell2<T> c20 = cell2_load(ex, ey, b + 2, s1);
      T G20[2][2];
      inv2(c20.F, G20);
      face2d(c00, c10, G00, G10, 0, lam1, lam2, want_g, GhA, GfA, FfA, JA);
      face2d(c10, c20, G10, G20, 0, lam1, lam2, want_g, GhB, GfB, FfB, JB);
      if (O.face_cons[0]) store_pair4(O.face_cons[0] + 4 * cl, GhA, GhB, second);
      if (O.cactive[0]) {
        T* o = O.cactive[0] + 2 * cl;
        o[0] = GhA[0][0]; o[1] = GhA[0][1];
        if (second) { o[2] = GhB[0][0]; o[3] = GhB[0][1]; }
      }
      if (want_g) {
        store_pair5(O.face_fwd[0] + 5 * cl, FfA, JA, FfB, JB, second);
        store_pair5(O.face_inv[0] + 5 * cl, GfA, JA, GfB, JB, second);
      }
    }
    {  // eta faces: (c00,c01) and (c10,c11)
      Cell2<T> c01 = cell2_load(ex, ey, b + s1, s1), c11 = cell2_load(ex, ey, b + s1 + 1, s1);
      T G01[2][2], G11[2][2];
      inv2(c01.F, G01);
      inv2(c11.F, G11);
      face2d(c00, c01, G00, G01, 1, lam1, lam2, want_g, GhA, GfA, FfA, JA);
      face2d(c10, c11, G10, G11, 1, lam1, lam2, want_g, GhB, GfB, FfB, JB);
      if (O.face_cons[1]) store_pair4(O.face_cons[1] + 4 * cl, GhA, GhB, second);
      if (O.cactive[1]) {
        T* o = O.cactive[1] + 2 * cl;
        o[0] = GhA[1][0]; o[1] = GhA[1][1];
        if (second) { o[2] = GhB[1][0]; o[3] = GhB[1][1]; }
      }
      if (want_g) {
        store_pair5(O.face_fwd[1] + 5 * cl, FfA, JA, FfB, JB, second);
        store_pair5(O.face_inv[1] + 5 * cl, GfA, JA, GfB, JB, second);
      }
    }
  }
}

// ===========================================================================
// 2-D, FULL profile, cell + face: a warp owns 32 consecutive cells of a row and walks along eta; the
// eight AoS arrays (40-B Metric / 32-B ConservedMetric elements) go through the warp's staging buffers
// and the TMA unit like the 3-D kernels ("Output staging" above).  Same arithmetic as k_metrics2d_pair.
// ===========================================================================
template <class T>
CG_D void put5(T* o, const T (&M)[2][2], T J) {
  o[0] = M[0][0]; o[1] = M[1][0]; o[2] = M[0][1]; o[3] = M[1][1]; o[4] = J;
}
template <class T>
CG_D void put4(T* o, const T (&M)[2][2]) {
  if constexpr (sizeof(T) == 8) {
    double2* p = reinterpret_cast<double2*>(o);
    p[0] = make_double2(M[0][0], M[1][0]);
    p[1] = make_double2(M[0][1], M[1][1]);
  } else {
    o[0] = M[0][0]; o[1] = M[1][0]; o[2] = M[0][1]; o[3] = M[1][1];
  }
}
template <class T> constexpr int stage2d_vals = 6 * stage_len<T, 5> + 2 * stage_len<T, 4>;
#ifndef CG_W2Y
#define CG_W2Y 4
#endif
// rows (warps) per block: the node rows of neighbouring warps are shared through L1 (4: 1.02 ms at 4096^2, 1: 1.13 ms)
constexpr int W2Y = CG_W2Y;
template <class T>
__global__ void __launch_bounds__(32 * W2Y) k_metrics2d_staged(Dims D, const T* __restrict__ ex, const T* __restrict__ ey,
                                                         MetricOut<T> O, int scheme, int jchunk) {
  extern __shared__ __align__(128) unsigned char cg_dyn_smem[];
  const T lam1 = lam1_of<T>(scheme), lam2 = lam2_of<T>(scheme);
  const int lane = threadIdx.x;
  const int64_t nw0 = D.nw[0], nw1 = D.nw[1], s1 = D.se[1];
  const int64_t i0 = (int64_t)blockIdx.x * 32;
  const int nvalid = (int)(nw0 - i0 < 32 ? nw0 - i0 : 32);
  const bool out_ok = lane < nvalid;
  const int64_t i = out_ok ? i0 + lane : nw0 - 1;
  const int64_t j0 = (W2Y == 1 ? (int64_t)blockIdx.y : (int64_t)blockIdx.y * blockDim.y + threadIdx.y) * jchunk;
  const int64_t j1 = j0 + jchunk < nw1 ? j0 + jchunk : nw1;
  T* const sg = reinterpret_cast<T*>(cg_dyn_smem) + (W2Y == 1 ? 0 : threadIdx.y) * stage2d_vals<T>;
  T* sb[8];
#pragma unroll
  for (int a = 0; a < 6; ++a) sb[a] = sg + a * stage_len<T, 5>;
  sb[6] = sg + 6 * stage_len<T, 5>;
  sb[7] = sb[6] + stage_len<T, 4>;
  const T dl0 = clamp_delta<T>(D.w0[0] + i, D.nhalo, D.gn[0]);
  for (int64_t j = j0; j < j1; ++j) {
    const int64_t b = (i + 1) + (j + 1) * s1;
    const int64_t cl0 = i0 + j * nw0;
    bulk_wait_read();  // the TMA unit has finished reading last step's staging buffers (lanes > 0 own no groups)
    __syncwarp();
    Stage<T, 5> Scf(O.cell_fwd, cl0, sb[0]), Sci(O.cell_inv, cl0, sb[1]), Sf0(O.face_fwd[0], cl0, sb[2]),
        Si0(O.face_inv[0], cl0, sb[3]), Sf1(O.face_fwd[1], cl0, sb[4]), Si1(O.face_inv[1], cl0, sb[5]);
    Stage<T, 4> Sc0(O.face_cons[0], cl0, sb[6]), Sc1(O.face_cons[1], cl0, sb[7]);
    const Cell2<T> c00 = cell2_load(ex, ey, b, s1), c10 = cell2_load(ex, ey, b + 1, s1), c01 = cell2_load(ex, ey, b + s1, s1);
    T G00[2][2], G10[2][2], G01[2][2];
    inv2(c00.F, G00);
    inv2(c10.F, G10);
    inv2(c01.F, G01);
    {
      const T dl1 = clamp_delta<T>(D.w0[1] + j, D.nhalo, D.gn[1]);
      T FA[2][2];
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        FA[q][0] = c00.F[q][0] + c00.tw[q] * dl1;
        FA[q][1] = c00.F[q][1] + c00.tw[q] * dl0;
      }
      const T JA = FA[0][0] * FA[1][1] - FA[0][1] * FA[1][0];
      put5(Scf.slot(lane), FA, JA);  // every lane stages (no branch); slots >= nvalid are never sent
      put5(Sci.slot(lane), G00, JA);
    }
    {
      T Gh[2][2], Gf[2][2], Ff[2][2], J = 0;
      face2d(c00, c10, G00, G10, 0, lam1, lam2, true, Gh, Gf, Ff, J);
      put4(Sc0.slot(lane), Gh);
      put5(Sf0.slot(lane), Ff, J);
      put5(Si0.slot(lane), Gf, J);
      face2d(c00, c01, G00, G01, 1, lam1, lam2, true, Gh, Gf, Ff, J);
      put4(Sc1.slot(lane), Gh);
      put5(Sf1.slot(lane), Ff, J);
      put5(Si1.slot(lane), Gf, J);
    }
    fence_async_smem();
    __syncwarp();
    Scf.issue(nvalid, lane);
    Sci.issue(nvalid, lane);
    Sf0.issue(nvalid, lane);
    Si0.issue(nvalid, lane);
    Sf1.issue(nvalid, lane);
    Si1.issue(nvalid, lane);
    Sc0.issue(nvalid, lane);
    Sc1.issue(nvalid, lane);
    if (lane == 0) bulk_commit();
  }
  if (lane == 0) bulk_wait_read();  // shared memory must outlive the TMA reads
}

// the staged kernel covers the FULL cell+face configuration at any width
template <class T>
inline bool staged2d_applies(const MetricOut<T>& O, int what) {
  static const bool use_staged = !(getenv("CG_2D_STAGED") && atoi(getenv("CG_2D_STAGED")) == 0);
  return use_staged && (what & 3) == 3 && O.cell_fwd && O.cell_inv && O.face_fwd[0] && O.face_inv[0] && O.face_cons[0] &&
         O.face_fwd[1] && O.face_inv[1] && O.face_cons[1] && !O.cjac && !O.cactive[0] && !O.cactive[1];
}
template <class T>
inline bool tiled2d_supported(const Dims& D, const MetricOut<T>& O, int what) {
  // pair stores need every row to start 16-B aligned: an even number of cells per row
  return D.dim == 2 && (staged2d_applies(O, what) || D.nw[0] % 2 == 0 || sizeof(T) == 4);
}
template <class T>
inline int launch_metrics2d_tiled(const Dims& D, const T* ex, const T* ey, const MetricOut<T>& O, int scheme, int what,
                                  cudaStream_t st) {
  if (staged2d_applies(O, what)) {
    const unsigned gx = (unsigned)((D.nw[0] + 31) / 32);
    // rows per warp: enough warps to fill the machine a few times over, long enough to amortise the set-up
    int jchunk = (int)((D.nw[1] * (int64_t)gx + 148LL * 64 - 1) / (148LL * 64));
    if (jchunk < 4) jchunk = 4;
    const int64_t nwarps_y = (D.nw[1] + jchunk - 1) / jchunk;
    const int smem = W2Y * stage2d_vals<T> * (int)sizeof(T);
    cudaFuncSetAttribute(k_metrics2d_staged<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    k_metrics2d_staged<T><<<dim3(gx, (unsigned)((nwarps_y + W2Y - 1) / W2Y)), dim3(32, W2Y), smem, st>>>(D, ex, ey, O, scheme, jchunk);
    return 1;
  }
  const int64_t n = ((D.nw[0] + 1) / 2) * D.nw[1];
  int64_t blocks = (n + 127) / 128;
  if (blocks > 148LL * 64) blocks = 148LL * 64;
  k_metrics2d_pair<T><<<(unsigned)blocks, 128, 0, st>>>(D, ex, ey, O, scheme, what);
  return 1;
}

// The three face kernels write disjoint arrays.  On small grids they are launched on `st` and two helper streams
// (fork / join with events) so that the last, partly filled wave of one kernel overlaps the first wave of the next.
// CG_CONCURRENT=0 serialises them always.
struct Fork {
  cudaStream_t s[2] = {nullptr, nullptr};
  cudaEvent_t start = nullptr, done[2] = {nullptr, nullptr};
  bool made = false;
  // The helper streams and events belong to ONE grid handle (a handle is single-threaded by contract, so two host
  // threads working on two handles of the same device never share an event).
  bool ensure() {
    static const int enabled = [] {
      const char* e = getenv("CG_CONCURRENT");  // debug switch (listed in include/curvigrid_cuda.h)
      return e ? atoi(e) : 1;
    }();
    if (!enabled) return false;
    if (made) return true;
    if (cudaStreamCreateWithFlags(&s[0], cudaStreamNonBlocking) != cudaSuccess) return false;
    if (cudaStreamCreateWithFlags(&s[1], cudaStreamNonBlocking) != cudaSuccess) return false;
    if (cudaEventCreateWithFlags(&start, cudaEventDisableTiming) != cudaSuccess) return false;
    if (cudaEventCreateWithFlags(&done[0], cudaEventDisableTiming) != cudaSuccess) return false;
    if (cudaEventCreateWithFlags(&done[1], cudaEventDisableTiming) != cudaSuccess) return false;
    made = true;
    return true;
  }
  void destroy() {
    for (int i = 0; i < 2; ++i) {
      if (s[i]) cudaStreamDestroy(s[i]);
      if (done[i]) cudaEventDestroy(done[i]);
      s[i] = nullptr;
      done[i] = nullptr;
    }
    if (start) cudaEventDestroy(start);
    start = nullptr;
    made = false;
  }
};

// mask: bit0 xi kernel (+cell), bit1 eta kernel, bit2 zeta kernel
template <class T, int SCH>
inline int launch3d_scheme(const Dims& D, const T* ex, const T* ey, const T* ez, const MetricOut<T>& O, int what,
                           int mask, cudaStream_t st, Fork* fork) {
  // enough blocks for ~30 waves (fine-grained tail, measured best at 512^3) as long as a chunk stays >= 64 steps
  // (the prologue of a chunk costs about three steps)
  const int64_t want = 148LL * 64;
  const int64_t nplanes = D.r1 - D.r0;  // planes of axis 2 to evaluate (the whole window unless cg_grid_eval_planes)
  if (nplanes <= 0) return 0;
  // chunk length along the marching axis: long enough to amortise the prologue
  auto chunk_of = [&](int64_t tiles, int64_t extent) {
    int c = (int)extent;
    if (tiles < want) {
      int64_t nchunks = (want + tiles - 1) / tiles;
      c = (int)((extent + nchunks - 1) / nchunks);
      if (c < 64) c = 64;
      if (c > extent) c = (int)extent;
    }
    return c;
  };
  int n = 0;
  const bool face = what & 2;
  // opt in to > 48 KB of shared memory (per device, cheap: set on every launch)
  // which compile-time configuration the outputs correspond to
  const bool cellface = (what & 3) == 3;
  const bool is_full = cellface && O.cell_fwd && O.cell_inv && !O.cjac && !O.cactive[0] && !O.cactive[1] && !O.cactive[2] &&
                       O.face_fwd[0] && O.face_inv[0] && O.face_cons[0] && O.face_fwd[1] && O.face_inv[1] && O.face_cons[1] &&
                       O.face_fwd[2] && O.face_inv[2] && O.face_cons[2];
  const bool is_compact = cellface && O.cjac && O.cactive[0] && O.cactive[1] && O.cactive[2] && !O.cell_fwd && !O.cell_inv &&
                          !O.face_cons[0] && !O.face_cons[1] && !O.face_cons[2];
  const int mode = is_full ? MODE_FULL : (is_compact ? MODE_COMPACT : MODE_GENERIC);
  // FULL-profile refreshes of one cache only (the reference's refresh_cell_metrics! / refresh_face_metrics!)
  const bool no_compact = !O.cjac && !O.cactive[0] && !O.cactive[1] && !O.cactive[2];
  const bool faces_full = no_compact && O.face_fwd[0] && O.face_inv[0] && O.face_cons[0] && O.face_fwd[1] && O.face_inv[1] &&
                          O.face_cons[1] && O.face_fwd[2] && O.face_inv[2] && O.face_cons[2];
  const bool is_face_only = (what & 3) == 2 && faces_full;
  const bool is_cell_only = (what & 3) == 1 && no_compact && O.cell_fwd && O.cell_inv;
  const int zmode = (is_full || is_face_only) ? MODE_FULL : mode;  // eta / zeta kernels
  bool attr_ok = true;
  auto launch_xi = [&](auto kern, int warp_vals, dim3 grid, int kchunk) {
    const int smem = XWY * warp_vals * (int)sizeof(T);
    attr_ok &= cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) == cudaSuccess;
    if (XWY == 1 && JFAST) grid = dim3(grid.y, grid.x, grid.z);
    kern<<<grid, dim3(32, XWY), smem, st>>>(D, ex, ey, ez, O, what, kchunk);
  };
  cudaStream_t zst = st;  // stream of the next eta / zeta launch
  auto launch_zeta = [&](auto kern, int warp_vals, dim3 grid, int chunk) {
    const int smem = ZWY * warp_vals * (int)sizeof(T);
    attr_ok &= cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) == cudaSuccess;
    if (ZWY == 1 && JFAST) grid = dim3(grid.y, grid.x, grid.z);
    kern<<<grid, dim3(32, ZWY), smem, zst>>>(D, ex, ey, ez, O, chunk);
  };
  // small grids only: a large grid fills the GPU for many waves per kernel and the one-warp blocks of the eta / zeta
  // kernels leave short tails; serial launches measured equal or better there (512^3: 28.6 vs 29.2 ms)
  const bool small_grid = D.nw[0] * D.nw[1] * nplanes < (int64_t)(1 << 24);
  Fork* fk = (face && (mask & 7) == 7 && small_grid && fork && fork->ensure()) ? fork : nullptr;
  if (fk) {
    cudaEventRecord(fk->start, st);
    cudaStreamWaitEvent(fk->s[0], fk->start, 0);
    cudaStreamWaitEvent(fk->s[1], fk->start, 0);
  }
  // FULL cell+face evaluation with the xi and zeta kernels fused (k_face_xz): an experiment, OFF by default.  It is
  // correct (the whole GPU suite passes on it) and executes 6 % fewer instructions, but it needs three staging phases
  // per step in the shared memory that is left and measures 21.2 ms at 512^3 against 10.7 + 8.4 ms for the two
  // kernels it replaces.  CG_FUSE_XZ=1 selects it.
  static const bool fuse_env = getenv("CG_FUSE_XZ") && atoi(getenv("CG_FUSE_XZ")) == 1;
  const bool fuse_xz = fuse_env && mode == MODE_FULL && (mask & 5) == 5;
  if (fuse_xz) {
    const unsigned gx = (unsigned)((D.nw[0] + WARP_OUT - 1) / WARP_OUT), gy = (unsigned)D.nw[1];
    const int kchunk = chunk_of((int64_t)gx * gy, nplanes);
    const dim3 grid(gx, gy, (unsigned)((nplanes + kchunk - 1) / kchunk));
    const int smem = xz_warp_vals<T> * (int)sizeof(T);
    attr_ok &= cudaFuncSetAttribute(k_face_xz<T, SCH>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) == cudaSuccess;
    k_face_xz<T, SCH><<<grid, 32, smem, st>>>(D, ex, ey, ez, O, kchunk);
    ++n;
  }
  if (!fuse_xz && (mask & 1) && ((what & 1) || face)) {
    const unsigned gx = (unsigned)((D.nw[0] + WARP_OUT - 1) / WARP_OUT), gy = (unsigned)((D.nw[1] + XWY - 1) / XWY);
    const int64_t xslots = 148LL * (mode == MODE_COMPACT ? CG_XI_COMPACT_MINB : CG_MINB) * 4 / XWY;
    const int kchunk = CG_PICK_CHUNK ? pick_chunk((int64_t)gx * gy, nplanes, xslots, 64, 1.5) : chunk_of((int64_t)gx * gy, nplanes);
    const dim3 grid(gx, gy, (unsigned)((nplanes + kchunk - 1) / kchunk));
    if (mode == MODE_FULL) launch_xi(k_face_xi<T, SCH, MODE_FULL>, xi_warp_vals<T, MODE_FULL>, grid, kchunk);
    else if (mode == MODE_COMPACT) launch_xi(k_face_xi<T, SCH, MODE_COMPACT>, xi_warp_vals<T, MODE_COMPACT>, grid, kchunk);
    else if (is_face_only) launch_xi(k_face_xi<T, SCH, MODE_FACE>, xi_warp_vals<T, MODE_FACE>, grid, kchunk);
    else if (is_cell_only) launch_xi(k_face_xi<T, SCH, MODE_CELL>, xi_warp_vals<T, MODE_CELL>, grid, kchunk);
    else launch_xi(k_face_xi<T, SCH, MODE_GENERIC>, xi_warp_vals<T, MODE_GENERIC>, grid, kchunk);
    ++n;
  }
  const unsigned gxz = (unsigned)((D.nw[0] + ZETA_OUT - 1) / ZETA_OUT);
  const int64_t zslots = 148LL * (zmode == MODE_COMPACT ? CG_COMPACT_MINB : CG_ZETA_MINB) * 4 / ZWY;
  if ((mask & 2) && face) {
    if (fk) zst = fk->s[0];
    // eta faces with the zeta kernel's code, marching along j: rows = k planes
    const unsigned gy = (unsigned)((nplanes + ZWY - 1) / ZWY);
    const int jchunk = CG_PICK_CHUNK ? pick_chunk((int64_t)gxz * gy, D.nw[1], zslots, 64, 3.0) : chunk_of((int64_t)gxz * gy, D.nw[1]);
    const dim3 grid(gxz, gy, (unsigned)((D.nw[1] + jchunk - 1) / jchunk));
    if (zmode == MODE_FULL) launch_zeta(k_face_zeta<T, SCH, true, MODE_FULL>, zeta_warp_vals<T, MODE_FULL>, grid, jchunk);
    else if (zmode == MODE_COMPACT) launch_zeta(k_face_zeta<T, SCH, true, MODE_COMPACT>, zeta_warp_vals<T, MODE_COMPACT>, grid, jchunk);
    else launch_zeta(k_face_zeta<T, SCH, true, MODE_GENERIC>, zeta_warp_vals<T, MODE_GENERIC>, grid, jchunk);
    ++n;
  }
  if (!fuse_xz && (mask & 4) && face) {
    if (fk) zst = fk->s[1];
    const unsigned gy = (unsigned)((D.nw[1] + ZWY - 1) / ZWY);
    const int kchunk = CG_PICK_CHUNK ? pick_chunk((int64_t)gxz * gy, nplanes, zslots, 64, 3.0) : chunk_of((int64_t)gxz * gy, nplanes);
    const dim3 grid(gxz, gy, (unsigned)((nplanes + kchunk - 1) / kchunk));
    if (zmode == MODE_FULL) launch_zeta(k_face_zeta<T, SCH, false, MODE_FULL>, zeta_warp_vals<T, MODE_FULL>, grid, kchunk);
    else if (zmode == MODE_COMPACT) launch_zeta(k_face_zeta<T, SCH, false, MODE_COMPACT>, zeta_warp_vals<T, MODE_COMPACT>, grid, kchunk);
    else launch_zeta(k_face_zeta<T, SCH, false, MODE_GENERIC>, zeta_warp_vals<T, MODE_GENERIC>, grid, kchunk);
    ++n;
  }
  if (fk) {
    cudaEventRecord(fk->done[0], fk->s[0]);
    cudaEventRecord(fk->done[1], fk->s[1]);
    cudaStreamWaitEvent(st, fk->done[0], 0);
    cudaStreamWaitEvent(st, fk->done[1], 0);
  }
  return attr_ok ? n : -1;
}

template <class T>
inline int launch_metrics3d_tiled(const Dims& D, const T* ex, const T* ey, const T* ez, const MetricOut<T>& O,
                                  int scheme, int what, int mask, cudaStream_t st, Fork* fork) {
  if (scheme == ENDPOINT) return launch3d_scheme<T, ENDPOINT>(D, ex, ey, ez, O, what, mask, st, fork);
  if (scheme == GRADIENT) return launch3d_scheme<T, GRADIENT>(D, ex, ey, ez, O, what, mask, st, fork);
  return launch3d_scheme<T, CURVATURE>(D, ex, ey, ez, O, what, mask, st, fork);
}

}  // namespace cg

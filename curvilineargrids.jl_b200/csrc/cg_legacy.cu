// Legacy MEG6 metric pipeline on the device: CurvilinearGrid3D(x, y, z, :meg6 | :meg6_symmetric)
// `update!(mesh, force, include_halo_region)`.
//
// Replaces (paths under /root/reference/src):
//   grids/curvilinear_mapped_grids/3d.jl:472-598            update!, _centroid_coordinates_kernel!
//   metric_schemes/MetricDiscretizationSchemes.jl:20-89     update_metrics!, update_edge_metrics!
//   metric_schemes/forward_metrics.jl:89-167                forward_metrics!, jacobian_3d_kernel!
//   metric_schemes/conservative_metrics.jl:13-175, 259-427  (symmetric_)conservative_metric(s)!
//   discretization_schemes/DiscretizationSchemes.jl:90-159  compute_{first,second}_derivatives!
//   discretization_schemes/derivative_kernels.jl:84-314     the KernelAbstractions GPU kernels
//   discretization_schemes/derivative_stencils.jl:1-193     stencils (Kahan, tol_diff, eps clips)
//   discretization_schemes/meg/*.jl                         MEG derivative + edge interpolation
//
// Design: the reference runs ~350 full-array KA launches per update, each with separate
// inner / lo-edge / hi-edge kernels.  Here every pass is ONE launch that selects the central or
// one-sided stencil from the cell's position in the domain, products/differences/divisions are
// folded into the neighbouring pass, and all temporaries live in four scratch arrays.  The
// thresholds (tol_diff, eps clips) are discontinuous, so this translation unit is compiled with
// -fmad=false and keeps the reference's operation order (Kahan sums included).
#include <cuda.h>
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <cfloat>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include <nvtx3/nvToolsExt.h>

#include "../../include/curvigrid_cuda.h"
#include "cg_legacy_ranges.h"

#ifndef CG_LEGACY_MARCH_DEFAULT
#define CG_LEGACY_MARCH_DEFAULT 1
#endif

namespace {

struct LBox {
  int lo[3], hi[3];
};
struct LShape {
  int n[3];
  long long st[3];
};

__device__ __forceinline__ double kahan3(double a, double b, double c) {
  double s = 0.0, comp = 0.0, y, t;
  y = a - comp; t = s + y; comp = (t - s) - y; s = t;
  y = b - comp; t = s + y; comp = (t - s) - y; s = t;
  y = c - comp; t = s + y; comp = (t - s) - y; s = t;
  return s + comp;
}
__device__ __forceinline__ double kahan6(const double* v) {
  double s = 0.0, comp = 0.0;
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    double y = v[i] - comp;
    double t = s + y;
    comp = (t - s) - y;
    s = t;
  }
  return s + comp;
}
__device__ __forceinline__ bool isapprox(double a, double b, double rtol, double atol) {
  if (a == b) return true;
  if (!isfinite(a) || !isfinite(b)) return false;
  return fabs(a - b) <= fmax(atol, rtol * fmax(fabs(a), fabs(b)));
}
// tol_diff(a, b) = isapprox(a, b; rtol = sqrt(eps), atol = 10 eps) ? (a - b) * 0 : a - b   (derivative_stencils.jl).
// Branch-free form with the same value for every input: isapprox is "a == b, or both finite and
// |a - b| <= max(atol, rtol max(|a|, |b|))".  max(|a|, |b|) is taken on the bit patterns (for non-negative doubles the
// unsigned integer order is the numeric order; sm_100a has no FP64 min/max instruction, fmax expands into ~10), a NaN
// or an infinity shows as an all-ones exponent of that maximum; a == b finite gives d = 0 <= atol, a == b infinite
// gives d = NaN on either path.
__device__ __forceinline__ double tol_diff(double a, double b) {
  const double d = a - b;
  const unsigned ahi = (unsigned)__double2hiint(a) & 0x7fffffffu, bhi = (unsigned)__double2hiint(b) & 0x7fffffffu;
  const unsigned alo = (unsigned)__double2loint(a), blo = (unsigned)__double2loint(b);
  const bool agt = ahi > bhi || (ahi == bhi && alo > blo);
  const unsigned mhi = agt ? ahi : bhi, mlo = agt ? alo : blo;  // bits of max(|a|, |b|)
  const bool fin = mhi < 0x7ff00000u;
  const double thr = 1.4901161193847656e-08 /* sqrt(eps) */ * __hiloint2double((int)mhi, (int)mlo);
  unsigned near;  // |d| <= thr || |d| <= atol as two predicated compares (the compiler would build max(thr, atol))
  asm("{\n\t.reg .pred p;\n\t.reg .f64 ad;\n\tabs.f64 ad, %1;\n\tsetp.le.f64 p, ad, %2;\n\t"
      "setp.le.or.f64 p, ad, %3, p;\n\tselp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(near)
      : "d"(d), "d"(thr), "d"(10 * DBL_EPSILON));
  return (fin && near) ? d * 0.0 : d;
}
__device__ __forceinline__ double clip(double v, double e) { return fabs(v) >= e ? v : v * 0.0; }
__device__ __forceinline__ double central6(double m3, double m2, double m1, double p1, double p2, double p3) {
  double d = kahan3((3.0 / 4.0) * tol_diff(p1, m1), (-3.0 / 20.0) * tol_diff(p2, m2), (1.0 / 60.0) * tol_diff(p3, m3));
  return clip(d, DBL_EPSILON);
}
// component `which` (0,1,2) of mixed_derivatives_loedge on the window p[0..5]
__device__ __forceinline__ double mixed_lo(const double* p, int which) {
  double v[6];
  if (which == 0) {
    v[0] = (-137.0 / 60.0) * p[0]; v[1] = 5 * p[1]; v[2] = -5 * p[2]; v[3] = (10.0 / 3.0) * p[3];
    v[4] = -(5.0 / 4.0) * p[4]; v[5] = (1.0 / 5.0) * p[5];
  } else if (which == 1) {
    v[0] = (-1.0 / 5.0) * p[0]; v[1] = -(13.0 / 12.0) * p[1]; v[2] = 2 * p[2]; v[3] = -p[3];
    v[4] = (1.0 / 3.0) * p[4]; v[5] = -(1.0 / 20.0) * p[5];
  } else {
    v[0] = (1.0 / 20.0) * p[0]; v[1] = -(1.0 / 2.0) * p[1]; v[2] = -(1.0 / 3.0) * p[2]; v[3] = p[3];
    v[4] = -(1.0 / 4.0) * p[4]; v[5] = (1.0 / 30.0) * p[5];
  }
  return clip(kahan6(v), 50 * DBL_EPSILON);
}
// derivative at p[3 + which] (which = 0,1,2) of mixed_derivatives_hiedge on the window p[0..5]
__device__ __forceinline__ double mixed_hi(const double* p, int which) {
  double v[6];
  if (which == 2) {
    v[0] = (-1.0 / 5.0) * p[0]; v[1] = (5.0 / 4.0) * p[1]; v[2] = -(10.0 / 3.0) * p[2]; v[3] = 5 * p[3];
    v[4] = -5 * p[4]; v[5] = (137.0 / 60.0) * p[5];
  } else if (which == 1) {
    v[0] = (1.0 / 20.0) * p[0]; v[1] = -(1.0 / 3.0) * p[1]; v[2] = p[2]; v[3] = -2 * p[3];
    v[4] = (13.0 / 12.0) * p[4]; v[5] = (1.0 / 5.0) * p[5];
  } else {
    v[0] = (-1.0 / 30.0) * p[0]; v[1] = (1.0 / 4.0) * p[1]; v[2] = -p[2]; v[3] = (1.0 / 3.0) * p[3];
    v[4] = (1.0 / 2.0) * p[4]; v[5] = -(1.0 / 20.0) * p[5];
  }
  return clip(kahan6(v), 50 * DBL_EPSILON);
}
__device__ __forceinline__ double phiL(double f, double d, double dd) { return f + ((1.0 / 2.0) * d + (1.0 / 12.0) * dd); }
__device__ __forceinline__ double phiR(double f, double d, double dd) { return f - ((1.0 / 2.0) * d + (1.0 / 12.0) * dd); }

// iterate a box: linear thread id -> (i,j,k), I = linear array index, pos/len along `axis`
#define LEGACY_FOR_BOX(B)                                                                                   \
  const long long e0_ = (B).hi[0] - (B).lo[0], e1_ = (B).hi[1] - (B).lo[1], e2_ = (B).hi[2] - (B).lo[2];    \
  const long long ntot_ = e0_ * e1_ * e2_;                                                                  \
  for (long long lin_ = blockIdx.x * (long long)blockDim.x + threadIdx.x; lin_ < ntot_;                     \
       lin_ += (long long)gridDim.x * blockDim.x)

#define LEGACY_INDEX(B, S)                                                        \
  int c_[3];                                                                      \
  c_[0] = (B).lo[0] + (int)(lin_ % e0_);                                          \
  c_[1] = (B).lo[1] + (int)((lin_ / e0_) % e1_);                                  \
  c_[2] = (B).lo[2] + (int)(lin_ / (e0_ * e1_));                                  \
  const long long I = c_[0] * (S).st[0] + c_[1] * (S).st[1] + c_[2] * (S).st[2];

// first derivative along `axis` over `dom` (inner central-6, first/last three cells one-sided);
// MODE 0: src = phi.  MODE 1: src = a*b (product folded in).  MODE 2: src = a*b - c*d (symmetric scheme).
template <int MODE>
__device__ __forceinline__ double src_val(const double* a, const double* b, const double* c, const double* d, long long I) {
  if (MODE == 0) return a[I];
  if (MODE == 1) return a[I] * b[I];
  return a[I] * b[I] - c[I] * d[I];
}

// materialise the (possibly product) source field over the WHOLE array (conservative_metrics.jl:39-41)
template <int MODE>
__global__ void k_legacy_source(LShape S, const double* a, const double* b, const double* c, const double* d, double* out) {
  const long long n = (long long)S.n[0] * S.n[1] * S.n[2];
  for (long long I = blockIdx.x * (long long)blockDim.x + threadIdx.x; I < n; I += (long long)gridDim.x * blockDim.x)
    out[I] = src_val<MODE>(a, b, c, d, I);
}

__global__ void k_legacy_first(LShape S, LBox dom, int axis, const double* __restrict__ phi, double* __restrict__ d1) {
  const long long s = S.st[axis];
  const int lo = dom.lo[axis], n = dom.hi[axis] - dom.lo[axis];
  LEGACY_FOR_BOX(dom) {
    LEGACY_INDEX(dom, S)
    const int pos = c_[axis] - lo;
    double r;
    if (pos < 3) {
      double p[6];
      const long long b = I - pos * s;
#pragma unroll
      for (int m = 0; m < 6; ++m) p[m] = phi[b + m * s];
      r = mixed_lo(p, pos);
    } else if (pos >= n - 3) {
      double p[6];
      const long long b = I + (n - 1 - pos) * s - 5 * s;
#pragma unroll
      for (int m = 0; m < 6; ++m) p[m] = phi[b + m * s];
      r = mixed_hi(p, pos - (n - 3));
    } else {
      r = central6(phi[I - 3 * s], phi[I - 2 * s], phi[I - s], phi[I + s], phi[I + 2 * s], phi[I + 3 * s]);
    }
    d1[I] = r;
  }
}

__global__ void k_legacy_second(LShape S, LBox dom, int axis, const double* __restrict__ phi,
                                const double* __restrict__ d1, double* __restrict__ d2) {
  const long long s = S.st[axis];
  const int lo = dom.lo[axis], n = dom.hi[axis] - dom.lo[axis];
  LEGACY_FOR_BOX(dom) {
    LEGACY_INDEX(dom, S)
    const int pos = c_[axis] - lo;
    double r;
    if (pos < 3) {
      double p[6];
      const long long b = I - pos * s;
#pragma unroll
      for (int m = 0; m < 6; ++m) p[m] = d1[b + m * s];
      r = mixed_lo(p, pos);
    } else if (pos >= n - 3) {
      double p[6];
      const long long b = I + (n - 1 - pos) * s - 5 * s;
#pragma unroll
      for (int m = 0; m < 6; ++m) p[m] = d1[b + m * s];
      r = mixed_hi(p, pos - (n - 3));
    } else {
      double dd = -tol_diff(d1[I + s], d1[I - s]) / 2;
      r = kahan3(2 * (phi[I + s] + phi[I - s]), -4 * phi[I], dd);
    }
    d2[I] = r;
  }
}

// MEG cell-centre derivative (cell_center_derivs.jl:80-122).
// POST 0: out = clip(v).  POST 1: out = clip(scale*(prev - v)) with prev = out's current content
// (the second outer derivative of a conservative metric: hat = scale*(outer1 - outer2), then the eps clip).
template <int POST>
__global__ void k_legacy_meg(LShape S, LBox dom, int axis, const double* __restrict__ phi, const double* __restrict__ d1,
                             const double* __restrict__ d2, double* __restrict__ out, double scale) {
  const long long s = S.st[axis];
  const int lo = dom.lo[axis], n = dom.hi[axis] - dom.lo[axis];
  const double e = 50 * DBL_EPSILON;
  LEGACY_FOR_BOX(dom) {
    LEGACY_INDEX(dom, S)
    const int pos = c_[axis] - lo;
    const double L0 = phiL(phi[I], d1[I], d2[I]), R0 = phiR(phi[I], d1[I], d2[I]);
    double hi, lw;
    if (pos == n - 1) hi = L0;
    else hi = (L0 + phiR(phi[I + s], d1[I + s], d2[I + s])) / 2;
    if (pos == 0) lw = R0;
    else lw = (phiL(phi[I - s], d1[I - s], d2[I - s]) + R0) / 2;
    double v = clip(hi - lw, e);
    if (POST == 0) out[I] = v;
    else {
      double h = scale * (out[I] - v);
      out[I] = clip(h, e);
    }
  }
}

// interpolate_to_edge! (interp_to_edges.jl:15-79): edge[I] holds the I+1/2 value
__global__ void k_legacy_interp(LShape S, LBox dom, int axis, int whole, const double* __restrict__ phi,
                                const double* __restrict__ d1, const double* __restrict__ d2, double* __restrict__ edge) {
  const long long s = S.st[axis];
  const int lo = dom.lo[axis], n = dom.hi[axis] - dom.lo[axis];
  // whole-array domain: inner = [lo+1, hi-1), lo index lo+1 writes edge[lo], hi index hi-2 writes edge[hi-2];
  // otherwise inner = dom, lo index lo writes edge[lo-1], hi index hi-1 writes edge[hi-1]
  const int first = whole ? 1 : 0, last = whole ? n - 2 : n - 1;
  LEGACY_FOR_BOX(dom) {
    LEGACY_INDEX(dom, S)
    const int pos = c_[axis] - lo;
    if (pos < first || pos > last) continue;
    const double L0 = phiL(phi[I], d1[I], d2[I]);
    if (pos == last) edge[I] = L0;
    else edge[I] = (L0 + phiR(phi[I + s], d1[I + s], d2[I + s])) / 2;
    if (pos == first) edge[I - s] = phiR(phi[I], d1[I], d2[I]);
  }
}

__global__ void k_legacy_centroid(LShape S, LBox dom, const double* __restrict__ xn, const double* __restrict__ yn,
                                  const double* __restrict__ zn, int off, int nn0, int nn1, double* cx, double* cy,
                                  double* cz) {
  LEGACY_FOR_BOX(dom) {
    LEGACY_INDEX(dom, S)
    const long long sn1 = nn0, sn2 = (long long)nn0 * nn1;
    const long long b = (c_[0] - off) + sn1 * (c_[1] - off) + sn2 * (c_[2] - off);
    const double* A[3] = {xn, yn, zn};
    double* C[3] = {cx, cy, cz};
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      const double* a = A[q] + b;
      // same summation order as 3d.jl:565-575
      C[q][I] = 0.125 * (a[0] + a[1] + a[1 + sn1] + a[sn1] + a[sn2] + a[1 + sn2] + a[1 + sn1 + sn2] + a[sn1 + sn2]);
    }
  }
}

// _check_valid_metrics (3d.jl:600-677) folded into the producers (fused path): a kernel that writes a checked array tests
// its own values inside the check box (cell.domain; for edge arrays shrunk by one cell along the edge axis) and raises
// vflags[0] (cell metrics) / vflags[1] (edge metrics); no separate pass over 46 arrays.  vflags == nullptr: no check.
struct LCheck {
  int* vflags;
  LBox box;
};
__device__ __forceinline__ bool in_box(const LBox& b, int c0, int c1, int c2) {
  return c0 >= b.lo[0] && c0 < b.hi[0] && c1 >= b.lo[1] && c1 < b.hi[1] && c2 >= b.lo[2] && c2 < b.hi[2];
}

// jacobian_3d_kernel! (forward_metrics.jl:155-167): StaticArrays det = x0 . (x1 x x2) over the columns
__global__ void k_legacy_det(LShape S, LBox dom, const double* __restrict__ fwd, long long nc, double* __restrict__ J,
                             LCheck chk = LCheck{nullptr, {}}) {
  LEGACY_FOR_BOX(dom) {
    LEGACY_INDEX(dom, S)
    const double a0 = fwd[0 * nc + I], b0 = fwd[1 * nc + I], c0 = fwd[2 * nc + I];   // x_xi, x_eta, x_zeta
    const double a1 = fwd[3 * nc + I], b1 = fwd[4 * nc + I], c1 = fwd[5 * nc + I];   // y_*
    const double a2 = fwd[6 * nc + I], b2 = fwd[7 * nc + I], c2 = fwd[8 * nc + I];   // z_*
    const double j = a0 * (b1 * c2 - b2 * c1) + a1 * (b2 * c0 - b0 * c2) + a2 * (b0 * c1 - b1 * c0);
    J[I] = j;
    if (chk.vflags && in_box(chk.box, c_[0], c_[1], c_[2]) && !(isfinite(j) && j > 0.0)) atomicOr(chk.vflags, 1);  // J finite and > 0
  }
}

__global__ void k_legacy_div9(LShape S, LBox dom, const double* __restrict__ hat, const double* __restrict__ J,
                              long long nc, double* __restrict__ inv, LCheck chk = LCheck{nullptr, {}}) {
  LEGACY_FOR_BOX(dom) {
    LEGACY_INDEX(dom, S)
    const double j = J[I];
    bool bad = false;
#pragma unroll
    for (int m = 0; m < 9; ++m) {
      const double v = hat[m * nc + I] / j;
      inv[m * nc + I] = v;
      bad |= !isfinite(v);
    }
    if (chk.vflags && bad && in_box(chk.box, c_[0], c_[1], c_[2])) atomicOr(chk.vflags, 1);
  }
}

// 2-D: _centroid_coordinates_kernel! (2d.jl:734-746), same summation order
__global__ void k_legacy_centroid2d(LShape S, LBox dom, const double* __restrict__ xn, const double* __restrict__ yn,
                                    int off, int nn0, double* cx, double* cy) {
  LEGACY_FOR_BOX(dom) {
    LEGACY_INDEX(dom, S)
    const long long sn1 = nn0;
    const long long b = (c_[0] - off) + sn1 * (c_[1] - off);
    const double *ax = xn + b, *ay = yn + b;
    cx[I] = 0.25 * (ax[0] + ax[1] + ax[1 + sn1] + ax[sn1]);
    cy[I] = 0.25 * (ay[0] + ay[1] + ay[1 + sn1] + ay[sn1]);
  }
}

// inverse_metric_2d_kernel! (conservative_metrics.jl:218-236, StaticArrays det / inv of [x_xi x_eta; y_xi y_eta]) and
// the hatted metrics = inverse * J (:205-210); fwd = [x_xi, x_eta, y_xi, y_eta][nc]
__global__ void k_legacy_inv2d(LShape S, LBox dom, const double* __restrict__ fwd, long long nc, double* __restrict__ J,
                               double* __restrict__ inv, double* __restrict__ hat) {
  LEGACY_FOR_BOX(dom) {
    LEGACY_INDEX(dom, S)
    const double a11 = fwd[I], a12 = fwd[nc + I], a21 = fwd[2 * nc + I], a22 = fwd[3 * nc + I];
    const double d = a11 * a22 - a12 * a21;
    const double v[4] = {a22 / d, -(a12 / d), -(a21 / d), a11 / d};  // xi_x, xi_y, eta_x, eta_y
    J[I] = d;
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      inv[m * nc + I] = v[m];
      hat[m * nc + I] = v[m] * d;
    }
  }
}

// ---------------------------------------------------------------------------
// Fused passes (round 2).  One block owns a tile of T cells along the derivative axis x LN lines and runs the three
// array passes of a derivative (first derivative, second derivative, MEG reconstruction or edge interpolation) through
// shared memory: the source field is read ONCE (products a*b / a*b - c*d formed on the fly, never materialised), d1 and
// the half-jump h = 1/2 d1 + 1/12 d2 never touch global memory.  Every value is produced by the same device functions,
// in the same order, as the pass-by-pass kernels above (bit-identical by construction; CG_LEGACY_FUSED=0 runs the old
// path, tests compare the two bit for bit).  The stencils are selected from the ABSOLUTE position along the axis, so a
// tile near the ends of the domain extends its ranges to the six-point one-sided windows:
//   h  on [t0-1, t1+1)                        (interp: [t0, t1+1))
//   d1 on the h range +-1,  and [0,6) / [n-6,n) when the h range touches the first / last three cells
//   phi on the d1 range +-3, and [0,6) / [n-6,n) likewise                       (all clamped to [0, n))
// axis != 0: lines = 32 consecutive xi (lanes, coalesced), T = 54 positions walked by the 8 warps;
// axis == 0: lines = 8 consecutive eta rows (warps), T = 118 positions along xi (lanes, coalesced).
// Several fields (jobs) share one launch through blockIdx.y.
// ---------------------------------------------------------------------------
// one-sided six-point stencils as ONE code path: row 0..2 = mixed_lo(which = row), row 3..5 = mixed_hi(which = row - 3).
// The products are the same doubles as in mixed_lo / mixed_hi (x * 1 and x * -1 are exact), so the result is bit-identical,
// but lanes of one warp that need different rows do not serialise three variants.
__constant__ double LEGACY_EDGE[6][6] = {
    {-137.0 / 60.0, 5.0, -5.0, 10.0 / 3.0, -(5.0 / 4.0), 1.0 / 5.0},
    {-1.0 / 5.0, -(13.0 / 12.0), 2.0, -1.0, 1.0 / 3.0, -(1.0 / 20.0)},
    {1.0 / 20.0, -(1.0 / 2.0), -(1.0 / 3.0), 1.0, -(1.0 / 4.0), 1.0 / 30.0},
    {-1.0 / 30.0, 1.0 / 4.0, -1.0, 1.0 / 3.0, 1.0 / 2.0, -(1.0 / 20.0)},
    {1.0 / 20.0, -(1.0 / 3.0), 1.0, -2.0, 13.0 / 12.0, 1.0 / 5.0},
    {-1.0 / 5.0, 5.0 / 4.0, -(10.0 / 3.0), 5.0, -5.0, 137.0 / 60.0}};
__device__ __forceinline__ double mixed_edge(const double* w, int row) {
  double v[6];
#pragma unroll
  for (int m = 0; m < 6; ++m) v[m] = LEGACY_EDGE[row][m] * w[m];
  return clip(kahan6(v), 50 * DBL_EPSILON);
}

struct LJob {
  const double *a, *b, *c, *d;  // source = a | a*b | a*b - c*d
  double* out;
  int map, idx;                 // TMA path: `a` is row `idx` of the array block behind tensor map `map`
  int chk;                      // 0: output not checked, 1: finite inside the check box -> cell flag, 2: -> edge flag
};
// The two array blocks a plain source field can live in ([rows][ncell]: the cell SoA and the centroid scratch), for the
// TMA tensor maps of a launch.
struct LTmaCtx {
  const double* base[2];
  int rows[2];
  long long nc;
};
constexpr int LMAXJ = 19;
struct LJobs {
  LJob j[LMAXJ];
};
enum { OP_CCD = 0, OP_CCD_POST = 1, OP_INTERP = 2 };

// TMA = true (plain source fields, even row length): the brick of the source field — 32 lines x 66 positions, or 8 lines
// x 128 positions along xi — arrives by ONE tensor-map TMA load (cp.async.bulk.tensor.4d, SASS UTMALDG) issued by one
// thread and signalled on an mbarrier, instead of a load + store per element; the tensor map covers the whole array
// block [row][k][j][i], out-of-range parts of a box are zero-filled by the hardware (they are never read).
template <bool AX0, int SRC, int OP, bool TMA>
__global__ void __launch_bounds__(256, 4) k_legacy_fused(LShape S, LBox dom, int axis, int whole, LJobs jobs, double scale,
                                                         const __grid_constant__ CUtensorMap tm0,
                                                         const __grid_constant__ CUtensorMap tm1, LCheck chk, int pos_lo,
                                                         int pos_hi) {
  constexpr int T = AX0 ? 118 : 54, LN = AX0 ? 8 : 32, PS = AX0 ? 32 : 8;
  constexpr int WP = T + 10, WD = T + 5, WH = T + 2;
  constexpr int LS = AX0 ? 1 : LN;  // shared-memory stride between consecutive positions of a line (d1, h)
  // The source brick has room for one more element at either end of the xi extent: a TMA box must start at a 16-byte
  // aligned address, i.e. at an EVEN xi index, so the box starts one cell early when the tile starts at an odd one
  // (`sft`); row pitch 34 (lines along xi) / 130 (positions along xi).
  constexpr int PLN = AX0 ? LN : LN + 2, PWP = AX0 ? WP + 2 : WP;
  constexpr int LSP = AX0 ? 1 : PLN;  // stride between consecutive positions of a line in the source brick
  constexpr int IT_P = (WP + PS - 1) / PS, IT_D = (WD + PS - 1) / PS, IT_H = (WH + PS - 1) / PS, IT_O = (T + PS - 1) / PS;
  __shared__ __align__(128) double sphi[PLN * PWP];
  __shared__ double sd1[LN * WD], sh[LN * WH];
  __shared__ __align__(8) unsigned long long mbar;
  const LJob J = jobs.j[blockIdx.y];
  const int n = dom.hi[axis] - dom.lo[axis];
  const int ntile = (pos_hi - pos_lo + T - 1) / T;  // tiles cover the positions [pos_lo, pos_hi) of the axis
  const int la = AX0 ? 1 : 0;                      // axis the lines are counted along
  const int oa = AX0 ? 2 : (axis == 1 ? 2 : 1);    // the remaining axis
  const int nl = dom.hi[la] - dom.lo[la], ngr = (nl + LN - 1) / LN;
  unsigned bb = blockIdx.x;
  const int tile = (int)(bb % (unsigned)ntile);
  bb /= (unsigned)ntile;
  const int gr = (int)(bb % (unsigned)ngr), oi = (int)(bb / (unsigned)ngr);
  const int t0 = pos_lo + tile * T, t1 = min(t0 + T, pos_hi);
  // ranges (absolute positions along the axis, [lo, hi)): cg_legacy_ranges.h
  const LegacyRanges rg = legacy_tile_ranges(n, t0, t1, OP == OP_INTERP);
  const int hl = rg.hl, hh = rg.hh, dl = rg.dl, dh = rg.dh, pl = rg.pl, ph = rg.ph;
  const int line = AX0 ? threadIdx.y : threadIdx.x, p0 = AX0 ? threadIdx.x : threadIdx.y;
  const int li = gr * LN + line;
  const bool lok = li < nl;
  const long long s = S.st[axis];
  const long long base = (long long)(dom.lo[la] + li) * S.st[la] + (long long)(dom.lo[oa] + oi) * S.st[oa] +
                         (long long)dom.lo[axis] * s;  // array index of position 0 of this line
  // per-line shared-memory rows; X_at(p) = address of absolute position p
  const int sft = TMA ? ((dom.lo[0] + (AX0 ? pl : gr * LN)) & 1) : 0;
  double* const phi_l = sphi + (AX0 ? line * PWP : line) + sft - pl * LSP;
  double* const d1_l = sd1 + (AX0 ? line * WD : line) - dl * LS;
  double* const h_l = sh + (AX0 ? line * WH : line) - hl * LS;
  // ---- phase 1: source field ----
  if (TMA) {
    const unsigned mb = (unsigned)__cvta_generic_to_shared(&mbar);
    const int tid = threadIdx.y * 32 + threadIdx.x;
    if (tid == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mb));
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
      // box origin: lines from dom.lo[la] + gr * LN, positions from dom.lo[axis] + pl, the third index, the array row
      int c[3];
      c[axis] = dom.lo[axis] + pl;
      c[la] = dom.lo[la] + gr * LN;
      c[oa] = dom.lo[oa] + oi;
      c[0] -= sft;
      const CUtensorMap* tm = J.map == 0 ? &tm0 : &tm1;
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"((unsigned)(PLN * PWP * sizeof(double)))
                   : "memory");
      asm volatile(
          "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(
              (unsigned)__cvta_generic_to_shared(sphi)),
          "l"(tm), "r"(c[0]), "r"(c[1]), "r"(c[2]), "r"(J.idx), "r"(mb)
          : "memory");
    }
    asm volatile(
        "{\n\t.reg .pred p;\n\tCG_TMA_WAIT:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra CG_TMA_DONE;\n\t"
        "bra CG_TMA_WAIT;\n\tCG_TMA_DONE:\n\t}" ::"r"(mb)
        : "memory");
  } else if (lok) {
    long long I = base + (long long)(pl + p0) * s;
#pragma unroll
    for (int k = 0; k < IT_P; ++k, I += PS * s) {
      const int p = pl + p0 + PS * k;
      if (p < ph) phi_l[p * LSP] = src_val<SRC>(J.a, J.b, J.c, J.d, I);
    }
  }
  __syncthreads();
  // Compute phases.  AX0 = false: warp w owns the CHUNK of 8 consecutive positions lo + 8 w .. lo + 8 w + 7 of its 32
  // lines; a chunk that lies in the central-stencil region reads its window once into registers (14 values for 8 first
  // derivatives instead of 48) — the register-blocked sweep along the derivative axis.  Chunks that touch the one-sided
  // regions or the end of a range, and the AX0 = true layout (lanes along the axis), take the per-position path.
  constexpr int CK = 8;
  auto d1_at = [&](int p) {  // k_legacy_first at absolute position p
    double r;
    if (p >= 3 && p < n - 3) {
      const double* q = phi_l + p * LSP;
      r = central6(q[-3 * LSP], q[-2 * LSP], q[-LSP], q[LSP], q[2 * LSP], q[3 * LSP]);
    } else {
      const bool low = p < 3;
      const double* q = phi_l + (low ? 0 : n - 6) * LSP;
      double w[6];
#pragma unroll
      for (int m = 0; m < 6; ++m) w[m] = q[m * LSP];
      r = mixed_edge(w, low ? p : p - (n - 6));
    }
    d1_l[p * LS] = r;
  };
  auto h_at = [&](int p) {  // k_legacy_second at p and the half-jump of the MEG states: phiL = f + h, phiR = f - h
    const double* d = d1_l + p * LS;
    double r;
    if (p >= 3 && p < n - 3) {
      const double* q = phi_l + p * LSP;
      const double dd = -tol_diff(d[LS], d[-LS]) / 2;
      r = kahan3(2 * (q[LSP] + q[-LSP]), -4 * q[0], dd);
    } else {
      const bool low = p < 3;
      const double* q = d1_l + (low ? 0 : n - 6) * LS;
      double w[6];
#pragma unroll
      for (int m = 0; m < 6; ++m) w[m] = q[m * LS];
      r = mixed_edge(w, low ? p : p - (n - 6));
    }
    h_l[p * LS] = (1.0 / 2.0) * d[0] + (1.0 / 12.0) * r;
  };
  const double e50 = 50 * DBL_EPSILON;
  const int first = whole ? 1 : 0, last = whole ? n - 2 : n - 1;  // interp: positions written (k_legacy_interp)
  // validity of a value just written at position p of this line (folded _check_valid_metrics)
  const bool checked = chk.vflags != nullptr && J.chk != 0;
  const int cl_ = dom.lo[la] + li, co_ = dom.lo[oa] + oi;
  const bool line_in = checked && cl_ >= chk.box.lo[la] && cl_ < chk.box.hi[la] && co_ >= chk.box.lo[oa] && co_ < chk.box.hi[oa];
  const int chk_lo = line_in ? chk.box.lo[axis] - dom.lo[axis] : 0, chk_hi = line_in ? chk.box.hi[axis] - dom.lo[axis] : 0;
  bool bad = false;  // raised in a register; ONE atomic per thread after the loops (an atomic inside them pins their order)
  auto check = [&](int p, double v) { bad |= p >= chk_lo && p < chk_hi && !(fabs(v) <= DBL_MAX); };
  auto out_at = [&](int p) {  // k_legacy_meg / k_legacy_interp at p
    const double* q = phi_l + p * LSP;
    const double* h = h_l + p * LS;
    double* o = J.out + base + (long long)p * s;
    const double L0 = q[0] + h[0];
    if (OP == OP_INTERP) {
      if (p < first || p > last) return;
      const double v = p == last ? L0 : (L0 + (q[LSP] - h[LS])) / 2;
      o[0] = v;
      check(p, v);
      if (p == first) o[-s] = q[0] - h[0];
    } else {
      const double R0 = q[0] - h[0];
      double hi, lw;
      if (p == n - 1) hi = L0;
      else hi = (L0 + (q[LSP] - h[LS])) / 2;
      if (p == 0) lw = R0;
      else lw = ((q[-LSP] + h[-LS]) + R0) / 2;
      const double v = clip(hi - lw, e50);
      const double w = OP == OP_CCD ? v : clip(scale * (o[0] - v), e50);
      o[0] = w;
      check(p, w);
    }
  };
  // ---- phase 2: first derivative ----
  if (lok) {
    if (!AX0) {
      const int pb = dl + CK * p0;
      if (pb >= 3 && pb + CK <= min(dh, n - 3)) {
        const double* q = phi_l + pb * LSP;
        double v[CK + 6];
#pragma unroll
        for (int m = 0; m < CK + 6; ++m) v[m] = q[(m - 3) * LSP];
#pragma unroll
        for (int k = 0; k < CK; ++k)
          d1_l[(pb + k) * LS] = central6(v[k], v[k + 1], v[k + 2], v[k + 4], v[k + 5], v[k + 6]);
      } else {
        for (int p = pb; p < min(pb + CK, dh); ++p) d1_at(p);
      }
    } else {
#pragma unroll
      for (int k = 0; k < IT_D; ++k) {
        const int p = dl + p0 + PS * k;
        if (p >= dh) break;
        d1_at(p);
      }
    }
  }
  __syncthreads();
  // ---- phase 3: second derivative and half-jump ----
  if (lok) {
    if (!AX0) {
      const int pb = hl + CK * p0;
      if (pb >= 3 && pb + CK <= min(hh, n - 3)) {
        const double *q = phi_l + pb * LSP, *d = d1_l + pb * LS;
        double f[CK + 2], g[CK + 2];
#pragma unroll
        for (int m = 0; m < CK + 2; ++m) {
          f[m] = q[(m - 1) * LSP];
          g[m] = d[(m - 1) * LS];
        }
#pragma unroll
        for (int k = 0; k < CK; ++k) {
          const double dd = -tol_diff(g[k + 2], g[k]) / 2;
          const double r = kahan3(2 * (f[k + 2] + f[k]), -4 * f[k + 1], dd);
          h_l[(pb + k) * LS] = (1.0 / 2.0) * g[k + 1] + (1.0 / 12.0) * r;
        }
      } else {
        for (int p = pb; p < min(pb + CK, hh); ++p) h_at(p);
      }
    } else {
#pragma unroll
      for (int k = 0; k < IT_H; ++k) {
        const int p = hl + p0 + PS * k;
        if (p >= hh) break;
        h_at(p);
      }
    }
  }
  __syncthreads();
  if (!lok) return;
  // ---- phase 4: MEG cell-centre derivative (k_legacy_meg) or edge interpolation (k_legacy_interp) ----
  if (!AX0) {
    const int pb = t0 + CK * p0;
    const bool fast = OP == OP_INTERP ? (pb >= first + 1 && pb + CK <= min(t1, last)) : (pb >= 1 && pb + CK <= min(t1, n - 1));
    if (pb >= t1) {
    } else if (fast) {
      const double *q = phi_l + pb * LSP, *h = h_l + pb * LS;
      double* o = J.out + base + (long long)pb * s;
      double Lm[CK + 2], Rm[CK + 2];  // states of the positions pb-1 .. pb+CK
#pragma unroll
      for (int m = OP == OP_INTERP ? 1 : 0; m < CK + 2; ++m) {  // (interp has no use for the state below the chunk)
        const double f = q[(m - 1) * LSP], hh_ = h[(m - 1) * LS];
        Lm[m] = f + hh_;
        Rm[m] = f - hh_;
      }
#pragma unroll
      for (int k = 0; k < CK; ++k, o += s) {
        const double hi = (Lm[k + 1] + Rm[k + 2]) / 2;
        if (OP == OP_INTERP) {
          o[0] = hi;
          check(pb + k, hi);
        } else {
          const double lw = (Lm[k] + Rm[k + 1]) / 2;
          const double v = clip(hi - lw, e50);
          const double w = OP == OP_CCD ? v : clip(scale * (o[0] - v), e50);
          o[0] = w;
          check(pb + k, w);
        }
      }
    } else {
      for (int p = pb; p < min(pb + CK, t1); ++p) out_at(p);
    }
  } else {
#pragma unroll
    for (int k = 0; k < IT_O; ++k) {
      const int p = t0 + p0 + PS * k;
      if (p >= t1) break;
      out_at(p);
    }
  }
  if (bad) atomicOr(chk.vflags + (J.chk - 1), 1);
}

// ---------------------------------------------------------------------------
// Marching variant for eta / zeta derivatives (CG_LEGACY_MARCH): a warp owns 32 consecutive xi (coalesced) and a segment
// of the derivative axis and SWEEPS along it with the whole stencil state in registers — seven source values, three first
// derivatives, two half-jumps, one face value — so every source value is loaded once, every intermediate is computed once
// (no tile halos), and there is no shared memory and no barrier at all: pushing phi(q) yields d1(q-3), h(q-4), the face
// value (L(q-5) + R(q-4)) / 2 and the output at q-5.  Only positions whose whole dependency cone uses CENTRAL stencils
// are marched, [5, n-5); the five cells at either end go through the tile kernel above (one-sided windows).  Same device
// functions, same operand order: bit-identical to both other paths.
// ---------------------------------------------------------------------------
#ifndef CG_MARCH_UNROLL
#define CG_MARCH_UNROLL 1
#endif
constexpr int MARCH_SEG = 64, MARCH_UNROLL = CG_MARCH_UNROLL;
template <int SRC, int OP>
__global__ void __launch_bounds__(128) k_legacy_march(LShape S, LBox dom, int axis, LJobs jobs, double scale, LCheck chk) {
  const LJob J = jobs.j[blockIdx.y];
  const int n = dom.hi[axis] - dom.lo[axis];
  const int oa = axis == 1 ? 2 : 1;
  const int nl = dom.hi[0] - dom.lo[0], ngr = (nl + 31) / 32;
  const int P0 = 5, P1 = n - 5, nseg = (P1 - P0 + MARCH_SEG - 1) / MARCH_SEG;
  const int no = dom.hi[oa] - dom.lo[oa];
  long long w = (long long)blockIdx.x * 4 + threadIdx.y;  // warp index -> (xi group, segment, other index)
  if (w >= (long long)ngr * nseg * no) return;
  const int gr = (int)(w % ngr);
  w /= ngr;
  const int seg = (int)(w % nseg), oi = (int)(w / nseg);
  const int li = gr * 32 + threadIdx.x;
  if (li >= nl) return;
  const int s0 = P0 + seg * MARCH_SEG, s1 = min(s0 + MARCH_SEG, P1);
  const long long s = S.st[axis];
  const long long base = (long long)(dom.lo[0] + li) + (long long)(dom.lo[oa] + oi) * S.st[oa] + (long long)dom.lo[axis] * s;
  const double e50 = 50 * DBL_EPSILON;
  // folded validity check (see LCheck)
  const bool checked = chk.vflags != nullptr && J.chk != 0;
  const int cl_ = dom.lo[0] + li, co_ = dom.lo[oa] + oi;
  const bool line_in = checked && cl_ >= chk.box.lo[0] && cl_ < chk.box.hi[0] && co_ >= chk.box.lo[oa] && co_ < chk.box.hi[oa];
  const int chk_lo = line_in ? chk.box.lo[axis] - dom.lo[axis] : 0, chk_hi = line_in ? chk.box.hi[axis] - dom.lo[axis] : 0;
  bool bad = false;
  double f0 = 0, f1 = 0, f2 = 0, f3 = 0, f4 = 0, f5 = 0, f6 = 0;  // phi(q-6 .. q)
  double d0 = 0, d1 = 0, d2 = 0;                                  // d1(q-5 .. q-3)
  double hA = 0, hB = 0, hiPrev = 0;                              // h(q-5), h(q-4), face value at q-6
  long long I = base + (long long)(s0 - 5) * s;
  double nxt = src_val<SRC>(J.a, J.b, J.c, J.d, I);
#pragma unroll(MARCH_UNROLL)
  for (int q = s0 - 5; q <= s1 + 4; ++q, I += s) {
    f0 = f1; f1 = f2; f2 = f3; f3 = f4; f4 = f5; f5 = f6; f6 = nxt;
    if (q < s1 + 4) nxt = src_val<SRC>(J.a, J.b, J.c, J.d, I + s);  // next position travels while this one is computed
    double prev = 0.0;
    if (OP == OP_CCD_POST && q >= s0 + 5) prev = J.out[I - 5 * s];
    if (q >= s0 + 1) {  // d1(q-3)
      d0 = d1; d1 = d2;
      d2 = central6(f0, f1, f2, f4, f5, f6);
    }
    if (q >= s0 + 3) {  // h(q-4): second derivative from d1(q-5), d1(q-3), phi(q-5 .. q-3)
      const double dd = -tol_diff(d2, d0) / 2;
      const double r = kahan3(2 * (f3 + f1), -4 * f2, dd);
      hA = hB;
      hB = (1.0 / 2.0) * d1 + (1.0 / 12.0) * r;
    }
    if (q >= s0 + 4) {  // face value between q-5 and q-4: (phiL(q-5) + phiR(q-4)) / 2
      const double hi = ((f1 + hA) + (f2 - hB)) / 2;
      if (q >= s0 + 5) {
        const int pp = q - 5;
        double v;
        if (OP == OP_INTERP) v = hi;
        else {
          v = clip(hi - hiPrev, e50);
          if (OP == OP_CCD_POST) v = clip(scale * (prev - v), e50);
        }
        J.out[I - 5 * s] = v;
        bad |= pp >= chk_lo && pp < chk_hi && !(fabs(v) <= DBL_MAX);
      }
      hiPrev = hi;
    }
  }
  if (bad) atomicOr(chk.vflags + (J.chk - 1), 1);
}

// Marching along XI (the contiguous axis): a lane cannot own a xi line and stay coalesced, so each warp moves its tile —
// 32 consecutive eta rows x one xi segment + 5 cells either side — through a warp-private shared-memory buffer: coalesced
// loads with the lanes along xi, then every lane sweeps ITS row out of shared memory exactly like k_legacy_march (row pitch
// odd: conflict-free), writes each output over the source value it no longer needs, and the warp stores the rows back
// coalesced.  Warp-private: __syncwarp only, no block barrier.  Central positions [5, n-5); the ends go to k_legacy_strip.
constexpr int MARCHX_SEG = 48, MARCHX_PITCH = MARCHX_SEG + 11;  // 59 doubles per row (odd)
template <int SRC, int OP>
__global__ void __launch_bounds__(64) k_legacy_march_x(LShape S, LBox dom, LJobs jobs, LCheck chk, int seglen) {
  static_assert(OP != OP_CCD_POST, "the second outer derivative along xi keeps the tile kernel");
  __shared__ double sbuf[2][32 * MARCHX_PITCH];
  double* const buf = sbuf[threadIdx.y];
  const LJob J = jobs.j[blockIdx.y];
  const int lane = threadIdx.x;
  const int n = dom.hi[0] - dom.lo[0];
  const int nl = dom.hi[1] - dom.lo[1], ngr = (nl + 31) / 32, no = dom.hi[2] - dom.lo[2];
  const int P0 = 5, P1 = n - 5, nseg = (P1 - P0 + seglen - 1) / seglen;
  long long w = (long long)blockIdx.x * 2 + threadIdx.y;  // warp index -> (segment, row group, k)
  if (w >= (long long)nseg * ngr * no) return;
  const int seg = (int)(w % nseg);
  w /= nseg;
  const int gr = (int)(w % ngr), oi = (int)(w / ngr);
  const int s0 = P0 + seg * seglen, s1 = min(s0 + seglen, P1);
  const int nrow = min(32, nl - gr * 32);
  const long long row0 = (long long)(dom.lo[1] + gr * 32) * S.st[1] + (long long)(dom.lo[2] + oi) * S.st[2] + dom.lo[0];
  const int len = s1 - s0 + 10;  // source values per row: positions s0-5 .. s1+4
  // ---- coalesced load: lanes along xi ----
  // (eight rows = sixteen loads in flight per lane: a row-by-row loop exposes the global latency 64 times per tile)
  static_assert(MARCHX_SEG + 10 <= 64, "two 32-lane pieces per row");
#pragma unroll 1
  for (int r0 = 0; r0 < nrow; r0 += 8) {
    double v[8][2];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const long long Ir = row0 + (long long)(r0 + u) * S.st[1] + (s0 - 5);
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        const int c = lane + 32 * t;
        v[u][t] = (r0 + u < nrow && c < len) ? src_val<SRC>(J.a, J.b, J.c, J.d, Ir + c) : 0.0;
      }
    }
#pragma unroll
    for (int u = 0; u < 8; ++u)
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        const int c = lane + 32 * t;
        if (r0 + u < nrow && c < len) buf[(r0 + u) * MARCHX_PITCH + c] = v[u][t];
      }
  }
  __syncwarp();
  // ---- sweep: lane = row ----
  bool bad = false;
  if (lane < nrow) {
    const double e50 = 50 * DBL_EPSILON;
    const bool checked = chk.vflags != nullptr && J.chk != 0;
    const int cl_ = dom.lo[1] + gr * 32 + lane, co_ = dom.lo[2] + oi;
    const bool line_in = checked && cl_ >= chk.box.lo[1] && cl_ < chk.box.hi[1] && co_ >= chk.box.lo[2] && co_ < chk.box.hi[2];
    const int chk_lo = line_in ? chk.box.lo[0] - dom.lo[0] : 0, chk_hi = line_in ? chk.box.hi[0] - dom.lo[0] : 0;
    double* const rowb = buf + lane * MARCHX_PITCH - (s0 - 5);  // rowb[q] = value at position q
    double f0 = 0, f1 = 0, f2 = 0, f3 = 0, f4 = 0, f5 = 0, f6 = 0, d0 = 0, d1 = 0, d2 = 0, hA = 0, hB = 0, hiPrev = 0;
    for (int q = s0 - 5; q <= s1 + 4; ++q) {
      f0 = f1; f1 = f2; f2 = f3; f3 = f4; f4 = f5; f5 = f6; f6 = rowb[q];
      if (q >= s0 + 1) {
        d0 = d1; d1 = d2;
        d2 = central6(f0, f1, f2, f4, f5, f6);
      }
      if (q >= s0 + 3) {
        const double dd = -tol_diff(d2, d0) / 2;
        const double r = kahan3(2 * (f3 + f1), -4 * f2, dd);
        hA = hB;
        hB = (1.0 / 2.0) * d1 + (1.0 / 12.0) * r;
      }
      if (q >= s0 + 4) {
        const double hi = ((f1 + hA) + (f2 - hB)) / 2;
        if (q >= s0 + 5) {
          const int pp = q - 5;
          const double v = OP == OP_INTERP ? hi : clip(hi - hiPrev, e50);
          rowb[pp] = v;  // the source value of position q-5 lives in f1 by now
          bad |= pp >= chk_lo && pp < chk_hi && !(fabs(v) <= DBL_MAX);
        }
        hiPrev = hi;
      }
    }
  }
  __syncwarp();
  // ---- coalesced store ----
  const int nout = s1 - s0;
#pragma unroll 4
  for (int r = 0; r < nrow; ++r) {
    double* const o = J.out + row0 + (long long)r * S.st[1] + s0;
#pragma unroll
    for (int t = 0; t < 2; ++t) {
      const int c = lane + 32 * t;
      if (c < nout) o[c] = buf[r * MARCHX_PITCH + 5 + c];
    }
  }
  if (bad) atomicOr(chk.vflags + (J.chk - 1), 1);
}

// The five cells at either end of a marched line (blockIdx.z = 0: positions 0..4, 1: n-5..n-1), one thread per line, all
// in registers: ten source values, seven first derivatives (three one-sided, four central), six half-jumps (three
// one-sided, three central), five outputs.  Local index i <-> absolute position a0 + i, a0 = 0 (low end) or n - 10 (high).
template <int SRC, int OP, bool high>
__device__ __forceinline__ void legacy_strip_body(const LShape& S, const LBox& dom, int axis, int whole, const LJobs& jobs, double scale,
                                                  const LCheck& chk) {
  const LJob J = jobs.j[blockIdx.y];
  const int n = dom.hi[axis] - dom.lo[axis];
  const int la = axis == 0 ? 1 : 0;               // axis the lines are counted along (lanes)
  const int oa = axis == 0 ? 2 : (axis == 1 ? 2 : 1);
  const int nl = dom.hi[la] - dom.lo[la], ngr = (nl + 31) / 32, no = dom.hi[oa] - dom.lo[oa];
  const long long w = (long long)blockIdx.x * 4 + threadIdx.y;
  if (w >= (long long)ngr * no) return;
  const int gr = (int)(w % ngr), oi = (int)(w / ngr);
  const int li = gr * 32 + threadIdx.x;
  if (li >= nl) return;
  const long long s = S.st[axis];
  const int a0 = high ? n - 10 : 0;
  const long long base = (long long)(dom.lo[la] + li) * S.st[la] + (long long)(dom.lo[oa] + oi) * S.st[oa] + (long long)dom.lo[axis] * s;
  double P[10], D[10], H[10];
#pragma unroll
  for (int i = 0; i < 10; ++i) P[i] = src_val<SRC>(J.a, J.b, J.c, J.d, base + (long long)(a0 + i) * s);
  // first derivatives: local 0..6 (low end) / 3..9 (high end)
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const bool need = high ? i >= 3 : i <= 6;
    const bool edge = high ? i >= 7 : i < 3;
    if (!need) { D[i] = 0.0; continue; }
    if (edge) {
      double wv[6];
#pragma unroll
      for (int m = 0; m < 6; ++m) wv[m] = P[(high ? 4 : 0) + m];
      D[i] = mixed_edge(wv, high ? 3 + (i - 7) : i);
    } else {
      D[i] = central6(P[i - 3 < 0 ? 0 : i - 3], P[i - 2 < 0 ? 0 : i - 2], P[i - 1 < 0 ? 0 : i - 1], P[i + 1 > 9 ? 9 : i + 1],
                      P[i + 2 > 9 ? 9 : i + 2], P[i + 3 > 9 ? 9 : i + 3]);
    }
  }
  // half-jumps: local 0..5 (low end) / 4..9 (high end)
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const bool need = high ? i >= 4 : i <= 5;
    const bool edge = high ? i >= 7 : i < 3;
    if (!need) { H[i] = 0.0; continue; }
    double r;
    if (edge) {
      double wv[6];
#pragma unroll
      for (int m = 0; m < 6; ++m) wv[m] = D[(high ? 4 : 0) + m];
      r = mixed_edge(wv, high ? 3 + (i - 7) : i);
    } else {
      const int im = i - 1 < 0 ? 0 : i - 1, ip = i + 1 > 9 ? 9 : i + 1;
      const double dd = -tol_diff(D[ip], D[im]) / 2;
      r = kahan3(2 * (P[ip] + P[im]), -4 * P[i], dd);
    }
    H[i] = (1.0 / 2.0) * D[i] + (1.0 / 12.0) * r;
  }
  const double e50 = 50 * DBL_EPSILON;
  const int first = whole ? 1 : 0, last = whole ? n - 2 : n - 1;
  const bool checked = chk.vflags != nullptr && J.chk != 0;
  const int cl_ = dom.lo[la] + li, co_ = dom.lo[oa] + oi;
  const bool line_in = checked && cl_ >= chk.box.lo[la] && cl_ < chk.box.hi[la] && co_ >= chk.box.lo[oa] && co_ < chk.box.hi[oa];
  const int chk_lo = line_in ? chk.box.lo[axis] - dom.lo[axis] : 0, chk_hi = line_in ? chk.box.hi[axis] - dom.lo[axis] : 0;
  bool bad = false;
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    const int i = (high ? 5 : 0) + k, p = a0 + i;  // local index, absolute position
    const int im = i - 1 < 0 ? 0 : i - 1, ip = i + 1 > 9 ? 9 : i + 1;
    double* o = J.out + base + (long long)p * s;
    const double L0 = P[i] + H[i];
    double v;
    if (OP == OP_INTERP) {
      if (p < first || p > last) continue;
      v = p == last ? L0 : (L0 + (P[ip] - H[ip])) / 2;
      o[0] = v;
      if (p == first) o[-s] = P[i] - H[i];
    } else {
      const double R0 = P[i] - H[i];
      const double hi = p == n - 1 ? L0 : (L0 + (P[ip] - H[ip])) / 2;
      const double lw = p == 0 ? R0 : ((P[im] + H[im]) + R0) / 2;
      v = clip(hi - lw, e50);
      if (OP == OP_CCD_POST) v = clip(scale * (o[0] - v), e50);
      o[0] = v;
    }
    bad |= p >= chk_lo && p < chk_hi && !(fabs(v) <= DBL_MAX);
  }
  if (bad) atomicOr(chk.vflags + (J.chk - 1), 1);
}
template <int SRC, int OP>
__global__ void __launch_bounds__(128) k_legacy_strip(LShape S, LBox dom, int axis, int whole, LJobs jobs, double scale, LCheck chk) {
  if (blockIdx.z == 0) legacy_strip_body<SRC, OP, false>(S, dom, axis, whole, jobs, scale, chk);
  else legacy_strip_body<SRC, OP, true>(S, dom, axis, whole, jobs, scale, chk);
}

// zero every cell of the arrays arr + blockIdx.z * nc that lies outside `keep` (the outputs are defined as zero outside
// the metric domain; the fused path overwrites everything inside, so only the shell has to be cleared).  One warp per
// (j, k) row: rows outside keep are cleared whole, the others only left and right of keep.
__global__ void k_legacy_zero_outside(LShape S, LBox keep, double* arr, long long nc) {
  double* a = arr + (long long)blockIdx.z * nc;
  const int lane = threadIdx.x & 31;
  const long long nrow = (long long)S.n[1] * S.n[2];
  for (long long row = blockIdx.x * (long long)(blockDim.x >> 5) + (threadIdx.x >> 5); row < nrow;
       row += (long long)gridDim.x * (blockDim.x >> 5)) {
    const int j = (int)(row % S.n[1]), k = (int)(row / S.n[1]);
    double* r = a + row * S.n[0];
    const bool in = j >= keep.lo[1] && j < keep.hi[1] && k >= keep.lo[2] && k < keep.hi[2];
    const int e0 = in ? keep.lo[0] : S.n[0];  // [0, e0) and [b1, n0) are cleared
    const int b1 = in ? keep.hi[0] : S.n[0];
    for (int i = lane; i < e0; i += 32) r[i] = 0.0;
    for (int i = b1 + lane; i < S.n[0]; i += 32) r[i] = 0.0;
  }
}

typedef CUresult (*LEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                              const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                              CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
LEncodeFn tensor_map_encoder() {
  const char* e = getenv("CG_LEGACY_TMA");  // debug switch (listed in include/curvigrid_cuda.h), read at every launch
  if (e && atoi(e) == 0) return nullptr;
  static const LEncodeFn fn = [] {
    void* h = dlopen("libcuda.so.1", RTLD_NOW | RTLD_LOCAL);
    return h ? (LEncodeFn)dlsym(h, "cuTensorMapEncodeTiled") : (LEncodeFn) nullptr;
  }();
  return fn;
}
// tensor map of an array block [rows][n2][n1][n0] with the box of a tile of the fused kernel along `axis`
bool make_tile_map(CUtensorMap& tm, const double* base, int rows, const LShape& S, int axis) {
  const LEncodeFn enc = tensor_map_encoder();
  if (!enc || !base || rows < 1 || (S.n[0] & 1) || (reinterpret_cast<uintptr_t>(base) & 15)) return false;
  const cuuint64_t dims[4] = {(cuuint64_t)S.n[0], (cuuint64_t)S.n[1], (cuuint64_t)S.n[2], (cuuint64_t)rows};
  const cuuint64_t strides[3] = {(cuuint64_t)S.n[0] * 8, (cuuint64_t)S.n[0] * S.n[1] * 8, (cuuint64_t)S.n[0] * S.n[1] * S.n[2] * 8};
  cuuint32_t box[4] = {1, 1, 1, 1};
  if (axis == 0) { box[0] = 130; box[1] = 8; }   // PWP x LN of k_legacy_fused<true>
  else { box[0] = 34; box[axis] = 64; }          // PLN x WP of k_legacy_fused<false>
  const cuuint32_t es[4] = {1, 1, 1, 1};
  return enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, (void*)base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int SRC, int OP>
void launch_tiles(const LShape& S, const LBox& md, int axis, int whole, const LJobs& jobs_in, int njobs, double scale,
                  cudaStream_t st, const LTmaCtx* tma, const LCheck chk, int pos_lo, int pos_hi) {
  const int n = pos_hi - pos_lo;  // positions the tiles cover
  const dim3 blk(32, 8);
  LJobs jobs = jobs_in;
  CUtensorMap tm[2];
  memset(tm, 0, sizeof tm);
  // xi derivatives keep the per-element loads: their 8-row x 130-value box measured slower than the loads it replaces
  // (256^3 interpolation launch 3.14 -> 3.36 ms), the 34 x 64 bricks of the eta / zeta derivatives faster (2.82 -> 2.40 ms)
  bool use_tma = SRC == 0 && tma != nullptr && axis != 0;
  bool need[2] = {false, false};
  for (int m = 0; m < njobs && use_tma; ++m) {  // every source field must be a row of one of the two array blocks
    bool found = false;
    for (int b = 0; b < 2 && !found; ++b) {
      const long long off = jobs.j[m].a - tma->base[b];
      if (tma->base[b] && off >= 0 && off % tma->nc == 0 && off / tma->nc < tma->rows[b]) {
        jobs.j[m].map = b;
        jobs.j[m].idx = (int)(off / tma->nc);
        need[b] = found = true;
      }
    }
    use_tma = found;
  }
  for (int b = 0; b < 2 && use_tma; ++b)
    if (need[b]) use_tma = make_tile_map(tm[b], tma->base[b], tma->rows[b], S, axis);
  if (axis == 0) {
    const long long nt = (n + 117) / 118, ng = (md.hi[1] - md.lo[1] + 7) / 8, no = md.hi[2] - md.lo[2];
    const dim3 grid((unsigned)(nt * ng * no), njobs);
    if (use_tma) k_legacy_fused<true, SRC, OP, SRC == 0><<<grid, blk, 0, st>>>(S, md, axis, whole, jobs, scale, tm[0], tm[1], chk, pos_lo, pos_hi);
    else k_legacy_fused<true, SRC, OP, false><<<grid, blk, 0, st>>>(S, md, axis, whole, jobs, scale, tm[0], tm[1], chk, pos_lo, pos_hi);
  } else {
    const int oa = axis == 1 ? 2 : 1;
    const long long nt = (n + 53) / 54, ng = (md.hi[0] - md.lo[0] + 31) / 32, no = md.hi[oa] - md.lo[oa];
    const dim3 grid((unsigned)(nt * ng * no), njobs);
    if (use_tma) k_legacy_fused<false, SRC, OP, SRC == 0><<<grid, blk, 0, st>>>(S, md, axis, whole, jobs, scale, tm[0], tm[1], chk, pos_lo, pos_hi);
    else k_legacy_fused<false, SRC, OP, false><<<grid, blk, 0, st>>>(S, md, axis, whole, jobs, scale, tm[0], tm[1], chk, pos_lo, pos_hi);
  }
}
bool legacy_march() {
  const char* e = getenv("CG_LEGACY_MARCH");  // debug switch (listed in include/curvigrid_cuda.h)
  return e ? atoi(e) != 0 : CG_LEGACY_MARCH_DEFAULT;
}
// one derivative pass over the metric domain: tiles everywhere, or (eta / zeta axes, long enough) the marching kernel on
// the central positions [5, n-5) + tiles on the five cells at either end.  Returns the number of launches.
template <int SRC, int OP>
int launch_fused(const LShape& S, const LBox& md, int axis, int whole, const LJobs& jobs, int njobs, double scale,
                 cudaStream_t st, const LTmaCtx* tma = nullptr, const LCheck chk = LCheck{nullptr, {}}) {
  const int n = md.hi[axis] - md.lo[axis];
  if (axis != 0 && n >= 42 && legacy_march()) {
    const int oa = axis == 1 ? 2 : 1;
    const long long ngr = (md.hi[0] - md.lo[0] + 31) / 32, nseg = (n - 10 + MARCH_SEG - 1) / MARCH_SEG, no = md.hi[oa] - md.lo[oa];
    // the five cells at either end first (a POST pass reads what the pass before wrote: same stream, any order is fine)
    k_legacy_strip<SRC, OP><<<dim3((unsigned)((ngr * no + 3) / 4), njobs, 2), dim3(32, 4), 0, st>>>(S, md, axis, whole, jobs, scale, chk);
    const long long warps = ngr * nseg * no;
    k_legacy_march<SRC, OP><<<dim3((unsigned)((warps + 3) / 4), njobs), dim3(32, 4), 0, st>>>(S, md, axis, jobs, scale, chk);
    (void)tma;
    return 2;
  }
  if constexpr (OP != OP_CCD_POST) {
    // experiment switch (header): measured 14.02 against 14.16 ms at 256^3 for the tiles it replaces, i.e. within the noise
    // (its sweep is as fast as k_legacy_march, the two shared-memory transposes cost what the tile halos cost): OFF
    static const bool marchx_env = getenv("CG_LEGACY_MARCHX") && atoi(getenv("CG_LEGACY_MARCHX")) == 1;
    if (axis == 0 && n >= 42 && md.hi[1] - md.lo[1] >= 8 && legacy_march() && marchx_env) {
      const long long ngr = (md.hi[1] - md.lo[1] + 31) / 32, no = md.hi[2] - md.lo[2];
      k_legacy_strip<SRC, OP><<<dim3((unsigned)((ngr * no + 3) / 4), njobs, 2), dim3(32, 4), 0, st>>>(S, md, axis, whole, jobs, scale, chk);
      const int nseg0 = (n - 10 + MARCHX_SEG - 1) / MARCHX_SEG, seglen = (n - 10 + nseg0 - 1) / nseg0;  // balanced segments
      const long long warps = (long long)((n - 10 + seglen - 1) / seglen) * ngr * no;
      k_legacy_march_x<SRC, OP><<<dim3((unsigned)((warps + 1) / 2), njobs), dim3(32, 2), 0, st>>>(S, md, jobs, chk, seglen);
      return 2;
    }
  }
  launch_tiles<SRC, OP>(S, md, axis, whole, jobs, njobs, scale, st, tma, chk, 0, n);
  return 1;
}
bool legacy_fused() {
  const char* e = getenv("CG_LEGACY_FUSED");  // debug switch (listed in include/curvigrid_cuda.h)
  return !(e && atoi(e) == 0);
}
bool box_equal(const LBox& a, const LBox& b) {
  for (int d = 0; d < 3; ++d)
    if (a.lo[d] != b.lo[d] || a.hi[d] != b.hi[d]) return false;
  return true;
}
// clear the shell outside `keep` of narr consecutive arrays (nothing to do when keep is the whole array)
void zero_outside(const LShape& S, const LBox& full, const LBox& keep, double* arr, int narr, long long nc, cudaStream_t st,
                  uint64_t& nl) {
  if (box_equal(full, keep) || narr == 0) return;
  long long b = ((long long)S.n[1] * S.n[2] + 7) / 8;
  if (b > 148LL * 8) b = 148LL * 8;
  k_legacy_zero_outside<<<dim3((unsigned)b, 1, narr), 256, 0, st>>>(S, keep, arr, nc);
  ++nl;
}
// the 19 (3-D) / 9 (2-D) / 3 (1-D) edge interpolations of one axis in one launch
void interp_axis_fused(const LShape& S, const LBox& full, const LBox& md, int axis, int whole, const double* const* phi,
                       int nphi, double* E, long long nc, cudaStream_t st, uint64_t& nl, const LTmaCtx* tma = nullptr,
                       LCheck chk = LCheck{nullptr, {}}, int first_checked = 1 << 30) {
  // written: md along the other axes; along `axis` positions [first-1, last] of md
  LBox keep = md;
  if (whole) keep.hi[axis] = md.hi[axis] - 1;
  else keep.lo[axis] = md.lo[axis] - 1;
  zero_outside(S, full, keep, E, nphi, nc, st, nl);
  LJobs jobs{};
  for (int m = 0; m < nphi; ++m) jobs.j[m] = {phi[m], nullptr, nullptr, nullptr, E + (long long)m * nc, 0, 0, m >= first_checked ? 2 : 0};
  chk.box.lo[axis] += 1;  // expand(domain, axis, -1)
  chk.box.hi[axis] -= 1;
  nl += launch_fused<0, OP_INTERP>(S, md, axis, whole, jobs, nphi, 1.0, st, tma, chk);
}

int blocks_for(long long n) {
  long long b = (n + 255) / 256;
  const long long cap = 148LL * 32;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}
long long box_count(const LBox& b) {
  return (long long)(b.hi[0] - b.lo[0]) * (b.hi[1] - b.lo[1]) * (b.hi[2] - b.lo[2]);
}

thread_local char g_legacy_err[256];

}  // namespace

// one persistent pair of flag words per device and host thread for cg_legacy_check_valid, another for the validity
// the fused update accumulates (a stream-ordered allocation per call made the memory pool trim and re-grow at every
// synchronisation: sporadic stalls of hundreds of ms)
namespace {
int* legacy_flag_words(int device, int which) {
  thread_local int* cache[2][64] = {};
  if (device < 0 || device >= 64) return nullptr;
  if (!cache[which][device] && cudaMalloc((void**)&cache[which][device], 2 * sizeof(int)) != cudaSuccess) return nullptr;
  return cache[which][device];
}
thread_local bool g_validity_ready[64] = {};
}  // namespace

extern "C" const char* cg_legacy_last_error(void) { return g_legacy_err; }

// see include/curvigrid_cuda.h
extern "C" int cg_legacy_validity(int device, int* cell_ok, int* edge_ok, void* stream) {
  g_legacy_err[0] = 0;
  if (!cell_ok || !edge_ok || device < 0 || device >= 64) {
    snprintf(g_legacy_err, sizeof g_legacy_err, "cg_legacy_validity: bad argument");
    return CG_ERR_ARG;
  }
  if (!g_validity_ready[device]) {
    snprintf(g_legacy_err, sizeof g_legacy_err, "cg_legacy_validity: no fused 3-D update has run on this thread and device");
    return CG_ERR_STATE;
  }
  if (cudaSetDevice(device) != cudaSuccess) return CG_ERR_CUDA;
  int h[2] = {1, 1};
  cudaMemcpyAsync(h, legacy_flag_words(device, 1), sizeof h, cudaMemcpyDeviceToHost, (cudaStream_t)stream);
  const cudaError_t ce = cudaStreamSynchronize((cudaStream_t)stream);
  if (ce != cudaSuccess) {
    snprintf(g_legacy_err, sizeof g_legacy_err, "cg_legacy_validity: %s", cudaGetErrorString(ce));
    return CG_ERR_CUDA;
  }
  *cell_ok = !h[0];
  *edge_ok = !h[1];
  return CG_OK;
}

// see include/curvigrid_cuda.h
extern "C" int cg_legacy_meg6_update_3d(int device, const double* x, const double* y, const double* z,
                                        const int64_t* nnode, int nhalo, int halo_coords_included, int symmetric,
                                        double* cell, double* edge, double* centroid, double* scratch,
                                        size_t scratch_bytes, void* stream, uint64_t* launches) {
  struct R { R() { nvtxRangePushA("cg_legacy_meg6_update_3d"); } ~R() { nvtxRangePop(); } } nvtx_range_;
  g_legacy_err[0] = 0;
  if (!x || !y || !z || !nnode || !cell || !edge || !scratch) {
    snprintf(g_legacy_err, sizeof g_legacy_err, "cg_legacy_meg6_update_3d: NULL argument");
    return CG_ERR_ARG;
  }
  if (nhalo < 5) {
    snprintf(g_legacy_err, sizeof g_legacy_err, "MEG6 needs nhalo >= 5 (DiscretizationSchemes.jl:53-60)");
    return CG_ERR_ARG;
  }
  LShape S;
  LBox full, dom;
  for (int d = 0; d < 3; ++d) {
    const int64_t nf = (halo_coords_included ? nnode[d] : nnode[d] + 2 * nhalo) - 1;
    if (nf - 2 * nhalo < 6) {
      snprintf(g_legacy_err, sizeof g_legacy_err, "MEG6 one-sided edge stencils need >= 6 interior cells per axis");
      return CG_ERR_ARG;
    }
    S.n[d] = (int)nf;
    full.lo[d] = 0;
    full.hi[d] = (int)nf;
    dom.lo[d] = nhalo;
    dom.hi[d] = (int)nf - nhalo;
  }
  S.st[0] = 1;
  S.st[1] = S.n[0];
  S.st[2] = (long long)S.n[0] * S.n[1];
  const long long nc = (long long)S.n[0] * S.n[1] * S.n[2];
  if (scratch_bytes < (size_t)(7 * nc) * sizeof(double)) {
    snprintf(g_legacy_err, sizeof g_legacy_err, "scratch too small: need %lld bytes", (long long)(7 * nc * 8));
    return CG_ERR_ARG;
  }
  cudaError_t ce = cudaSetDevice(device);
  if (ce != cudaSuccess) {
    snprintf(g_legacy_err, sizeof g_legacy_err, "cudaSetDevice: %s", cudaGetErrorString(ce));
    return CG_ERR_CUDA;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const LBox md = halo_coords_included ? full : dom;
  const int whole = halo_coords_included ? 1 : 0;
  uint64_t nl = 0;
  double *cx = scratch, *cy = scratch + nc, *cz = scratch + 2 * nc;
  double *d1 = scratch + 3 * nc, *d2 = scratch + 4 * nc, *src = scratch + 5 * nc, *tmp = scratch + 6 * nc;
  double* C[3] = {cx, cy, cz};
  double* J = cell;
  double* fwd = cell + nc;       // [q*3+axis]
  double* inv = cell + 10 * nc;  // [row*3+col]
  double* hat = cell + 19 * nc;
  const int bl = blocks_for(box_count(md)), bw = blocks_for(nc);
  if (legacy_fused()) {
    const LTmaCtx tma = {{cell, scratch}, {28, 3}, nc};  // source fields are rows of the cell SoA or the centroid scratch
    // folded _check_valid_metrics (3d.jl:600-677): J finite and > 0, forward and inverse metrics finite on cell.domain,
    // hatted edge metrics finite on expand(domain, axis, -1); read back with cg_legacy_validity
    LCheck chk{legacy_flag_words(device, 1), dom};
    if (chk.vflags) cudaMemsetAsync(chk.vflags, 0, 2 * sizeof(int), st);
    g_validity_ready[device < 64 && device >= 0 ? device : 0] = chk.vflags != nullptr;
    // 15 launches: every derivative is one fused launch (first + second derivative + MEG reconstruction through shared
    // memory), the fields that share an axis share the launch; only the shell outside the metric domain is cleared
    zero_outside(S, full, md, cell, 28, nc, st, nl);
    zero_outside(S, full, md, scratch, 3, nc, st, nl);
    k_legacy_centroid<<<bl, 256, 0, st>>>(S, md, x, y, z, halo_coords_included ? 0 : nhalo, (int)nnode[0], (int)nnode[1],
                                           cx, cy, cz);
    ++nl;
    if (centroid) cudaMemcpyAsync(centroid, scratch, (size_t)3 * nc * sizeof(double), cudaMemcpyDeviceToDevice, st);
    for (int ax = 0; ax < 3; ++ax) {  // forward metrics: d x_q / d xi_ax, q = 0..2
      LJobs jobs{};
      for (int q = 0; q < 3; ++q) jobs.j[q] = {C[q], nullptr, nullptr, nullptr, fwd + (long long)(q * 3 + ax) * nc, 0, 0, 1};
      nl += launch_fused<0, OP_CCD>(S, md, ax, whole, jobs, 3, 1.0, st, &tma, chk);
    }
    k_legacy_det<<<bl, 256, 0, st>>>(S, md, fwd, nc, J, chk);
    ++nl;
    for (int r = 0; r < 3; ++r) {  // conservative metrics, row r: the three columns share both launches
      const int s = (r + 1) % 3, u = (r + 2) % 3;
      LJobs j1{}, j2{};
      for (int c = 0; c < 3; ++c) {
        const int q1 = (c + 1) % 3, q2 = (c + 2) % 3;
        double* h = hat + (long long)(r * 3 + c) * nc;
        const double *a_s = fwd + (long long)(q1 * 3 + s) * nc, *a_u = fwd + (long long)(q1 * 3 + u) * nc;
        const double *b_s = fwd + (long long)(q2 * 3 + s) * nc, *b_u = fwd + (long long)(q2 * 3 + u) * nc;
        j1.j[c] = {a_s, C[q2], symmetric ? b_s : nullptr, symmetric ? C[q1] : nullptr, h, 0, 0, 0};
        j2.j[c] = {a_u, C[q2], symmetric ? b_u : nullptr, symmetric ? C[q1] : nullptr, h, 0, 0, 0};
      }
      if (symmetric) {
        nl += launch_fused<2, OP_CCD>(S, md, u, whole, j1, 3, 1.0, st);
        nl += launch_fused<2, OP_CCD_POST>(S, md, s, whole, j2, 3, 1.0 / 2.0, st);
      } else {
        nl += launch_fused<1, OP_CCD>(S, md, u, whole, j1, 3, 1.0, st);
        nl += launch_fused<1, OP_CCD_POST>(S, md, s, whole, j2, 3, 1.0, st);
      }
    }
    k_legacy_div9<<<bl, 256, 0, st>>>(S, md, hat, J, nc, inv, chk);
    ++nl;
    for (int ax = 0; ax < 3; ++ax) {
      const double* phi[19];
      for (int m = 0; m < 19; ++m) phi[m] = m == 0 ? J : cell + (long long)(9 + m) * nc;  // J, inv 0..8, hat 0..8
      interp_axis_fused(S, full, md, ax, whole, phi, 19, edge + (long long)ax * 19 * nc, nc, st, nl, &tma, chk, 10);
    }
  } else {
  if (device >= 0 && device < 64) g_validity_ready[device] = false;  // the pass-by-pass path does not accumulate validity
  cudaMemsetAsync(cell, 0, (size_t)28 * nc * sizeof(double), st);
  cudaMemsetAsync(edge, 0, (size_t)57 * nc * sizeof(double), st);
  cudaMemsetAsync(scratch, 0, (size_t)7 * nc * sizeof(double), st);
  k_legacy_centroid<<<bl, 256, 0, st>>>(S, md, x, y, z, halo_coords_included ? 0 : nhalo, (int)nnode[0], (int)nnode[1],
                                         cx, cy, cz);
  ++nl;
  if (centroid) cudaMemcpyAsync(centroid, scratch, (size_t)3 * nc * sizeof(double), cudaMemcpyDeviceToDevice, st);
  auto ccd = [&](const double* phi, int axis, double* out, int post, double scale) {
    k_legacy_first<<<bl, 256, 0, st>>>(S, md, axis, phi, d1);
    k_legacy_second<<<bl, 256, 0, st>>>(S, md, axis, phi, d1, d2);
    if (post == 0) k_legacy_meg<0><<<bl, 256, 0, st>>>(S, md, axis, phi, d1, d2, out, scale);
    else k_legacy_meg<1><<<bl, 256, 0, st>>>(S, md, axis, phi, d1, d2, out, scale);
    nl += 3;
  };
  for (int q = 0; q < 3; ++q)
    for (int ax = 0; ax < 3; ++ax) ccd(C[q], ax, fwd + (long long)(q * 3 + ax) * nc, 0, 1.0);
  k_legacy_det<<<bl, 256, 0, st>>>(S, md, fwd, nc, J);
  ++nl;
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) {
      const int q1 = (c + 1) % 3, q2 = (c + 2) % 3, s = (r + 1) % 3, u = (r + 2) % 3;
      double* h = hat + (long long)(r * 3 + c) * nc;
      const double *a_s = fwd + (long long)(q1 * 3 + s) * nc, *a_u = fwd + (long long)(q1 * 3 + u) * nc;
      const double *b_s = fwd + (long long)(q2 * 3 + s) * nc, *b_u = fwd + (long long)(q2 * 3 + u) * nc;
      // outer1 = (inner1)_u -> h ; then h = clip(scale*(h - (inner2)_s))
      if (symmetric) k_legacy_source<2><<<bw, 256, 0, st>>>(S, a_s, C[q2], b_s, C[q1], src);
      else k_legacy_source<1><<<bw, 256, 0, st>>>(S, a_s, C[q2], nullptr, nullptr, src);
      ccd(src, u, h, 0, 1.0);
      if (symmetric) k_legacy_source<2><<<bw, 256, 0, st>>>(S, a_u, C[q2], b_u, C[q1], src);
      else k_legacy_source<1><<<bw, 256, 0, st>>>(S, a_u, C[q2], nullptr, nullptr, src);
      ccd(src, s, h, 1, symmetric ? (1.0 / 2.0) : 1.0);
      nl += 2;
    }
  k_legacy_div9<<<bl, 256, 0, st>>>(S, md, hat, J, nc, inv);
  ++nl;
  for (int ax = 0; ax < 3; ++ax) {
    double* E = edge + (long long)ax * 19 * nc;
    for (int m = 0; m < 19; ++m) {
      const double* phi = m == 0 ? J : cell + (long long)(9 + m) * nc;  // J, inv 0..8, hat 0..8
      k_legacy_first<<<bl, 256, 0, st>>>(S, md, ax, phi, d1);
      k_legacy_second<<<bl, 256, 0, st>>>(S, md, ax, phi, d1, d2);
      k_legacy_interp<<<bl, 256, 0, st>>>(S, md, ax, whole, phi, d1, d2, E + (long long)m * nc);
      nl += 3;
    }
  }
  }
  (void)tmp;
  ce = cudaGetLastError();
  if (ce != cudaSuccess) {
    snprintf(g_legacy_err, sizeof g_legacy_err, "legacy kernels: %s", cudaGetErrorString(ce));
    return CG_ERR_CUDA;
  }
  if (launches) *launches = nl;
  return CG_OK;
}

// see include/curvigrid_cuda.h
extern "C" int cg_legacy_meg6_update_2d(int device, const double* x, const double* y, const int64_t* nnode, int nhalo,
                                        int halo_coords_included, double* cell, double* edge, double* centroid,
                                        double* scratch, size_t scratch_bytes, void* stream, uint64_t* launches) {
  struct R { R() { nvtxRangePushA("cg_legacy_meg6_update_2d"); } ~R() { nvtxRangePop(); } } nvtx_range_;
  g_legacy_err[0] = 0;
  if (!x || !y || !nnode || !cell || !edge || !scratch) {
    snprintf(g_legacy_err, sizeof g_legacy_err, "cg_legacy_meg6_update_2d: NULL argument");
    return CG_ERR_ARG;
  }
  if (nhalo < 5) {
    snprintf(g_legacy_err, sizeof g_legacy_err, "MEG6 needs nhalo >= 5 (DiscretizationSchemes.jl:53-60)");
    return CG_ERR_ARG;
  }
  LShape S;
  LBox full, dom;
  for (int d = 0; d < 2; ++d) {
    const int64_t nf = (halo_coords_included ? nnode[d] : nnode[d] + 2 * nhalo) - 1;
    if (nf - 2 * nhalo < 6) {
      snprintf(g_legacy_err, sizeof g_legacy_err, "MEG6 one-sided edge stencils need >= 6 interior cells per axis");
      return CG_ERR_ARG;
    }
    S.n[d] = (int)nf;
    full.lo[d] = 0;
    full.hi[d] = (int)nf;
    dom.lo[d] = nhalo;
    dom.hi[d] = (int)nf - nhalo;
  }
  S.n[2] = 1;
  full.lo[2] = dom.lo[2] = 0;
  full.hi[2] = dom.hi[2] = 1;
  S.st[0] = 1;
  S.st[1] = S.n[0];
  S.st[2] = (long long)S.n[0] * S.n[1];
  const long long nc = (long long)S.n[0] * S.n[1];
  if (scratch_bytes < (size_t)(4 * nc) * sizeof(double)) {
    snprintf(g_legacy_err, sizeof g_legacy_err, "scratch too small: need %lld bytes", (long long)(4 * nc * 8));
    return CG_ERR_ARG;
  }
  cudaError_t ce = cudaSetDevice(device);
  if (ce != cudaSuccess) {
    snprintf(g_legacy_err, sizeof g_legacy_err, "cudaSetDevice: %s", cudaGetErrorString(ce));
    return CG_ERR_CUDA;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const LBox md = halo_coords_included ? full : dom;
  const int whole = halo_coords_included ? 1 : 0;
  uint64_t nl = 0;
  double *cx = scratch, *cy = scratch + nc, *d1 = scratch + 2 * nc, *d2 = scratch + 3 * nc;
  double* C[2] = {cx, cy};
  double* J = cell;
  double* fwd = cell + nc;      // [q*2+axis]
  double* inv = cell + 5 * nc;  // xi_x, xi_y, eta_x, eta_y
  double* hat = cell + 9 * nc;
  const int bl = blocks_for(box_count(md));
  if (legacy_fused()) {
    const LTmaCtx tma = {{cell, scratch}, {13, 2}, nc};
    zero_outside(S, full, md, cell, 13, nc, st, nl);
    zero_outside(S, full, md, scratch, 2, nc, st, nl);
    k_legacy_centroid2d<<<bl, 256, 0, st>>>(S, md, x, y, halo_coords_included ? 0 : nhalo, (int)nnode[0], cx, cy);
    ++nl;
    if (centroid) cudaMemcpyAsync(centroid, scratch, (size_t)2 * nc * sizeof(double), cudaMemcpyDeviceToDevice, st);
    for (int ax = 0; ax < 2; ++ax) {  // forward_metrics.jl:36-60
      LJobs jobs{};
      for (int q = 0; q < 2; ++q) jobs.j[q] = {C[q], nullptr, nullptr, nullptr, fwd + (long long)(q * 2 + ax) * nc, 0, 0, 0};
      nl += launch_fused<0, OP_CCD>(S, md, ax, whole, jobs, 2, 1.0, st, &tma);
    }
    k_legacy_inv2d<<<bl, 256, 0, st>>>(S, md, fwd, nc, J, inv, hat);
    ++nl;
    for (int ax = 0; ax < 2; ++ax) {  // MetricDiscretizationSchemes.jl:58-89
      const double* phi[9];
      for (int m = 0; m < 9; ++m) phi[m] = m == 0 ? J : cell + (long long)(4 + m) * nc;  // J, inv 0..3, hat 0..3
      interp_axis_fused(S, full, md, ax, whole, phi, 9, edge + (long long)ax * 9 * nc, nc, st, nl, &tma);
    }
  } else {
  cudaMemsetAsync(cell, 0, (size_t)13 * nc * sizeof(double), st);
  cudaMemsetAsync(edge, 0, (size_t)18 * nc * sizeof(double), st);
  cudaMemsetAsync(scratch, 0, (size_t)4 * nc * sizeof(double), st);
  k_legacy_centroid2d<<<bl, 256, 0, st>>>(S, md, x, y, halo_coords_included ? 0 : nhalo, (int)nnode[0], cx, cy);
  ++nl;
  if (centroid) cudaMemcpyAsync(centroid, scratch, (size_t)2 * nc * sizeof(double), cudaMemcpyDeviceToDevice, st);
  for (int q = 0; q < 2; ++q)
    for (int ax = 0; ax < 2; ++ax) {  // forward_metrics.jl:36-60
      k_legacy_first<<<bl, 256, 0, st>>>(S, md, ax, C[q], d1);
      k_legacy_second<<<bl, 256, 0, st>>>(S, md, ax, C[q], d1, d2);
      k_legacy_meg<0><<<bl, 256, 0, st>>>(S, md, ax, C[q], d1, d2, fwd + (long long)(q * 2 + ax) * nc, 1.0);
      nl += 3;
    }
  k_legacy_inv2d<<<bl, 256, 0, st>>>(S, md, fwd, nc, J, inv, hat);
  ++nl;
  for (int ax = 0; ax < 2; ++ax) {  // MetricDiscretizationSchemes.jl:58-89
    double* E = edge + (long long)ax * 9 * nc;
    for (int m = 0; m < 9; ++m) {
      const double* phi = m == 0 ? J : cell + (long long)(4 + m) * nc;  // J, inv 0..3, hat 0..3
      k_legacy_first<<<bl, 256, 0, st>>>(S, md, ax, phi, d1);
      k_legacy_second<<<bl, 256, 0, st>>>(S, md, ax, phi, d1, d2);
      k_legacy_interp<<<bl, 256, 0, st>>>(S, md, ax, whole, phi, d1, d2, E + (long long)m * nc);
      nl += 3;
    }
  }
  }
  ce = cudaGetLastError();
  if (ce != cudaSuccess) {
    snprintf(g_legacy_err, sizeof g_legacy_err, "legacy kernels: %s", cudaGetErrorString(ce));
    return CG_ERR_CUDA;
  }
  if (launches) *launches = nl;
  return CG_OK;
}

namespace {
// centroids of a 1-D grid (1d.jl:445-455) on the metric domain
__global__ void k_legacy_centroid1d(LShape S, LBox dom, const double* __restrict__ xn, int off, double* cx) {
  LEGACY_FOR_BOX(dom) {
    LEGACY_INDEX(dom, S)
    const double* a = xn + (c_[0] - off);
    cx[I] = 0.5 * (a[0] + a[1]);
  }
}
// conservative_metrics! for a 1-D grid (conservative_metrics.jl:239-245): xi_x = 1 / x_xi, J = |x_xi|, hatted = xi_x J
__global__ void k_legacy_inv1d(LShape S, LBox dom, const double* __restrict__ xxi, double* __restrict__ J,
                               double* __restrict__ inv, double* __restrict__ hat) {
  LEGACY_FOR_BOX(dom) {
    LEGACY_INDEX(dom, S)
    const double f = xxi[I];
    const double g = 1.0 / f, j = fabs(f);
    J[I] = j;
    inv[I] = g;
    hat[I] = g * j;
  }
}
// _check_valid_metrics (1d.jl:457-482, 2d.jl:748-788, 3d.jl:600-677) as ONE fused reduction: the listed cell rows must
// be finite on cell.domain (and the row `positive_row` > 0), the listed edge rows of axis a finite on
// expand(domain, a, -1).  flags[0] |= 1: bad cell metrics, flags[1] |= 1: bad edge metrics.
struct LRows {
  int n_cell, cell[32], positive_row, n_edge, edge[32], edge_rows_per_axis, dim;
};
__global__ void k_legacy_check_valid(LShape S, LBox dom, LRows R, const double* __restrict__ cell,
                                     const double* __restrict__ edge, long long nc, int* flags) {
  bool bad_c = false, bad_e = false;
  LEGACY_FOR_BOX(dom) {
    LEGACY_INDEX(dom, S)
    for (int r = 0; r < R.n_cell; ++r) bad_c |= !isfinite(cell[(long long)R.cell[r] * nc + I]);
    if (R.positive_row >= 0) bad_c |= !(cell[(long long)R.positive_row * nc + I] > 0.0);
    for (int a = 0; a < R.dim; ++a) {
      if (c_[a] <= dom.lo[a] || c_[a] >= dom.hi[a] - 1) continue;  // expand(domain, a, -1)
      const double* E = edge + (long long)a * R.edge_rows_per_axis * nc;
      for (int r = 0; r < R.n_edge; ++r) bad_e |= !isfinite(E[(long long)R.edge[r] * nc + I]);
    }
  }
  if (__any_sync(0xffffffffu, bad_c) && (threadIdx.x & 31) == 0) atomicOr(flags, 1);
  if (__any_sync(0xffffffffu, bad_e) && (threadIdx.x & 31) == 0) atomicOr(flags + 1, 1);
}
}  // namespace

// see include/curvigrid_cuda.h
extern "C" int cg_legacy_meg6_update_1d(int device, const double* x, const int64_t* nnode, int nhalo, int halo_coords_included,
                                        double* cell, double* edge, double* centroid, double* scratch,
                                        size_t scratch_bytes, void* stream, uint64_t* launches) {
  struct R { R() { nvtxRangePushA("cg_legacy_meg6_update_1d"); } ~R() { nvtxRangePop(); } } nvtx_range_;
  g_legacy_err[0] = 0;
  if (!x || !nnode || !cell || !edge || !scratch) {
    snprintf(g_legacy_err, sizeof g_legacy_err, "cg_legacy_meg6_update_1d: NULL argument");
    return CG_ERR_ARG;
  }
  if (nhalo < 5) {
    snprintf(g_legacy_err, sizeof g_legacy_err, "MEG6 needs nhalo >= 5 (DiscretizationSchemes.jl:53-60)");
    return CG_ERR_ARG;
  }
  const int64_t nf = (halo_coords_included ? nnode[0] : nnode[0] + 2 * nhalo) - 1;
  if (nf - 2 * nhalo < 6) {
    snprintf(g_legacy_err, sizeof g_legacy_err, "MEG6 one-sided edge stencils need >= 6 interior cells per axis");
    return CG_ERR_ARG;
  }
  LShape S;
  LBox full, dom;
  S.n[0] = (int)nf; S.n[1] = S.n[2] = 1;
  S.st[0] = 1; S.st[1] = S.st[2] = nf;
  for (int d = 0; d < 3; ++d) {
    full.lo[d] = dom.lo[d] = 0;
    full.hi[d] = dom.hi[d] = d == 0 ? (int)nf : 1;
  }
  dom.lo[0] = nhalo;
  dom.hi[0] = (int)nf - nhalo;
  const long long nc = nf;
  if (scratch_bytes < (size_t)(3 * nc) * sizeof(double)) {
    snprintf(g_legacy_err, sizeof g_legacy_err, "scratch too small: need %lld bytes", (long long)(3 * nc * 8));
    return CG_ERR_ARG;
  }
  cudaError_t ce = cudaSetDevice(device);
  if (ce != cudaSuccess) {
    snprintf(g_legacy_err, sizeof g_legacy_err, "cudaSetDevice: %s", cudaGetErrorString(ce));
    return CG_ERR_CUDA;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const LBox md = halo_coords_included ? full : dom;
  const int whole = halo_coords_included ? 1 : 0;
  uint64_t nl = 0;
  double *cx = scratch, *d1 = scratch + nc, *d2 = scratch + 2 * nc;
  double *J = cell, *xxi = cell + nc, *inv = cell + 2 * nc, *hat = cell + 3 * nc;
  const int bl = blocks_for(box_count(md));
  if (legacy_fused()) {
    const LTmaCtx tma = {{cell, scratch}, {4, 1}, nc};
    zero_outside(S, full, md, cell, 4, nc, st, nl);
    zero_outside(S, full, md, scratch, 1, nc, st, nl);
    k_legacy_centroid1d<<<bl, 256, 0, st>>>(S, md, x, halo_coords_included ? 0 : nhalo, cx);
    ++nl;
    if (centroid) cudaMemcpyAsync(centroid, cx, (size_t)nc * sizeof(double), cudaMemcpyDeviceToDevice, st);
    LJobs jobs{};
    jobs.j[0] = {cx, nullptr, nullptr, nullptr, xxi, 0, 0, 0};  // forward_metrics.jl:11-27
    nl += launch_fused<0, OP_CCD>(S, md, 0, whole, jobs, 1, 1.0, st, &tma);
    k_legacy_inv1d<<<bl, 256, 0, st>>>(S, md, xxi, J, inv, hat);
    ++nl;
    const double* phi[3] = {J, inv, hat};  // MetricDiscretizationSchemes.jl:58-89
    interp_axis_fused(S, full, md, 0, whole, phi, 3, edge, nc, st, nl, &tma);
  } else {
  cudaMemsetAsync(cell, 0, (size_t)4 * nc * sizeof(double), st);
  cudaMemsetAsync(edge, 0, (size_t)3 * nc * sizeof(double), st);
  cudaMemsetAsync(scratch, 0, (size_t)3 * nc * sizeof(double), st);
  // 1d.jl:388-408: the centroids are refreshed on cell.domain only (unlike 2-D / 3-D); with halo_coords_included the
  // node array covers the halo and the metric domain is cell.full
  k_legacy_centroid1d<<<bl, 256, 0, st>>>(S, md, x, halo_coords_included ? 0 : nhalo, cx);
  ++nl;
  if (centroid) cudaMemcpyAsync(centroid, cx, (size_t)nc * sizeof(double), cudaMemcpyDeviceToDevice, st);
  k_legacy_first<<<bl, 256, 0, st>>>(S, md, 0, cx, d1);      // forward_metrics.jl:11-27
  k_legacy_second<<<bl, 256, 0, st>>>(S, md, 0, cx, d1, d2);
  k_legacy_meg<0><<<bl, 256, 0, st>>>(S, md, 0, cx, d1, d2, xxi, 1.0);
  k_legacy_inv1d<<<bl, 256, 0, st>>>(S, md, xxi, J, inv, hat);
  nl += 4;
  for (int m = 0; m < 3; ++m) {  // MetricDiscretizationSchemes.jl:58-89: J, xi.x1, xi^.x1 to the i+1/2 edges
    const double* phi = m == 0 ? J : (m == 1 ? inv : hat);
    k_legacy_first<<<bl, 256, 0, st>>>(S, md, 0, phi, d1);
    k_legacy_second<<<bl, 256, 0, st>>>(S, md, 0, phi, d1, d2);
    k_legacy_interp<<<bl, 256, 0, st>>>(S, md, 0, whole, phi, d1, d2, edge + (long long)m * nc);
    nl += 3;
  }
  }
  ce = cudaGetLastError();
  if (ce != cudaSuccess) {
    snprintf(g_legacy_err, sizeof g_legacy_err, "legacy kernels: %s", cudaGetErrorString(ce));
    return CG_ERR_CUDA;
  }
  if (launches) *launches = nl;
  return CG_OK;
}

extern "C" int cg_legacy_check_valid(int device, int dim, const int64_t* nfull, int nhalo, const double* cell,
                                     const int* cell_rows, int n_cell_rows, int positive_row, const double* edge,
                                     int edge_rows_per_axis, const int* edge_rows, int n_edge_rows, int* cell_ok,
                                     int* edge_ok, void* stream) {
  g_legacy_err[0] = 0;
  if (dim < 1 || dim > 3 || !nfull || !cell || !cell_ok || !edge_ok || n_cell_rows < 0 || n_cell_rows > 32 ||
      n_edge_rows < 0 || n_edge_rows > 32 || (n_cell_rows > 0 && !cell_rows) || (n_edge_rows > 0 && (!edge_rows || !edge))) {
    snprintf(g_legacy_err, sizeof g_legacy_err, "cg_legacy_check_valid: bad argument");
    return CG_ERR_ARG;
  }
  cudaError_t ce = cudaSetDevice(device);
  if (ce != cudaSuccess) return CG_ERR_CUDA;
  LShape S;
  LBox dom;
  long long nc = 1;
  for (int d = 0; d < 3; ++d) {
    S.n[d] = d < dim ? (int)nfull[d] : 1;
    S.st[d] = nc;
    nc *= S.n[d];
    dom.lo[d] = d < dim ? nhalo : 0;
    dom.hi[d] = d < dim ? (int)nfull[d] - nhalo : 1;
  }
  LRows R;
  R.n_cell = n_cell_rows;
  R.n_edge = n_edge_rows;
  R.positive_row = positive_row;
  R.edge_rows_per_axis = edge_rows_per_axis;
  R.dim = dim;
  for (int r = 0; r < n_cell_rows; ++r) R.cell[r] = cell_rows[r];
  for (int r = 0; r < n_edge_rows; ++r) R.edge[r] = edge_rows[r];
  cudaStream_t st = (cudaStream_t)stream;
  if (device < 0 || device >= 64) return CG_ERR_ARG;
  int* flags = legacy_flag_words(device, 0);
  if (!flags) return CG_ERR_OOM;
  cudaMemsetAsync(flags, 0, 2 * sizeof(int), st);
  k_legacy_check_valid<<<blocks_for(box_count(dom)), 256, 0, st>>>(S, dom, R, cell, edge, nc, flags);
  int h[2] = {0, 0};
  cudaMemcpyAsync(h, flags, sizeof h, cudaMemcpyDeviceToHost, st);
  ce = cudaStreamSynchronize(st);
  if (ce != cudaSuccess) {
    snprintf(g_legacy_err, sizeof g_legacy_err, "cg_legacy_check_valid: %s", cudaGetErrorString(ce));
    return CG_ERR_CUDA;
  }
  *cell_ok = !h[0];
  *edge_ok = !h[1];
  return CG_OK;
}

// ORACLE — TEST INFRASTRUCTURE ONLY.  Never linked into the product library.
//
// CPU restatement of the reference's LEGACY metric pipeline for
// CurvilinearGrid3D(x, y, z, :meg6 | :meg6_symmetric): centroids, forward metrics,
// Thomas-Lombard / Nonomura conservative metrics and the 66 edge interpolations
// (SURVEY.md §3.4, §8 rows a17-a21, Appendix A.5).  It is an array-pass-by-array-pass
// restatement: same passes, same domains, same operation order (Kahan sums,
// tol_diff, eps clips) as the reference, compiled with -ffp-contract=off because the
// thresholds are discontinuous.
//
// parity: PARTIALLY PINNED — no Julia runtime here; pinned by the reference's own
// legacy tests re-encoded in tests/test_oracle_pins.py (linear fields -> exact
// derivatives, test_discretizations.jl:41-146; noise suppression :193-221; rectilinear
// 3-D known answers, test_3d.jl:137-147).
//
// Reference files followed (under /root/reference/src):
//   discretization_schemes/derivative_stencils.jl:1-193      stencils
//   discretization_schemes/derivative_kernels.jl:16-314      array kernels
//   discretization_schemes/DiscretizationSchemes.jl:63-159   domains + drivers
//   discretization_schemes/meg/MonotoneExplicitGradients.jl:108-122   L/R states
//   discretization_schemes/meg/cell_center_derivs.jl:9-122   MEG cell-centre derivative
//   discretization_schemes/meg/interp_to_edges.jl:15-79      edge interpolation
//   metric_schemes/forward_metrics.jl:89-167                 forward metrics + det
//   metric_schemes/conservative_metrics.jl:13-175, 259-427   TL / symmetric metrics
//   metric_schemes/MetricDiscretizationSchemes.jl:20-89      drivers
//   grids/curvilinear_mapped_grids/3d.jl:472-598             update!, centroids
#pragma once
#include <cmath>
#include <cstdint>
#include <limits>
#include <vector>

namespace cgl {

constexpr double EPS = std::numeric_limits<double>::epsilon();

struct Box {  // half-open cell-index box in the full array
  int64_t lo[3], hi[3];
};
struct Shape {
  int64_t n[3];
  int64_t st[3];
  int64_t total() const { return n[0] * n[1] * n[2]; }
};

// derivative_stencils.jl:1-11
template <int N>
inline double kahan_sum(const double (&v)[N]) {
  double s = 0.0, c = 0.0;
  for (int i = 0; i < N; ++i) {
    double y = v[i] - c;
    double tmp = s + y;
    c = (tmp - s) - y;
    s = tmp;
  }
  return s + c;
}
// Base.isapprox(a, b; rtol, atol)
inline bool isapprox(double a, double b, double rtol, double atol) {
  if (a == b) return true;
  if (!std::isfinite(a) || !std::isfinite(b)) return false;
  return std::fabs(a - b) <= std::fmax(atol, rtol * std::fmax(std::fabs(a), std::fabs(b)));
}
// derivative_stencils.jl:13-16
inline double tol_diff(double a, double b) {
  double d = a - b;
  return isapprox(a, b, std::sqrt(EPS), 10 * EPS) ? d * 0.0 : d;
}
// derivative_stencils.jl:34-45 (6th-order central)
inline double central6(double m3, double m2, double m1, double p1, double p2, double p3) {
  const double v[3] = {(3.0 / 4.0) * tol_diff(p1, m1), (-3.0 / 20.0) * tol_diff(p2, m2), (1.0 / 60.0) * tol_diff(p3, m3)};
  double d = kahan_sum(v);
  return std::fabs(d) >= EPS ? d : d * 0.0;
}
// derivative_stencils.jl:97-135
inline void mixed_lo(const double* p, double* d) {  // p[0..5] -> d[0..2]
  const double e = 50 * EPS;
  const double a[6] = {(-137.0 / 60.0) * p[0], 5 * p[1], -5 * p[2], (10.0 / 3.0) * p[3], -(5.0 / 4.0) * p[4], (1.0 / 5.0) * p[5]};
  const double b[6] = {(-1.0 / 5.0) * p[0], -(13.0 / 12.0) * p[1], 2 * p[2], -p[3], (1.0 / 3.0) * p[4], -(1.0 / 20.0) * p[5]};
  const double c[6] = {(1.0 / 20.0) * p[0], -(1.0 / 2.0) * p[1], -(1.0 / 3.0) * p[2], p[3], -(1.0 / 4.0) * p[4], (1.0 / 30.0) * p[5]};
  double d1 = kahan_sum(a), d2 = kahan_sum(b), d3 = kahan_sum(c);
  d[0] = std::fabs(d1) >= e ? d1 : d1 * 0.0;
  d[1] = std::fabs(d2) >= e ? d2 : d2 * 0.0;
  d[2] = std::fabs(d3) >= e ? d3 : d3 * 0.0;
}
// derivative_stencils.jl:137-175; returns (d4, d5, d6) = derivatives at p[3], p[4], p[5]
inline void mixed_hi(const double* p, double* d) {
  const double e = 50 * EPS;
  const double a6[6] = {(-1.0 / 5.0) * p[0], (5.0 / 4.0) * p[1], -(10.0 / 3.0) * p[2], 5 * p[3], -5 * p[4], (137.0 / 60.0) * p[5]};
  const double a5[6] = {(1.0 / 20.0) * p[0], -(1.0 / 3.0) * p[1], p[2], -2 * p[3], (13.0 / 12.0) * p[4], (1.0 / 5.0) * p[5]};
  const double a4[6] = {(-1.0 / 30.0) * p[0], (1.0 / 4.0) * p[1], -p[2], (1.0 / 3.0) * p[3], (1.0 / 2.0) * p[4], -(1.0 / 20.0) * p[5]};
  double d6 = kahan_sum(a6), d5 = kahan_sum(a5), d4 = kahan_sum(a4);
  d[0] = std::fabs(d4) >= e ? d4 : d4 * 0.0;
  d[1] = std::fabs(d5) >= e ? d5 : d5 * 0.0;
  d[2] = std::fabs(d6) >= e ? d6 : d6 * 0.0;
}
// derivative_stencils.jl:180-193
inline double second_derivative(double m1, double c, double p1, double dm1, double dp1) {
  double dd = -tol_diff(dp1, dm1) / 2;
  const double v[3] = {2 * (p1 + m1), -4 * c, dd};
  return kahan_sum(v);
}
// MonotoneExplicitGradients.jl:108-122
inline double phiL_ip(double f, double d, double dd) { return f + ((1.0 / 2.0) * d + (1.0 / 12.0) * dd); }
inline double phiR_ip(double f, double d, double dd) { return f - ((1.0 / 2.0) * d + (1.0 / 12.0) * dd); }  // also phiR_{i-1/2}

struct Grid {
  Shape S;
  inline int64_t at(int64_t i, int64_t j, int64_t k) const { return i + S.n[0] * (j + S.n[1] * k); }
  template <class F>
  void each(const Box& b, F f) const {
#pragma omp parallel for collapse(2) schedule(static)
    for (int64_t k = b.lo[2]; k < b.hi[2]; ++k)
      for (int64_t j = b.lo[1]; j < b.hi[1]; ++j)
        for (int64_t i = b.lo[0]; i < b.hi[0]; ++i) f(at(i, j, k));
  }
};
inline Box expand(Box b, int axis, int64_t by) {
  b.lo[axis] -= by;
  b.hi[axis] += by;
  return b;
}
inline Box lower_boundary(Box b, int axis, int64_t off) {  // first index along axis (+off)
  b.lo[axis] += off;
  b.hi[axis] = b.lo[axis] + 1;
  return b;
}
inline Box upper_boundary(Box b, int axis, int64_t off) {  // last index along axis (+off)
  b.hi[axis] += off;
  b.lo[axis] = b.hi[axis] - 1;
  return b;
}

// DiscretizationSchemes.jl:90-120 (MEG6, use_one_sided_on_edges = true)
inline void first_derivatives(const Grid& g, double* dphi, const double* phi, int axis, const Box& dom) {
  const int64_t s = g.S.st[axis];
  g.each(expand(dom, axis, -3), [&](int64_t I) {
    dphi[I] = central6(phi[I - 3 * s], phi[I - 2 * s], phi[I - s], phi[I + s], phi[I + 2 * s], phi[I + 3 * s]);
  });
  g.each(lower_boundary(dom, axis, 0), [&](int64_t I) {
    double p[6], d[3];
    for (int m = 0; m < 6; ++m) p[m] = phi[I + m * s];
    mixed_lo(p, d);
    dphi[I] = d[0];
    dphi[I + s] = d[1];
    dphi[I + 2 * s] = d[2];
  });
  g.each(upper_boundary(dom, axis, 0), [&](int64_t I) {
    double p[6], d[3];
    for (int m = 0; m < 6; ++m) p[m] = phi[I - (5 - m) * s];
    mixed_hi(p, d);  // d = derivatives at I-2, I-1, I
    dphi[I] = d[2];
    dphi[I - s] = d[1];
    dphi[I - 2 * s] = d[0];
  });
}
// DiscretizationSchemes.jl:129-159
inline void second_derivatives(const Grid& g, double* d2, const double* dphi, const double* phi, int axis, const Box& dom) {
  const int64_t s = g.S.st[axis];
  g.each(expand(dom, axis, -3), [&](int64_t I) {
    d2[I] = second_derivative(phi[I - s], phi[I], phi[I + s], dphi[I - s], dphi[I + s]);
  });
  g.each(lower_boundary(dom, axis, 0), [&](int64_t I) {
    double p[6], d[3];
    for (int m = 0; m < 6; ++m) p[m] = dphi[I + m * s];
    mixed_lo(p, d);
    d2[I] = d[0];
    d2[I + s] = d[1];
    d2[I + 2 * s] = d[2];
  });
  g.each(upper_boundary(dom, axis, 0), [&](int64_t I) {
    double p[6], d[3];
    for (int m = 0; m < 6; ++m) p[m] = dphi[I - (5 - m) * s];
    mixed_hi(p, d);
    d2[I] = d[2];
    d2[I - s] = d[1];
    d2[I - 2 * s] = d[0];
  });
}
// cell_center_derivs.jl:24-122 (compute_gradients = true, use_one_sided_on_edges = true)
inline void cell_center_derivatives(const Grid& g, double* out, double* d2, double* d1, const double* phi, int axis,
                                    const Box& dom) {
  const double e = 50 * EPS;
  const int64_t s = g.S.st[axis];
  first_derivatives(g, d1, phi, axis, dom);
  second_derivatives(g, d2, d1, phi, axis, dom);
  g.each(expand(dom, axis, -1), [&](int64_t I) {
    double hi = (phiL_ip(phi[I], d1[I], d2[I]) + phiR_ip(phi[I + s], d1[I + s], d2[I + s])) / 2;
    double lo = (phiL_ip(phi[I - s], d1[I - s], d2[I - s]) + phiR_ip(phi[I], d1[I], d2[I])) / 2;
    double v = hi - lo;
    out[I] = std::fabs(v) >= e ? v : v * 0.0;
  });
  g.each(lower_boundary(dom, axis, 0), [&](int64_t I) {
    double hi = (phiL_ip(phi[I], d1[I], d2[I]) + phiR_ip(phi[I + s], d1[I + s], d2[I + s])) / 2;
    double lo = phiR_ip(phi[I], d1[I], d2[I]);
    double v = hi - lo;
    out[I] = std::fabs(v) >= e ? v : v * 0.0;
  });
  g.each(upper_boundary(dom, axis, 0), [&](int64_t I) {
    double hi = phiL_ip(phi[I], d1[I], d2[I]);
    double lo = (phiL_ip(phi[I - s], d1[I - s], d2[I - s]) + phiR_ip(phi[I], d1[I], d2[I])) / 2;
    double v = hi - lo;
    out[I] = std::fabs(v) >= e ? v : v * 0.0;
  });
}
// interp_to_edges.jl:15-79
inline void interpolate_to_edge(const Grid& g, double* edge, double* d2, double* d1, const double* phi, int axis,
                                const Box& dom) {
  const int64_t s = g.S.st[axis];
  first_derivatives(g, d1, phi, axis, dom);
  second_derivatives(g, d2, d1, phi, axis, dom);
  bool whole = true;
  for (int d = 0; d < 3; ++d) whole = whole && (dom.hi[d] - dom.lo[d] == g.S.n[d]);
  Box inner = whole ? expand(dom, axis, -1) : dom;
  Box lo = whole ? lower_boundary(dom, axis, +1) : lower_boundary(dom, axis, 0);
  Box hi = whole ? upper_boundary(dom, axis, -1) : upper_boundary(dom, axis, 0);
  g.each(inner, [&](int64_t I) {
    edge[I] = (phiL_ip(phi[I], d1[I], d2[I]) + phiR_ip(phi[I + s], d1[I + s], d2[I + s])) / 2;
  });
  g.each(lo, [&](int64_t I) { edge[I - s] = phiR_ip(phi[I], d1[I], d2[I]); });
  g.each(hi, [&](int64_t I) { edge[I] = phiL_ip(phi[I], d1[I], d2[I]); });
}

// Output block of the legacy SoA (metric_soa.jl:2-170), the never-written `.t` arrays omitted:
//   cell [28][ncell]:  0 J | 1..9 fwd[q*3+axis] = d x_q / d xi_axis | 10..18 inv[row*3+c] | 19..27 hat[row*3+c]
//   edge [3][19][ncell]: 0 J | 1..9 inv[row*3+c] | 10..18 hat[row*3+c]   (storage index I = face I + 1/2)
enum { L_J = 0, L_FWD = 1, L_INV = 10, L_HAT = 19, L_NCELL = 28, LE_J = 0, LE_INV = 1, LE_HAT = 10, L_NEDGE = 19 };

inline void legacy_meg6_3d(const double* xn, const double* yn, const double* zn, const int64_t* nnode, int nhalo,
                           bool halo_coords_included, bool symmetric, double* cell, double* edge, double* cent) {
  // node arrays are interior (nnode) or node.full; cell.full = nodes(full) - 1
  Grid g;
  int64_t nfull_nodes[3];
  for (int d = 0; d < 3; ++d) {
    nfull_nodes[d] = halo_coords_included ? nnode[d] : nnode[d] + 2 * nhalo;
    g.S.n[d] = nfull_nodes[d] - 1;
  }
  g.S.st[0] = 1;
  g.S.st[1] = g.S.n[0];
  g.S.st[2] = g.S.n[0] * g.S.n[1];
  const int64_t nc = g.S.total();
  Box full, dom;
  for (int d = 0; d < 3; ++d) {
    full.lo[d] = 0;
    full.hi[d] = g.S.n[d];
    dom.lo[d] = nhalo;
    dom.hi[d] = g.S.n[d] - nhalo;
  }
  const Box md = halo_coords_included ? full : dom;  // metric_domain (3d.jl:472-478)
  // node coordinates in node.full indexing (3d.jl:430-447)
  const int64_t off = halo_coords_included ? 0 : nhalo;
  const int64_t nn0 = nnode[0], nn1 = nnode[1];
  auto node = [&](const double* a, int64_t i, int64_t j, int64_t k) {
    return a[(i - off) + nn0 * ((j - off) + nn1 * (k - off))];
  };
  std::vector<double> cx(nc, 0.0), cy(nc, 0.0), cz(nc, 0.0);
  const double* N[3] = {xn, yn, zn};
  double* C[3] = {cx.data(), cy.data(), cz.data()};
  for (int q = 0; q < 3; ++q) {
    const double* a = N[q];
    double* c = C[q];
    // 3d.jl:555-598
    for (int64_t k = md.lo[2]; k < md.hi[2]; ++k)
      for (int64_t j = md.lo[1]; j < md.hi[1]; ++j)
        for (int64_t i = md.lo[0]; i < md.hi[0]; ++i)
          c[g.at(i, j, k)] = 0.125 * (node(a, i, j, k) + node(a, i + 1, j, k) + node(a, i + 1, j + 1, k) + node(a, i, j + 1, k) +
                                      node(a, i, j, k + 1) + node(a, i + 1, j, k + 1) + node(a, i + 1, j + 1, k + 1) +
                                      node(a, i, j + 1, k + 1));
    if (cent)
      for (int64_t I = 0; I < nc; ++I) cent[q * nc + I] = c[I];
  }
  for (int64_t I = 0; I < (int64_t)L_NCELL * nc; ++I) cell[I] = 0.0;
  for (int64_t I = 0; I < (int64_t)3 * L_NEDGE * nc; ++I) edge[I] = 0.0;
  std::vector<double> d1(nc, 0.0), d2(nc, 0.0), in1(nc, 0.0), in2(nc, 0.0), o1(nc, 0.0), o2(nc, 0.0);
  auto fwd = [&](int q, int ax) { return cell + (int64_t)(L_FWD + q * 3 + ax) * nc; };
  auto inv = [&](int row, int c) { return cell + (int64_t)(L_INV + row * 3 + c) * nc; };
  auto hat = [&](int row, int c) { return cell + (int64_t)(L_HAT + row * 3 + c) * nc; };
  double* J = cell;
  // forward_metrics.jl:120-152
  for (int q = 0; q < 3; ++q)
    for (int ax = 0; ax < 3; ++ax) cell_center_derivatives(g, fwd(q, ax), d2.data(), d1.data(), C[q], ax, md);
  g.each(md, [&](int64_t I) {
    // StaticArrays det of [xξ xη xζ; yξ yη yζ; zξ zη zζ]: x0 . (x1 × x2) over columns
    const double a0 = fwd(0, 0)[I], a1 = fwd(1, 0)[I], a2 = fwd(2, 0)[I];
    const double b0 = fwd(0, 1)[I], b1 = fwd(1, 1)[I], b2 = fwd(2, 1)[I];
    const double c0 = fwd(0, 2)[I], c1 = fwd(1, 2)[I], c2 = fwd(2, 2)[I];
    J[I] = a0 * (b1 * c2 - b2 * c1) + a1 * (b2 * c0 - b0 * c2) + a2 * (b0 * c1 - b1 * c0);
  });
  // conservative_metrics.jl:76-175 / 328-427.  row r of column c uses (q1,q2) = (c+1, c+2) mod 3 and the axis
  // pair (outer1, outer2) = ((r+2)%3, (r+1)%3):  hat = (q1_{s} q2)_{u} - (q1_{u} q2)_{s},  s=(r+1)%3, u=(r+2)%3
  const double e = 50 * EPS;
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) {
      const int q1 = (c + 1) % 3, q2 = (c + 2) % 3, s = (r + 1) % 3, u = (r + 2) % 3;
      const double *a_s = fwd(q1, s), *a_u = fwd(q1, u), *b_s = fwd(q2, s), *b_u = fwd(q2, u);
      const double *Q1 = C[q1], *Q2 = C[q2];
      for (int64_t I = 0; I < nc; ++I) {
        if (symmetric) {
          in1[I] = a_s[I] * Q2[I] - b_s[I] * Q1[I];
          in2[I] = a_u[I] * Q2[I] - b_u[I] * Q1[I];
        } else {
          in1[I] = a_s[I] * Q2[I];
          in2[I] = a_u[I] * Q2[I];
        }
      }
      cell_center_derivatives(g, o1.data(), d2.data(), d1.data(), in1.data(), u, md);
      cell_center_derivatives(g, o2.data(), d2.data(), d1.data(), in2.data(), s, md);
      double* h = hat(r, c);
      g.each(md, [&](int64_t I) {
        double v = symmetric ? (1.0 / 2.0) * (o1[I] - o2[I]) : (o1[I] - o2[I]);
        h[I] = std::fabs(v) >= e ? v : v * 0.0;
      });
    }
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) {
      double* x = inv(r, c);
      const double* h = hat(r, c);
      g.each(md, [&](int64_t I) { x[I] = h[I] / J[I]; });
    }
  // MetricDiscretizationSchemes.jl:58-89
  for (int ax = 0; ax < 3; ++ax) {
    double* E = edge + (int64_t)ax * L_NEDGE * nc;
    interpolate_to_edge(g, E + (int64_t)LE_J * nc, d2.data(), d1.data(), J, ax, md);
    for (int m = 0; m < 9; ++m) {
      interpolate_to_edge(g, E + (int64_t)(LE_INV + m) * nc, d2.data(), d1.data(), cell + (int64_t)(L_INV + m) * nc, ax, md);
      interpolate_to_edge(g, E + (int64_t)(LE_HAT + m) * nc, d2.data(), d1.data(), cell + (int64_t)(L_HAT + m) * nc, ax, md);
    }
  }
}


// ---------------------------------------------------------------------------
// 2-D legacy grid: CurvilinearGrid2D(x, y, :meg6 | :meg6_symmetric) `update!` (2d.jl:616-637, centroids :734-746):
// forward_metrics! (forward_metrics.jl:36-84), conservative_metrics! 2-D (conservative_metrics.jl:177-236; the
// symmetric variant forwards to it, :417-421) and update_edge_metrics! (MetricDiscretizationSchemes.jl:58-89).
// The arrays are handled as (ni, nj, 1) boxes by the same passes as the 3-D pipeline.
//   cell [13][ncell]:   0 J | 1..4 fwd[q*2+axis] | 5..8 inv (xi_x, xi_y, eta_x, eta_y) | 9..12 hat = inv * J
//   edge [2][9][ncell]: 0 J | 1..4 inv | 5..8 hat      (storage index I = face I + 1/2; `.t` arrays omitted)
// ---------------------------------------------------------------------------
enum { L2_J = 0, L2_FWD = 1, L2_INV = 5, L2_HAT = 9, L2_NCELL = 13, L2_NEDGE = 9 };

inline void legacy_meg6_2d(const double* xn, const double* yn, const int64_t* nnode, int nhalo, bool halo_coords_included,
                           double* cell, double* edge, double* cent) {
  Grid g;
  for (int d = 0; d < 2; ++d) g.S.n[d] = (halo_coords_included ? nnode[d] : nnode[d] + 2 * nhalo) - 1;
  g.S.n[2] = 1;
  g.S.st[0] = 1;
  g.S.st[1] = g.S.n[0];
  g.S.st[2] = g.S.n[0] * g.S.n[1];
  const int64_t nc = g.S.total();
  Box full, dom;
  for (int d = 0; d < 2; ++d) {
    full.lo[d] = 0;
    full.hi[d] = g.S.n[d];
    dom.lo[d] = nhalo;
    dom.hi[d] = g.S.n[d] - nhalo;
  }
  full.lo[2] = dom.lo[2] = 0;
  full.hi[2] = dom.hi[2] = 1;
  const Box md = halo_coords_included ? full : dom;
  const int64_t off = halo_coords_included ? 0 : nhalo;
  const int64_t nn0 = nnode[0];
  auto node = [&](const double* a, int64_t i, int64_t j) { return a[(i - off) + nn0 * (j - off)]; };
  std::vector<double> cx(nc, 0.0), cy(nc, 0.0);
  const double* N[2] = {xn, yn};
  double* C[2] = {cx.data(), cy.data()};
  for (int q = 0; q < 2; ++q) {
    const double* a = N[q];
    double* c = C[q];
    for (int64_t j = md.lo[1]; j < md.hi[1]; ++j)  // 2d.jl:734-746
      for (int64_t i = md.lo[0]; i < md.hi[0]; ++i)
        c[g.at(i, j, 0)] = 0.25 * (node(a, i, j) + node(a, i + 1, j) + node(a, i + 1, j + 1) + node(a, i, j + 1));
    if (cent)
      for (int64_t I = 0; I < nc; ++I) cent[q * nc + I] = c[I];
  }
  for (int64_t I = 0; I < (int64_t)L2_NCELL * nc; ++I) cell[I] = 0.0;
  for (int64_t I = 0; I < (int64_t)2 * L2_NEDGE * nc; ++I) edge[I] = 0.0;
  std::vector<double> d1(nc, 0.0), d2(nc, 0.0);
  auto fwd = [&](int q, int ax) { return cell + (int64_t)(L2_FWD + q * 2 + ax) * nc; };
  double* J = cell;
  for (int q = 0; q < 2; ++q)
    for (int ax = 0; ax < 2; ++ax) cell_center_derivatives(g, fwd(q, ax), d2.data(), d1.data(), C[q], ax, md);
  double* inv = cell + (int64_t)L2_INV * nc;
  double* hat = cell + (int64_t)L2_HAT * nc;
  g.each(md, [&](int64_t I) {
    // StaticArrays det / inv of [x_xi x_eta; y_xi y_eta] (inverse_metric_2d_kernel!, conservative_metrics.jl:218-236)
    const double a11 = fwd(0, 0)[I], a12 = fwd(0, 1)[I], a21 = fwd(1, 0)[I], a22 = fwd(1, 1)[I];
    const double d = a11 * a22 - a12 * a21;
    J[I] = d;
    inv[0 * nc + I] = a22 / d;      // xi_x
    inv[1 * nc + I] = -(a12 / d);   // xi_y
    inv[2 * nc + I] = -(a21 / d);   // eta_x
    inv[3 * nc + I] = a11 / d;      // eta_y
    for (int m = 0; m < 4; ++m) hat[m * nc + I] = inv[m * nc + I] * d;
  });
  for (int ax = 0; ax < 2; ++ax) {
    double* E = edge + (int64_t)ax * L2_NEDGE * nc;
    interpolate_to_edge(g, E, d2.data(), d1.data(), J, ax, md);
    for (int m = 0; m < 8; ++m)
      interpolate_to_edge(g, E + (int64_t)(1 + m) * nc, d2.data(), d1.data(), cell + (int64_t)(L2_INV + m) * nc, ax, md);
  }
}


// ---------------------------------------------------------------------------
// Legacy CurvilinearGrid1D(x, :meg6) update! (1d.jl:388-408; forward_metrics.jl:11-27; conservative_metrics.jl:239-245;
// MetricDiscretizationSchemes.jl:58-89).  cell = [4][nc] (J, x_xi, xi_x, hatted), edge = [3][nc] (J, xi_x, hatted at
// i+1/2), cent = [nc].
// ---------------------------------------------------------------------------
inline void legacy_meg6_1d(const double* xn, const int64_t* nnode, int nhalo, bool halo_coords_included, double* cell,
                           double* edge, double* cent) {
  Grid g;
  g.S.n[0] = (halo_coords_included ? nnode[0] : nnode[0] + 2 * nhalo) - 1;
  g.S.n[1] = g.S.n[2] = 1;
  g.S.st[0] = 1;
  g.S.st[1] = g.S.st[2] = g.S.n[0];
  const int64_t nc = g.S.total();
  Box full, dom;
  for (int d = 0; d < 3; ++d) {
    full.lo[d] = dom.lo[d] = 0;
    full.hi[d] = dom.hi[d] = d == 0 ? g.S.n[0] : 1;
  }
  dom.lo[0] = nhalo;
  dom.hi[0] = g.S.n[0] - nhalo;
  const Box md = halo_coords_included ? full : dom;
  const int64_t off = halo_coords_included ? 0 : nhalo;
  std::vector<double> cx(nc, 0.0), d1(nc, 0.0), d2(nc, 0.0);
  for (int64_t i = md.lo[0]; i < md.hi[0]; ++i) cx[i] = 0.5 * (xn[i - off] + xn[i - off + 1]);  // 1d.jl:445-455
  if (cent)
    for (int64_t I = 0; I < nc; ++I) cent[I] = cx[I];
  for (int64_t I = 0; I < 4 * nc; ++I) cell[I] = 0.0;
  for (int64_t I = 0; I < 3 * nc; ++I) edge[I] = 0.0;
  double *J = cell, *xxi = cell + nc, *inv = cell + 2 * nc, *hat = cell + 3 * nc;
  cell_center_derivatives(g, xxi, d2.data(), d1.data(), cx.data(), 0, md);
  g.each(md, [&](int64_t I) {
    inv[I] = 1.0 / xxi[I];
    J[I] = std::fabs(xxi[I]);
    hat[I] = inv[I] * J[I];
  });
  interpolate_to_edge(g, edge, d2.data(), d1.data(), J, 0, md);
  interpolate_to_edge(g, edge + nc, d2.data(), d1.data(), inv, 0, md);
  interpolate_to_edge(g, edge + 2 * nc, d2.data(), d1.data(), hat, 0, md);
}

}  // namespace cgl

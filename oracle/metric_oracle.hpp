// ORACLE — TEST INFRASTRUCTURE ONLY.  Never linked into the product library.
//
// CPU restatement of the unified-grid metric evaluation of
// smillerc/CurvilinearGrids.jl v0.13.3.  It follows the reference's *closure
// graph* literally: every metric is a closure of the mapping, and every
// derivative is taken by nested forward-mode AD at the point where the
// reference takes it.  Nothing here knows the closed forms the CUDA kernels
// use, so agreement between the two is an independent check.
//
// parity: PARTIALLY PINNED.  No Julia runtime exists in this environment and
// the reference ships no stored golden vectors; this file
// is pinned against the reference's own known-answer and property tests,
// re-encoded in tests/test_oracle_pins.py (rectilinear known answers incl.
// halo cells, GCL bounds, mapped-vs-discrete convergence).
//
// Reference files followed (paths under /root/reference/src/grids/unified_grids):
//   metrics.jl:141-178   _edge_reconstruct (three schemes)
//   metrics.jl:183-341   MetricCache 3-D / 2-D / 1-D
//   metrics.jl:428-510   edge_functions_{2d,1d}
//   metrics.jl:670-902   get_inverse_metric_terms (Thomas-Lombard functions)
//   metrics.jl:910-1113  ξ/η/ζ_derivs
//   metrics.jl:1120-1529 ϕ_{i,j,k}edge
//   metrics.jl:1536-1651 ∂ϕ_∂{ξ,η,ζ}_3d
//   metrics.jl:1716-1793 _inverse_and_normalized_edge_metrics
//   metrics.jl:1795-2053 cell / face fill kernels
//   discrete_grid.jl:55-169 interpolant mapping + gradient override
//   generation.jl:187-309 node / centroid / face-coordinate kernels
// Third-party arithmetic restated (not vendored in /root/reference):
//   Interpolations.jl 0.16.2  BSpline(Linear()) + extrapolate(..., Line()),
//                             Interpolations.gradient
//   ForwardDiff 1.3.3 / DifferentiationInterface 0.7.18 (see cg_dual.hpp)
//   StaticArrays 1.9.18       inv / det of 1..3-dim SMatrix
#pragma once
#include <array>
#include <cmath>
#include <cstdint>

#include "cg_dual.hpp"

namespace cgo {

template <class S>
using Pt = std::array<S, 3>;  // computational point (ξ,η,ζ); unused axes carry constants

template <class S, int K>
struct Vals {  // scalar (K=1) or column-major N×N matrix (K=N*N) closure result
  S e[K];
};

enum Scheme { ENDPOINT = 0, GRADIENT = 1, CURVATURE = 2 };

// ---------------------------------------------------------------------------
// DifferentiationInterface equivalents.  `f` is a generic callable
// Pt<Q> -> Vals<Q,K> for any (nested dual) number type Q.  All coordinates of a
// point share one number type; coordinates that are not differentiated are
// lifted with a zero derivative part, which is what ForwardDiff's tag
// promotion does when a Dual meets a plain number.
// ---------------------------------------------------------------------------
template <int K, class S, class F>
inline Vals<S, K> derivative(const F& f, int axis, const Pt<S>& p) {
  Pt<Dual<S>> q;
  for (int a = 0; a < 3; ++a) q[a] = {p[a], constant<S>(a == axis ? 1.0 : 0.0)};
  Vals<Dual<S>, K> r = f(q);
  Vals<S, K> out;
  for (int k = 0; k < K; ++k) out.e[k] = r.e[k].d;
  return out;
}

template <int K, class S, class F>
inline void value_and_derivative(const F& f, int axis, const Pt<S>& p, Vals<S, K>& v, Vals<S, K>& d) {
  Pt<Dual<S>> q;
  for (int a = 0; a < 3; ++a) q[a] = {p[a], constant<S>(a == axis ? 1.0 : 0.0)};
  Vals<Dual<S>, K> r = f(q);
  for (int k = 0; k < K; ++k) {
    v.e[k] = r.e[k].v;
    d.e[k] = r.e[k].d;
  }
}

template <int K, class S, class F>
inline void value_derivative_and_second_derivative(const F& f, int axis, const Pt<S>& p, Vals<S, K>& v,
                                                   Vals<S, K>& d, Vals<S, K>& dd) {
  using D1 = Dual<S>;
  using D2 = Dual<D1>;
  Pt<D2> q;
  for (int a = 0; a < 3; ++a) {
    const double s = (a == axis) ? 1.0 : 0.0;
    q[a] = D2{D1{p[a], constant<S>(s)}, D1{constant<S>(s), constant<S>(0.0)}};
  }
  Vals<D2, K> r = f(q);
  for (int k = 0; k < K; ++k) {
    v.e[k] = r.e[k].v.v;
    d.e[k] = r.e[k].v.d;
    dd.e[k] = r.e[k].d.d;
  }
}

// ---------------------------------------------------------------------------
// metrics.jl:141-178  _edge_reconstruct, Δξ = 1.  Note the +1/12 curvature
// weight on BOTH states (the legacy MEG path uses -1/12 on the right state).
// Face value between the points p and p+e_axis of a closure f.
// Also: edge_functions_{1,2,3}d (metrics.jl:343-510) and ϕ_{i,j,k}edge
// (metrics.jl:1120-1529) all reduce to this.
// ---------------------------------------------------------------------------
template <int SCH, int K, class S, class F>
inline Vals<S, K> edge_reconstruct(const F& f, int axis, const Pt<S>& p) {
  Pt<S> p1 = p;
  p1[axis] = p1[axis] + 1.0;
  Vals<S, K> out;
  if constexpr (SCH == ENDPOINT) {
    Vals<S, K> a = f(p), b = f(p1);
    for (int k = 0; k < K; ++k) out.e[k] = 0.5 * (a.e[k] + b.e[k]);
  } else if constexpr (SCH == GRADIENT) {
    Vals<S, K> a, da, b, db;
    value_and_derivative<K>(f, axis, p, a, da);
    value_and_derivative<K>(f, axis, p1, b, db);
    for (int k = 0; k < K; ++k) {
      S L = a.e[k] + 0.5 * da.e[k];
      S R = b.e[k] - 0.5 * db.e[k];
      out.e[k] = 0.5 * (L + R);
    }
  } else {
    Vals<S, K> a, da, dda, b, db, ddb;
    value_derivative_and_second_derivative<K>(f, axis, p, a, da, dda);
    value_derivative_and_second_derivative<K>(f, axis, p1, b, db, ddb);
    const double w = 1.0 / 12.0;
    for (int k = 0; k < K; ++k) {
      S L = a.e[k] + 0.5 * da.e[k] + w * dda.e[k];
      S R = b.e[k] - 0.5 * db.e[k] + w * ddb.e[k];
      out.e[k] = 0.5 * (L + R);
    }
  }
  return out;
}

// ---------------------------------------------------------------------------
// StaticArrays inv / det for column-major N×N (e[r + N*c] = M[r][c]).
// ---------------------------------------------------------------------------
template <class S>
inline S det2(const Vals<S, 4>& m) { return m.e[0] * m.e[3] - m.e[2] * m.e[1]; }
template <class S>
inline Vals<S, 4> inv2(const Vals<S, 4>& m) {
  S idet = 1.0 / det2(m);
  Vals<S, 4> o;
  o.e[0] = m.e[3] * idet;
  o.e[1] = -m.e[1] * idet;
  o.e[2] = -m.e[2] * idet;
  o.e[3] = m.e[0] * idet;
  return o;
}
template <class S>
inline S det3(const Vals<S, 9>& m) {
  // columns x0=(e0,e1,e2) x1=(e3,e4,e5) x2=(e6,e7,e8); det = x0 . (x1 × x2)
  S c0 = m.e[4] * m.e[8] - m.e[5] * m.e[7];
  S c1 = m.e[5] * m.e[6] - m.e[3] * m.e[8];
  S c2 = m.e[3] * m.e[7] - m.e[4] * m.e[6];
  return m.e[0] * c0 + m.e[1] * c1 + m.e[2] * c2;
}
template <class S>
inline Vals<S, 9> inv3(const Vals<S, 9>& m) {
  // rows of the inverse are (x1×x2, x2×x0, x0×x1)/det
  const S *x0 = &m.e[0], *x1 = &m.e[3], *x2 = &m.e[6];
  S y0[3] = {x1[1] * x2[2] - x1[2] * x2[1], x1[2] * x2[0] - x1[0] * x2[2], x1[0] * x2[1] - x1[1] * x2[0]};
  S y1[3] = {x2[1] * x0[2] - x2[2] * x0[1], x2[2] * x0[0] - x2[0] * x0[2], x2[0] * x0[1] - x2[1] * x0[0]};
  S y2[3] = {x0[1] * x1[2] - x0[2] * x1[1], x0[2] * x1[0] - x0[0] * x1[2], x0[0] * x1[1] - x0[1] * x1[0]};
  S idet = 1.0 / (x0[0] * y0[0] + x0[1] * y0[1] + x0[2] * y0[2]);
  Vals<S, 9> o;
  for (int c = 0; c < 3; ++c) {
    o.e[0 + 3 * c] = y0[c] * idet;
    o.e[1 + 3 * c] = y1[c] * idet;
    o.e[2 + 3 * c] = y2[c] * idet;
  }
  return o;
}

// ---------------------------------------------------------------------------
// Mapping concept:
//   template<class S> S coord(int c, double t, const Pt<S>& p) const;
//   void forward_jacobian(double t, const Pt<double>& p, double* F /*N*N col-major*/) const;
//   int dim() const;
// ---------------------------------------------------------------------------

// AD Jacobian of the mapping (DifferentiationInterface.jacobian of
// u -> SVector(x(t,u...,p), ...); metrics.jl:200-214, 259-268, 685-692).
template <int N, class M, class S>
inline Vals<S, N * N> ad_jacobian(const M& m, double t, const Pt<S>& p) {
  Vals<S, N * N> J;
  auto f = [&](const auto& q) {
    using Q = typename std::decay<decltype(q[0])>::type;
    Vals<Q, N> r;
    for (int c = 0; c < N; ++c) r.e[c] = m.template coord<Q>(c, t, q);
    return r;
  };
  for (int a = 0; a < N; ++a) {
    Vals<S, N> col = derivative<N>(f, a, p);
    for (int c = 0; c < N; ++c) J.e[c + N * a] = col.e[c];
  }
  return J;
}

// ============================ 3-D closure graph ============================
template <class M, int SCH>
struct Cache3 {
  const M& m;
  double t;

  template <class S>
  S coord(int c, const Pt<S>& p) const { return m.template coord<S>(c, t, p); }

  // metrics.jl:726-734  xξ … zζ = derivative(s -> q(t,…s…,p), backend, ·)
  template <class S>
  S dcoord(int c, int axis, const Pt<S>& p) const {
    auto f = [&](const auto& q) {
      using Q = typename std::decay<decltype(q[0])>::type;
      Vals<Q, 1> r;
      r.e[0] = this->template coord<Q>(c, q);
      return r;
    };
    return derivative<1>(f, axis, p).e[0];
  }

  // metrics.jl:736-744  potentials  (∂_b q1) * q2
  template <class S>
  S potential(int q1, int b, int q2, const Pt<S>& p) const { return dcoord(q1, b, p) * coord(q2, p); }

  // metrics.jl:1536-1651  ∂ϕ_∂a_3d(ϕ) = ϕ_{a+½}(p) − ϕ_{a+½}(p − e_a)
  template <class S>
  S D(int a, int q1, int b, int q2, const Pt<S>& p) const {
    auto f = [&](const auto& q) {
      using Q = typename std::decay<decltype(q[0])>::type;
      Vals<Q, 1> r;
      r.e[0] = this->template potential<Q>(q1, b, q2, q);
      return r;
    };
    Pt<S> pm = p;
    pm[a] = pm[a] - 1.0;
    return edge_reconstruct<SCH, 1>(f, a, p).e[0] - edge_reconstruct<SCH, 1>(f, a, pm).e[0];
  }

  // metrics.jl:774-782.  row r ∈ (ξ̂,η̂,ζ̂), column c ∈ (x,y,z).
  //   column x uses (q1,q2)=(y,z), y uses (z,x), z uses (x,y)
  //   ξ̂ = D_ζ[q1_η q2] − D_η[q1_ζ q2];  η̂ = D_ξ[q1_ζ q2] − D_ζ[q1_ξ q2];
  //   ζ̂ = D_η[q1_ξ q2] − D_ξ[q1_η q2]
  template <class S>
  S Mhat(int r, int c, const Pt<S>& p) const {
    const int q1 = (c + 1) % 3, q2 = (c + 2) % 3;
    const int s = (r + 1) % 3, u = (r + 2) % 3;
    return D(u, q1, s, q2, p) - D(s, q1, u, q2, p);
  }

  // metrics.jl:697-707  inverse_jacobian_matrix = inv(AD jacobian)
  template <class S>
  Vals<S, 9> Jinv(const Pt<S>& p) const { return inv3(ad_jacobian<3>(m, t, p)); }

  // metrics.jl:1751-1793 with 818-850: G = Jinv_{axis+½}, Ghat[r][c] = M̂_{axis+½}
  void face(int axis, const Pt<real>& p, real* G, real* Ghat) const {
    auto fj = [&](const auto& q) { return this->Jinv(q); };
    Vals<real, 9> g = edge_reconstruct<SCH, 9>(fj, axis, p);
    for (int k = 0; k < 9; ++k) G[k] = g.e[k];
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) {
        auto fm = [&](const auto& q) {
          using Q = typename std::decay<decltype(q[0])>::type;
          Vals<Q, 1> v;
          v.e[0] = this->template Mhat<Q>(r, c, q);
          return v;
        };
        Ghat[r + 3 * c] = edge_reconstruct<SCH, 1>(fm, axis, p).e[0];
      }
  }
};

// ============================ 2-D closure graph ============================
template <class M, int SCH>
struct Cache2 {
  const M& m;
  double t;
  // metrics.jl:271-282
  template <class S>
  Vals<S, 4> jinv(const Pt<S>& p) const { return inv2(ad_jacobian<2>(m, t, p)); }
  template <class S>
  Vals<S, 4> normalized_jinv(const Pt<S>& p) const {
    Vals<S, 4> jac = ad_jacobian<2>(m, t, p);
    S J = det2(jac);
    Vals<S, 4> ji = inv2(jac);
    for (int k = 0; k < 4; ++k) ji.e[k] = ji.e[k] * J;
    return ji;
  }
  // metrics.jl:1735-1749 with 577-582
  void face(int axis, const Pt<real>& p, real* G, real* Ghat) const {
    auto fj = [&](const auto& q) { return this->jinv(q); };
    auto fn = [&](const auto& q) { return this->normalized_jinv(q); };
    Vals<real, 4> g = edge_reconstruct<SCH, 4>(fj, axis, p);
    Vals<real, 4> h = edge_reconstruct<SCH, 4>(fn, axis, p);
    for (int k = 0; k < 4; ++k) {
      G[k] = g.e[k];
      Ghat[k] = h.e[k];
    }
  }
};

// ============================ 1-D closure graph ============================
template <class M, int SCH>
struct Cache1 {
  const M& m;
  double t;
  // metrics.jl:310-323
  template <class S>
  Vals<S, 1> jinv(const Pt<S>& p) const {
    Vals<S, 1> j = ad_jacobian<1>(m, t, p);
    Vals<S, 1> o;
    o.e[0] = 1.0 / j.e[0];
    return o;
  }
  template <class S>
  Vals<S, 1> normalized_jinv(const Pt<S>& p) const {
    Vals<S, 1> j = ad_jacobian<1>(m, t, p);
    Vals<S, 1> o;
    o.e[0] = (1.0 / j.e[0]) * j.e[0];
    return o;
  }
  void face(int axis, const Pt<real>& p, real* G, real* Ghat) const {
    auto fj = [&](const auto& q) { return this->jinv(q); };
    auto fn = [&](const auto& q) { return this->normalized_jinv(q); };
    G[0] = edge_reconstruct<SCH, 1>(fj, axis, p).e[0];
    Ghat[0] = edge_reconstruct<SCH, 1>(fn, axis, p).e[0];
  }
};

// ---------------------------------------------------------------------------
// Per-cell evaluation = bodies of the fill kernels.
//   cell  (metrics.jl:1795-1877): F = forward.jacobian, G = inverse.Jinv,
//         J = det(F); forward=(F,J), inverse=(G,J)   [inverse stores det(F)]
//   face  (metrics.jl:1879-2053): (G,Ghat); F = inv(G); J = det(F);
//         forward=(F,J), inverse=(G,J), conserved=Ghat
// Outputs use the Julia element layout: N*N column-major doubles then J.
// ---------------------------------------------------------------------------
template <int N>
inline real detN(const real* m) {
  if constexpr (N == 1) return m[0];
  if constexpr (N == 2) {
    Vals<real, 4> v;
    for (int k = 0; k < 4; ++k) v.e[k] = m[k];
    return det2(v);
  }
  if constexpr (N == 3) {
    Vals<real, 9> v;
    for (int k = 0; k < 9; ++k) v.e[k] = m[k];
    return det3(v);
  }
}
template <int N>
inline void invN(const real* m, real* o) {
  if constexpr (N == 1) o[0] = 1.0 / m[0];
  if constexpr (N == 2) {
    Vals<real, 4> v;
    for (int k = 0; k < 4; ++k) v.e[k] = m[k];
    Vals<real, 4> r = inv2(v);
    for (int k = 0; k < 4; ++k) o[k] = r.e[k];
  }
  if constexpr (N == 3) {
    Vals<real, 9> v;
    for (int k = 0; k < 9; ++k) v.e[k] = m[k];
    Vals<real, 9> r = inv3(v);
    for (int k = 0; k < 9; ++k) o[k] = r.e[k];
  }
}

template <int N, class M>
inline void cell_metric(const M& m, double t, const Pt<real>& p, real* fwd /*N*N+1*/, real* inv /*N*N+1*/) {
  real F[N * N], G[N * N];
  m.forward_jacobian(t, p, F);
  Vals<real, N * N> Jad = ad_jacobian<N>(m, t, p);
  invN<N>(Jad.e, G);
  real J = detN<N>(F);
  for (int k = 0; k < N * N; ++k) {
    fwd[k] = F[k];
    inv[k] = G[k];
  }
  fwd[N * N] = J;
  inv[N * N] = J;
}

template <int N, int SCH, class M>
inline void face_metric(const M& m, double t, int axis, const Pt<real>& p, real* fwd, real* inv, real* cons) {
  real G[N * N], Ghat[N * N], F[N * N];
  if constexpr (N == 3) {
    Cache3<M, SCH> c{m, t};
    c.face(axis, p, G, Ghat);
  } else if constexpr (N == 2) {
    Cache2<M, SCH> c{m, t};
    c.face(axis, p, G, Ghat);
  } else {
    Cache1<M, SCH> c{m, t};
    c.face(axis, p, G, Ghat);
  }
  invN<N>(G, F);
  real J = detN<N>(F);
  for (int k = 0; k < N * N; ++k) {
    fwd[k] = F[k];
    inv[k] = G[k];
    cons[k] = Ghat[k];
  }
  fwd[N * N] = J;
  inv[N * N] = J;
}

// ADThomasLombardMetric, 3-D (metrics.jl:607-668 closures, pair tables :2336-3080, fill :3082-3188).
//
// The reference does NOT evaluate the Thomas-Lombard closure graph here.  Per potential P = (d_b' q1) q2
// (_ad_thomas_lombard_potential_from_values, :2377-2399), face-normal axis a and difference axis b it tabulates, at
// every CELL CENTRE, the nine mixed derivatives (P, P_a, P_aa, P_b, P_ab, P_aab, P_bb, P_abb, P_aabb) by seeding
// ForwardDiff duals along a and b (:2676-2701), then
//   1. reconstructs value / b-derivative / second b-derivative to the a-face from the cells I and I + e_a
//      (_edge_potential_values_from_pair_table :2889-2905 with _edge_reconstruct, CurvatureCorrected),
//   2. takes the array derivative along b: edge_high - edge_low of those face values at I - e_b, I, I + e_b
//      (_array_derivative_from_pair_table :2907-2935).
// i.e. outer reconstruction first, cell-centred difference second — the opposite order of Cache3 (which is why the
// reference's test asserts the two to agree at 1e-12, test_3d_continuous.jl:296-297).  This struct restates THAT
// order; tests compare it with Cache3<M, CURVATURE> (the equivalence the reference asserts) and with the CUDA path.
template <class M>
struct AdtlPairTable {
  const M& m;
  double t;

  // potential (d_bp q1) q2 at p, for any (nested dual) number type
  template <class S>
  S potential(int q1, int bp, int q2, const Pt<S>& p) const {
    Cache3<M, CURVATURE> c{m, t};
    return c.template potential<S>(q1, bp, q2, p);
  }

  // the nine numbers at a cell centre: out[n][k] = d_b^k d_a^n P,  n, k in 0..2
  void nine(int q1, int bp, int q2, int a, int b, const Pt<real>& p, real (&out)[3][3]) const {
    // h_n(p) = d_a^n P(p) as a callable on any number type, by AD along a
    auto jets_a = [&](const auto& q) {
      using Q = typename std::decay<decltype(q[0])>::type;
      auto f = [&](const auto& r) {
        using R = typename std::decay<decltype(r[0])>::type;
        Vals<R, 1> v;
        v.e[0] = this->template potential<R>(q1, bp, q2, r);
        return v;
      };
      Vals<Q, 1> v0, v1, v2;
      value_derivative_and_second_derivative<1>(f, a, q, v0, v1, v2);
      Vals<Q, 3> o;
      o.e[0] = v0.e[0];
      o.e[1] = v1.e[0];
      o.e[2] = v2.e[0];
      return o;
    };
    Vals<real, 3> v, d, dd;
    value_derivative_and_second_derivative<3>(jets_a, b, p, v, d, dd);
    for (int n = 0; n < 3; ++n) {
      out[n][0] = v.e[n];
      out[n][1] = d.e[n];
      out[n][2] = dd.e[n];
    }
  }

  static real recon(real v0, real d0, real dd0, real v1, real d1, real dd1) {  // metrics.jl:141-178, Curvature
    const double w = 1.0 / 12.0;
    return 0.5 * ((v0 + 0.5 * d0 + w * dd0) + (v1 - 0.5 * d1 + w * dd1));
  }

  // (Q, Q_b, Q_bb) on the a-face between the cells at p and p + e_a
  void face_vals(int q1, int bp, int q2, int a, int b, const Pt<real>& p, real (&Q)[3]) const {
    real lo[3][3], hi[3][3];
    Pt<real> pa = p;
    pa[a] = pa[a] + 1.0;
    nine(q1, bp, q2, a, b, p, lo);
    nine(q1, bp, q2, a, b, pa, hi);
    for (int k = 0; k < 3; ++k) Q[k] = recon(lo[0][k], lo[1][k], lo[2][k], hi[0][k], hi[1][k], hi[2][k]);
  }

  // array derivative along b of the face-reconstructed potential (:2907-2935)
  real term(int q1, int bp, int q2, int a, int b, const Pt<real>& p) const {
    Pt<real> pm = p, pp = p;
    pm[b] = pm[b] - 1.0;
    pp[b] = pp[b] + 1.0;
    real Q0[3], Qp[3], Qm[3];
    face_vals(q1, bp, q2, a, b, p, Q0);
    face_vals(q1, bp, q2, a, b, pp, Qp);
    face_vals(q1, bp, q2, a, b, pm, Qm);
    return recon(Q0[0], Q0[1], Q0[2], Qp[0], Qp[1], Qp[2]) - recon(Qm[0], Qm[1], Qm[2], Q0[0], Q0[1], Q0[2]);
  }

  // active row of the conserved metrics on the face `axis`: row r = axis of metrics.jl:774-782,
  //   M^[r][c] = D_u[(d_s q1) q2] - D_s[(d_u q1) q2],  s = r + 1, u = r + 2 (cyclic), (q1, q2) = (c + 1, c + 2)
  real active(int axis, int c, const Pt<real>& p) const {
    const int q1 = (c + 1) % 3, q2 = (c + 2) % 3, s = (axis + 1) % 3, u = (axis + 2) % 3;
    return term(q1, s, q2, axis, u, p) - term(q1, u, q2, axis, s, p);
  }
};

// Face forward metric = the mapping's Jacobian AT THE FACE CENTRE (edge.jacobian(t, xi_f, eta_f, zeta_f), :3162-3168),
// inverse = inv(F), J = det(F); the conserved matrix carries only the active row (:3174-3182), assembled from the
// pair tables above.
template <class M>
inline void face_metric_adtl3(const M& m, double t, int axis, const Pt<real>& p, real* fwd, real* inv, real* cons) {
  AdtlPairTable<M> tab{m, t};
  Pt<real> pf = p;
  pf[axis] += 0.5;  // _face_center_3d (:3271-3283): the cell centre moved half a cell along the face axis
  const Vals<real, 9> F = ad_jacobian<3>(m, t, pf);
  real Fm[9], Gm[9];
  for (int k = 0; k < 9; ++k) Fm[k] = F.e[k];
  invN<3>(Fm, Gm);
  const real J = detN<3>(Fm);
  for (int k = 0; k < 9; ++k) {
    fwd[k] = Fm[k];
    inv[k] = Gm[k];
    cons[k] = (k % 3 == axis) ? tab.active(axis, k / 3, p) : real(0.0);  // column-major: row = k % 3, column = k / 3
  }
  fwd[9] = J;
  inv[9] = J;
}

// ---------------------------------------------------------------------------
// DiscreteGrid mapping: discrete_grid.jl:63-71 + Interpolations.jl semantics
// (SURVEY.md Appendix A.1).  `a[c]` is the INTERIOR node array of coordinate c,
// column-major with n[d] nodes along axis d; node index 1 sits at ξ = 1.
// ---------------------------------------------------------------------------
struct DiscreteMap {
  int N;
  const double* a[3];
  int64_t n[3];  // nodes per axis (1 for unused axes)

  int dim() const { return N; }

  // value and gradient of BSpline(Linear()) at an in-bounds position xs
  template <class S>
  void itp_value_gradient(const double* A, const Pt<S>& xs, S& val, S* g) const {
    int64_t i0[3] = {0, 0, 0};
    S del[3] = {constant<S>(0.0), constant<S>(0.0), constant<S>(0.0)};
    for (int d = 0; d < N; ++d) {
      double f = std::floor(primal(xs[d]));
      if (f > double(n[d] - 1)) f -= 1.0;  // upper bound belongs to the last cell (δ = 1)
      if (f < 1.0) f = 1.0;
      i0[d] = int64_t(f) - 1;  // 0-based
      del[d] = xs[d] - f;
    }
    val = constant<S>(0.0);
    for (int d = 0; d < 3; ++d) g[d] = constant<S>(0.0);
    const int ncorner = 1 << N;
    for (int cnr = 0; cnr < ncorner; ++cnr) {
      int64_t idx = 0, stride = 1;
      S w[3];
      for (int d = 0; d < N; ++d) {
        int bit = (cnr >> d) & 1;
        idx += (i0[d] + bit) * stride;
        stride *= n[d];
        w[d] = bit ? del[d] : (1.0 - del[d]);
      }
      const double av = A[idx];
      S wv = constant<S>(av);
      for (int d = 0; d < N; ++d) wv = wv * w[d];
      val = val + wv;
      for (int d = 0; d < N; ++d) {
        S gd = constant<S>(((cnr >> d) & 1) ? av : -av);
        for (int e = 0; e < N; ++e)
          if (e != d) gd = gd * w[e];
        g[d] = g[d] + gd;
      }
    }
  }

  // extrapolate(itp, Line()):  itp(xs) + Σ_d (x_d − xs_d) ∂_d itp(xs),  xs = clamp(x)
  template <class S>
  S coord(int c, double /*t*/, const Pt<S>& x) const {
    Pt<S> xs = x;
    for (int d = 0; d < N; ++d) {
      double pv = primal(x[d]);
      if (pv > double(n[d])) xs[d] = constant<S>(double(n[d]));   // clamp(::Dual) returns the bound as a constant
      else if (pv < 1.0) xs[d] = constant<S>(1.0);
    }
    S val, g[3];
    itp_value_gradient(a[c], xs, val, g);
    // (x_d − xs_d) is identically zero (value and all derivative parts) for an
    // unclamped axis, so only clamped axes contribute.
    for (int d = 0; d < N; ++d) val = val + (x[d] - xs[d]) * g[d];
    return val;
  }

  // discrete_grid.jl:113-142  forward Jacobian = Interpolations.gradient(etp, ξ…):
  // gradient of the interpolant at the CLAMPED point (Line() extrapolate_gradient = g).
  void forward_jacobian(double /*t*/, const Pt<real>& x, real* F) const {
    Pt<real> xs = x;
    for (int d = 0; d < N; ++d) {
      if (xs[d] > real(n[d])) xs[d] = real(n[d]);
      else if (xs[d] < 1.0) xs[d] = 1.0;
    }
    for (int c = 0; c < N; ++c) {
      real val, g[3];
      itp_value_gradient(a[c], xs, val, g);
      for (int d = 0; d < N; ++d) F[c + N * d] = g[d];
    }
  }
};

// ---------------------------------------------------------------------------
// MappedGrid mapping given as a separable term list
//   q_c(t,ξ,η,ζ) = Σ_m coef_m(t) · f_m1(ξ) f_m2(η) f_m3(ζ),
//   coef(t) = c0 + c1 t + c2 sin(w t + ph),
//   f ∈ {1, a + b s, sin(a + b s), cos(a + b s)}.
// Every analytic mapping in the reference's tests/README has this form
// (README.md:40-43, test_2d_continuous.jl:21-66, test_3d_continuous.jl:10-53,
// test_3d.jl:252-257, test_mapped_discrete_face_geometry.jl:92-256).  Here it is
// just an analytic closure evaluated on dual numbers, as the reference does.
// ---------------------------------------------------------------------------
struct TermFactor {
  int kind;  // 0: 1, 1: a+b*s, 2: sin(a+b*s), 3: cos(a+b*s)
  double a, b;
};
struct Term {
  int coord;
  double c0, c1, c2, w, ph;
  TermFactor f[3];
};
struct TermMap {
  int N;
  int nterms;
  const Term* terms;
  int dim() const { return N; }
  template <class S>
  S coord(int c, double t, const Pt<S>& p) const {
    S acc = constant<S>(0.0);
    for (int m = 0; m < nterms; ++m) {
      const Term& T = terms[m];
      if (T.coord != c) continue;
      double cf = T.c0 + T.c1 * t + T.c2 * std::sin(T.w * t + T.ph);
      S v = constant<S>(cf);
      for (int d = 0; d < N; ++d) {
        const TermFactor& f = T.f[d];
        if (f.kind == 0) continue;
        S arg = f.a + f.b * p[d];
        if (f.kind == 1) v = v * arg;
        else if (f.kind == 2) v = v * sin_(arg);
        else v = v * cos_(arg);
      }
      acc = acc + v;
    }
    return acc;
  }
  // MappedGrid keeps the AD Jacobian as forward.jacobian (metrics.jl:200-214)
  void forward_jacobian(double t, const Pt<real>& p, real* F) const {
    if (N == 1) { auto J = ad_jacobian<1>(*this, t, p); F[0] = J.e[0]; }
    else if (N == 2) { auto J = ad_jacobian<2>(*this, t, p); for (int k = 0; k < 4; ++k) F[k] = J.e[k]; }
    else { auto J = ad_jacobian<3>(*this, t, p); for (int k = 0; k < 9; ++k) F[k] = J.e[k]; }
  }
};

// ---------------------------------------------------------------------------
// Non-separable analytic test mappings (the reference's MappedGrid takes ANY closures, mapped_grid.jl:421-540): the
// checker's side of tests/test_gpu_generic_mapping.py, where the same formulas are given to the library as CUDA source
// (cg_grid_set_mapping_source).  Evaluated on dual numbers like every other closure here.
//   id 1 (3-D): x = xi + p0 sin(p1 (eta + zeta / 2) + 0.3 t)
//               y = eta (1 + p2 xi) / (1 + p3 zeta^2)
//               z = zeta + p0 exp(-p4 (xi - eta)^2) cos(t) + sqrt(1 + 0.01 xi^2)
//   id 2 (2-D): x = xi + p0 sin(p1 xi eta + t),  y = eta + p2 sqrt(1 + p3 xi^2 + eta^2 / 2)
//   id 3 (1-D): x = xi + p0 sin(p1 xi) exp(-p2 xi) + p3 t
// ---------------------------------------------------------------------------
struct FuncMap {
  int N, id;
  const double* prm;
  int dim() const { return N; }
  template <class S>
  S coord(int c, double t, const Pt<S>& p) const {
    const S &xi = p[0], &eta = p[1], &zeta = p[2];
    if (id == 1) {
      if (c == 0) return xi + prm[0] * sin_(prm[1] * (eta + 0.5 * zeta) + 0.3 * t);
      if (c == 1) return eta * (1.0 + prm[2] * xi) / (1.0 + prm[3] * zeta * zeta);
      return zeta + prm[0] * exp_(-prm[4] * (xi - eta) * (xi - eta)) * std::cos(t) + sqrt_(1.0 + 0.01 * xi * xi);
    }
    if (id == 2) {
      if (c == 0) return xi + prm[0] * sin_(prm[1] * xi * eta + t);
      return eta + prm[2] * sqrt_(1.0 + prm[3] * xi * xi + 0.5 * eta * eta);
    }
    return xi + prm[0] * sin_(prm[1] * xi) * exp_(-prm[2] * xi) + prm[3] * t;
  }
  void forward_jacobian(double t, const Pt<real>& p, real* F) const {
    if (N == 1) { auto J = ad_jacobian<1>(*this, t, p); F[0] = J.e[0]; }
    else if (N == 2) { auto J = ad_jacobian<2>(*this, t, p); for (int k = 0; k < 4; ++k) F[k] = J.e[k]; }
    else { auto J = ad_jacobian<3>(*this, t, p); for (int k = 0; k < 9; ++k) F[k] = J.e[k]; }
  }
};

}  // namespace cgo

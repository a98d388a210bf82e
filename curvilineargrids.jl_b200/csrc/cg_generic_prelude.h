// Device source of the GENERIC (non-separable) MappedGrid path, compiled at run time with NVRTC together with the
// caller's mapping (cg_grid_set_mapping_source).  Two raw strings: the prelude (jet arithmetic) goes BEFORE the user
// source, the kernels AFTER it.
//
// Replaces, for arbitrary analytic mappings, what the reference does with ForwardDiff on Julia closures
// (/root/reference/src/grids/unified_grids/metrics.jl:183-341 MetricCache, :670-902 Thomas-Lombard functions,
// :1536-1651 cell-centred differences, :141-178 _edge_reconstruct, :1716-1877 cell / face fill; mapped_grid.jl:421-540
// takes any closure).  A Julia closure cannot cross a C ABI; the mapping arrives as CUDA C++ source
//
//     template <class S> __device__ void cg_map(S xi, S eta, S zeta, double t, const double* p, S& x, S& y, S& z);
//
// (2-D / 1-D mappings ignore the unused arguments) and is instantiated on truncated-Taylor number types, which is exact
// differentiation like ForwardDiff's duals.  Formulation (not the reference's closure nesting): every operator of the
// closure graph is linear and translation invariant, so a face entry is
//     Rec_f[ D_u[(d_s q1) q2] - D_s[(d_u q1) q2] ]      Rec_a[g](P) = 1/2 (W+ g(P) + W- g(P + e_a)),
//     D_a[g](P) = Rec_a[g](P) - Rec_a[g](P - e_a),      W+- = 1 +- lam1 d_a + lam2 d_a^2,
// and ONE evaluation of the mapping on the number type  D1_b< J2_a< J2_f<double> > >  (first derivative along b, value /
// first / second derivative along a, the same along the face axis f: 18 doubles) at a stencil point serves the three
// columns of a row term at once: 18 mapping evaluations per face point, 108 per cell, instead of one nested-dual
// evaluation per entry and operator.
#pragma once

namespace cg {

static const char* const GENERIC_PRELUDE = R"CGSRC(
namespace cgg {
template <class T> struct D1 { T v, d; };          // value, first derivative (one direction)
template <class T> struct J2 { T v, d, dd; };      // value, first and second derivative (one direction)

template <class T> struct Tr { static __device__ __forceinline__ T from(double c) { return c; } };
template <class T> struct Tr<D1<T> > { static __device__ __forceinline__ D1<T> from(double c) { D1<T> r = {Tr<T>::from(c), Tr<T>::from(0.0)}; return r; } };
template <class T> struct Tr<J2<T> > { static __device__ __forceinline__ J2<T> from(double c) { J2<T> r = {Tr<T>::from(c), Tr<T>::from(0.0), Tr<T>::from(0.0)}; return r; } };

#define CGG_F __device__ __forceinline__
// the double overloads of the CUDA math library, visible next to the jet overloads below
using ::sin; using ::cos; using ::exp; using ::log; using ::sqrt; using ::tanh;
// ---- D1 ----
template <class T> CGG_F D1<T> operator+(const D1<T>& a, const D1<T>& b) { D1<T> r = {a.v + b.v, a.d + b.d}; return r; }
template <class T> CGG_F D1<T> operator-(const D1<T>& a, const D1<T>& b) { D1<T> r = {a.v - b.v, a.d - b.d}; return r; }
template <class T> CGG_F D1<T> operator-(const D1<T>& a) { D1<T> r = {-a.v, -a.d}; return r; }
template <class T> CGG_F D1<T> operator+(const D1<T>& a) { return a; }
template <class T> CGG_F D1<T> operator*(const D1<T>& a, const D1<T>& b) { D1<T> r = {a.v * b.v, a.v * b.d + a.d * b.v}; return r; }
template <class T> CGG_F D1<T> operator+(const D1<T>& a, double b) { D1<T> r = {a.v + b, a.d}; return r; }
template <class T> CGG_F D1<T> operator+(double a, const D1<T>& b) { D1<T> r = {a + b.v, b.d}; return r; }
template <class T> CGG_F D1<T> operator-(const D1<T>& a, double b) { D1<T> r = {a.v - b, a.d}; return r; }
template <class T> CGG_F D1<T> operator-(double a, const D1<T>& b) { D1<T> r = {a - b.v, -b.d}; return r; }
template <class T> CGG_F D1<T> operator*(const D1<T>& a, double b) { D1<T> r = {a.v * b, a.d * b}; return r; }
template <class T> CGG_F D1<T> operator*(double a, const D1<T>& b) { D1<T> r = {a * b.v, a * b.d}; return r; }
// ---- J2 ----
template <class T> CGG_F J2<T> operator+(const J2<T>& a, const J2<T>& b) { J2<T> r = {a.v + b.v, a.d + b.d, a.dd + b.dd}; return r; }
template <class T> CGG_F J2<T> operator-(const J2<T>& a, const J2<T>& b) { J2<T> r = {a.v - b.v, a.d - b.d, a.dd - b.dd}; return r; }
template <class T> CGG_F J2<T> operator-(const J2<T>& a) { J2<T> r = {-a.v, -a.d, -a.dd}; return r; }
template <class T> CGG_F J2<T> operator+(const J2<T>& a) { return a; }
template <class T> CGG_F J2<T> operator*(const J2<T>& a, const J2<T>& b) {
  J2<T> r = {a.v * b.v, a.v * b.d + a.d * b.v, a.v * b.dd + 2.0 * (a.d * b.d) + a.dd * b.v};
  return r;
}
template <class T> CGG_F J2<T> operator+(const J2<T>& a, double b) { J2<T> r = {a.v + b, a.d, a.dd}; return r; }
template <class T> CGG_F J2<T> operator+(double a, const J2<T>& b) { J2<T> r = {a + b.v, b.d, b.dd}; return r; }
template <class T> CGG_F J2<T> operator-(const J2<T>& a, double b) { J2<T> r = {a.v - b, a.d, a.dd}; return r; }
template <class T> CGG_F J2<T> operator-(double a, const J2<T>& b) { J2<T> r = {a - b.v, -b.d, -b.dd}; return r; }
template <class T> CGG_F J2<T> operator*(const J2<T>& a, double b) { J2<T> r = {a.v * b, a.d * b, a.dd * b}; return r; }
template <class T> CGG_F J2<T> operator*(double a, const J2<T>& b) { J2<T> r = {a * b.v, a * b.d, a * b.dd}; return r; }
// ---- composition with a scalar function: f, f', f'' at the value part ----
template <class T> CGG_F D1<T> chain(const D1<T>& g, const T& f0, const T& f1, const T&) { D1<T> r = {f0, f1 * g.d}; return r; }
template <class T> CGG_F J2<T> chain(const J2<T>& g, const T& f0, const T& f1, const T& f2) {
  J2<T> r = {f0, f1 * g.d, f2 * (g.d * g.d) + f1 * g.dd};
  return r;
}
// elementary functions on double come from the CUDA math library; these overloads are found by argument-dependent lookup
CGG_F double recip(double x) { return 1.0 / x; }
template <template <class> class W, class T> CGG_F W<T> recip(const W<T>& g) {
  const T r = recip(g.v), r2 = r * r;
  return chain(g, r, -r2, 2.0 * (r2 * r));
}
template <template <class> class W, class T> CGG_F W<T> sin(const W<T>& g) {
  const T s = sin(g.v), c = cos(g.v);
  return chain(g, s, c, -s);
}
template <template <class> class W, class T> CGG_F W<T> cos(const W<T>& g) {
  const T s = sin(g.v), c = cos(g.v);
  return chain(g, c, -s, -c);
}
template <template <class> class W, class T> CGG_F W<T> exp(const W<T>& g) {
  const T e = exp(g.v);
  return chain(g, e, e, e);
}
template <template <class> class W, class T> CGG_F W<T> log(const W<T>& g) {
  const T r = recip(g.v);
  return chain(g, log(g.v), r, -(r * r));
}
template <template <class> class W, class T> CGG_F W<T> sqrt(const W<T>& g) {
  const T s = sqrt(g.v), r = recip(g.v);
  const T f1 = 0.5 * (s * r);             // 1 / (2 sqrt x)
  return chain(g, s, f1, -0.5 * (f1 * r));  // -1 / (4 x sqrt x)
}
template <template <class> class W, class T> CGG_F W<T> tanh(const W<T>& g) {
  const T th = tanh(g.v), f1 = 1.0 - th * th;
  return chain(g, th, f1, -2.0 * (th * f1));
}
template <template <class> class W, class T> CGG_F W<T> operator/(const W<T>& a, const W<T>& b) { return a * recip(b); }
template <template <class> class W, class T> CGG_F W<T> operator/(const W<T>& a, double b) { return a * (1.0 / b); }
template <template <class> class W, class T> CGG_F W<T> operator/(double a, const W<T>& b) { return a * recip(b); }
}  // namespace cgg
using namespace cgg;
)CGSRC";

static const char* const GENERIC_KERNELS = R"CGSRC(
namespace cgg {
struct GD {
  int dim, nhalo, scheme, what;
  long long nw[3], w0[3];
};
struct GOut {
  void *cell_fwd, *cell_inv, *face_fwd[3], *face_inv[3], *face_cons[3], *cjac, *cactive[3];
};
struct GCoord {
  void *node[3], *centroid[3], *face[3][3];
};
typedef J2<double> JF;   // along the face axis
typedef J2<JF> JA;       // + along the outer difference axis
typedef D1<JA> S18;      // + first derivative along the inner axis
typedef D1<JF> S6;       // Jacobian column as a jet along the face axis
typedef D1<double> S2;

template <class S> CGG_F S cst(double c) { return Tr<S>::from(c); }
template <class S> CGG_F S det3(const S (&m)[3][3]) {
  return m[0][0] * (m[1][1] * m[2][2] - m[1][2] * m[2][1]) - m[0][1] * (m[1][0] * m[2][2] - m[1][2] * m[2][0]) +
         m[0][2] * (m[1][0] * m[2][1] - m[1][1] * m[2][0]);
}
template <class S> CGG_F void inv3(const S (&m)[3][3], S (&o)[3][3]) {
  const S c00 = m[1][1] * m[2][2] - m[1][2] * m[2][1], c01 = m[1][2] * m[2][0] - m[1][0] * m[2][2],
          c02 = m[1][0] * m[2][1] - m[1][1] * m[2][0];
  const S id = recip(m[0][0] * c00 + m[0][1] * c01 + m[0][2] * c02);
  o[0][0] = c00 * id; o[1][0] = c01 * id; o[2][0] = c02 * id;
  o[0][1] = (m[0][2] * m[2][1] - m[0][1] * m[2][2]) * id;
  o[1][1] = (m[0][0] * m[2][2] - m[0][2] * m[2][0]) * id;
  o[2][1] = (m[0][1] * m[2][0] - m[0][0] * m[2][1]) * id;
  o[0][2] = (m[0][1] * m[1][2] - m[0][2] * m[1][1]) * id;
  o[1][2] = (m[0][2] * m[1][0] - m[0][0] * m[1][2]) * id;
  o[2][2] = (m[0][0] * m[1][1] - m[0][1] * m[1][0]) * id;
}
template <class S> CGG_F void inv2(const S (&m)[2][2], S (&o)[2][2], S& det) {
  det = m[0][0] * m[1][1] - m[0][1] * m[1][0];
  const S id = recip(det);
  o[0][0] = m[1][1] * id; o[0][1] = -(m[0][1] * id); o[1][0] = -(m[1][0] * id); o[1][1] = m[0][0] * id;
}
CGG_F void lams(int scheme, double& l1, double& l2) {
  l1 = scheme == 0 ? 0.0 : 0.5;
  l2 = scheme == 2 ? 1.0 / 12.0 : 0.0;
}
// computational coordinate of the centre of window cell n along axis d (SURVEY A.0: xi = Iglobal - h + 1/2)
CGG_F double centre(const GD& D, int d, long long n) { return (double)(D.w0[d] + n + 1 - D.nhalo) + 0.5; }

// Jacobian F[q][d] = d x_q / d xi_d at P as a jet along `f` (f < 0: plain values in the .v parts)
template <class S, class SEED> CGG_F void jac_at(const double (&P)[3], int N, double t, const double* prm, SEED seed, S (&F)[3][3]) {
  for (int d = 0; d < N; ++d) {
    D1<S> X[3], Q[3];
    for (int e = 0; e < 3; ++e) { Q[e].v = seed(e, P[e]); Q[e].d = cst<S>(e == d ? 1.0 : 0.0); }
    cg_map(Q[0], Q[1], Q[2], t, prm, X[0], X[1], X[2]);
    for (int q = 0; q < N; ++q) F[q][d] = X[q].d;
  }
}

template <class T> CGG_F void put_metric(T* arr, long long cell, int N, const double (&M)[3][3], double J) {
  T* o = arr + cell * (N * N + 1);
  for (int c = 0; c < N; ++c)
    for (int r = 0; r < N; ++r) o[r + N * c] = (T)M[r][c];
  o[N * N] = (T)J;
}

// ---- node / centroid / face-centre coordinates (generation.jl:187-309) ----
template <class T> CGG_F void coords_body(const GD& D, double t, const double* prm, const GCoord& C) {
  long long nn[3] = {1, 1, 1};
  for (int d = 0; d < D.dim; ++d) nn[d] = D.nw[d] + 1;
  const long long ntot = nn[0] * nn[1] * nn[2];
  for (long long lin = blockIdx.x * (long long)blockDim.x + threadIdx.x; lin < ntot; lin += (long long)gridDim.x * blockDim.x) {
    long long n[3], r = lin;
    n[0] = r % nn[0]; r /= nn[0]; n[1] = r % nn[1]; n[2] = r / nn[1];
    double xi[3] = {0, 0, 0}, X[3];
    bool is_cell = true;
    long long clin = 0, cst_ = 1;
    for (int d = 0; d < D.dim; ++d) {
      xi[d] = (double)(D.w0[d] + n[d] + 1 - D.nhalo);
      if (n[d] >= D.nw[d]) is_cell = false;
      clin += n[d] * cst_;
      cst_ *= D.nw[d];
    }
    cg_map(xi[0], xi[1], xi[2], t, prm, X[0], X[1], X[2]);
    for (int q = 0; q < D.dim; ++q)
      if (C.node[q]) ((T*)C.node[q])[lin] = (T)X[q];
    if (!is_cell) continue;
    double c[3] = {0, 0, 0};
    for (int d = 0; d < D.dim; ++d) c[d] = xi[d] + 0.5;
    cg_map(c[0], c[1], c[2], t, prm, X[0], X[1], X[2]);
    for (int q = 0; q < D.dim; ++q)
      if (C.centroid[q]) ((T*)C.centroid[q])[clin] = (T)X[q];
    for (int a = 0; a < D.dim; ++a) {
      if (!C.face[a][0]) continue;
      double p[3] = {c[0], c[1], c[2]};
      p[a] += 0.5;
      cg_map(p[0], p[1], p[2], t, prm, X[0], X[1], X[2]);
      for (int q = 0; q < D.dim; ++q)
        if (C.face[a][q]) ((T*)C.face[a][q])[clin] = (T)X[q];
    }
  }
}

// ---- cell metrics (metrics.jl:1795-1877): F = AD Jacobian at the centre, G = inv F, J = det F ----
template <class T> CGG_F void cell_body(const GD& D, double t, const double* prm, const GOut& O) {
  const int N = D.dim;
  const long long ntot = D.nw[0] * D.nw[1] * D.nw[2];
  for (long long lin = blockIdx.x * (long long)blockDim.x + threadIdx.x; lin < ntot; lin += (long long)gridDim.x * blockDim.x) {
    long long n[3], r = lin;
    n[0] = r % D.nw[0]; r /= D.nw[0]; n[1] = r % D.nw[1]; n[2] = r / D.nw[1];
    double P[3] = {0, 0, 0};
    for (int d = 0; d < N; ++d) P[d] = centre(D, d, n[d]);
    double F[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}}, G[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}}, J;
    jac_at<double>(P, N, t, prm, [](int, double v) { return v; }, F);
    if (N == 3) { J = det3(F); inv3(F, G); }
    else if (N == 2) {
      double f2[2][2] = {{F[0][0], F[0][1]}, {F[1][0], F[1][1]}}, g2[2][2];
      inv2(f2, g2, J);
      G[0][0] = g2[0][0]; G[0][1] = g2[0][1]; G[1][0] = g2[1][0]; G[1][1] = g2[1][1];
    } else { J = F[0][0]; G[0][0] = 1.0 / F[0][0]; }
    if (O.cell_fwd) put_metric((T*)O.cell_fwd, lin, N, F, J);
    if (O.cell_inv) put_metric((T*)O.cell_inv, lin, N, G, J);
    if (O.cjac) ((T*)O.cjac)[lin] = (T)J;
  }
}

// ---- face metrics of ONE face axis per thread (metrics.jl:1716-1793, 1879-2053) ----
template <class T> CGG_F void face_body(const GD& D, double t, const double* prm, const GOut& O) {
  const int N = D.dim;
  double l1, l2;
  lams(D.scheme, l1, l2);
  const long long ncell = D.nw[0] * D.nw[1] * D.nw[2], ntot = ncell * N;
  for (long long w = blockIdx.x * (long long)blockDim.x + threadIdx.x; w < ntot; w += (long long)gridDim.x * blockDim.x) {
    const int f = (int)(w / ncell);
    const long long lin = w - f * ncell;
    long long n[3], r = lin;
    n[0] = r % D.nw[0]; r /= D.nw[0]; n[1] = r % D.nw[1]; n[2] = r / D.nw[1];
    double P0[3] = {0, 0, 0};
    for (int d = 0; d < N; ++d) P0[d] = centre(D, d, n[d]);
    const bool full = O.face_fwd[f] && O.face_inv[f];
    double Gf[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}}, Gh[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
    // reconstruction of the inverse Jacobian (and, in 1-D / 2-D, of J inv F: metrics.jl:271-323)
    for (int j = 0; j < 2; ++j) {
      double P[3] = {P0[0], P0[1], P0[2]};
      P[f] += j;
      const double sg = j == 0 ? l1 : -l1;
      if (!full && N == 3) break;
      JF F[3][3], G[3][3];
      for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) { F[a][b] = cst<JF>(a == b ? 1.0 : 0.0); G[a][b] = cst<JF>(0.0); }
      jac_at<JF>(P, N, t, prm, [f](int e, double v) { JF s = {v, e == f ? 1.0 : 0.0, 0.0}; return s; }, F);
      JF Jd = cst<JF>(1.0);
      if (N == 3) inv3(F, G);
      else if (N == 2) {
        JF f2[2][2] = {{F[0][0], F[0][1]}, {F[1][0], F[1][1]}}, g2[2][2];
        inv2(f2, g2, Jd);
        G[0][0] = g2[0][0]; G[0][1] = g2[0][1]; G[1][0] = g2[1][0]; G[1][1] = g2[1][1];
      } else { Jd = F[0][0]; G[0][0] = recip(F[0][0]); }
      for (int a = 0; a < N; ++a)
        for (int b = 0; b < N; ++b) {
          Gf[a][b] += 0.5 * (G[a][b].v + sg * G[a][b].d + l2 * G[a][b].dd);
          if (N < 3) {
            const JF h = G[a][b] * Jd;  // normalised inverse Jacobian J inv F
            Gh[a][b] += 0.5 * (h.v + sg * h.d + l2 * h.dd);
          }
        }
    }
    if (N == 3) {
      // conserved metrics: row r = D_u[(d_s q1) q2] - D_s[(d_u q1) q2], s = r + 1, u = r + 2, column c: (q1, q2) = (c + 1, c + 2)
      const int rlo = O.face_cons[f] ? 0 : f, rhi = O.face_cons[f] ? 3 : f + 1;  // COMPACT profile: the active row only
      for (int j = 0; j < 2; ++j) {
        double P[3] = {P0[0], P0[1], P0[2]};
        P[f] += j;
        const double sg = j == 0 ? l1 : -l1;
        for (int rr = rlo; rr < rhi; ++rr) {
          JF M[3] = {cst<JF>(0.0), cst<JF>(0.0), cst<JF>(0.0)};
          for (int term = 0; term < 2; ++term) {
            const int s = (rr + 1) % 3, u = (rr + 2) % 3;
            const int a = term == 0 ? u : s, b = term == 0 ? s : u;
            const double tsign = term == 0 ? 1.0 : -1.0;
            for (int k = -1; k <= 1; ++k) {
              if (k == 0 && l1 == 0.0) continue;
              S18 Q[3], X[3];
              for (int e = 0; e < 3; ++e) {
                const JF v0 = {P[e] + (e == a ? k : 0), e == f ? 1.0 : 0.0, 0.0};
                const JA v1 = {v0, cst<JF>(e == a ? 1.0 : 0.0), cst<JF>(0.0)};
                Q[e].v = v1;
                Q[e].d = cst<JA>(e == b ? 1.0 : 0.0);
              }
              cg_map(Q[0], Q[1], Q[2], t, prm, X[0], X[1], X[2]);
              for (int c = 0; c < 3; ++c) {
                const JA ph = X[(c + 1) % 3].d * X[(c + 2) % 3].v;  // potential (d_b q1) q2 as a jet along a (and f)
                JF g;
                if (k == 1) g = 0.5 * (ph.v - l1 * ph.d + l2 * ph.dd);
                else if (k == 0) g = l1 * ph.d;
                else g = -0.5 * (ph.v + l1 * ph.d + l2 * ph.dd);
                M[c] = M[c] + tsign * g;
              }
            }
          }
          for (int c = 0; c < 3; ++c) Gh[rr][c] += 0.5 * (M[c].v + sg * M[c].d + l2 * M[c].dd);
        }
      }
    }
    if (full) {
      double Ff[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}}, Jf;
      if (N == 3) { inv3(Gf, Ff); Jf = det3(Ff); }
      else if (N == 2) {
        double g2[2][2] = {{Gf[0][0], Gf[0][1]}, {Gf[1][0], Gf[1][1]}}, f2[2][2], dg;
        inv2(g2, f2, dg);
        Ff[0][0] = f2[0][0]; Ff[0][1] = f2[0][1]; Ff[1][0] = f2[1][0]; Ff[1][1] = f2[1][1];
        Jf = f2[0][0] * f2[1][1] - f2[0][1] * f2[1][0];
      } else { Ff[0][0] = 1.0 / Gf[0][0]; Jf = Ff[0][0]; }
      put_metric((T*)O.face_fwd[f], lin, N, Ff, Jf);
      put_metric((T*)O.face_inv[f], lin, N, Gf, Jf);
    }
    if (O.face_cons[f]) {
      T* o = (T*)O.face_cons[f] + lin * (N * N);
      for (int c = 0; c < N; ++c)
        for (int rr = 0; rr < N; ++rr) o[rr + N * c] = (T)Gh[rr][c];
    }
    if (O.cactive[f]) {
      T* o = (T*)O.cactive[f] + lin * N;
      for (int c = 0; c < N; ++c) o[c] = (T)Gh[f][c];
    }
  }
}
}  // namespace cgg

extern "C" __global__ void cgg_coords_f64(cgg::GD D, double t, const double* prm, cgg::GCoord C) { cgg::coords_body<double>(D, t, prm, C); }
extern "C" __global__ void cgg_coords_f32(cgg::GD D, double t, const double* prm, cgg::GCoord C) { cgg::coords_body<float>(D, t, prm, C); }
extern "C" __global__ void cgg_cell_f64(cgg::GD D, double t, const double* prm, cgg::GOut O) { cgg::cell_body<double>(D, t, prm, O); }
extern "C" __global__ void cgg_cell_f32(cgg::GD D, double t, const double* prm, cgg::GOut O) { cgg::cell_body<float>(D, t, prm, O); }
extern "C" __global__ void __launch_bounds__(128) cgg_face_f64(cgg::GD D, double t, const double* prm, cgg::GOut O) { cgg::face_body<double>(D, t, prm, O); }
extern "C" __global__ void __launch_bounds__(128) cgg_face_f32(cgg::GD D, double t, const double* prm, cgg::GOut O) { cgg::face_body<float>(D, t, prm, O); }
)CGSRC";

}  // namespace cg

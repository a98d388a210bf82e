// Closed-form building blocks of the DiscreteGrid metric kernels (host+device).
//
// The reference differentiates a multilinear interpolant with nested AD
// (/root/reference/src/grids/unified_grids/metrics.jl:670-902, 1536-1651,
// discrete_grid.jl:63-71).  Inside one cell the interpolant is the centred
// multilinear form  q(d) = sum_S q_S prod_{a in S} d_a , so every AD result is a
// polynomial identity in the 8 corner nodes.  DESIGN.md §3 derives the forms
// used here; tests/closed_form_numpy.py states them in numpy and
// tests/test_closed_form_vs_oracle.py checks them against the literal oracle.
#pragma once

#if defined(__CUDACC__)
#define CG_HD __host__ __device__ __forceinline__
#else
#define CG_HD inline
#endif

namespace cg {

enum Scheme { ENDPOINT = 0, GRADIENT = 1, CURVATURE = 2 };

// reconstruction weights: L = P + lam1 P' + lam2 P'',  R = P - lam1 P' + lam2 P''
// (metrics.jl:141-178; +1/12 on both sides for Curvature)
template <class T> CG_HD T lam1_of(int scheme) { return scheme == ENDPOINT ? T(0) : T(0.5); }
template <class T> CG_HD T lam2_of(int scheme) { return scheme == CURVATURE ? T(1.0 / 12.0) : T(0); }

// centred bilinear form of 4 face nodes n[bu][bv]:  c0 + cu*du + cv*dv + cuv*du*dv
//
// Round-off rule of every kernel in this library: a FIRST difference is taken between two neighbouring nodes
// (exact, or exact to eps * dx) and only then summed — never as a difference of node SUMS, which carries an
// absolute error eps * |x|.  The conserved metrics multiply such a difference (size dx) with a coordinate VALUE
// (size |x|): with differences of sums the relative error of the result is eps (|x| / dx)^2 (7e-12 on the 512^3
// BASELINE mesh — what the reference's own Float64 evaluation shows against a long-double evaluation), with
// neighbour differences it is eps |x| / dx.
template <class T>
struct Form4 {
  T c0, cu, cv, cuv;
};
template <class T>
CG_HD Form4<T> face_form(T n00, T n10, T n01, T n11) {
  T s0 = n10 + n00, d0 = n10 - n00, s1 = n11 + n01, d1 = n11 - n01;
  Form4<T> f;
  f.c0 = T(0.25) * (s0 + s1);
  f.cu = T(0.5) * (d0 + d1);
  f.cv = T(0.5) * ((n01 - n00) + (n11 - n10));
  f.cuv = d1 - d0;
  return f;
}
template <class T> CG_HD Form4<T> form_avg(const Form4<T>& a, const Form4<T>& b) {
  return {T(0.5) * (a.c0 + b.c0), T(0.5) * (a.cu + b.cu), T(0.5) * (a.cv + b.cv), T(0.5) * (a.cuv + b.cuv)};
}
template <class T> CG_HD Form4<T> form_dif(const Form4<T>& hi, const Form4<T>& lo) {
  return {hi.c0 - lo.c0, hi.cu - lo.cu, hi.cv - lo.cv, hi.cuv - lo.cuv};
}

// 3x3 helpers; M[r][c] row-major in registers
// 1/x on the device: MUFU.RCP64H seed (rcp.approx, ~2^-20) + two Newton steps, straight-line code (no
// special-case branch / slow-path call like the generic division or __drcp_rn; <= 1 ulp, x = 0 gives NaN)
CG_HD double rcp(double x) {
#if defined(__CUDA_ARCH__) && defined(CG_RCP_RN)
  return __drcp_rn(x);
#elif defined(__CUDA_ARCH__)
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double e = fma(-x, y, 1.0);
  e = fma(e, e, e);
  y = fma(y, e, y);
  e = fma(-x, y, 1.0);
  return fma(y, e, y);
#else
  return 1.0 / x;
#endif
}
CG_HD float rcp(float x) {
#if defined(__CUDA_ARCH__)
  return __frcp_rn(x);
#else
  return 1.0f / x;
#endif
}

template <class T>
CG_HD T det3(const T (&m)[3][3]) {
  return m[0][0] * (m[1][1] * m[2][2] - m[1][2] * m[2][1]) - m[0][1] * (m[1][0] * m[2][2] - m[1][2] * m[2][0]) +
         m[0][2] * (m[1][0] * m[2][1] - m[1][1] * m[2][0]);
}
template <class T>
CG_HD T inv3(const T (&m)[3][3], T (&o)[3][3]) {  // returns det
  T c00 = m[1][1] * m[2][2] - m[1][2] * m[2][1];
  T c01 = m[1][2] * m[2][0] - m[1][0] * m[2][2];
  T c02 = m[1][0] * m[2][1] - m[1][1] * m[2][0];
  T det = m[0][0] * c00 + m[0][1] * c01 + m[0][2] * c02;
  T id = rcp(det);
  o[0][0] = c00 * id;
  o[1][0] = c01 * id;
  o[2][0] = c02 * id;
  o[0][1] = (m[0][2] * m[2][1] - m[0][1] * m[2][2]) * id;
  o[1][1] = (m[0][0] * m[2][2] - m[0][2] * m[2][0]) * id;
  o[2][1] = (m[0][1] * m[2][0] - m[0][0] * m[2][1]) * id;
  o[0][2] = (m[0][1] * m[1][2] - m[0][2] * m[1][1]) * id;
  o[1][2] = (m[0][2] * m[1][0] - m[0][0] * m[1][2]) * id;
  o[2][2] = (m[0][0] * m[1][1] - m[0][1] * m[1][0]) * id;
  return det;
}
template <class T>
CG_HD void mul3(const T (&a)[3][3], const T (&b)[3][3], T (&o)[3][3]) {
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) o[r][c] = a[r][0] * b[0][c] + a[r][1] * b[1][c] + a[r][2] * b[2][c];
}
template <class T>
CG_HD T inv2(const T (&m)[2][2], T (&o)[2][2]) {
  T det = m[0][0] * m[1][1] - m[0][1] * m[1][0];
  T id = rcp(det);
  o[0][0] = m[1][1] * id;
  o[0][1] = -m[0][1] * id;
  o[1][0] = -m[1][0] * id;
  o[1][1] = m[0][0] * id;
  return det;
}
template <class T>
CG_HD void mul2(const T (&a)[2][2], const T (&b)[2][2], T (&o)[2][2]) {
#pragma unroll
  for (int r = 0; r < 2; ++r)
#pragma unroll
    for (int c = 0; c < 2; ++c) o[r][c] = a[r][0] * b[0][c] + a[r][1] * b[1][c];
}

// Left / right reconstruction states of G(d) = inv(F(d)) along one axis at the
// cell centre: G' = -G F' G, G'' = 2 G F' G F' G (F'' = 0).
//   L = G + lam1 G' + lam2 G'',  R = G - lam1 G' + lam2 G''
template <class T>
CG_HD void ginv_states3(const T (&G)[3][3], const T (&Fp)[3][3], T lam1, T lam2, T (&L)[3][3], T (&R)[3][3]) {
  T GFp[3][3], G1[3][3], G2[3][3];
  mul3(G, Fp, GFp);
  mul3(GFp, G, G1);  // = -G'
  mul3(GFp, G1, G2); // = +G''/2
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      T s = G[r][c] + (T(2) * lam2) * G2[r][c];
      T d = lam1 * G1[r][c];
      L[r][c] = s - d;
      R[r][c] = s + d;
    }
}
// same, for a derivative axis AX whose column of F' is identically zero (d/d(xi_AX) of
// dq/d(xi_AX) vanishes inside a multilinear cell): Fa, Fb are the columns (AX+1)%3, (AX+2)%3 of F'.
template <int AX, class T>
CG_HD void ginv_states3_ax(const T (&G)[3][3], const T (&Fa)[3], const T (&Fb)[3], T lam1, T lam2, T (&L)[3][3],
                           T (&R)[3][3]) {
  constexpr int A = (AX + 1) % 3, B = (AX + 2) % 3;
  T Pa[3], Pb[3];  // columns A and B of G*F' (column AX is zero)
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    Pa[r] = G[r][0] * Fa[0] + G[r][1] * Fa[1] + G[r][2] * Fa[2];
    Pb[r] = G[r][0] * Fb[0] + G[r][1] * Fb[1] + G[r][2] * Fb[2];
  }
  T G1[3][3];
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) G1[r][c] = Pa[r] * G[A][c] + Pb[r] * G[B][c];  // = -G'
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      T g2 = Pa[r] * G1[A][c] + Pb[r] * G1[B][c];  // = G''/2
      T s = G[r][c] + (T(2) * lam2) * g2;
      T d = lam1 * G1[r][c];
      L[r][c] = s - d;
      R[r][c] = s + d;
    }
}

template <class T>
CG_HD void ginv_states2(const T (&G)[2][2], const T (&Fp)[2][2], T lam1, T lam2, T (&L)[2][2], T (&R)[2][2]) {
  T GFp[2][2], G1[2][2], G2[2][2];
  mul2(G, Fp, GFp);
  mul2(GFp, G, G1);
  mul2(GFp, G1, G2);
#pragma unroll
  for (int r = 0; r < 2; ++r)
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      T s = G[r][c] + (T(2) * lam2) * G2[r][c];
      T d = lam1 * G1[r][c];
      L[r][c] = s - d;
      R[r][c] = s + d;
    }
}

// quadratic in one offset: e0 + e1 d + e2 d^2
template <class T>
struct Poly2 {
  T e0, e1, e2;
};
// W+-[Q] = 1/2 Q(0) +- 1/2 lam1 Q'(0) + 1/2 lam2 Q''(0)  (half of the L / R state)
template <class T> CG_HD T w_plus(const Poly2<T>& q, T lam1, T lam2) { return T(0.5) * q.e0 + (T(0.5) * lam1) * q.e1 + lam2 * q.e2; }
template <class T> CG_HD T w_minus(const Poly2<T>& q, T lam1, T lam2) { return T(0.5) * q.e0 - (T(0.5) * lam1) * q.e1 + lam2 * q.e2; }

// Line scalars of a potential P = p0 + p1 d + p2 d^2 along the face-normal axis
// (pairs whose difference axis equals the face normal):
//   face value = 1/2 { (V+ - 2U)_I - V+_{I-f} + (2U - V-)_{I+f} + V-_{I+2f} }
template <class T>
struct LineS {
  T vp, vm, u;
};
template <class T>
CG_HD LineS<T> line_scalars(T p0, T p1, T p2, T lam1, T lam2) {
  T a = T(2) * lam2 + lam1 * lam1, b = T(2) * lam2 - lam1 * lam1;
  LineS<T> s;
  T base = T(0.5) * p0 + a * p2;
  s.vp = base + lam1 * p1;
  s.vm = base - lam1 * p1;
  s.u = T(0.5) * p0 + b * p2;
  return s;
}

}  // namespace cg

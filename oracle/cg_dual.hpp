// ORACLE — TEST INFRASTRUCTURE ONLY.  Never linked into the product library.
//
// Forward-mode dual numbers, nestable to any depth, standing in for the
// ForwardDiff.jl `Dual` type that the reference differentiates its mapping
// closures with (DifferentiationInterface `derivative`, `value_and_derivative`,
// `value_derivative_and_second_derivative`; call sites in
// /root/reference/src/grids/unified_grids/metrics.jl:190-214, 251-268,
// 367-421, 919-1113, 1547-1651).  ForwardDiff 1.3.3 / DifferentiationInterface
// 0.7.18 are not vendored in /root/reference; forward-mode AD is exact
// differentiation, so the restatement is the textbook dual-number algebra.
#pragma once
#include <cmath>

// Base real type of the restatement.  Default double (= the reference's Float64 arithmetic).  The extended builds
// (liboracle_ld.so: long double, 64-bit mantissa; liboracle_q.so: __float128) evaluate the SAME closure graph in
// higher precision; tests use them as "truth" to tell which side owns a last-digits difference between the FP64
// oracle and the CUDA kernels (conserved metrics are O(dx^2) differences of O(|x| dx) products, so both FP64 sides
// carry round-off of order eps |x| / dx).
#ifndef CGO_REAL
#define CGO_REAL double
#endif
#if defined(CGO_REAL_IS_QUAD)
#include <quadmath.h>
#endif

namespace cgo {

using real = CGO_REAL;

template <class T>
struct Dual {
  T v;  // value
  T d;  // derivative part
};

// ---- construction of constants and access to the innermost primal ----------
template <class T>
struct dtraits {
  static T from(double c) { return c; }
  static double primal(const T& x) { return (double)x; }
};
template <class T>
struct dtraits<Dual<T>> {
  static Dual<T> from(double c) { return {dtraits<T>::from(c), dtraits<T>::from(0.0)}; }
  static double primal(const Dual<T>& x) { return dtraits<T>::primal(x.v); }
};
template <class T>
inline T constant(double c) { return dtraits<T>::from(c); }
template <class T>
inline double primal(const T& x) { return dtraits<T>::primal(x); }

// ---- arithmetic -------------------------------------------------------------
template <class T> inline Dual<T> operator+(const Dual<T>& a, const Dual<T>& b) { return {a.v + b.v, a.d + b.d}; }
template <class T> inline Dual<T> operator-(const Dual<T>& a, const Dual<T>& b) { return {a.v - b.v, a.d - b.d}; }
template <class T> inline Dual<T> operator-(const Dual<T>& a) { return {-a.v, -a.d}; }
template <class T> inline Dual<T> operator*(const Dual<T>& a, const Dual<T>& b) { return {a.v * b.v, a.v * b.d + a.d * b.v}; }
template <class T> inline Dual<T> operator/(const Dual<T>& a, const Dual<T>& b) {
  T q = a.v / b.v;
  return {q, (a.d - q * b.d) / b.v};
}
template <class T> inline Dual<T> operator+(const Dual<T>& a, double b) { return {a.v + b, a.d}; }
template <class T> inline Dual<T> operator+(double a, const Dual<T>& b) { return {a + b.v, b.d}; }
template <class T> inline Dual<T> operator-(const Dual<T>& a, double b) { return {a.v - b, a.d}; }
template <class T> inline Dual<T> operator-(double a, const Dual<T>& b) { return {a - b.v, -b.d}; }
template <class T> inline Dual<T> operator*(const Dual<T>& a, double b) { return {a.v * b, a.d * b}; }
template <class T> inline Dual<T> operator*(double a, const Dual<T>& b) { return {a * b.v, a * b.d}; }
template <class T> inline Dual<T> operator/(const Dual<T>& a, double b) { return {a.v / b, a.d / b}; }
template <class T> inline Dual<T> operator/(double a, const Dual<T>& b) {
  T q = a / b.v;
  return {q, -(q * b.d) / b.v};
}

// ---- elementary functions (the reference's test mappings use sin/cos/sinpi) -
inline double sin_(double x) { return std::sin(x); }
inline double cos_(double x) { return std::cos(x); }
#if defined(CGO_REAL_IS_LD)
inline long double sin_(long double x) { return sinl(x); }
inline long double cos_(long double x) { return cosl(x); }
#elif defined(CGO_REAL_IS_QUAD)
inline __float128 sin_(__float128 x) { return sinq(x); }
inline __float128 cos_(__float128 x) { return cosq(x); }
#endif
inline double exp_(double x) { return std::exp(x); }
inline double sqrt_(double x) { return std::sqrt(x); }
#if defined(CGO_REAL_IS_LD)
inline long double exp_(long double x) { return expl(x); }
inline long double sqrt_(long double x) { return sqrtl(x); }
#elif defined(CGO_REAL_IS_QUAD)
inline __float128 exp_(__float128 x) { return expq(x); }
inline __float128 sqrt_(__float128 x) { return sqrtq(x); }
#endif
template <class T> inline Dual<T> exp_(const Dual<T>& a) {
  T e = exp_(a.v);
  return {e, e * a.d};
}
template <class T> inline Dual<T> sqrt_(const Dual<T>& a) {
  T s = sqrt_(a.v);
  return {s, a.d / (2.0 * s)};
}
template <class T> inline Dual<T> sin_(const Dual<T>& a);
template <class T> inline Dual<T> cos_(const Dual<T>& a);
template <class T> inline Dual<T> sin_(const Dual<T>& a) { return {sin_(a.v), cos_(a.v) * a.d}; }
template <class T> inline Dual<T> cos_(const Dual<T>& a) { return {cos_(a.v), -(sin_(a.v) * a.d)}; }

}  // namespace cgo

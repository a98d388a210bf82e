// Ranges of a tile of k_legacy_fused (cg_legacy.cu) along the derivative axis, as absolute positions [lo, hi) in [0, n):
// which half-jumps h, first derivatives d1 and source values phi the outputs [t0, t1) need when the stencil is chosen from
// the ABSOLUTE position (central inside, six-point one-sided windows in the first / last three cells).  Host + device:
// tests/test_legacy_ranges_cpu.py checks the closure of these ranges exhaustively without a GPU.
#pragma once
#if defined(__CUDACC__)
#define CGL_HD __host__ __device__ __forceinline__
#else
#define CGL_HD inline
#endif

struct LegacyRanges {
  int hl, hh;  // h = 1/2 d1 + 1/12 d2
  int dl, dh;  // d1
  int pl, ph;  // phi
};

CGL_HD int cgl_min(int a, int b) { return a < b ? a : b; }
CGL_HD int cgl_max(int a, int b) { return a > b ? a : b; }

// interp: edge interpolation (needs h at p and p + 1); otherwise the MEG cell-centre derivative (h at p - 1 .. p + 1)
CGL_HD LegacyRanges legacy_tile_ranges(int n, int t0, int t1, bool interp) {
  LegacyRanges r;
  r.hl = interp ? t0 : cgl_max(0, t0 - 1);
  r.hh = cgl_min(n, t1 + 1);
  r.dl = cgl_max(0, r.hl - 1);
  r.dh = cgl_min(n, r.hh + 1);
  if (r.hl < 3) r.dh = cgl_max(r.dh, 6);  // a one-sided second derivative reads ALL of d1[0 .. 5] ...
  if (r.hh > n - 3) {                     // ... or of d1[n-6 .. n-1]
    r.dl = cgl_min(r.dl, n - 6);
    r.dh = n;
  }
  r.pl = cgl_max(0, r.dl - 3);
  r.ph = cgl_min(n, r.dh + 3);
  if (r.dl < 3) r.ph = cgl_max(r.ph, 6);  // one-sided first derivatives read phi[0 .. 5] / phi[n-6 .. n-1]
  if (r.dh > n - 3) r.pl = cgl_min(r.pl, n - 6);
  return r;
}

"""CPU check of the tile ranges of the fused legacy MEG6 kernel (csrc/cg_legacy_ranges.h, host + device): for every domain
length n, both tile lengths and both operations, every value an output needs — with the stencil chosen from the ABSOLUTE
position (central inside, six-point one-sided windows in the first / last three cells) — lies inside the ranges the tile
computes, and the ranges fit the shared-memory arrays.  (The case n = k T + 3 — a last tile of exactly three cells, whose
one-sided second derivatives reach back into the tile before — was wrong before this test existed.)"""
import os
import subprocess
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SRC = r"""
#include <cstdio>
#include "cg_legacy_ranges.h"
static bool in(int p, int lo, int hi) { return p >= lo && p < hi; }
int main() {
  const int Ts[2] = {54, 118};
  long checked = 0;
  for (int ti = 0; ti < 2; ++ti) {
    const int T = Ts[ti];
    for (int interp = 0; interp < 2; ++interp)
      for (int n = 6; n <= 700; ++n)
        for (int t0 = 0; t0 < n; t0 += T) {
          const int t1 = t0 + T < n ? t0 + T : n;
          const LegacyRanges r = legacy_tile_ranges(n, t0, t1, interp != 0);
          bool ok = r.ph - r.pl <= T + 10 && r.dh - r.dl <= T + 5 && r.hh - r.hl <= T + 2 && r.pl >= 0 && r.ph <= n &&
                    r.dl >= r.pl && r.dh <= r.ph && r.hl >= r.dl && r.hh <= r.dh;
          for (int p = t0; p < t1 && ok; ++p) {  // outputs -> h
            ok = in(p, r.hl, r.hh) && (p + 1 >= n || in(p + 1, r.hl, r.hh)) && (interp || p == 0 || in(p - 1, r.hl, r.hh));
          }
          for (int q = r.hl; q < r.hh && ok; ++q) {  // h -> d1 (and phi for the central second derivative)
            ok = in(q, r.dl, r.dh) && in(q, r.pl, r.ph);
            if (q >= 3 && q < n - 3) ok = ok && in(q - 1, r.dl, r.dh) && in(q + 1, r.dl, r.dh) && in(q - 1, r.pl, r.ph) && in(q + 1, r.pl, r.ph);
            else for (int m = 0; m < 6; ++m) ok = ok && in((q < 3 ? 0 : n - 6) + m, r.dl, r.dh);
          }
          for (int d = r.dl; d < r.dh && ok; ++d) {  // d1 -> phi
            if (d >= 3 && d < n - 3) ok = in(d - 3, r.pl, r.ph) && in(d + 3, r.pl, r.ph);
            else for (int m = 0; m < 6; ++m) ok = ok && in((d < 3 ? 0 : n - 6) + m, r.pl, r.ph);
          }
          if (!ok) { printf("FAIL T=%d interp=%d n=%d t0=%d: h[%d,%d) d1[%d,%d) phi[%d,%d)\n", T, interp, n, t0, r.hl, r.hh, r.dl, r.dh, r.pl, r.ph); return 1; }
          ++checked;
        }
  }
  printf("OK %ld tiles\n", checked);
  return 0;
}
"""


def test_fused_legacy_tile_ranges_are_closed():
    with tempfile.TemporaryDirectory() as d:
        src, exe = os.path.join(d, "r.cpp"), os.path.join(d, "r")
        open(src, "w").write(SRC)
        subprocess.check_call(["g++", "-O1", "-std=c++17", "-I", os.path.join(ROOT, "curvilineargrids.jl_b200", "csrc"), src, "-o", exe])
        out = subprocess.run([exe], capture_output=True, text=True)
        assert out.returncode == 0 and out.stdout.startswith("OK"), out.stdout + out.stderr


def test_branch_free_tol_diff_formula_equals_the_isapprox_definition():
    """csrc/cg_legacy.cu evaluates tol_diff(a, b) = isapprox(a, b; rtol = sqrt(eps), atol = 10 eps) ? (a - b) * 0 : a - b
    without branches: max(|a|, |b|) on the bit patterns, "not finite" as an all-ones exponent of that maximum, and
    |d| <= thr || |d| <= atol instead of |d| <= max(atol, thr).  The same formula in numpy agrees with the plain definition
    (derivative_stencils.jl:13-16, Base.isapprox) on random, nearly equal, equal, signed-zero, subnormal, infinite and NaN
    inputs — bit for bit, NaN for NaN."""
    import numpy as np
    rng = np.random.default_rng(0)
    base = rng.standard_normal(200000) * 10.0 ** rng.integers(-12, 12, 200000)
    rel = 1.0 + rng.standard_normal(200000) * 10.0 ** rng.integers(-17, -5, 200000)
    special = np.array([0.0, -0.0, 1.0, -1.0, np.inf, -np.inf, np.nan, 5e-324, -5e-324, 2.2e-308, 1e-15, 2.3e-15, 1.7e308])
    a = np.concatenate([base, base, np.repeat(special, len(special))])
    b = np.concatenate([base * rel, base, np.tile(special, len(special))])
    rtol, atol = 1.4901161193847656e-08, 10 * np.finfo(np.float64).eps
    with np.errstate(all="ignore"):
        d = a - b
        # plain definition
        approx = (a == b) | (np.isfinite(a) & np.isfinite(b) & (np.abs(d) <= np.maximum(atol, rtol * np.maximum(np.abs(a), np.abs(b)))))
        want = np.where(approx, d * 0.0, d)
        # branch-free formula of the kernel
        ua, ub = np.abs(a).view(np.uint64), np.abs(b).view(np.uint64)
        um = np.maximum(ua, ub)
        fin = (um >> np.uint64(32)) < np.uint64(0x7FF00000)
        thr = rtol * um.view(np.float64)
        near = (np.abs(d) <= thr) | (np.abs(d) <= atol)
        got = np.where(fin & near, d * 0.0, d)
    both_nan = np.isnan(want) & np.isnan(got)
    assert np.array_equal(got.view(np.uint64)[~both_nan], want.view(np.uint64)[~both_nan])
    assert int(approx.sum()) > 1000 and int((~approx).sum()) > 1000

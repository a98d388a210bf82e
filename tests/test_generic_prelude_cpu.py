"""CPU check of the truncated-Taylor number types the generic MappedGrid path differentiates mappings with
(csrc/cg_generic_prelude.h, the source NVRTC compiles for the device): the same text, built for the host with g++, must
reproduce the analytic derivatives (sympy) of a function that uses every overloaded operation, also when the types are
nested the way the face kernel nests them (D1<J2<J2<double>>>, both jets along the same direction: derivatives up to the
fifth order)."""
import os
import re
import subprocess
import tempfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

MAIN = r"""
template <class S> S fun(S x) {
  return sin(x * x) * exp(-x) / (1.0 + sqrt(x)) + log(x) * tanh(0.5 * x) - cos(x) / x + 2.0 * x - (x - 0.25) * 3.0 + (0.5 - x) / 4.0;
}
int main() {
  const double x0 = 0.7;
  // J2: value, first, second derivative
  J2<double> a = {x0, 1.0, 0.0};
  J2<double> r = fun(a);
  printf("%.17g %.17g %.17g\n", r.v, r.d, r.dd);
  // J2<J2>: both jets along x -> v.v f, v.d f', d.d f'', d.dd = dd.d f''', dd.dd f''''
  J2<J2<double> > b = {{x0, 1.0, 0.0}, {1.0, 0.0, 0.0}, {0.0, 0.0, 0.0}};
  J2<J2<double> > q = fun(b);
  printf("%.17g %.17g %.17g %.17g %.17g %.17g\n", q.v.v, q.v.d, q.d.d, q.d.dd, q.dd.d, q.dd.dd);
  // D1<J2<J2>>: one more first derivative on top -> .d.dd.dd is the fifth derivative
  D1<J2<J2<double> > > c = {b, {{1.0, 0.0, 0.0}, {0.0, 0.0, 0.0}, {0.0, 0.0, 0.0}}};
  D1<J2<J2<double> > > w = fun(c);
  printf("%.17g %.17g %.17g\n", w.v.v.v, w.d.v.v, w.d.dd.dd);
  return 0;
}
"""


def test_jet_arithmetic_matches_analytic_derivatives():
    sympy = pytest.importorskip("sympy")
    hdr = open(os.path.join(ROOT, "curvilineargrids.jl_b200", "csrc", "cg_generic_prelude.h")).read()
    prelude = re.search(r'GENERIC_PRELUDE = R"CGSRC\((.*?)\)CGSRC"', hdr, re.S).group(1)
    src = "#include <cmath>\n#include <cstdio>\n#define __device__\n#define __forceinline__ inline\n" + prelude + MAIN
    with tempfile.TemporaryDirectory() as d:
        cpp, exe = os.path.join(d, "j.cpp"), os.path.join(d, "j")
        open(cpp, "w").write(src)
        subprocess.check_call(["g++", "-O1", "-std=c++17", cpp, "-o", exe])
        out = subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split()
    got = [float(v) for v in out]
    x = sympy.symbols("x")
    f = (sympy.sin(x * x) * sympy.exp(-x) / (1 + sympy.sqrt(x)) + sympy.log(x) * sympy.tanh(x / 2) - sympy.cos(x) / x + 2 * x
         - (x - sympy.Rational(1, 4)) * 3 + (sympy.Rational(1, 2) - x) / 4)
    der = [float(sympy.diff(f, x, k).subs(x, sympy.Rational(7, 10)).evalf(30)) for k in range(6)]
    want = [der[0], der[1], der[2],                                   # J2
            der[0], der[1], der[2], der[3], der[3], der[4],           # J2<J2>
            der[0], der[1], der[5]]                                   # D1<J2<J2>>
    assert len(got) == len(want)
    for g, w in zip(got, want):
        assert abs(g - w) <= 2e-11 * max(1.0, abs(w)), (g, w)

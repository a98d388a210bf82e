"""The generic MappedGrid kernels (csrc/cg_generic_prelude.h: the text NVRTC compiles for the device) built for the HOST
with g++ and run serially (one "thread", grid-stride loops) against the oracle's nested-dual evaluation of the same
non-separable mappings: the closure-graph algebra of the NVRTC path is parity-checked without a GPU.  The GPU tests
(tests/test_gpu_generic_mapping.py) run the same comparison through the C ABI on the device."""
import os
import re
import subprocess
import sys
import tempfile

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from util import RTOL_F64, assert_metrics_close  # noqa: E402

SHIM = r"""
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#define __device__
#define __global__
#define __forceinline__ inline
#define __launch_bounds__(x)
struct Dim3 { unsigned x, y, z; };
static const Dim3 blockIdx = {0, 0, 0}, threadIdx = {0, 0, 0}, blockDim = {1, 1, 1}, gridDim = {1, 1, 1};
"""

MAPS = {
    1: """template <class S> void cg_map(S xi, S eta, S zeta, double t, const double* p, S& x, S& y, S& z) {
  x = xi + p[0] * sin(p[1] * (eta + 0.5 * zeta) + 0.3 * t);
  y = eta * (1.0 + p[2] * xi) / (1.0 + p[3] * zeta * zeta);
  z = zeta + p[0] * exp(-p[4] * (xi - eta) * (xi - eta)) * cos(t) + sqrt(1.0 + 0.01 * xi * xi);
}""",
    2: """template <class S> void cg_map(S xi, S eta, S zeta, double t, const double* p, S& x, S& y, S& z) {
  x = xi + p[0] * sin(p[1] * xi * eta + t);
  y = eta + p[2] * sqrt(1.0 + p[3] * xi * xi + 0.5 * eta * eta);
}""",
    3: """template <class S> void cg_map(S xi, S eta, S zeta, double t, const double* p, S& x, S& y, S& z) {
  x = xi + p[0] * sin(p[1] * xi) * exp(-p[2] * xi) + p[3] * t;
}""",
}

MAIN = r"""
int main(int argc, char** argv) {
  // argv: dim scheme nhalo n0 n1 n2 t nparams p... outfile
  cgg::GD D;
  D.dim = atoi(argv[1]); D.scheme = atoi(argv[2]); D.nhalo = atoi(argv[3]); D.what = 3;
  long long nc = 1;
  for (int d = 0; d < 3; ++d) { D.nw[d] = d < D.dim ? atoll(argv[4 + d]) + 2 * D.nhalo : 1; D.w0[d] = 0; nc *= D.nw[d]; }
  const double t = atof(argv[7]);
  const int np = atoi(argv[8]);
  std::vector<double> prm(np > 0 ? np : 1);
  for (int i = 0; i < np; ++i) prm[i] = atof(argv[9 + i]);
  const int N = D.dim, KM = N * N + 1, KC = N * N;
  std::vector<double> cf(nc * KM), ci(nc * KM), ff(3 * nc * KM), fi(3 * nc * KM), fc(3 * nc * KC);
  cgg::GOut O = {};
  O.cell_fwd = cf.data(); O.cell_inv = ci.data();
  for (int a = 0; a < N; ++a) { O.face_fwd[a] = ff.data() + a * nc * KM; O.face_inv[a] = fi.data() + a * nc * KM; O.face_cons[a] = fc.data() + a * nc * KC; }
  cgg_cell_f64(D, t, prm.data(), O);
  cgg_face_f64(D, t, prm.data(), O);
  FILE* f = fopen(argv[9 + np], "wb");
  fwrite(cf.data(), 8, nc * KM, f); fwrite(ci.data(), 8, nc * KM, f);
  fwrite(ff.data(), 8, (size_t)N * nc * KM, f); fwrite(fi.data(), 8, (size_t)N * nc * KM, f); fwrite(fc.data(), 8, (size_t)N * nc * KC, f);
  fclose(f);
  return 0;
}
"""

CASES = {1: ((5, 4, 3), [0.3, 0.4, 0.01, 0.001, 0.02]), 2: ((9, 7), [0.2, 0.05, 0.3, 0.1]), 3: ((17,), [0.1, 0.6, 0.02, 0.5])}


@pytest.mark.parametrize("scheme", ["endpoint", "gradient", "curvature"])
@pytest.mark.parametrize("map_id", [1, 2, 3])
def test_generic_kernels_on_the_host_match_the_oracle(map_id, scheme):
    from oracle import oracle as O
    hdr = open(os.path.join(ROOT, "curvilineargrids.jl_b200", "csrc", "cg_generic_prelude.h")).read()
    prelude = re.search(r'GENERIC_PRELUDE = R"CGSRC\((.*?)\)CGSRC"', hdr, re.S).group(1)
    kernels = re.search(r'GENERIC_KERNELS = R"CGSRC\((.*?)\)CGSRC"', hdr, re.S).group(1)
    celldims, prm = CASES[map_id]
    h, t = 2, 0.7
    N = len(celldims)
    with tempfile.TemporaryDirectory() as d:
        cpp, exe, out = os.path.join(d, "g.cpp"), os.path.join(d, "g"), os.path.join(d, "o.bin")
        open(cpp, "w").write(SHIM + prelude + MAPS[map_id] + kernels + MAIN)
        subprocess.check_call(["g++", "-O1", "-std=c++17", "-w", cpp, "-o", exe])
        dims = list(celldims) + [1] * (3 - N)
        args = [exe, str(N), str({"endpoint": 0, "gradient": 1, "curvature": 2}[scheme]), str(h)] + [str(v) for v in dims] + \
               [repr(t), str(len(prm))] + [repr(float(v)) for v in prm] + [out]
        subprocess.check_call(args)
        raw = np.fromfile(out, dtype=np.float64)
    nc = int(np.prod([c + 2 * h for c in celldims]))
    KM, KC = N * N + 1, N * N
    o = 0
    got = {}
    for key, cnt, shape in (("cell_fwd", nc * KM, (nc, KM)), ("cell_inv", nc * KM, (nc, KM)), ("face_fwd", N * nc * KM, (N, nc, KM)),
                            ("face_inv", N * nc * KM, (N, nc, KM)), ("face_cons", N * nc * KC, (N, nc, KC))):
        got[key] = raw[o:o + cnt].reshape(shape)
        o += cnt
    ref = O.funcmap_eval(N, map_id, prm, t, celldims, h, scheme)
    got["nfull"] = ref["nfull"]
    assert_metrics_close(got, ref, rtol=RTOL_F64, label=f"host build of the generic kernels, map {map_id} {scheme}")


def _run_host_build(src_map, N, scheme_id, h, celldims, t, prm):
    hdr = open(os.path.join(ROOT, "curvilineargrids.jl_b200", "csrc", "cg_generic_prelude.h")).read()
    prelude = re.search(r'GENERIC_PRELUDE = R"CGSRC\((.*?)\)CGSRC"', hdr, re.S).group(1)
    kernels = re.search(r'GENERIC_KERNELS = R"CGSRC\((.*?)\)CGSRC"', hdr, re.S).group(1)
    with tempfile.TemporaryDirectory() as d:
        cpp, exe, out = os.path.join(d, "g.cpp"), os.path.join(d, "g"), os.path.join(d, "o.bin")
        open(cpp, "w").write(SHIM + prelude + src_map + kernels + MAIN)
        subprocess.check_call(["g++", "-O1", "-std=c++17", "-w", cpp, "-o", exe])
        dims = list(celldims) + [1] * (3 - N)
        subprocess.check_call([exe, str(N), str(scheme_id), str(h)] + [str(v) for v in dims] + [repr(float(t)), str(len(prm))] +
                              [repr(float(v)) for v in prm] + [out])
        raw = np.fromfile(out, dtype=np.float64)
    nc = int(np.prod([c + 2 * h for c in celldims]))
    KM, KC = N * N + 1, N * N
    o, got = 0, {}
    for key, cnt, shape in (("cell_fwd", nc * KM, (nc, KM)), ("cell_inv", nc * KM, (nc, KM)), ("face_fwd", N * nc * KM, (N, nc, KM)),
                            ("face_inv", N * nc * KM, (N, nc, KM)), ("face_cons", N * nc * KC, (N, nc, KC))):
        got[key] = raw[o:o + cnt].reshape(shape)
        o += cnt
    return got


@pytest.mark.parametrize("name", ["readme2d", "wavy3d_t", "rtp3d"])
def test_term_lists_as_generated_source_match_the_oracle_on_the_host(name):
    """workloads.terms_to_cuda_source turns any separable term list into `cg_map` source: the reference's own test mappings
    (README 2-D mapping, the time-dependent wavy 3-D mapping of BASELINE config 4, the spherical (r, theta, phi) mapping) run
    through the generic kernels — built for the host — and agree with the oracle's evaluation of the term list."""
    import curvigrid_b200 as cg
    from oracle import oracle as O
    wl = cg.workloads
    terms, celldims, h = {"readme2d": (wl.readme_wavy_2d_terms(), (8, 7), 2),
                          "wavy3d_t": (wl.wavy_3d_terms((5, 4, 3), time_amp=0.1), (5, 4, 3), 2),
                          "rtp3d": (wl.spherical_rtp_3d_terms((4, 4, 4)), (4, 4, 4), 2)}[name]
    t = 0.37
    got = _run_host_build(wl.terms_to_cuda_source(terms), len(celldims), 2, h, celldims, t, [])
    ref = O.mapped_eval(terms, t, celldims, h, "curvature")
    got["nfull"] = ref["nfull"]
    assert_metrics_close(got, ref, rtol=RTOL_F64, label=f"{name}: generated source on the host vs oracle")

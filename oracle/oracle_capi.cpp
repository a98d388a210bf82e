// ORACLE — TEST INFRASTRUCTURE ONLY.  Never linked into the product library.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
// reference legs may load liboracle.so.
//
// C entry points (ctypes) around metric_oracle.hpp: the fill loops of
// /root/reference/src/grids/unified_grids/metrics.jl:2055-2334 (cell + face
// storage over `cell.full`), the coordinate kernels of generation.jl:187-309
// and the GCL functional of /root/reference/src/grids/gcl.jl:20-117.
//
// Index conventions: 0-based full cell index
// (i,j,k), ξ = (i+1) + global_offset − nhalo + ½; storage index I of
// face[axis] holds the face I+½ (high side).
#include <omp.h>

#include <cstdint>
#include <cstring>
#include <vector>

#include "legacy_meg6.hpp"
#include "metric_oracle.hpp"

using namespace cgo;

namespace {

template <int N, int SCH, class M>
void eval_cells(const M& m, double t, const int64_t* nfull, int nhalo, const int64_t* goff, int64_t nsample,
                const int64_t* sample, int what, double* cell_fwd, double* cell_inv, double* face_fwd,
                double* face_inv, double* face_cons) {
  const int KM = N * N + 1, KC = N * N;
  int64_t ntot = 1;
  for (int d = 0; d < N; ++d) ntot *= nfull[d];
  const int64_t nout = sample ? nsample : ntot;
#pragma omp parallel for schedule(dynamic, 4)
  for (int64_t s = 0; s < nout; ++s) {
    int64_t lin = sample ? sample[s] : s;
    int64_t idx[3] = {0, 0, 0};
    int64_t r = lin;
    for (int d = 0; d < N; ++d) {
      idx[d] = r % nfull[d];
      r /= nfull[d];
    }
    Pt<real> p = {0.0, 0.0, 0.0};
    for (int d = 0; d < N; ++d) p[d] = real(idx[d] + 1 + goff[d] - nhalo) + 0.5;
    // results are computed in `real` (double, or the extended type of liboracle_ld / liboracle_q) and rounded to
    // double on store
    real bf[N * N + 1], bi[N * N + 1], bc[N * N];
    if (what & 1) {
      cell_metric<N>(m, t, p, bf, bi);
      for (int k = 0; k < KM; ++k) {
        cell_fwd[s * KM + k] = (double)bf[k];
        cell_inv[s * KM + k] = (double)bi[k];
      }
    }
    if (what & 2) {
      for (int ax = 0; ax < N; ++ax) {
        if constexpr (SCH == 3) {  // ADThomasLombardMetric (3-D only; 1-D / 2-D fall back to Curvature, metrics.jl:1669-1673)
          if constexpr (N == 3) face_metric_adtl3(m, t, ax, p, bf, bi, bc);
          else face_metric<N, 2>(m, t, ax, p, bf, bi, bc);
        } else {
          face_metric<N, SCH>(m, t, ax, p, bf, bi, bc);
        }
        for (int k = 0; k < KM; ++k) {
          face_fwd[(ax * nout + s) * KM + k] = (double)bf[k];
          face_inv[(ax * nout + s) * KM + k] = (double)bi[k];
        }
        for (int k = 0; k < KC; ++k) face_cons[(ax * nout + s) * KC + k] = (double)bc[k];
      }
    }
  }
}

template <class M>
int dispatch(const M& m, int dim, int scheme, double t, const int64_t* nfull, int nhalo, const int64_t* goff,
             int64_t nsample, const int64_t* sample, int what, double* cf, double* ci, double* ff, double* fi,
             double* fc) {
#define CGO_CASE(NN, SS) \
  if (dim == NN && scheme == SS) { eval_cells<NN, SS>(m, t, nfull, nhalo, goff, nsample, sample, what, cf, ci, ff, fi, fc); return 0; }
  CGO_CASE(1, 0) CGO_CASE(1, 1) CGO_CASE(1, 2)
  CGO_CASE(2, 0) CGO_CASE(2, 1) CGO_CASE(2, 2)
  CGO_CASE(3, 0) CGO_CASE(3, 1) CGO_CASE(3, 2)
  CGO_CASE(1, 3) CGO_CASE(2, 3) CGO_CASE(3, 3)
#undef CGO_CASE
  return 1;
}

// generation.jl:187-309: which = 0 node (ξ = I − h), 1 centroid (+½ on all
// axes), 2+axis high-side face centre (+½ on all axes and +½ more on `axis`).
template <class M>
void eval_coords(const M& m, int dim, double t, const int64_t* ncell_full, int nhalo, const int64_t* goff, int which,
                 double* out /* [coord][points] */) {
  int64_t np[3] = {1, 1, 1};
  for (int d = 0; d < dim; ++d) np[d] = ncell_full[d] + (which == 0 ? 1 : 0);
  const int64_t ntot = np[0] * np[1] * np[2];
#pragma omp parallel for schedule(static)
  for (int64_t lin = 0; lin < ntot; ++lin) {
    int64_t idx[3], r = lin;
    for (int d = 0; d < 3; ++d) {
      idx[d] = r % np[d];
      r /= np[d];
    }
    Pt<real> p = {0.0, 0.0, 0.0};
    for (int d = 0; d < dim; ++d) {
      p[d] = real(idx[d] + 1 + goff[d] - nhalo);
      if (which >= 1) p[d] += 0.5;
      if (which >= 2 && d == which - 2) p[d] += 0.5;
    }
    for (int c = 0; c < dim; ++c) out[c * ntot + lin] = (double)m.template coord<real>(c, t, p);
  }
}

TermMap make_termmap(int dim, int nterms, const double* table, std::vector<Term>& storage) {
  storage.resize(nterms);
  for (int m = 0; m < nterms; ++m) {
    const double* r = table + 15 * m;
    Term& T = storage[m];
    T.coord = int(r[0]);
    T.c0 = r[1]; T.c1 = r[2]; T.c2 = r[3]; T.w = r[4]; T.ph = r[5];
    for (int d = 0; d < 3; ++d) {
      T.f[d].kind = int(r[6 + 3 * d]);
      T.f[d].a = r[7 + 3 * d];
      T.f[d].b = r[8 + 3 * d];
    }
  }
  return TermMap{dim, nterms, storage.data()};
}

}  // namespace

extern "C" {

int cgo_num_threads(void) { return omp_get_max_threads(); }
// bits of the significand of the base real type: 53 (liboracle.so), 64 (liboracle_ld.so), 113 (liboracle_q.so)
int cgo_real_digits(void) {
#if defined(CGO_REAL_IS_QUAD)
  return 113;
#elif defined(CGO_REAL_IS_LD)
  return 64;
#else
  return 53;
#endif
}
void cgo_set_num_threads(int n) { if (n > 0) omp_set_num_threads(n); }

// DiscreteGrid(x[,y[,z]], nhalo): x,y,z are INTERIOR node arrays (halo already
// stripped as discrete_grid.jl:55-61 does), column-major, nn[d] nodes.
// Outputs: sample == NULL → full arrays in Julia layout; else packed [nsample].
// face_* buffers are [axis][cell][K].
int cgo_discrete_eval(int dim, const double* x, const double* y, const double* z, const int64_t* nn, int nhalo,
                      int scheme, int64_t nsample, const int64_t* sample, int what, double* cell_fwd,
                      double* cell_inv, double* face_fwd, double* face_inv, double* face_cons) {
  DiscreteMap m;
  m.N = dim;
  m.a[0] = x; m.a[1] = y; m.a[2] = z;
  int64_t nfull[3] = {1, 1, 1}, goff[3] = {0, 0, 0};
  for (int d = 0; d < 3; ++d) m.n[d] = d < dim ? nn[d] : 1;
  for (int d = 0; d < dim; ++d) nfull[d] = nn[d] - 1 + 2 * nhalo;
  return dispatch(m, dim, scheme, 0.0, nfull, nhalo, goff, nsample, sample, what, cell_fwd, cell_inv, face_fwd,
                  face_inv, face_cons);
}

// MappedGrid with a term-list mapping (15 doubles per term:
// coord, c0, c1, c2, w, ph, then (kind,a,b) for each of 3 axes).
// goff = first(global_cell_indices) − 1 per axis (mapped_grid.jl:344-379).
int cgo_mapped_eval(int dim, int nterms, const double* termtable, double t, const int64_t* ncell, int nhalo,
                    const int64_t* goff, int scheme, int64_t nsample, const int64_t* sample, int what,
                    double* cell_fwd, double* cell_inv, double* face_fwd, double* face_inv, double* face_cons) {
  std::vector<Term> st;
  TermMap m = make_termmap(dim, nterms, termtable, st);
  int64_t nfull[3] = {1, 1, 1}, g[3] = {0, 0, 0};
  for (int d = 0; d < dim; ++d) {
    nfull[d] = ncell[d] + 2 * nhalo;
    if (goff) g[d] = goff[d];
  }
  return dispatch(m, dim, scheme, t, nfull, nhalo, g, nsample, sample, what, cell_fwd, cell_inv, face_fwd, face_inv,
                  face_cons);
}

// MappedGrid with one of the non-separable test mappings of metric_oracle.hpp (FuncMap)
int cgo_funcmap_eval(int dim, int id, const double* prm, double t, const int64_t* ncell, int nhalo, const int64_t* goff,
                     int scheme, int64_t nsample, const int64_t* sample, int what, double* cell_fwd, double* cell_inv,
                     double* face_fwd, double* face_inv, double* face_cons) {
  FuncMap m{dim, id, prm};
  int64_t nfull[3] = {1, 1, 1}, g[3] = {0, 0, 0};
  for (int d = 0; d < dim; ++d) {
    nfull[d] = ncell[d] + 2 * nhalo;
    if (goff) g[d] = goff[d];
  }
  return dispatch(m, dim, scheme, t, nfull, nhalo, g, nsample, sample, what, cell_fwd, cell_inv, face_fwd, face_inv,
                  face_cons);
}
int cgo_funcmap_coords(int dim, int id, const double* prm, double t, const int64_t* ncell, int nhalo, const int64_t* goff,
                       int which, double* out) {
  FuncMap m{dim, id, prm};
  int64_t nfull[3] = {1, 1, 1}, g[3] = {0, 0, 0};
  for (int d = 0; d < dim; ++d) {
    nfull[d] = ncell[d] + 2 * nhalo;
    if (goff) g[d] = goff[d];
  }
  eval_coords(m, dim, t, nfull, nhalo, g, which, out);
  return 0;
}

int cgo_discrete_coords(int dim, const double* x, const double* y, const double* z, const int64_t* nn, int nhalo,
                        int which, double* out) {
  DiscreteMap m;
  m.N = dim;
  m.a[0] = x; m.a[1] = y; m.a[2] = z;
  int64_t nfull[3] = {1, 1, 1}, goff[3] = {0, 0, 0};
  for (int d = 0; d < 3; ++d) m.n[d] = d < dim ? nn[d] : 1;
  for (int d = 0; d < dim; ++d) nfull[d] = nn[d] - 1 + 2 * nhalo;
  eval_coords(m, dim, 0.0, nfull, nhalo, goff, which, out);
  return 0;
}

int cgo_mapped_coords(int dim, int nterms, const double* termtable, double t, const int64_t* ncell, int nhalo,
                      const int64_t* goff, int which, double* out) {
  std::vector<Term> st;
  TermMap m = make_termmap(dim, nterms, termtable, st);
  int64_t nfull[3] = {1, 1, 1}, g[3] = {0, 0, 0};
  for (int d = 0; d < dim; ++d) {
    nfull[d] = ncell[d] + 2 * nhalo;
    if (goff) g[d] = goff[d];
  }
  eval_coords(m, dim, t, nfull, nhalo, g, which, out);
  return 0;
}

// gcl.jl:20-117 on cell.domain: I_c[idx] = Σ_axis (Ghat_axis[idx][axis,c] − Ghat_axis[idx−e_axis][axis,c]).
// cons = [axis][cell.full][N*N]; returns max |I_c| per c in out[0..dim).
int cgo_gcl_max(int dim, const double* cons, const int64_t* nfull, int nhalo, double* out) {
  const int K = dim * dim;
  int64_t st[3] = {1, 1, 1}, ntot = 1;
  for (int d = 0; d < dim; ++d) { st[d] = ntot; ntot *= nfull[d]; }
  int64_t lo[3] = {0, 0, 0}, hi[3] = {1, 1, 1};
  for (int d = 0; d < dim; ++d) { lo[d] = nhalo; hi[d] = nfull[d] - nhalo; }
  for (int c = 0; c < dim; ++c) out[c] = 0.0;
  for (int64_t k = lo[2]; k < hi[2]; ++k)
    for (int64_t j = lo[1]; j < hi[1]; ++j)
      for (int64_t i = lo[0]; i < hi[0]; ++i) {
        int64_t lin = i * st[0] + (dim > 1 ? j * st[1] : 0) + (dim > 2 ? k * st[2] : 0);
        for (int c = 0; c < dim; ++c) {
          double acc = 0.0;
          for (int ax = 0; ax < dim; ++ax) {
            const double* A = cons + ax * ntot * K;
            acc += A[lin * K + ax + dim * c] - A[(lin - st[ax]) * K + ax + dim * c];
          }
          if (std::fabs(acc) > out[c]) out[c] = std::fabs(acc);
        }
      }
  return 0;
}

// Legacy CurvilinearGrid3D(x, y, z, :meg6 | :meg6_symmetric) metric update (legacy_meg6.hpp).
// x,y,z: node arrays, nnode = their extents (interior nodes, or node.full when halo_coords_included).
// cell = [28][ncell.full], edge = [3][19][ncell.full], cent = [3][ncell.full] (may be NULL).
int cgo_legacy_meg6_3d(const double* x, const double* y, const double* z, const int64_t* nnode, int nhalo,
                       int halo_coords_included, int symmetric, double* cell, double* edge, double* cent) {
  cgl::legacy_meg6_3d(x, y, z, nnode, nhalo, halo_coords_included != 0, symmetric != 0, cell, edge, cent);
  return 0;
}

// Legacy CurvilinearGrid2D(x, y, :meg6 | :meg6_symmetric) metric update: cell = [13][ncell.full],
// edge = [2][9][ncell.full], cent = [2][ncell.full] (may be NULL).
int cgo_legacy_meg6_2d(const double* x, const double* y, const int64_t* nnode, int nhalo, int halo_coords_included,
                       double* cell, double* edge, double* cent) {
  cgl::legacy_meg6_2d(x, y, nnode, nhalo, halo_coords_included != 0, cell, edge, cent);
  return 0;
}

// Legacy CurvilinearGrid1D(x, :meg6) metric update: cell = [4][ncell.full], edge = [3][ncell.full], cent = [ncell.full].
int cgo_legacy_meg6_1d(const double* x, const int64_t* nnode, int nhalo, int halo_coords_included, double* cell,
                       double* edge, double* cent) {
  cgl::legacy_meg6_1d(x, nnode, nhalo, halo_coords_included != 0, cell, edge, cent);
  return 0;
}

// stencil probes for the pins (derivative_stencils.jl:34-45, 97-175, 180-193)
double cgo_legacy_central6(const double* v) { return cgl::central6(v[0], v[1], v[2], v[3], v[4], v[5]); }
void cgo_legacy_mixed(const double* v, double* lo3, double* hi3) {
  cgl::mixed_lo(v, lo3);
  cgl::mixed_hi(v, hi3);
}

}  // extern "C"

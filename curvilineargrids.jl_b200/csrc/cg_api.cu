// libcurvigrid_cuda.so — C ABI (include/curvigrid_cuda.h) over the CUDA kernels.
// Host-side handle, device memory, stream ordering, status codes.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>  // types only: the NCCL entry points are resolved at run time (see nccl_api)

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <type_traits>
#include <vector>

#include "../../include/curvigrid_cuda.h"
#include "cg_kernels_ref.cuh"
#include "cg_kernels_tiled.cuh"
#include "cg_mapped.cuh"
#include "cg_mapped_split.cuh"
#include "cg_generic.cuh"

// NVTX ranges around the entry points (tracing).  NVTX v3 is header-only and resolves the profiler's
// injection library lazily: without a profiler attached a range costs one predictable branch.
#include <nvtx3/nvToolsExt.h>
namespace {
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
};
}  // namespace
#define CG_NVTX(name) NvtxRange cg_nvtx_range_(name)

namespace {

thread_local std::string g_err;
std::atomic<uint64_t> g_launches{0};

int fail(int code, const char* fmt, ...) {
  char buf[4096];  // room for an NVRTC log
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}
#define CG_CUDA(expr)                                                                        \
  do {                                                                                       \
    cudaError_t e_ = (expr);                                                                 \
    if (e_ != cudaSuccess)                                                                   \
      return fail(e_ == cudaErrorMemoryAllocation ? CG_ERR_OOM : CG_ERR_CUDA, "%s: %s", #expr, \
                  cudaGetErrorString(e_));                                                   \
  } while (0)

struct Buf {
  void* p = nullptr;
  size_t bytes = 0;
  bool owned = false;
};

}  // namespace

struct cg_grid {
  int device = 0, dim = 0, dtype = 0, nhalo = 0, scheme = 2, profile = 0;
  bool adtl = false;  // ADThomasLombardMetric (3-D MappedGrid): Curvature evaluation + k_mapped_adtl_faces
  int64_t ncell[3] = {1, 1, 1};
  int64_t nw[3] = {1, 1, 1}, w0[3] = {0, 0, 0}, goff[3] = {0, 0, 0};
  int kind = 0;  // 0 unset, 1 discrete, 2 mapped
  int kernel_opt = 1;
  // node source (interior node planes), either owned or caller's device memory
  Buf src[3];
  int64_t src_origin[3] = {0, 0, 0}, src_count[3] = {1, 1, 1};
  Buf ext[3];
  bool ext_valid = false;
  Buf cell_fwd, cell_inv, face_fwd[3], face_inv[3], face_cons[3];
  Buf node[3], centroid[3], facec[3][3], cjac, cactive[3];
  bool valid_cell = false, valid_face = false, valid_coords = false;
  unsigned long long* gcl_bits = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  bool timed = false;
  cudaStream_t copy_stream = nullptr;  // cg_grid_update_from_host: H2D pieces
  cudaEvent_t copy_ev[2] = {nullptr, nullptr};
  cg::Fork fork;                       // helper streams of the small-grid face launches (per handle)
  cudaStream_t comm_stream = nullptr;  // cg_halo_exchange_*: ncclSend / ncclRecv of node planes
  cudaEvent_t comm_ev[2] = {nullptr, nullptr};
  Buf seam;                            // cg_grid_gcl_max_dist: the neighbour's last conserved plane
  std::vector<double> terms;
  cg::MappedTables mtab;
  cg::GenericMap gmap;  // generic (NVRTC-compiled) mapping: cg_grid_set_mapping_source
  // GCL residuals accumulated plane by plane during cg_grid_update_from_host (CG_WHAT_GCL): valid while no other
  // evaluation has rewritten the face metrics (face_epoch counts them)
  uint64_t face_epoch = 0, gcl_epoch = ~0ull;
  double t = 0.0;

  size_t esize() const { return dtype == CG_F64 ? 8 : 4; }
  int64_t ncells() const { return nw[0] * nw[1] * nw[2]; }
  int64_t nnodes() const {
    int64_t n = 1;
    for (int d = 0; d < dim; ++d) n *= nw[d] + 1;
    return n;
  }
  // padded ("ext") node array: nw + 4 nodes per active axis; the row pitch is a multiple of 16 bytes (the eta / zeta
  // kernels fetch row segments with 16-byte cp.async) and 64 values of slack follow the last row
  int64_t pitch0() const { return dim > 1 ? (nw[0] + 4 + 3) / 4 * 4 : nw[0] + 4; }
  int64_t next() const {
    int64_t n = pitch0();
    for (int d = 1; d < dim; ++d) n *= nw[d] + 4;
    return n + 64;
  }
  cg::Dims dims() const {
    cg::Dims D;
    D.dim = dim;
    D.nhalo = nhalo;
    int64_t s = 1;
    for (int d = 0; d < 3; ++d) {
      D.nw[d] = nw[d];
      D.ne[d] = d < dim ? nw[d] + 4 : 1;
      D.se[d] = s;
      s *= d == 0 ? pitch0() : D.ne[d];
      D.w0[d] = w0[d] + goff[d];
      D.gn[d] = ncell[d];
    }
    D.r0 = 0;
    D.r1 = nw[2];
    return D;
  }
};

namespace {

int ensure(cg_grid* g, Buf& b, size_t bytes) {
  if (b.p) {
    if (b.bytes < bytes) return fail(CG_ERR_ARG, "bound output buffer too small (%zu < %zu bytes)", b.bytes, bytes);
    return CG_OK;
  }
  CG_CUDA(cudaSetDevice(g->device));
  CG_CUDA(cudaMalloc(&b.p, bytes));
  b.bytes = bytes;
  b.owned = true;
  return CG_OK;
}
void release(Buf& b) {
  if (b.p && b.owned) cudaFree(b.p);
  b = Buf();
}

Buf* select(cg_grid* g, int array, int axis, size_t* bytes) {
  const int N = g->dim;
  const size_t es = g->esize();
  const size_t nc = (size_t)g->ncells();
  if (array != CG_ARR_FACE_COORD && array != CG_ARR_CELL_FORWARD && array != CG_ARR_CELL_INVERSE &&
      array != CG_ARR_COMPACT_J && (axis < 0 || axis >= N))
    return nullptr;
  switch (array) {
    case CG_ARR_CELL_FORWARD: *bytes = nc * (N * N + 1) * es; return &g->cell_fwd;
    case CG_ARR_CELL_INVERSE: *bytes = nc * (N * N + 1) * es; return &g->cell_inv;
    case CG_ARR_FACE_FORWARD: *bytes = nc * (N * N + 1) * es; return &g->face_fwd[axis];
    case CG_ARR_FACE_INVERSE: *bytes = nc * (N * N + 1) * es; return &g->face_inv[axis];
    case CG_ARR_FACE_CONSERVED: *bytes = nc * (N * N) * es; return &g->face_cons[axis];
    case CG_ARR_NODE_COORD: *bytes = (size_t)g->nnodes() * es; return &g->node[axis];
    case CG_ARR_CENTROID_COORD: *bytes = nc * es; return &g->centroid[axis];
    case CG_ARR_FACE_COORD:
      if (axis < 0 || axis >= 3 * N || axis % 3 >= N) return nullptr;
      *bytes = nc * es;
      return &g->facec[axis / 3][axis % 3];
    case CG_ARR_COMPACT_J: *bytes = nc * es; return &g->cjac;
    case CG_ARR_COMPACT_ACTIVE: *bytes = nc * N * es; return &g->cactive[axis];
  }
  return nullptr;
}

void invalidate_for(cg_grid* g, int array) {
  switch (array) {
    case CG_ARR_CELL_FORWARD: case CG_ARR_CELL_INVERSE: case CG_ARR_COMPACT_J: g->valid_cell = false; break;
    case CG_ARR_FACE_FORWARD: case CG_ARR_FACE_INVERSE: case CG_ARR_FACE_CONSERVED: case CG_ARR_COMPACT_ACTIVE:
      g->valid_face = false; break;
    default: g->valid_coords = false;
  }
}

int launch_blocks(int64_t n, int threads) {
  int64_t b = (n + threads - 1) / threads;
  const int64_t cap = 148LL * 64;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

template <class T>
int pad_nodes(cg_grid* g, cudaStream_t st, int64_t e_lo = 0, int64_t e_hi = -1) {
  const size_t bytes = (size_t)g->next() * sizeof(T);
  for (int q = 0; q < g->dim; ++q) {
    int rc = ensure(g, g->ext[q], bytes);
    if (rc) return rc;
  }
  cg::Dims D = g->dims();
  // discrete grids ignore global_offset: nodes are addressed by window position only
  for (int d = 0; d < 3; ++d) D.w0[d] = g->w0[d];
  cg::NodeSrc<T> S;
  int64_t s = 1;
  for (int d = 0; d < 3; ++d) {
    S.p[d] = (const T*)g->src[d].p;
    S.s0[d] = g->src_origin[d];
    S.sn[d] = g->src_count[d];
    S.st[d] = s;
    s *= g->src_count[d];
  }
  {
    if (e_hi < 0) e_hi = D.ne[2];  // padded planes [e_lo, e_hi) of axis 2 (3-D only; the whole array by default)
    const unsigned rows = (unsigned)(D.ne[1] * (e_hi - e_lo));
    T *e0 = (T*)g->ext[0].p, *e1 = (T*)g->ext[1].p, *e2 = (T*)g->ext[2].p;
    if (rows == 0) return CG_OK;
    if (D.dim == 3) cg::k_pad_nodes<T, 3><<<rows, 256, 0, st>>>(D, S, e0, e1, e2, (int)e_lo);
    else if (D.dim == 2) cg::k_pad_nodes<T, 2><<<rows, 256, 0, st>>>(D, S, e0, e1, e2, 0);
    else cg::k_pad_nodes<T, 1><<<rows, 256, 0, st>>>(D, S, e0, e1, e2, 0);
  }
  g_launches++;
  CG_CUDA(cudaGetLastError());
  g->ext_valid = true;
  return CG_OK;
}


template <class T>
int bind_coord_outputs(cg_grid* g, cg::CoordOut<T>& C) {
  const int N = g->dim;
  memset(&C, 0, sizeof C);
  size_t b;
  for (int q = 0; q < N; ++q) {
    Buf* nb = select(g, CG_ARR_NODE_COORD, q, &b);
    int rc = ensure(g, *nb, b);
    if (rc) return rc;
    C.node[q] = (T*)nb->p;
    Buf* cb = select(g, CG_ARR_CENTROID_COORD, q, &b);
    rc = ensure(g, *cb, b);
    if (rc) return rc;
    C.centroid[q] = (T*)cb->p;
    for (int a = 0; a < N; ++a) {
      Buf* fb = select(g, CG_ARR_FACE_COORD, 3 * a + q, &b);
      rc = ensure(g, *fb, b);
      if (rc) return rc;
      C.face[a][q] = (T*)fb->p;
    }
  }
  return CG_OK;
}

template <class T>
int bind_metric_outputs(cg_grid* g, int what, cg::MetricOut<T>& O) {
  const int N = g->dim;
  memset(&O, 0, sizeof O);
  size_t b;
  auto get = [&](int arr, int ax, T** dst) -> int {
    Buf* p = select(g, arr, ax, &b);
    int rc = ensure(g, *p, b);
    if (rc) return rc;
    *dst = (T*)p->p;
    return CG_OK;
  };
  int rc = CG_OK;
  if (g->profile == CG_PROFILE_FULL) {
    if (what & CG_WHAT_CELL) {
      if ((rc = get(CG_ARR_CELL_FORWARD, 0, &O.cell_fwd))) return rc;
      if ((rc = get(CG_ARR_CELL_INVERSE, 0, &O.cell_inv))) return rc;
    }
    if (what & CG_WHAT_FACE)
      for (int a = 0; a < N; ++a) {
        if ((rc = get(CG_ARR_FACE_FORWARD, a, &O.face_fwd[a]))) return rc;
        if ((rc = get(CG_ARR_FACE_INVERSE, a, &O.face_inv[a]))) return rc;
        if ((rc = get(CG_ARR_FACE_CONSERVED, a, &O.face_cons[a]))) return rc;
      }
  } else {
    if (what & CG_WHAT_CELL)
      if ((rc = get(CG_ARR_COMPACT_J, 0, &O.cjac))) return rc;
    if (what & CG_WHAT_FACE)
      for (int a = 0; a < N; ++a)
        if ((rc = get(CG_ARR_COMPACT_ACTIVE, a, &O.cactive[a]))) return rc;
  }
  return CG_OK;
}

int start_timer(cg_grid* g, cudaStream_t st) {
  if (!g->ev0) {
    CG_CUDA(cudaEventCreate(&g->ev0));
    CG_CUDA(cudaEventCreate(&g->ev1));
  }
  CG_CUDA(cudaEventRecord(g->ev0, st));
  return CG_OK;
}

// MappedGrid with a generic (non-separable) mapping: the NVRTC-compiled kernels of cg_generic_prelude.h
template <class T>
int eval_generic(cg_grid* g, int what, cudaStream_t st) {
  const int fi = sizeof(T) == 8 ? 0 : 1;
  cg::GD D;
  D.dim = g->dim;
  D.nhalo = g->nhalo;
  D.scheme = g->scheme;
  D.what = what;
  for (int d = 0; d < 3; ++d) {
    D.nw[d] = g->nw[d];
    D.w0[d] = g->w0[d] + g->goff[d];
  }
  double t = g->t;
  const double* prm = g->gmap.d_params;
  int rc;
  if (what & CG_WHAT_COORDS) {
    cg::CoordOut<T> C;
    if ((rc = bind_coord_outputs<T>(g, C))) return rc;
    cg::GCoord GC;
    for (int a = 0; a < 3; ++a) {
      GC.node[a] = C.node[a];
      GC.centroid[a] = C.centroid[a];
      for (int q = 0; q < 3; ++q) GC.face[a][q] = C.face[a][q];
    }
    void* args[] = {&D, &t, &prm, &GC};
    const std::string e = cg::generic_launch(g->gmap.coords[fi], g->nnodes(), 128, args, st);
    if (!e.empty()) return fail(CG_ERR_CUDA, "generic mapping, coordinates: %s", e.c_str());
    g_launches++;
    g->valid_coords = true;
  }
  if (!(what & (CG_WHAT_CELL | CG_WHAT_FACE))) return CG_OK;
  cg::MetricOut<T> O;
  if ((rc = bind_metric_outputs<T>(g, what, O))) return rc;
  if ((rc = start_timer(g, st))) return rc;
  cg::GOut GO;
  GO.cell_fwd = O.cell_fwd;
  GO.cell_inv = O.cell_inv;
  GO.cjac = O.cjac;
  for (int a = 0; a < 3; ++a) {
    GO.face_fwd[a] = O.face_fwd[a];
    GO.face_inv[a] = O.face_inv[a];
    GO.face_cons[a] = O.face_cons[a];
    GO.cactive[a] = O.cactive[a];
  }
  void* args[] = {&D, &t, &prm, &GO};
  if (what & CG_WHAT_CELL) {
    const std::string e = cg::generic_launch(g->gmap.cell[fi], g->ncells(), 128, args, st);
    if (!e.empty()) return fail(CG_ERR_CUDA, "generic mapping, cell metrics: %s", e.c_str());
    g_launches++;
  }
  if (what & CG_WHAT_FACE) {
    const std::string e = cg::generic_launch(g->gmap.face[fi], g->ncells() * g->dim, 128, args, st);
    if (!e.empty()) return fail(CG_ERR_CUDA, "generic mapping, face metrics: %s", e.c_str());
    g_launches++;
  }
  CG_CUDA(cudaEventRecord(g->ev1, st));
  g->timed = true;
  if (what & CG_WHAT_CELL) g->valid_cell = true;
  if (what & CG_WHAT_FACE) { g->valid_face = true; g->face_epoch++; }
  return CG_OK;
}

// MappedGrid: 1-D operator tables (once per update) + the table-driven metric kernel
template <class T>
int eval_mapped(cg_grid* g, int what, cudaStream_t st) {
  const int N = g->dim;
  if (g->gmap.active) {
    if (g->adtl) return fail(CG_ERR_ARG, "ADThomasLombardMetric needs a term-list mapping (cg_grid_set_mapping_terms)");
    return eval_generic<T>(g, what, st);
  }
  cg::Dims D = g->dims();
  int rc = cg::mapped_prepare<T>(g->mtab, g->terms, g->t, D, g->scheme, st, g->adtl);
  if (rc == 1) return fail(CG_ERR_ARG, "mapping has too many terms (max %d)", cg::M_MAXT - 1);
  if (rc == 2) return fail(CG_ERR_ARG, "mapping has too many term pairs (max %d)", cg::M_MAXP);
  if (rc) return fail(CG_ERR_CUDA, "mapped_prepare: %s", cudaGetErrorString(cudaGetLastError()));
  g_launches += g->mtab.launches;
  g->mtab.launches = 0;
  cg::MTerms TM = g->mtab.terms;
  if (what & CG_WHAT_COORDS) {
    cg::CoordOut<T> C;
    if ((rc = bind_coord_outputs<T>(g, C))) return rc;
    if (cg::mapped_prepare_coords<T>(g->mtab, D, st)) return fail(CG_ERR_CUDA, "mapped_prepare_coords: %s", cudaGetErrorString(cudaGetLastError()));
    g_launches += g->mtab.launches;
    g->mtab.launches = 0;
    if (g->nw[1] + 1 > 65535 || g->nw[2] + 1 > 65535) return fail(CG_ERR_ARG, "MappedGrid coordinates: more than 65534 cells along eta / zeta");
    cg::launch_mapped_coords<T>(D, TM, (const T*)g->mtab.ctab, g->mtab.L + 1, C, st);
    g_launches++;
    CG_CUDA(cudaGetLastError());
    g->valid_coords = true;
  }
  if (!(what & (CG_WHAT_CELL | CG_WHAT_FACE))) return CG_OK;
  cg::MetricOut<T> O;
  if ((rc = bind_metric_outputs<T>(g, what, O))) return rc;
  if ((rc = start_timer(g, st))) return rc;
  const int64_t nc = g->ncells();
  const T* tab = (const T*)g->mtab.tab;
  const T* ftab = (const T*)g->mtab.ftab;
  bool adtl_done = false;
  if (N == 1)
    cg::k_mapped_metrics1d<T><<<launch_blocks(nc, 128), 128, 0, st>>>(D, TM, ftab, g->mtab.L, O, g->scheme, what);
  else if (N == 2)
    cg::k_mapped_metrics2d<T><<<launch_blocks(nc, 128), 128, 0, st>>>(D, TM, g->mtab.pairs, tab, ftab, g->mtab.L, O,
                                                                      g->scheme, what);
  else {
    const bool full_cf = (what & 3) == 3 && O.cell_fwd && O.cell_inv && O.face_cons[0] && O.face_fwd[0] && O.face_inv[0] &&
                         !O.cjac && !O.cactive[0];
    bool ok = true;
    if (full_cf && g->adtl) {
      // ADThomasLombardMetric: its own kernel (active rows + face-centre Jacobians, no reconstruction of inv F)
      ok = cg::launch_mapped_adtl3d_staged<T>(D, TM, g->mtab.pairs, tab, ftab, (const T*)g->mtab.fhalf, g->mtab.L, O, st);
      adtl_done = true;
    } else if (full_cf) {
      // debug switch CG_MAPPED_KERNEL: staged = the round-1 kernel (every lane multiplies all three table rows), uni = one
      // kernel with the warp-uniform products factored out; default = the four light kernels of cg_mapped_split.cuh
      static const int which = [] {
        const char* e = getenv("CG_MAPPED_KERNEL");
        return !e ? 0 : (std::string(e) == "staged" ? 2 : (std::string(e) == "uni" ? 1 : 0));
      }();
      ok = false;
      if (which == 0) {
        const int nl = cg::launch_mapped_metrics3d_split<T>(D, TM, g->mtab.pairs, tab, (const T*)g->mtab.tabx, ftab, g->mtab.L, O, g->scheme, st);
        if (nl < 0) return fail(CG_ERR_CUDA, "MappedGrid kernels: cannot set the shared-memory size");
        ok = nl > 0;
        if (ok) g_launches += nl - 1;
      }
      if (!ok && which <= 1 && TM.n >= 1 && g->mtab.pairs.n >= 1)
        ok = cg::launch_mapped_metrics3d_uni<T>(D, TM, g->mtab.pairs, tab, ftab, g->mtab.L, O, g->scheme, st);
      if (!ok) ok = cg::launch_mapped_metrics3d_staged<T>(D, TM, g->mtab.pairs, tab, ftab, g->mtab.L, O, g->scheme, st);
    } else {
      cg::k_mapped_metrics3d<T><<<launch_blocks(nc, 128), 128, 0, st>>>(D, TM, g->mtab.pairs, tab, ftab, g->mtab.L, O,
                                                                        g->scheme, what);
    }
    if (!ok) return fail(CG_ERR_CUDA, "MappedGrid kernel: cannot set the shared-memory size");
  }
  g_launches++;
  // other output configurations of ADThomasLombardMetric (face-only refresh, COMPACT): CurvatureCorrected
  // evaluation, then the face arrays are overwritten (COMPACT stores the active rows only: nothing to overwrite)
  if (g->adtl && !adtl_done && (what & CG_WHAT_FACE) && N == 3 && g->profile == CG_PROFILE_FULL) {
    cg::k_mapped_adtl_faces<T><<<launch_blocks(nc, 128), 128, 0, st>>>(D, TM, ftab, (const T*)g->mtab.fhalf,
                                                                       g->mtab.L + 1, O);
    g_launches++;
  }
  CG_CUDA(cudaGetLastError());
  CG_CUDA(cudaEventRecord(g->ev1, st));
  g->timed = true;
  if (what & CG_WHAT_CELL) g->valid_cell = true;
  if (what & CG_WHAT_FACE) { g->valid_face = true; g->face_epoch++; }
  return CG_OK;
}

template <class T>
int eval_impl(cg_grid* g, int what, cudaStream_t st, int64_t k_lo = 0, int64_t k_hi = -1, bool final = true) {
  const int N = g->dim;
  if (g->kind == 0) return fail(CG_ERR_STATE, "cg_grid_eval: no nodes or mapping set");
  if (g->kind == 2) return eval_mapped<T>(g, what, st);
  if (!g->ext_valid) return fail(CG_ERR_STATE, "cg_grid_eval: node data not available");
  cg::Dims D = g->dims();
  if (g->kind == 1)
    for (int d = 0; d < 3; ++d) D.w0[d] = g->w0[d];
  const T *ex = (const T*)g->ext[0].p, *ey = (const T*)g->ext[1].p, *ez = (const T*)g->ext[2].p;
  const bool ranged = k_hi >= 0;
  if (ranged) {
    D.r0 = k_lo;
    D.r1 = k_hi;
  }

  if (what & CG_WHAT_COORDS) {
    cg::CoordOut<T> C;
    int rcc = bind_coord_outputs<T>(g, C);
    if (rcc) return rcc;
    cg::k_coords<T><<<launch_blocks(g->nnodes(), 256), 256, 0, st>>>(D, ex, ey, ez, C);
    g_launches++;
    CG_CUDA(cudaGetLastError());
    g->valid_coords = true;
  }
  if (!(what & (CG_WHAT_CELL | CG_WHAT_FACE))) return CG_OK;

  cg::MetricOut<T> O;
  int rco = bind_metric_outputs<T>(g, what, O);
  if (rco) return rco;
  if ((rco = start_timer(g, st))) return rco;
  const int64_t nc = g->ncells();
  int nl = 0;
  if (N == 1) {
    cg::k_metrics1d_ref<T><<<launch_blocks(nc, 256), 256, 0, st>>>(D, ex, O, g->scheme, what);
    nl = 1;
  } else if (N == 2) {
    if (g->kernel_opt == 1 && cg::tiled2d_supported<T>(D, O, what))
      nl = cg::launch_metrics2d_tiled<T>(D, ex, ey, O, g->scheme, what, st);
    else {
      cg::k_metrics2d_ref<T><<<launch_blocks(nc, 128), 128, 0, st>>>(D, ex, ey, O, g->scheme, what);
      nl = 1;
    }
  } else {
    // kernel_opt: 0 gather only, 1 all marching kernels, 16+mask = marching kernels for the axes in mask
    // (bit0 xi+cell, bit1 eta, bit2 zeta) and the gather kernel for the rest (development cross-checks)
    int mask = g->kernel_opt == 1 ? 7 : (g->kernel_opt >= 16 ? (g->kernel_opt - 16) & 7 : 0);
    if (!cg::tiled3d_supported<T>(D, O)) mask = 0;
    if (ranged && mask != 7) return fail(CG_ERR_ARG, "cg_grid_eval_planes needs the marching kernels (CG_OPT_KERNEL = 1)");
    if (mask != 7) {
      const int rest = (~mask & 7) | ((mask & 1) ? 0 : 8);
      cg::k_metrics3d_ref<T><<<launch_blocks(nc, 128), 128, 0, st>>>(D, ex, ey, ez, O, g->scheme, what, rest);
      nl = 1;
    }
    if (mask) {
      int n2 = cg::launch_metrics3d_tiled<T>(D, ex, ey, ez, O, g->scheme, what, mask, st, &g->fork);
      if (n2 < 0) return fail(CG_ERR_CUDA, "marching kernel launch failed");
      nl += n2;
    }
  }
  if (nl < 0) return fail(CG_ERR_CUDA, "tiled kernel launch failed");
  g_launches += nl;
  CG_CUDA(cudaGetLastError());
  CG_CUDA(cudaEventRecord(g->ev1, st));
  g->timed = true;
  if (final) {
    if (what & CG_WHAT_CELL) g->valid_cell = true;
    if (what & CG_WHAT_FACE) { g->valid_face = true; g->face_epoch++; }
  }
  return CG_OK;
}

// The reference recomputes node / centroid / face coordinates on every update! (discrete_grid.jl:605-647,
// mapped_grid.jl:551-587).  Coordinate storage is allocated lazily here, so the refresh runs when it exists.
int refresh_coords_if_stored(cg_grid* g, cudaStream_t st) {
  if (!g->node[0].p) return CG_OK;
  return g->dtype == CG_F64 ? eval_impl<double>(g, CG_WHAT_COORDS, st) : eval_impl<float>(g, CG_WHAT_COORDS, st);
}

bool is_device_ptr(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

}  // namespace

extern "C" {

int cg_version(void) { return 100; }
const char* cg_last_error(void) { return g_err.c_str(); }
uint64_t cg_launch_count(void) { return g_launches.load(); }

int cg_grid_create(int device, int dim, int dtype, const int64_t* ncell, int nhalo, int scheme, int profile,
                   const int64_t* window_origin, const int64_t* window_cells, const int64_t* global_offset,
                   cg_grid** out) {
  if (!out) return fail(CG_ERR_ARG, "cg_grid_create: out is NULL");
  *out = nullptr;
  if (dim < 1 || dim > 3) return fail(CG_ERR_ARG, "Unsupported metric dimension N=%d", dim);
  if (dtype != CG_F64 && dtype != CG_F32) return fail(CG_ERR_ARG, "unsupported dtype %d", dtype);
  if (nhalo < 0) return fail(CG_ERR_ARG, "nhalo must be >= 0");
  if (scheme < 0 || scheme > 3) return fail(CG_ERR_ARG, "unknown conserved_metric_scheme %d", scheme);
  if (profile != CG_PROFILE_FULL && profile != CG_PROFILE_COMPACT) return fail(CG_ERR_ARG, "unknown profile %d", profile);
  if (!ncell) return fail(CG_ERR_ARG, "ncell is NULL");
  int ndev = 0;
  CG_CUDA(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) return fail(CG_ERR_ARG, "device %d out of range (%d devices)", device, ndev);
  cg_grid* g = new cg_grid();
  g->device = device;
  g->dim = dim;
  g->dtype = dtype;
  g->nhalo = nhalo;
  // ADThomasLombardMetric: 3-D only; in 1-D / 2-D it simplifies to the direct CurvatureCorrected form (metrics.jl:1669-1673)
  g->adtl = scheme == CG_AD_THOMAS_LOMBARD && dim == 3;
  g->scheme = scheme == CG_AD_THOMAS_LOMBARD ? CG_CURVATURE_CORRECTED : scheme;
  g->profile = profile;
  for (int d = 0; d < dim; ++d) {
    if (ncell[d] < 1) {
      delete g;
      return fail(CG_ERR_ARG, "celldims[%d] = %lld must be >= 1", d, (long long)ncell[d]);
    }
    g->ncell[d] = ncell[d];
    const int64_t nfull = ncell[d] + 2 * nhalo;
    g->w0[d] = window_origin ? window_origin[d] : 0;
    g->nw[d] = window_cells ? window_cells[d] : nfull;
    g->goff[d] = global_offset ? global_offset[d] : 0;
    if (g->w0[d] < 0 || g->nw[d] < 1 || g->w0[d] + g->nw[d] > nfull) {
      delete g;
      return fail(CG_ERR_ARG, "window [%lld,+%lld) outside cell.full (0..%lld) on axis %d", (long long)g->w0[d],
                  (long long)g->nw[d], (long long)nfull, d);
    }
  }
  cudaError_t e = cudaSetDevice(device);
  if (e == cudaSuccess) e = cudaMalloc((void**)&g->gcl_bits, 3 * sizeof(unsigned long long));
  if (e != cudaSuccess) {
    delete g;
    return fail(CG_ERR_CUDA, "cg_grid_create: %s", cudaGetErrorString(e));
  }
  *out = g;
  return CG_OK;
}

int cg_grid_destroy(cg_grid* g) {
  if (!g) return CG_OK;
  cudaSetDevice(g->device);
  for (int a = 0; a < 3; ++a) {
    release(g->src[a]);
    release(g->ext[a]);
    release(g->face_fwd[a]);
    release(g->face_inv[a]);
    release(g->face_cons[a]);
    release(g->node[a]);
    release(g->centroid[a]);
    release(g->cactive[a]);
    for (int q = 0; q < 3; ++q) release(g->facec[a][q]);
  }
  release(g->cell_fwd);
  release(g->cell_inv);
  release(g->cjac);
  g->mtab.free_all();
  g->gmap.destroy();
  if (g->gcl_bits) cudaFree(g->gcl_bits);
  if (g->ev0) cudaEventDestroy(g->ev0);
  if (g->ev1) cudaEventDestroy(g->ev1);
  if (g->copy_stream) cudaStreamDestroy(g->copy_stream);
  if (g->comm_stream) cudaStreamDestroy(g->comm_stream);
  for (int e = 0; e < 2; ++e) {
    if (g->copy_ev[e]) cudaEventDestroy(g->copy_ev[e]);
    if (g->comm_ev[e]) cudaEventDestroy(g->comm_ev[e]);
  }
  g->fork.destroy();
  release(g->seam);
  delete g;
  return CG_OK;
}

int cg_grid_set_option(cg_grid* g, int option, int64_t value) {
  if (!g) return fail(CG_ERR_ARG, "NULL grid");
  if (option == CG_OPT_KERNEL) {
    if (value != 0 && value != 1 && !(value >= 16 && value < 24)) return fail(CG_ERR_ARG, "CG_OPT_KERNEL must be 0 or 1");
    g->kernel_opt = (int)value;
    return CG_OK;
  }
  return fail(CG_ERR_ARG, "unknown option %d", option);
}

int cg_grid_required_node_range(const cg_grid* g, int axis, int64_t* lo, int64_t* hi) {
  if (!g || axis < 0 || axis >= g->dim || !lo || !hi) return fail(CG_ERR_ARG, "bad argument");
  const int64_t gn = g->ncell[axis];
  int64_t lm = g->w0[axis] - 1 - g->nhalo, hm = g->w0[axis] + g->nw[axis] + 2 - g->nhalo;
  int64_t l = lm < 0 ? 0 : (lm > gn ? gn : lm), h = hm < 0 ? 0 : (hm > gn ? gn : hm);
  if (lm < 0 && h < 1) h = 1;
  if (hm > gn && l > gn - 1) l = gn - 1;
  *lo = l;
  *hi = h + 1;
  return CG_OK;
}

int cg_grid_set_nodes(cg_grid* g, const void* x, const void* y, const void* z, int halo_coords_included,
                      const int64_t* node_origin, const int64_t* node_count, void* stream) {
  CG_NVTX("cg_grid_set_nodes");
  if (!g) return fail(CG_ERR_ARG, "NULL grid");
  const void* in[3] = {x, y, z};
  for (int q = 0; q < g->dim; ++q)
    if (!in[q]) return fail(CG_ERR_ARG, "coordinate array %d is NULL", q);
  if (g->adtl) return fail(CG_ERR_ARG, "ADThomasLombardMetric is available for MappedGrid only (the reference runs it on analytic mappings, CPU-only)");
  CG_CUDA(cudaSetDevice(g->device));
  cudaStream_t st = (cudaStream_t)stream;
  int64_t origin[3] = {0, 0, 0}, count[3] = {1, 1, 1};
  for (int d = 0; d < g->dim; ++d) {
    if (node_origin && node_count) {
      origin[d] = node_origin[d];
      count[d] = node_count[d];
    } else if (halo_coords_included) {
      origin[d] = -g->nhalo;  // halo nodes are present in the buffer but never read
      count[d] = g->ncell[d] + 1 + 2 * g->nhalo;
    } else {
      origin[d] = 0;
      count[d] = g->ncell[d] + 1;
    }
    int64_t lo, hi;
    cg_grid_required_node_range(g, d, &lo, &hi);
    if (origin[d] > lo || origin[d] + count[d] < hi)
      return fail(CG_ERR_ARG, "node arrays cover [%lld,%lld) on axis %d but the window needs [%lld,%lld)",
                  (long long)origin[d], (long long)(origin[d] + count[d]), d, (long long)lo, (long long)hi);
    if (g->ncell[d] < 1) return fail(CG_ERR_ARG, "need at least 2 nodes per axis");
  }
  const size_t bytes = (size_t)(count[0] * count[1] * count[2]) * g->esize();
  for (int q = 0; q < g->dim; ++q) {
    if (is_device_ptr(in[q])) {
      release(g->src[q]);
      g->src[q].p = const_cast<void*>(in[q]);
      g->src[q].bytes = bytes;
      g->src[q].owned = false;
    } else {
      if (g->src[q].p && (!g->src[q].owned || g->src[q].bytes < bytes)) release(g->src[q]);
      int rc = ensure(g, g->src[q], bytes);
      if (rc) return rc;
      CG_CUDA(cudaMemcpyAsync(g->src[q].p, in[q], bytes, cudaMemcpyHostToDevice, st));
    }
  }
  for (int d = 0; d < 3; ++d) {
    g->src_origin[d] = origin[d];
    g->src_count[d] = count[d];
  }
  g->kind = 1;
  g->valid_cell = g->valid_face = g->valid_coords = false;
  int rc = g->dtype == CG_F64 ? pad_nodes<double>(g, st) : pad_nodes<float>(g, st);
  if (rc) return rc;
  return refresh_coords_if_stored(g, st);
}

int cg_grid_node_buffer(cg_grid* g, int coord, void** devptr, int64_t* origin3, int64_t* count3) {
  if (!g || coord < 0 || coord >= g->dim || !devptr) return fail(CG_ERR_ARG, "bad argument");
  if (!g->src[coord].p) return fail(CG_ERR_STATE, "no node buffer: call cg_grid_set_nodes first");
  *devptr = g->src[coord].p;
  for (int d = 0; d < 3; ++d) {
    if (origin3) origin3[d] = g->src_origin[d];
    if (count3) count3[d] = g->src_count[d];
  }
  return CG_OK;
}

int cg_grid_nodes_changed(cg_grid* g, void* stream) {
  if (!g) return fail(CG_ERR_ARG, "NULL grid");
  if (g->kind != 1) return fail(CG_ERR_STATE, "not a DiscreteGrid");
  CG_CUDA(cudaSetDevice(g->device));
  g->valid_cell = g->valid_face = g->valid_coords = false;
  int rc = g->dtype == CG_F64 ? pad_nodes<double>(g, (cudaStream_t)stream) : pad_nodes<float>(g, (cudaStream_t)stream);
  if (rc) return rc;
  return refresh_coords_if_stored(g, (cudaStream_t)stream);
}

int cg_grid_set_mapping_terms(cg_grid* g, int nterms, const double* table) {
  if (!g) return fail(CG_ERR_ARG, "NULL grid");
  if (nterms < 1 || !table) return fail(CG_ERR_ARG, "empty mapping");
  for (int m = 0; m < nterms; ++m) {
    const double* r = table + 15 * m;
    if (r[0] < 0 || r[0] >= g->dim) return fail(CG_ERR_ARG, "term %d: coord %g out of range", m, r[0]);
    for (int d = 0; d < 3; ++d)
      if (r[6 + 3 * d] < 0 || r[6 + 3 * d] > 3) return fail(CG_ERR_ARG, "term %d: bad factor kind", m);
  }
  g->terms.assign(table, table + 15 * (size_t)nterms);
  g->gmap.active = false;
  g->kind = 2;
  g->mtab.dirty = true;
  g->valid_cell = g->valid_face = g->valid_coords = false;
  return CG_OK;
}

int cg_grid_set_mapping_source(cg_grid* g, const char* cuda_source) {
  CG_NVTX("cg_grid_set_mapping_source (NVRTC)");
  if (!g) return fail(CG_ERR_ARG, "NULL grid");
  if (!cuda_source || !*cuda_source) return fail(CG_ERR_ARG, "empty mapping source");
  CG_CUDA(cudaSetDevice(g->device));
  const std::string err = cg::generic_load(g->gmap, cuda_source);
  if (!err.empty()) return fail(err.rfind("NVRTC:", 0) == 0 ? CG_ERR_ARG : CG_ERR_CUDA, "cg_grid_set_mapping_source: %s", err.c_str());
  g->kind = 2;
  g->terms.clear();
  g->valid_cell = g->valid_face = g->valid_coords = false;
  return CG_OK;
}

int cg_grid_set_mapping_params(cg_grid* g, const double* params, int nparams) {
  if (!g) return fail(CG_ERR_ARG, "NULL grid");
  if (nparams < 0 || (nparams > 0 && !params)) return fail(CG_ERR_ARG, "bad mapping parameters");
  CG_CUDA(cudaSetDevice(g->device));
  if (nparams > g->gmap.nparams || !g->gmap.d_params) {
    if (g->gmap.d_params) cudaFree(g->gmap.d_params);
    g->gmap.d_params = nullptr;
    CG_CUDA(cudaMalloc((void**)&g->gmap.d_params, sizeof(double) * (size_t)(nparams > 0 ? nparams : 1)));
  }
  g->gmap.nparams = nparams;
  // pageable source: the copy is synchronous with respect to the host buffer, ordered on the legacy default stream
  if (nparams > 0) CG_CUDA(cudaMemcpy(g->gmap.d_params, params, sizeof(double) * (size_t)nparams, cudaMemcpyHostToDevice));
  g->valid_cell = g->valid_face = g->valid_coords = false;
  return CG_OK;
}

int cg_generic_compile_check(const char* cuda_source, char* log, size_t log_bytes) {
  if (!cuda_source) return fail(CG_ERR_ARG, "NULL source");
  std::vector<char> cubin;
  const std::string err = cg::generic_compile(cuda_source, cubin);
  if (log && log_bytes) snprintf(log, log_bytes, "%s", err.c_str());
  if (!err.empty()) return fail(CG_ERR_ARG, "%s", err.c_str());
  return CG_OK;
}

int cg_grid_update(cg_grid* g, double t, void* stream) {
  CG_NVTX("cg_grid_update");
  if (!g) return fail(CG_ERR_ARG, "NULL grid");
  if (g->kind == 0) return fail(CG_ERR_STATE, "cg_grid_update: no nodes or mapping set");
  g->t = t;
  g->mtab.dirty = true;
  g->valid_cell = g->valid_face = g->valid_coords = false;
  CG_CUDA(cudaSetDevice(g->device));
  return refresh_coords_if_stored(g, (cudaStream_t)stream);
}

int cg_grid_eval(cg_grid* g, int what, void* stream) {
  CG_NVTX("cg_grid_eval");
  if (!g) return fail(CG_ERR_ARG, "NULL grid");
  if (what <= 0 || what > 7) return fail(CG_ERR_ARG, "bad what mask %d", what);
  CG_CUDA(cudaSetDevice(g->device));
  return g->dtype == CG_F64 ? eval_impl<double>(g, what, (cudaStream_t)stream)
                            : eval_impl<float>(g, what, (cudaStream_t)stream);
}

int cg_grid_pad_planes(cg_grid* g, int64_t e_lo, int64_t e_hi, void* stream) {
  if (!g) return fail(CG_ERR_ARG, "NULL grid");
  if (g->kind != 1 || g->dim != 3) return fail(CG_ERR_STATE, "cg_grid_pad_planes: 3-D DiscreteGrid only");
  if (e_lo < 0 || e_hi > g->nw[2] + 4 || e_lo > e_hi) return fail(CG_ERR_ARG, "bad padded plane range [%lld, %lld)", (long long)e_lo, (long long)e_hi);
  CG_CUDA(cudaSetDevice(g->device));
  g->valid_cell = g->valid_face = g->valid_coords = false;
  return g->dtype == CG_F64 ? pad_nodes<double>(g, (cudaStream_t)stream, e_lo, e_hi)
                            : pad_nodes<float>(g, (cudaStream_t)stream, e_lo, e_hi);
}

int cg_grid_eval_planes(cg_grid* g, int what, int64_t k_lo, int64_t k_hi, int final, void* stream) {
  CG_NVTX("cg_grid_eval_planes");
  if (!g) return fail(CG_ERR_ARG, "NULL grid");
  if (what <= 0 || what > 3) return fail(CG_ERR_ARG, "bad what mask %d (CELL | FACE)", what);
  if (g->kind != 1 || g->dim != 3) return fail(CG_ERR_STATE, "cg_grid_eval_planes: 3-D DiscreteGrid only");
  if (k_lo < 0 || k_hi > g->nw[2] || k_lo > k_hi) return fail(CG_ERR_ARG, "bad cell plane range [%lld, %lld)", (long long)k_lo, (long long)k_hi);
  CG_CUDA(cudaSetDevice(g->device));
  return g->dtype == CG_F64 ? eval_impl<double>(g, what, (cudaStream_t)stream, k_lo, k_hi, final != 0)
                            : eval_impl<float>(g, what, (cudaStream_t)stream, k_lo, k_hi, final != 0);
}

int cg_grid_valid(const cg_grid* g, int* cell_valid, int* face_valid) {
  if (!g) return fail(CG_ERR_ARG, "NULL grid");
  if (cell_valid) *cell_valid = g->valid_cell;
  if (face_valid) *face_valid = g->valid_face;
  return CG_OK;
}

int cg_grid_coords_valid(const cg_grid* g, int* coords_valid) {
  if (!g || !coords_valid) return fail(CG_ERR_ARG, "bad argument");
  *coords_valid = g->valid_coords;
  return CG_OK;
}

int cg_grid_invalidate(cg_grid* g, int what) {
  if (!g) return fail(CG_ERR_ARG, "NULL grid");
  if (what & CG_WHAT_CELL) g->valid_cell = false;
  if (what & CG_WHAT_FACE) g->valid_face = false;
  if (what & CG_WHAT_COORDS) g->valid_coords = false;
  return CG_OK;
}

int cg_grid_get_ptr(cg_grid* g, int array, int axis, void** devptr, size_t* nbytes) {
  if (!g || !devptr) return fail(CG_ERR_ARG, "bad argument");
  size_t bytes = 0;
  Buf* b = select(g, array, axis, &bytes);
  if (!b) return fail(CG_ERR_ARG, "unknown array %d / axis %d", array, axis);
  const bool fresh = b->p == nullptr;
  int rc = ensure(g, *b, bytes);
  if (rc) return rc;
  if (fresh) invalidate_for(g, array);  // a never-computed array: the cache it belongs to is not valid
  *devptr = b->p;
  if (nbytes) *nbytes = bytes;
  return CG_OK;
}

int cg_grid_bind_output(cg_grid* g, int array, int axis, void* devptr, size_t nbytes) {
  if (!g || !devptr) return fail(CG_ERR_ARG, "bad argument");
  size_t bytes = 0;
  Buf* b = select(g, array, axis, &bytes);
  if (!b) return fail(CG_ERR_ARG, "unknown array %d / axis %d", array, axis);
  if (nbytes < bytes) return fail(CG_ERR_ARG, "buffer too small: %zu < %zu bytes", nbytes, bytes);
  if (reinterpret_cast<uintptr_t>(devptr) & 15)  // 16-B vector stores and TMA bulk stores
    return fail(CG_ERR_ARG, "bound output buffers must be 16-byte aligned");
  if (b->p != devptr) invalidate_for(g, array);  // the new buffer holds nothing yet
  release(*b);
  b->p = devptr;
  b->bytes = nbytes;
  b->owned = false;
  return CG_OK;
}

int cg_grid_copy_out(cg_grid* g, int array, int axis, void* host_dst, size_t nbytes) {
  if (!g || !host_dst) return fail(CG_ERR_ARG, "bad argument");
  size_t bytes = 0;
  Buf* b = select(g, array, axis, &bytes);
  if (!b) return fail(CG_ERR_ARG, "unknown array %d / axis %d", array, axis);
  if (!b->p) return fail(CG_ERR_STATE, "array %d/%d has not been computed", array, axis);
  if (nbytes < bytes) return fail(CG_ERR_ARG, "host buffer too small: %zu < %zu bytes", nbytes, bytes);
  CG_CUDA(cudaSetDevice(g->device));
  CG_CUDA(cudaMemcpy(host_dst, b->p, bytes, cudaMemcpyDeviceToHost));
  return CG_OK;
}

extern "C++" {
namespace {
template <class T>
int face_flux_bundle_t(cg_grid* g, int axis, int csys, void* mv, void* area, void* normal, cudaStream_t st) {
  const int N = g->dim;
  const bool compact = g->profile == CG_PROFILE_COMPACT;
  const void* cons = compact ? g->cactive[axis].p : g->face_cons[axis].p;
  if (!cons) return fail(CG_ERR_STATE, "conserved face metrics of axis %d are not stored", axis);
  const T* fq[3] = {nullptr, nullptr, nullptr};
  if (csys != 0) {
    if (!g->valid_coords) return fail(CG_ERR_STATE, "face coordinates are not valid; call cg_grid_eval(CG_WHAT_COORDS) first");
    for (int c = 0; c < N; ++c) fq[c] = (const T*)g->facec[axis][c].p;
  }
  const int64_t n = g->ncells();
  cg::k_face_flux_bundle<T><<<launch_blocks(n, 256), 256, 0, st>>>(N, axis, csys, compact, n, (const T*)cons, fq[0], fq[1],
                                                                    fq[2], (T*)mv, (T*)area, (T*)normal);
  g_launches++;
  CG_CUDA(cudaGetLastError());
  return CG_OK;
}
template <class T>
int cell_volumes_t(cg_grid* g, int csys, void* vol, cudaStream_t st) {
  const int N = g->dim;
  const bool compact = g->profile == CG_PROFILE_COMPACT;
  const void* J = compact ? g->cjac.p : g->cell_fwd.p;
  if (!J) return fail(CG_ERR_STATE, "cell metrics are not stored");
  const T* cq[3] = {nullptr, nullptr, nullptr};
  if (csys != 0) {
    if (!g->valid_coords) return fail(CG_ERR_STATE, "centroid coordinates are not valid; call cg_grid_eval(CG_WHAT_COORDS) first");
    for (int c = 0; c < N; ++c) cq[c] = (const T*)g->centroid[c].p;
  }
  const int64_t n = g->ncells();
  cg::k_cell_volumes<T><<<launch_blocks(n, 256), 256, 0, st>>>(N, csys, compact, n, (const T*)J, cq[0], cq[1], cq[2], (T*)vol);
  g_launches++;
  CG_CUDA(cudaGetLastError());
  return CG_OK;
}
}  // namespace
}  // extern "C++"

int cg_grid_face_flux_geometry(cg_grid* g, int axis, int coordsys, void* metric_vector, void* area, void* normal,
                               void* stream) {
  if (!g || axis < 0 || axis >= g->dim) return fail(CG_ERR_ARG, "bad grid / face axis");
  if (coordsys < 0 || coordsys > 4) return fail(CG_ERR_ARG, "bad coordinate system id %d", coordsys);
  if (!g->valid_face) return fail(CG_ERR_STATE, "face metrics are not valid; call cg_grid_eval first");
  CG_CUDA(cudaSetDevice(g->device));
  return g->dtype == CG_F64 ? face_flux_bundle_t<double>(g, axis, coordsys, metric_vector, area, normal, (cudaStream_t)stream)
                            : face_flux_bundle_t<float>(g, axis, coordsys, metric_vector, area, normal, (cudaStream_t)stream);
}

int cg_grid_cell_volumes(cg_grid* g, int coordsys, void* volume, void* stream) {
  if (!g || !volume) return fail(CG_ERR_ARG, "bad argument");
  if (coordsys < 0 || coordsys > 4) return fail(CG_ERR_ARG, "bad coordinate system id %d", coordsys);
  if (!g->valid_cell) return fail(CG_ERR_STATE, "cell metrics are not valid; call cg_grid_eval first");
  CG_CUDA(cudaSetDevice(g->device));
  return g->dtype == CG_F64 ? cell_volumes_t<double>(g, coordsys, volume, (cudaStream_t)stream)
                            : cell_volumes_t<float>(g, coordsys, volume, (cudaStream_t)stream);
}

extern "C++" {
namespace {
// dst[dst_idx[p]] = src[src_idx[p]] for p < n: the scalar ghost-cell fill of a multiblock interface
// (exchange_interface!, src/multiblock/transfer.jl:25-117, with the pairs of interface_index_plan :157-195)
template <class E>
__global__ void __launch_bounds__(256) k_copy_index_pairs(E* __restrict__ dst, const E* __restrict__ src,
                                                          const int64_t* __restrict__ dst_idx,
                                                          const int64_t* __restrict__ src_idx, int64_t n) {
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x)
    dst[dst_idx[p]] = src[src_idx[p]];
}
}  // namespace
}  // extern "C++"

int cg_copy_index_pairs(void* dst, const void* src, const void* dst_idx, const void* src_idx, int64_t n, int elem_bytes,
                        void* stream) {
  if (n < 0 || (n > 0 && (!dst || !src || !dst_idx || !src_idx))) return fail(CG_ERR_ARG, "bad argument");
  if (elem_bytes != 4 && elem_bytes != 8) return fail(CG_ERR_ARG, "element size must be 4 or 8 bytes");
  if (n == 0) return CG_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (elem_bytes == 8)
    k_copy_index_pairs<double><<<launch_blocks(n, 256), 256, 0, st>>>((double*)dst, (const double*)src,
                                                                      (const int64_t*)dst_idx, (const int64_t*)src_idx, n);
  else
    k_copy_index_pairs<float><<<launch_blocks(n, 256), 256, 0, st>>>((float*)dst, (const float*)src,
                                                                     (const int64_t*)dst_idx, (const int64_t*)src_idx, n);
  g_launches++;
  CG_CUDA(cudaGetLastError());
  return CG_OK;
}

extern "C++" {
namespace {
// Vector / tensor ghost-cell fill (transfer.jl:461-488): element p of the plan carries its own basis-transfer matrix
// A_p (N x N, column-major like SMatrix); vector: v_dst = A v_src, tensor: T_dst = A T_src A^T.  Fields are AoS like
// Julia's Array{SVector{N,T}} / Array{SMatrix{N,N,T}}: K = N or N*N consecutive values per cell.  A == nullptr: identity.
template <class E, int N, int KIND>
__global__ void __launch_bounds__(256) k_transfer_index_pairs(E* __restrict__ dst, const E* __restrict__ src,
                                                              const int64_t* __restrict__ dst_idx,
                                                              const int64_t* __restrict__ src_idx,
                                                              const E* __restrict__ A, int64_t n) {
  constexpr int K = KIND == 1 ? N : N * N;
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x) {
    E v[K], a[N][N];
#pragma unroll
    for (int k = 0; k < K; ++k) v[k] = src[src_idx[p] * K + k];
#pragma unroll
    for (int c = 0; c < N; ++c)
#pragma unroll
      for (int r = 0; r < N; ++r) a[r][c] = A ? A[p * (N * N) + r + N * c] : (r == c ? E(1) : E(0));
    E o[K];
    if (KIND == 1) {
#pragma unroll
      for (int r = 0; r < N; ++r) {
        E acc = E(0);
#pragma unroll
        for (int c = 0; c < N; ++c) acc += a[r][c] * v[c];
        o[r] = acc;
      }
    } else {
      E t[N][N];  // A * T (T column-major: T[r][c] = v[r + N c])
#pragma unroll
      for (int r = 0; r < N; ++r)
#pragma unroll
        for (int c = 0; c < N; ++c) {
          E acc = E(0);
#pragma unroll
          for (int m = 0; m < N; ++m) acc += a[r][m] * v[m + N * c];
          t[r][c] = acc;
        }
#pragma unroll
      for (int r = 0; r < N; ++r)
#pragma unroll
        for (int c = 0; c < N; ++c) {
          E acc = E(0);
#pragma unroll
          for (int m = 0; m < N; ++m) acc += t[r][m] * a[c][m];
          o[r + N * c] = acc;
        }
    }
#pragma unroll
    for (int k = 0; k < K; ++k) dst[dst_idx[p] * K + k] = o[k];
  }
}

// basis -> Cartesian matrix Q (grid_traits_api.jl:215-258): identity for a Cartesian basis, the unit vectors
// (e_r, e_theta, e_phi) for SphericalCS + SphericalBasis (N = 2: [s c; c -s], N = 3: the usual spherical triad)
template <class E, int N>
__device__ __forceinline__ void basis_to_cartesian(int basis, const E (&q)[3], E (&Q)[N][N]) {
#pragma unroll
  for (int r = 0; r < N; ++r)
#pragma unroll
    for (int c = 0; c < N; ++c) Q[r][c] = r == c ? E(1) : E(0);
  if (basis != 1) return;
  if constexpr (N == 2) {
    const E s = sin(q[1]), c = cos(q[1]);
    Q[0][0] = s; Q[0][1] = c;
    Q[1][0] = c; Q[1][1] = -s;
  } else if constexpr (N == 3) {
    const E st = sin(q[1]), ct = cos(q[1]), sp = sin(q[2]), cp = cos(q[2]);
    Q[0][0] = st * cp; Q[0][1] = ct * cp; Q[0][2] = -sp;
    Q[1][0] = st * sp; Q[1][1] = ct * sp; Q[1][2] = cp;
    Q[2][0] = ct;      Q[2][1] = -st;     Q[2][2] = E(0);
  }
}
// A_p = transpose(Q_to(q_to)) * Q_from(q_from) (basis_transfer_matrix, grid_traits_api.jl:159-171) for the cell pairs
// (from_idx[p], to_idx[p]) of the two blocks' centroid arrays (_build_interface_transforms, cache.jl:55-108)
template <class E, int N>
__global__ void __launch_bounds__(256) k_basis_transfer(int from_basis, int to_basis, const E* fq0, const E* fq1,
                                                        const E* fq2, const int64_t* __restrict__ from_idx,
                                                        const E* tq0, const E* tq1, const E* tq2,
                                                        const int64_t* __restrict__ to_idx, E* __restrict__ A, int64_t n) {
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x) {
    const int64_t fi = from_idx[p], ti = to_idx[p];
    const E qf[3] = {fq0 ? fq0[fi] : E(0), fq1 ? fq1[fi] : E(0), fq2 ? fq2[fi] : E(0)};
    const E qt[3] = {tq0 ? tq0[ti] : E(0), tq1 ? tq1[ti] : E(0), tq2 ? tq2[ti] : E(0)};
    E Qf[N][N], Qt[N][N];
    basis_to_cartesian<E, N>(from_basis, qf, Qf);
    basis_to_cartesian<E, N>(to_basis, qt, Qt);
#pragma unroll
    for (int r = 0; r < N; ++r)
#pragma unroll
      for (int c = 0; c < N; ++c) {
        E acc = E(0);
#pragma unroll
        for (int m = 0; m < N; ++m) acc += Qt[m][r] * Qf[m][c];
        A[p * (N * N) + r + N * c] = acc;
      }
  }
}
template <class E>
int transfer_pairs_t(void* dst, const void* src, const void* di, const void* si, int64_t n, int N, int kind, const void* A,
                     cudaStream_t st) {
  const int blocks = launch_blocks(n, 256);
#define CG_TP(NN, KK)                                                                                                  \
  if (N == NN && kind == KK) {                                                                                         \
    k_transfer_index_pairs<E, NN, KK><<<blocks, 256, 0, st>>>((E*)dst, (const E*)src, (const int64_t*)di,              \
                                                              (const int64_t*)si, (const E*)A, n);                      \
    return CG_OK;                                                                                                      \
  }
  CG_TP(1, 1) CG_TP(2, 1) CG_TP(3, 1) CG_TP(1, 2) CG_TP(2, 2) CG_TP(3, 2)
#undef CG_TP
  return fail(CG_ERR_ARG, "cg_transfer_index_pairs: N must be 1..3 and field_kind 1 (vector) or 2 (tensor)");
}
}  // namespace
}  // extern "C++"

int cg_transfer_index_pairs(void* dst, const void* src, const void* dst_idx, const void* src_idx, int64_t n, int dim,
                            int field_kind, int elem_bytes, const void* transforms, void* stream) {
  if (n < 0 || (n > 0 && (!dst || !src || !dst_idx || !src_idx))) return fail(CG_ERR_ARG, "bad argument");
  if (elem_bytes != 4 && elem_bytes != 8) return fail(CG_ERR_ARG, "element size must be 4 or 8 bytes");
  if (n == 0) return CG_OK;
  int rc = elem_bytes == 8 ? transfer_pairs_t<double>(dst, src, dst_idx, src_idx, n, dim, field_kind, transforms, (cudaStream_t)stream)
                           : transfer_pairs_t<float>(dst, src, dst_idx, src_idx, n, dim, field_kind, transforms, (cudaStream_t)stream);
  if (rc) return rc;
  g_launches++;
  CG_CUDA(cudaGetLastError());
  return CG_OK;
}

int cg_basis_transfer_matrices(int dim, int64_t n, int from_basis, int to_basis, const void* const* from_centroid,
                               const void* from_idx, const void* const* to_centroid, const void* to_idx, int elem_bytes,
                               void* transforms, void* stream) {
  if (dim < 1 || dim > 3 || n < 0 || !transforms || !from_idx || !to_idx || !from_centroid || !to_centroid)
    return fail(CG_ERR_ARG, "bad argument");
  if ((from_basis != 0 && from_basis != 1) || (to_basis != 0 && to_basis != 1)) return fail(CG_ERR_ARG, "basis id must be 0 (Cartesian) or 1 (SphericalBasis)");
  if ((from_basis == 1 || to_basis == 1) && dim < 2)
    return fail(CG_ERR_ARG, "Unsupported public basis transfer for SphericalBasis and N=1.");  // grid_traits_api.jl:259-269
  if (elem_bytes != 4 && elem_bytes != 8) return fail(CG_ERR_ARG, "element size must be 4 or 8 bytes");
  if (n == 0) return CG_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int blocks = launch_blocks(n, 256);
  const void* f[3] = {from_centroid[0], dim > 1 ? from_centroid[1] : nullptr, dim > 2 ? from_centroid[2] : nullptr};
  const void* t[3] = {to_centroid[0], dim > 1 ? to_centroid[1] : nullptr, dim > 2 ? to_centroid[2] : nullptr};
#define CG_BT(EE, NN)                                                                                                   \
  k_basis_transfer<EE, NN><<<blocks, 256, 0, st>>>(from_basis, to_basis, (const EE*)f[0], (const EE*)f[1], (const EE*)f[2], \
                                                   (const int64_t*)from_idx, (const EE*)t[0], (const EE*)t[1],             \
                                                   (const EE*)t[2], (const int64_t*)to_idx, (EE*)transforms, n)
  if (elem_bytes == 8) {
    if (dim == 1) CG_BT(double, 1); else if (dim == 2) CG_BT(double, 2); else CG_BT(double, 3);
  } else {
    if (dim == 1) CG_BT(float, 1); else if (dim == 2) CG_BT(float, 2); else CG_BT(float, 3);
  }
#undef CG_BT
  g_launches++;
  CG_CUDA(cudaGetLastError());
  return CG_OK;
}

extern "C++" {
namespace {
// launches k_gcl over cell.domain of the global grid intersected with the window and leaves the per-column maxima as
// bit patterns in g->gcl_bits.  Without `low` the window-local plane 0 of the slab axis is skipped (it needs
// I - e_axis, which lives on the previous slab); `low` = that slab's last conserved plane of the slab axis.
// [p_lo, p_hi): restriction to window planes of the last axis (p_hi < 0: all); reset = clear the maxima first
int gcl_launch(cg_grid* g, const void* low, cudaStream_t st, bool* empty, int64_t p_lo = 0, int64_t p_hi = -1,
               bool reset = true) {
  const int N = g->dim;
  const bool compact = g->profile == CG_PROFILE_COMPACT;
  int64_t lo[3] = {0, 0, 0}, hi[3] = {1, 1, 1};
  *empty = false;
  for (int d = 0; d < N; ++d) {
    int64_t l = g->nhalo - g->w0[d], h = g->nhalo + g->ncell[d] - g->w0[d];
    const int64_t first = (low && d == N - 1) ? 0 : 1;
    lo[d] = l < first ? first : l;
    hi[d] = h > g->nw[d] ? g->nw[d] : h;
    if (d == N - 1 && p_hi >= 0) {
      if (lo[d] < p_lo) lo[d] = p_lo;
      if (hi[d] > p_hi) hi[d] = p_hi;
    }
    if (hi[d] <= lo[d]) *empty = true;
  }
  if (reset) CG_CUDA(cudaMemsetAsync(g->gcl_bits, 0, 3 * sizeof(unsigned long long), st));
  if (*empty) return CG_OK;
  const void* c[3] = {nullptr, nullptr, nullptr};
  for (int a = 0; a < N; ++a) c[a] = compact ? g->cactive[a].p : g->face_cons[a].p;
  cg::Dims D = g->dims();
  int64_t nrows = 1;
  for (int d = 1; d < N; ++d) nrows *= hi[d] - lo[d];
  const int64_t ext0 = hi[0] - lo[0];
  const dim3 grid((unsigned)((ext0 + 255) / 256 > 64 ? 64 : (ext0 + 255) / 256), (unsigned)(nrows > 65535 ? 65535 : nrows));
  auto go = [&](auto tag, auto ntag, auto ctag) {
    using T = decltype(tag);
    cg::k_gcl<T, decltype(ntag)::value, decltype(ctag)::value><<<grid, 256, 0, st>>>(
        D, (const T*)c[0], (const T*)c[1], (const T*)c[2], lo[0], lo[1], lo[2], hi[0], hi[1], hi[2], (const T*)low, g->gcl_bits);
  };
  auto by_profile = [&](auto tag, auto ntag) {
    if (compact) go(tag, ntag, std::true_type{});
    else go(tag, ntag, std::false_type{});
  };
  auto by_dim = [&](auto tag) {
    if (N == 1) by_profile(tag, std::integral_constant<int, 1>{});
    else if (N == 2) by_profile(tag, std::integral_constant<int, 2>{});
    else by_profile(tag, std::integral_constant<int, 3>{});
  };
  if (g->dtype == CG_F64) by_dim(double{});
  else by_dim(float{});
  g_launches++;
  CG_CUDA(cudaGetLastError());
  return CG_OK;
}
int gcl_read(cg_grid* g, double* out, cudaStream_t st) {
  unsigned long long bits[3];
  CG_CUDA(cudaMemcpyAsync(bits, g->gcl_bits, sizeof bits, cudaMemcpyDeviceToHost, st));
  CG_CUDA(cudaStreamSynchronize(st));
  for (int cc = 0; cc < g->dim; ++cc) memcpy(&out[cc], &bits[cc], 8);
  return CG_OK;
}
}  // namespace
}  // extern "C++"

int cg_grid_gcl_max_seam(cg_grid* g, const void* low_plane, double* out, void* stream) {
  CG_NVTX("cg_grid_gcl_max");
  if (!g || !out) return fail(CG_ERR_ARG, "bad argument");
  if (!g->valid_face) return fail(CG_ERR_STATE, "face metrics are not valid; call cg_grid_eval first");
  CG_CUDA(cudaSetDevice(g->device));
  cudaStream_t st = (cudaStream_t)stream;
  bool empty = false;
  // cg_grid_update_from_host with CG_WHAT_GCL has already reduced the planes >= 1: only the seam plane is left
  const bool acc = g->gcl_epoch == g->face_epoch;
  int rc = acc ? (low_plane ? gcl_launch(g, low_plane, st, &empty, 0, 1, false) : CG_OK) : gcl_launch(g, low_plane, st, &empty);
  if (rc) return rc;
  return gcl_read(g, out, st);
}

int cg_grid_gcl_max(cg_grid* g, double* out, void* stream) { return cg_grid_gcl_max_seam(g, nullptr, out, stream); }

int cg_grid_seam_plane(cg_grid* g, void** devptr, size_t* nbytes) {
  if (!g || !devptr) return fail(CG_ERR_ARG, "bad argument");
  const int N = g->dim, ax = N - 1;
  const bool compact = g->profile == CG_PROFILE_COMPACT;
  const Buf& b = compact ? g->cactive[ax] : g->face_cons[ax];
  if (!b.p) return fail(CG_ERR_STATE, "conserved face metrics of the slab axis are not stored");
  const size_t K = compact ? N : N * N;
  int64_t plane = 1;
  for (int d = 0; d < ax; ++d) plane *= g->nw[d];
  const size_t pb = (size_t)plane * K * g->esize();
  *devptr = (char*)b.p + pb * (size_t)(g->nw[ax] - 1);
  if (nbytes) *nbytes = pb;
  return CG_OK;
}

int cg_grid_sync(cg_grid* g, void* stream) {
  if (!g) return fail(CG_ERR_ARG, "NULL grid");
  CG_CUDA(cudaSetDevice(g->device));
  CG_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  return CG_OK;
}

int cg_grid_last_eval_ms(cg_grid* g, float* ms) {
  if (!g || !ms) return fail(CG_ERR_ARG, "bad argument");
  if (!g->timed) return fail(CG_ERR_STATE, "no eval has been timed");
  CG_CUDA(cudaEventSynchronize(g->ev1));
  CG_CUDA(cudaEventElapsedTime(ms, g->ev0, g->ev1));
  return CG_OK;
}

int cg_discrete_metrics_host(int device, int dim, int dtype, const void* x, const void* y, const void* z,
                             const int64_t* ncell, int nhalo, int halo_coords_included, int scheme, void* cell_fwd,
                             void* cell_inv, void* face_fwd, void* face_inv, void* face_cons) {
  cg_grid* g = nullptr;
  int rc = cg_grid_create(device, dim, dtype, ncell, nhalo, scheme, CG_PROFILE_FULL, nullptr, nullptr, nullptr, &g);
  if (rc) return rc;
  rc = cg_grid_set_nodes(g, x, y, z, halo_coords_included, nullptr, nullptr, nullptr);
  int what = ((cell_fwd || cell_inv) ? CG_WHAT_CELL : 0) | ((face_fwd || face_inv || face_cons) ? CG_WHAT_FACE : 0);
  if (!rc && what) rc = cg_grid_eval(g, what, nullptr);
  const size_t es = g->esize(), nc = (size_t)g->ncells();
  const size_t km = (size_t)(dim * dim + 1) * es * nc, kc = (size_t)(dim * dim) * es * nc;
  if (!rc && cell_fwd) rc = cg_grid_copy_out(g, CG_ARR_CELL_FORWARD, 0, cell_fwd, km);
  if (!rc && cell_inv) rc = cg_grid_copy_out(g, CG_ARR_CELL_INVERSE, 0, cell_inv, km);
  for (int a = 0; a < dim && !rc; ++a) {
    if (face_fwd) rc = cg_grid_copy_out(g, CG_ARR_FACE_FORWARD, a, (char*)face_fwd + a * km, km);
    if (!rc && face_inv) rc = cg_grid_copy_out(g, CG_ARR_FACE_INVERSE, a, (char*)face_inv + a * km, km);
    if (!rc && face_cons) rc = cg_grid_copy_out(g, CG_ARR_FACE_CONSERVED, a, (char*)face_cons + a * kc, kc);
  }
  std::string keep = g_err;
  cg_grid_destroy(g);
  g_err = keep;
  return rc;
}

// =============================================================================================================
// Multi-GPU: node-plane exchange with NCCL inside the boundary.
//
// The reference has no communication layer: local_grid (mapped_grid.jl:344-379) hands out rank-local analytic
// grids and interface_index_plan (src/multiblock/transfer.jl:157-236) lists index pairs "for a solver or MPI
// layer".  A sampled DiscreteGrid slab needs REAL neighbour node planes: 1 below and 3 above its cell planes.
// NCCL is not linked: the process that hosts this library (Julia with NCCL.jl, Python with torch) has normally
// loaded one already, and two NCCL copies in one process is asking for trouble.  The symbols are taken from the
// loaded libnccl.so.2 (dlopen RTLD_NOLOAD), else from $CG_NCCL_LIB, else from the system's libnccl.so.2.
// =============================================================================================================
extern "C++" {
namespace {
struct NcclApi {
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool ok = false;
  std::string why;
};
const NcclApi& nccl_api() {
  static const NcclApi api = [] {
    NcclApi a;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    if (!h) {
      const char* e = getenv("CG_NCCL_LIB");
      if (e) h = dlopen(e, RTLD_NOW | RTLD_LOCAL);
    }
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
    if (!h) {
      a.why = std::string("libnccl.so.2 not found: ") + (dlerror() ? dlerror() : "");
      return a;
    }
    bool all = true;
    auto sym = [&](const char* n) {
      void* p = dlsym(h, n);
      if (!p) {
        all = false;
        a.why = std::string("missing NCCL symbol ") + n;
      }
      return p;
    };
    a.GetUniqueId = (decltype(a.GetUniqueId))sym("ncclGetUniqueId");
    a.CommInitRank = (decltype(a.CommInitRank))sym("ncclCommInitRank");
    a.CommDestroy = (decltype(a.CommDestroy))sym("ncclCommDestroy");
    a.GroupStart = (decltype(a.GroupStart))sym("ncclGroupStart");
    a.GroupEnd = (decltype(a.GroupEnd))sym("ncclGroupEnd");
    a.Send = (decltype(a.Send))sym("ncclSend");
    a.Recv = (decltype(a.Recv))sym("ncclRecv");
    a.AllReduce = (decltype(a.AllReduce))sym("ncclAllReduce");
    a.GetErrorString = (decltype(a.GetErrorString))sym("ncclGetErrorString");
    a.ok = all;
    return a;
  }();
  return api;
}
#define CG_NCCL(expr)                                                                                    \
  do {                                                                                                   \
    ncclResult_t r_ = (expr);                                                                            \
    if (r_ != ncclSuccess) return fail(CG_ERR_NCCL, "%s: %s", #expr, nccl_api().GetErrorString(r_));     \
  } while (0)
int need_nccl() {
  if (!nccl_api().ok) return fail(CG_ERR_NCCL, "NCCL is not available: %s", nccl_api().why.c_str());
  return CG_OK;
}

// k-slab plan (same index algebra as slabs.py, which the CPU tests pin): cell planes [k0,k1) of the global cell.full
// array need the interior node planes required(k0,k1); node plane m is OWNED by the slab that holds cell plane
// m + nhalo (the first / last slab also own what lies below / above).
struct SlabRange { int64_t lo, hi; };
SlabRange slab_required(int64_t k0, int64_t k1, int64_t h, int64_t n) {
  const int64_t lm = k0 - 1 - h, hm = k1 + 2 - h;
  int64_t lo = lm < 0 ? 0 : (lm > n ? n : lm), hi = hm < 0 ? 0 : (hm > n ? n : hm);
  if (lm < 0 && hi < 1) hi = 1;
  if (hm > n && lo > n - 1) lo = n - 1;
  return {lo, hi + 1};
}
SlabRange slab_owned(int64_t k0, int64_t k1, int64_t h, int64_t n, bool first, bool last) {
  auto cl = [&](int64_t v) { return v < 0 ? 0 : (v > n + 1 ? n + 1 : v); };
  const int64_t lo = first ? 0 : cl(k0 - h), hi = last ? n + 1 : cl(k1 - h);
  return {lo, hi > lo ? hi : lo};
}
struct SlabPlanC {
  int ax = 0;
  int64_t k0 = 0, k1 = 0;
  std::vector<int> send_peer, recv_peer;
  std::vector<SlabRange> send, recv;
};
int make_slab_plan(const cg_grid* g, int rank, int nranks, const int64_t* bounds, SlabPlanC& P) {
  if (g->kind != 1) return fail(CG_ERR_STATE, "halo exchange: DiscreteGrid only (a MappedGrid slab needs no exchange)");
  if (nranks < 1 || rank < 0 || rank >= nranks || !bounds) return fail(CG_ERR_ARG, "bad rank / nranks / slab_bounds");
  const int ax = g->dim - 1;
  const int64_t h = g->nhalo, n = g->ncell[ax];
  if (bounds[0] != 0 || bounds[nranks] != n + 2 * h) return fail(CG_ERR_ARG, "slab_bounds must cover cell.full of the slab axis");
  for (int r = 0; r < nranks; ++r)
    if (bounds[r + 1] <= bounds[r]) return fail(CG_ERR_ARG, "slab_bounds must increase");
  if (g->w0[ax] != bounds[rank] || g->nw[ax] != bounds[rank + 1] - bounds[rank])
    return fail(CG_ERR_ARG, "the handle's window [%lld,+%lld) is not slab %d of slab_bounds", (long long)g->w0[ax],
                (long long)g->nw[ax], rank);
  P.ax = ax;
  P.k0 = bounds[rank];
  P.k1 = bounds[rank + 1];
  const SlabRange req_me = slab_required(P.k0, P.k1, h, n);
  const SlabRange own_me = slab_owned(P.k0, P.k1, h, n, rank == 0, rank == nranks - 1);
  if (g->src_origin[ax] > req_me.lo || g->src_origin[ax] + g->src_count[ax] < req_me.hi)
    return fail(CG_ERR_STATE, "the node buffer does not cover the planes this slab needs");
  for (int peer = 0; peer < nranks; ++peer) {
    if (peer == rank) continue;
    const SlabRange req_p = slab_required(bounds[peer], bounds[peer + 1], h, n);
    const SlabRange own_p = slab_owned(bounds[peer], bounds[peer + 1], h, n, peer == 0, peer == nranks - 1);
    int64_t lo = own_me.lo > req_p.lo ? own_me.lo : req_p.lo, hi = own_me.hi < req_p.hi ? own_me.hi : req_p.hi;
    if (hi > lo) {
      P.send_peer.push_back(peer);
      P.send.push_back({lo, hi});
    }
    lo = own_p.lo > req_me.lo ? own_p.lo : req_me.lo;
    hi = own_p.hi < req_me.hi ? own_p.hi : req_me.hi;
    if (hi > lo) {
      P.recv_peer.push_back(peer);
      P.recv.push_back({lo, hi});
    }
  }
  return CG_OK;
}
int ensure_comm_stream(cg_grid* g) {
  if (g->comm_stream) return CG_OK;
  CG_CUDA(cudaStreamCreateWithFlags(&g->comm_stream, cudaStreamNonBlocking));
  CG_CUDA(cudaEventCreateWithFlags(&g->comm_ev[0], cudaEventDisableTiming));
  CG_CUDA(cudaEventCreateWithFlags(&g->comm_ev[1], cudaEventDisableTiming));
  return CG_OK;
}
// grouped ncclSend / ncclRecv of the plan's node planes, straight out of / into the node buffer, on the handle's
// communication stream
int post_plane_exchange(cg_grid* g, ncclComm_t comm, const SlabPlanC& P) {
  const NcclApi& A = nccl_api();
  int64_t plane = 1;
  for (int d = 0; d < P.ax; ++d) plane *= g->src_count[d];
  const ncclDataType_t dt = g->dtype == CG_F64 ? ncclFloat64 : ncclFloat32;
  const size_t es = g->esize();
  if (P.send.empty() && P.recv.empty()) return CG_OK;
  CG_NCCL(A.GroupStart());
  for (int q = 0; q < g->dim; ++q) {
    char* base = (char*)g->src[q].p;
    for (size_t s = 0; s < P.send.size(); ++s)
      CG_NCCL(A.Send(base + (size_t)((P.send[s].lo - g->src_origin[P.ax]) * plane) * es,
                     (size_t)((P.send[s].hi - P.send[s].lo) * plane), dt, P.send_peer[s], comm, g->comm_stream));
    for (size_t r = 0; r < P.recv.size(); ++r)
      CG_NCCL(A.Recv(base + (size_t)((P.recv[r].lo - g->src_origin[P.ax]) * plane) * es,
                     (size_t)((P.recv[r].hi - P.recv[r].lo) * plane), dt, P.recv_peer[r], comm, g->comm_stream));
  }
  CG_NCCL(A.GroupEnd());
  return CG_OK;
}
// which padded node planes / cell planes of the slab depend on received node planes (slabs.overlap_schedule):
// padded plane e holds the interior node plane m = k0 + e - 1 - h, Line()-extrapolated outside [0, n] from
// c = clamp(m) and its inner neighbour; the cell plane k reads the padded planes k .. k+4
struct OverlapC { int64_t a, b, ke_lo, ke_hi, ne, nw; };
OverlapC overlap_schedule(const cg_grid* g, const SlabPlanC& P) {
  const int64_t h = g->nhalo, n = g->ncell[P.ax], nw = P.k1 - P.k0, ne = nw + 4;
  auto received = [&](int64_t m) {
    for (const SlabRange& r : P.recv)
      if (m >= r.lo && m < r.hi) return true;
    return false;
  };
  std::vector<char> dirty((size_t)ne);
  for (int64_t e = 0; e < ne; ++e) {
    const int64_t m = P.k0 + e - 1 - h, c = m < 0 ? 0 : (m > n ? n : m);
    dirty[(size_t)e] = received(c) || (m != c && received(m < c ? c + 1 : c - 1));
  }
  int64_t a = 0, b = ne;
  while (a < ne && dirty[(size_t)a]) ++a;
  while (b > a && dirty[(size_t)b - 1]) --b;
  for (int64_t e = a; e < b; ++e)
    if (dirty[(size_t)e]) a = b = ne;  // not a prefix + suffix (cannot happen with contiguous ownership)
  OverlapC o;
  o.a = a; o.b = b; o.ne = ne; o.nw = nw;
  o.ke_lo = a < nw ? a : nw;
  int64_t khi = b - 4 < nw ? b - 4 : nw;
  o.ke_hi = khi > o.ke_lo ? khi : o.ke_lo;
  return o;
}
template <class T>
int step_overlapped_t(cg_grid* g, ncclComm_t comm, const SlabPlanC& P, int what, cudaStream_t st) {
  int rc;
  const OverlapC S = overlap_schedule(g, P);
  g->valid_cell = g->valid_face = g->valid_coords = false;
  // the exchange reads / writes the node buffer after whatever `st` did to it before
  CG_CUDA(cudaEventRecord(g->comm_ev[0], st));
  CG_CUDA(cudaStreamWaitEvent(g->comm_stream, g->comm_ev[0], 0));
  if ((rc = post_plane_exchange(g, comm, P))) return rc;
  CG_CUDA(cudaEventRecord(g->comm_ev[1], g->comm_stream));
  // everything that depends on owned node planes only, while the neighbour planes travel
  if (S.b > S.a && (rc = pad_nodes<T>(g, st, S.a, S.b))) return rc;
  if (S.ke_hi > S.ke_lo && (rc = eval_impl<T>(g, what, st, S.ke_lo, S.ke_hi, false))) return rc;
  CG_CUDA(cudaStreamWaitEvent(st, g->comm_ev[1], 0));
  if (S.a > 0 && (rc = pad_nodes<T>(g, st, 0, S.a))) return rc;
  if (S.ne > S.b && (rc = pad_nodes<T>(g, st, S.b, S.ne))) return rc;
  if (S.ke_lo > 0 && (rc = eval_impl<T>(g, what, st, 0, S.ke_lo, false))) return rc;
  if (S.nw > S.ke_hi && (rc = eval_impl<T>(g, what, st, S.ke_hi, S.nw, false))) return rc;
  if (what & CG_WHAT_CELL) g->valid_cell = true;
  if (what & CG_WHAT_FACE) { g->valid_face = true; g->face_epoch++; }
  return refresh_coords_if_stored(g, st);
}
}  // namespace
}  // extern "C++"

// -------------------------------------------------------------------------------------------------------------
// update!(grid, x, y, z) from HOST arrays, pipelined; with a communicator also exchange-aware (k-slabs).
// -------------------------------------------------------------------------------------------------------------
extern "C++" {
namespace {
template <class T>
int update_from_host_t(cg_grid* g, const void* const* in, int what, int nchunks, ncclComm_t comm, const SlabPlanC* P,
                       int nranks, int rank, cudaStream_t st) {
  const int ax = 2;
  const int64_t nk = g->src_count[ax], plane = g->src_count[0] * g->src_count[1];
  const int64_t o = g->src_origin[ax], gn = g->ncell[ax], h = g->nhalo, nw = g->nw[ax], ne = nw + 4;
  const size_t es = g->esize();
  int rc;
  if (!g->copy_stream) {
    CG_CUDA(cudaStreamCreateWithFlags(&g->copy_stream, cudaStreamNonBlocking));
    CG_CUDA(cudaEventCreateWithFlags(&g->copy_ev[0], cudaEventDisableTiming));
    CG_CUDA(cudaEventCreateWithFlags(&g->copy_ev[1], cudaEventDisableTiming));
  }
  // planes of the node buffer this rank reads from ITS host arrays: everything without a communicator, the owned
  // planes of the slab otherwise (the others arrive from their owners)
  int64_t own_lo = o, own_hi = o + nk;
  if (P) {
    const SlabRange own = slab_owned(P->k0, P->k1, h, gn, rank == 0, rank == nranks - 1);
    own_lo = own.lo > o ? own.lo : o;
    own_hi = own.hi < o + nk ? own.hi : o + nk;
  }
  const bool do_gcl = (what & CG_WHAT_GCL) && (what & CG_WHAT_FACE);
  what &= CG_WHAT_CELL | CG_WHAT_FACE;
  std::vector<char> landed((size_t)nk, 0), padded((size_t)ne, 0), evald((size_t)nw, 0), gcld((size_t)nw, 0);
  if (do_gcl) {
    if (!g->gcl_bits) CG_CUDA(cudaMalloc((void**)&g->gcl_bits, 3 * sizeof(unsigned long long)));
    CG_CUDA(cudaMemsetAsync(g->gcl_bits, 0, 3 * sizeof(unsigned long long), st));
  }
  auto is_landed = [&](int64_t m) { return m >= o && m < o + nk && landed[(size_t)(m - o)]; };
  auto h2d = [&](int64_t lo, int64_t hi) -> int {  // interior node planes [lo, hi) on the copy stream
    for (int q = 0; q < 3; ++q)
      CG_CUDA(cudaMemcpyAsync((char*)g->src[q].p + (size_t)((lo - o) * plane) * es,
                              (const char*)in[q] + (size_t)((lo - o) * plane) * es, (size_t)((hi - lo) * plane) * es,
                              cudaMemcpyHostToDevice, g->copy_stream));
    return CG_OK;
  };
  // pad every padded plane whose source planes have landed, evaluate every cell plane whose five padded planes exist
  auto advance = [&]() -> int {
    for (int64_t e = 0; e < ne;) {
      auto ready = [&](int64_t ee) {
        if (padded[(size_t)ee]) return false;
        const int64_t m = g->w0[ax] + ee - 1 - h, c = m < 0 ? 0 : (m > gn ? gn : m);
        return is_landed(c) && (m == c || is_landed(m < c ? c + 1 : c - 1));
      };
      if (!ready(e)) { ++e; continue; }
      int64_t e1 = e;
      while (e1 < ne && ready(e1)) ++e1;
      if ((rc = pad_nodes<T>(g, st, e, e1))) return rc;
      for (int64_t i = e; i < e1; ++i) padded[(size_t)i] = 1;
      e = e1;
    }
    for (int64_t k = 0; k < nw;) {
      auto ready = [&](int64_t kk) {
        if (evald[(size_t)kk]) return false;
        for (int d = 0; d < 5; ++d)
          if (!padded[(size_t)(kk + d)]) return false;
        return true;
      };
      if (!ready(k)) { ++k; continue; }
      int64_t k1 = k;
      while (k1 < nw && ready(k1)) ++k1;
      if ((rc = eval_impl<T>(g, what, st, k, k1, false))) return rc;
      for (int64_t i = k; i < k1; ++i) evald[(size_t)i] = 1;
      k = k1;
    }
    // GCL residuals of the planes whose own and lower conserved planes exist (plane 0 is the seam: cg_grid_gcl_max*)
    for (int64_t k = 1; do_gcl && k < nw;) {
      auto ready = [&](int64_t kk) { return !gcld[(size_t)kk] && evald[(size_t)kk] && evald[(size_t)(kk - 1)]; };
      if (!ready(k)) { ++k; continue; }
      int64_t k1 = k;
      while (k1 < nw && ready(k1)) ++k1;
      bool empty = false;
      if ((rc = gcl_launch(g, nullptr, st, &empty, k, k1, false))) return rc;
      for (int64_t i = k; i < k1; ++i) gcld[(size_t)i] = 1;
      k = k1;
    }
    return CG_OK;
  };
  g->valid_cell = g->valid_face = g->valid_coords = false;
  // the copies overwrite the node buffer the previous evaluation's padding read
  CG_CUDA(cudaEventRecord(g->copy_ev[0], st));
  CG_CUDA(cudaStreamWaitEvent(g->copy_stream, g->copy_ev[0], 0));
  std::vector<char> copied((size_t)nk, 0);
  if (P && (!P->send.empty() || !P->recv.empty())) {
    // the planes the neighbours wait for go first; the exchange is posted as soon as they are on the device
    for (const SlabRange& r : P->send) {
      if ((rc = h2d(r.lo, r.hi))) return rc;
      for (int64_t m = r.lo; m < r.hi; ++m) copied[(size_t)(m - o)] = 1;
    }
    CG_CUDA(cudaEventRecord(g->comm_ev[0], g->copy_stream));
    CG_CUDA(cudaStreamWaitEvent(g->comm_stream, g->comm_ev[0], 0));
    if ((rc = post_plane_exchange(g, comm, *P))) return rc;
    CG_CUDA(cudaEventRecord(g->comm_ev[1], g->comm_stream));
  }
  const int64_t nown = own_hi - own_lo;
  if (nchunks < 1) nchunks = 1;
  if (nchunks > nown) nchunks = (int)(nown > 0 ? nown : 1);
  for (int c = 0; c < nchunks; ++c) {
    const int64_t p_lo = own_lo + nown * c / nchunks, p_hi = own_lo + nown * (c + 1) / nchunks;
    for (int64_t m = p_lo; m < p_hi;) {  // the runs of this piece that did not travel with the boundary planes
      if (copied[(size_t)(m - o)]) { ++m; continue; }
      int64_t m1 = m;
      while (m1 < p_hi && !copied[(size_t)(m1 - o)]) ++m1;
      if ((rc = h2d(m, m1))) return rc;
      m = m1;
    }
    CG_CUDA(cudaEventRecord(g->copy_ev[1], g->copy_stream));
    CG_CUDA(cudaStreamWaitEvent(st, g->copy_ev[1], 0));
    for (int64_t m = p_lo; m < p_hi; ++m) landed[(size_t)(m - o)] = 1;
    if ((rc = advance())) return rc;
  }
  if (P && (!P->send.empty() || !P->recv.empty())) {
    CG_CUDA(cudaStreamWaitEvent(st, g->comm_ev[1], 0));
    for (const SlabRange& r : P->recv)
      for (int64_t m = r.lo; m < r.hi; ++m)
        if (m >= o && m < o + nk) landed[(size_t)(m - o)] = 1;
    if ((rc = advance())) return rc;
  }
  for (int64_t k = 0; k < nw; ++k)
    if (!evald[(size_t)k]) return fail(CG_ERR_STATE, "node buffer does not cover the window (cell plane %lld)", (long long)k);
  if (what & CG_WHAT_CELL) g->valid_cell = true;
  if (what & CG_WHAT_FACE) { g->valid_face = true; g->face_epoch++; }
  if (do_gcl) g->gcl_epoch = g->face_epoch;
  return refresh_coords_if_stored(g, st);
}

int update_from_host_common(cg_grid* g, const void* x, const void* y, const void* z, int what, int nchunks, void* comm,
                            int rank, int nranks, const int64_t* slab_bounds, void* stream) {
  CG_NVTX("cg_grid_update_from_host");
  if (!g) return fail(CG_ERR_ARG, "NULL grid");
  if (g->kind != 1 || g->dim != 3) return fail(CG_ERR_STATE, "cg_grid_update_from_host: 3-D DiscreteGrid only");
  if (what <= 0 || (what & ~(CG_WHAT_CELL | CG_WHAT_FACE | CG_WHAT_GCL)) || !(what & 3))
    return fail(CG_ERR_ARG, "bad what mask %d (CELL | FACE [| GCL])", what);
  const void* in[3] = {x, y, z};
  for (int q = 0; q < 3; ++q) {
    if (!in[q]) return fail(CG_ERR_ARG, "coordinate array %d is NULL", q);
    if (!g->src[q].p) return fail(CG_ERR_STATE, "no node buffer: call cg_grid_set_nodes once first");
  }
  CG_CUDA(cudaSetDevice(g->device));
  // the copies land in the library's OWN node buffer: a caller's device array (cg_grid_set_nodes with device
  // pointers) is never overwritten behind the caller's back, the handle switches to a buffer of its own
  for (int q = 0; q < 3; ++q)
    if (!g->src[q].owned) {
      const size_t bytes = g->src[q].bytes;
      release(g->src[q]);
      int rc = ensure(g, g->src[q], bytes);
      if (rc) return rc;
    }
  SlabPlanC P;
  const bool distd = comm != nullptr && nranks > 1;
  if (distd) {
    int rc = need_nccl();
    if (rc) return rc;
    if ((rc = make_slab_plan(g, rank, nranks, slab_bounds, P))) return rc;
    if ((rc = ensure_comm_stream(g))) return rc;
  }
  return g->dtype == CG_F64
             ? update_from_host_t<double>(g, in, what, nchunks, (ncclComm_t)comm, distd ? &P : nullptr, nranks, rank, (cudaStream_t)stream)
             : update_from_host_t<float>(g, in, what, nchunks, (ncclComm_t)comm, distd ? &P : nullptr, nranks, rank, (cudaStream_t)stream);
}
}  // namespace
}  // extern "C++"

int cg_grid_update_from_host(cg_grid* g, const void* x, const void* y, const void* z, int what, int nchunks,
                             void* stream) {
  return update_from_host_common(g, x, y, z, what, nchunks, nullptr, 0, 1, nullptr, stream);
}

int cg_grid_update_from_host_dist(cg_grid* g, const void* x, const void* y, const void* z, int what, int nchunks,
                                  void* comm, int rank, int nranks, const int64_t* slab_bounds, void* stream) {
  if (!comm) return fail(CG_ERR_ARG, "cg_grid_update_from_host_dist: comm is NULL");
  return update_from_host_common(g, x, y, z, what, nchunks, comm, rank, nranks, slab_bounds, stream);
}

int cg_nccl_unique_id(void* id128) {
  if (!id128) return fail(CG_ERR_ARG, "bad argument");
  int rc = need_nccl();
  if (rc) return rc;
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  CG_NCCL(nccl_api().GetUniqueId((ncclUniqueId*)id128));
  return CG_OK;
}

int cg_nccl_comm_create(const void* id128, int nranks, int rank, int device, void** comm) {
  if (!id128 || !comm || nranks < 1 || rank < 0 || rank >= nranks) return fail(CG_ERR_ARG, "bad argument");
  int rc = need_nccl();
  if (rc) return rc;
  CG_CUDA(cudaSetDevice(device));
  ncclUniqueId id;
  memcpy(&id, id128, sizeof id);
  ncclComm_t c = nullptr;
  CG_NCCL(nccl_api().CommInitRank(&c, nranks, id, rank));
  *comm = (void*)c;
  return CG_OK;
}

int cg_nccl_comm_destroy(void* comm) {
  if (!comm) return CG_OK;
  int rc = need_nccl();
  if (rc) return rc;
  CG_NCCL(nccl_api().CommDestroy((ncclComm_t)comm));
  return CG_OK;
}

int cg_nccl_sendrecv(void* comm, const void* sendbuf, size_t send_bytes, int send_peer, void* recvbuf, size_t recv_bytes,
                     int recv_peer, void* stream) {
  if (!comm) return fail(CG_ERR_ARG, "comm is NULL");
  int rc = need_nccl();
  if (rc) return rc;
  const NcclApi& A = nccl_api();
  CG_NCCL(A.GroupStart());
  if (sendbuf && send_bytes > 0 && send_peer >= 0) CG_NCCL(A.Send(sendbuf, send_bytes, ncclChar, send_peer, (ncclComm_t)comm, (cudaStream_t)stream));
  if (recvbuf && recv_bytes > 0 && recv_peer >= 0) CG_NCCL(A.Recv(recvbuf, recv_bytes, ncclChar, recv_peer, (ncclComm_t)comm, (cudaStream_t)stream));
  CG_NCCL(A.GroupEnd());
  return CG_OK;
}

int cg_halo_exchange_begin(cg_grid* g, void* comm, int rank, int nranks, const int64_t* slab_bounds, void* stream) {
  CG_NVTX("cg_halo_exchange_begin");
  if (!g || !comm) return fail(CG_ERR_ARG, "bad argument");
  int rc = need_nccl();
  if (rc) return rc;
  SlabPlanC P;
  if ((rc = make_slab_plan(g, rank, nranks, slab_bounds, P))) return rc;
  CG_CUDA(cudaSetDevice(g->device));
  if ((rc = ensure_comm_stream(g))) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  CG_CUDA(cudaEventRecord(g->comm_ev[0], st));
  CG_CUDA(cudaStreamWaitEvent(g->comm_stream, g->comm_ev[0], 0));
  if ((rc = post_plane_exchange(g, (ncclComm_t)comm, P))) return rc;
  CG_CUDA(cudaEventRecord(g->comm_ev[1], g->comm_stream));
  g->valid_cell = g->valid_face = g->valid_coords = false;
  return CG_OK;
}

int cg_halo_exchange_end(cg_grid* g, void* stream) {
  if (!g) return fail(CG_ERR_ARG, "NULL grid");
  if (!g->comm_stream) return fail(CG_ERR_STATE, "cg_halo_exchange_end without cg_halo_exchange_begin");
  CG_CUDA(cudaSetDevice(g->device));
  CG_CUDA(cudaStreamWaitEvent((cudaStream_t)stream, g->comm_ev[1], 0));
  return CG_OK;
}

int cg_grid_step_overlapped(cg_grid* g, void* comm, int rank, int nranks, const int64_t* slab_bounds, int what,
                            void* stream) {
  CG_NVTX("cg_grid_step_overlapped");
  if (!g || !comm) return fail(CG_ERR_ARG, "bad argument");
  if (what <= 0 || what > 3) return fail(CG_ERR_ARG, "bad what mask %d (CELL | FACE)", what);
  if (g->dim != 3) return fail(CG_ERR_STATE, "cg_grid_step_overlapped: 3-D DiscreteGrid only");
  int rc = need_nccl();
  if (rc) return rc;
  SlabPlanC P;
  if ((rc = make_slab_plan(g, rank, nranks, slab_bounds, P))) return rc;
  CG_CUDA(cudaSetDevice(g->device));
  if ((rc = ensure_comm_stream(g))) return rc;
  return g->dtype == CG_F64 ? step_overlapped_t<double>(g, (ncclComm_t)comm, P, what, (cudaStream_t)stream)
                            : step_overlapped_t<float>(g, (ncclComm_t)comm, P, what, (cudaStream_t)stream);
}

int cg_grid_gcl_max_dist(cg_grid* g, void* comm, int rank, int nranks, double* out, void* stream) {
  CG_NVTX("cg_grid_gcl_max_dist");
  if (!g || !out || !comm) return fail(CG_ERR_ARG, "bad argument");
  if (!g->valid_face) return fail(CG_ERR_STATE, "face metrics are not valid; evaluate first");
  int rc = need_nccl();
  if (rc) return rc;
  const NcclApi& A = nccl_api();
  CG_CUDA(cudaSetDevice(g->device));
  cudaStream_t st = (cudaStream_t)stream;
  void* last = nullptr;
  size_t pb = 0;
  if ((rc = cg_grid_seam_plane(g, &last, &pb))) return rc;
  if (rank > 0 && (rc = ensure(g, g->seam, pb))) return rc;
  // my last conserved plane of the slab axis goes up, the lower neighbour's comes in (same stream: ordered after the
  // metric kernels)
  if (nranks > 1) {
    CG_NCCL(A.GroupStart());
    if (rank + 1 < nranks) CG_NCCL(A.Send(last, pb, ncclChar, rank + 1, (ncclComm_t)comm, st));
    if (rank > 0) CG_NCCL(A.Recv(g->seam.p, pb, ncclChar, rank - 1, (ncclComm_t)comm, st));
    CG_NCCL(A.GroupEnd());
  }
  bool empty = false;
  if (g->gcl_epoch == g->face_epoch) {  // planes >= 1 were reduced while the host update ran: seam plane only
    if (rank > 0 && (rc = gcl_launch(g, g->seam.p, st, &empty, 0, 1, false))) return rc;
  } else if ((rc = gcl_launch(g, rank > 0 ? g->seam.p : nullptr, st, &empty))) return rc;
  // non-negative doubles order like their bit patterns (a NaN pattern is the largest): max over ranks on the bits
  if (nranks > 1) CG_NCCL(A.AllReduce(g->gcl_bits, g->gcl_bits, 3, ncclUint64, ncclMax, (ncclComm_t)comm, st));
  return gcl_read(g, out, st);
}

}  // extern "C"

// MappedGrid (analytic term-list mapping) device path.  Placeholder: tables only.
#pragma once
#include <vector>
#include "cg_kernels_ref.cuh"
namespace cg {
struct MappedDev { const double* tab; int nterms; };
struct MappedTables {
  bool dirty = true;
  int launches = 0;
  double* d_tab = nullptr;
  int nterms = 0;
  MappedDev dev() const { return MappedDev{d_tab, nterms}; }
  void free_all() { if (d_tab) cudaFree(d_tab); d_tab = nullptr; }
};
template <class T>
inline int mapped_prepare(MappedTables&, const std::vector<double>&, double, const Dims&, cudaStream_t) { return 1; }
template <class T>
__global__ void k_mapped_nodes(Dims, MappedDev, T*, T*, T*) {}
}  // namespace cg

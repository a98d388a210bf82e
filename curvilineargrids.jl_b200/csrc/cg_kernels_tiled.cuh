// Tiled / marching kernels (the fast path).  Placeholder until the tiled kernels land:
// reports "unsupported" so cg_grid_eval uses the gather kernels.
#pragma once
#include "cg_kernels_ref.cuh"
namespace cg {
template <class T> inline bool tiled2d_supported(const Dims&, const MetricOut<T>&) { return false; }
template <class T> inline bool tiled3d_supported(const Dims&, const MetricOut<T>&) { return false; }
template <class T> inline int launch_metrics2d_tiled(const Dims&, const T*, const T*, const MetricOut<T>&, int, int, cudaStream_t) { return -1; }
template <class T> inline int launch_metrics3d_tiled(const Dims&, const T*, const T*, const T*, const MetricOut<T>&, int, int, cudaStream_t) { return -1; }
}  // namespace cg

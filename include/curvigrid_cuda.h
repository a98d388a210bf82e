/*
 * libcurvigrid_cuda.so — B200-native metric evaluation for CurvilinearGrids.jl
 * (C ABI; plain pointers and sizes only, no C++/torch types).
 *
 * The reference has no FFI.  Its only backend seam is Julia multiple dispatch on
 * a `backend::KernelAbstractions.Backend` first argument; each entry point below
 * names the reference method a Julia `ccall` shim would override with it
 * (paths relative to /root/reference/src/grids/unified_grids/).  INTEGRATION.md
 * shows the shim.
 *
 * Memory layouts are byte-identical to the reference's Julia arrays
 * (SURVEY.md Appendix B): column-major over cells (i fastest), element =
 * N*N matrix entries column-major (F11,F21,...) followed by J for `Metric`
 * (metrics.jl:19-22), N*N entries for `ConservedMetric` (metrics.jl:32-34).
 * Storage index I of a face array holds the face I+1/2 (high side).
 *
 * Every call returns a cg_status; cg_last_error() gives the message of the last
 * failure on the calling thread.  Nothing throws across the ABI.  A handle is
 * bound to one device and must not be used from two host threads at once (the
 * reference has no locking either).  `stream` arguments are cudaStream_t passed
 * as void* (NULL = the legacy default stream); calls are stream-ordered and
 * asynchronous unless stated otherwise.
 *
 * Environment variables read by the library (debug / experiment switches only; none is needed in production):
 *   CG_CONCURRENT=0      launch the three 3-D face kernels of a SMALL grid serially instead of on forked streams
 *   CG_2D_STAGED=0       2-D FULL cell+face: use the pair kernel instead of the staged (TMA) kernel
 *   CG_MAPPED_NCHUNK=n   MappedGrid 3-D kernel: number of chunks along k (default: wave-aware choice)
 *   CG_MAPPED_KERNEL=staged|uni  MappedGrid 3-D FULL: the round-1 kernel (every lane multiplies all three table rows) or
 *                        the single kernel with warp-uniform eta/zeta products, instead of the four light kernels
 *                        (k_mapped_cons_uni + k_mapped_face_uni<0,1,2>)
 *   CG_MAPPED_CONS=uni   MappedGrid 3-D FULL: one conserved kernel for all three face axes (k_mapped_cons_uni) instead of
 *                        one light kernel per face axis (k_mapped_cons_face)
 *   CG_MAPPED_FACE_RX=0  MappedGrid 3-D face kernels: re-read the xi factor rows from L1 in every step instead of holding
 *                        them in registers
 *   CG_MAPPED_FOLD=1     MappedGrid 3-D FULL: conserved metrics folded into the three face kernels (an experiment: slower
 *                        than the four light kernels)
 *   CG_FUSE_XZ=1         3-D FULL cell+face: the fused xi+zeta kernel (k_face_xz, an experiment: slower than the
 *                        two kernels it replaces) instead of k_face_xi + k_face_zeta
 *   CG_LEGACY_FUSED=0    legacy MEG6 pipeline: the pass-by-pass launches (273 per 3-D update) instead of the fused
 *                        ones (20); both produce the same bits (tests/test_gpu_legacy.py)
 *   CG_LEGACY_MARCH=0    legacy MEG6 pipeline: tile kernels for the eta / zeta derivatives too, instead of the marching
 *                        (register-only) kernel + the end-strip kernel
 *   CG_LEGACY_MARCHX=1   legacy MEG6 pipeline: xi derivatives through the warp-private transpose + sweep kernel
 *                        (k_legacy_march_x, an experiment: as fast as the tile kernel it replaces)
 *   CG_LEGACY_TMA=0      legacy MEG6 pipeline: per-element loads of the source bricks instead of tensor-map TMA loads
 *   CG_NCCL_LIB=path     libnccl to dlopen when the process has not loaded one (default libnccl.so.2)
 *   CG_NVRTC_LIB=path    libnvrtc to dlopen for cg_grid_set_mapping_source (default libnvrtc.so.12, then the CUDA toolkit's)
 */
#ifndef CURVIGRID_CUDA_H
#define CURVIGRID_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct cg_grid cg_grid; /* opaque */

enum cg_status {
  CG_OK = 0,
  CG_ERR_ARG = 1,    /* ArgumentError in the reference (unified_grids.jl:54-86) */
  CG_ERR_OOM = 2,
  CG_ERR_CUDA = 3,
  CG_ERR_STATE = 4,  /* e.g. metrics requested with storage disabled (cache_mode=:off) */
  CG_ERR_NCCL = 5
};

/* conserved_metric_scheme (metrics.jl:111-126) */
enum cg_scheme {
  CG_ENDPOINT_AVERAGE = 0,
  CG_GRADIENT_CORRECTED = 1,
  CG_CURVATURE_CORRECTED = 2,
  /* ADThomasLombardMetric (metrics.jl:135, 607-668, 3082-3188): 3-D MappedGrid, FULL profile: face forward metric =
   * Jacobian at the face centre, inverse = inv, J = det; conserved = active row only (= the CurvatureCorrected row).
   * In 1-D / 2-D it is the CurvatureCorrected scheme (metrics.jl:1669-1673).  Not available for DiscreteGrid. */
  CG_AD_THOMAS_LOMBARD = 3
};
enum cg_dtype { CG_F64 = 0, CG_F32 = 1 };
/* FULL: the reference's 107 doubles/cell (3-D).  COMPACT: J of the cell + the active conserved
 * row per axis = what face_flux_geometry/cellvolume read (face_geometry.jl:393-416). */
enum cg_profile { CG_PROFILE_FULL = 0, CG_PROFILE_COMPACT = 1 };
enum cg_what { CG_WHAT_CELL = 1, CG_WHAT_FACE = 2, CG_WHAT_COORDS = 4, CG_WHAT_GCL = 8 /* cg_grid_update_from_host* only */ };

/* array selectors for cg_grid_get_ptr / cg_grid_copy_out / cg_grid_bind_output */
enum cg_array {
  CG_ARR_CELL_FORWARD = 0,   /* cell_metrics(grid).forward            [cell][N*N+1] */
  CG_ARR_CELL_INVERSE = 1,   /* cell_metrics(grid).inverse            [cell][N*N+1] */
  CG_ARR_FACE_FORWARD = 2,   /* face_metrics(grid)[axis].forward      [cell][N*N+1] */
  CG_ARR_FACE_INVERSE = 3,   /* face_metrics(grid)[axis].inverse      [cell][N*N+1] */
  CG_ARR_FACE_CONSERVED = 4, /* face_metrics(grid)[axis].conserved    [cell][N*N]   */
  CG_ARR_NODE_COORD = 5,     /* grid.node_coordinates[axis]           [node.full]   */
  CG_ARR_CENTROID_COORD = 6, /* grid.centroid_coordinates[axis]       [cell.full]   */
  CG_ARR_FACE_COORD = 7,     /* grid.face_coordinates[axis/3][axis%3] [cell.full]; axis = 3*face_axis + coord */
  CG_ARR_COMPACT_J = 8,      /* compact profile: det(F) of the cell   [cell]        */
  CG_ARR_COMPACT_ACTIVE = 9  /* compact profile: conserved[axis][axis,:]  [cell][N] */
};

enum cg_option {
  CG_OPT_KERNEL = 0 /* 0 = gather reference kernels, 1 = tiled kernels (default) */
};

int cg_version(void);
const char* cg_last_error(void);
/* number of kernels this library has launched in this process (bench.py `gpu_launches`) */
uint64_t cg_launch_count(void);

/*
 * Grid storage + iterators.  Replaces _build_unified_components
 * (generation.jl:143-183): get_iterators (:3-33) and the KernelAbstractions.zeros
 * allocations (:37-139), which become lazily allocated device buffers.
 *   ncell[d]          interior cells of the WHOLE grid per axis (celldims)
 *   window_origin[d], window_cells[d]
 *                     which part of the global cell.full array this handle
 *                     stores (k-slab of a multi-GPU decomposition); NULL = all.
 *   global_offset[d]  first(global_cell_indices)-1 for a MappedGrid made by
 *                     local_grid (mapped_grid.jl:344-379); NULL = 0.
 */
int cg_grid_create(int device, int dim, int dtype, const int64_t* ncell, int nhalo, int scheme, int profile,
                   const int64_t* window_origin, const int64_t* window_cells, const int64_t* global_offset,
                   cg_grid** out);
int cg_grid_destroy(cg_grid* g);
int cg_grid_set_option(cg_grid* g, int option, int64_t value);

/*
 * DiscreteGrid(x[,y[,z]], nhalo; halo_coords_included) node input
 * (discrete_grid.jl:428-590; strips halos like _strip_halo_nodes :55-61 and
 * applies the Line() extrapolation of _linear_interpolant :63).
 * x,y,z: host or device pointers (detected), grid dtype, column-major.  Extents:
 *   whole grid:  ncell+1 per axis (or ncell+1+2*nhalo with halo_coords_included)
 *   slab window: pass node_origin/node_count = the range of INTERIOR node
 *                indices the arrays cover (must include the planes the window
 *                needs: window cells +-1/+3, clamped to the interior).
 * Marks cell and face metrics invalid (like update!, discrete_grid.jl:605-641).
 */
int cg_grid_set_nodes(cg_grid* g, const void* x, const void* y, const void* z, int halo_coords_included,
                      const int64_t* node_origin, const int64_t* node_count, void* stream);
/* device buffer the node planes live in before padding (for NCCL halo exchange into it) */
int cg_grid_node_buffer(cg_grid* g, int coord, void** devptr, int64_t* origin3, int64_t* count3);
/* (re)apply padding after the node buffer was written in place (e.g. by ncclRecv) */
int cg_grid_nodes_changed(cg_grid* g, void* stream);
/* interior node planes (along the slab axis) this window needs: [lo, hi) */
int cg_grid_required_node_range(const cg_grid* g, int axis, int64_t* lo, int64_t* hi);

/*
 * MappedGrid(x[,y[,z]], params, celldims, nhalo) with a separable term-list
 * mapping (mapped_grid.jl:421-540).  table = nterms rows of 15 doubles:
 *   [coord, c0, c1, c2, w, ph, (kind,a,b) x 3 axes],
 *   q_coord(t,xi) += (c0 + c1 t + c2 sin(w t + ph)) * prod_axis f(kind, a + b*xi_axis)
 *   kind: 0 = 1, 1 = a+b*s, 2 = sin(a+b*s), 3 = cos(a+b*s)
 */
int cg_grid_set_mapping_terms(cg_grid* g, int nterms, const double* table);
/*
 * MappedGrid with an ARBITRARY analytic mapping (the reference's MappedGrid takes any Julia closures,
 * src/grids/unified_grids/mapped_grid.jl:421-540, and differentiates them with ForwardDiff; a closure cannot cross a C
 * ABI, CUDA C++ source can).  `cuda_source` defines
 *
 *     template <class S> __device__ void cg_map(S xi, S eta, S zeta, double t, const double* p, S& x, S& y, S& z);
 *
 * (2-D / 1-D mappings ignore the unused arguments and outputs; + - * / sin cos exp log sqrt tanh and mixed arithmetic with
 * double are available on S).  The library compiles it with NVRTC for sm_100a together with truncated-Taylor number types
 * and generic coordinate / cell / face kernels (csrc/cg_generic_prelude.h) and evaluates the same closure graph as for a
 * term-list mapping: all three schemes, FULL and COMPACT profiles, FP64 / FP32 storage (the arithmetic is FP64),
 * windows (local_grid).  A fallback for mappings that are not separable, about two orders of magnitude slower than the
 * table-driven kernels.  A compile error returns CG_ERR_ARG with the NVRTC log in cg_last_error(); a missing libnvrtc /
 * libcuda returns CG_ERR_CUDA.  `params` (the reference's `params` argument of update!) are copied to the device;
 * t comes from cg_grid_update.  ADThomasLombardMetric needs a term-list mapping.
 */
int cg_grid_set_mapping_source(cg_grid* g, const char* cuda_source);
int cg_grid_set_mapping_params(cg_grid* g, const double* params, int nparams);
/* compile only (no GPU needed): CG_OK, or CG_ERR_ARG with the NVRTC log copied to `log` (may be NULL) */
int cg_generic_compile_check(const char* cuda_source, char* log, size_t log_bytes);

/*
 * update!(grid, t, params) (mapped_grid.jl:551-593, discrete_grid.jl:605-647):
 * refresh node/centroid/face coordinates (if coordinate storage is enabled) and
 * invalidate both metric caches (metric_cache_api.jl:16-41).
 */
int cg_grid_update(cg_grid* g, double t, void* stream);

/*
 * refresh_cell_metrics! / refresh_face_metrics! (metric_cache_api.jl:57-127) =
 * _fill_cell_metric_storage!(backend, ...) metrics.jl:2079 and
 * _fill_face_metric_storage!(backend, ...) metrics.jl:2167/2209, fused in one
 * launch; `what` = CG_WHAT_CELL | CG_WHAT_FACE | CG_WHAT_COORDS
 * (_compute_unified_{node,centroid,face}_coordinates!, generation.jl:319-598).
 */
int cg_grid_eval(cg_grid* g, int what, void* stream);

/*
 * Overlap of the node-plane exchange with interior work (3-D DiscreteGrid k-slabs, marching kernels):
 *   cg_grid_pad_planes  re-pads only the padded node planes [e_lo, e_hi) of the slab axis.  Padded plane e holds
 *                       the window's node plane e - 1; the cell plane k reads the padded planes k .. k+4.
 *   cg_grid_eval_planes fills the cell + face metrics of the window's cell planes [k_lo, k_hi) only; the caches are
 *                       marked valid when `final` != 0 (the caller covers every plane before that).
 * Typical step: post ncclSend/ncclRecv of the neighbour planes; pad + evaluate the planes that depend on owned nodes
 * only; wait for the exchange; pad + evaluate the 1 (low) / 3 (high) boundary planes.
 */
int cg_grid_pad_planes(cg_grid* g, int64_t e_lo, int64_t e_hi, void* stream);
int cg_grid_eval_planes(cg_grid* g, int what, int64_t k_lo, int64_t k_hi, int final, void* stream);

/*
 * update!(grid, x, y, z) of a 3-D DiscreteGrid from HOST coordinate arrays (discrete_grid.jl:605-647) followed by
 * refresh_cell_metrics! / refresh_face_metrics!, pipelined: the node planes (same layout and extents as the arrays
 * of the last cg_grid_set_nodes; pinned host memory for a truly asynchronous copy) go to the device in `nchunks`
 * pieces along the slab axis on an internal copy stream; each piece that lands releases the padded planes and cell
 * planes it completes on `stream` (cg_grid_pad_planes / cg_grid_eval_planes), so the host-to-device copy of piece
 * c+1 overlaps the metric kernels of piece c.  Marks the caches valid.  The copies land in a node buffer the
 * library owns: if cg_grid_set_nodes was given DEVICE arrays the handle stops aliasing them and allocates its own
 * (a caller's device array is never overwritten).  With CG_WHAT_GCL in `what` the GCL residuals of every plane are
 * reduced as soon as the plane and the one below it are complete, overlapped with the copy of the later pieces; the next
 * cg_grid_gcl_max / _seam / _dist then only adds the seam plane and reads the three maxima (same values, bit for bit).
 */
int cg_grid_update_from_host(cg_grid* g, const void* x, const void* y, const void* z, int what, int nchunks,
                             void* stream);

/* valid flags of UnifiedMetricCache (unified_grids.jl:34-52); invalidate_* (metric_cache_api.jl:16-41).
 * Binding a new buffer (cg_grid_bind_output) or first use of a never-computed array (cg_grid_get_ptr) clears the flag
 * of the cache the array belongs to. */
int cg_grid_valid(const cg_grid* g, int* cell_valid, int* face_valid);
/* node / centroid / face coordinates are up to date with the nodes / mapping state (they are re-evaluated by
 * cg_grid_set_nodes, cg_grid_nodes_changed, cg_grid_update, cg_grid_update_from_host and cg_grid_step_overlapped
 * whenever coordinate storage exists, like the reference's update!, discrete_grid.jl:605-647) */
int cg_grid_coords_valid(const cg_grid* g, int* coords_valid);
int cg_grid_invalidate(cg_grid* g, int what);

/* device pointer + size of an output array (allocated on first use) */
int cg_grid_get_ptr(cg_grid* g, int array, int axis, void** devptr, size_t* nbytes);
/* use caller-owned device storage (CUDA.jl / torch) for an output array; devptr must be 16-byte aligned
 * (CUDA allocations are): the kernels write with 16-B vector stores and TMA bulk copies */
int cg_grid_bind_output(cg_grid* g, int array, int axis, void* devptr, size_t nbytes);
/* synchronous device->host copy of an output array */
int cg_grid_copy_out(cg_grid* g, int array, int axis, void* host_dst, size_t nbytes);

/*
 * GridTypes.gcl(face_metrics(grid), cell.domain) reduced to max|I_c| per physical
 * column (../gcl.jl:20-117).  Synchronises the stream; out[0..dim).
 */
int cg_grid_gcl_max(cg_grid* g, double* out, void* stream);
/*
 * The same for a k-slab window.  cg_grid_gcl_max skips the window-local plane 0 of the slab axis (the last axis)
 * when that plane is not the first of cell.domain: its low face lives on the previous slab.  With `low_plane` =
 * the previous slab's LAST plane of the conserved metrics of the slab axis (device pointer, [cells of a plane][K];
 * cg_grid_seam_plane of the neighbour's handle, moved by the caller or by cg_grid_gcl_max_dist) the plane is
 * evaluated too, so the max over the slabs equals the max of the undivided grid.  A non-finite residual gives NaN
 * (maximum(abs.(I)) propagates NaN in the reference).
 */
int cg_grid_gcl_max_seam(cg_grid* g, const void* low_plane, double* out, void* stream);
int cg_grid_seam_plane(cg_grid* g, void** devptr, size_t* nbytes);

/*
 * Solver-facing bundles for every cell of the window, SoA, device pointers of the grid's dtype:
 *   face_flux_geometry(grid, I, hi side of `axis`) (face_geometry.jl:393-416):
 *     metric_vector [N][cell] = component_scale(q_face) .* conserved[axis][I][axis,:]
 *                               (conservation_face_metric_component_scale, face_geometry.jl:302-332),
 *     area [cell] = |metric_vector|, normal [N][cell]; the low face of cell I is minus the entry I - e_axis.
 *   cellvolume(grid, I) = conservation_cell_metric_scale(q_centroid) * |det F| (face_geometry.jl:277-300,
 *     generation.jl:615-726; the array form is _compute_unified_cell_volumes_kernel!, generation.jl:942-990).
 * Work on either profile (COMPACT holds exactly what they read).  coordsys != CG_CS_UNIT needs valid coordinate
 * storage (cg_grid_eval with CG_WHAT_COORDS).  Any output of cg_grid_face_flux_geometry may be NULL.
 */
enum cg_coordsys {
  CG_CS_UNIT = 0,           /* CartesianCS / CurvilinearCS: scale 1 */
  CG_CS_CYLINDRICAL = 1,    /* CylindricalCS + CartesianBasis: 2 pi |q1| */
  CG_CS_AXISYMMETRIC_X = 2, /* AxisymmetricCS{:x}: 2 pi |q2| */
  CG_CS_AXISYMMETRIC_Y = 3, /* AxisymmetricCS{:y}: 2 pi |q1| */
  CG_CS_SPHERICAL_BASIS = 4 /* SphericalCS + SphericalBasis: (r^2 sin th, r sin th, r) per component */
};
int cg_grid_face_flux_geometry(cg_grid* g, int axis, int coordsys, void* metric_vector, void* area, void* normal,
                               void* stream);
int cg_grid_cell_volumes(cg_grid* g, int coordsys, void* volume, void* stream);

/*
 * Scalar ghost-cell fill of a multiblock interface on the device (exchange_interface!, src/multiblock/transfer.jl:25-117):
 * dst[dst_idx[p]] = src[src_idx[p]], p < n, with the ordered pairs of interface_index_plan (transfer.jl:157-236) as
 * 0-based linear indices (int64, device memory).  elem_bytes = 4 or 8.  With dst_idx / src_idx = 0..n-1 the same call
 * packs / unpacks the values of a cross-rank exchange (ncclSend / ncclRecv of the packed buffer in between).
 */
int cg_copy_index_pairs(void* dst, const void* src, const void* dst_idx, const void* src_idx, int64_t n, int elem_bytes,
                        void* stream);
/*
 * Vector / tensor ghost-cell fill (exchange_interface! with field_kind = :vector | :tensor, transfer.jl:461-488):
 * fields are AoS like Julia's Array{SVector{N,T}} / Array{SMatrix{N,N,T}} (K = N or N*N consecutive values per cell,
 * tensors column-major); pair p carries its own basis-transfer matrix A_p (transforms[p][N*N], column-major):
 *   vector (field_kind 1): dst[dst_idx[p]] = A_p * src[src_idx[p]],  tensor (2): A_p * T * transpose(A_p).
 * transforms == NULL: identity (Cartesian bases).  With dst_idx / src_idx = 0..n-1 the same call unpacks a packed
 * buffer that arrived over NCCL (cg_nccl_sendrecv).
 * cg_basis_transfer_matrices builds A_p = transpose(Q_to(q_to)) * Q_from(q_from) (basis_transfer_matrix,
 * grid_traits_api.jl:159-171; _build_interface_transforms, multiblock/cache.jl:55-108) from the two blocks' centroid
 * coordinate arrays (SoA device pointers, one per coordinate) at the paired INTERIOR cells; basis id 0 = Cartesian
 * basis (Q = I), 1 = SphericalCS + SphericalBasis (Q = (e_r, e_theta, e_phi), N = 2 or 3).
 */
int cg_transfer_index_pairs(void* dst, const void* src, const void* dst_idx, const void* src_idx, int64_t n, int dim,
                            int field_kind, int elem_bytes, const void* transforms, void* stream);
int cg_basis_transfer_matrices(int dim, int64_t n, int from_basis, int to_basis, const void* const* from_centroid,
                               const void* from_idx, const void* const* to_centroid, const void* to_idx, int elem_bytes,
                               void* transforms, void* stream);

/*
 * Multi-GPU k-slabs with NCCL inside the boundary (one process per GPU).  The reference has no communication layer
 * (local_grid, mapped_grid.jl:344-379, only hands out rank-local ANALYTIC grids); a sampled DiscreteGrid slab needs
 * real neighbour node planes: 1 below and 3 above its cell planes.  `comm` is an ncclComm_t passed as void* (from
 * NCCL.jl / the host's own NCCL, or from cg_nccl_comm_create); libnccl is resolved at run time from the copy the
 * process has loaded.  `slab_bounds` = nranks + 1 increasing cell-plane indices of the GLOBAL cell.full along the
 * slab axis (the last axis): rank r stores the planes [slab_bounds[r], slab_bounds[r+1]) and the handle's window
 * must be exactly that slab.  Node plane m is owned by the slab that holds the cell plane m + nhalo.  Failures of
 * NCCL return CG_ERR_NCCL.
 *   cg_halo_exchange_begin   grouped ncclSend / ncclRecv of the node planes straight out of / into the handle's node
 *                            buffer on an internal stream (ordered after the work already queued on `stream`)
 *   cg_halo_exchange_end     makes `stream` wait for the exchange; follow with cg_grid_nodes_changed (or
 *                            cg_grid_pad_planes) and cg_grid_eval
 *   cg_grid_step_overlapped  one evaluation of a 3-D slab with the exchange hidden behind interior work: post the
 *                            exchange; pad + evaluate every plane that depends on owned nodes only; wait; pad +
 *                            evaluate the 1 (low) / 3 (high) boundary planes; mark the caches valid
 *   cg_grid_gcl_max_dist     GCL max over ALL slabs: the last conserved plane of the slab axis travels to the next
 *                            rank (seam faces are checked), ncclAllReduce(max) of the three residuals
 */
int cg_nccl_unique_id(void* id128 /* 128 bytes out: ncclGetUniqueId, to be broadcast by the host */);
int cg_nccl_comm_create(const void* id128, int nranks, int rank, int device, void** comm);
int cg_nccl_comm_destroy(void* comm);
/* grouped ncclSend / ncclRecv of raw bytes on `stream` (either half may be absent: NULL buffer or peer < 0): the
 * cross-rank leg of a multiblock interface exchange (packed donor values out, packed ghost values in) */
int cg_nccl_sendrecv(void* comm, const void* sendbuf, size_t send_bytes, int send_peer, void* recvbuf, size_t recv_bytes,
                     int recv_peer, void* stream);
int cg_halo_exchange_begin(cg_grid* g, void* comm, int rank, int nranks, const int64_t* slab_bounds, void* stream);
int cg_halo_exchange_end(cg_grid* g, void* stream);
int cg_grid_step_overlapped(cg_grid* g, void* comm, int rank, int nranks, const int64_t* slab_bounds, int what,
                            void* stream);
int cg_grid_gcl_max_dist(cg_grid* g, void* comm, int rank, int nranks, double* out, void* stream);
/*
 * cg_grid_update_from_host for a k-slab: only the node planes this rank OWNS are read from its host arrays (same
 * layout and extents as the arrays of cg_grid_set_nodes); the planes the neighbours wait for are copied first and
 * the NCCL exchange is posted as soon as they are on the device; the remaining pieces follow, each releasing the
 * padded planes and cell planes it completes; the planes that depend on received nodes are done after the exchange.
 */
int cg_grid_update_from_host_dist(cg_grid* g, const void* x, const void* y, const void* z, int what, int nchunks,
                                  void* comm, int rank, int nranks, const int64_t* slab_bounds, void* stream);

int cg_grid_sync(cg_grid* g, void* stream);
/* device time of the last cg_grid_eval on this handle in ms (CUDA events; synchronises) */
int cg_grid_last_eval_ms(cg_grid* g, float* ms);

/* one-shot host-buffer convenience: DiscreteGrid(...) eager fill, results copied to host
 * buffers (any output pointer may be NULL).  face_* = [axis][cell][K]. */
int cg_discrete_metrics_host(int device, int dim, int dtype, const void* x, const void* y, const void* z,
                             const int64_t* ncell, int nhalo, int halo_coords_included, int scheme,
                             void* cell_fwd, void* cell_inv, void* face_fwd, void* face_inv, void* face_cons);

/*
 * Legacy path: CurvilinearGrid3D(x, y, z, :meg6 | :meg6_symmetric) `update!(mesh, force, include_halo_region)`
 * (src/grids/curvilinear_mapped_grids/3d.jl:472-493 -> src/metric_schemes/MetricDiscretizationSchemes.jl:20-89):
 * centroids, MEG6 forward metrics, Thomas-Lombard (or Nonomura symmetric) conservative metrics and the 57 edge
 * interpolations, FP64.  x,y,z: DEVICE node arrays of extents nnode[3] (interior nodes, or node.full when
 * halo_coords_included).  With ncell = prod(cell.full dims), cell.full = node.full - 1:
 *   cell     [28][ncell]   0 J | 1..9 d x_q/d xi_a at [1+3q+a] | 10..18 inverse [10+3row+col] | 19..27 hatted
 *   edge     [3][19][ncell] per axis: 0 J | 1..9 inverse | 10..18 hatted; storage index I = face I+1/2
 *   centroid [3][ncell]    (may be NULL)
 *   scratch  >= 7*ncell doubles of device memory
 * (the SoA of metric_soa.jl:2-170 without the never-written `.t` arrays).  Values are defined on the metric
 * domain (cell.domain, or cell.full when halo_coords_included), zero elsewhere.
 */
int cg_legacy_meg6_update_3d(int device, const double* x, const double* y, const double* z, const int64_t* nnode,
                             int nhalo, int halo_coords_included, int symmetric, double* cell, double* edge,
                             double* centroid, double* scratch, size_t scratch_bytes, void* stream,
                             uint64_t* launches);
/*
 * Legacy 2-D path: CurvilinearGrid2D(x, y, :meg6 | :meg6_symmetric) `update!(mesh, force, include_halo_region)`
 * (src/grids/curvilinear_mapped_grids/2d.jl:616-637, centroids :734-746; forward_metrics.jl:36-84; the 2-D
 * conservative metrics are inverse * J, conservative_metrics.jl:177-236, and the symmetric scheme forwards to them,
 * :417-421).  x, y: DEVICE node arrays of extents nnode[2].  With ncell = prod(cell.full dims):
 *   cell     [13][ncell]   0 J | 1..4 d x_q/d xi_a at [1+2q+a] | 5..8 xi_x, xi_y, eta_x, eta_y | 9..12 hatted
 *   edge     [2][9][ncell] per axis: 0 J | 1..4 inverse | 5..8 hatted; storage index I = face I+1/2
 *   centroid [2][ncell]    (may be NULL)
 *   scratch  >= 4*ncell doubles of device memory
 */
int cg_legacy_meg6_update_2d(int device, const double* x, const double* y, const int64_t* nnode, int nhalo,
                             int halo_coords_included, double* cell, double* edge, double* centroid, double* scratch,
                             size_t scratch_bytes, void* stream, uint64_t* launches);
/*
 * Legacy 1-D path: CurvilinearGrid1D(x, :meg6) `update!(mesh, force, include_halo_region)`
 * (src/grids/curvilinear_mapped_grids/1d.jl:388-408; forward_metrics.jl:11-27; conservative_metrics.jl:239-245).
 * x: DEVICE node array of nnode[0] values.  With ncell = cell.full:
 *   cell [4][ncell]  0 J = |x_xi| | 1 x_xi | 2 xi_x = 1 / x_xi | 3 hatted = xi_x J
 *   edge [3][ncell]  J, xi_x, hatted interpolated to i+1/2;  centroid [ncell] (may be NULL);  scratch >= 3 * ncell doubles
 */
int cg_legacy_meg6_update_1d(int device, const double* x, const int64_t* nnode, int nhalo, int halo_coords_included,
                             double* cell, double* edge, double* centroid, double* scratch, size_t scratch_bytes,
                             void* stream, uint64_t* launches);
/*
 * _check_valid_metrics (1d.jl:457-482, 2d.jl:748-788, 3d.jl:600-677) as ONE fused reduction on the device instead of
 * the reference's ~60 all(isfinite.(...)) passes: the rows `cell_rows` of cell [rows][ncell] must be finite on
 * cell.domain and the row `positive_row` (-1: none) positive; the rows `edge_rows` of edge [dim][edge_rows_per_axis]
 * [ncell] finite on expand(domain, axis, -1).  Synchronises; *cell_ok / *edge_ok = 1 when valid.
 */
int cg_legacy_check_valid(int device, int dim, const int64_t* nfull, int nhalo, const double* cell, const int* cell_rows,
                          int n_cell_rows, int positive_row, const double* edge, int edge_rows_per_axis,
                          const int* edge_rows, int n_edge_rows, int* cell_ok, int* edge_ok, void* stream);
/*
 * Validity of the LAST fused cg_legacy_meg6_update_3d of this host thread on `device`: the producers test their own
 * values while they write them (J finite and > 0, forward and inverse metrics finite on cell.domain; hatted edge metrics
 * finite on expand(domain, axis, -1): 3d.jl:600-677), so no pass over the arrays is needed.  Synchronises `stream`;
 * CG_ERR_STATE when the last update ran pass-by-pass (CG_LEGACY_FUSED=0) — use cg_legacy_check_valid then.
 */
int cg_legacy_validity(int device, int* cell_ok, int* edge_ok, void* stream);
const char* cg_legacy_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* CURVIGRID_CUDA_H */

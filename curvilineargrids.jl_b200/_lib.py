"""ctypes binding of libcurvigrid_cuda.so (include/curvigrid_cuda.h).

There is no CPU fallback: if the shared library is missing or a call fails the
error is raised to the caller.
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# CG_LIB_PATH: development override (kernel variants under build/variants/)
LIB_PATH = os.environ.get("CG_LIB_PATH") or os.path.join(_HERE, "libcurvigrid_cuda.so")

CG_OK, CG_ERR_ARG, CG_ERR_OOM, CG_ERR_CUDA, CG_ERR_STATE, CG_ERR_NCCL = range(6)
CG_F64, CG_F32 = 0, 1
CG_PROFILE_FULL, CG_PROFILE_COMPACT = 0, 1
CG_WHAT_CELL, CG_WHAT_FACE, CG_WHAT_COORDS, CG_WHAT_GCL = 1, 2, 4, 8
(CG_ARR_CELL_FORWARD, CG_ARR_CELL_INVERSE, CG_ARR_FACE_FORWARD, CG_ARR_FACE_INVERSE, CG_ARR_FACE_CONSERVED,
 CG_ARR_NODE_COORD, CG_ARR_CENTROID_COORD, CG_ARR_FACE_COORD, CG_ARR_COMPACT_J, CG_ARR_COMPACT_ACTIVE) = range(10)
CG_OPT_KERNEL = 0
CG_CS_UNIT, CG_CS_CYLINDRICAL, CG_CS_AXISYMMETRIC_X, CG_CS_AXISYMMETRIC_Y, CG_CS_SPHERICAL_BASIS = range(5)

# every symbol include/curvigrid_cuda.h declares
EXPORTS = [
    "cg_version", "cg_last_error", "cg_launch_count", "cg_grid_create", "cg_grid_destroy", "cg_grid_set_option",
    "cg_grid_set_nodes", "cg_grid_node_buffer", "cg_grid_nodes_changed", "cg_grid_required_node_range",
    "cg_grid_set_mapping_terms", "cg_grid_update", "cg_grid_eval", "cg_grid_valid", "cg_grid_invalidate",
    "cg_grid_get_ptr", "cg_grid_bind_output", "cg_grid_copy_out", "cg_grid_gcl_max", "cg_grid_sync",
    "cg_grid_last_eval_ms", "cg_grid_face_flux_geometry", "cg_grid_cell_volumes", "cg_grid_pad_planes", "cg_grid_eval_planes", "cg_grid_update_from_host", "cg_copy_index_pairs", "cg_discrete_metrics_host", "cg_legacy_meg6_update_3d", "cg_legacy_meg6_update_2d", "cg_legacy_last_error",
    "cg_grid_coords_valid", "cg_grid_gcl_max_seam", "cg_grid_seam_plane", "cg_nccl_unique_id", "cg_nccl_comm_create",
    "cg_nccl_comm_destroy", "cg_halo_exchange_begin", "cg_halo_exchange_end", "cg_grid_step_overlapped",
    "cg_grid_gcl_max_dist", "cg_grid_update_from_host_dist", "cg_transfer_index_pairs", "cg_basis_transfer_matrices",
    "cg_nccl_sendrecv", "cg_legacy_meg6_update_1d", "cg_legacy_check_valid",
    "cg_grid_set_mapping_source", "cg_grid_set_mapping_params", "cg_generic_compile_check", "cg_legacy_validity",
]

_lib = None


class CurvigridError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"[cg_status {code}] {msg}")
        self.code = code


class CurvigridArgumentError(CurvigridError, ValueError):
    """CG_ERR_ARG: the reference throws ArgumentError here."""


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with ./build.sh (or __graft_entry__.build()). "
                "There is no CPU fallback for the metric kernels.")
        L = ctypes.CDLL(LIB_PATH)
        L.cg_last_error.restype = ctypes.c_char_p
        L.cg_launch_count.restype = ctypes.c_uint64
        L.cg_legacy_last_error.restype = ctypes.c_char_p
        _lib = L
    return _lib


def check(rc):
    if rc != CG_OK:
        msg = lib().cg_last_error().decode("utf-8", "replace")
        if rc == CG_ERR_ARG:
            raise CurvigridArgumentError(rc, msg)
        raise CurvigridError(rc, msg)


def i64x3(vals, fill=1):
    v = list(vals) + [fill] * (3 - len(vals))
    return (ctypes.c_int64 * 3)(*v)

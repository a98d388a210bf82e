"""curvigrid_b200 — B200-native metric evaluation behind CurvilinearGrids.jl's grid API.

The directory is named after the reference (`curvilineargrids.jl_b200`); import it
as `curvigrid_b200` (repo-root shim) because of the dot in the name.
"""
from . import _lib, legacy, multiblock, slabs, workloads  # noqa: F401
from ._lib import CurvigridArgumentError, CurvigridError  # noqa: F401
from .grids import (  # noqa: F401
    ADThomasLombardMetric, CurvatureCorrectedReconstruction, DiscreteGrid, EndpointAverageReconstruction, GradientCorrectedReconstruction,
    MappedGrid, cell_metrics, cell_metrics_valid, cellvolume, cellvolumes, face_flux_geometry, face_flux_geometry_arrays, face_metrics, face_metrics_valid, kernel_names_3d,
    fp64_instr_per_cell, gcl, invalidate_cell_metrics, invalidate_face_metrics, local_grid, refresh_cell_metrics, refresh_face_metrics,
    refresh_metrics, seam_plane, update)

__version__ = "0.1.0"

"""CPU-only checks of the boundary: the shared library loads, exports every symbol the header
declares, and reports errors through status codes (no compute calls without a GPU)."""
import ctypes
import os
import re

import curvigrid_b200 as cg

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "curvigrid_cuda.h")).read()
    declared = set(re.findall(r"\b(cg_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"cg_status", "cg_scheme", "cg_map"}   # cg_map: the caller-supplied device function of the generic mapping
    lib = cg._lib.lib()
    missing = [s for s in sorted(declared) if not hasattr(lib, s)]
    assert not missing, missing
    assert declared == set(cg._lib.EXPORTS), declared ^ set(cg._lib.EXPORTS)
    assert lib.cg_version() >= 100


def test_null_and_bad_arguments_return_status_codes():
    L = cg._lib
    lib = L.lib()
    assert lib.cg_grid_destroy(None) == L.CG_OK
    assert lib.cg_grid_eval(None, 3, None) == L.CG_ERR_ARG
    assert b"NULL" in lib.cg_last_error()
    h = ctypes.c_void_p()
    rc = lib.cg_grid_create(0, 7, 0, L.i64x3((4, 4, 4)), 2, 2, 0, None, None, None, ctypes.byref(h))
    assert rc == L.CG_ERR_ARG and h.value is None
    rc = lib.cg_grid_create(0, 2, 0, L.i64x3((4, 4)), -1, 2, 0, None, None, None, ctypes.byref(h))
    assert rc == L.CG_ERR_ARG


def test_no_cpu_fallback():
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError):
        cg.DiscreteGrid(*cg.workloads.wavy_nodes_2d(5, 5), 2)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "curvilineargrids.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.lower() or f == "cg_mapped.cuh" or "oracle" in src.lower() and all(
                    "import" not in ln and "#include" not in ln for ln in src.splitlines() if "oracle" in ln.lower()), f


def test_generic_mapping_source_compiles_without_a_gpu():
    """cg_generic_compile_check: NVRTC turns prelude + mapping + generic kernels into an sm_100a cubin on the CPU box;
    a broken mapping comes back as CG_ERR_ARG with the compiler's message.  (Skipped when libnvrtc is absent.)"""
    import pytest
    L = cg._lib
    lib = L.lib()
    good = cg.workloads.terms_to_cuda_source(cg.workloads.wavy_3d_terms((8, 8, 8), time_amp=0.1)).encode()
    log = ctypes.create_string_buffer(4096)
    rc = lib.cg_generic_compile_check(good, log, 4096)
    if rc != L.CG_OK and b"libnvrtc not found" in log.value:
        pytest.skip("libnvrtc is not installed here")
    assert rc == L.CG_OK, log.value.decode()
    bad = b"template <class S> __device__ void cg_map(S a, S b, S c, double t, const double* p, S& x, S& y, S& z) { x = undefined_fn(a); }"
    assert lib.cg_generic_compile_check(bad, log, 4096) == L.CG_ERR_ARG
    assert b"undefined_fn" in log.value and b"cg_map.cu" in log.value
    assert lib.cg_grid_set_mapping_source(None, good) == L.CG_ERR_ARG


def test_new_entry_points_report_errors_without_a_gpu():
    L = cg._lib
    lib = L.lib()
    assert lib.cg_grid_set_mapping_params(None, None, 0) == L.CG_ERR_ARG
    assert lib.cg_generic_compile_check(None, None, 0) == L.CG_ERR_ARG
    c, e = ctypes.c_int(), ctypes.c_int()
    assert lib.cg_legacy_validity(0, None, ctypes.byref(e), None) == L.CG_ERR_ARG
    # no fused 3-D update has run on this thread: a state error, not a stale answer
    assert lib.cg_legacy_validity(0, ctypes.byref(c), ctypes.byref(e), None) == L.CG_ERR_STATE
    assert b"no fused" in lib.cg_legacy_last_error()
    # CG_WHAT_GCL is a cg_grid_update_from_host* flag only
    assert lib.cg_grid_eval(None, L.CG_WHAT_CELL | L.CG_WHAT_GCL, None) == L.CG_ERR_ARG

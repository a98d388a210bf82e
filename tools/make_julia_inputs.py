"""Writes the INPUTS of the Julia golden cases (tests/golden/julia_inputs/<case>/): raw little-endian Float64 arrays in
Julia (column-major) memory order + a `case.txt` of key=value lines.  tools/julia_goldens.jl reads them with the stock
CurvilinearGrids.jl API and writes the reference's digits next to them (tests/golden/julia/<case>/<scheme>/...), which
tests/test_julia_goldens.py checks the oracle and the CUDA path against.  This is SURVEY §8(f) rank 4's purpose (an
on-disk exchange format for digit-level goldens from a real Julia run) without HDF5, which this image lacks; what must
round-trip is what /root/reference/src/io/to_h5.jl:103-623 stores: coordinates, nhalo, coordinate system, basis.
The `funcmap` cases are mappings that are NOT separable (id + parameters; the formulas are stated three times: as Julia
closures in tools/julia_goldens.jl, as C++ in oracle/metric_oracle.hpp FuncMap, as CUDA source in workloads.py).

Cases cover: curved boundaries (halo = Line() extrapolation, incl. the 2- and 3-face corner regions that settle SURVEY
Appendix C.1), halo_coords_included, 1-/2-/3-D, every conserved scheme, a time-dependent mapping and ADThomasLombard."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def _workloads():
    import importlib.util
    spec = importlib.util.spec_from_file_location("wl", os.path.join(ROOT, "curvilineargrids.jl_b200", "workloads.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def write_case(out, name, meta, arrays):
    d = os.path.join(out, name)
    os.makedirs(d, exist_ok=True)
    with open(os.path.join(d, "case.txt"), "w") as f:
        for k, v in meta.items():
            f.write(f"{k}={v}\n")
    for k, a in arrays.items():
        np.asarray(a, dtype="<f8").ravel(order="F").tofile(os.path.join(d, k + ".bin"))


def main():
    from util import mild_nodes
    wl = _workloads()
    out = os.path.join(ROOT, "tests", "golden", "julia_inputs")
    csv = lambda t: ",".join(str(int(v)) for v in t)
    all3 = "endpoint,gradient,curvature"

    def discrete(name, nodes, nhalo, schemes=all3, halo_included=False, cs="CurvilinearCS", basis="CartesianBasis"):
        meta = {"kind": "discrete", "dim": len(nodes), "nhalo": nhalo, "nn": csv(nodes[0].shape), "schemes": schemes,
                "halo_coords_included": int(halo_included), "coordinate_system": cs, "basis": basis}
        write_case(out, name, meta, {"xyz"[q]: a for q, a in enumerate(nodes)})

    def mapped(name, terms, celldims, nhalo, t, schemes=all3):
        tab = np.asarray(terms, dtype=np.float64).reshape(-1, 15)
        meta = {"kind": "mapped", "dim": len(celldims), "nhalo": nhalo, "celldims": csv(celldims), "schemes": schemes,
                "t": repr(float(t)), "nterms": tab.shape[0]}
        write_case(out, name, meta, {"terms": tab.T})   # Julia reads a 15 x nterms matrix

    def funcmap(name, map_id, params, celldims, nhalo, t, schemes=all3):
        # NON-separable analytic mappings (arbitrary closures on the Julia side, CUDA source on the device side)
        meta = {"kind": "funcmap", "dim": len(celldims), "nhalo": nhalo, "celldims": csv(celldims), "schemes": schemes,
                "t": repr(float(t)), "map_id": map_id, "params": ",".join(repr(float(v)) for v in params)}
        write_case(out, name, meta, {})

    discrete("d3_wavy", wl.wavy_nodes_3d(9, 8, 7), 3)                       # curved boundary, corner halo regions
    discrete("d3_mild", mild_nodes((8, 7, 6), seed=3), 2)
    inner = mild_nodes((7, 6, 5), seed=7)
    big = []
    for a in inner:
        b = np.zeros(tuple(s + 4 for s in a.shape))
        b[2:-2, 2:-2, 2:-2] = a
        big.append(b)
    discrete("d3_halo_included", big, 2, schemes="curvature", halo_included=True)
    discrete("d2_wavy", wl.wavy_nodes_2d(21, 17), 5)
    discrete("d2_mild", mild_nodes((12, 9), seed=2), 3)
    discrete("d1_mild", mild_nodes((9,), seed=1), 2)
    mapped("m2_readme", wl.readme_wavy_2d_terms(), (16, 12), 3, 0.0)
    mapped("m3_wavy_t", wl.wavy_3d_terms((9, 8, 7), time_amp=0.1), (9, 8, 7), 2, 0.37, schemes=all3 + ",adtl")
    mapped("m3_rtp", wl.spherical_rtp_3d_terms((8, 7, 6)), (8, 7, 6), 2, 0.0, schemes="curvature,adtl")
    funcmap("g3_nonseparable", 1, [0.3, 0.4, 0.01, 0.001, 0.02], (9, 8, 7), 3, 0.7)
    funcmap("g2_nonseparable", 2, [0.2, 0.05, 0.3, 0.1], (14, 11), 3, 0.7)
    funcmap("g1_nonseparable", 3, [0.1, 0.6, 0.02, 0.5], (23,), 3, 0.7, schemes="curvature")
    print("wrote", out)


if __name__ == "__main__":
    main()

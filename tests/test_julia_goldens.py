"""Digit-level parity against the REAL package, when its output is present.

tools/julia_goldens.jl (run by anyone with Julia + CurvilinearGrids v0.13.3) writes tests/golden/julia/<case>/<scheme>/
from the inputs committed under tests/golden/julia_inputs/ (tools/make_julia_inputs.py).  With those files present the
oracle (CPU, here) and the CUDA path (-m gpu) are compared with the reference's own Float64 digits; without them the
tests SKIP with the message "parity unpinned" — the oracle is then pinned only by the reference's re-encoded
known-answer / property tests (tests/test_oracle_pins.py).  SURVEY §8(f) rank 4, /root/reference/src/io/to_h5.jl."""
import glob
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INP = os.path.join(ROOT, "tests", "golden", "julia_inputs")
OUT = os.path.join(ROOT, "tests", "golden", "julia")
SCHEME_ID = {"endpoint": 0, "gradient": 1, "curvature": 2, "adtl": 3}


def read_case(name):
    meta = {}
    with open(os.path.join(INP, name, "case.txt")) as f:
        for ln in f:
            if "=" in ln:
                k, v = ln.strip().split("=", 1)
                meta[k] = v
    return meta


def cases():
    out = []
    for d in sorted(glob.glob(os.path.join(INP, "*"))):
        name = os.path.basename(d)
        for s in read_case(name)["schemes"].split(","):
            out.append((name, s))
    return out


def load_inputs(name):
    meta = read_case(name)
    dim, h = int(meta["dim"]), int(meta["nhalo"])
    if meta["kind"] == "discrete":
        nn = tuple(int(v) for v in meta["nn"].split(","))
        nodes = [np.fromfile(os.path.join(INP, name, "xyz"[q] + ".bin"), dtype="<f8").reshape(nn, order="F") for q in range(dim)]
        return meta, dim, h, nodes
    if meta["kind"] == "funcmap":
        return meta, dim, h, [float(v) for v in meta["params"].split(",")]
    nt = int(meta["nterms"])
    terms = np.fromfile(os.path.join(INP, name, "terms.bin"), dtype="<f8").reshape((15, nt), order="F").T.copy()
    return meta, dim, h, terms


def load_golden(name, scheme, dim):
    d = os.path.join(OUT, name, scheme)
    if not os.path.exists(os.path.join(d, "done.txt")):
        pytest.skip("parity unpinned: no Julia goldens under tests/golden/julia (run tools/julia_goldens.jl with "
                    "Julia and CurvilinearGrids v0.13.3; inputs are committed under tests/golden/julia_inputs)")
    KM, KC = dim * dim + 1, dim * dim
    rd = lambda f, k: np.fromfile(os.path.join(d, f), dtype="<f8").reshape(-1, k)
    return {"cell_fwd": rd("cell_forward.bin", KM), "cell_inv": rd("cell_inverse.bin", KM),
            "face_fwd": np.stack([rd(f"face{a}_forward.bin", KM) for a in range(dim)]),
            "face_inv": np.stack([rd(f"face{a}_inverse.bin", KM) for a in range(dim)]),
            "face_cons": np.stack([rd(f"face{a}_conserved.bin", KC) for a in range(dim)]),
            "centroid": [np.fromfile(os.path.join(d, f"centroid{c}.bin"), dtype="<f8") for c in range(dim)],
            "node": [np.fromfile(os.path.join(d, f"node{c}.bin"), dtype="<f8") for c in range(dim)]}


def test_inputs_are_committed_and_reproducible(tmp_path):
    """The inputs a Julia user needs exist, and tools/make_julia_inputs.py regenerates them bit for bit."""
    assert len(cases()) >= 20
    import subprocess
    import sys
    for name in ("d3_wavy", "m3_wavy_t"):
        before = {f: open(f, "rb").read() for f in glob.glob(os.path.join(INP, name, "*"))}
        assert before
    subprocess.check_call([sys.executable, os.path.join(ROOT, "tools", "make_julia_inputs.py")], stdout=subprocess.DEVNULL)
    for f, b in before.items():
        assert open(f, "rb").read() == b, f


@pytest.mark.parametrize("name,scheme", cases())
def test_oracle_matches_julia_digits(oracle, name, scheme):
    from util import assert_metrics_close
    meta, dim, h, data = load_inputs(name)
    gold = load_golden(name, scheme, dim)
    if meta["kind"] == "discrete":
        ref = oracle.discrete_eval(data, h, scheme, halo_coords_included=meta["halo_coords_included"] == "1")
        cen = oracle.discrete_coords(data, h, 1, halo_coords_included=meta["halo_coords_included"] == "1")
    elif meta["kind"] == "funcmap":
        celldims = tuple(int(v) for v in meta["celldims"].split(","))
        ref = oracle.funcmap_eval(dim, int(meta["map_id"]), data, float(meta["t"]), celldims, h, scheme)
        cen = oracle.funcmap_coords(dim, int(meta["map_id"]), data, float(meta["t"]), celldims, h, 1)
    else:
        celldims = tuple(int(v) for v in meta["celldims"].split(","))
        ref = oracle.mapped_eval(data, float(meta["t"]), celldims, h, scheme)
        cen = oracle.mapped_coords(data, float(meta["t"]), celldims, h, 1)
    assert_metrics_close(ref, gold, rtol=1e-12, label=f"oracle vs Julia {name}/{scheme}")
    for c in range(dim):
        assert np.abs(cen[c] - gold["centroid"][c]).max() <= 1e-13 * max(1.0, np.abs(gold["centroid"][c]).max())


@pytest.mark.gpu
@pytest.mark.parametrize("name,scheme", cases())
def test_cuda_matches_julia_digits(cg, name, scheme):
    from util import assert_metrics_close, grid_outputs
    meta, dim, h, data = load_inputs(name)
    gold = load_golden(name, scheme, dim)
    if meta["kind"] == "discrete":
        g = cg.DiscreteGrid(*data, h, conserved_metric_scheme=scheme,
                            halo_coords_included=meta["halo_coords_included"] == "1")
    elif meta["kind"] == "funcmap":   # the NVRTC path against the real package's ForwardDiff evaluation of the closures
        celldims = tuple(int(v) for v in meta["celldims"].split(","))
        g = cg.MappedGrid(cg.workloads.GENERIC_TEST_SOURCES[int(meta["map_id"])], data, celldims, h,
                          conserved_metric_scheme=scheme, cache_mode="lazy")
        cg.update(g, float(meta["t"]), data)
    else:
        celldims = tuple(int(v) for v in meta["celldims"].split(","))
        g = cg.MappedGrid(data, {}, celldims, h, conserved_metric_scheme=scheme, cache_mode="lazy")
        cg.update(g, float(meta["t"]))
    assert_metrics_close(grid_outputs(cg, g), gold, rtol=1e-12, label=f"CUDA vs Julia {name}/{scheme}")
    for c in range(dim):
        got = g.centroid_coordinates[c].reshape(-1).cpu().numpy()
        assert np.abs(got - gold["centroid"][c]).max() <= 1e-13 * max(1.0, np.abs(gold["centroid"][c]).max())

"""Generate the golden vectors under tests/golden/ from the oracle (run by hand; the .npz files
are committed).  The reference ships no stored vectors and cannot run here (no Julia), so these
pin the ORACLE's output: a later change to the oracle or to the kernels shows up as a diff.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import curvigrid_b200 as cg  # noqa: E402
from oracle import oracle as O  # noqa: E402
from util import mild_nodes  # noqa: E402

KEYS = ("cell_fwd", "cell_inv", "face_fwd", "face_inv", "face_cons")


def main():
    wl = cg.workloads
    # discrete grids: mild 3-D / 2-D / 1-D meshes, every scheme
    for name, shape, h in (("discrete3d", (6, 5, 5), 2), ("discrete2d", (9, 8), 3), ("discrete1d", (10,), 2)):
        nodes = mild_nodes(shape, seed=100 + len(shape))
        out = {f"node{q}": nodes[q] for q in range(len(shape))}
        out["nhalo"] = h
        for sch in ("endpoint", "gradient", "curvature"):
            r = O.discrete_eval(nodes, h, sch)
            for k in KEYS:
                out[f"{sch}_{k}"] = r[k]
        np.savez_compressed(os.path.join(HERE, f"{name}.npz"), **out)
    # mapped grids: README quick-start mapping (BASELINE config 1 at reduced size) and the 3-D wavy mapping
    for name, terms, celldims, h, t in (("mapped2d_readme", wl.readme_wavy_2d_terms(), (12, 10), 3, 0.0),
                                        ("mapped3d_wavy", wl.wavy_3d_terms((6, 5, 5), time_amp=0.1), (6, 5, 5), 2, 0.4)):
        out = {"terms": np.asarray(terms), "celldims": np.asarray(celldims), "nhalo": h, "t": t}
        for sch in ("endpoint", "gradient", "curvature"):
            r = O.mapped_eval(terms, t, celldims, h, sch)
            for k in KEYS:
                out[f"{sch}_{k}"] = r[k]
        np.savez_compressed(os.path.join(HERE, f"{name}.npz"), **out)
    print("wrote", sorted(f for f in os.listdir(HERE) if f.endswith(".npz")))


if __name__ == "__main__":
    main()

"""tools/time_legacy.py N [symmetric] : device time of one legacy MEG6 update! (cg_legacy_meg6_update_3d) on an N^3-cell
wavy mesh, fused launches and the pass-by-pass pipeline (CG_LEGACY_FUSED=0), CUDA events, median of 7 repetitions after 3 warm-ups.
Algorithmic traffic: 24 B node read + 3 centroids + 28 cell + 57 edge doubles written = 728 B written + 24 B read per
cell of cell.full (DESIGN.md §10 quotes 848 B with the centroid copy and scratch)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import curvigrid_b200 as cg  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
sym = len(sys.argv) > 2 and sys.argv[2] == "symmetric"
x, y, z = cg.workloads.wavy_nodes_3d(n + 1, n + 1, n + 1)
out = {"n": n, "scheme": "meg6_symmetric" if sym else "meg6"}
for label, env in (("fused", "1"), ("pass_by_pass", "0")):
    os.environ["CG_LEGACY_FUSED"] = env
    mesh = cg.legacy.CurvilinearGrid3D(x, y, z, out["scheme"])
    for _ in range(3):
        cg.legacy.update(mesh, True, False)
    torch.cuda.synchronize()
    reps, times = 7, []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        cg.legacy.update(mesh, True, False)
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    times.sort()
    ms = times[len(times) // 2]
    ncell = n ** 3
    out[label] = {"ms": ms, "ms_min": times[0], "ms_max": times[-1], "launches": mesh.launches, "Mcells_per_s": ncell / ms / 1e3,
                  "GBs_on_848B_per_cell": 848.0 * ncell / ms / 1e6}
    del mesh
print(json.dumps(out))

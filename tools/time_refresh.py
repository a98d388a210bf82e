"""Times the three refresh entry points of a 3-D DiscreteGrid on the current GPU:
refresh_cell_metrics!, refresh_face_metrics!, and the fused cell+face evaluation (kernel ms by CUDA events)."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
import bench
import curvigrid_b200 as cg

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
host = bench.wavy_slab_3d(n + 1, n + 1, n + 1, 0, n + 1)
dev = [torch.from_numpy(a).cuda() for a in host]
g = cg.DiscreteGrid(*dev, 5, cache_mode="lazy")
which = (("cell", cg.refresh_cell_metrics), ("face", cg.refresh_face_metrics), ("cell+face", cg.refresh_metrics))
if len(sys.argv) > 2 and sys.argv[2] == "fused":
    which = which[2:]
for name, fn in which:
    ts = []
    for _ in range(4):
        fn(g)
        ts.append(g.last_eval_ms())
    print(f"{name:10s} {np.mean(ts[1:]):8.3f} ms")

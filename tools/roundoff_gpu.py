"""Diagnostic (GPU box): per face axis / row error of the CUDA conserved metrics against the long-double oracle on the
512^3 BASELINE mesh, next to the FP64 oracle's own error.  Usage: python tools/roundoff_gpu.py [n] [kernel_opt]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import curvigrid_b200 as cg  # noqa: E402
from oracle import oracle as O  # noqa: E402
from util import sampled_outputs  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
kopt = int(sys.argv[2]) if len(sys.argv) > 2 else 1
h = 5
nodes = cg.workloads.wavy_nodes_3d(n + 1, n + 1, n + 1)
g = cg.DiscreteGrid(*nodes, h, cache_mode="lazy")
g.set_kernel(kopt)
cg.refresh_metrics(g)
nf = g.nfull
rng = np.random.default_rng(3)
# cells near the far corner of the domain (largest |x| / dx) and near the near corner
ii = np.concatenate([rng.integers(nf[0] - 40, nf[0] - 6, 24), rng.integers(6, 40, 24)])
jj = np.concatenate([rng.integers(nf[1] - 40, nf[1] - 6, 24), rng.integers(6, 40, 24)])
kk = np.concatenate([rng.integers(nf[2] - 40, nf[2] - 6, 24), rng.integers(6, 40, 24)])
sample = np.sort(ii + nf[0] * (jj + nf[1] * kk)).astype(np.int64)
got = sampled_outputs(cg, g, sample)
truth = O.discrete_eval(nodes, h, "curvature", sample=sample, precision="ld")
o64 = O.discrete_eval(nodes, h, "curvature", sample=sample)
for ax in range(3):
    t = truth["face_cons"][ax]
    sc = np.abs(t).max(axis=1, keepdims=True)
    eg = np.abs(got["face_cons"][ax] - t) / sc
    eo = np.abs(o64["face_cons"][ax] - t) / sc
    for r in range(3):
        cols = [r + 3 * c for c in range(3)]
        print(f"face {ax} row {r}: gpu {eg[:, cols].max():.2e}   oracle64 {eo[:, cols].max():.2e}   per column gpu",
              ["%.1e" % eg[:, r + 3 * c].max() for c in range(3)])
for k in ("cell_fwd", "cell_inv", "face_fwd", "face_inv"):
    t = truth[k]
    sc = np.abs(t[..., :9]).max(axis=-1, keepdims=True)
    print(k, "gpu %.2e  oracle64 %.2e" % ((np.abs(got[k][..., :9] - t[..., :9]) / sc).max(),
                                           (np.abs(o64[k][..., :9] - t[..., :9]) / sc).max()))

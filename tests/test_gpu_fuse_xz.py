"""The fused xi+zeta marching kernel (k_face_xz, CG_FUSE_XZ=1: an experiment that is OFF by default because it is slower
than the two kernels it replaces) stays correct: oracle parity on several chunk layouts.  The switch is
read once per process, so the check runs in a subprocess."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r"""
import sys
sys.path.insert(0, %r); sys.path.insert(0, %r)
import numpy as np
import curvigrid_b200 as cg
from oracle import oracle as O
from util import assert_metrics_close, grid_outputs, mild_nodes
for shape, h, scheme in (((37, 9, 8), 2, "curvature"), ((12, 11, 70), 3, "gradient"), ((8, 7, 6), 2, "endpoint")):
    nodes = mild_nodes(shape, seed=5)
    g = cg.DiscreteGrid(*nodes, h, conserved_metric_scheme=scheme)
    assert_metrics_close(grid_outputs(cg, g), O.discrete_eval(nodes, h, scheme), rtol=1e-12, label=str(shape))
    assert max(cg.gcl(g)) < 5e-13
print("fused xz OK")
"""


def test_fused_xz_kernel_matches_oracle():
    env = dict(os.environ, CG_FUSE_XZ="1")
    r = subprocess.run([sys.executable, "-c", SCRIPT % (ROOT, os.path.join(ROOT, "tests"))], capture_output=True, text=True,
                       env=env, timeout=600)
    assert r.returncode == 0 and "fused xz OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]

import sys; sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import numpy as np
import curvigrid_b200 as cg
from oracle import oracle as O
from util import grid_outputs, mild_nodes, metric_err
nodes = mild_nodes((40,5,131), seed=31)
for scheme in ("endpoint","curvature"):
    ref = O.discrete_eval(nodes, 2, scheme)
    for kern in (0,1):
        g = cg.DiscreteGrid(*nodes, 2, conserved_metric_scheme=scheme, cache_mode="lazy"); g.set_kernel(kern); cg.refresh_metrics(g)
        got = grid_outputs(cg, g)
        print(scheme, 'kernel', kern, {k: '%.2e' % metric_err(got[k], ref[k], has_j=k!='face_cons') for k in ("cell_fwd","cell_inv","face_fwd","face_inv","face_cons")})

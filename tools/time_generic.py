"""tools/time_generic.py N : device time of one cell+face evaluation of a 3-D MappedGrid with a non-separable mapping
through the NVRTC-compiled generic kernels, next to the table-driven path on a separable mapping of the same size."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import curvigrid_b200 as cg  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
SRC = """
template <class S> __device__ void cg_map(S xi, S eta, S zeta, double t, const double* p, S& x, S& y, S& z) {
  x = xi + p[0] * sin(p[1] * (eta + 0.5 * zeta) + 0.3 * t);
  y = eta * (1.0 + p[2] * xi) / (1.0 + p[3] * zeta * zeta);
  z = zeta + p[0] * exp(-p[4] * (xi - eta) * (xi - eta)) * cos(t) + sqrt(1.0 + 0.01 * xi * xi);
}"""
out = {"n": n}
for label, mapping, prm in (("generic_nvrtc", SRC, [0.3, 0.4, 0.001, 1e-5, 0.02]),
                            ("generic_nvrtc_on_separable", cg.workloads.terms_to_cuda_source(cg.workloads.wavy_3d_terms((n, n, n), time_amp=0.1)), None),
                            ("tables_separable", cg.workloads.wavy_3d_terms((n, n, n), time_amp=0.1), {})):
    g = cg.MappedGrid(mapping, prm, (n, n, n), 5, cache_mode="lazy")
    ts = []
    for i in range(3):
        cg.update(g, 0.1 * i)
        cg.refresh_metrics(g)
        ts.append(g.last_eval_ms())
    ncell = (n + 10) ** 3
    out[label] = {"ms": min(ts), "Mcells_per_s": ncell / min(ts) / 1e3}
print(json.dumps(out))

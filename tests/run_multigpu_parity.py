"""Multi-GPU parity script (launch with torchrun, one rank per GPU, NCCL):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29511 tests/run_multigpu_parity.py

Every rank owns a k-slab of a 3-D DiscreteGrid, knows only its OWN node planes, receives the
neighbour planes it needs through NCCL send/recv (slabs.exchange_node_planes), evaluates its slab
and checks it against the same planes of the whole grid evaluated on its own GPU."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import curvigrid_b200 as cg  # noqa: E402
from util import assert_metrics_close, grid_outputs, mild_nodes  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
    h = 5
    nodes = mild_nodes((21, 17, 8 * world + 9), seed=77)
    ni, nj, nk = (s - 1 for s in nodes[0].shape)
    for profile in ("full", "compact"):
        plan = cg.slabs.make_plan(rank, world, nk, h)
        lo, hi = plan.required
        olo, ohi = plan.owned
        local = []
        for a in nodes:
            t = torch.from_numpy(np.ascontiguousarray(a.transpose(2, 1, 0)[lo:hi])).cuda()
            t[: olo - lo] = float("nan")   # not mine: must arrive through the exchange
            t[ohi - lo:] = float("nan")
            local.append(t)
        # handle first (its node buffer aliases `local`), then ONE overlapped step: exchange posted, interior planes
        # padded + evaluated while the neighbour planes travel, boundary planes after the wait
        g = cg.DiscreteGrid(*local, h, profile=profile, cache_mode="lazy",
                            **cg.slabs.slab_grid_kwargs(plan, (ni, nj), h))
        sched = cg.slabs.step_overlapped(g, plan, local)
        assert world == 1 or sched["eval_late"], sched
        whole = cg.DiscreteGrid(*nodes, h, profile=profile)
        k0, k1 = plan.cells
        plane = (ni + 2 * h) * (nj + 2 * h)
        if profile == "full":
            got, ref = grid_outputs(cg, g), grid_outputs(cg, whole)
            want = {k: ref[k][..., k0 * plane:k1 * plane, :] for k in ("cell_fwd", "cell_inv", "face_fwd", "face_inv", "face_cons")}
            assert_metrics_close(got, want, rtol=5e-13, label=f"rank {rank}")  # chunk boundaries differ (1-ulp FMA contraction, amplified by the cancellation in the conserved metrics)
        else:
            a, b = cg.cell_metrics(g).cpu().numpy(), cg.cell_metrics(whole).cpu().numpy()[k0:k1]
            assert np.allclose(a, b, rtol=5e-13, atol=0)
        res = torch.tensor(cg.gcl(g), device="cuda", dtype=torch.float64)
        dist.all_reduce(res, op=dist.ReduceOp.MAX)   # GCL max over all slabs (SURVEY.md §8e reductions)
        assert float(res.max()) < 1e-13, res
        if rank == 0:
            print(f"multi-GPU parity OK: world={world} profile={profile} gcl={res.tolist()}", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

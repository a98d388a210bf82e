"""Multi-GPU parity script (launch with torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29511 tests/run_multigpu_parity.py

Every rank owns a k-slab of a 3-D DiscreteGrid, knows only its OWN node planes and gets the neighbour planes it needs
through NCCL *inside the C ABI* (cg_grid_step_overlapped / cg_halo_exchange_begin+end / cg_grid_update_from_host_dist;
torch.distributed only broadcasts the 128-byte NCCL id).  Checked per rank:
  * the slab equals the same planes of the whole grid evaluated on the rank's own GPU,
  * sampled cells against the oracle,
  * the GCL maximum over all slabs INCLUDING the seam planes (cg_grid_gcl_max_dist) equals the whole grid's,
  * a corrupted seam face on rank r is seen through rank r+1's plane 0.
tests/test_multigpu_launcher.py runs this under pytest -m gpu when the box has >= 2 GPUs."""
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
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    comm = cg.slabs.NcclComm(rank, world, local)
    from oracle import oracle as O
    O.set_num_threads(max(1, (os.cpu_count() or 8) // world))
    h = 5
    nodes = mild_nodes((21, 17, 8 * world + 9), seed=77)
    ni, nj, nk = (s - 1 for s in nodes[0].shape)
    nfull = (ni + 2 * h, nj + 2 * h, nk + 2 * h)
    plane = nfull[0] * nfull[1]
    for profile in ("full", "compact"):
        plan = cg.slabs.make_plan(rank, world, nk, h)
        lo, hi = plan.required
        olo, ohi = plan.owned
        k0, k1 = plan.cells
        host = []
        for a in nodes:
            t = np.ascontiguousarray(a.transpose(2, 1, 0)[lo:hi]).copy()
            t[: olo - lo] = np.nan   # not mine: must arrive through the exchange
            t[ohi - lo:] = np.nan
            host.append(t)
        g = cg.DiscreteGrid(*[t.transpose(2, 1, 0) for t in host], h, profile=profile, cache_mode="lazy",
                            **cg.slabs.slab_grid_kwargs(plan, (ni, nj), h))
        whole = cg.DiscreteGrid(*nodes, h, profile=profile)
        ref = grid_outputs(cg, whole) if profile == "full" else None

        def check(label):
            if profile == "full":
                got = grid_outputs(cg, g)
                want = {k: ref[k][..., k0 * plane:k1 * plane, :] for k in ("cell_fwd", "cell_inv", "face_fwd", "face_inv", "face_cons")}
                # chunk boundaries differ (1-ulp FMA contraction, amplified by the cancellation in the conserved metrics)
                assert_metrics_close(got, want, rtol=5e-13, label=f"rank {rank} {label}")
                assert not any(np.isnan(got[k]).any() for k in got if k != "nfull")
            else:
                a, b = cg.cell_metrics(g).cpu().numpy(), cg.cell_metrics(whole).cpu().numpy()[k0:k1]
                assert np.allclose(a, b, rtol=5e-13, atol=0), label
                for ax in range(3):
                    a, b = cg.face_metrics(g)[ax].cpu().numpy(), cg.face_metrics(whole)[ax].cpu().numpy()[k0:k1]
                    assert np.abs(a - b).max() <= 5e-13 * np.abs(b).max(), label

        # (1) one overlapped step through the C ABI
        cg.slabs.step_overlapped_nccl(g, plan, comm)
        check("step_overlapped")
        # (2) the plain sequence: begin / end / nodes_changed / eval
        L, lib = cg._lib, cg._lib.lib()
        L.check(lib.cg_halo_exchange_begin(g._h, comm._h, rank, world, comm.bounds(plan), g._stream()))
        L.check(lib.cg_halo_exchange_end(g._h, g._stream()))
        L.check(lib.cg_grid_nodes_changed(g._h, g._stream()))
        cg.refresh_metrics(g)
        check("begin/end")
        # (3) the exchange-aware pipelined host update, from pinned host arrays whose non-owned planes are NaN
        pinned = [torch.from_numpy(t).pin_memory() for t in host]
        for nchunks in (1, 3, 16):
            g.update_from_host(*pinned, nchunks=nchunks, comm=comm, plan=plan)
            check(f"update_from_host_dist nchunks={nchunks}")
        # (4) GCL over all slabs, seam planes included
        res = np.array(cg.slabs.gcl_dist(g, comm))
        single = np.array(cg.gcl(whole))
        assert res.max() < 1e-13 and np.abs(res - single).max() < 1e-15, (res, single)
        # (5) sampled cells against the oracle (FULL)
        if profile == "full":
            rng = np.random.default_rng(rank)
            sample = np.sort(rng.choice((k1 - k0) * plane, size=24, replace=False)) + k0 * plane
            o = O.discrete_eval(nodes, h, "curvature", sample=sample)
            got = grid_outputs(cg, g)
            pick = {k: got[k][..., sample - k0 * plane, :] for k in ("cell_fwd", "cell_inv", "face_fwd", "face_inv", "face_cons")}
            assert_metrics_close(pick, o, rtol=1e-12, label=f"rank {rank} oracle")
        # (6) a corrupted seam face: the last conserved zeta-plane of rank r is the low face of rank r+1's plane 0
        if world > 1:
            fm = cg.face_metrics(g)[2]
            t = fm if profile == "compact" else fm.conserved
            inside = k1 > h and k1 < nfull[2] - h      # the seam lies inside cell.domain along k
            if rank == 0:
                assert inside
                t[-1, h + 2, h + 3, 2 if profile == "compact" else 2 + 3 * 2] += 0.5
            torch.cuda.synchronize()
            seen = np.array(cg.slabs.gcl_dist(g, comm))
            assert seen[2] > 0.4, seen
            # without the seam plane rank 1 is blind to it (rank 0 still sees the face as the top of its own last cell)
            own = cg.gcl(g)
            if rank == 1:
                assert own[2] < 1e-13, own
        if rank == 0:
            print(f"multi-GPU parity OK: world={world} profile={profile} gcl={res.tolist()}", flush=True)
        g.close()
        whole.close()
    # multiblock interface with the two blocks on different ranks (rank 0: left block, rank 1: right block): scalar
    # and vector fields, packed values over NCCL inside the C ABI, transforms applied while unpacking
    if world >= 2 and rank < 2:
        mbk, wl = cg.multiblock, cg.workloads
        hh, cd = 2, (5, 6, 4)
        left = cg.MappedGrid(wl.sphere_sector_3d_terms(cd), {}, cd, hh, cache_mode="lazy")
        right = cg.MappedGrid(wl.spherical_rtp_3d_terms(cd), {}, cd, hh, cache_mode="lazy", coordinate_system="SphericalCS",
                              basis="SphericalBasis")
        full = tuple(c + 2 * hh for c in cd)
        dom = tuple((hh + 1, n - hh) for n in full)
        plan = mbk.interface_index_plan(dom, mbk.BlockFace(1, 1, "max"), dom, mbk.BlockFace(2, 1, "min"), (1, 2), (False, False))
        dplan = mbk.DevicePlan(plan, full, full, "cuda").build_transforms(left, right)
        gen = torch.Generator(device="cpu").manual_seed(5)
        fl = torch.randn(tuple(reversed(full)) + (3,), generator=gen, dtype=torch.float64).cuda()
        fr = torch.randn(tuple(reversed(full)) + (3,), generator=gen, dtype=torch.float64).cuda()
        want_l, want_r = fl.clone(), fr.clone()
        mbk.exchange_interface(want_l, want_r, dplan, field_kind="vector")          # both blocks on one GPU
        mine = fl if rank == 0 else fr
        mbk.exchange_interface_remote(mine, dplan, "left" if rank == 0 else "right", comm, 1 - rank, field_kind="vector")
        assert torch.equal(mine, want_l if rank == 0 else want_r)
        sl, sr = fl[..., 0].contiguous(), fr[..., 0].contiguous()
        ws_l, ws_r = sl.clone(), sr.clone()
        mbk.exchange_interface(ws_l, ws_r, dplan)
        ms = sl if rank == 0 else sr
        mbk.exchange_interface_remote(ms, dplan, "left" if rank == 0 else "right", comm, 1 - rank)
        assert torch.equal(ms, ws_l if rank == 0 else ws_r)
        if rank == 0:
            print("multiblock remote exchange OK (scalar + vector)", flush=True)
    dist.barrier()
    comm.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

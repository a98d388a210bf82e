#!/usr/bin/env python
"""bench.py — metric-eval throughput (BASELINE.json `metric`) on 1..8 B200.

A "step" is one pass of the hot path over one DiscreteGrid: Line() padding of the
node arrays + the fused cell+face metric kernel (+ the node-plane halo exchange
between k-slabs when N > 1).

  workload (default)  BASELINE configs[2]: 3-D DiscreteGrid 512^3 wavy mesh
                      (test_3d.jl:153-178 generator, L=4), nhalo=5, CurvilinearCS,
                      CurvatureCorrected, FULL cell + conserved face metrics, FP64.
                      N > 1: the grid grows along k (512 x 512 x 512*N cells), one
                      k-slab per GPU, node planes exchanged with NCCL send/recv
                      every step (weak scaling).
  value               Mcells/s, cells = entries of cell.full written, inputs resident
                      in HBM, device-timed (CUDA events), max over ranks.
  e2e                 same metric through the public API with HOST buffers: pinned
                      host nodes -> H2D -> (exchange) -> pad -> metrics -> GCL max
                      reduced on device -> D2H of the residuals, all inside the
                      timed region.
  other workloads     --workload 2d4096 (BASELINE configs[1], GradientCorrected), m3d512 (configs[3]: MappedGrid,
                      update!(grid, t) + both metric caches every step), sph1536 --profile compact (configs[4]:
                      spherical-basis DiscreteGrid, 1536 x 1536 x 192 cells per GPU, k-slabs, weak scaling)
  --impl reference    the CPU oracle (restatement of the reference algorithm; the
                      Julia package cannot run here) on all host threads over a
                      bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BYTES_PER_CELL = {  # algorithmic bytes per cell
    ("3d", "full"): 24 + 160 + 696,
    ("3d", "compact"): 24 + 80,
    ("2d", "full"): 16 + 80 + 224,
    ("2d", "compact"): 16 + 8 + 32,
    ("m3d", "full"): 24 + 24 + 72 + 856,  # MappedGrid update!: node/centroid/face coordinates + metrics, no node read
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--h2d-chunks", type=int, default=16, help="pieces of the pipelined host->device update (e2e)")
    ap.add_argument("--workload", default="3d512",
                    help="3d<N> | 2d<N> DiscreteGrid, m3d<N> MappedGrid with update! every step (interior cells per axis)")
    ap.add_argument("--scheme", default=None)
    ap.add_argument("--profile", default="full", choices=["full", "compact"])
    ap.add_argument("--dtype", default="f64", choices=["f64", "f32"])
    ap.add_argument("--kernel", type=int, default=1, help="1 tiled (default), 0 gather reference kernels")
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="target CPU time of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--strong", action="store_true",
                    help="strong scaling: the slab axis keeps N cells in total (N^3 split over the GPUs) instead of "
                         "one slab per GPU; the JSON line says \"scaling\": \"strong\"")
    return ap.parse_args()


def workload_spec(name, scheme):
    if name.startswith("3d"):
        n = int(name[2:])
        return dict(kind="3d", n=n, nhalo=5, scheme=scheme or "curvature",
                    desc=f"3-D DiscreteGrid {n}^3 wavy (test_3d.jl:153-178), nhalo=5, cell+face")
    if name.startswith("2d"):
        n = int(name[2:])
        return dict(kind="2d", n=n, nhalo=5, scheme=scheme or "gradient",
                    desc=f"2-D DiscreteGrid {n}^2 wavy (test_2d.jl:154-181), nhalo=5, cell+face")
    if name.startswith("sph"):
        # BASELINE config 5: spherical-basis DiscreteGrid ((r, theta, phi) node coordinates), k-slabs; per GPU
        # sph<N> = N x N x N/8 cells (sph1536: 1536 x 1536 x 192), COMPACT profile unless --profile full is given
        n = int(name[3:])
        return dict(kind="3d", gen="sph", n=n, nslab=max(1, n // 8), nhalo=5, scheme=scheme or "curvature",
                    desc=f"3-D spherical-basis DiscreteGrid {n}x{n}x{max(1, n // 8)} cells per GPU ((r,theta,phi) nodes of "
                         f"test_mapped_discrete_face_geometry.jl:196-226), nhalo=5, cell+face")
    if name.startswith("m3d"):
        n = int(name[3:])
        return dict(kind="m3d", n=n, nhalo=5, scheme=scheme or "curvature",
                    desc=f"3-D MappedGrid {n}^3 time-dependent wavy mapping (test_3d_continuous.jl:10-53, amplitudes x "
                         f"(1+0.1 sin t)), update!(grid,t) + cell+face metrics every step, nhalo=5")
    raise SystemExit(f"unknown workload {name}")


# ---------------------------------------------------------------------------
# node generators for one slab (numpy, host) — the reference's wavy meshes
# ---------------------------------------------------------------------------
def wavy_slab_3d(ni, nj, nk_total, k_lo, k_hi, L=4.0):
    """Node planes [k_lo,k_hi) of wavy_nodes_3d(ni,nj,nk_total) in memory order [k,j,i]."""
    dx, dy, dz = L / ni, L / nj, L / nk_total
    i = np.arange(ni, dtype=np.float64)[None, None, :]
    j = np.arange(nj, dtype=np.float64)[None, :, None]
    k = np.arange(k_lo, k_hi, dtype=np.float64)[:, None, None]
    sp = lambda a: np.sin(np.pi * a)
    shp = (k_hi - k_lo, nj, ni)
    x = -L / 2 + dx * (i + sp(j * dy) * sp(k * dz))
    y = -L / 2 + dy * (j + sp(k * dz) * sp(i * dx))
    z = -L / 2 + dz * (k + sp(i * dx) * sp(j * dy))
    return [np.ascontiguousarray(np.broadcast_to(a, shp)) for a in (x, y, z)]


def sph_slab_3d(ni, nj, nk_total, k_lo, k_hi):
    """Node planes [k_lo,k_hi) of the perturbed (r, theta, phi) mesh of workloads.spherical_rtp_3d_terms, memory order
    [k,j,i]; ni, nj, nk_total are node counts."""
    ci, cj, ck = ni - 1, nj - 1, nk_total - 1
    r0, th0, ph0 = 1.2, 0.45, 0.25
    dr, dth, dph = 0.4 / ci, 0.35 / cj, 0.45 / ck
    i = np.arange(ni, dtype=np.float64)[None, None, :]
    j = np.arange(nj, dtype=np.float64)[None, :, None]
    k = np.arange(k_lo, k_hi, dtype=np.float64)[:, None, None]
    s2i, s2j, s1k = np.sin(2 * np.pi * i / ci), np.sin(2 * np.pi * j / cj), np.sin(np.pi * k / ck)
    shp = (k_hi - k_lo, nj, ni)
    r = r0 + dr * i + 0.012 * s2j * s1k
    th = th0 + dth * j + 0.008 * s2i
    ph = ph0 + dph * k + 0.006 * s2i * s2j
    return [np.ascontiguousarray(np.broadcast_to(a, shp)) for a in (r, th, ph)]


def wavy_slab_2d(nx, ny_total, j_lo, j_hi, a0=0.1):
    i = np.arange(nx, dtype=np.float64)[None, :]
    j = np.arange(j_lo, j_hi, dtype=np.float64)[:, None]
    x1d, y1d = i / (nx - 1), j / (ny_total - 1)
    w = a0 * np.sin(2 * np.pi * x1d) * np.sin(2 * np.pi * y1d)
    shp = (j_hi - j_lo, nx)
    return [np.ascontiguousarray(np.broadcast_to(x1d + w, shp)), np.ascontiguousarray(np.broadcast_to(y1d + w, shp))]


# ---------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [s.strip() for s in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


# ---------------------------------------------------------------------------
def cpu_oracle_rate(spec, nodes_host_global, seconds, nthreads=None):
    """Oracle (CPU restatement of the reference) on a bounded random sample of cell.full."""
    from oracle import oracle as O
    if nthreads:
        O.set_num_threads(nthreads)
    cores = O.num_threads()
    dim = 3 if spec["kind"] == "3d" else 2
    nodes = nodes_host_global
    nfull = [s - 1 + 2 * spec["nhalo"] for s in nodes[0].shape]
    ntot = int(np.prod(nfull))
    rng = np.random.default_rng(0)
    # calibrate
    probe = np.sort(rng.choice(ntot, size=min(ntot, 64 * cores), replace=False))
    t0 = time.perf_counter()
    O.discrete_eval(nodes, spec["nhalo"], spec["scheme"], sample=probe)
    dt = max(time.perf_counter() - t0, 1e-6)
    rate = len(probe) / dt
    n = int(min(ntot, max(len(probe), rate * seconds)))
    sample = np.sort(rng.choice(ntot, size=n, replace=False))
    t0 = time.perf_counter()
    O.discrete_eval(nodes, spec["nhalo"], spec["scheme"], sample=sample)
    dt = time.perf_counter() - t0
    return n / dt / 1e6, cores, f"{n} random cells of cell.full ({dim}-D, {spec['scheme']}), {dt:.1f} s wall"


def global_nodes_for_cpu(spec, nranks):
    """Host node arrays of the whole workload as numpy [i,j,k]-indexed views (for the oracle)."""
    n = spec["n"]
    if spec.get("gen") == "sph":
        ns = spec["nslab"]
        arrs = sph_slab_3d(n + 1, n + 1, ns * nranks + 1, 0, ns * nranks + 1)
        return [a.transpose(2, 1, 0) for a in arrs]
    if spec["kind"] == "3d":
        arrs = wavy_slab_3d(n + 1, n + 1, n * nranks + 1, 0, n * nranks + 1)
        return [a.transpose(2, 1, 0) for a in arrs]
    arrs = wavy_slab_2d(n + 1, n * nranks + 1, 0, n * nranks + 1)
    return [a.transpose(1, 0) for a in arrs]


def run_reference(args, spec):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # bounded sample per step so that K steps + W warm-ups end within minutes
    nodes = global_nodes_for_cpu(spec, 1 if spec["kind"] == "3d" and spec["n"] > 256 else 1)
    per_step = max(1.0, min(args.cpu_seconds, 120.0 / max(1, args.steps + args.warmup)))
    vals, sample_desc, cores = [], "", 1
    for s in range(args.warmup + args.steps):
        v, cores, sample_desc = cpu_oracle_rate(spec, nodes, per_step)
        if s >= args.warmup:
            vals.append(v)
    value = float(np.mean(vals))
    line = {
        "impl": "reference", "metric": "metric-eval Mcells/s (cell+face, FP64)", "value": value, "unit": "Mcells/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": per_step * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": spec["desc"], "scheme": spec["scheme"], "profile": "full",
                   "note": "CPU oracle = C++ restatement of the reference algorithm (nested forward-mode AD through "
                           "the interpolant); the Julia package cannot run here (no Julia toolchain)"},
        "cpu_baseline": {"value": value, "unit": "Mcells/s", "cores": cores, "kind": "port", "sample": sample_desc},
        "e2e": {"value": value, "unit": "Mcells/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------
def run_ours(args, spec):
    import torch
    import torch.distributed as dist

    import curvigrid_b200 as cg

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local_rank)
    dev = torch.device(f"cuda:{local_rank}")
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    n, h = spec["n"], spec["nhalo"]
    dim = 3 if spec["kind"] == "3d" else 2
    T = np.float64 if args.dtype == "f64" else np.float32
    tdt = torch.float64 if args.dtype == "f64" else torch.float32
    nslab = spec.get("nslab", n)
    ncell_slab = nslab * world  # weak scaling: the grid grows along the slab axis
    if args.strong:
        ncell_slab = n          # strong scaling: n cells along the slab axis in total, split over the ranks
        nslab = n // world
    plan = cg.slabs.make_plan(rank, world, ncell_slab, h)
    lo, hi = plan.required
    if dim == 3:
        gen = sph_slab_3d if spec.get("gen") == "sph" else wavy_slab_3d
        host = gen(n + 1, n + 1, ncell_slab + 1, lo, hi)
        inplane = (n, n)
    else:
        host = wavy_slab_2d(n + 1, ncell_slab + 1, lo, hi)
        inplane = (n,)
    host_pinned = [torch.from_numpy(a.astype(T)).pin_memory() for a in host]
    # truth for non-owned planes comes from their owners; blank them so the exchange is real
    olo, ohi = plan.owned
    dev_nodes = [t.to(dev, non_blocking=True) for t in host_pinned]
    if world > 1:
        for t in dev_nodes:
            t[:olo - lo].zero_()
            t[ohi - lo:].zero_()
    kw = cg.slabs.slab_grid_kwargs(plan, inplane, h) if world > 1 else {}
    grid = cg.DiscreteGrid(*dev_nodes, h, T=T, conserved_metric_scheme=spec["scheme"], cache_mode="lazy",
                           profile=args.profile, **kw)
    grid.set_kernel(args.kernel)
    lib = cg._lib.lib()
    import ctypes
    cells_local = int(np.prod(grid.window_cells))
    cells_total = cells_local
    if world > 1:
        tcount = torch.tensor([cells_local], device=dev, dtype=torch.int64)
        dist.all_reduce(tcount)
        cells_total = int(tcount.item())

    overlap = world > 1 and dim == 3 and args.kernel == 1

    def step_plain():
        if world > 1:
            cg.slabs.exchange_node_planes(plan, dev_nodes)
        cg._lib.check(lib.cg_grid_nodes_changed(grid._h, grid._stream()))
        cg.refresh_metrics(grid)

    def step_resident():
        if overlap:
            # exchange posted first; planes that depend on owned nodes only are padded + evaluated while the
            # neighbour planes travel over NVLink; the 1 (low) / 3 (high) boundary planes follow the wait
            cg.slabs.step_overlapped(grid, plan, dev_nodes)
        else:
            step_plain()

    pipelined = world == 1 and dim == 3 and args.kernel == 1

    def step_e2e():
        if pipelined:
            # cg_grid_update_from_host: node planes in pieces, the H2D copy of a piece overlaps the kernels of the
            # piece before (pinned host arrays -> padded nodes -> metrics)
            grid.update_from_host(*host_pinned, nchunks=args.h2d_chunks)
        else:
            for d, hsrc in zip(dev_nodes, host_pinned):
                d[olo - lo:ohi - lo].copy_(hsrc[olo - lo:ohi - lo], non_blocking=True)
            step_resident()
        res = cg.gcl(grid)  # device reduction + D2H of the residuals (synchronises)
        return res

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            tm = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
            ms = float(tm.item())
        return ms

    for _ in range(max(args.warmup, 3)):
        step_resident()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = lib.cg_launch_count()
    kernel_ms = []
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step_resident()
        kernel_ms.append(None)
    e1.record()
    barrier()
    launches = lib.cg_launch_count() - launches0
    total_ms = e0.elapsed_time(e1)
    clocks = sampler.stop()
    if world > 1:
        tm = torch.tensor([total_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        total_ms = float(tm.item())
    ms_per_step = total_ms / args.steps
    value = cells_total / (ms_per_step * 1e-3) / 1e6

    # dominant kernel: the fused metric kernel, timed with CUDA events on its own stream
    kms = []
    for _ in range(min(args.steps, 5)):
        step_plain()  # one cg_grid_eval over the whole slab (the overlapped step splits it into plane ranges)
        kms.append(grid.last_eval_ms())
    k_ms = float(np.mean(kms))
    bpc = BYTES_PER_CELL[(spec["kind"], args.profile)] * (1.0 if args.dtype == "f64" else 0.5)
    peak, peak_src = measured_peak_gbs()
    achieved = bpc * cells_local / (k_ms * 1e-3) / 1e9
    traffic = None
    try:  # DRAM bytes of the metric kernels from the committed `ncu --set full` capture of this workload
        with open(os.path.join(ROOT, "profiles", "ncu_traffic_3d512.json")) as f:
            tj = json.load(f)
        if (args.workload, args.profile, args.dtype, spec["scheme"], args.kernel) == ("3d512", "full", "f64", "curvature", 1):
            traffic = tj["total"]
    except Exception:
        pass
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "algorithmic_bytes": bpc * cells_local,
                "kernel": ("k_face_xi + k_face_zeta<SWAP> (eta) + k_face_zeta (one cell+face evaluation = 3 launches)"
                           if dim == 3 and args.kernel == 1 else ("metrics3d" if dim == 3 else "metrics2d")),
                "kernel_ms": k_ms,
                "algorithmic_bytes_per_cell": bpc, "peak_source": f"MEASURED_PEAKS.json ({peak_src}, copy burst)",
                # also against the nominal ~8 TB/s of the north star
                "frac_of_nominal_8TBs": achieved / 8000.0}
    if dim == 3 and args.kernel == 1 and args.profile == "full" and args.dtype == "f64" and spec["scheme"] == "curvature":
        # second ceiling: FP64 instructions of the three marching kernels per cell (SASS of the loop
        # bodies: xi 736 per 29 cells, eta / zeta 697 per 31) against the measured FP64 issue rate (59 per clock and SM)
        ipc = 32.0 * (736.0 / 29.0 + 2.0 * 697.0 / 31.0)  # thread instructions: a warp instruction occupies 32 lanes
        rate = ipc * cells_local / (k_ms * 1e-3)
        roofline["fp64"] = {"instr_per_cell": ipc, "achieved_Tinstr_s": rate / 1e12,
                            "peak_Tinstr_s": 59.0 * 148 * 1.965e9 / 1e12,
                            "frac": rate / (59.0 * 148 * 1.965e9), "note": "peak at the maximum SM clock"}

    e2e = None
    if not args.no_e2e:
        for _ in range(2):
            step_e2e()
        ms = timed(step_e2e, args.steps) / args.steps
        h2d = sum(int(t[olo - lo:ohi - lo].numel() * t.element_size()) for t in host_pinned)
        e2e = {"value": cells_total / (ms * 1e-3) / 1e6, "unit": "Mcells/s", "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": 8 * dim, "ms_per_step": ms,
               "note": ("pinned host nodes -> H2D in %d pieces overlapped with pad + metrics of the piece before "
                        "(cg_grid_update_from_host)" % args.h2d_chunks if pipelined else
                        "pinned host nodes -> H2D -> pad -> metrics") + " -> device GCL max -> D2H; metric arrays "
                       "stay on the device like a KernelAbstractions GPU backend's storage"}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        nodes_cpu = [a.transpose(*reversed(range(dim))) for a in host]  # [i,j,k] views of the full node arrays
        v, cores, desc = cpu_oracle_rate(spec, nodes_cpu, args.cpu_seconds)
        cpu = {"value": v, "unit": "Mcells/s", "cores": cores, "kind": "port", "sample": desc}

    if rank == 0:
        line = {
            "metric": "metric-eval Mcells/s (cell+face, FP64)" if args.dtype == "f64" else "metric-eval Mcells/s (cell+face, FP32)",
            "value": value, "unit": "Mcells/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong" if args.strong else "weak", "vs_baseline": None,
            "dtype": args.dtype, "data": "synthetic",
            "config": {"workload": spec["desc"] + (f"; k-slabs over {world} GPUs, grid {n}x{n}x{ncell_slab}" if world > 1 and dim == 3 else ""),
                       "scheme": spec["scheme"], "profile": args.profile, "cells_written": cells_total,
                       "kernel": "tiled" if args.kernel == 1 else "gather",
                       "l2": "inputs+outputs per step far exceed the 126 MB L2 (no flush needed)",
                       "parallelism": f"slab{world}" if world > 1 else "single",
                       "halo_exchange": ("NCCL node planes, overlapped with interior planes" if overlap else
                                         ("NCCL node planes" if world > 1 else "none"))},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches),
            "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_mapped(args, spec):
    """BASELINE config 4: MappedGrid, update!(grid, t_k) + refresh of both metric caches every step (1 GPU)."""
    import torch

    import curvigrid_b200 as cg

    torch.cuda.set_device(0)
    n, h = spec["n"], spec["nhalo"]
    terms = cg.workloads.wavy_3d_terms((n, n, n), time_amp=0.1)
    grid = cg.MappedGrid(terms, {}, (n, n, n), h, cache_mode="lazy", conserved_metric_scheme=spec["scheme"])
    lib = cg._lib.lib()
    cells = int(np.prod(grid.window_cells))
    tk = [0]

    def step():
        tk[0] += 1
        cg.update(grid, 0.01 * tk[0])
        cg.refresh_metrics(grid)

    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()
    sampler = ClockSampler(0)
    sampler.start()
    launches0 = lib.cg_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    launches = lib.cg_launch_count() - launches0
    ms_per_step = e0.elapsed_time(e1) / args.steps
    clocks = sampler.stop()
    kms = []
    for _ in range(min(args.steps, 5)):
        step()
        kms.append(grid.last_eval_ms())
    k_ms = float(np.mean(kms))
    peak, peak_src = measured_peak_gbs()
    bpc = 856.0  # the metric kernel writes the FULL profile and reads only the 1-D tables
    achieved = bpc * cells / (k_ms * 1e-3) / 1e9
    line = {
        "metric": "metric-eval Mcells/s (cell+face, FP64)", "value": cells / (ms_per_step * 1e-3) / 1e6, "unit": "Mcells/s",
        "n_gpus": 1, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": spec["desc"], "scheme": spec["scheme"], "profile": "full", "cells_written": cells,
                   "l2": "outputs per step far exceed the 126 MB L2 (no flush needed)", "parallelism": "single"},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": None, "kernel": "k_mapped_metrics3d_staged", "kernel_ms": k_ms,
                     "algorithmic_bytes_per_cell": bpc, "peak_source": f"MEASURED_PEAKS.json ({peak_src}, copy burst)"},
        "cpu_baseline": None, "e2e": None, "gpu_launches": int(launches), "clocks": clocks,
    }
    print(json.dumps(line), flush=True)


def main():
    args = parse_args()
    spec = workload_spec(args.workload, args.scheme)
    if args.impl == "reference":
        run_reference(args, spec)
    elif spec["kind"] == "m3d":
        run_mapped(args, spec)
    else:
        run_ours(args, spec)


if __name__ == "__main__":
    main()

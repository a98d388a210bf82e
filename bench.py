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
    ap.add_argument("--h2d-chunks", type=int, default=32, help="pieces of the pipelined host->device update (e2e)")
    ap.add_argument("--workload", default="3d512",
                    help="3d<N> | 2d<N> DiscreteGrid, m3d<N> MappedGrid with update! every step (interior cells per axis)")
    ap.add_argument("--scheme", default=None)
    ap.add_argument("--profile", default="full", choices=["full", "compact"])
    ap.add_argument("--dtype", default="f64", choices=["f64", "f32"])
    ap.add_argument("--kernel", type=int, default=1, help="1 tiled (default), 0 gather reference kernels")
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="target CPU time of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-others", action="store_true", help="skip the short runs of the other BASELINE configs (other_configs)")
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
def _full(a, shp):
    """A writable C-contiguous array of shape `shp` holding `a` (broadcast)."""
    a = np.asarray(a)
    if a.shape == tuple(shp) and a.flags.c_contiguous and a.flags.writeable:
        return a
    out = np.empty(shp, dtype=a.dtype)
    out[...] = a
    return out


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
    return [_full(a, shp) for a in (x, y, z)]


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
    return [_full(a, shp) for a in (r, th, ph)]


def wavy_slab_2d(nx, ny_total, j_lo, j_hi, a0=0.1):
    i = np.arange(nx, dtype=np.float64)[None, :]
    j = np.arange(j_lo, j_hi, dtype=np.float64)[:, None]
    x1d, y1d = i / (nx - 1), j / (ny_total - 1)
    w = a0 * np.sin(2 * np.pi * x1d) * np.sin(2 * np.pi * y1d)
    shp = (j_hi - j_lo, nx)
    return [_full(x1d + w, shp), _full(y1d + w, shp)]


# ---------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []
        self.nvml, self.samples, self._stop = None, [], threading.Event()

    def _nvml_loop(self):
        n, h = self.nvml
        while not self._stop.is_set():
            try:
                self.samples.append((n.nvmlDeviceGetClockInfo(h, n.NVML_CLOCK_SM), n.nvmlDeviceGetMaxClockInfo(h, n.NVML_CLOCK_SM),
                                     n.nvmlDeviceGetCurrentClocksEventReasons(h)))
            except Exception:
                break
            time.sleep(0.005)

    def start(self):
        # NVML in-process (a sample every ~5 ms: dozens inside a 0.3 s timed region); nvidia-smi -lms 100 as the fallback
        try:
            import pynvml as n
            n.nvmlInit()
            h = None
            try:  # the CUDA device of this rank by UUID (NVML counts physical GPUs, CUDA_VISIBLE_DEVICES may renumber them)
                import torch
                u = str(torch.cuda.get_device_properties(torch.cuda.current_device()).uuid)
                h = n.nvmlDeviceGetHandleByUUID(("GPU-" + u) if not u.startswith("GPU-") else u)
            except Exception:
                h = None
            self.nvml = (n, h if h is not None else n.nvmlDeviceGetHandleByIndex(self.index))
            self.thread = threading.Thread(target=self._nvml_loop, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
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
        if self.nvml:
            self._stop.set()
            self.thread.join(timeout=1)
            bits = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
                    0x80: "hw_power_brake_slowdown"}
            reasons = sorted({nm for _, _, r in self.samples for b, nm in bits.items() if r & b})
            sm = [a for a, _, _ in self.samples]
            return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_min_mhz": float(min(sm)) if sm else None,
                    "sm_max_mhz": float(max(m for _, m, _ in self.samples)) if sm else None, "reasons": reasons,
                    "samples": len(sm), "source": "nvml, 5 ms period"}
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
def _host_threads():
    """All host threads, whatever OMP_NUM_THREADS says (torchrun sets it to 1, which shrank the CPU arm 29x in r1)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def cpu_oracle_rate(spec, nodes_host_global, seconds):
    """Oracle (CPU restatement of the reference) on a bounded random sample of cell.full, all host threads."""
    from oracle import oracle as O
    O.set_num_threads(_host_threads())
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
    return n / dt / 1e6, cores, f"{n} random cells of cell.full ({dim}-D, {spec['scheme']}), {dt:.1f} s wall", ntot


def global_nodes_for_cpu(spec):
    """Host node arrays of the one-GPU workload as numpy [i,j,k]-indexed views (for the oracle)."""
    n = spec["n"]
    if spec.get("gen") == "sph":
        ns = spec["nslab"]
        arrs = sph_slab_3d(n + 1, n + 1, ns + 1, 0, ns + 1)
        return [a.transpose(2, 1, 0) for a in arrs]
    if spec["kind"] == "3d":
        arrs = wavy_slab_3d(n + 1, n + 1, n + 1, 0, n + 1)
        return [a.transpose(2, 1, 0) for a in arrs]
    arrs = wavy_slab_2d(n + 1, n + 1, 0, n + 1)
    return [a.transpose(1, 0) for a in arrs]


def run_reference(args, spec):
    """The reference arm: the CPU oracle ("port": the Julia package cannot run here) on ALL host threads over a bounded
    sample of the one-GPU workload.  Under torchrun rank 0 alone runs; the other ranks exit 0 without work."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if spec["kind"] == "m3d":
        spec = workload_spec("3d%d" % spec["n"], spec["scheme"])
    nodes = global_nodes_for_cpu(spec)
    per_step = max(1.0, min(args.cpu_seconds, 120.0 / max(1, args.steps + args.warmup)))
    vals, sample_desc, cores, ntot = [], "", 1, 1
    for s in range(args.warmup + args.steps):
        v, cores, sample_desc, ntot = cpu_oracle_rate(spec, nodes, per_step)
        if s >= args.warmup:
            vals.append(v)
    value = float(np.mean(vals))
    line = {
        "impl": "reference", "metric": "metric-eval Mcells/s (cell+face, FP64)", "value": value, "unit": "Mcells/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        # one step of the workload = every cell of cell.full: extrapolated from the sampled rate
        "ms_per_step": ntot / (value * 1e6) * 1e3, "sample_seconds_per_step": per_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": spec["desc"], "scheme": spec["scheme"], "profile": "full",
                   "note": "CPU oracle = C++ restatement of the reference algorithm (nested forward-mode AD through "
                           "the interpolant); the Julia package cannot run here (no Julia toolchain).  value = sampled "
                           "cells per second on all host threads; ms_per_step = the whole step extrapolated from it"},
        "cpu_baseline": {"value": value, "unit": "Mcells/s", "cores": cores, "kind": "port", "sample": sample_desc},
        "e2e": {"value": value, "unit": "Mcells/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------
class Ctx:
    """Process-wide state of one bench run: ranks, device, the NCCL communicator of the C ABI."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if self.world != args.gpus and self.world > 1:
            raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={self.world}")
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device(f"cuda:{self.local_rank}")
        self.comm = None
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
            import curvigrid_b200 as cg
            # ncclComm_t owned by libcurvigrid_cuda.so: the node-plane exchange, the seam plane of the GCL functional and
            # its max-reduction are C-ABI calls; torch.distributed only carries the 128-byte id and the bench's timing
            self.comm = cg.slabs.NcclComm(self.rank, self.world, self.local_rank)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def allmax(self, x):
        if self.world == 1:
            return float(x)
        t = self.torch.tensor([float(x)], device=self.dev, dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def allsum_int(self, x):
        if self.world == 1:
            return int(x)
        t = self.torch.tensor([int(x)], device=self.dev, dtype=self.torch.int64)
        self.dist.all_reduce(t)
        return int(t.item())

    def close(self):
        if self.comm is not None:
            self.comm.close()
        if self.world > 1:
            self.dist.destroy_process_group()


def parity_block(cg, ctx, grid, host, plan, spec, profile, dtype, nsamples=64):
    """Sampled cells of this rank's slab against the oracle, inside the bench run: the CUDA values
    against the FP64 oracle (the reference's arithmetic) and both against the same closure graph in long double
    ("truth").  The oracle's grid is the rank's own node planes; sampled cells keep two planes away from slab seams so
    that the sub-grid's extrapolated halo is never what they read.  Max over ranks."""
    from oracle import oracle as O
    torch = ctx.torch
    O.set_num_threads(_host_threads())
    dim, h = grid.dim, spec["nhalo"]
    lo, hi = plan.required
    k0, k1 = plan.cells
    nw = list(grid.window_cells)
    rng = np.random.default_rng(1000 + ctx.rank)
    idx = [rng.integers(0, nw[d], nsamples) for d in range(dim - 1)]
    s_lo = 0 if ctx.rank == 0 else 2
    s_hi = nw[dim - 1] if ctx.rank == ctx.world - 1 else nw[dim - 1] - 3
    ks = rng.integers(s_lo, max(s_lo + 1, s_hi), nsamples)                     # window-local plane
    perm = list(range(dim - 1, -1, -1))
    sub_nodes = [np.asarray(a).transpose(*perm).astype(np.float64) for a in host]   # [i,j,k] views of the rank's planes
    nf = [nw[d] for d in range(dim - 1)] + [(hi - lo) - 1 + 2 * h]
    ko = ks + k0 - lo                                                          # the oracle grid's full cell plane
    lin_o = np.zeros(nsamples, dtype=np.int64)
    lin_w = np.zeros(nsamples, dtype=np.int64)
    so = sw = 1
    for d in range(dim):
        c = idx[d] if d < dim - 1 else None
        lin_o += (c if c is not None else ko) * so
        lin_w += (c if c is not None else ks) * sw
        so *= nf[d]
        sw *= nw[d]
    order = np.argsort(lin_o)
    lin_o, lin_w = lin_o[order], lin_w[order]
    o64 = O.discrete_eval(sub_nodes, h, spec["scheme"], sample=lin_o)
    truth = O.discrete_eval(sub_nodes, h, spec["scheme"], sample=lin_o, precision="ld") if dtype == "f64" else o64
    N = dim
    widx = torch.from_numpy(lin_w).to(ctx.dev)

    def rel(a, b, has_j):
        K = a.shape[-1] - (1 if has_j else 0)
        sc = np.abs(b[..., :K]).max(axis=-1, keepdims=True)
        sc = np.where(sc == 0, 1.0, sc)
        e = float((np.abs(a[..., :K] - b[..., :K]) / sc).max())
        if has_j:
            e = max(e, float((np.abs(a[..., K] - b[..., K]) / np.abs(b[..., K])).max()))
        return e

    # cell.domain (where BASELINE.md defines parity) and the halo region separately: halo nodes are Line()
    # extrapolations ROUNDED to the working precision, so both FP64 evaluations carry eps (|x|/dx)^2 there
    dom = np.ones(nsamples, dtype=bool)
    gidx = [idx[d][order] + grid.window_origin[d] for d in range(dim - 1)] + [ks[order] + k0]
    for d in range(dim):
        dom &= (gidx[d] >= h) & (gidx[d] < grid.nfull[d] - h)

    def region(a, b, has_j, m):
        return rel(a[..., m, :], b[..., m, :], has_j) if m.any() else 0.0

    out = {}
    if profile == "full":
        cm, fm = cg.cell_metrics(grid), cg.face_metrics(grid)
        pick = lambda t, k: t.reshape(-1, k)[widx].double().cpu().numpy()
        got = {"cell_fwd": pick(cm.forward, N * N + 1), "cell_inv": pick(cm.inverse, N * N + 1),
               "face_fwd": np.stack([pick(fm[a].forward, N * N + 1) for a in range(N)]),
               "face_inv": np.stack([pick(fm[a].inverse, N * N + 1) for a in range(N)]),
               "face_cons": np.stack([pick(fm[a].conserved, N * N) for a in range(N)])}
        jac = ("cell_fwd", "cell_inv", "face_fwd", "face_inv")
        for nm, m in (("domain", dom), ("halo", ~dom)):
            out[nm + "_jacobians_vs_truth"] = max(region(got[k], truth[k], True, m) for k in jac)
            out[nm + "_conserved_vs_truth"] = region(got["face_cons"], truth["face_cons"], False, m)
            out[nm + "_oracle64_conserved_vs_truth"] = region(o64["face_cons"], truth["face_cons"], False, m)
        out["conserved_vs_oracle64"] = rel(got["face_cons"], o64["face_cons"], False)
    else:
        J = cg.cell_metrics(grid).reshape(-1)[widx].double().cpu().numpy()
        act = [cg.face_metrics(grid)[a].reshape(-1, N)[widx].double().cpu().numpy() for a in range(N)]
        for nm, m in (("domain", dom), ("halo", ~dom)):
            if not m.any():
                out[nm + "_jacobians_vs_truth"] = out[nm + "_conserved_vs_truth"] = out[nm + "_oracle64_conserved_vs_truth"] = 0.0
                continue
            out[nm + "_jacobians_vs_truth"] = float(np.abs(J[m] / truth["cell_fwd"][m, N * N] - 1).max())
            ev = et = 0.0
            for a in range(N):
                t_ = np.stack([truth["face_cons"][a][m][:, a + N * c] for c in range(N)], axis=1)
                o_ = np.stack([o64["face_cons"][a][m][:, a + N * c] for c in range(N)], axis=1)
                sc = np.abs(truth["face_cons"][a][m]).max(axis=1, keepdims=True)
                ev, et = max(ev, float((np.abs(act[a][m] - t_) / sc).max())), max(et, float((np.abs(o_ - t_) / sc).max()))
            out[nm + "_conserved_vs_truth"], out[nm + "_oracle64_conserved_vs_truth"] = ev, et
    out = {k: ctx.allmax(v) for k, v in out.items()}
    # `max_rel_err`: CUDA against the exact value of the reference's algorithm on the same inputs (long-double truth)
    # on cell.domain, where BASELINE.md §4 defines parity
    out["max_rel_err"] = max(out["domain_jacobians_vs_truth"], out["domain_conserved_vs_truth"])
    out["cells"] = nsamples * ctx.world
    out["cells_in_domain"] = ctx.allsum_int(int(dom.sum()))
    out["truth"] = "oracle closure graph in long double" if dtype == "f64" else "FP64 oracle on the FP32-rounded nodes"
    return out


def measure_discrete(cg, ctx, args, spec, *, profile, dtype, steps, warmup, strong=False, do_e2e=True, do_cpu=True,
                     do_parity=True, workload_name=""):
    """One workload on the ranks of this run; returns the JSON line as a dict (rank 0 prints it or nests it)."""
    torch, lib = ctx.torch, cg._lib.lib()
    world, rank, dev = ctx.world, ctx.rank, ctx.dev
    n, h = spec["n"], spec["nhalo"]
    dim = 3 if spec["kind"] == "3d" else 2
    T = np.float64 if dtype == "f64" else np.float32
    nslab = spec.get("nslab", n)
    ncell_slab = nslab * world  # weak scaling: the grid grows along the slab axis
    if strong:
        ncell_slab = nslab      # strong scaling: the one-GPU grid split over the ranks
    plan = cg.slabs.make_plan(rank, world, ncell_slab, h)
    lo, hi = plan.required
    olo, ohi = plan.owned
    if dim == 3:
        gen = sph_slab_3d if spec.get("gen") == "sph" else wavy_slab_3d
        host = gen(n + 1, n + 1, ncell_slab + 1, lo, hi)
        inplane = (n, n)
    else:
        host = wavy_slab_2d(n + 1, ncell_slab + 1, lo, hi)
        inplane = (n,)
    host = [a if a.dtype == T else a.astype(T) for a in host]
    if world > 1:
        # the truth for the non-owned planes comes from their owners through the exchange: blank them on this rank
        for a in host:
            a[:olo - lo] = 0
            a[ohi - lo:] = 0
    perm = tuple(range(dim - 1, -1, -1))
    kw = cg.slabs.slab_grid_kwargs(plan, inplane, h) if world > 1 else {}
    # built from HOST arrays ([i,j,k]-indexed views of the [k,j,i] memory): the library owns the device node buffer
    grid = cg.DiscreteGrid(*[a.transpose(*perm) for a in host], h, T=T, conserved_metric_scheme=spec["scheme"],
                           cache_mode="lazy", profile=profile, **kw)
    host_pinned = [torch.from_numpy(a).pin_memory() for a in host] if do_e2e else None
    grid.set_kernel(args.kernel)
    comm = ctx.comm
    cells_local = int(np.prod(grid.window_cells))
    cells_total = ctx.allsum_int(cells_local)
    overlap = world > 1 and dim == 3 and args.kernel == 1
    pipelined = dim == 3 and args.kernel == 1
    st = grid._stream

    def exchange_then_eval():
        if world > 1:  # C ABI: grouped ncclSend / ncclRecv into the node buffer, then the stream waits
            cg._lib.check(lib.cg_halo_exchange_begin(grid._h, comm._h, rank, world, comm.bounds(plan), st()))
            cg._lib.check(lib.cg_halo_exchange_end(grid._h, st()))
        cg._lib.check(lib.cg_grid_nodes_changed(grid._h, st()))
        cg.refresh_metrics(grid)

    def step_resident():
        if overlap:
            # cg_grid_step_overlapped: exchange posted first; planes that depend on owned nodes only are padded +
            # evaluated while the neighbour planes travel over NVLink; the boundary planes follow the wait
            cg.slabs.step_overlapped_nccl(grid, plan, comm)
        else:
            exchange_then_eval()

    def gcl_all():
        return cg.slabs.gcl_dist(grid, comm) if world > 1 else cg.gcl(grid)

    def step_e2e():
        if pipelined:
            # cg_grid_update_from_host[_dist]: the owned node planes go to the device in pieces on a copy stream; the
            # planes the neighbours wait for first, then the NCCL exchange; each piece that lands releases the padded
            # planes and cell planes it completes
            grid.update_from_host(*host_pinned, nchunks=args.h2d_chunks, comm=comm, plan=plan, gcl=True)
        else:
            grid.set_nodes(*host_pinned)
            if world > 1:
                exchange_then_eval()
            else:
                cg.refresh_metrics(grid)
        return gcl_all()  # device reduction (+ seam plane and max over ranks) + D2H of the residuals (synchronises)

    def timed(fn, steps_):
        ctx.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps_):
            fn()
        e1.record()
        ctx.barrier()
        return ctx.allmax(e0.elapsed_time(e1))

    exchange_then_eval()  # the first evaluation fills the non-owned planes
    for _ in range(max(warmup, 3)):
        step_resident()
    ctx.barrier()
    sampler = ClockSampler(ctx.local_rank)
    sampler.start()
    launches0 = lib.cg_launch_count()
    total_ms = timed(step_resident, steps)
    launches = lib.cg_launch_count() - launches0
    clocks = sampler.stop()
    try:  # the metric kernels of the LAST timed step (the handle's own CUDA events): the in-loop, power-capped figure
        k_ms_in_loop = float(grid.last_eval_ms()) if not overlap else None
    except Exception:
        k_ms_in_loop = None
    ms_per_step = total_ms / steps
    value = cells_total / (ms_per_step * 1e-3) / 1e6

    # dominant kernel(s): one cg_grid_eval over the whole slab, timed with CUDA events on its own stream.  The step runs at
    # the GPU's power cap (FP64-heavy kernels: SM clock ~1750 of 1965 MHz), so WHEN the kernels are timed matters:
    #   kernel_ms                  5 evaluations timed alone, back to back, after a 0.3 s pause — the condition of every
    #                              earlier record of this repo (the nvidia-smi sampler's shutdown used to provide the pause)
    #   kernel_ms_burst            3 evaluations, each after its own 0.3 s pause: a kernel timed alone from a rested GPU,
    #                              the condition of the measured peak (a best-of-10 burst copy) and of an ncu capture
    #   kernel_ms_last_timed_step  the same launches inside the last step of the timed loop (sustained, power-capped)
    time.sleep(0.3)
    kms = []
    for _ in range(min(steps, 5)):
        exchange_then_eval()
        kms.append(grid.last_eval_ms())
    k_ms = float(np.mean(kms))
    kb = []
    for _ in range(3):
        time.sleep(0.3)
        exchange_then_eval()
        kb.append(grid.last_eval_ms())
    k_ms_burst = float(np.median(kb))
    bpc = BYTES_PER_CELL[(spec["kind"], profile)] * (1.0 if dtype == "f64" else 0.5)
    peak, peak_src = measured_peak_gbs()
    achieved = bpc * cells_local / (k_ms * 1e-3) / 1e9
    traffic, traffic_src = None, None
    try:  # DRAM bytes of the metric kernels: a CITATION of the committed `ncu --set full` capture, not measured in this run
        tp = os.path.join(ROOT, "profiles", "ncu_traffic_3d512.json")
        with open(tp) as f:
            tj = json.load(f)
        if (workload_name, profile, dtype, spec["scheme"], args.kernel, world) == ("3d512", "full", "f64", "curvature", 1, 1):
            traffic = tj["total"]
            traffic_src = "ncu --set full capture committed as profiles/ncu_traffic_3d512.json (%s)" % tj.get("capture", "round 1")
    except Exception:
        pass
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_source": traffic_src, "algorithmic_bytes": bpc * cells_local,
                "kernel": (cg.kernel_names_3d() if dim == 3 and args.kernel == 1 else ("metrics3d" if dim == 3 else "k_metrics2d_staged")),
                "kernel_ms": k_ms, "kernel_ms_burst": k_ms_burst, "frac_burst": bpc * cells_local / (k_ms_burst * 1e-3) / 1e9 / peak,
                "kernel_ms_last_timed_step": k_ms_in_loop,
                "frac_last_timed_step": (bpc * cells_local / (k_ms_in_loop * 1e-3) / 1e9 / peak) if k_ms_in_loop else None,
                "kernel_ms_note": "kernel_ms: mean of 5 evaluations timed alone, back to back, 0.3 s after the timed loop (as in every "
                                  "earlier record); kernel_ms_burst: each evaluation after its own 0.3 s pause (rested GPU, like the "
                                  "burst copy the peak was measured with); kernel_ms_last_timed_step: inside the last step of the timed "
                                  "loop (GPU at its power cap, SM clock in `clocks`)",
                "algorithmic_bytes_per_cell": bpc, "peak_source": f"MEASURED_PEAKS.json ({peak_src}, copy burst)",
                # also against the nominal ~8 TB/s of the north star
                "frac_of_nominal_8TBs": achieved / 8000.0}

    e2e = None
    if do_e2e:
        for _ in range(2):
            step_e2e()
        ms = timed(step_e2e, steps) / steps
        h2d = sum(int(t[olo - lo:ohi - lo].numel() * t.element_size()) for t in host_pinned)  # owned planes only
        e2e = {"value": cells_total / (ms * 1e-3) / 1e6, "unit": "Mcells/s", "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": 8 * dim, "ms_per_step": ms, "h2d_GBs_per_rank_if_copy_bound": h2d / (ms * 1e-3) / 1e9,
               "note": ("pinned host nodes -> H2D of the owned planes in %d pieces (boundary planes first, then the NCCL "
                        "exchange), each piece overlapped with pad + metrics of the planes it completes "
                        "(cg_grid_update_from_host%s)" % (args.h2d_chunks, "_dist" if world > 1 else "") if pipelined else
                        "pinned host nodes -> H2D -> pad -> metrics") + " -> device GCL max (seam planes and max over "
                       "ranks inside the C ABI) -> D2H of 3 doubles; the metric arrays stay on the device like a "
                       "KernelAbstractions GPU backend's storage"}

    parity = None
    if do_parity:
        step_resident()
        parity = parity_block(cg, ctx, grid, host, plan, spec, profile, dtype)
        parity["gcl_max"] = max(gcl_all())

    cpu = None
    if rank == 0 and world == 1 and do_cpu:
        nodes_cpu = [a.transpose(*reversed(range(dim))) for a in host]  # [i,j,k] views of the full node arrays
        v, cores, desc, _ = cpu_oracle_rate(spec, nodes_cpu, args.cpu_seconds)
        cpu = {"value": v, "unit": "Mcells/s", "cores": cores, "kind": "port", "sample": desc}

    line = {
        "metric": "metric-eval Mcells/s (cell+face, FP64)" if dtype == "f64" else "metric-eval Mcells/s (cell+face, FP32)",
        "value": value, "unit": "Mcells/s", "n_gpus": world, "steps": steps, "warmup": max(warmup, 3),
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None,
        "dtype": dtype, "data": "synthetic",
        "config": {"workload": spec["desc"] + (f"; k-slabs over {world} GPUs, grid {n}x{n}x{ncell_slab}" if world > 1 and dim == 3 else ""),
                   "scheme": spec["scheme"], "profile": profile, "cells_written": cells_total,
                   "kernel": "tiled" if args.kernel == 1 else "gather",
                   "l2": "inputs+outputs per step far exceed the 126 MB L2 (no flush needed)",
                   "parallelism": f"slab{world}" if world > 1 else "single",
                   "halo_exchange": ("NCCL node planes inside the C ABI (cg_grid_step_overlapped), overlapped with interior planes"
                                     if overlap else ("NCCL node planes (cg_halo_exchange_begin/end)" if world > 1 else "none"))},
        "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "parity": parity, "gpu_launches": int(launches),
        "clocks": clocks,
    }
    grid.close()
    del grid, host_pinned, host
    torch.cuda.empty_cache()
    return line


def measure_mapped(cg, ctx, args, spec, steps, warmup, do_parity=True):
    """BASELINE config 4: MappedGrid, update!(grid, t_k) — which refreshes the node / centroid / face coordinates like
    mapped_grid.jl:551-587 — + refresh of both metric caches every step (1 GPU; every rank runs its own replica)."""
    torch, lib = ctx.torch, cg._lib.lib()
    n, h = spec["n"], spec["nhalo"]
    terms = cg.workloads.wavy_3d_terms((n, n, n), time_amp=0.1)
    grid = cg.MappedGrid(terms, {}, (n, n, n), h, cache_mode="lazy", conserved_metric_scheme=spec["scheme"])
    _ = grid.node_coordinates   # coordinate storage exists from here on: every update! rewrites the 15 coordinate arrays
    cells = int(np.prod(grid.window_cells))
    tk = [0]

    def step():
        tk[0] += 1
        cg.update(grid, 0.01 * tk[0])
        cg.refresh_metrics(grid)

    for _ in range(max(warmup, 3)):
        step()
    torch.cuda.synchronize()
    sampler = ClockSampler(ctx.local_rank)
    sampler.start()
    launches0 = lib.cg_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    launches = lib.cg_launch_count() - launches0
    ms_per_step = e0.elapsed_time(e1) / steps
    clocks = sampler.stop()
    kms = []
    for _ in range(min(steps, 5)):
        step()
        kms.append(grid.last_eval_ms())
    k_ms = float(np.mean(kms))
    peak, peak_src = measured_peak_gbs()
    bpc = float(BYTES_PER_CELL[("m3d", "full")])  # 976: coordinates (24 + 24 + 72) + FULL metrics (856), no node read
    achieved = bpc * cells / (ms_per_step * 1e-3) / 1e9
    parity = None
    if do_parity:
        from oracle import oracle as O
        O.set_num_threads(_host_threads())
        t = 0.01 * tk[0]
        rng = np.random.default_rng(11)
        sample = np.sort(rng.choice(cells, size=32, replace=False))
        idx = torch.from_numpy(sample).to(ctx.dev)
        cm, fm = cg.cell_metrics(grid), cg.face_metrics(grid)
        pick = lambda a, k: a.reshape(-1, k)[idx].cpu().numpy()
        got = {"cell_fwd": pick(cm.forward, 10), "cell_inv": pick(cm.inverse, 10),
               "face_fwd": np.stack([pick(fm[a].forward, 10) for a in range(3)]),
               "face_inv": np.stack([pick(fm[a].inverse, 10) for a in range(3)]),
               "face_cons": np.stack([pick(fm[a].conserved, 9) for a in range(3)])}
        truth = O.mapped_eval(terms, t, (n, n, n), h, spec["scheme"], sample=sample, precision="ld")
        o64 = O.mapped_eval(terms, t, (n, n, n), h, spec["scheme"], sample=sample)

        def rel(a, b, has_j):
            K = a.shape[-1] - (1 if has_j else 0)
            sc = np.abs(b[..., :K]).max(axis=-1, keepdims=True)
            e = float((np.abs(a[..., :K] - b[..., :K]) / sc).max())
            return max(e, float((np.abs(a[..., K] - b[..., K]) / np.abs(b[..., K])).max())) if has_j else e
        jac = ("cell_fwd", "cell_inv", "face_fwd", "face_inv")
        parity = {"jacobians_vs_truth": max(rel(got[k], truth[k], True) for k in jac),
                  "conserved_vs_truth": rel(got["face_cons"], truth["face_cons"], False),
                  "conserved_vs_oracle64": rel(got["face_cons"], o64["face_cons"], False),
                  "oracle64_conserved_vs_truth": rel(o64["face_cons"], truth["face_cons"], False),
                  "cells": len(sample), "gcl_max": max(cg.gcl(grid)), "truth": "oracle closure graph in long double"}
        parity["max_rel_err"] = max(parity["jacobians_vs_truth"], parity["conserved_vs_truth"])
    line = {
        "metric": "metric-eval Mcells/s (cell+face, FP64)", "value": cells / (ms_per_step * 1e-3) / 1e6, "unit": "Mcells/s",
        "n_gpus": 1, "steps": steps, "warmup": max(warmup, 3), "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": spec["desc"], "scheme": spec["scheme"], "profile": "full", "cells_written": cells,
                   "l2": "outputs per step far exceed the 126 MB L2 (no flush needed)", "parallelism": "single",
                   "coordinates": "node / centroid / face coordinates are rewritten by every update! (inside the timed step)"},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": None, "kernel": "k_mapped_coords + k_mapped_tables/factors + k_mapped_metrics3d_staged (the whole update! + refresh step)",
                     "kernel_ms": ms_per_step, "metric_kernel_ms": k_ms,
                     "algorithmic_bytes_per_cell": bpc, "peak_source": f"MEASURED_PEAKS.json ({peak_src}, copy burst)"},
        "cpu_baseline": None, "e2e": None, "parity": parity, "gpu_launches": int(launches), "clocks": clocks,
    }
    grid.close()
    del grid
    torch.cuda.empty_cache()
    return line


def brief(line):
    """What an `other_configs` entry keeps of a full line."""
    r = line["roofline"]
    out = {"workload": line["config"]["workload"], "scheme": line["config"]["scheme"], "profile": line["config"]["profile"],
           "dtype": line["dtype"], "n_gpus": line["n_gpus"], "scaling": line["scaling"], "steps": line["steps"],
           "value": line["value"], "unit": line["unit"], "ms_per_step": line["ms_per_step"], "kernel_ms": r["kernel_ms"],
           "frac": r["frac"], "achieved_GBs": r["achieved"], "algorithmic_bytes_per_cell": r["algorithmic_bytes_per_cell"],
           "parity": line.get("parity"), "clocks": line["clocks"], "gpu_launches": line["gpu_launches"]}
    if line.get("e2e"):
        out["e2e"] = {"value": line["e2e"]["value"], "ms_per_step": line["e2e"]["ms_per_step"]}
    return out


def measure_legacy(cg, n=256, reps=5):
    """Legacy MEG6 pipeline (CurvilinearGrid3D(x, y, z, :meg6) update!) on an n^3-cell wavy mesh: the
    fused launches, CUDA events around each update (validity reduction included), median."""
    import torch
    x, y, z = cg.workloads.wavy_nodes_3d(n + 1, n + 1, n + 1)
    mesh = cg.legacy.CurvilinearGrid3D(x, y, z, "meg6")
    for _ in range(3):
        cg.legacy.update(mesh, True, False)
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        cg.legacy.update(mesh, True, False)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[len(ts) // 2]
    out = {"workload": f"legacy CurvilinearGrid3D {n}^3 wavy, :meg6, update! (centroids + 28 cell + 57 edge arrays)",
           "dtype": "f64", "steps": reps, "value": n ** 3 / ms / 1e3, "unit": "Mcells/s", "ms_per_step": ms,
           "gpu_launches": mesh.launches + 1, "algorithmic_bytes_per_cell": 848.0,
           "frac": 848.0 * n ** 3 / (ms * 1e-3) / 1e9 / measured_peak_gbs()[0],
           "bound": "FP64 issue (Kahan sums / isapprox thresholds of the reference, order-sensitive): DESIGN.md §10"}
    del mesh
    torch.cuda.empty_cache()
    return out


def measure_generic(cg, n=128, reps=3):
    """MappedGrid with a NON-separable mapping through the NVRTC-compiled generic kernels: update! +
    cell+face refresh, device time of the evaluation."""
    src = """
template <class S> __device__ void cg_map(S xi, S eta, S zeta, double t, const double* p, S& x, S& y, S& z) {
  x = xi + p[0] * sin(p[1] * (eta + 0.5 * zeta) + 0.3 * t);
  y = eta * (1.0 + p[2] * xi) / (1.0 + p[3] * zeta * zeta);
  z = zeta + p[0] * exp(-p[4] * (xi - eta) * (xi - eta)) * cos(t) + sqrt(1.0 + 0.01 * xi * xi);
}"""
    g = cg.MappedGrid(src, [0.3, 0.4, 0.001, 1e-5, 0.02], (n, n, n), 5, cache_mode="lazy")
    ts = []
    for i in range(reps + 1):
        cg.update(g, 0.1 * i)
        cg.refresh_metrics(g)
        ts.append(g.last_eval_ms())
    ms = sorted(ts[1:])[len(ts[1:]) // 2]
    gcl = max(cg.gcl(g))
    out = {"workload": f"3-D MappedGrid {n}^3, non-separable mapping as CUDA source (NVRTC), update! + cell+face",
           "scheme": "curvature", "dtype": "f64", "steps": reps, "value": (n + 10) ** 3 / ms / 1e3, "unit": "Mcells/s",
           "ms_per_step": ms, "gpu_launches": 2, "gcl_max": gcl,
           "bound": "FP64 issue: 108 mapping evaluations on 18-component jets per cell (fallback path)"}
    g.close()
    return out


def run_ours(args, spec):
    import curvigrid_b200 as cg
    ctx = Ctx(args)
    world = ctx.world
    steps, warmup = args.steps, args.warmup
    if spec["kind"] == "m3d":
        line = measure_mapped(cg, ctx, args, spec, steps, warmup)
    else:
        line = measure_discrete(cg, ctx, args, spec, profile=args.profile, dtype=args.dtype, steps=steps, warmup=warmup,
                                strong=args.strong, do_e2e=not args.no_e2e, do_cpu=not args.no_cpu_baseline,
                                workload_name=args.workload)
        if line["roofline"].get("kernel_ms") and args.workload == "3d512" and args.profile == "full" and args.dtype == "f64" \
                and spec["scheme"] == "curvature" and args.kernel == 1:
            fp = cg.fp64_instr_per_cell()
            if fp:
                cells_local = line["config"]["cells_written"] / world
                rate = fp * cells_local / (line["roofline"]["kernel_ms"] * 1e-3)
                line["roofline"]["fp64"] = {"instr_per_cell": fp, "achieved_Tinstr_s": rate / 1e12,
                                            "peak_Tinstr_s": 59.0 * 148 * 1.965e9 / 1e12,
                                            "frac": rate / (59.0 * 148 * 1.965e9), "note": "peak at the maximum SM clock"}
    # the other BASELINE configs in the same driver-visible line (short runs)
    if not args.no_others and args.workload == "3d512" and args.profile == "full" and args.dtype == "f64" and not args.strong:
        others = {}
        k = max(3, min(5, steps))

        def other(name, sp, **kw):
            try:
                others[name] = brief(measure_discrete(cg, ctx, args, sp, steps=k, warmup=3, do_cpu=False, workload_name=name, **kw))
            except Exception as e:  # a failed extra never takes the headline line with it
                others[name] = {"error": repr(e)[:300]}
        if world == 1:
            other("2d4096 (BASELINE config 2)", workload_spec("2d4096", None), profile="full", dtype="f64", do_e2e=False)
            try:
                others["m3d512 (BASELINE config 4)"] = brief(measure_mapped(cg, ctx, args, workload_spec("m3d512", None), k, 3))
            except Exception as e:
                others["m3d512 (BASELINE config 4)"] = {"error": repr(e)[:300]}
            other("3d512 compact", workload_spec("3d512", None), profile="compact", dtype="f64", do_e2e=False)
            other("3d512 f32", workload_spec("3d512", None), profile="full", dtype="f32", do_e2e=False)
            for name, fn in (("legacy meg6 256^3 (SURVEY rows a17-a21)", measure_legacy),
                             ("generic mapping 128^3 (NVRTC)", measure_generic)):
                try:
                    others[name] = fn(cg)
                except Exception as e:
                    others[name] = {"error": repr(e)[:300]}
        else:
            other("sph1536 compact weak (BASELINE config 5)", workload_spec("sph1536", None), profile="compact", dtype="f64",
                  do_e2e=False)
            other("3d512 strong", workload_spec("3d512", None), profile="full", dtype="f64", strong=True, do_e2e=False)
        line["other_configs"] = others
    if ctx.rank == 0:
        print(json.dumps(line), flush=True)
    ctx.close()


def main():
    args = parse_args()
    spec = workload_spec(args.workload, args.scheme)
    if args.impl == "reference":
        run_reference(args, spec)
    else:
        run_ours(args, spec)


if __name__ == "__main__":
    main()

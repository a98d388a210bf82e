"""k-slab decomposition of a DiscreteGrid over the GPUs of one box.

The reference has no communication layer: `local_grid`
(mapped_grid.jl:344-379) gives rank-local analytic grids, and multiblock
`interface_index_plan` (src/multiblock/transfer.jl:157-236) only lists index
pairs for "a solver or MPI layer".  For sampled meshes a rank needs *real*
neighbour node planes: 1 plane below and 3 above its cell
slab.  This module is that exchange: pure index algebra + torch.distributed
point-to-point (NCCL over NVLink on GPUs; gloo in the CPU tests).

Conventions: the slab axis is the slowest axis (k for 3-D, j for 2-D) so every
plane is one contiguous range of memory.  Cell planes are indices of the
GLOBAL cell.full array (0 .. nk+2h-1); node planes are indices of the INTERIOR
node array (0 .. nk).
"""
from __future__ import annotations

from dataclasses import dataclass


def partition_planes(nplanes, nranks):
    """Split `nplanes` cell planes into `nranks` contiguous [k0,k1) ranges, sizes differing by <= 1."""
    base, rem = divmod(nplanes, nranks)
    out, k = [], 0
    for r in range(nranks):
        n = base + (1 if r < rem else 0)
        out.append((k, k + n))
        k += n
    return out


def required_node_range(k0, k1, nhalo, ncell):
    """Interior node planes [lo,hi) a slab of cell planes [k0,k1) reads (same rule as
    cg_grid_required_node_range): padded planes k0-1 .. k1+2, clamped to the interior,
    plus the second boundary plane where Line() extrapolation needs a slope."""
    lm, hm = k0 - 1 - nhalo, k1 + 2 - nhalo
    lo = min(max(lm, 0), ncell)
    hi = min(max(hm, 0), ncell)
    if lm < 0 and hi < 1:
        hi = 1
    if hm > ncell and lo > ncell - 1:
        lo = ncell - 1
    return lo, hi + 1


def owned_node_range(k0, k1, nhalo, ncell, is_first, is_last):
    """Interior node planes a rank is the *source* of: plane m belongs to the slab holding
    cell plane m+nhalo; the first/last ranks also own what lies below/above."""
    lo = 0 if is_first else min(max(k0 - nhalo, 0), ncell + 1)
    hi = ncell + 1 if is_last else min(max(k1 - nhalo, 0), ncell + 1)
    return lo, max(hi, lo)


@dataclass
class SlabPlan:
    rank: int
    nranks: int
    nhalo: int
    ncell_slab: int       # interior cells along the slab axis (global)
    cells: tuple          # (k0, k1) cell planes of this rank
    required: tuple       # (lo, hi) interior node planes held locally
    owned: tuple          # (lo, hi) interior node planes this rank sources
    sends: list           # [(peer, lo, hi)] node planes (global indices) to send
    recvs: list           # [(peer, lo, hi)] node planes to receive


def make_plan(rank, nranks, ncell_slab, nhalo):
    parts = partition_planes(ncell_slab + 2 * nhalo, nranks)
    req = [required_node_range(k0, k1, nhalo, ncell_slab) for k0, k1 in parts]
    own = [owned_node_range(k0, k1, nhalo, ncell_slab, r == 0, r == nranks - 1) for r, (k0, k1) in enumerate(parts)]
    sends, recvs = [], []
    for peer in range(nranks):
        if peer == rank:
            continue
        lo, hi = max(own[rank][0], req[peer][0]), min(own[rank][1], req[peer][1])
        if hi > lo:
            sends.append((peer, lo, hi))
        lo, hi = max(own[peer][0], req[rank][0]), min(own[peer][1], req[rank][1])
        if hi > lo:
            recvs.append((peer, lo, hi))
    return SlabPlan(rank, nranks, nhalo, ncell_slab, parts[rank], req[rank], own[rank], sends, recvs)


def post_exchange(plan, buffers, dist=None, group=None):
    """Post the sends / receives of the node-plane exchange and return the pending requests WITHOUT waiting
    (NCCL: the transfers run on NCCL's stream while the caller keeps launching interior work)."""
    if plan.nranks == 1:
        return []
    if dist is None:
        import torch.distributed as dist
    ops = []
    base = plan.required[0]
    for buf in buffers:
        for peer, lo, hi in plan.sends:
            ops.append(dist.P2POp(dist.isend, buf[lo - base:hi - base], peer, group=group))
        for peer, lo, hi in plan.recvs:
            ops.append(dist.P2POp(dist.irecv, buf[lo - base:hi - base], peer, group=group))
    if not ops:
        return []
    return dist.batch_isend_irecv(ops)


def wait_exchange(reqs):
    """NCCL: makes the current stream wait for the transfers (the host does not block); gloo: blocks."""
    for r in reqs:
        r.wait()


def exchange_node_planes(plan, buffers, dist=None, group=None):
    """Fill the non-owned planes of each local node buffer from their owners.

    buffers: list of torch tensors whose FIRST dimension is the slab axis and covers
    plan.required (memory order, e.g. [nk_local, nj+1, ni+1]).  Returns the (completed) requests."""
    reqs = post_exchange(plan, buffers, dist, group)
    wait_exchange(reqs)
    return reqs


def overlap_schedule(plan):
    """Which padded node planes and cell planes of a rank's slab can be done BEFORE the exchange has landed.

    Padded plane e of the slab holds the global interior node plane m = k0 + e - 1 - nhalo, Line()-extrapolated
    outside [0, ncell] from the planes c = clamp(m) and c -+ 1; the cell plane k of the slab reads
    the padded planes k .. k+4 (cg_grid_eval_planes).  Returns a dict:
      pad_early / pad_late   lists of padded plane ranges [e_lo, e_hi)
      eval_early / eval_late lists of cell plane ranges   [k_lo, k_hi)   (window-local)
    pad_early + eval_early depend on owned node planes only."""
    k0, k1 = plan.cells
    nw, h, n = k1 - k0, plan.nhalo, plan.ncell_slab
    ne = nw + 4
    recv = [(lo, hi) for _, lo, hi in plan.recvs]

    def received(m):
        return any(lo <= m < hi for lo, hi in recv)

    dirty = []
    for e in range(ne):
        m = k0 + e - 1 - h
        c = min(max(m, 0), n)
        deps = [c] if m == c else [c, c + 1 if m < c else c - 1]
        dirty.append(any(received(d) for d in deps))
    a = 0
    while a < ne and dirty[a]:
        a += 1
    b = ne
    while b > a and dirty[b - 1]:
        b -= 1
    if any(dirty[a:b]):  # not a prefix + suffix (cannot happen with contiguous ownership): no early work
        a, b = ne, ne
    ke_lo = min(a, nw)
    ke_hi = max(ke_lo, min(b - 4, nw))
    return {"pad_early": [(a, b)] if b > a else [], "pad_late": [r for r in ((0, a), (b, ne)) if r[1] > r[0]],
            "eval_early": [(ke_lo, ke_hi)] if ke_hi > ke_lo else [],
            "eval_late": [r for r in ((0, ke_lo), (ke_hi, nw)) if r[1] > r[0]]}


def step_overlapped(grid, plan, buffers, what=3, dist=None, group=None):
    """One evaluation of a rank's slab with the node-plane exchange overlapped with interior work
    (3-D DiscreteGrid, marching kernels): post the exchange, pad + evaluate everything that depends on owned
    node planes only, wait for the exchange, pad + evaluate the 1 (low) / 3 (high) boundary planes."""
    import ctypes

    from . import _lib as L
    lib, st = L.lib(), grid._stream()
    i64 = ctypes.c_int64
    plan_ = overlap_schedule(plan)
    grid._bind_metric_outputs(what)
    reqs = post_exchange(plan, buffers, dist, group)
    for lo, hi in plan_["pad_early"]:
        L.check(lib.cg_grid_pad_planes(grid._h, i64(lo), i64(hi), st))
    for lo, hi in plan_["eval_early"]:
        L.check(lib.cg_grid_eval_planes(grid._h, what, i64(lo), i64(hi), 0, st))
    wait_exchange(reqs)
    for lo, hi in plan_["pad_late"]:
        L.check(lib.cg_grid_pad_planes(grid._h, i64(lo), i64(hi), st))
    for lo, hi in plan_["eval_late"]:
        L.check(lib.cg_grid_eval_planes(grid._h, what, i64(lo), i64(hi), 0, st))
    # every plane is covered: an empty final range marks the caches valid
    L.check(lib.cg_grid_eval_planes(grid._h, what, i64(0), i64(0), 1, st))
    return plan_


class NcclComm:
    """An ncclComm_t owned by libcurvigrid_cuda.so (cg_nccl_comm_create): what a Julia host would take from NCCL.jl.
    The 128-byte unique id is made on rank 0 (cg_nccl_unique_id) and broadcast with torch.distributed (any backend);
    from then on every exchange is a C-ABI call (cg_grid_step_overlapped, cg_halo_exchange_*, cg_grid_gcl_max_dist)."""

    def __init__(self, rank, nranks, device, dist=None, group=None):
        import ctypes

        import torch

        from . import _lib as L
        if dist is None:
            import torch.distributed as dist
        self.rank, self.nranks = int(rank), int(nranks)
        idbuf = (ctypes.c_ubyte * 128)()
        if rank == 0:
            L.check(L.lib().cg_nccl_unique_id(idbuf))
        if nranks > 1:
            obj = [bytes(idbuf)]
            dist.broadcast_object_list(obj, src=0, group=group)
            idbuf = (ctypes.c_ubyte * 128).from_buffer_copy(obj[0])
        self._h = ctypes.c_void_p()
        torch.cuda.set_device(device)
        L.check(L.lib().cg_nccl_comm_create(idbuf, self.nranks, self.rank, int(device), ctypes.byref(self._h)))

    def close(self):
        from . import _lib as L
        if getattr(self, "_h", None) and self._h.value:
            L.lib().cg_nccl_comm_destroy(self._h)
            self._h = None

    def bounds(self, plan):
        """slab_bounds argument of the C ABI for the partition make_plan uses."""
        import ctypes
        parts = partition_planes(plan.ncell_slab + 2 * plan.nhalo, plan.nranks)
        return (ctypes.c_int64 * (plan.nranks + 1))(*([p[0] for p in parts] + [parts[-1][1]]))


def step_overlapped_nccl(grid, plan, comm, what=3):
    """step_overlapped through the C ABI only (cg_grid_step_overlapped): NCCL send / recv of the node planes on the
    handle's communication stream, interior planes meanwhile, boundary planes after the wait."""
    from . import _lib as L
    grid._bind_metric_outputs(what)
    L.check(L.lib().cg_grid_step_overlapped(grid._h, comm._h, comm.rank, comm.nranks, comm.bounds(plan), what,
                                            grid._stream()))


def gcl_dist(grid, comm):
    """GCL max over all slabs, seam faces included (cg_grid_gcl_max_dist)."""
    import ctypes

    from . import _lib as L
    out = (ctypes.c_double * 3)()
    L.check(L.lib().cg_grid_gcl_max_dist(grid._h, comm._h, comm.rank, comm.nranks, out, grid._stream()))
    return tuple(out[c] for c in range(grid.dim))


def slab_grid_kwargs(plan, inplane_celldims, nhalo):
    """kwargs for DiscreteGrid(...) describing this rank's slab of a 2-D/3-D grid.
    `inplane_celldims`: interior cells of the non-slab axes, fastest first."""
    dim = len(inplane_celldims) + 1
    celldims = tuple(inplane_celldims) + (plan.ncell_slab,)
    nfull = tuple(c + 2 * nhalo for c in celldims)
    k0, k1 = plan.cells
    origin = (0,) * (dim - 1) + (k0,)
    cells = nfull[:-1] + (k1 - k0,)
    node_origin = (0,) * (dim - 1) + (plan.required[0],)
    return dict(window=(origin, cells), global_celldims=celldims, node_origin=node_origin)

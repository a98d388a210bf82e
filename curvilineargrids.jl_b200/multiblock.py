"""Multiblock interface plans and the scalar ghost-cell exchange on the device.

Host side: `interface_index_plan` restates /root/reference/src/multiblock/transfer.jl:157-195 (driver) and :302-359
(`_paired_interface_layer_indices`) with the helpers of validation.jl:223-283 (`_face_axis_index`,
`_build_face_cartesian_index`, `_map_tangential_indices`) and the checks of transfer.jl:259-293.  Pure index algebra;
indices are 1-based like the reference's `CartesianIndex` (so its known answers, test_multiblock.jl:261-273, can be
asserted literally).

Device side: `exchange_interface` fills receiver ghost cells from donor interior cells of cell-centred fields that live
on the GPU (transfer.jl:25-117) with one launch per direction: scalar fields by `cg_copy_index_pairs`, vector / tensor
fields (AoS like Julia's Array{SVector} / Array{SMatrix}) by `cg_transfer_index_pairs` with the per-pair
`basis_transfer_matrix` (transfer.jl:461-488) that `DevicePlan.build_transforms` tabulates on the device from the two
blocks' centroid arrays (`cg_basis_transfer_matrices`; multiblock/cache.jl:55-108).  The two blocks may be on the same
GPU or on different ranks: `exchange_interface_remote` packs the donor values, moves them with NCCL inside the C ABI
(`cg_nccl_sendrecv`) and applies the transform while unpacking.
"""
from __future__ import annotations

import ctypes
from collections import namedtuple

import numpy as np

from . import _lib as L

BlockFace = namedtuple("BlockFace", "block_id axis side")              # types.jl:11-25 (axis 1-based, side "min"/"max")
BlockInterface = namedtuple("BlockInterface", "left right permutation flips")   # types.jl (tangential map left -> right)
IndexPlan = namedtuple("IndexPlan", "left_to_right right_to_left")


def _check_face(face, N):
    if not 1 <= face.axis <= N:
        raise ValueError(f"Invalid `axis={face.axis}` for `N={N}`.")
    if face.side not in ("min", "max"):
        raise ValueError(f"Invalid `side={face.side}`. Expected `:min` or `:max`.")


def _tangential_ranges(domain, face):
    # transfer.jl:295-300
    return tuple(r for d, r in enumerate(domain, start=1) if d != face.axis)


def _face_axis_index(domain, face, ghost_layer):
    # validation.jl:223-231: ghost_layer <= 0 walks into the interior, >= 1 into the halo
    lo, hi = domain[face.axis - 1]
    return lo - ghost_layer if face.side == "min" else hi + ghost_layer


def _build_index(tangential, axis, axis_index, N):
    # validation.jl:233-251
    full, t = [], iter(tangential)
    for d in range(1, N + 1):
        full.append(axis_index if d == axis else next(t))
    return tuple(full)


def _map_tangential(left_tvals, left_ranges, right_ranges, permutation, flips):
    # validation.jl:253-278
    right = [0] * len(left_tvals)
    for i, tv in enumerate(left_tvals):
        rd = permutation[i] - 1
        offset = tv - left_ranges[i][0]
        if flips[i]:
            offset = (right_ranges[rd][1] - right_ranges[rd][0]) - offset
        right[rd] = right_ranges[rd][0] + offset
    return tuple(right)


def _in_domain(idx, domain):
    return domain is None or all(lo <= i <= hi for i, (lo, hi) in zip(idx, domain))


def interface_index_plan(left_domain, left_face, right_domain, right_face, permutation, flips, depth=1,
                         left_valid_domain=None, right_valid_domain=None):
    """Ordered `source => target` index pairs of an oriented block interface (transfer.jl:157-195).

    Domains are tuples of inclusive 1-based `(first, last)` ranges per axis (`CartesianIndices`).  Returns
    IndexPlan(left_to_right, right_to_left) of `(source_index, target_index)` tuples in the reference's order:
    halo layer by halo layer, tangential indices column-major over the left face."""
    N = len(left_domain)
    if depth < 1:
        raise ValueError("`depth` must be positive.")
    _check_face(left_face, N)
    _check_face(right_face, N)
    perm, flip = tuple(permutation), tuple(flips)
    if len(perm) != N - 1:
        raise ValueError(f"Invalid `permutation` length {len(perm)}; expected {N - 1}.")
    if len(flip) != N - 1:
        raise ValueError(f"Invalid `flips` length {len(flip)}; expected {N - 1}.")
    if sorted(perm) != list(range(1, N)):
        raise ValueError(f"Invalid tangential `permutation`. Expected a permutation of `1:{N - 1}`.")
    if not all(isinstance(f, (bool, np.bool_)) for f in flip):
        raise ValueError("`flips` entries must be booleans.")
    lt, rt = _tangential_ranges(left_domain, left_face), _tangential_ranges(right_domain, right_face)
    for i in range(N - 1):
        rd = perm[i] - 1
        if lt[i][1] - lt[i][0] != rt[rd][1] - rt[rd][0]:
            raise ValueError(f"Interface tangential range mismatch: left tangential axis {i + 1} has length "
                             f"{lt[i][1] - lt[i][0] + 1} but right tangential axis {rd + 1} has length "
                             f"{rt[rd][1] - rt[rd][0] + 1}")
    l2r, r2l = [], []
    for layer in range(1, depth + 1):
        src = -(layer - 1)
        la, ra = _face_axis_index(left_domain, left_face, src), _face_axis_index(right_domain, right_face, src)
        lg, rg = _face_axis_index(left_domain, left_face, layer), _face_axis_index(right_domain, right_face, layer)
        if N == 1:
            tvals = [()]
        else:  # CartesianIndices(left_tangential): first tangential axis fastest
            grids = np.meshgrid(*[np.arange(lo, hi + 1) for lo, hi in lt], indexing="ij")
            tvals = list(zip(*[g.ravel(order="F") for g in grids]))
        for ltv in tvals:
            ltv = tuple(int(v) for v in ltv)
            rtv = _map_tangential(ltv, lt, rt, perm, flip) if N > 1 else ()
            li, ri = _build_index(ltv, left_face.axis, la, N), _build_index(rtv, right_face.axis, ra, N)
            lgi, rgi = _build_index(ltv, left_face.axis, lg, N), _build_index(rtv, right_face.axis, rg, N)
            if _in_domain(li, left_valid_domain) and _in_domain(rgi, right_valid_domain):
                l2r.append((li, rgi))
            if _in_domain(ri, right_valid_domain) and _in_domain(lgi, left_valid_domain):
                r2l.append((ri, lgi))
    return IndexPlan(l2r, r2l)


def linear_indices(pairs, src_shape, dst_shape):
    """0-based column-major linear indices (numpy int64) of the pairs' sources / targets in arrays of the given
    `cell.full` shapes (Julia layout: first axis fastest)."""
    def lin(idx, shape):
        out, stride = 0, 1
        for i, n in zip(idx, shape):
            if not 1 <= i <= n:
                raise ValueError(f"index {idx} outside an array of size {shape}")
            out += (i - 1) * stride
            stride *= n
        return out
    s = np.array([lin(a, src_shape) for a, _ in pairs], dtype=np.int64)
    d = np.array([lin(b, dst_shape) for _, b in pairs], dtype=np.int64)
    return s, d


def _copy_pairs(dst, src, dst_idx, src_idx):
    import torch
    if not (dst.is_cuda and src.is_cuda and dst.is_contiguous() and src.is_contiguous() and dst.dtype == src.dtype):
        raise ValueError("exchange needs contiguous CUDA tensors of one dtype")
    esize = dst.element_size()
    st = ctypes.c_void_p(torch.cuda.current_stream(dst.device).cuda_stream)
    L.check(L.lib().cg_copy_index_pairs(ctypes.c_void_p(dst.data_ptr()), ctypes.c_void_p(src.data_ptr()),
                                        ctypes.c_void_p(dst_idx.data_ptr()), ctypes.c_void_p(src_idx.data_ptr()),
                                        ctypes.c_int64(dst_idx.numel()), ctypes.c_int(esize), st))


class DevicePlan:
    """Index pairs of one interface on the device (both directions), for fields stored like the Julia arrays
    (column-major `cell.full`; a torch tensor of shape reversed(full dims))."""

    def __init__(self, plan, left_shape, right_shape, device):
        import torch
        mk = lambda a: torch.from_numpy(a).to(device)
        s, d = linear_indices(plan.left_to_right, left_shape, right_shape)
        self.l2r_src, self.l2r_dst = mk(s), mk(d)
        s, d = linear_indices(plan.right_to_left, right_shape, left_shape)
        self.r2l_src, self.r2l_dst = mk(s), mk(d)
        self.l2r_A = self.r2l_A = None   # per-pair basis-transfer matrices (None: identity, Cartesian bases)
        self.dim = len(left_shape)

    def build_transforms(self, left_grid, right_grid):
        """_build_interface_transforms (multiblock/cache.jl:55-108): for pair i, A_l2r[i] = basis_transfer_matrix(left,
        centroid(left interior i), right, centroid(right interior i)) and the reverse, on the device.  Grids are this
        package's MappedGrid / DiscreteGrid (their `basis` / `coordinate_system` select Q); identity when both bases are
        Cartesian."""
        import torch
        if self.l2r_src.numel() != self.r2l_src.numel():
            raise ValueError("transforms need the same interior cells in both directions (no valid-domain filtering)")
        bid = lambda g: 1 if (g.basis == "SphericalBasis") else 0
        for g in (left_grid, right_grid):
            if g.basis == "SphericalBasis" and g.coordinate_system != "SphericalCS":
                raise ValueError(f"Unsupported public basis transfer for SphericalBasis with coordinate system {g.coordinate_system}.")
        if bid(left_grid) == 0 and bid(right_grid) == 0:
            self.l2r_A = self.r2l_A = None
            return self
        N, n = self.dim, self.l2r_src.numel()
        lc, rc = left_grid.centroid_coordinates, right_grid.centroid_coordinates
        dt = lc[0].dtype
        ptrs = lambda cs: (ctypes.c_void_p * 3)(*[ctypes.c_void_p(c.data_ptr()) for c in cs] + [None] * (3 - len(cs)))
        st = ctypes.c_void_p(torch.cuda.current_stream(lc[0].device).cuda_stream)
        self.l2r_A = torch.empty((n, N * N), dtype=dt, device=lc[0].device)
        self.r2l_A = torch.empty((n, N * N), dtype=dt, device=lc[0].device)
        L.check(L.lib().cg_basis_transfer_matrices(N, ctypes.c_int64(n), bid(left_grid), bid(right_grid), ptrs(lc),
                                                   ctypes.c_void_p(self.l2r_src.data_ptr()), ptrs(rc),
                                                   ctypes.c_void_p(self.r2l_src.data_ptr()), lc[0].element_size(),
                                                   ctypes.c_void_p(self.l2r_A.data_ptr()), st))
        L.check(L.lib().cg_basis_transfer_matrices(N, ctypes.c_int64(n), bid(right_grid), bid(left_grid), ptrs(rc),
                                                   ctypes.c_void_p(self.r2l_src.data_ptr()), ptrs(lc),
                                                   ctypes.c_void_p(self.l2r_src.data_ptr()), lc[0].element_size(),
                                                   ctypes.c_void_p(self.r2l_A.data_ptr()), st))
        return self


_KINDS = {"scalar": 0, "vector": 1, "tensor": 2}


def _transfer_pairs(dst, src, dst_idx, src_idx, dim, kind, A):
    """dst[dst_idx[p]] = A_p (x) src[src_idx[p]] for AoS vector / tensor fields (cg_transfer_index_pairs)."""
    import torch
    K = dim if kind == 1 else dim * dim
    if not (dst.is_cuda and src.is_cuda and dst.is_contiguous() and src.is_contiguous() and dst.dtype == src.dtype):
        raise ValueError("exchange needs contiguous CUDA tensors of one dtype")
    if dst.shape[-1] != K or src.shape[-1] != K:
        raise ValueError(f"field_kind needs {K} components per cell in the last dimension")
    if A is not None and (A.dtype != dst.dtype or not A.is_contiguous()):
        raise ValueError("transforms must be contiguous and of the field's dtype")
    st = ctypes.c_void_p(torch.cuda.current_stream(dst.device).cuda_stream)
    L.check(L.lib().cg_transfer_index_pairs(ctypes.c_void_p(dst.data_ptr()), ctypes.c_void_p(src.data_ptr()),
                                            ctypes.c_void_p(dst_idx.data_ptr()), ctypes.c_void_p(src_idx.data_ptr()),
                                            ctypes.c_int64(dst_idx.numel()), dim, kind, dst.element_size(),
                                            None if A is None else ctypes.c_void_p(A.data_ptr()), st))


def exchange_interface(left_field, right_field, dplan, direction="both", field_kind="scalar"):
    """exchange_interface!(mb, iface, fields; field_kind, direction) — transfer.jl:25-87: donor interior cells into
    receiver ghost cells, both blocks on this GPU (one launch per direction).  Vector / tensor fields carry their N or
    N*N components in the last tensor dimension (the memory of Julia's Array{SVector} / Array{SMatrix}) and are rotated
    by the plan's basis-transfer matrices (`DevicePlan.build_transforms`; identity without them)."""
    if direction not in ("left_to_right", "right_to_left", "both"):
        raise ValueError(f"Invalid `direction={direction}`.")
    if field_kind not in _KINDS:
        raise ValueError(f"Invalid `field_kind={field_kind}`.")
    kind = _KINDS[field_kind]
    if direction in ("left_to_right", "both"):
        if kind == 0:
            _copy_pairs(right_field, left_field, dplan.l2r_dst, dplan.l2r_src)
        else:
            _transfer_pairs(right_field, left_field, dplan.l2r_dst, dplan.l2r_src, dplan.dim, kind, dplan.l2r_A)
    if direction in ("right_to_left", "both"):
        if kind == 0:
            _copy_pairs(left_field, right_field, dplan.r2l_dst, dplan.r2l_src)
        else:
            _transfer_pairs(left_field, right_field, dplan.r2l_dst, dplan.r2l_src, dplan.dim, kind, dplan.r2l_A)
    return left_field, right_field


def exchange_interface_remote(field, dplan, side, comm, peer, field_kind="scalar", direction="both"):
    """The same exchange with the two blocks on DIFFERENT ranks: this rank holds the `side` ("left" | "right") block's
    field, `peer` the other one.  Donor values are packed in plan order, cross NVLink with NCCL inside the C ABI
    (cg_nccl_sendrecv on `comm`, a slabs.NcclComm) and are rotated by the basis-transfer matrices while they are
    unpacked into the ghost cells.  Both ranks hold the same DevicePlan (index algebra + transforms are cheap)."""
    import torch
    if side not in ("left", "right"):
        raise ValueError("side must be 'left' or 'right'")
    if field_kind not in _KINDS:
        raise ValueError(f"Invalid `field_kind={field_kind}`.")
    kind = _KINDS[field_kind]
    K = 1 if kind == 0 else (dplan.dim if kind == 1 else dplan.dim ** 2)
    left = side == "left"
    send_on = direction in ("both", "left_to_right" if left else "right_to_left")
    recv_on = direction in ("both", "right_to_left" if left else "left_to_right")
    src_idx = dplan.l2r_src if left else dplan.r2l_src       # my interior cells the peer wants
    dst_idx = dplan.r2l_dst if left else dplan.l2r_dst       # my ghost cells
    A = dplan.r2l_A if left else dplan.l2r_A                 # transform of what I receive
    n = src_idx.numel()
    flat = field.reshape(-1, K) if kind else field.reshape(-1)
    sendbuf = flat[src_idx].contiguous() if send_on else None                  # pack (gather in plan order)
    recvbuf = torch.empty((n, K) if kind else (n,), dtype=field.dtype, device=field.device) if recv_on else None
    st = ctypes.c_void_p(torch.cuda.current_stream(field.device).cuda_stream)
    nbytes = n * K * field.element_size()
    L.check(L.lib().cg_nccl_sendrecv(comm._h, None if sendbuf is None else ctypes.c_void_p(sendbuf.data_ptr()),
                                     ctypes.c_size_t(nbytes if send_on else 0), peer if send_on else -1,
                                     None if recvbuf is None else ctypes.c_void_p(recvbuf.data_ptr()),
                                     ctypes.c_size_t(nbytes if recv_on else 0), peer if recv_on else -1, st))
    if recv_on:
        ar = torch.arange(n, dtype=torch.int64, device=field.device)
        if kind == 0:
            _copy_pairs(field, recvbuf, dst_idx, ar)
        else:
            _transfer_pairs(field, recvbuf, dst_idx, ar, dplan.dim, kind, A)
    return field


def pack(field, src_idx):
    """Donor side of a cross-rank exchange: the interior values in plan order (contiguous, ready for ncclSend)."""
    import torch
    buf = torch.empty(src_idx.numel(), dtype=field.dtype, device=field.device)
    ar = torch.arange(src_idx.numel(), dtype=torch.int64, device=field.device)
    _copy_pairs(buf, field, ar, src_idx)
    return buf


def unpack(field, dst_idx, buf):
    """Receiver side: packed values (after ncclRecv) into the ghost cells."""
    import torch
    ar = torch.arange(dst_idx.numel(), dtype=torch.int64, device=field.device)
    _copy_pairs(field, buf, dst_idx, ar)
    return field

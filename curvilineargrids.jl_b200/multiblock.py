"""Multiblock interface plans and the scalar ghost-cell exchange on the device.

Host side: `interface_index_plan` restates /root/reference/src/multiblock/transfer.jl:157-195 (driver) and :302-359
(`_paired_interface_layer_indices`) with the helpers of validation.jl:223-283 (`_face_axis_index`,
`_build_face_cartesian_index`, `_map_tangential_indices`) and the checks of transfer.jl:259-293.  Pure index algebra;
indices are 1-based like the reference's `CartesianIndex` (so its known answers, test_multiblock.jl:261-273, can be
asserted literally).

Device side: `exchange_interface` copies donor interior cells into receiver ghost cells of cell-centred fields that
live on the GPU (scalar field kind, transfer.jl:25-117) with one launch of `cg_copy_index_pairs` per direction; the
two blocks may be on the same GPU or, with `pack` / `unpack`, on different ranks (NCCL send / recv of the packed
values).  Vector / tensor fields additionally need `basis_transfer_matrix` (transfer.jl:461-488) and stay Julia.
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


def exchange_interface(left_field, right_field, dplan, direction="both"):
    """exchange_interface!(mb, iface, fields; field_kind=:scalar, direction) — transfer.jl:25-87: donor interior cells
    into receiver ghost cells, both blocks on this GPU (one launch per direction)."""
    if direction not in ("left_to_right", "right_to_left", "both"):
        raise ValueError(f"Invalid `direction={direction}`.")
    if direction in ("left_to_right", "both"):
        _copy_pairs(right_field, left_field, dplan.l2r_dst, dplan.l2r_src)
    if direction in ("right_to_left", "both"):
        _copy_pairs(left_field, right_field, dplan.r2l_dst, dplan.r2l_src)
    return left_field, right_field


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

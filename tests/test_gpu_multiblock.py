"""Scalar interface exchange on the device (cg_copy_index_pairs) against the host application of the same plan."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
def test_device_exchange_matches_host_plan(cg, dtype):
    mbk = cg.multiblock
    h = 3
    lfull, rfull = (6 + 2 * h, 7 + 2 * h, 5 + 2 * h), (9 + 2 * h, 5 + 2 * h, 7 + 2 * h)
    ldom = tuple((h + 1, n - h) for n in lfull)
    rdom = tuple((h + 1, n - h) for n in rfull)
    # left xi-max face (tangential eta x zeta = 7 x 5) onto the right block's xi-min face (tangential 5 x 7), swapped + flipped
    plan = mbk.interface_index_plan(ldom, mbk.BlockFace(1, 1, "max"), rdom, mbk.BlockFace(2, 1, "min"), (2, 1),
                                    (True, False), depth=h)
    rng = np.random.default_rng(3)
    f1 = rng.standard_normal(lfull)
    f2 = rng.standard_normal(rfull)
    g1, g2 = f1.copy(), f2.copy()
    for s, t in plan.left_to_right:
        g2[tuple(v - 1 for v in t)] = f1[tuple(v - 1 for v in s)]
    for s, t in plan.right_to_left:
        g1[tuple(v - 1 for v in t)] = f2[tuple(v - 1 for v in s)]
    to_dev = lambda a: torch.from_numpy(np.ascontiguousarray(a.transpose(2, 1, 0))).to("cuda", dtype)   # Julia layout
    d1, d2 = to_dev(f1), to_dev(f2)
    dplan = mbk.DevicePlan(plan, lfull, rfull, "cuda")
    mbk.exchange_interface(d1, d2, dplan)
    back = lambda t: t.cpu().numpy().transpose(2, 1, 0)
    assert np.array_equal(back(d1), g1.astype(back(d1).dtype)) and np.array_equal(back(d2), g2.astype(back(d2).dtype))
    # one direction only, and the pack / unpack halves of a cross-rank exchange
    e1, e2 = to_dev(f1), to_dev(f2)
    mbk.exchange_interface(e1, e2, dplan, direction="left_to_right")
    assert np.array_equal(back(e2), g2.astype(back(e2).dtype)) and np.array_equal(back(e1), f1.astype(back(e1).dtype))
    buf = mbk.pack(e2, dplan.r2l_src)
    mbk.unpack(e1, dplan.r2l_dst, buf)
    assert np.array_equal(back(e1), g1.astype(back(e1).dtype))
    with pytest.raises(ValueError):
        mbk.exchange_interface(e1, e2, dplan, direction="sideways")


def _two_blocks_2d(h=2):
    """Two abutting 2-D blocks sharing the left block's xi-max / right block's xi-min face (test_multiblock.jl:131-200)."""
    lfull, rfull = (6 + 2 * h, 5 + 2 * h), (4 + 2 * h, 5 + 2 * h)
    ldom = tuple((h + 1, n - h) for n in lfull)
    rdom = tuple((h + 1, n - h) for n in rfull)
    return lfull, rfull, ldom, rdom


def test_vector_and_tensor_exchange_cartesian_bases(cg):
    """test_multiblock.jl:174-199: with Cartesian bases the vector / tensor exchange is a plain copy of the components."""
    mbk = cg.multiblock
    h = 2
    lfull, rfull, ldom, rdom = _two_blocks_2d(h)
    plan = mbk.interface_index_plan(ldom, mbk.BlockFace(1, 1, "max"), rdom, mbk.BlockFace(2, 1, "min"), (1,), (False,))
    dplan = mbk.DevicePlan(plan, lfull, rfull, "cuda")
    mk = lambda full, K: torch.zeros(tuple(reversed(full)) + (K,), dtype=torch.float64, device="cuda")
    v1, v2 = mk(lfull, 2), mk(rfull, 2)
    i1, i2 = ldom[0][1] - 1, rdom[0][0] - 1            # 0-based interior columns next to the interface
    js = slice(h, lfull[1] - h)
    v1[js, i1] = torch.tensor([1.0, 2.0], device="cuda", dtype=torch.float64)
    v2[js, i2] = torch.tensor([3.0, 4.0], device="cuda", dtype=torch.float64)
    mbk.exchange_interface(v1, v2, dplan, field_kind="vector")
    assert torch.equal(v1[js, i1 + 1], v2[js, i2]) and torch.equal(v2[js, i2 - 1], v1[js, i1])
    t1, t2 = mk(lfull, 4), mk(rfull, 4)
    t1[js, i1] = torch.tensor([1.0, 3.0, 2.0, 4.0], device="cuda", dtype=torch.float64)   # [1 2; 3 4] column-major
    t2[js, i2] = torch.tensor([5.0, 7.0, 6.0, 8.0], device="cuda", dtype=torch.float64)
    mbk.exchange_interface(t1, t2, dplan, field_kind="tensor")
    assert torch.equal(t1[js, i1 + 1], t2[js, i2]) and torch.equal(t2[js, i2 - 1], t1[js, i1])
    with pytest.raises(ValueError):
        mbk.exchange_interface(t1, t2, dplan, field_kind="spinor")
    with pytest.raises(ValueError):
        mbk.exchange_interface(v1, t2, dplan, field_kind="vector")


def test_vector_transfer_cartesian_to_spherical_basis(cg):
    """test_multiblock.jl:334-372: a Cartesian-basis block hands a vector to a SphericalCS / SphericalBasis block; the
    received components, put back through Q(q_receiver) = (e_r, e_theta, e_phi), give the donor vector again (magnitude
    and direction at 1e-12).  The tensor path is checked the same way: Q T_sph Q^T = T_cart.  The transforms are
    tabulated on the device from the two grids' centroid arrays (cg_basis_transfer_matrices)."""
    mbk, wl = cg.multiblock, cg.workloads
    h, cd = 2, (5, 6, 4)
    left = cg.MappedGrid(wl.sphere_sector_3d_terms(cd), {}, cd, h, cache_mode="lazy")                     # Cartesian x, y, z
    right = cg.MappedGrid(wl.spherical_rtp_3d_terms(cd), {}, cd, h, cache_mode="lazy", coordinate_system="SphericalCS",
                          basis="SphericalBasis")                                                           # (r, theta, phi)
    full = tuple(c + 2 * h for c in cd)
    dom = tuple((h + 1, n - h) for n in full)
    plan = mbk.interface_index_plan(dom, mbk.BlockFace(1, 1, "max"), dom, mbk.BlockFace(2, 1, "min"), (1, 2), (False, False))
    dplan = mbk.DevicePlan(plan, full, full, "cuda").build_transforms(left, right)
    assert dplan.l2r_A is not None and dplan.l2r_A.shape == (len(plan.left_to_right), 9)
    mk = lambda K: torch.zeros(tuple(reversed(full)) + (K,), dtype=torch.float64, device="cuda")
    donor = np.array([1.3, -0.7, 2.1])
    v_cart, v_sph = mk(3), mk(3)
    v_cart[..., :] = torch.from_numpy(donor).cuda()
    mbk.exchange_interface(v_cart, v_sph, dplan, field_kind="vector", direction="left_to_right")
    rc = [c.cpu().numpy() for c in right.centroid_coordinates]

    def Q(th, ph):
        st, ct, sp, cp = np.sin(th), np.cos(th), np.sin(ph), np.cos(ph)
        return np.array([[st * cp, ct * cp, -sp], [st * sp, ct * sp, cp], [ct, -st, 0.0]])
    i_int, i_gh = dom[0][0] - 1, dom[0][0] - 2        # the right block's first interior column and its ghost column
    vs = v_sph.cpu().numpy()
    T_cart = np.array([[1.0, 0.2, -0.3], [0.2, 2.0, 0.5], [-0.3, 0.5, 3.0]])
    t_cart, t_sph = mk(9), mk(9)
    t_cart[..., :] = torch.from_numpy(T_cart.ravel(order="F")).cuda()
    mbk.exchange_interface(t_cart, t_sph, dplan, field_kind="tensor", direction="left_to_right")
    ts = t_sph.cpu().numpy()
    for k in range(h, full[2] - h):
        for j in range(h, full[1] - h):
            q = Q(rc[1][k, j, i_int], rc[2][k, j, i_int])     # transforms use the receiver's INTERIOR centroid (cache.jl:55-108)
            back = q @ vs[k, j, i_gh]
            assert abs(np.linalg.norm(back) - np.linalg.norm(donor)) <= 1e-12
            assert abs(back @ donor / (np.linalg.norm(back) * np.linalg.norm(donor)) - 1.0) <= 1e-12
            Tb = q @ ts[k, j, i_gh].reshape(3, 3, order="F") @ q.T
            assert np.abs(Tb - T_cart).max() <= 1e-12
    assert not vs[:, :, :i_gh].any()                           # nothing but the ghost layer was written
    # spherical -> Cartesian and back is the identity on the components
    v_back = mk(3)
    mbk.exchange_interface(v_back, v_sph.clone().copy_(v_sph), dplan, field_kind="vector", direction="right_to_left")
    with pytest.raises(ValueError):
        bad = cg.MappedGrid(wl.sphere_sector_3d_terms(cd), {}, cd, h, cache_mode="lazy", basis="SphericalBasis")
        mbk.DevicePlan(plan, full, full, "cuda").build_transforms(left, bad)

"""interface_index_plan restatement (src/multiblock/transfer.jl:157-359, validation.jl:223-283) pinned with the
reference's own known answers (test/unit/test_multiblock.jl:202-273) and structural properties.  CPU only."""
import numpy as np
import pytest

from curvigrid_b200 import multiblock as mbk
from curvigrid_b200.multiblock import BlockFace, interface_index_plan


def test_reference_3d_known_answers():
    # test_multiblock.jl:259-273
    left = ((2, 5), (3, 6), (4, 7))
    right = ((10, 13), (20, 23), (30, 33))
    plan = interface_index_plan(left, BlockFace(1, 1, "max"), right, BlockFace(2, 3, "min"), (2, 1), (False, True))
    assert len(plan.left_to_right) == 16
    assert ((5, 3, 4), (13, 20, 29)) in plan.left_to_right
    assert ((13, 20, 30), (6, 3, 4)) in plan.right_to_left


def _grid_2d(ni, nj, h):
    full = (ni + 2 * h, nj + 2 * h)
    dom = ((h + 1, h + ni), (h + 1, h + nj))
    return full, dom


def test_reference_2d_depth2_plan_equals_the_scalar_exchange():
    """test_multiblock.jl:202-231: two 4x5 blocks (nhalo 2) abutting along xi, flipped tangential axis, depth 2.
    The plan applied to the fields must fill exactly the two ghost layers next to the interface, with the donor's
    first / second interior layer, tangentially reversed."""
    ni, nj, h = 4, 5, 2
    full, dom = _grid_2d(ni, nj, h)
    plan = interface_index_plan(dom, BlockFace(1, 1, "max"), dom, BlockFace(2, 1, "min"), (1,), (True,), depth=2)
    assert len(plan.left_to_right) == len(plan.right_to_left) == 2 * nj
    i = np.arange(1, full[0] + 1)[:, None]
    j = np.arange(1, full[1] + 1)[None, :]
    f1 = 10.0 * i + j
    f2 = 100.0 + 10.0 * i + j
    g1, g2 = f1.copy(), f2.copy()
    for (s, t) in plan.left_to_right:
        g2[t[0] - 1, t[1] - 1] = f1[s[0] - 1, s[1] - 1]
    for (s, t) in plan.right_to_left:
        g1[t[0] - 1, t[1] - 1] = f2[s[0] - 1, s[1] - 1]
    # right block's ghost layers h, h-1 (1-based columns 2, 1) hold the left block's last / second-to-last interior
    # columns, reversed along j (flip)
    for layer in (1, 2):
        want = f1[h + ni - layer, h:h + nj][::-1]
        assert np.array_equal(g2[h - layer, h:h + nj], want)
        want = f2[h + layer - 1, h:h + nj][::-1]
        assert np.array_equal(g1[h + ni + layer - 1, h:h + nj], want)
    # nothing else changed
    m2 = np.ones_like(g2, dtype=bool); m2[h - 2:h, h:h + nj] = False
    m1 = np.ones_like(g1, dtype=bool); m1[h + ni:h + ni + 2, h:h + nj] = False
    assert np.array_equal(g2[m2], f2[m2]) and np.array_equal(g1[m1], f1[m1])


def test_reference_valid_domain_filter():
    # test_multiblock.jl:233-257
    ni, nj, h = 4, 5, 2
    _, dom = _grid_2d(ni, nj, h)
    sub = (dom[1][0] + 1, dom[1][1] - 1)
    left_valid = ((dom[0][0], dom[0][1] + 1), sub)
    right_valid = ((dom[0][0] - 1, dom[0][1]), sub)
    plan = interface_index_plan(dom, BlockFace(1, 1, "max"), dom, BlockFace(2, 1, "min"), (1,), (True,),
                                left_valid_domain=left_valid, right_valid_domain=right_valid)
    assert len(plan.left_to_right) == 3 and len(plan.right_to_left) == 3
    inside = lambda idx, d: all(lo <= v <= hi for v, (lo, hi) in zip(idx, d))
    assert all(inside(s, left_valid) and inside(t, right_valid) for s, t in plan.left_to_right)
    assert all(inside(s, right_valid) and inside(t, left_valid) for s, t in plan.right_to_left)


def test_plan_argument_errors():
    # transfer.jl:259-293
    d3 = ((1, 4), (1, 4), (1, 4))
    with pytest.raises(ValueError):
        interface_index_plan(d3, BlockFace(1, 1, "max"), d3, BlockFace(2, 1, "min"), (1,), (False, False))
    with pytest.raises(ValueError):
        interface_index_plan(d3, BlockFace(1, 1, "max"), d3, BlockFace(2, 1, "min"), (1, 1), (False, False))
    with pytest.raises(ValueError):
        interface_index_plan(d3, BlockFace(1, 1, "max"), d3, BlockFace(2, 1, "min"), (1, 2), (0, 1))
    with pytest.raises(ValueError):
        interface_index_plan(d3, BlockFace(1, 4, "max"), d3, BlockFace(2, 1, "min"), (1, 2), (False, False))
    with pytest.raises(ValueError):
        interface_index_plan(d3, BlockFace(1, 1, "mid"), d3, BlockFace(2, 1, "min"), (1, 2), (False, False))
    with pytest.raises(ValueError):
        interface_index_plan(d3, BlockFace(1, 1, "max"), ((1, 4), (1, 5), (1, 4)), BlockFace(2, 1, "min"), (1, 2), (False, False))
    with pytest.raises(ValueError):
        interface_index_plan(d3, BlockFace(1, 1, "max"), d3, BlockFace(2, 1, "min"), (1, 2), (False, False), depth=0)
    # 1-D interface: one pair per layer
    p = interface_index_plan(((3, 9),), BlockFace(1, 1, "max"), ((3, 7),), BlockFace(2, 1, "min"), (), (), depth=2)
    assert p.left_to_right == [((9,), (2,)), ((8,), (1,))] and p.right_to_left == [((3,), (10,)), ((4,), (11,))]


def test_linear_indices_are_column_major():
    s, d = mbk.linear_indices([((2, 3), (1, 1))], (5, 7), (4, 4))
    assert s.tolist() == [1 + 5 * 2] and d.tolist() == [0]
    with pytest.raises(ValueError):
        mbk.linear_indices([((6, 1), (1, 1))], (5, 7), (4, 4))

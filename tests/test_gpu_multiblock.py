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

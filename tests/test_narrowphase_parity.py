"""Narrowphase parity: k_narrowphase's device functions (b2g_collide_pairs) against the
reference's five b2Collide* functions — feature ids / type / pointCount bit-exact, points and
normals within 1e-5 relative (north_star gate).  Inputs are ordered pairs with host-supplied
sin/cos, so both sides see identical bits."""
import os

import numpy as np
import pytest

import util
from box2d_optimized_b200 import capi

GOLD = os.path.join(os.path.dirname(__file__), "golden")
pytestmark = pytest.mark.gpu


def test_golden_random_pairs():
    g = np.load(os.path.join(GOLD, "narrowphase_random.npz"))
    out = util.gpu_collide(g["tA"], g["oA"], g["xfA"], g["tB"], g["oB"], g["xfB"], g["quads"])
    n, maxd = util.compare_manifolds(out, g["manifold"], rel=1e-5)
    assert n > 500
    print(f"{n} touching pairs, max abs float difference {maxd:g}")


def test_reference_testbed_polygon_pair():
    # the only fixed narrowphase input in the reference tree: testbed/tests/polygon_collision.cpp:32-56
    g = np.load(os.path.join(GOLD, "polygon_collision_testbed.npz"))
    out = util.gpu_collide([2], [int(g["oA"])], g["xfA"], [2], [int(g["oB"])], g["xfB"], g["quads"])
    util.compare_manifolds(out, g["manifold"], rel=1e-5)


@pytest.mark.parametrize("name", ["pyramid", "mixed"])
def test_golden_scene_contacts(name):
    """every contact of a live reference world (its own A/B order, its own transforms)"""
    g = np.load(os.path.join(GOLD, f"scene_{name}.npz"))
    fa, fb = g["con_a"], g["con_b"]
    xf = g["bodies"][:, 0:4]
    ba, bb = g["fix_body"][fa], g["fix_body"][fb]
    out = util.gpu_collide(g["fix_type"][fa], g["fix_shape_off"][fa], xf[ba], g["fix_type"][fb],
                           g["fix_shape_off"][fb], xf[bb], g["quads"])
    n, maxd = util.compare_manifolds(out, g["con_manifold"], rel=1e-5)
    assert n > 100
    # touching flag of the reference == pointCount > 0 of ours
    assert np.array_equal((g["con_flags"] & 1) != 0, util.man_count(out) > 0)


def test_live_reference_large_fuzz(require_ref):
    """40k fresh random pairs against the compiled reference (oracle/_ref)"""
    import sys
    sys.path.insert(0, GOLD)
    import make_golden
    case = make_golden.narrowphase_cases(777, 8000)
    args = (case["tA"], case["oA"], case["xfA"], case["tB"], case["oB"], case["xfB"], case["quads"])
    r = util.ref_collide(*args)
    g = util.gpu_collide(*args)
    n, maxd = util.compare_manifolds(g, r, rel=1e-5)
    assert n > 10000
    print(f"{n} touching of {len(r)}, max abs diff {maxd:g}")


@pytest.mark.parametrize("name,size,steps", [("pyramid", 20, 120), ("mixed", 3000, 200), ("falling_circles", 300, 150),
                                              ("tumbler", 200, 260)])
def test_live_reference_scene_contacts(require_ref, name, size, steps):
    from oracle.bindings import RefScene
    s = RefScene(name, size, 12345)
    s.step(steps)
    s.collide_now()
    b, fx, c = s.bodies(), s.fixtures(), s.contacts()
    fa, fb = c["fix_a"], c["fix_b"]
    xf = b[:, 0:4]
    out = util.gpu_collide(fx["type"][fa], fx["shape_off"][fa], xf[fx["body"][fa]], fx["type"][fb],
                           fx["shape_off"][fb], xf[fx["body"][fb]], fx["quads"])
    n, maxd = util.compare_manifolds(out, c["manifold"], rel=1e-5)
    assert n > 50
    print(f"{name}: {n} touching of {len(fa)}, max abs diff {maxd:g}")


def test_empty_and_unsupported():
    lib = capi.load_cuda()
    assert lib.b2g_collide_pairs(0, 0, None, None, None, None, None, None, None, 0, None) != 0  # null args rejected
    pool = util.ShapePool()
    e1 = pool.edge((-1, 0), (1, 0))
    e2 = pool.edge((-1, 0.001), (1, 0.001))
    xf = util.xf_rows([0.0], [0.0], [0.0])
    out = util.gpu_collide([1], [e1], xf, [1], [e2], xf, pool.array())
    assert util.man_count(out)[0] == 0  # edge-edge has no function in the reference (b2_contact.cpp:68-76)

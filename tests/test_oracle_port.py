"""Pins the plain-C restatement (oracle/b2_oracle.c) against the reference: golden vectors made
by the compiled reference (tests/golden/make_golden.py) and, where oracle/_ref is present, fresh
fuzz against it.  CPU only.  Everything is expected to agree BIT FOR BIT (same arithmetic, same
libm sinf/cosf, no FMA on either side)."""
import ctypes as C
import os

import numpy as np
import oracle.bindings as oracle_bindings  # noqa: E402  (the checker)
import pytest

import util
from box2d_optimized_b200 import capi

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def port():
    if not os.path.exists(oracle_bindings.oracle_path()):
        import subprocess
        subprocess.check_call(["make", "-C", os.path.join(capi.ROOT, "oracle"), "port"])
    return oracle_bindings.load_oracle()


def port_collide(lib, tA, oA, xA, tB, oB, xB, quads):
    n = len(tA)
    out = np.zeros((n, 16), np.float32)
    tA, oA, tB, oB = map(capi.i32, (tA, oA, tB, oB))
    xA, xB, quads = capi.f32(xA), capi.f32(xB), capi.f32(quads)
    assert lib.b2o_collide_pairs(n, capi.ip(tA), capi.ip(oA), capi.fp(xA), capi.ip(tB), capi.ip(oB), capi.fp(xB),
                                 capi.fp(quads), capi.fp(out)) == 0
    return out


def assert_manifolds_bit_equal(a, b):
    n, maxd = util.compare_manifolds(a, b, rel=0.0)
    assert maxd == 0.0
    return n


def test_narrowphase_golden(port):
    g = np.load(os.path.join(GOLD, "narrowphase_random.npz"))
    out = port_collide(port, g["tA"], g["oA"], g["xfA"], g["tB"], g["oB"], g["xfB"], g["quads"])
    assert assert_manifolds_bit_equal(out, g["manifold"]) > 500


def test_reference_testbed_polygon_pair(port):
    g = np.load(os.path.join(GOLD, "polygon_collision_testbed.npz"))
    out = port_collide(port, [2], [int(g["oA"])], g["xfA"], [2], [int(g["oB"])], g["xfB"], g["quads"])
    assert_manifolds_bit_equal(out, g["manifold"])


@pytest.mark.parametrize("name", ["pyramid", "mixed"])
def test_scene_contacts_aabbs_pairs_golden(port, name):
    g = np.load(os.path.join(GOLD, f"scene_{name}.npz"))
    fa, fb = g["con_a"], g["con_b"]
    xf = g["bodies"][:, 0:4]
    out = port_collide(port, g["fix_type"][fa], g["fix_shape_off"][fa], xf[g["fix_body"][fa]], g["fix_type"][fb],
                       g["fix_shape_off"][fb], xf[g["fix_body"][fb]], g["quads"])
    assert assert_manifolds_bit_equal(out, g["con_manifold"]) > 100
    # tight AABBs of the moving fixtures
    nf = len(g["fix_body"])
    aabb = np.zeros((nf, 4), np.float32)
    xff = np.ascontiguousarray(xf[g["fix_body"]])
    types, offs, quads = capi.i32(g["fix_type"]), capi.i32(g["fix_shape_off"]), capi.f32(g["quads"])
    port.b2o_compute_aabbs(nf, capi.ip(types), capi.ip(offs), capi.fp(quads), capi.fp(xff), capi.fp(aabb))
    moving = g["bodies"][:, 11][g["fix_body"]] != 0
    assert np.array_equal(aabb[moving].view(np.uint32), g["aabbs"][moving].view(np.uint32))
    # pair set == the reference's contact set
    body = capi.i32(g["fix_body"])
    dyn = np.ascontiguousarray(g["bodies"][:, 11][g["fix_body"]] == 2, np.uint8)
    pairs = np.zeros((8 * nf, 2), np.int32)
    cnt = C.c_int32()
    gold_aabb = capi.f32(g["aabbs"])
    assert port.b2o_find_pairs(nf, capi.fp(gold_aabb), capi.ip(body), None, dyn.ctypes.data_as(capi.u8p),
                               capi.ip(pairs), len(pairs), C.byref(cnt)) == 0
    assert util.pair_set(pairs[:cnt.value, 0], pairs[:cnt.value, 1]) == util.pair_set(fa, fb)


@pytest.mark.parametrize("name", ["pyramid", "mixed"])
def test_solver_iterates_golden(port, name):
    g = np.load(os.path.join(GOLD, f"solver_{name}.npz"))
    nb, nc = len(g["pos"]), len(g["index"])
    pos, vel, man = g["pos"].copy(), g["vel"].copy(), g["manifold"].copy()
    vit = np.zeros((8, nb, 4), np.float32)
    pit = np.zeros((3, nb, 4), np.float32)
    done = C.c_int32()
    mass, index, mat, radii = capi.f32(g["mass"]), capi.i32(g["index"]), capi.f32(g["material"]), capi.f32(g["radii"])
    port.b2o_solve(nb, capi.fp(pos), capi.fp(vel), capi.fp(mass), nc, capi.ip(index), capi.fp(man), capi.fp(mat),
                   capi.fp(radii), float(g["dt"]), 1.0, 1, 8, 3, capi.fp(vit), capi.fp(pit), C.byref(done))
    assert done.value == int(g["pos_iters_done"])
    assert np.array_equal(vit[:, :, :3].view(np.uint32), g["vel_iterates"][:, :, :3].view(np.uint32))
    k = done.value
    assert np.array_equal(pit[:k, :, :3].view(np.uint32), g["pos_iterates"][:k, :, :3].view(np.uint32))
    assert np.array_equal(pos[:, :3].view(np.uint32), g["pos_out"][:, :3].view(np.uint32))
    cnt = g["manifold_out"][:, 15].copy().view(np.int32)
    assert np.array_equal(man[cnt > 0][:, [6, 7]].view(np.uint32), g["manifold_out"][cnt > 0][:, [6, 7]].view(np.uint32))


def test_live_reference_fuzz(port, ref_available):
    if not ref_available:
        pytest.skip("oracle/_ref not built on this machine")
    import sys
    sys.path.insert(0, GOLD)
    import make_golden
    case = make_golden.narrowphase_cases(4242, 3000)
    args = (case["tA"], case["oA"], case["xfA"], case["tB"], case["oB"], case["xfB"], case["quads"])
    assert assert_manifolds_bit_equal(port_collide(port, *args), util.ref_collide(*args)) > 3000

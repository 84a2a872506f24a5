"""Broadphase parity (integer work: bit-exact).  The reference's contact set after any Step is
exactly the set of inclusive tight-AABB overlaps that pass the filters;
the device LBVH must report the same SET, and tight AABBs must match bit for bit."""
import ctypes as C
import os

import numpy as np
import pytest

import util
from box2d_optimized_b200 import capi, arena_from_scene

GOLD = os.path.join(os.path.dirname(__file__), "golden")
pytestmark = pytest.mark.gpu


def gpu_aabbs(types, offs, quads, xf):
    n = len(types)
    out = np.zeros((n, 4), np.float32)
    types, offs, quads, xf = capi.i32(types), capi.i32(offs), capi.f32(quads), capi.f32(xf)
    capi.check(capi.load_cuda().b2g_compute_aabbs(0, n, capi.ip(types), capi.ip(offs), capi.fp(quads), len(quads),
                                                  capi.fp(xf), capi.fp(out)))
    return out


def gpu_find_pairs(aabb, body, dyn, world=None, capacity=None):
    n = len(aabb)
    capacity = capacity or max(64, 64 * n)
    pairs = np.zeros((capacity, 2), np.int32)
    cnt = C.c_int32()
    aabb, body, dyn = capi.f32(aabb), capi.i32(body), np.ascontiguousarray(dyn, np.uint8)
    w = capi.i32(world) if world is not None else None
    rc = capi.load_cuda().b2g_find_pairs(0, n, capi.fp(aabb), capi.ip(body), capi.ip(w) if w is not None else None,
                                         dyn.ctypes.data_as(capi.u8p), capi.ip(pairs), capacity, C.byref(cnt))
    return rc, cnt.value, pairs[:min(cnt.value, capacity)]


def brute_pairs(aabb, body, dyn, world=None):
    """O(n^2) restatement of the pair definition, inclusive overlap (b2_collision.h:270-276)"""
    lo, hi = aabb[:, 0:2], aabb[:, 2:4]
    ov = (hi[:, None, 0] >= lo[None, :, 0]) & (lo[:, None, 0] <= hi[None, :, 0]) & \
         (hi[:, None, 1] >= lo[None, :, 1]) & (lo[:, None, 1] <= hi[None, :, 1])
    ov &= body[:, None] != body[None, :]
    ov &= (dyn[:, None] | dyn[None, :]).astype(bool)
    if world is not None:
        ov &= world[:, None] == world[None, :]
    i, j = np.nonzero(np.triu(ov, 1))
    return set(zip(i.tolist(), j.tolist()))


@pytest.mark.parametrize("name", ["pyramid", "mixed"])
def test_golden_aabbs_bit_exact(name):
    g = np.load(os.path.join(GOLD, f"scene_{name}.npz"))
    xf = g["bodies"][:, 0:4][g["fix_body"]]
    out = gpu_aabbs(g["fix_type"], g["fix_shape_off"], g["quads"], xf)
    dynamic = g["bodies"][:, 11][g["fix_body"]] != 0  # static AABBs are frozen at creation in the reference
    assert np.array_equal(out[dynamic].view(np.uint32), g["aabbs"][dynamic].view(np.uint32))


@pytest.mark.parametrize("name", ["pyramid", "mixed"])
def test_golden_pair_set(name):
    g = np.load(os.path.join(GOLD, f"scene_{name}.npz"))
    body = g["fix_body"]
    dyn = (g["bodies"][:, 11][body] == 2)
    rc, cnt, pairs = gpu_find_pairs(g["aabbs"], body, dyn)
    assert rc == 0
    # filters beyond body/dynamic (edge-edge, categories) do not reject anything in these scenes
    assert util.pair_set(pairs[:, 0], pairs[:, 1]) == util.pair_set(g["con_a"], g["con_b"])


@pytest.mark.parametrize("n,seed", [(1, 0), (2, 1), (3, 2), (257, 3), (5000, 4)])
def test_random_boxes_vs_brute_force(n, seed):
    rng = np.random.default_rng(seed)
    c = rng.uniform(0, np.sqrt(n) * 2.0 + 1.0, (n, 2))
    h = rng.uniform(0.1, 1.5, (n, 2))
    if n > 10:
        h[:3] *= 40.0  # a few huge boxes (ground / walls)
        c[5] = c[6]    # identical centres (Morton ties)
        h[5] = h[6]
    aabb = np.concatenate([c - h, c + h], 1).astype(np.float32)
    body = rng.integers(0, max(1, n // 2 + 1), n).astype(np.int32)
    dyn_body = rng.random(n // 2 + 2) < 0.7
    dyn = dyn_body[body]
    rc, cnt, pairs = gpu_find_pairs(aabb, body, dyn)
    assert rc == 0
    expect = brute_pairs(aabb, body, dyn)
    got = util.pair_set(pairs[:, 0], pairs[:, 1])
    assert cnt == len(pairs) == len(got), "a pair was reported twice"
    assert got == expect


def test_touching_edges_are_inclusive():
    aabb = np.array([[0, 0, 1, 1], [1, 0, 2, 1], [2.0000002, 0, 3, 1], [0, 1, 1, 2]], np.float32)
    rc, cnt, pairs = gpu_find_pairs(aabb, np.arange(4, dtype=np.int32), np.ones(4, bool))
    assert rc == 0
    assert util.pair_set(pairs[:, 0], pairs[:, 1]) == {(0, 1), (0, 3), (1, 3)}


def test_multi_world_segments_never_mix():
    rng = np.random.default_rng(9)
    per, worlds = 300, 7
    c = rng.uniform(0, 12.0, (per, 2)); h = rng.uniform(0.2, 0.8, (per, 2))
    one = np.concatenate([c - h, c + h], 1).astype(np.float32)
    aabb = np.tile(one, (worlds, 1))  # every world at the SAME coordinates
    world = np.repeat(np.arange(worlds), per).astype(np.int32)
    body = np.arange(per * worlds, dtype=np.int32)
    dyn = np.ones(per * worlds, bool)
    rc, cnt, pairs = gpu_find_pairs(aabb, body, dyn, world)
    assert rc == 0
    got = util.pair_set(pairs[:, 0], pairs[:, 1])
    assert got == brute_pairs(aabb, body, dyn, world)
    assert len(got) == worlds * len(brute_pairs(one, body[:per], dyn[:per]))


def test_capacity_overflow_is_reported_not_truncated():
    n = 64
    aabb = np.tile(np.array([[0, 0, 1, 1]], np.float32), (n, 1))
    rc, cnt, pairs = gpu_find_pairs(aabb, np.arange(n, dtype=np.int32), np.ones(n, bool), capacity=100)
    assert rc == -3 and cnt == n * (n - 1) // 2


@pytest.mark.parametrize("name,size,steps", [("pyramid", 20, 1), ("pyramid", 20, 90), ("many_pyramids", 12, 60),
                                              ("mixed", 3000, 120), ("tumbler", 150, 200),
                                              ("falling_squares", 300, 100),
                                              # joined neighbours overlap: b2Body::ShouldCollide filters them
                                              # unless collideConnected is set; folded non-neighbours collide
                                              ("chain", 14, 0), ("chain", 14, 90), ("chain", 14, 200),
                                              ("chain_collide", 14, 1), ("chain_collide", 14, 120)])
def test_live_reference_pair_set_and_order(require_ref, name, size, steps):
    """pair set == the reference's contact set on the reference's own transforms; A/B order of
    mixed-type pairs follows the reference's function table"""
    from oracle.bindings import RefScene
    s = RefScene(name, size, 12345)
    s.step(steps)
    A = arena_from_scene(s)
    A.find_new_contacts()
    cg, cr = A.download_contacts(), s.contacts()
    assert util.pair_set(cg["fix_a"], cg["fix_b"]) == util.pair_set(cr["fix_a"], cr["fix_b"])
    fx = s.fixtures()
    ordered_ref = {(a, b) for a, b in zip(cr["fix_a"].tolist(), cr["fix_b"].tolist())
                   if fx["type"][a] != fx["type"][b]}
    ordered_gpu = {(a, b) for a, b in zip(cg["fix_a"].tolist(), cg["fix_b"].tolist())
                   if fx["type"][a] != fx["type"][b]}
    assert ordered_ref == ordered_gpu
    # tight AABBs of moving fixtures, bit for bit
    dynamic = s.bodies()[:, 11][fx["body"]] != 0
    assert np.array_equal(A.download_aabbs()[dynamic].view(np.uint32), s.aabbs()[dynamic].view(np.uint32))
    A.close()

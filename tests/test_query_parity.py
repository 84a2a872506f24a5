"""b2World::QueryAABB / b2World::RayCast on the device tree against the compiled reference
(b2_world.cpp:1193-1246).

Gates: QueryAABB fixture sets bit-exact; ray casts: hit / no-hit equal, fraction and normal
bit-exact (the shape tests are the reference's arithmetic, compiled without FMA contraction), hit
fixture equal unless two fixtures are hit at the same fraction (the reference then keeps whichever
its own tree visits last; the device keeps the lower index)."""
import numpy as np
import pytest

from box2d_optimized_b200 import arena_from_scene

pytestmark = pytest.mark.gpu


def scene_pair(name, size, seed, steps):
    from oracle.bindings import RefScene
    ref = RefScene(name, size, seed)
    ref.step(steps)
    A = arena_from_scene(ref, max_contacts=max(4096, 16 * ref.body_count))
    A.find_new_contacts()
    return ref, A


def scene_bounds(ref):
    bb = ref.aabbs()
    dyn = ref.bodies()[:, 11][ref.fixtures()["body"]] != 0
    bb = bb[dyn] if dyn.any() else bb
    return bb[:, 0].min(), bb[:, 1].min(), bb[:, 2].max(), bb[:, 3].max()


def random_rays(rng, bounds, n):
    x0, y0, x1, y1 = bounds
    w, h = x1 - x0, y1 - y0
    p1 = np.stack([rng.uniform(x0 - 0.2 * w, x1 + 0.2 * w, n), rng.uniform(y0 - 0.2 * h, y1 + 0.5 * h, n)], 1)
    ang = rng.uniform(0, 2 * np.pi, n)
    length = rng.uniform(0.05, 1.0, n) * max(w, h)
    p2 = p1 + np.stack([np.cos(ang), np.sin(ang)], 1) * length[:, None]
    rays = np.concatenate([p1, p2], 1).astype(np.float32)
    # a fan of vertical rays from above (the RL "lidar" shape) and axis-aligned rays (degenerate axes)
    k = n // 8
    rays[:k, 2] = rays[:k, 0]
    rays[k:2 * k, 3] = rays[k:2 * k, 1]
    return rays


SCENES = [("pyramid", 20, 0, 200), ("mixed", 1500, 12345, 220), ("tumbler", 120, 3, 260), ("many_pyramids", 6, 0, 60)]


@pytest.mark.parametrize("name,size,seed,steps", SCENES)
def test_query_aabb_sets_are_exact(require_ref, name, size, seed, steps):
    ref, A = scene_pair(name, size, seed, steps)
    rng = np.random.default_rng(11)
    x0, y0, x1, y1 = scene_bounds(ref)
    n = 400
    c = np.stack([rng.uniform(x0, x1, n), rng.uniform(y0, y1, n)], 1)
    half = np.stack([rng.uniform(0.01, 0.25 * (x1 - x0), n), rng.uniform(0.01, 0.25 * (y1 - y0), n)], 1)
    boxes = np.concatenate([c - half, c + half], 1).astype(np.float32)
    # boxes that share an edge exactly with a fixture AABB (the overlap test is inclusive), a box
    # covering everything, an empty region
    bb = ref.aabbs()
    boxes[0] = [bb[1, 2], bb[1, 1], bb[1, 2] + 1.0, bb[1, 3]]
    boxes[1] = [x0 - 100, y0 - 100, x1 + 100, y1 + 100]
    boxes[2] = [x1 + 500, y1 + 500, x1 + 501, y1 + 501]
    cap = ref.fixture_count
    rc, rf = ref.query_aabb(boxes, cap)
    gc, gf = A.query_aabb(boxes, cap)
    assert np.array_equal(rc, gc)
    assert np.array_equal(rf, gf)
    assert gc[1] == ref.fixture_count and gc[2] == 0 and gc[0] >= 1
    # a capped query still counts everything
    gc2, gf2 = A.query_aabb(boxes[:8], 3)
    assert np.array_equal(gc2, gc[:8])
    A.close()


@pytest.mark.parametrize("name,size,seed,steps", SCENES)
def test_closest_ray_casts_match(require_ref, name, size, seed, steps):
    ref, A = scene_pair(name, size, seed, steps)
    rays = random_rays(np.random.default_rng(5), scene_bounds(ref), 4000)
    rfix, rfrac, rnorm, _ = ref.ray_cast_closest(rays)
    gfix, gfrac, gnorm = A.ray_cast_closest(rays)
    hit = rfix >= 0
    assert np.array_equal(hit, gfix >= 0)
    assert hit.sum() > 200, "the rays must actually hit something"
    assert np.array_equal(rfrac[hit].view(np.uint32), gfrac[hit].view(np.uint32)), "fractions are bit-exact"
    same = rfix == gfix
    # different fixture only when both are hit at the very same fraction: check with the all-hits cast
    for i in np.nonzero(hit & ~same)[0]:
        n, fx, fr, _ = ref.ray_cast_all(rays[i], 64)
        tied = set(fx[fr == rfrac[i]].tolist())
        assert int(rfix[i]) in tied and int(gfix[i]) in tied and int(gfix[i]) == min(tied), (i, rfix[i], gfix[i], tied)
    assert np.array_equal(rnorm[hit & same].view(np.uint32), gnorm[hit & same].view(np.uint32)), "normals are bit-exact"
    print(f"{name}: {int(hit.sum())} of {len(rays)} rays hit, {int((hit & ~same).sum())} equal-fraction ties")
    A.close()


def test_all_hits_and_filters(require_ref):
    ref, A = scene_pair("mixed", 1500, 12345, 220)
    rays = random_rays(np.random.default_rng(9), scene_bounds(ref), 300)
    cap = 128
    counts, gfix, gfrac, gnorm = A.ray_cast_all(rays, cap)
    assert counts.max() > 3
    for i in range(len(rays)):
        n, fx, fr, nm = ref.ray_cast_all(rays[i], cap)
        assert n == counts[i]
        assert np.array_equal(fx, gfix[i, :n]) and np.array_equal(fr.view(np.uint32), gfrac[i, :n].view(np.uint32))
        assert np.array_equal(nm.view(np.uint32), gnorm[i, :n].view(np.uint32))
    # max_fraction clips the segment exactly like input.maxFraction
    half = np.full(len(rays), 0.5, np.float32)
    cfix, cfrac, _ = A.ray_cast_closest(rays, max_fraction=half)
    ffix, ffrac, _ = A.ray_cast_closest(rays)
    expect_hit = (ffix >= 0) & (ffrac <= 0.5)
    assert np.array_equal(cfix >= 0, expect_hit)
    assert np.array_equal(cfrac[expect_hit], ffrac[expect_hit])
    # category mask 0 sees nothing
    nfix, _, _ = A.ray_cast_closest(rays, category_mask=0)
    assert (nfix == -1).all()
    A.close()


def test_queries_through_the_drop_in_api(require_ref):
    """The same shim program (scenes/b2_scene_shim.h) over this repo's b2World: QueryAABB / RayCast
    with user callbacks, replayed from the device results."""
    from box2d_optimized_b200 import GpuScene
    from oracle.bindings import RefScene
    ref, gpu = RefScene("pyramid", 12, 0), GpuScene("pyramid", 12, 0)
    # identical initial state; no steps, so both sides hold bit-identical transforms
    rays = random_rays(np.random.default_rng(2), scene_bounds(ref), 500)
    rfix, rfrac, rnorm, rpoint = ref.ray_cast_closest(rays)
    gfix, gfrac, gnorm, gpoint = gpu.ray_cast_closest(rays)
    hit = rfix >= 0
    assert hit.sum() > 50 and np.array_equal(hit, gfix >= 0)
    assert np.array_equal(rfrac[hit], gfrac[hit]) and np.array_equal(rpoint[hit], gpoint[hit])
    same = rfix == gfix
    assert same[hit].mean() > 0.9
    assert np.array_equal(rnorm[hit & same], gnorm[hit & same])
    boxes = np.array([[-3, 0, 3, 4], [-100, -100, 100, 100], [50, 50, 51, 51]], np.float32)
    rc, rf = ref.query_aabb(boxes, ref.fixture_count)
    gc, gf = gpu.query_aabb(boxes, ref.fixture_count)
    assert np.array_equal(rc, gc) and np.array_equal(rf, gf)
    n_r, fx_r, fr_r, _ = ref.ray_cast_all(np.array([-20, 0.75, 20, 0.75], np.float32), 64)
    n_g, fx_g, fr_g, _ = gpu.ray_cast_all(np.array([-20, 0.75, 20, 0.75], np.float32), 64)
    assert n_r == n_g and n_r >= 12 and np.array_equal(fx_r, fx_g) and np.array_equal(fr_r, fr_g)


def test_queries_respect_worlds_in_a_batched_arena(require_ref):
    """Three copies of one scene in one arena: a query tagged with a world sees only that world's
    fixtures (offset indices), an untagged query sees the three copies."""
    from oracle.bindings import RefScene
    ref = RefScene("pyramid", 10, 0)
    ref.step(30)
    nf = ref.fixture_count
    A = arena_from_scene(ref, max_contacts=4096 * 3, num_worlds=3, copies=3)
    A.find_new_contacts()
    boxes = np.array([[-3, 0, 3, 4], [-100, -100, 100, 100]], np.float32)
    rc, rf = ref.query_aabb(boxes, nf)
    for w in range(3):
        gc, gf = A.query_aabb(boxes, 3 * nf, world=np.full(len(boxes), w, np.int32))
        assert np.array_equal(gc, rc)
        for q in range(len(boxes)):
            assert np.array_equal(gf[q, :gc[q]], rf[q, :rc[q]] + w * nf)
    gc, gf = A.query_aabb(boxes, 3 * nf)
    assert np.array_equal(gc, 3 * rc)
    rays = random_rays(np.random.default_rng(3), scene_bounds(ref), 500)
    rfix, rfrac, _, _ = ref.ray_cast_closest(rays)
    gfix, gfrac, _ = A.ray_cast_closest(rays, world=np.full(len(rays), 2, np.int32))
    hit = rfix >= 0
    assert np.array_equal(hit, gfix >= 0) and np.array_equal(rfrac[hit], gfrac[hit])
    assert np.array_equal(gfix[hit], rfix[hit] + 2 * nf) or (gfix[hit] >= 2 * nf).all()
    A.close()

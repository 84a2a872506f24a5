"""Batched independent worlds (BASELINE config 4): an arena holding `copies` worlds must give
every world EXACTLY the floats that world gets when it is stepped alone — bodies of different worlds never
interact, island roots, colour priorities (world-local pair keys) and serial-bucket order are functions of a
world's own indices, so the result is bit-identical, world by world, in both solver modes.

Also: the device category / mask / group filter on a scene that uses all three group signs."""
import numpy as np
import pytest

import util
from box2d_optimized_b200 import capi, GpuScene, Arena, arena_from_scene

pytestmark = pytest.mark.gpu

STEPS = 200
COPIES = 8


def _host_scene(name, size, seed, prestep):
    s = GpuScene(name, size, seed)
    if prestep:
        s.step(prestep)
    return s


def _run(scene, copies, mode, steps):
    A = arena_from_scene(scene, copies=copies, num_worlds=copies)
    A.find_new_contacts()
    P = Arena.params(solver_mode=mode)
    st = capi.StepStats()
    contacts = []
    for _ in range(steps):
        A.step(P, st)
        contacts.append(st.num_contacts)
    d = A.download_bodies(what=("pos", "vel", "flags", "force"))
    A.close()
    return d, contacts


@pytest.mark.parametrize("mode", [capi.SOLVER_COLOURED, capi.SOLVER_SEQUENTIAL])
@pytest.mark.parametrize("name,size,seed,prestep", [("tumbler", 150, 0, 170), ("pyramid", 12, 0, 0),
                                                    ("mixed", 700, 12345, 0)])
def test_world_k_of_a_batched_arena_equals_that_world_alone(name, size, seed, prestep, mode):
    scene = _host_scene(name, size, seed, prestep)
    nb = scene.body_count
    one, c1 = _run(scene, 1, mode, STEPS)
    many, c8 = _run(scene, COPIES, mode, STEPS)
    assert [COPIES * c for c in c1] == c8, "contact counts per step"
    for k in range(COPIES):
        sl = slice(k * nb, (k + 1) * nb)
        for what in ("pos", "vel", "force"):
            a, b = one[what].view(np.uint32), many[what][sl].view(np.uint32)
            assert np.array_equal(a, b), f"world {k}: {what} differs (max |d| = {np.abs(one[what] - many[what][sl]).max():g})"
        assert np.array_equal(one["flags"], many["flags"][sl]), f"world {k}: flags differ"
    assert np.isfinite(one["pos"]).all()


def test_different_worlds_in_one_arena_equal_their_single_world_runs():
    """an arena of DIFFERENT worlds (tumbler spawn variants, what bench.py's config 4 steps): world k equals
    the single-world run of its own variant, bit for bit"""
    variants = [_host_scene("tumbler", 100, seed, 120) for seed in (1, 2, 3)]
    nb = variants[0].body_count
    many, _ = _run(variants, 6, capi.SOLVER_COLOURED, 150)
    for v, scene in enumerate(variants):
        one, _ = _run(scene, 1, capi.SOLVER_COLOURED, 150)
        for k in (v, v + 3):
            sl = slice(k * nb, (k + 1) * nb)
            assert np.array_equal(one["pos"].view(np.uint32), many["pos"][sl].view(np.uint32)), f"world {k}"
            assert np.array_equal(one["vel"].view(np.uint32), many["vel"][sl].view(np.uint32)), f"world {k}"
    # and the variants really are different worlds
    a, b = many["pos"][0:nb], many["pos"][nb:2 * nb]
    assert np.abs(a - b).max() > 0.1


# ---------------------------------------------------------------------------------------------
# device contact filter: categoryBits / maskBits / groupIndex (b2_world_callbacks.cpp:28-40)
# ---------------------------------------------------------------------------------------------
def _should_collide(fa, fb):
    """the reference's rule, restated for the test's own bookkeeping"""
    if fa[2] == fb[2] and fa[2] != 0:
        return fa[2] > 0
    return (fa[1] & fb[0]) != 0 and (fa[0] & fb[1]) != 0


@pytest.mark.parametrize("steps", [0, 40, 150, 400])
def test_filter_scene_pair_set_equals_the_reference(require_ref, steps):
    from oracle.bindings import RefScene
    s = RefScene("filters", 120, 99)
    s.step(steps)
    A = arena_from_scene(s)
    A.find_new_contacts()
    cg, cr = A.download_contacts(), s.contacts()
    got, ref = util.pair_set(cg["fix_a"], cg["fix_b"]), util.pair_set(cr["fix_a"], cr["fix_b"])
    assert got == ref
    A.close()
    # the scene exercises every branch of the rule: some overlapping pairs are rejected by each of them
    fx = s.fixtures()
    filt = fx["filter"]
    groups = set(int(g) for g in filt[:, 2])
    assert {1, -2, 0} <= groups
    for a, b in got:
        assert _should_collide(filt[a], filt[b])
    if steps >= 150:
        kinds = set()
        for a, b in got:
            if filt[a][2] == filt[b][2] == 1:
                kinds.add("same positive group")
            if filt[a][2] == 0 and filt[b][2] == 0:
                kinds.add("category/mask")
        assert kinds == {"same positive group", "category/mask"}
        # rejected overlaps exist too: AABB-overlapping pairs that the filter refused
        aabb = s.aabbs()
        body = fx["body"]
        n = len(aabb)
        rejected = 0
        for a in range(n):
            ov = np.nonzero((aabb[a, 0] <= aabb[:, 2]) & (aabb[:, 0] <= aabb[a, 2]) & (aabb[a, 1] <= aabb[:, 3]) &
                            (aabb[:, 1] <= aabb[a, 3]))[0]
            for b in ov:
                if b > a and body[a] != body[b] and not _should_collide(filt[a], filt[b]):
                    rejected += 1
        assert rejected > 10


def test_filter_scene_runs_free_like_the_reference(require_ref):
    """400 free-running steps in the production mode: same contact count, same resting heights per filter
    class (the -2 group piles up inside itself, ghosts fall through everything but the floor, only the blue
    ones stay on the shelf)"""
    from oracle.bindings import RefScene
    r, g = RefScene("filters", 120, 99), GpuScene("filters", 120, 99)
    r.step(400)
    g.step(400)
    rb, gb = r.bodies(), g.bodies()
    assert abs(r.contact_count - g.contact_count) <= 6
    on_shelf_r = set(np.nonzero(rb[1:, 5] > 6.0)[0].tolist())
    on_shelf_g = set(np.nonzero(gb[1:, 5] > 6.0)[0].tolist())
    assert len(on_shelf_r) > 5
    assert len(on_shelf_r ^ on_shelf_g) <= 2
    for cls in range(5):
        idx = 1 + np.arange(cls, 120, 5)
        assert abs(float(np.mean(rb[idx, 5])) - float(np.mean(gb[idx, 5]))) < 0.15, f"filter class {cls}"


def test_indexed_uploads_equal_the_contiguous_ones():
    """b2g_upload_{bodies,fixtures,shapes}_indexed (the scatter uploads a b2WorldBatch uses): an arena filled row by
    row in a shuffled order, in three calls, steps bit for bit like the arena filled with the contiguous calls."""
    from box2d_optimized_b200.arena import _scene_upload_arrays
    scene = _host_scene("mixed", 700, 12345, 0)
    ref, _ = _run(scene, 1, capi.SOLVER_COLOURED, 120)
    a = _scene_upload_arrays(scene)
    nb, nf, nq = a["nb"], a["nf"], a["nq"]
    A = Arena(nb, nf, nq, max(1024, 8 * nb), num_worlds=1, max_joints=1)
    rng = np.random.default_rng(7)
    pb, pf, pq = rng.permutation(nb), rng.permutation(nf), rng.permutation(nq)
    fx = a["fx"]
    A.upload_shapes_indexed(pq, fx["quads"].reshape(-1, 4)[pq])
    A.upload_bodies_indexed(pb, a["pos"][pb], a["vel"][pb], a["xf"][pb], a["massq"][pb], a["center"][pb], a["force"][pb],
                            a["flags"][pb], np.zeros(nb, np.int32))
    A.upload_fixtures_indexed(pf, fx["body"][pf], fx["shape_off"][pf], a["tf"][pf], a["filt"][pf], fx["material"][pf])
    A.find_new_contacts()
    P = Arena.params(solver_mode=capi.SOLVER_COLOURED)
    st = capi.StepStats()
    for _ in range(120):
        A.step(P, st)
    d = A.download_bodies(what=("pos", "vel", "flags", "force"))
    A.close()
    for k in ("pos", "vel", "flags", "force"):
        assert np.array_equal(np.asarray(d[k]).view(np.uint32), np.asarray(ref[k]).view(np.uint32)), k

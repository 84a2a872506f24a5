"""World-level single-step parity (teacher forced).

Every step starts from the reference's EXACT state — body sweeps/velocities/transforms/sleep
timers, its contact list with A/B order, manifolds and warm-start impulses, its previous 1/dt —
mirrored into the arena through the C-ABI (b2g_upload_bodies / b2g_upload_contacts).  The GPU then
runs ONE complete b2g_step in the sequential mode, visiting the constraints in the reference's own
island order (PostSolve tap), and everything the step produces is compared with what the
reference's b2World::Step produced from the same state:

  bodies    c, a, v, w        <= 1e-4 relative (north_star gate for the sequential mode);
                              observed: 0 — bit-exact whenever the host libm is the glibc variant
                              whose sinf/cosf the device restates (then asserted exactly)
  awake flags, sleep timers   equal / 1e-6
  contacts  pair set equal (bit-exact integer gate); touching flags, feature ids equal;
            accumulated impulses <= 1e-4 relative

This exercises narrowphase -> islands -> solver -> integration -> sleep -> broadphase together,
with no tolerance for ordering effects because there are none left."""
import numpy as np
import pytest

import util
from box2d_optimized_b200 import capi, Arena, arena_from_scene, body_flags

pytestmark = pytest.mark.gpu


def mirror_reference_state(A, ref, params, inv):
    b = ref.bodies()
    st = ref.sleep_times()
    nb = len(b)
    z = np.zeros(nb, np.float32)
    pos = np.stack([b[:, 4], b[:, 5], b[:, 6], z], 1)
    vel = np.stack([b[:, 7], b[:, 8], b[:, 9], z], 1)
    force = np.stack([z, z, z, st], 1)
    flags = np.array([body_flags(int(t), awake=bool(a), allow_sleep=bool(s)) for t, a, s in
                      zip(b[:, 11], b[:, 10], params[:, 7])], np.uint32)
    mass = np.stack([inv[:, 0], inv[:, 1], params[:, 0], params[:, 6]], 1)
    A.upload_bodies(0, pos=pos, vel=vel, xf=b[:, 0:4], mass=mass, force=force, flags=flags)
    c = ref.contacts()
    fl = np.where(c["flags"] & 1, capi.CONTACT_TOUCHING, 0) | np.where(c["flags"] & 2, capi.CONTACT_ENABLED, 0)
    A.upload_contacts(c["fix_a"], c["fix_b"], fl.astype(np.uint32), c["manifold"], c["material"])
    A.set_inv_dt0(ref.inv_dt0())
    j = ref.joints()
    if len(j["bodies"]):
        A.upload_joints(j["bodies"], j["anchors"], j["params"], state=ref.joint_state())
    return c


def device_rotations_match_libm():
    """True when this host's libm is the glibc variant rot_set restates (x86-64 with FMA)."""
    from test_rotation_parity import libm_sincos
    angles = np.random.default_rng(1).uniform(-100, 100, 20000).astype(np.float32)
    got = np.empty((len(angles), 2), np.float32)
    capi.check(capi.load_cuda().b2g_rotations(0, len(angles), capi.fp(angles), capi.fp(got)))
    return bool((got.view(np.uint32) == libm_sincos(angles).view(np.uint32)).all())


def rel(a, b):
    return float(np.max(np.abs(a - b) / np.maximum(1.0, np.abs(b)))) if a.size else 0.0


@pytest.mark.parametrize("name,size,seed,steps", [("pyramid", 12, 0, 150), ("mixed", 700, 12345, 260),
                                                   ("falling_squares", 200, 7, 160), ("tumbler", 80, 3, 200),
                                                   ("pendulum_limit", 6, 0, 120), ("pendulum_motor", 6, 0, 120),
                                                   # sensor zones, a kinematic sensor paddle and probes: touching of sensor
                                                   # contacts = b2TestOverlap (GJK), compared contact by contact every step
                                                   ("sensors", 300, 0, 240),
                                                   # several joints in one island: the device walks them in descending
                                                   # index order, which is the reference's DFS order for these chains
                                                   ("chain", 14, 0, 300), ("chain_collide", 14, 0, 200),
                                                   # distance joints: rods, springs, limited ropes, cross-linked
                                                   ("springs", 8, 0, 240),
                                                   # weld joints: rigid and soft cantilevers, a welded compound falling on them
                                                   ("welds", 6, 0, 240),
                                                   # prismatic joints: motorised piston, lifts between limits,
                                                   # free and locked rails, carriage on carriage
                                                   ("sliders", 6, 0, 240),
                                                   # wheel joints: sprung motorised cars over ramps and crates
                                                   ("cars", 4, 0, 300),
                                                   # the same cars with crates that fall asleep and are woken by
                                                   # the impact: ring-order corner, tolerance gate only (see below)
                                                   ("cars", 4, 1, 300),
                                                   # friction joints (braked boxes) and motor joints (driven platforms)
                                                   ("drags", 6, 0, 240),
                                                   # mouse joints: bodies dragged to targets through loose boxes
                                                   ("mice", 8, 0, 240)])
def test_every_step_from_the_reference_state(require_ref, name, size, seed, steps):
    from oracle.bindings import RefScene
    ref = RefScene(name, size, seed)
    # the reference creates its first contacts inside the first Step; the tumbler spawns one body per step
    ref.step(size + 2 if name == "tumbler" else 1)
    A = arena_from_scene(ref, max_contacts=max(4096, 16 * ref.body_count))
    params = ref.body_params()
    inv = ref.body_inv()
    P = Arena.params(solver_mode=capi.SOLVER_SEQUENTIAL)
    stats = capi.StepStats()
    worst = dict(pos=0.0, vel=0.0, imp=0.0, sleep=0.0, joint=0.0)
    nj = len(ref.joints()["bodies"])   # joints are visited in the reference's island DFS order (oracle tap)
    solved_total = 0
    for k in range(steps):
        before = mirror_reference_state(A, ref, params, inv)
        if nj:
            A.set_sequential_joint_order(ref.next_step_joint_order())
        fa, fb = ref.step_recording_order()      # the reference advances one Step
        A.set_sequential_order(fa, fb)
        A.step(P, stats)
        solved_total += len(fa)
        assert stats.num_constraints == len(fa), f"step {k}: {stats.num_constraints} constraints vs {len(fa)} solved by the reference"
        rb = ref.bodies()
        gb = A.download_bodies(what=("pos", "vel", "flags", "force"))
        worst["pos"] = max(worst["pos"], rel(gb["pos"][:, :3], rb[:, 4:7]))
        worst["vel"] = max(worst["vel"], rel(gb["vel"][:, :3], rb[:, 7:10]))
        awake_g = (gb["flags"] & capi.BODY_AWAKE) != 0
        assert np.array_equal(awake_g, rb[:, 10] != 0), f"step {k}: awake flags differ"
        worst["sleep"] = max(worst["sleep"], float(np.abs(gb["force"][:, 3] - ref.sleep_times()).max()))
        if nj:
            worst["joint"] = max(worst["joint"], rel(A.download_joints(nj), ref.joint_state()))
        # contacts after the step
        cg, cr = A.download_contacts(), ref.contacts()
        assert util.pair_set(cg["fix_a"], cg["fix_b"]) == util.pair_set(cr["fix_a"], cr["fix_b"]), f"step {k}: pair sets differ"
        # contacts that existed before the step keep the reference's A/B order: compare them in full
        key_g = {(a, b): i for i, (a, b) in enumerate(zip(cg["fix_a"].tolist(), cg["fix_b"].tolist()))}
        idx_r, idx_g = [], []
        old = set(zip(before["fix_a"].tolist(), before["fix_b"].tolist()))
        for i, (a, b) in enumerate(zip(cr["fix_a"].tolist(), cr["fix_b"].tolist())):
            if (a, b) in old and (a, b) in key_g:
                idx_r.append(i)
                idx_g.append(key_g[(a, b)])
        idx_r, idx_g = np.array(idx_r, int), np.array(idx_g, int)
        if len(idx_r):
            tg = (cg["flags"][idx_g] & capi.CONTACT_TOUCHING) != 0
            tr = (cr["flags"][idx_r] & 1) != 0
            assert np.array_equal(tg, tr), f"step {k}: touching flags differ"
            mg, mr = cg["manifold"][idx_g], cr["manifold"][idx_r]
            util.compare_manifolds(mg[tr], mr[tr], rel=1e-5)
            cnt = util.man_count(mr)
            one = tr & (cnt > 0)
            two = tr & (cnt > 1)
            worst["imp"] = max(worst["imp"], rel(mg[one][:, [6, 7]], mr[one][:, [6, 7]]),
                               rel(mg[two][:, [10, 11]], mr[two][:, [10, 11]]))
    print(f"{name}: {steps} teacher-forced steps, {solved_total} constraint solves; worst relative error "
          f"pos {worst['pos']:.3g} vel {worst['vel']:.3g} impulses {worst['imp']:.3g} sleepTime {worst['sleep']:.3g} "
          f"joint impulses {worst['joint']:.3g}")
    assert solved_total > 1000 or nj
    assert max(worst["pos"], worst["vel"], worst["imp"], worst["joint"]) <= 1e-4 and worst["sleep"] <= 1e-6
    # A body woken in the middle of b2ContactManager::Collide: the reference re-evaluates its remaining
    # contacts only if they come LATER in the world's contact ring than the waking contact; the device
    # narrowphase takes the awake flags of the start of the pass, so such a contact keeps last step's
    # manifold for one step (1e-6-level difference in its local points).  Within the gates, not bit-exact.
    ring_order_sensitive = (name, seed) == ("cars", 1)
    if device_rotations_match_libm() and not ring_order_sensitive:
        # same libm algorithm on both sides -> every float of the step is reproduced bit for bit
        assert worst == dict(pos=0.0, vel=0.0, imp=0.0, sleep=0.0, joint=0.0), worst
    A.close()

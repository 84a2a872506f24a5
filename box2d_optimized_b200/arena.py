"""numpy-level wrapper of the device arena (the C-ABI of include/b2cuda.h).

This is the host-side mirror used from Python: scenes are described as plain SoA numpy arrays
(the same arrays a maintainer's binding inside the reference would hand over),
uploaded once, and stepped on the GPU.  No oracle code is reachable from here.
"""
import ctypes as C

import numpy as np

from . import capi
from .capi import f32, i32, u32


def body_flags(btype, awake=True, allow_sleep=True, enabled=True, fixed_rotation=False):
    f = (int(btype) << capi.BODY_TYPE_SHIFT)
    if awake and btype != capi.STATIC:
        f |= capi.BODY_AWAKE
    if allow_sleep:
        f |= capi.BODY_AUTOSLEEP
    if enabled:
        f |= capi.BODY_ENABLED
    if fixed_rotation:
        f |= capi.BODY_FIXED_ROTATION
    return f


class Arena:
    def __init__(self, max_bodies, max_fixtures, max_shape_quads, max_contacts, num_worlds=1, max_joints=0, device=0):
        self.lib = capi.load_cuda()
        d = capi.ArenaDef(device, num_worlds, max_bodies, max_fixtures, max_shape_quads, max_contacts, max_joints, 0)
        self.h = C.c_void_p()
        capi.check(self.lib.b2g_arena_create(C.byref(d), C.byref(self.h)), "b2g_arena_create")
        self.num_bodies = 0
        self.num_fixtures = 0
        self.device = device
        self._keep = []

    def close(self):
        if self.h:
            self.lib.b2g_arena_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- uploads ---------------------------------------------------------------------------
    def upload_bodies(self, first=0, pos=None, vel=None, xf=None, mass=None, center=None, force=None, flags=None,
                      world=None):
        arrs = dict(pos=pos, vel=vel, xf=xf, mass=mass, center=center, force=force)
        n = None
        a = capi.BodyArrays()
        keep = []
        for k, v in arrs.items():
            if v is not None:
                v = f32(v).reshape(-1, 4)
                n = len(v)
                keep.append(v)
                setattr(a, k, capi.fp(v))
        if flags is not None:
            flags = u32(flags)
            n = len(flags)
            keep.append(flags)
            a.flags = capi.up(flags)
        if world is not None:
            world = i32(world)
            n = len(world)
            keep.append(world)
            a.world = capi.ip(world)
        capi.check(self.lib.b2g_upload_bodies(self.h, first, n, C.byref(a)), "b2g_upload_bodies")
        self.num_bodies = max(self.num_bodies, first + n)

    def upload_fixtures(self, first=0, body=None, shape_off=None, type_flags=None, filter=None, material=None):
        a = capi.FixtureArrays()
        keep = []
        n = None
        if body is not None:
            body = i32(body); n = len(body); keep.append(body); a.body = capi.ip(body)
        if shape_off is not None:
            shape_off = i32(shape_off); n = len(shape_off); keep.append(shape_off); a.shape_off = capi.ip(shape_off)
        if type_flags is not None:
            type_flags = u32(type_flags); n = len(type_flags); keep.append(type_flags)
            a.type_flags = capi.up(type_flags)
        if filter is not None:
            filter = u32(filter).reshape(-1, 2); n = len(filter); keep.append(filter); a.filter = capi.up(filter)
        if material is not None:
            material = f32(material).reshape(-1, 4); n = len(material); keep.append(material)
            a.material = capi.fp(material)
        capi.check(self.lib.b2g_upload_fixtures(self.h, first, n, C.byref(a)), "b2g_upload_fixtures")
        self.num_fixtures = max(self.num_fixtures, first + n)

    def upload_bodies_indexed(self, index, pos, vel, xf, mass, center, force, flags, world):
        """row i of the arrays goes to body index[i] (b2g_upload_bodies_indexed)"""
        index = i32(index)
        a = capi.BodyArrays()
        keep = [index]
        for k, v in dict(pos=pos, vel=vel, xf=xf, mass=mass, center=center, force=force).items():
            v = f32(v).reshape(-1, 4)
            keep.append(v)
            setattr(a, k, capi.fp(v))
        flags, world = u32(flags), i32(world)
        keep += [flags, world]
        a.flags, a.world = capi.up(flags), capi.ip(world)
        capi.check(self.lib.b2g_upload_bodies_indexed(self.h, len(index), capi.ip(index), C.byref(a)), "b2g_upload_bodies_indexed")
        if len(index):
            self.num_bodies = max(self.num_bodies, int(index.max()) + 1)

    def upload_fixtures_indexed(self, index, body, shape_off, type_flags, filter, material):
        index, body, shape_off = i32(index), i32(body), i32(shape_off)
        type_flags, filter, material = u32(type_flags), u32(filter).reshape(-1, 2), f32(material).reshape(-1, 4)
        a = capi.FixtureArrays()
        a.body, a.shape_off, a.type_flags = capi.ip(body), capi.ip(shape_off), capi.up(type_flags)
        a.filter, a.material = capi.up(filter), capi.fp(material)
        capi.check(self.lib.b2g_upload_fixtures_indexed(self.h, len(index), capi.ip(index), C.byref(a)), "b2g_upload_fixtures_indexed")
        if len(index):
            self.num_fixtures = max(self.num_fixtures, int(index.max()) + 1)

    def upload_shapes_indexed(self, index, quads):
        index, quads = i32(index), f32(quads).reshape(-1, 4)
        capi.check(self.lib.b2g_upload_shapes_indexed(self.h, len(index), capi.ip(index), capi.fp(quads)), "b2g_upload_shapes_indexed")

    def upload_shapes(self, quads, first=0):
        quads = f32(quads).reshape(-1, 4)
        capi.check(self.lib.b2g_upload_shapes(self.h, first, len(quads), capi.fp(quads)), "b2g_upload_shapes")

    def upload_joints(self, bodies, anchors, params, first=0, state=None):
        bodies = i32(bodies).reshape(-1, 2)
        anchors = f32(anchors).reshape(-1, 4)
        params = f32(params).reshape(-1, 12)
        state = None if state is None else f32(state).reshape(-1, 5)
        a = capi.JointArrays(capi.ip(bodies), capi.fp(anchors), capi.fp(params), None if state is None else capi.fp(state))
        capi.check(self.lib.b2g_upload_joints(self.h, first, len(bodies), C.byref(a)), "b2g_upload_joints")

    # ---- spatial queries on the broadphase tree (include/b2cuda.h "Spatial queries") ----
    @staticmethod
    def _ptr(a):
        return None if a is None else C.c_void_p(a.ctypes.data)

    def query_aabb(self, aabbs, cap, world=None):
        """aabbs [n,4] -> (counts [n], fixtures [n,cap] sorted ascending, -1 padded)"""
        aabbs = f32(aabbs).reshape(-1, 4)
        n = len(aabbs)
        world = None if world is None else i32(world)
        counts = np.zeros(n, np.int32)
        fixtures = np.full((n, max(cap, 1)), -1, np.int32)
        capi.check(self.lib.b2g_query_aabb(self.h, n, self._ptr(aabbs), self._ptr(world), cap, self._ptr(counts),
                                           self._ptr(fixtures), 0), "b2g_query_aabb")
        big = np.iinfo(np.int32).max
        k = np.arange(fixtures.shape[1])[None, :] < np.minimum(counts, cap)[:, None]
        fixtures = np.where(k, fixtures, big)
        fixtures.sort(axis=1)
        fixtures[fixtures == big] = -1
        return counts, fixtures

    def ray_cast_closest(self, rays, max_fraction=None, world=None, category_mask=0xFFFF):
        """rays [n,4] = p1, p2 -> (fixture [n] or -1, fraction [n], normal [n,2])"""
        rays = f32(rays).reshape(-1, 4)
        n = len(rays)
        mf = None if max_fraction is None else f32(max_fraction)
        world = None if world is None else i32(world)
        fixture = np.full(n, -1, np.int32)
        fraction = np.zeros(n, np.float32)
        normal = np.zeros((n, 2), np.float32)
        capi.check(self.lib.b2g_ray_cast_closest(self.h, n, self._ptr(rays), self._ptr(mf), self._ptr(world),
                                                 category_mask, self._ptr(fixture), self._ptr(fraction),
                                                 self._ptr(normal), 0), "b2g_ray_cast_closest")
        return fixture, fraction, normal

    def ray_cast_all(self, rays, cap, max_fraction=None, world=None, category_mask=0xFFFF):
        """every hit per ray, sorted by (fraction, fixture): (counts [n], fixture [n,cap], fraction [n,cap], normal [n,cap,2])"""
        rays = f32(rays).reshape(-1, 4)
        n = len(rays)
        mf = None if max_fraction is None else f32(max_fraction)
        world = None if world is None else i32(world)
        counts = np.zeros(n, np.int32)
        fixture = np.full((n, max(cap, 1)), -1, np.int32)
        fraction = np.full((n, max(cap, 1)), np.inf, np.float32)
        normal = np.zeros((n, max(cap, 1), 2), np.float32)
        capi.check(self.lib.b2g_ray_cast_all(self.h, n, self._ptr(rays), self._ptr(mf), self._ptr(world), category_mask,
                                             cap, self._ptr(counts), self._ptr(fixture), self._ptr(fraction),
                                             self._ptr(normal), 0), "b2g_ray_cast_all")
        valid = np.arange(fixture.shape[1])[None, :] < np.minimum(counts, cap)[:, None]
        fraction = np.where(valid, fraction, np.inf)
        fixture = np.where(valid, fixture, np.iinfo(np.int32).max)
        order = np.lexsort((fixture, fraction), axis=1)
        fixture = np.take_along_axis(fixture, order, 1)
        fraction = np.take_along_axis(fraction, order, 1)
        normal = np.take_along_axis(normal, order[:, :, None], 1)
        fixture[~np.take_along_axis(valid, order, 1)] = -1
        return counts, fixture, fraction, normal

    def set_sequential_joint_order(self, joints):
        joints = i32(joints)
        capi.check(self.lib.b2g_set_sequential_joint_order(self.h, len(joints), capi.ip(joints)),
                   "b2g_set_sequential_joint_order")

    def joint_colours(self, count):
        """colour of each joint in the production mode's joint colouring (b2g_debug_joint_colours)"""
        out = np.zeros(max(count, 1), np.int32)
        capi.check(self.lib.b2g_debug_joint_colours(self.h, capi.ip(out)), "b2g_debug_joint_colours")
        return out[:count]

    def download_joints(self, count, first=0):
        """accumulated joint impulses [count, 5] = impulse.xy, motor, lower, upper"""
        state = np.zeros((count, 5), np.float32)
        capi.check(self.lib.b2g_download_joints(self.h, first, count, capi.fp(state)), "b2g_download_joints")
        return state

    def upload_contacts(self, fix_a, fix_b, flags, manifold, material):
        fix_a, fix_b, flags = i32(fix_a), i32(fix_b), u32(flags)
        manifold, material = f32(manifold).reshape(-1, 16), f32(material).reshape(-1, 4)
        a = capi.ContactArrays(capi.ip(fix_a), capi.ip(fix_b), capi.up(flags), capi.fp(manifold), capi.fp(material), None)
        capi.check(self.lib.b2g_upload_contacts(self.h, len(fix_a), C.byref(a)), "b2g_upload_contacts")

    def set_sequential_order(self, fix_a, fix_b):
        fix_a, fix_b = i32(fix_a), i32(fix_b)
        capi.check(self.lib.b2g_set_sequential_order(self.h, len(fix_a), capi.ip(fix_a), capi.ip(fix_b)),
                   "b2g_set_sequential_order")

    def upload_forces(self, force_ptr, first, count):
        capi.check(self.lib.b2g_upload_forces(self.h, first, count, force_ptr), "b2g_upload_forces")

    # ---- stepping --------------------------------------------------------------------------
    @staticmethod
    def params(dt=1.0 / 60.0, vel_iters=8, pos_iters=3, gravity=(0.0, -10.0), warm_starting=True, allow_sleep=True,
               clear_forces=True, solver_mode=capi.SOLVER_COLOURED, record_events=False):
        return capi.StepParams(dt, vel_iters, pos_iters, gravity[0], gravity[1], int(warm_starting),
                               int(allow_sleep), int(clear_forces), int(solver_mode), int(record_events))

    def find_new_contacts(self):
        capi.check(self.lib.b2g_find_new_contacts(self.h), "b2g_find_new_contacts")

    def step(self, params, stats=None):
        capi.check(self.lib.b2g_step(self.h, C.byref(params), C.byref(stats) if stats is not None else None),
                   "b2g_step")

    def synchronize(self):
        capi.check(self.lib.b2g_synchronize(self.h), "b2g_synchronize")

    def stream(self):
        return self.lib.b2g_stream(self.h)

    def set_profiling(self, on=True):
        self.lib.b2g_set_profiling(self.h, int(on))

    def device_views(self):
        v = capi.DeviceViews()
        capi.check(self.lib.b2g_device_views(self.h, C.byref(v)), "b2g_device_views")
        return v

    def set_kernel_timing(self, on=True):
        self.lib.b2g_set_kernel_timing(self.h, int(on))

    def kernel_timing(self):
        """{class name: (total ms, launches, summed work items)} since timing was switched on"""
        out = {}
        for c in range(self.lib.b2g_kernel_class_count()):
            ms, n, u = C.c_double(), C.c_int64(), C.c_double()
            self.lib.b2g_get_kernel_timing(self.h, c, C.byref(ms), C.byref(n), C.byref(u))
            out[self.lib.b2g_kernel_class_name(c).decode()] = (ms.value, n.value, u.value)
        return out

    def set_inv_dt0(self, v):
        self.lib.b2g_set_inv_dt0(self.h, float(v))

    # ---- readback --------------------------------------------------------------------------
    def download_bodies(self, first=0, count=None, what=("pos", "vel", "xf", "flags", "force")):
        n = self.num_bodies - first if count is None else count
        out = {}
        a = capi.BodyArrays()
        for k in what:
            if k == "flags":
                out[k] = np.zeros(n, np.uint32)
                a.flags = capi.up(out[k])
            elif k == "world":
                out[k] = np.zeros(n, np.int32)
                a.world = capi.ip(out[k])
            else:
                out[k] = np.zeros((n, 4), np.float32)
                setattr(a, k, capi.fp(out[k]))
        if n:
            capi.check(self.lib.b2g_download_bodies(self.h, first, n, C.byref(a)), "b2g_download_bodies")
        return out

    def download_aabbs(self):
        out = np.zeros((self.num_fixtures, 4), np.float32)
        if len(out):
            capi.check(self.lib.b2g_download_fixture_aabbs(self.h, 0, len(out), capi.fp(out)), "aabbs")
        return out

    def contact_count(self):
        n = C.c_int32()
        capi.check(self.lib.b2g_contact_count(self.h, C.byref(n)), "b2g_contact_count")
        return n.value

    def download_contacts(self):
        n = self.contact_count()
        d = dict(fix_a=np.zeros(n, np.int32), fix_b=np.zeros(n, np.int32), flags=np.zeros(n, np.uint32),
                 manifold=np.zeros((n, 16), np.float32), material=np.zeros((n, 4), np.float32),
                 colour=np.zeros(n, np.int32))
        if n:
            a = capi.ContactArrays(capi.ip(d["fix_a"]), capi.ip(d["fix_b"]), capi.up(d["flags"]),
                                   capi.fp(d["manifold"]), capi.fp(d["material"]), capi.ip(d["colour"]))
            capi.check(self.lib.b2g_download_contacts(self.h, 0, n, C.byref(a)), "b2g_download_contacts")
        return d


def _scene_upload_arrays(scene):
    """numpy SoA of one scene in the layout the upload calls take"""
    b = scene.bodies()
    p = scene.body_params()
    fx = scene.fixtures()
    jn = scene.joints()
    nb = len(b)
    btype = b[:, 11].astype(np.int32)
    mass = p[:, 0]
    inertia_origin = p[:, 1]
    lc = p[:, 2:4]
    inv_mass = np.where(mass > 0, np.float32(1.0) / np.where(mass > 0, mass, 1).astype(np.float32), 0).astype(np.float32)
    # b2Body keeps I about the centre of mass: GetInertia() = I + m*|lc|^2 (b2_body.h:693-696)
    i_center = (inertia_origin - mass * (lc[:, 0] * lc[:, 0] + lc[:, 1] * lc[:, 1])).astype(np.float32)
    inv_i = np.where(i_center > 0, np.float32(1.0) / np.where(i_center > 0, i_center, 1).astype(np.float32), 0).astype(np.float32)
    z = np.zeros(nb, np.float32)
    tf = fx["type"].astype(np.uint32) | np.where(fx["sensor"] != 0, capi.FIX_SENSOR, 0).astype(np.uint32)
    filt = np.stack([(fx["filter"][:, 0].astype(np.uint32) & 0xffff) | ((fx["filter"][:, 1].astype(np.uint32) & 0xffff) << 16),
                     fx["filter"][:, 2].astype(np.int32).view(np.uint32)], 1)
    return dict(
        nb=nb, nf=len(fx["body"]), nq=len(fx["quads"]), nj=len(jn["bodies"]), fx=fx, jn=jn, tf=tf, filt=filt,
        pos=np.stack([b[:, 4], b[:, 5], b[:, 6], z], 1), vel=np.stack([b[:, 7], b[:, 8], b[:, 9], z], 1),
        xf=b[:, 0:4], massq=np.stack([inv_mass, inv_i, mass, p[:, 6]], 1),
        center=np.stack([lc[:, 0], lc[:, 1], p[:, 4], p[:, 5]], 1), force=np.zeros((nb, 4), np.float32),
        flags=np.array([body_flags(int(t), awake=bool(a), allow_sleep=bool(s)) for t, a, s in
                        zip(btype, b[:, 10], p[:, 7])], np.uint32),
        inv=(inv_mass, inv_i))


def arena_from_scene(scene, max_contacts=None, device=0, num_worlds=1, copies=1):
    """Builds an arena holding `copies` worlds' bodies and fixtures, each in its own world when
    num_worlds > 1.  `scene` is a (reference or drop-in) scene, or a LIST of scenes with equal body /
    fixture / shape / joint counts: world k is then scenes[k % len] (batched worlds that are different
    worlds, not clones).  Body state is taken bit-for-bit from the scene, so the arena starts exactly
    where that scene currently is."""
    scenes = list(scene) if isinstance(scene, (list, tuple)) else [scene]
    arrs = [_scene_upload_arrays(s) for s in scenes]
    a0 = arrs[0]
    nb, nf, nq, nj = a0["nb"], a0["nf"], a0["nq"], a0["nj"]
    for a in arrs[1:]:
        if (a["nb"], a["nf"], a["nq"], a["nj"]) != (nb, nf, nq, nj):
            raise ValueError("arena_from_scene: the scenes of one batched arena must have equal counts")
    if max_contacts is None:
        max_contacts = max(1024, 8 * nb * copies)
    A = Arena(nb * copies, nf * copies, nq * copies, max_contacts, num_worlds=num_worlds, device=device,
              max_joints=max(nj * copies, 1))
    for k in range(copies):
        a = arrs[k % len(arrs)]
        fx, jn = a["fx"], a["jn"]
        A.upload_bodies(k * nb, pos=a["pos"], vel=a["vel"], xf=a["xf"], mass=a["massq"], center=a["center"],
                        force=a["force"], flags=a["flags"], world=np.full(nb, k if num_worlds > 1 else 0, np.int32))
        A.upload_shapes(fx["quads"], first=k * nq)
        A.upload_fixtures(k * nf, body=fx["body"] + k * nb, shape_off=fx["shape_off"] + k * nq, type_flags=a["tf"],
                          filter=a["filt"], material=fx["material"])
        if nj:
            A.upload_joints(jn["bodies"] + k * nb, jn["anchors"], jn["params"], first=k * nj)
    A.scene_inv = a0["inv"]
    return A

"""Python views of the two scene shims (same C functions, prefixes b2gpu_ / b2ref_).

`GpuScene` drives this repo's drop-in C++ API (and through it the CUDA step).  The view of the
compiled reference (`RefScene`) is test/baseline infrastructure and lives in oracle/bindings.py.
"""
import ctypes as C

import numpy as np

from . import capi


class _Scene:
    prefix = None

    def __init__(self, lib, name, size=0, seed=0):
        self.lib = lib
        self._f = lambda n: getattr(lib, self.prefix + n)
        self.h = self._f("scene_create")(name.encode(), int(size), int(seed))
        if not self.h:
            raise ValueError(f"unknown scene {name!r}")
        self.name = name

    def close(self):
        if self.h:
            self._f("scene_destroy")(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def step(self, n=1):
        self._f("scene_step")(self.h, int(n))

    def time_steps(self, n):
        """wall-clock ms for n steps"""
        return float(self._f("scene_time_steps")(self.h, int(n)))

    def set_iterations(self, vi, pi):
        self._f("scene_set_iterations")(self.h, int(vi), int(pi))

    def set_flags(self, allow_sleep=True, warm_starting=True):
        self._f("scene_set_flags")(self.h, int(allow_sleep), int(warm_starting))

    @property
    def body_count(self):
        return self._f("scene_body_count")(self.h)

    @property
    def fixture_count(self):
        return self._f("scene_fixture_count")(self.h)

    @property
    def contact_count(self):
        return self._f("scene_contact_count")(self.h)

    def bodies(self):
        """[n,12] = xf.p.xy, xf.q.s, xf.q.c, c.xy, a, v.xy, w, awake, type"""
        out = np.zeros((self.body_count, 12), np.float32)
        if len(out):
            self._f("scene_get_bodies")(self.h, capi.fp(out))
        return out

    def body_params(self):
        """[n,8] = mass, inertia(origin), localCenter.xy, linDamp, angDamp, gravityScale, sleepingAllowed"""
        out = np.zeros((self.body_count, 8), np.float32)
        if len(out):
            self._f("scene_get_body_params")(self.h, capi.fp(out))
        return out

    def fixtures(self):
        n = self.fixture_count
        nq = self._f("scene_shape_quad_total")(self.h)
        d = dict(body=np.zeros(n, np.int32), type=np.zeros(n, np.int32), shape_off=np.zeros(n, np.int32),
                 filter=np.zeros((n, 3), np.int32), material=np.zeros((n, 4), np.float32),
                 sensor=np.zeros(n, np.int32), quads=np.zeros((max(nq, 1), 4), np.float32))
        if n:
            self._f("scene_get_fixtures")(self.h, capi.ip(d["body"]), capi.ip(d["type"]), capi.ip(d["shape_off"]),
                                          capi.ip(d["filter"]), capi.fp(d["material"]), capi.ip(d["sensor"]),
                                          capi.fp(d["quads"]))
        return d

    def aabbs(self):
        out = np.zeros((self.fixture_count, 4), np.float32)
        if len(out):
            self._f("scene_get_aabbs")(self.h, capi.fp(out))
        return out

    # ---- spatial queries through the public C++ API (b2World::QueryAABB / RayCast) ----
    def query_aabb(self, aabbs, cap):
        aabbs = np.ascontiguousarray(aabbs, np.float32).reshape(-1, 4)
        counts = np.zeros(len(aabbs), np.int32)
        fixtures = np.full((len(aabbs), max(cap, 1)), -1, np.int32)
        self._f("scene_query_aabb")(self.h, len(aabbs), capi.fp(aabbs), cap, capi.ip(counts), capi.ip(fixtures))
        return counts, fixtures

    def ray_cast_closest(self, rays):
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 4)
        n = len(rays)
        fixture = np.zeros(n, np.int32)
        fraction = np.zeros(n, np.float32)
        normal = np.zeros((n, 2), np.float32)
        point = np.zeros((n, 2), np.float32)
        self._f("scene_ray_cast_closest")(self.h, n, capi.fp(rays), capi.ip(fixture), capi.fp(fraction), capi.fp(normal),
                                          capi.fp(point))
        return fixture, fraction, normal, point

    def ray_cast_all(self, ray, cap):
        ray = np.ascontiguousarray(ray, np.float32).reshape(4)
        fixture = np.full(max(cap, 1), -1, np.int32)
        fraction = np.zeros(max(cap, 1), np.float32)
        normal = np.zeros((max(cap, 1), 2), np.float32)
        n = self._f("scene_ray_cast_all")(self.h, capi.fp(ray), cap, capi.ip(fixture), capi.fp(fraction), capi.fp(normal))
        return n, fixture[:min(n, cap)], fraction[:min(n, cap)], normal[:min(n, cap)]

    def time_ray_casts(self, rays):
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 4)
        return self._f("scene_time_ray_casts")(self.h, len(rays), capi.fp(rays))

    def joints(self):
        """revolute joints: bodies [n,2], anchors [n,4], params [n,8] (include/b2cuda.h b2gJointArrays)"""
        cap = max(self._f("scene_joint_count")(self.h), 1)
        bodies = np.zeros((cap, 2), np.int32)
        anchors = np.zeros((cap, 4), np.float32)
        params = np.zeros((cap, 12), np.float32)
        n = self._f("scene_get_joints")(self.h, cap, capi.ip(bodies), capi.fp(anchors), capi.fp(params))
        return dict(bodies=bodies[:n], anchors=anchors[:n], params=params[:n])

    def contacts(self):
        cap = max(self.contact_count, 1)
        fa = np.zeros(cap, np.int32)
        fb = np.zeros(cap, np.int32)
        fl = np.zeros(cap, np.int32)
        man = np.zeros((cap, 16), np.float32)
        mat = np.zeros((cap, 4), np.float32)
        n = self._f("scene_get_contacts")(self.h, cap, capi.ip(fa), capi.ip(fb), capi.ip(fl), capi.fp(man),
                                          capi.fp(mat))
        return dict(fix_a=fa[:n], fix_b=fb[:n], flags=fl[:n], manifold=man[:n], material=mat[:n])


class GpuScene(_Scene):
    prefix = "b2gpu_"

    def __init__(self, name, size=0, seed=0, solver_mode=capi.SOLVER_COLOURED, capacity=None, device=None):
        lib = capi.load_gpu_scenes()
        if capacity is not None:
            lib.b2gpu_set_default_capacity(*[int(c) for c in capacity])
        if device is not None:
            lib.b2gpu_set_default_device(int(device))
        super().__init__(lib, name, size, seed)
        lib.b2gpu_scene_set_solver_mode(self.h, int(solver_mode))

    def set_profiling(self, on=True):
        self.lib.b2gpu_scene_set_profiling(self.h, int(on))

    def profile(self):
        out = np.zeros(5, np.float32)
        self.lib.b2gpu_scene_get_profile(self.h, capi.fp(out))
        return dict(zip(("step", "collide", "solve", "broadphase", "solveTOI"), out.tolist()))

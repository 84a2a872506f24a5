"""ctypes binding of the C-ABI (include/b2cuda.h) and of the two scene shims.

PyTorch is not involved here: arrays cross as numpy buffers / raw pointers.  The product
library is libb2cuda.so; nothing here knows about oracle/ (the checker's bindings live in
oracle/bindings.py and are loaded only by tests, smoke() and bench.py's baseline legs).
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)

OK = 0
BODY_AWAKE, BODY_AUTOSLEEP, BODY_BULLET, BODY_FIXED_ROTATION, BODY_ENABLED = 0x2, 0x4, 0x8, 0x10, 0x20
BODY_TYPE_SHIFT = 16
STATIC, KINEMATIC, DYNAMIC = 0, 1, 2
SHAPE_CIRCLE, SHAPE_EDGE, SHAPE_POLYGON = 0, 1, 2
FIX_SENSOR, FIX_DEAD = 0x100, 0x200
CONTACT_TOUCHING, CONTACT_ENABLED = 0x8, 0x10
SOLVER_COLOURED, SOLVER_SEQUENTIAL = 0, 1

f32p = C.POINTER(C.c_float)
i32p = C.POINTER(C.c_int32)
u32p = C.POINTER(C.c_uint32)
u8p = C.POINTER(C.c_uint8)


class ArenaDef(C.Structure):
    _fields_ = [("device", C.c_int32), ("num_worlds", C.c_int32), ("max_bodies", C.c_int32),
                ("max_fixtures", C.c_int32), ("max_shape_quads", C.c_int32), ("max_contacts", C.c_int32),
                ("max_joints", C.c_int32), ("reserved", C.c_int32)]


class BodyArrays(C.Structure):
    _fields_ = [("pos", f32p), ("vel", f32p), ("xf", f32p), ("mass", f32p), ("center", f32p), ("force", f32p),
                ("flags", u32p), ("world", i32p)]


class FixtureArrays(C.Structure):
    _fields_ = [("body", i32p), ("shape_off", i32p), ("type_flags", u32p), ("filter", u32p), ("material", f32p)]


class ContactArrays(C.Structure):
    _fields_ = [("fixture_a", i32p), ("fixture_b", i32p), ("flags", u32p), ("manifold", f32p), ("material", f32p),
                ("colour", i32p)]


class JointArrays(C.Structure):
    _fields_ = [("bodies", i32p), ("anchors", f32p), ("params", f32p), ("state", f32p)]


class DeviceViews(C.Structure):
    _fields_ = [("pos", C.c_void_p), ("vel", C.c_void_p), ("xf", C.c_void_p), ("force", C.c_void_p),
                ("flags", C.c_void_p), ("capacity", C.c_int32), ("device", C.c_int32)]


class StepParams(C.Structure):
    _fields_ = [("dt", C.c_float), ("velocity_iterations", C.c_int32), ("position_iterations", C.c_int32),
                ("gravity_x", C.c_float), ("gravity_y", C.c_float), ("warm_starting", C.c_int32),
                ("allow_sleep", C.c_int32), ("clear_forces", C.c_int32), ("solver_mode", C.c_int32),
                ("record_events", C.c_int32)]


class StepStats(C.Structure):
    _fields_ = [("num_bodies", C.c_int32), ("num_fixtures", C.c_int32), ("num_contacts", C.c_int32),
                ("num_touching", C.c_int32), ("num_constraints", C.c_int32), ("num_colours", C.c_int32),
                ("num_overflow", C.c_int32), ("num_awake", C.c_int32), ("num_pairs", C.c_int32),
                ("colour_rounds", C.c_int32), ("num_launches", C.c_int32), ("bp_max_visits", C.c_int32),
                ("ms_collide", C.c_float), ("ms_solve", C.c_float), ("ms_broadphase", C.c_float),
                ("ms_step", C.c_float), ("bp_mean_visits", C.c_float), ("bp_rebuilt", C.c_int32)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_ }


# every symbol include/b2cuda.h declares (checked by tests/test_capi_symbols.py)
CUDA_SYMBOLS = [
    "b2g_last_error", "b2g_device_count", "b2g_arena_create", "b2g_arena_destroy", "b2g_upload_bodies",
    "b2g_upload_fixtures", "b2g_upload_shapes", "b2g_upload_joints", "b2g_set_counts", "b2g_upload_forces",
    "b2g_step", "b2g_step_download", "b2g_step_collide", "b2g_step_solve", "b2g_find_new_contacts", "b2g_download_bodies",
    "b2g_download_body_state_async", "b2g_download_fixture_aabbs", "b2g_contact_count", "b2g_download_contacts",
    "b2g_upload_contact_overrides", "b2g_download_events", "b2g_synchronize", "b2g_stream", "b2g_set_profiling",
    "b2g_set_inv_dt0", "b2g_set_kernel_timing", "b2g_kernel_class_count", "b2g_kernel_class_name",
    "b2g_get_kernel_timing", "b2g_device_views", "b2g_upload_contacts", "b2g_set_sequential_order", "b2g_set_sequential_joint_order", "b2g_host_alloc", "b2g_host_free", "b2g_download_new_pairs", "b2g_set_pair_vetoes", "b2g_download_veto_seen", "b2g_query_aabb", "b2g_ray_cast_closest", "b2g_ray_cast_all", "b2g_download_joints", "b2g_rotations", "b2g_compute_aabbs", "b2g_collide_pairs",
    "b2g_find_pairs", "b2g_solve_sequential",
    "b2g_halo_set_lists", "b2g_halo_pack", "b2g_halo_recv_buffer", "b2g_halo_unpack", "b2g_debug_tile_state",
    "b2g_debug_joint_colours", "b2g_upload_bodies_indexed", "b2g_upload_fixtures_indexed", "b2g_upload_shapes_indexed",
]
DIST_SYMBOLS = ["b2g_dist_unique_id", "b2g_dist_init", "b2g_dist_exchange", "b2g_dist_destroy"]

_cuda = None
_gpu_scenes = None
_ref = None


class B2GError(RuntimeError):
    pass


def lib_path(name="libb2cuda.so"):
    return os.path.join(HERE, name)


def load_cuda():
    """Loads the product library.  Fails loudly if it has not been built: there is no fallback."""
    global _cuda
    if _cuda is None:
        p = os.environ.get("B2G_CUDA_LIB") or lib_path()   # B2G_CUDA_LIB: a debug build (scripts/gpu_big_trace.py)
        if not os.path.exists(p):
            raise B2GError(f"{p} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(the CUDA extension is mandatory, there is no CPU fallback)")
        lib = C.CDLL(p)
        lib.b2g_last_error.restype = C.c_char_p
        lib.b2g_stream.restype = C.c_void_p
        lib.b2g_stream.argtypes = [C.c_void_p]
        for fn in ("b2g_arena_destroy", "b2g_synchronize"):
            getattr(lib, fn).argtypes = [C.c_void_p]
        lib.b2g_arena_create.argtypes = [C.POINTER(ArenaDef), C.POINTER(C.c_void_p)]
        lib.b2g_upload_bodies.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.POINTER(BodyArrays)]
        lib.b2g_download_bodies.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.POINTER(BodyArrays)]
        lib.b2g_upload_fixtures.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.POINTER(FixtureArrays)]
        lib.b2g_upload_shapes.argtypes = [C.c_void_p, C.c_int32, C.c_int32, f32p]
        lib.b2g_upload_joints.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.POINTER(JointArrays)]
        lib.b2g_set_counts.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32]
        lib.b2g_upload_forces.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]
        lib.b2g_step.argtypes = [C.c_void_p, C.POINTER(StepParams), C.POINTER(StepStats)]
        lib.b2g_step_download.argtypes = [C.c_void_p, C.POINTER(StepParams), C.POINTER(StepStats), C.c_int32, C.c_int32,
                                          C.c_void_p]
        lib.b2g_step_collide.argtypes = [C.c_void_p, C.POINTER(StepParams)]
        lib.b2g_step_solve.argtypes = [C.c_void_p, C.POINTER(StepParams), C.POINTER(StepStats)]
        lib.b2g_find_new_contacts.argtypes = [C.c_void_p]
        lib.b2g_download_body_state_async.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]
        lib.b2g_download_fixture_aabbs.argtypes = [C.c_void_p, C.c_int32, C.c_int32, f32p]
        lib.b2g_contact_count.argtypes = [C.c_void_p, i32p]
        lib.b2g_download_contacts.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.POINTER(ContactArrays)]
        lib.b2g_upload_contact_overrides.argtypes = [C.c_void_p, C.c_int32, C.c_int32, u32p, f32p]
        lib.b2g_download_events.argtypes = [C.c_void_p, i32p, i32p, i32p, i32p, C.c_int32]
        lib.b2g_set_profiling.argtypes = [C.c_void_p, C.c_int32]
        lib.b2g_set_inv_dt0.argtypes = [C.c_void_p, C.c_float]
        lib.b2g_device_views.argtypes = [C.c_void_p, C.POINTER(DeviceViews)]
        lib.b2g_halo_set_lists.argtypes = [C.c_void_p, C.c_int32, C.c_int32, i32p, C.c_int32, i32p]
        lib.b2g_halo_pack.argtypes = [C.c_void_p, C.c_int32, C.POINTER(C.c_void_p), C.POINTER(C.c_int64)]
        lib.b2g_halo_recv_buffer.argtypes = [C.c_void_p, C.c_int32, C.POINTER(C.c_void_p), C.POINTER(C.c_int64)]
        lib.b2g_halo_unpack.argtypes = [C.c_void_p, C.c_int32, C.c_void_p]
        lib.b2g_upload_contacts.argtypes = [C.c_void_p, C.c_int32, C.POINTER(ContactArrays)]
        lib.b2g_set_sequential_order.argtypes = [C.c_void_p, C.c_int32, i32p, i32p]
        lib.b2g_set_sequential_joint_order.argtypes = [C.c_void_p, C.c_int32, i32p]
        lib.b2g_set_kernel_timing.argtypes = [C.c_void_p, C.c_int32]
        lib.b2g_kernel_class_name.restype = C.c_char_p
        lib.b2g_kernel_class_name.argtypes = [C.c_int32]
        lib.b2g_get_kernel_timing.argtypes = [C.c_void_p, C.c_int32, C.POINTER(C.c_double), C.POINTER(C.c_int64),
                                              C.POINTER(C.c_double)]
        lib.b2g_host_alloc.argtypes = [C.POINTER(C.c_void_p), C.c_uint64]
        lib.b2g_host_free.argtypes = [C.c_void_p]
        lib.b2g_download_new_pairs.argtypes = [C.c_void_p, C.c_int32, i32p, i32p, i32p]
        lib.b2g_set_pair_vetoes.argtypes = [C.c_void_p, C.c_int32, i32p, i32p]
        lib.b2g_download_veto_seen.argtypes = [C.c_void_p, C.c_int32, u8p]
        lib.b2g_query_aabb.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32]
        lib.b2g_ray_cast_closest.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32,
                                             C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32]
        lib.b2g_ray_cast_all.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_int32,
                                         C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32]
        lib.b2g_download_joints.argtypes = [C.c_void_p, C.c_int32, C.c_int32, f32p]
        lib.b2g_debug_joint_colours.argtypes = [C.c_void_p, i32p]
        lib.b2g_upload_bodies_indexed.argtypes = [C.c_void_p, C.c_int32, i32p, C.POINTER(BodyArrays)]
        lib.b2g_upload_fixtures_indexed.argtypes = [C.c_void_p, C.c_int32, i32p, C.POINTER(FixtureArrays)]
        lib.b2g_upload_shapes_indexed.argtypes = [C.c_void_p, C.c_int32, i32p, f32p]
        lib.b2g_rotations.argtypes = [C.c_int32, C.c_int32, f32p, f32p]
        lib.b2g_compute_aabbs.argtypes = [C.c_int32, C.c_int32, i32p, i32p, f32p, C.c_int32, f32p, f32p]
        lib.b2g_collide_pairs.argtypes = [C.c_int32, C.c_int32, i32p, i32p, f32p, i32p, i32p, f32p, f32p, C.c_int32,
                                          f32p]
        lib.b2g_find_pairs.argtypes = [C.c_int32, C.c_int32, f32p, i32p, i32p, u8p, i32p, C.c_int32, i32p]
        lib.b2g_solve_sequential.argtypes = [C.c_int32, C.c_int32, f32p, f32p, f32p, C.c_int32, i32p, f32p, f32p,
                                             f32p, C.c_float, C.c_float, C.c_int32, C.c_int32, C.c_int32, f32p, f32p,
                                             i32p]
        _cuda = lib
    return _cuda


def check(rc, what="b2cuda"):
    if rc != OK:
        raise B2GError(f"{what} failed ({rc}): {load_cuda().b2g_last_error().decode()}")


def _declare_shim(lib, prefix):
    g = lambda n: getattr(lib, prefix + n)
    g("scene_create").restype = C.c_void_p
    g("scene_create").argtypes = [C.c_char_p, C.c_int, C.c_int]
    g("scene_destroy").argtypes = [C.c_void_p]
    g("scene_step").argtypes = [C.c_void_p, C.c_int]
    g("scene_time_steps").restype = C.c_double
    g("scene_time_steps").argtypes = [C.c_void_p, C.c_int]
    g("scene_set_iterations").argtypes = [C.c_void_p, C.c_int, C.c_int]
    g("scene_set_flags").argtypes = [C.c_void_p, C.c_int, C.c_int]
    for n in ("scene_body_count", "scene_fixture_count", "scene_contact_count", "scene_shape_quad_total"):
        g(n).argtypes = [C.c_void_p]
    g("scene_get_bodies").argtypes = [C.c_void_p, f32p]
    g("scene_get_body_params").argtypes = [C.c_void_p, f32p]
    g("scene_get_fixtures").argtypes = [C.c_void_p, i32p, i32p, i32p, i32p, f32p, i32p, f32p]
    g("scene_get_aabbs").argtypes = [C.c_void_p, f32p]
    g("scene_get_contacts").argtypes = [C.c_void_p, C.c_int, i32p, i32p, i32p, f32p, f32p]
    g("scene_query_aabb").argtypes = [C.c_void_p, C.c_int, f32p, C.c_int, i32p, i32p]
    g("scene_ray_cast_closest").argtypes = [C.c_void_p, C.c_int, f32p, i32p, f32p, f32p, f32p]
    g("scene_ray_cast_all").argtypes = [C.c_void_p, f32p, C.c_int, i32p, f32p, f32p]
    g("scene_time_ray_casts").restype = C.c_double
    g("scene_time_ray_casts").argtypes = [C.c_void_p, C.c_int, f32p]
    g("scene_joint_count").argtypes = [C.c_void_p]
    g("scene_get_joints").argtypes = [C.c_void_p, C.c_int, i32p, f32p, f32p]


_dist = None


def load_dist():
    """libb2cuda_dist.so: the NCCL transport of the halo exchange (one process per GPU)"""
    global _dist
    if _dist is None:
        load_cuda()
        p = lib_path("libb2cuda_dist.so")
        if not os.path.exists(p):
            raise B2GError(f"{p} is missing: run __graft_entry__.build()")
        lib = C.CDLL(p)
        lib.b2g_dist_unique_id.argtypes = [C.c_void_p]
        lib.b2g_dist_init.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_void_p)]
        lib.b2g_dist_exchange.argtypes = [C.c_void_p, C.c_void_p]
        lib.b2g_dist_destroy.argtypes = [C.c_void_p]
        _dist = lib
    return _dist


def load_gpu_scenes():
    """The drop-in C++ API (include/box2d) + scenes, over libb2cuda.so."""
    global _gpu_scenes
    if _gpu_scenes is None:
        load_cuda()
        p = lib_path("libb2gpu_scenes.so")
        if not os.path.exists(p):
            raise B2GError(f"{p} is missing: run __graft_entry__.build()")
        lib = C.CDLL(p)
        _declare_shim(lib, "b2gpu_")
        lib.b2gpu_scene_set_solver_mode.argtypes = [C.c_void_p, C.c_int]
        lib.b2gpu_scene_set_profiling.argtypes = [C.c_void_p, C.c_int]
        lib.b2gpu_scene_get_profile.argtypes = [C.c_void_p, f32p]
        _gpu_scenes = lib
    return _gpu_scenes


def fp(a):
    return a.ctypes.data_as(f32p) if a is not None else None


def ip(a):
    return a.ctypes.data_as(i32p) if a is not None else None


def up(a):
    return a.ctypes.data_as(u32p) if a is not None else None


def f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def u32(a):
    return np.ascontiguousarray(a, dtype=np.uint32)

"""oracle/bindings.py — ctypes views of the CHECKER: the compiled reference (oracle/_ref/libb2ref.so)
and the plain-C restatement (oracle/libb2oracle.so).

TEST / BASELINE INFRASTRUCTURE ONLY.  Imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py; nothing under box2d_optimized_b200/ imports it.
"""
import ctypes as C
import os

import numpy as np

from box2d_optimized_b200 import capi
from box2d_optimized_b200.capi import B2GError, f32p, i32p, u8p
from box2d_optimized_b200.scene import _Scene

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_ref = None


def ref_path():
    return os.path.join(ROOT, "oracle", "_ref", "libb2ref.so")


def load_ref():
    """TEST/BASELINE ONLY: the compiled reference (oracle/_ref).  Never used by the product path."""
    global _ref
    if _ref is None:
        p = ref_path()
        if not os.path.exists(p):
            raise B2GError(f"{p} is missing: `make -C oracle ref` (needs /root/reference)")
        lib = C.CDLL(p)
        capi._declare_shim(lib, "b2ref_")
        lib.b2ref_polygon_set.argtypes = [f32p, C.c_int, f32p]
        lib.b2ref_shape_mass.argtypes = [C.c_int, f32p, C.c_float, f32p]
        lib.b2ref_compute_aabbs.argtypes = [C.c_int, i32p, i32p, f32p, f32p, f32p]
        lib.b2ref_collide_pairs.argtypes = [C.c_int, i32p, i32p, f32p, i32p, i32p, f32p, f32p, f32p]
        lib.b2ref_world_collide.argtypes = [C.c_void_p]
        lib.b2ref_get_body_inv.argtypes = [C.c_void_p, f32p]
        lib.b2ref_get_inv_dt0.restype = C.c_float
        lib.b2ref_get_inv_dt0.argtypes = [C.c_void_p]
        lib.b2ref_get_sleep_times.argtypes = [C.c_void_p, f32p]
        lib.b2ref_get_joint_state.argtypes = [C.c_void_p, C.c_int, f32p]
        lib.b2ref_next_step_joint_order.argtypes = [C.c_void_p, C.c_int, i32p]
        lib.b2ref_step_recording_order.argtypes = [C.c_void_p, C.c_int, i32p, i32p]
        lib.b2ref_solve.argtypes = [C.c_int, f32p, f32p, f32p, C.c_int, i32p, f32p, f32p, f32p, C.c_float, C.c_float,
                                    C.c_int, C.c_int, C.c_int, f32p, f32p, i32p]
        _ref = lib
    return _ref


_oracle = None


def oracle_path():
    return os.path.join(ROOT, "oracle", "libb2oracle.so")


def load_oracle():
    """TEST ONLY: the plain-C restatement (oracle/b2_oracle.c).  Never used by the product path."""
    global _oracle
    if _oracle is None:
        p = oracle_path()
        if not os.path.exists(p):
            raise B2GError(f"{p} is missing: `make -C oracle port`")
        lib = C.CDLL(p)
        lib.b2o_compute_aabbs.argtypes = [C.c_int, i32p, i32p, f32p, f32p, f32p]
        lib.b2o_collide_pairs.argtypes = [C.c_int, i32p, i32p, f32p, i32p, i32p, f32p, f32p, f32p]
        lib.b2o_find_pairs.argtypes = [C.c_int, f32p, i32p, i32p, u8p, i32p, C.c_int, i32p]
        lib.b2o_solve.argtypes = [C.c_int, f32p, f32p, f32p, C.c_int, i32p, f32p, f32p, f32p, C.c_float, C.c_float,
                                  C.c_int, C.c_int, C.c_int, f32p, f32p, i32p]
        _oracle = lib
    return _oracle


class RefScene(_Scene):
    """The reference's own CPU b2World::Step on the same scene.  TEST / BASELINE ONLY."""
    prefix = "b2ref_"

    def __init__(self, name, size=0, seed=0):
        super().__init__(load_ref(), name, size, seed)

    def collide_now(self):
        self.lib.b2ref_world_collide(self.h)

    def body_inv(self):
        out = np.zeros((self.body_count, 2), np.float32)
        self.lib.b2ref_get_body_inv(self.h, capi.fp(out))
        return out

    def inv_dt0(self):
        return float(self.lib.b2ref_get_inv_dt0(self.h))

    def step_recording_order(self):
        """one Step with a PostSolve tap: returns the ordered (fixA, fixB) pairs in the order the
        reference's island solver visited them"""
        cap = max(self.contact_count, 1) + 16
        fa = np.zeros(cap, np.int32)
        fb = np.zeros(cap, np.int32)
        n = self.lib.b2ref_step_recording_order(self.h, cap, capi.ip(fa), capi.ip(fb))
        return fa[:n], fb[:n]

    def joint_state(self):
        """accumulated impulses of the revolute joints [n, 5] (white-box: b2_revolute_joint.h:178-181)"""
        n = self.lib.b2ref_scene_joint_count(self.h)
        out = np.zeros((max(n, 1), 5), np.float32)
        n = self.lib.b2ref_get_joint_state(self.h, n, capi.fp(out))
        return out[:n]

    def next_step_joint_order(self):
        """joint indices in the order the NEXT Step's island DFS will add them (oracle/ref_harness.cpp);
        runs the head of that Step (pair refresh + Collide), which the Step then repeats unchanged"""
        n = self.lib.b2ref_scene_joint_count(self.h)
        out = np.zeros(max(n, 1), np.int32)
        k = self.lib.b2ref_next_step_joint_order(self.h, n, capi.ip(out))
        return out[:k]

    def sleep_times(self):
        out = np.zeros(self.body_count, np.float32)
        self.lib.b2ref_get_sleep_times(self.h, capi.fp(out))
        return out

"""b2Rot::Set on the device against the host libm the reference links (b2_math.h:313-318).

Bit-exact gate: the reference's body transforms come from glibc's sinf/cosf, and one ulp in a
rotation is amplified by resting contacts, so the device restates glibc's algorithm
(box2d_optimized_b200/csrc/b2g_math.cuh rot_set).  numpy's own SIMD sin/cos are NOT the reference's
functions, so the check goes through ctypes to libm's sincosf."""
import ctypes as C
import ctypes.util

import numpy as np
import pytest

from box2d_optimized_b200 import capi

pytestmark = pytest.mark.gpu


def libm_sincos(angles):
    libm = C.CDLL(ctypes.util.find_library("m") or "libm.so.6")
    libm.sinf.restype = C.c_float
    libm.sinf.argtypes = [C.c_float]
    libm.cosf.restype = C.c_float
    libm.cosf.argtypes = [C.c_float]
    out = np.empty((len(angles), 2), np.float32)
    for i, a in enumerate(angles.tolist()):
        out[i, 0] = libm.sinf(a)
        out[i, 1] = libm.cosf(a)
    return out


def test_rotations_match_host_libm_bit_for_bit():
    rng = np.random.default_rng(5)
    n = 200000
    angles = np.concatenate([
        rng.uniform(-8, 8, n), rng.uniform(-119.9, 119.9, n // 2), rng.normal(0, 1e-3, n // 8),
        rng.uniform(-0.8, 0.8, n // 4), np.arange(-16, 17) * (np.pi / 4), np.array([0.0, -0.0, 0.75, 0.785, 2.0 ** -12, 119.99, -119.99]),
        rng.uniform(-1000, 1000, 64),  # beyond 120: glibc's large-argument path, fallback on the device
    ]).astype(np.float32)
    got = np.empty((len(angles), 2), np.float32)
    capi.check(capi.load_cuda().b2g_rotations(0, len(angles), capi.fp(angles), capi.fp(got)))
    want = libm_sincos(angles)
    small = np.abs(angles) < 120.0
    diff = got.view(np.uint32)[small] != want.view(np.uint32)[small]
    assert not diff.any(), f"{int(diff.sum())} of {int(small.sum()) * 2} differ, first at angle {angles[small][np.argwhere(diff)[0][0]]!r}"
    # large arguments: at most one ulp
    np.testing.assert_allclose(got[~small], want[~small], rtol=0, atol=1.2e-7)

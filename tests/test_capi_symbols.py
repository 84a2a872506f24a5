"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol
include/b2cuda.h declares; with no GPU the product path fails loudly instead of falling back."""
import ctypes
import os
import re

import pytest

from box2d_optimized_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "b2cuda.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b2g_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree():
    assert declared_symbols() == sorted(capi.CUDA_SYMBOLS + capi.DIST_SYMBOLS)


def test_library_exports_every_declared_symbol():
    """b2g_dist_* (the NCCL transport of the halo exchange) live in libb2cuda_dist.so, everything else in
    libb2cuda.so"""
    lib = ctypes.CDLL(capi.lib_path())
    dist = ctypes.CDLL(capi.lib_path("libb2cuda_dist.so"))
    for name in declared_symbols():
        home = dist if name in capi.DIST_SYMBOLS else lib
        assert hasattr(home, name), f"{name} declared in include/b2cuda.h but not exported"


def test_host_library_exports_scene_shim():
    lib = capi.load_gpu_scenes()
    for name in ("scene_create", "scene_step", "scene_get_bodies", "scene_get_contacts", "scene_get_fixtures"):
        assert hasattr(lib, "b2gpu_" + name)


def test_no_cpu_fallback_without_device():
    lib = capi.load_cuda()
    if lib.b2g_device_count() > 0:
        pytest.skip("a CUDA device is present")
    d = capi.ArenaDef(0, 1, 16, 16, 16, 16, 0, 0)
    h = ctypes.c_void_p()
    rc = lib.b2g_arena_create(ctypes.byref(d), ctypes.byref(h))
    assert rc == -4  # B2G_ERR_NO_DEVICE
    assert b"no CPU fallback" in lib.b2g_last_error()

"""box2d_optimized_b200 — B200-native b2World::Step.

Layout (only what the hot path needs):
  csrc/   CUDA kernels for sm_100a + the C-ABI of include/b2cuda.h  -> libb2cuda.so
  host/   drop-in C++ API (include/box2d/*.h) over the C-ABI         -> libb2gpu_scenes.so
  capi.py / arena.py / scene.py   ctypes + numpy views used by tests and bench.py
"""
from . import capi  # noqa: F401
from .arena import Arena, arena_from_scene, body_flags  # noqa: F401
from .scene import GpuScene  # noqa: F401

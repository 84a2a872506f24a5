"""Workload for compute-sanitizer (memcheck / racecheck): a few Steps through every solver driver.
usage: compute-sanitizer --tool memcheck python scripts/gpu_sanitize.py [steps-scale [scene,scene...]]
  pyramid(10)   -> k_solve_bins_fused (islands in shared-memory bins)
  mixed(3000)   -> one island over B2G_BIG_ISLAND bodies once it has piled up: tile plan / k_big_tiles
  tumbler(120)  -> revolute motor joint, serial overflow bucket (the container touches most boxes)
  chain(150)    -> joint colouring, colour-parallel joint passes"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from box2d_optimized_b200 import GpuScene

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
only = sys.argv[2].split(",") if len(sys.argv) > 2 else None
for name, size, steps in (("pyramid", 10, 60), ("mixed", 3000, 160), ("tumbler", 120, 200), ("chain", 150, 40)):
    if only and name not in only:
        continue
    g = GpuScene(name, size, 12345 if name == "mixed" else 0)
    g.step(max(2, int(steps * scale)))
    b = g.bodies()
    assert np.isfinite(b).all()
    print(f"{name}({size}): {g.body_count} bodies, {g.contact_count} contacts after {max(2, int(steps * scale))} steps", flush=True)
    g.close()
print("sanitize workload done")

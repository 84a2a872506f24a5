"""Run-to-run reproducibility at the headline size: two arenas built from the same scene, stepped side by side,
compared bit for bit every 20 steps (positions, velocities, contact count)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from box2d_optimized_b200 import GpuScene, capi, Arena, arena_from_scene

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 340
scene = GpuScene("mixed", n, 12345)
A = arena_from_scene(scene, max_contacts=8 * n)
B = arena_from_scene(scene, max_contacts=8 * n)
A.find_new_contacts(); B.find_new_contacts()
P = Arena.params()
sa, sb = capi.StepStats(), capi.StepStats()
ok = True
for k in range(steps):
    A.step(P, sa); B.step(P, sb)
    if (k + 1) % 20 == 0 or sa.num_contacts != sb.num_contacts:
        a = A.download_bodies(what=("pos", "vel")); b = B.download_bodies(what=("pos", "vel"))
        same = np.array_equal(a["pos"].view(np.uint32), b["pos"].view(np.uint32)) and np.array_equal(a["vel"].view(np.uint32), b["vel"].view(np.uint32))
        print(f"step {k + 1}: contacts {sa.num_contacts} / {sb.num_contacts}, constraints {sa.num_constraints} / {sb.num_constraints}, "
              f"colours {sa.num_colours} / {sb.num_colours}, identical {same}", flush=True)
        if not same:
            d = np.abs(a["pos"] - b["pos"]).max(axis=1)
            print("   bodies differing", int((d > 0).sum()), "max |dpos|", float(d.max()))
            ok = False
            break
print("REPRODUCIBLE" if ok else "NOT REPRODUCIBLE")
sys.exit(0 if ok else 1)

import sys; sys.path.insert(0, ".")
import numpy as np
from box2d_optimized_b200 import capi, Arena, arena_from_scene, GpuScene
sc = GpuScene("many_pyramids", 100, 0)
A = arena_from_scene(sc, max_contacts=8 * sc.body_count)
A.find_new_contacts()
P = Arena.params(); st = capi.StepStats()
for k in range(60): A.step(P, st)
for k in range(20):
    A.set_kernel_timing(True)
    A.step(P, st)
    kt = A.kernel_timing()
    A.set_kernel_timing(False)
    print(k, {n: round(v[0] * 1000, 1) for n, v in kt.items() if v[0] > 0 and n in ("bp_traverse", "bp_build", "sort_scan")}, st.num_pairs, "visits mean", round(st.bp_mean_visits, 1), "max", st.bp_max_visits, "rebuilt", st.bp_rebuilt)

import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from box2d_optimized_b200 import GpuScene, capi, Arena, arena_from_scene
def run(scene, copies, mode, steps):
    A = arena_from_scene(scene, copies=copies, num_worlds=copies); A.find_new_contacts()
    P = Arena.params(solver_mode=mode); st = capi.StepStats(); out = []
    for _ in range(steps):
        A.step(P, st); out.append((st.num_contacts, st.num_constraints, st.num_colours, st.num_overflow, st.colour_rounds))
    d = A.download_bodies(what=("pos",)); A.close(); return d["pos"], out
which = sys.argv[1:] or ["tumbler", "pyramid"]
if "tumbler" in which:
    s = GpuScene("tumbler", 150, 0); s.step(170); run(s, 1, 0, 200); run(s, 8, 0, 200); print("tumbler done")
if "pyramid" in which:
    s2 = GpuScene("pyramid", 12, 0); run(s2, 1, 0, 200); run(s2, 8, 0, 200); print("pyramid done")
if "single8" in which:
    s3 = GpuScene("mixed", 700, 12345); run(s3, 8, 0, 100); print("warm 8 done")
m = GpuScene("mixed", 700, 12345)
p1, o1 = run(m, 1, 0, 120); p8, o8 = run(m, 8, 0, 120)
for k, (a, b) in enumerate(zip(o1, o8)):
    if a[0] * 8 != b[0] or a[1] * 8 != b[1] or a[2] != b[2]:
        print("first stats mismatch at step", k + 1, "single", a, "batched", b); break
else:
    print("stats identical")
nb = m.body_count
print("positions identical:", [bool(np.array_equal(p1.view(np.uint32), p8[j*nb:(j+1)*nb].view(np.uint32))) for j in range(8)])

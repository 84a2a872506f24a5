import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from box2d_optimized_b200 import capi, GpuScene
from oracle.bindings import RefScene
name, size = sys.argv[1], int(sys.argv[2])
mode = int(sys.argv[3]) if len(sys.argv) > 3 else 0
a = GpuScene(name, size, 99, solver_mode=mode); b = GpuScene(name, size, 99, solver_mode=mode)
for step in range(200):
    a.step(1); b.step(1)
    ba, bb = a.bodies(), b.bodies()
    ca, cb = a.contacts(), b.contacts()
    same_b = np.array_equal(ba.view(np.uint32), bb.view(np.uint32))
    same_c = len(ca["fix_a"]) == len(cb["fix_a"]) and np.array_equal(ca["fix_a"], cb["fix_a"]) and np.array_equal(ca["fix_b"], cb["fix_b"])
    same_m = same_c and np.array_equal(ca["manifold"].view(np.uint32), cb["manifold"].view(np.uint32))
    if not (same_b and same_c and same_m):
        print("first divergence at step", step, "bodies", same_b, "contact list", same_c, "manifolds", same_m)
        if not same_b:
            d = np.nonzero((ba.view(np.uint32) != bb.view(np.uint32)).any(1))[0]
            print("bodies differing:", len(d), d[:10]); print(ba[d[0]], bb[d[0]])
        if same_c and not same_m:
            d = np.nonzero((ca["manifold"].view(np.uint32) != cb["manifold"].view(np.uint32)).any(1))[0]
            print("manifolds differing", len(d), d[:10]); i=d[0]; print(ca["manifold"][i], cb["manifold"][i], ca["fix_a"][i], ca["fix_b"][i])
        break
else:
    print("no divergence in 200 steps")

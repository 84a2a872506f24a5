"""Real-timeline split of a step (events on the step's stream, b2Profile fields): collide / solve / broadphase."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from box2d_optimized_b200 import GpuScene
for name, size, warm in (("mixed", 100000, 300), ("many_pyramids", 100, 100)):
    g = GpuScene(name, size, 12345 if name == "mixed" else 0)
    g.step(warm)
    g.set_profiling(True)
    acc = []
    for k in range(50):
        g.step(1)
        p = g.profile()
        acc.append([p["step"], p["collide"], p["solve"], p["broadphase"]])
    a = np.array(acc)
    print(name, size, "ms mean step %.4f collide %.4f solve %.4f broadphase %.4f | median step %.4f" % (*a.mean(axis=0), np.median(a[:, 0])))

import sys, numpy as np
sys.path.insert(0, ".")
from box2d_optimized_b200 import GpuScene, capi
from oracle.bindings import RefScene
for name, size in ((sys.argv[1] if len(sys.argv) > 1 else "welds", int(sys.argv[2]) if len(sys.argv) > 2 else 6),):
    for mode in (capi.SOLVER_COLOURED, capi.SOLVER_SEQUENTIAL):
        ref, gpu = RefScene(name, size, 0), GpuScene(name, size, 0, solver_mode=mode)
        for k in range(6):
            ref.step(40); gpu.step(40)
            rb, gb = ref.bodies(), gpu.bodies()
            e = np.abs(rb[:, 4:6] - gb[:, 4:6]).max(axis=1)
            print(name, "mode", mode, "step", 40 * (k + 1), "err per body", np.round(e, 3).tolist())

import sys; sys.path.insert(0, ".")
import numpy as np
from box2d_optimized_b200 import GpuScene
from oracle.bindings import RefScene
n = int(sys.argv[1]) if len(sys.argv) > 1 else 150
ref, gpu = RefScene("chain", n, 0), GpuScene("chain", n, 0)
for k in range(1, 31):
    ref.step(1); gpu.step(1)
    rb, gb = ref.bodies(), gpu.bodies()
    d = np.abs(rb[:, 4:6] - gb[:, 4:6]).max(1)
    i = int(np.argmax(d))
    if k <= 6 or k % 5 == 0:
        print(k, "max diff %.5f at body %d" % (d[i], i), "ref", rb[i, 4:7], "gpu", gb[i, 4:7], "contacts", ref.contact_count, gpu.contact_count)

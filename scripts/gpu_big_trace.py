"""Where does k_big_solve spend a pass?  Needs libb2cuda.so built with -DB2G_BIG_TRACE (per-block
globaltimer stamps at every grid_arrive / grid_wait exit).  Prints, per barrier, the work time of the
median and the slowest block and the time from the last arrival to the release."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from box2d_optimized_b200 import GpuScene, capi

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 300
g = GpuScene("mixed", n, 12345)
g.step(steps)
g.bodies()
lib = capi.load_cuda()
CAP, B = 1024, 148
buf = np.zeros((B, CAP, 2), np.uint64)
lib.b2g_debug_big_trace.argtypes = [C.c_void_p, C.c_int]
rc = lib.b2g_debug_big_trace(buf.ctypes.data_as(C.c_void_p), B)
assert rc == 0
t = buf.astype(np.int64)
nb = int((t[0, :, 0] > 0).sum())
arr, rel = t[:, :nb, 0], t[:, :nb, 1]
t0 = arr.min()
print("barriers", nb, "kernel span us", (rel.max() - t0) / 1e3)
work = arr[:, 1:] - rel[:, :-1]          # release of barrier i-1 -> arrival at barrier i, per block
lastArr = arr.max(axis=0)
relMin, relMax = rel.min(axis=0), rel.max(axis=0)
print("per pass (us): work median-block %.2f, slowest-block %.2f | last arrival -> first release %.2f, -> last release %.2f | period %.2f" % (
    np.median(work, axis=0).mean() / 1e3, work.max(axis=0).mean() / 1e3, (relMin - lastArr).mean() / 1e3,
    (relMax - lastArr).mean() / 1e3, np.diff(lastArr).mean() / 1e3))
for i in range(0, nb - 1, max(1, nb // 60)):
    print(i, "work med %.2f max %.2f argmax %d | rel-lastArr %.2f..%.2f" % (
        np.median(work[:, i]) / 1e3, work[:, i].max() / 1e3, int(work[:, i].argmax()), (relMin[i + 1] - lastArr[i + 1]) / 1e3,
        (relMax[i + 1] - lastArr[i + 1]) / 1e3))

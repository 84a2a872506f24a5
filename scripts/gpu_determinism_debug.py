"""first divergence between two identical arenas: colours or body state?"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from box2d_optimized_b200 import GpuScene, capi, Arena, arena_from_scene
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
start = int(sys.argv[2]) if len(sys.argv) > 2 else 60
scene = GpuScene("mixed", n, 12345)
A = arena_from_scene(scene, max_contacts=8 * n); B = arena_from_scene(scene, max_contacts=8 * n)
A.find_new_contacts(); B.find_new_contacts()
P = Arena.params(); sa, sb = capi.StepStats(), capi.StepStats()
def contacts(X):
    c = X.download_contacts()
    key = c["fix_a"].astype(np.int64) * (1 << 32) + c["fix_b"]
    o = np.argsort(key)
    return key[o], c["colour"][o], c["manifold"][o]
for k in range(140):
    A.step(P, sa); B.step(P, sb)
    if k + 1 < start:
        continue
    a = A.download_bodies(what=("pos", "vel")); b = B.download_bodies(what=("pos", "vel"))
    same = np.array_equal(a["pos"].view(np.uint32), b["pos"].view(np.uint32)) and np.array_equal(a["vel"].view(np.uint32), b["vel"].view(np.uint32))
    ka, ca, ma = contacts(A); kb, cb, mb = contacts(B)
    samek = np.array_equal(ka, kb)
    samec = samek and np.array_equal(ca, cb)
    print(f"step {k+1}: bodies identical {same}; contact keys identical {samek}; colours identical {samec}; colours {sa.num_colours}/{sb.num_colours} "
          f"rounds {sa.colour_rounds}/{sb.colour_rounds} serial {sa.num_overflow}/{sb.num_overflow} constraints {sa.num_constraints}", flush=True)
    if samek and not samec:
        d = np.nonzero(ca != cb)[0]
        print("   colour differs on", len(d), "contacts; e.g.", [(int(ca[i]), int(cb[i])) for i in d[:12]])
    import ctypes as C
    def tstate(X):
        plan = np.zeros(11, np.uint32); tc = np.zeros(160, np.int32); ts = np.zeros(X.num_bodies, np.int32)
        X.lib.b2g_debug_tile_state.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        X.lib.b2g_debug_tile_state(X.h, plan.ctypes.data_as(C.c_void_p), tc.ctypes.data_as(C.c_void_p), ts.ctypes.data_as(C.c_void_p))
        return plan, tc, ts
    pa, tca, tsa = tstate(A); pb, tcb, tsb = tstate(B)
    print("   plan identical", np.array_equal(pa, pb), "S,R", pa[5:7], pb[5:7], "count", pa[4], pb[4], "| tile counts identical", np.array_equal(tca, tcb),
          "| tile of body identical", np.array_equal(tsa // 2048, tsb // 2048), "tiles used", int((tca > 0).sum()), int((tcb > 0).sum()))
    if not np.array_equal(pa, pb):
        print("   planA", pa, pa[7:].view(np.float32)); print("   planB", pb, pb[7:].view(np.float32))
    if not same:
        d = np.abs(a["pos"] - b["pos"]).max(axis=1); dv = np.abs(a["vel"] - b["vel"]).max(axis=1)
        print("   bodies differing: pos", int((d > 0).sum()), "vel", int((dv > 0).sum()), "max", float(d.max()), float(dv.max()))
        break

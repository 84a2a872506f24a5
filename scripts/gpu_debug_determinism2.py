import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from box2d_optimized_b200 import capi, Arena, arena_from_scene
from oracle.bindings import RefScene
mode = int(sys.argv[1]) if len(sys.argv) > 1 else 1
ref = RefScene("mixed", 1500, 99)
A = arena_from_scene(ref); B = arena_from_scene(ref)
A.find_new_contacts(); B.find_new_contacts()
P = Arena.params(solver_mode=mode)
lib = capi.load_cuda()
def cmp(tag, step):
    a = A.download_bodies(); b = B.download_bodies()
    ok = True
    for k in a:
        if not np.array_equal(a[k].view(np.uint32), b[k].view(np.uint32)):
            d = np.nonzero((a[k].view(np.uint32) != b[k].view(np.uint32)).reshape(len(a[k]), -1).any(1))[0]
            print(f"step {step} {tag}: body array {k} differs at {len(d)} bodies {d[:8]}", a[k][d[0]], b[k][d[0]]); ok = False
    ca, cb = A.download_contacts(), B.download_contacts()
    if len(ca["fix_a"]) != len(cb["fix_a"]):
        print(f"step {step} {tag}: contact count differs {len(ca['fix_a'])} {len(cb['fix_a'])}"); return False
    for k in ca:
        if not np.array_equal(ca[k].view(np.uint32), cb[k].view(np.uint32)):
            d = np.nonzero((ca[k].view(np.uint32) != cb[k].view(np.uint32)).reshape(len(ca[k]), -1).any(1))[0]
            i = d[0]
            print(f"step {step} {tag}: contact array {k} differs at {len(d)} contacts {d[:8]}; first: fixA {ca['fix_a'][i]} fixB {ca['fix_b'][i]}\n", ca[k][i], "\n", cb[k][i]); ok = False
    return ok
for step in range(120):
    capi.check(lib.b2g_step_collide(A.h, C.byref(P))); capi.check(lib.b2g_step_collide(B.h, C.byref(P)))
    A.synchronize(); B.synchronize()
    if not cmp("after collide", step): break
    capi.check(lib.b2g_step_solve(A.h, C.byref(P), None)); capi.check(lib.b2g_step_solve(B.h, C.byref(P), None))
    if not cmp("after solve", step): break
else:
    print("no divergence")

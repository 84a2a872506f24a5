"""First-contact GPU sanity run: exercises every stage once and prints diagnostics."""
import sys, os, time, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ctypes as C
from box2d_optimized_b200 import capi, Arena, arena_from_scene, GpuScene
from oracle.bindings import RefScene

def section(name):
    print(f"\n=== {name} ===", flush=True)

def pairs_of(c):
    a = np.minimum(c["fix_a"], c["fix_b"]).astype(np.int64); b = np.maximum(c["fix_a"], c["fix_b"]).astype(np.int64)
    return set(zip(a.tolist(), b.tolist()))

try:
    section("device")
    lib = capi.load_cuda()
    print("devices", lib.b2g_device_count())

    section("arena from ref pyramid, find_new_contacts")
    ref = RefScene("pyramid", 20)
    ref.step(1)  # reference creates contacts at first step
    A = arena_from_scene(ref)
    A.find_new_contacts()
    cg = A.download_contacts()
    cr = ref.contacts()
    print("gpu contacts", len(cg["fix_a"]), "ref contacts", len(cr["fix_a"]))
    pg, pr = pairs_of(cg), pairs_of(cr)
    print("pair set equal:", pg == pr, "only gpu", len(pg - pr), "only ref", len(pr - pg))
    aabb_g = A.download_aabbs(); aabb_r = ref.aabbs()
    print("aabb max abs diff", np.abs(aabb_g - aabb_r).max())

    section("step arena 60 steps coloured")
    P = Arena.params()
    st = capi.StepStats()
    for i in range(60):
        A.step(P, st)
        if i % 10 == 0 or i == 59:
            print(i, st.as_dict())
    bd = A.download_bodies()
    print("pos[1]", bd["pos"][1], "vel[1]", bd["vel"][1])
    ref.step(60)
    rb = ref.bodies()
    print("ref c[1]", rb[1, 4:7], "max |dpos| vs ref", np.abs(bd["pos"][:, :2] - rb[:, 4:6]).max())

    section("GpuScene pyramid via C++ API")
    g = GpuScene("pyramid", 20)
    t = time.time(); g.step(100); print("100 steps wall", time.time() - t)
    gb = g.bodies()
    r2 = RefScene("pyramid", 20); r2.step(100); rb2 = r2.bodies()
    print("max |dpos| gpu-api vs ref after 100:", np.abs(gb[:, 4:6] - rb2[:, 4:6]).max(), "contacts", g.contact_count, r2.contact_count)

    section("sequential mode")
    g2 = GpuScene("pyramid", 20, solver_mode=capi.SOLVER_SEQUENTIAL)
    g2.step(100); gb2 = g2.bodies()
    print("max |dpos| seq vs ref after 100:", np.abs(gb2[:, 4:6] - rb2[:, 4:6]).max())

    section("many_pyramids timing")
    g3 = GpuScene("many_pyramids", 100)
    g3.step(20)
    ms = g3.time_steps(100)
    print("many_pyramids 100 steps: %.3f ms/step" % (ms / 100), "bodies", g3.body_count, "contacts", g3.contact_count)
    g3.set_profiling(True); g3.step(1); print(g3.profile())

    section("mixed 10000")
    g4 = GpuScene("mixed", 10000, 12345)
    g4.step(50)
    ms = g4.time_steps(100)
    print("mixed10k: %.3f ms/step" % (ms / 100), "contacts", g4.contact_count)
    print("ALL OK")
except Exception:
    traceback.print_exc()
    sys.exit(1)

"""Where does the oversize-island kernel (k_big_tiles, or k_big_solve with B2G_NO_TILES=1) spend its time?
Needs a library built with -DB2G_BIG_TRACE (per-block globaltimer stamps at every grid_arrive / grid_wait exit):

    python -c "from box2d_optimized_b200 import build as b; b.build_cuda(True, 'libb2cuda_trace.so', ['-DB2G_BIG_TRACE'])"
    B2G_CUDA_LIB=box2d_optimized_b200/libb2cuda_trace.so python scripts/gpu_big_trace.py [bodies] [steps]

Prints, per grid barrier, the work time (release of the previous barrier -> arrival) of the median and the slowest
block and the time from the last arrival to the release, plus the colour histograms of both domains."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from box2d_optimized_b200 import GpuScene, capi, Arena, arena_from_scene

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 310
scene = GpuScene("mixed", n, 12345)
A = arena_from_scene(scene, max_contacts=8 * n)
A.find_new_contacts()
P = Arena.params()
st = capi.StepStats()
for _ in range(steps):
    A.step(P, st)
A.synchronize()
print("constraints", st.num_constraints, "colours", st.num_colours, "serial", st.num_overflow, "rounds", st.colour_rounds)
col = A.download_contacts()["colour"]
act = col[col >= 0]
print("interior colour histogram:", np.bincount(act[act < 32], minlength=25).tolist())
print("cut colour histogram     :", np.bincount(act[act >= 32] - 32, minlength=25).tolist())
lib = capi.load_cuda()
if hasattr(lib, "b2g_debug_tile_marks"):
    marks = np.zeros((320, 64), np.uint64)
    lib.b2g_debug_tile_marks.argtypes = [C.c_void_p]
    if lib.b2g_debug_tile_marks(marks.ctypes.data_as(C.c_void_p)) == 0 and marks[0, 0] > 0:
        m = marks[marks[:, 0] > 0].astype(np.int64)
        t0 = m[:, 0].min()
        names = {0: "kernel start", 1: "tile loaded", 2: "barrier 0 passed", 3: "constraints prepared", 4: "serial prepared",
                 5: "warm start done", 31: "positions integrated", 40: "end"}
        for i in range(6, 14):
            names[i] = f"velocity sweep {i - 6} done"
        for i in range(32, 35):
            names[i] = f"position sweep {i - 32} done"
        for i in range(44, 56):
            names[i] = f"  interior of sweep {i - 44} done"
        names.update({56: "    sweep 4: boundary published", 57: "    sweep 4: thread 0's cut constraints done", 58: "    sweep 4: boundary reloaded"})
        order = [0, 1, 2, 3, 4, 44, 5] + [x for it in range(8) for x in ((45 + it, 56, 57, 58, 6 + it) if it == 4 else (45 + it, 6 + it))] + [31] + [x for it in range(3) for x in (53 + it, 32 + it)] + [40]
        prev = None
        print("phase marks of k_big_tiles, us since the first block started (min / median / max over blocks; delta of medians):")
        for i in order:
            col = m[:, i]
            if (col <= 0).all():
                continue
            med = float(np.median(col - t0)) / 1e3
            print(f"  {names[i]:44s} {float((col - t0).min()) / 1e3:8.2f} {med:8.2f} {float((col - t0).max()) / 1e3:8.2f}   +{(med - prev) if prev is not None else 0.0:7.2f}")
            prev = med
if hasattr(lib, "b2g_debug_tile_marks") and "--tiles" in sys.argv:
    # per tile: interior time of velocity sweep 4 against what the tile holds
    nbod = scene.body_count
    plan = np.zeros(11, np.uint32); tc = np.zeros(320, np.int32); ts = np.zeros(nbod, np.int32)
    lib.b2g_debug_tile_state.argtypes = [C.c_void_p] * 4
    lib.b2g_debug_tile_state(A.h, plan.ctypes.data_as(C.c_void_p), tc.ctypes.data_as(C.c_void_p), ts.ctypes.data_as(C.c_void_p))
    ntile = int((marks[:, 0] > 0).sum())
    cap = int(ts.max()) // ntile + 1
    cap = 1 << (cap - 1).bit_length()
    con = A.download_contacts()
    fx = scene.fixtures()
    ba, bb = fx["body"][con["fix_a"]], fx["body"][con["fix_b"]]
    ok = (con["colour"] >= 0) & (con["colour"] < 32)
    ta = np.where(ts[ba] >= 0, ts[ba] // cap, np.where(ts[bb] >= 0, ts[bb] // cap, -1))
    m2 = marks[:ntile].astype(np.int64)
    tin = (m2[:, 49] - m2[:, 9]) / 1e3      # previous sweep done -> interior of sweep 4 done (this block)
    rows = []
    for t_ in range(ntile):
        sel = ok & (ta == t_)
        cols = con["colour"][sel]
        hist = np.bincount(cols, minlength=25)
        rows.append((tin[t_], int(tc[t_]), int(sel.sum()), int((hist > 0).sum()), int(hist.max())))
    rows = np.array(rows)
    order = np.argsort(rows[:, 0])
    print("tile interior time of velocity sweep 4 (us) | bodies | interior constraints | colours used | largest colour")
    for k in list(order[:5]) + list(order[len(order) // 2 - 2: len(order) // 2 + 3]) + list(order[-8:]):
        print("  %6.2f | %5d | %5d | %2d | %4d" % tuple(rows[k]))
    S_, R_ = int(plan[5]), int(plan[6])
    print("plan: S", S_, "R", R_, "count", int(plan[4]), "bounds x", plan[7:8].view(np.float32), "invDx", plan[8:9].view(np.float32), "y0", plan[9:10].view(np.float32), "invDy", plan[10:11].view(np.float32))
    grid = tc[:S_ * R_].reshape(S_, R_)
    print("bodies per strip:", grid.sum(axis=1).tolist())
    print("strip 0 rows:", grid[0].tolist())
    print("strip %d rows:" % (S_ // 2), grid[S_ // 2].tolist())
    print("correlation of interior time with: bodies %.2f, constraints %.2f, colours %.2f, largest colour %.2f" % tuple(
        np.corrcoef(rows[:, 0], rows[:, j])[0, 1] for j in (1, 2, 3, 4)))
CAP, B = 1024, 148
buf = np.zeros((B, CAP, 2), np.uint64)
lib.b2g_debug_big_trace.argtypes = [C.c_void_p, C.c_int]
rc = lib.b2g_debug_big_trace(buf.ctypes.data_as(C.c_void_p), B)
assert rc == 0
t = buf.astype(np.int64)
nb = int((t[0, :, 0] > 0).sum())
arr, rel = t[:, :nb, 0], t[:, :nb, 1]
t0 = arr.min()
print("barriers", nb, "first arrival -> last release us", (rel.max() - t0) / 1e3)
work = arr[:, 1:] - rel[:, :-1]          # release of barrier i-1 -> arrival at barrier i, per block
lastArr = arr.max(axis=0)
relMin, relMax = rel.min(axis=0), rel.max(axis=0)
print("per barrier (us): work median-block %.2f, slowest-block %.2f | last arrival -> first release %.2f, -> last release %.2f | period %.2f" % (
    np.median(work, axis=0).mean() / 1e3, work.max(axis=0).mean() / 1e3, (relMin - lastArr).mean() / 1e3,
    (relMax - lastArr).mean() / 1e3, np.diff(lastArr).mean() / 1e3))
print("sum over barriers: slowest-block work %.1f us, barrier latency %.1f us" % (
    work.max(axis=0).sum() / 1e3, (relMax - lastArr).sum() / 1e3))
for i in range(nb - 1):
    print(i + 1, "work med %.2f max %.2f min %.2f argmax %d | rel-lastArr %.2f..%.2f" % (
        np.median(work[:, i]) / 1e3, work[:, i].max() / 1e3, work[:, i].min() / 1e3, int(work[:, i].argmax()),
        (relMin[i + 1] - lastArr[i + 1]) / 1e3, (relMax[i + 1] - lastArr[i + 1]) / 1e3))

import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from box2d_optimized_b200 import GpuScene, capi, Arena, arena_from_scene
scene = GpuScene("mixed", 700, 12345)
nb = scene.body_count; nf = scene.fixture_count
COP = 8
A = arena_from_scene(scene, copies=COP, num_worlds=COP); A.find_new_contacts()
S = arena_from_scene(scene, copies=1, num_worlds=1); S.find_new_contacts()
P = Arena.params(); st = capi.StepStats(); ss = capi.StepStats()
for k in range(120):
    A.step(P, st); S.step(P, ss)
    d = A.download_bodies(what=("pos", "vel")); e = S.download_bodies(what=("pos", "vel"))
    c = A.download_contacts(); cs = S.download_contacts()
    pos = d["pos"].reshape(COP, nb, 4)
    same = [bool(np.array_equal(e["pos"].view(np.uint32), pos[j].view(np.uint32))) for j in range(COP)]
    bad = not all(same) or st.num_contacts != len(c["fix_a"]) or ss.num_contacts != len(cs["fix_a"]) or st.num_contacts != COP * ss.num_contacts
    if bad:
        print(f"step {k+1}: stats contacts batched {st.num_contacts} (downloaded {len(c['fix_a'])}), single {ss.num_contacts} (downloaded {len(cs['fix_a'])}); "
              f"positions equal to single: {same}; constraints {st.num_constraints}/{ss.num_constraints} colours {st.num_colours}/{ss.num_colours} "
              f"serial {st.num_overflow}/{ss.num_overflow} rounds {st.colour_rounds}/{ss.colour_rounds}")
        w = c["fix_a"] // nf
        print("   per world", np.bincount(w, minlength=COP).tolist())
        m = w == 0
        k0 = (c["fix_a"][m]).astype(np.int64) * (1 << 32) + c["fix_b"][m]; o0 = np.argsort(k0)
        k1 = cs["fix_a"].astype(np.int64) * (1 << 32) + cs["fix_b"]; o1 = np.argsort(k1)
        print("   keys equal", np.array_equal(k0[o0], k1[o1]))
        if np.array_equal(k0[o0], k1[o1]):
            c0, c1 = c["colour"][m][o0], cs["colour"][o1]
            dd = np.nonzero(c0 != c1)[0]; print("   colour diffs", len(dd), [(int(c0[i]), int(c1[i])) for i in dd[:10]])
        break
else:
    print("single and batched identical for 120 steps")

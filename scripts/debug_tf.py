import sys, numpy as np
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import util
from box2d_optimized_b200 import capi, Arena, arena_from_scene
from oracle.bindings import RefScene
from test_world_step_parity import mirror_reference_state
name, size, seed, steps = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
ref = RefScene(name, size, seed); ref.step(size + 2 if name == "tumbler" else 1)
A = arena_from_scene(ref, max_contacts=max(4096, 16 * ref.body_count))
params, inv = ref.body_params(), ref.body_inv()
import os
PI = int(os.environ.get("POSITERS", "3"))
ref.set_iterations(8, PI)
P = Arena.params(solver_mode=capi.SOLVER_SEQUENTIAL, pos_iters=PI); stats = capi.StepStats()
fixbody = ref.fixtures()["body"]
import os
fresh = int(os.environ.get("FRESH", "-1"))
EPS = float(os.environ.get("EPS", "2e-5"))
for k in range(steps):
    if k == fresh:
        A.close(); A = arena_from_scene(ref, max_contacts=max(4096, 16 * ref.body_count)); print("fresh arena at", k)
    before = mirror_reference_state(A, ref, params, inv)
    b0 = ref.bodies()
    jo = ref.next_step_joint_order(); A.set_sequential_joint_order(jo)
    fa, fb = ref.step_recording_order(); A.set_sequential_order(fa, fb); A.step(P, stats)
    if k < 3: print('joint order', jo.tolist())
    if stats.num_constraints != len(fa): print("step", k, "constraints", stats.num_constraints, "ref", len(fa))
    cg, cr = A.download_contacts(), ref.contacts()
    sg, sr = util.pair_set(cg["fix_a"], cg["fix_b"]), util.pair_set(cr["fix_a"], cr["fix_b"])
    if sg != sr: print("step", k, "pairs gpu-only", sorted(sg - sr), "ref-only", sorted(sr - sg))
    rb = ref.bodies(); gb = A.download_bodies(what=("pos", "vel", "flags", "force"))
    ev = np.abs(gb["vel"][:, :3] - rb[:, 7:10]) / np.maximum(1, np.abs(rb[:, 7:10]))
    ep = np.abs(gb["pos"][:, :3] - rb[:, 4:7]) / np.maximum(1, np.abs(rb[:, 4:7]))
    if ev.max() > EPS or ep.max() > EPS:
        print("bodies with error", np.nonzero((ev.max(1) > EPS) | (ep.max(1) > EPS))[0])
        js = A.download_joints(len(ref.joint_state()))
        if js is not None: print('joint state diff', np.abs(js - ref.joint_state()).max(axis=1))
        i = int(np.argmax(np.maximum(ev.max(1), ep.max(1))))
        nc = [(int(a), int(b)) for a, b in zip(fa, fb) if fixbody[a] == i or fixbody[b] == i]
        print(f"step {k} body {i} type {rb[i,11]} ev {ev[i]} ep {ep[i]} gpu vel {gb['vel'][i,:3]} ref vel {rb[i,7:10]} before vel {b0[i,7:10]} a {b0[i,6]} constraints {nc}")
        if EPS == 0:
            for q in range(len(cr["fix_a"])):
                if fixbody[cr["fix_a"][q]] == i or fixbody[cr["fix_b"][q]] == i:
                    for q2 in range(len(cg["fix_a"])):
                        if cg["fix_a"][q2] == cr["fix_a"][q] and cg["fix_b"][q2] == cr["fix_b"][q]:
                            d = cg["manifold"][q2].view(np.uint32) != cr["manifold"][q].view(np.uint32)
                            print("gpu manifold differs at", np.nonzero(d)[0].tolist(), cg["manifold"][q2][d].tolist(), cr["manifold"][q][d].tolist())
                    print("ref contact", int(cr["fix_a"][q]), int(cr["fix_b"][q]), "flags", hex(int(cr["flags"][q])), "manifold", cr["manifold"][q].tolist())
            print("gpu pos", gb["pos"][i, :3].tolist(), "ref pos", rb[i, 4:7].tolist(), "before", b0[i, 4:7].tolist())
            break

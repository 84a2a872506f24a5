"""Solver parity, sequential mode: b2g_solve_sequential (one thread, the oracle's constraint
order) against b2ContactSolver driven by the reference harness on identical inputs.  north_star
gate: velocity iterates after every velocity iteration and position iterates after every
position iteration within 1e-4 (measured: ~1e-6 or better; the only non-identical operations
are sinf/cosf of the body angles)."""
import ctypes as C
import os

import numpy as np
import oracle.bindings as oracle_bindings  # noqa: E402  (the checker)
import pytest

from box2d_optimized_b200 import capi

GOLD = os.path.join(os.path.dirname(__file__), "golden")
pytestmark = pytest.mark.gpu
TOL = 1e-4


def gpu_solve(g, vi=8, pi=3, warm=1, dt_ratio=1.0):
    nb, nc = len(g["pos"]), len(g["index"])
    pos, vel, man = g["pos"].copy(), g["vel"].copy(), g["manifold"].copy()
    vit = np.zeros((vi, nb, 4), np.float32)
    pit = np.zeros((pi, nb, 4), np.float32)
    done = C.c_int32()
    mass, index, mat, radii = capi.f32(g["mass"]), capi.i32(g["index"]), capi.f32(g["material"]), capi.f32(g["radii"])
    capi.check(capi.load_cuda().b2g_solve_sequential(
        0, nb, capi.fp(pos), capi.fp(vel), capi.fp(mass), nc, capi.ip(index), capi.fp(man), capi.fp(mat),
        capi.fp(radii), float(g["dt"]), dt_ratio, warm, vi, pi, capi.fp(vit), capi.fp(pit), C.byref(done)))
    return pos, vel, man, vit, pit, done.value


def impulse_err(man, ref):
    """stored impulses of the points that exist (a 1-point manifold's second slot is stale memory
    in the reference)"""
    cnt = ref[:, 15].copy().view(np.int32)
    e = rel_err(man[cnt > 0][:, [6, 7]], ref[cnt > 0][:, [6, 7]])
    if (cnt > 1).any():
        e = max(e, rel_err(man[cnt > 1][:, [10, 11]], ref[cnt > 1][:, [10, 11]]))
    return e


def rel_err(a, b):
    return float(np.max(np.abs(a - b) / np.maximum(1.0, np.abs(b))))


@pytest.mark.parametrize("name", ["pyramid", "mixed"])
def test_golden_iterates(name):
    g = np.load(os.path.join(GOLD, f"solver_{name}.npz"))
    pos, vel, man, vit, pit, done = gpu_solve(g)
    assert done == int(g["pos_iters_done"])
    worst = 0.0
    for k in range(vit.shape[0]):
        e = rel_err(vit[k][:, :3], g["vel_iterates"][k][:, :3])
        assert e <= TOL, f"velocity iterate {k}: {e}"
        worst = max(worst, e)
    for k in range(done):
        e = rel_err(pit[k][:, :3], g["pos_iterates"][k][:, :3])
        assert e <= TOL, f"position iterate {k}: {e}"
        worst = max(worst, e)
    assert rel_err(pos[:, :3], g["pos_out"][:, :3]) <= TOL
    assert rel_err(vel[:, :3], g["vel_out"][:, :3]) <= TOL
    assert impulse_err(man, g["manifold_out"]) <= TOL
    print(f"{name}: worst relative iterate error {worst:g}")


@pytest.mark.parametrize("name,size,steps,warm", [("pyramid", 20, 30, 1), ("pyramid", 20, 200, 1), ("mixed", 2000, 90, 1),
                                                   ("mixed", 2000, 90, 0), ("tumbler", 150, 220, 1)])
def test_live_reference_iterates(require_ref, name, size, steps, warm):
    import sys
    sys.path.insert(0, GOLD)
    import make_golden
    g = make_golden.solver_case(name, size, 12345, steps)
    if not warm:
        # recompute the reference iterates with warm starting off
        nb = len(g["pos"])
        vit = np.zeros((8, nb, 4), np.float32); pit = np.zeros((3, nb, 4), np.float32)
        pos_o, vel_o, man_o = g["pos"].copy(), g["vel"].copy(), g["manifold"].copy()
        done = C.c_int32()
        oracle_bindings.load_ref().b2ref_solve(nb, capi.fp(pos_o), capi.fp(vel_o), capi.fp(g["mass"]), len(g["index"]),
                                    capi.ip(g["index"]), capi.fp(man_o), capi.fp(g["material"]), capi.fp(g["radii"]),
                                    float(g["dt"]), 1.0, 0, 8, 3, capi.fp(vit), capi.fp(pit), C.byref(done))
        g.update(vel_iterates=vit, pos_iterates=pit, pos_out=pos_o, vel_out=vel_o, manifold_out=man_o,
                 pos_iters_done=np.int32(done.value))
    pos, vel, man, vit, pit, done = gpu_solve(g, warm=warm)
    assert done == int(g["pos_iters_done"])
    worst = max(max(rel_err(vit[k][:, :3], g["vel_iterates"][k][:, :3]) for k in range(8)),
                max(rel_err(pit[k][:, :3], g["pos_iterates"][k][:, :3]) for k in range(done)))
    assert worst <= TOL, worst
    assert impulse_err(man, g["manifold_out"]) <= TOL
    assert len(g["index"]) > 100
    print(f"{name}@{steps}: {len(g['index'])} constraints, worst relative iterate error {worst:g}")


def test_no_constraints_just_integrates():
    g = dict(pos=np.zeros((3, 4), np.float32), vel=np.array([[1, 2, 3, 0]] * 3, np.float32),
             mass=np.ones((3, 4), np.float32), index=np.zeros((0, 2), np.int32), manifold=np.zeros((0, 16), np.float32),
             material=np.zeros((0, 4), np.float32), radii=np.zeros((0, 2), np.float32), dt=np.float32(0.5))
    pos, vel, man, vit, pit, done = gpu_solve(g)
    assert done == 1  # contactsOkay on the first position iteration
    assert np.allclose(pos[:, :3], [[0.5, 1.0, 1.5]] * 3)

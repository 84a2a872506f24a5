"""Generates the committed golden vectors from the compiled reference (oracle/_ref/libb2ref.so).

Run HERE (where /root/reference exists and `make -C oracle ref` has been done):
    python tests/golden/make_golden.py
Writes small .npz fixtures next to this script.  They pin (a) the oracle restatement on CPU and
(b) the CUDA kernels on the GPU box even if oracle/_ref did not travel.
"""
import os
import sys

import numpy as np
import oracle.bindings as oracle_bindings  # noqa: E402  (the checker)

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

from box2d_optimized_b200 import capi  # noqa: E402
from oracle.bindings import RefScene
import util  # noqa: E402


def random_polygon(rng, lib):
    nv = int(rng.integers(3, 9))
    ang = np.sort(rng.uniform(0, 2 * np.pi, nv))
    rad = rng.uniform(0.3, 1.0, nv)
    pts = np.stack([rad * np.cos(ang), rad * np.sin(ang)], 1).astype(np.float32)
    pts += rng.uniform(-0.2, 0.2, 2).astype(np.float32)
    rec = np.zeros((9, 4), np.float32)
    cnt = lib.b2ref_polygon_set(capi.fp(pts), nv, capi.fp(rec))
    return rec[:1 + cnt]


def narrowphase_cases(seed, n_per_type):
    """random ordered pairs of every supported type, mostly near contact"""
    rng = np.random.default_rng(seed)
    lib = oracle_bindings.load_ref()
    pool = util.ShapePool()
    tA, oA, tB, oB, xa, xb = [], [], [], [], [], []
    types = [(0, 0), (2, 0), (2, 2), (1, 0), (1, 2)]
    for (ta, tb) in types:
        for _ in range(n_per_type):
            def make(t):
                if t == 0:
                    return pool.circle(rng.uniform(-0.2, 0.2, 2), float(rng.uniform(0.1, 0.8))), 0.6
                if t == 1:
                    L = float(rng.uniform(0.5, 3.0))
                    one = bool(rng.integers(0, 2))
                    v0 = (-L - float(rng.uniform(0.5, 2)), float(rng.uniform(-1, 1)))
                    v3 = (L + float(rng.uniform(0.5, 2)), float(rng.uniform(-1, 1)))
                    return pool.edge((-L, 0.0), (L, 0.0), v0, v3, one), 0.3
                if rng.random() < 0.3:
                    return pool.polygon_record(util.box_record(float(rng.uniform(0.2, 1)), float(rng.uniform(0.2, 1)))), 0.7
                return pool.polygon_record(random_polygon(rng, lib)), 0.7
            offa, ra = make(ta)
            offb, rb = make(tb)
            tA.append(ta); oA.append(offa); tB.append(tb); oB.append(offb)
            pa = rng.uniform(-50, 50, 2)
            d = rng.uniform(0.0, 1.15) * (ra + rb)
            th = rng.uniform(0, 2 * np.pi)
            pb = pa + d * np.array([np.cos(th), np.sin(th)])
            if ta == 1:  # keep the other shape near the segment
                pb = pa + np.array([rng.uniform(-3.5, 3.5), rng.uniform(-0.9, 0.9)])
            aa, ab = rng.uniform(-np.pi, np.pi, 2)
            if rng.random() < 0.25:
                aa = ab = 0.0  # axis-aligned stacks: the degenerate, tie-heavy case
            xa.append([pa[0], pa[1], aa]); xb.append([pb[0], pb[1], ab])
    xa = np.array(xa, np.float32); xb = np.array(xb, np.float32)
    xfA = util.xf_rows(xa[:, 0], xa[:, 1], xa[:, 2]); xfB = util.xf_rows(xb[:, 0], xb[:, 1], xb[:, 2])
    return dict(tA=np.array(tA, np.int32), oA=np.array(oA, np.int32), tB=np.array(tB, np.int32),
                oB=np.array(oB, np.int32), xfA=xfA, xfB=xfB, quads=pool.array())


def solver_case(scene_name, size, seed, steps):
    """solver inputs harvested from a live reference world + the reference's iterates"""
    s = RefScene(scene_name, size, seed)
    s.step(steps)
    s.collide_now()
    b = s.bodies(); p = s.body_params(); inv = s.body_inv(); fx = s.fixtures(); c = s.contacts()
    sel = (c["flags"] & 3) == 3
    sel &= util.man_count(c["manifold"]) > 0
    fa, fb = c["fix_a"][sel], c["fix_b"][sel]
    index = np.stack([fx["body"][fa], fx["body"][fb]], 1).astype(np.int32)
    radius = np.where(fx["type"] == 0, fx["quads"][fx["shape_off"], 2], np.float32(0.01)).astype(np.float32)
    radii = np.stack([radius[fa], radius[fb]], 1).astype(np.float32)
    nb = len(b)
    dt = np.float32(1.0 / 60.0)
    pos = np.zeros((nb, 4), np.float32); pos[:, 0:2] = b[:, 4:6]; pos[:, 2] = b[:, 6]
    vel = np.zeros((nb, 4), np.float32); vel[:, 0:2] = b[:, 7:9]; vel[:, 2] = b[:, 9]
    dyn = b[:, 11] == 2
    vel[dyn, 1] += dt * np.float32(-10.0)
    mass = np.zeros((nb, 4), np.float32); mass[:, 0:2] = inv; mass[:, 2:4] = p[:, 2:4]
    man = c["manifold"][sel].copy(); mat = c["material"][sel].copy()
    vi, pi = 8, 3
    vit = np.zeros((vi, nb, 4), np.float32); pit = np.zeros((pi, nb, 4), np.float32)
    pos_o, vel_o, man_o = pos.copy(), vel.copy(), man.copy()
    done = capi.C.c_int32() if hasattr(capi, "C") else None
    import ctypes
    done = ctypes.c_int32()
    oracle_bindings.load_ref().b2ref_solve(nb, capi.fp(pos_o), capi.fp(vel_o), capi.fp(mass), len(index), capi.ip(index),
                                capi.fp(man_o), capi.fp(mat), capi.fp(radii), float(dt), 1.0, 1, vi, pi,
                                capi.fp(vit), capi.fp(pit), ctypes.byref(done))
    return dict(pos=pos, vel=vel, mass=mass, index=index, manifold=man, material=mat, radii=radii, dt=dt,
                vel_iterates=vit, pos_iterates=pit, pos_out=pos_o, vel_out=vel_o, manifold_out=man_o,
                pos_iters_done=np.int32(done.value))


def scene_case(name, size, seed, steps):
    """body transforms, fixture records, tight AABBs, the contact pair list (ordered A,B) and the
    manifolds the reference computes for those transforms"""
    s = RefScene(name, size, seed)
    s.step(steps)
    s.collide_now()
    b = s.bodies(); fx = s.fixtures(); c = s.contacts()
    return dict(bodies=b, body_params=s.body_params(), fix_body=fx["body"], fix_type=fx["type"],
                fix_shape_off=fx["shape_off"], fix_filter=fx["filter"], fix_material=fx["material"],
                fix_sensor=fx["sensor"], quads=fx["quads"], aabbs=s.aabbs(), con_a=c["fix_a"], con_b=c["fix_b"],
                con_flags=c["flags"], con_manifold=c["manifold"])


def main():
    case = narrowphase_cases(2024, 400)
    m = util.ref_collide(case["tA"], case["oA"], case["xfA"], case["tB"], case["oB"], case["xfB"], case["quads"])
    np.savez_compressed(os.path.join(HERE, "narrowphase_random.npz"), manifold=m, **case)
    print("narrowphase_random:", len(m), "pairs,", int((util.man_count(m) > 0).sum()), "touching")
    for nm, sz, sd, st in (("pyramid", 20, 0, 45), ("mixed", 600, 12345, 150)):
        d = solver_case(nm, sz, sd, st)
        np.savez_compressed(os.path.join(HERE, f"solver_{nm}.npz"), **d)
        print(f"solver_{nm}: {len(d['index'])} constraints, {len(d['pos'])} bodies, pos iters {d['pos_iters_done']}")
        d = scene_case(nm, sz, sd, st)
        np.savez_compressed(os.path.join(HERE, f"scene_{nm}.npz"), **d)
        print(f"scene_{nm}: {len(d['con_a'])} contacts, {len(d['fix_body'])} fixtures")
    # the one fixed narrowphase input the reference's own tree contains (testbed/tests/polygon_collision.cpp:32-56)
    pool = util.ShapePool()
    oa = pool.polygon_record(util.box_record(0.2, 0.4)); ob = pool.polygon_record(util.box_record(0.5, 0.5))
    xfA = util.xf_rows([0.0], [0.0], [0.0]); xfB = util.xf_rows([19.345284], [1.5632932], [1.9160721])
    m = util.ref_collide([2], [oa], xfA, [2], [ob], xfB, pool.array())
    np.savez_compressed(os.path.join(HERE, "polygon_collision_testbed.npz"), manifold=m, quads=pool.array(), xfA=xfA,
                        xfB=xfB, oA=np.int32(oa), oB=np.int32(ob))


if __name__ == "__main__":
    main()

"""Long-horizon outcome check against the reference's CPU Step (config 3 semantics: run until
>= 95 % asleep or 3000 steps) and config 1/2 settle-to-sleep."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from box2d_optimized_b200 import GpuScene
from oracle.bindings import RefScene

def pe(b, p):
    dyn = b[:, 11] == 2
    return float(np.sum(p[dyn, 0] * 10.0 * b[dyn, 5]))

out = {}
for name, size, seed, maxsteps in (("mixed", 20000, 12345, 3000), ("many_pyramids", 10, 0, 1000)):
    g = GpuScene(name, size, seed); r = RefScene(name, size, seed)
    t0 = time.time(); tg = tr = 0.0
    steps = 0
    rec = []
    while steps < maxsteps:
        a = time.time(); g.step(100); gb = g.bodies(); tg += time.time() - a
        a = time.time(); r.step(100); rb = r.bodies(); tr += time.time() - a
        steps += 100
        dyn = gb[:, 11] == 2
        ag, ar = 1 - gb[dyn, 10].mean(), 1 - rb[dyn, 10].mean()
        rec.append((steps, float(ag), float(ar)))
        if ag >= 0.95 and ar >= 0.95:
            break
    gp, rp = g.body_params(), r.body_params()
    res = dict(steps=steps, asleep_gpu=rec[-1][1], asleep_ref=rec[-1][2], pe_gpu=pe(gb, gp), pe_ref=pe(rb, rp),
               finite=bool(np.isfinite(gb).all()), min_y=float(gb[gb[:, 11] == 2, 5].min()),
               contacts_gpu=g.contact_count, contacts_ref=r.contact_count, wall_gpu_s=tg, wall_ref_s=tr,
               sleep_curve=rec[::3])
    out[f"{name}_{size}"] = res
    print(name, size, json.dumps(res))
json.dump(out, open("gpurun_out/long_validation.json", "w"), indent=1)

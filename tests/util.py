"""Shared helpers for the parity tests: shape-pool builders and manifold field access."""
import numpy as np
import oracle.bindings as oracle_bindings  # noqa: E402  (the checker)

from box2d_optimized_b200 import capi


def man_ids(m):
    m = np.ascontiguousarray(m, np.float32).reshape(-1, 16)
    return m[:, 12:14].copy().view(np.uint32)


def man_type(m):
    return np.ascontiguousarray(m, np.float32).reshape(-1, 16)[:, 14].copy().view(np.int32)


def man_count(m):
    return np.ascontiguousarray(m, np.float32).reshape(-1, 16)[:, 15].copy().view(np.int32)


def pair_set(fa, fb):
    a = np.minimum(fa, fb).astype(np.int64)
    b = np.maximum(fa, fb).astype(np.int64)
    return set(zip(a.tolist(), b.tolist()))


class ShapePool:
    """Builds the float4 shape pool of include/b2cuda.h from python descriptions."""

    def __init__(self):
        self.quads = []
        self.n = 0

    def _add(self, rows):
        off = self.n
        self.quads.extend(rows)
        self.n += len(rows)
        return off

    def circle(self, p, r):
        return self._add([[p[0], p[1], r, 0.0]])

    def edge(self, v1, v2, v0=(0, 0), v3=(0, 0), one_sided=False, radius=0.01):
        return self._add([[v1[0], v1[1], v2[0], v2[1]], [v0[0], v0[1], v3[0], v3[1]],
                          [radius, 1.0 if one_sided else 0.0, 0.0, 0.0]])

    def polygon_record(self, rec):
        """rec = [1 + count, 4] array as produced by b2ref_polygon_set or box()"""
        return self._add([list(map(float, r)) for r in rec])

    def array(self):
        return np.array(self.quads if self.quads else [[0, 0, 0, 0]], np.float32)


def box_record(hx, hy, radius=0.01):
    verts = [(-hx, -hy), (hx, -hy), (hx, hy), (-hx, hy)]
    normals = [(0, -1), (1, 0), (0, 1), (-1, 0)]
    rec = [[0.0, 0.0, radius, 4.0]]
    for v, n in zip(verts, normals):
        rec.append([v[0], v[1], n[0], n[1]])
    return np.array(rec, np.float32)


def xf_rows(px, py, angle):
    """host-computed transforms (float32 sin/cos of numpy == glibc for the test's purposes: both
    sides receive THESE bits, so the comparison does not depend on whose sinf is used)"""
    a = np.asarray(angle, np.float32)
    return np.stack([np.asarray(px, np.float32), np.asarray(py, np.float32), np.sin(a).astype(np.float32),
                     np.cos(a).astype(np.float32)], 1)


def gpu_collide(tA, oA, xA, tB, oB, xB, quads, device=0):
    n = len(tA)
    out = np.zeros((n, 16), np.float32)
    tA, oA, tB, oB = map(capi.i32, (tA, oA, tB, oB))
    xA, xB, quads = capi.f32(xA), capi.f32(xB), capi.f32(quads)
    capi.check(capi.load_cuda().b2g_collide_pairs(device, n, capi.ip(tA), capi.ip(oA), capi.fp(xA), capi.ip(tB),
                                                  capi.ip(oB), capi.fp(xB), capi.fp(quads), len(quads),
                                                  capi.fp(out)), "b2g_collide_pairs")
    return out


def ref_collide(tA, oA, xA, tB, oB, xB, quads):
    n = len(tA)
    out = np.zeros((n, 16), np.float32)
    tA, oA, tB, oB = map(capi.i32, (tA, oA, tB, oB))
    xA, xB, quads = capi.f32(xA), capi.f32(xB), capi.f32(quads)
    rc = oracle_bindings.load_ref().b2ref_collide_pairs(n, capi.ip(tA), capi.ip(oA), capi.fp(xA), capi.ip(tB), capi.ip(oB),
                                             capi.fp(xB), capi.fp(quads), capi.fp(out))
    assert rc == 0
    return out


def compare_manifolds(g, r, rel=1e-5):
    """bit-exact on integer fields (pointCount, type, feature ids), `rel` on points and normals.
    Returns (number compared, max abs float difference)."""
    cg, cr = man_count(g), man_count(r)
    assert np.array_equal(cg, cr), f"pointCount differs at {np.nonzero(cg != cr)[0][:10]}"
    touching = cr > 0
    assert np.array_equal(man_type(g)[touching], man_type(r)[touching]), "manifold type differs"
    ig, ir = man_ids(g), man_ids(r)
    assert np.array_equal(ig[touching, 0], ir[touching, 0]), "feature id of point 0 differs"
    two = cr > 1
    assert np.array_equal(ig[two, 1], ir[two, 1]), "feature id of point 1 differs"
    maxd = 0.0
    # localNormal, localPoint, points[0].localPoint
    for cols, mask in (((0, 1, 2, 3, 4, 5), touching), ((8, 9), two)):
        a = g[mask][:, cols]
        b = r[mask][:, cols]
        if a.size:
            d = np.abs(a - b)
            tol = rel * np.maximum(1.0, np.abs(b))
            assert (d <= tol).all(), f"manifold floats differ: max {d.max()}"
            maxd = max(maxd, float(d.max()))
    return int(touching.sum()), maxd

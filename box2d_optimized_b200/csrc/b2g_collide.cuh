// b2g_collide.cuh — narrowphase manifold generation on the device.
//
// One call of collide_dispatch() replaces b2Contact::Evaluate, i.e. the 4x4 function table
// of src/dynamics/b2_contact.cpp:46-56 and the five b2Collide* functions behind it:
//   circle-circle   src/collision/b2_collide_circle.cpp:27-53
//   polygon-circle  src/collision/b2_collide_circle.cpp:55-158
//   polygon-polygon src/collision/b2_collide_polygon.cpp:27-243
//   edge-circle     src/collision/b2_collide_edge.cpp:31-165
//   edge-polygon    src/collision/b2_collide_edge.cpp:167-524
// plus b2ClipSegmentToLine (src/collision/b2_collision.cpp:205-237) and
// b2WorldManifold::Initialize (src/collision/b2_collision.cpp:26-90).
//
// Shapes are read from the float4 shape pool described in include/b2cuda.h.  Feature ids are
// bit-exact with the reference (integer work); points/normals follow the reference's
// operation order so they agree to the last bit when sin/cos of the transforms are equal.
#pragma once
#include "b2g_math.cuh"

#define B2G_FEATURE_VERTEX 0u
#define B2G_FEATURE_FACE 1u

// b2ContactID (b2_collision.h:44-68): little-endian {indexA, indexB, typeA, typeB}
B2G_HD uint32_t feature_key(uint32_t indexA, uint32_t indexB, uint32_t typeA, uint32_t typeB) {
  return (indexA & 0xffu) | ((indexB & 0xffu) << 8) | ((typeA & 0xffu) << 16) | ((typeB & 0xffu) << 24);
}
B2G_HD uint32_t feature_swap(uint32_t key) {
  // swap A <-> B features
  uint32_t iA = key & 0xffu, iB = (key >> 8) & 0xffu, tA = (key >> 16) & 0xffu, tB = (key >> 24) & 0xffu;
  return feature_key(iB, iA, tB, tA);
}

// b2Manifold (b2_collision.h:100-117)
struct Manifold {
  float2 localNormal;
  float2 localPoint;
  float2 lp[2];        // points[i].localPoint
  float normalImp[2];  // points[i].normalImpulse
  float tangentImp[2]; // points[i].tangentImpulse
  uint32_t id[2];      // points[i].id.key
  int type;
  int pointCount;
};

// 64-byte packed form: 4 x float4, see b2gContactArrays.manifold in include/b2cuda.h
B2G_HD void manifold_pack(const Manifold& m, float4& q0, float4& q1, float4& q2, float4& q3) {
  q0 = make_float4(m.localNormal.x, m.localNormal.y, m.localPoint.x, m.localPoint.y);
  q1 = make_float4(m.lp[0].x, m.lp[0].y, m.normalImp[0], m.tangentImp[0]);
  q2 = make_float4(m.lp[1].x, m.lp[1].y, m.normalImp[1], m.tangentImp[1]);
#ifdef __CUDA_ARCH__
  q3 = make_float4(__uint_as_float(m.id[0]), __uint_as_float(m.id[1]), __int_as_float(m.type),
                   __int_as_float(m.pointCount));
#else
  union {
    uint32_t u;
    float f;
  } a, b, c, d;
  a.u = m.id[0];
  b.u = m.id[1];
  c.u = (uint32_t)m.type;
  d.u = (uint32_t)m.pointCount;
  q3 = make_float4(a.f, b.f, c.f, d.f);
#endif
}
#ifdef __CUDACC__
__device__ __forceinline__ void manifold_unpack(Manifold& m, float4 q0, float4 q1, float4 q2, float4 q3) {
  m.localNormal = make_float2(q0.x, q0.y);
  m.localPoint = make_float2(q0.z, q0.w);
  m.lp[0] = make_float2(q1.x, q1.y);
  m.normalImp[0] = q1.z;
  m.tangentImp[0] = q1.w;
  m.lp[1] = make_float2(q2.x, q2.y);
  m.normalImp[1] = q2.z;
  m.tangentImp[1] = q2.w;
  m.id[0] = __float_as_uint(q3.x);
  m.id[1] = __float_as_uint(q3.y);
  m.type = __float_as_int(q3.z);
  m.pointCount = __float_as_int(q3.w);
}

// ---- shape records -------------------------------------------------------------------
struct Circle {
  float2 p;
  float radius;
};
struct Edge {
  float2 v0, v1, v2, v3;
  float radius;
  bool oneSided;
};
struct Poly {
  float2 v[B2G_MAX_POLY_VERTS];
  float2 n[B2G_MAX_POLY_VERTS];
  float2 centroid;
  float radius;
  int count;
};

__device__ __forceinline__ Circle load_circle(const float4* __restrict__ pool, int off) {
  float4 a = __ldg(pool + off);
  Circle c;
  c.p = make_float2(a.x, a.y);
  c.radius = a.z;
  return c;
}
__device__ __forceinline__ Edge load_edge(const float4* __restrict__ pool, int off) {
  float4 a = __ldg(pool + off), b = __ldg(pool + off + 1), c = __ldg(pool + off + 2);
  Edge e;
  e.v1 = make_float2(a.x, a.y);
  e.v2 = make_float2(a.z, a.w);
  e.v0 = make_float2(b.x, b.y);
  e.v3 = make_float2(b.z, b.w);
  e.radius = c.x;
  e.oneSided = c.y != 0.0f;
  return e;
}
__device__ __forceinline__ void load_poly(Poly& P, const float4* __restrict__ pool, int off) {
  float4 h = __ldg(pool + off);
  P.centroid = make_float2(h.x, h.y);
  P.radius = h.z;
  P.count = (int)h.w;
#pragma unroll
  for (int i = 0; i < B2G_MAX_POLY_VERTS; ++i) {
    if (i < P.count) {
      float4 q = __ldg(pool + off + 1 + i);
      P.v[i] = make_float2(q.x, q.y);
      P.n[i] = make_float2(q.z, q.w);
    }
  }
}

// ---- AABBs (b2_polygon_shape.cpp:373-388, b2_circle_shape.cpp:91-96, b2_edge_shape.cpp:156-167)
__device__ __forceinline__ float4 shape_aabb(const float4* __restrict__ pool, int type, int off, Xf xf) {
  float2 lower, upper;
  float r;
  if (type == 0) {
    Circle c = load_circle(pool, off);
    float2 p = xf.p + rot_mul(xf.q, c.p);
    return make_float4(p.x - c.radius, p.y - c.radius, p.x + c.radius, p.y + c.radius);
  } else if (type == 1) {
    float4 a = __ldg(pool + off);
    float4 c = __ldg(pool + off + 2);
    float2 v1 = xf_mul(xf, make_float2(a.x, a.y));
    float2 v2 = xf_mul(xf, make_float2(a.z, a.w));
    lower = make_float2(minf_(v1.x, v2.x), minf_(v1.y, v2.y));
    upper = make_float2(maxf_(v1.x, v2.x), maxf_(v1.y, v2.y));
    r = c.x;
  } else {
    float4 h = __ldg(pool + off);
    int count = (int)h.w;
    r = h.z;
    float4 q = __ldg(pool + off + 1);
    lower = xf_mul(xf, make_float2(q.x, q.y));
    upper = lower;
    for (int i = 1; i < count; ++i) {
      q = __ldg(pool + off + 1 + i);
      float2 v = xf_mul(xf, make_float2(q.x, q.y));
      lower = make_float2(minf_(lower.x, v.x), minf_(lower.y, v.y));
      upper = make_float2(maxf_(upper.x, v.x), maxf_(upper.y, v.y));
    }
  }
  return make_float4(lower.x - r, lower.y - r, upper.x + r, upper.y + r);
}

// ---- circle vs circle ------------------------------------------------------------------
__device__ __forceinline__ void collide_circles(Manifold& m, Circle A, Xf xfA, Circle B, Xf xfB) {
  m.pointCount = 0;
  float2 pA = xf_mul(xfA, A.p);
  float2 pB = xf_mul(xfB, B.p);
  float2 d = pB - pA;
  float distSqr = dot2(d, d);
  float radius = A.radius + B.radius;
  if (distSqr > radius * radius) return;
  m.type = 0;
  m.localPoint = A.p;
  m.localNormal = make_float2(0.0f, 0.0f);
  m.pointCount = 1;
  m.lp[0] = B.p;
  m.id[0] = 0;
}

// ---- polygon vs circle -----------------------------------------------------------------
__device__ __forceinline__ void collide_polygon_circle(Manifold& m, const Poly& A, Xf xfA, Circle B, Xf xfB) {
  m.pointCount = 0;
  float2 c = xf_mul(xfB, B.p);
  float2 cLocal = xf_mulT(xfA, c);

  int normalIndex = 0;
  float separation = -B2G_MAX_FLOAT;
  float radius = A.radius + B.radius;
  for (int i = 0; i < A.count; ++i) {
    float s = dot2(A.n[i], cLocal - A.v[i]);
    if (s > radius) return;
    if (s > separation) {
      separation = s;
      normalIndex = i;
    }
  }
  int vertIndex1 = normalIndex;
  int vertIndex2 = vertIndex1 + 1 < A.count ? vertIndex1 + 1 : 0;
  float2 v1 = A.v[vertIndex1];
  float2 v2 = A.v[vertIndex2];

  m.type = 1;
  m.lp[0] = B.p;
  m.id[0] = 0;
  if (separation < B2G_EPSILON) {
    // centre inside the polygon
    m.pointCount = 1;
    m.localNormal = A.n[normalIndex];
    m.localPoint = 0.5f * (v1 + v2);
    return;
  }
  float u1 = dot2(cLocal - v1, v2 - v1);
  float u2 = dot2(cLocal - v2, v1 - v2);
  if (u1 <= 0.0f) {
    if (dist_sq(cLocal, v1) > radius * radius) return;
    m.pointCount = 1;
    m.localNormal = cLocal - v1;
    normalize2(m.localNormal);
    m.localPoint = v1;
  } else if (u2 <= 0.0f) {
    if (dist_sq(cLocal, v2) > radius * radius) return;
    m.pointCount = 1;
    m.localNormal = cLocal - v2;
    normalize2(m.localNormal);
    m.localPoint = v2;
  } else {
    float2 faceCenter = 0.5f * (v1 + v2);
    float s = dot2(cLocal - faceCenter, A.n[vertIndex1]);
    if (s > radius) return;
    m.pointCount = 1;
    m.localNormal = A.n[vertIndex1];
    m.localPoint = faceCenter;
  }
}

// ---- Sutherland-Hodgman clip of a 2-vertex segment (b2_collision.cpp:205-237) -----------
struct ClipVertex {
  float2 v;
  uint32_t id;
};
__device__ __forceinline__ int clip_segment(ClipVertex out[2], const ClipVertex in[2], float2 normal, float offset,
                                            int vertexIndexA) {
  int count = 0;
  float distance0 = dot2(normal, in[0].v) - offset;
  float distance1 = dot2(normal, in[1].v) - offset;
  if (distance0 <= 0.0f) out[count++] = in[0];
  if (distance1 <= 0.0f) out[count++] = in[1];
  if (distance0 * distance1 < 0.0f) {
    float interp = distance0 / (distance0 - distance1);
    out[count].v = in[0].v + interp * (in[1].v - in[0].v);
    // vertexA is hitting edgeB
    out[count].id = feature_key((uint32_t)vertexIndexA, (in[0].id >> 8) & 0xffu, B2G_FEATURE_VERTEX, B2G_FEATURE_FACE);
    ++count;
  }
  return count;
}

// ---- polygon vs polygon ----------------------------------------------------------------
// max over poly1's edge normals of the min over poly2's vertices (SAT), b2_collide_polygon.cpp:27-66
__device__ __forceinline__ float poly_max_separation(int& edgeIndex, const Poly& p1, Xf xf1, const Poly& p2, Xf xf2) {
  Xf xf = xf_mulT(xf2, xf1);
  int bestIndex = 0;
  float maxSeparation = -B2G_MAX_FLOAT;
  for (int i = 0; i < p1.count; ++i) {
    float2 n = rot_mul(xf.q, p1.n[i]);
    float2 v1 = xf_mul(xf, p1.v[i]);
    float si = B2G_MAX_FLOAT;
    for (int j = 0; j < p2.count; ++j) {
      float sij = dot2(n, p2.v[j] - v1);
      if (sij < si) si = sij;
    }
    if (si > maxSeparation) {
      maxSeparation = si;
      bestIndex = i;
    }
  }
  edgeIndex = bestIndex;
  return maxSeparation;
}

__device__ __forceinline__ void collide_polygons(Manifold& m, const Poly& A, Xf xfA, const Poly& B, Xf xfB) {
  m.pointCount = 0;
  float totalRadius = A.radius + B.radius;

  int edgeA = 0;
  float separationA = poly_max_separation(edgeA, A, xfA, B, xfB);
  if (separationA > totalRadius) return;
  int edgeB = 0;
  float separationB = poly_max_separation(edgeB, B, xfB, A, xfA);
  if (separationB > totalRadius) return;

  const float k_tol = 0.1f * B2G_LINEAR_SLOP;
  const bool flip = separationB > separationA + k_tol;
  const Poly& p1 = flip ? B : A;  // reference polygon
  const Poly& p2 = flip ? A : B;  // incident polygon
  Xf xf1 = flip ? xfB : xfA;
  Xf xf2 = flip ? xfA : xfB;
  int edge1 = flip ? edgeB : edgeA;
  m.type = flip ? 2 : 1;

  // incident edge on poly2: most anti-parallel normal (b2_collide_polygon.cpp:68-111)
  ClipVertex incident[2];
  {
    float2 normal1 = rot_mulT(xf2.q, rot_mul(xf1.q, p1.n[edge1]));
    int index = 0;
    float minDot = B2G_MAX_FLOAT;
    for (int i = 0; i < p2.count; ++i) {
      float d = dot2(normal1, p2.n[i]);
      if (d < minDot) {
        minDot = d;
        index = i;
      }
    }
    int i1 = index;
    int i2 = i1 + 1 < p2.count ? i1 + 1 : 0;
    incident[0].v = xf_mul(xf2, p2.v[i1]);
    incident[0].id = feature_key((uint32_t)edge1, (uint32_t)i1, B2G_FEATURE_FACE, B2G_FEATURE_VERTEX);
    incident[1].v = xf_mul(xf2, p2.v[i2]);
    incident[1].id = feature_key((uint32_t)edge1, (uint32_t)i2, B2G_FEATURE_FACE, B2G_FEATURE_VERTEX);
  }

  int iv1 = edge1;
  int iv2 = edge1 + 1 < p1.count ? edge1 + 1 : 0;
  float2 v11 = p1.v[iv1];
  float2 v12 = p1.v[iv2];

  float2 localTangent = v12 - v11;
  normalize2(localTangent);
  float2 localNormal = cross_vs(localTangent, 1.0f);
  float2 planePoint = 0.5f * (v11 + v12);

  float2 tangent = rot_mul(xf1.q, localTangent);
  float2 normal = cross_vs(tangent, 1.0f);

  v11 = xf_mul(xf1, v11);
  v12 = xf_mul(xf1, v12);

  float frontOffset = dot2(normal, v11);
  float sideOffset1 = -dot2(tangent, v11) + totalRadius;
  float sideOffset2 = dot2(tangent, v12) + totalRadius;

  ClipVertex clip1[2], clip2[2];
  int np = clip_segment(clip1, incident, -tangent, sideOffset1, iv1);
  if (np < 2) return;
  np = clip_segment(clip2, clip1, tangent, sideOffset2, iv2);
  if (np < 2) return;

  m.localNormal = localNormal;
  m.localPoint = planePoint;
  int pointCount = 0;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    float separation = dot2(normal, clip2[i].v) - frontOffset;
    if (separation <= totalRadius) {
      m.lp[pointCount] = xf_mulT(xf2, clip2[i].v);
      m.id[pointCount] = flip ? feature_swap(clip2[i].id) : clip2[i].id;
      ++pointCount;
    }
  }
  m.pointCount = pointCount;
}

// ---- edge vs circle --------------------------------------------------------------------
__device__ __forceinline__ void collide_edge_circle(Manifold& m, const Edge& E, Xf xfA, Circle B, Xf xfB) {
  m.pointCount = 0;
  float2 Q = xf_mulT(xfA, xf_mul(xfB, B.p));
  float2 A = E.v1, Bv = E.v2;
  float2 e = Bv - A;
  float2 n = make_float2(e.y, -e.x);
  float offset = dot2(n, Q - A);
  if (E.oneSided && offset < 0.0f) return;

  float u = dot2(e, Bv - Q);
  float v = dot2(e, Q - A);
  float radius = E.radius + B.radius;

  if (v <= 0.0f) {  // region A
    float2 d = Q - A;
    if (dot2(d, d) > radius * radius) return;
    if (E.oneSided) {
      float2 e1 = A - E.v0;
      float u1 = dot2(e1, A - Q);
      if (u1 > 0.0f) return;
    }
    m.pointCount = 1;
    m.type = 0;
    m.localNormal = make_float2(0.0f, 0.0f);
    m.localPoint = A;
    m.id[0] = feature_key(0, 0, B2G_FEATURE_VERTEX, B2G_FEATURE_VERTEX);
    m.lp[0] = B.p;
    return;
  }
  if (u <= 0.0f) {  // region B
    float2 d = Q - Bv;
    if (dot2(d, d) > radius * radius) return;
    if (E.oneSided) {
      float2 e2 = E.v3 - Bv;
      float v2_ = dot2(e2, Q - Bv);
      if (v2_ > 0.0f) return;
    }
    m.pointCount = 1;
    m.type = 0;
    m.localNormal = make_float2(0.0f, 0.0f);
    m.localPoint = Bv;
    m.id[0] = feature_key(1, 0, B2G_FEATURE_VERTEX, B2G_FEATURE_VERTEX);
    m.lp[0] = B.p;
    return;
  }
  // region AB
  float den = dot2(e, e);
  float2 P = (1.0f / den) * (u * A + v * Bv);
  float2 d = Q - P;
  if (dot2(d, d) > radius * radius) return;
  if (offset < 0.0f) n = make_float2(-n.x, -n.y);
  normalize2(n);
  m.pointCount = 1;
  m.type = 1;
  m.localNormal = n;
  m.localPoint = A;
  m.id[0] = feature_key(0, 0, B2G_FEATURE_FACE, B2G_FEATURE_VERTEX);
  m.lp[0] = B.p;
}

// ---- edge vs polygon (b2_collide_edge.cpp:167-524) --------------------------------------
__device__ __forceinline__ void collide_edge_polygon(Manifold& m, const Edge& E, Xf xfA, const Poly& B, Xf xfB) {
  m.pointCount = 0;
  Xf xf = xf_mulT(xfA, xfB);
  float2 centroidB = xf_mul(xf, B.centroid);
  float2 v1 = E.v1, v2 = E.v2;
  float2 edge1 = v2 - v1;
  normalize2(edge1);
  float2 normal1 = make_float2(edge1.y, -edge1.x);
  float offset1 = dot2(normal1, centroidB - v1);
  if (E.oneSided && offset1 < 0.0f) return;

  // polygon B in the edge frame
  float2 tv[B2G_MAX_POLY_VERTS], tn[B2G_MAX_POLY_VERTS];
  const int count = B.count;
  for (int i = 0; i < count; ++i) {
    tv[i] = xf_mul(xf, B.v[i]);
    tn[i] = rot_mul(xf.q, B.n[i]);
  }
  float radius = B.radius + E.radius;

  // separating axis candidates: +-edge normal (min-max), then polygon normals
  int edgeAxisIndex = -1;
  float edgeSep = -B2G_MAX_FLOAT;
  float2 edgeNormal = make_float2(0.0f, 0.0f);
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    float2 axis = j == 0 ? normal1 : -normal1;
    float sj = B2G_MAX_FLOAT;
    for (int i = 0; i < count; ++i) {
      float si = dot2(axis, tv[i] - v1);
      if (si < sj) sj = si;
    }
    if (sj > edgeSep) {
      edgeAxisIndex = j;
      edgeSep = sj;
      edgeNormal = axis;
    }
  }
  (void)edgeAxisIndex;
  if (edgeSep > radius) return;

  int polyIndex = -1;
  float polySep = -B2G_MAX_FLOAT;
  float2 polyNormal = make_float2(0.0f, 0.0f);
  bool polyValid = false;
  for (int i = 0; i < count; ++i) {
    float2 n = -tn[i];
    float s1 = dot2(n, tv[i] - v1);
    float s2 = dot2(n, tv[i] - v2);
    float s = minf_(s1, s2);
    if (s > polySep) {
      polyValid = true;
      polyIndex = i;
      polySep = s;
      polyNormal = n;
    }
  }
  if (polySep > radius) return;

  // hysteresis for jitter reduction
  const float k_relativeTol = 0.98f;
  const float k_absoluteTol = 0.001f;
  bool usePoly = (polySep - radius > k_relativeTol * (edgeSep - radius) + k_absoluteTol);
  // primary axis; type "unknown" (no polygon axis found) behaves like edgeB in the reference's
  // else-branch below, which cannot happen for count >= 1 — polyValid keeps that explicit.
  bool primaryIsEdgeA = !usePoly;
  float2 primaryNormal = usePoly ? polyNormal : edgeNormal;
  int primaryIndex = usePoly ? polyIndex : 0;
  (void)polyValid;

  if (E.oneSided) {
    // smooth collision via the Gauss map (ghost vertices)
    float2 edge0 = v1 - E.v0;
    normalize2(edge0);
    float2 normal0 = make_float2(edge0.y, -edge0.x);
    bool convex1 = cross2(edge0, edge1) >= 0.0f;

    float2 edge2 = E.v3 - v2;
    normalize2(edge2);
    float2 normal2 = make_float2(edge2.y, -edge2.x);
    bool convex2 = cross2(edge1, edge2) >= 0.0f;

    const float sinTol = 0.1f;
    bool side1 = dot2(primaryNormal, edge1) <= 0.0f;
    if (side1) {
      if (convex1) {
        if (cross2(primaryNormal, normal0) > sinTol) return;  // skip region
      } else {
        primaryIsEdgeA = true;  // snap region
        primaryNormal = edgeNormal;
        primaryIndex = 0;
      }
    } else {
      if (convex2) {
        if (cross2(normal2, primaryNormal) > sinTol) return;
      } else {
        primaryIsEdgeA = true;
        primaryNormal = edgeNormal;
        primaryIndex = 0;
      }
    }
  }

  ClipVertex clipPoints[2];
  int ref_i1, ref_i2;
  float2 ref_v1, ref_v2, ref_normal, sideNormal1, sideNormal2;
  if (primaryIsEdgeA) {
    m.type = 1;
    int bestIndex = 0;
    float bestValue = dot2(primaryNormal, tn[0]);
    for (int i = 1; i < count; ++i) {
      float value = dot2(primaryNormal, tn[i]);
      if (value < bestValue) {
        bestValue = value;
        bestIndex = i;
      }
    }
    int i1 = bestIndex;
    int i2 = i1 + 1 < count ? i1 + 1 : 0;
    clipPoints[0].v = tv[i1];
    clipPoints[0].id = feature_key(0, (uint32_t)i1, B2G_FEATURE_FACE, B2G_FEATURE_VERTEX);
    clipPoints[1].v = tv[i2];
    clipPoints[1].id = feature_key(0, (uint32_t)i2, B2G_FEATURE_FACE, B2G_FEATURE_VERTEX);
    ref_i1 = 0;
    ref_i2 = 1;
    ref_v1 = v1;
    ref_v2 = v2;
    ref_normal = primaryNormal;
    sideNormal1 = -edge1;
    sideNormal2 = edge1;
  } else {
    m.type = 2;
    clipPoints[0].v = v2;
    clipPoints[0].id = feature_key(1, (uint32_t)primaryIndex, B2G_FEATURE_VERTEX, B2G_FEATURE_FACE);
    clipPoints[1].v = v1;
    clipPoints[1].id = feature_key(0, (uint32_t)primaryIndex, B2G_FEATURE_VERTEX, B2G_FEATURE_FACE);
    ref_i1 = primaryIndex;
    ref_i2 = ref_i1 + 1 < count ? ref_i1 + 1 : 0;
    ref_v1 = tv[ref_i1];
    ref_v2 = tv[ref_i2];
    ref_normal = tn[ref_i1];
    sideNormal1 = make_float2(ref_normal.y, -ref_normal.x);
    sideNormal2 = -sideNormal1;
  }
  float sideOffset1 = dot2(sideNormal1, ref_v1);
  float sideOffset2 = dot2(sideNormal2, ref_v2);

  ClipVertex clip1[2], clip2[2];
  int np = clip_segment(clip1, clipPoints, sideNormal1, sideOffset1, ref_i1);
  if (np < 2) return;
  np = clip_segment(clip2, clip1, sideNormal2, sideOffset2, ref_i2);
  if (np < 2) return;

  if (primaryIsEdgeA) {
    m.localNormal = ref_normal;
    m.localPoint = ref_v1;
  } else {
    m.localNormal = B.n[ref_i1];
    m.localPoint = B.v[ref_i1];
  }
  int pointCount = 0;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    float separation = dot2(ref_normal, clip2[i].v - ref_v1);
    if (separation <= radius) {
      if (primaryIsEdgeA) {
        m.lp[pointCount] = xf_mulT(xf, clip2[i].v);
        m.id[pointCount] = clip2[i].id;
      } else {
        m.lp[pointCount] = clip2[i].v;
        m.id[pointCount] = feature_swap(clip2[i].id);
      }
      ++pointCount;
    }
  }
  m.pointCount = pointCount;
}

// ---- dispatch: b2Contact::functions[typeA][typeB] (b2_contact.cpp:46-56) ----------------
// Returns false when the reference has no function for the ordered pair (e.g. edge-edge).
__device__ __forceinline__ void manifold_clear(Manifold& m) {
  m.pointCount = 0;
  m.type = 0;
  m.localNormal = make_float2(0.0f, 0.0f);
  m.localPoint = make_float2(0.0f, 0.0f);
  m.lp[0] = m.lp[1] = make_float2(0.0f, 0.0f);
  m.normalImp[0] = m.normalImp[1] = 0.0f;
  m.tangentImp[0] = m.tangentImp[1] = 0.0f;
  m.id[0] = m.id[1] = 0;
}
__device__ __forceinline__ bool collide_dispatch(Manifold& m, const float4* __restrict__ pool, int typeA, int offA,
                                                 Xf xfA, int typeB, int offB, Xf xfB) {
  manifold_clear(m);
  if (typeA == 0 && typeB == 0) {
    collide_circles(m, load_circle(pool, offA), xfA, load_circle(pool, offB), xfB);
  } else if (typeA == 2 && typeB == 0) {
    Poly A;
    load_poly(A, pool, offA);
    collide_polygon_circle(m, A, xfA, load_circle(pool, offB), xfB);
  } else if (typeA == 2 && typeB == 2) {
    Poly A, B;
    load_poly(A, pool, offA);
    load_poly(B, pool, offB);
    collide_polygons(m, A, xfA, B, xfB);
  } else if (typeA == 1 && typeB == 0) {
    collide_edge_circle(m, load_edge(pool, offA), xfA, load_circle(pool, offB), xfB);
  } else if (typeA == 1 && typeB == 2) {
    Poly B;
    load_poly(B, pool, offB);
    collide_edge_polygon(m, load_edge(pool, offA), xfA, B, xfB);
  } else {
    return false;
  }
  return true;
}

// ---- world manifold (b2_collision.cpp:26-90) --------------------------------------------
struct WorldManifold {
  float2 normal;
  float2 points[2];
  float separations[2];
};
__device__ __forceinline__ void world_manifold(WorldManifold& w, const Manifold& m, Xf xfA, float radiusA, Xf xfB,
                                               float radiusB) {
  if (m.pointCount == 0) return;
  if (m.type == 0) {
    w.normal = make_float2(1.0f, 0.0f);
    float2 pointA = xf_mul(xfA, m.localPoint);
    float2 pointB = xf_mul(xfB, m.lp[0]);
    if (dist_sq(pointA, pointB) > B2G_EPSILON * B2G_EPSILON) {
      w.normal = pointB - pointA;
      normalize2(w.normal);
    }
    float2 cA = pointA + radiusA * w.normal;
    float2 cB = pointB - radiusB * w.normal;
    w.points[0] = 0.5f * (cA + cB);
    w.separations[0] = dot2(cB - cA, w.normal);
  } else if (m.type == 1) {
    w.normal = rot_mul(xfA.q, m.localNormal);
    float2 planePoint = xf_mul(xfA, m.localPoint);
    for (int i = 0; i < m.pointCount; ++i) {
      float2 clipPoint = xf_mul(xfB, m.lp[i]);
      float2 cA = clipPoint + (radiusA - dot2(clipPoint - planePoint, w.normal)) * w.normal;
      float2 cB = clipPoint - radiusB * w.normal;
      w.points[i] = 0.5f * (cA + cB);
      w.separations[i] = dot2(cB - cA, w.normal);
    }
  } else {
    w.normal = rot_mul(xfB.q, m.localNormal);
    float2 planePoint = xf_mul(xfB, m.localPoint);
    for (int i = 0; i < m.pointCount; ++i) {
      float2 clipPoint = xf_mul(xfA, m.lp[i]);
      float2 cB = clipPoint + (radiusB - dot2(clipPoint - planePoint, w.normal)) * w.normal;
      float2 cA = clipPoint - radiusA * w.normal;
      w.points[i] = 0.5f * (cA + cB);
      w.separations[i] = dot2(cA - cB, w.normal);
    }
    w.normal = -w.normal;
  }
}
#endif  // __CUDACC__

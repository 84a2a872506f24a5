// -----------------------------------------------------------------------------------------------
// Third-party notice.  To stay source- and result-compatible with box2d-optimized, parts of this
// file restate declarations, inline math and creation-time algorithms of that library (itself a
// fork of Box2D).  Those parts are covered by the MIT License:
//   Copyright (c) 2019 Erin Catto, Copyright (c) 2020 Manolis Tsamis
// The full licence text and permission notice are in LICENSES/box2d-optimized-MIT.txt.
// -----------------------------------------------------------------------------------------------
// b2_shapes.cpp — host-side shape geometry of the drop-in C++ API.
//
// Same results as the reference's src/collision/b2_circle_shape.cpp:91-105,
// b2_edge_shape.cpp:27-176, b2_polygon_shape.cpp:36-468 and b2_collision.cpp:26-134 for the
// functions a scene builder touches: hull construction, mass properties, tight AABBs, and the
// world manifold.  These run once per shape at creation time; the per-step work is on the device.
#include <climits>
#include "box2d/b2g_shapes.h"

const b2Vec2 b2Vec2_zero(0.0f, 0.0f);

// ---- circle -----------------------------------------------------------------------------------
bool b2CircleShape::TestPoint(const b2Transform& xf, const b2Vec2& p) const {
  b2Vec2 center = xf.p + b2Mul(xf.q, m_p);
  b2Vec2 d = p - center;
  return b2Dot(d, d) <= m_radius * m_radius;
}
void b2CircleShape::ComputeAABB(b2AABB* aabb, const b2Transform& xf) const {
  b2Vec2 p = xf.p + b2Mul(xf.q, m_p);
  aabb->lowerBound.Set(p.x - m_radius, p.y - m_radius);
  aabb->upperBound.Set(p.x + m_radius, p.y + m_radius);
}
void b2CircleShape::ComputeMass(b2MassData* md, float density) const {
  md->mass = density * b2_pi * m_radius * m_radius;
  md->center = m_p;
  md->I = md->mass * (0.5f * m_radius * m_radius + b2Dot(m_p, m_p));
}
void b2CircleShape::WriteDeviceQuads(float* q) const {
  q[0] = m_p.x;
  q[1] = m_p.y;
  q[2] = m_radius;
  q[3] = 0.0f;
}

// ---- edge -------------------------------------------------------------------------------------
void b2EdgeShape::SetOneSided(const b2Vec2& v0, const b2Vec2& v1, const b2Vec2& v2, const b2Vec2& v3) {
  m_vertex0 = v0;
  m_vertex1 = v1;
  m_vertex2 = v2;
  m_vertex3 = v3;
  m_oneSided = true;
}
void b2EdgeShape::SetTwoSided(const b2Vec2& v1, const b2Vec2& v2) {
  m_vertex1 = v1;
  m_vertex2 = v2;
  m_oneSided = false;
}
bool b2EdgeShape::TestPoint(const b2Transform&, const b2Vec2&) const { return false; }
void b2EdgeShape::ComputeAABB(b2AABB* aabb, const b2Transform& xf) const {
  b2Vec2 v1 = b2Mul(xf, m_vertex1);
  b2Vec2 v2 = b2Mul(xf, m_vertex2);
  b2Vec2 lower = b2Min(v1, v2);
  b2Vec2 upper = b2Max(v1, v2);
  b2Vec2 r(m_radius, m_radius);
  aabb->lowerBound = lower - r;
  aabb->upperBound = upper + r;
}
void b2EdgeShape::ComputeMass(b2MassData* md, float) const {
  md->mass = 0.0f;
  md->center = 0.5f * (m_vertex1 + m_vertex2);
  md->I = 0.0f;
}
void b2EdgeShape::WriteDeviceQuads(float* q) const {
  q[0] = m_vertex1.x; q[1] = m_vertex1.y; q[2] = m_vertex2.x; q[3] = m_vertex2.y;
  q[4] = m_vertex0.x; q[5] = m_vertex0.y; q[6] = m_vertex3.x; q[7] = m_vertex3.y;
  q[8] = m_radius;    q[9] = m_oneSided ? 1.0f : 0.0f; q[10] = 0.0f; q[11] = 0.0f;
}

// ---- polygon ----------------------------------------------------------------------------------
void b2PolygonShape::SetAsBox(float hx, float hy) {
  m_count = 4;
  m_vertices[0].Set(-hx, -hy);
  m_vertices[1].Set(hx, -hy);
  m_vertices[2].Set(hx, hy);
  m_vertices[3].Set(-hx, hy);
  m_normals[0].Set(0.0f, -1.0f);
  m_normals[1].Set(1.0f, 0.0f);
  m_normals[2].Set(0.0f, 1.0f);
  m_normals[3].Set(-1.0f, 0.0f);
  m_centroid.SetZero();
}

void b2PolygonShape::SetAsBox(float hx, float hy, const b2Vec2& center, float angle) {
  SetAsBox(hx, hy);
  m_centroid = center;
  b2Transform xf;
  xf.p = center;
  xf.q.Set(angle);
  for (int32 i = 0; i < m_count; ++i) {
    m_vertices[i] = b2Mul(xf, m_vertices[i]);
    m_normals[i] = b2Mul(xf.q, m_normals[i]);
  }
}

// area-weighted centroid by fanning triangles from the first vertex
static b2Vec2 PolygonCentroid(const b2Vec2* vs, int32 count) {
  b2Vec2 c(0.0f, 0.0f);
  float area = 0.0f;
  const b2Vec2 s = vs[0];
  const float inv3 = 1.0f / 3.0f;
  for (int32 i = 0; i < count; ++i) {
    b2Vec2 p1 = vs[0] - s;
    b2Vec2 p2 = vs[i] - s;
    b2Vec2 p3 = i + 1 < count ? vs[i + 1] - s : vs[0] - s;
    b2Vec2 e1 = p2 - p1;
    b2Vec2 e2 = p3 - p1;
    float D = b2Cross(e1, e2);
    float triangleArea = 0.5f * D;
    area += triangleArea;
    c += triangleArea * inv3 * (p1 + p2 + p3);
  }
  c = (1.0f / area) * c + s;
  return c;
}

void b2PolygonShape::Set(const b2Vec2* vertices, int32 count) {
  if (count < 3) {
    SetAsBox(1.0f, 1.0f);
    return;
  }
  int32 n = b2Min(count, (int32)b2_maxPolygonVertices);

  // weld near-duplicate input points
  b2Vec2 ps[b2_maxPolygonVertices];
  int32 kept = 0;
  const float weldSq = (0.5f * b2_linearSlop) * (0.5f * b2_linearSlop);
  for (int32 i = 0; i < n; ++i) {
    bool unique = true;
    for (int32 j = 0; j < kept; ++j) {
      if (b2DistanceSquared(vertices[i], ps[j]) < weldSq) {
        unique = false;
        break;
      }
    }
    if (unique) ps[kept++] = vertices[i];
  }
  n = kept;
  if (n < 3) {
    SetAsBox(1.0f, 1.0f);
    return;
  }

  // gift wrapping from the right-most (then lowest) point, counter-clockwise
  int32 start = 0;
  float x0 = ps[0].x;
  for (int32 i = 1; i < n; ++i) {
    float x = ps[i].x;
    if (x > x0 || (x == x0 && ps[i].y < ps[start].y)) {
      start = i;
      x0 = x;
    }
  }
  int32 hull[b2_maxPolygonVertices];
  int32 m = 0;
  int32 ih = start;
  for (;;) {
    hull[m] = ih;
    int32 ie = 0;
    for (int32 j = 1; j < n; ++j) {
      if (ie == ih) {
        ie = j;
        continue;
      }
      b2Vec2 r = ps[ie] - ps[hull[m]];
      b2Vec2 v = ps[j] - ps[hull[m]];
      float c = b2Cross(r, v);
      if (c < 0.0f) ie = j;
      if (c == 0.0f && v.LengthSquared() > r.LengthSquared()) ie = j;  // collinear: take the farther
    }
    ++m;
    ih = ie;
    if (ie == start) break;
  }
  if (m < 3) {
    SetAsBox(1.0f, 1.0f);
    return;
  }
  m_count = m;
  for (int32 i = 0; i < m; ++i) m_vertices[i] = ps[hull[i]];
  for (int32 i = 0; i < m; ++i) {
    int32 i2 = i + 1 < m ? i + 1 : 0;
    b2Vec2 edge = m_vertices[i2] - m_vertices[i];
    m_normals[i] = b2Cross(edge, 1.0f);
    m_normals[i].Normalize();
  }
  m_centroid = PolygonCentroid(m_vertices, m);
}

bool b2PolygonShape::TestPoint(const b2Transform& xf, const b2Vec2& p) const {
  b2Vec2 pLocal = b2MulT(xf.q, p - xf.p);
  for (int32 i = 0; i < m_count; ++i) {
    float dot = b2Dot(m_normals[i], pLocal - m_vertices[i]);
    if (dot > 0.0f) return false;
  }
  return true;
}

void b2PolygonShape::ComputeAABB(b2AABB* aabb, const b2Transform& xf) const {
  b2Vec2 lower = b2Mul(xf, m_vertices[0]);
  b2Vec2 upper = lower;
  for (int32 i = 1; i < m_count; ++i) {
    b2Vec2 v = b2Mul(xf, m_vertices[i]);
    lower = b2Min(lower, v);
    upper = b2Max(upper, v);
  }
  b2Vec2 r(m_radius, m_radius);
  aabb->lowerBound = lower - r;
  aabb->upperBound = upper + r;
}

void b2PolygonShape::ComputeMass(b2MassData* md, float density) const {
  // triangle fan about vertex 0; I about that vertex, then shifted (parallel axis) to the origin
  b2Vec2 center(0.0f, 0.0f);
  float area = 0.0f;
  float I = 0.0f;
  const b2Vec2 s = m_vertices[0];
  const float k_inv3 = 1.0f / 3.0f;
  for (int32 i = 0; i < m_count; ++i) {
    b2Vec2 e1 = m_vertices[i] - s;
    b2Vec2 e2 = i + 1 < m_count ? m_vertices[i + 1] - s : m_vertices[0] - s;
    float D = b2Cross(e1, e2);
    float triangleArea = 0.5f * D;
    area += triangleArea;
    center += triangleArea * k_inv3 * (e1 + e2);
    float ex1 = e1.x, ey1 = e1.y;
    float ex2 = e2.x, ey2 = e2.y;
    float intx2 = ex1 * ex1 + ex2 * ex1 + ex2 * ex2;
    float inty2 = ey1 * ey1 + ey2 * ey1 + ey2 * ey2;
    I += (0.25f * k_inv3 * D) * (intx2 + inty2);
  }
  md->mass = density * area;
  center *= 1.0f / area;
  md->center = center + s;
  md->I = density * I;
  md->I += md->mass * (b2Dot(md->center, md->center) - b2Dot(center, center));
}

bool b2PolygonShape::Validate() const {
  for (int32 i = 0; i < m_count; ++i) {
    int32 i2 = i < m_count - 1 ? i + 1 : 0;
    b2Vec2 p = m_vertices[i];
    b2Vec2 e = m_vertices[i2] - p;
    for (int32 j = 0; j < m_count; ++j) {
      if (j == i || j == i2) continue;
      b2Vec2 v = m_vertices[j] - p;
      if (b2Cross(e, v) < 0.0f) return false;
    }
  }
  return true;
}

void b2PolygonShape::WriteDeviceQuads(float* q) const {
  q[0] = m_centroid.x;
  q[1] = m_centroid.y;
  q[2] = m_radius;
  q[3] = (float)m_count;
  for (int32 i = 0; i < m_count; ++i) {
    q[4 + 4 * i] = m_vertices[i].x;
    q[5 + 4 * i] = m_vertices[i].y;
    q[6 + 4 * i] = m_normals[i].x;
    q[7 + 4 * i] = m_normals[i].y;
  }
}

// ---- world manifold (b2_collision.cpp:26-90) -----------------------------------------------------
void b2WorldManifold::Initialize(const b2Manifold* manifold, const b2Transform& xfA, float radiusA,
                                 const b2Transform& xfB, float radiusB) {
  if (manifold->pointCount == 0) return;
  switch (manifold->type) {
    case b2Manifold::e_circles: {
      normal.Set(1.0f, 0.0f);
      b2Vec2 pointA = b2Mul(xfA, manifold->localPoint);
      b2Vec2 pointB = b2Mul(xfB, manifold->points[0].localPoint);
      if (b2DistanceSquared(pointA, pointB) > b2_epsilon * b2_epsilon) {
        normal = pointB - pointA;
        normal.Normalize();
      }
      b2Vec2 cA = pointA + radiusA * normal;
      b2Vec2 cB = pointB - radiusB * normal;
      points[0] = 0.5f * (cA + cB);
      separations[0] = b2Dot(cB - cA, normal);
    } break;
    case b2Manifold::e_faceA: {
      normal = b2Mul(xfA.q, manifold->localNormal);
      b2Vec2 planePoint = b2Mul(xfA, manifold->localPoint);
      for (int32 i = 0; i < manifold->pointCount; ++i) {
        b2Vec2 clipPoint = b2Mul(xfB, manifold->points[i].localPoint);
        b2Vec2 cA = clipPoint + (radiusA - b2Dot(clipPoint - planePoint, normal)) * normal;
        b2Vec2 cB = clipPoint - radiusB * normal;
        points[i] = 0.5f * (cA + cB);
        separations[i] = b2Dot(cB - cA, normal);
      }
    } break;
    case b2Manifold::e_faceB: {
      normal = b2Mul(xfB.q, manifold->localNormal);
      b2Vec2 planePoint = b2Mul(xfB, manifold->localPoint);
      for (int32 i = 0; i < manifold->pointCount; ++i) {
        b2Vec2 clipPoint = b2Mul(xfA, manifold->points[i].localPoint);
        b2Vec2 cB = clipPoint + (radiusB - b2Dot(clipPoint - planePoint, normal)) * normal;
        b2Vec2 cA = clipPoint - radiusA * normal;
        points[i] = 0.5f * (cA + cB);
        separations[i] = b2Dot(cA - cB, normal);
      }
      normal = -normal;
    } break;
  }
}

void b2GetPointStates(b2PointState state1[b2_maxManifoldPoints], b2PointState state2[b2_maxManifoldPoints],
                      const b2Manifold* manifold1, const b2Manifold* manifold2) {
  for (int32 i = 0; i < b2_maxManifoldPoints; ++i) state1[i] = state2[i] = b2_nullState;
  for (int32 i = 0; i < manifold1->pointCount; ++i) {
    state1[i] = b2_removeState;
    for (int32 j = 0; j < manifold2->pointCount; ++j)
      if (manifold2->points[j].id.key == manifold1->points[i].id.key) {
        state1[i] = b2_persistState;
        break;
      }
  }
  for (int32 i = 0; i < manifold2->pointCount; ++i) {
    state2[i] = b2_addState;
    for (int32 j = 0; j < manifold1->pointCount; ++j)
      if (manifold1->points[j].id.key == manifold2->points[i].id.key) {
        state2[i] = b2_persistState;
        break;
      }
  }
}

// ---- single-shape ray casts (host).  Same equations as the device versions in
// csrc/b2g_query.cuh; restated from b2_circle_shape.cpp:56-89, b2_edge_shape.cpp:89-154,
// b2_polygon_shape.cpp:303-371. ---------------------------------------------------------------
bool b2CircleShape::RayCast(b2RayCastOutput* output, const b2RayCastInput& input, const b2Transform& xf) const {
  const b2Vec2 centre = xf.p + b2Mul(xf.q, m_p);
  const b2Vec2 s = input.p1 - centre;
  const b2Vec2 r = input.p2 - input.p1;
  const float b = b2Dot(s, s) - m_radius * m_radius;
  const float c = b2Dot(s, r);
  const float rr = b2Dot(r, r);
  const float sigma = c * c - rr * b;
  if (sigma < 0.0f || rr < b2_epsilon) return false;  // line misses the circle, or degenerate segment
  float a = -(c + b2Sqrt(sigma));
  if (a < 0.0f || a > input.maxFraction * rr) return false;
  a /= rr;
  output->fraction = a;
  output->normal = s + a * r;
  output->normal.Normalize();
  return true;
}

bool b2EdgeShape::RayCast(b2RayCastOutput* output, const b2RayCastInput& input, const b2Transform& xf) const {
  // ray in the edge's frame
  const b2Vec2 p1 = b2MulT(xf.q, input.p1 - xf.p);
  const b2Vec2 p2 = b2MulT(xf.q, input.p2 - xf.p);
  const b2Vec2 d = p2 - p1;
  const b2Vec2 e = m_vertex2 - m_vertex1;
  b2Vec2 normal(e.y, -e.x);  // to the right of v1 -> v2
  normal.Normalize();
  const float numerator = b2Dot(normal, m_vertex1 - p1);
  if (m_oneSided && numerator > 0.0f) return false;
  const float denominator = b2Dot(normal, d);
  if (denominator == 0.0f) return false;
  const float t = numerator / denominator;
  if (t < 0.0f || input.maxFraction < t) return false;
  const b2Vec2 q = p1 + t * d;
  const float rr = b2Dot(e, e);
  if (rr == 0.0f) return false;
  const float along = b2Dot(q - m_vertex1, e) / rr;
  if (along < 0.0f || 1.0f < along) return false;
  output->fraction = t;
  output->normal = numerator > 0.0f ? -b2Mul(xf.q, normal) : b2Mul(xf.q, normal);
  return true;
}

bool b2PolygonShape::RayCast(b2RayCastOutput* output, const b2RayCastInput& input, const b2Transform& xf) const {
  const b2Vec2 p1 = b2MulT(xf.q, input.p1 - xf.p);
  const b2Vec2 p2 = b2MulT(xf.q, input.p2 - xf.p);
  const b2Vec2 d = p2 - p1;
  float enter = 0.0f, leave = input.maxFraction;
  int32 face = -1;
  for (int32 i = 0; i < m_count; ++i) {
    // clip the segment against the half-plane of face i: dot(n, p1 + a d - v) <= 0
    const float numerator = b2Dot(m_normals[i], m_vertices[i] - p1);
    const float denominator = b2Dot(m_normals[i], d);
    if (denominator == 0.0f) {
      if (numerator < 0.0f) return false;  // parallel and outside
    } else if (denominator < 0.0f && numerator < enter * denominator) {
      enter = numerator / denominator;
      face = i;
    } else if (denominator > 0.0f && numerator < leave * denominator) {
      leave = numerator / denominator;
    }
    if (leave < enter) return false;
  }
  if (face < 0) return false;
  output->fraction = enter;
  output->normal = b2Mul(xf.q, m_normals[face]);
  return true;
}

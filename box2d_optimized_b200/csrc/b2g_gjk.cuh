// b2g_gjk.cuh — GJK overlap test for sensor contacts.
//
// b2Contact::Update decides `touching` of a SENSOR contact with b2TestOverlap
// (src/collision/b2_collision.cpp:239-258) = b2Distance (src/collision/b2_distance.cpp:432-584) from an
// empty simplex cache with useRadii, overlap iff the distance is below 10 epsilon.  That differs
// from "the manifold has points" (SAT with a 2 * polygonRadius skin) when two corners pass each other
// inside the skin, so the sensor path runs the same algorithm: proxies of at most 8 vertices
// (b2DistanceProxy::Set, :31-66), support points by first-best dot product
// (include/box2d/b2_distance.h:137-152), the Voronoi-region simplex solvers (:249-418), at most 20
// support evaluations, duplicate-support termination.  SURVEY.md §8(f) rank 3.
#pragma once
#include "b2g_collide.cuh"

struct GjkProxy {
  float2 v[B2G_MAX_POLY_VERTS];
  int count;
  float radius;
};

__device__ __forceinline__ void gjk_load_proxy(GjkProxy& P, const float4* __restrict__ pool, int type, int off) {
  if (type == 0) {  // circle: its centre, radius = the circle's
    float4 a = __ldg(pool + off);
    P.v[0] = make_float2(a.x, a.y);
    P.count = 1;
    P.radius = a.z;
  } else if (type == 1) {  // edge: v1, v2
    float4 a = __ldg(pool + off), c = __ldg(pool + off + 2);
    P.v[0] = make_float2(a.x, a.y);
    P.v[1] = make_float2(a.z, a.w);
    P.count = 2;
    P.radius = c.x;
  } else {
    float4 h = __ldg(pool + off);
    P.count = (int)h.w;
    P.radius = h.z;
    for (int i = 0; i < P.count; ++i) {
      float4 q = __ldg(pool + off + 1 + i);
      P.v[i] = make_float2(q.x, q.y);
    }
  }
}

__device__ __forceinline__ int gjk_support(const GjkProxy& P, float2 d) {
  int best = 0;
  float bestValue = dot2(P.v[0], d);
  for (int i = 1; i < P.count; ++i) {
    float value = dot2(P.v[i], d);
    if (value > bestValue) {
      best = i;
      bestValue = value;
    }
  }
  return best;
}

struct GjkVertex {
  float2 wA, wB, w;  // support points and their difference wB - wA
  float a;           // barycentric weight of the closest point
  int iA, iB;
};

// closest point of the segment w1-w2 to the origin (b2Simplex::Solve2, :249-279)
__device__ __forceinline__ void gjk_solve2(GjkVertex* s, int& count) {
  float2 w1 = s[0].w, w2 = s[1].w;
  float2 e12 = w2 - w1;
  float d12_2 = -dot2(w1, e12);
  if (d12_2 <= 0.0f) {  // vertex 1 region
    s[0].a = 1.0f;
    count = 1;
    return;
  }
  float d12_1 = dot2(w2, e12);
  if (d12_1 <= 0.0f) {  // vertex 2 region
    s[1].a = 1.0f;
    count = 1;
    s[0] = s[1];
    return;
  }
  float inv = 1.0f / (d12_1 + d12_2);
  s[0].a = d12_1 * inv;
  s[1].a = d12_2 * inv;
  count = 2;
}

// closest point of the triangle to the origin (b2Simplex::Solve3, :286-418)
__device__ __forceinline__ void gjk_solve3(GjkVertex* s, int& count) {
  float2 w1 = s[0].w, w2 = s[1].w, w3 = s[2].w;
  float2 e12 = w2 - w1;
  float d12_1 = dot2(w2, e12), d12_2 = -dot2(w1, e12);
  float2 e13 = w3 - w1;
  float d13_1 = dot2(w3, e13), d13_2 = -dot2(w1, e13);
  float2 e23 = w3 - w2;
  float d23_1 = dot2(w3, e23), d23_2 = -dot2(w2, e23);
  float n123 = cross2(e12, e13);
  float d123_1 = n123 * cross2(w2, w3);
  float d123_2 = n123 * cross2(w3, w1);
  float d123_3 = n123 * cross2(w1, w2);
  if (d12_2 <= 0.0f && d13_2 <= 0.0f) {  // vertex 1
    s[0].a = 1.0f;
    count = 1;
    return;
  }
  if (d12_1 > 0.0f && d12_2 > 0.0f && d123_3 <= 0.0f) {  // edge 1-2
    float inv = 1.0f / (d12_1 + d12_2);
    s[0].a = d12_1 * inv;
    s[1].a = d12_2 * inv;
    count = 2;
    return;
  }
  if (d13_1 > 0.0f && d13_2 > 0.0f && d123_2 <= 0.0f) {  // edge 1-3
    float inv = 1.0f / (d13_1 + d13_2);
    s[0].a = d13_1 * inv;
    s[2].a = d13_2 * inv;
    count = 2;
    s[1] = s[2];
    return;
  }
  if (d12_1 <= 0.0f && d23_2 <= 0.0f) {  // vertex 2
    s[1].a = 1.0f;
    count = 1;
    s[0] = s[1];
    return;
  }
  if (d13_1 <= 0.0f && d23_1 <= 0.0f) {  // vertex 3
    s[2].a = 1.0f;
    count = 1;
    s[0] = s[2];
    return;
  }
  if (d23_1 > 0.0f && d23_2 > 0.0f && d123_1 <= 0.0f) {  // edge 2-3
    float inv = 1.0f / (d23_1 + d23_2);
    s[1].a = d23_1 * inv;
    s[2].a = d23_2 * inv;
    count = 2;
    s[0] = s[2];
    return;
  }
  float inv = 1.0f / (d123_1 + d123_2 + d123_3);  // inside
  s[0].a = d123_1 * inv;
  s[1].a = d123_2 * inv;
  s[2].a = d123_3 * inv;
  count = 3;
}

// b2TestOverlap(shapeA, shapeB, xfA, xfB).  Not inlined: sensor contacts are rare and the narrowphase
// kernel's register budget belongs to the manifold functions.
__device__ __noinline__ bool gjk_test_overlap(const float4* __restrict__ pool, int typeA, int offA, Xf xfA, int typeB,
                                                 int offB, Xf xfB) {
  GjkProxy A, B;
  gjk_load_proxy(A, pool, typeA, offA);
  gjk_load_proxy(B, pool, typeB, offB);
  GjkVertex s[3];
  int count = 1;  // empty cache: start from vertex 0 of both proxies (b2Simplex::ReadCache, :120-133)
  s[0].iA = 0;
  s[0].iB = 0;
  s[0].wA = xf_mul(xfA, A.v[0]);
  s[0].wB = xf_mul(xfB, B.v[0]);
  s[0].w = s[0].wB - s[0].wA;
  s[0].a = 1.0f;
  int saveA[3], saveB[3];
  int iter = 0;
  while (iter < 20) {
    const int saveCount = count;
    for (int i = 0; i < saveCount; ++i) {
      saveA[i] = s[i].iA;
      saveB[i] = s[i].iB;
    }
    if (count == 2) gjk_solve2(s, count);
    else if (count == 3) gjk_solve3(s, count);
    if (count == 3) break;  // the origin is inside the triangle
    float2 d;
    if (count == 1) {
      d = -s[0].w;
    } else {
      float2 e12 = s[1].w - s[0].w;
      float sgn = cross2(e12, -s[0].w);
      d = sgn > 0.0f ? cross_sv(1.0f, e12) : cross_vs(e12, 1.0f);
    }
    if (dot2(d, d) < B2G_EPSILON * B2G_EPSILON) break;  // origin on the simplex: overlapped
    GjkVertex& nv = s[count];
    nv.iA = gjk_support(A, rot_mulT(xfA.q, -d));
    nv.wA = xf_mul(xfA, A.v[nv.iA]);
    nv.iB = gjk_support(B, rot_mulT(xfB.q, d));
    nv.wB = xf_mul(xfB, B.v[nv.iB]);
    nv.w = nv.wB - nv.wA;
    ++iter;
    bool duplicate = false;
    for (int i = 0; i < saveCount; ++i)
      if (nv.iA == saveA[i] && nv.iB == saveB[i]) {
        duplicate = true;
        break;
      }
    if (duplicate) break;
    ++count;
  }
  // witness points (b2Simplex::GetWitnessPoints, :199-226) and their distance
  float2 pA, pB;
  if (count == 1) {
    pA = s[0].wA;
    pB = s[0].wB;
  } else if (count == 2) {
    pA = s[0].a * s[0].wA + s[1].a * s[1].wA;
    pB = s[0].a * s[0].wB + s[1].a * s[1].wB;
  } else {
    pA = s[0].a * s[0].wA + s[1].a * s[1].wA + s[2].a * s[2].wA;
    pB = pA;
  }
  float distance = len2(pB - pA);
  float rA = A.radius, rB = B.radius;
  if (distance > rA + rB && distance > B2G_EPSILON) distance -= rA + rB;  // useRadii (:550-571)
  else distance = 0.0f;
  return distance < 10.0f * B2G_EPSILON;
}

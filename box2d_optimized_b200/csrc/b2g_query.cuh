// b2g_query.cuh — read-only consumers of the broadphase BVH:
//   b2World::QueryAABB  (src/dynamics/b2_world.cpp:1193-1207 -> b2BroadPhase::Query, b2_broad_phase.h:622-643)
//   b2World::RayCast    (src/dynamics/b2_world.cpp:1226-1246 -> b2BroadPhase::RayCast, b2_broad_phase.h:645-716,
//                        b2Fixture::RayCast -> b2{Circle,Edge,Polygon}Shape::RayCast)
// One thread per query walks the same implicit 8-wide tree the pair finder walks.  Batches of
// queries are what an RL observation (a fan of rays per agent, thousands of agents) looks like.
//
// Results are sets / minima, so they do not depend on the traversal order; where the reference's
// outcome does depend on its own tree order (which of two hits with EQUAL fraction a "closest"
// callback keeps) the lower fixture index wins here.
#pragma once
#include "b2g_broadphase.cuh"

// ---- shape ray casts: fraction in [0, maxFraction], normal in world space ----------------------
// b2CircleShape::RayCast (src/collision/b2_circle_shape.cpp:56-89)
__device__ __forceinline__ bool ray_cast_circle(const float4* __restrict__ pool, int off, Xf xf, float2 p1, float2 p2,
                                                float maxFraction, float& fraction, float2& normal) {
  Circle c = load_circle(pool, off);
  float2 position = xf.p + rot_mul(xf.q, c.p);
  float2 s = p1 - position;
  float b = dot2(s, s) - c.radius * c.radius;
  float2 r = p2 - p1;
  float cc = dot2(s, r);
  float rr = dot2(r, r);
  float sigma = cc * cc - rr * b;
  if (sigma < 0.0f || rr < B2G_EPSILON) return false;
  float a = -(cc + sqrtf(sigma));
  if (0.0f <= a && a <= maxFraction * rr) {
    a /= rr;
    fraction = a;
    normal = s + a * r;
    normalize2(normal);
    return true;
  }
  return false;
}

// b2EdgeShape::RayCast (src/collision/b2_edge_shape.cpp:89-154)
__device__ __forceinline__ bool ray_cast_edge(const float4* __restrict__ pool, int off, Xf xf, float2 wp1, float2 wp2,
                                              float maxFraction, float& fraction, float2& normal) {
  Edge e = load_edge(pool, off);
  float2 p1 = rot_mulT(xf.q, wp1 - xf.p);
  float2 p2 = rot_mulT(xf.q, wp2 - xf.p);
  float2 d = p2 - p1;
  float2 ev = e.v2 - e.v1;
  float2 n = make_float2(ev.y, -ev.x);
  normalize2(n);
  float numerator = dot2(n, e.v1 - p1);
  if (e.oneSided && numerator > 0.0f) return false;
  float denominator = dot2(n, d);
  if (denominator == 0.0f) return false;
  float t = numerator / denominator;
  if (t < 0.0f || maxFraction < t) return false;
  float2 q = p1 + t * d;
  float rr = dot2(ev, ev);
  if (rr == 0.0f) return false;
  float s = dot2(q - e.v1, ev) / rr;
  if (s < 0.0f || 1.0f < s) return false;
  fraction = t;
  float2 wn = rot_mul(xf.q, n);
  normal = numerator > 0.0f ? -wn : wn;
  return true;
}

// b2PolygonShape::RayCast (src/collision/b2_polygon_shape.cpp:303-371)
__device__ __forceinline__ bool ray_cast_polygon(const float4* __restrict__ pool, int off, Xf xf, float2 wp1, float2 wp2,
                                                 float maxFraction, float& fraction, float2& normal) {
  float2 p1 = rot_mulT(xf.q, wp1 - xf.p);
  float2 p2 = rot_mulT(xf.q, wp2 - xf.p);
  float2 d = p2 - p1;
  float lower = 0.0f, upper = maxFraction;
  int index = -1;
  float2 bestNormal = make_float2(0.0f, 0.0f);
  const int count = (int)__ldg(pool + off).w;
  for (int i = 0; i < count; ++i) {
    float4 vn = __ldg(pool + off + 1 + i);  // v.x, v.y, n.x, n.y
    float2 v = make_float2(vn.x, vn.y), nrm = make_float2(vn.z, vn.w);
    float numerator = dot2(nrm, v - p1);
    float denominator = dot2(nrm, d);
    if (denominator == 0.0f) {
      if (numerator < 0.0f) return false;
    } else {
      if (denominator < 0.0f && numerator < lower * denominator) {
        lower = numerator / denominator;  // the segment enters this half-space
        index = i;
        bestNormal = nrm;
      } else if (denominator > 0.0f && numerator < upper * denominator) {
        upper = numerator / denominator;  // the segment leaves this half-space
      }
    }
    if (upper < lower) return false;
  }
  if (index >= 0) {
    fraction = lower;
    normal = rot_mul(xf.q, bestNormal);
    return true;
  }
  return false;
}

__device__ __forceinline__ bool ray_cast_fixture(const float4* __restrict__ pool, int type, int off, Xf xf, float2 p1,
                                                 float2 p2, float maxFraction, float& fraction, float2& normal) {
  if (type == 0) return ray_cast_circle(pool, off, xf, p1, p2, maxFraction, fraction, normal);
  if (type == 1) return ray_cast_edge(pool, off, xf, p1, p2, maxFraction, fraction, normal);
  return ray_cast_polygon(pool, off, xf, p1, p2, maxFraction, fraction, normal);
}

// ---- QueryAABB ---------------------------------------------------------------------------------
// counts[q] = number of fixtures whose tight AABB overlaps box q (inclusive test); the first
// `cap` of them (arbitrary order; the host sorts) go to fixtures[q*cap ..].  world[q] >= 0
// restricts the query to one world of a multi-world arena.
__global__ void __launch_bounds__(128)
k_query_aabb(int nq, const float4* __restrict__ qbox, const int* __restrict__ qworld, WideBvh T,
             const float4* __restrict__ leafBox, const int4* __restrict__ leafInfo,
             const unsigned long long* __restrict__ leafKey, const int* __restrict__ worldFirst,
             const int* __restrict__ worldLast, int numWorlds, int cap, int* counts, int* fixtures) {
  B2G_PDL_ENTER();
  int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= nq) return;
  const int n = T.count[0];
  float4 box = qbox[q];
  int ws = 0, we = n - 1;
  if (qworld && numWorlds > 1 && qworld[q] >= 0) {
    ws = worldFirst[qworld[q]];
    we = worldLast[qworld[q]];
  }
  int found = 0;
  int* out = fixtures + (size_t)q * cap;
  if (n == 1) {
    // the reference reports a lone root leaf without testing it (b2_broad_phase.h:631-633)
    int4 li = leafInfo[0];
    if (!((unsigned int)li.z & 16u)) {
      if (found < cap) out[found] = li.x;
      ++found;
    }
    counts[q] = found;
    return;
  }
  int stack[72];
  int sp = 0;
  if (n > 1) stack[sp++] = (T.levels << 27);
  WideChildren ch;
  while (sp > 0) {
    int e = stack[--sp];
    const int l = e >> 27, j = e & 0x7ffffff;
    wide_load(T, leafBox, leafKey, l, j, ch);
    unsigned int hits = 0;
#pragma unroll
    for (int c = 0; c < B2G_BVH_W; ++c)
      if (c < ch.count && aabb_overlap(box, ch.box[c])) hits |= 1u << c;
    while (hits) {
      {
        const int ci = ch.base + __ffs(hits) - 1;
        hits &= hits - 1;
        if (l == 1) {
          if (ci >= ws && ci <= we) {
            int4 li = leafInfo[ci];
            if (!((unsigned int)li.z & 16u)) {
              if (found < cap) out[found] = li.x;
              ++found;
            }
          }
        } else if (wide_in_range(l - 1, ci, ws, we)) {
          stack[sp++] = ((l - 1) << 27) | ci;
        }
      }
    }
  }
  counts[q] = found;
}

// ---- RayCast -----------------------------------------------------------------------------------
// segment vs box as b2BroadPhase::RayCast does it: segment AABB overlap, then the separating-axis
// test against the segment's normal (b2_broad_phase.h:676-688)
__device__ __forceinline__ bool ray_hits_box(float4 box, float4 segBox, float2 p1, float2 v, float2 absV) {
  if (!aabb_overlap(box, segBox)) return false;
  float2 c = make_float2(0.5f * (box.x + box.z), 0.5f * (box.y + box.w));
  float2 h = make_float2(0.5f * (box.z - box.x), 0.5f * (box.w - box.y));
  float separation = absf_(dot2(v, p1 - c)) - dot2(absV, h);
  return !(separation > 0.0f);
}
__device__ __forceinline__ float4 segment_box(float2 p1, float2 p2, float maxFraction) {
  float2 t = p1 + maxFraction * (p2 - p1);
  return make_float4(minf_(p1.x, t.x), minf_(p1.y, t.y), maxf_(p1.x, t.x), maxf_(p1.y, t.y));
}

// mode 0: closest hit per ray -> hitFixture[r] (-1 = none), hitFraction[r], hitNormal[r]
//         (a b2RayCastCallback that returns `fraction`; equal fractions: lower fixture index)
// mode 1: every hit with fraction <= maxFraction[r] (a callback that returns 1): counts[r] and the
//         first `cap` hits per ray, unordered
// categoryMask: hits are taken only from fixtures with (categoryBits & categoryMask) != 0 — the
// usual filter a callback applies by returning -1; 0xFFFF takes everything.
__global__ void __launch_bounds__(128)
k_ray_cast(int nr_, const float4* __restrict__ rays, const float* __restrict__ maxFractionIn,
           const int* __restrict__ rworld, int mode, uint32_t categoryMask, WideBvh T,
           const float4* __restrict__ leafBox, const int4* __restrict__ leafInfo,
           const unsigned long long* __restrict__ leafKey, const int* __restrict__ worldFirst,
           const int* __restrict__ worldLast, int numWorlds, const int* __restrict__ fShapeOff,
           const float4* __restrict__ shapes, const float4* __restrict__ xf, int cap, int* counts, int* hitFixture,
           float* hitFraction, float2* hitNormal) {
  B2G_PDL_ENTER();
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= nr_) return;
  const int n = T.count[0];
  float4 ray = rays[r];
  float2 p1 = make_float2(ray.x, ray.y), p2 = make_float2(ray.z, ray.w);
  float maxFraction = maxFractionIn ? maxFractionIn[r] : 1.0f;
  int ws = 0, we = n - 1;
  if (rworld && numWorlds > 1 && rworld[r] >= 0) {
    ws = worldFirst[rworld[r]];
    we = worldLast[rworld[r]];
  }
  float2 dir = p2 - p1;
  normalize2(dir);
  float2 v = make_float2(-dir.y, dir.x);  // b2Cross(1, r)
  float2 absV = make_float2(absf_(v.x), absf_(v.y));
  float4 segBox = segment_box(p1, p2, maxFraction);

  int bestFixture = -1, found = 0;
  float bestFraction = 0.0f;
  float2 bestNormal = make_float2(0.0f, 0.0f);
  const size_t base = (size_t)r * cap;

  auto visit_leaf = [&](int leaf) {
    int4 li = leafInfo[leaf];
    unsigned int z = (unsigned int)li.z;
    if (z & 16u) return;
    if (!(((unsigned int)li.w & 0xffffu) & categoryMask)) return;
    float fraction;
    float2 normal;
    if (!ray_cast_fixture(shapes, (int)(z & 3u), fShapeOff[li.x], xf_from4(xf[li.y]), p1, p2, maxFraction, fraction,
                          normal))
      return;
    if (mode == 0) {
      if (bestFixture < 0 || fraction < bestFraction || (fraction == bestFraction && li.x < bestFixture)) {
        bestFixture = li.x;
        bestFraction = fraction;
        bestNormal = normal;
        maxFraction = fraction;  // the callback's return value clips the ray
        segBox = segment_box(p1, p2, maxFraction);
      }
    } else {
      if (found < cap) {
        hitFixture[base + found] = li.x;
        hitFraction[base + found] = fraction;
        hitNormal[base + found] = normal;
      }
      ++found;
    }
  };

  if (n >= 1) {
    int stack[72];
    int sp = 0;
    stack[sp++] = (T.levels << 27);
    WideChildren ch;
    while (sp > 0) {
      int e = stack[--sp];
      const int l = e >> 27, j = e & 0x7ffffff;
      wide_load(T, leafBox, leafKey, l, j, ch);
      unsigned int hits = 0;
#pragma unroll
      for (int c = 0; c < B2G_BVH_W; ++c)
        if (c < ch.count && ray_hits_box(ch.box[c], segBox, p1, v, absV)) hits |= 1u << c;
      while (hits) {
        // (a leaf visited earlier in this loop may have clipped the segment: a stale hit only costs a visit)
        {
          const int ci = ch.base + __ffs(hits) - 1;
          hits &= hits - 1;
          if (l == 1) {
            if (ci >= ws && ci <= we) visit_leaf(ci);
          } else if (wide_in_range(l - 1, ci, ws, we)) {
            stack[sp++] = ((l - 1) << 27) | ci;
          }
        }
      }
    }
  }
  if (mode == 0) {
    hitFixture[r] = bestFixture;
    hitFraction[r] = bestFraction;
    hitNormal[r] = bestNormal;
  } else {
    counts[r] = found;
  }
}

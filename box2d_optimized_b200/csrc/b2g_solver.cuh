// b2g_solver.cuh — contact-constraint device functions (one constraint per call).
//
// Each function is the body of one loop iteration of b2ContactSolver
// (src/dynamics/b2_contact_solver.cpp):
//   prepare_constraint        Initialize :65-152 + InitializeVelocityConstraints :155-258
//   warm_start_constraint     WarmStart :260-299
//   solve_velocity_constraint SolveVelocityConstraints :301-639 (friction rows, then the
//                             1-point row or the 2-point block LCP by total enumeration)
//   store_impulses_constraint StoreImpulses :641-657
//   solve_position_constraint SolvePositionConstraints :711-787 + b2PositionSolverManifold :659-708
// The SAME functions are used by the graph-coloured production kernels (one thread per
// constraint of a colour) and by the sequential parity kernel (one thread walks the list in
// order), so parity of the arithmetic is established once and the production mode differs
// only in visiting order.
//
// Constraint storage is structure-of-arrays in float4 planes indexed by solver slot s, so a
// warp reads 32 consecutive float4 per plane (512 B, fully coalesced); body velocities and
// positions are gathered / scattered as one float4 each.
#pragma once
#include "b2g_collide.cuh"

struct SolverPlanes {
  // velocity constraint (b2ContactVelocityConstraint, b2_contact_solver.h:35-76)
  float4* nf;    // normal.x, normal.y, friction, tangentSpeed
  float4* r1;    // rA1.x, rA1.y, rB1.x, rB1.y
  float4* r2;    // rA2.x, rA2.y, rB2.x, rB2.y
  float4* m1;    // normalMass1, tangentMass1, velocityBias1, normalMass2
  float4* m2;    // tangentMass2, velocityBias2, K11, K12
  float4* kk;    // K22, invK11, invK12, invK22   (normalMass = K^-1, symmetric)
  float4* mass;  // invMassA, invIA, invMassB, invIB
  int4* idx;     // bodyA, bodyB, pointCount (velocity rows), contact index
  float4* imp;   // normalImpulse1, tangentImpulse1, normalImpulse2, tangentImpulse2
  // position constraint (b2ContactPositionConstraint, b2_contact_solver.cpp:36-48)
  float4* pn;    // localNormal.x, localNormal.y, localPoint.x, localPoint.y
  float4* pp;    // localPoints[0].xy, localPoints[1].xy
  float4* pc;    // localCenterA.xy, localCenterB.xy
  float4* pr;    // radiusA, radiusB, bits(type | pointCount << 8), bits(island root)
};

#ifdef __CUDACC__
// ---- body state accessors ----------------------------------------------------------------------
// The solver functions are templated on where body velocities / positions live:
//   GlobalBodies  plain arrays in HBM, index = global body index (per-colour launches, seq mode)
//   TileBodies    the fused per-island-bin kernel: index >= 0 is a slot of the CTA's shared-memory
//                 tile, index < 0 is ~globalIndex of a body outside the tile (static bodies: read
//                 only, never stored because they are not movable)
struct GlobalBodies {
  float4* a;
  __device__ __forceinline__ float4 load(int i) const { return a[i]; }
  __device__ __forceinline__ void store(int i, float4 v) const { a[i] = v; }
};
struct TileBodies {
  float4* tile;
  const float4* __restrict__ global;
  __device__ __forceinline__ float4 load(int i) const { return i >= 0 ? tile[i] : global[~i]; }
  __device__ __forceinline__ void store(int i, float4 v) const { tile[i] = v; }
};

// ---- prepare -------------------------------------------------------------------------------
// bodyPos = (c.x, c.y, a, _), bodyVel = (v.x, v.y, w, _), bodyMass = (invMass, invI, _, _),
// bodyCenter = (localCenter.x, localCenter.y, _, _).
// idxA/idxB are what the iteration kernels will use to address the bodies (global index, or tile
// slot / ~global in the fused kernel); bodyA/bodyB are always global (mass, centre).
template <class PosAccess, class VelAccess>
__device__ __forceinline__ void prepare_constraint(const SolverPlanes& S, int s, int contactIndex, const Manifold& m,
                                                   int bodyA, int bodyB, int idxA, int idxB, float4 material,
                                                   float radiusA, float radiusB, const PosAccess& bodyPos,
                                                   const VelAccess& bodyVel, const float4* __restrict__ bodyMass,
                                                   const float4* __restrict__ bodyCenter, float dtRatio,
                                                   bool warmStarting, int root = 0) {
  float4 mAq = bodyMass[bodyA], mBq = bodyMass[bodyB];
  float mA = mAq.x, iA = mAq.y, mB = mBq.x, iB = mBq.y;
  float4 cenA = bodyCenter[bodyA], cenB = bodyCenter[bodyB];
  float2 localCenterA = make_float2(cenA.x, cenA.y), localCenterB = make_float2(cenB.x, cenB.y);
  float4 pA = bodyPos.load(idxA), pB = bodyPos.load(idxB);
  float4 vAq = bodyVel.load(idxA), vBq = bodyVel.load(idxB);
  float2 cA = make_float2(pA.x, pA.y), cB = make_float2(pB.x, pB.y);
  float aA = pA.z, aB = pB.z;
  float2 vA = make_float2(vAq.x, vAq.y), vB = make_float2(vBq.x, vBq.y);
  float wA = vAq.z, wB = vBq.z;

  const float friction = material.x, restitution = material.y, threshold = material.z, tangentSpeed = material.w;
  int pointCount = m.pointCount;

  // position constraint
  S.pn[s] = make_float4(m.localNormal.x, m.localNormal.y, m.localPoint.x, m.localPoint.y);
  S.pp[s] = make_float4(m.lp[0].x, m.lp[0].y, m.lp[1].x, m.lp[1].y);
  S.pc[s] = make_float4(localCenterA.x, localCenterA.y, localCenterB.x, localCenterB.y);
  // type and point count share a word; the freed lane carries the island root, so that a position pass
  // finds "is my island done?" in the constants it has staged anyway
  S.pr[s] = make_float4(radiusA, radiusB, __int_as_float(m.type | (pointCount << 8)), __int_as_float(root));
  S.mass[s] = make_float4(mA, iA, mB, iB);

  float imp[4];
  for (int j = 0; j < 2; ++j) {
    if (warmStarting && j < pointCount) {
      imp[2 * j] = dtRatio * m.normalImp[j];
      imp[2 * j + 1] = dtRatio * m.tangentImp[j];
    } else {
      imp[2 * j] = 0.0f;
      imp[2 * j + 1] = 0.0f;
    }
  }
  S.imp[s] = make_float4(imp[0], imp[1], imp[2], imp[3]);

  // InitializeVelocityConstraints
  Xf xfA = xf_from_sweep(cA, aA, localCenterA);
  Xf xfB = xf_from_sweep(cB, aB, localCenterB);
  WorldManifold wm;
  wm.points[1] = make_float2(0.0f, 0.0f);
  world_manifold(wm, m, xfA, radiusA, xfB, radiusB);
  float2 normal = wm.normal;
  float2 tangent = cross_vs(normal, 1.0f);

  float2 rA[2], rB[2];
  float normalMass[2], tangentMass[2], velocityBias[2];
  for (int j = 0; j < 2; ++j) {
    rA[j] = rB[j] = make_float2(0.0f, 0.0f);
    normalMass[j] = tangentMass[j] = velocityBias[j] = 0.0f;
    if (j < pointCount) {
      rA[j] = wm.points[j] - cA;
      rB[j] = wm.points[j] - cB;
      float rnA = cross2(rA[j], normal);
      float rnB = cross2(rB[j], normal);
      float kNormal = mA + mB + iA * rnA * rnA + iB * rnB * rnB;
      normalMass[j] = kNormal > 0.0f ? 1.0f / kNormal : 0.0f;
      float rtA = cross2(rA[j], tangent);
      float rtB = cross2(rB[j], tangent);
      float kTangent = mA + mB + iA * rtA * rtA + iB * rtB * rtB;
      tangentMass[j] = kTangent > 0.0f ? 1.0f / kTangent : 0.0f;
      // restitution bias from the relative normal velocity before warm starting
      float vRel = dot2(normal, vB + cross_sv(wB, rB[j]) - (vA + cross_sv(wA, rA[j])));
      if (vRel < -threshold) velocityBias[j] = -restitution * vRel;
    }
  }

  float k11 = 0.0f, k12 = 0.0f, k22 = 0.0f, n11 = 0.0f, n12 = 0.0f, n22 = 0.0f;
  int velPointCount = pointCount;
  if (pointCount == 2) {
    float rn1A = cross2(rA[0], normal);
    float rn1B = cross2(rB[0], normal);
    float rn2A = cross2(rA[1], normal);
    float rn2B = cross2(rB[1], normal);
    k11 = mA + mB + iA * rn1A * rn1A + iB * rn1B * rn1B;
    k22 = mA + mB + iA * rn2A * rn2A + iB * rn2B * rn2B;
    k12 = mA + mB + iA * rn1A * rn2A + iB * rn1B * rn2B;
    const float k_maxConditionNumber = 1000.0f;
    if (k11 * k11 < k_maxConditionNumber * (k11 * k22 - k12 * k12)) {
      // b2Mat22::GetInverse (b2_math.h:204-216) with ex=(k11,k12), ey=(k12,k22)
      float a = k11, b = k12, c = k12, d = k22;
      float det = a * d - b * c;
      if (det != 0.0f) det = 1.0f / det;
      n11 = det * d;
      n12 = -det * b;
      n22 = det * a;
    } else {
      velPointCount = 1;  // redundant rows: keep one (pc keeps 2 points)
    }
  }

  S.nf[s] = make_float4(normal.x, normal.y, friction, tangentSpeed);
  S.r1[s] = make_float4(rA[0].x, rA[0].y, rB[0].x, rB[0].y);
  S.r2[s] = make_float4(rA[1].x, rA[1].y, rB[1].x, rB[1].y);
  S.m1[s] = make_float4(normalMass[0], tangentMass[0], velocityBias[0], normalMass[1]);
  S.m2[s] = make_float4(tangentMass[1], velocityBias[1], k11, k12);
  S.kk[s] = make_float4(k22, n11, n12, n22);
  S.idx[s] = make_int4(idxA, idxB, velPointCount, contactIndex);
}

// A body whose inverse mass and inertia are both zero (static, kinematic, massless) is never
// changed by a contact row; skipping its store keeps coloured launches race-free even though
// such bodies are shared by many constraints of one colour.
__device__ __forceinline__ bool movable(float invM, float invI) { return invM != 0.0f || invI != 0.0f; }

template <class VelAccess>
__device__ __forceinline__ void warm_start_constraint(const SolverPlanes& S, int s, const VelAccess& bodyVel) {
  int4 ix = S.idx[s];
  float4 ms = S.mass[s];
  float mA = ms.x, iA = ms.y, mB = ms.z, iB = ms.w;
  float4 nf = S.nf[s];
  float2 normal = make_float2(nf.x, nf.y);
  float2 tangent = cross_vs(normal, 1.0f);
  float4 r1 = S.r1[s], r2 = S.r2[s], imp = S.imp[s];
  float4 vAq = bodyVel.load(ix.x), vBq = bodyVel.load(ix.y);
  float2 vA = make_float2(vAq.x, vAq.y), vB = make_float2(vBq.x, vBq.y);
  float wA = vAq.z, wB = vBq.z;
  {
    float2 P = imp.x * normal + imp.y * tangent;
    float2 rA = make_float2(r1.x, r1.y), rB = make_float2(r1.z, r1.w);
    wA -= iA * cross2(rA, P);
    vA -= mA * P;
    wB += iB * cross2(rB, P);
    vB += mB * P;
  }
  if (ix.z == 2) {
    float2 P = imp.z * normal + imp.w * tangent;
    float2 rA = make_float2(r2.x, r2.y), rB = make_float2(r2.z, r2.w);
    wA -= iA * cross2(rA, P);
    vA -= mA * P;
    wB += iB * cross2(rB, P);
    vB += mB * P;
  }
  if (movable(mA, iA)) bodyVel.store(ix.x, make_float4(vA.x, vA.y, wA, vAq.w));
  if (movable(mB, iB)) bodyVel.store(ix.y, make_float4(vB.x, vB.y, wB, vBq.w));
}

template <class VelAccess>
__device__ __forceinline__ void solve_velocity_constraint(const SolverPlanes& S, int s, const VelAccess& bodyVel) {
  int4 ix = S.idx[s];
  float4 ms = S.mass[s];
  float mA = ms.x, iA = ms.y, mB = ms.z, iB = ms.w;
  float4 nf = S.nf[s];
  float2 normal = make_float2(nf.x, nf.y);
  float friction = nf.z, tangentSpeed = nf.w;
  float4 r1 = S.r1[s], r2 = S.r2[s], m1 = S.m1[s], m2 = S.m2[s], imp = S.imp[s];
  const int pointCount = ix.z;
  float2 rA1 = make_float2(r1.x, r1.y), rB1 = make_float2(r1.z, r1.w);
  float2 rA2 = make_float2(r2.x, r2.y), rB2 = make_float2(r2.z, r2.w);

  float4 vAq = bodyVel.load(ix.x), vBq = bodyVel.load(ix.y);
  float2 vA = make_float2(vAq.x, vAq.y), vB = make_float2(vBq.x, vBq.y);
  float wA = vAq.z, wB = vBq.z;

  float2 tangent = cross_vs(normal, 1.0f);
  // friction rows first: non-penetration is more important than friction
  {
    float2 dv = vB + cross_sv(wB, rB1) - (vA + cross_sv(wA, rA1));
    float vt = dot2(dv, tangent) - tangentSpeed;
    float lambda = m1.y * (-vt);
    float maxFriction = friction * imp.x;
    float newImpulse = clampf(imp.y + lambda, -maxFriction, maxFriction);
    lambda = newImpulse - imp.y;
    imp.y = newImpulse;
    float2 P = lambda * tangent;
    vA -= mA * P;
    wA -= iA * cross2(rA1, P);
    vB += mB * P;
    wB += iB * cross2(rB1, P);
  }
  if (pointCount == 2) {
    float2 dv = vB + cross_sv(wB, rB2) - (vA + cross_sv(wA, rA2));
    float vt = dot2(dv, tangent) - tangentSpeed;
    float lambda = m2.x * (-vt);
    float maxFriction = friction * imp.z;
    float newImpulse = clampf(imp.w + lambda, -maxFriction, maxFriction);
    lambda = newImpulse - imp.w;
    imp.w = newImpulse;
    float2 P = lambda * tangent;
    vA -= mA * P;
    wA -= iA * cross2(rA2, P);
    vB += mB * P;
    wB += iB * cross2(rB2, P);
  }

  if (pointCount == 1) {
    float2 dv = vB + cross_sv(wB, rB1) - (vA + cross_sv(wA, rA1));
    float vn = dot2(dv, normal);
    float lambda = -m1.x * (vn - m1.z);
    float newImpulse = maxf_(imp.x + lambda, 0.0f);
    lambda = newImpulse - imp.x;
    imp.x = newImpulse;
    float2 P = lambda * normal;
    vA -= mA * P;
    wA -= iA * cross2(rA1, P);
    vB += mB * P;
    wB += iB * cross2(rB1, P);
  } else {
    // 2-point block solver: vn = K x + b', x >= 0, vn >= 0, complementary; enumerate 4 cases
    float4 kk = S.kk[s];
    const float k11 = m2.z, k12 = m2.w, k22 = kk.x;
    const float n11 = kk.y, n12 = kk.z, n22 = kk.w;
    float2 a = make_float2(imp.x, imp.z);
    float2 dv1 = vB + cross_sv(wB, rB1) - (vA + cross_sv(wA, rA1));
    float2 dv2 = vB + cross_sv(wB, rB2) - (vA + cross_sv(wA, rA2));
    float vn1 = dot2(dv1, normal);
    float vn2 = dot2(dv2, normal);
    float2 b = make_float2(vn1 - m1.z, vn2 - m2.y);
    // b -= K a   (b2Mul(Mat22, Vec2): ex.x*v.x + ey.x*v.y , ex.y*v.x + ey.y*v.y)
    b.x -= k11 * a.x + k12 * a.y;
    b.y -= k12 * a.x + k22 * a.y;

    float2 x;
    bool solved = false;
    // case 1: both rows active, x = -K^-1 b'
    x.x = -(n11 * b.x + n12 * b.y);
    x.y = -(n12 * b.x + n22 * b.y);
    if (x.x >= 0.0f && x.y >= 0.0f) solved = true;
    if (!solved) {
      // case 2: vn1 = 0, x2 = 0
      x.x = -m1.x * b.x;
      x.y = 0.0f;
      vn2 = k12 * x.x + b.y;
      if (x.x >= 0.0f && vn2 >= 0.0f) solved = true;
    }
    if (!solved) {
      // case 3: vn2 = 0, x1 = 0
      x.x = 0.0f;
      x.y = -m1.w * b.y;
      vn1 = k12 * x.y + b.x;
      if (x.y >= 0.0f && vn1 >= 0.0f) solved = true;
    }
    if (!solved) {
      // case 4: x = 0
      x.x = 0.0f;
      x.y = 0.0f;
      if (b.x >= 0.0f && b.y >= 0.0f) solved = true;
    }
    if (solved) {
      float2 d = x - a;
      float2 P1 = d.x * normal;
      float2 P2 = d.y * normal;
      vA -= mA * (P1 + P2);
      wA -= iA * (cross2(rA1, P1) + cross2(rA2, P2));
      vB += mB * (P1 + P2);
      wB += iB * (cross2(rB1, P1) + cross2(rB2, P2));
      imp.x = x.x;
      imp.z = x.y;
    }
    // else: no solution, keep the impulses (b2_contact_solver.cpp:629-630)
  }

  S.imp[s] = imp;
  if (movable(mA, iA)) bodyVel.store(ix.x, make_float4(vA.x, vA.y, wA, vAq.w));
  if (movable(mB, iB)) bodyVel.store(ix.y, make_float4(vB.x, vB.y, wB, vBq.w));
}

// returns the smallest separation seen (<= 0 contributes to the island's minSeparation)
template <class PosAccess>
__device__ __forceinline__ float solve_position_constraint(const SolverPlanes& S, int s, const PosAccess& bodyPos) {
  int4 ix = S.idx[s];
  float4 ms = S.mass[s];
  float mA = ms.x, iA = ms.y, mB = ms.z, iB = ms.w;
  float4 pn = S.pn[s], pp = S.pp[s], pcen = S.pc[s], pr = S.pr[s];
  float2 localNormal = make_float2(pn.x, pn.y), localPoint = make_float2(pn.z, pn.w);
  float2 localCenterA = make_float2(pcen.x, pcen.y), localCenterB = make_float2(pcen.z, pcen.w);
  float radiusA = pr.x, radiusB = pr.y;
  const int type = __float_as_int(pr.z) & 0xff;
  const int pointCount = __float_as_int(pr.z) >> 8;

  float4 pAq = bodyPos.load(ix.x), pBq = bodyPos.load(ix.y);
  float2 cA = make_float2(pAq.x, pAq.y), cB = make_float2(pBq.x, pBq.y);
  float aA = pAq.z, aB = pBq.z;
  float minSeparation = 0.0f;

  // b2PositionSolverManifold::Initialize recomputes both transforms from (c, a) for every point (b2_contact_solver.cpp
  // :746-752).  A transform is a pure function of (c, a), and a point whose correction is zero leaves (c, a) of both
  // bodies bit-for-bit unchanged, so the second point may reuse the first point's transform of any body that did
  // not move: same floats, half the sin / cos evaluations on a settled pile.
  Xf xfA = xf_from_sweep(cA, aA, localCenterA);
  Xf xfB = xf_from_sweep(cB, aB, localCenterB);
  for (int j = 0; j < pointCount; ++j) {
    float2 lpj = j == 0 ? make_float2(pp.x, pp.y) : make_float2(pp.z, pp.w);
    float2 normal, point;
    float separation;
    if (type == 0) {
      float2 pointA = xf_mul(xfA, localPoint);
      float2 pointB = xf_mul(xfB, make_float2(pp.x, pp.y));
      normal = pointB - pointA;
      normalize2(normal);
      point = 0.5f * (pointA + pointB);
      separation = dot2(pointB - pointA, normal) - radiusA - radiusB;
    } else if (type == 1) {
      normal = rot_mul(xfA.q, localNormal);
      float2 planePoint = xf_mul(xfA, localPoint);
      float2 clipPoint = xf_mul(xfB, lpj);
      separation = dot2(clipPoint - planePoint, normal) - radiusA - radiusB;
      point = clipPoint;
    } else {
      normal = rot_mul(xfB.q, localNormal);
      float2 planePoint = xf_mul(xfB, localPoint);
      float2 clipPoint = xf_mul(xfA, lpj);
      separation = dot2(clipPoint - planePoint, normal) - radiusA - radiusB;
      point = clipPoint;
      normal = -normal;
    }
    float2 rA = point - cA;
    float2 rB = point - cB;
    minSeparation = minf_(minSeparation, separation);
    float C = clampf(B2G_BAUMGARTE * (separation + B2G_LINEAR_SLOP), -B2G_MAX_LINEAR_CORRECTION, 0.0f);
    float rnA = cross2(rA, normal);
    float rnB = cross2(rB, normal);
    float K = mA + mB + iA * rnA * rnA + iB * rnB * rnB;
    float impulse = K > 0.0f ? -C / K : 0.0f;
    float2 P = impulse * normal;
    cA -= mA * P;
    aA -= iA * cross2(rA, P);
    cB += mB * P;
    aB += iB * cross2(rB, P);
    if (j + 1 < pointCount && impulse != 0.0f) {
      if (movable(mA, iA)) xfA = xf_from_sweep(cA, aA, localCenterA);
      if (movable(mB, iB)) xfB = xf_from_sweep(cB, aB, localCenterB);
    }
  }
  if (movable(mA, iA)) bodyPos.store(ix.x, make_float4(cA.x, cA.y, aA, pAq.w));
  if (movable(mB, iB)) bodyPos.store(ix.y, make_float4(cB.x, cB.y, aB, pBq.w));
  return minSeparation;
}
#endif  // __CUDACC__

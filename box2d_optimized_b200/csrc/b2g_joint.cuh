// b2g_joint.cuh — revolute joint (with motor and limits) on the device.
//
// One call = the body of b2RevoluteJoint::{InitVelocityConstraints, SolveVelocityConstraints,
// SolvePositionConstraints} (src/dynamics/b2_revolute_joint.cpp:73-321) for one joint, with the
// same operation order.  Needed by the tumbler config (testbed/benchmarks/benchmarks.h:177-186);
// SURVEY.md §8(f) rank 1.  Joints are few (one per tumbler world): the thread that owns an island's
// serial work walks that island's joints in joint-index order — the reference's per-island loop
// (b2_island.cpp:323-338, 396-401) — before the contact colours in every velocity iteration and
// after them in every position iteration.
#pragma once
#include "b2g_solver.cuh"

#define B2G_JOINT_LIMIT 1u
#define B2G_JOINT_MOTOR 2u
#define B2G_JOINT_COLLIDE_CONNECTED 4u

// per-step work area of one joint (plain struct in global memory; one thread touches it)
struct JointWork {
  float2 rA, rB;
  float k11, k12, k22;  // K = [k11 k12; k12 k22]
  float axialMass, angle;
  float mA, mB, iA, iB;
  float2 lcA, lcB;
  int ia, ib;  // body addresses for the accessors (tile slot, ~global, or global index)
};

struct JointArraysDev {
  const int2* bodies;
  const float4* anchors;   // localAnchorA.xy, localAnchorB.xy
  const float4* params0;   // referenceAngle, lowerAngle, upperAngle, maxMotorTorque
  const float4* params1;   // motorSpeed, bits(flags), 0, 0
  float4* state;           // impulse.x, impulse.y, motorImpulse, lowerImpulse
  float* upper;            // upperImpulse
  JointWork* work;
};

#ifdef __CUDACC__
__device__ __forceinline__ float2 mat22_solve(float a11, float a12, float a21, float a22, float2 b) {
  float det = a11 * a22 - a12 * a21;
  if (det != 0.0f) det = 1.0f / det;
  return make_float2(det * (a22 * b.x - a12 * b.y), det * (a11 * b.y - a21 * b.x));
}

template <class PosAccess, class VelAccess>
__device__ __forceinline__ void joint_init(const JointArraysDev& J, int j, int ia, int ib, const PosAccess& pos,
                                           const VelAccess& vel, const float4* __restrict__ bodyMass,
                                           const float4* __restrict__ bodyCenter, float dtRatio, bool warmStarting) {
  JointWork w;
  int2 bd = J.bodies[j];
  float4 mAq = bodyMass[bd.x], mBq = bodyMass[bd.y];
  float4 cAq = bodyCenter[bd.x], cBq = bodyCenter[bd.y];
  w.ia = ia;
  w.ib = ib;
  w.mA = mAq.x; w.iA = mAq.y; w.mB = mBq.x; w.iB = mBq.y;
  w.lcA = make_float2(cAq.x, cAq.y);
  w.lcB = make_float2(cBq.x, cBq.y);
  float4 an = J.anchors[j], p0 = J.params0[j], p1 = J.params1[j];
  uint32_t flags = __float_as_uint(p1.y);
  float4 pA = pos.load(ia), pB = pos.load(ib);
  float4 vAq = vel.load(ia), vBq = vel.load(ib);
  float aA = pA.z, aB = pB.z;
  float2 vA = make_float2(vAq.x, vAq.y), vB = make_float2(vBq.x, vBq.y);
  float wA = vAq.z, wB = vBq.z;
  Rot qA = rot_set(aA), qB = rot_set(aB);
  w.rA = rot_mul(qA, make_float2(an.x, an.y) - w.lcA);
  w.rB = rot_mul(qB, make_float2(an.z, an.w) - w.lcB);
  float mA = w.mA, mB = w.mB, iA = w.iA, iB = w.iB;
  w.k11 = mA + mB + w.rA.y * w.rA.y * iA + w.rB.y * w.rB.y * iB;
  w.k12 = -w.rA.y * w.rA.x * iA - w.rB.y * w.rB.x * iB;
  w.k22 = mA + mB + w.rA.x * w.rA.x * iA + w.rB.x * w.rB.x * iB;
  w.axialMass = iA + iB;
  bool fixedRotation;
  if (w.axialMass > 0.0f) {
    w.axialMass = 1.0f / w.axialMass;
    fixedRotation = false;
  } else {
    fixedRotation = true;
  }
  w.angle = aB - aA - p0.x;
  float4 st = J.state[j];
  float upper = J.upper[j];
  if (!(flags & B2G_JOINT_LIMIT) || fixedRotation) {
    st.w = 0.0f;
    upper = 0.0f;
  }
  if (!(flags & B2G_JOINT_MOTOR) || fixedRotation) st.z = 0.0f;
  if (warmStarting) {
    st.x *= dtRatio;
    st.y *= dtRatio;
    st.z *= dtRatio;
    st.w *= dtRatio;
    upper *= dtRatio;
    float axialImpulse = st.z + st.w - upper;
    float2 P = make_float2(st.x, st.y);
    vA -= mA * P;
    wA -= iA * (cross2(w.rA, P) + axialImpulse);
    vB += mB * P;
    wB += iB * (cross2(w.rB, P) + axialImpulse);
  } else {
    st = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    upper = 0.0f;
  }
  J.state[j] = st;
  J.upper[j] = upper;
  J.work[j] = w;
  if (movable(mA, iA)) vel.store(ia, make_float4(vA.x, vA.y, wA, vAq.w));
  if (movable(mB, iB)) vel.store(ib, make_float4(vB.x, vB.y, wB, vBq.w));
}

template <class VelAccess>
__device__ __forceinline__ void joint_solve_velocity(const JointArraysDev& J, int j, const VelAccess& vel, float dt,
                                                     float inv_dt) {
  JointWork w = J.work[j];
  float4 p0 = J.params0[j], p1 = J.params1[j];
  uint32_t flags = __float_as_uint(p1.y);
  float4 st = J.state[j];
  float upper = J.upper[j];
  float4 vAq = vel.load(w.ia), vBq = vel.load(w.ib);
  float2 vA = make_float2(vAq.x, vAq.y), vB = make_float2(vBq.x, vBq.y);
  float wA = vAq.z, wB = vBq.z;
  float mA = w.mA, mB = w.mB, iA = w.iA, iB = w.iB;
  bool fixedRotation = (iA + iB == 0.0f);
  if ((flags & B2G_JOINT_MOTOR) && !fixedRotation) {
    float Cdot = wB - wA - p1.x;
    float impulse = -w.axialMass * Cdot;
    float oldImpulse = st.z;
    float maxImpulse = dt * p0.w;
    st.z = clampf(st.z + impulse, -maxImpulse, maxImpulse);
    impulse = st.z - oldImpulse;
    wA -= iA * impulse;
    wB += iB * impulse;
  }
  if ((flags & B2G_JOINT_LIMIT) && !fixedRotation) {
    {
      float C = w.angle - p0.y;
      float Cdot = wB - wA;
      float impulse = -w.axialMass * (Cdot + maxf_(C, 0.0f) * inv_dt);
      float oldImpulse = st.w;
      st.w = maxf_(st.w + impulse, 0.0f);
      impulse = st.w - oldImpulse;
      wA -= iA * impulse;
      wB += iB * impulse;
    }
    {
      float C = p0.z - w.angle;
      float Cdot = wA - wB;
      float impulse = -w.axialMass * (Cdot + maxf_(C, 0.0f) * inv_dt);
      float oldImpulse = upper;
      upper = maxf_(upper + impulse, 0.0f);
      impulse = upper - oldImpulse;
      wA += iA * impulse;
      wB -= iB * impulse;
    }
  }
  {
    float2 Cdot = vB + cross_sv(wB, w.rB) - vA - cross_sv(wA, w.rA);
    float2 impulse = mat22_solve(w.k11, w.k12, w.k12, w.k22, -Cdot);
    st.x += impulse.x;
    st.y += impulse.y;
    vA -= mA * impulse;
    wA -= iA * cross2(w.rA, impulse);
    vB += mB * impulse;
    wB += iB * cross2(w.rB, impulse);
  }
  J.state[j] = st;
  J.upper[j] = upper;
  if (movable(mA, iA)) vel.store(w.ia, make_float4(vA.x, vA.y, wA, vAq.w));
  if (movable(mB, iB)) vel.store(w.ib, make_float4(vB.x, vB.y, wB, vBq.w));
}

// returns true when the joint's position error is within tolerance (jointOkay)
template <class PosAccess>
__device__ __forceinline__ bool joint_solve_position(const JointArraysDev& J, int j, const PosAccess& pos) {
  JointWork w = J.work[j];
  float4 an = J.anchors[j], p0 = J.params0[j], p1 = J.params1[j];
  uint32_t flags = __float_as_uint(p1.y);
  float4 pAq = pos.load(w.ia), pBq = pos.load(w.ib);
  float2 cA = make_float2(pAq.x, pAq.y), cB = make_float2(pBq.x, pBq.y);
  float aA = pAq.z, aB = pBq.z;
  float angularError = 0.0f, positionError = 0.0f;
  bool fixedRotation = (w.iA + w.iB == 0.0f);
  if ((flags & B2G_JOINT_LIMIT) && !fixedRotation) {
    float angle = aB - aA - p0.x;
    float C = 0.0f;
    if (absf_(p0.z - p0.y) < 2.0f * B2G_ANGULAR_SLOP) {
      C = clampf(angle - p0.y, -B2G_MAX_ANGULAR_CORRECTION, B2G_MAX_ANGULAR_CORRECTION);
    } else if (angle <= p0.y) {
      C = clampf(angle - p0.y + B2G_ANGULAR_SLOP, -B2G_MAX_ANGULAR_CORRECTION, 0.0f);
    } else if (angle >= p0.z) {
      C = clampf(angle - p0.z - B2G_ANGULAR_SLOP, 0.0f, B2G_MAX_ANGULAR_CORRECTION);
    }
    float limitImpulse = -w.axialMass * C;
    aA -= w.iA * limitImpulse;
    aB += w.iB * limitImpulse;
    angularError = absf_(C);
  }
  {
    Rot qA = rot_set(aA), qB = rot_set(aB);
    float2 rA = rot_mul(qA, make_float2(an.x, an.y) - w.lcA);
    float2 rB = rot_mul(qB, make_float2(an.z, an.w) - w.lcB);
    float2 C = cB + rB - cA - rA;
    positionError = len2(C);
    float mA = w.mA, mB = w.mB, iA = w.iA, iB = w.iB;
    float k11 = mA + mB + iA * rA.y * rA.y + iB * rB.y * rB.y;
    float k21 = -iA * rA.x * rA.y - iB * rB.x * rB.y;
    float k22 = mA + mB + iA * rA.x * rA.x + iB * rB.x * rB.x;
    float2 impulse = -mat22_solve(k11, k21, k21, k22, C);
    cA -= mA * impulse;
    aA -= iA * cross2(rA, impulse);
    cB += mB * impulse;
    aB += iB * cross2(rB, impulse);
  }
  if (movable(w.mA, w.iA)) pos.store(w.ia, make_float4(cA.x, cA.y, aA, pAq.w));
  if (movable(w.mB, w.iB)) pos.store(w.ib, make_float4(cB.x, cB.y, aB, pBq.w));
  return positionError <= B2G_LINEAR_SLOP && angularError <= B2G_ANGULAR_SLOP;
}
#endif

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
// joint type in bits 8-11 of the flags word (include/b2cuda.h b2gJointArrays)
#define B2G_JOINT_TYPE(flags) (((flags) >> 8) & 0xFu)
#define B2G_JOINT_REVOLUTE 0u
#define B2G_JOINT_DISTANCE 1u
#define B2G_JOINT_WELD 2u
#define B2G_JOINT_PRISMATIC 3u
#define B2G_JOINT_WHEEL 4u
#define B2G_JOINT_FRICTION 5u
#define B2G_JOINT_MOTOR_JOINT 6u
#define B2G_JOINT_MOUSE 7u

// per-step work area of one joint (plain struct in global memory; one thread touches it)
struct JointWork {
  float2 rA, rB;
  float k11, k12, k22;  // K = [k11 k12; k12 k22]
  float axialMass, angle;
  float mA, mB, iA, iB;
  float2 lcA, lcB;
  int ia, ib;  // body addresses for the accessors (tile slot, ~global, or global index)
  // distance joint (b2_distance_joint.h:152-169): axis, its effective masses, soft-constraint terms
  float2 u;
  float dMass, softMass, gamma, bias, currentLength;
  // weld joint (b2_weld_joint.h:112-126): 3x3 effective mass, columns ex, ey, ez
  float3 wex, wey, wez;
  // prismatic joint (b2_prismatic_joint.h:182-190): world axis and its perpendicular, their lever arms
  // (k11/k12/k22, axialMass and `angle` = translation are shared with the revolute fields)
  // friction / motor joint (b2_friction_joint.h:103-104, b2_motor_joint.h:120-128): k11/k12/k22 = m_linearMass
  // (symmetric), axialMass = m_angularMass, u = m_linearError, angle = m_angularError
  // wheel joint (b2_wheel_joint.h:219-229): axis = m_ax, perp = m_ay, a1/a2 = m_sAx/m_sBx, s1/s2 = m_sAy/m_sBy,
  // k11 = m_mass, dMass = m_motorMass, softMass = m_springMass, bias, gamma, angle = m_translation
  float2 axis, perp;
  float a1, a2, s1, s2;
};

struct JointArraysDev {
  const int2* bodies;
  const float4* anchors;   // localAnchorA.xy, localAnchorB.xy
  const float4* params0;   // referenceAngle, lowerAngle, upperAngle, maxMotorTorque
  const float4* params1;   // motorSpeed, bits(flags), 0, 0
  const float4* params2;   // third parameter quad (wheel joints)
  float4* state;           // impulse.x, impulse.y, motorImpulse, lowerImpulse
  float* upper;            // upperImpulse
  JointWork* work;
  float h;                 // this step's dt (soft constraints)
};
// mouse joints: anchors = targetA.x, targetA.y, localAnchorB.x, localAnchorB.y; params0 = maxForce, stiffness,
// damping, 0; params1 = 0, bits(flags | 7 << 8), 0, 0; state = impulse.x, impulse.y, -, -.  Only bodyB is moved.
// friction joints: params0 = maxForce, maxTorque, 0, 0; params1 = 0, bits(flags | 5 << 8), 0, 0;
// state = linearImpulse.x, linearImpulse.y, angularImpulse, -
// motor joints: anchors = linearOffset.x, linearOffset.y, 0, 0; params0 = maxForce, maxTorque, correctionFactor,
// angularOffset; params1 = 0, bits(flags | 6 << 8), 0, 0; state as the friction joint
// wheel joints: params0 = stiffness, lowerTranslation, upperTranslation, maxMotorTorque;
// params1 = motorSpeed, bits(flags | 4 << 8), localXAxisA.x, localXAxisA.y (as given: the reference does not
// normalise it); params2 = damping, 0, 0, 0;
// state = impulse, springImpulse, motorImpulse, lowerImpulse; upper = upperImpulse
// prismatic joints: params0 = referenceAngle, lowerTranslation, upperTranslation, maxMotorForce;
// params1 = motorSpeed, bits(flags | 3 << 8), localXAxisA.x, localXAxisA.y (unit length);
// state = impulse.x, impulse.y, motorImpulse, lowerImpulse; upper = upperImpulse
// weld joints: params0 = referenceAngle, stiffness, damping, 0; params1 = 0, bits(flags | 2 << 8), 0, 0;
// state = impulse.x, impulse.y, impulse.z, -
// distance joints reuse the arrays: params0 = length, minLength, maxLength, stiffness;
// params1 = damping, bits(flags | type << 8), 0, 0; state = impulse, -, -, lowerImpulse; upper = upperImpulse

#ifdef __CUDACC__
__device__ __forceinline__ float2 mat22_solve(float a11, float a12, float a21, float a22, float2 b) {
  float det = a11 * a22 - a12 * a21;
  if (det != 0.0f) det = 1.0f / det;
  return make_float2(det * (a22 * b.x - a12 * b.y), det * (a11 * b.y - a21 * b.x));
}

template <class PosAccess, class VelAccess>
__device__ __forceinline__ void revolute_init(const JointArraysDev& J, int j, int ia, int ib, const PosAccess& pos,
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
__device__ __forceinline__ void revolute_solve_velocity(const JointArraysDev& J, int j, const VelAccess& vel, float dt,
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
__device__ __forceinline__ bool revolute_solve_position(const JointArraysDev& J, int j, const PosAccess& pos) {
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

// ---- distance joint: b2DistanceJoint::{InitVelocityConstraints, SolveVelocityConstraints,
// SolvePositionConstraints} (src/dynamics/b2_distance_joint.cpp:76-303), rigid, soft (stiffness /
// damping) and with min / max length limits ------------------------------------------------------
template <class PosAccess, class VelAccess>
__device__ __forceinline__ void distance_init(const JointArraysDev& J, int j, int ia, int ib, const PosAccess& pos,
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
  w.k11 = w.k12 = w.k22 = w.axialMass = w.angle = 0.0f;
  float4 an = J.anchors[j], p0 = J.params0[j], p1 = J.params1[j];
  const float length = p0.x, minLength = p0.y, maxLength = p0.z, stiffness = p0.w, damping = p1.x;
  float4 pA = pos.load(ia), pB = pos.load(ib);
  float4 vAq = vel.load(ia), vBq = vel.load(ib);
  float2 cA = make_float2(pA.x, pA.y), cB = make_float2(pB.x, pB.y);
  float2 vA = make_float2(vAq.x, vAq.y), vB = make_float2(vBq.x, vBq.y);
  float wA = vAq.z, wB = vBq.z;
  Rot qA = rot_set(pA.z), qB = rot_set(pB.z);
  w.rA = rot_mul(qA, make_float2(an.x, an.y) - w.lcA);
  w.rB = rot_mul(qB, make_float2(an.z, an.w) - w.lcB);
  w.u = cB + w.rB - cA - w.rA;
  float4 st = J.state[j];
  float upper = J.upper[j];
  w.currentLength = len2(w.u);
  if (w.currentLength > B2G_LINEAR_SLOP) {
    w.u = (1.0f / w.currentLength) * w.u;
  } else {  // singular: the anchors coincide
    w.u = make_float2(0.0f, 0.0f);
    st.x = 0.0f;
    st.w = 0.0f;
    upper = 0.0f;
  }
  float crAu = cross2(w.rA, w.u), crBu = cross2(w.rB, w.u);
  float invMass = w.mA + w.iA * crAu * crAu + w.mB + w.iB * crBu * crBu;
  w.dMass = invMass != 0.0f ? 1.0f / invMass : 0.0f;
  if (stiffness > 0.0f && minLength < maxLength) {  // soft
    float C = w.currentLength - length;
    float h = J.h;
    w.gamma = h * (damping + h * stiffness);
    w.gamma = w.gamma != 0.0f ? 1.0f / w.gamma : 0.0f;
    w.bias = C * h * stiffness * w.gamma;
    invMass += w.gamma;
    w.softMass = invMass != 0.0f ? 1.0f / invMass : 0.0f;
  } else {  // rigid
    w.gamma = 0.0f;
    w.bias = 0.0f;
    w.softMass = w.dMass;
  }
  if (warmStarting) {
    st.x *= dtRatio;
    st.w *= dtRatio;
    upper *= dtRatio;
    float2 P = (st.x + st.w - upper) * w.u;
    vA -= w.mA * P;
    wA -= w.iA * cross2(w.rA, P);
    vB += w.mB * P;
    wB += w.iB * cross2(w.rB, P);
  } else {
    st.x = 0.0f;
  }
  J.state[j] = st;
  J.upper[j] = upper;
  J.work[j] = w;
  if (movable(w.mA, w.iA)) vel.store(ia, make_float4(vA.x, vA.y, wA, vAq.w));
  if (movable(w.mB, w.iB)) vel.store(ib, make_float4(vB.x, vB.y, wB, vBq.w));
}

template <class VelAccess>
__device__ __forceinline__ void distance_solve_velocity(const JointArraysDev& J, int j, const VelAccess& vel, float inv_dt) {
  JointWork w = J.work[j];
  float4 p0 = J.params0[j];
  const float minLength = p0.y, maxLength = p0.z, stiffness = p0.w;
  float4 st = J.state[j];
  float upper = J.upper[j];
  float4 vAq = vel.load(w.ia), vBq = vel.load(w.ib);
  float2 vA = make_float2(vAq.x, vAq.y), vB = make_float2(vBq.x, vBq.y);
  float wA = vAq.z, wB = vBq.z;
  if (minLength < maxLength) {
    if (stiffness > 0.0f) {
      float2 vpA = vA + cross_sv(wA, w.rA), vpB = vB + cross_sv(wB, w.rB);
      float Cdot = dot2(w.u, vpB - vpA);
      float impulse = -w.softMass * (Cdot + w.bias + w.gamma * st.x);
      st.x += impulse;
      float2 P = impulse * w.u;
      vA -= w.mA * P;
      wA -= w.iA * cross2(w.rA, P);
      vB += w.mB * P;
      wB += w.iB * cross2(w.rB, P);
    }
    {  // lower limit
      float C = w.currentLength - minLength;
      float bias = maxf_(0.0f, C) * inv_dt;
      float2 vpA = vA + cross_sv(wA, w.rA), vpB = vB + cross_sv(wB, w.rB);
      float Cdot = dot2(w.u, vpB - vpA);
      float impulse = -w.dMass * (Cdot + bias);
      float oldImpulse = st.w;
      st.w = maxf_(0.0f, st.w + impulse);
      impulse = st.w - oldImpulse;
      float2 P = impulse * w.u;
      vA -= w.mA * P;
      wA -= w.iA * cross2(w.rA, P);
      vB += w.mB * P;
      wB += w.iB * cross2(w.rB, P);
    }
    {  // upper limit
      float C = maxLength - w.currentLength;
      float bias = maxf_(0.0f, C) * inv_dt;
      float2 vpA = vA + cross_sv(wA, w.rA), vpB = vB + cross_sv(wB, w.rB);
      float Cdot = dot2(w.u, vpA - vpB);
      float impulse = -w.dMass * (Cdot + bias);
      float oldImpulse = upper;
      upper = maxf_(0.0f, upper + impulse);
      impulse = upper - oldImpulse;
      float2 P = -impulse * w.u;
      vA -= w.mA * P;
      wA -= w.iA * cross2(w.rA, P);
      vB += w.mB * P;
      wB += w.iB * cross2(w.rB, P);
    }
  } else {  // equal limits: a rigid rod
    float2 vpA = vA + cross_sv(wA, w.rA), vpB = vB + cross_sv(wB, w.rB);
    float Cdot = dot2(w.u, vpB - vpA);
    float impulse = -w.dMass * Cdot;
    st.x += impulse;
    float2 P = impulse * w.u;
    vA -= w.mA * P;
    wA -= w.iA * cross2(w.rA, P);
    vB += w.mB * P;
    wB += w.iB * cross2(w.rB, P);
  }
  J.state[j] = st;
  J.upper[j] = upper;
  if (movable(w.mA, w.iA)) vel.store(w.ia, make_float4(vA.x, vA.y, wA, vAq.w));
  if (movable(w.mB, w.iB)) vel.store(w.ib, make_float4(vB.x, vB.y, wB, vBq.w));
}

template <class PosAccess>
__device__ __forceinline__ bool distance_solve_position(const JointArraysDev& J, int j, const PosAccess& pos) {
  JointWork w = J.work[j];
  float4 an = J.anchors[j], p0 = J.params0[j];
  const float minLength = p0.y, maxLength = p0.z;
  float4 pAq = pos.load(w.ia), pBq = pos.load(w.ib);
  float2 cA = make_float2(pAq.x, pAq.y), cB = make_float2(pBq.x, pBq.y);
  float aA = pAq.z, aB = pBq.z;
  Rot qA = rot_set(aA), qB = rot_set(aB);
  float2 rA = rot_mul(qA, make_float2(an.x, an.y) - w.lcA);
  float2 rB = rot_mul(qB, make_float2(an.z, an.w) - w.lcB);
  float2 u = cB + rB - cA - rA;
  float length = normalize2(u);
  float C;
  if (minLength == maxLength) C = length - minLength;
  else if (length < minLength) C = length - minLength;
  else if (maxLength < length) C = length - maxLength;
  else return true;
  float impulse = -w.dMass * C;
  float2 P = impulse * u;
  cA -= w.mA * P;
  aA -= w.iA * cross2(rA, P);
  cB += w.mB * P;
  aB += w.iB * cross2(rB, P);
  if (movable(w.mA, w.iA)) pos.store(w.ia, make_float4(cA.x, cA.y, aA, pAq.w));
  if (movable(w.mB, w.iB)) pos.store(w.ib, make_float4(cB.x, cB.y, aB, pBq.w));
  return absf_(C) < B2G_LINEAR_SLOP;
}

// ---- weld joint: b2WeldJoint::{InitVelocityConstraints, SolveVelocityConstraints,
// SolvePositionConstraints} (src/dynamics/b2_weld_joint.cpp:62-305) with the b2Mat33 helpers it uses
// (src/common/b2_math.cpp:29-98) ------------------------------------------------------------------
__device__ __forceinline__ float3 cross3(float3 a, float3 b) {
  return make_float3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
__device__ __forceinline__ float dot3(float3 a, float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
struct Mat33 {
  float3 ex, ey, ez;
};
__device__ __forceinline__ Mat33 weld_K(float mA, float mB, float iA, float iB, float2 rA, float2 rB) {
  Mat33 K;
  K.ex.x = mA + mB + rA.y * rA.y * iA + rB.y * rB.y * iB;
  K.ey.x = -rA.y * rA.x * iA - rB.y * rB.x * iB;
  K.ez.x = -rA.y * iA - rB.y * iB;
  K.ex.y = K.ey.x;
  K.ey.y = mA + mB + rA.x * rA.x * iA + rB.x * rB.x * iB;
  K.ez.y = rA.x * iA + rB.x * iB;
  K.ex.z = K.ez.x;
  K.ey.z = K.ez.y;
  K.ez.z = iA + iB;
  return K;
}
__device__ __forceinline__ Mat33 mat33_inverse22(const Mat33& K) {
  float a = K.ex.x, b = K.ey.x, c = K.ex.y, d = K.ey.y;
  float det = a * d - b * c;
  if (det != 0.0f) det = 1.0f / det;
  Mat33 M;
  M.ex = make_float3(det * d, -det * c, 0.0f);
  M.ey = make_float3(-det * b, det * a, 0.0f);
  M.ez = make_float3(0.0f, 0.0f, 0.0f);
  return M;
}
__device__ __forceinline__ Mat33 mat33_sym_inverse33(const Mat33& K) {
  float det = dot3(K.ex, cross3(K.ey, K.ez));
  if (det != 0.0f) det = 1.0f / det;
  float a11 = K.ex.x, a12 = K.ey.x, a13 = K.ez.x, a22 = K.ey.y, a23 = K.ez.y, a33 = K.ez.z;
  Mat33 M;
  M.ex.x = det * (a22 * a33 - a23 * a23);
  M.ex.y = det * (a13 * a23 - a12 * a33);
  M.ex.z = det * (a12 * a23 - a13 * a22);
  M.ey.x = M.ex.y;
  M.ey.y = det * (a11 * a33 - a13 * a13);
  M.ey.z = det * (a13 * a12 - a11 * a23);
  M.ez.x = M.ex.z;
  M.ez.y = M.ey.z;
  M.ez.z = det * (a11 * a22 - a12 * a12);
  return M;
}
__device__ __forceinline__ float3 mat33_solve33(const Mat33& K, float3 b) {
  float det = dot3(K.ex, cross3(K.ey, K.ez));
  if (det != 0.0f) det = 1.0f / det;
  return make_float3(det * dot3(b, cross3(K.ey, K.ez)), det * dot3(K.ex, cross3(b, K.ez)), det * dot3(K.ex, cross3(K.ey, b)));
}
__device__ __forceinline__ float2 mat33_solve22(const Mat33& K, float2 b) {
  return mat22_solve(K.ex.x, K.ey.x, K.ex.y, K.ey.y, b);
}

template <class PosAccess, class VelAccess>
__device__ __forceinline__ void weld_init(const JointArraysDev& J, int j, int ia, int ib, const PosAccess& pos,
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
  w.k11 = w.k12 = w.k22 = w.axialMass = w.angle = 0.0f;
  w.u = make_float2(0.0f, 0.0f);
  w.dMass = w.softMass = w.currentLength = 0.0f;
  float4 an = J.anchors[j], p0 = J.params0[j];
  const float referenceAngle = p0.x, stiffness = p0.y, damping = p0.z;
  float4 pA = pos.load(ia), pB = pos.load(ib);
  float4 vAq = vel.load(ia), vBq = vel.load(ib);
  float aA = pA.z, aB = pB.z;
  float2 vA = make_float2(vAq.x, vAq.y), vB = make_float2(vBq.x, vBq.y);
  float wA = vAq.z, wB = vBq.z;
  Rot qA = rot_set(aA), qB = rot_set(aB);
  w.rA = rot_mul(qA, make_float2(an.x, an.y) - w.lcA);
  w.rB = rot_mul(qB, make_float2(an.z, an.w) - w.lcB);
  float mA = w.mA, mB = w.mB, iA = w.iA, iB = w.iB;
  Mat33 K = weld_K(mA, mB, iA, iB, w.rA, w.rB);
  Mat33 M;
  if (stiffness > 0.0f) {
    M = mat33_inverse22(K);
    float invM = iA + iB;
    float C = aB - aA - referenceAngle;
    float h = J.h;
    w.gamma = h * (damping + h * stiffness);
    w.gamma = w.gamma != 0.0f ? 1.0f / w.gamma : 0.0f;
    w.bias = C * h * stiffness * w.gamma;
    invM += w.gamma;
    M.ez.z = invM != 0.0f ? 1.0f / invM : 0.0f;
  } else if (K.ez.z == 0.0f) {
    M = mat33_inverse22(K);
    w.gamma = 0.0f;
    w.bias = 0.0f;
  } else {
    M = mat33_sym_inverse33(K);
    w.gamma = 0.0f;
    w.bias = 0.0f;
  }
  w.wex = M.ex;
  w.wey = M.ey;
  w.wez = M.ez;
  float4 st = J.state[j];
  if (warmStarting) {
    st.x *= dtRatio;
    st.y *= dtRatio;
    st.z *= dtRatio;
    float2 P = make_float2(st.x, st.y);
    vA -= mA * P;
    wA -= iA * (cross2(w.rA, P) + st.z);
    vB += mB * P;
    wB += iB * (cross2(w.rB, P) + st.z);
  } else {
    st.x = st.y = st.z = 0.0f;
  }
  J.state[j] = st;
  J.work[j] = w;
  if (movable(mA, iA)) vel.store(ia, make_float4(vA.x, vA.y, wA, vAq.w));
  if (movable(mB, iB)) vel.store(ib, make_float4(vB.x, vB.y, wB, vBq.w));
}

template <class VelAccess>
__device__ __forceinline__ void weld_solve_velocity(const JointArraysDev& J, int j, const VelAccess& vel) {
  JointWork w = J.work[j];
  const float stiffness = J.params0[j].y;
  float4 st = J.state[j];
  float4 vAq = vel.load(w.ia), vBq = vel.load(w.ib);
  float2 vA = make_float2(vAq.x, vAq.y), vB = make_float2(vBq.x, vBq.y);
  float wA = vAq.z, wB = vBq.z;
  float mA = w.mA, mB = w.mB, iA = w.iA, iB = w.iB;
  if (stiffness > 0.0f) {
    float Cdot2 = wB - wA;
    float impulse2 = -w.wez.z * (Cdot2 + w.bias + w.gamma * st.z);
    st.z += impulse2;
    wA -= iA * impulse2;
    wB += iB * impulse2;
    float2 Cdot1 = vB + cross_sv(wB, w.rB) - vA - cross_sv(wA, w.rA);
    // b2Mul22(m_mass, Cdot1), negated
    float2 impulse1 = -make_float2(w.wex.x * Cdot1.x + w.wey.x * Cdot1.y, w.wex.y * Cdot1.x + w.wey.y * Cdot1.y);
    st.x += impulse1.x;
    st.y += impulse1.y;
    float2 P = impulse1;
    vA -= mA * P;
    wA -= iA * cross2(w.rA, P);
    vB += mB * P;
    wB += iB * cross2(w.rB, P);
  } else {
    float2 Cdot1 = vB + cross_sv(wB, w.rB) - vA - cross_sv(wA, w.rA);
    float Cdot2 = wB - wA;
    // b2Mul(m_mass, Cdot) = Cdot.x * ex + Cdot.y * ey + Cdot.z * ez, negated
    float3 impulse = make_float3(-((Cdot1.x * w.wex.x + Cdot1.y * w.wey.x) + Cdot2 * w.wez.x),
                                 -((Cdot1.x * w.wex.y + Cdot1.y * w.wey.y) + Cdot2 * w.wez.y),
                                 -((Cdot1.x * w.wex.z + Cdot1.y * w.wey.z) + Cdot2 * w.wez.z));
    st.x += impulse.x;
    st.y += impulse.y;
    st.z += impulse.z;
    float2 P = make_float2(impulse.x, impulse.y);
    vA -= mA * P;
    wA -= iA * (cross2(w.rA, P) + impulse.z);
    vB += mB * P;
    wB += iB * (cross2(w.rB, P) + impulse.z);
  }
  J.state[j] = st;
  if (movable(mA, iA)) vel.store(w.ia, make_float4(vA.x, vA.y, wA, vAq.w));
  if (movable(mB, iB)) vel.store(w.ib, make_float4(vB.x, vB.y, wB, vBq.w));
}

template <class PosAccess>
__device__ __forceinline__ bool weld_solve_position(const JointArraysDev& J, int j, const PosAccess& pos) {
  JointWork w = J.work[j];
  float4 an = J.anchors[j], p0 = J.params0[j];
  const float referenceAngle = p0.x, stiffness = p0.y;
  float4 pAq = pos.load(w.ia), pBq = pos.load(w.ib);
  float2 cA = make_float2(pAq.x, pAq.y), cB = make_float2(pBq.x, pBq.y);
  float aA = pAq.z, aB = pBq.z;
  Rot qA = rot_set(aA), qB = rot_set(aB);
  float mA = w.mA, mB = w.mB, iA = w.iA, iB = w.iB;
  float2 rA = rot_mul(qA, make_float2(an.x, an.y) - w.lcA);
  float2 rB = rot_mul(qB, make_float2(an.z, an.w) - w.lcB);
  float positionError, angularError;
  Mat33 K = weld_K(mA, mB, iA, iB, rA, rB);
  if (stiffness > 0.0f) {
    float2 C1 = cB + rB - cA - rA;
    positionError = len2(C1);
    angularError = 0.0f;
    float2 P = -mat33_solve22(K, C1);
    cA -= mA * P;
    aA -= iA * cross2(rA, P);
    cB += mB * P;
    aB += iB * cross2(rB, P);
  } else {
    float2 C1 = cB + rB - cA - rA;
    float C2 = aB - aA - referenceAngle;
    positionError = len2(C1);
    angularError = absf_(C2);
    float3 impulse;
    if (K.ez.z > 0.0f) {
      float3 x = mat33_solve33(K, make_float3(C1.x, C1.y, C2));
      impulse = make_float3(-x.x, -x.y, -x.z);
    } else {
      float2 i2 = -mat33_solve22(K, C1);
      impulse = make_float3(i2.x, i2.y, 0.0f);
    }
    float2 P = make_float2(impulse.x, impulse.y);
    cA -= mA * P;
    aA -= iA * (cross2(rA, P) + impulse.z);
    cB += mB * P;
    aB += iB * (cross2(rB, P) + impulse.z);
  }
  if (movable(mA, iA)) pos.store(w.ia, make_float4(cA.x, cA.y, aA, pAq.w));
  if (movable(mB, iB)) pos.store(w.ib, make_float4(cB.x, cB.y, aB, pBq.w));
  return positionError <= B2G_LINEAR_SLOP && angularError <= B2G_ANGULAR_SLOP;
}

// ---- prismatic joint: b2PrismaticJoint::{InitVelocityConstraints, SolveVelocityConstraints,
// SolvePositionConstraints} (src/dynamics/b2_prismatic_joint.cpp:114-451): motor, translation limits,
// the 2-row perpendicular + angular block ----------------------------------------------------------
template <class PosAccess, class VelAccess>
__device__ __forceinline__ void prismatic_init(const JointArraysDev& J, int j, int ia, int ib, const PosAccess& pos,
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
  w.u = make_float2(0.0f, 0.0f);
  w.dMass = w.softMass = w.gamma = w.bias = w.currentLength = 0.0f;
  w.wex = w.wey = w.wez = make_float3(0.0f, 0.0f, 0.0f);
  float4 an = J.anchors[j], p1 = J.params1[j];
  const uint32_t flags = __float_as_uint(p1.y);
  const float2 localX = make_float2(p1.z, p1.w);
  const float2 localY = cross_sv(1.0f, localX);
  float4 pA = pos.load(ia), pB = pos.load(ib);
  float4 vAq = vel.load(ia), vBq = vel.load(ib);
  float2 cA = make_float2(pA.x, pA.y), cB = make_float2(pB.x, pB.y);
  float2 vA = make_float2(vAq.x, vAq.y), vB = make_float2(vBq.x, vBq.y);
  float wA = vAq.z, wB = vBq.z;
  Rot qA = rot_set(pA.z), qB = rot_set(pB.z);
  w.rA = rot_mul(qA, make_float2(an.x, an.y) - w.lcA);
  w.rB = rot_mul(qB, make_float2(an.z, an.w) - w.lcB);
  float2 d = (cB - cA) + w.rB - w.rA;
  float mA = w.mA, mB = w.mB, iA = w.iA, iB = w.iB;
  {
    w.axis = rot_mul(qA, localX);
    w.a1 = cross2(d + w.rA, w.axis);
    w.a2 = cross2(w.rB, w.axis);
    w.axialMass = mA + mB + iA * w.a1 * w.a1 + iB * w.a2 * w.a2;
    if (w.axialMass > 0.0f) w.axialMass = 1.0f / w.axialMass;
  }
  {
    w.perp = rot_mul(qA, localY);
    w.s1 = cross2(d + w.rA, w.perp);
    w.s2 = cross2(w.rB, w.perp);
    w.k11 = mA + mB + iA * w.s1 * w.s1 + iB * w.s2 * w.s2;
    w.k12 = iA * w.s1 + iB * w.s2;
    w.k22 = iA + iB;
    if (w.k22 == 0.0f) w.k22 = 1.0f;  // bodies with fixed rotation
  }
  float4 st = J.state[j];
  float upper = J.upper[j];
  w.angle = 0.0f;  // translation
  if (flags & B2G_JOINT_LIMIT) {
    w.angle = dot2(w.axis, d);
  } else {
    st.w = 0.0f;
    upper = 0.0f;
  }
  if (!(flags & B2G_JOINT_MOTOR)) st.z = 0.0f;
  if (warmStarting) {
    st.x *= dtRatio;
    st.y *= dtRatio;
    st.z *= dtRatio;
    st.w *= dtRatio;
    upper *= dtRatio;
    float axialImpulse = st.z + st.w - upper;
    float2 P = st.x * w.perp + axialImpulse * w.axis;
    float LA = st.x * w.s1 + st.y + axialImpulse * w.a1;
    float LB = st.x * w.s2 + st.y + axialImpulse * w.a2;
    vA -= mA * P;
    wA -= iA * LA;
    vB += mB * P;
    wB += iB * LB;
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
__device__ __forceinline__ void prismatic_solve_velocity(const JointArraysDev& J, int j, const VelAccess& vel, float dt,
                                                         float inv_dt) {
  JointWork w = J.work[j];
  float4 p0 = J.params0[j], p1 = J.params1[j];
  const uint32_t flags = __float_as_uint(p1.y);
  const float lowerTranslation = p0.y, upperTranslation = p0.z, maxMotorForce = p0.w, motorSpeed = p1.x;
  float4 st = J.state[j];
  float upper = J.upper[j];
  float4 vAq = vel.load(w.ia), vBq = vel.load(w.ib);
  float2 vA = make_float2(vAq.x, vAq.y), vB = make_float2(vBq.x, vBq.y);
  float wA = vAq.z, wB = vBq.z;
  float mA = w.mA, mB = w.mB, iA = w.iA, iB = w.iB;
  if (flags & B2G_JOINT_MOTOR) {
    float Cdot = dot2(w.axis, vB - vA) + w.a2 * wB - w.a1 * wA;
    float impulse = w.axialMass * (motorSpeed - Cdot);
    float oldImpulse = st.z;
    float maxImpulse = dt * maxMotorForce;
    st.z = clampf(st.z + impulse, -maxImpulse, maxImpulse);
    impulse = st.z - oldImpulse;
    float2 P = impulse * w.axis;
    float LA = impulse * w.a1, LB = impulse * w.a2;
    vA -= mA * P;
    wA -= iA * LA;
    vB += mB * P;
    wB += iB * LB;
  }
  if (flags & B2G_JOINT_LIMIT) {
    {  // lower
      float C = w.angle - lowerTranslation;
      float Cdot = dot2(w.axis, vB - vA) + w.a2 * wB - w.a1 * wA;
      float impulse = -w.axialMass * (Cdot + maxf_(C, 0.0f) * inv_dt);
      float oldImpulse = st.w;
      st.w = maxf_(st.w + impulse, 0.0f);
      impulse = st.w - oldImpulse;
      float2 P = impulse * w.axis;
      float LA = impulse * w.a1, LB = impulse * w.a2;
      vA -= mA * P;
      wA -= iA * LA;
      vB += mB * P;
      wB += iB * LB;
    }
    {  // upper (signs flipped so that C and the impulse stay positive)
      float C = upperTranslation - w.angle;
      float Cdot = dot2(w.axis, vA - vB) + w.a1 * wA - w.a2 * wB;
      float impulse = -w.axialMass * (Cdot + maxf_(C, 0.0f) * inv_dt);
      float oldImpulse = upper;
      upper = maxf_(upper + impulse, 0.0f);
      impulse = upper - oldImpulse;
      float2 P = impulse * w.axis;
      float LA = impulse * w.a1, LB = impulse * w.a2;
      vA += mA * P;
      wA += iA * LA;
      vB -= mB * P;
      wB -= iB * LB;
    }
  }
  {  // the prismatic constraint in block form
    float2 Cdot;
    Cdot.x = dot2(w.perp, vB - vA) + w.s2 * wB - w.s1 * wA;
    Cdot.y = wB - wA;
    float2 df = mat22_solve(w.k11, w.k12, w.k12, w.k22, -Cdot);
    st.x += df.x;
    st.y += df.y;
    float2 P = df.x * w.perp;
    float LA = df.x * w.s1 + df.y;
    float LB = df.x * w.s2 + df.y;
    vA -= mA * P;
    wA -= iA * LA;
    vB += mB * P;
    wB += iB * LB;
  }
  J.state[j] = st;
  J.upper[j] = upper;
  if (movable(mA, iA)) vel.store(w.ia, make_float4(vA.x, vA.y, wA, vAq.w));
  if (movable(mB, iB)) vel.store(w.ib, make_float4(vB.x, vB.y, wB, vBq.w));
}

template <class PosAccess>
__device__ __forceinline__ bool prismatic_solve_position(const JointArraysDev& J, int j, const PosAccess& pos) {
  JointWork w = J.work[j];
  float4 an = J.anchors[j], p0 = J.params0[j], p1 = J.params1[j];
  const uint32_t flags = __float_as_uint(p1.y);
  const float referenceAngle = p0.x, lowerTranslation = p0.y, upperTranslation = p0.z;
  const float2 localX = make_float2(p1.z, p1.w);
  const float2 localY = cross_sv(1.0f, localX);
  float4 pAq = pos.load(w.ia), pBq = pos.load(w.ib);
  float2 cA = make_float2(pAq.x, pAq.y), cB = make_float2(pBq.x, pBq.y);
  float aA = pAq.z, aB = pBq.z;
  Rot qA = rot_set(aA), qB = rot_set(aB);
  float mA = w.mA, mB = w.mB, iA = w.iA, iB = w.iB;
  float2 rA = rot_mul(qA, make_float2(an.x, an.y) - w.lcA);
  float2 rB = rot_mul(qB, make_float2(an.z, an.w) - w.lcB);
  float2 d = cB + rB - cA - rA;
  float2 axis = rot_mul(qA, localX);
  float a1 = cross2(d + rA, axis);
  float a2 = cross2(rB, axis);
  float2 perp = rot_mul(qA, localY);
  float s1 = cross2(d + rA, perp);
  float s2 = cross2(rB, perp);
  float3 impulse;
  float2 C1 = make_float2(dot2(perp, d), aB - aA - referenceAngle);
  float linearError = absf_(C1.x);
  float angularError = absf_(C1.y);
  bool active = false;
  float C2 = 0.0f;
  if (flags & B2G_JOINT_LIMIT) {
    float translation = dot2(axis, d);
    if (absf_(upperTranslation - lowerTranslation) < 2.0f * B2G_LINEAR_SLOP) {
      C2 = translation;
      linearError = maxf_(linearError, absf_(translation));
      active = true;
    } else if (translation <= lowerTranslation) {
      C2 = minf_(translation - lowerTranslation, 0.0f);
      linearError = maxf_(linearError, lowerTranslation - translation);
      active = true;
    } else if (translation >= upperTranslation) {
      C2 = maxf_(translation - upperTranslation, 0.0f);
      linearError = maxf_(linearError, translation - upperTranslation);
      active = true;
    }
  }
  if (active) {
    float k11 = mA + mB + iA * s1 * s1 + iB * s2 * s2;
    float k12 = iA * s1 + iB * s2;
    float k13 = iA * s1 * a1 + iB * s2 * a2;
    float k22 = iA + iB;
    if (k22 == 0.0f) k22 = 1.0f;
    float k23 = iA * a1 + iB * a2;
    float k33 = mA + mB + iA * a1 * a1 + iB * a2 * a2;
    Mat33 K;
    K.ex = make_float3(k11, k12, k13);
    K.ey = make_float3(k12, k22, k23);
    K.ez = make_float3(k13, k23, k33);
    impulse = mat33_solve33(K, make_float3(-C1.x, -C1.y, -C2));
  } else {
    float k11 = mA + mB + iA * s1 * s1 + iB * s2 * s2;
    float k12 = iA * s1 + iB * s2;
    float k22 = iA + iB;
    if (k22 == 0.0f) k22 = 1.0f;
    float2 impulse1 = mat22_solve(k11, k12, k12, k22, -C1);
    impulse = make_float3(impulse1.x, impulse1.y, 0.0f);
  }
  float2 P = impulse.x * perp + impulse.z * axis;
  float LA = impulse.x * s1 + impulse.y + impulse.z * a1;
  float LB = impulse.x * s2 + impulse.y + impulse.z * a2;
  cA -= mA * P;
  aA -= iA * LA;
  cB += mB * P;
  aB += iB * LB;
  if (movable(mA, iA)) pos.store(w.ia, make_float4(cA.x, cA.y, aA, pAq.w));
  if (movable(mB, iB)) pos.store(w.ib, make_float4(cB.x, cB.y, aB, pBq.w));
  return linearError <= B2G_LINEAR_SLOP && angularError <= B2G_ANGULAR_SLOP;
}

// ---- wheel joint: b2WheelJoint::{InitVelocityConstraints, SolveVelocityConstraints,
// SolvePositionConstraints} (src/dynamics/b2_wheel_joint.cpp:87-446): point-to-line row, suspension
// spring along the axis, rotational motor, translation limits -------------------------------------------
template <class PosAccess, class VelAccess>
__device__ __forceinline__ void wheel_init(const JointArraysDev& J, int j, int ia, int ib, const PosAccess& pos,
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
  w.u = make_float2(0.0f, 0.0f);
  w.k12 = w.k22 = w.currentLength = 0.0f;
  w.wex = w.wey = w.wez = make_float3(0.0f, 0.0f, 0.0f);
  float4 an = J.anchors[j], p0 = J.params0[j], p1 = J.params1[j], p2 = J.params2[j];
  const uint32_t flags = __float_as_uint(p1.y);
  const float stiffness = p0.x, damping = p2.x;
  const float2 localX = make_float2(p1.z, p1.w);
  const float2 localY = cross_sv(1.0f, localX);
  float mA = w.mA, mB = w.mB, iA = w.iA, iB = w.iB;
  float4 pA = pos.load(ia), pB = pos.load(ib);
  float4 vAq = vel.load(ia), vBq = vel.load(ib);
  float2 cA = make_float2(pA.x, pA.y), cB = make_float2(pB.x, pB.y);
  float2 vA = make_float2(vAq.x, vAq.y), vB = make_float2(vBq.x, vBq.y);
  float wA = vAq.z, wB = vBq.z;
  Rot qA = rot_set(pA.z), qB = rot_set(pB.z);
  w.rA = rot_mul(qA, make_float2(an.x, an.y) - w.lcA);
  w.rB = rot_mul(qB, make_float2(an.z, an.w) - w.lcB);
  float2 d = cB + w.rB - cA - w.rA;
  {  // point to line constraint
    w.perp = rot_mul(qA, localY);
    w.s1 = cross2(d + w.rA, w.perp);
    w.s2 = cross2(w.rB, w.perp);
    w.k11 = mA + mB + iA * w.s1 * w.s1 + iB * w.s2 * w.s2;
    if (w.k11 > 0.0f) w.k11 = 1.0f / w.k11;
  }
  // spring constraint
  w.axis = rot_mul(qA, localX);
  w.a1 = cross2(d + w.rA, w.axis);
  w.a2 = cross2(w.rB, w.axis);
  const float invMass = mA + mB + iA * w.a1 * w.a1 + iB * w.a2 * w.a2;
  w.axialMass = invMass > 0.0f ? 1.0f / invMass : 0.0f;
  w.softMass = 0.0f;
  w.bias = 0.0f;
  w.gamma = 0.0f;
  float4 st = J.state[j];  // impulse, springImpulse, motorImpulse, lowerImpulse
  float upper = J.upper[j];
  if (stiffness > 0.0f && invMass > 0.0f) {
    w.softMass = 1.0f / invMass;
    float C = dot2(d, w.axis);
    float h = J.h;
    w.gamma = h * (damping + h * stiffness);
    if (w.gamma > 0.0f) w.gamma = 1.0f / w.gamma;
    w.bias = C * h * stiffness * w.gamma;
    w.softMass = invMass + w.gamma;
    if (w.softMass > 0.0f) w.softMass = 1.0f / w.softMass;
  } else {
    st.y = 0.0f;
  }
  w.angle = 0.0f;  // translation
  if (flags & B2G_JOINT_LIMIT) {
    w.angle = dot2(w.axis, d);
  } else {
    st.w = 0.0f;
    upper = 0.0f;
  }
  if (flags & B2G_JOINT_MOTOR) {
    w.dMass = iA + iB;
    if (w.dMass > 0.0f) w.dMass = 1.0f / w.dMass;
  } else {
    w.dMass = 0.0f;
    st.z = 0.0f;
  }
  if (warmStarting) {
    // the limit impulses are not scaled by dtRatio (b2_wheel_joint.cpp:202-205)
    st.x *= dtRatio;
    st.y *= dtRatio;
    st.z *= dtRatio;
    float axialImpulse = st.y + st.w - upper;
    float2 P = st.x * w.perp + axialImpulse * w.axis;
    float LA = st.x * w.s1 + axialImpulse * w.a1 + st.z;
    float LB = st.x * w.s2 + axialImpulse * w.a2 + st.z;
    vA -= mA * P;
    wA -= iA * LA;
    vB += mB * P;
    wB += iB * LB;
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
__device__ __forceinline__ void wheel_solve_velocity(const JointArraysDev& J, int j, const VelAccess& vel, float dt,
                                                     float inv_dt) {
  JointWork w = J.work[j];
  float4 p0 = J.params0[j], p1 = J.params1[j];
  const uint32_t flags = __float_as_uint(p1.y);
  const float lowerTranslation = p0.y, upperTranslation = p0.z, maxMotorTorque = p0.w, motorSpeed = p1.x;
  float4 st = J.state[j];
  float upper = J.upper[j];
  float4 vAq = vel.load(w.ia), vBq = vel.load(w.ib);
  float2 vA = make_float2(vAq.x, vAq.y), vB = make_float2(vBq.x, vBq.y);
  float wA = vAq.z, wB = vBq.z;
  float mA = w.mA, mB = w.mB, iA = w.iA, iB = w.iB;
  {  // spring
    float Cdot = dot2(w.axis, vB - vA) + w.a2 * wB - w.a1 * wA;
    float impulse = -w.softMass * (Cdot + w.bias + w.gamma * st.y);
    st.y += impulse;
    float2 P = impulse * w.axis;
    float LA = impulse * w.a1, LB = impulse * w.a2;
    vA -= mA * P;
    wA -= iA * LA;
    vB += mB * P;
    wB += iB * LB;
  }
  {  // rotational motor (runs with motorMass = 0 when the motor is off)
    float Cdot = wB - wA - motorSpeed;
    float impulse = -w.dMass * Cdot;
    float oldImpulse = st.z;
    float maxImpulse = dt * maxMotorTorque;
    st.z = clampf(st.z + impulse, -maxImpulse, maxImpulse);
    impulse = st.z - oldImpulse;
    wA -= iA * impulse;
    wB += iB * impulse;
  }
  if (flags & B2G_JOINT_LIMIT) {
    {  // lower
      float C = w.angle - lowerTranslation;
      float Cdot = dot2(w.axis, vB - vA) + w.a2 * wB - w.a1 * wA;
      float impulse = -w.axialMass * (Cdot + maxf_(C, 0.0f) * inv_dt);
      float oldImpulse = st.w;
      st.w = maxf_(st.w + impulse, 0.0f);
      impulse = st.w - oldImpulse;
      float2 P = impulse * w.axis;
      float LA = impulse * w.a1, LB = impulse * w.a2;
      vA -= mA * P;
      wA -= iA * LA;
      vB += mB * P;
      wB += iB * LB;
    }
    {  // upper (signs flipped)
      float C = upperTranslation - w.angle;
      float Cdot = dot2(w.axis, vA - vB) + w.a1 * wA - w.a2 * wB;
      float impulse = -w.axialMass * (Cdot + maxf_(C, 0.0f) * inv_dt);
      float oldImpulse = upper;
      upper = maxf_(upper + impulse, 0.0f);
      impulse = upper - oldImpulse;
      float2 P = impulse * w.axis;
      float LA = impulse * w.a1, LB = impulse * w.a2;
      vA += mA * P;
      wA += iA * LA;
      vB -= mB * P;
      wB -= iB * LB;
    }
  }
  {  // point to line
    float Cdot = dot2(w.perp, vB - vA) + w.s2 * wB - w.s1 * wA;
    float impulse = -w.k11 * Cdot;
    st.x += impulse;
    float2 P = impulse * w.perp;
    float LA = impulse * w.s1, LB = impulse * w.s2;
    vA -= mA * P;
    wA -= iA * LA;
    vB += mB * P;
    wB += iB * LB;
  }
  J.state[j] = st;
  J.upper[j] = upper;
  if (movable(mA, iA)) vel.store(w.ia, make_float4(vA.x, vA.y, wA, vAq.w));
  if (movable(mB, iB)) vel.store(w.ib, make_float4(vB.x, vB.y, wB, vBq.w));
}

template <class PosAccess>
__device__ __forceinline__ bool wheel_solve_position(const JointArraysDev& J, int j, const PosAccess& pos) {
  JointWork w = J.work[j];
  float4 an = J.anchors[j], p0 = J.params0[j], p1 = J.params1[j];
  const uint32_t flags = __float_as_uint(p1.y);
  const float lowerTranslation = p0.y, upperTranslation = p0.z;
  const float2 localX = make_float2(p1.z, p1.w);
  const float2 localY = cross_sv(1.0f, localX);
  float4 pAq = pos.load(w.ia), pBq = pos.load(w.ib);
  float2 cA = make_float2(pAq.x, pAq.y), cB = make_float2(pBq.x, pBq.y);
  float aA = pAq.z, aB = pBq.z;
  float mA = w.mA, mB = w.mB, iA = w.iA, iB = w.iB;
  float linearError = 0.0f;
  if (flags & B2G_JOINT_LIMIT) {
    Rot qA = rot_set(aA), qB = rot_set(aB);
    float2 rA = rot_mul(qA, make_float2(an.x, an.y) - w.lcA);
    float2 rB = rot_mul(qB, make_float2(an.z, an.w) - w.lcB);
    float2 d = (cB - cA) + rB - rA;
    float2 ax = rot_mul(qA, localX);
    // the reference takes the lever arms about the axis of InitVelocityConstraints here (m_ax), :363-364
    float sAx = cross2(d + rA, w.axis);
    float sBx = cross2(rB, w.axis);
    float C = 0.0f;
    float translation = dot2(ax, d);
    if (absf_(upperTranslation - lowerTranslation) < 2.0f * B2G_LINEAR_SLOP) C = translation;
    else if (translation <= lowerTranslation) C = minf_(translation - lowerTranslation, 0.0f);
    else if (translation >= upperTranslation) C = maxf_(translation - upperTranslation, 0.0f);
    if (C != 0.0f) {
      float invMass = mA + mB + iA * sAx * sAx + iB * sBx * sBx;
      float impulse = 0.0f;
      if (invMass != 0.0f) impulse = -C / invMass;
      float2 P = impulse * ax;
      float LA = impulse * sAx, LB = impulse * sBx;
      cA -= mA * P;
      aA -= iA * LA;
      cB += mB * P;
      aB += iB * LB;
      linearError = absf_(C);
    }
  }
  {  // perpendicular constraint
    Rot qA = rot_set(aA), qB = rot_set(aB);
    float2 rA = rot_mul(qA, make_float2(an.x, an.y) - w.lcA);
    float2 rB = rot_mul(qB, make_float2(an.z, an.w) - w.lcB);
    float2 d = (cB - cA) + rB - rA;
    float2 ay = rot_mul(qA, localY);
    float sAy = cross2(d + rA, ay);
    float sBy = cross2(rB, ay);
    float C = dot2(d, ay);
    // the effective mass uses the lever arms of InitVelocityConstraints (m_sAy, m_sBy), :416
    float invMass = mA + mB + iA * w.s1 * w.s1 + iB * w.s2 * w.s2;
    float impulse = 0.0f;
    if (invMass != 0.0f) impulse = -C / invMass;
    float2 P = impulse * ay;
    float LA = impulse * sAy, LB = impulse * sBy;
    cA -= mA * P;
    aA -= iA * LA;
    cB += mB * P;
    aB += iB * LB;
    linearError = maxf_(linearError, absf_(C));
  }
  if (movable(mA, iA)) pos.store(w.ia, make_float4(cA.x, cA.y, aA, pAq.w));
  if (movable(mB, iB)) pos.store(w.ib, make_float4(cB.x, cB.y, aB, pBq.w));
  return linearError <= B2G_LINEAR_SLOP;
}

// ---- friction joint and motor joint: b2FrictionJoint / b2MotorJoint::{InitVelocityConstraints,
// SolveVelocityConstraints} (src/dynamics/b2_friction_joint.cpp:65-181, b2_motor_joint.cpp:70-208): a clamped
// angular row and a clamped 2-D linear row; the motor joint adds a position-error feed (correctionFactor) and
// takes its lever arms from the linear offset.  Neither has a position solve (both return true). ------------
template <bool MOTOR, class PosAccess, class VelAccess>
__device__ __forceinline__ void drag_init(const JointArraysDev& J, int j, int ia, int ib, const PosAccess& pos,
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
  w.dMass = w.softMass = w.gamma = w.bias = w.currentLength = 0.0f;
  w.wex = w.wey = w.wez = make_float3(0.0f, 0.0f, 0.0f);
  w.axis = w.perp = make_float2(0.0f, 0.0f);
  w.a1 = w.a2 = w.s1 = w.s2 = 0.0f;
  float4 an = J.anchors[j], p0 = J.params0[j];
  float4 pA = pos.load(ia), pB = pos.load(ib);
  float4 vAq = vel.load(ia), vBq = vel.load(ib);
  float2 vA = make_float2(vAq.x, vAq.y), vB = make_float2(vBq.x, vBq.y);
  float wA = vAq.z, wB = vBq.z;
  Rot qA = rot_set(pA.z), qB = rot_set(pB.z);
  if (MOTOR) {
    w.rA = rot_mul(qA, make_float2(an.x, an.y) - w.lcA);
    w.rB = rot_mul(qB, -w.lcB);
  } else {
    w.rA = rot_mul(qA, make_float2(an.x, an.y) - w.lcA);
    w.rB = rot_mul(qB, make_float2(an.z, an.w) - w.lcB);
  }
  float mA = w.mA, mB = w.mB, iA = w.iA, iB = w.iB;
  {  // m_linearMass = K.GetInverse() (b2_math.h:204-216)
    float a = mA + mB + iA * w.rA.y * w.rA.y + iB * w.rB.y * w.rB.y;
    float b = -iA * w.rA.x * w.rA.y - iB * w.rB.x * w.rB.y;
    float d = mA + mB + iA * w.rA.x * w.rA.x + iB * w.rB.x * w.rB.x;
    float det = a * d - b * b;
    if (det != 0.0f) det = 1.0f / det;
    w.k11 = det * d;
    w.k12 = -det * b;
    w.k22 = det * a;
  }
  w.axialMass = iA + iB;
  if (w.axialMass > 0.0f) w.axialMass = 1.0f / w.axialMass;
  w.u = make_float2(0.0f, 0.0f);
  w.angle = 0.0f;
  if (MOTOR) {
    float2 cA = make_float2(pA.x, pA.y), cB = make_float2(pB.x, pB.y);
    w.u = cB + w.rB - cA - w.rA;
    w.angle = pB.z - pA.z - p0.w;
  }
  float4 st = J.state[j];
  if (warmStarting) {
    st.x *= dtRatio;
    st.y *= dtRatio;
    st.z *= dtRatio;
    float2 P = make_float2(st.x, st.y);
    vA -= mA * P;
    wA -= iA * (cross2(w.rA, P) + st.z);
    vB += mB * P;
    wB += iB * (cross2(w.rB, P) + st.z);
  } else {
    st = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
  }
  J.state[j] = st;
  J.work[j] = w;
  if (movable(mA, iA)) vel.store(ia, make_float4(vA.x, vA.y, wA, vAq.w));
  if (movable(mB, iB)) vel.store(ib, make_float4(vB.x, vB.y, wB, vBq.w));
}

template <bool MOTOR, class VelAccess>
__device__ __forceinline__ void drag_solve_velocity(const JointArraysDev& J, int j, const VelAccess& vel, float h,
                                                    float inv_h) {
  JointWork w = J.work[j];
  float4 p0 = J.params0[j];
  const float maxForce = p0.x, maxTorque = p0.y, correctionFactor = p0.z;
  float4 st = J.state[j];
  float4 vAq = vel.load(w.ia), vBq = vel.load(w.ib);
  float2 vA = make_float2(vAq.x, vAq.y), vB = make_float2(vBq.x, vBq.y);
  float wA = vAq.z, wB = vBq.z;
  float mA = w.mA, mB = w.mB, iA = w.iA, iB = w.iB;
  {  // angular row
    float Cdot = wB - wA;
    if (MOTOR) Cdot = Cdot + inv_h * correctionFactor * w.angle;
    float impulse = -w.axialMass * Cdot;
    float oldImpulse = st.z;
    float maxImpulse = h * maxTorque;
    st.z = clampf(st.z + impulse, -maxImpulse, maxImpulse);
    impulse = st.z - oldImpulse;
    wA -= iA * impulse;
    wB += iB * impulse;
  }
  {  // linear rows
    float2 Cdot = vB + cross_sv(wB, w.rB) - vA - cross_sv(wA, w.rA);
    if (MOTOR) Cdot = Cdot + (inv_h * correctionFactor) * w.u;
    float2 impulse = -make_float2(w.k11 * Cdot.x + w.k12 * Cdot.y, w.k12 * Cdot.x + w.k22 * Cdot.y);
    float2 oldImpulse = make_float2(st.x, st.y);
    float2 acc = oldImpulse + impulse;
    float maxImpulse = h * maxForce;
    if (dot2(acc, acc) > maxImpulse * maxImpulse) {
      normalize2(acc);
      acc = maxImpulse * acc;   // b2Vec2::operator*=(float): component * scalar
    }
    st.x = acc.x;
    st.y = acc.y;
    impulse = acc - oldImpulse;
    vA -= mA * impulse;
    wA -= iA * cross2(w.rA, impulse);
    vB += mB * impulse;
    wB += iB * cross2(w.rB, impulse);
  }
  J.state[j] = st;
  if (movable(mA, iA)) vel.store(w.ia, make_float4(vA.x, vA.y, wA, vAq.w));
  if (movable(mB, iB)) vel.store(w.ib, make_float4(vB.x, vB.y, wB, vBq.w));
}

// ---- mouse joint: b2MouseJoint::{InitVelocityConstraints, SolveVelocityConstraints}
// (src/dynamics/b2_mouse_joint.cpp:77-160): a soft point constraint dragging bodyB's anchor to a world
// target with bounded force; bodyA takes no part in the arithmetic -------------------------------------------
template <class PosAccess, class VelAccess>
__device__ __forceinline__ void mouse_init(const JointArraysDev& J, int j, int ia, int ib, const PosAccess& pos,
                                           const VelAccess& vel, const float4* __restrict__ bodyMass,
                                           const float4* __restrict__ bodyCenter, float dtRatio, bool warmStarting) {
  JointWork w;
  int2 bd = J.bodies[j];
  float4 mBq = bodyMass[bd.y];
  float4 cBq = bodyCenter[bd.y];
  w.ia = ia;
  w.ib = ib;
  w.mA = 0.0f; w.iA = 0.0f; w.mB = mBq.x; w.iB = mBq.y;
  w.lcA = make_float2(0.0f, 0.0f);
  w.lcB = make_float2(cBq.x, cBq.y);
  w.rA = make_float2(0.0f, 0.0f);
  w.axialMass = w.angle = 0.0f;
  w.dMass = w.softMass = w.currentLength = 0.0f;
  w.wex = w.wey = w.wez = make_float3(0.0f, 0.0f, 0.0f);
  w.axis = w.perp = make_float2(0.0f, 0.0f);
  w.a1 = w.a2 = w.s1 = w.s2 = 0.0f;
  float4 an = J.anchors[j], p0 = J.params0[j];
  const float k = p0.y, d = p0.z;
  float4 pB = pos.load(ib);
  float4 vBq = vel.load(ib);
  float2 cB = make_float2(pB.x, pB.y);
  float2 vB = make_float2(vBq.x, vBq.y);
  float wB = vBq.z;
  Rot qB = rot_set(pB.z);
  float h = J.h;
  w.gamma = h * (d + h * k);
  if (w.gamma != 0.0f) w.gamma = 1.0f / w.gamma;
  w.bias = h * k * w.gamma;  // m_beta
  w.rB = rot_mul(qB, make_float2(an.z, an.w) - w.lcB);
  {  // m_mass = K.GetInverse()
    float a = w.mB + w.iB * w.rB.y * w.rB.y + w.gamma;
    float b = -w.iB * w.rB.x * w.rB.y;
    float dd = w.mB + w.iB * w.rB.x * w.rB.x + w.gamma;
    float det = a * dd - b * b;
    if (det != 0.0f) det = 1.0f / det;
    w.k11 = det * dd;
    w.k12 = -det * b;
    w.k22 = det * a;
  }
  w.u = cB + w.rB - make_float2(an.x, an.y);  // m_C
  w.u.x *= w.bias;
  w.u.y *= w.bias;
  wB *= 0.98f;  // cheat with some damping
  float4 st = J.state[j];
  if (warmStarting) {
    st.x *= dtRatio;
    st.y *= dtRatio;
    float2 P = make_float2(st.x, st.y);
    vB += w.mB * P;
    wB += w.iB * cross2(w.rB, P);
  } else {
    st = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
  }
  J.state[j] = st;
  J.work[j] = w;
  // the reference writes bodyB's velocity back unconditionally; a body without mass keeps it too
  vel.store(ib, make_float4(vB.x, vB.y, wB, vBq.w));
}

template <class VelAccess>
__device__ __forceinline__ void mouse_solve_velocity(const JointArraysDev& J, int j, const VelAccess& vel, float h) {
  JointWork w = J.work[j];
  const float maxForce = J.params0[j].x;
  float4 st = J.state[j];
  float4 vBq = vel.load(w.ib);
  float2 vB = make_float2(vBq.x, vBq.y);
  float wB = vBq.z;
  float2 Cdot = vB + cross_sv(wB, w.rB);
  float2 imp0 = make_float2(st.x, st.y);
  float2 t = -(Cdot + w.u + w.gamma * imp0);
  float2 impulse = make_float2(w.k11 * t.x + w.k12 * t.y, w.k12 * t.x + w.k22 * t.y);
  float2 acc = imp0 + impulse;
  float maxImpulse = h * maxForce;
  if (dot2(acc, acc) > maxImpulse * maxImpulse) {
    float s = maxImpulse / len2(acc);
    acc.x *= s;
    acc.y *= s;
  }
  st.x = acc.x;
  st.y = acc.y;
  impulse = acc - imp0;
  vB += w.mB * impulse;
  wB += w.iB * cross2(w.rB, impulse);
  J.state[j] = st;
  vel.store(w.ib, make_float4(vB.x, vB.y, wB, vBq.w));
}

// ---- dispatch on the joint type -------------------------------------------------------------------
template <class PosAccess, class VelAccess>
__device__ __forceinline__ void joint_init(const JointArraysDev& J, int j, int ia, int ib, const PosAccess& pos,
                                           const VelAccess& vel, const float4* __restrict__ bodyMass,
                                           const float4* __restrict__ bodyCenter, float dtRatio, bool warmStarting) {
  const uint32_t type = B2G_JOINT_TYPE(__float_as_uint(J.params1[j].y));
  if (type == B2G_JOINT_DISTANCE) distance_init(J, j, ia, ib, pos, vel, bodyMass, bodyCenter, dtRatio, warmStarting);
  else if (type == B2G_JOINT_WELD) weld_init(J, j, ia, ib, pos, vel, bodyMass, bodyCenter, dtRatio, warmStarting);
  else if (type == B2G_JOINT_PRISMATIC) prismatic_init(J, j, ia, ib, pos, vel, bodyMass, bodyCenter, dtRatio, warmStarting);
  else if (type == B2G_JOINT_WHEEL) wheel_init(J, j, ia, ib, pos, vel, bodyMass, bodyCenter, dtRatio, warmStarting);
  else if (type == B2G_JOINT_MOUSE) mouse_init(J, j, ia, ib, pos, vel, bodyMass, bodyCenter, dtRatio, warmStarting);
  else if (type == B2G_JOINT_FRICTION) drag_init<false>(J, j, ia, ib, pos, vel, bodyMass, bodyCenter, dtRatio, warmStarting);
  else if (type == B2G_JOINT_MOTOR_JOINT) drag_init<true>(J, j, ia, ib, pos, vel, bodyMass, bodyCenter, dtRatio, warmStarting);
  else revolute_init(J, j, ia, ib, pos, vel, bodyMass, bodyCenter, dtRatio, warmStarting);
}
template <class VelAccess>
__device__ __forceinline__ void joint_solve_velocity(const JointArraysDev& J, int j, const VelAccess& vel, float dt,
                                                     float inv_dt) {
  const uint32_t type = B2G_JOINT_TYPE(__float_as_uint(J.params1[j].y));
  if (type == B2G_JOINT_DISTANCE) distance_solve_velocity(J, j, vel, inv_dt);
  else if (type == B2G_JOINT_WELD) weld_solve_velocity(J, j, vel);
  else if (type == B2G_JOINT_PRISMATIC) prismatic_solve_velocity(J, j, vel, dt, inv_dt);
  else if (type == B2G_JOINT_WHEEL) wheel_solve_velocity(J, j, vel, dt, inv_dt);
  else if (type == B2G_JOINT_MOUSE) mouse_solve_velocity(J, j, vel, dt);
  else if (type == B2G_JOINT_FRICTION) drag_solve_velocity<false>(J, j, vel, dt, inv_dt);
  else if (type == B2G_JOINT_MOTOR_JOINT) drag_solve_velocity<true>(J, j, vel, dt, inv_dt);
  else revolute_solve_velocity(J, j, vel, dt, inv_dt);
}
template <class PosAccess>
__device__ __forceinline__ bool joint_solve_position(const JointArraysDev& J, int j, const PosAccess& pos) {
  const uint32_t type = B2G_JOINT_TYPE(__float_as_uint(J.params1[j].y));
  if (type == B2G_JOINT_DISTANCE) return distance_solve_position(J, j, pos);
  if (type == B2G_JOINT_WELD) return weld_solve_position(J, j, pos);
  if (type == B2G_JOINT_PRISMATIC) return prismatic_solve_position(J, j, pos);
  if (type == B2G_JOINT_WHEEL) return wheel_solve_position(J, j, pos);
  if (type == B2G_JOINT_FRICTION || type == B2G_JOINT_MOTOR_JOINT || type == B2G_JOINT_MOUSE) return true;
  return revolute_solve_position(J, j, pos);
}
#endif

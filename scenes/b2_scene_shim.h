// b2_scene_shim.h — extern "C" access to the scenes of b2_scenes.h, for ctypes.
//
// Included by exactly two translation units, each defining SHIM(name) first:
//   oracle/ref_harness.cpp                     SHIM(x) = b2ref_##x  (the compiled reference)
//   box2d_optimized_b200/host/gpu_scene_shim.cpp  SHIM(x) = b2gpu_##x  (the drop-in API over CUDA)
// Only public Box2D API is used here, so both expansions are the same program.
#ifndef B2_SCENE_SHIM_H
#define B2_SCENE_SHIM_H

#include "b2_scenes.h"

extern "C" {

void* SHIM(scene_create)(const char* name, int size, int seed) { return scene_build(name, size, seed); }
void SHIM(scene_destroy)(void* h) { delete static_cast<Scene*>(h); }
void SHIM(scene_step)(void* h, int n) {
  Scene* s = static_cast<Scene*>(h);
  for (int i = 0; i < n; ++i) s->step();
}
// wall-clock milliseconds for n steps (CPU baseline timing, std::chrono::steady_clock)
double SHIM(scene_time_steps)(void* h, int n) {
  Scene* s = static_cast<Scene*>(h);
  auto t0 = std::chrono::steady_clock::now();
  for (int i = 0; i < n; ++i) s->step();
  auto t1 = std::chrono::steady_clock::now();
  return std::chrono::duration<double, std::milli>(t1 - t0).count();
}
void SHIM(scene_set_iterations)(void* h, int vi, int pi) {
  Scene* s = static_cast<Scene*>(h);
  s->velocityIterations = vi;
  s->positionIterations = pi;
}
void SHIM(scene_set_flags)(void* h, int allowSleep, int warmStarting) {
  Scene* s = static_cast<Scene*>(h);
  s->world->SetAllowSleeping(allowSleep != 0);
  s->world->SetWarmStarting(warmStarting != 0);
}
int SHIM(scene_body_count)(void* h) { return (int)static_cast<Scene*>(h)->bodies.size(); }
int SHIM(scene_fixture_count)(void* h) { return (int)static_cast<Scene*>(h)->fixtures.size(); }
int SHIM(scene_contact_count)(void* h) { return static_cast<Scene*>(h)->world->GetContactCount(); }

// out[n][12] = xf.p.x, xf.p.y, xf.q.s, xf.q.c, c.x, c.y, a, v.x, v.y, w, awake, type
void SHIM(scene_get_bodies)(void* h, float* out) {
  Scene* s = static_cast<Scene*>(h);
  for (size_t i = 0; i < s->bodies.size(); ++i) {
    b2Body* b = s->bodies[i];
    float* o = out + i * 12;
    const b2Transform& xf = b->GetTransform();
    o[0] = xf.p.x; o[1] = xf.p.y; o[2] = xf.q.s; o[3] = xf.q.c;
    o[4] = b->GetWorldCenter().x; o[5] = b->GetWorldCenter().y; o[6] = b->GetAngle();
    o[7] = b->GetLinearVelocity().x; o[8] = b->GetLinearVelocity().y; o[9] = b->GetAngularVelocity();
    o[10] = b->IsAwake() ? 1.0f : 0.0f;
    o[11] = (float)b->GetType();
  }
}
// out[n][8] = mass, inertia about the origin, localCenter.x, localCenter.y, linearDamping,
//             angularDamping, gravityScale, sleepingAllowed
void SHIM(scene_get_body_params)(void* h, float* out) {
  Scene* s = static_cast<Scene*>(h);
  for (size_t i = 0; i < s->bodies.size(); ++i) {
    b2Body* b = s->bodies[i];
    float* o = out + i * 8;
    o[0] = b->GetMass(); o[1] = b->GetInertia(); o[2] = b->GetLocalCenter().x; o[3] = b->GetLocalCenter().y;
    o[4] = b->GetLinearDamping(); o[5] = b->GetAngularDamping(); o[6] = b->GetGravityScale();
    o[7] = b->IsSleepingAllowed() ? 1.0f : 0.0f;
  }
}
int SHIM(scene_shape_quad_total)(void* h) {
  Scene* s = static_cast<Scene*>(h);
  int total = 0;
  for (b2Fixture* f : s->fixtures) total += scene_shape_quad_count(f->GetShape());
  return total;
}
// body[n], type[n], shapeOff[n], filter[n][3] = category, mask, group, material[n][4] =
// friction, restitution, restitutionThreshold, density; sensor[n]; quads = shape pool in fixture order
void SHIM(scene_get_fixtures)(void* h, int* body, int* type, int* shapeOff, int* filter, float* material, int* sensor,
                              float* quads) {
  Scene* s = static_cast<Scene*>(h);
  int off = 0;
  for (size_t i = 0; i < s->fixtures.size(); ++i) {
    b2Fixture* f = s->fixtures[i];
    body[i] = s->bodyIndex[f->GetBody()];
    type[i] = (int)f->GetType();
    shapeOff[i] = off;
    const b2Filter& fl = f->GetFilterData();
    filter[3 * i] = fl.categoryBits; filter[3 * i + 1] = fl.maskBits; filter[3 * i + 2] = fl.groupIndex;
    material[4 * i] = f->GetFriction(); material[4 * i + 1] = f->GetRestitution();
    material[4 * i + 2] = f->GetRestitutionThreshold(); material[4 * i + 3] = f->GetDensity();
    sensor[i] = f->IsSensor() ? 1 : 0;
    scene_write_shape_quads(f->GetShape(), quads + 4 * off);
    off += scene_shape_quad_count(f->GetShape());
  }
}
// aabb[n][4] of every fixture as the broadphase sees it
void SHIM(scene_get_aabbs)(void* h, float* aabb) {
  Scene* s = static_cast<Scene*>(h);
  for (size_t i = 0; i < s->fixtures.size(); ++i) {
    const b2AABB& bb = s->fixtures[i]->GetAABB();
    aabb[4 * i] = bb.lowerBound.x; aabb[4 * i + 1] = bb.lowerBound.y;
    aabb[4 * i + 2] = bb.upperBound.x; aabb[4 * i + 3] = bb.upperBound.y;
  }
}
// world contact list in list order: fixA[n], fixB[n] (creation-order fixture indices),
// flags[n] (bit 0 touching, bit 1 enabled), manifold[n][16] in the include/b2cuda.h layout,
// material[n][4] = friction, restitution, restitutionThreshold, tangentSpeed.  Returns n.
int SHIM(scene_get_contacts)(void* h, int cap, int* fixA, int* fixB, int* flags, float* manifold, float* material) {
  Scene* s = static_cast<Scene*>(h);
  int n = 0;
  for (b2Contact* c = s->world->GetContactListStart(); c != s->world->GetContactListEnd(); c = c->GetNext()) {
    if (n >= cap) break;
    fixA[n] = s->fixtureIndex[c->GetFixtureA()];
    fixB[n] = s->fixtureIndex[c->GetFixtureB()];
    flags[n] = (c->IsTouching() ? 1 : 0) | (c->IsEnabled() ? 2 : 0);
    const b2Manifold* m = c->GetManifold();
    float* q = manifold + 16 * n;
    q[0] = m->localNormal.x; q[1] = m->localNormal.y; q[2] = m->localPoint.x; q[3] = m->localPoint.y;
    for (int k = 0; k < 2; ++k) {
      q[4 + 4 * k] = m->points[k].localPoint.x; q[5 + 4 * k] = m->points[k].localPoint.y;
      q[6 + 4 * k] = m->points[k].normalImpulse; q[7 + 4 * k] = m->points[k].tangentImpulse;
      uint32_t key = m->points[k].id.key;
      memcpy(&q[12 + k], &key, 4);
    }
    int32_t type = (int32_t)m->type, count = m->pointCount;
    memcpy(&q[14], &type, 4);
    memcpy(&q[15], &count, 4);
    if (material) {
      material[4 * n] = c->GetFriction(); material[4 * n + 1] = c->GetRestitution();
      material[4 * n + 2] = c->GetRestitutionThreshold(); material[4 * n + 3] = c->GetTangentSpeed();
    }
    ++n;
  }
  return n;
}

// ---- spatial queries through the public API (b2World::QueryAABB / RayCast) ----------------------
namespace {
struct ShimQueryAll : b2QueryCallback {
  Scene* s;
  std::vector<int> found;
  bool ReportFixture(b2Fixture* f) override {
    found.push_back(s->fixtureIndex[f]);
    return true;
  }
};
struct ShimRayClosest : b2RayCastCallback {
  Scene* s;
  int fixture = -1;
  float fraction = 0.0f;
  b2Vec2 normal, point;
  float ReportFixture(b2Fixture* f, const b2Vec2& p, const b2Vec2& n, float fr) override {
    fixture = s->fixtureIndex[f];
    fraction = fr;
    normal = n;
    point = p;
    return fr;
  }
};
struct ShimRayClosestLean : b2RayCastCallback {  // what a user's closest-hit callback costs: no index lookup
  b2Fixture* fixture = nullptr;
  float fraction = 0.0f;
  float ReportFixture(b2Fixture* f, const b2Vec2&, const b2Vec2&, float fr) override {
    fixture = f;
    fraction = fr;
    return fr;
  }
};
struct ShimRayAll : b2RayCastCallback {
  Scene* s;
  std::vector<int> fixtures;
  std::vector<float> fractions, normals;
  float ReportFixture(b2Fixture* f, const b2Vec2&, const b2Vec2& n, float fr) override {
    fixtures.push_back(s->fixtureIndex[f]);
    fractions.push_back(fr);
    normals.push_back(n.x);
    normals.push_back(n.y);
    return 1.0f;
  }
};
}  // namespace
// aabbs[n][4]; counts[n]; fixtures[n][cap] ascending.  Returns the largest count.
int SHIM(scene_query_aabb)(void* h, int n, const float* aabbs, int cap, int* counts, int* fixtures) {
  Scene* s = static_cast<Scene*>(h);
  int most = 0;
  for (int i = 0; i < n; ++i) {
    ShimQueryAll cb;
    cb.s = s;
    b2AABB box;
    box.lowerBound.Set(aabbs[4 * i], aabbs[4 * i + 1]);
    box.upperBound.Set(aabbs[4 * i + 2], aabbs[4 * i + 3]);
    s->world->QueryAABB(&cb, box);
    std::sort(cb.found.begin(), cb.found.end());
    counts[i] = (int)cb.found.size();
    if (counts[i] > most) most = counts[i];
    for (int k = 0; k < counts[i] && k < cap; ++k) fixtures[(size_t)i * cap + k] = cb.found[k];
  }
  return most;
}
// rays[n][4]; fixture[n] (-1 = no hit), fraction[n], normal[n][2], point[n][2]
void SHIM(scene_ray_cast_closest)(void* h, int n, const float* rays, int* fixture, float* fraction, float* normal,
                                  float* point) {
  Scene* s = static_cast<Scene*>(h);
  for (int i = 0; i < n; ++i) {
    ShimRayClosest cb;
    cb.s = s;
    cb.normal.SetZero();
    cb.point.SetZero();
    s->world->RayCast(&cb, b2Vec2(rays[4 * i], rays[4 * i + 1]), b2Vec2(rays[4 * i + 2], rays[4 * i + 3]));
    fixture[i] = cb.fixture;
    fraction[i] = cb.fraction;
    normal[2 * i] = cb.normal.x; normal[2 * i + 1] = cb.normal.y;
    point[2 * i] = cb.point.x; point[2 * i + 1] = cb.point.y;
  }
}
// every hit of ONE ray (callback returns 1), sorted by (fraction, fixture).  Returns the count.
int SHIM(scene_ray_cast_all)(void* h, const float* ray, int cap, int* fixtures, float* fractions, float* normals) {
  Scene* s = static_cast<Scene*>(h);
  ShimRayAll cb;
  cb.s = s;
  s->world->RayCast(&cb, b2Vec2(ray[0], ray[1]), b2Vec2(ray[2], ray[3]));
  int n = (int)cb.fixtures.size();
  std::vector<int> order(n);
  for (int i = 0; i < n; ++i) order[i] = i;
  std::sort(order.begin(), order.end(), [&](int a, int b) {
    return cb.fractions[a] < cb.fractions[b] || (cb.fractions[a] == cb.fractions[b] && cb.fixtures[a] < cb.fixtures[b]);
  });
  for (int k = 0; k < n && k < cap; ++k) {
    int i = order[k];
    fixtures[k] = cb.fixtures[i];
    fractions[k] = cb.fractions[i];
    normals[2 * k] = cb.normals[2 * i];
    normals[2 * k + 1] = cb.normals[2 * i + 1];
  }
  return n;
}
// wall-clock milliseconds for n closest-hit ray casts (CPU baseline of the query path)
double SHIM(scene_time_ray_casts)(void* h, int n, const float* rays) {
  Scene* s = static_cast<Scene*>(h);
  auto t0 = std::chrono::steady_clock::now();
  int hits = 0;
  for (int i = 0; i < n; ++i) {
    ShimRayClosestLean cb;
    s->world->RayCast(&cb, b2Vec2(rays[4 * i], rays[4 * i + 1]), b2Vec2(rays[4 * i + 2], rays[4 * i + 3]));
    hits += cb.fixture != nullptr;
  }
  auto t1 = std::chrono::steady_clock::now();
  return std::chrono::duration<double, std::milli>(t1 - t0).count() + 0.0 * hits;
}

int SHIM(scene_joint_count)(void* h) { return static_cast<Scene*>(h)->world->GetJointCount(); }
// joints in creation order: bodies[n][2], anchors[n][4] = localAnchorA.xy, localAnchorB.xy, params[n][12]
// in the include/b2cuda.h b2gJointArrays layout:
//   revolute: referenceAngle, lowerAngle, upperAngle, maxMotorTorque, motorSpeed, bits(flags), 0, 0
//   distance: length, minLength, maxLength, stiffness, damping, bits(flags | 1 << 8), 0, 0
//   weld:     referenceAngle, stiffness, damping, 0, 0, bits(flags | 2 << 8), 0, 0
//   mouse:    maxForce, stiffness, damping, 0, 0, bits(flags | 7 << 8), 0...; anchors = target.xy, localAnchorB.xy
//   friction: maxForce, maxTorque, 0, 0, 0, bits(flags | 5 << 8), 0...
//   motor:    maxForce, maxTorque, correctionFactor, angularOffset, 0, bits(flags | 6 << 8), 0...;
//             anchors = linearOffset.xy, 0, 0
//   wheel:    stiffness, lowerTranslation, upperTranslation, maxMotorTorque, motorSpeed, bits(flags | 4 << 8),
//             localAxisA.x, localAxisA.y, damping, 0, 0, 0
//   prismatic: referenceAngle, lowerTranslation, upperTranslation, maxMotorForce, motorSpeed,
//             bits(flags | 3 << 8), localAxisA.x, localAxisA.y
// flags: 1 = enableLimit, 2 = enableMotor, 4 = collideConnected.  Other joint types are skipped.
int SHIM(scene_get_joints)(void* h, int cap, int* bodies, float* anchors, float* params) {
  Scene* s = static_cast<Scene*>(h);
  std::vector<b2Joint*> js;
  for (b2Joint* j = s->world->GetJointList(); j; j = j->GetNext()) js.push_back(j);
  int n = 0;
  for (auto it = js.rbegin(); it != js.rend() && n < cap; ++it) {  // the list is newest-first
    b2Joint* j = *it;
    float* p = params + 12 * n;
    float p6 = 0.0f, p7 = 0.0f;
    p[8] = p[9] = p[10] = p[11] = 0.0f;
    uint32_t fl = j->GetCollideConnected() ? 4u : 0u;
    if (j->GetType() == e_revoluteJoint) {
      b2RevoluteJoint* r = static_cast<b2RevoluteJoint*>(j);
      anchors[4 * n] = r->GetLocalAnchorA().x; anchors[4 * n + 1] = r->GetLocalAnchorA().y;
      anchors[4 * n + 2] = r->GetLocalAnchorB().x; anchors[4 * n + 3] = r->GetLocalAnchorB().y;
      p[0] = r->GetReferenceAngle(); p[1] = r->GetLowerLimit(); p[2] = r->GetUpperLimit();
      p[3] = r->GetMaxMotorTorque(); p[4] = r->GetMotorSpeed();
      fl |= (r->IsLimitEnabled() ? 1u : 0u) | (r->IsMotorEnabled() ? 2u : 0u);
    } else if (j->GetType() == e_distanceJoint) {
      b2DistanceJoint* d = static_cast<b2DistanceJoint*>(j);
      anchors[4 * n] = d->GetLocalAnchorA().x; anchors[4 * n + 1] = d->GetLocalAnchorA().y;
      anchors[4 * n + 2] = d->GetLocalAnchorB().x; anchors[4 * n + 3] = d->GetLocalAnchorB().y;
      p[0] = d->GetLength(); p[1] = d->GetMinLength(); p[2] = d->GetMaxLength();
      p[3] = d->GetStiffness(); p[4] = d->GetDamping();
      fl |= 1u << 8;
    } else if (j->GetType() == e_weldJoint) {
      b2WeldJoint* wj = static_cast<b2WeldJoint*>(j);
      anchors[4 * n] = wj->GetLocalAnchorA().x; anchors[4 * n + 1] = wj->GetLocalAnchorA().y;
      anchors[4 * n + 2] = wj->GetLocalAnchorB().x; anchors[4 * n + 3] = wj->GetLocalAnchorB().y;
      p[0] = wj->GetReferenceAngle(); p[1] = wj->GetStiffness(); p[2] = wj->GetDamping(); p[3] = 0.0f; p[4] = 0.0f;
      fl |= 2u << 8;
    } else if (j->GetType() == e_mouseJoint) {
      b2MouseJoint* mo = static_cast<b2MouseJoint*>(j);
#ifdef B2G_WORLD_H
      const b2Vec2 lb = mo->GetLocalAnchorB();
#else  // the reference keeps it protected; the oracle harness is compiled with -fno-access-control
      const b2Vec2 lb = mo->m_localAnchorB;
#endif
      anchors[4 * n] = mo->GetTarget().x; anchors[4 * n + 1] = mo->GetTarget().y;
      anchors[4 * n + 2] = lb.x; anchors[4 * n + 3] = lb.y;
      p[0] = mo->GetMaxForce(); p[1] = mo->GetStiffness(); p[2] = mo->GetDamping(); p[3] = p[4] = 0.0f;
      fl |= 7u << 8;
    } else if (j->GetType() == e_frictionJoint) {
      b2FrictionJoint* fj = static_cast<b2FrictionJoint*>(j);
      anchors[4 * n] = fj->GetLocalAnchorA().x; anchors[4 * n + 1] = fj->GetLocalAnchorA().y;
      anchors[4 * n + 2] = fj->GetLocalAnchorB().x; anchors[4 * n + 3] = fj->GetLocalAnchorB().y;
      p[0] = fj->GetMaxForce(); p[1] = fj->GetMaxTorque(); p[2] = p[3] = p[4] = 0.0f;
      fl |= 5u << 8;
    } else if (j->GetType() == e_motorJoint) {
      b2MotorJoint* mj = static_cast<b2MotorJoint*>(j);
      anchors[4 * n] = mj->GetLinearOffset().x; anchors[4 * n + 1] = mj->GetLinearOffset().y;
      anchors[4 * n + 2] = anchors[4 * n + 3] = 0.0f;
      p[0] = mj->GetMaxForce(); p[1] = mj->GetMaxTorque(); p[2] = mj->GetCorrectionFactor(); p[3] = mj->GetAngularOffset();
      p[4] = 0.0f;
      fl |= 6u << 8;
    } else if (j->GetType() == e_wheelJoint) {
      b2WheelJoint* wh = static_cast<b2WheelJoint*>(j);
      anchors[4 * n] = wh->GetLocalAnchorA().x; anchors[4 * n + 1] = wh->GetLocalAnchorA().y;
      anchors[4 * n + 2] = wh->GetLocalAnchorB().x; anchors[4 * n + 3] = wh->GetLocalAnchorB().y;
      p[0] = wh->GetStiffness(); p[1] = wh->GetLowerLimit(); p[2] = wh->GetUpperLimit();
      p[3] = wh->GetMaxMotorTorque(); p[4] = wh->GetMotorSpeed();
      p6 = wh->GetLocalAxisA().x; p7 = wh->GetLocalAxisA().y;
      p[8] = wh->GetDamping();
      fl |= (wh->IsLimitEnabled() ? 1u : 0u) | (wh->IsMotorEnabled() ? 2u : 0u) | (4u << 8);
    } else if (j->GetType() == e_prismaticJoint) {
      b2PrismaticJoint* pj = static_cast<b2PrismaticJoint*>(j);
      anchors[4 * n] = pj->GetLocalAnchorA().x; anchors[4 * n + 1] = pj->GetLocalAnchorA().y;
      anchors[4 * n + 2] = pj->GetLocalAnchorB().x; anchors[4 * n + 3] = pj->GetLocalAnchorB().y;
      p[0] = pj->GetReferenceAngle(); p[1] = pj->GetLowerLimit(); p[2] = pj->GetUpperLimit();
      p[3] = pj->GetMaxMotorForce(); p[4] = pj->GetMotorSpeed();
      p6 = pj->GetLocalAxisA().x; p7 = pj->GetLocalAxisA().y;
      fl |= (pj->IsLimitEnabled() ? 1u : 0u) | (pj->IsMotorEnabled() ? 2u : 0u) | (3u << 8);
    } else {
      continue;
    }
    bodies[2 * n] = s->bodyIndex[j->GetBodyA()];
    bodies[2 * n + 1] = s->bodyIndex[j->GetBodyB()];
    memcpy(&p[5], &fl, 4);
    p[6] = p6;
    p[7] = p7;
    ++n;
  }
  return n;
}

}  // extern "C"
#endif

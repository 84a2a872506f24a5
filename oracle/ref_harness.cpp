// ref_harness.cpp — TEST INFRASTRUCTURE ONLY.  Never linked into the product.
//
// Compiled by oracle/Makefile together with the UNMODIFIED reference sources where they lie under
// /root/reference (src/collision, src/common, src/dynamics, src/particle) into
// oracle/_ref/libb2ref.so.  Nothing from the reference is copied into this repository: this
// file only *calls* it.  It exposes
//   (1) the shared scene shim (scenes/b2_scene_shim.h) expanded with the b2ref_ prefix, i.e. the
//       reference's own b2World::Step on the BASELINE scenes — the parity oracle and the
//       `cpu_baseline.kind = "reference"` timing arm;
//   (2) white-box taps (built with -fno-access-control, as SURVEY.md §8c / Appendix C describe):
//       the five b2Collide* functions on explicit shapes, b2PolygonShape::Set, ComputeAABB,
//       b2ContactManager::Collide on a live world, and b2ContactSolver driven on explicit arrays
//       with per-iteration iterates dumped.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs load it.
#include <cstring>
#include <vector>
#include "box2d/box2d.h"
#include "dynamics/b2_contact_solver.h"  // private header, reachable through -I/root/reference/src

#define SHIM(name) b2ref_##name
#include "b2_scene_shim.h"

namespace {

struct ShapeBox {
  b2CircleShape circle;
  b2EdgeShape edge;
  b2PolygonShape poly;
  b2Shape* get(int type) {
    if (type == 0) return &circle;
    if (type == 1) return &edge;
    return &poly;
  }
};

void shape_from_quads(ShapeBox& s, int type, const float* q) {
  if (type == 0) {
    s.circle.m_p.Set(q[0], q[1]);
    s.circle.m_radius = q[2];
  } else if (type == 1) {
    s.edge.m_vertex1.Set(q[0], q[1]);
    s.edge.m_vertex2.Set(q[2], q[3]);
    s.edge.m_vertex0.Set(q[4], q[5]);
    s.edge.m_vertex3.Set(q[6], q[7]);
    s.edge.m_radius = q[8];
    s.edge.m_oneSided = q[9] != 0.0f;
  } else {
    s.poly.m_centroid.Set(q[0], q[1]);
    s.poly.m_radius = q[2];
    s.poly.m_count = (int32)q[3];
    for (int i = 0; i < s.poly.m_count; ++i) {
      s.poly.m_vertices[i].Set(q[4 + 4 * i], q[5 + 4 * i]);
      s.poly.m_normals[i].Set(q[6 + 4 * i], q[7 + 4 * i]);
    }
  }
}

b2Transform xf_from(const float* x) {
  b2Transform t;
  t.p.Set(x[0], x[1]);
  t.q.s = x[2];
  t.q.c = x[3];
  return t;
}

void manifold_to_floats(const b2Manifold& m, float* q) {
  q[0] = m.localNormal.x; q[1] = m.localNormal.y; q[2] = m.localPoint.x; q[3] = m.localPoint.y;
  for (int k = 0; k < 2; ++k) {
    q[4 + 4 * k] = m.points[k].localPoint.x; q[5 + 4 * k] = m.points[k].localPoint.y;
    q[6 + 4 * k] = m.points[k].normalImpulse; q[7 + 4 * k] = m.points[k].tangentImpulse;
    uint32_t key = m.points[k].id.key;
    memcpy(&q[12 + k], &key, 4);
  }
  int32_t type = (int32_t)m.type, count = m.pointCount;
  memcpy(&q[14], &type, 4);
  memcpy(&q[15], &count, 4);
}

void manifold_from_floats(b2Manifold& m, const float* q) {
  m.localNormal.Set(q[0], q[1]);
  m.localPoint.Set(q[2], q[3]);
  for (int k = 0; k < 2; ++k) {
    m.points[k].localPoint.Set(q[4 + 4 * k], q[5 + 4 * k]);
    m.points[k].normalImpulse = q[6 + 4 * k];
    m.points[k].tangentImpulse = q[7 + 4 * k];
    memcpy(&m.points[k].id.key, &q[12 + k], 4);
  }
  int32_t type, count;
  memcpy(&type, &q[14], 4);
  memcpy(&count, &q[15], 4);
  m.type = (b2Manifold::Type)type;
  m.pointCount = count;
}

}  // namespace

extern "C" {

// b2PolygonShape::Set on `count` points -> shape-pool record (1 + m_count quads). Returns m_count.
int b2ref_polygon_set(const float* points, int count, float* quads) {
  b2Vec2 pts[b2_maxPolygonVertices];
  int n = count < b2_maxPolygonVertices ? count : b2_maxPolygonVertices;
  for (int i = 0; i < n; ++i) pts[i].Set(points[2 * i], points[2 * i + 1]);
  b2PolygonShape p;
  p.Set(pts, n);
  scene_write_shape_quads(&p, quads);
  return p.m_count;
}

// polygon mass data through the reference: out = mass, center.x, center.y, I
void b2ref_shape_mass(int type, const float* quads, float density, float* out) {
  ShapeBox s;
  shape_from_quads(s, type, quads);
  b2MassData md;
  s.get(type)->ComputeMass(&md, density);
  out[0] = md.mass; out[1] = md.center.x; out[2] = md.center.y; out[3] = md.I;
}

void b2ref_compute_aabbs(int n, const int* type, const int* shapeOff, const float* quads, const float* xf, float* aabb) {
  for (int i = 0; i < n; ++i) {
    ShapeBox s;
    shape_from_quads(s, type[i], quads + 4 * shapeOff[i]);
    b2AABB bb;
    s.get(type[i])->ComputeAABB(&bb, xf_from(xf + 4 * i));
    aabb[4 * i] = bb.lowerBound.x; aabb[4 * i + 1] = bb.lowerBound.y;
    aabb[4 * i + 2] = bb.upperBound.x; aabb[4 * i + 3] = bb.upperBound.y;
  }
}

// the five b2Collide* functions on n ordered pairs; returns -1 if a pair has no function
int b2ref_collide_pairs(int n, const int* typeA, const int* offA, const float* xfA, const int* typeB, const int* offB,
                        const float* xfB, const float* quads, float* manifold) {
  for (int i = 0; i < n; ++i) {
    ShapeBox a, b;
    shape_from_quads(a, typeA[i], quads + 4 * offA[i]);
    shape_from_quads(b, typeB[i], quads + 4 * offB[i]);
    b2Transform ta = xf_from(xfA + 4 * i), tb = xf_from(xfB + 4 * i);
    b2Manifold m;
    memset(&m, 0, sizeof(m));
    if (typeA[i] == 0 && typeB[i] == 0) b2CollideCircles(&m, &a.circle, ta, &b.circle, tb);
    else if (typeA[i] == 2 && typeB[i] == 0) b2CollidePolygonAndCircle(&m, &a.poly, ta, &b.circle, tb);
    else if (typeA[i] == 2 && typeB[i] == 2) b2CollidePolygons(&m, &a.poly, ta, &b.poly, tb);
    else if (typeA[i] == 1 && typeB[i] == 0) b2CollideEdgeAndCircle(&m, &a.edge, ta, &b.circle, tb);
    else if (typeA[i] == 1 && typeB[i] == 2) b2CollideEdgeAndPolygon(&m, &a.edge, ta, &b.poly, tb);
    else return -1;
    manifold_to_floats(m, manifold + 16 * i);
  }
  return 0;
}

// run the narrowphase of the NEXT step on a live world now (b2ContactManager::Collide)
void b2ref_world_collide(void* h) { static_cast<Scene*>(h)->world->m_contactManager.Collide(); }

// out[n][2] = m_invMass, m_invI (private members; exact solver inputs)
void b2ref_get_body_inv(void* h, float* out) {
  Scene* s = static_cast<Scene*>(h);
  for (size_t i = 0; i < s->bodies.size(); ++i) {
    out[2 * i] = s->bodies[i]->m_invMass;
    out[2 * i + 1] = s->bodies[i]->m_invI;
  }
}
float b2ref_get_inv_dt0(void* h) { return static_cast<Scene*>(h)->world->m_inv_dt0; }
// out[n] = sleep timers
void b2ref_get_sleep_times(void* h, float* out) {
  Scene* s = static_cast<Scene*>(h);
  for (size_t i = 0; i < s->bodies.size(); ++i) out[i] = s->bodies[i]->m_sleepTime;
}

// revolute joints in creation order (the order of scene_get_joints): out[n][5] = m_impulse.xy,
// m_motorImpulse, m_lowerImpulse, m_upperImpulse (b2_revolute_joint.h:178-181).  Returns n.
int b2ref_get_joint_state(void* h, int cap, float* out) {
  Scene* s = static_cast<Scene*>(h);
  std::vector<b2Joint*> js;
  for (b2Joint* j = s->world->GetJointList(); j; j = j->GetNext()) js.push_back(j);
  int n = 0;
  for (auto it = js.rbegin(); it != js.rend() && n < cap; ++it) {
    float* o = out + 5 * n;
    if ((*it)->GetType() == e_revoluteJoint) {
      b2RevoluteJoint* r = static_cast<b2RevoluteJoint*>(*it);
      o[0] = r->m_impulse.x; o[1] = r->m_impulse.y; o[2] = r->m_motorImpulse; o[3] = r->m_lowerImpulse; o[4] = r->m_upperImpulse;
    } else if ((*it)->GetType() == e_distanceJoint) {  // b2_distance_joint.h:157-159
      b2DistanceJoint* d = static_cast<b2DistanceJoint*>(*it);
      o[0] = d->m_impulse; o[1] = 0.0f; o[2] = 0.0f; o[3] = d->m_lowerImpulse; o[4] = d->m_upperImpulse;
    } else if ((*it)->GetType() == e_weldJoint) {  // b2_weld_joint.h:112
      b2WeldJoint* wj = static_cast<b2WeldJoint*>(*it);
      o[0] = wj->m_impulse.x; o[1] = wj->m_impulse.y; o[2] = wj->m_impulse.z; o[3] = 0.0f; o[4] = 0.0f;
    } else if ((*it)->GetType() == e_mouseJoint) {  // b2_mouse_joint.h:113
      b2MouseJoint* mo = static_cast<b2MouseJoint*>(*it);
      o[0] = mo->m_impulse.x; o[1] = mo->m_impulse.y; o[2] = 0.0f; o[3] = 0.0f; o[4] = 0.0f;
    } else if ((*it)->GetType() == e_frictionJoint) {  // b2_friction_joint.h:86-87
      b2FrictionJoint* fj = static_cast<b2FrictionJoint*>(*it);
      o[0] = fj->m_linearImpulse.x; o[1] = fj->m_linearImpulse.y; o[2] = fj->m_angularImpulse; o[3] = 0.0f; o[4] = 0.0f;
    } else if ((*it)->GetType() == e_motorJoint) {  // b2_motor_joint.h:104-105
      b2MotorJoint* mj = static_cast<b2MotorJoint*>(*it);
      o[0] = mj->m_linearImpulse.x; o[1] = mj->m_linearImpulse.y; o[2] = mj->m_angularImpulse; o[3] = 0.0f; o[4] = 0.0f;
    } else if ((*it)->GetType() == e_wheelJoint) {  // b2_wheel_joint.h:196-200
      b2WheelJoint* wh = static_cast<b2WheelJoint*>(*it);
      o[0] = wh->m_impulse; o[1] = wh->m_springImpulse; o[2] = wh->m_motorImpulse; o[3] = wh->m_lowerImpulse; o[4] = wh->m_upperImpulse;
    } else if ((*it)->GetType() == e_prismaticJoint) {  // b2_prismatic_joint.h:164-167
      b2PrismaticJoint* pj = static_cast<b2PrismaticJoint*>(*it);
      o[0] = pj->m_impulse.x; o[1] = pj->m_impulse.y; o[2] = pj->m_motorImpulse; o[3] = pj->m_lowerImpulse; o[4] = pj->m_upperImpulse;
    } else {
      continue;
    }
    ++n;
  }
  return n;
}

// The order in which the NEXT b2World::Step will add joints to its islands (island.Add(joint),
// b2_world.cpp:622-647).  Joints have no PostSolve, so the head of Step is run here — the pending pair
// refresh and b2ContactManager::Collide (b2_world.cpp:1114-1138), which Step will simply repeat with the
// same outcome — and then the traversal of b2World::Solve (:522-659) is walked on the reference's own
// body list, per-body contact arrays and joint edge lists, with exactly the flags Solve will see.
// out = creation-order joint indices (the order of scene_get_joints).  Returns the count.
int b2ref_next_step_joint_order(void* h, int cap, int* out) {
  Scene* s = static_cast<Scene*>(h);
  b2World* w = s->world;
  if (w->m_newContacts) {
    w->m_contactManager.FindNewContacts();
    w->m_newContacts = false;
  }
  w->m_contactManager.Collide();
  std::vector<b2Joint*> js;
  for (b2Joint* j = w->GetJointList(); j; j = j->GetNext()) js.push_back(j);
  std::unordered_map<const b2Joint*, int> jointIndex;
  {
    int n = 0;
    for (auto it = js.rbegin(); it != js.rend(); ++it)
      if ((*it)->GetType() == e_revoluteJoint || (*it)->GetType() == e_distanceJoint || (*it)->GetType() == e_weldJoint ||
          (*it)->GetType() == e_prismaticJoint || (*it)->GetType() == e_wheelJoint ||
          (*it)->GetType() == e_frictionJoint || (*it)->GetType() == e_motorJoint || (*it)->GetType() == e_mouseJoint)
        jointIndex[*it] = n++;
  }
  std::unordered_map<const b2Body*, bool> bodySeen;
  std::unordered_map<const b2Contact*, bool> contactSeen;
  std::unordered_map<const b2Joint*, bool> jointSeen;
  std::vector<b2Body*> stack;
  int n = 0;
  for (b2Body* seed = w->m_bodyListHead; seed; seed = seed->m_next) {
    if (bodySeen[seed]) continue;
    if (!seed->IsEnabled()) continue;
    if (!seed->IsAwake()) continue;
    if (seed->GetType() == b2_staticBody) break;
    if (seed->GetContactCount() == 0 && seed->GetJointList() == nullptr) {
      bodySeen[seed] = true;
      continue;
    }
    std::vector<b2Body*> islandBodies;
    stack.clear();
    stack.push_back(seed);
    bodySeen[seed] = true;
    while (!stack.empty()) {
      b2Body* b = stack.back();
      stack.pop_back();
      islandBodies.push_back(b);
      if (b->GetType() == b2_staticBody) continue;
      for (int32 i = 0; i < b->GetContactCount(); ++i) {
        b2Contact* c = b->GetContact(i);
        if (contactSeen[c]) continue;
        if (!c->IsEnabled() || !c->IsTouching()) continue;
        if (c->GetFixtureA()->IsSensor() || c->GetFixtureB()->IsSensor()) continue;
        contactSeen[c] = true;
        b2Body* bA = c->GetFixtureA()->GetBody();
        b2Body* bB = c->GetFixtureB()->GetBody();
        b2Body* other = bA == b ? bB : bA;
        if (bodySeen[other]) continue;
        stack.push_back(other);
        bodySeen[other] = true;
      }
      for (b2JointEdge* je = b->GetJointList(); je; je = je->next) {
        if (jointSeen[je->joint]) continue;
        b2Body* other = je->other;
        if (!other->IsEnabled()) continue;
        jointSeen[je->joint] = true;
        auto it = jointIndex.find(je->joint);
        if (it != jointIndex.end() && n < cap) out[n++] = it->second;
        if (bodySeen[other]) continue;
        stack.push_back(other);
        bodySeen[other] = true;
      }
    }
    for (b2Body* b : islandBodies)
      if (b->GetType() == b2_staticBody) bodySeen[b] = false;  // static bodies take part in other islands too
  }
  return n;
}

// one b2World::Step with a PostSolve tap: the island solver's visiting order (SURVEY Appendix C:
// PostSolve order IS the solver order).  Returns the number of solved contacts.
namespace {
struct OrderTap : b2ContactListener {
  Scene* scene;
  std::vector<std::pair<int, int>> order;
  void PostSolve(b2Contact* c, const b2ContactImpulse*) override {
    order.push_back(std::make_pair(scene->fixtureIndex[c->GetFixtureA()], scene->fixtureIndex[c->GetFixtureB()]));
  }
};
}  // namespace
int b2ref_step_recording_order(void* h, int cap, int* fixA, int* fixB) {
  Scene* s = static_cast<Scene*>(h);
  OrderTap tap;
  tap.scene = s;
  s->world->SetContactListener(&tap);
  s->step();
  s->world->SetContactListener(nullptr);
  int n = (int)tap.order.size() < cap ? (int)tap.order.size() : cap;
  for (int i = 0; i < n; ++i) {
    fixA[i] = tap.order[i].first;
    fixB[i] = tap.order[i].second;
  }
  return (int)tap.order.size();
}

// b2ContactSolver on explicit arrays, driven as b2Island::Solve drives it (b2_island.cpp:306-409).
// Same signature and array layouts as b2g_solve_sequential in include/b2cuda.h (minus `device`).
int b2ref_solve(int nb, float* pos, float* vel, const float* mass, int nc, const int* index, float* manifold,
                const float* material, const float* radii, float dt, float dtRatio, int warmStarting, int velIters,
                int posIters, float* velIterates, float* posIterates, int* posItersDone) {
  std::vector<b2Position> positions(nb);
  std::vector<b2Velocity> velocities(nb);
  for (int i = 0; i < nb; ++i) {
    positions[i].c.Set(pos[4 * i], pos[4 * i + 1]);
    positions[i].a = pos[4 * i + 2];
    velocities[i].v.Set(vel[4 * i], vel[4 * i + 1]);
    velocities[i].w = vel[4 * i + 2];
  }
  // stand-in bodies / fixtures / contacts carrying exactly the fields b2ContactSolver::Initialize reads
  b2BodyDef bd;
  bd.type = b2_dynamicBody;
  std::vector<b2Body*> bodies(nb);
  for (int i = 0; i < nb; ++i) {
    bodies[i] = new b2Body(&bd, nullptr);
    bodies[i]->m_islandIndex = i;
    bodies[i]->m_invMass = mass[4 * i];
    bodies[i]->m_invI = mass[4 * i + 1];
    bodies[i]->m_sweep.localCenter.Set(mass[4 * i + 2], mass[4 * i + 3]);
  }
  std::vector<b2CircleShape> shapes(2 * (size_t)nc);
  std::vector<b2Fixture*> fixtures(2 * (size_t)nc);
  std::vector<b2Contact*> contacts(nc);
  for (int i = 0; i < nc; ++i) {
    for (int k = 0; k < 2; ++k) {
      shapes[2 * i + k].m_radius = radii[2 * i + k];
      b2Fixture* f = new b2Fixture();
      f->m_shape = &shapes[2 * i + k];
      f->m_body = bodies[index[2 * i + k]];
      f->m_friction = 0.0f;
      f->m_restitution = 0.0f;
      f->m_restitutionThreshold = 0.0f;
      f->m_isSensor = false;
      fixtures[2 * i + k] = f;
    }
    void* mem = malloc(sizeof(b2Contact));
    b2Contact* c = new (mem) b2Contact(fixtures[2 * i], fixtures[2 * i + 1], nullptr);
    c->m_friction = material[4 * i];
    c->m_restitution = material[4 * i + 1];
    c->m_restitutionThreshold = material[4 * i + 2];
    c->m_tangentSpeed = material[4 * i + 3];
    manifold_from_floats(c->m_manifold, manifold + 16 * i);
    contacts[i] = c;
  }

  b2StackAllocator allocator;
  b2TimeStep step;
  step.dt = dt;
  step.inv_dt = dt > 0.0f ? 1.0f / dt : 0.0f;
  step.dtRatio = dtRatio;
  step.velocityIterations = velIters;
  step.positionIterations = posIters;
  step.particleIterations = 1;
  step.warmStarting = warmStarting != 0;
  int done = 0;
  {
    b2ContactSolver solver;
    if (nc > 0) {
      b2ContactSolverDef def;
      def.step = step;
      def.contacts = contacts.data();
      def.count = nc;
      def.positions = positions.data();
      def.velocities = velocities.data();
      def.allocator = &allocator;
      solver.Initialize(&def);
      solver.InitializeVelocityConstraints();
      if (step.warmStarting) solver.WarmStart();
    }
    for (int it = 0; it < velIters; ++it) {
      if (nc > 0) solver.SolveVelocityConstraints();
      if (velIterates)
        for (int i = 0; i < nb; ++i) {
          float* o = velIterates + ((size_t)it * nb + i) * 4;
          o[0] = velocities[i].v.x; o[1] = velocities[i].v.y; o[2] = velocities[i].w; o[3] = 0.0f;
        }
    }
    if (nc > 0) solver.StoreImpulses();
    // integrate positions exactly as b2_island.cpp:353-385
    for (int i = 0; i < nb; ++i) {
      b2Vec2 c = positions[i].c;
      float a = positions[i].a;
      b2Vec2 v = velocities[i].v;
      float w = velocities[i].w;
      b2Vec2 translation = dt * v;
      if (b2Dot(translation, translation) > b2_maxTranslationSquared) {
        float ratio = b2_maxTranslation / translation.Length();
        v *= ratio;
      }
      float rotation = dt * w;
      if (rotation * rotation > b2_maxRotationSquared) {
        float ratio = b2_maxRotation / b2Abs(rotation);
        w *= ratio;
      }
      c += dt * v;
      a += dt * w;
      positions[i].c = c;
      positions[i].a = a;
      velocities[i].v = v;
      velocities[i].w = w;
    }
    for (int it = 0; it < posIters; ++it) {
      bool ok = nc > 0 ? solver.SolvePositionConstraints() : true;
      ++done;
      if (posIterates)
        for (int i = 0; i < nb; ++i) {
          float* o = posIterates + ((size_t)it * nb + i) * 4;
          o[0] = positions[i].c.x; o[1] = positions[i].c.y; o[2] = positions[i].a; o[3] = 0.0f;
        }
      if (ok) break;
    }
  }
  for (int i = 0; i < nb; ++i) {
    pos[4 * i] = positions[i].c.x; pos[4 * i + 1] = positions[i].c.y; pos[4 * i + 2] = positions[i].a;
    vel[4 * i] = velocities[i].v.x; vel[4 * i + 1] = velocities[i].v.y; vel[4 * i + 2] = velocities[i].w;
  }
  for (int i = 0; i < nc; ++i) {
    manifold_to_floats(contacts[i]->m_manifold, manifold + 16 * i);
    free(contacts[i]);
    delete fixtures[2 * i];
    delete fixtures[2 * i + 1];
  }
  for (int i = 0; i < nb; ++i) {
    b2Free(bodies[i]->m_contactList);
    delete bodies[i];
  }
  if (posItersDone) *posItersDone = done;
  return 0;
}

}  // extern "C"

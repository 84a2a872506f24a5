// -----------------------------------------------------------------------------------------------
// Third-party notice.  To stay source- and result-compatible with box2d-optimized, parts of this
// file restate declarations, inline math and creation-time algorithms of that library (itself a
// fork of Box2D).  Those parts are covered by the MIT License:
//   Copyright (c) 2019 Erin Catto, Copyright (c) 2020 Manolis Tsamis
// The full licence text and permission notice are in LICENSES/box2d-optimized-MIT.txt.
// -----------------------------------------------------------------------------------------------
// b2_scenes.h — the BASELINE.json scenes, written ONCE against the public Box2D API.
//
// This file includes only "box2d/box2d.h" and uses only b2World / b2Body / b2Fixture / shape
// calls, so the very same source compiles against
//   (a) the reference (-I/root/reference/include, linked with its sources) -> oracle/_ref/libb2ref.so
//   (b) this repo's drop-in API (-Iinclude, linked with libb2cuda)         -> libb2gpu_scenes.so
// which is the drop-in claim in executable form.  Scene definitions follow SURVEY.md §8(d):
//   pyramid        testbed/tests/pyramid.cpp:33-72 (config 1)
//   many_pyramids  `size` copies of it 30 m apart on one ground edge (config 2)
//   mixed          circles + convex polygons dropped into a 3-box container, LCG seed (config 3)
//   tumbler        testbed/benchmarks/benchmarks.h:137-204 (b3, config 4)
//   mixed_linked   mixed + a revolute joint between every 16th body and its neighbour (joints in an oversize island)
//   filters        category / mask / group filtering (b2_world_callbacks.cpp:28-40), all three group signs
//   chain / chain_collide   testbed/tests/chain.cpp:31-66 shape (collideConnected filter)
//   welds          weld joints: cantilever beams, a welded compound (b2_weld_joint.cpp)
//   cars           wheel joints: sprung, motorised cars driving over ramps and loose boxes (b2_wheel_joint.cpp)
//   mice           mouse joints: bodies dragged to world targets, soft and stiff, strong and too weak (b2_mouse_joint.cpp)
//   drags          friction joints (braked falling / spinning boxes) and motor joints (platforms driven to an
//                  offset, carrying boxes) (b2_friction_joint.cpp, b2_motor_joint.cpp)
//   sliders        prismatic joints: motorised pistons, limited rails, a free slider (b2_prismatic_joint.cpp)
//   springs        distance joints: rods, springs, limited ropes (b2_distance_joint.cpp)
//   sensors        sensor zones / paddle / probes in a rain of shapes (b2TestOverlap path)
//   hello          unit-test/hello_world.cpp:33-112
//   falling_squares / falling_circles   benchmarks.h:57-135 (b1, b2)
#ifndef B2_SCENES_H
#define B2_SCENES_H

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>
#include "box2d/box2d.h"

struct SceneLCG {
  uint32_t state;
  explicit SceneLCG(uint32_t seed) : state(seed) {}
  float next() {  // U[0,1)
    state = state * 1664525u + 1013904223u;
    return (float)(state >> 8) / 16777216.0f;
  }
  float range(float lo, float hi) { return lo + (hi - lo) * next(); }
};

struct Scene {
  b2World* world = nullptr;
  std::vector<b2Body*> bodies;        // creation order = device body index
  std::vector<b2Fixture*> fixtures;   // creation order = device fixture index
  std::unordered_map<const b2Fixture*, int> fixtureIndex;
  std::unordered_map<const b2Body*, int> bodyIndex;
  float dt = 1.0f / 60.0f;
  int velocityIterations = 8;
  int positionIterations = 3;
  std::string kind;
  int spawnTarget = 0, spawned = 0;  // tumbler: one box per step until spawnTarget
  // tumbler variants (seed > 0): every box is spawned at (0, 10) + U[-2, 2]^2 from an LCG seeded per
  // world, so batched worlds are different worlds, not clones; seed 0 = the reference's b3 scene
  bool spawnJitter = false;
  SceneLCG spawnRng{1u};
  int steps = 0;

  ~Scene() { delete world; }

  b2Body* addBody(const b2BodyDef& bd) {
    b2Body* b = world->CreateBody(&bd);
    bodyIndex[b] = (int)bodies.size();
    bodies.push_back(b);
    return b;
  }
  b2Fixture* addFixture(b2Body* b, const b2FixtureDef& fd) {
    b2Fixture* f = b->CreateFixture(&fd);
    fixtureIndex[f] = (int)fixtures.size();
    fixtures.push_back(f);
    return f;
  }
  b2Fixture* addFixture(b2Body* b, const b2Shape& shape, float density) {
    b2FixtureDef fd;
    fd.shape = &shape;
    fd.density = density;
    return addFixture(b, fd);
  }

  void step() {
    world->Step(dt, velocityIterations, positionIterations);
    if (kind == "tumbler" && spawned < spawnTarget) {
      b2BodyDef bd;
      bd.type = b2_dynamicBody;
      bd.position.Set(0.0f, 10.0f);
      if (spawnJitter) {
        float dx = spawnRng.range(-2.0f, 2.0f), dy = spawnRng.range(-2.0f, 2.0f);
        bd.position.Set(dx, 10.0f + dy);
      }
      b2Body* body = addBody(bd);
      b2PolygonShape shape;
      shape.SetAsBox(0.125f, 0.125f);
      addFixture(body, shape, 1.0f);
      ++spawned;
    }
    ++steps;
  }
};

inline void scene_add_pyramid(Scene& s, int rows, float xOffset) {
  float a = 0.5f;
  b2PolygonShape shape;
  shape.SetAsBox(a, a);
  b2Vec2 x(-7.0f + xOffset, 0.75f);
  b2Vec2 y;
  b2Vec2 deltaX(0.5625f, 1.25f);
  b2Vec2 deltaY(1.125f, 0.0f);
  for (int i = 0; i < rows; ++i) {
    y = x;
    for (int j = i; j < rows; ++j) {
      b2BodyDef bd;
      bd.type = b2_dynamicBody;
      bd.position = y;
      b2Body* body = s.addBody(bd);
      s.addFixture(body, shape, 5.0f);
      y += deltaY;
    }
    x += deltaX;
  }
}

inline Scene* scene_build(const std::string& name, int size, int seed) {
  Scene* s = new Scene();
  s->kind = name;
  s->world = new b2World(b2Vec2(0.0f, -10.0f));
  s->world->SetContinuousPhysics(false);  // as both reference benchmark mains do (single.cpp:44)
  if (name == "pyramid" || name == "many_pyramids") {
    int rows = name == "pyramid" ? (size > 0 ? size : 20) : 20;
    int copies = name == "pyramid" ? 1 : (size > 0 ? size : 100);
    {
      b2BodyDef bd;
      b2Body* ground = s->addBody(bd);
      b2EdgeShape shape;
      shape.SetTwoSided(b2Vec2(-40.0f, 0.0f), b2Vec2(40.0f + 30.0f * (float)(copies - 1), 0.0f));
      s->addFixture(ground, shape, 0.0f);
    }
    for (int k = 0; k < copies; ++k) scene_add_pyramid(*s, rows, 30.0f * (float)k);
  } else if (name == "mixed" || name == "mixed_linked") {
    int n = size > 0 ? size : 1000;
    SceneLCG rng((uint32_t)(seed > 0 ? seed : 12345));
    int cols = (int)std::ceil(1.5 * std::sqrt((double)n));
    float pitch = 1.1f;
    float width = pitch * (float)cols;
    int rowsNeeded = (n + cols - 1) / cols;
    float height = pitch * (float)rowsNeeded + 10.0f;
    {
      b2BodyDef bd;
      b2Body* container = s->addBody(bd);
      b2PolygonShape floor, wallL, wallR;
      floor.SetAsBox(0.5f * width + 2.0f, 1.0f, b2Vec2(0.5f * width, -1.0f), 0.0f);
      wallL.SetAsBox(1.0f, 0.5f * height + 1.0f, b2Vec2(-1.5f, 0.5f * height), 0.0f);
      wallR.SetAsBox(1.0f, 0.5f * height + 1.0f, b2Vec2(width + 1.5f, 0.5f * height), 0.0f);
      s->addFixture(container, floor, 0.0f);
      s->addFixture(container, wallL, 0.0f);
      s->addFixture(container, wallR, 0.0f);
    }
    for (int i = 0; i < n; ++i) {
      int cx = i % cols, cy = i / cols;
      b2BodyDef bd;
      bd.type = b2_dynamicBody;
      bd.position.Set(0.55f + pitch * (float)cx, 0.6f + pitch * (float)cy);
      bd.angle = rng.range(0.0f, 2.0f * b2_pi);
      b2Body* body = s->addBody(bd);
      b2FixtureDef fd;
      fd.density = 1.0f;
      fd.friction = 0.2f;
      if (i & 1) {
        b2CircleShape circle;
        circle.m_radius = rng.range(0.25f, 0.5f);
        fd.shape = &circle;
        s->addFixture(body, fd);
      } else {
        int nv = 3 + (int)(rng.next() * 6.0f);
        if (nv > 8) nv = 8;
        float rx = rng.range(0.3f, 0.5f), ry = rng.range(0.3f, 0.5f);
        b2Vec2 pts[8];
        for (int k = 0; k < nv; ++k) {
          float ang = 2.0f * b2_pi * (float)k / (float)nv;
          pts[k].Set(rx * cosf(ang), ry * sinf(ang));
        }
        b2PolygonShape poly;
        poly.Set(pts, nv);
        fd.shape = &poly;
        s->addFixture(body, fd);
      }
    }
    if (name == "mixed_linked") {
      // every 16th body is hinged to its right-hand grid neighbour (collideConnected = false): joints inside
      // an oversize island
      for (int i = 0; i + 1 < n; i += 16) {
        if ((i % cols) + 1 >= cols) continue;
        b2Body* a = s->bodies[1 + i];
        b2Body* b = s->bodies[2 + i];
        b2RevoluteJointDef jd;
        jd.Initialize(a, b, 0.5f * (a->GetPosition() + b->GetPosition()));
        s->world->CreateJoint(&jd);
      }
    }
  } else if (name == "tumbler") {
    s->spawnTarget = size > 0 ? size : 500;
    if (seed > 0) {
      s->spawnJitter = true;
      s->spawnRng = SceneLCG(0x9e3779b9u * (uint32_t)seed + 12345u);
    }
    b2Body* ground;
    {
      b2BodyDef bd;
      ground = s->addBody(bd);
    }
    {
      b2BodyDef bd;
      bd.type = b2_dynamicBody;
      bd.position.Set(0.0f, 10.0f);
      b2Body* body = s->addBody(bd);
      b2PolygonShape shape;
      shape.SetAsBox(0.5f, 10.0f, b2Vec2(10.0f, 0.0f), 0.0);
      s->addFixture(body, shape, 5.0f);
      shape.SetAsBox(0.5f, 10.0f, b2Vec2(-10.0f, 0.0f), 0.0);
      s->addFixture(body, shape, 5.0f);
      shape.SetAsBox(10.0f, 0.5f, b2Vec2(0.0f, 10.0f), 0.0);
      s->addFixture(body, shape, 5.0f);
      shape.SetAsBox(10.0f, 0.5f, b2Vec2(0.0f, -10.0f), 0.0);
      s->addFixture(body, shape, 5.0f);
      b2RevoluteJointDef jd;
      jd.bodyA = ground;
      jd.bodyB = body;
      jd.localAnchorA.Set(0.0f, 10.0f);
      jd.localAnchorB.Set(0.0f, 0.0f);
      jd.referenceAngle = 0.0f;
      jd.motorSpeed = 0.05f * b2_pi;
      jd.maxMotorTorque = 1e8f;
      jd.enableMotor = true;
      s->world->CreateJoint(&jd);
    }
  } else if (name == "pendulum" || name == "pendulum_limit" || name == "pendulum_motor") {
    // revolute joint checks (cf. unit-test/joint_test.cpp:27-106): `size` boxes, each hinged to the
    // static ground 2 m to the side of its centre, no contacts between them (they are 10 m apart)
    int n = size > 0 ? size : 3;
    b2BodyDef gd;
    b2Body* ground = s->addBody(gd);
    for (int i = 0; i < n; ++i) {
      b2BodyDef bd;
      bd.type = b2_dynamicBody;
      bd.position.Set(10.0f * (float)i + 2.0f, 5.0f);
      bd.angularDamping = 0.05f * (float)i;
      b2Body* body = s->addBody(bd);
      b2PolygonShape box;
      box.SetAsBox(0.5f + 0.1f * (float)i, 0.25f);
      s->addFixture(body, box, 1.0f + (float)i);
      b2RevoluteJointDef jd;
      jd.Initialize(ground, body, b2Vec2(10.0f * (float)i, 5.0f));
      if (name == "pendulum_limit") {
        jd.enableLimit = true;
        jd.lowerAngle = -0.4f;
        jd.upperAngle = 0.3f;
      }
      if (name == "pendulum_motor") {
        jd.enableMotor = true;
        jd.motorSpeed = 1.0f;
        jd.maxMotorTorque = 40.0f;
      }
      s->world->CreateJoint(&jd);
    }
  } else if (name == "chain" || name == "chain_collide") {
    // a chain of overlapping links hinged end to end (cf. testbed/tests/chain.cpp:31-66), swinging
    // down onto a ground edge: joined neighbours overlap but must not collide unless
    // collideConnected is set (b2Body::ShouldCollide, b2_body.cpp:396-419); non-neighbours do
    int n = size > 0 ? size : 12;
    b2BodyDef gd;
    b2Body* ground = s->addBody(gd);
    b2EdgeShape edge;
    edge.SetTwoSided(b2Vec2(-40.0f, 0.0f), b2Vec2(40.0f, 0.0f));
    s->addFixture(ground, edge, 0.0f);
    b2PolygonShape link;
    link.SetAsBox(0.6f, 0.125f);
    b2FixtureDef fd;
    fd.shape = &link;
    fd.density = 20.0f;
    fd.friction = 0.2f;
    b2Body* prev = ground;
    for (int i = 0; i < n; ++i) {
      b2BodyDef bd;
      bd.type = b2_dynamicBody;
      bd.position.Set(0.5f + (float)i, 6.0f);
      b2Body* body = s->addBody(bd);
      s->addFixture(body, fd);
      b2RevoluteJointDef jd;
      jd.Initialize(prev, body, b2Vec2((float)i, 6.0f));
      jd.collideConnected = (name == "chain_collide");
      s->world->CreateJoint(&jd);
      prev = body;
    }
  } else if (name == "welds") {
    // weld joints (b2_weld_joint.cpp:62-305): cantilever beams of welded boxes sticking out of a wall
    // (rigid and with a rotational spring), a welded compound tumbling onto them
    int n = size > 0 ? size : 6;
    b2BodyDef gd;
    b2Body* ground = s->addBody(gd);
    b2EdgeShape edge;
    edge.SetTwoSided(b2Vec2(-40.0f, 0.0f), b2Vec2(40.0f, 0.0f));
    s->addFixture(ground, edge, 0.0f);
    b2PolygonShape plank;
    plank.SetAsBox(0.5f, 0.125f);
    for (int beam = 0; beam < 2; ++beam) {
      b2Body* prev = ground;
      float y = 3.0f + 3.0f * (float)beam;
      for (int i = 0; i < n; ++i) {
        b2BodyDef bd;
        bd.type = b2_dynamicBody;
        bd.position.Set(-10.0f + 0.5f + 1.0f * (float)i, y);
        b2Body* body = s->addBody(bd);
        s->addFixture(body, plank, 20.0f);
        b2WeldJointDef jd;
        jd.Initialize(prev, body, b2Vec2(-10.0f + 1.0f * (float)i, y));
        if (beam == 1) b2AngularStiffness(jd.stiffness, jd.damping, 5.0f, 0.7f, prev, body);
        s->world->CreateJoint(&jd);
        prev = body;
      }
    }
    {  // an L-shaped compound of three welded boxes dropped onto the lower beam
      b2PolygonShape box;
      box.SetAsBox(0.4f, 0.4f);
      b2Body* parts[3];
      const float px[3] = {-7.0f, -6.2f, -6.2f}, py[3] = {9.0f, 9.0f, 9.8f};
      for (int k = 0; k < 3; ++k) {
        b2BodyDef bd;
        bd.type = b2_dynamicBody;
        bd.position.Set(px[k], py[k]);
        parts[k] = s->addBody(bd);
        s->addFixture(parts[k], box, 2.0f);
      }
      b2WeldJointDef jd;
      jd.Initialize(parts[0], parts[1], b2Vec2(-6.6f, 9.0f));
      s->world->CreateJoint(&jd);
      jd.Initialize(parts[1], parts[2], b2Vec2(-6.2f, 9.4f));
      s->world->CreateJoint(&jd);
    }
  } else if (name == "cars") {
    // wheel joints (b2_wheel_joint.cpp:87-446): `size` cars, each a chassis on two sprung wheels (rear wheel
    // driven by the joint motor, front wheel free), suspension travel limited on every second car and
    // rigid (stiffness 0) on every third; they drive over ramps and push loose boxes along a ground edge
    int n = size > 0 ? size : 4;
    b2BodyDef gd;
    b2Body* ground = s->addBody(gd);
    b2EdgeShape edge;
    edge.SetTwoSided(b2Vec2(-80.0f, 0.0f), b2Vec2(80.0f, 0.0f));
    s->addFixture(ground, edge, 0.0f);
    b2PolygonShape ramp;
    for (int i = 0; i < n; ++i) {
      b2Vec2 tri[3] = {b2Vec2(0.0f, 0.0f), b2Vec2(3.0f, 0.0f), b2Vec2(3.0f, 0.4f + 0.1f * (float)(i % 3))};
      for (int k = 0; k < 3; ++k) tri[k].x += -30.0f + 12.0f * (float)i + 6.0f;
      ramp.Set(tri, 3);
      s->addFixture(ground, ramp, 0.0f);
    }
    b2PolygonShape chassis;
    b2Vec2 hull[6] = {b2Vec2(-1.5f, -0.5f), b2Vec2(1.5f, -0.5f), b2Vec2(1.5f, 0.0f),
                      b2Vec2(0.0f, 0.9f),   b2Vec2(-1.15f, 0.9f), b2Vec2(-1.5f, 0.2f)};
    chassis.Set(hull, 6);
    b2CircleShape tyre;
    tyre.m_radius = 0.4f;
    b2PolygonShape crate;
    crate.SetAsBox(0.25f, 0.25f);
    for (int i = 0; i < n; ++i) {
      const float x = -30.0f + 12.0f * (float)i;
      b2BodyDef bd;
      bd.type = b2_dynamicBody;
      bd.position.Set(x, 1.0f);
      b2Body* car = s->addBody(bd);
      s->addFixture(car, chassis, 1.0f);
      b2FixtureDef fd;
      fd.shape = &tyre;
      fd.density = 1.0f;
      fd.friction = 0.9f;
      bd.position.Set(x - 1.0f, 0.35f);
      b2Body* rear = s->addBody(bd);
      s->addFixture(rear, fd);
      bd.position.Set(x + 1.0f, 0.4f);
      b2Body* front = s->addBody(bd);
      s->addFixture(front, fd);
      b2WheelJointDef jd;
      const b2Vec2 axis(0.0f, 1.0f);
      const float hertz = 4.0f, ratio = 0.7f;
      jd.Initialize(car, rear, rear->GetPosition(), axis);
      jd.motorSpeed = -10.0f - 5.0f * (float)(i % 2);
      jd.maxMotorTorque = 20.0f;
      jd.enableMotor = true;
      if (i % 3 != 2) b2LinearStiffness(jd.stiffness, jd.damping, hertz, ratio, car, rear);
      jd.lowerTranslation = -0.25f;
      jd.upperTranslation = 0.25f;
      jd.enableLimit = (i % 2) == 1;
      s->world->CreateJoint(&jd);
      jd.Initialize(car, front, front->GetPosition(), axis);
      jd.motorSpeed = 0.0f;
      jd.maxMotorTorque = 10.0f;
      jd.enableMotor = false;
      if (i % 3 != 2) b2LinearStiffness(jd.stiffness, jd.damping, hertz, ratio, car, front);
      s->world->CreateJoint(&jd);
      for (int k = 0; k < 3; ++k) {  // loose crates ahead of the car
        b2BodyDef cd;
        cd.type = b2_dynamicBody;
        cd.position.Set(x + 3.0f + 0.6f * (float)k, 0.26f + 0.52f * (float)(k % 2));
        // never asleep: a sleeping crate hit by a car is woken in the middle of b2ContactManager::Collide,
        // and whether its ground contact is re-evaluated in that same pass depends on the contact ring
        // order (ring-order corner); the joint parity test stays clear of that
        cd.allowSleep = seed == 0 ? false : true;
        s->addFixture(s->addBody(cd), crate, 0.5f);
      }
    }
  } else if (name == "mice") {
    // mouse joints (b2_mouse_joint.cpp:77-160): boxes and balls resting on the ground are dragged to targets
    // above and beside them — anchored off-centre, with spring rates from 1 to 5 Hz, force budgets from less
    // than the weight to a thousand times it — through each other and a row of loose boxes
    int n = size > 0 ? size : 8;
    b2BodyDef gd;
    b2Body* floorBody = s->addBody(gd);
    b2EdgeShape edge;
    edge.SetTwoSided(b2Vec2(-60.0f, 0.0f), b2Vec2(60.0f, 0.0f));
    s->addFixture(floorBody, edge, 0.0f);
    // the joints hang on a second, shapeless static body: a joint without collideConnected switches off the
    // contacts between its two bodies, and the dragged bodies must keep colliding with the floor
    b2Body* ground = s->addBody(gd);
    b2PolygonShape box;
    box.SetAsBox(0.4f, 0.4f);
    b2CircleShape ball;
    ball.m_radius = 0.35f;
    for (int i = 0; i < n; ++i) {
      float x = -12.0f + 3.0f * (float)i;
      b2BodyDef bd;
      bd.type = b2_dynamicBody;
      bd.position.Set(x, 0.41f);
      bd.allowSleep = seed == 0 ? false : true;
      b2Body* body = s->addBody(bd);
      if (i % 2) s->addFixture(body, ball, 2.0f);
      else s->addFixture(body, box, 2.0f);
      b2MouseJointDef jd;
      jd.bodyA = ground;
      jd.bodyB = body;
      jd.target.Set(x + 0.3f, 0.6f);                       // grabbed off-centre
      jd.maxForce = (i % 4 == 3 ? 0.5f : 1000.0f) * body->GetMass() * 10.0f;
      b2LinearStiffness(jd.stiffness, jd.damping, 1.0f + (float)(i % 5), 0.7f, ground, body);
      b2MouseJoint* mj = static_cast<b2MouseJoint*>(s->world->CreateJoint(&jd));
      mj->SetTarget(b2Vec2(x + 2.0f - 1.0f * (float)(i % 3), 3.0f + 0.5f * (float)(i % 4)));
      b2BodyDef cd;
      cd.type = b2_dynamicBody;
      cd.position.Set(x + 1.5f, 0.41f);
      cd.allowSleep = seed == 0 ? false : true;             // see "cars": keeps the parity run off the ring-order corner
      s->addFixture(s->addBody(cd), box, 1.0f);             // a loose box in the way
    }
  } else if (name == "drags") {
    // friction joints (b2_friction_joint.cpp:65-181): boxes tied to the ground by a force / torque budget —
    // below their weight (they sink slowly, spinning down), above it (they hang), off-centre anchors; motor
    // joints (b2_motor_joint.cpp:70-208): platforms pulled to a linear + angular offset with bounded force,
    // soft and stiff correction factors, loose boxes dropped onto them, one platform riding on another
    int n = size > 0 ? size : 6;
    b2BodyDef gd;
    b2Body* ground = s->addBody(gd);
    b2EdgeShape edge;
    edge.SetTwoSided(b2Vec2(-60.0f, 0.0f), b2Vec2(60.0f, 0.0f));
    s->addFixture(ground, edge, 0.0f);
    b2PolygonShape box;
    box.SetAsBox(0.5f, 0.5f);
    b2PolygonShape plate;
    plate.SetAsBox(1.2f, 0.15f);
    b2CircleShape ball;
    ball.m_radius = 0.3f;
    for (int i = 0; i < n; ++i) {  // braked boxes, mass 1
      float x = -20.0f + 2.5f * (float)i;
      b2BodyDef bd;
      bd.type = b2_dynamicBody;
      bd.position.Set(x, 6.0f + 0.5f * (float)(i % 3));
      bd.angularVelocity = 4.0f - 1.5f * (float)i;
      bd.linearVelocity.Set(1.0f - 0.5f * (float)(i % 4), 0.0f);
      b2Body* body = s->addBody(bd);
      s->addFixture(body, box, 1.0f);
      b2FrictionJointDef jd;
      jd.Initialize(ground, body, b2Vec2(x + 0.2f * (float)(i % 2), bd.position.y));
      jd.maxForce = 4.0f + 2.5f * (float)i;      // weight is 10: the first ones sink, the last ones hang
      jd.maxTorque = 0.5f * (float)(i + 1);
      jd.collideConnected = true;
      s->world->CreateJoint(&jd);
    }
    b2Body* prevPlatform = nullptr;
    for (int i = 0; i < n; ++i) {  // driven platforms
      float x = 0.0f + 4.0f * (float)i;
      b2BodyDef bd;
      bd.type = b2_dynamicBody;
      bd.position.Set(x, 2.0f);
      b2Body* platform = s->addBody(bd);
      s->addFixture(platform, plate, 2.0f);
      b2MotorJointDef jd;
      b2Body* base = (i % 3 == 2 && prevPlatform) ? prevPlatform : ground;
      jd.Initialize(base, platform);
      jd.linearOffset += b2Vec2(0.5f, 1.5f);            // target: up and to the right of where it starts
      jd.angularOffset = 0.15f * (float)(i % 3) - 0.15f;
      jd.maxForce = 150.0f + 50.0f * (float)i;
      jd.maxTorque = 100.0f;
      jd.correctionFactor = (i % 2) ? 0.9f : 0.3f;
      jd.collideConnected = true;
      s->world->CreateJoint(&jd);
      b2BodyDef cd;
      cd.type = b2_dynamicBody;
      cd.position.Set(x - 0.4f, 3.2f);
      s->addFixture(s->addBody(cd), box, 1.0f);
      cd.position.Set(x + 0.6f, 3.0f);
      s->addFixture(s->addBody(cd), ball, 1.0f);
      prevPlatform = platform;
    }
  } else if (name == "sliders") {
    // prismatic joints (b2_prismatic_joint.cpp:114-451) in their regimes: a motorised piston pushing a
    // pile of boxes along the ground against a limit, vertical lifts between a lower and an upper
    // limit carrying loose boxes, tilted free rails (no motor, no limit) and a carriage chained to a
    // carriage (dynamic-dynamic), all colliding with each other and a ground edge
    int n = size > 0 ? size : 6;
    b2BodyDef gd;
    b2Body* ground = s->addBody(gd);
    b2EdgeShape edge;
    edge.SetTwoSided(b2Vec2(-60.0f, 0.0f), b2Vec2(60.0f, 0.0f));
    s->addFixture(ground, edge, 0.0f);
    b2PolygonShape box;
    box.SetAsBox(0.4f, 0.4f);
    b2PolygonShape plate;
    plate.SetAsBox(1.0f, 0.2f);
    b2CircleShape ball;
    ball.m_radius = 0.3f;
    {  // the piston: a tall block driven along +x by a motor, limit [0, 6]
      b2PolygonShape ram;
      ram.SetAsBox(0.3f, 1.0f);
      b2BodyDef bd;
      bd.type = b2_dynamicBody;
      bd.position.Set(-20.0f, 1.05f);
      b2Body* body = s->addBody(bd);
      s->addFixture(body, ram, 5.0f);
      b2PrismaticJointDef jd;
      jd.Initialize(ground, body, b2Vec2(-20.0f, 1.05f), b2Vec2(1.0f, 0.0f));
      jd.enableMotor = true;
      jd.motorSpeed = 2.0f;
      jd.maxMotorForce = 400.0f;
      jd.enableLimit = true;
      jd.lowerTranslation = 0.0f;
      jd.upperTranslation = 6.0f;
      s->world->CreateJoint(&jd);
      for (int i = 0; i < n; ++i) {
        b2BodyDef cd;
        cd.type = b2_dynamicBody;
        cd.position.Set(-18.5f + 0.9f * (float)(i % 3), 0.45f + 0.85f * (float)(i / 3));
        s->addFixture(s->addBody(cd), box, 1.0f);
      }
    }
    for (int i = 0; i < n; ++i) {  // lifts: plate on a vertical rail, motor pushing up against gravity, limits
      float x = -10.0f + 3.0f * (float)i;
      b2BodyDef bd;
      bd.type = b2_dynamicBody;
      bd.position.Set(x, 2.0f);
      b2Body* lift = s->addBody(bd);
      s->addFixture(lift, plate, 2.0f);
      b2PrismaticJointDef jd;
      jd.Initialize(ground, lift, b2Vec2(x, 2.0f), b2Vec2(0.0f, 1.0f));
      jd.enableLimit = true;
      jd.lowerTranslation = -1.0f;
      jd.upperTranslation = 0.5f + 0.25f * (float)i;
      jd.enableMotor = (i % 2) == 0;
      jd.motorSpeed = 1.0f;
      jd.maxMotorForce = 60.0f + 30.0f * (float)i;
      s->world->CreateJoint(&jd);
      b2BodyDef cd;
      cd.type = b2_dynamicBody;
      cd.position.Set(x - 0.3f, 2.65f);
      s->addFixture(s->addBody(cd), box, 1.0f);
      cd.position.Set(x + 0.45f, 2.6f);
      s->addFixture(s->addBody(cd), ball, 1.0f);
    }
    for (int i = 0; i < n; ++i) {  // tilted free rails with a second carriage hanging off the first
      float x = -10.0f + 3.0f * (float)i, y = 8.0f;
      b2Vec2 axis(0.8f, i % 2 ? -0.6f : 0.6f);
      b2BodyDef bd;
      bd.type = b2_dynamicBody;
      bd.position.Set(x, y);
      bd.angle = 0.1f * (float)i;
      b2Body* first = s->addBody(bd);
      s->addFixture(first, box, 1.0f);
      b2PrismaticJointDef jd;
      jd.Initialize(ground, first, b2Vec2(x, y), axis);
      if (i % 3 == 0) {  // lower == upper: the locked-slider branch of the position solver
        jd.enableLimit = true;
        jd.lowerTranslation = jd.upperTranslation = 0.0f;
      }
      s->world->CreateJoint(&jd);
      bd.position.Set(x + 0.2f, y - 1.2f);
      bd.angle = 0.0f;
      b2Body* second = s->addBody(bd);
      s->addFixture(second, ball, 3.0f);
      jd = b2PrismaticJointDef();
      jd.Initialize(first, second, b2Vec2(x, y - 0.6f), b2Vec2(0.0f, 1.0f));
      jd.enableLimit = true;
      jd.lowerTranslation = -0.5f;
      jd.upperTranslation = 0.3f;
      jd.collideConnected = (i % 2) == 0;
      s->world->CreateJoint(&jd);
    }
  } else if (name == "springs") {
    // distance joints in their three regimes (b2_distance_joint.cpp:76-303): rigid rods (a hanging
    // net of boxes), soft springs (stiffness / damping from b2LinearStiffness) and slack ropes with
    // min / max length limits, all bumping into each other and a ground edge
    int n = size > 0 ? size : 8;
    b2BodyDef gd;
    b2Body* ground = s->addBody(gd);
    b2EdgeShape edge;
    edge.SetTwoSided(b2Vec2(-40.0f, 0.0f), b2Vec2(40.0f, 0.0f));
    s->addFixture(ground, edge, 0.0f);
    b2PolygonShape box;
    box.SetAsBox(0.3f, 0.3f);
    b2CircleShape ball;
    ball.m_radius = 0.35f;
    b2Body* prevRow[64];
    for (int i = 0; i < n && i < 64; ++i) {
      // column i: ceiling anchor -> rigid rod -> box -> spring -> ball -> limited rope -> box
      float x = -0.9f * (float)n + 1.8f * (float)i;
      b2BodyDef bd;
      bd.type = b2_dynamicBody;
      bd.position.Set(x, 4.3f);
      b2Body* a = s->addBody(bd);
      s->addFixture(a, box, 1.0f);
      bd.position.Set(x + 0.4f, 2.3f);
      b2Body* b = s->addBody(bd);
      s->addFixture(b, ball, 0.8f);
      bd.position.Set(x - 0.3f, 0.5f);
      bd.angle = 0.3f * (float)i;
      b2Body* c = s->addBody(bd);
      s->addFixture(c, box, 1.5f);
      bd.angle = 0.0f;
      b2DistanceJointDef rod;
      rod.Initialize(ground, a, b2Vec2(x, 6.3f), b2Vec2(x, 4.6f));
      s->world->CreateJoint(&rod);
      b2DistanceJointDef spring;
      spring.Initialize(a, b, b2Vec2(x, 4.0f), b2Vec2(x + 0.4f, 2.3f));
      b2LinearStiffness(spring.stiffness, spring.damping, 2.0f + 0.25f * (float)i, 0.1f + 0.05f * (float)i, a, b);
      spring.minLength = 0.5f * spring.length;
      spring.maxLength = 1.6f * spring.length;
      s->world->CreateJoint(&spring);
      b2DistanceJointDef rope;
      rope.Initialize(b, c, b2Vec2(x + 0.4f, 2.3f), b2Vec2(x - 0.3f, 0.5f));
      rope.minLength = 0.25f * rope.length;
      rope.maxLength = 1.1f * rope.length;   // slack: only the limits act
      s->world->CreateJoint(&rope);
      if (i > 0) {  // rigid cross-links between neighbouring columns: one big jointed island
        b2DistanceJointDef link;
        link.Initialize(prevRow[i - 1], a, prevRow[i - 1]->GetPosition(), a->GetPosition());
        link.collideConnected = true;
        s->world->CreateJoint(&link);
      }
      prevRow[i] = a;
    }
  } else if (name == "sensors") {
    // sensor fixtures (static regions, a rotating paddle and probes riding on bodies) crossed by a
    // rain of circles and polygons: b2Contact::Update takes `touching` of a sensor contact from
    // b2TestOverlap = GJK distance (b2_contact.cpp:145-151, b2_collision.cpp:239-258)
    int n = size > 0 ? size : 300;
    SceneLCG rng((uint32_t)(seed > 0 ? seed : 4242));
    b2BodyDef gd;
    b2Body* ground = s->addBody(gd);
    b2EdgeShape edge;
    edge.SetTwoSided(b2Vec2(-40.0f, 0.0f), b2Vec2(40.0f, 0.0f));
    s->addFixture(ground, edge, 0.0f);
    {
      b2FixtureDef sd;
      sd.isSensor = true;
      b2PolygonShape zone;
      zone.SetAsBox(4.0f, 0.75f, b2Vec2(0.0f, 6.0f), 0.3f);
      sd.shape = &zone;
      s->addFixture(ground, sd);
      b2CircleShape disc;
      disc.m_p.Set(-7.0f, 4.0f);
      disc.m_radius = 1.5f;
      sd.shape = &disc;
      s->addFixture(ground, sd);
      b2EdgeShape wire;
      wire.SetTwoSided(b2Vec2(4.0f, 2.0f), b2Vec2(10.0f, 5.0f));
      sd.shape = &wire;
      s->addFixture(ground, sd);
      b2PolygonShape tri;
      b2Vec2 pts[3] = {b2Vec2(-3.0f, 1.0f), b2Vec2(-1.0f, 1.2f), b2Vec2(-2.2f, 2.6f)};
      tri.Set(pts, 3);
      sd.shape = &tri;
      s->addFixture(ground, sd);
    }
    {
      // a kinematic paddle that is one long thin sensor, sweeping through the rain
      b2BodyDef pd;
      pd.type = b2_kinematicBody;
      pd.position.Set(6.0f, 8.0f);
      pd.angularVelocity = 1.3f;
      b2Body* paddle = s->addBody(pd);
      b2PolygonShape blade;
      blade.SetAsBox(3.0f, 0.1f);
      b2FixtureDef sd;
      sd.isSensor = true;
      sd.shape = &blade;
      s->addFixture(paddle, sd);
    }
    int cols = 24;
    for (int i = 0; i < n; ++i) {
      b2BodyDef bd;
      bd.type = b2_dynamicBody;
      bd.position.Set(-12.0f + 1.05f * (float)(i % cols) + rng.range(-0.2f, 0.2f), 9.0f + 1.1f * (float)(i / cols));
      bd.angle = rng.range(0.0f, 2.0f * b2_pi);
      bd.angularVelocity = rng.range(-3.0f, 3.0f);
      b2Body* body = s->addBody(bd);
      b2FixtureDef fd;
      fd.density = 1.0f;
      fd.friction = 0.3f;
      if (i % 3 == 0) {
        b2CircleShape c;
        c.m_radius = rng.range(0.15f, 0.4f);
        fd.shape = &c;
        s->addFixture(body, fd);
      } else {
        int nv = 3 + (int)(rng.next() * 5.0f);
        float rx = rng.range(0.2f, 0.45f), ry = rng.range(0.2f, 0.45f);
        b2Vec2 pts[8];
        for (int k = 0; k < nv; ++k) {
          float ang = 2.0f * b2_pi * (float)k / (float)nv;
          pts[k].Set(rx * cosf(ang), ry * sinf(ang));
        }
        b2PolygonShape poly;
        poly.Set(pts, nv);
        fd.shape = &poly;
        s->addFixture(body, fd);
      }
      if (i % 7 == 0) {
        // a probe: a sensor disc riding on the body, overlapping its neighbours' solid shapes
        b2CircleShape probe;
        probe.m_radius = 0.8f;
        b2FixtureDef sd;
        sd.isSensor = true;
        sd.shape = &probe;
        s->addFixture(body, sd);
      }
    }
  } else if (name == "filters") {
    // b2ContactFilter::ShouldCollide (b2_world_callbacks.cpp:28-40): category / mask bits and group indices.
    // Boxes and balls fall through each other or collide depending on their filter:
    //   group +1 : always collide with each other (even though their masks would say no)
    //   group -2 : never collide with each other (they pile up INSIDE one another on the floor)
    //   group  0 : category/mask rule: "red" (0x2) sees floor + blue, "blue" (0x4) sees floor + red + blue,
    //              "ghost" (0x8, mask 0x1) sees only the floor, and nobody's mask names the ghosts
    // plus a shelf (category 0x10) that only the blue ones rest on.
    int n = size > 0 ? size : 120;
    {
      b2BodyDef bd;
      b2Body* ground = s->addBody(bd);
      b2PolygonShape floor;
      floor.SetAsBox(30.0f, 1.0f, b2Vec2(0.0f, -1.0f), 0.0f);
      b2FixtureDef fd;
      fd.shape = &floor;
      fd.filter.categoryBits = 0x0001;
      fd.filter.maskBits = 0xFFFF;
      s->addFixture(ground, fd);
      b2PolygonShape shelf;
      shelf.SetAsBox(12.0f, 0.25f, b2Vec2(0.0f, 6.0f), 0.0f);
      fd.shape = &shelf;
      fd.filter.categoryBits = 0x0010;
      fd.filter.maskBits = 0x0004;
      s->addFixture(ground, fd);
    }
    SceneLCG rng((uint32_t)(seed > 0 ? seed : 99));
    int cols = 12;
    for (int i = 0; i < n; ++i) {
      b2BodyDef bd;
      bd.type = b2_dynamicBody;
      bd.position.Set(-10.0f + 1.7f * (float)(i % cols) + rng.range(-0.3f, 0.3f), 8.0f + 1.4f * (float)(i / cols));
      bd.angle = rng.range(-0.5f, 0.5f);
      b2Body* body = s->addBody(bd);
      b2FixtureDef fd;
      fd.density = 1.0f;
      fd.friction = 0.3f;
      b2PolygonShape box;
      b2CircleShape ball;
      if (i & 1) {
        ball.m_radius = rng.range(0.3f, 0.55f);
        fd.shape = &ball;
      } else {
        box.SetAsBox(rng.range(0.3f, 0.6f), rng.range(0.3f, 0.6f));
        fd.shape = &box;
      }
      switch (i % 5) {
        case 0: fd.filter.categoryBits = 0x0002; fd.filter.maskBits = 0x0001 | 0x0004; break;            // red
        case 1: fd.filter.categoryBits = 0x0004; fd.filter.maskBits = 0x0001 | 0x0002 | 0x0004 | 0x0010; break;  // blue
        case 2: fd.filter.categoryBits = 0x0008; fd.filter.maskBits = 0x0001; break;                     // ghost
        case 3: fd.filter.categoryBits = 0x0020; fd.filter.maskBits = 0x0001; fd.filter.groupIndex = 1; break;   // group +1
        default: fd.filter.categoryBits = 0x0040; fd.filter.maskBits = 0xFFFF; fd.filter.groupIndex = -2; break;  // group -2
      }
      s->addFixture(body, fd);
    }
  } else if (name == "hello") {
    s->velocityIterations = 6;
    s->positionIterations = 2;
    b2BodyDef groundBodyDef;
    groundBodyDef.position.Set(0.0f, -10.0f);
    b2Body* groundBody = s->addBody(groundBodyDef);
    b2PolygonShape groundBox;
    groundBox.SetAsBox(50.0f, 10.0f);
    s->addFixture(groundBody, groundBox, 0.0f);
    b2BodyDef bodyDef;
    bodyDef.type = b2_dynamicBody;
    bodyDef.position.Set(0.0f, 4.0f);
    b2Body* body = s->addBody(bodyDef);
    b2PolygonShape dynamicBox;
    dynamicBox.SetAsBox(1.0f, 1.0f);
    b2FixtureDef fixtureDef;
    fixtureDef.shape = &dynamicBox;
    fixtureDef.density = 1.0f;
    fixtureDef.friction = 0.3f;
    s->addFixture(body, fixtureDef);
  } else if (name == "falling_squares" || name == "falling_circles") {
    int n = size > 0 ? size : 300;
    {
      b2BodyDef bd;
      b2Body* ground = s->addBody(bd);
      b2PolygonShape shape;
      shape.SetAsBox(200.0f, 10.0f, b2Vec2(0.0f, -10.0f), 0.0f);
      s->addFixture(ground, shape, 0.0f);
    }
    SceneLCG rng((uint32_t)(seed > 0 ? seed : 7));
    int cols = 30;
    for (int i = 0; i < n; ++i) {
      b2BodyDef bd;
      bd.type = b2_dynamicBody;
      bd.position.Set(-18.0f + 1.25f * (float)(i % cols) + rng.range(-0.1f, 0.1f), 1.0f + 1.25f * (float)(i / cols));
      b2Body* body = s->addBody(bd);
      if (name == "falling_squares") {
        b2PolygonShape shape;
        shape.SetAsBox(0.5f, 0.5f);
        s->addFixture(body, shape, 1.0f);
      } else {
        b2CircleShape shape;
        shape.m_radius = 0.5f;
        s->addFixture(body, shape, 1.0f);
      }
    }
  } else {
    delete s;
    return nullptr;
  }
  return s;
}

// ---- export helpers (public shape members only) ---------------------------------------------
inline int scene_shape_quad_count(const b2Shape* sh) {
  switch (sh->GetType()) {
    case b2Shape::e_circle: return 1;
    case b2Shape::e_edge: return 3;
    case b2Shape::e_polygon: return 1 + static_cast<const b2PolygonShape*>(sh)->m_count;
    default: return 0;
  }
}
inline void scene_write_shape_quads(const b2Shape* sh, float* q) {
  switch (sh->GetType()) {
    case b2Shape::e_circle: {
      const b2CircleShape* c = static_cast<const b2CircleShape*>(sh);
      q[0] = c->m_p.x; q[1] = c->m_p.y; q[2] = c->m_radius; q[3] = 0.0f;
    } break;
    case b2Shape::e_edge: {
      const b2EdgeShape* e = static_cast<const b2EdgeShape*>(sh);
      q[0] = e->m_vertex1.x; q[1] = e->m_vertex1.y; q[2] = e->m_vertex2.x; q[3] = e->m_vertex2.y;
      q[4] = e->m_vertex0.x; q[5] = e->m_vertex0.y; q[6] = e->m_vertex3.x; q[7] = e->m_vertex3.y;
      q[8] = e->m_radius; q[9] = e->m_oneSided ? 1.0f : 0.0f; q[10] = 0.0f; q[11] = 0.0f;
    } break;
    case b2Shape::e_polygon: {
      const b2PolygonShape* p = static_cast<const b2PolygonShape*>(sh);
      q[0] = p->m_centroid.x; q[1] = p->m_centroid.y; q[2] = p->m_radius; q[3] = (float)p->m_count;
      for (int i = 0; i < p->m_count; ++i) {
        q[4 + 4 * i] = p->m_vertices[i].x; q[5 + 4 * i] = p->m_vertices[i].y;
        q[6 + 4 * i] = p->m_normals[i].x;  q[7 + 4 * i] = p->m_normals[i].y;
      }
    } break;
    default: break;
  }
}

#endif

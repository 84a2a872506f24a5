// -----------------------------------------------------------------------------------------------
// Third-party notice.  To stay source- and result-compatible with box2d-optimized, parts of this
// file restate declarations, inline math and creation-time algorithms of that library (itself a
// fork of Box2D).  Those parts are covered by the MIT License:
//   Copyright (c) 2019 Erin Catto, Copyright (c) 2020 Manolis Tsamis
// The full licence text and permission notice are in LICENSES/box2d-optimized-MIT.txt.
// -----------------------------------------------------------------------------------------------
// api_tests.cpp — acceptance tests of the public C++ API, compiled twice:
//   against the reference (-I/root/reference/include + its sources): proves the tests are right
//   against this repo's drop-in API (-Iinclude + libb2gpu_scenes.so): proves the drop-in
// The cases restate the reference's own unit tests (unit-test/hello_world.cpp:33-112,
// world_test.cpp:38-73, collision_test.cpp:28-81, math_test.cpp:27-54) without doctest.
#include <cstdio>
#include <cmath>
#include <cfloat>
#include "box2d/box2d.h"

static int g_failed = 0;
#define CHECK(cond)                                                        \
  do {                                                                     \
    if (!(cond)) {                                                         \
      printf("CHECK FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond);       \
      ++g_failed;                                                          \
    }                                                                      \
  } while (0)

static void hello_world() {
  b2Vec2 gravity(0.0f, -10.0f);
  b2World world(gravity);
  b2BodyDef groundBodyDef;
  groundBodyDef.position.Set(0.0f, -10.0f);
  b2Body* groundBody = world.CreateBody(&groundBodyDef);
  b2PolygonShape groundBox;
  groundBox.SetAsBox(50.0f, 10.0f);
  groundBody->CreateFixture(&groundBox, 0.0f);
  b2BodyDef bodyDef;
  bodyDef.type = b2_dynamicBody;
  bodyDef.position.Set(0.0f, 4.0f);
  b2Body* body = world.CreateBody(&bodyDef);
  b2PolygonShape dynamicBox;
  dynamicBox.SetAsBox(1.0f, 1.0f);
  b2FixtureDef fixtureDef;
  fixtureDef.shape = &dynamicBox;
  fixtureDef.density = 1.0f;
  fixtureDef.friction = 0.3f;
  body->CreateFixture(&fixtureDef);
  float timeStep = 1.0f / 60.0f;
  b2Vec2 position = body->GetPosition();
  float angle = body->GetAngle();
  for (int32 i = 0; i < 60; ++i) {
    world.Step(timeStep, 6, 2);
    position = body->GetPosition();
    angle = body->GetAngle();
  }
  printf("hello world: %4.3f %4.3f %4.3f\n", position.x, position.y, angle);
  CHECK(b2Abs(position.x) < 0.01f);
  CHECK(b2Abs(position.y - 1.01f) < 0.01f);
  CHECK(b2Abs(angle) < 0.01f);
}

static bool begin_contact = false;
static int end_contacts = 0, pre_solves = 0, post_solves = 0;
class MyContactListener : public b2ContactListener {
 public:
  void BeginContact(b2Contact* contact) override {
    begin_contact = true;
    CHECK(contact->GetFixtureA() != nullptr && contact->GetFixtureB() != nullptr);
    CHECK(contact->IsTouching());
  }
  void EndContact(b2Contact*) override { ++end_contacts; }
  void PreSolve(b2Contact* c, const b2Manifold* oldManifold) override {
    ++pre_solves;
    CHECK(oldManifold != nullptr);
    CHECK(c->GetManifold()->pointCount > 0);
  }
  void PostSolve(b2Contact*, const b2ContactImpulse* impulse) override {
    ++post_solves;
    CHECK(impulse->count > 0);
  }
};

static void begin_contact_test() {
  b2World world = b2World(b2Vec2(0.0f, -10.0f));
  MyContactListener listener;
  world.SetContactListener(&listener);
  b2CircleShape circle;
  circle.m_radius = 5.f;
  b2BodyDef bodyDef;
  bodyDef.type = b2_dynamicBody;
  b2Body* bodyA = world.CreateBody(&bodyDef);
  b2Body* bodyB = world.CreateBody(&bodyDef);
  bodyA->CreateFixture(&circle, 0.0f);
  bodyB->CreateFixture(&circle, 0.0f);
  bodyA->SetTransform(b2Vec2(0.f, 0.f), 0.f);
  bodyB->SetTransform(b2Vec2(100.f, 0.f), 0.f);
  const float timeStep = 1.f / 60.f;
  world.Step(timeStep, 6, 2);
  CHECK(world.GetContactListStart() == world.GetContactListEnd());
  CHECK(begin_contact == false);
  bodyB->SetTransform(b2Vec2(1.f, 0.f), 0.f);
  world.Step(timeStep, 6, 2);
  CHECK(world.GetContactListStart() != world.GetContactListEnd());
  CHECK(begin_contact == true);
  CHECK(pre_solves > 0);
  CHECK(post_solves > 0);
  CHECK(bodyA->GetContactCount() == 1);
  CHECK(bodyA->GetContact(0) == world.GetContactListStart());
  // separate again: the contact ends
  bodyB->SetTransform(b2Vec2(100.f, 0.f), 0.f);
  world.Step(timeStep, 6, 2);
  world.Step(timeStep, 6, 2);
  CHECK(end_contacts > 0);
  CHECK(world.GetContactCount() == 0);
}

static void polygon_mass_data() {
  const b2Vec2 center(100.0f, -50.0f);
  const float hx = 0.5f, hy = 1.5f;
  const float angle1 = 0.25f;
  b2PolygonShape polygon1;
  polygon1.SetAsBox(hx, hy, center, angle1);
  const float absTol = 2.0f * b2_epsilon;
  const float relTol = 2.0f * b2_epsilon;
  CHECK(b2Abs(polygon1.m_centroid.x - center.x) < absTol + relTol * b2Abs(center.x));
  CHECK(b2Abs(polygon1.m_centroid.y - center.y) < absTol + relTol * b2Abs(center.y));
  b2Vec2 vertices[4];
  vertices[0].Set(center.x - hx, center.y - hy);
  vertices[1].Set(center.x + hx, center.y - hy);
  vertices[2].Set(center.x - hx, center.y + hy);
  vertices[3].Set(center.x + hx, center.y + hy);
  b2PolygonShape polygon2;
  polygon2.Set(vertices, 4);
  CHECK(b2Abs(polygon2.m_centroid.x - center.x) < absTol + relTol * b2Abs(center.x));
  CHECK(b2Abs(polygon2.m_centroid.y - center.y) < absTol + relTol * b2Abs(center.y));
  const float mass = 4.0f * hx * hy;
  const float inertia = (mass / 3.0f) * (hx * hx + hy * hy) + mass * b2Dot(center, center);
  b2MassData massData1;
  polygon1.ComputeMass(&massData1, 1.0f);
  CHECK(b2Abs(massData1.center.x - center.x) < absTol + relTol * b2Abs(center.x));
  CHECK(b2Abs(massData1.center.y - center.y) < absTol + relTol * b2Abs(center.y));
  CHECK(b2Abs(massData1.mass - mass) < 20.0f * (absTol + relTol * mass));
  CHECK(b2Abs(massData1.I - inertia) < 40.0f * (absTol + relTol * inertia));
  b2MassData massData2;
  polygon2.ComputeMass(&massData2, 1.0f);
  CHECK(b2Abs(massData2.center.x - center.x) < absTol + relTol * b2Abs(center.x));
  CHECK(b2Abs(massData2.center.y - center.y) < absTol + relTol * b2Abs(center.y));
  CHECK(b2Abs(massData2.mass - mass) < 20.0f * (absTol + relTol * mass));
  CHECK(b2Abs(massData2.I - inertia) < 40.0f * (absTol + relTol * inertia));
}

static void sweep_math() {
  b2Sweep sweep;
  sweep.localCenter.SetZero();
  sweep.c0.Set(-2.0f, 4.0f);
  sweep.c.Set(3.0f, 8.0f);
  sweep.a0 = 0.5f;
  sweep.a = 5.0f;
  sweep.alpha0 = 0.0f;
  b2Transform transform;
  sweep.GetTransform(&transform, 0.0f);
  CHECK(transform.p.x == sweep.c0.x);
  CHECK(transform.p.y == sweep.c0.y);
  CHECK(transform.q.c == cosf(sweep.a0));
  CHECK(transform.q.s == sinf(sweep.a0));
  sweep.GetTransform(&transform, 1.0f);
  CHECK(transform.p.x == sweep.c.x);
  CHECK(transform.p.y == sweep.c.y);
  CHECK(transform.q.c == cosf(sweep.a));
  CHECK(transform.q.s == sinf(sweep.a));
}

// locked-world error convention: creation calls made from inside a callback are silent no-ops
static b2World* g_world = nullptr;
static b2Body* g_created = (b2Body*)1;
class LockProbe : public b2ContactListener {
 public:
  void BeginContact(b2Contact*) override {
    CHECK(g_world->IsLocked());
    b2BodyDef bd;
    g_created = g_world->CreateBody(&bd);
  }
};
static void locked_world_is_silent() {
  b2World world(b2Vec2(0.0f, -10.0f));
  g_world = &world;
  LockProbe probe;
  world.SetContactListener(&probe);
  b2CircleShape circle;
  circle.m_radius = 1.0f;
  b2BodyDef bd;
  bd.type = b2_dynamicBody;
  b2Body* a = world.CreateBody(&bd);
  bd.position.Set(0.5f, 0.0f);
  b2Body* b = world.CreateBody(&bd);
  a->CreateFixture(&circle, 1.0f);
  b->CreateFixture(&circle, 1.0f);
  int32 before = world.GetBodyCount();
  world.Step(1.0f / 60.0f, 6, 2);
  world.Step(1.0f / 60.0f, 6, 2);
  CHECK(g_created == nullptr);
  CHECK(world.GetBodyCount() == before);
  CHECK(!world.IsLocked());
}

// body list order: non-static bodies at the head in reverse creation order, static at the tail
static void body_list_order() {
  b2World world(b2Vec2(0.0f, 0.0f));
  b2BodyDef sd;
  b2BodyDef dd;
  dd.type = b2_dynamicBody;
  b2Body* s1 = world.CreateBody(&sd);
  b2Body* d1 = world.CreateBody(&dd);
  b2Body* d2 = world.CreateBody(&dd);
  b2Body* s2 = world.CreateBody(&sd);
  b2Body* expect[4] = {d2, d1, s1, s2};
  int i = 0;
  for (b2Body* b = world.GetBodyList(); b; b = b->GetNext()) CHECK(i < 4 && b == expect[i++]);
  CHECK(i == 4);
  CHECK(world.GetBodyCount() == 4);
  world.DestroyBody(d1);
  CHECK(world.GetBodyCount() == 3);
}

// revolute joint: reactions, motor and limit setters (b2_revolute_joint.cpp:333-447)
static void revolute_joint_api() {
  b2World world(b2Vec2(0.0f, -10.0f));
  b2BodyDef gd;
  b2Body* ground = world.CreateBody(&gd);
  b2BodyDef bd;
  bd.type = b2_dynamicBody;
  bd.position.Set(0.0f, -2.0f);  // hangs straight down from the pivot at the origin
  b2Body* rod = world.CreateBody(&bd);
  b2PolygonShape box;
  box.SetAsBox(0.1f, 2.0f);
  rod->CreateFixture(&box, 1.0f);  // mass 0.8
  b2RevoluteJointDef jd;
  jd.Initialize(ground, rod, b2Vec2(0.0f, 0.0f));
  b2RevoluteJoint* j = static_cast<b2RevoluteJoint*>(world.CreateJoint(&jd));
  CHECK(j != nullptr && world.GetJointCount() == 1);
  for (int i = 0; i < 60; ++i) world.Step(1.0f / 60.0f, 8, 3);
  b2Vec2 F = j->GetReactionForce(60.0f);
  CHECK(fabsf(F.y - 8.0f) < 0.05f);
  CHECK(fabsf(F.x) < 1e-3f);
  CHECK(fabsf(j->GetJointAngle()) < 1e-4f);
  CHECK(j->GetReactionTorque(60.0f) == 0.0f);
  CHECK(j->GetMotorTorque(60.0f) == 0.0f);
  // motor: drives the joint speed and reports the torque that holds the rod against gravity
  j->SetMaxMotorTorque(1000.0f);
  j->SetMotorSpeed(1.0f);
  j->EnableMotor(true);
  CHECK(j->IsMotorEnabled() && j->GetMotorSpeed() == 1.0f && j->GetMaxMotorTorque() == 1000.0f);
  CHECK(rod->IsAwake());
  for (int i = 0; i < 30; ++i) world.Step(1.0f / 60.0f, 8, 3);
  CHECK(fabsf(j->GetJointSpeed() - 1.0f) < 1e-3f);
  CHECK(j->GetJointAngle() > 0.45f && j->GetJointAngle() < 0.55f);
  float torque = j->GetMotorTorque(60.0f);
  CHECK(torque > 5.0f && torque < 12.0f);  // m g l sin(angle) = 0.8 * 10 * 2 * sin(0.5) = 7.7
  CHECK(j->GetReactionTorque(60.0f) == torque);
  // limit: the motor is released and the joint pushed back inside [-0.2, 0.2]
  j->EnableMotor(false);
  j->SetLimits(-0.2f, 0.2f);
  j->EnableLimit(true);
  CHECK(j->IsLimitEnabled() && j->GetLowerLimit() == -0.2f && j->GetUpperLimit() == 0.2f);
  float lo = 10.0f, hi = -10.0f;
  for (int i = 0; i < 240; ++i) {
    world.Step(1.0f / 60.0f, 8, 3);
    if (i >= 60) {
      lo = b2Min(lo, j->GetJointAngle());
      hi = b2Max(hi, j->GetJointAngle());
    }
  }
  CHECK(hi <= 0.2f + 0.04f && lo >= -0.2f - 0.04f);
  CHECK(hi - lo > 0.05f);  // still swinging between the stops
  printf("revolute: F=(%.9g, %.9g) torque=%.9g swing=[%.9g, %.9g] angle=%.9g\n", F.x, F.y, torque, lo, hi, j->GetJointAngle());
  world.DestroyJoint(j);
  CHECK(world.GetJointCount() == 0);
  world.Step(1.0f / 60.0f, 8, 3);
  CHECK(rod->GetLinearVelocity().y < 0.0f || rod->GetPosition().y < -2.0f);  // free fall now
}

// a pyramid settles to the same place whether or not the contact buffers had to grow on the way
// (reference: b2BlockAllocator / b2GrowableStack grow transparently; §8(b) "grow-and-retry, never drop")
static float pyramid_top_height(int rows) {
  b2World world(b2Vec2(0.0f, -10.0f));
  b2BodyDef gd;
  b2Body* ground = world.CreateBody(&gd);
  b2EdgeShape edge;
  edge.SetTwoSided(b2Vec2(-40.0f, 0.0f), b2Vec2(40.0f, 0.0f));
  ground->CreateFixture(&edge, 0.0f);
  b2PolygonShape box;
  box.SetAsBox(0.5f, 0.5f);
  b2Body* top = nullptr;
  for (int i = 0; i < rows; ++i)
    for (int j = i; j < rows; ++j) {
      b2BodyDef bd;
      bd.type = b2_dynamicBody;
      bd.position.Set(-7.0f + 0.5625f * (float)i + 1.125f * (float)(j - i), 0.75f + 1.25f * (float)i);
      top = world.CreateBody(&bd);
      top->CreateFixture(&box, 5.0f);
    }
  for (int i = 0; i < 200; ++i) world.Step(1.0f / 60.0f, 8, 3);
  CHECK(world.GetContactCount() > rows * rows / 2);
  return top->GetPosition().y;
}
static void contact_buffers_grow() {
  const int rows = 12;
#ifdef B2G_WORLD_H
  b2World::SetDefaultCapacity(1 << 10, 1 << 10, 64);  // 78 boxes need ~220 contacts: grows twice
#endif
  float small = pyramid_top_height(rows);
#ifdef B2G_WORLD_H
  b2World::SetDefaultCapacity(1 << 16, 1 << 16, 1 << 18);
#endif
  float large = pyramid_top_height(rows);
  CHECK(small > 11.5f && small < 11.8f);  // 12 rows of unit boxes plus their skins (reference: 11.6345)
  CHECK(fabsf(small - large) < 0.01f);
  printf("pyramid top: %.6f (grown buffers) %.6f (large buffers)\n", small, large);
}

// A scripted session of world edits between steps.  Every TRACE line is compared numerically
// between the build against the reference and the build against the drop-in (tests/test_host_api.py).
static void trace(const char* tag, b2World& world, b2Body** b, int n) {
  printf("TRACE %s contacts=%d", tag, world.GetContactCount());
  for (int i = 0; i < n; ++i) {
    if (!b[i]) continue;
    printf(" | %.5f %.5f %.5f %d", b[i]->GetPosition().x, b[i]->GetPosition().y, b[i]->GetAngle(), b[i]->IsAwake() ? 1 : 0);
  }
  printf("\n");
}
static void world_editing_session() {
  b2World world(b2Vec2(0.0f, -10.0f));
  b2BodyDef gd;
  b2Body* ground = world.CreateBody(&gd);
  b2EdgeShape edge;
  edge.SetTwoSided(b2Vec2(-30.0f, 0.0f), b2Vec2(30.0f, 0.0f));
  ground->CreateFixture(&edge, 0.0f);
  b2PolygonShape box;
  box.SetAsBox(0.5f, 0.5f);
  b2Body* b[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  for (int i = 0; i < 5; ++i) {  // a stack of five boxes
    b2BodyDef bd;
    bd.type = b2_dynamicBody;
    bd.position.Set(0.0f, 0.55f + 1.05f * (float)i);
    b[i] = world.CreateBody(&bd);
    b[i]->CreateFixture(&box, 2.0f);
  }
  {  // a ball to the side and a two-fixture dumbbell
    b2BodyDef bd;
    bd.type = b2_dynamicBody;
    bd.position.Set(4.0f, 3.0f);
    b[5] = world.CreateBody(&bd);
    b2CircleShape ball;
    ball.m_radius = 0.4f;
    b[5]->CreateFixture(&ball, 1.0f);
    bd.position.Set(-5.0f, 2.0f);
    b[6] = world.CreateBody(&bd);
    b2CircleShape end;
    end.m_radius = 0.3f;
    end.m_p.Set(-0.8f, 0.0f);
    b[6]->CreateFixture(&end, 1.0f);
    end.m_p.Set(0.8f, 0.0f);
    b[6]->CreateFixture(&end, 1.0f);
  }
  const float dt = 1.0f / 60.0f;
  for (int i = 0; i < 60; ++i) world.Step(dt, 8, 3);
  trace("settled", world, b, 7);
  world.ShiftOrigin(b2Vec2(100.0f, -50.0f));                     // floating origin: there and back again
  CHECK(fabsf(b[5]->GetPosition().x - (4.0f - 100.0f)) < 1e-4f && fabsf(ground->GetPosition().y - 50.0f) < 1e-4f);
  world.Step(dt, 8, 3);
  world.ShiftOrigin(b2Vec2(-100.0f, 50.0f));
  for (int i = 0; i < 29; ++i) world.Step(dt, 8, 3);
  trace("shifted", world, b, 7);
  b[4]->ApplyLinearImpulseToCenter(b2Vec2(0.4f, 0.0f), true);   // nudge the top box: it slides a little
  b[5]->ApplyAngularImpulse(0.05f, true);                        // and roll the ball
  for (int i = 0; i < 30; ++i) world.Step(dt, 8, 3);
  trace("impulse", world, b, 7);
  world.DestroyBody(b[4]);                                       // take the top box away
  b[4] = nullptr;
  CHECK(world.GetBodyCount() == 7);
  for (int i = 0; i < 60; ++i) world.Step(dt, 8, 3);
  trace("destroyed", world, b, 7);
  b[5]->SetTransform(b2Vec2(8.0f, 2.0f), 0.4f);                  // drop the ball somewhere else
  b[5]->SetLinearVelocity(b2Vec2(0.5f, -2.0f));
  b[5]->SetAngularVelocity(0.0f);
  for (int i = 0; i < 45; ++i) world.Step(dt, 8, 3);
  trace("teleported", world, b, 7);
  b[6]->DestroyFixture(b[6]->GetFixtureList());                  // the dumbbell loses one end and tips over
  for (int i = 0; i < 30; ++i) world.Step(dt, 8, 3);
  trace("fixture_destroyed", world, b, 7);
  b[3]->SetTransform(b2Vec2(-10.0f, 3.0f), 0.0f);                // lift a box aside, freeze it in the air, thaw it
  b[3]->SetType(b2_staticBody);
  for (int i = 0; i < 10; ++i) world.Step(dt, 8, 3);
  CHECK(b[3]->GetLinearVelocity().Length() == 0.0f && b[3]->GetPosition().y == 3.0f);
  trace("frozen", world, b, 7);
  b[3]->SetType(b2_dynamicBody);
  for (int i = 0; i < 60; ++i) world.Step(dt, 8, 3);
  trace("retyped", world, b, 7);
  b[3]->SetTransform(b2Vec2(-12.0f, 3.0f), 0.0f);                // a disabled body is never the seed of an island:
  b[3]->SetEnabled(false);                                       // alone in the air it just hangs there
  for (int i = 0; i < 30; ++i) world.Step(dt, 8, 3);
  CHECK(b[3]->GetPosition().y == 3.0f);
  trace("disabled", world, b, 7);
  b[3]->SetEnabled(true);
  for (int i = 0; i < 60; ++i) world.Step(dt, 8, 3);
  CHECK(b[3]->GetPosition().y < 0.6f);
  trace("enabled", world, b, 7);
  b[2]->SetGravityScale(-0.5f);                                  // the (now) top box floats up against a steady wind
  for (int i = 0; i < 40; ++i) {
    b[2]->ApplyForceToCenter(b2Vec2(1.5f, 0.0f), true);
    world.Step(dt, 8, 3);
  }
  trace("forces", world, b, 7);
  b[2]->SetGravityScale(1.0f);
  b[2]->SetFixedRotation(true);
  b[2]->SetAngularVelocity(3.0f);
  world.Step(dt, 8, 3);
  CHECK(b[2]->GetAngularVelocity() == 0.0f || b[2]->GetInertia() == 0.0f);
  for (int i = 0; i < 240; ++i) world.Step(dt, 8, 3);            // everything comes to rest and falls asleep
  trace("resting", world, b, 7);
}

// a user b2ContactFilter: boxes whose user data carry the same non-zero team pass through each other
struct TeamFilter : b2ContactFilter {
  int calls = 0;
  bool ShouldCollide(b2Fixture* a, b2Fixture* b) override {
    ++calls;
    uintptr_t ta = a->GetBody()->GetUserData().pointer, tb = b->GetBody()->GetUserData().pointer;
    if (ta != 0 && ta == tb) return false;
    return b2ContactFilter::ShouldCollide(a, b);
  }
};
static void user_contact_filter() {
  b2World world(b2Vec2(0.0f, -10.0f));
  TeamFilter filter;
  world.SetContactFilter(&filter);
  b2BodyDef gd;
  b2Body* ground = world.CreateBody(&gd);
  b2EdgeShape edge;
  edge.SetTwoSided(b2Vec2(-20.0f, 0.0f), b2Vec2(20.0f, 0.0f));
  ground->CreateFixture(&edge, 0.0f);
  b2PolygonShape box;
  box.SetAsBox(0.5f, 0.5f);
  b2Body* b[3];
  for (int i = 0; i < 3; ++i) {  // two overlapping boxes of team 7 side by side on the ground, a neutral one on top
    b2BodyDef bd;
    bd.type = b2_dynamicBody;
    bd.position.Set(i == 0 ? 0.0f : (i == 1 ? 0.4f : 0.2f), i < 2 ? 0.6f : 1.8f);
    bd.userData.pointer = i < 2 ? 7 : 0;
    b[i] = world.CreateBody(&bd);
    b[i]->CreateFixture(&box, 1.0f);
  }
  for (int i = 0; i < 180; ++i) world.Step(1.0f / 60.0f, 8, 3);
  CHECK(filter.calls > 0);
  // team mates interpenetrate: both rest on the ground; the neutral box rests on them
  CHECK(fabsf(b[0]->GetPosition().y - 0.51f) < 0.02f);
  CHECK(fabsf(b[1]->GetPosition().y - 0.51f) < 0.02f);
  CHECK(fabsf(b[1]->GetPosition().x - b[0]->GetPosition().x - 0.4f) < 0.02f);  // still overlapping by 0.6
  CHECK(fabsf(b[2]->GetPosition().y - 1.52f) < 0.03f);
  for (b2Contact* c = world.GetContactListStart(); c != world.GetContactListEnd(); c = c->GetNext()) {
    uintptr_t ta = c->GetFixtureA()->GetBody()->GetUserData().pointer, tb = c->GetFixtureB()->GetBody()->GetUserData().pointer;
    CHECK(!(ta != 0 && ta == tb));
  }
  int contactsWith = world.GetContactCount();
  // change sides: box 1 leaves the team, Refilter offers its contacts to the filter again and the
  // (still overlapping) rejected pair is accepted at the next step
  b[1]->GetUserData().pointer = 0;
  b[1]->GetFixtureList()->Refilter();
  b[1]->SetAwake(true);  // Refilter does not wake anybody, and a contact between sleepers is never evaluated
  for (int i = 0; i < 120; ++i) world.Step(1.0f / 60.0f, 8, 3);
  CHECK(world.GetContactCount() == contactsWith + 1);  // the former team mates now have a contact
  float gap = b[1]->GetPosition().x - b[0]->GetPosition().x;
  CHECK(gap > 0.95f);                                   // and have been pushed apart
  printf("contact filter: %d calls, contacts %d -> %d, former team mates %.3f apart\n", filter.calls, contactsWith,
         world.GetContactCount(), gap);
}

// distance joint: a rigid rod keeps its length, a spring oscillates at the frequency it was tuned to,
// a rope with limits stops at its maximum length (b2_distance_joint.cpp)
static void distance_joint_api() {
  b2World world(b2Vec2(0.0f, -10.0f));
  b2BodyDef gd;
  b2Body* ground = world.CreateBody(&gd);
  b2CircleShape ball;
  ball.m_radius = 0.25f;
  b2BodyDef bd;
  bd.type = b2_dynamicBody;
  bd.position.Set(2.0f, 0.0f);   // rod: starts horizontal, swings like a pendulum
  b2Body* bob = world.CreateBody(&bd);
  bob->CreateFixture(&ball, 1.0f);
  b2DistanceJointDef rod;
  rod.Initialize(ground, bob, b2Vec2(0.0f, 0.0f), bob->GetPosition());
  b2DistanceJoint* jr = static_cast<b2DistanceJoint*>(world.CreateJoint(&rod));
  CHECK(jr != nullptr && fabsf(jr->GetLength() - 2.0f) < 1e-6f && jr->GetMinLength() == jr->GetMaxLength());
  bd.position.Set(10.0f, -1.0f);  // spring: 2 Hz, lightly damped, released from rest at its rest length
  b2Body* mass = world.CreateBody(&bd);
  mass->CreateFixture(&ball, 1.0f);
  b2DistanceJointDef sp;
  sp.Initialize(ground, mass, b2Vec2(10.0f, 0.0f), mass->GetPosition());
  b2LinearStiffness(sp.stiffness, sp.damping, 2.0f, 0.05f, ground, mass);
  sp.minLength = 0.1f;
  sp.maxLength = 5.0f;
  b2DistanceJoint* js = static_cast<b2DistanceJoint*>(world.CreateJoint(&sp));
  CHECK(js->GetStiffness() > 0.0f && js->GetDamping() > 0.0f);
  bd.position.Set(20.0f, -1.0f);  // rope: slack between 0.5 and 1.5, the ball falls until it is taut
  b2Body* hung = world.CreateBody(&bd);
  hung->CreateFixture(&ball, 1.0f);
  b2DistanceJointDef rp;
  rp.Initialize(ground, hung, b2Vec2(20.0f, 0.0f), hung->GetPosition());
  rp.minLength = 0.5f;
  rp.maxLength = 1.5f;
  b2DistanceJoint* jp = static_cast<b2DistanceJoint*>(world.CreateJoint(&rp));
  int crossings = 0;
  float prev = 0.0f, lowest = 0.0f, rodErr = 0.0f;
  for (int i = 0; i < 240; ++i) {
    world.Step(1.0f / 60.0f, 8, 3);
    rodErr = b2Max(rodErr, fabsf(jr->GetCurrentLength() - 2.0f));
    float v = mass->GetLinearVelocity().y;
    if (i > 0 && ((prev < 0.0f && v >= 0.0f) || (prev > 0.0f && v <= 0.0f))) ++crossings;
    prev = v;
    lowest = b2Min(lowest, hung->GetPosition().y);
  }
  CHECK(rodErr < 0.02f);
  CHECK(crossings >= 8 && crossings <= 17);         // 2 Hz: at most 16 velocity zero crossings in 4 s, fewer once damped out
  CHECK(lowest < -1.45f && lowest > -1.56f);        // stopped by the upper limit
  CHECK(fabsf(jp->GetCurrentLength() - 1.5f) < 0.02f);
  b2Vec2 F = jp->GetReactionForce(60.0f);           // the taut rope carries the ball's weight
  float m = hung->GetMass();
  CHECK(fabsf(F.y - 10.0f * m) < 0.05f * 10.0f * m);
  CHECK(jp->SetMaxLength(1.0f) == 1.0f && jp->SetLength(0.7f) == 0.7f && jp->SetMinLength(2.0f) == 1.0f);
  hung->SetAwake(true);  // the setters do not wake anybody
  for (int i = 0; i < 120; ++i) world.Step(1.0f / 60.0f, 8, 3);
  CHECK(fabsf(jp->GetCurrentLength() - 1.0f) < 0.02f);
  printf("distance: rod error %.5f, %d zero crossings, rope bottom %.4f, rope force %.4f\n", rodErr, crossings, lowest, F.y);
}

// weld joint: a welded cantilever carries its own weight (reaction force and torque at the root)
static void weld_joint_api() {
  b2World world(b2Vec2(0.0f, -10.0f));
  b2BodyDef gd;
  b2Body* ground = world.CreateBody(&gd);
  b2PolygonShape plank;
  plank.SetAsBox(1.0f, 0.1f);
  b2BodyDef bd;
  bd.type = b2_dynamicBody;
  bd.position.Set(1.0f, 5.0f);
  b2Body* beam = world.CreateBody(&bd);
  beam->CreateFixture(&plank, 5.0f);   // mass 2 x 0.2 x 5 = 2
  b2WeldJointDef jd;
  jd.Initialize(ground, beam, b2Vec2(0.0f, 5.0f));
  b2WeldJoint* j = static_cast<b2WeldJoint*>(world.CreateJoint(&jd));
  CHECK(j != nullptr && j->GetReferenceAngle() == 0.0f && j->GetStiffness() == 0.0f);
  for (int i = 0; i < 120; ++i) world.Step(1.0f / 60.0f, 8, 3);
  CHECK(fabsf(beam->GetPosition().y - 5.0f) < 0.02f && fabsf(beam->GetAngle()) < 0.02f);
  b2Vec2 F = j->GetReactionForce(60.0f);
  float T = j->GetReactionTorque(60.0f);
  CHECK(fabsf(F.y - 20.0f) < 0.5f);          // the weight of the beam
  CHECK(fabsf(T - 20.0f) < 1.0f);            // weight x lever arm of 1 m
  // a soft weld sags and swings
  b2AngularStiffness(jd.stiffness, jd.damping, 1.0f, 0.2f, ground, beam);
  j->SetStiffness(jd.stiffness);
  j->SetDamping(jd.damping);
  beam->SetAwake(true);
  float lowest = 0.0f;
  for (int i = 0; i < 120; ++i) {
    world.Step(1.0f / 60.0f, 8, 3);
    lowest = b2Min(lowest, beam->GetAngle());
  }
  CHECK(lowest < -0.1f);
  printf("weld: F=(%.4f, %.4f) T=%.4f soft sag %.4f\n", F.x, F.y, T, lowest);
}

// prismatic joint: a motorised lift stops at its upper limit carrying its own weight, then falls to the
// lower limit when the motor is switched off; relative rotation stays locked
static void prismatic_joint_api() {
  b2World world(b2Vec2(0.0f, -10.0f));
  b2BodyDef gd;
  b2Body* ground = world.CreateBody(&gd);
  b2PolygonShape plate;
  plate.SetAsBox(1.0f, 0.25f);
  b2BodyDef bd;
  bd.type = b2_dynamicBody;
  bd.position.Set(0.5f, 3.0f);   // anchor off-centre: the rail must also carry a torque
  b2Body* lift = world.CreateBody(&bd);
  lift->CreateFixture(&plate, 2.0f);   // mass 2 x 0.5 x 2 = 2
  b2PrismaticJointDef jd;
  jd.Initialize(ground, lift, b2Vec2(0.0f, 3.0f), b2Vec2(0.0f, 2.0f));
  jd.enableLimit = true;
  jd.lowerTranslation = -1.0f;
  jd.upperTranslation = 1.5f;
  jd.enableMotor = true;
  jd.motorSpeed = 1.0f;
  jd.maxMotorForce = 100.0f;
  b2PrismaticJoint* j = static_cast<b2PrismaticJoint*>(world.CreateJoint(&jd));
  CHECK(j != nullptr && j->IsLimitEnabled() && j->IsMotorEnabled() && j->GetUpperLimit() == 1.5f);
  CHECK(fabsf(j->GetLocalAxisA().y - 1.0f) < 1e-6f);   // the constructor normalises the axis
  for (int i = 0; i < 60; ++i) world.Step(1.0f / 60.0f, 8, 3);
  float speed = j->GetJointSpeed();
  CHECK(fabsf(j->GetJointTranslation() - 1.0f) < 0.05f && fabsf(speed - 1.0f) < 0.02f);
  float motorForce = j->GetMotorForce(60.0f);
  CHECK(fabsf(motorForce - 20.0f) < 0.5f);   // the motor carries the weight
  for (int i = 0; i < 120; ++i) world.Step(1.0f / 60.0f, 8, 3);
  CHECK(fabsf(j->GetJointTranslation() - 1.5f) < 0.02f);
  CHECK(fabsf(lift->GetAngle()) < 0.01f && fabsf(lift->GetPosition().x - 0.5f) < 0.01f);
  float T = j->GetReactionTorque(60.0f);
  CHECK(fabsf(fabsf(T) - 10.0f) < 0.5f);     // weight x 0.5 m lever arm
  j->EnableMotor(false);
  for (int i = 0; i < 180; ++i) world.Step(1.0f / 60.0f, 8, 3);
  CHECK(fabsf(j->GetJointTranslation() + 1.0f) < 0.02f);
  b2Vec2 F = j->GetReactionForce(60.0f);
  CHECK(fabsf(F.y - 20.0f) < 0.5f && fabsf(F.x) < 0.5f);   // the lower limit carries the weight
  j->SetLimits(-2.0f, 1.5f);
  for (int i = 0; i < 120; ++i) world.Step(1.0f / 60.0f, 8, 3);
  CHECK(fabsf(j->GetJointTranslation() + 2.0f) < 0.02f);
  printf("prismatic: speed %.4f motor force %.4f torque %.4f F=(%.4f, %.4f)\n", speed, motorForce, T, F.x, F.y);
}

// wheel joint: a wheel hanging on its suspension spring settles where the spring carries its weight,
// the motor spins it up, the limits stop the travel
static void wheel_joint_api() {
  b2World world(b2Vec2(0.0f, -10.0f));
  b2BodyDef gd;
  b2Body* ground = world.CreateBody(&gd);
  b2CircleShape tyre;
  tyre.m_radius = 0.5f;
  b2BodyDef bd;
  bd.type = b2_dynamicBody;
  bd.position.Set(0.0f, 5.0f);
  b2Body* wheel = world.CreateBody(&bd);
  wheel->CreateFixture(&tyre, 2.0f);   // mass 2 x pi x 0.25 = 1.5708
  b2WheelJointDef jd;
  jd.Initialize(ground, wheel, wheel->GetPosition(), b2Vec2(0.0f, 1.0f));
  b2LinearStiffness(jd.stiffness, jd.damping, 2.0f, 0.7f, ground, wheel);
  jd.enableMotor = true;
  jd.motorSpeed = 3.0f;
  jd.maxMotorTorque = 50.0f;
  b2WheelJoint* j = static_cast<b2WheelJoint*>(world.CreateJoint(&jd));
  CHECK(j != nullptr && j->IsMotorEnabled() && !j->IsLimitEnabled() && j->GetStiffness() == jd.stiffness);
  for (int i = 0; i < 240; ++i) world.Step(1.0f / 60.0f, 8, 3);
  // static sag = m g / k
  float sag = j->GetJointTranslation(), expect = -wheel->GetMass() * 10.0f / jd.stiffness;
  CHECK(fabsf(sag - expect) < 0.01f);
  CHECK(fabsf(j->GetJointAngularSpeed() - 3.0f) < 0.01f && fabsf(wheel->GetPosition().x) < 1e-3f);
  b2Vec2 F = j->GetReactionForce(60.0f);
  CHECK(fabsf(F.y - wheel->GetMass() * 10.0f) < 0.2f);   // the spring carries the weight
  // a rigid wheel joint with limits: drops to the lower stop
  j->SetStiffness(0.0f);
  j->SetDamping(0.0f);
  j->EnableMotor(false);
  j->SetLimits(-1.0f, 0.5f);
  j->EnableLimit(true);
  wheel->SetAwake(true);
  for (int i = 0; i < 180; ++i) world.Step(1.0f / 60.0f, 8, 3);
  CHECK(fabsf(j->GetJointTranslation() + 1.0f) < 0.02f);
  CHECK(fabsf(j->GetJointLinearSpeed()) < 0.01f);
  printf("wheel: sag %.4f (m g / k = %.4f) spin %.4f F=(%.4f, %.4f) stop %.4f\n", sag, expect, j->GetJointAngularSpeed(), F.x, F.y,
         j->GetJointTranslation());
}

// friction joint: a puck sliding in a world without gravity is braked by exactly maxForce and stops;
// motor joint: a body is pulled to a linear and angular offset and held there against gravity
static void friction_and_motor_joint_api() {
  {
    b2World world(b2Vec2(0.0f, 0.0f));
    b2BodyDef gd;
    b2Body* ground = world.CreateBody(&gd);
    b2PolygonShape box;
    box.SetAsBox(0.5f, 0.5f);
    b2BodyDef bd;
    bd.type = b2_dynamicBody;
    bd.position.Set(0.0f, 0.0f);
    bd.linearVelocity.Set(5.0f, 0.0f);
    bd.angularVelocity = 2.0f;
    b2Body* puck = world.CreateBody(&bd);
    puck->CreateFixture(&box, 1.0f);   // mass 1, inertia 1/6
    b2FrictionJointDef jd;
    jd.Initialize(ground, puck, puck->GetPosition());
    jd.maxForce = 10.0f;
    jd.maxTorque = 1.0f;
    b2FrictionJoint* j = static_cast<b2FrictionJoint*>(world.CreateJoint(&jd));
    CHECK(j != nullptr && j->GetMaxForce() == 10.0f && j->GetMaxTorque() == 1.0f);
    world.Step(1.0f / 60.0f, 8, 3);
    b2Vec2 F = j->GetReactionForce(60.0f);
    CHECK(fabsf(F.x + 10.0f) < 1e-3f && fabsf(F.y) < 1e-3f);                 // braking with the whole budget
    CHECK(fabsf(puck->GetLinearVelocity().x - (5.0f - 10.0f / 60.0f)) < 1e-4f);
    for (int i = 0; i < 59; ++i) world.Step(1.0f / 60.0f, 8, 3);
    CHECK(fabsf(puck->GetLinearVelocity().x) < 1e-4f && fabsf(puck->GetAngularVelocity()) < 1e-4f);
    CHECK(fabsf(puck->GetPosition().x - 1.2083f) < 0.01f);                   // v0^2 / (2 a), semi-implicit
    j->SetMaxForce(0.0f);
    j->SetMaxTorque(0.0f);
    puck->SetLinearVelocity(b2Vec2(1.0f, 0.0f));
    for (int i = 0; i < 60; ++i) world.Step(1.0f / 60.0f, 8, 3);
    CHECK(fabsf(puck->GetLinearVelocity().x - 1.0f) < 1e-5f);                // no budget, no braking
    printf("friction: F=(%.4f, %.4f) stop at x=%.4f\n", F.x, F.y, puck->GetPosition().x - 1.0f);
  }
  {
    b2World world(b2Vec2(0.0f, -10.0f));
    b2BodyDef gd;
    b2Body* ground = world.CreateBody(&gd);
    b2PolygonShape box;
    box.SetAsBox(0.5f, 0.5f);
    b2BodyDef bd;
    bd.type = b2_dynamicBody;
    bd.position.Set(0.0f, 4.0f);
    b2Body* body = world.CreateBody(&bd);
    body->CreateFixture(&box, 2.0f);   // mass 2
    b2MotorJointDef jd;
    jd.Initialize(ground, body);
    CHECK(jd.linearOffset.x == 0.0f && jd.linearOffset.y == 4.0f && jd.angularOffset == 0.0f);
    jd.maxForce = 1000.0f;
    jd.maxTorque = 1000.0f;
    b2MotorJoint* j = static_cast<b2MotorJoint*>(world.CreateJoint(&jd));
    CHECK(j != nullptr && j->GetCorrectionFactor() == 0.3f);
    j->SetLinearOffset(b2Vec2(2.0f, 5.0f));
    j->SetAngularOffset(1.0f);
    for (int i = 0; i < 180; ++i) world.Step(1.0f / 60.0f, 8, 3);
    b2Vec2 p = body->GetPosition();
    CHECK(fabsf(p.x - 2.0f) < 0.01f && fabsf(p.y - 5.0f) < 0.02f && fabsf(body->GetAngle() - 1.0f) < 0.01f);
    b2Vec2 F = j->GetReactionForce(60.0f);
    CHECK(fabsf(F.y - 20.0f) < 0.2f && fabsf(F.x) < 0.2f);                   // holds the weight
    j->SetMaxForce(5.0f);                                                     // a quarter of the weight: it falls
    body->SetAwake(true);                                                     // (the setter does not wake it)
    for (int i = 0; i < 60; ++i) world.Step(1.0f / 60.0f, 8, 3);
    CHECK(body->GetPosition().y < 2.0f);
    printf("motor: at (%.4f, %.4f) angle %.4f F=(%.4f, %.4f) then y=%.4f\n", p.x, p.y, body->GetAngle(), F.x, F.y,
           body->GetPosition().y);
  }
}

// mouse joint: a box is dragged to a target and hangs there, sagging by m g / k; too weak a grip cannot lift it
static void mouse_joint_api() {
  b2World world(b2Vec2(0.0f, -10.0f));
  b2BodyDef gd;
  b2Body* ground = world.CreateBody(&gd);
  b2EdgeShape edge;
  edge.SetTwoSided(b2Vec2(-20.0f, 0.0f), b2Vec2(20.0f, 0.0f));
  ground->CreateFixture(&edge, 0.0f);
  b2Body* anchor = world.CreateBody(&gd);   // shapeless: the joint must not switch off the floor contact
  b2PolygonShape box;
  box.SetAsBox(0.5f, 0.5f);
  b2BodyDef bd;
  bd.type = b2_dynamicBody;
  bd.position.Set(0.0f, 0.51f);
  b2Body* body = world.CreateBody(&bd);
  body->CreateFixture(&box, 2.0f);   // mass 2
  b2MouseJointDef jd;
  jd.bodyA = anchor;
  jd.bodyB = body;
  jd.target = body->GetPosition();
  jd.maxForce = 1000.0f * body->GetMass();
  b2LinearStiffness(jd.stiffness, jd.damping, 5.0f, 0.7f, anchor, body);
  b2MouseJoint* j = static_cast<b2MouseJoint*>(world.CreateJoint(&jd));
  CHECK(j != nullptr && j->GetTarget().y == 0.51f && j->GetStiffness() == jd.stiffness);
  j->SetTarget(b2Vec2(3.0f, 4.0f));
  for (int i = 0; i < 240; ++i) world.Step(1.0f / 60.0f, 8, 3);
  b2Vec2 p = body->GetPosition();
  float sag = body->GetMass() * 10.0f / jd.stiffness;
  CHECK(fabsf(p.x - 3.0f) < 0.01f && fabsf(p.y - (4.0f - sag)) < 0.01f);
  b2Vec2 F = j->GetReactionForce(60.0f);
  CHECK(fabsf(F.y - 20.0f) < 0.1f && fabsf(F.x) < 0.1f && j->GetReactionTorque(60.0f) == 0.0f);
  j->SetMaxForce(10.0f);   // half the weight
  body->SetAwake(true);    // (it hangs still and has fallen asleep; the setter does not wake it)
  for (int i = 0; i < 120; ++i) world.Step(1.0f / 60.0f, 8, 3);
  CHECK(body->GetPosition().y < 0.6f && body->GetPosition().y > 0.4f);   // back on the floor
  b2Vec2 F2 = j->GetReactionForce(60.0f);
  CHECK(fabsf(F2.Length() - 10.0f) < 1e-3f);
  printf("mouse: at (%.4f, %.4f), sag %.4f, F=(%.4f, %.4f), weak grip |F|=%.4f y=%.4f\n", p.x, p.y, sag, F.x, F.y, F2.Length(),
         body->GetPosition().y);
}

// Body edits made INSIDE contact callbacks (the world is locked, but impulses and velocity setters are
// allowed there, b2_body.h): an impulse applied in BeginContact acts in the same step's solve, a velocity set
// in PostSolve survives the step (b2_contact.cpp:197-209, b2_island.cpp:430-441).
struct KickListener : public b2ContactListener {
  b2Body* ball = nullptr;
  b2Body* roller = nullptr;
  int kicks = 0, spins = 0;
  float vyAtBegin = 0.0f, rollerYAtPostSolve = 0.0f, rollerVyAtPostSolve = 1.0f;
  void BeginContact(b2Contact* c) override {
    b2Body* a = c->GetFixtureA()->GetBody();
    b2Body* b = c->GetFixtureB()->GetBody();
    if ((a == ball || b == ball) && kicks == 0) {
      vyAtBegin = ball->GetLinearVelocity().y;
      ball->ApplyLinearImpulse(b2Vec2(0.0f, 12.0f * ball->GetMass()), ball->GetWorldCenter(), true);
      ++kicks;
    }
  }
  void PostSolve(b2Contact* c, const b2ContactImpulse*) override {
    b2Body* a = c->GetFixtureA()->GetBody();
    b2Body* b = c->GetFixtureB()->GetBody();
    if ((a == roller || b == roller) && spins == 0) {
      rollerYAtPostSolve = roller->GetPosition().y;          // post-solve state: resting on the ground
      rollerVyAtPostSolve = roller->GetLinearVelocity().y;
      roller->SetAngularVelocity(-3.0f);
      ++spins;
    }
  }
};
static void edits_inside_callbacks() {
  b2World world(b2Vec2(0.0f, -10.0f));
  KickListener L;
  world.SetContactListener(&L);
  b2BodyDef gd;
  b2Body* ground = world.CreateBody(&gd);
  b2EdgeShape edge;
  edge.SetTwoSided(b2Vec2(-30.0f, 0.0f), b2Vec2(30.0f, 0.0f));
  ground->CreateFixture(&edge, 0.0f);
  b2CircleShape circle;
  circle.m_radius = 0.5f;
  b2BodyDef bd;
  bd.type = b2_dynamicBody;
  bd.position.Set(0.0f, 2.0f);
  b2Body* ball = world.CreateBody(&bd);
  ball->CreateFixture(&circle, 1.0f);
  bd.position.Set(6.0f, 1.0f);
  b2Body* roller = world.CreateBody(&bd);
  b2FixtureDef fd;
  fd.shape = &circle;
  fd.density = 1.0f;
  fd.friction = 0.8f;
  roller->CreateFixture(&fd);
  L.ball = ball;
  L.roller = roller;
  b2Body* b[2] = {ball, roller};
  float vyAfterKick = 0.0f, top = 0.0f;
  for (int i = 0; i < 120; ++i) {
    int before = L.kicks;
    world.Step(1.0f / 60.0f, 8, 3);
    if (before == 0 && L.kicks == 1) vyAfterKick = ball->GetLinearVelocity().y;
    if (ball->GetPosition().y > top) top = ball->GetPosition().y;
    if (i % 20 == 19) {
      char tag[32];
      snprintf(tag, sizeof(tag), "callbacks%d", i + 1);
      trace(tag, world, b, 2);
    }
  }
  CHECK(L.kicks == 1 && L.spins == 1);
  CHECK(L.vyAtBegin < -4.0f);                 // it was falling when it touched
  CHECK(vyAfterKick > 5.0f);                  // the impulse acted in the SAME step: the ball leaves upwards
  CHECK(top > 2.3f);                          // ... and flies above its drop height (6.33^2 / 2g + 0.5)
  CHECK(fabsf(L.rollerYAtPostSolve - 0.5f) < 0.05f && fabsf(L.rollerVyAtPostSolve) < 0.5f);  // PostSolve sees the solved state
  CHECK(roller->GetPosition().x > 6.5f);      // the spin set in PostSolve survived: it rolled to the right
  printf("callbacks: vy at begin %.4f, after the kick %.4f, apex %.4f, roller x %.4f\n", L.vyAtBegin, vyAfterKick, top,
         roller->GetPosition().x);
}

// A world that keeps creating and destroying bodies (projectiles, debris): every newcomer lands on the box left by
// the one before it.  The results must not depend on whether a body sits on a fresh or on a reused device row, and in
// the drop-in the device tables must stay bounded (the reference frees the memory in b2World::DestroyBody).
static void spawn_and_destroy_churn() {
  b2World world(b2Vec2(0.0f, -10.0f));
  b2BodyDef gd;
  b2Body* ground = world.CreateBody(&gd);
  b2PolygonShape gbox;
  gbox.SetAsBox(20.0f, 0.5f, b2Vec2(0.0f, -0.5f), 0.0f);
  ground->CreateFixture(&gbox, 0.0f);
  b2PolygonShape box;
  box.SetAsBox(0.5f, 0.5f);
  b2CircleShape ball;
  ball.m_radius = 0.4f;
  const int ALIVE = 10;
  b2Body* ring[ALIVE] = {nullptr};
  float landed = 0.0f;
  for (int i = 0; i < 120; ++i) {
    b2BodyDef bd;
    bd.type = b2_dynamicBody;
    bd.position.Set(1.5f * (float)(i % ALIVE) - 7.0f, 3.0f);
    b2Body* b = world.CreateBody(&bd);
    if (i % 3 == 2) b->CreateFixture(&ball, 1.0f);
    else b->CreateFixture(&box, 1.0f);
    for (int s = 0; s < 6; ++s) world.Step(1.0f / 60.0f, 8, 3);
    b2Body*& slot = ring[i % ALIVE];
    if (slot) {
      if (i == 118) landed = slot->GetPosition().y;  // body 108: a box, on the ground for most of its 60 steps
      world.DestroyBody(slot);
    }
    slot = b;
  }
  for (int s = 0; s < 90; ++s) world.Step(1.0f / 60.0f, 8, 3);
  CHECK(world.GetBodyCount() == 1 + ALIVE);
  CHECK(fabsf(landed - 0.505f) < 0.02f);
  for (int k = 0; k < ALIVE; ++k) {  // bodies 110..119 rest on the ground: boxes at 0.5, balls at 0.4 (+ slop)
    const int i = 110 + k;
    const float want = (i % 3 == 2) ? 0.4f : 0.5f;
    b2Body* b = ring[i % ALIVE];
    CHECK(fabsf(b->GetPosition().y - want) < 0.03f && fabsf(b->GetLinearVelocity().y) < 0.05f);
  }
  CHECK(world.GetContactCount() == ALIVE);
#ifdef B2G_WORLD_H
  CHECK(world.GetBodyIndexCount() <= 1 + ALIVE + 3);  // ground + the bodies alive at a time + rows waiting for a step
  printf("churn: %d device body rows for 121 bodies created\n", world.GetBodyIndexCount());
#endif
  printf("churn: body 108 rested at y = %.4f when it was destroyed, %d contacts at the end\n", landed, world.GetContactCount());
}

int main() {
  hello_world();
  begin_contact_test();
  polygon_mass_data();
  sweep_math();
  locked_world_is_silent();
  body_list_order();
  revolute_joint_api();
  distance_joint_api();
  weld_joint_api();
  prismatic_joint_api();
  wheel_joint_api();
  friction_and_motor_joint_api();
  mouse_joint_api();
  contact_buffers_grow();
  world_editing_session();
  user_contact_filter();
  edits_inside_callbacks();
  spawn_and_destroy_churn();
  printf(g_failed ? "FAILED %d checks\n" : "all API checks passed\n", g_failed);
  return g_failed ? 1 : 0;
}

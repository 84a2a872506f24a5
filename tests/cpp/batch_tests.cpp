// b2WorldBatch acceptance program (GPU build only: the reference has no batch).  K different small worlds — a
// motor-driven paddle in a box of boxes and balls, bodies spawned while running, one impulse, one destroyed body —
// are stepped (a) each as a b2World of its own and (b) as members of one b2WorldBatch.  World k of the batch must
// get EXACTLY the floats it gets alone (bodies of different worlds never interact, and nothing a world computes
// depends on where its rows sit in the arena), so the comparison is bitwise.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>

#include "box2d/box2d.h"

static int g_failures = 0;
#define CHECK(cond, ...)                       \
  do {                                         \
    if (!(cond)) {                             \
      ++g_failures;                            \
      printf("FAIL %s:%d: ", __FILE__, __LINE__); \
      printf(__VA_ARGS__);                     \
      printf("\n");                            \
    }                                          \
  } while (0)

struct Built {
  std::vector<b2Body*> bodies;
  b2Body* paddle = nullptr;
  b2RevoluteJoint* hinge = nullptr;
};

static Built build(b2World* w, int variant) {
  Built B;
  b2BodyDef gd;
  b2Body* ground = w->CreateBody(&gd);
  b2PolygonShape box;
  box.SetAsBox(12.0f, 0.5f, b2Vec2(0.0f, -0.5f), 0.0f);
  ground->CreateFixture(&box, 0.0f);
  box.SetAsBox(0.5f, 8.0f, b2Vec2(-12.0f, 8.0f), 0.0f);
  ground->CreateFixture(&box, 0.0f);
  box.SetAsBox(0.5f, 8.0f, b2Vec2(12.0f, 8.0f), 0.0f);
  ground->CreateFixture(&box, 0.0f);
  B.bodies.push_back(ground);
  {
    b2BodyDef bd;
    bd.type = b2_dynamicBody;
    bd.position.Set(0.0f, 3.0f);
    bd.allowSleep = false;
    B.paddle = w->CreateBody(&bd);
    b2PolygonShape blade;
    blade.SetAsBox(4.0f, 0.25f);
    B.paddle->CreateFixture(&blade, 5.0f);
    b2RevoluteJointDef jd;
    jd.Initialize(ground, B.paddle, b2Vec2(0.0f, 3.0f));
    jd.enableMotor = true;
    jd.motorSpeed = 0.6f + 0.1f * (float)variant;
    jd.maxMotorTorque = 1e5f;
    B.hinge = static_cast<b2RevoluteJoint*>(w->CreateJoint(&jd));
    B.bodies.push_back(B.paddle);
  }
  const int n = 24 + 3 * variant;
  for (int i = 0; i < n; ++i) {
    b2BodyDef bd;
    bd.type = b2_dynamicBody;
    bd.position.Set(-8.0f + 1.1f * (float)(i % 15), 5.0f + 1.2f * (float)(i / 15) + 0.05f * (float)variant);
    b2Body* b = w->CreateBody(&bd);
    if ((i + variant) % 3 == 0) {
      b2CircleShape c;
      c.m_radius = 0.4f;
      b->CreateFixture(&c, 1.0f);
    } else {
      b2PolygonShape p;
      p.SetAsBox(0.45f, 0.35f);
      b->CreateFixture(&p, 1.0f);
    }
    B.bodies.push_back(b);
  }
  return B;
}

// what happens to world `variant` at step `s` (identical in both runs)
static void script(b2World* w, Built& B, int variant, int s) {
  if (s < 20 && s % 2 == 0) {  // spawn
    b2BodyDef bd;
    bd.type = b2_dynamicBody;
    bd.position.Set(-6.0f + 0.7f * (float)(s % 11) + 0.01f * (float)variant, 12.0f);
    b2Body* b = w->CreateBody(&bd);
    b2PolygonShape p;
    p.SetAsBox(0.3f, 0.3f);
    b->CreateFixture(&p, 2.0f);
    B.bodies.push_back(b);
  }
  if (s == 60) B.bodies[5 + variant]->ApplyLinearImpulseToCenter(b2Vec2(3.0f, 8.0f), true);
  if (s == 90 && variant % 2 == 1) {
    w->DestroyBody(B.bodies[4]);
    B.bodies[4] = nullptr;
  }
  if (s == 120) B.hinge->SetMotorSpeed(-B.hinge->GetMotorSpeed());
}

struct State {
  std::vector<float> v;
  int contacts = 0;
};
static State snapshot(b2World* w, const Built& B) {
  State S;
  for (b2Body* b : B.bodies) {
    if (!b) continue;
    const b2Vec2 p = b->GetPosition(), lv = b->GetLinearVelocity();
    S.v.insert(S.v.end(), {p.x, p.y, b->GetAngle(), lv.x, lv.y, b->GetAngularVelocity(), b->IsAwake() ? 1.0f : 0.0f});
  }
  S.contacts = w->GetContactCount();
  return S;
}

struct CountFixtures : b2QueryCallback {
  b2World* world;
  int count = 0, foreign = 0;
  bool ReportFixture(b2Fixture* f) override {
    ++count;
    if (f->GetBody()->GetWorld() != world) ++foreign;
    return true;
  }
};

int main() {
  const int K = 6, STEPS = 200;
  const float dt = 1.0f / 60.0f;
  // (a) alone
  std::vector<State> alone(K);
  std::vector<std::vector<State>> aloneAt(K);
  for (int k = 0; k < K; ++k) {
    b2World w(b2Vec2(0.0f, -10.0f));
    Built B = build(&w, k);
    for (int s = 0; s < STEPS; ++s) {
      script(&w, B, k, s);
      w.Step(dt, 8, 3);
      if (s == 59 || s == 119) aloneAt[k].push_back(snapshot(&w, B));
    }
    alone[k] = snapshot(&w, B);
  }
  // (b) batched
  {
    std::vector<b2World*> worlds;
    std::vector<Built> built(K);
    b2WorldBatch batch;
    for (int k = 0; k < K; ++k) {
      worlds.push_back(new b2World(b2Vec2(0.0f, -10.0f)));
      built[k] = build(worlds[k], k);
      CHECK(batch.Add(worlds[k]), "Add(%d)", k);
    }
    CHECK(!batch.Add(worlds[0]), "a world cannot join twice");
    CHECK(batch.GetWorldCount() == K && batch.GetWorld(2) == worlds[2], "members");
    std::vector<std::vector<State>> at(K);
    for (int s = 0; s < STEPS; ++s) {
      for (int k = 0; k < K; ++k) script(worlds[k], built[k], k, s);
      batch.Step(dt, 8, 3);
      if (s == 59 || s == 119)
        for (int k = 0; k < K; ++k) at[k].push_back(snapshot(worlds[k], built[k]));
    }
    for (int k = 0; k < K; ++k) {
      State S = snapshot(worlds[k], built[k]);
      CHECK(S.v.size() == alone[k].v.size(), "world %d: body count", k);
      CHECK(S.contacts == alone[k].contacts, "world %d: contacts %d batched, %d alone", k, S.contacts, alone[k].contacts);
      float worst = 0.0f;
      for (size_t i = 0; i < S.v.size() && i < alone[k].v.size(); ++i) worst = std::fmax(worst, std::fabs(S.v[i] - alone[k].v[i]));
      const bool same = S.v.size() == alone[k].v.size() && memcmp(S.v.data(), alone[k].v.data(), S.v.size() * sizeof(float)) == 0;
      CHECK(same, "world %d: state after %d steps differs from the world stepped alone (max |d| %g)", k, STEPS, worst);
      for (size_t c = 0; c < at[k].size(); ++c)
        CHECK(at[k][c].v == aloneAt[k][c].v && at[k][c].contacts == aloneAt[k][c].contacts, "world %d: checkpoint %zu", k, c);
      printf("TRACE world %d bodies=%zu contacts=%d max|d|=%g hinge=%.5f\n", k, S.v.size() / 7, S.contacts, worst,
             built[k].hinge->GetJointAngle());
    }
    // a member's queries see its own world only
    CountFixtures q;
    q.world = worlds[3];
    b2AABB all;
    all.lowerBound.Set(-100.0f, -100.0f);
    all.upperBound.Set(100.0f, 100.0f);
    worlds[3]->QueryAABB(&q, all);
    CHECK(q.count == worlds[3]->GetProxyCount() && q.foreign == 0, "QueryAABB: %d fixtures, %d foreign, world has %d", q.count,
          q.foreign, worlds[3]->GetProxyCount());
    // a member cannot be stepped on its own
    const b2Vec2 before = built[0].paddle->GetPosition();
    worlds[0]->Step(dt, 8, 3);
    CHECK(built[0].paddle->GetPosition().x == before.x, "b2World::Step of a batched world is refused");
    // growth beyond a world's slot: the arena is rebuilt, the run goes on
    for (int i = 0; i < 80; ++i) {
      b2BodyDef bd;
      bd.type = b2_dynamicBody;
      bd.position.Set(-9.0f + 0.22f * (float)i, 14.0f + 0.5f * (float)(i % 3));
      b2Body* b = worlds[1]->CreateBody(&bd);
      b2CircleShape c;
      c.m_radius = 0.2f;
      b->CreateFixture(&c, 1.0f);
      built[1].bodies.push_back(b);
    }
    const float yBefore = built[2].bodies[10]->GetPosition().y;
    for (int s = 0; s < 120; ++s) batch.Step(dt, 8, 3);
    State grown = snapshot(worlds[1], built[1]);
    bool finite = true, inside = true;
    for (size_t i = 0; i < grown.v.size(); i += 7) {
      finite = finite && std::isfinite(grown.v[i]) && std::isfinite(grown.v[i + 1]);
      inside = inside && grown.v[i + 1] > -1.0f && std::fabs(grown.v[i]) < 13.0f;
    }
    CHECK(finite && inside, "after slot growth every body of world 1 is finite and inside its box");
    CHECK(std::fabs(built[2].bodies[10]->GetPosition().y - yBefore) < 1.0f, "a resting body of world 2 stayed where it was across the rebuild");
    printf("TRACE grown world 1 bodies=%zu contacts=%d\n", grown.v.size() / 7, grown.contacts);
    // a world leaves the batch by being destroyed; the others go on
    delete worlds[4];
    worlds[4] = nullptr;
    for (int s = 0; s < 10; ++s) batch.Step(dt, 8, 3);
    CHECK(batch.GetWorld(4) == nullptr && batch.GetWorld(5) == worlds[5], "destroyed member");
    CHECK(std::isfinite(built[5].paddle->GetPosition().x), "world 5 still steps");
    for (b2World* w : worlds) delete w;
  }
  if (g_failures == 0) printf("all batch checks passed\n");
  return g_failures == 0 ? 0 : 1;
}

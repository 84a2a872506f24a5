// -----------------------------------------------------------------------------------------------
// Third-party notice.  To stay source- and result-compatible with box2d-optimized, parts of this
// file restate declarations, inline math and creation-time algorithms of that library (itself a
// fork of Box2D).  Those parts are covered by the MIT License:
//   Copyright (c) 2019 Erin Catto, Copyright (c) 2020 Manolis Tsamis
// The full licence text and permission notice are in LICENSES/box2d-optimized-MIT.txt.
// -----------------------------------------------------------------------------------------------
// b2_world_host.cpp — b2World / b2Body / b2Fixture / b2Contact host handles over the C-ABI.
//
// The reference keeps all state in these objects and walks it on one CPU thread
// (src/dynamics/b2_world.cpp, b2_body.cpp, b2_fixture.cpp, b2_contact.cpp).  Here they are thin
// mirrors: creation and edits write a host copy and mark it dirty, b2World::Step uploads what
// changed and calls b2g_step(), and getters lazily pull device state back.  Semantics kept from
// the reference: silent no-op / nullptr while locked (b2_world.cpp:142-146 etc.), non-static
// bodies at the head of the body list and static ones at the tail (b2_world.cpp:153-171),
// fixture lists in reverse creation order (b2_body.cpp:244-246), mass recomputation on fixture
// changes (b2_body.cpp:352-416), process-global fixture ids (b2_fixture.cpp:40-41).
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <unordered_map>
#include <typeinfo>
#include <vector>
#include "b2cuda.h"
#include "box2d/box2d.h"

static int32 g_defaultBodies = 1 << 16, g_defaultFixtures = 1 << 16, g_defaultContacts = 1 << 18;
static int32 g_defaultDevice = -1;
static uint32 g_fixtureIdCounter = 0;

static void b2gCheck(int rc, const char* what) {
  if (rc != B2G_OK) {
    // error convention of the reference: assert in debug builds, log and carry on otherwise
    fprintf(stderr, "[b2cuda] %s failed (%d): %s\n", what, rc, b2g_last_error());
    b2Assert(false);
    abort();  // never continue on a silently wrong device state
  }
}

struct b2WorldBatchImpl;

struct b2WorldImpl {
  b2World* world = nullptr;
  b2gArena* arena = nullptr;
  // Member of a b2WorldBatch: the arena is the batch's (num_worlds > 1), this world owns the slot
  // [bodyBase, bodyBase + slot) of it (likewise fixtures and shape quads; joints are packed) and is world
  // `worldId` inside it.  Every device index of this world is host index + base.
  b2WorldBatchImpl* batch = nullptr;
  int32 worldId = 0, bodyBase = 0, fixtureBase = 0, quadBase = 0, jointBase = 0;
  int32 capBodies = 0, capFixtures = 0, capContacts = 0, capQuads = 0, capJoints = 0;
  std::vector<b2Body*> bodies;      // by device index (creation order); nullptr once destroyed
  std::vector<b2Fixture*> fixtures; // by device index
  std::vector<b2Joint*> joints;
  std::vector<float> shapePool;     // float4 records, append-only
  int32 shapesUploaded = 0;         // quads already on the device
  int32 bodiesOnDevice = 0;         // bodies the arena knows about (created-but-not-flushed ones are host only)
  int32 bodyDirtyLo = INT32_MAX, bodyDirtyHi = 0;
  int32 fixtureDirtyLo = INT32_MAX, fixtureDirtyHi = 0;
  bool jointsDirty = false;
  bool jointsStale = false;    // device has newer joint impulses than the host copies
  int32 jointsOnDevice = 0;    // joints the arena holds, in the order of `joints` at the last flush
  bool bodiesStale = false;    // device has newer body state than the host copies
  bool contactsStale = true;   // host contact list does not reflect the device
  bool profiling = false;
  bool deadFixtures = false;   // a fixture was destroyed and the device may still hold contacts naming it
  float lastInvDt = 0.0f;      // b2World::m_inv_dt0 (b2_world.cpp:1162): survives an arena re-creation
  bool arenaFresh = false;     // the arena was (re-)created since the last flush
  // Index recycling.  A destroyed body / fixture / shape record leaves a placeholder row on the device; the row is
  // reused by a later Create* — but only after a Step, whose pair refresh has retired every contact that named it.
  struct ShapeSlot { int32 off, quads; };
  std::vector<int32> freeBodies, freeFixtures, retiredBodies, retiredFixtures;
  std::vector<ShapeSlot> freeShapes, retiredShapes, dirtyShapes;
  void releaseRetired() {
    freeBodies.insert(freeBodies.end(), retiredBodies.begin(), retiredBodies.end());
    freeFixtures.insert(freeFixtures.end(), retiredFixtures.begin(), retiredFixtures.end());
    freeShapes.insert(freeShapes.end(), retiredShapes.begin(), retiredShapes.end());
    retiredBodies.clear();
    retiredFixtures.clear();
    retiredShapes.clear();
  }
  std::vector<b2Contact*> contacts;                       // current list, device order
  std::unordered_map<uint64_t, b2Contact*> contactPool;   // (fixA,fixB) -> handle, stable across steps
  b2Contact sentinel;
  std::vector<b2Contact*> graveyard;  // dead handles, kept valid until the end of the step's callbacks
  b2gStepStats lastStats;

  void touchBody(int32 i) {
    bodyDirtyLo = std::min(bodyDirtyLo, i);
    bodyDirtyHi = std::max(bodyDirtyHi, i + 1);
  }
  void touchFixture(int32 i) {
    fixtureDirtyLo = std::min(fixtureDirtyLo, i);
    fixtureDirtyHi = std::max(fixtureDirtyHi, i + 1);
  }
  static uint64_t key(int32 a, int32 b) { return ((uint64_t)(uint32)a << 32) | (uint32)b; }

  // device-only state carried across an arena re-creation (contacts keep manifolds + impulses)
  struct SavedContacts {
    std::vector<int32_t> fa, fb;
    std::vector<uint32_t> flags;
    std::vector<float> man, mat;
  } saved;
  void saveDeviceState();
  void growContacts();
  // user b2ContactFilter: pairs it rejected (fixture index pairs, sorted) and fixtures whose
  // contacts must be offered to it again (b2Fixture::Refilter)
  std::vector<std::pair<int32_t, int32_t>> vetoes;
  std::vector<int32_t> refilter;
  bool customFilter() const;
  void applyContactFilter();
  // rows of the SoA upload calls, built from the host handles
  struct BodyRows {
    std::vector<float> pos, vel, xf, mass, center, force;
    std::vector<uint32_t> flags;
    std::vector<int32_t> wid, index;
  };
  struct FixtureRows {
    std::vector<int32_t> body, off, index;
    std::vector<uint32_t> tf, filter;
    std::vector<float> mat;
  };
  void appendBodyRows(int32 lo, int32 n, BodyRows& R) const;
  void appendFixtureRows(int32 lo, int32 n, FixtureRows& R) const;
  void ensureArena();
  void flush();
  void pullBodies();
  void applyBodyRows(int32 n, const float* pos, const float* vel, const float* xf, const float* force, const uint32_t* flags);
  void pullJoints();
  void pullContacts();
  void rebuildContacts(int32 n, const int32_t* fa, const int32_t* fb, const uint32_t* flags, const float* man,
                       const float* mat, const int32_t* deviceIndex);
  b2Contact* findContact(int32 fa, int32 fb);
  void pushContactOverrides();
};

// b2WorldBatch: N b2World objects in ONE arena (include/b2cuda.h: num_worlds > 1), stepped by one device pass.
struct b2WorldBatchImpl {
  std::vector<b2WorldImpl*> members;  // by world id; nullptr once the world was destroyed
  b2gArena* arena = nullptr;
  int32 slotBodies = 0, slotFixtures = 0, slotQuads = 0, capContacts = 0, capJoints = 0, worldsOnDevice = 0;
  int32 jointsOnDevice = 0;
  float lastInvDt = 0.0f;
  bool contactsStale = true, profiling = false;
  b2gStepStats lastStats;
  b2WorldImpl::SavedContacts saved;  // arena-wide fixture indices of the arena that was torn down
  int32 savedSlotFixtures = 0;
  // members pulled one by one since the last Step; when that was many, the next step's first pull downloads
  // every member at once (an RL loop that reads or edits every environment every step)
  int32 memberPulls = 0;
  bool bulkPull = false;
  bool notePull() {
    ++memberPulls;
    return bulkPull;
  }
  void pullAllBodies();
  void ensureArena();
  void teardown();
  void growContacts();
  void flush();
  void pullContacts();
  int32 live() const {
    int32 n = 0;
    for (b2WorldImpl* m : members) n += m != nullptr;
    return n;
  }
};

void b2WorldImpl::ensureArena() {
  if (batch) {
    batch->ensureArena();
    return;
  }
  int32 needBodies = (int32)bodies.size(), needFixtures = (int32)fixtures.size();
  int32 needQuads = (int32)(shapePool.size() / 4), needJoints = (int32)joints.size();
  if (arena && needBodies <= capBodies && needFixtures <= capFixtures && needQuads <= capQuads &&
      needJoints <= capJoints)
    return;
  if (arena) {
    // grow: pull everything the host does not own (body state, joint impulses, contacts with their
    // manifolds and warm-start impulses), recreate, re-upload at the end of flush()
    saveDeviceState();
    b2gCheck(b2g_arena_destroy(arena), "b2g_arena_destroy");
    arena = nullptr;
    contactsStale = true;
  }
  auto grow = [](int32 cap, int32 need, int32 dflt) {
    int32 c = std::max(cap, dflt);
    while (c < need) c *= 2;
    return c;
  };
  capBodies = grow(capBodies, needBodies, g_defaultBodies);
  capFixtures = grow(capFixtures, needFixtures, g_defaultFixtures);
  capQuads = grow(capQuads, needQuads, std::max(g_defaultFixtures * 5, 1024));
  capContacts = std::max(capContacts, g_defaultContacts);
  while (capContacts < (int32)saved.fa.size() * 2) capContacts *= 2;
  capJoints = grow(capJoints, needJoints, 64);
  b2gArenaDef def;
  memset(&def, 0, sizeof(def));
  int dev = g_defaultDevice;
  if (dev < 0) {
    const char* e = getenv("B2G_DEVICE");
    dev = e ? atoi(e) : 0;
  }
  def.device = dev;
  def.num_worlds = 1;
  def.max_bodies = capBodies;
  def.max_fixtures = capFixtures;
  def.max_shape_quads = capQuads;
  def.max_contacts = capContacts;
  def.max_joints = capJoints;
  b2gCheck(b2g_arena_create(&def, &arena), "b2g_arena_create");
  b2g_set_profiling(arena, profiling ? 1 : 0);
  if (getenv("B2G_KERNEL_TIMING")) b2g_set_kernel_timing(arena, 1);  // diagnostics: table printed by ~b2World
  shapesUploaded = 0;
  bodiesOnDevice = 0;
  bodyDirtyLo = 0;
  bodyDirtyHi = needBodies;
  fixtureDirtyLo = 0;
  fixtureDirtyHi = needFixtures;
  jointsDirty = needJoints > 0;
  jointsOnDevice = 0;
  arenaFresh = true;
  world->m_newContacts = true;
}

bool b2WorldImpl::customFilter() const {
  b2ContactFilter* F = world->m_contactFilter;
  return F != nullptr && typeid(*F) != typeid(b2ContactFilter);
}

// b2ContactFilter::ShouldCollide for a user subclass (b2_contact_manager.cpp:163-170, 84-96).  The
// device applies the default rule and inserts every pair that passes it; after each pair refresh
// the host offers the newly inserted pairs (and, as the reference does every step, the rejected
// pairs that still overlap, and the contacts of re-filtered fixtures) to the user's callback and
// hands the rejected ones back as a veto list.  Their contacts are removed before any narrowphase.
void b2WorldImpl::applyContactFilter() {
  if (!arena) return;
  if (!customFilter()) {
    if (!vetoes.empty()) {
      vetoes.clear();
      b2gCheck(b2g_set_pair_vetoes(arena, 0, nullptr, nullptr), "b2g_set_pair_vetoes");
      world->m_newContacts = true;
    }
    refilter.clear();
    return;
  }
  b2ContactFilter* F = world->m_contactFilter;
  auto alive = [&](int32_t f) { return f >= 0 && f < (int32_t)fixtures.size() && fixtures[f] != nullptr; };
  bool changed = false, released = false;
  std::vector<std::pair<int32_t, int32_t>> next;
  if (!vetoes.empty()) {
    std::vector<uint8_t> seen(vetoes.size());
    b2gCheck(b2g_download_veto_seen(arena, (int32_t)vetoes.size(), seen.data()), "b2g_download_veto_seen");
    for (size_t i = 0; i < vetoes.size(); ++i) {
      auto v = vetoes[i];
      if (seen[i] && alive(v.first) && alive(v.second) && !F->ShouldCollide(fixtures[v.first], fixtures[v.second])) {
        next.push_back(v);  // still overlapping, still rejected
      } else {
        changed = true;     // separated, destroyed, or accepted now: the pair finder may report it again
        released = released || seen[i];
      }
    }
  }
  int32_t nNew = 0;
  b2gCheck(b2g_download_new_pairs(arena, 0, nullptr, nullptr, &nNew), "b2g_download_new_pairs");
  if (nNew > 0) {
    std::vector<int32_t> fa(nNew), fb(nNew);
    b2gCheck(b2g_download_new_pairs(arena, nNew, fa.data(), fb.data(), &nNew), "b2g_download_new_pairs");
    for (int32_t i = 0; i < nNew; ++i) {
      if (!alive(fa[i]) || !alive(fb[i])) continue;
      if (!F->ShouldCollide(fixtures[fa[i]], fixtures[fb[i]])) {
        next.emplace_back(std::min(fa[i], fb[i]), std::max(fa[i], fb[i]));
        changed = true;
      }
    }
  }
  if (!refilter.empty()) {
    // b2Fixture::Refilter: every contact of the fixture is offered to the filter again
    contactsStale = true;
    pullContacts();
    std::sort(refilter.begin(), refilter.end());
    for (b2Contact* c : contacts) {
      int32_t a = c->m_fixtureA->m_index, b = c->m_fixtureB->m_index;
      if (!std::binary_search(refilter.begin(), refilter.end(), a) && !std::binary_search(refilter.begin(), refilter.end(), b))
        continue;
      if (!F->ShouldCollide(c->m_fixtureA, c->m_fixtureB)) {
        next.emplace_back(std::min(a, b), std::max(a, b));
        changed = true;
      }
    }
    refilter.clear();
  }
  if (!changed) return;
  std::sort(next.begin(), next.end());
  next.erase(std::unique(next.begin(), next.end()), next.end());
  vetoes.swap(next);
  std::vector<int32_t> fa(vetoes.size()), fb(vetoes.size());
  for (size_t i = 0; i < vetoes.size(); ++i) {
    fa[i] = vetoes[i].first;
    fb[i] = vetoes[i].second;
  }
  b2gCheck(b2g_set_pair_vetoes(arena, (int32_t)vetoes.size(), fa.data(), fb.data()), "b2g_set_pair_vetoes");
  contactsStale = true;
  if (released) world->m_newContacts = true;  // a released pair gets its contact at the next step's pair refresh
}

void b2WorldImpl::saveDeviceState() {
  pullBodies();
  pullJoints();
  int32 n = 0;
  b2gCheck(b2g_contact_count(arena, &n), "b2g_contact_count");
  saved.fa.assign(n, 0);
  saved.fb.assign(n, 0);
  saved.flags.assign(n, 0);
  saved.man.assign((size_t)n * 16, 0.0f);
  saved.mat.assign((size_t)n * 4, 0.0f);
  if (n > 0) {
    b2gContactArrays a;
    memset(&a, 0, sizeof(a));
    a.fixture_a = saved.fa.data(); a.fixture_b = saved.fb.data(); a.flags = saved.flags.data();
    a.manifold = saved.man.data(); a.material = saved.mat.data();
    b2gCheck(b2g_download_contacts(arena, 0, n, &a), "b2g_download_contacts");
  }
}

// max_contacts was too small for the pairs found at the end of the step that just ran (the step
// itself is complete): double it, keeping every contact; the next Step's pair refresh inserts the
// pairs that did not fit (include/b2cuda.h, "B2G_ERR_CAPACITY from a step")
void b2WorldImpl::growContacts() {
  if (batch) {
    batch->growContacts();
    return;
  }
  bodiesStale = true;
  jointsStale = true;
  saveDeviceState();
  b2gCheck(b2g_arena_destroy(arena), "b2g_arena_destroy");
  arena = nullptr;
  capContacts *= 2;
  fprintf(stderr, "[b2cuda] contact capacity grown to %d\n", capContacts);
  flush();
  contactsStale = true;
}

void b2WorldImpl::appendBodyRows(int32 lo, int32 n, BodyRows& R) const {
  const size_t at = R.flags.size();
  for (std::vector<float>* v : {&R.pos, &R.vel, &R.xf, &R.mass, &R.center, &R.force}) v->resize((at + (size_t)n) * 4, 0.0f);
  R.flags.resize(at + n, 0u);
  R.wid.resize(at + n, worldId);
  R.index.resize(at + n);
  for (int32 k = 0; k < n; ++k) {
    const size_t r = at + (size_t)k;
    R.index[r] = bodyBase + lo + k;
    b2Body* b = bodies[lo + k];
    if (!b) continue;  // destroyed: a disabled static placeholder (flags 0)
    b->m_onDevice = true;
    float* p = &R.pos[r * 4];
    p[0] = b->m_sweep.c.x; p[1] = b->m_sweep.c.y; p[2] = b->m_sweep.a; p[3] = 0.0f;
    float* v = &R.vel[r * 4];
    v[0] = b->m_linearVelocity.x; v[1] = b->m_linearVelocity.y; v[2] = b->m_angularVelocity; v[3] = 0.0f;
    float* x = &R.xf[r * 4];
    x[0] = b->m_xf.p.x; x[1] = b->m_xf.p.y; x[2] = b->m_xf.q.s; x[3] = b->m_xf.q.c;
    float* m = &R.mass[r * 4];
    m[0] = b->m_invMass; m[1] = b->m_invI; m[2] = b->m_mass; m[3] = b->m_gravityScale;
    float* c = &R.center[r * 4];
    c[0] = b->m_sweep.localCenter.x; c[1] = b->m_sweep.localCenter.y; c[2] = b->m_linearDamping;
    c[3] = b->m_angularDamping;
    float* f = &R.force[r * 4];
    f[0] = b->m_force.x; f[1] = b->m_force.y; f[2] = b->m_torque; f[3] = b->m_sleepTime;
    R.flags[r] = (uint32_t)(b->m_flags & ~b2Body::e_islandFlag) | ((uint32_t)b->m_type << B2G_BODY_TYPE_SHIFT);
  }
}

void b2WorldImpl::appendFixtureRows(int32 lo, int32 n, FixtureRows& R) const {
  const size_t at = R.body.size();
  R.body.resize(at + n);
  R.off.resize(at + n);
  R.index.resize(at + n);
  R.tf.resize(at + n);
  R.filter.resize((at + (size_t)n) * 2, 0u);
  R.mat.resize((at + (size_t)n) * 4, 0.0f);
  for (int32 k = 0; k < n; ++k) {
    const size_t r = at + (size_t)k;
    R.index[r] = fixtureBase + lo + k;
    b2Fixture* f = fixtures[lo + k];
    if (!f) {
      R.body[r] = bodyBase; R.off[r] = quadBase; R.tf[r] = B2G_FIX_DEAD;
      continue;
    }
    R.body[r] = bodyBase + f->m_body->m_index;
    R.off[r] = quadBase + f->m_shapeOff;
    R.tf[r] = (uint32_t)f->m_shape->GetType() | (f->m_isSensor ? B2G_FIX_SENSOR : 0u);
    R.filter[r * 2] = (uint32_t)f->m_filter.categoryBits | ((uint32_t)f->m_filter.maskBits << 16);
    R.filter[r * 2 + 1] = (uint32_t)(int32_t)f->m_filter.groupIndex;
    R.mat[r * 4] = f->m_friction; R.mat[r * 4 + 1] = f->m_restitution;
    R.mat[r * 4 + 2] = f->m_restitutionThreshold; R.mat[r * 4 + 3] = f->m_density;
  }
}

// upload everything the host changed since the last step
void b2WorldImpl::flush() {
  ensureArena();
  int32 nq = (int32)(shapePool.size() / 4);
  if (nq > shapesUploaded) {
    b2gCheck(b2g_upload_shapes(arena, quadBase + shapesUploaded, nq - shapesUploaded,
                               shapePool.data() + (size_t)shapesUploaded * 4),
             "b2g_upload_shapes");
    shapesUploaded = nq;
  }
  for (const ShapeSlot& d : dirtyShapes)
    b2gCheck(b2g_upload_shapes(arena, quadBase + d.off, d.quads, shapePool.data() + (size_t)d.off * 4), "b2g_upload_shapes");
  dirtyShapes.clear();
  if (bodyDirtyHi > bodyDirtyLo) {
    int32 lo = bodyDirtyLo, n = bodyDirtyHi - bodyDirtyLo;
    BodyRows R;
    appendBodyRows(lo, n, R);
    b2gBodyArrays a;
    a.pos = R.pos.data(); a.vel = R.vel.data(); a.xf = R.xf.data(); a.mass = R.mass.data(); a.center = R.center.data();
    a.force = R.force.data(); a.flags = R.flags.data(); a.world = R.wid.data();
    b2gCheck(b2g_upload_bodies(arena, bodyBase + lo, n, &a), "b2g_upload_bodies");
    bodiesOnDevice = std::max(bodiesOnDevice, lo + n);
    bodyDirtyLo = INT32_MAX;
    bodyDirtyHi = 0;
  }
  if (fixtureDirtyHi > fixtureDirtyLo) {
    int32 lo = fixtureDirtyLo, n = fixtureDirtyHi - fixtureDirtyLo;
    FixtureRows R;
    appendFixtureRows(lo, n, R);
    b2gFixtureArrays a;
    a.body = R.body.data(); a.shape_off = R.off.data(); a.type_flags = R.tf.data(); a.filter = R.filter.data();
    a.material = R.mat.data();
    b2gCheck(b2g_upload_fixtures(arena, fixtureBase + lo, n, &a), "b2g_upload_fixtures");
    fixtureDirtyLo = INT32_MAX;
    fixtureDirtyHi = 0;
  }
  if (jointsDirty && !batch) {  // (a batch packs its members' joints into one table: b2WorldBatchImpl::flush)
    int32 n = (int32)joints.size();
    std::vector<int32_t> jb((size_t)n * 2);
    std::vector<float> anchors((size_t)n * 4), params((size_t)n * 12, 0.0f), state((size_t)n * 5);
    for (int32 k = 0; k < n; ++k) {
      b2Joint* j = joints[k];
      jb[(size_t)k * 2] = j->m_bodyA->m_index;
      jb[(size_t)k * 2 + 1] = j->m_bodyB->m_index;
      j->WriteDevice(&anchors[(size_t)k * 4], &params[(size_t)k * 12], &state[(size_t)k * 5]);
    }
    b2gJointArrays a;
    a.bodies = jb.data(); a.anchors = anchors.data(); a.params = params.data(); a.state = state.data();
    if (n > 0) b2gCheck(b2g_upload_joints(arena, 0, n, &a), "b2g_upload_joints");
    jointsOnDevice = n;
    jointsDirty = false;
  }
  if (!saved.fa.empty() && !batch) {
    // contacts carried over from the previous arena; pairs whose fixture died meanwhile are retired
    // by the pair refresh that starts the next step
    b2gContactArrays a;
    memset(&a, 0, sizeof(a));
    a.fixture_a = saved.fa.data(); a.fixture_b = saved.fb.data(); a.flags = saved.flags.data();
    a.manifold = saved.man.data(); a.material = saved.mat.data();
    b2gCheck(b2g_upload_contacts(arena, (int32_t)saved.fa.size(), &a), "b2g_upload_contacts");
    saved = SavedContacts();
    world->m_newContacts = true;
  }
  if (arenaFresh && !batch) {
    // warm starting scales last step's impulses by dt * inv_dt0 (b2_world.cpp:1130): a new arena starts at
    // the world's inv_dt0, not at zero, or the first step after any growth would drop every impulse
    b2gCheck(b2g_set_inv_dt0(arena, lastInvDt), "b2g_set_inv_dt0");
    arenaFresh = false;
  }
}

// every site that edits the joint table calls this first, so `joints` and the device agree on order
void b2WorldImpl::pullJoints() {
  if (!jointsStale || !arena) return;
  jointsStale = false;
  int32 n = std::min((int32)joints.size(), jointsOnDevice);
  if (n == 0) return;
  std::vector<float> state((size_t)n * 5);
  b2gCheck(b2g_download_joints(arena, jointBase, n, state.data()), "b2g_download_joints");
  for (int32 k = 0; k < n; ++k) joints[k]->ReadDeviceState(&state[(size_t)k * 5]);
}

void b2WorldImpl::pullBodies() {
  if (!bodiesStale || !arena) return;
  if (batch && batch->notePull()) {
    batch->pullAllBodies();  // many members are being read every step: one download for all of them
    return;
  }
  bodiesStale = false;
  int32 n = std::min((int32)bodies.size(), bodiesOnDevice);
  if (n == 0) return;
  std::vector<float> pos((size_t)n * 4), vel((size_t)n * 4), xf((size_t)n * 4), force((size_t)n * 4);
  std::vector<uint32_t> flags(n);
  b2gBodyArrays a;
  memset(&a, 0, sizeof(a));
  a.pos = pos.data(); a.vel = vel.data(); a.xf = xf.data(); a.force = force.data(); a.flags = flags.data();
  b2gCheck(b2g_download_bodies(arena, bodyBase, n, &a), "b2g_download_bodies");
  applyBodyRows(n, pos.data(), vel.data(), xf.data(), force.data(), flags.data());
}

// downloaded device state of this world's first n bodies into the host handles
void b2WorldImpl::applyBodyRows(int32 n, const float* pos, const float* vel, const float* xf, const float* force,
                                const uint32_t* flags) {
  for (int32 i = 0; i < n; ++i) {
    b2Body* b = bodies[i];
    if (!b || !b->m_onDevice) continue;  // (a body created on a reused index: the device row is still the placeholder)
    b->m_sweep.c.Set(pos[(size_t)i * 4], pos[(size_t)i * 4 + 1]);
    b->m_sweep.a = pos[(size_t)i * 4 + 2];
    b->m_sweep.c0 = b->m_sweep.c;
    b->m_sweep.a0 = b->m_sweep.a;
    b->m_linearVelocity.Set(vel[(size_t)i * 4], vel[(size_t)i * 4 + 1]);
    b->m_angularVelocity = vel[(size_t)i * 4 + 2];
    b->m_xf.p.Set(xf[(size_t)i * 4], xf[(size_t)i * 4 + 1]);
    b->m_xf.q.s = xf[(size_t)i * 4 + 2];
    b->m_xf.q.c = xf[(size_t)i * 4 + 3];
    b->m_force.Set(force[(size_t)i * 4], force[(size_t)i * 4 + 1]);
    b->m_torque = force[(size_t)i * 4 + 2];
    b->m_sleepTime = force[(size_t)i * 4 + 3];
    // the device's pending SetAwake(true) (set by Collide, consumed by Solve) rides along, so that an upload
    // of this body between the two halves of a step does not drop it
    uint16 keep = b->m_flags & ~(uint16)(b2Body::e_awakeFlag | B2G_BODY_WAKE_REQUEST);
    b->m_flags = keep | (uint16)(flags[i] & (B2G_BODY_AWAKE | B2G_BODY_WAKE_REQUEST));
  }
}

static void unpackManifold(b2Manifold& m, const float* q) {
  m.localNormal.Set(q[0], q[1]);
  m.localPoint.Set(q[2], q[3]);
  for (int k = 0; k < 2; ++k) {
    m.points[k].localPoint.Set(q[4 + 4 * k], q[5 + 4 * k]);
    m.points[k].normalImpulse = q[6 + 4 * k];
    m.points[k].tangentImpulse = q[7 + 4 * k];
    memcpy(&m.points[k].id.key, &q[12 + k], 4);
  }
  int32 type, count;
  memcpy(&type, &q[14], 4);
  memcpy(&count, &q[15], 4);
  m.type = (b2Manifold::Type)type;
  m.pointCount = count;
}

// rebuild the host contact list (b2World::GetContactListStart, b2Body::GetContact) from the device
void b2WorldImpl::pullContacts() {
  if (batch) {
    batch->pullContacts();
    return;
  }
  if (!contactsStale) return;
  contactsStale = false;
  if (deadFixtures && arena && !world->m_locked) {
    // b2Body::DestroyFixture / b2World::DestroyBody remove the fixture's contacts at once in the reference
    // (b2_body.cpp:244-259).  Here they die with the next pair refresh, so run it now: the list handed out
    // must never name a destroyed fixture.
    deadFixtures = false;
    flush();
    int rc = b2g_find_new_contacts(arena);
    for (int attempt = 0; rc == B2G_ERR_CAPACITY && attempt < 8; ++attempt) {
      growContacts();
      rc = b2g_find_new_contacts(arena);
    }
    b2gCheck(rc, "b2g_find_new_contacts");
    world->m_newContacts = false;
    contactsStale = false;
    applyContactFilter();
    contactsStale = false;
  }
  for (b2Body* b : bodies)
    if (b) b->m_contacts.clear();
  int32 n = 0;
  if (arena) b2gCheck(b2g_contact_count(arena, &n), "b2g_contact_count");
  std::vector<int32_t> fa(n), fb(n);
  std::vector<uint32_t> flags(n);
  std::vector<float> man((size_t)n * 16), mat((size_t)n * 4);
  if (n > 0) {
    b2gContactArrays a;
    memset(&a, 0, sizeof(a));
    a.fixture_a = fa.data(); a.fixture_b = fb.data(); a.flags = flags.data(); a.manifold = man.data();
    a.material = mat.data();
    b2gCheck(b2g_download_contacts(arena, 0, n, &a), "b2g_download_contacts");
  }
  rebuildContacts(n, fa.data(), fb.data(), flags.data(), man.data(), mat.data(), nullptr);
}

// host contact handles from downloaded records (fixture indices local to this world)
void b2WorldImpl::rebuildContacts(int32 n, const int32_t* fa, const int32_t* fb, const uint32_t* flags, const float* man,
                                  const float* mat, const int32_t* deviceIndex) {
  for (b2Body* b : bodies)
    if (b) b->m_contacts.clear();
  std::unordered_map<uint64_t, b2Contact*> next;
  next.reserve((size_t)n * 2 + 1);
  contacts.assign(n, nullptr);
  for (int32 i = 0; i < n; ++i) {
    uint64_t k = key(fa[i], fb[i]);
    b2Contact* c;
    auto it = contactPool.find(k);
    if (it != contactPool.end()) {
      c = it->second;
      contactPool.erase(it);
    } else {
      c = new b2Contact();
      c->m_world = world;
      c->m_fixtureA = fixtures[fa[i]];
      c->m_fixtureB = fixtures[fb[i]];
    }
    c->m_flags = flags[i];
    unpackManifold(c->m_manifold, &man[(size_t)i * 16]);
    c->m_friction = mat[(size_t)i * 4];
    c->m_restitution = mat[(size_t)i * 4 + 1];
    c->m_restitutionThreshold = mat[(size_t)i * 4 + 2];
    c->m_tangentSpeed = mat[(size_t)i * 4 + 3];
    c->m_deviceIndex = deviceIndex ? deviceIndex[i] : i;
    c->m_overridden = false;
    contacts[i] = c;
    next.emplace(k, c);
    c->m_fixtureA->m_body->m_contacts.push_back(c);
    c->m_fixtureB->m_body->m_contacts.push_back(c);
  }
  for (auto& kv : contactPool) graveyard.push_back(kv.second);  // contacts that died: freed by Step
  contactPool.swap(next);
  // ring: sentinel <-> c0 <-> c1 ... <-> sentinel (b2_contact_manager.h: Start()/End())
  b2Contact* prev = &sentinel;
  for (b2Contact* c : contacts) {
    prev->m_next = c;
    c->m_prev = prev;
    prev = c;
  }
  prev->m_next = &sentinel;
  sentinel.m_prev = prev;
}

b2Contact* b2WorldImpl::findContact(int32 fa, int32 fb) {
  auto it = contactPool.find(key(fa, fb));
  return it == contactPool.end() ? nullptr : it->second;
}

void b2WorldImpl::pushContactOverrides() {
  int32 n = (int32)contacts.size();
  int32 lo = n, hi = 0;
  for (int32 i = 0; i < n; ++i)
    if (contacts[i]->m_overridden) {
      lo = std::min(lo, i);
      hi = std::max(hi, i + 1);
    }
  if (hi <= lo) return;
  std::vector<uint32_t> flags(hi - lo);
  std::vector<float> mat((size_t)(hi - lo) * 4);
  for (int32 i = lo; i < hi; ++i) {
    b2Contact* c = contacts[i];
    flags[i - lo] = c->m_flags;
    float* m = &mat[(size_t)(i - lo) * 4];
    m[0] = c->m_friction; m[1] = c->m_restitution; m[2] = c->m_restitutionThreshold; m[3] = c->m_tangentSpeed;
    c->m_overridden = false;
  }
  b2gCheck(b2g_upload_contact_overrides(arena, lo, hi - lo, flags.data(), mat.data()), "b2g_upload_contact_overrides");
}

// ================================================================================================
// b2World
// ================================================================================================
void b2World::SetDefaultCapacity(int32 bodies, int32 fixtures, int32 contacts) {
  g_defaultBodies = std::max(bodies, 16);
  g_defaultFixtures = std::max(fixtures, 16);
  g_defaultContacts = std::max(contacts, 64);
}
void b2World::SetDefaultDevice(int32 device) { g_defaultDevice = device; }

b2World::b2World(const b2Vec2& gravity) {
  m_impl = new b2WorldImpl();
  m_impl->world = this;
  memset(&m_impl->lastStats, 0, sizeof(b2gStepStats));
  m_impl->sentinel.m_next = m_impl->sentinel.m_prev = &m_impl->sentinel;
  m_impl->sentinel.m_fixtureA = m_impl->sentinel.m_fixtureB = nullptr;
  m_bodyListHead = m_bodyListTail = nullptr;
  m_jointList = nullptr;
  m_bodyCount = 0;
  m_jointCount = 0;
  m_gravity = gravity;
  m_allowSleep = true;
  m_destructionListener = nullptr;
  m_contactFilter = nullptr;
  m_contactListener = nullptr;
  m_newContacts = false;
  m_locked = false;
  m_clearForces = true;
  m_warmStarting = true;
  m_continuousPhysics = true;
  m_subStepping = false;
  m_solverMode = B2G_SOLVER_COLOURED;
  memset(&m_profile, 0, sizeof(m_profile));
}

b2World::~b2World() {
  for (b2Body* b : m_impl->bodies) {
    if (!b) continue;
    b2Fixture* f = b->m_fixtureList;
    while (f) {
      b2Fixture* nx = f->m_next;
      delete f->m_shape;
      delete f;
      f = nx;
    }
    delete b;
  }
  for (b2Joint* j : m_impl->joints) delete j;
  for (auto& kv : m_impl->contactPool) delete kv.second;
  for (b2Contact* c : m_impl->graveyard) delete c;
  if (m_impl->batch) {
    // leave the batch: the slot stays behind as disabled placeholders
    b2WorldBatchImpl* B = m_impl->batch;
    for (int32 i = 0; i < (int32)m_impl->bodies.size(); ++i) {
      m_impl->bodies[i] = nullptr;
      m_impl->touchBody(i);
    }
    for (int32 i = 0; i < (int32)m_impl->fixtures.size(); ++i) {
      m_impl->fixtures[i] = nullptr;
      m_impl->touchFixture(i);
    }
    m_impl->joints.clear();
    m_impl->jointsDirty = true;
    if (m_impl->arena) {
      m_impl->flush();
      B->flush();
    }
    B->members[m_impl->worldId] = nullptr;
    B->contactsStale = true;
    m_impl->arena = nullptr;
  }
  if (m_impl->arena && getenv("B2G_KERNEL_TIMING")) {
    b2g_synchronize(m_impl->arena);
    fprintf(stderr, "[b2cuda] kernel classes of this world (since its last arena re-creation):\n");
    for (int c = 0; c < b2g_kernel_class_count(); ++c) {
      double ms = 0.0, units = 0.0;
      int64_t launches = 0;
      if (b2g_get_kernel_timing(m_impl->arena, c, &ms, &launches, &units) == B2G_OK && launches > 0)
        fprintf(stderr, "[b2cuda]   %-16s %10.2f ms %8lld launches %10.2f us/launch\n", b2g_kernel_class_name(c), ms,
                (long long)launches, 1000.0 * ms / (double)launches);
    }
  }
  if (m_impl->arena) b2g_arena_destroy(m_impl->arena);
  delete m_impl;
}

void b2World::SetProfiling(bool flag) {
  m_impl->profiling = flag;
  if (m_impl->arena) b2g_set_profiling(m_impl->arena, flag ? 1 : 0);
}

void b2World::SetAllowSleeping(bool flag) {
  if (flag == m_allowSleep) return;
  m_allowSleep = flag;
  if (!m_allowSleep)
    for (b2Body* b = m_bodyListHead; b; b = b->m_next) b->SetAwake(true);
}

b2Body* b2World::CreateBody(const b2BodyDef* def) {
  if (IsLocked()) return nullptr;
  b2Body* b = new b2Body(def, this);
  b->m_onDevice = false;
  if (!m_impl->freeBodies.empty()) {
    b->m_index = m_impl->freeBodies.back();
    m_impl->freeBodies.pop_back();
    m_impl->bodies[b->m_index] = b;
  } else {
    b->m_index = (int32)m_impl->bodies.size();
    m_impl->bodies.push_back(b);
  }
  m_impl->touchBody(b->m_index);
  // static bodies go to the tail, everything else to the head (b2_world.cpp:153-171)
  if (m_bodyListHead == nullptr) {
    m_bodyListHead = m_bodyListTail = b;
  } else if (def->type == b2_staticBody) {
    b->m_prev = m_bodyListTail;
    m_bodyListTail->m_next = b;
    m_bodyListTail = b;
  } else {
    b->m_next = m_bodyListHead;
    m_bodyListHead->m_prev = b;
    m_bodyListHead = b;
  }
  ++m_bodyCount;
  return b;
}

void b2World::DestroyBody(b2Body* b) {
  if (IsLocked() || !b) return;
  m_impl->pullBodies();
  // joints attached to the body go first (b2_world.cpp:190-204)
  b2JointEdge* je = b->m_jointList;
  while (je) {
    b2JointEdge* je0 = je;
    je = je->next;
    if (m_destructionListener) m_destructionListener->SayGoodbye(je0->joint);
    DestroyJoint(je0->joint);
  }
  b2Fixture* f = b->m_fixtureList;
  while (f) {
    b2Fixture* nx = f->m_next;
    if (m_destructionListener) m_destructionListener->SayGoodbye(f);
    m_impl->fixtures[f->m_index] = nullptr;
    m_impl->deadFixtures = true;
    m_impl->touchFixture(f->m_index);
    m_impl->retiredFixtures.push_back(f->m_index);
    m_impl->retiredShapes.push_back({f->m_shapeOff, f->m_shape->DeviceQuadCount()});
    delete f->m_shape;
    delete f;
    f = nx;
  }
  if (b->m_prev) b->m_prev->m_next = b->m_next;
  if (b->m_next) b->m_next->m_prev = b->m_prev;
  if (b == m_bodyListHead) m_bodyListHead = b->m_next;
  if (b == m_bodyListTail) m_bodyListTail = b->m_prev;
  m_impl->bodies[b->m_index] = nullptr;
  m_impl->touchBody(b->m_index);
  m_impl->retiredBodies.push_back(b->m_index);
  // host contact handles may point at the dead fixtures: drop them all, they are rebuilt lazily
  for (auto& kv : m_impl->contactPool) delete kv.second;
  m_impl->contactPool.clear();
  m_impl->contacts.clear();
  m_impl->contactsStale = true;
  m_newContacts = true;
  --m_bodyCount;
  delete b;
}

b2Joint* b2World::CreateJoint(const b2JointDef* def) {
  if (IsLocked()) return nullptr;
  if (def->type != e_revoluteJoint && def->type != e_distanceJoint && def->type != e_weldJoint &&
      def->type != e_prismaticJoint && def->type != e_wheelJoint && def->type != e_frictionJoint &&
      def->type != e_motorJoint && def->type != e_mouseJoint) {
    fprintf(stderr, "[b2cuda] only revolute, distance, weld, prismatic, wheel, friction, motor and mouse joints run on the device (SURVEY.md §8f); joint type %d ignored\n",
            (int)def->type);
    return nullptr;
  }
  m_impl->pullJoints();
  b2Joint* j;
  if (def->type == e_revoluteJoint) j = new b2RevoluteJoint(static_cast<const b2RevoluteJointDef*>(def));
  else if (def->type == e_distanceJoint) j = new b2DistanceJoint(static_cast<const b2DistanceJointDef*>(def));
  else if (def->type == e_prismaticJoint) j = new b2PrismaticJoint(static_cast<const b2PrismaticJointDef*>(def));
  else if (def->type == e_wheelJoint) j = new b2WheelJoint(static_cast<const b2WheelJointDef*>(def));
  else if (def->type == e_frictionJoint) j = new b2FrictionJoint(static_cast<const b2FrictionJointDef*>(def));
  else if (def->type == e_motorJoint) j = new b2MotorJoint(static_cast<const b2MotorJointDef*>(def));
  else if (def->type == e_mouseJoint) j = new b2MouseJoint(static_cast<const b2MouseJointDef*>(def));
  else j = new b2WeldJoint(static_cast<const b2WeldJointDef*>(def));
  j->m_index = (int32)m_impl->joints.size();
  m_impl->joints.push_back(j);
  m_impl->jointsDirty = true;
  j->m_prev = nullptr;
  j->m_next = m_jointList;
  if (m_jointList) m_jointList->m_prev = j;
  m_jointList = j;
  ++m_jointCount;
  // connect to the bodies' joint lists (b2_world.cpp:291-305)
  j->m_edgeA.joint = j;
  j->m_edgeA.other = j->m_bodyB;
  j->m_edgeA.prev = nullptr;
  j->m_edgeA.next = j->m_bodyA->m_jointList;
  if (j->m_bodyA->m_jointList) j->m_bodyA->m_jointList->prev = &j->m_edgeA;
  j->m_bodyA->m_jointList = &j->m_edgeA;
  j->m_edgeB.joint = j;
  j->m_edgeB.other = j->m_bodyA;
  j->m_edgeB.prev = nullptr;
  j->m_edgeB.next = j->m_bodyB->m_jointList;
  if (j->m_bodyB->m_jointList) j->m_bodyB->m_jointList->prev = &j->m_edgeB;
  j->m_bodyB->m_jointList = &j->m_edgeB;
  return j;
}

// ---- spatial queries (b2_world.cpp:1193-1246) on the device tree -------------------------------
void b2World::ShiftOrigin(const b2Vec2& newOrigin) {
  if (IsLocked()) return;
  m_impl->pullBodies();
  for (b2Body* b = m_bodyListHead; b; b = b->m_next) {
    b->m_xf.p -= newOrigin;
    b->m_sweep.c0 -= newOrigin;
    b->m_sweep.c -= newOrigin;
    b->UpdateAABBs();
    m_impl->touchBody(b->m_index);
  }
  // joints hold local anchors only, except the mouse joint's world target (b2_mouse_joint.cpp:187-190)
  for (b2Joint* j : m_impl->joints) {
    if (j && j->GetType() == e_mouseJoint) {
      m_impl->pullJoints();
      static_cast<b2MouseJoint*>(j)->m_targetA -= newOrigin;
      m_impl->jointsDirty = true;
    }
  }
  m_newContacts = true;
}

void b2World::QueryAABB(b2QueryCallback* callback, const b2AABB& aabb) {
  b2WorldImpl* I = m_impl;
  if (!callback) return;
  I->flush();
  if (I->arena == nullptr || I->fixtures.empty()) return;
  const int32 nf = (int32)I->fixtures.size();
  std::vector<int32_t> found((size_t)nf);
  int32_t count = 0;
  const float box[4] = {aabb.lowerBound.x, aabb.lowerBound.y, aabb.upperBound.x, aabb.upperBound.y};
  const int32_t wid = I->worldId;
  b2gCheck(b2g_query_aabb(I->arena, 1, box, I->batch ? &wid : nullptr, nf, &count, found.data(), 0), "b2g_query_aabb");
  count = std::min(count, nf);
  std::sort(found.begin(), found.begin() + count);
  for (int32 k = 0; k < count; ++k) {
    b2Fixture* f = I->fixtures[found[k] - I->fixtureBase];
    if (f && !callback->ReportFixture(f)) return;
  }
}

void b2World::RayCast(b2RayCastCallback* callback, const b2Vec2& point1, const b2Vec2& point2) {
  b2WorldImpl* I = m_impl;
  if (!callback) return;
  I->flush();
  if (I->arena == nullptr || I->fixtures.empty()) return;
  const int32 nf = (int32)I->fixtures.size();
  std::vector<int32_t> fix((size_t)nf);
  std::vector<float> frac((size_t)nf), nrm((size_t)nf * 2);
  int32_t count = 0;
  const float ray[4] = {point1.x, point1.y, point2.x, point2.y};
  const int32_t wid = I->worldId;
  b2gCheck(b2g_ray_cast_all(I->arena, 1, ray, nullptr, I->batch ? &wid : nullptr, 0xFFFFu, nf, &count, fix.data(),
                            frac.data(), nrm.data(), 0),
           "b2g_ray_cast_all");
  count = std::min(count, nf);
  std::vector<int32> order((size_t)count);
  for (int32 k = 0; k < count; ++k) order[k] = k;
  std::sort(order.begin(), order.end(), [&](int32 a, int32 b) {
    return frac[a] < frac[b] || (frac[a] == frac[b] && fix[a] < fix[b]);
  });
  // replay into the user's callback, nearest first, honouring its clip value (b2_broad_phase.h:690-707)
  float maxFraction = 1.0f;
  for (int32 k : order) {
    if (frac[k] > maxFraction) break;
    b2Fixture* f = I->fixtures[fix[k] - I->fixtureBase];
    if (!f) continue;
    const float fraction = frac[k];
    const b2Vec2 point = (1.0f - fraction) * point1 + fraction * point2;
    const float value = callback->ReportFixture(f, point, b2Vec2(nrm[2 * k], nrm[2 * k + 1]), fraction);
    if (value == 0.0f) return;
    if (value > 0.0f) maxFraction = value;
  }
}

void b2World::DestroyJoint(b2Joint* j) {
  if (IsLocked() || !j) return;
  m_impl->pullJoints();
  if (j->m_prev) j->m_prev->m_next = j->m_next;
  if (j->m_next) j->m_next->m_prev = j->m_prev;
  if (j == m_jointList) m_jointList = j->m_next;
  b2Body* bodyA = j->m_bodyA;
  b2Body* bodyB = j->m_bodyB;
  bodyA->SetAwake(true);
  bodyB->SetAwake(true);
  if (j->m_edgeA.prev) j->m_edgeA.prev->next = j->m_edgeA.next;
  if (j->m_edgeA.next) j->m_edgeA.next->prev = j->m_edgeA.prev;
  if (&j->m_edgeA == bodyA->m_jointList) bodyA->m_jointList = j->m_edgeA.next;
  if (j->m_edgeB.prev) j->m_edgeB.prev->next = j->m_edgeB.next;
  if (j->m_edgeB.next) j->m_edgeB.next->prev = j->m_edgeB.prev;
  if (&j->m_edgeB == bodyB->m_jointList) bodyB->m_jointList = j->m_edgeB.next;
  auto& js = m_impl->joints;
  js.erase(std::find(js.begin(), js.end(), j));
  for (int32 i = 0; i < (int32)js.size(); ++i) js[i]->m_index = i;
  m_impl->jointsDirty = true;
  if (m_impl->arena && !m_impl->batch)  // (a batch re-packs its joint table at the next flush)
    b2g_set_counts(m_impl->arena, (int32)m_impl->bodies.size(), (int32)m_impl->fixtures.size(), (int32)js.size());
  --m_jointCount;
  delete j;
}

void b2World::Step(float dt, int32 velocityIterations, int32 positionIterations, int32 particleIterations) {
  B2_NOT_USED(particleIterations);
  if (m_locked) return;
  if (m_impl->batch) {
    fprintf(stderr, "[b2cuda] b2World::Step: this world belongs to a b2WorldBatch; call b2WorldBatch::Step\n");
    return;
  }
  m_locked = true;
  b2WorldImpl* I = m_impl;
  I->flush();
  if (I->arena == nullptr) {
    m_locked = false;
    return;
  }
  // new fixtures / moved bodies: refresh the pair list first (b2_world.cpp:1114-1122)
  if (m_newContacts) {
    int rc = b2g_find_new_contacts(I->arena);
    for (int attempt = 0; rc == B2G_ERR_CAPACITY && attempt < 8; ++attempt) {
      I->growContacts();  // keeps every contact, doubles max_contacts
      rc = b2g_find_new_contacts(I->arena);
    }
    b2gCheck(rc, "b2g_find_new_contacts");
    m_newContacts = false;
    I->contactsStale = true;
    I->applyContactFilter();
  }
  b2gStepParams P;
  P.dt = dt;
  P.velocity_iterations = velocityIterations;
  P.position_iterations = positionIterations;
  P.gravity_x = m_gravity.x;
  P.gravity_y = m_gravity.y;
  P.warm_starting = m_warmStarting ? 1 : 0;
  P.allow_sleep = m_allowSleep ? 1 : 0;
  P.clear_forces = m_clearForces ? 1 : 0;
  P.solver_mode = m_solverMode;
  P.record_events = m_contactListener ? 1 : 0;

  if (m_contactListener == nullptr) {
    int rc = b2g_step(I->arena, &P, &I->lastStats);
    I->bodiesStale = true;
    I->jointsStale = true;
    if (dt > 0.0f) I->lastInvDt = 1.0f / dt;
    if (rc == B2G_ERR_CAPACITY) {
      I->growContacts();
      rc = B2G_OK;
    }
    b2gCheck(rc, "b2g_step");
    I->applyContactFilter();  // the pairs this step's closing refresh inserted
  } else {
    // callbacks need the contact list between Collide and Solve (b2_contact.cpp:197-209)
    I->pullContacts();  // previous manifolds, for PreSolve's oldManifold
    std::unordered_map<b2Contact*, b2Manifold> oldManifolds;
    for (b2Contact* c : I->contacts) oldManifolds.emplace(c, c->m_manifold);
    b2gCheck(b2g_step_collide(I->arena, &P), "b2g_step_collide");
    b2gCheck(b2g_synchronize(I->arena), "b2g_synchronize");
    std::vector<uint8_t> wasTouching(I->contacts.size());
    for (size_t i = 0; i < I->contacts.size(); ++i) wasTouching[i] = I->contacts[i]->IsTouching();
    I->contactsStale = true;
    I->pullContacts();
    I->bodiesStale = true;
    for (size_t i = 0; i < I->contacts.size(); ++i) {
      b2Contact* c = I->contacts[i];
      bool touching = c->IsTouching();
      bool sensor = c->m_fixtureA->IsSensor() || c->m_fixtureB->IsSensor();
      if (!wasTouching[i] && touching) m_contactListener->BeginContact(c);
      if (wasTouching[i] && !touching) m_contactListener->EndContact(c);
      if (!sensor && touching) {
        b2Body* bA = c->m_fixtureA->m_body;
        b2Body* bB = c->m_fixtureB->m_body;
        bool active = (bA->IsAwake() && bA->m_type != b2_staticBody) || (bB->IsAwake() && bB->m_type != b2_staticBody);
        if (active) {
          auto it = oldManifolds.find(c);
          b2Manifold empty;
          memset(&empty, 0, sizeof(empty));
          m_contactListener->PreSolve(c, it != oldManifolds.end() ? &it->second : &empty);
        }
      }
    }
    I->pushContactOverrides();
    // impulses, forces and velocities the callbacks applied act in THIS step's solve, as in the reference
    // (the listener runs inside Collide, before Solve): upload what they touched.  The host copies are the
    // device's post-Collide state (every setter pulls before it edits), so the range upload is consistent.
    if (I->bodyDirtyHi > I->bodyDirtyLo) I->flush();
    int rc = b2g_step_solve(I->arena, &P, &I->lastStats);
    // from here on getters must see the solved state (PostSolve / EndContact callbacks below read it, and
    // whatever they edit is an edit of that state, uploaded by the next Step)
    I->bodiesStale = true;
    I->jointsStale = true;
    if (dt > 0.0f) I->lastInvDt = 1.0f / dt;
    if (rc == B2G_ERR_CAPACITY) {
      I->growContacts();
      rc = B2G_OK;
    }
    b2gCheck(rc, "b2g_step_solve");
    I->applyContactFilter();  // the pairs this step's closing refresh inserted
    // the broadphase at the end of the step rebuilt the list; handles of contacts that died are
    // parked in the graveyard (still valid) until the callbacks below have run
    I->contactsStale = true;
    I->pullContacts();
    // PostSolve for the contacts the solver processed (b2_island.cpp:621-647), device order
    for (b2Contact* c : I->contacts) {
      if (c->IsTouching() && c->IsEnabled() && !c->m_fixtureA->IsSensor() && !c->m_fixtureB->IsSensor()) {
        b2ContactImpulse imp;
        imp.count = c->m_manifold.pointCount;
        for (int32 k = 0; k < imp.count; ++k) {
          imp.normalImpulses[k] = c->m_manifold.points[k].normalImpulse;
          imp.tangentImpulses[k] = c->m_manifold.points[k].tangentImpulse;
        }
        m_contactListener->PostSolve(c, &imp);
      }
    }
    // EndContact for touching contacts destroyed by the broadphase (b2_contact_manager.cpp:48-51)
    for (b2Contact* c : I->graveyard)
      if (c->IsTouching()) m_contactListener->EndContact(c);
  }
  for (b2Contact* c : I->graveyard) delete c;
  I->graveyard.clear();
  I->contactsStale = true;
  I->releaseRetired();  // this step's pair refresh has retired the contacts of everything destroyed before it
  if (I->profiling) {
    m_profile.step = I->lastStats.ms_step;
    m_profile.collide = I->lastStats.ms_collide;
    m_profile.solve = I->lastStats.ms_solve + I->lastStats.ms_broadphase;  // the reference's solve includes it
    m_profile.broadphase = I->lastStats.ms_broadphase;
    m_profile.solveTOI = 0.0f;
  }
  m_locked = false;
}

void b2World::ClearForces() {
  m_impl->pullBodies();
  for (b2Body* b = m_bodyListHead; b; b = b->m_next) {
    if (b->m_force.x != 0.0f || b->m_force.y != 0.0f || b->m_torque != 0.0f) {
      b->m_force.SetZero();
      b->m_torque = 0.0f;
      m_impl->touchBody(b->m_index);
    }
  }
}

b2Contact* b2World::GetContactListStart() {
  if (m_newContacts && !m_locked) {
    // the reference creates contacts lazily at the next Step; the handles only exist after it
  }
  m_impl->pullContacts();
  return m_impl->sentinel.m_next;
}
b2Contact* b2World::GetContactListEnd() { return &m_impl->sentinel; }
int32 b2World::GetContactCount() const {
  if (m_impl->batch) {
    m_impl->pullContacts();
    return (int32)m_impl->contacts.size();
  }
  int32 n = 0;
  if (m_impl->arena) b2g_contact_count(m_impl->arena, &n);
  return n;
}
int32 b2World::GetBodyIndexCount() const { return (int32)m_impl->bodies.size(); }
int32 b2World::GetProxyCount() const {
  int32 n = 0;
  for (b2Fixture* f : m_impl->fixtures)
    if (f) ++n;
  return n;
}

// ================================================================================================
// b2WorldBatch (extension): many independent b2World objects stepped by one device pass (BASELINE config 4:
// thousands of small worlds).  Each member keeps its whole public API (bodies, fixtures, joints, getters,
// setters, queries); what a batch does not do is call contact listeners or user contact filters.
// ================================================================================================
void b2WorldBatchImpl::teardown() {
  // pull what only the device holds (body state, joint impulses, contacts with their manifolds and warm-start
  // impulses) before the arena goes
  for (b2WorldImpl* m : members) {
    if (!m) continue;
    m->bodiesStale = true;
    m->jointsStale = true;
    m->pullBodies();
    m->pullJoints();
  }
  int32 n = 0;
  b2gCheck(b2g_contact_count(arena, &n), "b2g_contact_count");
  saved.fa.assign(n, 0);
  saved.fb.assign(n, 0);
  saved.flags.assign(n, 0);
  saved.man.assign((size_t)n * 16, 0.0f);
  saved.mat.assign((size_t)n * 4, 0.0f);
  if (n > 0) {
    b2gContactArrays a;
    memset(&a, 0, sizeof(a));
    a.fixture_a = saved.fa.data(); a.fixture_b = saved.fb.data(); a.flags = saved.flags.data();
    a.manifold = saved.man.data(); a.material = saved.mat.data();
    b2gCheck(b2g_download_contacts(arena, 0, n, &a), "b2g_download_contacts");
  }
  savedSlotFixtures = slotFixtures;
  b2gCheck(b2g_arena_destroy(arena), "b2g_arena_destroy");
  arena = nullptr;
  for (b2WorldImpl* m : members)
    if (m) m->arena = nullptr;
  contactsStale = true;
}

void b2WorldBatchImpl::growContacts() {
  teardown();
  capContacts *= 2;
  fprintf(stderr, "[b2cuda] batch contact capacity grown to %d\n", capContacts);
  flush();
}

void b2WorldBatchImpl::ensureArena() {
  int32 needB = 1, needF = 1, needQ = 1, needJ = 0;
  for (b2WorldImpl* m : members) {
    if (!m) continue;
    needB = std::max(needB, (int32)m->bodies.size());
    needF = std::max(needF, (int32)m->fixtures.size());
    needQ = std::max(needQ, (int32)(m->shapePool.size() / 4));
    needJ += (int32)m->joints.size();
  }
  const int32 nw = (int32)members.size();
  if (arena && needB <= slotBodies && needF <= slotFixtures && needQ <= slotQuads && needJ <= capJoints && nw <= worldsOnDevice)
    return;
  if (arena) teardown();
  // slots: equal for every world, with headroom (worlds that spawn bodies while they run, like the tumbler)
  auto grow = [](int32 slot, int32 need) {
    int32 c = std::max(slot, 16);
    while (c < need) c *= 2;
    return c;
  };
  slotBodies = grow(slotBodies, needB);
  slotFixtures = grow(slotFixtures, needF);
  slotQuads = grow(slotQuads, needQ);
  capJoints = std::max(grow(capJoints, needJ), 1);
  worldsOnDevice = nw;
  const int64_t contactsWanted = std::max<int64_t>((int64_t)saved.fa.size() * 2, (int64_t)nw * slotBodies * 4);
  capContacts = std::max(capContacts, 1024);
  while (capContacts < contactsWanted) capContacts *= 2;
  b2gArenaDef def;
  memset(&def, 0, sizeof(def));
  int dev = g_defaultDevice;
  if (dev < 0) {
    const char* e = getenv("B2G_DEVICE");
    dev = e ? atoi(e) : 0;
  }
  def.device = dev;
  def.num_worlds = nw;
  def.max_bodies = nw * slotBodies;
  def.max_fixtures = nw * slotFixtures;
  def.max_shape_quads = nw * slotQuads;
  def.max_contacts = capContacts;
  def.max_joints = capJoints;
  b2gCheck(b2g_arena_create(&def, &arena), "b2g_arena_create");
  b2g_set_profiling(arena, profiling ? 1 : 0);
  if (getenv("B2G_KERNEL_TIMING")) b2g_set_kernel_timing(arena, 1);
  {
    // every slot starts as placeholders of its world: disabled static bodies, dead fixtures
    const int32 nb = nw * slotBodies, nf = nw * slotFixtures;
    std::vector<float> zero4((size_t)std::max(nb, nf) * 4, 0.0f);
    std::vector<uint32_t> flags(nb, 0u), tf(nf, B2G_FIX_DEAD), filter((size_t)nf * 2, 0u);
    std::vector<int32_t> wid(nb), body(nf), off(nf);
    for (int32 i = 0; i < nb; ++i) wid[i] = i / slotBodies;
    for (int32 i = 0; i < nf; ++i) {
      body[i] = (i / slotFixtures) * slotBodies;
      off[i] = (i / slotFixtures) * slotQuads;
    }
    b2gBodyArrays a;
    a.pos = zero4.data(); a.vel = zero4.data(); a.xf = zero4.data(); a.mass = zero4.data(); a.center = zero4.data();
    a.force = zero4.data(); a.flags = flags.data(); a.world = wid.data();
    b2gCheck(b2g_upload_bodies(arena, 0, nb, &a), "b2g_upload_bodies");
    b2gFixtureArrays f;
    f.body = body.data(); f.shape_off = off.data(); f.type_flags = tf.data(); f.filter = filter.data();
    f.material = zero4.data();
    b2gCheck(b2g_upload_fixtures(arena, 0, nf, &f), "b2g_upload_fixtures");
  }
  jointsOnDevice = 0;
  for (int32 w = 0; w < nw; ++w) {
    b2WorldImpl* m = members[w];
    if (!m) continue;
    m->arena = arena;
    m->worldId = w;
    m->bodyBase = w * slotBodies;
    m->fixtureBase = w * slotFixtures;
    m->quadBase = w * slotQuads;
    m->shapesUploaded = 0;
    m->bodiesOnDevice = 0;
    m->bodyDirtyLo = 0;
    m->bodyDirtyHi = (int32)m->bodies.size();
    m->fixtureDirtyLo = 0;
    m->fixtureDirtyHi = (int32)m->fixtures.size();
    m->jointsDirty = !m->joints.empty();
    m->jointsOnDevice = 0;
    m->contactsStale = true;
    m->world->m_newContacts = true;
  }
}

// upload everything any member changed since the last step
void b2WorldBatchImpl::flush() {
  ensureArena();
  bool jointsDirty = false, fresh = false;
  {
    // every member's new shapes and edited bodies / fixtures in three scatter uploads, however many worlds
    // contributed a row (the tumbler benchmark spawns one box per world per step)
    b2WorldImpl::BodyRows B;
    b2WorldImpl::FixtureRows F;
    std::vector<float> quads;
    std::vector<int32_t> quadIndex;
    for (b2WorldImpl* m : members) {
      if (!m) continue;
      fresh = fresh || m->bodiesOnDevice == 0;
      jointsDirty = jointsDirty || m->jointsDirty;
      const int32 nq = (int32)(m->shapePool.size() / 4);
      for (int32 q = m->shapesUploaded; q < nq; ++q) {
        quadIndex.push_back(m->quadBase + q);
        quads.insert(quads.end(), m->shapePool.begin() + (size_t)q * 4, m->shapePool.begin() + (size_t)q * 4 + 4);
      }
      m->shapesUploaded = nq;
      for (const b2WorldImpl::ShapeSlot& d : m->dirtyShapes)
        for (int32 q = d.off; q < d.off + d.quads; ++q) {
          quadIndex.push_back(m->quadBase + q);
          quads.insert(quads.end(), m->shapePool.begin() + (size_t)q * 4, m->shapePool.begin() + (size_t)q * 4 + 4);
        }
      m->dirtyShapes.clear();
      if (m->bodyDirtyHi > m->bodyDirtyLo) {
        m->appendBodyRows(m->bodyDirtyLo, m->bodyDirtyHi - m->bodyDirtyLo, B);
        m->bodiesOnDevice = std::max(m->bodiesOnDevice, m->bodyDirtyHi);
        m->bodyDirtyLo = INT32_MAX;
        m->bodyDirtyHi = 0;
      }
      if (m->fixtureDirtyHi > m->fixtureDirtyLo) {
        m->appendFixtureRows(m->fixtureDirtyLo, m->fixtureDirtyHi - m->fixtureDirtyLo, F);
        m->fixtureDirtyLo = INT32_MAX;
        m->fixtureDirtyHi = 0;
      }
    }
    if (!quadIndex.empty())
      b2gCheck(b2g_upload_shapes_indexed(arena, (int32_t)quadIndex.size(), quadIndex.data(), quads.data()), "b2g_upload_shapes_indexed");
    if (!B.index.empty()) {
      b2gBodyArrays a;
      a.pos = B.pos.data(); a.vel = B.vel.data(); a.xf = B.xf.data(); a.mass = B.mass.data(); a.center = B.center.data();
      a.force = B.force.data(); a.flags = B.flags.data(); a.world = B.wid.data();
      b2gCheck(b2g_upload_bodies_indexed(arena, (int32_t)B.index.size(), B.index.data(), &a), "b2g_upload_bodies_indexed");
    }
    if (!F.index.empty()) {
      b2gFixtureArrays a;
      a.body = F.body.data(); a.shape_off = F.off.data(); a.type_flags = F.tf.data(); a.filter = F.filter.data();
      a.material = F.mat.data();
      b2gCheck(b2g_upload_fixtures_indexed(arena, (int32_t)F.index.size(), F.index.data(), &a), "b2g_upload_fixtures_indexed");
    }
  }
  if (jointsDirty) {
    // the joint table is packed (world after world): one member's change moves the others' rows.  Every site
    // that edits a member's joint list has pulled that member's impulses first (pullJoints); pull the rest.
    int32 total = 0;
    for (b2WorldImpl* m : members) {
      if (!m) continue;
      m->pullJoints();
      total += (int32)m->joints.size();
    }
    std::vector<int32_t> jb((size_t)total * 2);
    std::vector<float> anchors((size_t)total * 4), params((size_t)total * 12, 0.0f), state((size_t)total * 5);
    int32 k = 0;
    for (b2WorldImpl* m : members) {
      if (!m) continue;
      m->jointBase = k;
      for (b2Joint* j : m->joints) {
        jb[(size_t)k * 2] = m->bodyBase + j->m_bodyA->m_index;
        jb[(size_t)k * 2 + 1] = m->bodyBase + j->m_bodyB->m_index;
        j->WriteDevice(&anchors[(size_t)k * 4], &params[(size_t)k * 12], &state[(size_t)k * 5]);
        ++k;
      }
      m->jointsOnDevice = (int32)m->joints.size();
      m->jointsDirty = false;
      m->jointsStale = false;
    }
    b2gJointArrays a;
    a.bodies = jb.data(); a.anchors = anchors.data(); a.params = params.data(); a.state = state.data();
    if (total > 0) b2gCheck(b2g_upload_joints(arena, 0, total, &a), "b2g_upload_joints");
    jointsOnDevice = total;
    b2gCheck(b2g_set_counts(arena, worldsOnDevice * slotBodies, worldsOnDevice * slotFixtures, total), "b2g_set_counts");
  }
  if (!saved.fa.empty()) {
    // contacts of the arena that was torn down, moved to the new slots
    for (size_t i = 0; i < saved.fa.size(); ++i) {
      saved.fa[i] = (saved.fa[i] / savedSlotFixtures) * slotFixtures + saved.fa[i] % savedSlotFixtures;
      saved.fb[i] = (saved.fb[i] / savedSlotFixtures) * slotFixtures + saved.fb[i] % savedSlotFixtures;
    }
    b2gContactArrays a;
    memset(&a, 0, sizeof(a));
    a.fixture_a = saved.fa.data(); a.fixture_b = saved.fb.data(); a.flags = saved.flags.data();
    a.manifold = saved.man.data(); a.material = saved.mat.data();
    b2gCheck(b2g_upload_contacts(arena, (int32_t)saved.fa.size(), &a), "b2g_upload_contacts");
    saved = b2WorldImpl::SavedContacts();
    for (b2WorldImpl* m : members)
      if (m) m->world->m_newContacts = true;
  }
  if (fresh) b2gCheck(b2g_set_inv_dt0(arena, lastInvDt), "b2g_set_inv_dt0");
}

void b2WorldBatchImpl::pullAllBodies() {
  if (!arena) return;
  const int32 n = worldsOnDevice * slotBodies;
  std::vector<float> pos((size_t)n * 4), vel((size_t)n * 4), xf((size_t)n * 4), force((size_t)n * 4);
  std::vector<uint32_t> flags(n);
  b2gBodyArrays a;
  memset(&a, 0, sizeof(a));
  a.pos = pos.data(); a.vel = vel.data(); a.xf = xf.data(); a.force = force.data(); a.flags = flags.data();
  b2gCheck(b2g_download_bodies(arena, 0, n, &a), "b2g_download_bodies");
  for (b2WorldImpl* m : members) {
    if (!m || !m->bodiesStale || m->arena != arena) continue;
    m->bodiesStale = false;
    const int32 k = std::min((int32)m->bodies.size(), m->bodiesOnDevice);
    const size_t o = (size_t)m->bodyBase;
    m->applyBodyRows(k, &pos[o * 4], &vel[o * 4], &xf[o * 4], &force[o * 4], &flags[o]);
  }
}

void b2WorldBatchImpl::pullContacts() {
  if (!contactsStale) return;
  contactsStale = false;
  int32 n = 0;
  if (arena) b2gCheck(b2g_contact_count(arena, &n), "b2g_contact_count");
  std::vector<int32_t> fa(n), fb(n);
  std::vector<uint32_t> flags(n);
  std::vector<float> man((size_t)n * 16), mat((size_t)n * 4);
  if (n > 0) {
    b2gContactArrays a;
    memset(&a, 0, sizeof(a));
    a.fixture_a = fa.data(); a.fixture_b = fb.data(); a.flags = flags.data(); a.manifold = man.data();
    a.material = mat.data();
    b2gCheck(b2g_download_contacts(arena, 0, n, &a), "b2g_download_contacts");
  }
  // split by world (a contact's fixtures are in the same slot), keeping device order inside a world
  const int32 nw = (int32)members.size();
  std::vector<std::vector<int32_t>> of(nw);
  for (int32 i = 0; i < n; ++i) of[fa[i] / slotFixtures].push_back(i);
  for (int32 w = 0; w < nw; ++w) {
    b2WorldImpl* m = members[w];
    if (!m) continue;
    const std::vector<int32_t>& idx = of[w];
    const int32 k = (int32)idx.size();
    std::vector<int32_t> la(k), lb(k);
    std::vector<uint32_t> lf(k);
    std::vector<float> lman((size_t)k * 16), lmat((size_t)k * 4);
    int32 kept = 0;
    for (int32 t = 0; t < k; ++t) {
      const int32 i = idx[t];
      const int32 a = fa[i] - m->fixtureBase, b = fb[i] - m->fixtureBase;
      // a fixture destroyed since the last step still has its contacts on the device until the next pair refresh
      if (a < 0 || b < 0 || a >= (int32)m->fixtures.size() || b >= (int32)m->fixtures.size() || !m->fixtures[a] || !m->fixtures[b])
        continue;
      la[kept] = a;
      lb[kept] = b;
      lf[kept] = flags[i];
      memcpy(&lman[(size_t)kept * 16], &man[(size_t)i * 16], 64);
      memcpy(&lmat[(size_t)kept * 4], &mat[(size_t)i * 4], 16);
      const_cast<std::vector<int32_t>&>(idx)[kept] = i;
      ++kept;
    }
    m->rebuildContacts(kept, la.data(), lb.data(), lf.data(), lman.data(), lmat.data(), idx.data());
    for (b2Contact* c : m->graveyard) delete c;  // no listener runs in a batch: dead handles are not kept
    m->graveyard.clear();
    m->contactsStale = false;
  }
}

b2WorldBatch::b2WorldBatch() {
  m_impl = new b2WorldBatchImpl();
  memset(&m_impl->lastStats, 0, sizeof(b2gStepStats));
}

b2WorldBatch::~b2WorldBatch() {
  // members that outlive the batch become worlds without a device state again (they would re-upload at their
  // next Step); normally the batch is destroyed after its worlds
  if (m_impl->arena) {
    for (b2WorldImpl* m : m_impl->members) {
      if (!m) continue;
      m->bodiesStale = true;
      m->jointsStale = true;
      m->pullBodies();
      m->pullJoints();
    }
    b2g_arena_destroy(m_impl->arena);
  }
  for (b2WorldImpl* m : m_impl->members) {
    if (!m) continue;
    m->batch = nullptr;
    m->arena = nullptr;
    m->worldId = m->bodyBase = m->fixtureBase = m->quadBase = m->jointBase = 0;
    m->shapesUploaded = 0;
    m->bodiesOnDevice = 0;
    m->jointsOnDevice = 0;
    m->contactsStale = true;
  }
  delete m_impl;
}

bool b2WorldBatch::Add(b2World* world) {
  b2WorldImpl* m = world ? world->GetImpl() : nullptr;
  if (!m || m->batch || m->arena || world->IsLocked()) return false;  // only worlds that have not been stepped yet
  m->batch = m_impl;
  m->worldId = (int32)m_impl->members.size();
  m_impl->members.push_back(m);
  return true;
}

int32 b2WorldBatch::GetWorldCount() const { return (int32)m_impl->members.size(); }

b2World* b2WorldBatch::GetWorld(int32 index) const {
  if (index < 0 || index >= (int32)m_impl->members.size() || !m_impl->members[index]) return nullptr;
  return m_impl->members[index]->world;
}

void b2WorldBatch::SetProfiling(bool flag) {
  m_impl->profiling = flag;
  if (m_impl->arena) b2g_set_profiling(m_impl->arena, flag ? 1 : 0);
}

float b2WorldBatch::GetLastStepMilliseconds() const { return m_impl->lastStats.ms_step; }

// One b2World::Step (b2_world.cpp:1108-1171) of every member.  Gravity, the world flags and the solver mode are
// the first live member's: a batch is for many instances of one kind of world.
void b2WorldBatch::Step(float dt, int32 velocityIterations, int32 positionIterations) {
  b2WorldBatchImpl* B = m_impl;
  b2World* first = nullptr;
  for (b2WorldImpl* m : B->members)
    if (m) {
      if (m->world->m_locked) return;
      if (!first) first = m->world;
    }
  if (!first) return;
  for (b2WorldImpl* m : B->members)
    if (m) m->world->m_locked = true;
  {
    static bool warned = false;
    if (!warned)
      for (b2WorldImpl* m : B->members)
        if (m && (m->world->m_contactListener != nullptr || m->customFilter())) {
          fprintf(stderr, "[b2cuda] b2WorldBatch::Step: contact listeners and user contact filters are not called for "
                          "batched worlds (step such a world on its own)\n");
          warned = true;
          break;
        }
  }
  B->flush();
  bool newContacts = false;
  for (b2WorldImpl* m : B->members)
    if (m) newContacts = newContacts || m->world->m_newContacts;
  if (newContacts) {
    int rc = b2g_find_new_contacts(B->arena);
    for (int attempt = 0; rc == B2G_ERR_CAPACITY && attempt < 8; ++attempt) {
      B->growContacts();
      rc = b2g_find_new_contacts(B->arena);
    }
    b2gCheck(rc, "b2g_find_new_contacts");
    for (b2WorldImpl* m : B->members)
      if (m) m->world->m_newContacts = false;
  }
  b2gStepParams P;
  P.dt = dt;
  P.velocity_iterations = velocityIterations;
  P.position_iterations = positionIterations;
  P.gravity_x = first->m_gravity.x;
  P.gravity_y = first->m_gravity.y;
  P.warm_starting = first->m_warmStarting ? 1 : 0;
  P.allow_sleep = first->m_allowSleep ? 1 : 0;
  P.clear_forces = first->m_clearForces ? 1 : 0;
  P.solver_mode = first->m_solverMode;
  P.record_events = 0;
  int rc = b2g_step(B->arena, &P, &B->lastStats);
  if (dt > 0.0f) B->lastInvDt = 1.0f / dt;
  for (b2WorldImpl* m : B->members) {
    if (!m) continue;
    m->bodiesStale = true;
    m->jointsStale = true;
    m->contactsStale = true;
    if (dt > 0.0f) m->lastInvDt = 1.0f / dt;
  }
  B->contactsStale = true;
  for (b2WorldImpl* m : B->members)
    if (m) m->releaseRetired();
  B->bulkPull = B->memberPulls >= std::max(4, B->live() / 8);
  B->memberPulls = 0;
  if (rc == B2G_ERR_CAPACITY) {
    B->growContacts();
    rc = B2G_OK;
  }
  b2gCheck(rc, "b2g_step");
  for (b2WorldImpl* m : B->members)
    if (m) m->world->m_locked = false;
}

// ================================================================================================
// b2Body
// ================================================================================================
b2Body::b2Body(const b2BodyDef* bd, b2World* world) {
  m_flags = 0;
  if (bd->bullet) m_flags |= e_bulletFlag;
  if (bd->fixedRotation) m_flags |= e_fixedRotationFlag;
  if (bd->allowSleep) m_flags |= e_autoSleepFlag;
  if (bd->awake && bd->type != b2_staticBody) m_flags |= e_awakeFlag;
  if (bd->enabled) m_flags |= e_enabledFlag;
  m_world = world;
  m_xf.p = bd->position;
  m_xf.q.Set(bd->angle);
  m_sweep.localCenter.SetZero();
  m_sweep.c0 = m_xf.p;
  m_sweep.c = m_xf.p;
  m_sweep.a0 = bd->angle;
  m_sweep.a = bd->angle;
  m_sweep.alpha0 = 0.0f;
  m_jointList = nullptr;
  m_prev = nullptr;
  m_next = nullptr;
  m_linearVelocity = bd->linearVelocity;
  m_angularVelocity = bd->angularVelocity;
  m_linearDamping = bd->linearDamping;
  m_angularDamping = bd->angularDamping;
  m_gravityScale = bd->gravityScale;
  m_force.SetZero();
  m_torque = 0.0f;
  m_sleepTime = 0.0f;
  m_type = bd->type;
  m_mass = 0.0f;
  m_invMass = 0.0f;
  m_I = 0.0f;
  m_invI = 0.0f;
  m_userData = bd->userData;
  m_fixtureList = nullptr;
  m_fixtureCount = 0;
  m_index = -1;
}

void b2Body::SyncIn() const { m_world->m_impl->pullBodies(); }
void b2Body::Touch() {
  // a body the device has not seen yet (created since the last step) has nothing to pull: its host copy is the
  // only one.  (bodiesStale stays set, so the first edit of an older body still pulls before it marks its range.)
  if (m_onDevice) m_world->m_impl->pullBodies();
  m_world->m_impl->touchBody(m_index);
}

b2Fixture* b2Body::CreateFixture(const b2FixtureDef* def) {
  if (m_world->IsLocked()) return nullptr;
  Touch();
  b2WorldImpl* I = m_world->m_impl;
  b2Fixture* f = new b2Fixture();
  f->m_userData = def->userData;
  f->m_friction = def->friction;
  f->m_restitution = def->restitution;
  f->m_restitutionThreshold = def->restitutionThreshold;
  f->m_body = this;
  f->m_filter = def->filter;
  f->m_isSensor = def->isSensor;
  f->m_shape = def->shape->Clone();
  f->m_density = def->density;
  f->m_id = g_fixtureIdCounter++;
  if (!I->freeFixtures.empty()) {
    f->m_index = I->freeFixtures.back();
    I->freeFixtures.pop_back();
    I->fixtures[f->m_index] = f;
  } else {
    f->m_index = (int32)I->fixtures.size();
    I->fixtures.push_back(f);
  }
  I->touchFixture(f->m_index);
  int32 nq = f->m_shape->DeviceQuadCount();
  f->m_shapeOff = -1;
  for (size_t k = I->freeShapes.size(); k-- > 0;) {  // a free shape record of the same length, else a new one
    if (I->freeShapes[k].quads == nq) {
      f->m_shapeOff = I->freeShapes[k].off;
      I->freeShapes.erase(I->freeShapes.begin() + (ptrdiff_t)k);
      break;
    }
  }
  if (f->m_shapeOff < 0) {
    f->m_shapeOff = (int32)(I->shapePool.size() / 4);
    I->shapePool.resize(I->shapePool.size() + (size_t)nq * 4);
  } else {
    I->dirtyShapes.push_back({f->m_shapeOff, nq});  // a reused record below the pool's uploaded mark
  }
  f->m_shape->WriteDeviceQuads(&I->shapePool[(size_t)f->m_shapeOff * 4]);
  f->m_next = m_fixtureList;
  m_fixtureList = f;
  ++m_fixtureCount;
  if (f->m_density > 0.0f) ResetMassData();
  f->m_shape->ComputeAABB(&f->m_aabb, m_xf);
  m_world->m_newContacts = true;
  return f;
}

b2Fixture* b2Body::CreateFixture(const b2Shape* shape, float density) {
  b2FixtureDef def;
  def.shape = shape;
  def.density = density;
  return CreateFixture(&def);
}

void b2Body::DestroyFixture(b2Fixture* fixture) {
  if (fixture == nullptr || m_world->IsLocked()) return;
  Touch();
  b2WorldImpl* I = m_world->m_impl;
  b2Fixture** node = &m_fixtureList;
  while (*node != nullptr) {
    if (*node == fixture) {
      *node = fixture->m_next;
      break;
    }
    node = &(*node)->m_next;
  }
  I->fixtures[fixture->m_index] = nullptr;
  I->deadFixtures = true;
  I->touchFixture(fixture->m_index);
  I->retiredFixtures.push_back(fixture->m_index);
  I->retiredShapes.push_back({fixture->m_shapeOff, fixture->m_shape->DeviceQuadCount()});
  for (auto& kv : I->contactPool) delete kv.second;
  I->contactPool.clear();
  I->contacts.clear();
  I->contactsStale = true;
  m_world->m_newContacts = true;
  delete fixture->m_shape;
  delete fixture;
  --m_fixtureCount;
  ResetMassData();
}

void b2Body::ResetMassData() {
  Touch();
  m_mass = 0.0f;
  m_invMass = 0.0f;
  m_I = 0.0f;
  m_invI = 0.0f;
  m_sweep.localCenter.SetZero();
  if (m_type == b2_staticBody || m_type == b2_kinematicBody) {
    m_sweep.c0 = m_xf.p;
    m_sweep.c = m_xf.p;
    m_sweep.a0 = m_sweep.a;
    return;
  }
  b2Vec2 localCenter = b2Vec2_zero;
  for (b2Fixture* f = m_fixtureList; f; f = f->m_next) {
    if (f->m_density == 0.0f) continue;
    b2MassData massData;
    f->GetMassData(&massData);
    m_mass += massData.mass;
    localCenter += massData.mass * massData.center;
    m_I += massData.I;
  }
  if (m_mass > 0.0f) {
    m_invMass = 1.0f / m_mass;
    localCenter *= m_invMass;
  }
  if (m_I > 0.0f && (m_flags & e_fixedRotationFlag) == 0) {
    m_I -= m_mass * b2Dot(localCenter, localCenter);
    m_invI = 1.0f / m_I;
  } else {
    m_I = 0.0f;
    m_invI = 0.0f;
  }
  b2Vec2 oldCenter = m_sweep.c;
  m_sweep.localCenter = localCenter;
  m_sweep.c0 = m_sweep.c = b2Mul(m_xf, m_sweep.localCenter);
  m_linearVelocity += b2Cross(m_angularVelocity, m_sweep.c - oldCenter);
}

void b2Body::SetMassData(const b2MassData* massData) {
  if (m_world->IsLocked() || m_type != b2_dynamicBody) return;
  Touch();
  m_invMass = 0.0f;
  m_I = 0.0f;
  m_invI = 0.0f;
  m_mass = massData->mass;
  if (m_mass <= 0.0f) m_mass = 1.0f;
  m_invMass = 1.0f / m_mass;
  if (massData->I > 0.0f && (m_flags & e_fixedRotationFlag) == 0) {
    m_I = massData->I - m_mass * b2Dot(massData->center, massData->center);
    m_invI = 1.0f / m_I;
  }
  b2Vec2 oldCenter = m_sweep.c;
  m_sweep.localCenter = massData->center;
  m_sweep.c0 = m_sweep.c = b2Mul(m_xf, m_sweep.localCenter);
  m_linearVelocity += b2Cross(m_angularVelocity, m_sweep.c - oldCenter);
}

void b2Body::GetMassData(b2MassData* data) const {
  data->mass = m_mass;
  data->I = m_I + m_mass * b2Dot(m_sweep.localCenter, m_sweep.localCenter);
  data->center = m_sweep.localCenter;
}
float b2Body::GetInertia() const { return m_I + m_mass * b2Dot(m_sweep.localCenter, m_sweep.localCenter); }

void b2Body::SetTransform(const b2Vec2& position, float angle) {
  if (m_world->IsLocked()) return;
  Touch();
  m_xf.q.Set(angle);
  m_xf.p = position;
  m_sweep.c = b2Mul(m_xf, m_sweep.localCenter);
  m_sweep.a = angle;
  m_sweep.c0 = m_sweep.c;
  m_sweep.a0 = angle;
  UpdateAABBs();
  m_world->m_newContacts = true;
}

void b2Body::UpdateAABBs() {
  for (b2Fixture* f = m_fixtureList; f; f = f->m_next) f->m_shape->ComputeAABB(&f->m_aabb, m_xf);
}

const b2Transform& b2Body::GetTransform() const { SyncIn(); return m_xf; }
const b2Vec2& b2Body::GetPosition() const { SyncIn(); return m_xf.p; }
float b2Body::GetAngle() const { SyncIn(); return m_sweep.a; }
const b2Vec2& b2Body::GetWorldCenter() const { SyncIn(); return m_sweep.c; }
const b2Vec2& b2Body::GetLocalCenter() const { return m_sweep.localCenter; }
const b2Vec2& b2Body::GetLinearVelocity() const { SyncIn(); return m_linearVelocity; }
float b2Body::GetAngularVelocity() const { SyncIn(); return m_angularVelocity; }
b2Vec2 b2Body::GetLinearVelocityFromWorldPoint(const b2Vec2& worldPoint) const {
  SyncIn();
  return m_linearVelocity + b2Cross(m_angularVelocity, worldPoint - m_sweep.c);
}

void b2Body::SetLinearVelocity(const b2Vec2& v) {
  if (m_type == b2_staticBody) return;
  Touch();
  if (b2Dot(v, v) > 0.0f) SetAwake(true);
  m_linearVelocity = v;
}
void b2Body::SetAngularVelocity(float w) {
  if (m_type == b2_staticBody) return;
  Touch();
  if (w * w > 0.0f) SetAwake(true);
  m_angularVelocity = w;
}
void b2Body::ApplyForce(const b2Vec2& force, const b2Vec2& point, bool wake) {
  if (m_type != b2_dynamicBody) return;
  Touch();
  if (wake && (m_flags & e_awakeFlag) == 0) SetAwake(true);
  if (m_flags & e_awakeFlag) {
    m_force += force;
    m_torque += b2Cross(point - m_sweep.c, force);
  }
}
void b2Body::ApplyForceToCenter(const b2Vec2& force, bool wake) {
  if (m_type != b2_dynamicBody) return;
  Touch();
  if (wake && (m_flags & e_awakeFlag) == 0) SetAwake(true);
  if (m_flags & e_awakeFlag) m_force += force;
}
void b2Body::ApplyTorque(float torque, bool wake) {
  if (m_type != b2_dynamicBody) return;
  Touch();
  if (wake && (m_flags & e_awakeFlag) == 0) SetAwake(true);
  if (m_flags & e_awakeFlag) m_torque += torque;
}
void b2Body::ApplyLinearImpulse(const b2Vec2& impulse, const b2Vec2& point, bool wake) {
  if (m_type != b2_dynamicBody) return;
  Touch();
  if (wake && (m_flags & e_awakeFlag) == 0) SetAwake(true);
  if (m_flags & e_awakeFlag) {
    m_linearVelocity += m_invMass * impulse;
    m_angularVelocity += m_invI * b2Cross(point - m_sweep.c, impulse);
  }
}
void b2Body::ApplyLinearImpulseToCenter(const b2Vec2& impulse, bool wake) {
  if (m_type != b2_dynamicBody) return;
  Touch();
  if (wake && (m_flags & e_awakeFlag) == 0) SetAwake(true);
  if (m_flags & e_awakeFlag) m_linearVelocity += m_invMass * impulse;
}
void b2Body::ApplyAngularImpulse(float impulse, bool wake) {
  if (m_type != b2_dynamicBody) return;
  Touch();
  if (wake && (m_flags & e_awakeFlag) == 0) SetAwake(true);
  if (m_flags & e_awakeFlag) m_angularVelocity += m_invI * impulse;
}
void b2Body::SetLinearDamping(float v) { Touch(); m_linearDamping = v; }
void b2Body::SetAngularDamping(float v) { Touch(); m_angularDamping = v; }
void b2Body::SetGravityScale(float v) { Touch(); m_gravityScale = v; }

void b2Body::SetType(b2BodyType type) {
  if (m_world->IsLocked() || m_type == type) return;
  Touch();
  b2World* w = m_world;
  // keep static bodies at the tail of the list (b2_body.cpp:140-188)
  if (m_prev) m_prev->m_next = m_next;
  if (m_next) m_next->m_prev = m_prev;
  if (this == w->m_bodyListHead) w->m_bodyListHead = m_next;
  if (this == w->m_bodyListTail) w->m_bodyListTail = m_prev;
  m_prev = m_next = nullptr;
  if (w->m_bodyListHead == nullptr) {
    w->m_bodyListHead = w->m_bodyListTail = this;
  } else if (type == b2_staticBody) {
    m_prev = w->m_bodyListTail;
    w->m_bodyListTail->m_next = this;
    w->m_bodyListTail = this;
  } else {
    m_next = w->m_bodyListHead;
    w->m_bodyListHead->m_prev = this;
    w->m_bodyListHead = this;
  }
  m_type = type;
  ResetMassData();
  if (m_type == b2_staticBody) {
    m_linearVelocity.SetZero();
    m_angularVelocity = 0.0f;
    m_sweep.a0 = m_sweep.a;
    m_sweep.c0 = m_sweep.c;
    m_flags &= ~e_awakeFlag;
    UpdateAABBs();
  }
  SetAwake(true);
  m_force.SetZero();
  m_torque = 0.0f;
  w->m_newContacts = true;
}

void b2Body::SetBullet(bool flag) { if (flag) m_flags |= e_bulletFlag; else m_flags &= ~e_bulletFlag; }
bool b2Body::IsBullet() const { return (m_flags & e_bulletFlag) == e_bulletFlag; }
void b2Body::SetSleepingAllowed(bool flag) {
  Touch();
  if (flag) {
    m_flags |= e_autoSleepFlag;
  } else {
    m_flags &= ~e_autoSleepFlag;
    SetAwake(true);
  }
}
bool b2Body::IsSleepingAllowed() const { return (m_flags & e_autoSleepFlag) == e_autoSleepFlag; }
void b2Body::SetAwake(bool flag) {
  if (m_type == b2_staticBody) return;
  Touch();
  if (flag) {
    m_flags |= e_awakeFlag;
    m_sleepTime = 0.0f;
  } else {
    m_flags &= ~e_awakeFlag;
    m_sleepTime = 0.0f;
    m_linearVelocity.SetZero();
    m_angularVelocity = 0.0f;
    m_force.SetZero();
    m_torque = 0.0f;
  }
}
bool b2Body::IsAwake() const { SyncIn(); return (m_flags & e_awakeFlag) == e_awakeFlag; }
void b2Body::SetEnabled(bool flag) {
  if (flag == IsEnabled()) return;
  Touch();
  if (flag) m_flags |= e_enabledFlag; else m_flags &= ~e_enabledFlag;
  m_world->m_newContacts = true;
}
bool b2Body::IsEnabled() const { return (m_flags & e_enabledFlag) == e_enabledFlag; }
void b2Body::SetFixedRotation(bool flag) {
  bool status = (m_flags & e_fixedRotationFlag) == e_fixedRotationFlag;
  if (status == flag) return;
  Touch();
  if (flag) m_flags |= e_fixedRotationFlag; else m_flags &= ~e_fixedRotationFlag;
  m_angularVelocity = 0.0f;
  ResetMassData();
}
bool b2Body::IsFixedRotation() const { return (m_flags & e_fixedRotationFlag) == e_fixedRotationFlag; }
int32 b2Body::GetContactCount() {
  m_world->m_impl->pullContacts();
  return (int32)m_contacts.size();
}
b2Contact* b2Body::GetContact(int32 idx) {
  m_world->m_impl->pullContacts();
  return m_contacts[idx];
}
bool b2Body::ShouldCollide(const b2Body* other) const {
  if (m_type != b2_dynamicBody && other->m_type != b2_dynamicBody) return false;
  for (b2JointEdge* jn = m_jointList; jn; jn = jn->next)
    if (jn->other == other && jn->joint->GetCollideConnected() == false) return false;
  return true;
}

// ================================================================================================
// b2Fixture / b2Contact / joints / filter
// ================================================================================================
void b2Fixture::SetSensor(bool sensor) {
  if (sensor == m_isSensor) return;
  m_body->Touch();
  m_isSensor = sensor;
  m_body->m_world->m_impl->touchFixture(m_index);
}
void b2Fixture::SetFilterData(const b2Filter& filter) {
  m_filter = filter;
  Refilter();
}
void b2Contact::FlagForFiltering() {
  b2WorldImpl* I = m_fixtureA->m_body->m_world->m_impl;
  I->refilter.push_back(m_fixtureA->m_index);
  m_fixtureA->m_body->m_world->m_newContacts = true;
}
void b2Fixture::Refilter() {
  // the device re-evaluates the filter for every pair on every broadphase pass, which is the
  // effect of the reference's e_filterFlag protocol (b2_fixture.cpp:100-136)
  m_body->m_world->m_impl->touchFixture(m_index);
  m_body->m_world->m_impl->refilter.push_back(m_index);
  m_body->m_world->m_newContacts = true;
}
bool b2Fixture::TestPoint(const b2Vec2& p) const { return m_shape->TestPoint(m_body->GetTransform(), p); }
bool b2Fixture::RayCast(b2RayCastOutput* output, const b2RayCastInput& input) const {
  return m_shape->RayCast(output, input, m_body->GetTransform());
}
void b2Fixture::SetFriction(float v) { m_friction = v; m_body->m_world->m_impl->touchFixture(m_index); }
void b2Fixture::SetRestitution(float v) { m_restitution = v; m_body->m_world->m_impl->touchFixture(m_index); }
void b2Fixture::SetRestitutionThreshold(float v) {
  m_restitutionThreshold = v;
  m_body->m_world->m_impl->touchFixture(m_index);
}
void b2Fixture::UpdateAABB() { m_shape->ComputeAABB(&m_aabb, m_body->GetTransform()); }
const b2AABB& b2Fixture::GetAABB() const {
  // static fixtures keep the AABB of their creation transform
  if (m_body->m_type != b2_staticBody) m_shape->ComputeAABB(&m_aabb, m_body->GetTransform());
  return m_aabb;
}

void b2Contact::GetWorldManifold(b2WorldManifold* wm) const {
  const b2Body* bodyA = m_fixtureA->GetBody();
  const b2Body* bodyB = m_fixtureB->GetBody();
  wm->Initialize(&m_manifold, bodyA->GetTransform(), m_fixtureA->GetShape()->m_radius, bodyB->GetTransform(),
                 m_fixtureB->GetShape()->m_radius);
}
void b2Contact::SetEnabled(bool flag) {
  if (flag) m_flags |= e_enabledFlag; else m_flags &= ~e_enabledFlag;
  m_overridden = true;
}
void b2Contact::SetFriction(float v) { m_friction = v; m_overridden = true; }
void b2Contact::ResetFriction() { SetFriction(b2MixFriction(m_fixtureA->m_friction, m_fixtureB->m_friction)); }
void b2Contact::SetRestitution(float v) { m_restitution = v; m_overridden = true; }
void b2Contact::ResetRestitution() {
  SetRestitution(b2MixRestitution(m_fixtureA->m_restitution, m_fixtureB->m_restitution));
}
void b2Contact::SetRestitutionThreshold(float v) { m_restitutionThreshold = v; m_overridden = true; }
void b2Contact::ResetRestitutionThreshold() {
  SetRestitutionThreshold(
      b2MixRestitutionThreshold(m_fixtureA->m_restitutionThreshold, m_fixtureB->m_restitutionThreshold));
}
void b2Contact::SetTangentSpeed(float v) { m_tangentSpeed = v; m_overridden = true; }

void b2World::SetContactFilter(b2ContactFilter* filter) {
  m_contactFilter = filter;
  m_newContacts = true;  // the next step starts with a pair refresh, after which the filter is consulted
}
bool b2ContactFilter::ShouldCollide(b2Fixture* fixtureA, b2Fixture* fixtureB) {
  const b2Filter& filterA = fixtureA->GetFilterData();
  const b2Filter& filterB = fixtureB->GetFilterData();
  if (filterA.groupIndex == filterB.groupIndex && filterA.groupIndex != 0) return filterA.groupIndex > 0;
  return (filterA.maskBits & filterB.categoryBits) != 0 && (filterA.categoryBits & filterB.maskBits) != 0;
}

b2Joint::b2Joint(const b2JointDef* def) {
  m_type = def->type;
  m_prev = m_next = nullptr;
  m_bodyA = def->bodyA;
  m_bodyB = def->bodyB;
  m_index = 0;
  m_collideConnected = def->collideConnected;
  m_userData = def->userData;
  m_edgeA.joint = m_edgeB.joint = nullptr;
  m_edgeA.other = m_edgeB.other = nullptr;
  m_edgeA.prev = m_edgeA.next = m_edgeB.prev = m_edgeB.next = nullptr;
}

void b2RevoluteJointDef::Initialize(b2Body* bA, b2Body* bB, const b2Vec2& anchor) {
  bodyA = bA;
  bodyB = bB;
  localAnchorA = bodyA->GetLocalPoint(anchor);
  localAnchorB = bodyB->GetLocalPoint(anchor);
  referenceAngle = bodyB->GetAngle() - bodyA->GetAngle();
}

b2RevoluteJoint::b2RevoluteJoint(const b2RevoluteJointDef* def) : b2Joint(def) {
  m_localAnchorA = def->localAnchorA;
  m_localAnchorB = def->localAnchorB;
  m_referenceAngle = def->referenceAngle;
  m_lowerAngle = def->lowerAngle;
  m_upperAngle = def->upperAngle;
  m_maxMotorTorque = def->maxMotorTorque;
  m_motorSpeed = def->motorSpeed;
  m_enableLimit = def->enableLimit;
  m_enableMotor = def->enableMotor;
  m_impulse.SetZero();
  m_motorImpulse = 0.0f;
  m_lowerImpulse = 0.0f;
  m_upperImpulse = 0.0f;
}
void b2Joint::Touch(bool wake) {
  b2WorldImpl* I = m_bodyA->GetWorld()->GetImpl();
  I->pullJoints();
  if (wake) {
    m_bodyA->SetAwake(true);
    m_bodyB->SetAwake(true);
  }
  I->jointsDirty = true;
}
void b2RevoluteJoint::WriteDevice(float* anchors, float* p, float* st) const {
  anchors[0] = m_localAnchorA.x; anchors[1] = m_localAnchorA.y; anchors[2] = m_localAnchorB.x; anchors[3] = m_localAnchorB.y;
  p[0] = m_referenceAngle; p[1] = m_lowerAngle; p[2] = m_upperAngle; p[3] = m_maxMotorTorque;
  p[4] = m_motorSpeed;
  uint32_t fl = (m_enableLimit ? 1u : 0u) | (m_enableMotor ? 2u : 0u) | (m_collideConnected ? 4u : 0u);  // type 0
  memcpy(&p[5], &fl, 4);
  p[6] = p[7] = 0.0f;
  st[0] = m_impulse.x; st[1] = m_impulse.y; st[2] = m_motorImpulse; st[3] = m_lowerImpulse; st[4] = m_upperImpulse;
}
void b2RevoluteJoint::ReadDeviceState(const float* st) {
  m_impulse.Set(st[0], st[1]);
  m_motorImpulse = st[2];
  m_lowerImpulse = st[3];
  m_upperImpulse = st[4];
}

// ---- b2PrismaticJoint (b2_prismatic_joint.cpp:77-112, 453-601) ----------------------------------------
void b2PrismaticJointDef::Initialize(b2Body* bA, b2Body* bB, const b2Vec2& anchor, const b2Vec2& axis) {
  bodyA = bA;
  bodyB = bB;
  localAnchorA = bodyA->GetLocalPoint(anchor);
  localAnchorB = bodyB->GetLocalPoint(anchor);
  localAxisA = bodyA->GetLocalVector(axis);
  referenceAngle = bodyB->GetAngle() - bodyA->GetAngle();
}
b2PrismaticJoint::b2PrismaticJoint(const b2PrismaticJointDef* def) : b2Joint(def) {
  m_localAnchorA = def->localAnchorA;
  m_localAnchorB = def->localAnchorB;
  m_localXAxisA = def->localAxisA;
  m_localXAxisA.Normalize();
  m_localYAxisA = b2Cross(1.0f, m_localXAxisA);
  m_referenceAngle = def->referenceAngle;
  m_impulse.SetZero();
  m_motorImpulse = 0.0f;
  m_lowerImpulse = 0.0f;
  m_upperImpulse = 0.0f;
  m_lowerTranslation = def->lowerTranslation;
  m_upperTranslation = def->upperTranslation;
  m_maxMotorForce = def->maxMotorForce;
  m_motorSpeed = def->motorSpeed;
  m_enableLimit = def->enableLimit;
  m_enableMotor = def->enableMotor;
}
void b2PrismaticJoint::WriteDevice(float* anchors, float* p, float* st) const {
  anchors[0] = m_localAnchorA.x; anchors[1] = m_localAnchorA.y; anchors[2] = m_localAnchorB.x; anchors[3] = m_localAnchorB.y;
  p[0] = m_referenceAngle; p[1] = m_lowerTranslation; p[2] = m_upperTranslation; p[3] = m_maxMotorForce;
  p[4] = m_motorSpeed;
  uint32_t fl = (m_enableLimit ? 1u : 0u) | (m_enableMotor ? 2u : 0u) | (m_collideConnected ? 4u : 0u) | (3u << 8);  // type 3
  memcpy(&p[5], &fl, 4);
  p[6] = m_localXAxisA.x; p[7] = m_localXAxisA.y;
  st[0] = m_impulse.x; st[1] = m_impulse.y; st[2] = m_motorImpulse; st[3] = m_lowerImpulse; st[4] = m_upperImpulse;
}
void b2PrismaticJoint::ReadDeviceState(const float* st) {
  m_impulse.Set(st[0], st[1]);
  m_motorImpulse = st[2];
  m_lowerImpulse = st[3];
  m_upperImpulse = st[4];
}
b2Vec2 b2PrismaticJoint::GetAnchorA() const { return m_bodyA->GetWorldPoint(m_localAnchorA); }
b2Vec2 b2PrismaticJoint::GetAnchorB() const { return m_bodyB->GetWorldPoint(m_localAnchorB); }
b2Vec2 b2PrismaticJoint::GetReactionForce(float inv_dt) const {
  m_bodyA->GetWorld()->GetImpl()->pullJoints();
  b2Vec2 axis = m_bodyA->GetWorldVector(m_localXAxisA), perp = m_bodyA->GetWorldVector(m_localYAxisA);
  return inv_dt * (m_impulse.x * perp + (m_motorImpulse + m_lowerImpulse - m_upperImpulse) * axis);
}
float b2PrismaticJoint::GetReactionTorque(float inv_dt) const {
  m_bodyA->GetWorld()->GetImpl()->pullJoints();
  return inv_dt * m_impulse.y;
}
float b2PrismaticJoint::GetMotorForce(float inv_dt) const {
  m_bodyA->GetWorld()->GetImpl()->pullJoints();
  return inv_dt * m_motorImpulse;
}
float b2PrismaticJoint::GetJointTranslation() const {
  b2Vec2 d = m_bodyB->GetWorldPoint(m_localAnchorB) - m_bodyA->GetWorldPoint(m_localAnchorA);
  return b2Dot(d, m_bodyA->GetWorldVector(m_localXAxisA));
}
float b2PrismaticJoint::GetJointSpeed() const {
  b2Vec2 rA = b2Mul(m_bodyA->GetTransform().q, m_localAnchorA - m_bodyA->GetLocalCenter());
  b2Vec2 rB = b2Mul(m_bodyB->GetTransform().q, m_localAnchorB - m_bodyB->GetLocalCenter());
  b2Vec2 d = (m_bodyB->GetWorldCenter() + rB) - (m_bodyA->GetWorldCenter() + rA);
  b2Vec2 axis = b2Mul(m_bodyA->GetTransform().q, m_localXAxisA);
  b2Vec2 vA = m_bodyA->GetLinearVelocity(), vB = m_bodyB->GetLinearVelocity();
  float wA = m_bodyA->GetAngularVelocity(), wB = m_bodyB->GetAngularVelocity();
  return b2Dot(d, b2Cross(wA, axis)) + b2Dot(axis, vB + b2Cross(wB, rB) - vA - b2Cross(wA, rA));
}
void b2PrismaticJoint::EnableLimit(bool flag) {
  if (flag == m_enableLimit) return;
  Touch();
  m_enableLimit = flag;
  m_lowerImpulse = 0.0f;
  m_upperImpulse = 0.0f;
}
void b2PrismaticJoint::SetLimits(float lower, float upper) {
  if (lower == m_lowerTranslation && upper == m_upperTranslation) return;
  Touch();
  m_lowerTranslation = lower;
  m_upperTranslation = upper;
  m_lowerImpulse = 0.0f;
  m_upperImpulse = 0.0f;
}
void b2PrismaticJoint::EnableMotor(bool flag) {
  if (flag == m_enableMotor) return;
  Touch();
  m_enableMotor = flag;
}
void b2PrismaticJoint::SetMotorSpeed(float speed) {
  if (speed == m_motorSpeed) return;
  Touch();
  m_motorSpeed = speed;
}
void b2PrismaticJoint::SetMaxMotorForce(float force) {
  if (force == m_maxMotorForce) return;
  Touch();
  m_maxMotorForce = force;
}

// ---- b2MouseJoint (b2_mouse_joint.cpp:35-75, 162-190) --------------------------------------------------
b2MouseJoint::b2MouseJoint(const b2MouseJointDef* def) : b2Joint(def) {
  m_targetA = def->target;
  m_localAnchorB = b2MulT(m_bodyB->GetTransform(), m_targetA);
  m_maxForce = def->maxForce;
  m_stiffness = def->stiffness;
  m_damping = def->damping;
  m_impulse.SetZero();
}
void b2MouseJoint::WriteDevice(float* anchors, float* p, float* st) const {
  anchors[0] = m_targetA.x; anchors[1] = m_targetA.y; anchors[2] = m_localAnchorB.x; anchors[3] = m_localAnchorB.y;
  p[0] = m_maxForce; p[1] = m_stiffness; p[2] = m_damping; p[3] = p[4] = 0.0f;
  uint32_t fl = (m_collideConnected ? 4u : 0u) | (7u << 8);  // type 7
  memcpy(&p[5], &fl, 4);
  for (int k = 6; k < 12; ++k) p[k] = 0.0f;
  st[0] = m_impulse.x; st[1] = m_impulse.y; st[2] = st[3] = st[4] = 0.0f;
}
void b2MouseJoint::ReadDeviceState(const float* st) { m_impulse.Set(st[0], st[1]); }
b2Vec2 b2MouseJoint::GetAnchorA() const { return m_targetA; }
b2Vec2 b2MouseJoint::GetAnchorB() const { return m_bodyB->GetWorldPoint(m_localAnchorB); }
b2Vec2 b2MouseJoint::GetReactionForce(float inv_dt) const {
  m_bodyA->GetWorld()->GetImpl()->pullJoints();
  return inv_dt * m_impulse;
}
float b2MouseJoint::GetReactionTorque(float inv_dt) const { return inv_dt * 0.0f; }
void b2MouseJoint::SetTarget(const b2Vec2& target) {
  if (target.x == m_targetA.x && target.y == m_targetA.y) return;
  Touch(false);
  m_bodyB->SetAwake(true);  // only the dragged body is woken (b2_mouse_joint.cpp:49-56)
  m_targetA = target;
}
void b2MouseJoint::SetMaxForce(float force) {
  Touch(false);
  m_maxForce = force;
}
void b2MouseJoint::SetStiffness(float stiffness) {
  Touch(false);
  m_stiffness = stiffness;
}
void b2MouseJoint::SetDamping(float damping) {
  Touch(false);
  m_damping = damping;
}

// ---- b2FrictionJoint (b2_friction_joint.cpp:39-63, 183-226) -------------------------------------------
void b2FrictionJointDef::Initialize(b2Body* bA, b2Body* bB, const b2Vec2& anchor) {
  bodyA = bA;
  bodyB = bB;
  localAnchorA = bodyA->GetLocalPoint(anchor);
  localAnchorB = bodyB->GetLocalPoint(anchor);
}
b2FrictionJoint::b2FrictionJoint(const b2FrictionJointDef* def) : b2Joint(def) {
  m_localAnchorA = def->localAnchorA;
  m_localAnchorB = def->localAnchorB;
  m_linearImpulse.SetZero();
  m_angularImpulse = 0.0f;
  m_maxForce = def->maxForce;
  m_maxTorque = def->maxTorque;
}
void b2FrictionJoint::WriteDevice(float* anchors, float* p, float* st) const {
  anchors[0] = m_localAnchorA.x; anchors[1] = m_localAnchorA.y; anchors[2] = m_localAnchorB.x; anchors[3] = m_localAnchorB.y;
  p[0] = m_maxForce; p[1] = m_maxTorque; p[2] = p[3] = p[4] = 0.0f;
  uint32_t fl = (m_collideConnected ? 4u : 0u) | (5u << 8);  // type 5
  memcpy(&p[5], &fl, 4);
  for (int k = 6; k < 12; ++k) p[k] = 0.0f;
  st[0] = m_linearImpulse.x; st[1] = m_linearImpulse.y; st[2] = m_angularImpulse; st[3] = st[4] = 0.0f;
}
void b2FrictionJoint::ReadDeviceState(const float* st) {
  m_linearImpulse.Set(st[0], st[1]);
  m_angularImpulse = st[2];
}
b2Vec2 b2FrictionJoint::GetAnchorA() const { return m_bodyA->GetWorldPoint(m_localAnchorA); }
b2Vec2 b2FrictionJoint::GetAnchorB() const { return m_bodyB->GetWorldPoint(m_localAnchorB); }
b2Vec2 b2FrictionJoint::GetReactionForce(float inv_dt) const {
  m_bodyA->GetWorld()->GetImpl()->pullJoints();
  return inv_dt * m_linearImpulse;
}
float b2FrictionJoint::GetReactionTorque(float inv_dt) const {
  m_bodyA->GetWorld()->GetImpl()->pullJoints();
  return inv_dt * m_angularImpulse;
}
void b2FrictionJoint::SetMaxForce(float force) {
  Touch(false);  // the reference's setters do not wake the bodies
  m_maxForce = force;
}
void b2FrictionJoint::SetMaxTorque(float torque) {
  Touch(false);
  m_maxTorque = torque;
}

// ---- b2MotorJoint (b2_motor_joint.cpp:40-68, 210-297) -------------------------------------------------
void b2MotorJointDef::Initialize(b2Body* bA, b2Body* bB) {
  bodyA = bA;
  bodyB = bB;
  b2Vec2 xB = bodyB->GetPosition();
  linearOffset = bodyA->GetLocalPoint(xB);
  angularOffset = bodyB->GetAngle() - bodyA->GetAngle();
}
b2MotorJoint::b2MotorJoint(const b2MotorJointDef* def) : b2Joint(def) {
  m_linearOffset = def->linearOffset;
  m_angularOffset = def->angularOffset;
  m_linearImpulse.SetZero();
  m_angularImpulse = 0.0f;
  m_maxForce = def->maxForce;
  m_maxTorque = def->maxTorque;
  m_correctionFactor = def->correctionFactor;
}
void b2MotorJoint::WriteDevice(float* anchors, float* p, float* st) const {
  anchors[0] = m_linearOffset.x; anchors[1] = m_linearOffset.y; anchors[2] = anchors[3] = 0.0f;
  p[0] = m_maxForce; p[1] = m_maxTorque; p[2] = m_correctionFactor; p[3] = m_angularOffset; p[4] = 0.0f;
  uint32_t fl = (m_collideConnected ? 4u : 0u) | (6u << 8);  // type 6
  memcpy(&p[5], &fl, 4);
  for (int k = 6; k < 12; ++k) p[k] = 0.0f;
  st[0] = m_linearImpulse.x; st[1] = m_linearImpulse.y; st[2] = m_angularImpulse; st[3] = st[4] = 0.0f;
}
void b2MotorJoint::ReadDeviceState(const float* st) {
  m_linearImpulse.Set(st[0], st[1]);
  m_angularImpulse = st[2];
}
b2Vec2 b2MotorJoint::GetAnchorA() const { return m_bodyA->GetPosition(); }
b2Vec2 b2MotorJoint::GetAnchorB() const { return m_bodyB->GetPosition(); }
b2Vec2 b2MotorJoint::GetReactionForce(float inv_dt) const {
  m_bodyA->GetWorld()->GetImpl()->pullJoints();
  return inv_dt * m_linearImpulse;
}
float b2MotorJoint::GetReactionTorque(float inv_dt) const {
  m_bodyA->GetWorld()->GetImpl()->pullJoints();
  return inv_dt * m_angularImpulse;
}
void b2MotorJoint::SetLinearOffset(const b2Vec2& linearOffset) {
  if (linearOffset.x == m_linearOffset.x && linearOffset.y == m_linearOffset.y) return;
  Touch();
  m_linearOffset = linearOffset;
}
void b2MotorJoint::SetAngularOffset(float angularOffset) {
  if (angularOffset == m_angularOffset) return;
  Touch();
  m_angularOffset = angularOffset;
}
void b2MotorJoint::SetMaxForce(float force) {
  Touch(false);
  m_maxForce = force;
}
void b2MotorJoint::SetMaxTorque(float torque) {
  Touch(false);
  m_maxTorque = torque;
}
void b2MotorJoint::SetCorrectionFactor(float factor) {
  Touch(false);
  m_correctionFactor = factor;
}

// ---- b2WheelJoint (b2_wheel_joint.cpp:40-85, 448-628) -------------------------------------------------
void b2WheelJointDef::Initialize(b2Body* bA, b2Body* bB, const b2Vec2& anchor, const b2Vec2& axis) {
  bodyA = bA;
  bodyB = bB;
  localAnchorA = bodyA->GetLocalPoint(anchor);
  localAnchorB = bodyB->GetLocalPoint(anchor);
  localAxisA = bodyA->GetLocalVector(axis);
}
b2WheelJoint::b2WheelJoint(const b2WheelJointDef* def) : b2Joint(def) {
  m_localAnchorA = def->localAnchorA;
  m_localAnchorB = def->localAnchorB;
  m_localXAxisA = def->localAxisA;  // not normalised, as in the reference
  m_localYAxisA = b2Cross(1.0f, m_localXAxisA);
  m_impulse = m_springImpulse = m_motorImpulse = m_lowerImpulse = m_upperImpulse = 0.0f;
  m_lowerTranslation = def->lowerTranslation;
  m_upperTranslation = def->upperTranslation;
  m_enableLimit = def->enableLimit;
  m_maxMotorTorque = def->maxMotorTorque;
  m_motorSpeed = def->motorSpeed;
  m_enableMotor = def->enableMotor;
  m_stiffness = def->stiffness;
  m_damping = def->damping;
}
void b2WheelJoint::WriteDevice(float* anchors, float* p, float* st) const {
  anchors[0] = m_localAnchorA.x; anchors[1] = m_localAnchorA.y; anchors[2] = m_localAnchorB.x; anchors[3] = m_localAnchorB.y;
  p[0] = m_stiffness; p[1] = m_lowerTranslation; p[2] = m_upperTranslation; p[3] = m_maxMotorTorque;
  p[4] = m_motorSpeed;
  uint32_t fl = (m_enableLimit ? 1u : 0u) | (m_enableMotor ? 2u : 0u) | (m_collideConnected ? 4u : 0u) | (4u << 8);  // type 4
  memcpy(&p[5], &fl, 4);
  p[6] = m_localXAxisA.x; p[7] = m_localXAxisA.y;
  p[8] = m_damping; p[9] = p[10] = p[11] = 0.0f;
  st[0] = m_impulse; st[1] = m_springImpulse; st[2] = m_motorImpulse; st[3] = m_lowerImpulse; st[4] = m_upperImpulse;
}
void b2WheelJoint::ReadDeviceState(const float* st) {
  m_impulse = st[0];
  m_springImpulse = st[1];
  m_motorImpulse = st[2];
  m_lowerImpulse = st[3];
  m_upperImpulse = st[4];
}
b2Vec2 b2WheelJoint::GetAnchorA() const { return m_bodyA->GetWorldPoint(m_localAnchorA); }
b2Vec2 b2WheelJoint::GetAnchorB() const { return m_bodyB->GetWorldPoint(m_localAnchorB); }
b2Vec2 b2WheelJoint::GetReactionForce(float inv_dt) const {
  m_bodyA->GetWorld()->GetImpl()->pullJoints();
  b2Vec2 ax = m_bodyA->GetWorldVector(m_localXAxisA), ay = m_bodyA->GetWorldVector(m_localYAxisA);
  return inv_dt * (m_impulse * ay + (m_springImpulse + m_lowerImpulse - m_upperImpulse) * ax);
}
float b2WheelJoint::GetReactionTorque(float inv_dt) const {
  m_bodyA->GetWorld()->GetImpl()->pullJoints();
  return inv_dt * m_motorImpulse;
}
float b2WheelJoint::GetMotorTorque(float inv_dt) const {
  m_bodyA->GetWorld()->GetImpl()->pullJoints();
  return inv_dt * m_motorImpulse;
}
float b2WheelJoint::GetJointTranslation() const {
  b2Vec2 d = m_bodyB->GetWorldPoint(m_localAnchorB) - m_bodyA->GetWorldPoint(m_localAnchorA);
  return b2Dot(d, m_bodyA->GetWorldVector(m_localXAxisA));
}
float b2WheelJoint::GetJointLinearSpeed() const {
  b2Vec2 rA = b2Mul(m_bodyA->GetTransform().q, m_localAnchorA - m_bodyA->GetLocalCenter());
  b2Vec2 rB = b2Mul(m_bodyB->GetTransform().q, m_localAnchorB - m_bodyB->GetLocalCenter());
  b2Vec2 d = (m_bodyB->GetWorldCenter() + rB) - (m_bodyA->GetWorldCenter() + rA);
  b2Vec2 axis = b2Mul(m_bodyA->GetTransform().q, m_localXAxisA);
  b2Vec2 vA = m_bodyA->GetLinearVelocity(), vB = m_bodyB->GetLinearVelocity();
  float wA = m_bodyA->GetAngularVelocity(), wB = m_bodyB->GetAngularVelocity();
  return b2Dot(d, b2Cross(wA, axis)) + b2Dot(axis, vB + b2Cross(wB, rB) - vA - b2Cross(wA, rA));
}
float b2WheelJoint::GetJointAngle() const { return m_bodyB->GetAngle() - m_bodyA->GetAngle(); }
float b2WheelJoint::GetJointAngularSpeed() const { return m_bodyB->GetAngularVelocity() - m_bodyA->GetAngularVelocity(); }
void b2WheelJoint::EnableLimit(bool flag) {
  if (flag == m_enableLimit) return;
  Touch();
  m_enableLimit = flag;
  m_lowerImpulse = 0.0f;
  m_upperImpulse = 0.0f;
}
void b2WheelJoint::SetLimits(float lower, float upper) {
  if (lower == m_lowerTranslation && upper == m_upperTranslation) return;
  Touch();
  m_lowerTranslation = lower;
  m_upperTranslation = upper;
  m_lowerImpulse = 0.0f;
  m_upperImpulse = 0.0f;
}
void b2WheelJoint::EnableMotor(bool flag) {
  if (flag == m_enableMotor) return;
  Touch();
  m_enableMotor = flag;
}
void b2WheelJoint::SetMotorSpeed(float speed) {
  if (speed == m_motorSpeed) return;
  Touch();
  m_motorSpeed = speed;
}
void b2WheelJoint::SetMaxMotorTorque(float torque) {
  if (torque == m_maxMotorTorque) return;
  Touch();
  m_maxMotorTorque = torque;
}
void b2WheelJoint::SetStiffness(float stiffness) {
  Touch(false);  // the reference's spring setters do not wake the bodies
  m_stiffness = stiffness;
}
void b2WheelJoint::SetDamping(float damping) {
  Touch(false);
  m_damping = damping;
}

// ---- b2WeldJoint (b2_weld_joint.cpp:38-60, 307-330) ---------------------------------------------------
void b2WeldJointDef::Initialize(b2Body* bA, b2Body* bB, const b2Vec2& anchor) {
  bodyA = bA;
  bodyB = bB;
  localAnchorA = bodyA->GetLocalPoint(anchor);
  localAnchorB = bodyB->GetLocalPoint(anchor);
  referenceAngle = bodyB->GetAngle() - bodyA->GetAngle();
}
b2WeldJoint::b2WeldJoint(const b2WeldJointDef* def) : b2Joint(def) {
  m_localAnchorA = def->localAnchorA;
  m_localAnchorB = def->localAnchorB;
  m_referenceAngle = def->referenceAngle;
  m_stiffness = def->stiffness;
  m_damping = def->damping;
  m_impulse[0] = m_impulse[1] = m_impulse[2] = 0.0f;
}
void b2WeldJoint::WriteDevice(float* anchors, float* p, float* st) const {
  anchors[0] = m_localAnchorA.x; anchors[1] = m_localAnchorA.y; anchors[2] = m_localAnchorB.x; anchors[3] = m_localAnchorB.y;
  p[0] = m_referenceAngle; p[1] = m_stiffness; p[2] = m_damping; p[3] = 0.0f;
  p[4] = 0.0f;
  uint32_t fl = (m_collideConnected ? 4u : 0u) | (2u << 8);  // type 2 = weld
  memcpy(&p[5], &fl, 4);
  p[6] = p[7] = 0.0f;
  st[0] = m_impulse[0]; st[1] = m_impulse[1]; st[2] = m_impulse[2]; st[3] = 0.0f; st[4] = 0.0f;
}
void b2WeldJoint::ReadDeviceState(const float* st) {
  m_impulse[0] = st[0];
  m_impulse[1] = st[1];
  m_impulse[2] = st[2];
}
b2Vec2 b2WeldJoint::GetAnchorA() const { return m_bodyA->GetWorldPoint(m_localAnchorA); }
b2Vec2 b2WeldJoint::GetAnchorB() const { return m_bodyB->GetWorldPoint(m_localAnchorB); }
b2Vec2 b2WeldJoint::GetReactionForce(float inv_dt) const {
  m_bodyA->GetWorld()->GetImpl()->pullJoints();
  return inv_dt * b2Vec2(m_impulse[0], m_impulse[1]);
}
float b2WeldJoint::GetReactionTorque(float inv_dt) const {
  m_bodyA->GetWorld()->GetImpl()->pullJoints();
  return inv_dt * m_impulse[2];
}
void b2WeldJoint::SetStiffness(float stiffness) {
  Touch(false);
  m_stiffness = stiffness;
}
void b2WeldJoint::SetDamping(float damping) {
  Touch(false);
  m_damping = damping;
}

// ---- b2DistanceJoint (b2_distance_joint.cpp:44-74, 305-366) -----------------------------------------
void b2LinearStiffness(float& stiffness, float& damping, float frequencyHertz, float dampingRatio, const b2Body* bodyA,
                       const b2Body* bodyB) {
  const float massA = bodyA->GetMass(), massB = bodyB->GetMass();
  float mass = massB;  // reduced mass of the pair, or the only finite one
  if (massA > 0.0f && massB > 0.0f) mass = massA * massB / (massA + massB);
  else if (massA > 0.0f) mass = massA;
  const float omega = 2.0f * b2_pi * frequencyHertz;
  stiffness = mass * omega * omega;
  damping = 2.0f * mass * dampingRatio * omega;
}
void b2AngularStiffness(float& stiffness, float& damping, float frequencyHertz, float dampingRatio, const b2Body* bodyA,
                        const b2Body* bodyB) {
  const float IA = bodyA->GetInertia(), IB = bodyB->GetInertia();
  float I = IB;
  if (IA > 0.0f && IB > 0.0f) I = IA * IB / (IA + IB);
  else if (IA > 0.0f) I = IA;
  const float omega = 2.0f * b2_pi * frequencyHertz;
  stiffness = I * omega * omega;
  damping = 2.0f * I * dampingRatio * omega;
}
void b2DistanceJointDef::Initialize(b2Body* b1, b2Body* b2, const b2Vec2& anchor1, const b2Vec2& anchor2) {
  bodyA = b1;
  bodyB = b2;
  localAnchorA = bodyA->GetLocalPoint(anchor1);
  localAnchorB = bodyB->GetLocalPoint(anchor2);
  length = b2Max((anchor2 - anchor1).Length(), b2_linearSlop);
  minLength = length;
  maxLength = length;
}
b2DistanceJoint::b2DistanceJoint(const b2DistanceJointDef* def) : b2Joint(def) {
  m_localAnchorA = def->localAnchorA;
  m_localAnchorB = def->localAnchorB;
  m_length = b2Max(def->length, b2_linearSlop);
  m_minLength = b2Max(def->minLength, b2_linearSlop);
  m_maxLength = b2Max(def->maxLength, m_minLength);
  m_stiffness = def->stiffness;
  m_damping = def->damping;
  m_impulse = m_lowerImpulse = m_upperImpulse = 0.0f;
}
void b2DistanceJoint::WriteDevice(float* anchors, float* p, float* st) const {
  anchors[0] = m_localAnchorA.x; anchors[1] = m_localAnchorA.y; anchors[2] = m_localAnchorB.x; anchors[3] = m_localAnchorB.y;
  p[0] = m_length; p[1] = m_minLength; p[2] = m_maxLength; p[3] = m_stiffness;
  p[4] = m_damping;
  uint32_t fl = (m_collideConnected ? 4u : 0u) | (1u << 8);  // type 1 = distance
  memcpy(&p[5], &fl, 4);
  p[6] = p[7] = 0.0f;
  st[0] = m_impulse; st[1] = 0.0f; st[2] = 0.0f; st[3] = m_lowerImpulse; st[4] = m_upperImpulse;
}
void b2DistanceJoint::ReadDeviceState(const float* st) {
  m_impulse = st[0];
  m_lowerImpulse = st[3];
  m_upperImpulse = st[4];
}
b2Vec2 b2DistanceJoint::GetAnchorA() const { return m_bodyA->GetWorldPoint(m_localAnchorA); }
b2Vec2 b2DistanceJoint::GetAnchorB() const { return m_bodyB->GetWorldPoint(m_localAnchorB); }
b2Vec2 b2DistanceJoint::GetReactionForce(float inv_dt) const {
  m_bodyA->GetWorld()->GetImpl()->pullJoints();
  b2Vec2 u = GetAnchorB() - GetAnchorA();
  u.Normalize();
  return inv_dt * (m_impulse + m_lowerImpulse - m_upperImpulse) * u;
}
float b2DistanceJoint::GetReactionTorque(float) const { return 0.0f; }
float b2DistanceJoint::GetCurrentLength() const { return (GetAnchorB() - GetAnchorA()).Length(); }
float b2DistanceJoint::SetLength(float length) {
  Touch(false);  // the reference's setters do not wake the bodies
  m_impulse = 0.0f;
  m_length = b2Max(b2_linearSlop, length);
  return m_length;
}
float b2DistanceJoint::SetMinLength(float minLength) {
  Touch(false);  // the reference's setters do not wake the bodies
  m_lowerImpulse = 0.0f;
  m_minLength = b2Clamp(minLength, b2_linearSlop, m_maxLength);
  return m_minLength;
}
float b2DistanceJoint::SetMaxLength(float maxLength) {
  Touch(false);  // the reference's setters do not wake the bodies
  m_upperImpulse = 0.0f;
  m_maxLength = b2Max(maxLength, m_minLength);
  return m_maxLength;
}
void b2DistanceJoint::SetStiffness(float stiffness) {
  Touch(false);  // the reference's setters do not wake the bodies
  m_stiffness = stiffness;
}
void b2DistanceJoint::SetDamping(float damping) {
  Touch(false);  // the reference's setters do not wake the bodies
  m_damping = damping;
}
b2Vec2 b2RevoluteJoint::GetReactionForce(float inv_dt) const {
  m_bodyA->GetWorld()->GetImpl()->pullJoints();
  return inv_dt * m_impulse;
}
float b2RevoluteJoint::GetReactionTorque(float inv_dt) const {
  m_bodyA->GetWorld()->GetImpl()->pullJoints();
  return inv_dt * (m_motorImpulse + m_lowerImpulse - m_upperImpulse);
}
float b2RevoluteJoint::GetMotorTorque(float inv_dt) const {
  m_bodyA->GetWorld()->GetImpl()->pullJoints();
  return inv_dt * m_motorImpulse;
}
void b2RevoluteJoint::EnableLimit(bool flag) {
  if (flag == m_enableLimit) return;
  Touch();
  m_enableLimit = flag;
  m_lowerImpulse = 0.0f;
  m_upperImpulse = 0.0f;
}
void b2RevoluteJoint::SetLimits(float lower, float upper) {
  if (lower == m_lowerAngle && upper == m_upperAngle) return;
  Touch();
  m_lowerImpulse = 0.0f;
  m_upperImpulse = 0.0f;
  m_lowerAngle = lower;
  m_upperAngle = upper;
}
void b2RevoluteJoint::EnableMotor(bool flag) {
  if (flag == m_enableMotor) return;
  Touch();
  m_enableMotor = flag;
}
void b2RevoluteJoint::SetMotorSpeed(float speed) {
  if (speed == m_motorSpeed) return;
  Touch();
  m_motorSpeed = speed;
}
void b2RevoluteJoint::SetMaxMotorTorque(float torque) {
  if (torque == m_maxMotorTorque) return;
  Touch();
  m_maxMotorTorque = torque;
}
b2Vec2 b2RevoluteJoint::GetAnchorA() const { return m_bodyA->GetWorldPoint(m_localAnchorA); }
b2Vec2 b2RevoluteJoint::GetAnchorB() const { return m_bodyB->GetWorldPoint(m_localAnchorB); }
float b2RevoluteJoint::GetJointAngle() const { return m_bodyB->GetAngle() - m_bodyA->GetAngle() - m_referenceAngle; }
float b2RevoluteJoint::GetJointSpeed() const { return m_bodyB->GetAngularVelocity() - m_bodyA->GetAngularVelocity(); }

/*
 * b2cuda.h — C-ABI of the B200-native b2World::Step hot path.
 *
 * This is the drop-in boundary: a host program (the C++ b2World/b2Body/b2Fixture
 * mirror under include/box2d/, a ctypes binding, or a maintainer's patch to the
 * reference itself) hands body / fixture / shape state over as
 * plain structure-of-arrays host pointers, calls b2g_step(), and reads state back.
 * No torch types, no C++ types, no device pointers cross this boundary.
 *
 * The reference has no FFI layer of its own; every entry point below
 * cites the reference interface (file:line under /root/reference) whose work it
 * replaces.
 *
 * All functions return 0 (B2G_OK) on success or a negative B2G_ERR_* code.  There is
 * NO CPU fallback: if no CUDA device is usable, b2g_arena_create() fails loudly.
 */
#ifndef B2CUDA_H
#define B2CUDA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B2G_OK 0
#define B2G_ERR_INVALID (-1)  /* bad argument */
#define B2G_ERR_CUDA (-2)     /* CUDA runtime error, see b2g_last_error() */
#define B2G_ERR_CAPACITY (-3) /* an arena capacity was exceeded (never silently truncated) */
#define B2G_ERR_NO_DEVICE (-4)

/* ---- body flags (bit layout mirrors include/box2d/b2_body.h:476-485 where it overlaps) */
#define B2G_BODY_AWAKE 0x0002u
#define B2G_BODY_AUTOSLEEP 0x0004u
#define B2G_BODY_BULLET 0x0008u
#define B2G_BODY_FIXED_ROTATION 0x0010u
#define B2G_BODY_ENABLED 0x0020u
#define B2G_BODY_WAKE_REQUEST 0x0100u /* device-internal: SetAwake(true) requested during Collide */
#define B2G_BODY_TYPE_SHIFT 16        /* bits 16-17: 0 static, 1 kinematic, 2 dynamic (b2_body.h:45-50) */
#define B2G_BODY_TYPE_MASK 0x30000u

/* ---- shape types (include/box2d/b2_shape.h:52-59) */
#define B2G_SHAPE_CIRCLE 0
#define B2G_SHAPE_EDGE 1
#define B2G_SHAPE_POLYGON 2

/* ---- fixture type_flags word */
#define B2G_FIX_TYPE_MASK 0x3u
#define B2G_FIX_SENSOR 0x100u
#define B2G_FIX_DEAD 0x200u /* destroyed fixture: ignored by the broadphase */

/* ---- contact flags (include/box2d/b2_contact.h:155-173) */
#define B2G_CONTACT_ALIVE 0x0001u /* device-internal: the slot holds a live contact */
#define B2G_CONTACT_TOUCHING 0x0008u
#define B2G_CONTACT_ENABLED 0x0010u

/* ---- manifold types (include/box2d/b2_collision.h:103-108) */
#define B2G_MANIFOLD_CIRCLES 0
#define B2G_MANIFOLD_FACE_A 1
#define B2G_MANIFOLD_FACE_B 2

/* ---- solver modes */
#define B2G_SOLVER_COLOURED 0   /* production: graph-coloured Gauss-Seidel */
#define B2G_SOLVER_SEQUENTIAL 1 /* deterministic single-colour sequential order (parity vehicle) */

typedef struct b2gArena b2gArena; /* opaque: all body/fixture/contact state in HBM */

/* Capacities of one arena.  An arena holds num_worlds independent worlds concatenated
 * into one SoA (bodies of different worlds never collide); a plain b2World is an arena
 * with num_worlds == 1.  Replaces the allocators owned by b2World
 * (src/dynamics/b2_world.cpp:42-73). */
typedef struct b2gArenaDef {
  int32_t device;        /* CUDA device ordinal */
  int32_t num_worlds;    /* >= 1 */
  int32_t max_bodies;
  int32_t max_fixtures;
  int32_t max_shape_quads; /* shape pool size in float4 units, see "shape pool" below */
  int32_t max_contacts;    /* AABB-overlapping fixture pairs */
  int32_t max_joints;
  int32_t reserved;
} b2gArenaDef;

/* Structure-of-arrays views.  Any pointer may be NULL = "leave that array alone".
 * Layouts (all little-endian float32 / int32):
 *   pos    [n][4] = sweep.c.x, sweep.c.y, sweep.a, 0            (b2_body.h:506)
 *   vel    [n][4] = linearVelocity.x, .y, angularVelocity, 0     (b2_body.h:508-509)
 *   xf     [n][4] = xf.p.x, xf.p.y, xf.q.s, xf.q.c               (b2_body.h:504)
 *   mass   [n][4] = invMass, invI, mass, gravityScale            (b2_body.h:527-539)
 *   center [n][4] = sweep.localCenter.x, .y, linearDamping, angularDamping
 *   force  [n][4] = force.x, force.y, torque, sleepTime          (b2_body.h:511-512,542)
 *   flags  [n]    = B2G_BODY_* | type << B2G_BODY_TYPE_SHIFT
 *   world  [n]    = world id inside the arena, 0 <= id < num_worlds
 */
typedef struct b2gBodyArrays {
  float* pos;
  float* vel;
  float* xf;
  float* mass;
  float* center;
  float* force;
  uint32_t* flags;
  int32_t* world;
} b2gBodyArrays;

/*   body       [n]    = owning body index                        (b2_fixture.h:249)
 *   shape_off  [n]    = offset of the shape record in the shape pool (float4 units)
 *   type_flags [n]    = B2G_SHAPE_* | B2G_FIX_SENSOR | B2G_FIX_DEAD
 *   filter     [n][2] = categoryBits | maskBits << 16 , groupIndex (sign-extended int16)
 *                                                                (b2_fixture.h:37-57)
 *   material   [n][4] = friction, restitution, restitutionThreshold, density
 *                                                                (b2_fixture.h:243-268)
 * Shape pool records (float4 units):
 *   circle : { p.x, p.y, radius, 0 }                              (b2_circle_shape.h:57)
 *   edge   : { v1.x, v1.y, v2.x, v2.y } { v0.x, v0.y, v3.x, v3.y } { radius, oneSided, 0, 0 }
 *                                                                (b2_edge_shape.h:66-72)
 *   polygon: { centroid.x, centroid.y, radius, count } then count x { v.x, v.y, n.x, n.y }
 *                                                                (b2_polygon_shape.h:81-84)
 */
typedef struct b2gFixtureArrays {
  int32_t* body;
  int32_t* shape_off;
  uint32_t* type_flags;
  uint32_t* filter;
  float* material;
} b2gFixtureArrays;

/* Contacts as the device holds them (one per AABB-overlapping, filter-passing pair;
 * b2_contact.h:62-64).  manifold [n][16] =
 *   localNormal.xy, localPoint.xy,
 *   p0.localPoint.xy, p0.normalImpulse, p0.tangentImpulse,
 *   p1.localPoint.xy, p1.normalImpulse, p1.tangentImpulse,
 *   bits(p0.id.key), bits(p1.id.key), bits(type), bits(pointCount)   (b2_collision.h:75-117)
 * material [n][4] = friction, restitution, restitutionThreshold, tangentSpeed (b2_contact.h:205-216)
 */
typedef struct b2gContactArrays {
  int32_t* fixture_a;
  int32_t* fixture_b;
  uint32_t* flags;
  float* manifold;
  float* material;
  int32_t* colour; /* solver colour of the last step, -1 = not in the solver */
} b2gContactArrays;

/* Joints, SURVEY §8(f) rank 1: revolute (src/dynamics/b2_revolute_joint.cpp:73-321), distance
 * (src/dynamics/b2_distance_joint.cpp:76-303), weld (src/dynamics/b2_weld_joint.cpp:62-305),
 * prismatic (src/dynamics/b2_prismatic_joint.cpp:114-451), wheel (src/dynamics/b2_wheel_joint.cpp:87-446),
 * friction (src/dynamics/b2_friction_joint.cpp:65-181), motor (src/dynamics/b2_motor_joint.cpp:70-208) and
 * mouse (src/dynamics/b2_mouse_joint.cpp:77-160).  The type sits in bits 8-11 of the flags word.
 *   bodies  [n][2] = bodyA, bodyB
 *   anchors [n][4] = localAnchorA.xy, localAnchorB.xy
 *   params  [n][12] (unused trailing entries 0), revolute (type 0): referenceAngle, lowerAngle, upperAngle, maxMotorTorque,
 *                    motorSpeed, bits(flags), 0, 0
 *                  distance (type 1): length, minLength, maxLength, stiffness, damping,
 *                    bits(flags | 1 << 8), 0, 0
 *                  weld (type 2): referenceAngle, stiffness, damping, 0, 0, bits(flags | 2 << 8), 0, 0
 *                  prismatic (type 3): referenceAngle, lowerTranslation, upperTranslation, maxMotorForce,
 *                    motorSpeed, bits(flags | 3 << 8), localXAxisA.x, localXAxisA.y (unit length)
 *                  wheel (type 4): stiffness, lowerTranslation, upperTranslation, maxMotorTorque, motorSpeed,
 *                    bits(flags | 4 << 8), localXAxisA.x, localXAxisA.y, damping, 0, 0, 0
 *                  mouse (type 7): maxForce, stiffness, damping, 0, 0, bits(flags | 7 << 8), 0...; its anchors record
 *                    holds targetA.xy, localAnchorB.xy
 *                  friction (type 5): maxForce, maxTorque, 0, 0, 0, bits(flags | 5 << 8), 0...
 *                  motor (type 6): maxForce, maxTorque, correctionFactor, angularOffset, 0, bits(flags | 6 << 8), 0...;
 *                    its anchors record holds linearOffset.xy, 0, 0
 *                  flags: 1 enableLimit, 2 enableMotor, 4 collideConnected
 *   state   [n][5], revolute: m_impulse.x, m_impulse.y, m_motorImpulse, m_lowerImpulse, m_upperImpulse
 *                    (include/box2d/b2_revolute_joint.h:178-181)
 *                  distance: m_impulse, 0, 0, m_lowerImpulse, m_upperImpulse
 *                    (include/box2d/b2_distance_joint.h:157-159)
 *                  weld: m_impulse.x, .y, .z, 0, 0 (include/box2d/b2_weld_joint.h:112)
 *                  mouse: m_impulse.x, .y, 0, 0, 0 (include/box2d/b2_mouse_joint.h:113)
 *                  friction, motor: m_linearImpulse.x, .y, m_angularImpulse, 0, 0
 *                    (include/box2d/b2_friction_joint.h:86-87, b2_motor_joint.h:104-105)
 *                  wheel: m_impulse, m_springImpulse, m_motorImpulse, m_lowerImpulse, m_upperImpulse
 *                    (include/box2d/b2_wheel_joint.h:196-200)
 *                  prismatic: m_impulse.x, m_impulse.y, m_motorImpulse, m_lowerImpulse, m_upperImpulse
 *                    (include/box2d/b2_prismatic_joint.h:164-167)
 *                  the warm-start accumulators.  NULL on upload = start from zero, as the joints'
 *                  constructors do.
 */
typedef struct b2gJointArrays {
  int32_t* bodies;
  float* anchors;
  float* params;
  float* state;
} b2gJointArrays;

/* Per-step parameters = b2TimeStep (include/box2d/b2_time_step.h:39-48) + the world
 * flags b2World::Step reads (src/dynamics/b2_world.cpp:1108-1171). */
typedef struct b2gStepParams {
  float dt;
  int32_t velocity_iterations;
  int32_t position_iterations;
  float gravity_x, gravity_y;
  int32_t warm_starting; /* b2World::SetWarmStarting */
  int32_t allow_sleep;   /* b2World::SetAllowSleeping */
  int32_t clear_forces;  /* b2World::SetAutoClearForces */
  int32_t solver_mode;   /* B2G_SOLVER_* */
  int32_t record_events; /* 1: fill the begin/end-contact event lists */
} b2gStepParams;

/* Filled by b2g_step (b2Profile, include/box2d/b2_time_step.h:29-36, plus counters). */
typedef struct b2gStepStats {
  int32_t num_bodies, num_fixtures, num_contacts, num_touching;
  int32_t num_constraints; /* touching contacts in awake islands = solver rows */
  int32_t num_colours;     /* colours used by the production solver this step */
  int32_t num_overflow;    /* constraints that fell into the serial overflow colour */
  int32_t num_awake;       /* awake non-static bodies after the step */
  int32_t num_pairs;       /* candidate pairs reported by the broadphase */
  int32_t colour_rounds;   /* colouring rounds launched */
  int32_t num_launches;    /* kernels launched by this step (ours; excludes CUB's) */
  int32_t bp_max_visits;   /* tree nodes visited by the longest pair-finder walk (b2World::GetTreeHeight's role: tree quality) */
  float ms_collide, ms_solve, ms_broadphase, ms_step; /* CUDA-event times, 0 unless profiling on */
  float bp_mean_visits;    /* tree nodes visited per fixture by the pair finder */
  int32_t bp_rebuilt;      /* 1 when this step re-sorted and rebuilt the tree instead of refitting it */
} b2gStepStats;

const char* b2g_last_error(void);
int b2g_device_count(void);

/* b2World::b2World / ~b2World (src/dynamics/b2_world.cpp:42-101). */
int b2g_arena_create(const b2gArenaDef* def, b2gArena** out);
int b2g_arena_destroy(b2gArena* arena);

/* b2World::CreateBody + b2Body ctor (b2_world.cpp:140-176, b2_body.cpp:29-120) and every
 * b2Body setter: the host writes [first, first+count) of the body SoA. */
int b2g_upload_bodies(b2gArena* arena, int32_t first, int32_t count, const b2gBodyArrays* src);
/* b2Body::CreateFixture + b2Fixture::Create (b2_body.cpp:230-269, b2_fixture.cpp:44-75). */
int b2g_upload_fixtures(b2gArena* arena, int32_t first, int32_t count, const b2gFixtureArrays* src);
int b2g_upload_shapes(b2gArena* arena, int32_t first_quad, int32_t count_quads, const float* quads);
/* Sparse variants: row i of the given arrays goes to index[i] (indices distinct).  One host-to-device copy per
 * array and one scatter kernel whatever the spread of the indices — what a b2WorldBatch needs when each of a
 * thousand worlds edited one body (b2Body setters, CreateBody / CreateFixture while running).  Every array
 * pointer must be given (there is no "leave this column alone"). */
int b2g_upload_bodies_indexed(b2gArena* arena, int32_t count, const int32_t* index, const b2gBodyArrays* rows);
int b2g_upload_fixtures_indexed(b2gArena* arena, int32_t count, const int32_t* index, const b2gFixtureArrays* rows);
int b2g_upload_shapes_indexed(b2gArena* arena, int32_t count_quads, const int32_t* index, const float* quads);
/* b2World::CreateJoint for revolute joints (b2_world.cpp:268-323). */
int b2g_upload_joints(b2gArena* arena, int32_t first, int32_t count, const b2gJointArrays* src);
/* The joints' accumulated impulses after the last step, state[count][5] as above
 * (b2RevoluteJoint::GetReactionForce / GetReactionTorque / GetMotorTorque read them,
 * b2_revolute_joint.cpp:333-342, 374-377). */
int b2g_download_joints(b2gArena* arena, int32_t first, int32_t count, float* state);
/* Truncate counts (b2World::DestroyBody of trailing bodies; full compaction is host-side). */
int b2g_set_counts(b2gArena* arena, int32_t num_bodies, int32_t num_fixtures, int32_t num_joints);

/* Per-step user input: only the force accumulators (b2Body::ApplyForce*, b2_body.h:742-800),
 * force [n][4] as above except that element 3 (sleepTime) is ignored. */
int b2g_upload_forces(b2gArena* arena, int32_t first, int32_t count, const float* force);

/* b2World::Step (src/dynamics/b2_world.cpp:1108-1171): Collide -> Solve (islands, contact
 * solver, integration, sleep) -> FindNewContacts -> ClearForces, all on the device. */
int b2g_step(b2gArena* arena, const b2gStepParams* params, b2gStepStats* stats);
/* B2G_ERR_CAPACITY from a step means max_contacts was too small for the pairs found at the END of
 * the step: bodies, joints and the surviving contacts are final and consistent, only the new
 * pairs were not inserted.  Nothing is dropped silently: download the state, create a larger
 * arena, upload (b2g_upload_contacts keeps manifolds and impulses) and carry on — the next step's
 * pair refresh creates the missing contacts exactly when the reference would first use them.
 * host/b2_world_host.cpp does this automatically. */

/* b2g_step followed by b2g_download_body_state_async + b2g_synchronize for bodies [first,
 * first+count), except that the readback ([count][8] = xf(4), vel(4), pinned dst recommended)
 * starts as soon as Solve has written the final body state and overlaps the pair refresh at the
 * end of the step.  When the call returns, dst is complete. */
int b2g_step_download(b2gArena* arena, const b2gStepParams* params, b2gStepStats* stats, int32_t first,
                      int32_t count, float* dst);

/* The same step in two halves, for hosts that must run b2ContactListener::PreSolve between
 * the narrowphase and the solver (b2_contact.cpp:197-209 fires inside Collide):
 *   b2g_step_collide = b2ContactManager::Collide (b2_world.cpp:1134-1138)
 *   b2g_step_solve   = Solve + FindNewContacts + ClearForces (b2_world.cpp:1140-1167)
 * b2g_step(a,p,s) == b2g_step_collide(a,p) then b2g_step_solve(a,p,s). */
int b2g_step_collide(b2gArena* arena, const b2gStepParams* params);
int b2g_step_solve(b2gArena* arena, const b2gStepParams* params, b2gStepStats* stats);

/* The pair refresh Step runs first when fixtures were added or moved
 * (m_newContacts, b2_world.cpp:1114-1122): UpdateAndQuery + RemoveDeadContacts only. */
int b2g_find_new_contacts(b2gArena* arena);

/* Readback for b2Body getters (b2_body.h:187-217). dst arrays sized [count]. */
int b2g_download_bodies(b2gArena* arena, int32_t first, int32_t count, const b2gBodyArrays* dst);
/* Pinned-host fast path used by the e2e benchmark: xf+vel of [first,first+count) packed as
 * [count][8] = xf(4), vel(4).  dst must stay valid until b2g_synchronize(). */
int b2g_download_body_state_async(b2gArena* arena, int32_t first, int32_t count, float* dst);
int b2g_download_fixture_aabbs(b2gArena* arena, int32_t first, int32_t count, float* aabb /*[n][4]*/);
/* b2World::GetContactListStart/GetContactCount (b2_world.h:187-220). */
int b2g_contact_count(b2gArena* arena, int32_t* out);
int b2g_download_contacts(b2gArena* arena, int32_t first, int32_t count, const b2gContactArrays* dst);
/* b2Contact::SetEnabled/SetFriction/SetRestitution/SetTangentSpeed from PreSolve
 * (b2_contact.h:80-140): overwrite flags/material of contacts [first, first+count). */
int b2g_upload_contact_overrides(b2gArena* arena, int32_t first, int32_t count,
                                 const uint32_t* flags, const float* material);

/* Replaces the whole contact set with `count` contacts given in the b2gContactArrays layout
 * (fixture_a/fixture_b as ordered A,B; flags = B2G_CONTACT_TOUCHING | B2G_CONTACT_ENABLED;
 * manifold incl. warm-start impulses; material).  Lets a host mirror an existing b2World — its
 * contact list (b2_contact_manager.h:60-76) — into the arena mid-simulation; the parity tests use
 * it to start every step from the reference's exact state. */
int b2g_upload_contacts(b2gArena* arena, int32_t count, const b2gContactArrays* src);

/* B2G_SOLVER_SEQUENTIAL visits constraints in ascending pair-key order by default.  This call
 * imposes an explicit order for the NEXT step only: the listed ordered pairs first, in the given
 * order (e.g. the reference's island order as reported by b2ContactListener::PostSolve,
 * b2_island.cpp:621-647), every other constraint after them by pair key. */
int b2g_set_sequential_order(b2gArena* arena, int32_t count, const int32_t* fixture_a, const int32_t* fixture_b);
/* The same for joints: the order in which B2G_SOLVER_SEQUENTIAL visits the joints in the NEXT step
 * only (the reference's island DFS order, b2_world.cpp:622-647); joints not listed are skipped.
 * Without it joints are visited in descending index order. */
int b2g_set_sequential_joint_order(b2gArena* arena, int32_t count, const int32_t* joints);

/* b2ContactListener::BeginContact / EndContact (b2_contact.cpp:197-204,
 * b2_contact_manager.cpp:48-51) recorded by the last step when record_events was set:
 * each event = fixtureA, fixtureB.  Returns the number available through *count. */
int b2g_download_events(b2gArena* arena, int32_t* begin_pairs, int32_t* begin_count, int32_t* end_pairs,
                        int32_t* end_count, int32_t capacity_each);

/* Raw device addresses of the body state arrays ([capacity][4] float32 each, flags [capacity]
 * uint32), for zero-copy interop on the same device: the slab decomposition of one large world
 *  packs boundary bodies straight out of these arrays into its NCCL send buffers and
 * scatters the received ghost states straight back in.  The caller must order its own accesses
 * against the arena's stream (b2g_synchronize / b2g_stream). */
typedef struct b2gDeviceViews {
  void* pos;
  void* vel;
  void* xf;
  void* force;
  void* flags;
  int32_t capacity;
  int32_t device;
} b2gDeviceViews;
int b2g_device_views(b2gArena* arena, b2gDeviceViews* out);

/* ---- Halo exchange of a spatially decomposed world (BASELINE config 5: one very large world cut
 * into x-slabs, one arena per GPU).  Each arena holds the bodies its rank owns plus GHOST copies of the
 * neighbours' bodies near the cut; after every step the owners' states overwrite the ghosts.  `slot` 0 is the
 * lower-x neighbour, 1 the upper-x one.  A message is 64 bytes per body (pos, vel, xf, flags) in the order of
 * the lists; pack / unpack are kernels on the arena's stream, so with a stream-ordered transport (NCCL on
 * b2g_stream(), see b2g_dist_exchange) an exchange needs no host synchronisation.
 * Replaces nothing in the reference (it has no multi-device path); the body fields exchanged are the ones
 * b2Island::Solve writes back (src/dynamics/b2_island.cpp:430-438). */
int b2g_halo_set_lists(b2gArena* arena, int32_t slot, int32_t n_send, const int32_t* send_bodies, int32_t n_recv,
                       const int32_t* recv_bodies);
/* gathers the send list into the arena's send buffer; returns that device buffer */
int b2g_halo_pack(b2gArena* arena, int32_t slot, void** message, int64_t* bytes);
/* device buffer the transport should receive the neighbour's message into */
int b2g_halo_recv_buffer(b2gArena* arena, int32_t slot, void** message, int64_t* bytes);
/* scatters a received message (NULL = the arena's own receive buffer) into the ghost bodies */
int b2g_halo_unpack(b2gArena* arena, int32_t slot, const void* message);

/* ---- libb2cuda_dist.so: the NCCL transport of the halo exchange (one process per GPU; links NCCL, so it is a
 * separate library).  b2g_dist_unique_id on rank 0 -> broadcast the 128 bytes by any means -> b2g_dist_init on
 * every rank.  b2g_dist_exchange = pack, one ncclSend + ncclRecv per neighbour in one group, unpack, all
 * enqueued on the arena's stream.  Rank r's neighbours are r - 1 (slot 0) and r + 1 (slot 1). */
typedef struct b2gDist b2gDist;
int b2g_dist_unique_id(void* out128);
int b2g_dist_init(const void* unique_id128, int32_t rank, int32_t nranks, int32_t device, b2gDist** out);
int b2g_dist_exchange(b2gDist* dist, b2gArena* arena);
int b2g_dist_destroy(b2gDist* dist);

/* Tests / diagnostics: how an oversize island is currently cut into per-SM tiles (csrc/b2g_tiles.cuh): the plan
 * (11 words: bounds lo[2] hi[2] as order-preserving uints, body count, strips, rows, x0, 1/dx, y0, 1/dy), the
 * bodies per tile of the last step (160 ints) and every body's tile slot (tile * 2048 + slot, -1 = none). */
int b2g_debug_tile_state(b2gArena* arena, uint32_t* plan, int32_t* tile_count, int32_t* tile_slot);
/* Tests / diagnostics: colour[num_joints] of the production mode's joint colouring (csrc/b2g_fused.cuh,
 * k_joint_colour): joints of one colour share no movable body and are solved side by side where the reference
 * walks the island's joint list (b2_island.cpp:323-338, 392-401); 60 = a hub's joints beyond the colours,
 * walked by one thread. */
int b2g_debug_joint_colours(b2gArena* arena, int32_t* colour);

int b2g_synchronize(b2gArena* arena);
/* cudaStream_t the arena launches on (for external CUDA-event timing). */
void* b2g_stream(b2gArena* arena);
int b2g_set_profiling(b2gArena* arena, int32_t on);
/* Per-kernel-class CUDA-event timing on the arena's stream, for the roofline line of bench.py
 * (the reference's only profile is b2Profile's four wall-clock phases, b2_time_step.h:29-36).
 * While on, every launch is bracketed by two events; totals accumulate until switched on again.
 * `units` returns the summed work items (bodies / fixtures / contacts / constraints). */
int b2g_set_kernel_timing(b2gArena* arena, int32_t on);
int b2g_kernel_class_count(void);
const char* b2g_kernel_class_name(int32_t cls);
int b2g_get_kernel_timing(b2gArena* arena, int32_t cls, double* total_ms, int64_t* launches, double* units);
/* previous step's 1/dt (m_inv_dt0, b2_world.cpp:69,1162) — exposed for tests. */
int b2g_set_inv_dt0(b2gArena* arena, float inv_dt0);

/* Pinned host memory helpers for the e2e path. */
int b2g_host_alloc(void** out, uint64_t bytes);
int b2g_host_free(void* p);

/* ------------------------------------------------------------------------------------------
 * User contact filter, b2ContactFilter::ShouldCollide (b2_world_callbacks.h:60-69; consulted by
 * b2ContactManager::QueryCallback, b2_contact_manager.cpp:163-170, whenever an overlapping pair has
 * no contact).  The default category / mask / group rule and the joints' collideConnected run on
 * the device; a user callback cannot, so the host mediates:
 *   b2g_download_new_pairs   the pairs the last pair refresh inserted as contacts
 *   b2g_set_pair_vetoes      the complete list of rejected pairs; their contacts are removed at
 *                            once and the pair finder skips them from now on
 *   b2g_download_veto_seen   seen[i] = 1 when vetoed pair i (sorted by lower fixture, then upper)
 *                            still overlapped in the last refresh — ask the filter again, as the
 *                            reference does every step; pairs not seen have separated: drop them
 * ---------------------------------------------------------------------------------------- */
int b2g_download_new_pairs(b2gArena* arena, int32_t capacity, int32_t* fixture_a, int32_t* fixture_b, int32_t* count);
int b2g_set_pair_vetoes(b2gArena* arena, int32_t count, const int32_t* fixture_a, const int32_t* fixture_b);
int b2g_download_veto_seen(b2gArena* arena, int32_t count, uint8_t* seen);

/* ------------------------------------------------------------------------------------------
 * Spatial queries on the broadphase tree, batched.  All pointers are host
 * pointers, or device pointers on the arena's device when on_device != 0 (results then stay in
 * HBM; the call still returns after the kernel has finished).  world[i] >= 0 restricts query i to
 * one world of a multi-world arena; NULL or -1 = every world.  The tree reflects the last step or
 * pair refresh; pending uploads are applied first (b2BroadPhase::EnsureBuiltTree).
 * ---------------------------------------------------------------------------------------- */

/* b2World::QueryAABB (b2_world.cpp:1193-1207; b2BroadPhase::Query, b2_broad_phase.h:622-643):
 * aabbs[n][4] = lower.xy, upper.xy.  counts[n] = fixtures whose AABB overlaps (inclusive test,
 * b2TestOverlap); fixtures[n][cap] = the first cap of them per query, in no particular order. */
int b2g_query_aabb(b2gArena* arena, int32_t n, const float* aabbs, const int32_t* world, int32_t cap,
                   int32_t* counts, int32_t* fixtures, int32_t on_device);

/* b2World::RayCast (b2_world.cpp:1226-1246; b2BroadPhase::RayCast, b2_broad_phase.h:645-716;
 * shape tests b2_circle_shape.cpp:56-89, b2_edge_shape.cpp:89-154, b2_polygon_shape.cpp:303-371)
 * with the callback every "closest hit" user writes (return fraction): rays[n][4] = p1.xy, p2.xy;
 * max_fraction[n] or NULL (= 1).  fixture[n] = hit fixture or -1, fraction[n], normal[n][2].
 * Only fixtures with (categoryBits & category_mask) != 0 are considered.  Equal fractions: the
 * lower fixture index wins (the reference keeps whichever its tree visits last). */
int b2g_ray_cast_closest(b2gArena* arena, int32_t n, const float* rays, const float* max_fraction,
                         const int32_t* world, uint32_t category_mask, int32_t* fixture, float* fraction,
                         float* normal, int32_t on_device);

/* The same cast with the callback that returns 1 (report everything): counts[n] hits per ray and
 * the first cap of them in fixture/fraction/normal[n][cap](,[2]), unordered.  The host-side
 * b2World::RayCast replays these, sorted by fraction, into the user's b2RayCastCallback. */
int b2g_ray_cast_all(b2gArena* arena, int32_t n, const float* rays, const float* max_fraction,
                     const int32_t* world, uint32_t category_mask, int32_t cap, int32_t* counts, int32_t* fixture,
                     float* fraction, float* normal, int32_t on_device);

/* ------------------------------------------------------------------------------------------
 * Kernel-level entry points: each runs ONE production device function over explicit host
 * arrays.  They exist so the parity tests can feed the GPU the oracle's exact ordered inputs
 * (SURVEY.md §7 "Hard parts": A/B order and solver order are traversal dependent).
 * ---------------------------------------------------------------------------------------- */

/* b2Rot::Set (b2_math.h:313-318): angle[n] -> sin_cos[n][2].  The device evaluates the host
 * libm's sinf/cosf algorithm so that body transforms match the reference bit for bit. */
int b2g_rotations(int32_t device, int32_t n, const float* angle, float* sin_cos);

/* b2Fixture::UpdateAABB -> b2{Polygon,Circle,Edge}Shape::ComputeAABB
 * (b2_fixture.cpp:138-140, b2_polygon_shape.cpp:373-388, b2_circle_shape.cpp:91-96,
 * b2_edge_shape.cpp:156-167).  type[n], shape_off[n], xf[n][4] -> aabb[n][4]. */
int b2g_compute_aabbs(int32_t device, int32_t n, const int32_t* type, const int32_t* shape_off,
                      const float* shape_quads, int32_t num_quads, const float* xf, float* aabb);

/* b2Contact::Update's evaluate step = the five b2Collide* functions
 * (b2_collide_circle.cpp:27-158, b2_collide_polygon.cpp:120-243, b2_collide_edge.cpp:31-524)
 * on n ordered pairs (A first).  manifold[n][16] as in b2gContactArrays (impulses 0). */
int b2g_collide_pairs(int32_t device, int32_t n, const int32_t* type_a, const int32_t* shape_off_a,
                      const float* xf_a, const int32_t* type_b, const int32_t* shape_off_b,
                      const float* xf_b, const float* shape_quads, int32_t num_quads, float* manifold);

/* b2BroadPhase::UpdateAndQuery's pair set (include/box2d/b2_broad_phase.h:252-511) for n
 * leaves: aabb[n][4], body[n], world[n] (may be NULL), dyn[n] (1 = owning body is dynamic).
 * Reports every inclusive-overlap pair of different bodies with >= 1 dynamic body, as sorted
 * (lo, hi) leaf indices.  *num_pairs returns the full count even if it exceeds capacity. */
int b2g_find_pairs(int32_t device, int32_t n, const float* aabb, const int32_t* body, const int32_t* world,
                   const uint8_t* dyn, int32_t* pairs, int32_t capacity, int32_t* num_pairs);

/* b2ContactSolver (src/dynamics/b2_contact_solver.cpp:65-787) driven exactly as
 * b2Island::Solve drives it (b2_island.cpp:306-409) on nb bodies and nc ordered constraints,
 * one thread, constraint order as given:
 *   pos,vel [nb][4] in/out; mass [nb][4] = invMass, invI, localCenter.xy
 *   index [nc][2]; manifold [nc][16] in/out (impulses); material [nc][4]; radii [nc][2]
 *   vel_iterates [vel_iters][nb][4], pos_iterates [pos_iters][nb][4] (may be NULL)
 * The caller has already integrated velocities (b2_island.cpp:257-293); this entry runs
 * Initialize, InitializeVelocityConstraints, WarmStart, vel_iters x SolveVelocityConstraints,
 * StoreImpulses, position integration (b2_island.cpp:353-385), <= pos_iters x
 * SolvePositionConstraints with the early exit.  *pos_iters_done returns iterations run. */
int b2g_solve_sequential(int32_t device, int32_t nb, float* pos, float* vel, const float* mass, int32_t nc,
                         const int32_t* index, float* manifold, const float* material, const float* radii,
                         float dt, float dt_ratio, int32_t warm_starting, int32_t vel_iters,
                         int32_t pos_iters, float* vel_iterates, float* pos_iterates,
                         int32_t* pos_iters_done);

#ifdef __cplusplus
}
#endif
#endif /* B2CUDA_H */

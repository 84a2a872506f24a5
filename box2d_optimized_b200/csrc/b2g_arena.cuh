// b2g_arena.cuh — the device-resident world state (structure-of-arrays in HBM).
//
// Replaces the pointer-linked AoS state of the reference: b2Body (include/box2d/b2_body.h:498-547),
// b2Fixture (include/box2d/b2_fixture.h:243-268), b2Contact (include/box2d/b2_contact.h:176-230),
// the broadphase node pool (include/box2d/b2_broad_phase.h:32-42) and the per-island solver
// scratch (src/dynamics/b2_island.h:57-95).  SURVEY.md Appendix A lists the field mapping.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/b2cuda.h"
#include "b2g_solver.cuh"
#include "b2g_joint.cuh"

#define B2G_MAX_COLOURS 24          // colours solved by parallel launches
#define B2G_OVERFLOW_COLOUR B2G_MAX_COLOURS  // serial overflow bucket
#define B2G_MAX_POS_ITERS 16
#define B2G_ISLAND_EXACT_PERIOD 16  // steps between exact recomputations of oversize islands' labels
#define B2G_BVH_REBUILD_PERIOD 64  // longest run of refit-only steps; the leaf order is re-sorted earlier when walks get longer
#define B2G_KT_MAX 2048  // timed launches per step when per-kernel timing is on

// device-side counters, zeroed at the start of every step; mirrored to pinned host memory
struct StepCounts {
  int numPairs;      // candidate pairs emitted by the broadphase traversal
  int numActive;     // solver constraints this step
  int remaining;     // constraints still uncoloured after the rounds launched so far
  int numTouching;
  int numAwake;
  int beginCount, endCount;
  int lastUsefulRound;  // 1 + index of the last colouring round that coloured something
  int numBig;           // constraints of islands too large for a fused tile
  int numColours, numOverflow, maxIslandBodies;
  int slotCursor;       // next free slot of the island-sorted body order
  int numDead;          // contacts retired by this step's sweep
  int freeTopRead;      // host copy of the free-stack height (filled by the readback)
  int jointOverflow;    // joints that did not fit a tile's joint list (reported as an error)
  unsigned long long bpVisits;  // internal nodes visited by this step's pair traversal (tree quality)
  int bpMaxVisits;              // longest single walk
  int numBigBodies;             // bodies of oversize islands (an island can be oversize through joints alone)
  int colourCount[B2G_MAX_COLOURS + 1];
  int spillCount;               // bodies of a tiled oversize island that did not fit their tile
  int maxTileCount;             // bodies in the fullest tile
  int bigJoints;                // joints between bodies of oversize islands
  int worklistCount;            // uncoloured active constraints of this step (k_mark_active_bins -> k_colour_worklist)
  int worklistLeft[200];        // per colouring round: somebody is still uncoloured (grid mode)
  unsigned int boundsLo[2], boundsHi[2];  // ordered-int encoded min/max of 2*centre
};

// (fixLo << 32 | fixHi) -> contact slot, open addressing; see b2g_broadphase.cuh
struct ContactHash {
  unsigned long long* keys;
  int* vals;
  unsigned int mask;  // capacity - 1 (capacity is a power of two)
  // b2Body::ShouldCollide (b2_body.cpp:396-419): body pairs joined by a joint with
  // collideConnected == false, as sorted keys bodyLo << 32 | bodyHi (binary search; usually empty)
  const unsigned long long* ncKeys;
  int ncCount;
  // pairs a user b2ContactFilter::ShouldCollide rejected (b2_contact_manager.cpp:163-170), sorted
  // fixLo << 32 | fixHi; a rejected pair that is still overlapping marks vetoSeen so that the host
  // can ask the filter again, as the reference does on every step the pair is found without a contact
  const unsigned long long* vetoKeys;
  uint8_t* vetoSeen;
  int vetoCount;
};

// contacts live in stable slots; a slot is live when its flags carry B2G_CONTACT_ALIVE
struct ContactBuf {
  unsigned long long* key;  // fixLo << 32 | fixHi (the contact's identity)
  int2* fix;                // fixtureA, fixtureB (A/B by the type table, b2_contact.cpp:58-77)
  int2* body;               // bodyA, bodyB
  uint32_t* flags;          // B2G_CONTACT_*
  float4* material;         // friction, restitution, restitutionThreshold, tangentSpeed
  float4* m0;               // manifold planes, see b2gContactArrays.manifold
  float4* m1;
  float4* m2;
  float4* m3;
  int* colour;              // persistent solver colour, -1 = none
};

struct b2gArena {
  int device;
  cudaStream_t stream;
  cudaStream_t copyStream;      // body-state readback of b2g_step_download, overlapped with the pair refresh
  cudaEvent_t evPacked;
  float* pendingStateDst;       // host destination of the step in flight (b2g_step_download)
  int pendingStateFirst, pendingStateCount;
  cudaEvent_t ev[5];
  int profiling;
  int numWorlds, worldBits;
  int capBodies, capFixtures, capQuads, capContacts, capJoints;
  int nBodies, nFixtures, nJoints, nContacts;
  int fixBits;        // bits per fixture index in the pair key
  int aabbAllDirty;   // recompute static AABBs too (after fixture / body upload)
  int newFixtures;    // b2World::m_newContacts: fixtures were added, the next step starts with FindNewContacts
  int bvhLeaves, bvhAge;  // leaves of the current leaf order, steps since it was sorted
  float bvhVisitsFresh, bvhVisitsLast;  // mean nodes visited per pair-finder walk: right after the last sort / last step
  int recolour;       // drop every persistent colour (new arena, contacts uploaded by the host)
  uint8_t* worldRecolour;  // [numWorlds] 1 = a body of this world had its mass / type edited: the world is coloured afresh
  int recolourWorlds;      // any byte of worldRecolour set
  int roundsHint;     // colouring rounds to launch before the first check
  float invDt0;
  long long launches;
  long long launchesAtStepStart;
  // per-kernel-class CUDA-event timing (off by default: it serialises launches)
  int kernelTiming, ktCount;
  cudaEvent_t ktEv[2 * B2G_KT_MAX];
  int ktClass[B2G_KT_MAX];
  double ktUnits[B2G_KT_MAX];
  double ktMs[20], ktUnitsSum[20];
  long long ktLaunches[20];

  // bodies
  float4 *pos, *vel, *xf, *mass, *center, *force;
  uint32_t* bflags;
  int* bworld;
  int* worldFixMin;        // [numWorlds] smallest fixture index of each world
  int* bodyFixBase;        // per body: worldFixMin of its world (colour priorities use world-local pair keys)
  int fixBaseDirty;
  int* islandParent;       // union-find forest (lock-free unions)
  int* island;             // flattened: island id = smallest body index of the component
  uint32_t* islandAwake;   // per root: some member is awake
  uint32_t* islandMinSleep;  // per root: float bits of min sleepTime
  uint32_t* islandPen;     // [B2G_MAX_POS_ITERS][capBodies] float bits of max penetration per iteration
  unsigned long long* colourMask;  // per body: colours used by its constraints
  unsigned long long* bodyBest;    // per body: best proposal this round
  // fused solver: island-sorted body slots and bins
  int *islandCount, *islandStart, *islandCursor, *bodySlot, *slotBody, *binFirst, *binEnd, *cbin, *bucketCount, *bucketStart;
  unsigned int *conKeys, *conKeysSorted;
  int *conVals;
  int nbinsMax, bigMode, lastMaxIsland, lastNumBig;
  float stepDt;  // dt of the step in flight (soft joint constraints)
  int* jointOrder;   // [capJoints] explicit joint order of the sequential mode (one step)
  int jointOrderCount, jointOrderActive;
  int lastOverflow, lastActive;  // serial-bucket constraints / solver rows of the previous step
  int islandsValid;   // island[] of the previous step may seed this step's union-find
  uint8_t* islandDirty;  // per island root: an edge was removed since the labels were computed
  uint8_t* islandWasBig; // per island root: last step it was too large for a tile
  long long stepCount;
  size_t fusedSmemSet;
  int bigGrid;  // co-resident grid of the persistent big-island kernel
  unsigned int* bigBarrier;  // its grid-barrier counter (zeroed before every launch)
  int colourGrid;            // co-resident grid of k_colour_worklist
  unsigned int* colourBarrier;
  // oversize islands cut into per-SM tiles (b2g_tiles.cuh)
  struct TilePlan* tilePlan;
  int *tileStripOfX, *tileRowOfY, *tileHistX, *tileHistY;
  int *tileSlot, *tileBodies, *tileCount, *spillList;
  uint8_t* tileBoundary;
  int2* tileCutSeq;          // per solver slot of a cut constraint: its turn on either body (b2g_tiles.cuh)
  int tileNoSequencing;      // B2G_TILE_BARRIERS=1: grid barriers between the cut colours (measurements)
  unsigned int* tileBarrier;
  int tileGrid;              // co-resident grid of k_big_tiles (= SM count)
  int tilePlanValid, tilePlanAge;
  int lastBigBodies, lastSpill, lastMaxTile;
  int tilePlanBodies;        // oversize bodies when the plan was made
  int tilesDisabled;         // B2G_NO_TILES=1: keep the grid-pass path (measurements)

  // fixtures + shapes
  int* fBody;
  int* fShapeOff;
  uint32_t* fTypeFlags;
  uint2* fFilter;
  float4* fMaterial;
  float4* fAabb;
  float* fRadius;
  float4* shapes;

  // joints (revolute)
  int2* jBodies;
  float4* jAnchors;
  float4* jParams0;  // referenceAngle, lower, upper, maxMotorTorque
  float4* jParams1;  // motorSpeed, bits(flags), 0, 0
  float4* jParams2;  // wheel: damping, 0, 0, 0
  float4* jState;    // impulse.x, impulse.y, motorImpulse, lowerImpulse
  float* jUpper;     // upperImpulse
  unsigned long long* ncKeys;     // [capJoints] unsorted, then sorted into ncKeysSorted
  unsigned long long* ncKeysSorted;
  uint8_t* bodyNoCollide;         // [capBodies] body has at least one collideConnected == false joint
  int jointFilterDirty;
  float4* forceStage; // [capBodies] host forces land here in one linear copy, then merge into force.xyz
  float4* stateStage; // [capBodies][2] packed xf+vel for b2g_download_body_state_async (one linear D2H)
  JointWork* jWork;  // per-step scratch
  // joint colouring of the production mode (k_joint_colour), redone when the joint table or a body's mass changes
  int* jColour;                     // [capJoints]
  int* jSorted;                     // [capJoints] joints by colour
  int* jCstart;                     // [B2G_JOINT_COLOURS + 2]
  unsigned long long* jBodyMask;    // [capBodies] scratch of the colouring
  unsigned long long* jBodyBest;    // [capBodies]
  int jointColourDirty;

  // contacts
  ContactBuf cb[1];           // stable slots; nContacts = slot high-water mark, nAlive = live contacts
  unsigned long long* orderKey;  // per slot: visiting order imposed for the next sequential step
  int seqOrderActive;
  unsigned long long* seqKeys;  // [2][capContacts] scratch: key sort of the sequential mode's list
  uint8_t* persist;           // per slot: pair re-reported by this step's broadphase
  int* freeStack;             // free slots (LIFO)
  int* dFreeTop;              // device: entries in freeStack
  int nAlive, tombstones;
  ContactHash hash;
  int* downloadSlots;  // host: slot of each contact in the order of the last download
  int downloadCount;

  // broadphase scratch
  unsigned long long *mortonKeys, *mortonKeysSorted;
  int *leafFixture, *leafFixtureSorted;
  float4* leafBox;
  int4* leafInfo;
  unsigned long long* leafKey;  // (AABB size bits, sorted position): who reports a pair
  int* worldFirst;
  int* worldLast;
  float4* bvhBox;               // implicit 8-wide tree: boxes of the internal levels (b2g_broadphase.cuh)
  unsigned long long* bvhKey;   // and the largest leaf key below each node
  int* bvhDone;                 // last-block-done counter of the refit
  unsigned long long *pairKeys;  // new pairs (no live contact yet) reported by the traversal
  int lastNewPairs;              // how many of them the last pair refresh inserted
  unsigned long long* vetoKeys;  // see ContactHash
  uint8_t* vetoSeen;
  int vetoCap;

  // solver scratch
  uint8_t* activeFlag;
  int* activeList;
  int* sortedList;
  uint8_t *colourKey, *colourKeySorted;
  int* croot;
  SolverPlanes planes;

  // halo exchange of a spatially decomposed world (slot 0 = lower neighbour, 1 = upper neighbour)
  int* haloSend[2];    // bodies whose state the neighbour holds as ghosts
  int* haloRecv[2];    // ghost bodies refreshed from the neighbour
  int haloNumSend[2], haloNumRecv[2];
  float4* haloOut[2];  // packed messages: 4 quads per body (pos, vel, xf, flags)
  float4* haloIn[2];

  // events
  int2 *beginEvents, *endEvents;

  StepCounts* dCounts;
  StepCounts* hCounts;  // pinned
  void* cubTemp;
  size_t cubTempBytes;
  float* hostStage;  // pinned staging for small downloads
  unsigned char* scatterStage;  // device staging of the indexed uploads (grown on demand)
  size_t scatterStageBytes;
};

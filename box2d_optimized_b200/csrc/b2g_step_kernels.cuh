// b2g_step_kernels.cuh — the kernels of one b2World::Step on the device.
//
// Order and meaning follow src/dynamics/b2_world.cpp:1108-1171 (Step) and :492-670 (Solve):
//   k_narrowphase            b2ContactManager::Collide + b2Contact::Update
//   k_body_begin .. k_island_flatten   island discovery (DFS of b2World::Solve :522-659)
//   k_integrate_velocities   b2Island::Solve :257-293 / SolveOrphan :187-213
//   k_mark_active, k_colour_*          constraint list + greedy graph colouring (new)
//   k_prepare .. k_store_impulses      b2ContactSolver
//   k_integrate_positions    b2Island::Solve :353-385
//   k_solve_position         b2ContactSolver::SolvePositionConstraints
//   k_finalize_bodies, k_sleep         write-back, SynchronizeTransform, island sleep :430-483
//   broadphase kernels live in b2g_broadphase.cuh
#pragma once
#include "b2g_arena.cuh"
#include "b2g_gjk.cuh"
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

#define B2G_BODY_TYPE(f) (((f) >> B2G_BODY_TYPE_SHIFT) & 3u)
#define B2G_STATIC 0u
#define B2G_KINEMATIC 1u
#define B2G_DYNAMIC 2u

__device__ __forceinline__ unsigned int float_flip(float f) {
  // order-preserving float -> uint map (for atomicMin/atomicMax on floats of any sign)
  unsigned int u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float float_unflip(unsigned int u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}
__device__ __forceinline__ unsigned int hash32(unsigned int x) {
  x ^= x >> 16;
  x *= 0x7feb352du;
  x ^= x >> 15;
  x *= 0x846ca68bu;
  x ^= x >> 16;
  return x;
}

// ---------------------------------------------------------------------------------------------
// Narrowphase: one thread per contact slot.  b2ContactManager::Collide (b2_contact_manager.cpp:66-116) +
// b2Contact::Update (b2_contact.cpp:126-210).
// Slots are stable (free list), so a block's 128 consecutive slots hold a mix of shape-type pairs, and a warp
// that runs b2CollideCircles, b2CollidePolygonAndCircle and b2CollidePolygons one after the other pays for all
// three.  The block therefore regroups its slots by type pair first (counting sort in shared memory): thread t
// takes the t-th slot in (type pair) order, so all but a few warps of a block are type-uniform, while the
// block still reads and writes one contiguous window of the contact arrays.  Which thread computes a slot
// does not touch the result.
// ---------------------------------------------------------------------------------------------
#define B2G_NP_THREADS 128
#define B2G_NP_CLASSES 18  // 16 type pairs, sensors, nothing to do
#ifndef B2G_NP_MIN_BLOCKS
#define B2G_NP_MIN_BLOCKS 8  // 64 registers: 32 resident warps hide the shape / transform gathers (82 -> 77 us on mixed_100k)
#endif
__global__ void __launch_bounds__(B2G_NP_THREADS, B2G_NP_MIN_BLOCKS)
k_narrowphase(int nc, ContactBuf C, const uint32_t* bflags, const float4* __restrict__ xf,
              const int* __restrict__ fShapeOff, const uint32_t* __restrict__ fTypeFlags,
              const float4* __restrict__ shapes, uint32_t* bflagsRW, StepCounts* counts, int recordEvents,
              int2* beginEvents, int2* endEvents, int eventCap, const int* __restrict__ islandPrev,
              uint8_t* islandDirty) {
  B2G_PDL_ENTER();
  __shared__ int sCount[B2G_NP_CLASSES], sStart[B2G_NP_CLASSES];
  __shared__ unsigned char sOrder[B2G_NP_THREADS];
  const int tid = threadIdx.x, base = blockIdx.x * blockDim.x;
  if (tid < B2G_NP_CLASSES) sCount[tid] = 0;
  __syncthreads();
  int cls = B2G_NP_CLASSES - 1;
  {
    const int i0 = base + tid;
    if (i0 < nc) {
      const uint32_t fl = C.flags[i0];
      if (fl & B2G_CONTACT_ALIVE) {
        const int2 bd0 = C.body[i0];
        const uint32_t fa0 = bflags[bd0.x], fb0 = bflags[bd0.y];
        const bool act = ((fa0 & B2G_BODY_AWAKE) && B2G_BODY_TYPE(fa0) != B2G_STATIC) ||
                         ((fb0 & B2G_BODY_AWAKE) && B2G_BODY_TYPE(fb0) != B2G_STATIC);
        if (act) {
          const int2 fx0 = C.fix[i0];
          const uint32_t ta = fTypeFlags[fx0.x], tb = fTypeFlags[fx0.y];
          cls = ((ta | tb) & B2G_FIX_SENSOR) ? 16 : (int)(((ta & 3u) << 2) | (tb & 3u));
        } else if (fl & B2G_CONTACT_TOUCHING) {
          auto g = cg::coalesced_threads();
          if (g.thread_rank() == 0) atomicAdd(&counts->numTouching, (int)g.size());
        }
      }
    }
  }
  const int rank = atomicAdd(&sCount[cls], 1);
  __syncthreads();
  if (tid == 0) {
    int run = 0;
    for (int c = 0; c < B2G_NP_CLASSES; ++c) {
      sStart[c] = run;
      run += sCount[c];
    }
  }
  __syncthreads();
  sOrder[sStart[cls] + rank] = (unsigned char)tid;
  __syncthreads();
  if (tid >= sStart[B2G_NP_CLASSES - 1]) return;  // free slots, sleeping pairs, the tail of the array
  const int i = base + sOrder[tid];
  uint32_t flags = C.flags[i];
  int2 bd = C.body[i];
  uint32_t fa = bflags[bd.x], fb = bflags[bd.y];
  int2 fx = C.fix[i];
  uint32_t tfA = fTypeFlags[fx.x], tfB = fTypeFlags[fx.y];
  bool sensor = ((tfA | tfB) & B2G_FIX_SENSOR) != 0;
  bool wasTouching = (flags & B2G_CONTACT_TOUCHING) != 0;
  flags |= B2G_CONTACT_ENABLED;

  Xf xfA = xf_from4(xf[bd.x]), xfB = xf_from4(xf[bd.y]);
  Manifold m;
  bool touching;
  if (sensor) {
    // sensors report overlap but carry no manifold (b2_contact.cpp:145-151): b2TestOverlap = GJK
    // distance with radii (b2g_gjk.cuh)
    manifold_clear(m);
    touching = gjk_test_overlap(shapes, (int)(tfA & 3u), fShapeOff[fx.x], xfA, (int)(tfB & 3u), fShapeOff[fx.y], xfB);
  } else {
    collide_dispatch(m, shapes, (int)(tfA & 3u), fShapeOff[fx.x], xfA, (int)(tfB & 3u), fShapeOff[fx.y], xfB);
    touching = m.pointCount > 0;
    // carry warm-start impulses across by feature id (b2_contact.cpp:158-181)
    float4 o1 = C.m1[i], o2 = C.m2[i], o3 = C.m3[i];
    int oldCount = __float_as_int(o3.w);
    uint32_t oid0 = __float_as_uint(o3.x), oid1 = __float_as_uint(o3.y);
    for (int k = 0; k < m.pointCount; ++k) {
      uint32_t id = m.id[k];
      if (oldCount > 0 && oid0 == id) {
        m.normalImp[k] = o1.z;
        m.tangentImp[k] = o1.w;
      } else if (oldCount > 1 && oid1 == id) {
        m.normalImp[k] = o2.z;
        m.tangentImp[k] = o2.w;
      }
    }
    if (touching != wasTouching) {
      if (B2G_BODY_TYPE(fa) != B2G_STATIC) atomicOr(&bflagsRW[bd.x], B2G_BODY_WAKE_REQUEST);
      if (B2G_BODY_TYPE(fb) != B2G_STATIC) atomicOr(&bflagsRW[bd.y], B2G_BODY_WAKE_REQUEST);
      // an island edge disappeared: last step's island labels of this island cannot seed the
      // union-find any more (it may have split)
      if (!touching) islandDirty[islandPrev[B2G_BODY_TYPE(fa) != B2G_STATIC ? bd.x : bd.y]] = 1;
    }
  }
  float4 q0, q1, q2, q3;
  manifold_pack(m, q0, q1, q2, q3);
  C.m0[i] = q0;
  C.m1[i] = q1;
  C.m2[i] = q2;
  C.m3[i] = q3;
  flags = touching ? (flags | B2G_CONTACT_TOUCHING) : (flags & ~B2G_CONTACT_TOUCHING);
  C.flags[i] = flags;
  if (touching) {
    auto g = cg::coalesced_threads();
    if (g.thread_rank() == 0) atomicAdd(&counts->numTouching, (int)g.size());
  }
  if (recordEvents && touching != wasTouching) {
    if (touching) {
      int k = atomicAdd(&counts->beginCount, 1);
      if (k < eventCap) beginEvents[k] = fx;
    } else {
      int k = atomicAdd(&counts->endCount, 1);
      if (k < eventCap) endEvents[k] = fx;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Islands: connected components over touching, enabled, non-sensor contacts between non-static
// bodies (b2_world.cpp:560-647).  Lock-free union-find, smaller index wins, so the root of an
// island is its smallest body index — deterministic whatever the thread order.
// ---------------------------------------------------------------------------------------------
__global__ void k_body_begin(int nb, uint32_t* bflags, float4* force, int* islandParent, uint32_t* islandAwake,
                             const int* __restrict__ islandPrev, const uint8_t* __restrict__ islandDirty, int labelsValid,
                             const uint8_t* __restrict__ islandWasBig, int exactStep,
                             uint32_t* islandMinSleep, uint32_t* islandPen, int penStride, int posIters,
                             unsigned long long* colourMask, unsigned long long* bodyBest, int* islandCount,
                             int* islandCursor, int* binFirst, int* binEnd, int nbinsPlus, int* bucketCount,
                             int nbuckets) {
  B2G_PDL_ENTER();
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  // per-step tables of the fused solver, cleared here instead of by separate memsets
  for (int k = b; k < nbinsPlus; k += gridDim.x * blockDim.x) {
    binFirst[k] = 0x7f7f7f7f;
    binEnd[k] = 0;
  }
  for (int k = b; k < nbuckets; k += gridDim.x * blockDim.x) bucketCount[k] = 0;
  if (b >= nb) return;
  islandCount[b] = 0;
  islandCursor[b] = 0;
  uint32_t f = bflags[b];
  if (f & B2G_BODY_WAKE_REQUEST) {
    // b2Body::SetAwake(true): b2_body.h:726-730
    f = (f | B2G_BODY_AWAKE) & ~B2G_BODY_WAKE_REQUEST;
    bflags[b] = f;
    float4 fo = force[b];
    fo.w = 0.0f;
    force[b] = fo;
  }
  // Seed the union-find with last step's labels unless that island lost an edge: surviving islands
  // start flattened (depth 1), so the union pass over their old edges is two loads per edge.
  int seed = b;
  if (labelsValid) {
    int r = islandPrev[b];
    // Tile-sized islands: exact — rebuilt from their edges whenever one disappeared.  Oversize
    // islands (a settled pile) lose and gain edges every step while staying connected, and
    // re-uniting 100k bodies under one root is the most contended thing in the step, so their
    // labels are kept between exact recomputations every B2G_ISLAND_EXACT_PERIOD steps.  A stale
    // label can only keep a piece that just broke off attached a few steps longer (it then shares
    // the pile's sleep timer and position-iteration early exit): conservative, never unsafe.
    bool reuse = islandWasBig[r] ? !exactStep : !islandDirty[r];
    if (reuse) seed = r;
  }
  islandParent[b] = seed;
  islandAwake[b] = 0;
  islandMinSleep[b] = __float_as_uint(B2G_MAX_FLOAT);
  for (int k = 0; k < posIters; ++k) islandPen[(size_t)k * penStride + b] = 0;
  colourMask[b] = 0ull;
  bodyBest[b] = 0ull;
}

__device__ __forceinline__ int uf_find(int* parent, int x) {
  int p = parent[x];
  while (p != x) {
    int gp = parent[p];
    if (gp != p) parent[x] = gp;  // path halving
    x = p;
    p = gp;
  }
  return x;
}
__device__ __forceinline__ void uf_union(int* parent, int a, int b) {
  while (true) {
    a = uf_find(parent, a);
    b = uf_find(parent, b);
    if (a == b) return;
    if (a > b) {
      int t = a;
      a = b;
      b = t;
    }
    // hook the larger root under the smaller one
    int old = atomicCAS(&parent[b], b, a);
    if (old == b) return;
  }
}

__global__ void k_island_union(int nc, ContactBuf C, const uint32_t* __restrict__ bflags,
                               const uint32_t* __restrict__ fTypeFlags, int* island) {
  B2G_PDL_ENTER();
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nc) return;
  // Every load is issued before the first test: the kernel is a chain of dependent round trips
  // (flags -> fixtures -> bodies -> parents), so the three contact words go out together, then the
  // six words they index, instead of one round trip per early-out.  Dead slots keep in-range indices.
  uint32_t flags = C.flags[i];
  int2 fx = C.fix[i];
  int2 bd = C.body[i];
  uint32_t tfa = fTypeFlags[fx.x], tfb = fTypeFlags[fx.y];
  uint32_t bfa = bflags[bd.x], bfb = bflags[bd.y];
  int pa = island[bd.x], pb = island[bd.y];  // first hop of both finds (a body's parent is itself or an ancestor)
  if ((flags & (B2G_CONTACT_TOUCHING | B2G_CONTACT_ENABLED)) != (B2G_CONTACT_TOUCHING | B2G_CONTACT_ENABLED)) return;
  if ((tfa | tfb) & B2G_FIX_SENSOR) return;
  if (B2G_BODY_TYPE(bfa) == B2G_STATIC || B2G_BODY_TYPE(bfb) == B2G_STATIC) return;
  uf_union(island, pa, pb);
}

__global__ void k_island_union_joints(int nj, const int2* __restrict__ jBodies, const uint32_t* __restrict__ bflags,
                                      int* island) {
  B2G_PDL_ENTER();
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= nj) return;
  int2 bd = jBodies[j];
  uint32_t fa = bflags[bd.x], fb = bflags[bd.y];
  if (!(fa & B2G_BODY_ENABLED) || !(fb & B2G_BODY_ENABLED)) return;
  if (B2G_BODY_TYPE(fa) == B2G_STATIC || B2G_BODY_TYPE(fb) == B2G_STATIC) return;
  uf_union(island, bd.x, bd.y);
}

// Flatten into a SEPARATE array with a read-only find: compressing paths in place here would
// race with other threads' finds (a late path-halving store can overwrite a finished entry with
// a non-root ancestor), which made island ids — and therefore the whole step — nondeterministic.
__global__ void k_island_flatten(int nb, const uint32_t* __restrict__ bflags, const int* __restrict__ parent,
                                 int* island, uint32_t* islandAwake, int* islandCount, StepCounts* counts,
                                 uint8_t* islandDirty) {
  B2G_PDL_ENTER();
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  int root = -1;
  uint32_t f = 0;
  if (b < nb) {
    islandDirty[b] = 0;  // consumed by k_body_begin of this step
    root = b;
    for (int p = parent[root]; p != root; p = parent[root]) root = p;
    island[b] = root;
    f = bflags[b];
  }
  // island size = its non-static members (all of them are simulated once the island is awake).  One atomic per
  // warp and island: the 100 000 bodies of a settled pile would otherwise all hit one counter.
  const bool counted = b < nb && B2G_BODY_TYPE(f) != B2G_STATIC;
  const unsigned int peers = __match_any_sync(0xffffffffu, counted ? root : -1);
  if (counted && (int)(threadIdx.x & 31) == __ffs(peers) - 1) {
    int c = atomicAdd(&islandCount[root], __popc(peers)) + __popc(peers);
    if (c > counts->maxIslandBodies) atomicMax(&counts->maxIslandBodies, c);
  }
  // an island is simulated when any member could seed it (b2_world.cpp:526-545)
  if (b < nb && (f & B2G_BODY_AWAKE) && (f & B2G_BODY_ENABLED) && B2G_BODY_TYPE(f) != B2G_STATIC && !islandAwake[root])
    islandAwake[root] = 1u;
}

// ---------------------------------------------------------------------------------------------
// Velocity integration for every body of an awake island (b2_island.cpp:257-293); reached
// bodies are forced awake without resetting their sleep timer (b2_world.cpp:574).
// ---------------------------------------------------------------------------------------------
__global__ void k_integrate_velocities(int nb, uint32_t* bflags, const int* __restrict__ island,
                                       const uint32_t* __restrict__ islandAwake, float4* vel,
                                       const float4* __restrict__ mass, const float4* __restrict__ center,
                                       const float4* __restrict__ force, float h, float2 gravity,
                                       StepCounts* counts, const int* __restrict__ bodySlot, int onlyBig) {
  B2G_PDL_ENTER();
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nb) return;
  if (onlyBig && bodySlot[b] != -2) return;  // bodies of tile-sized islands are integrated by the fused kernel
  uint32_t f = bflags[b];
  uint32_t type = B2G_BODY_TYPE(f);
  if (type == B2G_STATIC) return;
  if (!islandAwake[island[b]]) return;
  if (!(f & B2G_BODY_AWAKE)) bflags[b] = f | B2G_BODY_AWAKE;
  if (type != B2G_DYNAMIC) return;
  float4 v4 = vel[b];
  float4 m4 = mass[b];    // invMass, invI, mass, gravityScale
  float4 c4 = center[b];  // lc.x, lc.y, linDamp, angDamp
  float4 f4 = force[b];   // f.x, f.y, torque, sleepTime
  float2 v = make_float2(v4.x, v4.y);
  float w = v4.z;
  float t = h * m4.x;
  float gs = m4.w * m4.z;
  v.x += t * (gs * gravity.x + f4.x);
  v.y += t * (gs * gravity.y + f4.y);
  w += h * m4.y * f4.z;
  float dl = 1.0f + h * c4.z;
  float da = 1.0f + h * c4.w;
  v.x /= dl;
  v.y /= dl;
  w /= da;
  vel[b] = make_float4(v.x, v.y, w, v4.w);
}

// ---------------------------------------------------------------------------------------------
// Solver constraint list: touching, enabled, non-sensor contacts of awake islands
// (b2_world.cpp:586-599).  Inactive contacts lose their colour.
// ---------------------------------------------------------------------------------------------
__global__ void k_mark_active(int nc, ContactBuf C, const uint32_t* __restrict__ fTypeFlags,
                              const int* __restrict__ island, const uint32_t* __restrict__ islandAwake,
                              uint8_t* activeFlag, int dropColours) {
  B2G_PDL_ENTER();
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nc) return;
  uint32_t flags = C.flags[i];
  bool active = (flags & (B2G_CONTACT_TOUCHING | B2G_CONTACT_ENABLED)) == (B2G_CONTACT_TOUCHING | B2G_CONTACT_ENABLED);
  if (active) {
    int2 fx = C.fix[i];
    if ((fTypeFlags[fx.x] | fTypeFlags[fx.y]) & B2G_FIX_SENSOR) active = false;
  }
  if (active) {
    int2 bd = C.body[i];
    active = (islandAwake[island[bd.x]] | islandAwake[island[bd.y]]) != 0;
  }
  activeFlag[i] = active ? 1 : 0;
  if (!active || dropColours) C.colour[i] = -1;
}

// ---------------------------------------------------------------------------------------------
// Greedy graph colouring of the constraint graph (two constraints conflict when they share a
// body the solver may move).  Colours persist from step to step on the contact; only new
// constraints are coloured, by Luby-style rounds: an uncoloured constraint that holds the
// highest hashed priority on both of its bodies takes the lowest colour free on both.  Every
// decision is a pure function of the constraint list, so the colouring is deterministic.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ bool body_movable(float4 m) { return m.x != 0.0f || m.y != 0.0f; }

__global__ void k_colour_begin(const int* __restrict__ numActivePtr, const int* __restrict__ activeList, ContactBuf C,
                               const float4* __restrict__ mass, unsigned long long* colourMask, StepCounts* counts) {
  B2G_PDL_ENTER();
  int n = *numActivePtr;
  for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < n; s += gridDim.x * blockDim.x) {
    int i = activeList[s];
    int c = C.colour[i];
    if (c >= B2G_MAX_COLOURS) {
      C.colour[i] = -1;  // overflow constraints retry every step
      c = -1;
    }
    if (c >= 0) {
      int2 bd = C.body[i];
      unsigned long long bit = 1ull << c;
      if (body_movable(mass[bd.x])) atomicOr(&colourMask[bd.x], bit);
      if (body_movable(mass[bd.y])) atomicOr(&colourMask[bd.y], bit);
      atomicAdd(&counts->colourCount[c], 1);
    }
  }
}

// Unique per constraint (a bijection of the 56-bit pair key) so two constraints on one body can
// never tie, and a function of the pair alone so colours do not depend on contact slot numbers.
__device__ __forceinline__ unsigned long long colour_key56(unsigned long long key) {
  const unsigned long long M56 = (1ull << 56) - 1ull;
  unsigned long long k = ((key >> 32) << 28 | (key & 0xfffffffull)) & M56;  // fixture indices < 2^28
  k = (k * 0x9e3779b97f4a7c15ull) & M56;  // odd multiplier: bijective mod 2^56
  k ^= k >> 29;                          // xor-shift: bijective
  k = (k * 0xbf58476d1ce4e5b9ull) & M56;
  k ^= k >> 31;
  return k;
}
__device__ __forceinline__ unsigned long long colour_priority(int round, int s, unsigned long long key) {
  (void)s;
  return ((unsigned long long)(round + 1) << 56) | colour_key56(key);
}

__global__ void k_colour_propose(const int* __restrict__ numActivePtr, const int* __restrict__ activeList, ContactBuf C,
                                 const float4* __restrict__ mass, unsigned long long* bodyBest, int round) {
  B2G_PDL_ENTER();
  int n = *numActivePtr;
  for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < n; s += gridDim.x * blockDim.x) {
    int i = activeList[s];
    if (C.colour[i] >= 0) continue;
    int2 bd = C.body[i];
    unsigned long long pr = colour_priority(round, s, C.key[i]);
    if (body_movable(mass[bd.x])) atomicMax(&bodyBest[bd.x], pr);
    if (body_movable(mass[bd.y])) atomicMax(&bodyBest[bd.y], pr);
  }
}

__global__ void k_colour_commit(const int* __restrict__ numActivePtr, const int* __restrict__ activeList, ContactBuf C,
                                const float4* __restrict__ mass, unsigned long long* colourMask,
                                const unsigned long long* __restrict__ bodyBest, int round, StepCounts* counts,
                                int lastOfBatch) {
  B2G_PDL_ENTER();
  int n = *numActivePtr;
  for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < n; s += gridDim.x * blockDim.x) {
    int i = activeList[s];
    if (C.colour[i] >= 0) continue;
    int2 bd = C.body[i];
    unsigned long long pr = colour_priority(round, s, C.key[i]);
    bool movA = body_movable(mass[bd.x]), movB = body_movable(mass[bd.y]);
    bool win = (!movA || bodyBest[bd.x] == pr) && (!movB || bodyBest[bd.y] == pr);
    if (win) {
      unsigned long long used = (movA ? colourMask[bd.x] : 0ull) | (movB ? colourMask[bd.y] : 0ull);
      unsigned long long freeBits = ~used & ((1ull << B2G_MAX_COLOURS) - 1ull);
      int c = freeBits ? (__ffsll((long long)freeBits) - 1) : B2G_OVERFLOW_COLOUR;
      C.colour[i] = c;
      if (c < B2G_MAX_COLOURS) {
        unsigned long long bit = 1ull << c;
        // the winner is unique on each movable body, so plain read-modify-write is race free
        if (movA) colourMask[bd.x] |= bit;
        if (movB) colourMask[bd.y] |= bit;
      }
      atomicAdd(&counts->colourCount[c], 1);
      if (counts->lastUsefulRound < round + 1) atomicMax(&counts->lastUsefulRound, round + 1);
    } else if (lastOfBatch) {
      atomicAdd(&counts->remaining, 1);
    }
  }
}

__global__ void k_colour_keys(const int* __restrict__ numActivePtr, const int* __restrict__ activeList, ContactBuf C,
                              uint8_t* colourKey) {
  B2G_PDL_ENTER();
  int n = *numActivePtr;
  for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < n; s += gridDim.x * blockDim.x)
    colourKey[s] = (uint8_t)C.colour[activeList[s]];
}

__global__ void k_gather_keys(int n, const int* __restrict__ list, ContactBuf C, unsigned long long* keys) {
  B2G_PDL_ENTER();
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < n) keys[s] = C.key[list[s]];
}

__global__ void k_gather_order(int n, const int* __restrict__ list, const unsigned long long* __restrict__ orderKey,
                               unsigned long long* keys) {
  B2G_PDL_ENTER();
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < n) keys[s] = orderKey[list[s]];
}

// self-check used by the tests: counts pairs of same-colour constraints that share a movable body
__global__ void k_colour_validate(int n, const int* __restrict__ sortedList, ContactBuf C,
                                  const float4* __restrict__ mass, int* bodyStamp, int* violations) {
  B2G_PDL_ENTER();
  // bodyStamp[b * stride + colour] would be exact; instead each constraint claims (body, colour)
  // with atomicExch on a per-body slot tagged by colour, one colour range per launch (host loops)
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  int i = sortedList[s];
  int2 bd = C.body[i];
  if (body_movable(mass[bd.x])) {
    int old = atomicExch(&bodyStamp[bd.x], i + 1);
    if (old != 0) atomicAdd(violations, 1);
  }
  if (body_movable(mass[bd.y])) {
    int old = atomicExch(&bodyStamp[bd.y], i + 1);
    if (old != 0) atomicAdd(violations, 1);
  }
}

// ---------------------------------------------------------------------------------------------
// Contact solver kernels.  Slot s of the solver planes is the s-th entry of sortedList (colour
// order in the production mode, contact order in the sequential mode).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
k_prepare(int first, int n, const int* __restrict__ sortedList, ContactBuf C, const float* __restrict__ fRadius,
          const uint32_t* __restrict__ bflags, const int* __restrict__ island, SolverPlanes S, int* croot,
          const float4* __restrict__ pos, const float4* __restrict__ vel, const float4* __restrict__ mass,
          const float4* __restrict__ center, float dtRatio, int warmStarting) {
  B2G_PDL_ENTER();
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  s += first;
  int i = sortedList[s];
  Manifold m;
  manifold_unpack(m, C.m0[i], C.m1[i], C.m2[i], C.m3[i]);
  int2 bd = C.body[i];
  int2 fx = C.fix[i];
  prepare_constraint(S, s, i, m, bd.x, bd.y, bd.x, bd.y, C.material[i], fRadius[fx.x], fRadius[fx.y],
                     GlobalBodies{const_cast<float4*>(pos)}, GlobalBodies{const_cast<float4*>(vel)}, mass, center,
                     dtRatio, warmStarting != 0);
  croot[s] = B2G_BODY_TYPE(bflags[bd.x]) != B2G_STATIC ? island[bd.x] : island[bd.y];
}

__global__ void __launch_bounds__(256) k_warm_start(int first, int last, SolverPlanes S, float4* vel) {
  B2G_PDL_ENTER();
  int s = first + blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= last) return;
  warm_start_constraint(S, s, GlobalBodies{vel});
}

__global__ void __launch_bounds__(256) k_solve_velocity(int first, int last, SolverPlanes S, float4* vel) {
  B2G_PDL_ENTER();
  int s = first + blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= last) return;
  solve_velocity_constraint(S, s, GlobalBodies{vel});
}

// island has converged in an earlier position iteration? (b2_island.cpp:391-409 early exit)
__device__ __forceinline__ bool island_done(const uint32_t* __restrict__ islandPen, int penStride, int iter, int root) {
  if (iter == 0) return false;
  float pen = __uint_as_float(islandPen[(size_t)(iter - 1) * penStride + root]);
  return pen <= 3.0f * B2G_LINEAR_SLOP;
}

__global__ void __launch_bounds__(256)
k_solve_position(int first, int last, SolverPlanes S, float4* pos, const int* __restrict__ croot, uint32_t* islandPen,
                 int penStride, int iter) {
  B2G_PDL_ENTER();
  int s = first + blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= last) return;
  int root = croot[s];
  if (island_done(islandPen, penStride, iter, root)) return;
  float minSep = solve_position_constraint(S, s, GlobalBodies{pos});
  // penetration >= +0.0 (never -0.0, whose bit pattern would win the max), so the float bit
  // pattern is monotone as an unsigned integer
  float pen = minSep < 0.0f ? -minSep : 0.0f;
  atomicMax(&islandPen[(size_t)iter * penStride + root], __float_as_uint(pen));
}

// sequential single-thread variants: same device functions, list order (parity vehicle and the
// serial overflow colour)
__global__ void k_warm_start_seq(int first, int last, SolverPlanes S, float4* vel) {
  B2G_PDL_ENTER();
  for (int s = first; s < last; ++s) warm_start_constraint(S, s, GlobalBodies{vel});
}
__global__ void k_solve_velocity_seq(int first, int last, SolverPlanes S, float4* vel) {
  B2G_PDL_ENTER();
  for (int s = first; s < last; ++s) solve_velocity_constraint(S, s, GlobalBodies{vel});
}
__global__ void k_solve_position_seq(int first, int last, SolverPlanes S, float4* pos, const int* __restrict__ croot,
                                     uint32_t* islandPen, int penStride, int iter) {
  B2G_PDL_ENTER();
  for (int s = first; s < last; ++s) {
    int root = croot[s];
    if (island_done(islandPen, penStride, iter, root)) continue;
    float minSep = solve_position_constraint(S, s, GlobalBodies{pos});
    uint32_t* slot = &islandPen[(size_t)iter * penStride + root];
    uint32_t v = __float_as_uint(minSep < 0.0f ? -minSep : 0.0f);
    if (v > *slot) *slot = v;
  }
}

__global__ void k_store_impulses(int first, int n, SolverPlanes S, ContactBuf C) {
  B2G_PDL_ENTER();
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  s += first;
  int4 ix = S.idx[s];
  float4 imp = S.imp[s];
  int i = ix.w;
  float4 q1 = C.m1[i];
  q1.z = imp.x;
  q1.w = imp.y;
  C.m1[i] = q1;
  if (ix.z == 2) {
    float4 q2 = C.m2[i];
    q2.z = imp.z;
    q2.w = imp.w;
    C.m2[i] = q2;
  }
}

__device__ __forceinline__ bool body_simulated(uint32_t f, const int* __restrict__ island,
                                               const uint32_t* __restrict__ islandAwake, int b) {
  // A DISABLED body cannot seed an island (b2_world.cpp:531) but in this fork it keeps its fixtures
  // and contacts (b2Body::SetEnabled only flips the flag, b2_body.cpp:512-559), so it is pulled into
  // — and simulated with — any island that reaches it through a contact.
  return B2G_BODY_TYPE(f) != B2G_STATIC && islandAwake[island[b]];
}

__global__ void k_integrate_positions(int nb, const uint32_t* __restrict__ bflags, const int* __restrict__ island,
                                      const uint32_t* __restrict__ islandAwake, float4* pos, float4* vel, float h,
                                      const int* __restrict__ bodySlot, int onlyBig) {
  B2G_PDL_ENTER();
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nb) return;
  if (onlyBig && bodySlot[b] != -2) return;
  if (!body_simulated(bflags[b], island, islandAwake, b)) return;
  float4 p4 = pos[b], v4 = vel[b];
  float2 v = make_float2(v4.x, v4.y);
  float w = v4.z;
  float2 translation = h * v;
  if (dot2(translation, translation) > B2G_MAX_TRANSLATION_SQ) {
    float ratio = B2G_MAX_TRANSLATION / len2(translation);
    v.x *= ratio;
    v.y *= ratio;
  }
  float rotation = h * w;
  if (rotation * rotation > B2G_MAX_ROTATION_SQ) {
    float ratio = B2G_MAX_ROTATION / absf_(rotation);
    w *= ratio;
  }
  p4.x += h * v.x;
  p4.y += h * v.y;
  p4.z += h * w;
  pos[b] = p4;
  vel[b] = make_float4(v.x, v.y, w, v4.w);
}

// write-back + SynchronizeTransform (b2_island.cpp:430-438) and the per-body half of the sleep
// bookkeeping (b2_island.cpp:452-472)
__global__ void k_finalize_bodies(int nb, const uint32_t* __restrict__ bflags, const int* __restrict__ island,
                                  const uint32_t* __restrict__ islandAwake, const float4* __restrict__ pos,
                                  const float4* __restrict__ vel, const float4* __restrict__ center, float4* xf,
                                  float4* force, uint32_t* islandMinSleep, float h, int allowSleep,
                                  const int* __restrict__ bodySlot, int onlyBig) {
  B2G_PDL_ENTER();
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nb) return;
  if (onlyBig && bodySlot[b] != -2) return;
  uint32_t f = bflags[b];
  if (!body_simulated(f, island, islandAwake, b)) return;
  float4 p4 = pos[b], c4 = center[b];
  Xf T = xf_from_sweep(make_float2(p4.x, p4.y), p4.z, make_float2(c4.x, c4.y));
  xf[b] = xf_to4(T);
  if (allowSleep) {
    float4 v4 = vel[b];
    float4 fo = force[b];
    const float linTolSqr = B2G_LINEAR_SLEEP_TOL * B2G_LINEAR_SLEEP_TOL;
    const float angTolSqr = B2G_ANGULAR_SLEEP_TOL * B2G_ANGULAR_SLEEP_TOL;
    float minSleep;
    if (!(f & B2G_BODY_AUTOSLEEP) || v4.z * v4.z > angTolSqr || v4.x * v4.x + v4.y * v4.y > linTolSqr) {
      fo.w = 0.0f;
      minSleep = 0.0f;
    } else {
      fo.w += h;
      minSleep = fo.w;
    }
    force[b] = fo;
    atomicMin(&islandMinSleep[island[b]], __float_as_uint(minSleep));
  }
}

// island sleep (b2_island.cpp:474-482), b2Body::SetAwake(false) (b2_body.h:731-739) and
// b2World::ClearForces (b2_world.cpp:1173-1180)
__global__ void k_sleep_and_clear(int nb, uint32_t* bflags, const int* __restrict__ island,
                                  const uint32_t* __restrict__ islandAwake,
                                  const uint32_t* __restrict__ islandMinSleep,
                                  const uint32_t* __restrict__ islandPen, int penStride, int posIters, float4* vel,
                                  float4* force, int allowSleep, int clearForces, StepCounts* counts,
                                  const int* __restrict__ bodySlot, int onlyBig) {
  B2G_PDL_ENTER();
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nb) return;
  if (onlyBig && bodySlot[b] != -2) return;
  uint32_t f = bflags[b];
  float4 fo = force[b];
  bool dirty = false;
  if (allowSleep && body_simulated(f, island, islandAwake, b)) {
    int root = island[b];
    float minSleep = __uint_as_float(islandMinSleep[root]);
    bool positionSolved =
        posIters > 0 && __uint_as_float(islandPen[(size_t)(posIters - 1) * penStride + root]) <= 3.0f * B2G_LINEAR_SLOP;
    if (minSleep >= B2G_TIME_TO_SLEEP && positionSolved) {
      f &= ~B2G_BODY_AWAKE;
      bflags[b] = f;
      vel[b] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
      fo = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
      dirty = true;
    }
  }
  if (clearForces && (fo.x != 0.0f || fo.y != 0.0f || fo.z != 0.0f)) {
    fo.x = fo.y = fo.z = 0.0f;
    dirty = true;
  }
  if (dirty) force[b] = fo;
  if ((f & B2G_BODY_AWAKE) && B2G_BODY_TYPE(f) != B2G_STATIC) {
    auto g = cg::coalesced_threads();
    if (g.thread_rank() == 0) atomicAdd(&counts->numAwake, (int)g.size());
  }
}

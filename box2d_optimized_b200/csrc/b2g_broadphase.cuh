// b2g_broadphase.cuh — pair finding on the device.
//
// Replaces b2BroadPhase::UpdateAndQuery / BuildAndQuery / QueryAll
// (include/box2d/b2_broad_phase.h:252-511, 587-620), the pair sink
// b2ContactManager::QueryCallback (src/dynamics/b2_contact_manager.cpp:136-188) with its filters
// (b2Body::ShouldCollide b2_body.cpp:459-480, b2ContactFilter::ShouldCollide
// b2_world_callbacks.cpp:28-40, the b2Contact::Create type table b2_contact.cpp:58-77) and the
// persist / dead-contact protocol (b2_contact_manager.cpp:97,123-134; b2_world.cpp:125-138).
//
// The reference rebuilds a median-split BVH top-down and queries it during the build; the pair
// SET it reports is exactly "all fixture pairs with inclusive tight-AABB overlap" (SURVEY
// Appendix B.20), independent of tree shape.  The device path therefore builds a linear BVH
// (Karras 2012) over Morton-sorted fixture AABBs every step and lets every leaf traverse it,
// reporting each pair once (by its smaller-AABB member); a hash probe either flags the pair's
// existing contact as persisting or queues the pair (warp-aggregated compaction) for creation.
#pragma once
#include <cooperative_groups/reduce.h>
#include "b2g_step_kernels.cuh"

// tight AABBs for fixtures of non-static bodies (b2_broad_phase.h:265-276); static AABBs are
// frozen at creation and only recomputed when `all` is set.  Also reduces
// the bounds of 2*centre for the Morton normalisation.
__global__ void __launch_bounds__(256)
k_update_aabbs(int nf, const int* __restrict__ fBody, const int* __restrict__ fShapeOff,
               const uint32_t* __restrict__ fTypeFlags, const float4* __restrict__ shapes,
               const uint32_t* __restrict__ bflags, const float4* __restrict__ xf, float4* fAabb, float* fRadius,
               int all, StepCounts* counts) {
  B2G_PDL_ENTER();
  int f = blockIdx.x * blockDim.x + threadIdx.x;
  float cx = 0.0f, cy = 0.0f;
  bool valid = false;
  if (f < nf) {
    uint32_t tf = fTypeFlags[f];
    if (!(tf & B2G_FIX_DEAD)) {
      int b = fBody[f];
      uint32_t bf = bflags[b];
      float4 box;
      // all: 0 = refresh non-static fixtures, 1 = refresh everything, 2 = boxes given, refresh nothing
      if (all == 1 || (all == 0 && B2G_BODY_TYPE(bf) != B2G_STATIC)) {
        int type = (int)(tf & 3u), off = fShapeOff[f];
        box = shape_aabb(shapes, type, off, xf_from4(xf[b]));
        fAabb[f] = box;
        if (all == 1) {
          float r;
          if (type == 0) r = __ldg(shapes + off).z;
          else if (type == 1) r = __ldg(shapes + off + 2).x;
          else r = __ldg(shapes + off).z;
          fRadius[f] = r;
        }
      } else {
        box = fAabb[f];
      }
      cx = box.x + box.z;
      cy = box.y + box.w;
      valid = true;
    }
  }
  // block reduce min/max of the doubled centres, then 4 atomics per block
  unsigned int lox = valid ? float_flip(cx) : 0xffffffffu, loy = valid ? float_flip(cy) : 0xffffffffu;
  unsigned int hix = valid ? float_flip(cx) : 0u, hiy = valid ? float_flip(cy) : 0u;
  for (int o = 16; o > 0; o >>= 1) {
    lox = min(lox, __shfl_xor_sync(0xffffffffu, lox, o));
    loy = min(loy, __shfl_xor_sync(0xffffffffu, loy, o));
    hix = max(hix, __shfl_xor_sync(0xffffffffu, hix, o));
    hiy = max(hiy, __shfl_xor_sync(0xffffffffu, hiy, o));
  }
  __shared__ unsigned int sm[4][8];
  int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) {
    sm[0][wid] = lox;
    sm[1][wid] = loy;
    sm[2][wid] = hix;
    sm[3][wid] = hiy;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int nw = blockDim.x >> 5;
    for (int k = 1; k < nw; ++k) {
      lox = min(lox, sm[0][k]);
      loy = min(loy, sm[1][k]);
      hix = max(hix, sm[2][k]);
      hiy = max(hiy, sm[3][k]);
    }
    lox = min(sm[0][0], lox);
    loy = min(sm[1][0], loy);
    hix = max(sm[2][0], hix);
    hiy = max(sm[3][0], hiy);
    atomicMin(&counts->boundsLo[0], lox);
    atomicMin(&counts->boundsLo[1], loy);
    atomicMax(&counts->boundsHi[0], hix);
    atomicMax(&counts->boundsHi[1], hiy);
  }
}

__device__ __forceinline__ unsigned int expand_bits16(unsigned int v) {
  v &= 0xffffu;
  v = (v | (v << 8)) & 0x00ff00ffu;
  v = (v | (v << 4)) & 0x0f0f0f0fu;
  v = (v | (v << 2)) & 0x33333333u;
  v = (v | (v << 1)) & 0x55555555u;
  return v;
}

// sort key = world id (high) | 32-bit Morton code of the AABB centre (low); dead fixtures sort last
__global__ void k_morton_keys(int nf, const float4* __restrict__ fAabb, const uint32_t* __restrict__ fTypeFlags,
                              const int* __restrict__ fBody, const int* __restrict__ bworld,
                              const StepCounts* __restrict__ counts, unsigned long long* keys, int* leafFixture,
                              int numWorlds) {
  B2G_PDL_ENTER();
  int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= nf) return;
  leafFixture[f] = f;
  if (fTypeFlags[f] & B2G_FIX_DEAD) {
    keys[f] = ((unsigned long long)numWorlds << 32) | 0xffffffffull;
    return;
  }
  float lox = float_unflip(counts->boundsLo[0]), loy = float_unflip(counts->boundsLo[1]);
  float hix = float_unflip(counts->boundsHi[0]), hiy = float_unflip(counts->boundsHi[1]);
  float ex = hix - lox, ey = hiy - loy;
  // ONE scale for both axes: square cells.  Per-axis scaling made the cells of a wide, flat world
  // (100 pyramids side by side: 3 km x 20 m) 150 times wider than tall, so the y bits below a
  // box height were noise that scrambled the order along x — runs of consecutive leaves spanned
  // ~90 m and every query walked three to four nodes per tree level.
  float em = fmaxf(ex, ey);
  float sx = em > 0.0f ? 65535.0f / em : 0.0f;
  float sy = sx;
  float4 box = fAabb[f];
  float cx = box.x + box.z, cy = box.y + box.w;
  unsigned int qx = (unsigned int)fminf(fmaxf((cx - lox) * sx, 0.0f), 65535.0f);
  unsigned int qy = (unsigned int)fminf(fmaxf((cy - loy) * sy, 0.0f), 65535.0f);
  unsigned int morton = (expand_bits16(qx) << 1) | expand_bits16(qy);
  unsigned long long w = numWorlds > 1 ? (unsigned long long)bworld[fBody[f]] : 0ull;
  keys[f] = (w << 32) | morton;
}

// joints with collideConnected == false -> body-pair keys + per-body marks (rebuilt when the joint
// table changes; b2World::CreateJoint / DestroyJoint flag the pair's contacts for filtering,
// b2_world.cpp:307-323, 389-405)
__global__ void k_joint_filter_build(int nj, const int2* __restrict__ jBodies, const float4* __restrict__ jParams1,
                                     unsigned long long* keys, uint8_t* bodyNoCollide) {
  B2G_PDL_ENTER();
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= nj) return;
  int2 bd = jBodies[j];
  uint32_t flags = __float_as_uint(jParams1[j].y);
  if (flags & 4u) {  // collideConnected
    keys[j] = 0xffffffffffffffffull;
    return;
  }
  keys[j] = ((unsigned long long)min(bd.x, bd.y) << 32) | (unsigned long long)max(bd.x, bd.y);
  bodyNoCollide[bd.x] = 1;
  bodyNoCollide[bd.y] = 1;
}

// leaves in sorted order: box + everything the pair filter needs in one int4
//   x = fixture, y = body, z = type | sensor<<2 | dynamic<<3 | dead<<4 | noCollideJoint<<5 | group<<16,
//   w = category | mask<<16
__global__ void k_leaf_gather(int nf, const int* __restrict__ leafFixtureSorted,
                              const unsigned long long* __restrict__ keysSorted, const float4* __restrict__ fAabb,
                              const int* __restrict__ fBody, const uint32_t* __restrict__ fTypeFlags,
                              const uint2* __restrict__ fFilter, const uint32_t* __restrict__ bflags,
                              const uint8_t* __restrict__ bodyNoCollide, float4* leafBox,
                              int4* leafInfo, unsigned long long* leafKey, int* worldFirst, int* worldLast,
                              int numWorlds) {
  B2G_PDL_ENTER();
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= nf) return;
  int f = leafFixtureSorted[p];
  uint32_t tf = fTypeFlags[f];
  int b = fBody[f];
  uint32_t bf = bflags[b];
  uint2 fl = fFilter[f];
  bool dead = (tf & B2G_FIX_DEAD) != 0;
  leafBox[p] = dead ? make_float4(B2G_MAX_FLOAT, B2G_MAX_FLOAT, -B2G_MAX_FLOAT, -B2G_MAX_FLOAT) : fAabb[f];
  unsigned int z = (tf & 3u) | ((tf & B2G_FIX_SENSOR) ? 4u : 0u) | (B2G_BODY_TYPE(bf) == B2G_DYNAMIC ? 8u : 0u) |
                   (dead ? 16u : 0u) | (bodyNoCollide[b] ? 32u : 0u) | ((fl.y & 0xffffu) << 16);
  leafInfo[p] = make_int4(f, b, (int)z, (int)fl.x);
  // reporting order: the leaf with the SMALLER (size, position) key reports the pair, so a huge
  // AABB (ground edge, container wall) never walks the tree for its thousands of partners —
  // they each find it instead.  size = half perimeter as non-negative float bits (monotone).
  float4 bx = leafBox[p];
  float size = dead ? 0.0f : (bx.z - bx.x) + (bx.w - bx.y);
  leafKey[p] = ((unsigned long long)__float_as_uint(fmaxf(size, 0.0f)) << 32) | (unsigned int)p;
  if (numWorlds > 1) {
    unsigned int w = (unsigned int)(keysSorted[p] >> 32);
    bool lastOfWorld = (p == nf - 1) || ((unsigned int)(keysSorted[p + 1] >> 32) != w);
    bool firstOfWorld = (p == 0) || ((unsigned int)(keysSorted[p - 1] >> 32) != w);
    if (lastOfWorld && w < (unsigned int)numWorlds) worldLast[w] = p;
    if (firstOfWorld && w < (unsigned int)numWorlds) worldFirst[w] = p;
  }
}

// ---- implicit 8-wide BVH over the Morton-sorted leaves ------------------------------------------
// Level 0 = the leaves in sorted order; node j of level l covers children [8j, 8j+8) of level l-1,
// i.e. leaves [j * 8^l, (j+1) * 8^l).  No topology is stored: a node IS the 8 consecutive
// (box, max-key) entries of the level below — 128 + 64 contiguous bytes — so a visit is one round
// trip to L2 with eight independent loads, and a 21 k-leaf tree is 5 visits deep instead of the
// ~30 dependent visits of a binary LBVH (the walks are pure latency: ~4 warps per SM).  Any valid
// bounding hierarchy reports the same exact pair set; splitting by count instead of by the highest
// differing Morton bit costs a little overlap and buys an index-free layout, a refit that is a
// plain reduction (no atomics per level) and children that can be fetched before their parent is
// tested.  The leaf order is refreshed by re-sorting every few steps; in between only boxes move.
#define B2G_BVH_W 8
#define B2G_BVH_MAX_LEVELS 12
struct WideBvh {
  int levels;                        // internal levels; the root level has one node (0 when n == 0)
  int count[B2G_BVH_MAX_LEVELS];     // entries per level, count[0] = leaves
  int offset[B2G_BVH_MAX_LEVELS];    // where level l >= 1 starts inside box[] / key[]
  float4* box;
  unsigned long long* key;           // largest leaf key below the node (reporting-order pruning)
  int* done;                         // blocks finished (last-block-done hand-over in the refit)
};
__host__ __device__ inline void wide_bvh_layout(WideBvh& T, int n) {
  T.levels = 0;
  T.count[0] = n;
  T.offset[0] = 0;
  int off = 0, c = n;
  while (c > 1 && T.levels + 1 < B2G_BVH_MAX_LEVELS) {
    c = (c + B2G_BVH_W - 1) / B2G_BVH_W;
    ++T.levels;
    T.count[T.levels] = c;
    T.offset[T.levels] = off;
    off += (c + 7) & ~7;  // keep every level 128-byte aligned
  }
  if (n == 1) {  // a lone leaf still gets a root so that queries have something to visit
    T.levels = 1;
    T.count[1] = 1;
    T.offset[1] = 0;
  }
}
__host__ __device__ inline int wide_bvh_nodes(int n) {
  WideBvh T;
  wide_bvh_layout(T, n);
  return T.levels > 0 ? T.offset[T.levels] + 8 : 8;
}

__device__ __forceinline__ float4 box_union(float4 a, float4 b) {
  return make_float4(fminf(a.x, b.x), fminf(a.y, b.y), fmaxf(a.z, b.z), fmaxf(a.w, b.w));
}
#define B2G_EMPTY_BOX make_float4(B2G_MAX_FLOAT, B2G_MAX_FLOAT, -B2G_MAX_FLOAT, -B2G_MAX_FLOAT)

// One block = 256 consecutive leaves = 32 level-1 nodes = 4 level-2 nodes, all reduced in shared
// memory; the last block to finish reduces the few remaining upper levels.  With REFRESH the leaf
// boxes are first recomputed from the body transforms (the refit-only steps between re-sorts:
// replaces k_update_aabbs + k_morton_keys + sort + k_leaf_gather on those steps).
template <bool REFRESH>
__global__ void __launch_bounds__(256)
k_wide_refit(WideBvh T, const int* __restrict__ leafFixtureSorted, const int* __restrict__ fBody,
             const int* __restrict__ fShapeOff, const uint32_t* __restrict__ fTypeFlags,
             const float4* __restrict__ shapes, const uint32_t* __restrict__ bflags, const float4* __restrict__ xf,
             float4* fAabb, float4* leafBox, unsigned long long* leafKey) {
  B2G_PDL_ENTER();
  __shared__ float4 sBox[256];
  __shared__ unsigned long long sKey[256];
  __shared__ float4 s1Box[32];
  __shared__ unsigned long long s1Key[32];
  __shared__ int sLast;
  const int n = T.count[0];
  const int tid = threadIdx.x;
  const int p = blockIdx.x * 256 + tid;
  float4 box = B2G_EMPTY_BOX;
  unsigned long long key = 0ull;
  if (p < n) {
    box = leafBox[p];  // dead / static leaves keep the box of the last re-sort
    key = leafKey[p];
    if (REFRESH) {
      int f = leafFixtureSorted[p];
      uint32_t tf = fTypeFlags[f];
      if (!(tf & B2G_FIX_DEAD)) {
        int b = fBody[f];
        if (B2G_BODY_TYPE(bflags[b]) != B2G_STATIC) {
          box = shape_aabb(shapes, (int)(tf & 3u), fShapeOff[f], xf_from4(xf[b]));
          fAabb[f] = box;
          leafBox[p] = box;
          float size = (box.z - box.x) + (box.w - box.y);
          key = ((unsigned long long)__float_as_uint(fmaxf(size, 0.0f)) << 32) | (unsigned int)p;
          leafKey[p] = key;
        }
      }
    }
  }
  sBox[tid] = box;
  sKey[tid] = key;
  __syncthreads();
  if (tid < 32 && T.levels >= 1) {
    float4 u = B2G_EMPTY_BOX;
    unsigned long long k = 0ull;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      u = box_union(u, sBox[tid * 8 + c]);
      unsigned long long kc = sKey[tid * 8 + c];
      k = kc > k ? kc : k;
    }
    s1Box[tid] = u;
    s1Key[tid] = k;
    int j = blockIdx.x * 32 + tid;
    if (j < T.count[1]) {
      T.box[T.offset[1] + j] = u;
      T.key[T.offset[1] + j] = k;
    }
  }
  __syncthreads();
  if (tid < 4 && T.levels >= 2) {
    float4 u = B2G_EMPTY_BOX;
    unsigned long long k = 0ull;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      u = box_union(u, s1Box[tid * 8 + c]);
      unsigned long long kc = s1Key[tid * 8 + c];
      k = kc > k ? kc : k;
    }
    int j = blockIdx.x * 4 + tid;
    if (j < T.count[2]) {
      T.box[T.offset[2] + j] = u;
      T.key[T.offset[2] + j] = k;
    }
  }
  if (T.levels < 3) return;
  // upper levels: whoever finishes last has every level-2 entry visible
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    int old = atomicAdd(T.done, 1);
    sLast = (old == (int)gridDim.x - 1);
    if (sLast) *T.done = 0;  // ready for the next launch
  }
  __syncthreads();
  if (!sLast) return;
  __threadfence();
  for (int l = 3; l <= T.levels; ++l) {
    const float4* cb = T.box + T.offset[l - 1];
    const unsigned long long* ck = T.key + T.offset[l - 1];
    const int nc = T.count[l - 1];
    for (int j = tid; j < T.count[l]; j += 256) {
      float4 u = B2G_EMPTY_BOX;
      unsigned long long k = 0ull;
      for (int c = 0; c < 8; ++c) {
        int ci = j * 8 + c;
        if (ci < nc) {
          u = box_union(u, __ldcg(cb + ci));
          unsigned long long kc = __ldcg(ck + ci);
          k = kc > k ? kc : k;
        }
      }
      T.box[T.offset[l] + j] = u;
      T.key[T.offset[l] + j] = k;
    }
    __threadfence();
    __syncthreads();
  }
}

// inclusive AABB overlap, b2TestOverlap (include/box2d/b2_collision.h:270-276)
__device__ __forceinline__ bool aabb_overlap(float4 a, float4 b) {
  return (a.z >= b.x) & (a.x <= b.z) & (a.w >= b.y) & (a.y <= b.w);
}

// QueryCallback's filters, without the user-filter virtual call (default b2ContactFilter only)
__device__ __forceinline__ bool pair_passes(int4 a, int4 b) {
  if (a.y == b.y) return false;                          // same body
  unsigned int za = (unsigned int)a.z, zb = (unsigned int)b.z;
  if (((za | zb) & 8u) == 0) return false;               // at least one dynamic body
  if ((za | zb) & 16u) return false;                     // dead fixture
  unsigned int ta = za & 3u, tb = zb & 3u;
  if (ta == B2G_SHAPE_EDGE && tb == B2G_SHAPE_EDGE) return false;  // no edge-edge function
  short ga = (short)(za >> 16), gb = (short)(zb >> 16);
  if (ga == gb && ga != 0) return ga > 0;
  unsigned int fa = (unsigned int)a.w, fb = (unsigned int)b.w;
  unsigned int catA = fa & 0xffffu, maskA = fa >> 16, catB = fb & 0xffffu, maskB = fb >> 16;
  return (maskA & catB) != 0 && (catA & maskB) != 0;
}

// contacts are bucketed by shape-type pair so narrowphase warps stay uniform
__device__ __forceinline__ unsigned int type_bucket(unsigned int ta, unsigned int tb) {
  unsigned int lo = min(ta, tb), hi = max(ta, tb);
  // (0,0) circles=0, (0,2) polygon-circle=1, (2,2) polygons=2, (0,1) edge-circle=3, (1,2) edge-polygon=4
  if (lo == 0 && hi == 0) return 0;
  if (lo == 0 && hi == 2) return 1;
  if (lo == 2) return 2;
  if (lo == 0 && hi == 1) return 3;
  return 4;
}

// ---- contact identity: open-addressing hash table  (fixLo << 32 | fixHi) -> contact slot --------
// The reference looks an existing contact up by scanning the body's contact array
// (b2_contact_manager.cpp:145-158); here it is one hash probe.  Empty key = ~0, value -1 = tombstone.
#define B2G_HASH_EMPTY 0xffffffffffffffffull
__device__ __forceinline__ unsigned int hash_slot(unsigned long long k, unsigned int mask) {
  k ^= k >> 33;
  k *= 0xff51afd7ed558ccdull;
  k ^= k >> 33;
  k *= 0xc4ceb9fe1a85ec53ull;
  k ^= k >> 33;
  return (unsigned int)k & mask;
}
// returns the contact slot, or -1 when the pair has no live contact
__device__ __forceinline__ int hash_find(const ContactHash& H, unsigned long long key) {
  unsigned int p = hash_slot(key, H.mask);
  while (true) {
    unsigned long long k = H.keys[p];
    if (k == key) return H.vals[p];
    if (k == B2G_HASH_EMPTY) return -1;
    p = (p + 1) & H.mask;
  }
}
__device__ __forceinline__ void hash_insert(const ContactHash& H, unsigned long long key, int slot) {
  unsigned int p = hash_slot(key, H.mask);
  while (true) {
    unsigned long long k = H.keys[p];
    if (k == key) {  // tombstone of the same pair: revive
      H.vals[p] = slot;
      return;
    }
    if (k == B2G_HASH_EMPTY) {
      unsigned long long old = atomicCAS(&H.keys[p], B2G_HASH_EMPTY, key);
      if (old == B2G_HASH_EMPTY || old == key) {
        H.vals[p] = slot;
        return;
      }
    }
    p = (p + 1) & H.mask;
  }
}
__device__ __forceinline__ void hash_erase(const ContactHash& H, unsigned long long key) {
  unsigned int p = hash_slot(key, H.mask);
  while (true) {
    unsigned long long k = H.keys[p];
    if (k == key) {
      H.vals[p] = -1;
      return;
    }
    if (k == B2G_HASH_EMPTY) return;
    p = (p + 1) & H.mask;
  }
}

// b2ContactManager::QueryCallback (b2_contact_manager.cpp:136-188): an overlapping, filter-passing
// pair either persists its existing contact or is queued for creation
__device__ __forceinline__ void emit_pair(int4 a, int4 b, const ContactHash& H, uint8_t* persist,
                                          unsigned long long* newPairs, int capacity, StepCounts* counts) {
  if (!pair_passes(a, b)) return;
  if (((unsigned int)a.z & (unsigned int)b.z & 32u) && H.ncCount > 0) {
    // both bodies carry a collideConnected == false joint: is it the same joint?
    unsigned long long bk = ((unsigned long long)min(a.y, b.y) << 32) | (unsigned long long)max(a.y, b.y);
    int l = 0, r = H.ncCount - 1;
    while (l <= r) {
      int m = (l + r) >> 1;
      unsigned long long v = H.ncKeys[m];
      if (v == bk) return;
      if (v < bk) l = m + 1; else r = m - 1;
    }
  }
  unsigned long long lo = (unsigned long long)min(a.x, b.x), hi = (unsigned long long)max(a.x, b.x);
  unsigned long long key = (lo << 32) | hi;
  if (H.vetoCount > 0) {  // rejected by the user's contact filter: no contact, but remember it still overlaps
    int l = 0, r = H.vetoCount - 1;
    while (l <= r) {
      int m = (l + r) >> 1;
      unsigned long long v = H.vetoKeys[m];
      if (v == key) {
        H.vetoSeen[m] = 1;
        return;
      }
      if (v < key) l = m + 1; else r = m - 1;
    }
  }
  int slot = hash_find(H, key);
  if (slot >= 0) {
    persist[slot] = 1;
    return;
  }
  auto g = cg::coalesced_threads();
  int base = 0;
  if (g.thread_rank() == 0) base = atomicAdd(&counts->numPairs, (int)g.size());
  base = g.shfl(base, 0);
  int k = base + (int)g.thread_rank();
  if (k < capacity) newPairs[k] = key;
}

// Children of node (level l, index j) = entries [8j, 8j+8) of level l-1; all eight (box, key) pairs
// are fetched before any is tested (independent loads, one latency).
struct WideChildren {
  float4 box[B2G_BVH_W];
  unsigned long long key[B2G_BVH_W];
  int base, count;
};
__device__ __forceinline__ void wide_load(const WideBvh& T, const float4* __restrict__ leafBox,
                                          const unsigned long long* __restrict__ leafKey, int l, int j,
                                          WideChildren& ch) {
  const float4* cb = l == 1 ? leafBox : T.box + T.offset[l - 1];
  const unsigned long long* ck = l == 1 ? leafKey : T.key + T.offset[l - 1];
  ch.base = j * B2G_BVH_W;
  ch.count = min(B2G_BVH_W, T.count[l - 1] - ch.base);
#pragma unroll
  for (int c = 0; c < B2G_BVH_W; ++c) {
    bool ok = c < ch.count;
    ch.box[c] = ok ? __ldg(cb + ch.base + c) : B2G_EMPTY_BOX;
    ch.key[c] = ok ? __ldg(ck + ch.base + c) : 0ull;
  }
}
// leaves covered by entry c of level l: [c << 3l, ((c + 1) << 3l) - 1]
__device__ __forceinline__ bool wide_in_range(int l, int c, int ws, int we) {
  long long lo = (long long)c << (3 * l), hi = (((long long)c + 1) << (3 * l)) - 1;
  return lo <= we && hi >= ws;
}

// FOUR lanes per query leaf.  Leaf i reports partner j iff key(j) > key(i) (each pair once, by its
// smaller member); subtrees whose largest key is not above key(i) are pruned, and the walk never
// leaves the query's own world segment [ws, we] of the sorted order.
// The walks are latency chains and a 20 k-fixture world fills only ~4 warps per SM with one thread
// per query, so each instruction waited ~10 cycles for its predecessor.  With four lanes per query
// every lane fetches and tests two of a node's eight children, the hit masks are combined with two
// 4-lane ballots, all four lanes keep identical stacks (no synchronisation), and at the leaf level
// each lane resolves its own hits, so a node's pair lookups run in parallel: a quarter of the
// instructions per visit on four times as many warps.
// LANES = 4 for worlds below ~64 k fixtures (latency bound), 1 above (throughput bound: there the
// extra lanes only add instructions: 1024 tumbler worlds 3.47 ms/step with one lane, 3.73 with four).
template <int LANES>
__global__ void __launch_bounds__(128)
k_bp_traverse(WideBvh T, const float4* __restrict__ leafBox, const int4* __restrict__ leafInfo,
              const unsigned long long* __restrict__ leafKey, const int* __restrict__ worldFirst,
              const int* __restrict__ worldLast, const unsigned long long* __restrict__ keysSorted, int numWorlds,
              ContactHash H, uint8_t* persist, unsigned long long* newPairs, int capacity, StepCounts* counts) {
  B2G_PDL_ENTER();
  constexpr int PER = B2G_BVH_W / LANES;  // children per lane
  const int n = T.count[0];
  auto g = cg::tiled_partition<LANES>(cg::this_thread_block());
  const int r = (int)g.thread_rank();
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) / LANES;  // uniform inside the tile
  if (i >= n || n < 2) return;
  int4 me = leafInfo[i];
  if ((unsigned int)me.z & 16u) return;
  float4 qbox = leafBox[i];
  const unsigned long long myKey = leafKey[i];
  int ws = 0, we = n - 1;
  if (numWorlds > 1) {
    unsigned int w = (unsigned int)(keysSorted[i] >> 32);
    ws = worldFirst[w];
    we = worldLast[w];
  }
  int stack[72];  // (level << 27) | index; at most 7 pending siblings per level; identical in all lanes
  int sp = 0;
  int visits = 0;
  stack[sp++] = (T.levels << 27);
  while (sp > 0) {
    const int e = stack[--sp];
    const int l = e >> 27, j = e & 0x7ffffff;
    ++visits;
    const float4* cb = l == 1 ? leafBox : T.box + T.offset[l - 1];
    const unsigned long long* ck = l == 1 ? leafKey : T.key + T.offset[l - 1];
    const int base = j * B2G_BVH_W;
    const int cnt = T.count[l - 1] - base;  // children that exist (>= 1)
    // this lane's children: r, r + LANES, ...  All loads are issued before any test.
    float4 bx[PER];
    unsigned long long ky[PER];
#pragma unroll
    for (int m = 0; m < PER; ++m) {
      const int c = r + LANES * m;
      const bool in = c < cnt;
      bx[m] = in ? __ldg(cb + base + c) : B2G_EMPTY_BOX;
      ky[m] = in ? __ldg(ck + base + c) : 0ull;
    }
    unsigned int own = 0;  // bit m: this lane's m-th child survives
#pragma unroll
    for (int m = 0; m < PER; ++m)
      if (ky[m] > myKey && aabb_overlap(qbox, bx[m])) own |= 1u << m;
    if (l == 1) {
      // leaves: each lane resolves its own hits (with four lanes a node's pair lookups run side by
      // side); one call site, emit_pair is large
      while (own) {
        const int ci = base + r + LANES * (__ffs(own) - 1);
        own &= own - 1;
        if (ci >= ws && ci <= we) emit_pair(me, leafInfo[ci], H, persist, newPairs, capacity, counts);
      }
    } else {
      unsigned int hits = 0;  // bit c: child c survives (the same value in every lane)
      if (LANES == 1) {
        hits = own;
      } else {
#pragma unroll
        for (int m = 0; m < PER; ++m) hits |= g.ballot((own >> m) & 1u) << (LANES * m);
      }
      while (hits) {
        const int ci = base + __ffs(hits) - 1;
        hits &= hits - 1;
        if (wide_in_range(l - 1, ci, ws, we)) stack[sp++] = ((l - 1) << 27) | ci;
      }
    }
  }
  // tree-quality counters (one atomic per warp, one lane per query counts)
  {
    auto cgp = cg::coalesced_threads();
    int mine = r == 0 ? visits : 0;
    int total = cg::reduce(cgp, mine, cg::plus<int>());
    int most = cg::reduce(cgp, mine, cg::greater<int>());
    if (cgp.thread_rank() == 0) {
      atomicAdd(&counts->bpVisits, (unsigned long long)total);
      atomicMax(&counts->bpMaxVisits, most);
    }
  }
}

// ---- contact list maintenance -------------------------------------------------------------------
// Contacts live in STABLE slots (no per-step sort or copy).  After the traversal has flagged every
// still-overlapping pair, k_contact_sweep retires the rest and k_contact_insert creates the new
// ones in free slots.  Slot numbers are not deterministic (free-list order), and nothing depends
// on them: islands, colours and the solver are functions of bodies and pair keys only.

// contacts of pairs the user's filter has just rejected: they were inserted by the refresh that
// found the pair and are taken out again before any narrowphase saw them (no wake-up, no event —
// in the reference such a contact is never created)
__global__ void k_contacts_remove(int n, const unsigned long long* __restrict__ keys, ContactBuf C, ContactHash H,
                                  int* freeStack, int* freeTop, int* removed) {
  B2G_PDL_ENTER();
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  unsigned long long key = keys[i];
  int slot = hash_find(H, key);
  if (slot < 0) return;
  C.flags[slot] = 0;
  C.colour[slot] = -1;
  hash_erase(H, key);
  int t = atomicAdd(freeTop, 1);
  freeStack[t] = slot;
  atomicAdd(removed, 1);
}

// contacts that were not re-reported die: b2ContactManager::Destroy (b2_contact_manager.cpp:48-61)
// fires EndContact if touching, b2Contact::Destroy (b2_contact.cpp:79-94) wakes both bodies if the
// manifold had points
__global__ void k_contact_sweep(int nSlots, ContactBuf C, uint8_t* persist, ContactHash H,
                                const uint32_t* __restrict__ fTypeFlags, uint32_t* bflags, float4* force,
                                int* freeStack, int* freeTop, StepCounts* counts, int recordEvents, int2* endEvents,
                                int eventCap, const int* __restrict__ island, uint8_t* islandDirty) {
  B2G_PDL_ENTER();
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= nSlots) return;
  uint32_t flags = C.flags[j];
  if (!(flags & B2G_CONTACT_ALIVE)) return;
  if (persist[j]) {
    persist[j] = 0;
    return;
  }
  int2 fx = C.fix[j];
  if (recordEvents && (flags & B2G_CONTACT_TOUCHING)) {
    int k = atomicAdd(&counts->endCount, 1);
    if (k < eventCap) endEvents[k] = fx;
  }
  if (flags & B2G_CONTACT_TOUCHING) {
    // a touching contact dies = an island edge disappears (see k_narrowphase)
    int2 bd0 = C.body[j];
    islandDirty[island[B2G_BODY_TYPE(bflags[bd0.x]) != B2G_STATIC ? bd0.x : bd0.y]] = 1;
  }
  int pointCount = __float_as_int(C.m3[j].w);
  if (pointCount > 0 && !((fTypeFlags[fx.x] | fTypeFlags[fx.y]) & B2G_FIX_SENSOR)) {
    // SetAwake(true) takes effect immediately (the next Collide must see the body awake);
    // concurrent writers all store the same values
    int2 bd = C.body[j];
    if (B2G_BODY_TYPE(bflags[bd.x]) != B2G_STATIC) {
      atomicOr(&bflags[bd.x], B2G_BODY_AWAKE);
      force[bd.x].w = 0.0f;
    }
    if (B2G_BODY_TYPE(bflags[bd.y]) != B2G_STATIC) {
      atomicOr(&bflags[bd.y], B2G_BODY_AWAKE);
      force[bd.y].w = 0.0f;
    }
  }
  C.flags[j] = 0;
  C.colour[j] = -1;
  hash_erase(H, C.key[j]);
  int t = atomicAdd(freeTop, 1);
  freeStack[t] = j;
  atomicAdd(&counts->numDead, 1);
}

// new contacts (b2Contact::Create + ctor, b2_contact.cpp:58-122): slot from the free stack, else
// appended past the high-water mark
__global__ void __launch_bounds__(256)
k_contact_insert(int nNew, const unsigned long long* __restrict__ newPairs, int freeTopBefore, int highWater,
                 ContactBuf C, uint8_t* persist, ContactHash H, const int* __restrict__ freeStack, int* freeTop,
                 const int* __restrict__ fBody, const uint32_t* __restrict__ fTypeFlags,
                 const float4* __restrict__ fMaterial) {
  B2G_PDL_ENTER();
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) *freeTop = freeTopBefore > nNew ? freeTopBefore - nNew : 0;
  if (i >= nNew) return;
  unsigned long long key = newPairs[i];
  int slot = i < freeTopBefore ? freeStack[freeTopBefore - 1 - i] : highWater + (i - freeTopBefore);
  int lo = (int)(key >> 32), hi = (int)(key & 0xffffffffull);
  unsigned int tl = fTypeFlags[lo] & 3u, th = fTypeFlags[hi] & 3u;
  // A/B order from the function table: [circle][circle] [polygon][circle] [polygon][polygon]
  // [edge][circle] [edge][polygon]; same-type pairs keep the lower fixture index as A
  bool loFirst = (tl == th) || (tl == B2G_SHAPE_POLYGON && th == B2G_SHAPE_CIRCLE) ||
                 (tl == B2G_SHAPE_EDGE && th == B2G_SHAPE_CIRCLE) || (tl == B2G_SHAPE_EDGE && th == B2G_SHAPE_POLYGON);
  int fa = loFirst ? lo : hi, fb = loFirst ? hi : lo;
  C.key[slot] = key;
  C.fix[slot] = make_int2(fa, fb);
  C.body[slot] = make_int2(fBody[fa], fBody[fb]);
  C.flags[slot] = B2G_CONTACT_ENABLED | B2G_CONTACT_ALIVE;
  float4 ma = fMaterial[fa], mb = fMaterial[fb];
  // mixing laws, include/box2d/b2_contact.h:42-60
  float friction = sqrtf(ma.x * mb.x);
  float restitution = ma.y > mb.y ? ma.y : mb.y;
  float threshold = ma.z < mb.z ? ma.z : mb.z;
  C.material[slot] = make_float4(friction, restitution, threshold, 0.0f);
  float4 z = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
  C.m0[slot] = z;
  C.m1[slot] = z;
  C.m2[slot] = z;
  C.m3[slot] = z;
  C.colour[slot] = -1;
  persist[slot] = 0;
  hash_insert(H, key, slot);
}

// rebuild the table from the live contacts (drops tombstones)
__global__ void k_hash_rebuild(int nSlots, ContactBuf C, ContactHash H) {
  B2G_PDL_ENTER();
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= nSlots) return;
  if (C.flags[j] & B2G_CONTACT_ALIVE) hash_insert(H, C.key[j], j);
}

// b2g_fused.cuh — the fused per-island solver of the production (graph-coloured) mode.
//
// b2Island::Solve (src/dynamics/b2_island.cpp:253-484) handles one island at a time on one CPU
// thread.  Here islands are packed into *bins* (consecutive islands, in root order, until a bin
// holds ~binSize bodies) and ONE thread block solves one bin from start to finish:
//   velocity integration -> constraint preparation -> warm start -> velocityIterations x colours
//   -> store impulses -> position integration -> positionIterations x colours with the per-island
//   early exit -> write-back, SynchronizeTransform, sleep bookkeeping, island sleep, ClearForces,
// with the bin's body velocities and positions resident in SHARED MEMORY for the whole solve and
// __syncthreads() between colours instead of kernel relaunches.  Constraints of different islands
// never share a movable body, so the global colouring stays valid inside a bin.
//
// Islands larger than `bigThreshold` bodies do not fit a tile; they are routed to the "big" bin
// and solved by the per-colour kernels of b2g_step_kernels.cuh over the whole GPU.
#pragma once
#include "b2g_step_kernels.cuh"
#include <cooperative_groups/scan.h>
#include <cooperative_groups/reduce.h>

#define B2G_FUSED_THREADS 256
#define B2G_COLOUR_BITS 5  // 24 colours + overflow < 32
#define B2G_SLOT_NONE (-1)
#define B2G_SLOT_BIG (-2)

// Slot range of every tile-sized island: the island's root claims `count` consecutive slots from a
// global cursor.  The order in which islands claim (hence which islands share a bin) varies from
// run to run, but islands are independent and every per-island reduction is order-free, so the
// simulation result does not depend on it.
__global__ void k_island_alloc(int nb, const int* __restrict__ island, const uint32_t* __restrict__ islandAwake,
                               const int* __restrict__ islandCount, int* islandStart, int* binFirst, int* binEnd,
                               int binSize, int bigThreshold, StepCounts* counts, uint8_t* islandWasBig) {
  B2G_PDL_ENTER();
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nb) return;
  int cnt = islandCount[b];
  islandWasBig[b] = (island[b] == b && cnt > bigThreshold) ? 1 : 0;
  if (cnt <= 0 || cnt > bigThreshold || island[b] != b || !islandAwake[b]) return;
  // one atomic per warp: a scene of many small islands would otherwise hammer the single cursor
  auto g = cg::coalesced_threads();
  int prefix = cg::exclusive_scan(g, cnt, cg::plus<int>());
  int total = cg::reduce(g, cnt, cg::plus<int>());
  int base = 0;
  if (g.thread_rank() == 0) base = atomicAdd(&counts->slotCursor, total);
  int start = g.shfl(base, 0) + prefix;
  islandStart[b] = start;
  int bin = start / binSize;
  atomicMin(&binFirst[bin], start);
  atomicMax(&binEnd[bin], start + cnt);
}

// slot of every simulated body inside its island's range
__device__ __forceinline__ void body_scatter_one(int b, const uint32_t* __restrict__ bflags, const int* __restrict__ island,
                                                 const uint32_t* __restrict__ islandAwake,
                                                 const int* __restrict__ islandCount, const int* __restrict__ islandStart,
                                                 int* islandCursor, int* bodySlot, int* slotBody, int bigThreshold,
                                                 StepCounts* counts) {
  if (!body_simulated(bflags[b], island, islandAwake, b)) {
    bodySlot[b] = B2G_SLOT_NONE;
    return;
  }
  int root = island[b];
  if (islandCount[root] > bigThreshold) {
    bodySlot[b] = B2G_SLOT_BIG;
    auto g = cg::coalesced_threads();
    if (g.thread_rank() == 0) atomicAdd(&counts->numBigBodies, (int)g.size());
    return;
  }
  int slot = islandStart[root] + atomicAdd(&islandCursor[root], 1);
  bodySlot[b] = slot;
  slotBody[slot] = b;
}
__global__ void k_body_scatter(int nb, const uint32_t* __restrict__ bflags, const int* __restrict__ island,
                               const uint32_t* __restrict__ islandAwake, const int* __restrict__ islandCount,
                               const int* __restrict__ islandStart, int* islandCursor, int* bodySlot, int* slotBody,
                               int bigThreshold, StepCounts* counts) {
  B2G_PDL_ENTER();
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nb) return;
  body_scatter_one(b, bflags, island, islandAwake, islandCount, islandStart, islandCursor, bodySlot, slotBody, bigThreshold,
                   counts);
}

// ---- colouring -------------------------------------------------------------------------------------
// Colours persist on the contact from step to step; what has to be coloured in a step is only what was
// born (or lost its colour) since the last one.  k_mark_active_bins classifies every contact slot, re-
// publishes the persisting colours and appends the uncoloured active constraints to a WORKLIST;
// k_colour_worklist then runs every Luby round of the step over that list inside one launch.
//
// Two colour domains share the 64-bit per-body mask: bits 0..23 for constraints solved inside one CTA
// (island bins, tiles of an oversize island), bits 32..55 for the CUT constraints of a tiled oversize
// island, which are solved in separate grid-wide passes and therefore only have to be conflict-free
// among themselves.  The stored colour is the bit index; its low five bits select the bucket.
#define B2G_CUT_DOMAIN_SHIFT 32
__device__ __forceinline__ unsigned long long colour_domain_mask(int domain) {
  return ((1ull << B2G_MAX_COLOURS) - 1ull) << (domain ? B2G_CUT_DOMAIN_SHIFT : 0);
}
// Pair key with WORLD-LOCAL fixture indices: world k of a batched arena then draws the same priorities,
// hence the same colours and the same floats, as that world stepped alone.
__device__ __forceinline__ unsigned long long local_pair_key(unsigned long long key, const int* __restrict__ bodyFixBase,
                                                            int body) {
  if (!bodyFixBase) return key;
  unsigned long long base = (unsigned long long)(unsigned int)bodyFixBase[body];
  return key - ((base << 32) | base);
}
__global__ void k_world_fix_min(int nf, const int* __restrict__ fBody, const int* __restrict__ bworld, int* worldFixMin) {
  B2G_PDL_ENTER();
  int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= nf) return;
  atomicMin(&worldFixMin[bworld[fBody[f]]], f);
}
__global__ void k_body_fix_base(int nb, const int* __restrict__ bworld, const int* __restrict__ worldFixMin,
                                int* bodyFixBase) {
  B2G_PDL_ENTER();
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nb) return;
  int m = worldFixMin[bworld[b]];
  bodyFixBase[b] = m == 0x7f7f7f7f ? 0 : m;
}

struct MarkArgs {
  int nc, nb;
  int dropColours, binSize, bigThreshold, bigBin;
  int tileBin0;   // bin of tile 0 of a tiled oversize island (tile t -> tileBin0 + t), or -1: not tiled
  int cutBin;     // bin of the cut constraints of the tiled island (the second colour domain)
  int tileCap;
};

// Also carries two neighbours that depend on the same inputs and on nothing else, to save their
// launches: the body scatter (thread i handles body i) and round 0 of the colouring proposals.
__global__ void k_mark_active_bins(MarkArgs M, ContactBuf C, const uint32_t* __restrict__ fTypeFlags,
                                   const uint32_t* __restrict__ bflags, const int* __restrict__ island,
                                   const uint32_t* __restrict__ islandAwake, const int* __restrict__ islandCount,
                                   const int* __restrict__ islandStart, int* cbin, StepCounts* counts,
                                   const float4* __restrict__ mass, unsigned long long* colourMask, int* islandCursor,
                                   int* bodySlot, int* slotBody, unsigned long long* bodyBest,
                                   const int* __restrict__ bodyFixBase, int* worklist, int* bucketCount, int* rank,
                                   const int* __restrict__ tileSlot, uint8_t* tileBoundary,
                                   const uint8_t* __restrict__ worldRecolour, const int* __restrict__ bworld) {
  B2G_PDL_ENTER();
  const int nc = M.nc, nb = M.nb, bigThreshold = M.bigThreshold, bigBin = M.bigBin;
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nb) body_scatter_one(i, bflags, island, islandAwake, islandCount, islandStart, islandCursor, bodySlot, slotBody,
                               bigThreshold, counts);
  if (i >= nc) return;
  // loads first, tests afterwards: three rounds of independent loads instead of a six-deep chain
  // (dead slots keep in-range indices)
  uint32_t flags = C.flags[i];
  int2 bd = C.body[i];
  int2 fx = C.fix[i];
  int c = C.colour[i];
  uint32_t tfa = fTypeFlags[fx.x], tfb = fTypeFlags[fx.y];
  uint32_t bfa = bflags[bd.x], bfb = bflags[bd.y];
  int ia = island[bd.x], ib = island[bd.y];
  float4 massA = mass[bd.x], massB = mass[bd.y];
  int root = B2G_BODY_TYPE(bfa) != B2G_STATIC ? ia : ib;
  uint32_t awake = islandAwake[root];
  int icount = islandCount[root], istart = islandStart[root];
  bool active = (flags & (B2G_CONTACT_TOUCHING | B2G_CONTACT_ENABLED)) == (B2G_CONTACT_TOUCHING | B2G_CONTACT_ENABLED);
  if ((tfa | tfb) & B2G_FIX_SENSOR) active = false;
  if (active) active = awake != 0;
  int bin = -1;
  if (active) {
    if (icount > bigThreshold) {
      bin = bigBin;
      if (M.tileBin0 >= 0) {
        // tiled oversize island: interior to a tile when every non-static body of the constraint sits in
        // that tile; anything else (two tiles, a body that did not fit its tile) is a cut constraint
        int ta = B2G_BODY_TYPE(bfa) != B2G_STATIC ? tileSlot[bd.x] : -2;
        int tb = B2G_BODY_TYPE(bfb) != B2G_STATIC ? tileSlot[bd.y] : -2;
        int tA = ta >= 0 ? ta / M.tileCap : ta, tB = tb >= 0 ? tb / M.tileCap : tb;
        if (tA >= 0 && (tB == tA || tB == -2)) bin = M.tileBin0 + tA;
        else if (tA == -2 && tB >= 0) bin = M.tileBin0 + tB;
        else {
          bin = M.cutBin;
          if (ta >= 0) tileBoundary[ta] = 1;
          if (tb >= 0) tileBoundary[tb] = 1;
        }
      }
      atomicAdd(&counts->numBig, 1);
    } else {
      bin = istart / M.binSize;
    }
    auto g = cg::coalesced_threads();
    if (g.thread_rank() == 0) atomicAdd(&counts->numActive, (int)g.size());
  }
  cbin[i] = bin;
  const int domain = (active && bin == M.cutBin && M.cutBin >= 0) ? 1 : 0;
  const bool edited = worldRecolour != nullptr && worldRecolour[bworld[bd.x]] != 0;  // (both bodies: same world)
  if (!active || M.dropColours || edited || (c & 31) >= B2G_MAX_COLOURS || (c >> 5) != domain) {
    // inactive contacts lose their colour; overflow constraints retry every step; a constraint that
    // changed sides (tile interior <-> cut) is coloured again in its new domain
    if (c != -1) C.colour[i] = -1;
    c = -1;
  }
  // colours persist from step to step: publish the ones still in use, and count them into their bucket
  if (c >= 0) {
    unsigned long long bit = 1ull << c;
    if (body_movable(massA)) atomicOr(&colourMask[bd.x], bit);
    if (body_movable(massB)) atomicOr(&colourMask[bd.y], bit);
    if (bin == bigBin) atomicAdd(&counts->colourCount[c & 31], 1);
    if ((c & 31) + 1 > counts->numColours) atomicMax(&counts->numColours, (c & 31) + 1);
    rank[i] = atomicAdd(&bucketCount[(bin << B2G_COLOUR_BITS) | (c & 31)], 1);
  } else if (active) {
    // to the worklist + round 0 of the proposals
    {
      auto g = cg::coalesced_threads();
      int base = 0;
      if (g.thread_rank() == 0) base = atomicAdd(&counts->worklistCount, (int)g.size());
      worklist[g.shfl(base, 0) + g.thread_rank()] = i;
    }
    unsigned long long pr = colour_priority(0, i, local_pair_key(C.key[i], bodyFixBase, bd.x));
    if (body_movable(massA)) atomicMax(&bodyBest[bd.x], pr);
    if (body_movable(massB)) atomicMax(&bodyBest[bd.y], pr);
  }
}

// Grid barrier of the persistent kernels (split into arrive / wait further down); forward declarations
__device__ __forceinline__ void grid_arrive(unsigned int* counter, unsigned int& target);
__device__ __forceinline__ void grid_wait(unsigned int* counter, unsigned int target);

// Every colouring round of the step in ONE launch.  Deterministic Luby rounds: an uncoloured constraint
// that holds the highest priority on both of its movable bodies takes the lowest colour free on both (of
// its domain).  Priorities carry the round number in their top bits, so bodyBest never has to be cleared.
// A short worklist (the steady state: a few hundred births per step) is handled by block 0 alone with CTA
// barriers; a long one (first step, an avalanche) by the whole co-resident grid with grid barriers.
// The last pass counts the newly coloured constraints into their (bin, colour) buckets.
#define B2G_WL_THREADS 1024
#define B2G_WL_SINGLE_MAX 2048
#define B2G_WL_MAX_ROUNDS 200
#define B2G_CUT_SECTORS 6  // direction classes a cut constraint's preferred colour is drawn from
__global__ void __launch_bounds__(B2G_WL_THREADS)
k_colour_worklist(ContactBuf C, const int* __restrict__ cbin, const float4* __restrict__ mass,
                  unsigned long long* colourMask, unsigned long long* bodyBest, const int* __restrict__ bodyFixBase,
                  const int* __restrict__ worklist, StepCounts* counts, int bigBin, int cutBin, int* bucketCount,
                  int* rank, unsigned int* barrier, int singleMax, const float4* __restrict__ pos, int cutSectors) {
  const int n = __ldcg(&counts->worklistCount);
  if (n == 0) return;
  const bool single = n <= singleMax;
  if (single && blockIdx.x != 0) return;
  const int stride = single ? blockDim.x : gridDim.x * blockDim.x;
  const int t0 = single ? threadIdx.x : blockIdx.x * blockDim.x + threadIdx.x;
  unsigned int target = 0;
  auto sync_all = [&]() {
    if (single) {
      __syncthreads();
    } else {
      grid_arrive(barrier, target);
      grid_wait(barrier, target);
    }
  };
  // A thread's worklist entries (k = t0, t0 + stride, ...) and everything about them that does not change
  // during the launch live in registers: a round is then one look at the bodies' proposals, not a chain of
  // dependent loads.  (Two entries per thread cover 2 048 constraints in single-CTA mode and ~300 k in grid
  // mode; a longer list is still handled, from memory.)
  struct Entry {
    int i, a, b, bin;
    int pref;  // preferred colour bit (cut domain), -1 = none
    unsigned long long k56;
    bool movA, movB, open;  // open = still uncoloured
  };
  auto fetch = [&](int k) {
    Entry e;
    e.i = worklist[k];
    const int2 bd = C.body[e.i];
    e.a = bd.x;
    e.b = bd.y;
    e.bin = cbin[e.i];
    e.k56 = colour_key56(local_pair_key(C.key[e.i], bodyFixBase, bd.x));
    e.movA = body_movable(mass[bd.x]);
    e.movB = body_movable(mass[bd.y]);
    e.open = true;
    e.pref = -1;
    if (e.bin == cutBin && cutBin >= 0) {
      // Cut constraints are solved in turn order = colour order, and what costs is the DEPTH of that order (one
      // L2 round trip per turn).  First fit on random priorities needs up to 2 deg - 1 colours; the geometry
      // knows better: across a straight seam a body meets at most one neighbour per ~30 degrees of direction,
      // so the undirected direction of the contact (6 sectors of 30 degrees) is tried first and the lowest free
      // colour only when that one is taken.  Bodies on a seam then carry ~3 cut colours, at corners ~6.
      const float4 pa = pos[bd.x], pb = pos[bd.y];
      float dx = pb.x - pa.x, dy = pb.y - pa.y;
      if (dy < 0.0f || (dy == 0.0f && dx < 0.0f)) {
        dx = -dx;
        dy = -dy;
      }
      int sector = (int)(atan2f(dy, dx) * ((float)cutSectors / 3.14159265f));  // [0, pi) -> 0 .. cutSectors - 1
      sector = sector < 0 ? 0 : (sector > cutSectors - 1 ? cutSectors - 1 : sector);
      e.pref = B2G_CUT_DOMAIN_SHIFT + sector;
    }
    return e;
  };
  constexpr int CACHED = 2;
  Entry cache[CACHED];
#pragma unroll
  for (int j = 0; j < CACHED; ++j) {
    cache[j].open = false;
    if (t0 + j * stride < n) cache[j] = fetch(t0 + j * stride);
  }
  const int kRest = t0 + CACHED * stride;  // first entry of this thread that is not cached
  // A constraint that finds no free colour on its bodies goes to its bin's serial bucket without waiting for
  // its turn to win (a hub body — the tumbler's container touches ~120 boxes — would otherwise cost one round
  // per contact).  Masks only grow, so "none free" is final; the test is made only while the masks are
  // stable (before round 0, and in the propose phases — never next to a commit), which keeps the colouring a
  // pure function of the constraint set.
  auto full_mask = [&](Entry& e) {
    const int domain = (e.bin == cutBin && cutBin >= 0) ? 1 : 0;
    const unsigned long long used = (e.movA ? __ldcg(&colourMask[e.a]) : 0ull) | (e.movB ? __ldcg(&colourMask[e.b]) : 0ull);
    if (~used & colour_domain_mask(domain)) return false;
    const int c = B2G_OVERFLOW_COLOUR + (domain ? B2G_CUT_DOMAIN_SHIFT : 0);
    __stcg(&C.colour[e.i], c);
    e.open = false;
    atomicAdd(&counts->numOverflow, 1);
    if (e.bin == bigBin) atomicAdd(&counts->colourCount[c & 31], 1);
    return true;
  };
  // commit: the winner of this round's proposals takes the lowest colour free on both bodies
  auto commit = [&](Entry& e, int round) {
    const unsigned long long pr = ((unsigned long long)(round + 1) << 56) | e.k56;
    const bool win = (!e.movA || __ldcg(&bodyBest[e.a]) == pr) && (!e.movB || __ldcg(&bodyBest[e.b]) == pr);
    if (!win) return false;
    // the winner is unique on each of its movable bodies: nobody else touches their masks this round
    const int domain = (e.bin == cutBin && cutBin >= 0) ? 1 : 0;
    const unsigned long long ma = e.movA ? __ldcg(&colourMask[e.a]) : 0ull, mb = e.movB ? __ldcg(&colourMask[e.b]) : 0ull;
    const unsigned long long freeBits = ~(ma | mb) & colour_domain_mask(domain);
    int c = __ffsll((long long)freeBits) - 1;  // never empty: checked while the masks were stable
    if (e.pref >= 0 && ((freeBits >> e.pref) & 1ull)) c = e.pref;
    const unsigned long long bit = 1ull << c;
    if (e.movA) __stcg(&colourMask[e.a], ma | bit);
    if (e.movB) __stcg(&colourMask[e.b], mb | bit);
    if ((c & 31) + 1 > counts->numColours) atomicMax(&counts->numColours, (c & 31) + 1);
    __stcg(&C.colour[e.i], c);
    e.open = false;
    if (e.bin == bigBin) atomicAdd(&counts->colourCount[c & 31], 1);
    return true;
  };
  auto propose = [&](Entry& e, int round) {
    if (full_mask(e)) return;
    const unsigned long long pr = ((unsigned long long)(round + 1) << 56) | e.k56;
    if (e.movA) atomicMax(&bodyBest[e.a], pr);
    if (e.movB) atomicMax(&bodyBest[e.b], pr);
  };
#pragma unroll
  for (int j = 0; j < CACHED; ++j)
    if (cache[j].open) full_mask(cache[j]);
  for (int k = kRest; k < n; k += stride) {
    Entry e = fetch(k);
    full_mask(e);
  }
  sync_all();
  int round = 0;
  for (;; ++round) {
    int left = 0;
#pragma unroll
    for (int j = 0; j < CACHED; ++j)
      if (cache[j].open && !commit(cache[j], round)) left = 1;
    for (int k = kRest; k < n; k += stride) {
      if (__ldcg(&C.colour[worklist[k]]) >= 0) continue;
      Entry e = fetch(k);
      if (!commit(e, round)) left = 1;
    }
    if (round + 1 >= B2G_WL_MAX_ROUNDS) break;
    // anything left anywhere?
    int any;
    if (single) {
      any = __syncthreads_or(left);
    } else {
      if (left) atomicOr(&counts->worklistLeft[round], 1);
      sync_all();
      any = __ldcg(&counts->worklistLeft[round]);
    }
    if (!any) break;
    // propose for the next round (the masks are stable here)
#pragma unroll
    for (int j = 0; j < CACHED; ++j)
      if (cache[j].open) propose(cache[j], round + 1);
    for (int k = kRest; k < n; k += stride) {
      if (__ldcg(&C.colour[worklist[k]]) >= 0) continue;
      Entry e = fetch(k);
      propose(e, round + 1);
    }
    sync_all();
  }
  sync_all();
  // bucket counting of what this launch coloured (constraints left uncoloured after the round limit go to
  // their bin's serial bucket for this step and retry in the next one)
  int leftover = 0;
  for (int k = t0; k < n; k += stride) {
    const int i = worklist[k];
    int c = __ldcg(&C.colour[i]);
    if (c < 0) {
      c = B2G_OVERFLOW_COLOUR;
      ++leftover;
    }
    rank[i] = atomicAdd(&bucketCount[(cbin[i] << B2G_COLOUR_BITS) | (c & 31)], 1);
  }
  if (leftover) atomicAdd(&counts->remaining, leftover);
  if (t0 == 0) counts->lastUsefulRound = round + 1;
}

// exclusive scan of the bucket counts by one block (buckets = bins x 32: ~10 k entries for one
// world, ~500 k for 1024 batched worlds).  Tiles of 8192: every thread scans 8 consecutive counts in
// registers (two 16-byte loads), one block-wide scan of the 1024 thread sums per tile.
__global__ void __launch_bounds__(1024) k_bucket_scan(int n, const int* __restrict__ bucketCount, int* bucketStart) {
  B2G_PDL_ENTER();
  __shared__ int warpSums[32];
  __shared__ int carryS;
  if (threadIdx.x == 0) carryS = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int base = 0; base < n; base += 8192) {
    const int i0 = base + threadIdx.x * 8;
    int v[8];
    if (i0 + 8 <= n) {
      int4 a = *reinterpret_cast<const int4*>(bucketCount + i0);
      int4 b = *reinterpret_cast<const int4*>(bucketCount + i0 + 4);
      v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else {
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = i0 + k < n ? bucketCount[i0 + k] : 0;
    }
    int sum = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      int t = v[k];
      v[k] = sum;  // exclusive within the thread
      sum += t;
    }
    int x = sum;
    for (int o = 1; o < 32; o <<= 1) {
      int y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) warpSums[wid] = x;
    __syncthreads();
    if (wid == 0) {
      int w = warpSums[lane];
      for (int o = 1; o < 32; o <<= 1) {
        int y = __shfl_up_sync(0xffffffffu, w, o);
        if (lane >= o) w += y;
      }
      warpSums[lane] = w;
    }
    __syncthreads();
    const int carry = carryS;
    const int prefix = carry + (wid > 0 ? warpSums[wid - 1] : 0) + x - sum;  // exclusive prefix of this thread
    if (i0 + 8 <= n) {
      *reinterpret_cast<int4*>(bucketStart + i0) = make_int4(prefix + v[0], prefix + v[1], prefix + v[2], prefix + v[3]);
      *reinterpret_cast<int4*>(bucketStart + i0 + 4) = make_int4(prefix + v[4], prefix + v[5], prefix + v[6], prefix + v[7]);
    } else {
#pragma unroll
      for (int k = 0; k < 8; ++k)
        if (i0 + k < n) bucketStart[i0 + k] = prefix + v[k];
    }
    __syncthreads();
    if (threadIdx.x == 1023) carryS = prefix + sum;
    __syncthreads();
  }
  if (threadIdx.x == 0) bucketStart[n] = carryS;
}

__global__ void k_bucket_scatter(int nc, const int* __restrict__ cbin, ContactBuf C, const int* __restrict__ bucketStart,
                                 const int* __restrict__ rank, int* sortedList) {
  B2G_PDL_ENTER();
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nc) return;
  int bin = cbin[i];
  if (bin < 0) return;
  int c = C.colour[i];
  if (c < 0) c = B2G_OVERFLOW_COLOUR;
  sortedList[bucketStart[(bin << B2G_COLOUR_BITS) | (c & 31)] + rank[i]] = i;
}

// The overflow bucket is the only place where constraint ORDER matters (one thread, Gauss-Seidel in
// list order), and the counting sort above leaves bucket contents in arbitrary order: rank-sort it
// by pair key so runs stay bit-reproducible.  Block-wide, O(n^2) key compares, n = overflow size.
// block-wide ascending sort of n ints in place (bitonic network, "flip" form: see order_bucket_by_key)
__device__ __forceinline__ void sort_ints_ascending(int* a, int n) {
  if (n <= 1) return;  // block-uniform
  int m = 1;
  while (m < n) m <<= 1;
  auto exchange = [&](int i, int l) {
    if (l < n) {
      const int x = a[i], y = a[l];
      if (x > y) {
        a[i] = y;
        a[l] = x;
      }
    }
  };
  for (int k = 2; k <= m; k <<= 1) {
    const int hk = k >> 1;
    for (int t = threadIdx.x; t < (m >> 1); t += blockDim.x) {
      const int i = (t / hk) * k + (t % hk);
      exchange(i, i ^ (k - 1));
    }
    __syncthreads();
    for (int j = k >> 2; j > 0; j >>= 1) {
      for (int t = threadIdx.x; t < (m >> 1); t += blockDim.x) {
        const int i = (t / j) * 2 * j + (t % j);
        exchange(i, i + j);
      }
      __syncthreads();
    }
  }
}

__device__ __forceinline__ void order_bucket_by_key(int o0, int o1, int* sortedList, int* scratch, const ContactBuf& C) {
  const int n = o1 - o0;
  if (n <= 1) return;  // block-uniform
  if (n <= 96) {
    // the usual case (a hub body's few dozen extra contacts): rank by counting, O(n^2) compares on cached keys
    for (int t = threadIdx.x; t < n; t += blockDim.x) {
      int i = sortedList[o0 + t];
      unsigned long long ki = C.key[i];
      int rank = 0;
      for (int j = 0; j < n; ++j) rank += C.key[sortedList[o0 + j]] < ki ? 1 : 0;
      scratch[o0 + rank] = i;
    }
    __syncthreads();
    for (int t = threadIdx.x; t < n; t += blockDim.x) sortedList[o0 + t] = scratch[o0 + t];
    __syncthreads();
    return;
  }
  // A large bucket (bodies that start overlapping hundreds of others: the reference's "n^2" benchmarks put tens of
  // thousands of constraints here) would cost n^2 = billions of compares: bitonic network in place instead,
  // O(n log^2 n).  Every compare-exchange puts the smaller key at the lower index (the "flip" form of the
  // network), so the virtual +inf padding up to the next power of two never has to move and is simply skipped.
  // Keys are unique, so the result is the same order the rank sort gives.
  int* a = sortedList + o0;
  int m = 1;
  while (m < n) m <<= 1;
  auto exchange = [&](int i, int l) {
    if (l < n) {
      const int ia = a[i], il = a[l];
      if (C.key[ia] > C.key[il]) {
        a[i] = il;
        a[l] = ia;
      }
    }
  };
  for (int k = 2; k <= m; k <<= 1) {
    const int hk = k >> 1;
    for (int t = threadIdx.x; t < (m >> 1); t += blockDim.x) {
      const int i = (t / hk) * k + (t % hk);
      exchange(i, i ^ (k - 1));
    }
    __syncthreads();
    for (int j = k >> 2; j > 0; j >>= 1) {
      for (int t = threadIdx.x; t < (m >> 1); t += blockDim.x) {
        const int i = (t / j) * 2 * j + (t % j);
        exchange(i, i + j);
      }
      __syncthreads();
    }
  }
}

__global__ void k_order_overflow(int o0, int o1, int* sortedList, int* scratch, ContactBuf C) {
  B2G_PDL_ENTER();
  order_bucket_by_key(o0, o1, sortedList, scratch, C);
}

struct FusedParams {
  int nc;              // contacts (length of the sorted key / index arrays)
  int binSize;
  float h, dtRatio;
  float2 gravity;
  int velIters, posIters, warmStarting, allowSleep, clearForces;
  int tileCap;         // bodies the shared-memory tile can hold
  int conCap;          // constraints whose 13 planes fit in shared memory next to the tile
  int nj;              // joints in the arena
  float invH;
  const int* jColour;  // [nj] colour of each joint (k_joint_colour); B2G_JOINT_COLOURS = serial tail
  const int* jSorted;  // [nj] joints by colour
  const int* jCstart;  // [B2G_JOINT_COLOURS + 2]
};
#define B2G_TILE_JOINTS 64
#define B2G_JOINT_COLOURS 60  // parallel joint colours (k_joint_colour); joints of a hub beyond them are walked serially
#define B2G_JOINT_COLOUR_THREADS 1024
#define B2G_TILE_JOINTS_SERIAL 64  // a tile (or the set of oversize islands) with at most this many joints walks them in list order
#define B2G_PLANES 13
#define B2G_LEVELS_MIN_BUCKET 128  // serial buckets from this size on are level-scheduled when that pays
#define B2G_LEVELS_MAX 96          // levels of the schedule (three borrowed 32-bit words per body)

// dynamic shared memory layout for a tile of `cap` bodies
struct FusedTile {
  float4* vel;
  float4* pos;
  int* body;        // global body index of each slot
  int* head;        // tile slot of the first body of this slot's island (per-island state lives there)
  unsigned int* pen;     // per island head: max penetration of the current position iteration
  unsigned int* sleepMin;  // per island head: float bits of min sleep time
  int* done;        // per island head: position solver converged
  // rounded up so the constraint planes that follow stay 16-byte aligned
  __host__ __device__ static size_t bytes(int cap) { return (((size_t)cap * (16 + 16 + 4 + 4 + 4 + 4 + 4)) + 15) & ~(size_t)15; }
};

__global__ void __launch_bounds__(B2G_FUSED_THREADS)
k_solve_bins_fused(FusedParams P, const int* __restrict__ binFirst, const int* __restrict__ binEnd,
                   const int* __restrict__ slotBody, const int* __restrict__ bodySlot,
                   const int* __restrict__ island, const int* __restrict__ islandStart,
                   const int* __restrict__ bucketStart, int* sortedList, int* orderScratch, ContactBuf C,
                   const float* __restrict__ fRadius, SolverPlanes S, uint32_t* bflags, float4* gpos, float4* gvel,
                   float4* gxf, float4* gforce, const float4* __restrict__ gmass, const float4* __restrict__ gcenter,
                   StepCounts* counts, JointArraysDev J, int* levelScratch, int* permScratch) {
  B2G_PDL_ENTER();
  const int bin = blockIdx.x;
  const int first = binFirst[bin];
  const int nbod = binEnd[bin] - first;
  if (nbod <= 0 || first >= 0x7f000000) return;
  const int tid = threadIdx.x, nt = blockDim.x;

  extern __shared__ __align__(16) unsigned char smemRaw[];
  FusedTile T;
  {
    unsigned char* p = smemRaw;
    T.vel = (float4*)p;  p += (size_t)P.tileCap * 16;
    T.pos = (float4*)p;  p += (size_t)P.tileCap * 16;
    T.body = (int*)p;    p += (size_t)P.tileCap * 4;
    T.head = (int*)p;    p += (size_t)P.tileCap * 4;
    T.pen = (unsigned int*)p;       p += (size_t)P.tileCap * 4;
    T.sleepMin = (unsigned int*)p;  p += (size_t)P.tileCap * 4;
    T.done = (int*)p;
  }
  __shared__ int cstart[B2G_MAX_COLOURS + 3];
  __shared__ int sJoint[B2G_TILE_JOINTS];
  __shared__ unsigned char sJointCol[B2G_TILE_JOINTS];
  __shared__ int sJointCount;
  __shared__ unsigned long long sJointColours;  // colours present among this tile's joints (bit B2G_JOINT_COLOURS: tail)
  if (tid == 0) sJointCount = 0, sJointColours = 0ull;

  // constraint ranges of this bin, one per colour (+ overflow), from the bucket table
  if (tid <= B2G_MAX_COLOURS + 1) cstart[tid] = P.nc > 0 ? bucketStart[(bin << B2G_COLOUR_BITS) + tid] : 0;

  // ---- phase 0: load the tile, integrate velocities (b2_island.cpp:257-293) --------------------
  const float h = P.h;
  for (int l = tid; l < nbod; l += nt) {
    int b = slotBody[first + l];
    T.body[l] = b;
    T.head[l] = islandStart[island[b]] - first;
    T.pen[l] = 0u;
    T.sleepMin[l] = __float_as_uint(B2G_MAX_FLOAT);
    T.done[l] = 0;
    uint32_t f = bflags[b];
    if (!(f & B2G_BODY_AWAKE)) bflags[b] = f | B2G_BODY_AWAKE;  // reached bodies are woken, timer kept
    float4 v4 = gvel[b];
    if (B2G_BODY_TYPE(f) == B2G_DYNAMIC) {
      float4 m4 = gmass[b], c4 = gcenter[b], f4 = gforce[b];
      float t = h * m4.x;
      float gs = m4.w * m4.z;
      v4.x += t * (gs * P.gravity.x + f4.x);
      v4.y += t * (gs * P.gravity.y + f4.y);
      v4.z += h * m4.y * f4.z;
      float dl = 1.0f + h * c4.z;
      float da = 1.0f + h * c4.w;
      v4.x /= dl;
      v4.y /= dl;
      v4.z /= da;
    }
    T.vel[l] = v4;
    T.pos[l] = gpos[b];
  }
  __syncthreads();
  // joints whose island lives in this tile (b2_island.cpp:323-325 walks the island's joint list)
  if (P.nj > 0) {
    for (int j = tid; j < P.nj; j += nt) {
      int2 bd = J.bodies[j];
      int sa = bodySlot[bd.x], sb = bodySlot[bd.y];
      int sl = sa >= 0 ? sa : sb;
      if (sl >= first && sl < first + nbod) {
        int k = atomicAdd(&sJointCount, 1);
        if (k < B2G_TILE_JOINTS) sJoint[k] = j;  // more than that: the walks below go through the arena's table
        atomicOr(&sJointColours, 1ull << P.jColour[j]);
      }
    }
    __syncthreads();
    if (tid == 0) {  // joint-index order, whatever order the atomics produced
      int n = min(sJointCount, B2G_TILE_JOINTS);
      for (int a = 1; a < n; ++a) {
        int v = sJoint[a], b = a - 1;
        while (b >= 0 && sJoint[b] > v) {
          sJoint[b + 1] = sJoint[b];
          --b;
        }
        sJoint[b + 1] = v;
      }
      for (int a = 0; a < n; ++a) sJointCol[a] = (unsigned char)P.jColour[sJoint[a]];
    }
    __syncthreads();
  }
  const int njTile = P.nj > 0 ? sJointCount : 0;
  const unsigned long long jointColours = P.nj > 0 ? sJointColours : 0ull;
  // The tile's joints, colour by colour (k_joint_colour: joints of a colour share no movable body), every thread
  // of the block taking part, a CTA barrier behind each colour; the joints of a hub beyond the colours follow in
  // descending index order on one thread (the list walk of b2Island::Solve).  Up to B2G_TILE_JOINTS joints come
  // from the shared list; a joint-heavy tile (a long chain, a crowd of ragdolls) goes through the arena's
  // colour-sorted table instead, slower but without a capacity.  Called by ALL threads.
  auto in_tile = [&](int j) {
    const int2 bd = J.bodies[j];
    const int sa = bodySlot[bd.x], sb = bodySlot[bd.y];
    const int sl = sa >= 0 ? sa : sb;
    return sl >= first && sl < first + nbod;
  };
  auto for_tile_joints = [&](auto&& f) {
    // up to a few dozen joints (ragdolls, a chain, the tumbler's hinge): one thread, descending index — the
    // order the reference's island DFS finds a chain built root to tip.  A sweep along a chain carries a
    // correction over its whole length where a two-colour sweep moves it two links per iteration (measured on
    // a 150-link chain: largest hinge gap 0.075 m coloured, under 0.05 m in list order), and costs ~0.5 us per
    // joint at this size
    const bool serial = njTile <= B2G_TILE_JOINTS_SERIAL;
    const bool listed = njTile <= B2G_TILE_JOINTS;
    for (unsigned long long m = serial ? 1ull : jointColours; m; m &= m - 1ull) {
      const int c = __ffsll((long long)m) - 1;
      const bool one = serial || c == B2G_JOINT_COLOURS;  // one thread, descending index
      const int k0 = listed ? 0 : P.jCstart[c], n = (listed ? njTile : P.jCstart[c + 1]) - k0;
      for (int t = one ? (tid == 0 ? 0 : n) : tid; t < n; t += one ? 1 : nt) {
        int j;
        bool mine;
        if (listed) {
          const int k = one ? n - 1 - t : t;
          j = sJoint[k];
          mine = serial || sJointCol[k] == c;
        } else {
          j = P.jSorted[k0 + t];  // (the tail of the table is in descending index order already)
          mine = in_tile(j);
        }
        if (mine) f(j);
      }
      __syncthreads();
    }
  };
  // the (few) non-empty parallel colours of this bin, so the pass loops do not walk all 24
  __shared__ int usedS0[B2G_MAX_COLOURS], usedS1[B2G_MAX_COLOURS];
  __shared__ int usedCount;
  if (tid == 0) {
    int k = 0;
    for (int c = 0; c < B2G_MAX_COLOURS; ++c)
      if (cstart[c] != cstart[c + 1]) {
        usedS0[k] = cstart[c];
        usedS1[k] = cstart[c + 1];
        ++k;
      }
    usedCount = k;
  }
  __syncthreads();
  const int nUsed = usedCount;
  order_bucket_by_key(cstart[B2G_MAX_COLOURS], cstart[B2G_MAX_COLOURS + 1], sortedList, orderScratch, C);
  const int cAll0 = cstart[0], cAll1 = cstart[B2G_MAX_COLOURS + 1];
  const TileBodies velAcc{T.vel, gvel};
  const TileBodies posAcc{T.pos, gpos};
  // Constraint planes: in shared memory when the whole bin fits (they are re-read by every one of
  // the 1 + velIters + posIters passes, each on the critical path between two barriers), else in
  // HBM/L2.  Slot s of a plane is addressed as plane[s] either way (shared base is pre-offset).
  if (cAll1 - cAll0 <= P.conCap) {
    float4* base = (float4*)(smemRaw + FusedTile::bytes(P.tileCap));
    const ptrdiff_t cc = P.conCap;
    S.nf = base + 0 * cc - cAll0;
    S.r1 = base + 1 * cc - cAll0;
    S.r2 = base + 2 * cc - cAll0;
    S.m1 = base + 3 * cc - cAll0;
    S.m2 = base + 4 * cc - cAll0;
    S.kk = base + 5 * cc - cAll0;
    S.mass = base + 6 * cc - cAll0;
    S.idx = (int4*)(base + 7 * cc) - cAll0;
    S.imp = base + 8 * cc - cAll0;
    S.pn = base + 9 * cc - cAll0;
    S.pp = base + 10 * cc - cAll0;
    S.pc = base + 11 * cc - cAll0;
    S.pr = base + 12 * cc - cAll0;
  }
  // The SERIAL bucket (constraints of a hub body beyond the 24 colours: the tumbler's container touches ~65
  // boxes) is one thread walking a chain, one dependent L2 / HBM round trip per constraint just to fetch its
  // constants.  When the planes live in global memory the otherwise unused plane area of this block holds
  // the bucket's constants instead: velocity planes for warm start + velocity iterations (impulses written
  // back once), then the position planes.
  const int ov0 = cstart[B2G_MAX_COLOURS], ov1 = cstart[B2G_MAX_COLOURS + 1];
  SolverPlanes V = S;  // view of the staged serial bucket (slots ov0 .. ov0 + nStaged)
  int nStaged = 0;
  if (cAll1 - cAll0 > P.conCap && ov1 > ov0) {
    nStaged = min(ov1 - ov0, (P.conCap * B2G_PLANES) / 9);
    float4* base = (float4*)(smemRaw + FusedTile::bytes(P.tileCap));
    const ptrdiff_t cc = nStaged;
    V.idx = (int4*)(base + 0 * cc) - ov0;
    V.mass = base + 1 * cc - ov0;
    V.nf = base + 2 * cc - ov0;
    V.r1 = base + 3 * cc - ov0;
    V.r2 = base + 4 * cc - ov0;
    V.m1 = base + 5 * cc - ov0;
    V.m2 = base + 6 * cc - ov0;
    V.kk = base + 7 * cc - ov0;
    V.imp = base + 8 * cc - ov0;
    V.pn = base + 2 * cc - ov0;  // the position planes reuse the velocity slots
    V.pp = base + 3 * cc - ov0;
    V.pc = base + 4 * cc - ov0;
    V.pr = base + 5 * cc - ov0;
  }
  const int ovStaged = ov0 + nStaged;  // serial slots below this one are read from V, the others from S

  // ---- phase 1: prepare every constraint of the bin --------------------------------------------
  for (int s = cAll0 + tid; s < cAll1; s += nt) {
    int i = sortedList[s];
    Manifold m;
    manifold_unpack(m, C.m0[i], C.m1[i], C.m2[i], C.m3[i]);
    int2 bd = C.body[i];
    int2 fx = C.fix[i];
    int sa = bodySlot[bd.x], sb = bodySlot[bd.y];
    int ia = sa >= 0 ? sa - first : ~bd.x;
    int ib = sb >= 0 ? sb - first : ~bd.y;
    prepare_constraint(S, s, i, m, bd.x, bd.y, ia, ib, C.material[i], fRadius[fx.x], fRadius[fx.y], posAcc, velAcc,
                       gmass, gcenter, P.dtRatio, P.warmStarting != 0);
  }
  __syncthreads();
  if (nStaged > 0) {
    for (int t = tid; t < nStaged; t += nt) {
      const int sl = ov0 + t;
      V.idx[sl] = S.idx[sl];
      V.mass[sl] = S.mass[sl];
      V.nf[sl] = S.nf[sl];
      V.r1[sl] = S.r1[sl];
      V.r2[sl] = S.r2[sl];
      V.m1[sl] = S.m1[sl];
      V.m2[sl] = S.m2[sl];
      V.kk[sl] = S.kk[sl];
      V.imp[sl] = S.imp[sl];
    }
    __syncthreads();
  }

  // ---- level schedule of a LARGE serial bucket ----------------------------------------------------------------
  // Bodies that overlap dozens of others at once (the reference's "n^2" and multi-fixture benchmarks) put
  // thousands of constraints beyond the 24 colours, and walking them with one thread is what such a step then
  // costs.  They are given LEVELS here — a second, per-bin colouring with 96 more colours: lane 0 of warp 0
  // walks the bucket in key order and gives each constraint the lowest level free on both of its bodies (the
  // bodies' used levels are 96-bit sets in three borrowed per-body words of the tile); constraints of one level
  // share no body and are solved by the whole block, level after level.  What finds none of the 96 free (a body
  // with more than ~100 extra contacts) gets a chain level above them: one more than the highest chain level
  // either of its bodies has reached.  Deterministic: levels are a function of the key-ordered bucket, the
  // order inside a level is immaterial.  Used when it pays (a hub's chain — the tumbler container's ~40 extra
  // contacts — stays below the size limit and keeps the plain walk).
  __shared__ int sLevels, sTail;  // levels walked by the whole block; constraints of the serial tail behind them
  const int nOv = ov1 - ov0;
  int* const lvl = orderScratch;             // [slot] level (the key sort is done with its scratch)
  int* const perm = permScratch;             // [ov0 + k] the bucket's slots grouped by level
  int* const levelStart = levelScratch + ov0;  // [levels + 2] starts (the tail is group `levels`), then [levels + 1] cursors
  if (tid == 0) sLevels = 0, sTail = 0;
  __syncthreads();  // (every thread reads sLevels below, also in blocks that skip the scheduling)
  if (nOv >= B2G_LEVELS_MIN_BUCKET) {  // block-uniform
    for (int t = tid; t < nOv; t += nt) {
      const int4 ix = S.idx[ov0 + t];
      perm[ov0 + t] = (int)(((unsigned int)(ix.x >= 0 ? ix.x : 0xFFFF) << 16) | (unsigned int)(ix.y >= 0 ? ix.y : 0xFFFF));
    }
    for (int l = tid; l < nbod; l += nt) T.sleepMin[l] = 0u, T.head[l] = 0;  // borrowed with T.pen and T.done (both zero here)
    __syncthreads();
    if (tid < 32) {
      // 32 entries per coalesced load; the serial part is a few shared-memory words per constraint
      int top = -1, chainTop = -1, beyond = 0;  // highest first-fit level, highest chain level, constraints beyond the 96
      for (int base = ov0; base < ov1; base += 32) {
        const int mine = base + tid < ov1 ? perm[base + tid] : -1;
        int myLevel = 0;
        const int cnt = ov1 - base < 32 ? ov1 - base : 32;
        for (int u = 0; u < cnt; ++u) {
          const unsigned int pk = (unsigned int)__shfl_sync(0xffffffffu, mine, u);
          int L = 0;
          if (tid == 0) {
            const int a = (int)(pk >> 16), b = (int)(pk & 0xFFFFu);
            unsigned int u0 = 0u, u1 = 0u, u2 = 0u;
            if (a != 0xFFFF) u0 |= T.pen[a], u1 |= T.sleepMin[a], u2 |= (unsigned int)T.done[a];
            if (b != 0xFFFF) u0 |= T.pen[b], u1 |= T.sleepMin[b], u2 |= (unsigned int)T.done[b];
            L = ~u0 ? __ffs((int)~u0) - 1 : (~u1 ? 32 + __ffs((int)~u1) - 1 : (~u2 ? 64 + __ffs((int)~u2) - 1 : B2G_LEVELS_MAX));
            if (L == B2G_LEVELS_MAX) {
              // no level left on these bodies: chain levels above the 96 — one more than the highest chain level
              // either body has reached (T.head borrowed as that counter)
              const int ca = a != 0xFFFF ? T.head[a] : 0, cb = b != 0xFFFF ? T.head[b] : 0;
              const int c = ca > cb ? ca : cb;
              L = B2G_LEVELS_MAX + c;
              if (a != 0xFFFF) T.head[a] = c + 1;
              if (b != 0xFFFF) T.head[b] = c + 1;
              chainTop = c > chainTop ? c : chainTop;
              ++beyond;
            } else if (L < B2G_LEVELS_MAX) {
              const unsigned int bit = 1u << (L & 31);
              if (a != 0xFFFF) {
                if (L < 32) T.pen[a] |= bit;
                else if (L < 64) T.sleepMin[a] |= bit;
                else T.done[a] = (int)((unsigned int)T.done[a] | bit);
              }
              if (b != 0xFFFF) {
                if (L < 32) T.pen[b] |= bit;
                else if (L < 64) T.sleepMin[b] |= bit;
                else T.done[b] = (int)((unsigned int)T.done[b] | bit);
              }
              top = L > top ? L : top;
            }
          }
          L = __shfl_sync(0xffffffffu, L, 0);
          if (tid == u) myLevel = L;
        }
        if (base + tid < ov1) lvl[base + tid] = myLevel;
      }
      // one level pass costs about four serial visits
      if (tid == 0 && top >= 0) {
        // What lies beyond the 96 levels is either walked as chain levels (wide chains: bodies piled on one spot)
        // or by one thread in key order (narrow ones: dozens of contacts between the same two bodies); a level
        // pass costs about four serial visits.
        const bool chain = beyond > 0 && (long long)(chainTop + 1) * 4 <= (long long)beyond;
        const int levels = chain ? B2G_LEVELS_MAX + chainTop + 1 : top + 1;
        const int tail = chain ? 0 : beyond;
        if ((long long)levels * 4 + tail <= (long long)nOv * 3 / 4) sLevels = levels, sTail = tail;
      }
    }
    __syncthreads();
    for (int l = tid; l < nbod; l += nt) {  // give the borrowed words back as the tile load left them
      T.pen[l] = 0u;
      T.sleepMin[l] = __float_as_uint(B2G_MAX_FLOAT);
      T.done[l] = 0;
      T.head[l] = islandStart[island[T.body[l]]] - first;
    }
    const int nl = sLevels;
    if (nl > 0) {
      const bool hasTail = sTail > 0;  // then every level >= 96 is the tail = group `nl`
      for (int t = tid; t <= 2 * nl + 2; t += nt) levelStart[t] = 0;
      __syncthreads();
      for (int t = tid; t < nOv; t += nt) {
        const int L = lvl[ov0 + t];
        atomicAdd(&levelStart[(hasTail && L >= B2G_LEVELS_MAX ? nl : L) + 1], 1);
      }
      __syncthreads();
      if (tid == 0)
        for (int L = 0; L <= nl; ++L) levelStart[L + 1] += levelStart[L];
      __syncthreads();
      int* const cursor = levelStart + nl + 2;
      for (int t = tid; t < nOv; t += nt) {  // order inside a level is immaterial: its constraints share no body
        const int L0 = lvl[ov0 + t];
        const int L = hasTail && L0 >= B2G_LEVELS_MAX ? nl : L0;
        perm[ov0 + levelStart[L] + atomicAdd(&cursor[L], 1)] = ov0 + t;
      }
      __syncthreads();
      // the tail is walked by one thread, so its order matters: ascending slot = key order
      sort_ints_ascending(perm + ov0 + levelStart[nl], hasTail ? sTail : 0);
    }
    __syncthreads();
  }
  const int nLevels = sLevels;
  // one pass over the serial bucket, by ALL threads: level by level, or (no schedule) one thread in key order
  auto serial_bucket = [&](auto&& visit) {
    if (nLevels > 0) {
      // levels 0 .. nLevels - 1 with the whole block; group nLevels is the tail (empty without one): one thread
      for (int L = 0; L <= nLevels; ++L) {
        const int k1 = levelStart[L + 1];
        const bool tail = L == nLevels;
        for (int k = tail ? (tid == 0 ? levelStart[L] : k1) : levelStart[L] + tid; k < k1; k += tail ? 1 : nt) {
          const int s = perm[ov0 + k];
          visit(s < ovStaged ? V : S, s);
        }
        __syncthreads();
      }
    } else {
      if (tid == 0) {
        for (int half = 0; half < 2; ++half) {  // the staged slots, then the rest: one call site for both
          const SolverPlanes& PL = half == 0 ? V : S;
          const int s1 = half == 0 ? ovStaged : ov1;
          for (int s = half == 0 ? ov0 : ovStaged; s < s1; ++s) visit(PL, s);
        }
      }
      __syncthreads();
    }
  };

  // ---- phase 2: warm start, colour by colour ---------------------------------------------------
  if (P.warmStarting) {
    for (int k = 0; k < nUsed; ++k) {
      int s0 = usedS0[k], s1 = usedS1[k];
      for (int s = s0 + tid; s < s1; s += nt) warm_start_constraint(S, s, velAcc);
      __syncthreads();
    }
    if (nOv > 0) serial_bucket([&](const SolverPlanes& PL, int s) { warm_start_constraint(PL, s, velAcc); });
  }

  // joints: InitVelocityConstraints incl. their warm start (b2_island.cpp:323-325), after the contacts'
  if (njTile > 0) {
    for_tile_joints([&](int j) {
      int2 bd = J.bodies[j];
      int sa = bodySlot[bd.x], sb = bodySlot[bd.y];
      joint_init(J, j, sa >= 0 ? sa - first : ~bd.x, sb >= 0 ? sb - first : ~bd.y, posAcc, velAcc, gmass, gcenter,
                 P.dtRatio, P.warmStarting != 0);
    });
  }

  // ---- phase 3: velocity iterations ---------------------------------------------------------------
  for (int it = 0; it < P.velIters; ++it) {
    if (njTile > 0)  // joints first, then contacts (b2_island.cpp:330-338)
      for_tile_joints([&](int j) { joint_solve_velocity(J, j, velAcc, P.h, P.invH); });
    for (int k = 0; k < nUsed; ++k) {
      int s0 = usedS0[k], s1 = usedS1[k];
      for (int s = s0 + tid; s < s1; s += nt) solve_velocity_constraint(S, s, velAcc);
      __syncthreads();
    }
    if (nOv > 0) serial_bucket([&](const SolverPlanes& PL, int s) { solve_velocity_constraint(PL, s, velAcc); });
  }

  // ---- phase 4: store impulses (b2_contact_solver.cpp:641-657) -----------------------------------
  for (int s = cAll0 + tid; s < cAll1; s += nt) {
    int4 ix = S.idx[s];
    float4 imp = (s >= ov0 && s < ovStaged) ? V.imp[s] : S.imp[s];
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

  // ---- phase 5: integrate positions (b2_island.cpp:353-385) --------------------------------------
  for (int l = tid; l < nbod; l += nt) {
    float4 p4 = T.pos[l], v4 = T.vel[l];
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
    T.pos[l] = p4;
    T.vel[l] = make_float4(v.x, v.y, w, v4.w);
  }
  __syncthreads();  // (also: phase 4 has read the staged impulses)
  if (nStaged > 0 && P.posIters > 0) {
    for (int t = tid; t < nStaged; t += nt) {
      const int sl = ov0 + t;
      V.pn[sl] = S.pn[sl];
      V.pp[sl] = S.pp[sl];
      V.pc[sl] = S.pc[sl];
      V.pr[sl] = S.pr[sl];
    }
    __syncthreads();
  }

  // ---- phase 6: position iterations with the per-island early exit (b2_island.cpp:391-409) --------
  for (int it = 0; it < P.posIters; ++it) {
    auto position_visit = [&](const SolverPlanes& PL, int s) {
      int4 ix = PL.idx[s];
      int slot = ix.x >= 0 ? ix.x : ix.y;  // a tile member of the island (the other may be static)
      int hd = T.head[slot];
      if (T.done[hd]) return;
      float minSep = solve_position_constraint(PL, s, posAcc);
      float pen = minSep < 0.0f ? -minSep : 0.0f;
      atomicMax(&T.pen[hd], __float_as_uint(pen));
    };
    for (int k = 0; k < nUsed; ++k) {
      const int s1 = usedS1[k];
      for (int s = usedS0[k] + tid; s < s1; s += nt) position_visit(S, s);
      __syncthreads();
    }
    if (nOv > 0) serial_bucket(position_visit);
    if (njTile > 0) {  // contacts first, then joints (b2_island.cpp:392-401); a joint that is not
                       // okay keeps its island iterating, expressed as a large "penetration"
      for_tile_joints([&](int j) {
        const int ia = J.work[j].ia, ib = J.work[j].ib;
        int slot = ia >= 0 ? ia : ib;
        int hd = T.head[slot];
        if (T.done[hd]) return;
        if (!joint_solve_position(J, j, posAcc)) atomicMax(&T.pen[hd], __float_as_uint(1.0f));
      });
    }
    // island converged? (contactsOkay && jointsOkay: minSeparation >= -3 slop)
    for (int l = tid; l < nbod; l += nt) {
      if (T.head[l] == l && !T.done[l]) {
        if (__uint_as_float(T.pen[l]) <= 3.0f * B2G_LINEAR_SLOP) T.done[l] = 1;
        T.pen[l] = 0u;
      }
    }
    __syncthreads();
  }

  // ---- phase 7: write back, SynchronizeTransform, sleep (b2_island.cpp:430-483), ClearForces ------
  const float linTolSqr = B2G_LINEAR_SLEEP_TOL * B2G_LINEAR_SLEEP_TOL;
  const float angTolSqr = B2G_ANGULAR_SLEEP_TOL * B2G_ANGULAR_SLEEP_TOL;
  if (P.allowSleep) {
    for (int l = tid; l < nbod; l += nt) {
      int b = T.body[l];
      uint32_t f = bflags[b];
      float4 v4 = T.vel[l];
      float st = gforce[b].w;
      float minSleep;
      if (!(f & B2G_BODY_AUTOSLEEP) || v4.z * v4.z > angTolSqr || v4.x * v4.x + v4.y * v4.y > linTolSqr) {
        st = 0.0f;
        minSleep = 0.0f;
      } else {
        st += h;
        minSleep = st;
      }
      T.vel[l].w = st;  // park the new sleep time in the unused lane
      atomicMin(&T.sleepMin[T.head[l]], __float_as_uint(minSleep));
    }
    __syncthreads();
  }
  int awake = 0;
  for (int l = tid; l < nbod; l += nt) {
    int b = T.body[l];
    float4 p4 = T.pos[l], v4 = T.vel[l], c4 = gcenter[b];
    Xf X = xf_from_sweep(make_float2(p4.x, p4.y), p4.z, make_float2(c4.x, c4.y));
    gpos[b] = p4;
    gxf[b] = xf_to4(X);
    float4 fo = gforce[b];
    bool sleepNow = false;
    if (P.allowSleep) {
      int hd = T.head[l];
      fo.w = v4.w;
      sleepNow = __uint_as_float(T.sleepMin[hd]) >= B2G_TIME_TO_SLEEP && T.done[hd] != 0;
    }
    if (sleepNow) {
      // b2Body::SetAwake(false), b2_body.h:731-739
      bflags[b] &= ~B2G_BODY_AWAKE;
      gvel[b] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
      gforce[b] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    } else {
      gvel[b] = make_float4(v4.x, v4.y, v4.z, 0.0f);
      if (P.clearForces) fo.x = fo.y = fo.z = 0.0f;
      gforce[b] = fo;
      ++awake;
    }
  }
  if (awake) atomicAdd(&counts->numAwake, awake);
}


// ---------------------------------------------------------------------------------------------
// Oversize islands (more bodies than a tile holds, e.g. a settled 100k-body pile = ONE island):
// a persistent cooperative kernel walks warm start, the velocity iterations, position integration
// and the position iterations colour by colour with a grid-wide barrier between colours, instead
// of one launch per colour per iteration (14 colours x 12 passes = 170 launches of ~16k
// constraints each were launch-latency bound).  Body state stays in global memory / L2.
// ---------------------------------------------------------------------------------------------
// ---- the persistent big-island kernel ----------------------------------------------------------------
// Body state of an oversize island lives in L2/HBM and is exchanged between SMs every colour pass.
struct CoherentBodies {  // L2 loads / stores: the other SMs' writes of the previous pass, never a stale L1 line
  float4* a;
  __device__ __forceinline__ float4 load(int i) const { return __ldcg(a + i); }
  __device__ __forceinline__ void store(int i, float4 v) const { __stcg(a + i, v); }
};
// ---- joints of islands that are solved through the global arrays (big islands, sequential mode):
// one thread, joint-index order
struct JointWalk {
  int nj, onlyBig;
  const uint32_t* bflags;
  const int* island;
  const uint32_t* islandAwake;
  const int* bodySlot;
  const int* order;  // explicit visiting order (b2g_set_sequential_joint_order), nullptr = descending index
  int norder;
  // production mode: joints sorted by colour (k_joint_colour); cstart[c] .. cstart[c + 1] = colour c,
  // cstart[B2G_JOINT_COLOURS] .. cstart[B2G_JOINT_COLOURS + 1] = joints beyond the colours (index order)
  const int* sorted;
  const int* cstart;
};
// k-th joint of the walk
__device__ __forceinline__ int joint_walk_count(const JointWalk& W) { return W.order ? W.norder : W.nj; }
__device__ __forceinline__ int joint_walk_at(const JointWalk& W, int k) { return W.order ? W.order[k] : W.nj - 1 - k; }
__device__ __forceinline__ int joint_owner_body(const JointWalk& W, const JointArraysDev& J, int j) {
  int2 bd = J.bodies[j];
  uint32_t fa = W.bflags[bd.x], fb = W.bflags[bd.y];
  if (!(fa & B2G_BODY_ENABLED) || !(fb & B2G_BODY_ENABLED)) return -1;
  int s = B2G_BODY_TYPE(fa) != B2G_STATIC ? bd.x : bd.y;
  if (B2G_BODY_TYPE(W.bflags[s]) == B2G_STATIC) return -1;
  if (!W.islandAwake[W.island[s]]) return -1;
  if (W.onlyBig && W.bodySlot[s] != B2G_SLOT_BIG) return -1;
  return s;
}
template <class Acc = GlobalBodies>
__device__ __forceinline__ void joints_init_global(const JointWalk& W, const JointArraysDev& J, float4* pos, float4* vel,
                                                   const float4* __restrict__ mass, const float4* __restrict__ center,
                                                   float dtRatio, int warm) {
  for (int k = 0; k < joint_walk_count(W); ++k) {
    const int j = joint_walk_at(W, k);
    if (joint_owner_body(W, J, j) < 0) continue;
    int2 bd = J.bodies[j];
    joint_init(J, j, bd.x, bd.y, Acc{pos}, Acc{vel}, mass, center, dtRatio, warm != 0);
  }
}
template <class Acc = GlobalBodies>
__device__ __forceinline__ void joints_velocity_global(const JointWalk& W, const JointArraysDev& J, float4* vel, float h,
                                                       float invH) {
  for (int k = 0; k < joint_walk_count(W); ++k) {
    const int j = joint_walk_at(W, k);
    if (joint_owner_body(W, J, j) >= 0) joint_solve_velocity(J, j, Acc{vel}, h, invH);
  }
}
template <class Acc = GlobalBodies>
__device__ __forceinline__ void joints_position_global(const JointWalk& W, const JointArraysDev& J, float4* pos,
                                                       uint32_t* islandPen, int penStride, int iter) {
  for (int k = 0; k < joint_walk_count(W); ++k) {
    const int j = joint_walk_at(W, k);
    int s = joint_owner_body(W, J, j);
    if (s < 0) continue;
    int root = W.island[s];
    if (island_done(islandPen, penStride, iter, root)) continue;
    if (!joint_solve_position(J, j, Acc{pos}))
      atomicMax(&islandPen[(size_t)iter * penStride + root], __float_as_uint(1.0f));
  }
}
// ---- joint colouring (production mode) -------------------------------------------------------------------------
// The reference walks an island's joints one after the other (b2_island.cpp:323-338, 392-401).  Like the contacts,
// joints that share no movable body commute, so the joint table is greedily coloured whenever it (or a body's
// mass) changes and a colour is then solved by many threads at once: a 4 000-link mobile (the reference's
// "Big mobile" benchmark, a binary tree of revolute joints) needs 3 colours instead of a 4 000-joint serial walk
// per iteration.  One block, Jones-Plassmann rounds on hashed priorities: an uncoloured joint that holds the
// highest priority on both of its movable bodies takes the lowest colour free on both.  A body with more than
// B2G_JOINT_COLOURS joints (a hub) leaves the rest to the serial tail, walked by one thread in index order.
__device__ __forceinline__ bool joint_moves_body(uint32_t type, int side, float4 m) {
  // the mouse joint writes bodyB back unconditionally (b2_mouse_joint.cpp:139-158)
  return body_movable(m) || (type == B2G_JOINT_MOUSE && side == 1);
}
__global__ void __launch_bounds__(B2G_JOINT_COLOUR_THREADS)
k_joint_colour(int nj, int nb, const int2* __restrict__ jBodies, const float4* __restrict__ jParams1,
               const float4* __restrict__ mass, unsigned long long* bodyMask, unsigned long long* bodyBest, int* colour,
               int* sorted, int* cstart) {
  B2G_PDL_ENTER();
  const int tid = threadIdx.x, nt = blockDim.x;
  __shared__ int sRemaining, sCount[B2G_JOINT_COLOURS + 2], sCursor[B2G_JOINT_COLOURS + 2];
  for (int b = tid; b < nb; b += nt) bodyMask[b] = 0ull, bodyBest[b] = 0ull;
  for (int j = tid; j < nj; j += nt) colour[j] = -1;
  if (tid <= B2G_JOINT_COLOURS + 1) sCount[tid] = 0;
  __syncthreads();
  auto prio = [](int j) {
    unsigned int hsh = (unsigned int)j * 2654435761u;
    hsh ^= hsh >> 15;
    hsh *= 2246822519u;
    hsh ^= hsh >> 13;
    return ((unsigned long long)hsh << 32) | (unsigned long long)(unsigned int)(j + 1);
  };
  for (int round = 0; round < nj + 1; ++round) {
    if (tid == 0) sRemaining = 0;
    __syncthreads();
    for (int j = tid; j < nj; j += nt) {
      if (colour[j] != -1) continue;
      const int2 bd = jBodies[j];
      const uint32_t type = B2G_JOINT_TYPE(__float_as_uint(jParams1[j].y));
      const unsigned long long pr = prio(j);
      if (joint_moves_body(type, 0, mass[bd.x])) atomicMax(&bodyBest[bd.x], pr);
      if (joint_moves_body(type, 1, mass[bd.y])) atomicMax(&bodyBest[bd.y], pr);
    }
    __syncthreads();
    for (int j = tid; j < nj; j += nt) {
      if (colour[j] != -1) continue;
      const int2 bd = jBodies[j];
      const uint32_t type = B2G_JOINT_TYPE(__float_as_uint(jParams1[j].y));
      const unsigned long long pr = prio(j);
      const bool mvA = joint_moves_body(type, 0, mass[bd.x]), mvB = joint_moves_body(type, 1, mass[bd.y]) && bd.y != bd.x;
      if ((mvA && bodyBest[bd.x] != pr) || (mvB && bodyBest[bd.y] != pr)) {
        sRemaining = 1;
        continue;
      }
      const unsigned long long used = (mvA ? bodyMask[bd.x] : 0ull) | (mvB ? bodyMask[bd.y] : 0ull);
      const unsigned long long freeBits = ~used & ((1ull << B2G_JOINT_COLOURS) - 1ull);
      const int c = freeBits ? __ffsll((long long)freeBits) - 1 : B2G_JOINT_COLOURS;
      colour[j] = c;
      if (c < B2G_JOINT_COLOURS) {  // the winner is alone on both bodies this round: plain stores
        if (mvA) bodyMask[bd.x] = bodyMask[bd.x] | (1ull << c);
        if (mvB) bodyMask[bd.y] = bodyMask[bd.y] | (1ull << c);
      }
      atomicAdd(&sCount[c], 1);
    }
    __syncthreads();
    if (!sRemaining) break;
    for (int j = tid; j < nj; j += nt) {  // the losers' bodies forget this round's winner
      if (colour[j] != -1) continue;
      const int2 bd = jBodies[j];
      bodyBest[bd.x] = 0ull;
      bodyBest[bd.y] = 0ull;
    }
    __syncthreads();
  }
  if (tid == 0) {
    int run = 0;
    for (int c = 0; c <= B2G_JOINT_COLOURS; ++c) {
      cstart[c] = run;
      sCursor[c] = run;
      run += sCount[c];
    }
    cstart[B2G_JOINT_COLOURS + 1] = run;
  }
  __syncthreads();
  for (int j = tid; j < nj; j += nt) {
    const int c = colour[j];
    if (c < B2G_JOINT_COLOURS) sorted[atomicAdd(&sCursor[c], 1)] = j;  // order inside a colour is immaterial
  }
  if (tid == 0 && sCount[B2G_JOINT_COLOURS] > 0) {  // serial tail: descending index, like the list walk
    int k = cstart[B2G_JOINT_COLOURS];
    for (int j = nj - 1; j >= 0; --j)
      if (colour[j] == B2G_JOINT_COLOURS) sorted[k++] = j;
  }
}

// Colour by colour over the whole grid; `sync` is the caller's grid barrier.  cstart is the same for every
// thread, so all of them take the same barriers.
template <class Sync, class F>
__device__ __forceinline__ void joints_coloured_grid(const JointWalk& W, const JointArraysDev& J, int gtid, int gsize,
                                                     Sync&& sync, F&& f) {
  const bool serial = W.nj <= B2G_TILE_JOINTS_SERIAL;  // a handful of joints: list order (descending index), one thread
  for (int c = serial ? B2G_JOINT_COLOURS : 0; c <= B2G_JOINT_COLOURS; ++c) {
    const int k0 = serial ? 0 : W.cstart[c], n = (serial ? W.nj : W.cstart[c + 1]) - k0;
    if (n == 0) continue;
    const bool one = c == B2G_JOINT_COLOURS;
    for (int t = one ? (gtid == 0 ? 0 : n) : gtid; t < n; t += one ? 1 : gsize) {
      const int j = serial ? W.nj - 1 - t : W.sorted[k0 + t];
      const int s = joint_owner_body(W, J, j);
      if (s >= 0) f(j, s);
    }
    sync();
  }
}
template <class Acc, class Sync>
__device__ __forceinline__ void joints_init_coloured(const JointWalk& W, const JointArraysDev& J, int gtid, int gsize,
                                                     Sync&& sync, float4* pos, float4* vel,
                                                     const float4* __restrict__ mass, const float4* __restrict__ center,
                                                     float dtRatio, int warm) {
  joints_coloured_grid(W, J, gtid, gsize, sync, [&](int j, int) {
    const int2 bd = J.bodies[j];
    joint_init(J, j, bd.x, bd.y, Acc{pos}, Acc{vel}, mass, center, dtRatio, warm != 0);
  });
}
template <class Acc, class Sync>
__device__ __forceinline__ void joints_velocity_coloured(const JointWalk& W, const JointArraysDev& J, int gtid, int gsize,
                                                         Sync&& sync, float4* vel, float h, float invH) {
  joints_coloured_grid(W, J, gtid, gsize, sync, [&](int j, int) { joint_solve_velocity(J, j, Acc{vel}, h, invH); });
}
template <class Acc, class Sync>
__device__ __forceinline__ void joints_position_coloured(const JointWalk& W, const JointArraysDev& J, int gtid, int gsize,
                                                         Sync&& sync, float4* pos, uint32_t* islandPen, int penStride,
                                                         int iter) {
  joints_coloured_grid(W, J, gtid, gsize, sync, [&](int j, int s) {
    const int root = W.island[s];
    if (island_done(islandPen, penStride, iter, root)) return;
    if (!joint_solve_position(J, j, Acc{pos})) atomicMax(&islandPen[(size_t)iter * penStride + root], __float_as_uint(1.0f));
  });
}

__global__ void k_joints_init_seq(JointWalk W, JointArraysDev J, float4* pos, float4* vel, const float4* mass,
                                  const float4* center, float dtRatio, int warm) {
  B2G_PDL_ENTER();
  joints_init_global(W, J, pos, vel, mass, center, dtRatio, warm);
}
__global__ void k_joints_velocity_seq(JointWalk W, JointArraysDev J, float4* vel, float h, float invH) {
  B2G_PDL_ENTER();
  joints_velocity_global(W, J, vel, h, invH);
}
__global__ void k_joints_position_seq(JointWalk W, JointArraysDev J, float4* pos, uint32_t* islandPen, int penStride,
                                      int iter) {
  B2G_PDL_ENTER();
  joints_position_global(W, J, pos, islandPen, penStride, iter);
}

// Split grid barrier on a monotonic counter (scripts/micro/grid_barrier.cu: 1.28 us vs 1.49 us for
// cooperative groups at 148 blocks).  arrive: the block's stores are ordered before thread 0's release
// increment by the CTA barrier (cumulativity); wait: thread 0 spins with acquire loads, the CTA barrier
// hands the observation to the rest of the block.  Work placed between the two halves (the prefetch of
// the next pass's constraint constants) overlaps the wait.
#ifdef B2G_BIG_TRACE  // debug build: per-block timestamps of every barrier (scripts/gpu_big_trace.py)
#define B2G_TRACE_CAP 1024
__device__ unsigned long long g_bigTrace[160 * B2G_TRACE_CAP * 2];
__device__ __forceinline__ unsigned long long trace_now() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
#endif
__device__ __forceinline__ void grid_arrive(unsigned int* counter, unsigned int& target) {
  __syncthreads();
  target += gridDim.x;
#ifdef B2G_BIG_TRACE
  if (threadIdx.x == 0 && target / gridDim.x <= B2G_TRACE_CAP)
    g_bigTrace[((size_t)blockIdx.x * B2G_TRACE_CAP + target / gridDim.x - 1) * 2] = trace_now();
#endif
  if (threadIdx.x == 0) asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
}
__device__ __forceinline__ void grid_wait(unsigned int* counter, unsigned int target) {
  if (threadIdx.x == 0) {
    unsigned int v;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
    } while ((int)(v - target) < 0);
#ifdef B2G_BIG_TRACE
    if (target / gridDim.x <= B2G_TRACE_CAP)
      g_bigTrace[((size_t)blockIdx.x * B2G_TRACE_CAP + target / gridDim.x - 1) * 2 + 1] = trace_now();
#endif
  }
  __syncthreads();
}
__device__ __forceinline__ void stage_copy16(void* smemDst, const void* globalSrc) {
  unsigned int d = (unsigned int)__cvta_generic_to_shared(smemDst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(globalSrc) : "memory");
}
__device__ __forceinline__ void stage_wait() { asm volatile("cp.async.wait_all;" ::: "memory"); }

#define B2G_BIG_THREADS 512
#define B2G_BIG_STAGE_PLANES 9
#define B2G_STAGE_VELOCITY 0  // idx, mass, nf, r1, r2, m1, m2, kk, imp  (warm start + velocity rows)
#define B2G_STAGE_POSITION 1  // idx, mass, pn, pp, pc, pr
// Each thread owns one 16-byte slot per plane in shared memory and keeps there the constants of the ONE
// constraint it will visit in the next colour pass, fetched with cp.async while the grid barrier of the
// current pass is still collecting arrivals.  After the barrier only the two body records (L2, written
// by other SMs a moment ago) stand between the thread and the arithmetic.  `staged` = solver slot whose
// planes are in the thread's slots (-1 none), so a wrong guess only costs a synchronous refill.
struct BigStage {
  SolverPlanes T;  // views of the shared-memory slots, indexed by threadIdx.x
  int staged, kind;
};
__device__ __forceinline__ void stage_issue(BigStage& G, const SolverPlanes& S, int s, int kind) {
  const int t = threadIdx.x;
  stage_copy16(G.T.idx + t, S.idx + s);
  stage_copy16(G.T.mass + t, S.mass + s);
  if (kind == B2G_STAGE_VELOCITY) {
    stage_copy16(G.T.nf + t, S.nf + s);
    stage_copy16(G.T.r1 + t, S.r1 + s);
    stage_copy16(G.T.r2 + t, S.r2 + s);
    stage_copy16(G.T.m1 + t, S.m1 + s);
    stage_copy16(G.T.m2 + t, S.m2 + s);
    stage_copy16(G.T.kk + t, S.kk + s);
    stage_copy16(G.T.imp + t, S.imp + s);
  } else {
    stage_copy16(G.T.pn + t, S.pn + s);
    stage_copy16(G.T.pp + t, S.pp + s);
    stage_copy16(G.T.pc + t, S.pc + s);
    stage_copy16(G.T.pr + t, S.pr + s);
  }
  G.staged = s;
  G.kind = kind;
}
__device__ __forceinline__ void stage_prefetch(BigStage& G, const SolverPlanes& S, int s, int kind) {
  if (G.staged != s || G.kind != kind) stage_issue(G, S, s, kind);  // same slot again: the staged copy is current
}
// make slot s of `kind` available in the thread's shared-memory slots (no copy when the prefetch guessed right)
__device__ __forceinline__ void stage_acquire(BigStage& G, const SolverPlanes& S, int s, int kind) {
  stage_prefetch(G, S, s, kind);
  stage_wait();
}
// island_done (b2g_step_kernels.cuh) on the penetrations other SMs published during this launch
__device__ __forceinline__ bool island_done_l2(const uint32_t* islandPen, int penStride, int iter, int root) {
  if (iter == 0) return false;
  float pen = __uint_as_float(__ldcg(islandPen + (size_t)(iter - 1) * penStride + root));
  return pen <= 3.0f * B2G_LINEAR_SLOP;
}

struct BigRanges {
  int first[B2G_MAX_COLOURS + 2];  // first[c]..first[c+1] = slots of colour c; [MAX] = overflow bucket
  int numColours;
};
// first colour >= from with constraints, or limit
__device__ __forceinline__ int big_next_colour(const BigRanges& R, int from, int limit) {
  while (from < limit && R.first[from] == R.first[from + 1]) ++from;
  return from;
}

#define B2G_BIG_WARM 0
#define B2G_BIG_VELOCITY 1
#define B2G_BIG_POSITION 2
struct BigPassArgs {
  const int* croot;
  uint32_t* islandPen;
  int penStride, it;
};
// one constraint of a pass: from the thread's staged copy, or straight from the planes.  Returns the
// penetration (position passes).
template <int PHASE, bool STAGED>
__device__ __forceinline__ float big_visit(BigStage& G, const SolverPlanes& S, int s, const CoherentBodies& velAcc,
                                           const CoherentBodies& posAcc) {
  const int t = threadIdx.x;
  if (STAGED) stage_acquire(G, S, s, PHASE == B2G_BIG_POSITION ? B2G_STAGE_POSITION : B2G_STAGE_VELOCITY);
  if (PHASE == B2G_BIG_WARM) {
    if (STAGED) warm_start_constraint(G.T, t, velAcc);
    else warm_start_constraint(S, s, velAcc);
  } else if (PHASE == B2G_BIG_VELOCITY) {
    if (STAGED) {
      solve_velocity_constraint(G.T, t, velAcc);
      S.imp[s] = G.T.imp[t];  // the staged copy stays current for a thread that revisits the same slot
    } else {
      solve_velocity_constraint(S, s, velAcc);
    }
  } else {
    float minSep = STAGED ? solve_position_constraint(G.T, t, posAcc) : solve_position_constraint(S, s, posAcc);
    return minSep < 0.0f ? -minSep : 0.0f;
  }
  return 0.0f;
}
// this thread's share of one colour: slot sFirst (staged), then sFirst + stride, ... (direct).  Must be
// called by whole warps (the position passes publish one penetration per warp and island root: every
// constraint of a 100 k-body island would otherwise hit the same islandPen word).
template <int PHASE>
__device__ __forceinline__ void big_colour(BigStage& G, const SolverPlanes& S, int sFirst, int s1, int stride,
                                           const CoherentBodies& velAcc, const CoherentBodies& posAcc,
                                           const BigPassArgs& Q) {
  int root = -1;
  float pen = 0.0f;
  if (sFirst < s1) {
    if (PHASE == B2G_BIG_POSITION) {
      root = Q.croot[sFirst];
      if (island_done_l2(Q.islandPen, Q.penStride, Q.it, root)) root = -1;
    }
    if (PHASE != B2G_BIG_POSITION || root >= 0) pen = big_visit<PHASE, true>(G, S, sFirst, velAcc, posAcc);
  }
  if (PHASE == B2G_BIG_POSITION) {
    const unsigned int peers = __match_any_sync(0xffffffffu, root);
    const unsigned int worst = __reduce_max_sync(peers, __float_as_uint(pen));
    if (root >= 0 && worst != 0u && (int)(threadIdx.x & 31) == __ffs(peers) - 1) {
      uint32_t* slot = &Q.islandPen[(size_t)Q.it * Q.penStride + root];
      if (__ldcg(slot) < worst) atomicMax(slot, worst);
    }
  }
  if (stride > 0 && sFirst < s1) {
    for (int s = sFirst + stride; s < s1; s += stride) {
      if (PHASE == B2G_BIG_POSITION) {
        int r2 = Q.croot[s];
        if (island_done_l2(Q.islandPen, Q.penStride, Q.it, r2)) continue;
        float p2 = big_visit<PHASE, false>(G, S, s, velAcc, posAcc);
        if (p2 > 0.0f) atomicMax(&Q.islandPen[(size_t)Q.it * Q.penStride + r2], __float_as_uint(p2));
      } else {
        big_visit<PHASE, false>(G, S, s, velAcc, posAcc);
      }
    }
  }
}

struct BigGrid {
  unsigned int* barrier;
  unsigned int target;
  int gtid, gsize;
};
// One sweep over all colours (one solver iteration of PHASE).  `again` = another sweep over the same
// kind of planes follows (warm start -> velocity iterations, velocity -> velocity, position -> position):
// its first pass is prefetched behind this sweep's last barrier.
template <int PHASE>
__device__ __forceinline__ void big_sweep(BigGrid& Z, BigStage& G, const SolverPlanes& S, const BigRanges& R,
                                          const CoherentBodies& velAcc, const CoherentBodies& posAcc,
                                          const BigPassArgs& Q, bool again) {
  const int kind = PHASE == B2G_BIG_POSITION ? B2G_STAGE_POSITION : B2G_STAGE_VELOCITY;
  const int ct = R.numColours;
  const int cFirst = big_next_colour(R, 0, ct);  // first pass of a sweep (ct when there is none)
  for (int c = cFirst; c < ct;) {
    big_colour<PHASE>(G, S, R.first[c] + Z.gtid, R.first[c + 1], Z.gsize, velAcc, posAcc, Q);
    const int cn = big_next_colour(R, c + 1, ct);
    // what this thread visits next: the next pass, else the first pass of the following sweep
    int sn = 0, sl = 0;
    if (cn < ct) {
      sn = R.first[cn] + Z.gtid;
      sl = R.first[cn + 1];
    } else if (again) {
      sn = R.first[cFirst] + Z.gtid;
      sl = R.first[cFirst + 1];
    }
    grid_arrive(Z.barrier, Z.target);
    if (sn < sl) stage_prefetch(G, S, sn, kind);
    grid_wait(Z.barrier, Z.target);
    c = cn;
  }
  const int ov0 = R.first[B2G_MAX_COLOURS], ov1 = R.first[B2G_MAX_COLOURS + 1];
  if (ov1 > ov0) {  // serial overflow bucket
    if (Z.gtid == 0) {
      for (int s = ov0; s < ov1; ++s) {
        if (PHASE == B2G_BIG_POSITION) {
          int root = Q.croot[s];
          if (island_done_l2(Q.islandPen, Q.penStride, Q.it, root)) continue;
          float pen = big_visit<PHASE, false>(G, S, s, velAcc, posAcc);
          atomicMax(&Q.islandPen[(size_t)Q.it * Q.penStride + root], __float_as_uint(pen));
        } else {
          big_visit<PHASE, false>(G, S, s, velAcc, posAcc);
        }
      }
    }
    grid_arrive(Z.barrier, Z.target);
    grid_wait(Z.barrier, Z.target);
  }
}

__global__ void __launch_bounds__(B2G_BIG_THREADS, 1)
k_big_solve(BigRanges R, SolverPlanes S, ContactBuf C, float4* vel, float4* pos, const int* __restrict__ croot,
            uint32_t* islandPen, int penStride, int nb, const uint32_t* __restrict__ bflags,
            const int* __restrict__ island, const uint32_t* __restrict__ islandAwake,
            const int* __restrict__ bodySlot, float h, int velIters, int posIters, int warmStarting, JointWalk W,
            JointArraysDev J, const float4* __restrict__ mass, const float4* __restrict__ center, float dtRatio,
            unsigned int* barrier) {
  extern __shared__ float4 stageMem[];
  BigGrid Z;
  Z.barrier = barrier;
  Z.target = 0;
  // consecutive groups of 32 constraints go to DIFFERENT blocks (warp w of block b is global warp
  // w * gridDim + b), so a colour of a few thousand constraints keeps one or two warps busy on every SM
  // instead of sixteen warps on a handful of SMs
  Z.gtid = ((threadIdx.x >> 5) * gridDim.x + blockIdx.x) * 32 + (threadIdx.x & 31);
  Z.gsize = gridDim.x * blockDim.x;
  const int gtid = Z.gtid, gsize = Z.gsize;
  const CoherentBodies velAcc{vel};
  const CoherentBodies posAcc{pos};
  const int ov1 = R.first[B2G_MAX_COLOURS + 1];
  BigStage G;
  {
    float4* m = stageMem;
    const int B = B2G_BIG_THREADS;
    G.T.idx = (int4*)m;
    G.T.mass = m + B;
    G.T.nf = m + 2 * B;
    G.T.r1 = m + 3 * B;
    G.T.r2 = m + 4 * B;
    G.T.m1 = m + 5 * B;
    G.T.m2 = m + 6 * B;
    G.T.kk = m + 7 * B;
    G.T.imp = m + 8 * B;
    G.T.pn = m + 2 * B;  // the position planes reuse the velocity slots
    G.T.pp = m + 3 * B;
    G.T.pc = m + 4 * B;
    G.T.pr = m + 5 * B;
    G.staged = -1;
    G.kind = B2G_STAGE_VELOCITY;
  }
  BigPassArgs Q;
  Q.croot = croot;
  Q.islandPen = islandPen;
  Q.penStride = penStride;
  Q.it = 0;

  if (warmStarting) big_sweep<B2G_BIG_WARM>(Z, G, S, R, velAcc, posAcc, Q, velIters > 0);
  const float invH = h > 0.0f ? 1.0f / h : 0.0f;
  auto jsync = [&]() {
    grid_arrive(Z.barrier, Z.target);
    grid_wait(Z.barrier, Z.target);
  };
  if (W.nj > 0) joints_init_coloured<CoherentBodies>(W, J, gtid, gsize, jsync, pos, vel, mass, center, dtRatio, warmStarting);
  for (int it = 0; it < velIters; ++it) {
    if (W.nj > 0) joints_velocity_coloured<CoherentBodies>(W, J, gtid, gsize, jsync, vel, h, invH);
    big_sweep<B2G_BIG_VELOCITY>(Z, G, S, R, velAcc, posAcc, Q, it + 1 < velIters);
  }
  // store impulses (b2_contact_solver.cpp:641-657)
  for (int s = R.first[0] + gtid; s < ov1; s += gsize) {
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
  // integrate positions of the big islands' bodies (b2_island.cpp:353-385)
  for (int b = gtid; b < nb; b += gsize) {
    if (bodySlot[b] != B2G_SLOT_BIG) continue;
    if (!body_simulated(bflags[b], island, islandAwake, b)) continue;
    float4 p4 = posAcc.load(b), v4 = velAcc.load(b);
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
    posAcc.store(b, p4);
    velAcc.store(b, make_float4(v.x, v.y, w, v4.w));
  }
  {
    // behind this barrier: the position planes of the thread's first position pass
    int sn = 0, sl = 0;
    if (posIters > 0) {
      const int cf = big_next_colour(R, 0, R.numColours);
      if (cf < R.numColours) {
        sn = R.first[cf] + gtid;
        sl = R.first[cf + 1];
      }
    }
    grid_arrive(Z.barrier, Z.target);
    if (sn < sl) stage_prefetch(G, S, sn, B2G_STAGE_POSITION);
    grid_wait(Z.barrier, Z.target);
  }
  for (int it = 0; it < posIters; ++it) {
    Q.it = it;
    big_sweep<B2G_BIG_POSITION>(Z, G, S, R, velAcc, posAcc, Q, it + 1 < posIters);
    if (W.nj > 0) joints_position_coloured<CoherentBodies>(W, J, gtid, gsize, jsync, pos, islandPen, penStride, it);
  }
}

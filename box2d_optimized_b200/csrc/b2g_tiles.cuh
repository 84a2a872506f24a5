// b2g_tiles.cuh — oversize islands (a settled 100k-body pile = ONE island) cut into per-SM TILES.
//
// b2Island::Solve (src/dynamics/b2_island.cpp:253-484) sweeps an island's constraints one after the other.
// The graph-coloured solver sweeps them colour by colour; for an island that does not fit one CTA every
// colour used to be a grid-wide pass (k_big_solve: ~13 colours x 12 sweeps = 157 grid barriers per step at
// ~3 us each).  Here the island's bodies are partitioned in space into one tile per SM (S vertical strips of
// equal body count, each cut into R rows of equal body count, S x R <= #SMs, re-planned every few steps from
// two histograms), and one persistent CTA per tile keeps its tile's bodies in SHARED MEMORY for the whole
// solve, exactly as k_solve_bins_fused does for small islands:
//   * a constraint whose non-static bodies sit in one tile is INTERIOR: swept colour by colour with CTA
//     barriers, its constants staged through shared memory one pass ahead (cp.async);
//   * the others are CUT constraints (~5-10 %): they have their own colour domain (2-4 colours) and are swept
//     after the interior ones of each iteration in grid-wide passes over the bodies' global copies; the tiles
//     publish their boundary bodies before and reload them after.
// Grid barriers per step: 2 + (1 + velIters + posIters) x (1 + cut colours) ~ 50 instead of 157.
// Any order of visiting constraints is a Gauss-Seidel order; the arithmetic per constraint is the same device
// code as every other mode (b2g_solver.cuh).
#pragma once
#include "b2g_fused.cuh"

#ifndef B2G_TILES_PER_SM
#define B2G_TILES_PER_SM 2       // persistent CTAs per SM: two half-size tiles hide each other's barrier / L2 waits
#endif
#define B2G_TILES_MAX (160 * B2G_TILES_PER_SM)   // >= SM count x tiles per SM
#define B2G_TILE_CAP (2048 / B2G_TILES_PER_SM)   // bodies a tile can hold in shared memory
#define B2G_TILE_THREADS (512 / B2G_TILES_PER_SM)
#define B2G_TILE_XBINS 4096
#define B2G_TILE_YBINS 8192        // a pile with stragglers high above it: the rows must still resolve the pile itself
#define B2G_TILE_MIN_BODIES 128  // do not cut an island into tiles smaller than this
#define B2G_TILE_PLAN_PERIOD 16  // steps between re-plans (sooner when tiles overflow).  A pile that still grows puts every newcomer above the planned range into the top row: 405 bodies there against 237 elsewhere after 32 steps; balance itself buys little (a tile's sweep time follows its colour count, not its size), headroom against overflow does
#define B2G_TILE_BODY_BYTES (16 + 16 + 4 + 2 + 1 + 1)  // shared memory per tile body: vel, pos, index, boundary list + degree (+ pad)

struct TilePlan {
  unsigned int lo[2], hi[2];  // float_flip()ed bounds of the oversize islands' body centres
  int count;                  // their number when the plan was made
  int S, R;                   // strips x rows
  float x0, invDx, y0, invDy;
};

__device__ __forceinline__ bool tile_is_big_body(int b, const uint32_t* __restrict__ bflags, const int* __restrict__ island,
                                                 const uint32_t* __restrict__ islandAwake,
                                                 const int* __restrict__ islandCount, int bigThreshold) {
  if (B2G_BODY_TYPE(bflags[b]) == B2G_STATIC) return false;
  const int root = island[b];
  return islandAwake[root] != 0 && islandCount[root] > bigThreshold;
}

// ---- planning (every B2G_TILE_PLAN_PERIOD steps) -------------------------------------------------------
__global__ void k_tile_bounds(int nb, const float4* __restrict__ pos, const uint32_t* __restrict__ bflags,
                              const int* __restrict__ island, const uint32_t* __restrict__ islandAwake,
                              const int* __restrict__ islandCount, int bigThreshold, TilePlan* plan) {
  B2G_PDL_ENTER();
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nb || !tile_is_big_body(b, bflags, island, islandAwake, islandCount, bigThreshold)) return;
  const float4 p = pos[b];
  auto g = cg::coalesced_threads();
  const unsigned int fx = float_flip(p.x), fy = float_flip(p.y);
  const unsigned int lx = cg::reduce(g, fx, cg::less<unsigned int>()), hx = cg::reduce(g, fx, cg::greater<unsigned int>());
  const unsigned int ly = cg::reduce(g, fy, cg::less<unsigned int>()), hy = cg::reduce(g, fy, cg::greater<unsigned int>());
  if (g.thread_rank() == 0) {
    atomicMin(&plan->lo[0], lx);
    atomicMax(&plan->hi[0], hx);
    atomicMin(&plan->lo[1], ly);
    atomicMax(&plan->hi[1], hy);
    atomicAdd(&plan->count, (int)g.size());
  }
}

// strips x rows: as many tiles as there are SMs, as square as the pile's aspect ratio allows
__global__ void k_tile_plan_begin(TilePlan* plan, int maxTiles) {
  B2G_PDL_ENTER();
  if (blockIdx.x != 0 || threadIdx.x >= 32) return;  // one warp: the lanes share the candidate strip counts
  const int count = plan->count;
  float x0 = float_unflip(plan->lo[0]), x1 = float_unflip(plan->hi[0]);
  float y0 = float_unflip(plan->lo[1]), y1 = float_unflip(plan->hi[1]);
  if (count <= 0) {
    x0 = y0 = 0.0f;
    x1 = y1 = 1.0f;
  }
  const float w = fmaxf(x1 - x0, 1.0f), h = fmaxf(y1 - y0, 1.0f);
  int T = count / B2G_TILE_MIN_BODIES;
  T = T < 1 ? 1 : (T > maxTiles ? maxTiles : T);
  int bestS = 1, bestR = 1;
  float bestScore = -1e30f;
  for (int S = 1 + (int)threadIdx.x; S <= T; S += 32) {
    const int R = T / S;
    const float used = (float)(S * R) / (float)T;
    const float aspect = (w / (float)S) / (h / (float)R);
    const float score = -fabsf(log2f(aspect)) - 8.0f * (1.0f - used);
    if (score > bestScore) {
      bestScore = score;
      bestS = S;
      bestR = R;
    }
  }
  for (int o = 16; o > 0; o >>= 1) {  // best score, the smaller strip count on a tie (what the serial scan kept)
    const float os = __shfl_xor_sync(0xffffffffu, bestScore, o);
    const int oS = __shfl_xor_sync(0xffffffffu, bestS, o), oR = __shfl_xor_sync(0xffffffffu, bestR, o);
    if (os > bestScore || (os == bestScore && oS < bestS)) {
      bestScore = os;
      bestS = oS;
      bestR = oR;
    }
  }
  if (threadIdx.x != 0) return;
  plan->S = bestS;
  plan->R = bestR;
  plan->x0 = x0;
  plan->invDx = (float)B2G_TILE_XBINS / w;
  plan->y0 = y0;
  plan->invDy = (float)B2G_TILE_YBINS / h;
}
__device__ __forceinline__ int tile_xbin(const TilePlan* P, float x) {
  int k = (int)((x - P->x0) * P->invDx);
  return k < 0 ? 0 : (k >= B2G_TILE_XBINS ? B2G_TILE_XBINS - 1 : k);
}
__device__ __forceinline__ int tile_ybin(const TilePlan* P, float y) {
  int k = (int)((y - P->y0) * P->invDy);
  return k < 0 ? 0 : (k >= B2G_TILE_YBINS ? B2G_TILE_YBINS - 1 : k);
}
__global__ void k_tile_xhist(int nb, const float4* __restrict__ pos, const uint32_t* __restrict__ bflags,
                             const int* __restrict__ island, const uint32_t* __restrict__ islandAwake,
                             const int* __restrict__ islandCount, int bigThreshold, const TilePlan* __restrict__ plan,
                             int* histX) {
  B2G_PDL_ENTER();
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nb || !tile_is_big_body(b, bflags, island, islandAwake, islandCount, bigThreshold)) return;
  atomicAdd(&histX[tile_xbin(plan, pos[b].x)], 1);
}
// exclusive prefix of `n` (<= 8192) counts by one block of 1024 threads; out[k] = min(parts - 1, prefix * parts / total)
__device__ __forceinline__ void tile_partition(const int* __restrict__ hist, int n, int parts, int* out) {
  __shared__ int warpSums[32];
  const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
  const int per = (n + 1023) / 1024;  // <= 8
  int v[8], sum = 0;
  for (int k = 0; k < 8; ++k) {
    const int i = t * per + k;
    v[k] = (k < per && i < n) ? hist[i] : 0;
    sum += v[k];
  }
  int x = sum;
  for (int o = 1; o < 32; o <<= 1) {
    int y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += y;
  }
  if (lane == 31) warpSums[wid] = x;
  __syncthreads();
  if (wid == 0) {
    int ws = warpSums[lane];
    for (int o = 1; o < 32; o <<= 1) {
      int y = __shfl_up_sync(0xffffffffu, ws, o);
      if (lane >= o) ws += y;
    }
    warpSums[lane] = ws;
  }
  __syncthreads();
  const int total = warpSums[31];
  int prefix = (wid > 0 ? warpSums[wid - 1] : 0) + x - sum;
  for (int k = 0; k < 8; ++k) {
    const int i = t * per + k;
    if (k < per && i < n) {
      int p = total > 0 ? (int)(((long long)prefix * parts) / total) : 0;
      out[i] = p >= parts ? parts - 1 : p;
    }
    prefix += v[k];
  }
  __syncthreads();
}
__global__ void __launch_bounds__(1024) k_tile_xplan(const TilePlan* __restrict__ plan, const int* __restrict__ histX,
                                                     int* stripOfX) {
  B2G_PDL_ENTER();
  tile_partition(histX, B2G_TILE_XBINS, plan->S, stripOfX);
}
__global__ void k_tile_yhist(int nb, const float4* __restrict__ pos, const uint32_t* __restrict__ bflags,
                             const int* __restrict__ island, const uint32_t* __restrict__ islandAwake,
                             const int* __restrict__ islandCount, int bigThreshold, const TilePlan* __restrict__ plan,
                             const int* __restrict__ stripOfX, int* histY) {
  B2G_PDL_ENTER();
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nb || !tile_is_big_body(b, bflags, island, islandAwake, islandCount, bigThreshold)) return;
  const float4 p = pos[b];
  const int strip = stripOfX[tile_xbin(plan, p.x)];
  atomicAdd(&histY[strip * B2G_TILE_YBINS + tile_ybin(plan, p.y)], 1);
}
__global__ void __launch_bounds__(1024) k_tile_yplan(const TilePlan* __restrict__ plan, const int* __restrict__ histY,
                                                     int* rowOfY) {
  B2G_PDL_ENTER();
  const int strip = blockIdx.x;
  if (strip >= plan->S) return;
  tile_partition(histY + strip * B2G_TILE_YBINS, B2G_TILE_YBINS, plan->R, rowOfY + strip * B2G_TILE_YBINS);
}

// ---- every step: body -> tile slot -------------------------------------------------------------------
// Slot order inside a tile is whatever the atomics give; nothing depends on it (constraints of one colour
// never share a movable body, per-island reductions are max / min).
__global__ void k_tile_assign(int nb, const float4* __restrict__ pos, const uint32_t* __restrict__ bflags,
                              const int* __restrict__ island, const uint32_t* __restrict__ islandAwake,
                              const int* __restrict__ islandCount, int bigThreshold, const TilePlan* __restrict__ plan,
                              const int* __restrict__ stripOfX, const int* __restrict__ rowOfY, int* tileSlot,
                              int* tileBodies, int* tileCount, int* spillList, StepCounts* counts) {
  B2G_PDL_ENTER();
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nb) return;
  int ts = -1;
  if (tile_is_big_body(b, bflags, island, islandAwake, islandCount, bigThreshold)) {
    const float4 p = pos[b];
    const int strip = stripOfX[tile_xbin(plan, p.x)];
    const int row = rowOfY[strip * B2G_TILE_YBINS + tile_ybin(plan, p.y)];
    const int tile = strip * plan->R + row;
    const int l = atomicAdd(&tileCount[tile], 1);
    if (l + 1 > counts->maxTileCount) atomicMax(&counts->maxTileCount, l + 1);
    if (l < B2G_TILE_CAP) {
      ts = tile * B2G_TILE_CAP + l;
      tileBodies[ts] = b;
    } else {
      spillList[atomicAdd(&counts->spillCount, 1)] = b;  // lives in global memory only; its constraints are cut
    }
  }
  tileSlot[b] = ts;
}
// A tile that received more bodies than it can hold gives ALL of them up (they live in global memory for this
// step and their constraints are cut constraints): which bodies came first is decided by atomics, and the result
// of a step must not depend on that.  The host re-plans long before a tile fills up, so this is a safety net.
__global__ void k_tile_overflow_fix(int nb, int* tileSlot, const int* __restrict__ tileCount, int* spillList,
                                    StepCounts* counts) {
  B2G_PDL_ENTER();
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nb) return;
  const int ts = tileSlot[b];
  if (ts >= 0 && tileCount[ts / B2G_TILE_CAP] > B2G_TILE_CAP) {
    tileSlot[b] = -1;
    spillList[atomicAdd(&counts->spillCount, 1)] = b;
  }
}
// bodies of oversize islands that carry a joint are exchanged through global memory like boundary bodies
__global__ void k_tile_joint_marks(int nj, const int2* __restrict__ jBodies, const int* __restrict__ tileSlot,
                                   uint8_t* tileBoundary, StepCounts* counts) {
  B2G_PDL_ENTER();
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= nj) return;
  const int2 bd = jBodies[j];
  const int sa = tileSlot[bd.x], sb = tileSlot[bd.y];
  if (sa >= 0) tileBoundary[sa] = 1;
  if (sb >= 0) tileBoundary[sb] = 1;
  if (sa >= 0 || sb >= 0) atomicAdd(&counts->bigJoints, 1);
}

// ---- the persistent tile kernel ---------------------------------------------------------------------------
#ifdef B2G_BIG_TRACE  // debug build: phase marks of every block (scripts/gpu_big_trace.py)
#define B2G_TILE_MARKS 64
__device__ unsigned long long g_tileMarks[B2G_TILES_MAX * B2G_TILE_MARKS];
#define TILE_MARK(id)                                                                        \
  do {                                                                                       \
    __syncthreads();                                                                         \
    if (threadIdx.x == 0 && (id) < B2G_TILE_MARKS) g_tileMarks[blockIdx.x * B2G_TILE_MARKS + (id)] = trace_now(); \
  } while (0)
#else
#define TILE_MARK(id) do {} while (0)
#endif
struct TileArgs {
  int tileBin0, cutBin, nb, nj;
  float h, invH, dtRatio;
  float2 gravity;
  int velIters, posIters, warmStarting, allowSleep, clearForces;
  int noSequencing;  // B2G_TILE_BARRIERS=1: always use grid barriers between the cut colours (measurements)
  const TilePlan* plan;
  const int* tileCount;
  const int* tileBodies;
  const uint8_t* tileBoundary;
  const int* tileSlot;
  const int* spillList;
  const int* bucketStart;
  int* sortedList;
  int* orderScratch;
  int* croot;
  int2* cutSeq;
  const unsigned long long* colourMask;
  uint32_t* islandPen;
  int penStride;
  uint32_t* islandMinSleep;
  uint32_t* bflags;
  float4 *pos, *vel, *xf, *force;
  const float4 *mass, *center;
  const float* fRadius;
  const int* island;
  const uint32_t* islandAwake;
  const int* bodySlot;
  StepCounts* counts;
  unsigned int* barrier;
};

// body state of a tile: index >= 0 is a slot of the CTA's shared-memory tile, index < 0 is ~globalIndex of a
// static body (read only)
struct TileSmemBodies {
  float4* tile;
  const float4* global;
  __device__ __forceinline__ float4 load(int i) const { return i >= 0 ? tile[i] : global[~i]; }
  __device__ __forceinline__ void store(int i, float4 v) const { tile[i] = v; }
};

// ---- cut constraints without grid barriers: SEQUENCED bodies ------------------------------------------------
// During the solve the unused fourth lane of a boundary body's global velocity (position) record carries a
// sequence number, written together with the state by one aligned 16-byte store.  A body that carries `deg` cut
// constraints goes, in sweep q, through the values
//     q (deg + 1)            published by its tile after the interior colours of sweep q
//     q (deg + 1) + r + 1    left by its cut constraint of rank r (rank = order of the constraint's colour
//                            among the body's cut colours: colour order is the Gauss-Seidel order)
// so a cut constraint simply waits until both of its bodies carry the number that says "your turn", and the
// tile takes a body back when it reads q (deg + 1) + deg.  Every wait is on something that happens earlier in
// (sweep, colour) order and all CTAs are co-resident, so nothing can wait in a circle; tiles that do not touch
// each other never wait for each other at all.  (A body that did not fit a tile has no publisher: the last
// cut constraint of a sweep leaves the next sweep's base.)
__device__ __forceinline__ float4 ld_body_relaxed(const float4* p) {
  float4 v;
  asm volatile("ld.relaxed.gpu.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_body_relaxed(float4* p, float4 v) {
  asm volatile("st.relaxed.gpu.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
#define B2G_SEQ_MARKER (-1000)  // published before a phase's first sweep: never a valid turn
struct SeqBodies {
  float4* a;
  int ia, ib;      // the constraint's bodies (global indices)
  int ea, eb;      // the numbers they must carry before the constraint may read them (-1: body is never written)
  int na, nb;      // the numbers to leave behind
  float4 va, vb;   // their records, once it is this constraint's turn on both (acquire())
  // both bodies are polled together: one L2 round trip per look, not one per body
  __device__ __forceinline__ void acquire() {
    va = ld_body_relaxed(a + ia);
    vb = ld_body_relaxed(a + ib);
    for (;;) {
      const bool okA = ea < 0 || __float_as_int(va.w) == ea;
      const bool okB = eb < 0 || __float_as_int(vb.w) == eb;
      if (okA && okB) break;
      if (!okA) va = ld_body_relaxed(a + ia);
      if (!okB) vb = ld_body_relaxed(a + ib);
    }
  }
  __device__ __forceinline__ float4 load(int i) const { return i == ia ? va : vb; }
  __device__ __forceinline__ void store(int i, float4 v) const {
    v.w = __int_as_float(i == ia ? na : nb);
    st_body_relaxed(a + i, v);
  }
  // the constraint is skipped (its island has converged): hand the bodies on unchanged
  __device__ __forceinline__ void pass() const {
    if (ea >= 0) store(ia, va);
    if (eb >= 0) store(ib, vb);
  }
};
// per cut constraint and body: rank | deg << 8 | (no tile publishes this body) << 16, or -1 for a body no
// constraint ever writes
__device__ __forceinline__ int seq_pack(unsigned long long mask, int colourBit, bool movable, bool unowned) {
  if (!movable) return -1;
  const unsigned int m = (unsigned int)(mask >> B2G_CUT_DOMAIN_SHIFT) & ((1u << B2G_MAX_COLOURS) - 1u);
  const int c = colourBit - B2G_CUT_DOMAIN_SHIFT;
  const int rank = __popc(m & ((1u << c) - 1u)), deg = __popc(m);
  return rank | (deg << 8) | (unowned ? 1 << 16 : 0);
}
__device__ __forceinline__ SeqBodies seq_bodies(float4* arr, int ia, int ib, int2 packed, int q) {
  SeqBodies B;
  B.a = arr;
  B.ia = ia;
  B.ib = ib;
  auto one = [&](int p, int& e, int& n) {
    if (p < 0) {
      e = n = -1;
      return;
    }
    const int rank = p & 0xff, deg = (p >> 8) & 0xff;
    e = q * (deg + 1) + rank;
    n = e + 1;
    if ((p >> 16) && rank == deg - 1) n = (q + 1) * (deg + 1);
  };
  one(packed.x, B.ea, B.na);
  one(packed.y, B.eb, B.nb);
  return B;
}

#define B2G_TILE_ROOTS 4
struct TileCtx {
  int tid, nt, gtid, gsize;
  int nbod, nBnd;
  int first;                 // tile * B2G_TILE_CAP
  float4 *vel, *pos;         // shared
  int* body;                 // shared
  unsigned short* bnd;       // shared: local slots of the boundary bodies
  unsigned char* bndDeg;     // shared: their cut degrees
  const int* cstart;         // shared [B2G_MAX_COLOURS + 2]: interior ranges of this tile
  const int* cut;            // shared [B2G_MAX_COLOURS + 2]: cut ranges (the same for every tile)
  const int *usedS0, *usedS1;
  int nUsed;
  const int *cutS0, *cutS1;
  int nCut;
  bool anyCut;               // cut colours or a cut overflow bucket exist
  bool seqMode;              // cut colours sequenced through the bodies instead of grid barriers
  int seqV, seqP;            // sweeps of each kind made so far
  // the (few) island roots of this tile's bodies, their "converged" flags and this iteration's penetration
  int* roots;                // shared [B2G_TILE_ROOTS], -1 = free
  int* doneS;                // shared
  unsigned int* penS;        // shared
  unsigned int* barrier;
  unsigned int target;
};
__device__ __forceinline__ void tile_grid_sync(TileCtx& X) {
  grid_arrive(X.barrier, X.target);
  grid_wait(X.barrier, X.target);
}
__device__ __forceinline__ int tile_root_index(const TileCtx& X, int root) {
#pragma unroll
  for (int k = 0; k < B2G_TILE_ROOTS; ++k)
    if (X.roots[k] == root) return k;
  return -1;
}
__device__ __forceinline__ bool tile_root_done(const TileCtx& X, const BigPassArgs& Q, int root, int k) {
  return k >= 0 ? X.doneS[k] != 0 : island_done_l2(Q.islandPen, Q.penStride, Q.it, root);
}
// penetration of a position visit: one shared-memory atomic per warp and island root (global when the root is
// not one of the tile's own)
__device__ __forceinline__ void tile_note_pen(const TileCtx& X, const BigPassArgs& Q, int root, int k, float pen) {
  const unsigned int peers = __match_any_sync(0xffffffffu, root);
  const unsigned int worst = __reduce_max_sync(peers, __float_as_uint(pen));
  if (root >= 0 && worst != 0u && (int)(threadIdx.x & 31) == __ffs(peers) - 1) {
    if (k >= 0) atomicMax(&X.penS[k], worst);
    else atomicMax(&Q.islandPen[(size_t)Q.it * Q.penStride + root], worst);
  }
}
// start of a position iteration (after a grid barrier): which of the tile's islands have converged
__device__ __forceinline__ void tile_roots_begin(const TileCtx& X, const BigPassArgs& Q) {
  if (X.tid < B2G_TILE_ROOTS) {
    const int root = X.roots[X.tid];
    X.doneS[X.tid] = root >= 0 && island_done_l2(Q.islandPen, Q.penStride, Q.it, root) ? 1 : 0;
    X.penS[X.tid] = 0u;
  }
  __syncthreads();
}
// before the grid barrier that ends a position iteration
__device__ __forceinline__ void tile_roots_flush(const TileCtx& X, const BigPassArgs& Q) {
  __syncthreads();
  if (X.tid < B2G_TILE_ROOTS) {
    const int root = X.roots[X.tid];
    const unsigned int w = X.penS[X.tid];
    if (root >= 0 && w != 0u) atomicMax(&Q.islandPen[(size_t)Q.it * Q.penStride + root], w);
  }
}

// one constraint of a pass.  STAGED: the constants are (or are made) present in the thread's shared-memory
// slots; else they are read from the planes.  Returns the penetration (position passes, -1 = skipped).
template <int PHASE, bool STAGED, class VelAcc, class PosAcc>
__device__ __forceinline__ float tile_visit(const TileCtx& X, BigStage& G, const SolverPlanes& S, int s, const VelAcc& velAcc,
                                            const PosAcc& posAcc, const BigPassArgs& Q, int& root, int& rootIdx) {
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
    root = __float_as_int(STAGED ? G.T.pr[t].w : S.pr[s].w);
    rootIdx = tile_root_index(X, root);
    if (tile_root_done(X, Q, root, rootIdx)) return -1.0f;
    float minSep = STAGED ? solve_position_constraint(G.T, t, posAcc) : solve_position_constraint(S, s, posAcc);
    return minSep < 0.0f ? -minSep : 0.0f;
  }
  return 0.0f;
}
// this thread's share of one INTERIOR colour (or of a cut colour in barrier mode): slot sFirst (staged), then
// sFirst + stride, ... (direct).  Called by whole warps.
template <int PHASE, class VelAcc, class PosAcc>
__device__ __forceinline__ void tile_colour(const TileCtx& X, BigStage& G, const SolverPlanes& S, int sFirst, int s1, int stride,
                                            const VelAcc& velAcc, const PosAcc& posAcc, const BigPassArgs& Q) {
  int root = -1, k = -1;
  float pen = 0.0f;
  if (sFirst < s1) {
    pen = tile_visit<PHASE, true>(X, G, S, sFirst, velAcc, posAcc, Q, root, k);
    if (pen < 0.0f) {
      pen = 0.0f;
      root = -1;
    }
  }
  if (PHASE == B2G_BIG_POSITION) tile_note_pen(X, Q, root, k, pen);
  if (stride > 0 && sFirst < s1) {
    for (int s = sFirst + stride; s < s1; s += stride) {
      int r2 = -1, k2 = -1;
      const float p2 = tile_visit<PHASE, false>(X, G, S, s, velAcc, posAcc, Q, r2, k2);
      if (PHASE == B2G_BIG_POSITION && p2 > 0.0f) {
        if (k2 >= 0) atomicMax(&X.penS[k2], __float_as_uint(p2));
        else atomicMax(&Q.islandPen[(size_t)Q.it * Q.penStride + r2], __float_as_uint(p2));
      }
    }
  }
}

__device__ __forceinline__ void tile_publish(const TileCtx& X, const float4* sm, float4* g) {
  for (int k = X.tid; k < X.nBnd; k += X.nt) {
    const int l = X.bnd[k];
    __stcg(g + X.body[l], sm[l]);
  }
}
__device__ __forceinline__ void tile_reload(const TileCtx& X, float4* sm, const float4* g) {
  for (int k = X.tid; k < X.nBnd; k += X.nt) {
    const int l = X.bnd[k];
    sm[l] = __ldcg(g + X.body[l]);
  }
  __syncthreads();
}
// sequenced mode: publish with the sweep's base number / take the bodies back when their last cut constraint
// of the sweep has left its number
__device__ __forceinline__ void tile_publish_seq(const TileCtx& X, const float4* sm, float4* g, int q, bool marker) {
  for (int k = X.tid; k < X.nBnd; k += X.nt) {
    const int l = X.bnd[k];
    float4 v = sm[l];
    v.w = __int_as_float(marker ? B2G_SEQ_MARKER : q * ((int)X.bndDeg[k] + 1));
    st_body_relaxed(g + X.body[l], v);
  }
}
__device__ __forceinline__ void tile_reload_seq(const TileCtx& X, float4* sm, const float4* g, int q) {
  for (int k = X.tid; k < X.nBnd; k += X.nt) {
    const int l = X.bnd[k];
    const int deg = X.bndDeg[k];
    const int want = q * (deg + 1) + deg;
    const float4* p = g + X.body[l];
    float4 v = ld_body_relaxed(p);
    while (__float_as_int(v.w) != want) v = ld_body_relaxed(p);
    sm[l] = v;
  }
  __syncthreads();
}

// One solver iteration of PHASE: the tile's interior colours with CTA barriers, then the cut colours over the
// bodies' global copies — sequenced through the bodies themselves, or (joints in the island, a cut constraint
// without a colour) with grid barriers — then the boundary bodies come back.
template <int PHASE>
__device__ __forceinline__ void tile_sweep(TileCtx& X, BigStage& G, const SolverPlanes& S, const TileArgs& A,
                                           const BigPassArgs& Q, bool again) {
  const int kind = PHASE == B2G_BIG_POSITION ? B2G_STAGE_POSITION : B2G_STAGE_VELOCITY;
  const TileSmemBodies velT{X.vel, A.vel}, posT{X.pos, A.pos};
  const CoherentBodies velG{A.vel}, posG{A.pos};
  if (PHASE == B2G_BIG_POSITION) tile_roots_begin(X, Q);
  // ---- interior
  for (int k = 0; k < X.nUsed; ++k) {
    tile_colour<PHASE>(X, G, S, X.usedS0[k] + X.tid, X.usedS1[k], X.nt, velT, posT, Q);
    // what this thread visits next: the next interior colour, else its first cut constraint, else the first
    // interior colour of the following sweep
    int sn = 0, sl = 0;
    if (k + 1 < X.nUsed) {
      sn = X.usedS0[k + 1] + X.tid;
      sl = X.usedS1[k + 1];
    } else if (X.nCut > 0) {
      sn = (X.seqMode ? X.cut[0] : X.cutS0[0]) + X.gtid;
      sl = X.seqMode ? X.cut[B2G_MAX_COLOURS] : X.cutS1[0];
    } else if (again) {
      sn = X.usedS0[0] + X.tid;
      sl = X.usedS1[0];
    }
    if (sn < sl) stage_prefetch(G, S, sn, kind);
    __syncthreads();
  }
  {
    const int o0 = X.cstart[B2G_MAX_COLOURS], o1 = X.cstart[B2G_MAX_COLOURS + 1];
    if (o1 > o0) {  // the tile's serial bucket: one thread, key order
      if (X.tid == 0) {
        for (int s = o0; s < o1; ++s) {
          int r = -1, k = -1;
          const float pen = tile_visit<PHASE, false>(X, G, S, s, velT, posT, Q, r, k);
          if (PHASE == B2G_BIG_POSITION && pen > 0.0f) {
            if (k >= 0) atomicMax(&X.penS[k], __float_as_uint(pen));
            else atomicMax(&Q.islandPen[(size_t)Q.it * Q.penStride + r], __float_as_uint(pen));
          }
        }
      }
      __syncthreads();
    }
  }
  float4* sm = PHASE == B2G_BIG_POSITION ? X.pos : X.vel;
  float4* gl = PHASE == B2G_BIG_POSITION ? A.pos : A.vel;
  TILE_MARK(44 + (PHASE == B2G_BIG_POSITION ? 9 + X.seqP : X.seqV));  // interior part of this sweep done
  // ---- cut
  if (X.anyCut && X.seqMode) {
    const int q = PHASE == B2G_BIG_POSITION ? X.seqP++ : X.seqV++;
    tile_publish_seq(X, sm, gl, q, false);
    if (PHASE == B2G_BIG_VELOCITY && q == 5) TILE_MARK(56);  // (trace build) boundary bodies published
    // every cut constraint has its own thread (slot cut[0] + gtid; a second one, + gsize, only beyond ~75 k
    // cut constraints, in slot = colour order), so all colours wait for their turn at the same time and the
    // critical path is the depth of the turn order, not a thread's list
    {
      const int c0 = X.cut[0], c1 = X.cut[B2G_MAX_COLOURS];
      bool first = true;
      for (int s = c0 + X.gtid; s < c1; s += X.gsize, first = false) {
        if (first) stage_acquire(G, S, s, kind);
        const int4 ix = first ? G.T.idx[threadIdx.x] : S.idx[s];
        SeqBodies acc = seq_bodies(gl, ix.x, ix.y, A.cutSeq[s], q);
        int r = -1, kk = -1;
        bool skip = false;
        if (PHASE == B2G_BIG_POSITION) {
          r = __float_as_int(first ? G.T.pr[threadIdx.x].w : S.pr[s].w);
          kk = tile_root_index(X, r);
          skip = tile_root_done(X, Q, r, kk);
        }
        acc.acquire();
        if (skip) {
          acc.pass();  // converged island: the bodies still have to move on
        } else {
          float pen;
          if (first) pen = tile_visit<PHASE, true>(X, G, S, s, acc, acc, Q, r, kk);
          else pen = tile_visit<PHASE, false>(X, G, S, s, acc, acc, Q, r, kk);
          if (PHASE == B2G_BIG_POSITION && pen > 0.0f) {
            if (kk >= 0) atomicMax(&X.penS[kk], __float_as_uint(pen));
            else atomicMax(&Q.islandPen[(size_t)Q.it * Q.penStride + r], __float_as_uint(pen));
          }
        }
      }
      if (again && X.nUsed > 0 && X.usedS0[0] + X.tid < X.usedS1[0]) stage_prefetch(G, S, X.usedS0[0] + X.tid, kind);
    }
    if (PHASE == B2G_BIG_VELOCITY && q == 5) TILE_MARK(57);  // thread 0's cut constraints done
    tile_reload_seq(X, sm, gl, q);
    if (PHASE == B2G_BIG_VELOCITY && q == 5) TILE_MARK(58);  // boundary bodies back
    if (PHASE == B2G_BIG_POSITION) {
      tile_roots_flush(X, Q);
      tile_grid_sync(X);  // the next iteration's early-exit test reads every tile's penetration
    }
  } else if (X.anyCut) {
    tile_publish(X, sm, gl);
    tile_grid_sync(X);
    for (int k = 0; k < X.nCut; ++k) {
      tile_colour<PHASE>(X, G, S, X.cutS0[k] + X.gtid, X.cutS1[k], X.gsize, velG, posG, Q);
      int sn = 0, sl = 0;
      if (k + 1 < X.nCut) {
        sn = X.cutS0[k + 1] + X.gtid;
        sl = X.cutS1[k + 1];
      } else if (again && X.nUsed > 0) {
        sn = X.usedS0[0] + X.tid;
        sl = X.usedS1[0];
      }
      const int o0 = X.cut[B2G_MAX_COLOURS], o1 = X.cut[B2G_MAX_COLOURS + 1];
      if (PHASE == B2G_BIG_POSITION && k + 1 == X.nCut && o1 == o0) tile_roots_flush(X, Q);
      grid_arrive(X.barrier, X.target);
      if (sn < sl) stage_prefetch(G, S, sn, kind);
      grid_wait(X.barrier, X.target);
    }
    const int o0 = X.cut[B2G_MAX_COLOURS], o1 = X.cut[B2G_MAX_COLOURS + 1];
    if (o1 > o0 || X.nCut == 0) {  // serial bucket of the cut domain: one thread of the grid
      if (X.gtid == 0) {
        for (int s = o0; s < o1; ++s) {
          int r = -1, k = -1;
          const float pen = tile_visit<PHASE, false>(X, G, S, s, velG, posG, Q, r, k);
          if (PHASE == B2G_BIG_POSITION && pen > 0.0f) {
            if (k >= 0) atomicMax(&X.penS[k], __float_as_uint(pen));
            else atomicMax(&Q.islandPen[(size_t)Q.it * Q.penStride + r], __float_as_uint(pen));
          }
        }
      }
      if (PHASE == B2G_BIG_POSITION) tile_roots_flush(X, Q);
      tile_grid_sync(X);
    }
    tile_reload(X, sm, gl);
  } else if (PHASE == B2G_BIG_POSITION) {
    tile_roots_flush(X, Q);
    tile_grid_sync(X);  // the next iteration's early-exit test reads every tile's penetration
  }
}

// joints of oversize islands: one thread of the grid walks them over the global body copies (the joints'
// bodies are boundary bodies, see k_tile_joint_marks), between a publish and a reload
template <class F>
__device__ __forceinline__ void tile_joint_phase(TileCtx& X, float4* sm, float4* gl, F&& walk) {
  tile_publish(X, sm, gl);
  tile_grid_sync(X);
  walk([&]() { tile_grid_sync(X); });  // colour by colour over the grid, a grid barrier behind each colour
  tile_reload(X, sm, gl);
}

__global__ void __launch_bounds__(B2G_TILE_THREADS, B2G_TILES_PER_SM)
k_big_tiles(TileArgs A, SolverPlanes S, ContactBuf C, JointWalk W, JointArraysDev J) {
  extern __shared__ __align__(16) unsigned char tileSmem[];
  __shared__ int cstart[B2G_MAX_COLOURS + 2], cut[B2G_MAX_COLOURS + 2];
  __shared__ int usedS0[B2G_MAX_COLOURS], usedS1[B2G_MAX_COLOURS], cutS0[B2G_MAX_COLOURS], cutS1[B2G_MAX_COLOURS];
  __shared__ int sCounts[4];  // nUsed, nCut, nBnd, -
  __shared__ int sRoots[B2G_TILE_ROOTS], sDone[B2G_TILE_ROOTS];
  __shared__ unsigned int sPen[B2G_TILE_ROOTS];

  TileCtx X;
  X.tid = threadIdx.x;
  X.nt = blockDim.x;
  // consecutive groups of 32 cut constraints go to DIFFERENT blocks (warp w of block b is global warp
  // w * gridDim + b): a cut colour of a few thousand constraints keeps a warp or two busy on every SM
  X.gtid = ((threadIdx.x >> 5) * gridDim.x + blockIdx.x) * 32 + (threadIdx.x & 31);
  X.gsize = gridDim.x * blockDim.x;
  X.barrier = A.barrier;
  X.target = 0;
  X.seqV = X.seqP = 0;
  X.roots = sRoots;
  X.doneS = sDone;
  X.penS = sPen;
  const int tid = X.tid, nt = X.nt;
  const int numTiles = A.plan->S * A.plan->R;
  const int tile = blockIdx.x;
  X.first = tile * B2G_TILE_CAP;
  {
    int n = tile < numTiles ? A.tileCount[tile] : 0;
    X.nbod = n > B2G_TILE_CAP ? 0 : n;  // an overfull tile has given all its bodies up (k_tile_overflow_fix)
  }
  const int nbod = X.nbod;
  {
    unsigned char* p = tileSmem;
    X.vel = (float4*)p;            p += (size_t)B2G_TILE_CAP * 16;
    X.pos = (float4*)p;            p += (size_t)B2G_TILE_CAP * 16;
    X.body = (int*)p;              p += (size_t)B2G_TILE_CAP * 4;
    X.bnd = (unsigned short*)p;    p += (size_t)B2G_TILE_CAP * 2;
    X.bndDeg = p;                  p += (size_t)B2G_TILE_CAP;
  }
  BigStage G;
  {
    float4* m = (float4*)(tileSmem + (size_t)B2G_TILE_CAP * B2G_TILE_BODY_BYTES);
    const int B = B2G_TILE_THREADS;
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
  if (tid <= B2G_MAX_COLOURS + 1) {
    cstart[tid] = tile < numTiles ? A.bucketStart[((A.tileBin0 + tile) << B2G_COLOUR_BITS) + tid] : 0;
    cut[tid] = A.bucketStart[(A.cutBin << B2G_COLOUR_BITS) + tid];
  }
  if (tid < 4) sCounts[tid] = 0;
  if (tid < B2G_TILE_ROOTS) sRoots[tid] = -1, sDone[tid] = 0, sPen[tid] = 0u;
  __syncthreads();
  if (tid == 0) {
    int k = 0, m = 0;
    for (int c = 0; c < B2G_MAX_COLOURS; ++c) {
      if (cstart[c] != cstart[c + 1]) {
        usedS0[k] = cstart[c];
        usedS1[k] = cstart[c + 1];
        ++k;
      }
      if (cut[c] != cut[c + 1]) {
        cutS0[m] = cut[c];
        cutS1[m] = cut[c + 1];
        ++m;
      }
    }
    sCounts[0] = k;
    sCounts[1] = m;
  }
  X.cstart = cstart;
  X.cut = cut;
  X.usedS0 = usedS0;
  X.usedS1 = usedS1;
  X.cutS0 = cutS0;
  X.cutS1 = cutS1;

  const float h = A.h;
  const int nSpill = A.counts->spillCount;
  const bool bigJoints = A.nj > 0 && A.counts->bigJoints > 0;
  TILE_MARK(0);
  // ---- phase 0: load the tile, integrate velocities (b2_island.cpp:257-293) ------------------------------
  auto integrate_velocity = [&](uint32_t f, float4 v4, int b) {
    if (B2G_BODY_TYPE(f) == B2G_DYNAMIC) {
      float4 m4 = A.mass[b], c4 = A.center[b], f4 = A.force[b];
      float t = h * m4.x;
      float gs = m4.w * m4.z;
      v4.x += t * (gs * A.gravity.x + f4.x);
      v4.y += t * (gs * A.gravity.y + f4.y);
      v4.z += h * m4.y * f4.z;
      float dl = 1.0f + h * c4.z;
      float da = 1.0f + h * c4.w;
      v4.x /= dl;
      v4.y /= dl;
      v4.z /= da;
    }
    return v4;
  };
  for (int l = tid; l < nbod; l += nt) {
    const int b = A.tileBodies[X.first + l];
    X.body[l] = b;
    const uint32_t f = A.bflags[b];
    if (!(f & B2G_BODY_AWAKE)) A.bflags[b] = f | B2G_BODY_AWAKE;  // reached bodies are woken, timer kept
    X.vel[l] = integrate_velocity(f, A.vel[b], b);
    X.pos[l] = A.pos[b];
    if (A.tileBoundary[X.first + l]) {
      const int k = atomicAdd(&sCounts[2], 1);
      X.bnd[k] = (unsigned short)l;
      X.bndDeg[k] = (unsigned char)__popc((unsigned int)(A.colourMask[b] >> B2G_CUT_DOMAIN_SHIFT) & ((1u << B2G_MAX_COLOURS) - 1u));
    }
    // the (few) islands this tile's bodies belong to
    const int root = A.island[b];
    int k = 0;
    for (; k < B2G_TILE_ROOTS; ++k) {
      const int old = atomicCAS(&sRoots[k], -1, root);
      if (old == -1 || old == root) break;
    }
  }
  for (int k = X.gtid; k < nSpill; k += X.gsize) {  // bodies that did not fit their tile: global memory only
    const int b = A.spillList[k];
    const uint32_t f = A.bflags[b];
    if (!(f & B2G_BODY_AWAKE)) A.bflags[b] = f | B2G_BODY_AWAKE;
    float4 v4 = integrate_velocity(f, A.vel[b], b);
    v4.w = 0.0f;  // the sequence lane starts at zero
    __stcg(A.vel + b, v4);
  }
  __syncthreads();
  X.nUsed = sCounts[0];
  X.nCut = sCounts[1];
  X.nBnd = sCounts[2];
  const bool cutOverflow = cut[B2G_MAX_COLOURS] != cut[B2G_MAX_COLOURS + 1];
  X.anyCut = X.nCut > 0 || cutOverflow;
  X.seqMode = !A.noSequencing && !bigJoints && !cutOverflow;
  // integrated velocities of the boundary bodies, for the cut constraints' preparation (not yet anybody's turn)
  if (X.seqMode) tile_publish_seq(X, X.vel, A.vel, 0, true);
  else tile_publish(X, X.vel, A.vel);
  order_bucket_by_key(cstart[B2G_MAX_COLOURS], cstart[B2G_MAX_COLOURS + 1], A.sortedList, A.orderScratch, C);
  if (blockIdx.x == 0) order_bucket_by_key(cut[B2G_MAX_COLOURS], cut[B2G_MAX_COLOURS + 1], A.sortedList, A.orderScratch, C);
  TILE_MARK(1);
  tile_grid_sync(X);
  TILE_MARK(2);

  // ---- phase 1: prepare the constraints ---------------------------------------------------------------------
  const TileSmemBodies velT{X.vel, A.vel}, posT{X.pos, A.pos};
  const CoherentBodies velG{A.vel}, posG{A.pos};
  for (int s = cstart[0] + tid; s < cstart[B2G_MAX_COLOURS + 1]; s += nt) {
    const int i = A.sortedList[s];
    Manifold m;
    manifold_unpack(m, C.m0[i], C.m1[i], C.m2[i], C.m3[i]);
    const int2 bd = C.body[i];
    const int2 fx = C.fix[i];
    const int sa = A.tileSlot[bd.x], sb = A.tileSlot[bd.y];
    const int ia = (sa >= X.first && sa < X.first + B2G_TILE_CAP) ? sa - X.first : ~bd.x;
    const int ib = (sb >= X.first && sb < X.first + B2G_TILE_CAP) ? sb - X.first : ~bd.y;
    const int root = B2G_BODY_TYPE(A.bflags[bd.x]) != B2G_STATIC ? A.island[bd.x] : A.island[bd.y];
    prepare_constraint(S, s, i, m, bd.x, bd.y, ia, ib, C.material[i], A.fRadius[fx.x], A.fRadius[fx.y], posT, velT, A.mass,
                       A.center, A.dtRatio, A.warmStarting != 0, root);
    A.croot[s] = root;
  }
  // In sequenced mode a cut constraint is prepared, swept and stored by the SAME thread (slot cut[0] + gtid, + gsize,
  // ...), so its planes never travel between threads (in barrier mode grid barriers stand between the three).
  auto prepare_cut = [&](int s) {
    const int i = A.sortedList[s];
    Manifold m;
    manifold_unpack(m, C.m0[i], C.m1[i], C.m2[i], C.m3[i]);
    const int2 bd = C.body[i];
    const int2 fx = C.fix[i];
    const int root = B2G_BODY_TYPE(A.bflags[bd.x]) != B2G_STATIC ? A.island[bd.x] : A.island[bd.y];
    prepare_constraint(S, s, i, m, bd.x, bd.y, bd.x, bd.y, C.material[i], A.fRadius[fx.x], A.fRadius[fx.y], posG, velG, A.mass,
                       A.center, A.dtRatio, A.warmStarting != 0, root);
    A.croot[s] = root;
    const int c = C.colour[i];
    if (c >= B2G_CUT_DOMAIN_SHIFT && (c & 31) < B2G_MAX_COLOURS) {
      const float4 ma = A.mass[bd.x], mb = A.mass[bd.y];
      A.cutSeq[s] = make_int2(seq_pack(A.colourMask[bd.x], c, body_movable(ma), A.tileSlot[bd.x] < 0),
                              seq_pack(A.colourMask[bd.y], c, body_movable(mb), A.tileSlot[bd.y] < 0));
    }
  };
  for (int s = cut[0] + X.gtid; s < cut[B2G_MAX_COLOURS]; s += X.gsize) prepare_cut(s);
  TILE_MARK(3);
  if (X.gtid == 0)
    for (int s = cut[B2G_MAX_COLOURS]; s < cut[B2G_MAX_COLOURS + 1]; ++s) prepare_cut(s);
  // the sweeps fetch a constraint's constants with cp.async.cg, i.e. from L2, and usually from another thread of
  // the CTA than the one that prepared them: make the planes visible there, not just to the CTA
  __threadfence();
  __syncthreads();
  TILE_MARK(4);

  BigPassArgs Q;
  Q.croot = A.croot;
  Q.islandPen = A.islandPen;
  Q.penStride = A.penStride;
  Q.it = 0;

  // ---- phase 2: warm start; joints' InitVelocityConstraints after the contacts' (b2_island.cpp:323-325) -------
  if (A.warmStarting) tile_sweep<B2G_BIG_WARM>(X, G, S, A, Q, A.velIters > 0);
  TILE_MARK(5);
  if (bigJoints)
    tile_joint_phase(X, X.vel, A.vel, [&](auto&& sync) {
      joints_init_coloured<CoherentBodies>(W, J, X.gtid, X.gsize, sync, A.pos, A.vel, A.mass, A.center, A.dtRatio, A.warmStarting);
    });

  // ---- phase 3: velocity iterations: joints, then contacts (b2_island.cpp:330-338) -----------------------------
  for (int it = 0; it < A.velIters; ++it) {
    if (bigJoints)
      tile_joint_phase(X, X.vel, A.vel, [&](auto&& sync) {
        joints_velocity_coloured<CoherentBodies>(W, J, X.gtid, X.gsize, sync, A.vel, A.h, A.invH);
      });
    tile_sweep<B2G_BIG_VELOCITY>(X, G, S, A, Q, it + 1 < A.velIters);
    TILE_MARK(6 + it);
  }

  // ---- phase 4: store impulses (b2_contact_solver.cpp:641-657) ---------------------------------------------------
  auto store_impulses = [&](int s) {
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
  };
  for (int s = cstart[0] + tid; s < cstart[B2G_MAX_COLOURS + 1]; s += nt) store_impulses(s);
  for (int s = cut[0] + X.gtid; s < cut[B2G_MAX_COLOURS]; s += X.gsize) store_impulses(s);
  if (X.gtid == 0)
    for (int s = cut[B2G_MAX_COLOURS]; s < cut[B2G_MAX_COLOURS + 1]; ++s) store_impulses(s);

  // ---- phase 5: integrate positions (b2_island.cpp:353-385) ------------------------------------------------------
  auto integrate_position = [&](float4& p4, float4& v4) {
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
    v4 = make_float4(v.x, v.y, w, v4.w);
  };
  for (int l = tid; l < nbod; l += nt) {
    float4 p4 = X.pos[l], v4 = X.vel[l];
    integrate_position(p4, v4);
    X.pos[l] = p4;
    X.vel[l] = v4;
  }
  if (nSpill > 0 && X.seqMode) tile_grid_sync(X);  // the unowned bodies' last velocity turn must be over everywhere
  for (int k = X.gtid; k < nSpill; k += X.gsize) {
    const int b = A.spillList[k];
    float4 p4 = __ldcg(A.pos + b), v4 = __ldcg(A.vel + b);
    integrate_position(p4, v4);
    p4.w = 0.0f;
    __stcg(A.pos + b, p4);
    __stcg(A.vel + b, v4);
  }
  __syncthreads();
  if (X.seqMode && X.anyCut && A.posIters > 0) {
    // the boundary bodies' global positions are last step's: mark them "nobody's turn" before any tile can
    // start polling them
    tile_publish_seq(X, X.pos, A.pos, 0, true);
    tile_grid_sync(X);
  }

  TILE_MARK(31);
  // ---- phase 6: position iterations: contacts, then joints, per-island early exit (b2_island.cpp:391-409) ----------
  for (int it = 0; it < A.posIters; ++it) {
    Q.it = it;
    tile_sweep<B2G_BIG_POSITION>(X, G, S, A, Q, it + 1 < A.posIters);
    TILE_MARK(32 + it);
    if (bigJoints)
      tile_joint_phase(X, X.pos, A.pos, [&](auto&& sync) {
        joints_position_coloured<CoherentBodies>(W, J, X.gtid, X.gsize, sync, A.pos, A.islandPen, A.penStride, it);
      });
  }

  // ---- phase 7: write back, SynchronizeTransform, sleep (b2_island.cpp:430-483), ClearForces -----------------------
  const float linTolSqr = B2G_LINEAR_SLEEP_TOL * B2G_LINEAR_SLEEP_TOL;
  const float angTolSqr = B2G_ANGULAR_SLEEP_TOL * B2G_ANGULAR_SLEEP_TOL;
  auto sleep_time = [&](uint32_t f, float4 v4, float st, float& minSleep) {
    if (!(f & B2G_BODY_AUTOSLEEP) || v4.z * v4.z > angTolSqr || v4.x * v4.x + v4.y * v4.y > linTolSqr) {
      st = 0.0f;
      minSleep = 0.0f;
    } else {
      st += h;
      minSleep = st;
    }
    return st;
  };
  // per island: min over its bodies of the new sleep time, one atomic per warp and island root
  auto publish_min_sleep = [&](int root, float minSleep) {
    const unsigned int peers = __match_any_sync(0xffffffffu, root);
    const unsigned int bits = __reduce_min_sync(peers, __float_as_uint(minSleep));
    if (root >= 0 && (int)(threadIdx.x & 31) == __ffs(peers) - 1) atomicMin(&A.islandMinSleep[root], bits);
  };
  if (A.allowSleep) {
    for (int l0 = 0; l0 < nbod; l0 += nt) {
      const int l = l0 + tid;
      int root = -1;
      float minSleep = 0.0f;
      if (l < nbod) {
        const int b = X.body[l];
        root = A.island[b];
        const float st = sleep_time(A.bflags[b], X.vel[l], A.force[b].w, minSleep);
        X.vel[l].w = st;  // park the new sleep time in the unused lane
      }
      publish_min_sleep(root, minSleep);
    }
    for (int k0 = 0; k0 < nSpill; k0 += X.gsize) {
      const int k = k0 + X.gtid;
      int root = -1;
      float minSleep = 0.0f;
      if (k < nSpill) {
        const int b = A.spillList[k];
        root = A.island[b];
        float4 v4 = __ldcg(A.vel + b);
        v4.w = sleep_time(A.bflags[b], v4, A.force[b].w, minSleep);
        __stcg(A.vel + b, v4);
      }
      publish_min_sleep(root, minSleep);
    }
    tile_grid_sync(X);
  }
  int awake = 0;
  auto finish_body = [&](int b, float4 p4, float4 v4) {
    const float4 c4 = A.center[b];
    Xf T = xf_from_sweep(make_float2(p4.x, p4.y), p4.z, make_float2(c4.x, c4.y));
    p4.w = 0.0f;  // the sequence lane goes back to zero
    A.pos[b] = p4;
    A.xf[b] = xf_to4(T);
    float4 fo = A.force[b];
    bool sleepNow = false;
    if (A.allowSleep) {
      const int root = A.island[b];
      fo.w = v4.w;
      const bool positionSolved =
          A.posIters > 0 && __uint_as_float(__ldcg(&A.islandPen[(size_t)(A.posIters - 1) * A.penStride + root])) <= 3.0f * B2G_LINEAR_SLOP;
      sleepNow = __uint_as_float(__ldcg(&A.islandMinSleep[root])) >= B2G_TIME_TO_SLEEP && positionSolved;
    }
    if (sleepNow) {
      // b2Body::SetAwake(false), b2_body.h:731-739
      A.bflags[b] &= ~B2G_BODY_AWAKE;
      A.vel[b] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
      A.force[b] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    } else {
      A.vel[b] = make_float4(v4.x, v4.y, v4.z, 0.0f);
      if (A.clearForces) fo.x = fo.y = fo.z = 0.0f;
      A.force[b] = fo;
      ++awake;
    }
  };
  for (int l = tid; l < nbod; l += nt) finish_body(X.body[l], X.pos[l], X.vel[l]);
  for (int k = X.gtid; k < nSpill; k += X.gsize) {
    const int b = A.spillList[k];
    finish_body(b, __ldcg(A.pos + b), __ldcg(A.vel + b));
  }
  if (awake) atomicAdd(&A.counts->numAwake, awake);
  TILE_MARK(40);
}

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

#define B2G_TILES_MAX 160        // >= SM count (one tile per SM)
#define B2G_TILE_CAP 2048        // bodies a tile can hold in shared memory
#define B2G_TILE_THREADS 512
#define B2G_TILE_XBINS 4096
#define B2G_TILE_YBINS 1024
#define B2G_TILE_MIN_BODIES 128  // do not cut an island into tiles smaller than this
#define B2G_TILE_PLAN_PERIOD 32  // steps between re-plans (sooner when tiles overflow)

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
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
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
  for (int S = 1; S <= T; ++S) {
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
// exclusive prefix of `n` (<= 4096) counts by one block of 1024 threads; out[k] = min(parts - 1, prefix * parts / total)
__device__ __forceinline__ void tile_partition(const int* __restrict__ hist, int n, int parts, int* out) {
  __shared__ int warpSums[32];
  const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
  const int per = (n + 1023) / 1024;  // <= 4
  int v[4], sum = 0;
  for (int k = 0; k < 4; ++k) {
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
  for (int k = 0; k < 4; ++k) {
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
    if (l < B2G_TILE_CAP) {
      ts = tile * B2G_TILE_CAP + l;
      tileBodies[ts] = b;
    } else {
      spillList[atomicAdd(&counts->spillCount, 1)] = b;  // lives in global memory only; its constraints are cut
    }
  }
  tileSlot[b] = ts;
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
struct TileArgs {
  int tileBin0, cutBin, nb, nj;
  float h, invH, dtRatio;
  float2 gravity;
  int velIters, posIters, warmStarting, allowSleep, clearForces;
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

template <int PHASE, bool STAGED, class VelAcc, class PosAcc>
__device__ __forceinline__ float tile_visit(BigStage& G, const SolverPlanes& S, int s, const VelAcc& velAcc,
                                            const PosAcc& posAcc) {
  const int t = threadIdx.x;
  if (STAGED) stage_acquire(G, S, s, PHASE == B2G_BIG_POSITION ? B2G_STAGE_POSITION : B2G_STAGE_VELOCITY);
  if (PHASE == B2G_BIG_WARM) {
    if (STAGED) warm_start_constraint(G.T, t, velAcc);
    else warm_start_constraint(S, s, velAcc);
  } else if (PHASE == B2G_BIG_VELOCITY) {
    if (STAGED) {
      solve_velocity_constraint(G.T, t, velAcc);
      S.imp[s] = G.T.imp[t];
    } else {
      solve_velocity_constraint(S, s, velAcc);
    }
  } else {
    float minSep = STAGED ? solve_position_constraint(G.T, t, posAcc) : solve_position_constraint(S, s, posAcc);
    return minSep < 0.0f ? -minSep : 0.0f;
  }
  return 0.0f;
}
// this thread's share of one colour: slot sFirst (staged), then sFirst + stride, ... (direct).  Called by
// whole warps (the position passes publish one penetration per warp and island root).
template <int PHASE, class VelAcc, class PosAcc>
__device__ __forceinline__ void tile_colour(BigStage& G, const SolverPlanes& S, int sFirst, int s1, int stride,
                                            const VelAcc& velAcc, const PosAcc& posAcc, const BigPassArgs& Q) {
  int root = -1;
  float pen = 0.0f;
  if (sFirst < s1) {
    if (PHASE == B2G_BIG_POSITION) {
      root = Q.croot[sFirst];
      if (island_done_l2(Q.islandPen, Q.penStride, Q.it, root)) root = -1;
    }
    if (PHASE != B2G_BIG_POSITION || root >= 0) pen = tile_visit<PHASE, true>(G, S, sFirst, velAcc, posAcc);
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
        float p2 = tile_visit<PHASE, false>(G, S, s, velAcc, posAcc);
        if (p2 > 0.0f) {
          uint32_t* slot = &Q.islandPen[(size_t)Q.it * Q.penStride + r2];
          if (__ldcg(slot) < __float_as_uint(p2)) atomicMax(slot, __float_as_uint(p2));
        }
      } else {
        tile_visit<PHASE, false>(G, S, s, velAcc, posAcc);
      }
    }
  }
}

struct TileCtx {
  int tid, nt, gtid, gsize;
  int nbod, nBnd;
  int first;                 // tile * B2G_TILE_CAP
  float4 *vel, *pos;         // shared
  int* body;                 // shared
  unsigned short* bnd;       // shared: local slots of the boundary bodies
  const int* cstart;         // shared [B2G_MAX_COLOURS + 2]: interior ranges of this tile
  const int* cut;            // shared [B2G_MAX_COLOURS + 2]: cut ranges (the same for every tile)
  const int *usedS0, *usedS1;
  int nUsed;
  const int *cutS0, *cutS1;
  int nCut;
  bool anyCut;               // cut colours or a cut overflow bucket exist
  unsigned int* barrier;
  unsigned int target;
};
__device__ __forceinline__ void tile_grid_sync(TileCtx& X) {
  grid_arrive(X.barrier, X.target);
  grid_wait(X.barrier, X.target);
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

// One solver iteration of PHASE: the tile's interior colours with CTA barriers, then — after the boundary
// bodies have been published — the cut colours with grid barriers, then the boundary bodies come back.
template <int PHASE>
__device__ __forceinline__ void tile_sweep(TileCtx& X, BigStage& G, const SolverPlanes& S, const TileArgs& A,
                                           const BigPassArgs& Q, bool again) {
  const int kind = PHASE == B2G_BIG_POSITION ? B2G_STAGE_POSITION : B2G_STAGE_VELOCITY;
  const TileSmemBodies velT{X.vel, A.vel}, posT{X.pos, A.pos};
  const CoherentBodies velG{A.vel}, posG{A.pos};
  // ---- interior
  for (int k = 0; k < X.nUsed; ++k) {
    tile_colour<PHASE>(G, S, X.usedS0[k] + X.tid, X.usedS1[k], X.nt, velT, posT, Q);
    // what this thread visits next: the next interior colour, else its first cut constraint, else the first
    // interior colour of the following sweep
    int sn = 0, sl = 0;
    if (k + 1 < X.nUsed) {
      sn = X.usedS0[k + 1] + X.tid;
      sl = X.usedS1[k + 1];
    } else if (X.nCut > 0) {
      sn = X.cutS0[0] + X.gtid;
      sl = X.cutS1[0];
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
          if (PHASE == B2G_BIG_POSITION) {
            const int root = Q.croot[s];
            if (island_done_l2(Q.islandPen, Q.penStride, Q.it, root)) continue;
            const float pen = tile_visit<PHASE, false>(G, S, s, velT, posT);
            atomicMax(&Q.islandPen[(size_t)Q.it * Q.penStride + root], __float_as_uint(pen));
          } else {
            tile_visit<PHASE, false>(G, S, s, velT, posT);
          }
        }
      }
      __syncthreads();
    }
  }
  // ---- cut
  if (X.anyCut) {
    float4* sm = PHASE == B2G_BIG_POSITION ? X.pos : X.vel;
    float4* gl = PHASE == B2G_BIG_POSITION ? A.pos : A.vel;
    tile_publish(X, sm, gl);
    tile_grid_sync(X);
    for (int k = 0; k < X.nCut; ++k) {
      tile_colour<PHASE>(G, S, X.cutS0[k] + X.gtid, X.cutS1[k], X.gsize, velG, posG, Q);
      int sn = 0, sl = 0;
      if (k + 1 < X.nCut) {
        sn = X.cutS0[k + 1] + X.gtid;
        sl = X.cutS1[k + 1];
      } else if (again && X.nUsed > 0) {
        sn = X.usedS0[0] + X.tid;
        sl = X.usedS1[0];
      }
      grid_arrive(X.barrier, X.target);
      if (sn < sl) stage_prefetch(G, S, sn, kind);
      grid_wait(X.barrier, X.target);
    }
    const int o0 = X.cut[B2G_MAX_COLOURS], o1 = X.cut[B2G_MAX_COLOURS + 1];
    if (o1 > o0) {  // serial bucket of the cut domain: one thread of the grid
      if (X.gtid == 0) {
        for (int s = o0; s < o1; ++s) {
          if (PHASE == B2G_BIG_POSITION) {
            const int root = Q.croot[s];
            if (island_done_l2(Q.islandPen, Q.penStride, Q.it, root)) continue;
            const float pen = tile_visit<PHASE, false>(G, S, s, velG, posG);
            atomicMax(&Q.islandPen[(size_t)Q.it * Q.penStride + root], __float_as_uint(pen));
          } else {
            tile_visit<PHASE, false>(G, S, s, velG, posG);
          }
        }
      }
      tile_grid_sync(X);
    }
    tile_reload(X, sm, gl);
  } else if (PHASE == B2G_BIG_POSITION) {
    tile_grid_sync(X);  // the next iteration's early-exit test reads every tile's penetration
  }
}

// joints of oversize islands: one thread of the grid walks them over the global body copies (the joints'
// bodies are boundary bodies, see k_tile_joint_marks), between a publish and a reload
template <class F>
__device__ __forceinline__ void tile_joint_phase(TileCtx& X, float4* sm, float4* gl, F&& walk) {
  tile_publish(X, sm, gl);
  tile_grid_sync(X);
  if (X.gtid == 0) walk();
  tile_grid_sync(X);
  tile_reload(X, sm, gl);
}

__global__ void __launch_bounds__(B2G_TILE_THREADS, 1)
k_big_tiles(TileArgs A, SolverPlanes S, ContactBuf C, JointWalk W, JointArraysDev J) {
  extern __shared__ __align__(16) unsigned char tileSmem[];
  __shared__ int cstart[B2G_MAX_COLOURS + 2], cut[B2G_MAX_COLOURS + 2];
  __shared__ int usedS0[B2G_MAX_COLOURS], usedS1[B2G_MAX_COLOURS], cutS0[B2G_MAX_COLOURS], cutS1[B2G_MAX_COLOURS];
  __shared__ int sCounts[4];  // nUsed, nCut, nBnd, awake

  TileCtx X;
  X.tid = threadIdx.x;
  X.nt = blockDim.x;
  // consecutive groups of 32 cut constraints go to DIFFERENT blocks (warp w of block b is global warp
  // w * gridDim + b): a cut colour of a few thousand constraints keeps a warp or two busy on every SM
  X.gtid = ((threadIdx.x >> 5) * gridDim.x + blockIdx.x) * 32 + (threadIdx.x & 31);
  X.gsize = gridDim.x * blockDim.x;
  X.barrier = A.barrier;
  X.target = 0;
  const int tid = X.tid, nt = X.nt;
  const int numTiles = A.plan->S * A.plan->R;
  const int tile = blockIdx.x;
  X.first = tile * B2G_TILE_CAP;
  {
    int n = tile < numTiles ? A.tileCount[tile] : 0;
    X.nbod = n > B2G_TILE_CAP ? B2G_TILE_CAP : n;
  }
  const int nbod = X.nbod;
  {
    unsigned char* p = tileSmem;
    X.vel = (float4*)p;            p += (size_t)B2G_TILE_CAP * 16;
    X.pos = (float4*)p;            p += (size_t)B2G_TILE_CAP * 16;
    X.body = (int*)p;              p += (size_t)B2G_TILE_CAP * 4;
    X.bnd = (unsigned short*)p;    p += (size_t)B2G_TILE_CAP * 2;
  }
  BigStage G;
  {
    float4* m = (float4*)(tileSmem + (size_t)B2G_TILE_CAP * (16 + 16 + 4 + 2));
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
  if (tid == 0) sCounts[2] = 0, sCounts[3] = 0;
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
    if (A.tileBoundary[X.first + l]) X.bnd[atomicAdd(&sCounts[2], 1)] = (unsigned short)l;
  }
  for (int k = X.gtid; k < nSpill; k += X.gsize) {  // bodies that did not fit their tile: global memory only
    const int b = A.spillList[k];
    const uint32_t f = A.bflags[b];
    if (!(f & B2G_BODY_AWAKE)) A.bflags[b] = f | B2G_BODY_AWAKE;
    __stcg(A.vel + b, integrate_velocity(f, A.vel[b], b));
  }
  __syncthreads();
  X.nUsed = sCounts[0];
  X.nCut = sCounts[1];
  X.nBnd = sCounts[2];
  X.anyCut = X.nCut > 0 || cut[B2G_MAX_COLOURS] != cut[B2G_MAX_COLOURS + 1];
  const bool bigJoints = A.nj > 0 && A.counts->bigJoints > 0;
  tile_publish(X, X.vel, A.vel);  // integrated velocities of the boundary bodies, for the cut constraints' preparation
  order_bucket_by_key(cstart[B2G_MAX_COLOURS], cstart[B2G_MAX_COLOURS + 1], A.sortedList, A.orderScratch, C);
  if (blockIdx.x == 0) order_bucket_by_key(cut[B2G_MAX_COLOURS], cut[B2G_MAX_COLOURS + 1], A.sortedList, A.orderScratch, C);
  tile_grid_sync(X);

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
    prepare_constraint(S, s, i, m, bd.x, bd.y, ia, ib, C.material[i], A.fRadius[fx.x], A.fRadius[fx.y], posT, velT, A.mass,
                       A.center, A.dtRatio, A.warmStarting != 0);
    A.croot[s] = B2G_BODY_TYPE(A.bflags[bd.x]) != B2G_STATIC ? A.island[bd.x] : A.island[bd.y];
  }
  for (int s = cut[0] + X.gtid; s < cut[B2G_MAX_COLOURS + 1]; s += X.gsize) {
    const int i = A.sortedList[s];
    Manifold m;
    manifold_unpack(m, C.m0[i], C.m1[i], C.m2[i], C.m3[i]);
    const int2 bd = C.body[i];
    const int2 fx = C.fix[i];
    prepare_constraint(S, s, i, m, bd.x, bd.y, bd.x, bd.y, C.material[i], A.fRadius[fx.x], A.fRadius[fx.y], posG, velG, A.mass,
                       A.center, A.dtRatio, A.warmStarting != 0);
    A.croot[s] = B2G_BODY_TYPE(A.bflags[bd.x]) != B2G_STATIC ? A.island[bd.x] : A.island[bd.y];
  }
  __syncthreads();  // the cut constraints are first read after the grid barrier that follows the interior warm start

  BigPassArgs Q;
  Q.croot = A.croot;
  Q.islandPen = A.islandPen;
  Q.penStride = A.penStride;
  Q.it = 0;

  // ---- phase 2: warm start; joints' InitVelocityConstraints after the contacts' (b2_island.cpp:323-325) -------
  if (A.warmStarting) tile_sweep<B2G_BIG_WARM>(X, G, S, A, Q, A.velIters > 0);
  else if (X.anyCut) tile_grid_sync(X);  // the cut constraints' planes must be complete before their first sweep
  if (bigJoints)
    tile_joint_phase(X, X.vel, A.vel, [&]() { joints_init_global<CoherentBodies>(W, J, A.pos, A.vel, A.mass, A.center, A.dtRatio, A.warmStarting); });

  // ---- phase 3: velocity iterations: joints, then contacts (b2_island.cpp:330-338) -----------------------------
  for (int it = 0; it < A.velIters; ++it) {
    if (bigJoints) tile_joint_phase(X, X.vel, A.vel, [&]() { joints_velocity_global<CoherentBodies>(W, J, A.vel, A.h, A.invH); });
    tile_sweep<B2G_BIG_VELOCITY>(X, G, S, A, Q, it + 1 < A.velIters);
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
  for (int s = cut[0] + X.gtid; s < cut[B2G_MAX_COLOURS + 1]; s += X.gsize) store_impulses(s);

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
  for (int k = X.gtid; k < nSpill; k += X.gsize) {
    const int b = A.spillList[k];
    float4 p4 = __ldcg(A.pos + b), v4 = __ldcg(A.vel + b);
    integrate_position(p4, v4);
    __stcg(A.pos + b, p4);
    __stcg(A.vel + b, v4);
  }
  __syncthreads();

  // ---- phase 6: position iterations: contacts, then joints, per-island early exit (b2_island.cpp:391-409) ----------
  for (int it = 0; it < A.posIters; ++it) {
    Q.it = it;
    tile_sweep<B2G_BIG_POSITION>(X, G, S, A, Q, it + 1 < A.posIters);
    if (bigJoints)
      tile_joint_phase(X, X.pos, A.pos, [&]() { joints_position_global<CoherentBodies>(W, J, A.pos, A.islandPen, A.penStride, it); });
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
    const unsigned int peers = __match_any_sync(__activemask(), root);
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
}

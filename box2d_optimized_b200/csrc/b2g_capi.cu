// b2g_capi.cu — arena management, the per-step launch sequence and the C-ABI of include/b2cuda.h.
//
// b2g_step() is b2World::Step (src/dynamics/b2_world.cpp:1108-1171) re-expressed as a fixed
// sequence of kernels on one CUDA stream; see b2g_step_kernels.cuh / b2g_broadphase.cuh for the
// kernel <-> reference-loop mapping.  There is no CPU fallback anywhere in this file.
#include <cub/cub.cuh>
#include <thrust/iterator/counting_iterator.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <algorithm>
#include "b2g_broadphase.cuh"
#include "b2g_tiles.cuh"
#include "b2g_query.cuh"
#include <thrust/iterator/transform_iterator.h>

static thread_local char g_err[512] = "";
static int set_err(const char* what, const char* detail) {
  snprintf(g_err, sizeof(g_err), "%s: %s", what, detail ? detail : "");
  return 0;
}
#define CK(call)                                              \
  do {                                                        \
    cudaError_t e_ = (call);                                  \
    if (e_ != cudaSuccess) {                                  \
      set_err(#call, cudaGetErrorString(e_));                 \
      return B2G_ERR_CUDA;                                    \
    }                                                         \
  } while (0)

static inline int div_up(int a, int b) { return (a + b - 1) / b; }

struct DevBuf {
  void* p = nullptr;
  ~DevBuf() {
    if (p) cudaFree(p);
  }
  template <typename T>
  T* as() {
    return (T*)p;
  }
  cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 1); }
  cudaError_t upload(const void* src, size_t bytes) {
    cudaError_t e = alloc(bytes);
    if (e != cudaSuccess) return e;
    return cudaMemcpy(p, src, bytes, cudaMemcpyHostToDevice);
  }
};

static int use_device(int device) {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    set_err("b2cuda", "no CUDA device: this library has no CPU fallback");
    return B2G_ERR_NO_DEVICE;
  }
  if (device < 0 || device >= ndev) return B2G_ERR_INVALID;
  CK(cudaSetDevice(device));
  return B2G_OK;
}

// kernel classes for the per-kernel timing of bench.py's roofline line (b2g_get_kernel_timing)
enum KClass {
  KC_NARROWPHASE, KC_ISLANDS, KC_INTEGRATE, KC_COLOUR, KC_PREPARE, KC_WARM_START, KC_SOLVE_VELOCITY,
  KC_SOLVE_POSITION, KC_STORE_IMPULSES, KC_FINALIZE, KC_BP_BUILD, KC_BP_TRAVERSE, KC_CONTACT_MERGE, KC_SORT_SCAN,
  KC_FUSED_SOLVE, KC_QUERY, KC_BIG_SOLVE, KC_BIG_TILES, KC_COUNT
};
static const char* kClassNames[KC_COUNT] = {
    "narrowphase", "islands", "integrate", "colour", "prepare", "warm_start", "solve_velocity", "solve_position",
    "store_impulses", "finalize", "bp_build", "bp_traverse", "contact_merge", "sort_scan", "fused_solve", "query",
    "big_solve", "big_tiles"};

static inline void ktime_begin(b2gArena* A, int cls, double units) {
  if (!A->kernelTiming || A->ktCount >= B2G_KT_MAX) return;
  cudaEventRecord(A->ktEv[2 * A->ktCount], A->stream);
  A->ktClass[A->ktCount] = cls;
  A->ktUnits[A->ktCount] = units;
}
static inline void ktime_end(b2gArena* A) {
  if (!A->kernelTiming || A->ktCount >= B2G_KT_MAX) return;
  cudaEventRecord(A->ktEv[2 * A->ktCount + 1], A->stream);
  A->ktCount++;
}
// called after a stream sync: fold this step's event pairs into the per-class totals
static void ktime_collect(b2gArena* A) {
  for (int i = 0; i < A->ktCount; ++i) {
    float ms = 0.0f;
    cudaEventElapsedTime(&ms, A->ktEv[2 * i], A->ktEv[2 * i + 1]);
    int c = A->ktClass[i];
    A->ktMs[c] += ms;
    A->ktLaunches[c] += 1;
    A->ktUnitsSum[c] += A->ktUnits[i];
  }
  A->ktCount = 0;
}

// `units` = the work items of this launch in the unit SURVEY §8(d) quotes bytes for
// every step kernel is launched with programmatic stream serialization (B2G_PDL_ENTER in
// b2g_math.cuh); B2G_PDL=0 in the environment falls back to plain stream order (A/B measurements)
static int pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("B2G_PDL");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v;
}
template <class... KArgs, class... Args>
static inline cudaError_t launch_pdl(cudaStream_t st, dim3 grid, dim3 block, size_t smem, void (*kernel)(KArgs...),
                                     Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#define LAUNCH(A, cls, units, kernel, grid, block, ...)                      \
  do {                                                                       \
    ktime_begin((A), (cls), (double)(units));                                \
    launch_pdl((A)->stream, dim3(grid), dim3(block), 0, kernel, __VA_ARGS__); \
    ktime_end((A));                                                          \
    (A)->launches++;                                                         \
  } while (0)
#define TIMED(A, cls, units, stmt) \
  do {                             \
    ktime_begin((A), (cls), (double)(units)); \
    stmt;                          \
    ktime_end((A));                \
  } while (0)

extern "C" const char* b2g_last_error(void) { return g_err; }

extern "C" int b2g_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

template <typename T>
static cudaError_t dalloc(T** p, size_t n) {
  cudaError_t e = cudaMalloc((void**)p, (n ? n : 1) * sizeof(T));
  if (e == cudaSuccess) e = cudaMemset(*p, 0, (n ? n : 1) * sizeof(T));
  return e;
}

static int alloc_contact_buf(ContactBuf& c, int cap) {
  CK(dalloc(&c.key, cap));
  CK(dalloc(&c.fix, cap));
  CK(dalloc(&c.body, cap));
  CK(dalloc(&c.flags, cap));
  CK(dalloc(&c.material, cap));
  CK(dalloc(&c.m0, cap));
  CK(dalloc(&c.m1, cap));
  CK(dalloc(&c.m2, cap));
  CK(dalloc(&c.m3, cap));
  CK(dalloc(&c.colour, cap));
  return B2G_OK;
}
static void free_contact_buf(ContactBuf& c) {
  cudaFree(c.key);
  cudaFree(c.fix);
  cudaFree(c.body);
  cudaFree(c.flags);
  cudaFree(c.material);
  cudaFree(c.m0);
  cudaFree(c.m1);
  cudaFree(c.m2);
  cudaFree(c.m3);
  cudaFree(c.colour);
}

static int bits_for(int n) {
  int b = 1;
  while ((1ll << b) < (long long)n) ++b;
  return b;
}

extern "C" int b2g_arena_create(const b2gArenaDef* def, b2gArena** out) {
  if (!def || !out || def->max_bodies <= 0 || def->max_fixtures <= 0 || def->max_contacts <= 0 ||
      def->max_shape_quads <= 0 || def->num_worlds <= 0) {
    set_err("b2g_arena_create", "invalid definition");
    return B2G_ERR_INVALID;
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    set_err("b2g_arena_create", "no CUDA device: this library has no CPU fallback");
    return B2G_ERR_NO_DEVICE;
  }
  if (def->device < 0 || def->device >= ndev) {
    set_err("b2g_arena_create", "device ordinal out of range");
    return B2G_ERR_INVALID;
  }
  CK(cudaSetDevice(def->device));
  b2gArena* A = (b2gArena*)calloc(1, sizeof(b2gArena));
  A->device = def->device;
  A->numWorlds = def->num_worlds;
  A->worldBits = bits_for(def->num_worlds + 1);
  A->capBodies = def->max_bodies;
  A->capFixtures = def->max_fixtures;
  A->capQuads = def->max_shape_quads;
  A->capContacts = def->max_contacts;
  A->capJoints = def->max_joints > 0 ? def->max_joints : 1;
  A->fixBits = bits_for(def->max_fixtures);
  A->aabbAllDirty = 1;
  A->recolour = 1;
  A->roundsHint = 8;
  CK(cudaStreamCreateWithFlags(&A->stream, cudaStreamNonBlocking));
  CK(cudaStreamCreateWithFlags(&A->copyStream, cudaStreamNonBlocking));
  CK(cudaEventCreateWithFlags(&A->evPacked, cudaEventDisableTiming));
  for (int i = 0; i < 5; ++i) CK(cudaEventCreate(&A->ev[i]));
  for (int i = 0; i < 2 * B2G_KT_MAX; ++i) CK(cudaEventCreate(&A->ktEv[i]));

  const int nb = A->capBodies, nf = A->capFixtures, nc = A->capContacts, nj = A->capJoints;
  CK(dalloc(&A->pos, nb));
  CK(dalloc(&A->vel, nb));
  CK(dalloc(&A->xf, nb));
  CK(dalloc(&A->mass, nb));
  CK(dalloc(&A->center, nb));
  CK(dalloc(&A->force, nb));
  CK(dalloc(&A->bflags, nb));
  CK(dalloc(&A->bworld, nb));
  CK(dalloc(&A->worldFixMin, A->numWorlds + 1));
  CK(dalloc(&A->bodyFixBase, nb));
  A->fixBaseDirty = 1;
  CK(dalloc(&A->island, nb));
  CK(dalloc(&A->islandParent, nb));
  CK(dalloc(&A->islandDirty, nb));
  CK(dalloc(&A->islandWasBig, nb));
  CK(dalloc(&A->islandAwake, nb));
  CK(dalloc(&A->islandMinSleep, nb));
  CK(dalloc(&A->islandPen, (size_t)nb * B2G_MAX_POS_ITERS));
  CK(dalloc(&A->colourMask, nb));
  CK(dalloc(&A->bodyBest, nb));
  CK(dalloc(&A->islandCount, nb));
  CK(dalloc(&A->islandStart, nb));
  CK(dalloc(&A->islandCursor, nb));
  CK(dalloc(&A->bodySlot, nb));
  CK(dalloc(&A->slotBody, nb));
  A->nbinsMax = nb / 32 + 8 + B2G_TILES_MAX + 2;
  CK(dalloc(&A->binFirst, A->nbinsMax));
  CK(dalloc(&A->binEnd, A->nbinsMax));
  CK(dalloc(&A->bucketCount, (size_t)(A->nbinsMax + 1) * 32 + 1));
  CK(dalloc(&A->bucketStart, (size_t)(A->nbinsMax + 1) * 32 + 1));

  CK(dalloc(&A->fBody, nf));
  CK(dalloc(&A->fShapeOff, nf));
  CK(dalloc(&A->fTypeFlags, nf));
  CK(dalloc(&A->fFilter, nf));
  CK(dalloc(&A->fMaterial, nf));
  CK(dalloc(&A->fAabb, nf));
  CK(dalloc(&A->fRadius, nf));
  CK(dalloc(&A->shapes, A->capQuads));

  CK(dalloc(&A->jBodies, nj));
  CK(dalloc(&A->jAnchors, nj));
  CK(dalloc(&A->jParams0, nj));
  CK(dalloc(&A->jParams1, nj));
  CK(dalloc(&A->jParams2, nj));
  CK(dalloc(&A->jState, nj));
  CK(dalloc(&A->jUpper, nj));
  CK(dalloc(&A->jWork, nj));
  CK(dalloc(&A->worldRecolour, A->numWorlds));
  CK(cudaMemset(A->worldRecolour, 0, (size_t)A->numWorlds));
  CK(dalloc(&A->jColour, nj));
  CK(dalloc(&A->jSorted, nj));
  CK(dalloc(&A->jCstart, B2G_JOINT_COLOURS + 2));
  CK(dalloc(&A->jBodyMask, nb));
  CK(dalloc(&A->jBodyBest, nb));
  A->jointColourDirty = 1;
  CK(dalloc(&A->stateStage, (size_t)nb * 2));
  CK(dalloc(&A->forceStage, (size_t)nb));
  CK(dalloc(&A->jointOrder, nj));
  CK(dalloc(&A->ncKeys, nj));
  CK(dalloc(&A->ncKeysSorted, nj));
  CK(dalloc(&A->bodyNoCollide, nb));

  int rc = alloc_contact_buf(A->cb[0], nc);
  if (rc) return rc;
  CK(dalloc(&A->persist, nc));
  CK(dalloc(&A->seqKeys, (size_t)2 * nc));
  CK(dalloc(&A->orderKey, nc));
  CK(dalloc(&A->freeStack, nc));
  CK(dalloc(&A->dFreeTop, 1));
  {
    // >= 4 x max_contacts: live entries (<= max_contacts) plus tombstones (rebuilt away at cap / 4, see
    // find_new_contacts) can never fill the table, so every probe sequence meets an empty cell
    unsigned int cap = 1024;
    while (cap < 4u * (unsigned int)nc) cap <<= 1;
    A->hash.mask = cap - 1;
    CK(cudaMalloc((void**)&A->hash.keys, (size_t)cap * 8));
    CK(cudaMalloc((void**)&A->hash.vals, (size_t)cap * 4));
    CK(cudaMemset(A->hash.keys, 0xff, (size_t)cap * 8));
    CK(cudaMemset(A->hash.vals, 0xff, (size_t)cap * 4));
  }

  CK(dalloc(&A->mortonKeys, nf));
  CK(dalloc(&A->mortonKeysSorted, nf));
  CK(dalloc(&A->leafFixture, nf));
  CK(dalloc(&A->leafFixtureSorted, nf));
  CK(dalloc(&A->leafBox, nf));
  CK(dalloc(&A->leafInfo, nf));
  CK(dalloc(&A->leafKey, nf));
  CK(dalloc(&A->worldFirst, A->numWorlds + 1));
  CK(dalloc(&A->worldLast, A->numWorlds + 1));
  CK(dalloc(&A->bvhBox, (size_t)wide_bvh_nodes(nf)));
  CK(dalloc(&A->bvhKey, (size_t)wide_bvh_nodes(nf)));
  CK(dalloc(&A->bvhDone, 1));
  CK(dalloc(&A->pairKeys, nc));

  CK(dalloc(&A->activeFlag, nc));
  CK(dalloc(&A->activeList, nc));
  CK(dalloc(&A->sortedList, nc));
  CK(dalloc(&A->colourKey, nc));
  CK(dalloc(&A->colourKeySorted, nc));
  CK(dalloc(&A->croot, nc));
  CK(dalloc(&A->cbin, nc));
  CK(dalloc(&A->conKeys, nc));
  CK(dalloc(&A->conKeysSorted, nc));
  CK(dalloc(&A->conVals, nc));
  SolverPlanes& S = A->planes;
  CK(dalloc(&S.nf, nc));
  CK(dalloc(&S.r1, nc));
  CK(dalloc(&S.r2, nc));
  CK(dalloc(&S.m1, nc));
  CK(dalloc(&S.m2, nc));
  CK(dalloc(&S.kk, nc));
  CK(dalloc(&S.mass, nc));
  CK(dalloc(&S.idx, nc));
  CK(dalloc(&S.imp, nc));
  CK(dalloc(&S.pn, nc));
  CK(dalloc(&S.pp, nc));
  CK(dalloc(&S.pc, nc));
  CK(dalloc(&S.pr, nc));

  CK(dalloc(&A->beginEvents, nc));
  CK(dalloc(&A->endEvents, nc));
  CK(dalloc(&A->dCounts, 1));
  CK(dalloc(&A->bigBarrier, 1));
  CK(dalloc(&A->colourBarrier, 1));
  CK(dalloc(&A->tilePlan, 1));
  CK(dalloc(&A->tileStripOfX, B2G_TILE_XBINS));
  CK(dalloc(&A->tileHistX, B2G_TILE_XBINS));
  CK(dalloc(&A->tileRowOfY, (size_t)B2G_TILES_MAX * B2G_TILE_YBINS));
  CK(dalloc(&A->tileHistY, (size_t)B2G_TILES_MAX * B2G_TILE_YBINS));
  CK(dalloc(&A->tileSlot, nb));
  CK(dalloc(&A->spillList, nb));
  CK(dalloc(&A->tileBodies, (size_t)B2G_TILES_MAX * B2G_TILE_CAP));
  CK(dalloc(&A->tileBoundary, (size_t)B2G_TILES_MAX * B2G_TILE_CAP));
  CK(dalloc(&A->tileCount, B2G_TILES_MAX));
  CK(dalloc(&A->tileBarrier, 1));
  CK(dalloc(&A->tileCutSeq, nc));
  {
    const char* e = getenv("B2G_NO_TILES");
    A->tilesDisabled = e ? atoi(e) : 0;
    e = getenv("B2G_TILE_BARRIERS");
    A->tileNoSequencing = e ? atoi(e) : 0;
  }
  CK(cudaMallocHost((void**)&A->hCounts, sizeof(StepCounts)));
  memset(A->hCounts, 0, sizeof(StepCounts));
  CK(cudaMallocHost((void**)&A->hostStage, 4096));

  // one CUB scratch buffer, sized for the largest of the four primitives used per step
  size_t need = 0, t = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, t, A->mortonKeys, A->mortonKeysSorted, A->leafFixture,
                                  A->leafFixtureSorted, nf, 0, 64, A->stream);
  need = t > need ? t : need;
  cub::DeviceRadixSort::SortPairs(nullptr, t, A->colourKey, A->colourKeySorted, A->activeList, A->sortedList, nc, 0, 8,
                                  A->stream);
  need = t > need ? t : need;
  cub::DeviceSelect::Flagged(nullptr, t, thrust::counting_iterator<int>(0), A->activeFlag, A->activeList,
                             &A->dCounts->numActive, nc, A->stream);
  need = t > need ? t : need;
  cub::DeviceRadixSort::SortPairs(nullptr, t, A->seqKeys, A->seqKeys + nc, A->activeList, A->sortedList, nc, 0, 64,
                                  A->stream);
  need = t > need ? t : need;
  A->cubTempBytes = need + 256;
  CK(cudaMalloc(&A->cubTemp, A->cubTempBytes));
  CK(cudaStreamSynchronize(A->stream));
  *out = A;
  return B2G_OK;
}

extern "C" int b2g_arena_destroy(b2gArena* A) {
  if (!A) return B2G_ERR_INVALID;
  cudaSetDevice(A->device);
  cudaStreamSynchronize(A->stream);
  void* ptrs[] = {A->pos, A->vel, A->xf, A->mass, A->center, A->force, A->bflags, A->bworld, A->worldFixMin, A->bodyFixBase, A->island, A->islandParent, A->islandDirty, A->islandWasBig,
                  A->islandAwake, A->islandMinSleep, A->islandPen, A->colourMask, A->bodyBest, A->islandCount, A->islandStart, A->islandCursor, A->bodySlot, A->slotBody,
                  A->binFirst, A->binEnd, A->bucketCount, A->bucketStart, A->cbin, A->conKeys, A->conKeysSorted, A->conVals, A->fBody,
                  A->fShapeOff, A->fTypeFlags, A->fFilter, A->fMaterial, A->fAabb, A->fRadius, A->shapes,
                  A->jBodies, A->jAnchors, A->jParams0, A->jParams1, A->jParams2, A->jState, A->jUpper, A->jWork, A->worldRecolour, A->jColour, A->jSorted, A->jCstart, A->jBodyMask, A->jBodyBest, A->stateStage, A->forceStage, A->jointOrder, A->ncKeys, A->ncKeysSorted, A->bodyNoCollide, A->seqKeys, A->persist, A->freeStack, A->dFreeTop, A->hash.keys, A->hash.vals, A->mortonKeys,
                  A->mortonKeysSorted, A->leafFixture, A->leafFixtureSorted, A->leafBox, A->leafInfo,
                  A->leafKey, A->worldFirst, A->worldLast, A->bvhBox, A->bvhKey, A->bvhDone,
                  A->pairKeys, A->activeFlag, A->activeList, A->sortedList, A->colourKey,
                  A->colourKeySorted, A->croot, A->planes.nf, A->planes.r1, A->planes.r2, A->planes.m1,
                  A->planes.m2, A->planes.kk, A->planes.mass, A->planes.idx, A->planes.imp, A->planes.pn,
                  A->planes.pp, A->planes.pc, A->planes.pr, A->beginEvents, A->endEvents, A->dCounts, A->bigBarrier, A->colourBarrier, A->tilePlan, A->tileStripOfX, A->tileHistX, A->tileRowOfY, A->tileHistY, A->tileSlot, A->spillList, A->tileBodies, A->tileBoundary, A->tileCount, A->tileBarrier, A->tileCutSeq, A->cubTemp};
  for (void* p : ptrs) cudaFree(p);
  for (int k = 0; k < 2; ++k) {
    cudaFree(A->haloSend[k]);
    cudaFree(A->haloRecv[k]);
    cudaFree(A->haloOut[k]);
    cudaFree(A->haloIn[k]);
  }
  free_contact_buf(A->cb[0]);
  free(A->downloadSlots);
  cudaFreeHost(A->hCounts);
  cudaFreeHost(A->hostStage);
  if (A->scatterStage) cudaFree(A->scatterStage);
  for (int i = 0; i < 5; ++i) cudaEventDestroy(A->ev[i]);
  for (int i = 0; i < 2 * B2G_KT_MAX; ++i) cudaEventDestroy(A->ktEv[i]);
  cudaStreamDestroy(A->stream);
  if (A->vetoKeys) cudaFree(A->vetoKeys);
  if (A->vetoSeen) cudaFree(A->vetoSeen);
  cudaStreamDestroy(A->copyStream);
  cudaEventDestroy(A->evPacked);
  free(A);
  return B2G_OK;
}

__global__ void k_flag_worlds(int first, int count, const int* __restrict__ bworld, uint8_t* worldFlag) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) worldFlag[bworld[first + i]] = 1;
}

#define UP(dst, src, elems, type)                                                                      \
  if (src) CK(cudaMemcpyAsync((dst) + (size_t)first * (elems), (src), (size_t)count * (elems) * sizeof(type), \
                              cudaMemcpyHostToDevice, A->stream))

extern "C" int b2g_upload_bodies(b2gArena* A, int32_t first, int32_t count, const b2gBodyArrays* s) {
  if (!A || !s || first < 0 || count < 0) return B2G_ERR_INVALID;
  if (first + count > A->capBodies) {
    set_err("b2g_upload_bodies", "max_bodies exceeded");
    return B2G_ERR_CAPACITY;
  }
  CK(cudaSetDevice(A->device));
  UP((float*)A->pos, s->pos, 4, float);
  UP((float*)A->vel, s->vel, 4, float);
  UP((float*)A->xf, s->xf, 4, float);
  UP((float*)A->mass, s->mass, 4, float);
  UP((float*)A->center, s->center, 4, float);
  UP((float*)A->force, s->force, 4, float);
  UP(A->bflags, s->flags, 1, uint32_t);
  UP(A->bworld, s->world, 1, int32_t);
  CK(cudaStreamSynchronize(A->stream));
  if (first + count > A->nBodies) A->nBodies = first + count;
  A->aabbAllDirty = 1;
  A->islandsValid = 0;
  if ((s->mass || s->flags) && count > 0) {
    // Which constraints may share a colour depends on which bodies are movable, so the persistent colours of the
    // edited bodies' WORLD are dropped and that world is coloured afresh — that world only: an edit in one world
    // of a batched arena leaves the other worlds' colours, and with them their floats, alone.
    k_flag_worlds<<<div_up(count, 256), 256, 0, A->stream>>>(first, count, A->bworld, A->worldRecolour);
    CK(cudaGetLastError());
    A->recolourWorlds = 1;
  }
  if (s->mass) A->jointColourDirty = 1;  // which bodies a joint moves decides which joints may share a colour
  if (s->world) A->fixBaseDirty = 1;
  return B2G_OK;
}

extern "C" int b2g_upload_fixtures(b2gArena* A, int32_t first, int32_t count, const b2gFixtureArrays* s) {
  if (!A || !s || first < 0 || count < 0) return B2G_ERR_INVALID;
  if (first + count > A->capFixtures) {
    set_err("b2g_upload_fixtures", "max_fixtures exceeded");
    return B2G_ERR_CAPACITY;
  }
  CK(cudaSetDevice(A->device));
  UP(A->fBody, s->body, 1, int32_t);
  UP(A->fShapeOff, s->shape_off, 1, int32_t);
  UP(A->fTypeFlags, s->type_flags, 1, uint32_t);
  UP((uint32_t*)A->fFilter, s->filter, 2, uint32_t);
  UP((float*)A->fMaterial, s->material, 4, float);
  CK(cudaStreamSynchronize(A->stream));
  if (first + count > A->nFixtures) A->nFixtures = first + count;
  A->aabbAllDirty = 1;
  A->newFixtures = 1;
  A->fixBaseDirty = 1;
  A->islandsValid = 0;
  return B2G_OK;
}

extern "C" int b2g_upload_shapes(b2gArena* A, int32_t first, int32_t count, const float* quads) {
  if (!A || !quads || first < 0 || count < 0) return B2G_ERR_INVALID;
  if (first + count > A->capQuads) {
    set_err("b2g_upload_shapes", "max_shape_quads exceeded");
    return B2G_ERR_CAPACITY;
  }
  CK(cudaSetDevice(A->device));
  UP((float*)A->shapes, quads, 4, float);
  CK(cudaStreamSynchronize(A->stream));
  A->aabbAllDirty = 1;
  A->newFixtures = 1;
  A->fixBaseDirty = 1;
  return B2G_OK;
}

// ---- indexed uploads (b2WorldBatch: one edited row in each of many worlds) ------------------------------------
static int scatter_stage(b2gArena* A, size_t bytes) {
  if (bytes <= A->scatterStageBytes) return B2G_OK;
  if (A->scatterStage) {
    CK(cudaStreamSynchronize(A->stream));
    CK(cudaFree(A->scatterStage));
    A->scatterStage = nullptr;
  }
  size_t cap = A->scatterStageBytes ? A->scatterStageBytes : 1 << 16;
  while (cap < bytes) cap *= 2;
  CK(cudaMalloc((void**)&A->scatterStage, cap));
  A->scatterStageBytes = cap;
  return B2G_OK;
}
__global__ void k_scatter_bodies(int n, const int* __restrict__ index, const float4* __restrict__ rows, const uint32_t* __restrict__ flags,
                                 const int* __restrict__ world, float4* pos, float4* vel, float4* xf, float4* mass,
                                 float4* center, float4* force, uint32_t* bflags, int* bworld, uint8_t* worldFlag) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int b = index[i];
  pos[b] = rows[i];
  vel[b] = rows[n + i];
  xf[b] = rows[2 * n + i];
  mass[b] = rows[3 * n + i];
  center[b] = rows[4 * n + i];
  force[b] = rows[5 * n + i];
  bflags[b] = flags[i];
  bworld[b] = world[i];
  worldFlag[world[i]] = 1;  // mass / type may have changed: this world is coloured afresh (see b2g_upload_bodies)
}
__global__ void k_scatter_fixtures(int n, const int* __restrict__ index, const int* __restrict__ body, const int* __restrict__ off,
                                   const uint32_t* __restrict__ tf, const uint2* __restrict__ filter,
                                   const float4* __restrict__ material, int* fBody, int* fShapeOff, uint32_t* fTypeFlags,
                                   uint2* fFilter, float4* fMaterial) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int f = index[i];
  fBody[f] = body[i];
  fShapeOff[f] = off[i];
  fTypeFlags[f] = tf[i];
  fFilter[f] = filter[i];
  fMaterial[f] = material[i];
}
__global__ void k_scatter_quads(int n, const int* __restrict__ index, const float4* __restrict__ quads, float4* shapes) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) shapes[index[i]] = quads[i];
}

extern "C" int b2g_upload_bodies_indexed(b2gArena* A, int32_t count, const int32_t* index, const b2gBodyArrays* s) {
  if (!A || !s || count < 0 || (count > 0 && (!index || !s->pos || !s->vel || !s->xf || !s->mass || !s->center || !s->force ||
                                              !s->flags || !s->world)))
    return B2G_ERR_INVALID;
  if (count == 0) return B2G_OK;
  int top = 0;
  for (int i = 0; i < count; ++i) {
    if (index[i] < 0 || index[i] >= A->capBodies || s->world[i] < 0 || s->world[i] >= A->numWorlds) return B2G_ERR_INVALID;
    top = index[i] + 1 > top ? index[i] + 1 : top;
  }
  CK(cudaSetDevice(A->device));
  const size_t n = (size_t)count, rowBytes = 6 * 16 + 4 + 4 + 4;
  int rc = scatter_stage(A, n * rowBytes);
  if (rc) return rc;
  float4* dRows = (float4*)A->scatterStage;
  uint32_t* dFlags = (uint32_t*)(dRows + 6 * n);
  int* dWorld = (int*)(dFlags + n);
  int* dIndex = dWorld + n;
  const float* cols[6] = {s->pos, s->vel, s->xf, s->mass, s->center, s->force};
  for (int k = 0; k < 6; ++k)
    CK(cudaMemcpyAsync(dRows + k * n, cols[k], n * 16, cudaMemcpyHostToDevice, A->stream));
  CK(cudaMemcpyAsync(dFlags, s->flags, n * 4, cudaMemcpyHostToDevice, A->stream));
  CK(cudaMemcpyAsync(dWorld, s->world, n * 4, cudaMemcpyHostToDevice, A->stream));
  CK(cudaMemcpyAsync(dIndex, index, n * 4, cudaMemcpyHostToDevice, A->stream));
  k_scatter_bodies<<<div_up(count, 256), 256, 0, A->stream>>>(count, dIndex, dRows, dFlags, dWorld, A->pos, A->vel, A->xf, A->mass,
                                                              A->center, A->force, A->bflags, A->bworld, A->worldRecolour);
  CK(cudaGetLastError());
  A->launches++;
  if (top > A->nBodies) A->nBodies = top;
  A->aabbAllDirty = 1;
  A->islandsValid = 0;
  A->recolourWorlds = 1;
  A->jointColourDirty = 1;
  A->fixBaseDirty = 1;
  return B2G_OK;
}

extern "C" int b2g_upload_fixtures_indexed(b2gArena* A, int32_t count, const int32_t* index, const b2gFixtureArrays* s) {
  if (!A || !s || count < 0 || (count > 0 && (!index || !s->body || !s->shape_off || !s->type_flags || !s->filter || !s->material)))
    return B2G_ERR_INVALID;
  if (count == 0) return B2G_OK;
  int top = 0;
  for (int i = 0; i < count; ++i) {
    if (index[i] < 0 || index[i] >= A->capFixtures) return B2G_ERR_INVALID;
    top = index[i] + 1 > top ? index[i] + 1 : top;
  }
  CK(cudaSetDevice(A->device));
  const size_t n = (size_t)count;
  int rc = scatter_stage(A, n * (16 + 8 + 4 * 4));
  if (rc) return rc;
  float4* dMat = (float4*)A->scatterStage;
  uint2* dFilter = (uint2*)(dMat + n);
  int* dBody = (int*)(dFilter + n);
  int* dOff = dBody + n;
  uint32_t* dTf = (uint32_t*)(dOff + n);
  int* dIndex = (int*)(dTf + n);
  CK(cudaMemcpyAsync(dMat, s->material, n * 16, cudaMemcpyHostToDevice, A->stream));
  CK(cudaMemcpyAsync(dFilter, s->filter, n * 8, cudaMemcpyHostToDevice, A->stream));
  CK(cudaMemcpyAsync(dBody, s->body, n * 4, cudaMemcpyHostToDevice, A->stream));
  CK(cudaMemcpyAsync(dOff, s->shape_off, n * 4, cudaMemcpyHostToDevice, A->stream));
  CK(cudaMemcpyAsync(dTf, s->type_flags, n * 4, cudaMemcpyHostToDevice, A->stream));
  CK(cudaMemcpyAsync(dIndex, index, n * 4, cudaMemcpyHostToDevice, A->stream));
  k_scatter_fixtures<<<div_up(count, 256), 256, 0, A->stream>>>(count, dIndex, dBody, dOff, dTf, dFilter, dMat, A->fBody, A->fShapeOff,
                                                                A->fTypeFlags, (uint2*)A->fFilter, A->fMaterial);
  CK(cudaGetLastError());
  A->launches++;
  if (top > A->nFixtures) A->nFixtures = top;
  A->aabbAllDirty = 1;
  A->newFixtures = 1;
  A->fixBaseDirty = 1;
  A->islandsValid = 0;
  return B2G_OK;
}

extern "C" int b2g_upload_shapes_indexed(b2gArena* A, int32_t count, const int32_t* index, const float* quads) {
  if (!A || count < 0 || (count > 0 && (!index || !quads))) return B2G_ERR_INVALID;
  if (count == 0) return B2G_OK;
  for (int i = 0; i < count; ++i)
    if (index[i] < 0 || index[i] >= A->capQuads) return B2G_ERR_INVALID;
  CK(cudaSetDevice(A->device));
  const size_t n = (size_t)count;
  int rc = scatter_stage(A, n * 20);
  if (rc) return rc;
  float4* dQuads = (float4*)A->scatterStage;
  int* dIndex = (int*)(dQuads + n);
  CK(cudaMemcpyAsync(dQuads, quads, n * 16, cudaMemcpyHostToDevice, A->stream));
  CK(cudaMemcpyAsync(dIndex, index, n * 4, cudaMemcpyHostToDevice, A->stream));
  k_scatter_quads<<<div_up(count, 256), 256, 0, A->stream>>>(count, dIndex, dQuads, A->shapes);
  CK(cudaGetLastError());
  A->launches++;
  A->aabbAllDirty = 1;
  A->newFixtures = 1;
  A->fixBaseDirty = 1;
  return B2G_OK;
}

extern "C" int b2g_upload_joints(b2gArena* A, int32_t first, int32_t count, const b2gJointArrays* s) {
  if (!A || !s || first < 0 || count < 0) return B2G_ERR_INVALID;
  if (first + count > A->capJoints) {
    set_err("b2g_upload_joints", "max_joints exceeded");
    return B2G_ERR_CAPACITY;
  }
  CK(cudaSetDevice(A->device));
  UP((int32_t*)A->jBodies, s->bodies, 2, int32_t);
  UP((float*)A->jAnchors, s->anchors, 4, float);
  if (s->params) {
    // split [n][12] into three float4 planes
    std::vector<float> p0((size_t)count * 4), p1((size_t)count * 4), p2((size_t)count * 4);
    for (int i = 0; i < count; ++i) {
      for (int k = 0; k < 4; ++k) {
        p0[(size_t)i * 4 + k] = s->params[(size_t)i * 12 + k];
        p1[(size_t)i * 4 + k] = s->params[(size_t)i * 12 + 4 + k];
        p2[(size_t)i * 4 + k] = s->params[(size_t)i * 12 + 8 + k];
      }
    }
    CK(cudaMemcpyAsync((float*)A->jParams2 + (size_t)first * 4, p2.data(), p2.size() * sizeof(float),
                       cudaMemcpyHostToDevice, A->stream));
    CK(cudaMemcpyAsync((float*)A->jParams0 + (size_t)first * 4, p0.data(), p0.size() * sizeof(float),
                       cudaMemcpyHostToDevice, A->stream));
    CK(cudaMemcpyAsync((float*)A->jParams1 + (size_t)first * 4, p1.data(), p1.size() * sizeof(float),
                       cudaMemcpyHostToDevice, A->stream));
    CK(cudaStreamSynchronize(A->stream));
  }
  {
    std::vector<float> st((size_t)count * 4, 0.0f), up((size_t)count, 0.0f);
    if (s->state) {
      for (int i = 0; i < count; ++i) {
        for (int k = 0; k < 4; ++k) st[(size_t)i * 4 + k] = s->state[(size_t)i * 5 + k];
        up[i] = s->state[(size_t)i * 5 + 4];
      }
    }
    if (s->state || s->params) {  // a joint created by this call starts from zero impulses
      CK(cudaMemcpyAsync((float*)A->jState + (size_t)first * 4, st.data(), st.size() * sizeof(float),
                         cudaMemcpyHostToDevice, A->stream));
      CK(cudaMemcpyAsync(A->jUpper + first, up.data(), up.size() * sizeof(float), cudaMemcpyHostToDevice, A->stream));
    }
    CK(cudaStreamSynchronize(A->stream));
  }
  if (first + count > A->nJoints) A->nJoints = first + count;
  A->islandsValid = 0;
  if (s->params || s->bodies) {
    // the joint table changed: contacts between the joined bodies are re-filtered before the next
    // Collide (b2World::CreateJoint flags them, b2_world.cpp:307-323)
    A->jointFilterDirty = 1;
    A->jointColourDirty = 1;
    A->aabbAllDirty = 1;
    A->newFixtures = 1;
    A->fixBaseDirty = 1;
  }
  return B2G_OK;
}

extern "C" int b2g_download_joints(b2gArena* A, int32_t first, int32_t count, float* state) {
  if (!A || !state || first < 0 || count < 0 || first + count > A->nJoints) return B2G_ERR_INVALID;
  if (count == 0) return B2G_OK;
  CK(cudaSetDevice(A->device));
  std::vector<float> st((size_t)count * 4), up((size_t)count);
  CK(cudaMemcpyAsync(st.data(), (float*)A->jState + (size_t)first * 4, st.size() * sizeof(float),
                     cudaMemcpyDeviceToHost, A->stream));
  CK(cudaMemcpyAsync(up.data(), A->jUpper + first, up.size() * sizeof(float), cudaMemcpyDeviceToHost, A->stream));
  CK(cudaStreamSynchronize(A->stream));
  for (int i = 0; i < count; ++i) {
    for (int k = 0; k < 4; ++k) state[(size_t)i * 5 + k] = st[(size_t)i * 4 + k];
    state[(size_t)i * 5 + 4] = up[i];
  }
  return B2G_OK;
}

extern "C" int b2g_set_counts(b2gArena* A, int32_t nb, int32_t nf, int32_t nj) {
  if (!A || nb < 0 || nf < 0 || nj < 0 || nb > A->capBodies || nf > A->capFixtures || nj > A->capJoints)
    return B2G_ERR_INVALID;
  A->nBodies = nb;
  A->nFixtures = nf;
  if (nj != A->nJoints) {
    A->jointFilterDirty = 1;
    A->jointColourDirty = 1;
    A->newFixtures = 1;
    A->fixBaseDirty = 1;
  }
  A->nJoints = nj;
  A->aabbAllDirty = 1;
  A->islandsValid = 0;
  return B2G_OK;
}

__global__ void k_merge_forces(int first, int count, const float4* __restrict__ stage, float4* force) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  float4 in = stage[i];
  float4 f = force[first + i];
  force[first + i] = make_float4(in.x, in.y, in.z, f.w);  // column 3 is the device-owned sleep timer
}

extern "C" int b2g_upload_forces(b2gArena* A, int32_t first, int32_t count, const float* force) {
  if (!A || !force || first < 0 || count < 0 || first + count > A->nBodies) return B2G_ERR_INVALID;
  if (count == 0) return B2G_OK;
  CK(cudaSetDevice(A->device));
  // one linear copy + a merge kernel: a strided copy of 12-byte rows costs a DMA descriptor per body
  CK(cudaMemcpyAsync(A->forceStage, force, (size_t)count * 16, cudaMemcpyHostToDevice, A->stream));
  k_merge_forces<<<div_up(count, 256), 256, 0, A->stream>>>(first, count, A->forceStage, A->force);
  CK(cudaGetLastError());
  A->launches++;
  return B2G_OK;
}

// ---------------------------------------------------------------------------------------------
// broadphase + contact list rebuild: b2ContactManager::FindNewContacts + RemoveDeadContacts
// ---------------------------------------------------------------------------------------------
static int read_counts(b2gArena* A) {
  CK(cudaMemcpyAsync(A->hCounts, A->dCounts, sizeof(StepCounts), cudaMemcpyDeviceToHost, A->stream));
  CK(cudaStreamSynchronize(A->stream));
  return B2G_OK;
}

static int reset_bounds(b2gArena* A) {
  CK(cudaMemsetAsync(A->dCounts->boundsLo, 0xff, sizeof(unsigned int) * 2, A->stream));
  CK(cudaMemsetAsync(A->dCounts->boundsHi, 0x00, sizeof(unsigned int) * 2, A->stream));
  return B2G_OK;
}

static WideBvh wide_tree(b2gArena* A) {
  WideBvh T;
  wide_bvh_layout(T, A->bvhLeaves);
  T.box = A->bvhBox;
  T.key = A->bvhKey;
  T.done = A->bvhDone;
  return T;
}

static int find_new_contacts(b2gArena* A, int recordEvents) {
  const int nf = A->nFixtures;
  ContactBuf& C = A->cb[0];
  const int nSlots = A->nContacts;  // slot high-water mark
  int nNew = 0;
  A->newFixtures = 0;
  if (nf > 0) {
    // The leaf order is re-sorted (Morton keys + radix sort) when fixtures were added / edited, when
    // it has degraded, or every B2G_BVH_REBUILD_PERIOD steps; in between only the boxes are refit.  The reported
    // pair set is exact either way — only traversal cost depends on tree quality.
    if (A->jointFilterDirty) {
      const int nj = A->nJoints;
      CK(cudaMemsetAsync(A->bodyNoCollide, 0, (size_t)A->capBodies, A->stream));
      if (nj > 0) {
        LAUNCH(A, KC_BP_BUILD, nj, k_joint_filter_build, div_up(nj, 256), 256, nj, A->jBodies, A->jParams1, A->ncKeys,
               A->bodyNoCollide);
        size_t tbj = A->cubTempBytes;
        CK(cub::DeviceRadixSort::SortKeys(A->cubTemp, tbj, A->ncKeys, A->ncKeysSorted, nj, 0, 64, A->stream));
      }
      A->hash.ncKeys = A->ncKeysSorted;
      A->hash.ncCount = nj;  // collideConnected joints sort last as ~0 keys, which match no body pair
      A->jointFilterDirty = 0;
      A->aabbAllDirty = A->aabbAllDirty ? A->aabbAllDirty : 1;  // leaf records carry the per-body mark
    }
    // re-sort when fixtures changed, when the order has aged, or as soon as the walks have become
    // 15 % longer than they were on the freshly sorted order (bodies have moved past each other)
    const bool degraded = A->bvhAge >= 2 && A->bvhVisitsLast > 1.15f * A->bvhVisitsFresh + 0.5f;
    const bool rebuild = A->aabbAllDirty != 0 || A->bvhLeaves != nf || A->bvhAge >= B2G_BVH_REBUILD_PERIOD || degraded;
    if (rebuild) {
      int rc = reset_bounds(A);
      if (rc) return rc;
      LAUNCH(A, KC_BP_BUILD, nf, k_update_aabbs, div_up(nf, 256), 256, nf, A->fBody, A->fShapeOff, A->fTypeFlags,
             A->shapes, A->bflags, A->xf, A->fAabb, A->fRadius, A->aabbAllDirty, A->dCounts);
      A->aabbAllDirty = 0;
      LAUNCH(A, KC_BP_BUILD, nf, k_morton_keys, div_up(nf, 256), 256, nf, A->fAabb, A->fTypeFlags, A->fBody, A->bworld,
             A->dCounts, A->mortonKeys, A->leafFixture, A->numWorlds);
      size_t tb = A->cubTempBytes;
      TIMED(A, KC_SORT_SCAN, nf,
            CK(cub::DeviceRadixSort::SortPairs(A->cubTemp, tb, A->mortonKeys, A->mortonKeysSorted, A->leafFixture,
                                               A->leafFixtureSorted, nf, 0,
                                               32 + (A->numWorlds > 1 ? A->worldBits : 0), A->stream)));
      LAUNCH(A, KC_BP_BUILD, nf, k_leaf_gather, div_up(nf, 256), 256, nf, A->leafFixtureSorted, A->mortonKeysSorted,
             A->fAabb, A->fBody, A->fTypeFlags, A->fFilter, A->bflags, A->bodyNoCollide, A->leafBox, A->leafInfo, A->leafKey,
             A->worldFirst, A->worldLast, A->numWorlds);
      A->bvhLeaves = nf;
      A->bvhAge = 0;
    }
    const WideBvh T = wide_tree(A);
    if (rebuild) {
      LAUNCH(A, KC_BP_BUILD, nf, k_wide_refit<false>, div_up(nf, 256), 256, T, A->leafFixtureSorted, A->fBody,
             A->fShapeOff, A->fTypeFlags, A->shapes, A->bflags, A->xf, A->fAabb, A->leafBox, A->leafKey);
    } else {
      LAUNCH(A, KC_BP_BUILD, nf, k_wide_refit<true>, div_up(nf, 256), 256, T, A->leafFixtureSorted, A->fBody,
             A->fShapeOff, A->fTypeFlags, A->shapes, A->bflags, A->xf, A->fAabb, A->leafBox, A->leafKey);
      A->bvhAge++;
    }
    if (nf > 1) {
      static int lanesOverride = -1;  // B2G_BP_LANES=1|4 (measurements)
      if (lanesOverride < 0) {
        const char* e = getenv("B2G_BP_LANES");
        lanesOverride = e ? atoi(e) : 0;
      }
      if (lanesOverride == 4 || (lanesOverride == 0 && nf < 65536))
        LAUNCH(A, KC_BP_TRAVERSE, nf, k_bp_traverse<4>, div_up(nf * 4, 128), 128, T, A->leafBox, A->leafInfo, A->leafKey,
               A->worldFirst, A->worldLast, A->mortonKeysSorted, A->numWorlds, A->hash, A->persist, A->pairKeys,
               A->capContacts, A->dCounts);
      else
        LAUNCH(A, KC_BP_TRAVERSE, nf, k_bp_traverse<1>, div_up(nf, 128), 128, T, A->leafBox, A->leafInfo, A->leafKey,
               A->worldFirst, A->worldLast, A->mortonKeysSorted, A->numWorlds, A->hash, A->persist, A->pairKeys,
               A->capContacts, A->dCounts);
    }
  }
  // retire contacts whose pair was not re-reported (all of them when there are no fixtures left)
  if (nSlots > 0) {
    LAUNCH(A, KC_CONTACT_MERGE, nSlots, k_contact_sweep, div_up(nSlots, 256), 256, nSlots, C, A->persist, A->hash,
           A->fTypeFlags, A->bflags, A->force, A->freeStack, A->dFreeTop, A->dCounts, recordEvents, A->endEvents,
           A->capContacts, A->island, A->islandDirty);
  }
  // the end-of-broadphase readback: how many pairs are new, how many contacts died
  CK(cudaMemcpyAsync(A->hCounts, A->dCounts, sizeof(StepCounts), cudaMemcpyDeviceToHost, A->stream));
  CK(cudaMemcpyAsync(&A->hCounts->freeTopRead, A->dFreeTop, sizeof(int), cudaMemcpyDeviceToHost, A->stream));
  CK(cudaStreamSynchronize(A->stream));
  nNew = A->hCounts->numPairs;
  A->lastNewPairs = nNew <= A->capContacts ? nNew : 0;
  if (nf > 1) {
    A->bvhVisitsLast = (float)((double)A->hCounts->bpVisits / nf);
    if (A->bvhAge == 0) A->bvhVisitsFresh = A->bvhVisitsLast;
  }
  const int freeTop = A->hCounts->freeTopRead;
  A->tombstones += A->hCounts->numDead;
  const int appended = nNew > freeTop ? nNew - freeTop : 0;
  if (nNew > A->capContacts || nSlots + appended > A->capContacts) {
    char msg[160];
    snprintf(msg, sizeof(msg), "broadphase needs %d contact slots (%d new pairs), max_contacts is %d",
             nSlots + appended, nNew, A->capContacts);
    set_err("b2g_step", msg);
    // the sweep above already ran: keep the live count right so the caller can download the
    // surviving contacts, grow, and let the next pair refresh find the pairs dropped here
    A->nAlive -= A->hCounts->numDead;
    return B2G_ERR_CAPACITY;
  }
  A->nAlive -= A->hCounts->numDead;
  // Tombstones keep their cell (a dead pair that comes back revives it) and slow probes down: rebuild the
  // table from the live contacts once they rival the live entries — BEFORE this step's inserts, so that
  // live + dead + new entries stay below half of the table whatever died this step.
  if (A->tombstones > (int)((A->hash.mask + 1) / 4) ||
      (long long)A->tombstones + A->nAlive + nNew > (long long)(A->hash.mask + 1) / 2) {
    CK(cudaMemsetAsync(A->hash.keys, 0xff, (size_t)(A->hash.mask + 1) * 8, A->stream));
    CK(cudaMemsetAsync(A->hash.vals, 0xff, (size_t)(A->hash.mask + 1) * 4, A->stream));
    if (nSlots > 0)
      LAUNCH(A, KC_CONTACT_MERGE, nSlots, k_hash_rebuild, div_up(nSlots, 256), 256, nSlots, C, A->hash);
    A->tombstones = 0;
  }
  if (nNew > 0) {
    LAUNCH(A, KC_CONTACT_MERGE, nNew, k_contact_insert, div_up(nNew, 256), 256, nNew, A->pairKeys, freeTop, nSlots, C,
           A->persist, A->hash, A->freeStack, A->dFreeTop, A->fBody, A->fTypeFlags, A->fMaterial);
  }
  A->nContacts = nSlots + appended;
  A->nAlive += nNew;
  return B2G_OK;
}

extern "C" int b2g_find_new_contacts(b2gArena* A) {
  if (!A) return B2G_ERR_INVALID;
  CK(cudaSetDevice(A->device));
  CK(cudaMemsetAsync(A->dCounts, 0, sizeof(StepCounts), A->stream));
  int rc = find_new_contacts(A, 0);
  if (rc) return rc;
  CK(cudaStreamSynchronize(A->stream));
  return B2G_OK;
}

// ---------------------------------------------------------------------------------------------
// the step
// ---------------------------------------------------------------------------------------------
// ---------------------------------------------------------------------------------------------
// Solve, list-order / per-colour-launch path: used by B2G_SOLVER_SEQUENTIAL (parity vehicle).  One
// kernel per reference loop, constraints addressed through global body arrays.
// ---------------------------------------------------------------------------------------------
static JointArraysDev joint_views(b2gArena* A) {
  JointArraysDev J;
  J.bodies = A->jBodies;
  J.anchors = A->jAnchors;
  J.params0 = A->jParams0;
  J.params1 = A->jParams1;
  J.params2 = A->jParams2;
  J.state = A->jState;
  J.upper = A->jUpper;
  J.work = A->jWork;
  J.h = A->stepDt;
  return J;
}
static JointWalk joint_walk(b2gArena* A, int onlyBig) {
  JointWalk W;
  W.nj = A->nJoints;
  W.onlyBig = onlyBig;
  W.bflags = A->bflags;
  W.island = A->island;
  W.islandAwake = A->islandAwake;
  W.bodySlot = A->bodySlot;
  W.order = nullptr;
  W.norder = 0;
  W.sorted = A->jSorted;
  W.cstart = A->jCstart;
  return W;
}

struct SolveOut {
  int numActive = 0, numColours = 0, numOverflow = 0, rounds = 0;
};

static int solve_legacy(b2gArena* A, const b2gStepParams* P, SolveOut& out) {
  const int nb = A->nBodies, nc = A->nContacts, nj = A->nJoints;
  const float h = P->dt;
  const float dtRatio = A->invDt0 * h;
  ContactBuf& C = A->cb[0];
  int numActive = 0, numColours = 0, numOverflow = 0, rounds = 0;
  int colourFirst[B2G_MAX_COLOURS + 2];
  memset(colourFirst, 0, sizeof(colourFirst));
  {
    // ---- islands ---------------------------------------------------------------------
    LAUNCH(A, KC_ISLANDS, nb, k_body_begin, div_up(nb, 256), 256, nb, A->bflags, A->force, A->islandParent, A->islandAwake,
           A->island, A->islandDirty, A->islandsValid, A->islandWasBig, 1,
           A->islandMinSleep, A->islandPen, A->capBodies, P->position_iterations, A->colourMask, A->bodyBest,
           A->islandCount, A->islandCursor, A->binFirst, A->binEnd, 0, A->bucketCount, 0);
    if (nc > 0) LAUNCH(A, KC_ISLANDS, nc, k_island_union, div_up(nc, 256), 256, nc, C, A->bflags, A->fTypeFlags, A->islandParent);
    if (nj > 0) LAUNCH(A, KC_ISLANDS, nj, k_island_union_joints, div_up(nj, 256), 256, nj, A->jBodies, A->bflags, A->islandParent);
    LAUNCH(A, KC_ISLANDS, nb, k_island_flatten, div_up(nb, 256), 256, nb, A->bflags, A->islandParent, A->island, A->islandAwake,
           A->islandCount, A->dCounts, A->islandDirty);
    LAUNCH(A, KC_INTEGRATE, nb, k_integrate_velocities, div_up(nb, 256), 256, nb, A->bflags, A->island, A->islandAwake, A->vel, A->mass,
           A->center, A->force, h, make_float2(P->gravity_x, P->gravity_y), A->dCounts, nullptr, 0);

    // ---- constraint list + colouring -------------------------------------------------
    if (nc > 0) {
      LAUNCH(A, KC_COLOUR, nc, k_mark_active, div_up(nc, 256), 256, nc, C, A->fTypeFlags, A->island, A->islandAwake, A->activeFlag,
             A->recolour | A->recolourWorlds);
      A->recolour = 0;
      if (A->recolourWorlds) {
        CK(cudaMemsetAsync(A->worldRecolour, 0, (size_t)A->numWorlds, A->stream));
        A->recolourWorlds = 0;
      }
      size_t tb = A->cubTempBytes;
      TIMED(A, KC_SORT_SCAN, nc,
            CK(cub::DeviceSelect::Flagged(A->cubTemp, tb, thrust::counting_iterator<int>(0), A->activeFlag,
                                          A->activeList, &A->dCounts->numActive, nc, A->stream)));
      const int* nAct = &A->dCounts->numActive;
      if (P->solver_mode != B2G_SOLVER_SEQUENTIAL) {
        int grid = div_up(nc, 256);
        if (grid > 148 * 8) grid = 148 * 8;
        LAUNCH(A, KC_COLOUR, nc, k_colour_begin, grid, 256, nAct, A->activeList, C, A->mass, A->colourMask, A->dCounts);
        int round = 0;
        int batch = A->roundsHint;
        while (true) {
          CK(cudaMemsetAsync(&A->dCounts->remaining, 0, sizeof(int), A->stream));
          for (int r = 0; r < batch; ++r, ++round) {
            LAUNCH(A, KC_COLOUR, nc, k_colour_propose, grid, 256, nAct, A->activeList, C, A->mass, A->bodyBest, round);
            LAUNCH(A, KC_COLOUR, nc, k_colour_commit, grid, 256, nAct, A->activeList, C, A->mass, A->colourMask, A->bodyBest, round,
                   A->dCounts, r == batch - 1);
          }
          int rc = read_counts(A);
          if (rc) return rc;
          if (A->hCounts->remaining == 0) break;
          if (round > 250) {
            set_err("b2g_step", "graph colouring did not converge");
            return B2G_ERR_CUDA;
          }
          batch = 4;
        }
        rounds = round;
        // adapt the first batch to what was actually needed: a settled scene colours nothing new
        // (one probing round), a cold start needs ~10
        {
          int useful = A->hCounts->lastUsefulRound;
          A->roundsHint = useful < 1 ? 1 : (useful + 1 < 16 ? useful + 1 : 16);
        }
        numActive = A->hCounts->numActive;
        int acc = 0;
        for (int c = 0; c <= B2G_MAX_COLOURS; ++c) {
          colourFirst[c] = acc;
          acc += A->hCounts->colourCount[c];
          if (c < B2G_MAX_COLOURS && A->hCounts->colourCount[c] > 0) numColours = c + 1;
        }
        colourFirst[B2G_MAX_COLOURS + 1] = acc;
        numOverflow = A->hCounts->colourCount[B2G_MAX_COLOURS];
        if (numActive > 0) {
          LAUNCH(A, KC_COLOUR, numActive, k_colour_keys, grid, 256, nAct, A->activeList, C, A->colourKey);
          size_t tb2 = A->cubTempBytes;
          TIMED(A, KC_SORT_SCAN, numActive,
                CK(cub::DeviceRadixSort::SortPairs(A->cubTemp, tb2, A->colourKey, A->colourKeySorted, A->activeList,
                                                   A->sortedList, numActive, 0, 8, A->stream)));
        }
      } else {
        int rc = read_counts(A);
        if (rc) return rc;
        numActive = A->hCounts->numActive;
        if (numActive > 0) {
          // contact slots are unordered; the sequential mode's list order is ascending pair key
          if (A->seqOrderActive) {  // explicit order imposed by b2g_set_sequential_order (this step only)
            LAUNCH(A, KC_COLOUR, numActive, k_gather_order, div_up(numActive, 256), 256, numActive, A->activeList,
                   A->orderKey, A->seqKeys);
            A->seqOrderActive = 0;
          } else {
            LAUNCH(A, KC_COLOUR, numActive, k_gather_keys, div_up(numActive, 256), 256, numActive, A->activeList, C,
                   A->seqKeys);
          }
          size_t tb3 = A->cubTempBytes;
          CK(cub::DeviceRadixSort::SortPairs(A->cubTemp, tb3, A->seqKeys, A->seqKeys + A->capContacts, A->activeList,
                                             A->sortedList, numActive, 0, 64, A->stream));
        }
      }
    }

    // ---- contact solver --------------------------------------------------------------
    SolverPlanes& S = A->planes;
    const bool coloured = P->solver_mode != B2G_SOLVER_SEQUENTIAL;
    JointWalk JW = joint_walk(A, 0);
    if (A->jointOrderActive && P->solver_mode == B2G_SOLVER_SEQUENTIAL) {  // b2g_set_sequential_joint_order, this step only
      JW.order = A->jointOrder;
      JW.norder = A->jointOrderCount;
      A->jointOrderActive = 0;
    }
    const JointArraysDev JV = joint_views(A);
    const float invH = h > 0.0f ? 1.0f / h : 0.0f;
    if (numActive > 0) {
      LAUNCH(A, KC_PREPARE, numActive, k_prepare, div_up(numActive, 128), 128, 0, numActive, A->sortedList, C, A->fRadius, A->bflags, A->island,
             S, A->croot, A->pos, A->vel, A->mass, A->center, dtRatio, P->warm_starting);
      if (P->warm_starting) {
        if (coloured) {
          for (int c = 0; c < numColours; ++c) {
            int n = colourFirst[c + 1] - colourFirst[c];
            if (n > 0) LAUNCH(A, KC_WARM_START, n, k_warm_start, div_up(n, 256), 256, colourFirst[c], colourFirst[c + 1], S, A->vel);
          }
          if (numOverflow > 0)
            LAUNCH(A, KC_WARM_START, numOverflow, k_warm_start_seq, 1, 1, colourFirst[B2G_MAX_COLOURS], colourFirst[B2G_MAX_COLOURS + 1], S,
                   A->vel);
        } else {
          LAUNCH(A, KC_WARM_START, numActive, k_warm_start_seq, 1, 1, 0, numActive, S, A->vel);
        }
      }
    }
    // joints: InitVelocityConstraints after the contacts' warm start (b2_island.cpp:323-325)
    if (nj > 0)
      LAUNCH(A, KC_WARM_START, nj, k_joints_init_seq, 1, 1, JW, JV, A->pos, A->vel, A->mass, A->center, dtRatio,
             P->warm_starting);
    for (int it = 0; it < P->velocity_iterations; ++it) {
      // joints first, then contacts (b2_island.cpp:330-338)
      if (nj > 0) LAUNCH(A, KC_SOLVE_VELOCITY, nj, k_joints_velocity_seq, 1, 1, JW, JV, A->vel, h, invH);
      if (numActive > 0) {
        if (coloured) {
          for (int c = 0; c < numColours; ++c) {
            int n = colourFirst[c + 1] - colourFirst[c];
            if (n > 0)
              LAUNCH(A, KC_SOLVE_VELOCITY, n, k_solve_velocity, div_up(n, 256), 256, colourFirst[c], colourFirst[c + 1], S, A->vel);
          }
          if (numOverflow > 0)
            LAUNCH(A, KC_SOLVE_VELOCITY, numOverflow, k_solve_velocity_seq, 1, 1, colourFirst[B2G_MAX_COLOURS], colourFirst[B2G_MAX_COLOURS + 1], S,
                   A->vel);
        } else {
          LAUNCH(A, KC_SOLVE_VELOCITY, numActive, k_solve_velocity_seq, 1, 1, 0, numActive, S, A->vel);
        }
      }
    }
    if (numActive > 0)
      LAUNCH(A, KC_STORE_IMPULSES, numActive, k_store_impulses, div_up(numActive, 256), 256, 0, numActive, S, C);
    LAUNCH(A, KC_INTEGRATE, nb, k_integrate_positions, div_up(nb, 256), 256, nb, A->bflags, A->island, A->islandAwake, A->pos, A->vel,
           h, nullptr, 0);
    for (int it = 0; it < P->position_iterations; ++it) {
      // contacts first, then joints (b2_island.cpp:392-401)
      if (numActive > 0) {
        if (coloured) {
          for (int c = 0; c < numColours; ++c) {
            int n = colourFirst[c + 1] - colourFirst[c];
            if (n > 0)
              LAUNCH(A, KC_SOLVE_POSITION, n, k_solve_position, div_up(n, 256), 256, colourFirst[c], colourFirst[c + 1], S, A->pos,
                     A->croot, A->islandPen, A->capBodies, it);
          }
          if (numOverflow > 0)
            LAUNCH(A, KC_SOLVE_POSITION, numOverflow, k_solve_position_seq, 1, 1, colourFirst[B2G_MAX_COLOURS], colourFirst[B2G_MAX_COLOURS + 1], S,
                   A->pos, A->croot, A->islandPen, A->capBodies, it);
        } else {
          LAUNCH(A, KC_SOLVE_POSITION, numActive, k_solve_position_seq, 1, 1, 0, numActive, S, A->pos, A->croot, A->islandPen, A->capBodies, it);
        }
      }
      if (nj > 0)
        LAUNCH(A, KC_SOLVE_POSITION, nj, k_joints_position_seq, 1, 1, JW, JV, A->pos, A->islandPen, A->capBodies, it);
    }
    LAUNCH(A, KC_FINALIZE, nb, k_finalize_bodies, div_up(nb, 256), 256, nb, A->bflags, A->island, A->islandAwake, A->pos, A->vel,
           A->center, A->xf, A->force, A->islandMinSleep, h, P->allow_sleep, nullptr, 0);
    LAUNCH(A, KC_FINALIZE, nb, k_sleep_and_clear, div_up(nb, 256), 256, nb, A->bflags, A->island, A->islandAwake, A->islandMinSleep,
           A->islandPen, A->capBodies, P->position_iterations, A->vel, A->force, P->allow_sleep, P->clear_forces,
           A->dCounts, nullptr, 0);
  }
  out.numActive = numActive;
  out.numColours = numColours;
  out.numOverflow = numOverflow;
  out.rounds = rounds;
  return B2G_OK;
}

// ---------------------------------------------------------------------------------------------
// Solve, production path: islands packed into bins, one thread block per bin solves it start to
// finish in shared memory (b2g_fused.cuh); islands too large for a tile take the per-colour
// whole-GPU kernels.  One mid-step readback (colouring convergence + size of the big set).
// ---------------------------------------------------------------------------------------------
#define B2G_BIG_ISLAND 1024  // bodies; larger islands do not go through a shared-memory tile

static int solve_fused(b2gArena* A, const b2gStepParams* P, SolveOut& out) {
  const int nb = A->nBodies, nc = A->nContacts, nj = A->nJoints;
  const float h = P->dt;
  const float dtRatio = A->invDt0 * h;
  ContactBuf& C = A->cb[0];
  SolverPlanes& S = A->planes;
  if (nj > 0 && A->jointColourDirty) {
    // joints that share no movable body get the same colour and are solved side by side (b2g_fused.cuh)
    LAUNCH(A, KC_COLOUR, nj, k_joint_colour, 1, B2G_JOINT_COLOUR_THREADS, nj, nb, A->jBodies, A->jParams1, A->mass,
           A->jBodyMask, A->jBodyBest, A->jColour, A->jSorted, A->jCstart);
    A->jointColourDirty = 0;
  }
  // Tile capacity follows the largest island of the previous step (x1.5 + slack): small islands
  // leave shared memory for the constraint planes and a second resident block per SM.  An island
  // that outgrows the cap within one step is simply routed to the big path for that step.
  int bigThr = A->lastMaxIsland + A->lastMaxIsland / 2 + 64;
  if (bigThr > B2G_BIG_ISLAND) bigThr = B2G_BIG_ISLAND;
  // nothing known yet (first step, or nothing was awake): allow the largest tile rather than sending a mid-size
  // island through the grid-pass kernel, whose serial bucket is walked by ONE thread of the grid
  if (A->lastMaxIsland == 0) bigThr = B2G_BIG_ISLAND;
  // aim at >= 2 bins per SM so the fused kernel fills the chip, within what a tile can hold
  int binSize = nb / (2 * 148);
  if (binSize < 32) binSize = 32;
  if (binSize > 512) binSize = 512;
  // Resident blocks per SM.  An island's colour passes and its serial bucket are latency chains
  // (one barrier / one constraint deep), so with many mid-size islands (batched tumbler worlds:
  // one 500-body island with a 100-constraint serial chain per world) throughput comes from
  // overlapping MANY islands per SM, not from keeping one island's planes in shared memory:
  // smaller blocks (128 threads, 4 per SM) with just the body tile in shared memory and the planes
  // in L2: 1024 tumbler worlds 6.3 -> 3.5 ms/step.  Pyramid-like islands (6 colours, no serial
  // bucket) are faster with their planes in shared memory: two or one fat blocks per SM as before.
  // B2G_FUSED_CTAS=n overrides the choice (measurements).
  static int ctasOverride = -1;
  if (ctasOverride < 0) {
    const char* e = getenv("B2G_FUSED_CTAS");
    ctasOverride = e ? atoi(e) : 0;
  }
  int fusedThreads = B2G_FUSED_THREADS;
  size_t budget = 0;
  {
    int want = ctasOverride;
    // chain-heavy = at least 2 % of last step's solver rows sat in serial buckets (hub bodies)
    const bool chainHeavy = nb / binSize + 1 >= 4 * 148 && bigThr >= 128 && A->lastActive > 0 &&
                            (long long)A->lastOverflow * 50 >= A->lastActive;
    if (want == 0) want = chainHeavy ? 7 : 2;
    // Try the densest packing first: 7 blocks of 96 threads per SM hold 1036 islands at once (1024 batched
    // tumbler worlds = ONE wave instead of two), then 6 and 4 blocks of 128 threads.  What decides is the
    // tile: one island of up to bigThr bodies (+ a few small ones) must fit the block's share of shared
    // memory.  With many equal islands their size is stable, so the cap follows it tightly (an island that
    // outgrows it within one step takes the grid-pass kernel for that step).
    const int tight = A->lastMaxIsland + A->lastMaxIsland / 16 + 8;
    for (int w = want; w > 2; w = (w > 6 ? 6 : (w > 4 ? 4 : 2))) {
      const size_t per = (size_t)(227 * 1024) / w - 2048;
      const int cap = (int)(per / (FusedTile::bytes(1024) / 1024));  // bodies a block's tile can hold
      int thr = bigThr;
      if (cap - thr < 32 && w > 4 && tight >= 128 && tight < thr) thr = tight;
      const int fit = cap - thr;  // bodies of small islands that still fit next to one big one
      if (fit >= 32) {
        bigThr = thr;
        if (binSize > fit) binSize = fit;
        budget = per;
        fusedThreads = w > 6 ? 96 : 128;
        break;
      }
    }
  }
  const int nbins = nb / binSize + 1;
  const int bigBin = nbins;  // sorts after every fused bin
  // Oversize islands: cut into per-SM tiles (b2g_tiles.cuh) when last step's oversize bodies fit the tiles'
  // shared memory with room to drift; the grid-pass kernel (k_big_solve) otherwise, and on the step an
  // oversize island first appears (no plan yet).
  if (A->tileGrid == 0) {
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, A->device));
    A->tileGrid = sms * B2G_TILES_PER_SM < B2G_TILES_MAX ? sms * B2G_TILES_PER_SM : B2G_TILES_MAX;
  }
  // Only islands beyond the hard cap are tiled: an island that is "oversize" just because the adaptive
  // threshold lags behind its growth (a transient of a step or two) takes the grid-pass kernel, whose
  // visiting order does not depend on what else is in the arena (batched worlds stay bit-identical to
  // single ones; tiles are cut through space, so they mix worlds that overlap in space).
  const bool useTiles = !A->tilesDisabled && A->lastNumBig > 0 && A->lastBigBodies > 0 &&
                        A->lastMaxIsland > B2G_BIG_ISLAND &&
                        (long long)A->lastBigBodies * 4 <= (long long)A->tileGrid * B2G_TILE_CAP * 3;
  {
    static int dbg = -1;
    if (dbg < 0) dbg = getenv("B2G_DEBUG_TILES") ? 1 : 0;
    if (dbg && (A->stepCount % 50 == 0 || (A->stepCount >= 74 && A->stepCount <= 80)))
      fprintf(stderr, "[tiles] step %lld useTiles %d lastNumBig %d lastBigBodies %d lastMaxIsland %d lastSpill %d tileGrid %d\n",
              A->stepCount, (int)useTiles, A->lastNumBig, A->lastBigBodies, A->lastMaxIsland, A->lastSpill, A->tileGrid);
  }
  const int tileBin0 = useTiles ? nbins + 1 : -1;
  const int cutBin = useTiles ? nbins + 1 + A->tileGrid : -1;
  const int nbinsAll = useTiles ? cutBin + 1 : nbins + 1;  // bins whose buckets the sort covers
  if (nbinsAll > A->nbinsMax) {
    set_err("b2g_step", "internal: bin table too small");
    return B2G_ERR_CAPACITY;
  }
  const int tileCap = binSize - 1 + bigThr;
  const size_t tileBytes = FusedTile::bytes(tileCap);
  // otherwise: two blocks per SM when the tile allows it (113 KB each), or one large block
  if (budget == 0) budget = tileBytes + 32 * 1024 <= 113 * 1024 ? 113 * 1024 : 225 * 1024;
  int conCap = budget > tileBytes ? (int)((budget - tileBytes) / (B2G_PLANES * 16)) : 0;
  if (conCap < 0) conCap = 0;
  const size_t smem = tileBytes + (size_t)conCap * B2G_PLANES * 16;

  const int nbuckets = nbinsAll << B2G_COLOUR_BITS;
  LAUNCH(A, KC_ISLANDS, nb, k_body_begin, div_up(nb, 256), 256, nb, A->bflags, A->force, A->islandParent,
         A->islandAwake, A->island, A->islandDirty, A->islandsValid, A->islandWasBig,
         (A->stepCount % B2G_ISLAND_EXACT_PERIOD) == 0, A->islandMinSleep, A->islandPen, A->capBodies, P->position_iterations, A->colourMask,
         A->bodyBest, A->islandCount, A->islandCursor, A->binFirst, A->binEnd, nbinsAll, A->bucketCount, nbuckets);
  if (nc > 0)
    LAUNCH(A, KC_ISLANDS, nc, k_island_union, div_up(nc, 256), 256, nc, C, A->bflags, A->fTypeFlags, A->islandParent);
  if (nj > 0)
    LAUNCH(A, KC_ISLANDS, nj, k_island_union_joints, div_up(nj, 256), 256, nj, A->jBodies, A->bflags, A->islandParent);
  LAUNCH(A, KC_ISLANDS, nb, k_island_flatten, div_up(nb, 256), 256, nb, A->bflags, A->islandParent, A->island,
         A->islandAwake, A->islandCount, A->dCounts, A->islandDirty);
  LAUNCH(A, KC_ISLANDS, nb, k_island_alloc, div_up(nb, 256), 256, nb, A->island, A->islandAwake, A->islandCount, A->islandStart,
         A->binFirst, A->binEnd, binSize, bigThr, A->dCounts, A->islandWasBig);
  if (useTiles) {
    // re-plan periodically, when the oversize bodies have grown by a quarter since the plan, and long before
    // a tile can fill up (a full tile gives all its bodies up for the step: correct, but slow)
    const bool replan = !A->tilePlanValid || A->tilePlanAge >= B2G_TILE_PLAN_PERIOD || A->lastSpill > 0 ||
                        A->lastMaxTile * 5 > B2G_TILE_CAP * 3 ||
                        (long long)A->lastBigBodies * 4 > (long long)A->tilePlanBodies * 5;
    if (replan) {
      TilePlan init;
      memset(&init, 0, sizeof(init));
      init.lo[0] = init.lo[1] = 0xffffffffu;
      CK(cudaMemcpyAsync(A->tilePlan, &init, sizeof(init), cudaMemcpyHostToDevice, A->stream));
      CK(cudaMemsetAsync(A->tileHistX, 0, sizeof(int) * B2G_TILE_XBINS, A->stream));
      CK(cudaMemsetAsync(A->tileHistY, 0, sizeof(int) * (size_t)B2G_TILES_MAX * B2G_TILE_YBINS, A->stream));
      LAUNCH(A, KC_ISLANDS, nb, k_tile_bounds, div_up(nb, 256), 256, nb, A->pos, A->bflags, A->island, A->islandAwake,
             A->islandCount, bigThr, A->tilePlan);
      LAUNCH(A, KC_ISLANDS, 1, k_tile_plan_begin, 1, 32, A->tilePlan, A->tileGrid);
      LAUNCH(A, KC_ISLANDS, nb, k_tile_xhist, div_up(nb, 256), 256, nb, A->pos, A->bflags, A->island, A->islandAwake,
             A->islandCount, bigThr, A->tilePlan, A->tileHistX);
      LAUNCH(A, KC_ISLANDS, B2G_TILE_XBINS, k_tile_xplan, 1, 1024, A->tilePlan, A->tileHistX, A->tileStripOfX);
      LAUNCH(A, KC_ISLANDS, nb, k_tile_yhist, div_up(nb, 256), 256, nb, A->pos, A->bflags, A->island, A->islandAwake,
             A->islandCount, bigThr, A->tilePlan, A->tileStripOfX, A->tileHistY);
      LAUNCH(A, KC_ISLANDS, B2G_TILE_YBINS, k_tile_yplan, A->tileGrid, 1024, A->tilePlan, A->tileHistY, A->tileRowOfY);
      A->tilePlanValid = 1;
      A->tilePlanAge = 0;
      A->tilePlanBodies = A->lastBigBodies;
    }
    A->tilePlanAge++;
    CK(cudaMemsetAsync(A->tileCount, 0, sizeof(int) * B2G_TILES_MAX, A->stream));
    CK(cudaMemsetAsync(A->tileBoundary, 0, (size_t)B2G_TILES_MAX * B2G_TILE_CAP, A->stream));
    LAUNCH(A, KC_ISLANDS, nb, k_tile_assign, div_up(nb, 256), 256, nb, A->pos, A->bflags, A->island, A->islandAwake,
           A->islandCount, bigThr, A->tilePlan, A->tileStripOfX, A->tileRowOfY, A->tileSlot, A->tileBodies, A->tileCount,
           A->spillList, A->dCounts);
    LAUNCH(A, KC_ISLANDS, nb, k_tile_overflow_fix, div_up(nb, 256), 256, nb, A->tileSlot, A->tileCount, A->spillList, A->dCounts);
    if (nj > 0)
      LAUNCH(A, KC_ISLANDS, nj, k_tile_joint_marks, div_up(nj, 256), 256, nj, A->jBodies, A->tileSlot, A->tileBoundary,
             A->dCounts);
  } else {
    A->tilePlanValid = 0;
  }
  if (nc == 0)  // otherwise the scatter rides in k_mark_active_bins
    LAUNCH(A, KC_ISLANDS, nb, k_body_scatter, div_up(nb, 256), 256, nb, A->bflags, A->island, A->islandAwake,
           A->islandCount, A->islandStart, A->islandCursor, A->bodySlot, A->slotBody, bigThr, A->dCounts);

  int numActive = 0, numBig = 0, rounds = 0;
  // Colouring: k_mark_active_bins publishes the persisting colours and lists what is uncoloured,
  // k_colour_worklist runs every round of the step over that list in one launch, the counting pass of the
  // (bin, colour) sort rides in both.  The step's counters are copied to the host behind them and only
  // waited for after the fused kernel has been queued (the answer decides nothing before that point).
  if (nc > 0) {
    if (A->numWorlds > 1 && A->fixBaseDirty) {
      CK(cudaMemsetAsync(A->worldFixMin, 0x7f, sizeof(int) * (size_t)(A->numWorlds + 1), A->stream));
      LAUNCH(A, KC_COLOUR, A->nFixtures, k_world_fix_min, div_up(A->nFixtures, 256), 256, A->nFixtures, A->fBody, A->bworld,
             A->worldFixMin);
      LAUNCH(A, KC_COLOUR, nb, k_body_fix_base, div_up(nb, 256), 256, nb, A->bworld, A->worldFixMin, A->bodyFixBase);
      A->fixBaseDirty = 0;
    }
    const int* fixBase = A->numWorlds > 1 ? A->bodyFixBase : nullptr;
    const int nmax = nb > nc ? nb : nc;
    MarkArgs M;
    M.nc = nc;
    M.nb = nb;
    M.dropColours = A->recolour;
    M.binSize = binSize;
    M.bigThreshold = bigThr;
    M.bigBin = bigBin;
    M.tileBin0 = tileBin0;
    M.cutBin = cutBin;
    M.tileCap = B2G_TILE_CAP;
    LAUNCH(A, KC_COLOUR, nc, k_mark_active_bins, div_up(nmax, 256), 256, M, C, A->fTypeFlags, A->bflags, A->island,
           A->islandAwake, A->islandCount, A->islandStart, A->cbin, A->dCounts, A->mass, A->colourMask, A->islandCursor,
           A->bodySlot, A->slotBody, A->bodyBest, fixBase, A->activeList, A->bucketCount, A->conVals,
           (const int*)A->tileSlot, A->tileBoundary, A->recolourWorlds ? (const uint8_t*)A->worldRecolour : nullptr,
           (const int*)A->bworld);
    A->recolour = 0;
    if (A->recolourWorlds) {
      CK(cudaMemsetAsync(A->worldRecolour, 0, (size_t)A->numWorlds, A->stream));
      A->recolourWorlds = 0;
    }
    {
      if (A->colourGrid == 0) {
        int perSM = 0, sms = 0;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, k_colour_worklist, B2G_WL_THREADS, 0));
        CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, A->device));
        if (perSM < 1) return B2G_ERR_CUDA;
        A->colourGrid = sms;
        const char* e = getenv("B2G_WL_GRID");  // blocks of the worklist colouring (measurements)
        if (e && atoi(e) > 0 && atoi(e) <= sms * perSM) A->colourGrid = atoi(e);
      }
      int cutBin = M.cutBin, bb = bigBin;
      static int singleMax = -1;  // B2G_WL_SINGLE_MAX=n: worklists up to n entries are coloured by one CTA (measurements)
      if (singleMax < 0) {
        const char* e = getenv("B2G_WL_SINGLE_MAX");
        singleMax = e ? atoi(e) : B2G_WL_SINGLE_MAX;
      }
      static int cutSectors = 0;
      if (cutSectors == 0) {
        const char* e = getenv("B2G_CUT_SECTORS");
        cutSectors = e ? atoi(e) : B2G_CUT_SECTORS;
        if (cutSectors < 1) cutSectors = 1;
      }
      void* args[] = {&C, &A->cbin, &A->mass, &A->colourMask, &A->bodyBest, &fixBase, &A->activeList, &A->dCounts,
                      &bb, &cutBin, &A->bucketCount, &A->conVals, &A->colourBarrier, &singleMax, &A->pos, &cutSectors};
      CK(cudaMemsetAsync(A->colourBarrier, 0, sizeof(unsigned int), A->stream));
      ktime_begin(A, KC_COLOUR, nc);
      // cooperative launch for the co-residency of its grid barrier (the long-worklist mode)
      CK(cudaLaunchCooperativeKernel((void*)k_colour_worklist, dim3(A->colourGrid), dim3(B2G_WL_THREADS), args, 0, A->stream));
      ktime_end(A);
      A->launches++;
    }
    CK(cudaMemcpyAsync(A->hCounts, A->dCounts, sizeof(StepCounts), cudaMemcpyDeviceToHost, A->stream));
    CK(cudaEventRecord(A->ev[4], A->stream));
    LAUNCH(A, KC_COLOUR, nbuckets, k_bucket_scan, 1, 1024, nbuckets, A->bucketCount, A->bucketStart);
    LAUNCH(A, KC_COLOUR, nc, k_bucket_scatter, div_up(nc, 256), 256, nc, A->cbin, C, A->bucketStart, A->conVals,
           A->sortedList);
  }

  // ---- every island that fits a tile: one launch -----------------------------------------------
  int fusedKt = -1;
  {
    FusedParams FP;
    FP.nc = nc;
    FP.binSize = binSize;
    FP.h = h;
    FP.dtRatio = dtRatio;
    FP.gravity = make_float2(P->gravity_x, P->gravity_y);
    FP.velIters = P->velocity_iterations;
    FP.posIters = P->position_iterations;
    FP.warmStarting = P->warm_starting;
    FP.allowSleep = P->allow_sleep;
    FP.clearForces = P->clear_forces;
    FP.tileCap = tileCap;
    FP.conCap = conCap;
    FP.nj = nj;
    FP.invH = h > 0.0f ? 1.0f / h : 0.0f;
    FP.jColour = A->jColour;
    FP.jSorted = A->jSorted;
    FP.jCstart = A->jCstart;
    if (smem > A->fusedSmemSet) {
      CK(cudaFuncSetAttribute(k_solve_bins_fused, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      A->fusedSmemSet = smem;
    }
    fusedKt = A->kernelTiming ? A->ktCount : -1;  // its unit count is only known after the readback below
    ktime_begin(A, KC_FUSED_SOLVE, 0.0);
    launch_pdl(A->stream, dim3(nbins), dim3(fusedThreads), smem, k_solve_bins_fused,
        FP, A->binFirst, A->binEnd, A->slotBody, A->bodySlot, A->island, A->islandStart, A->bucketStart,
        A->sortedList, (int*)A->conKeys, C, A->fRadius, S, A->bflags, A->pos, A->vel, A->xf, A->force, A->mass, A->center, A->dCounts,
        joint_views(A), A->conVals, (int*)A->conKeysSorted);
    ktime_end(A);
    A->launches++;
  }

  // ---- oversize islands, tiled: one persistent launch, queued before the host looks at any counter ---------
  int tilesKt = -1;
  if (useTiles) {
    TileArgs T;
    T.tileBin0 = tileBin0;
    T.cutBin = cutBin;
    T.nb = nb;
    T.nj = nj;
    T.h = h;
    T.invH = h > 0.0f ? 1.0f / h : 0.0f;
    T.dtRatio = dtRatio;
    T.gravity = make_float2(P->gravity_x, P->gravity_y);
    T.velIters = P->velocity_iterations;
    T.posIters = P->position_iterations;
    T.warmStarting = P->warm_starting;
    T.allowSleep = P->allow_sleep;
    T.clearForces = P->clear_forces;
    T.noSequencing = A->tileNoSequencing;
    T.cutSeq = A->tileCutSeq;
    T.colourMask = A->colourMask;
    T.plan = A->tilePlan;
    T.tileCount = A->tileCount;
    T.tileBodies = A->tileBodies;
    T.tileBoundary = A->tileBoundary;
    T.tileSlot = A->tileSlot;
    T.spillList = A->spillList;
    T.bucketStart = A->bucketStart;
    T.sortedList = A->sortedList;
    T.orderScratch = (int*)A->conKeys;
    T.croot = A->croot;
    T.islandPen = A->islandPen;
    T.penStride = A->capBodies;
    T.islandMinSleep = A->islandMinSleep;
    T.bflags = A->bflags;
    T.pos = A->pos;
    T.vel = A->vel;
    T.xf = A->xf;
    T.force = A->force;
    T.mass = A->mass;
    T.center = A->center;
    T.fRadius = A->fRadius;
    T.island = A->island;
    T.islandAwake = A->islandAwake;
    T.bodySlot = A->bodySlot;
    T.counts = A->dCounts;
    T.barrier = A->tileBarrier;
    const size_t tileSmemBytes = (size_t)B2G_TILE_CAP * B2G_TILE_BODY_BYTES + (size_t)B2G_BIG_STAGE_PLANES * B2G_TILE_THREADS * sizeof(float4);
    static bool attrSet = false;
    if (!attrSet) {
      CK(cudaFuncSetAttribute(k_big_tiles, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tileSmemBytes));
      int perSM = 0;
      CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, k_big_tiles, B2G_TILE_THREADS, tileSmemBytes));
      if (perSM < B2G_TILES_PER_SM) {
        set_err("b2g_step", "k_big_tiles: the tiles of an SM do not fit it together");
        return B2G_ERR_CUDA;
      }
      attrSet = true;
    }
    JointWalk JW = joint_walk(A, 1);
    JointArraysDev JV = joint_views(A);
    void* args[] = {&T, &S, &C, &JW, &JV};
    CK(cudaMemsetAsync(A->tileBarrier, 0, sizeof(unsigned int), A->stream));
    tilesKt = A->kernelTiming ? A->ktCount : -1;
    // units = constraints of the tiled islands (known after the readback below); the launch covers all of
    // b2Island::Solve for them (bench.py charges SURVEY 8d's 2480 B per constraint at 8/3 iterations)
    ktime_begin(A, KC_BIG_TILES, 0.0);
    CK(cudaLaunchCooperativeKernel((void*)k_big_tiles, dim3(A->tileGrid), dim3(B2G_TILE_THREADS), args, tileSmemBytes,
                                   A->stream));
    ktime_end(A);
    A->launches++;
  }

  if (nc > 0) {
    CK(cudaEventSynchronize(A->ev[4]));  // the copy finished long ago; the fused kernel is still running
    rounds = A->hCounts->lastUsefulRound;
    numActive = A->hCounts->numActive;
    numBig = A->hCounts->numBig;
    if (fusedKt >= 0 && fusedKt < B2G_KT_MAX) A->ktUnits[fusedKt] = (double)numActive - numBig;
    if (tilesKt >= 0 && tilesKt < B2G_KT_MAX) A->ktUnits[tilesKt] = (double)numBig;
    out.numColours = A->hCounts->numColours;
    out.numOverflow = A->hCounts->numOverflow + A->hCounts->remaining;
    A->lastOverflow = out.numOverflow;
    A->lastActive = numActive;
  } else if (nj > 0) {
    // no contacts at all, but joints can still form an oversize island (a chain in free fall)
    int rc = read_counts(A);
    if (rc) return rc;
  }
  // bodies of oversize islands are advanced by the big path even when those islands hold no
  // contact constraint (joints only)
  const int bigBodies = (nc > 0 || nj > 0) ? A->hCounts->numBigBodies : 0;
  A->lastNumBig = numBig > 0 ? numBig : bigBodies;
  A->lastMaxIsland = (nc > 0 || nj > 0) ? A->hCounts->maxIslandBodies : 0;
  A->lastBigBodies = bigBodies;
  A->lastSpill = (nc > 0 || nj > 0) ? A->hCounts->spillCount : 0;
  A->lastMaxTile = (nc > 0 || nj > 0) ? A->hCounts->maxTileCount : 0;

  // ---- oversize islands, not tiled (first appearance, or too many bodies for the tiles): grid-wide passes ----
  if (!useTiles && (numBig > 0 || bigBodies > 0)) {
    const int bigStart = numActive - numBig;
    int colourFirst[B2G_MAX_COLOURS + 2];
    int acc = bigStart, numColours = 0;
    for (int c = 0; c <= B2G_MAX_COLOURS; ++c) {
      colourFirst[c] = acc;
      acc += A->hCounts->colourCount[c];
      if (c < B2G_MAX_COLOURS && A->hCounts->colourCount[c] > 0) numColours = c + 1;
    }
    colourFirst[B2G_MAX_COLOURS + 1] = acc;
    const int numOverflow = A->hCounts->colourCount[B2G_MAX_COLOURS];
#ifdef B2G_BIG_TRACE
    if (A->stepCount % 50 == 0) {
      fprintf(stderr, "[big] step %lld colours:", A->stepCount);
      for (int c = 0; c <= B2G_MAX_COLOURS; ++c) fprintf(stderr, " %d", A->hCounts->colourCount[c]);
      fprintf(stderr, "\n");
    }
#endif
    if (numOverflow > 1)
      LAUNCH(A, KC_COLOUR, numOverflow, k_order_overflow, 1, 256, colourFirst[B2G_MAX_COLOURS],
             colourFirst[B2G_MAX_COLOURS + 1], A->sortedList, (int*)A->conKeys, C);
    LAUNCH(A, KC_INTEGRATE, nb, k_integrate_velocities, div_up(nb, 256), 256, nb, A->bflags, A->island,
           A->islandAwake, A->vel, A->mass, A->center, A->force, h, make_float2(P->gravity_x, P->gravity_y),
           A->dCounts, A->bodySlot, 1);
    if (numBig > 0)
      LAUNCH(A, KC_PREPARE, numBig, k_prepare, div_up(numBig, 128), 128, bigStart, numBig, A->sortedList, C, A->fRadius,
             A->bflags, A->island, S, A->croot, A->pos, A->vel, A->mass, A->center, dtRatio, P->warm_starting);
    {
      // one persistent cooperative launch: all colours x all iterations, grid barrier in between
      BigRanges R;
      for (int c = 0; c <= B2G_MAX_COLOURS + 1; ++c) R.first[c] = colourFirst[c];
      R.numColours = numColours;
      const size_t stageBytes = (size_t)B2G_BIG_STAGE_PLANES * B2G_BIG_THREADS * sizeof(float4);
      if (A->bigGrid == 0) {
        int perSM = 0, sms = 0;
        CK(cudaFuncSetAttribute(k_big_solve, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)stageBytes));
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, k_big_solve, B2G_BIG_THREADS, stageBytes));
        CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, A->device));
        if (perSM < 1) return B2G_ERR_CUDA;
        A->bigGrid = sms;  // one block per SM: the grid barrier costs grow with the block count
      }
      int velIters = P->velocity_iterations, posIters = P->position_iterations, warm = P->warm_starting;
      int penStride = A->capBodies, nbodies = nb;
      float hh = h, dtr = dtRatio;
      JointWalk JW = joint_walk(A, 1);
      JointArraysDev JV = joint_views(A);
      void* args[] = {&R, &S, &C, &A->vel, &A->pos, &A->croot, &A->islandPen, &penStride, &nbodies, &A->bflags,
                      &A->island, &A->islandAwake, &A->bodySlot, &hh, &velIters, &posIters, &warm, &JW, &JV,
                      &A->mass, &A->center, &dtr, &A->bigBarrier};
      CK(cudaMemsetAsync(A->bigBarrier, 0, sizeof(unsigned int), A->stream));
      // units = constraints; the launch is warm start + velIters velocity + posIters position sweeps + the
      // impulse store (bench.py charges 180 + 8 x 196 + 3 x 136 + 32 B per constraint at 8/3 iterations)
      ktime_begin(A, KC_BIG_SOLVE, (double)numBig);
      // cooperative launch for the co-residency guarantee; the barrier itself is grid_arrive / grid_wait
      CK(cudaLaunchCooperativeKernel((void*)k_big_solve, dim3(A->bigGrid), dim3(B2G_BIG_THREADS), args, stageBytes,
                                     A->stream));
      ktime_end(A);
      A->launches++;
    }
    LAUNCH(A, KC_FINALIZE, nb, k_finalize_bodies, div_up(nb, 256), 256, nb, A->bflags, A->island, A->islandAwake,
           A->pos, A->vel, A->center, A->xf, A->force, A->islandMinSleep, h, P->allow_sleep, A->bodySlot, 1);
    LAUNCH(A, KC_FINALIZE, nb, k_sleep_and_clear, div_up(nb, 256), 256, nb, A->bflags, A->island, A->islandAwake,
           A->islandMinSleep, A->islandPen, A->capBodies, P->position_iterations, A->vel, A->force, P->allow_sleep,
           P->clear_forces, A->dCounts, A->bodySlot, 1);
  }
  out.numActive = numActive;
  out.rounds = rounds;
  return B2G_OK;
}

__global__ void k_pack_body_state(int first, int count, const float4* __restrict__ xf, const float4* __restrict__ vel,
                                  float4* out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  out[2 * i] = xf[first + i];
  out[2 * i + 1] = vel[first + i];
}

static int step_check(b2gArena* A, const b2gStepParams* P) {
  if (!A || !P) return B2G_ERR_INVALID;
  if (P->position_iterations > B2G_MAX_POS_ITERS || P->position_iterations < 0 || P->velocity_iterations < 0) {
    set_err("b2g_step", "iteration counts out of range");
    return B2G_ERR_INVALID;
  }
  return B2G_OK;
}

// first half of Step: b2ContactManager::Collide (b2_world.cpp:1134-1138)
extern "C" int b2g_step_collide(b2gArena* A, const b2gStepParams* P) {
  int rc0 = step_check(A, P);
  if (rc0) return rc0;
  CK(cudaSetDevice(A->device));
  A->launchesAtStepStart = A->launches;
  if (A->newFixtures) {
    // b2World::Step starts with FindNewContacts when fixtures were added (m_newContacts,
    // b2_world.cpp:1114-1122); this also fills the fixture radii the solver reads
    CK(cudaMemsetAsync(A->dCounts, 0, sizeof(StepCounts), A->stream));
    int rcn = find_new_contacts(A, 0);
    if (rcn) return rcn;
  }
  const int nc = A->nContacts;
  ContactBuf& C = A->cb[0];
  CK(cudaMemsetAsync(A->dCounts, 0, sizeof(StepCounts), A->stream));
  if (A->profiling) CK(cudaEventRecord(A->ev[0], A->stream));
  if (nc > 0) {
    LAUNCH(A, KC_NARROWPHASE, nc, k_narrowphase, div_up(nc, B2G_NP_THREADS), B2G_NP_THREADS, nc, C, A->bflags, A->xf, A->fShapeOff, A->fTypeFlags, A->shapes,
           A->bflags, A->dCounts, P->record_events, A->beginEvents, A->endEvents, A->capContacts, A->island,
           A->islandDirty);
  }
  if (A->profiling) CK(cudaEventRecord(A->ev[1], A->stream));
  return B2G_OK;
}

// second half of Step: Solve + FindNewContacts + ClearForces (b2_world.cpp:1140-1167)
extern "C" int b2g_step_solve(b2gArena* A, const b2gStepParams* P, b2gStepStats* stats) {
  int rc0 = step_check(A, P);
  if (rc0) return rc0;
  CK(cudaSetDevice(A->device));
  const long long launches0 = A->launchesAtStepStart;
  const int nb = A->nBodies;
  const float h = P->dt;
  A->stepDt = h;
  const float inv_dt = h > 0.0f ? 1.0f / h : 0.0f;
  const int prof = A->profiling;
  int numActive = 0, numColours = 0, numOverflow = 0, rounds = 0;

  if (h > 0.0f && nb > 0) {
    SolveOut so;
    int rcs = P->solver_mode == B2G_SOLVER_COLOURED ? solve_fused(A, P, so) : solve_legacy(A, P, so);
    if (rcs) return rcs;
    A->islandsValid = 1;
    A->stepCount++;
    numActive = so.numActive;
    numColours = so.numColours;
    numOverflow = so.numOverflow;
    rounds = so.rounds;
    if (prof) CK(cudaEventRecord(A->ev[2], A->stream));

    // body state is final here (the pair refresh only reads transforms): a requested readback is
    // packed now and copied on a second stream while the broadphase runs
    if (A->pendingStateDst && A->pendingStateCount > 0) {
      k_pack_body_state<<<div_up(A->pendingStateCount, 256), 256, 0, A->stream>>>(
          A->pendingStateFirst, A->pendingStateCount, A->xf, A->vel, A->stateStage);
      CK(cudaGetLastError());
      A->launches++;
      CK(cudaEventRecord(A->evPacked, A->stream));
      CK(cudaStreamWaitEvent(A->copyStream, A->evPacked, 0));
      CK(cudaMemcpyAsync(A->pendingStateDst, A->stateStage, (size_t)A->pendingStateCount * 32, cudaMemcpyDeviceToHost,
                         A->copyStream));
    }

    // ---- FindNewContacts (end of Solve, b2_world.cpp:663-669) ------------------------
    int rc = find_new_contacts(A, P->record_events);
    A->invDt0 = inv_dt;  // the solve is complete even if the pair list could not grow
    if (rc) return rc;
    if (prof) CK(cudaEventRecord(A->ev[3], A->stream));
  } else {
    if (prof) {
      CK(cudaEventRecord(A->ev[2], A->stream));
      CK(cudaEventRecord(A->ev[3], A->stream));
    }
    int rc = read_counts(A);
    if (rc) return rc;
  }

  if (A->kernelTiming) {
    CK(cudaStreamSynchronize(A->stream));
    ktime_collect(A);
  }
  if (stats) {
    memset(stats, 0, sizeof(*stats));
    // hCounts was last read inside find_new_contacts, after every counter of this step was final
    stats->num_bodies = nb;
    stats->num_fixtures = A->nFixtures;
    stats->num_contacts = A->nAlive;
    stats->num_touching = A->hCounts->numTouching;
    stats->num_constraints = numActive;
    stats->num_colours = numColours;
    stats->num_overflow = numOverflow;
    stats->num_awake = A->hCounts->numAwake;
    stats->num_pairs = A->hCounts->numPairs;
    stats->bp_max_visits = A->hCounts->bpMaxVisits;
    stats->bp_mean_visits = A->nFixtures > 0 ? (float)((double)A->hCounts->bpVisits / A->nFixtures) : 0.0f;
    stats->bp_rebuilt = A->bvhAge == 0 ? 1 : 0;
    stats->colour_rounds = rounds;
    stats->num_launches = (int)(A->launches - launches0);
    if (prof) {
      CK(cudaStreamSynchronize(A->stream));
      cudaEventElapsedTime(&stats->ms_collide, A->ev[0], A->ev[1]);
      cudaEventElapsedTime(&stats->ms_solve, A->ev[1], A->ev[2]);
      cudaEventElapsedTime(&stats->ms_broadphase, A->ev[2], A->ev[3]);
      cudaEventElapsedTime(&stats->ms_step, A->ev[0], A->ev[3]);
    }
  }
  return B2G_OK;
}

extern "C" int b2g_step(b2gArena* A, const b2gStepParams* P, b2gStepStats* stats) {
  int rc = b2g_step_collide(A, P);
  if (rc) return rc;
  return b2g_step_solve(A, P, stats);
}

extern "C" int b2g_step_download(b2gArena* A, const b2gStepParams* P, b2gStepStats* stats, int32_t first, int32_t count,
                                 float* dst) {
  if (!A || !dst || first < 0 || count < 0 || first + count > A->nBodies) return B2G_ERR_INVALID;
  int rc = b2g_step_collide(A, P);
  if (rc) return rc;
  A->pendingStateDst = dst;
  A->pendingStateFirst = first;
  A->pendingStateCount = count;
  rc = b2g_step_solve(A, P, stats);
  A->pendingStateDst = nullptr;
  cudaError_t e = cudaStreamSynchronize(A->copyStream);
  if (rc) return rc;
  CK(e);
  return B2G_OK;
}

// ---------------------------------------------------------------------------------------------
// readback
// ---------------------------------------------------------------------------------------------
#define DOWN(dst, src, elems, type)                                                                          \
  if (dst) CK(cudaMemcpyAsync((dst), (src) + (size_t)first * (elems), (size_t)count * (elems) * sizeof(type), \
                              cudaMemcpyDeviceToHost, A->stream))

extern "C" int b2g_download_bodies(b2gArena* A, int32_t first, int32_t count, const b2gBodyArrays* d) {
  if (!A || !d || first < 0 || count < 0 || first + count > A->nBodies) return B2G_ERR_INVALID;
  CK(cudaSetDevice(A->device));
  DOWN(d->pos, (float*)A->pos, 4, float);
  DOWN(d->vel, (float*)A->vel, 4, float);
  DOWN(d->xf, (float*)A->xf, 4, float);
  DOWN(d->mass, (float*)A->mass, 4, float);
  DOWN(d->center, (float*)A->center, 4, float);
  DOWN(d->force, (float*)A->force, 4, float);
  DOWN(d->flags, A->bflags, 1, uint32_t);
  DOWN(d->world, A->bworld, 1, int32_t);
  CK(cudaStreamSynchronize(A->stream));
  return B2G_OK;
}


extern "C" int b2g_download_body_state_async(b2gArena* A, int32_t first, int32_t count, float* dst) {
  if (!A || !dst || first < 0 || count < 0 || first + count > A->nBodies) return B2G_ERR_INVALID;
  CK(cudaSetDevice(A->device));
  if (count == 0) return B2G_OK;
  // interleave on the device, then ONE linear copy: strided 2-D copies of 16-byte rows cost a DMA
  // descriptor per row and were slower than the step itself at 100k bodies
  k_pack_body_state<<<div_up(count, 256), 256, 0, A->stream>>>(first, count, A->xf, A->vel, A->stateStage);
  CK(cudaGetLastError());
  A->launches++;
  CK(cudaMemcpyAsync(dst, A->stateStage, (size_t)count * 32, cudaMemcpyDeviceToHost, A->stream));
  return B2G_OK;
}

extern "C" int b2g_download_fixture_aabbs(b2gArena* A, int32_t first, int32_t count, float* aabb) {
  if (!A || !aabb || first < 0 || count < 0 || first + count > A->nFixtures) return B2G_ERR_INVALID;
  CK(cudaSetDevice(A->device));
  CK(cudaMemcpyAsync(aabb, (float*)A->fAabb + (size_t)first * 4, (size_t)count * 16, cudaMemcpyDeviceToHost,
                     A->stream));
  CK(cudaStreamSynchronize(A->stream));
  return B2G_OK;
}

extern "C" int b2g_contact_count(b2gArena* A, int32_t* out) {
  if (!A || !out) return B2G_ERR_INVALID;
  *out = A->nAlive;
  return B2G_OK;
}

// gathers the live contacts (slots are sparse and unordered) into dense arrays
__global__ void k_pack_alive(int nSlots, ContactBuf C, int* counter, int* slots, unsigned long long* keys,
                             float4* man, int* fa, int* fb, uint32_t* flags, float4* material, int* colour) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= nSlots) return;
  uint32_t f = C.flags[j];
  if (!(f & B2G_CONTACT_ALIVE)) return;
  int i = atomicAdd(counter, 1);
  slots[i] = j;
  keys[i] = C.key[j];
  man[4 * i] = C.m0[j];
  man[4 * i + 1] = C.m1[j];
  man[4 * i + 2] = C.m2[j];
  man[4 * i + 3] = C.m3[j];
  int2 fx = C.fix[j];
  fa[i] = fx.x;
  fb[i] = fx.y;
  flags[i] = f & ~B2G_CONTACT_ALIVE;
  material[i] = C.material[j];
  colour[i] = C.colour[j];
}

struct PackedContacts {
  std::vector<int> slots, fa, fb, colour;
  std::vector<unsigned long long> keys;
  std::vector<uint32_t> flags;
  std::vector<float> man, material;
  std::vector<int> order;  // permutation that sorts by key: the order the C-ABI presents
};

// live contacts in ascending pair-key order (deterministic for the caller, whatever the slots)
static int pack_contacts(b2gArena* A, PackedContacts& out) {
  const int nSlots = A->nContacts, nAlive = A->nAlive;
  out = PackedContacts();
  if (nAlive <= 0) return B2G_OK;
  ContactBuf& C = A->cb[0];
  DevBuf dCnt, dSlots, dKeys, dMan, dFa, dFb, dFlags, dMat, dCol;
  CK(dCnt.alloc(4));
  CK(cudaMemsetAsync(dCnt.p, 0, 4, A->stream));
  CK(dSlots.alloc((size_t)nAlive * 4));
  CK(dKeys.alloc((size_t)nAlive * 8));
  CK(dMan.alloc((size_t)nAlive * 64));
  CK(dFa.alloc((size_t)nAlive * 4));
  CK(dFb.alloc((size_t)nAlive * 4));
  CK(dFlags.alloc((size_t)nAlive * 4));
  CK(dMat.alloc((size_t)nAlive * 16));
  CK(dCol.alloc((size_t)nAlive * 4));
  k_pack_alive<<<div_up(nSlots, 256), 256, 0, A->stream>>>(nSlots, C, dCnt.as<int>(), dSlots.as<int>(),
                                                           dKeys.as<unsigned long long>(), dMan.as<float4>(),
                                                           dFa.as<int>(), dFb.as<int>(), dFlags.as<uint32_t>(),
                                                           dMat.as<float4>(), dCol.as<int>());
  out.slots.resize(nAlive);
  out.keys.resize(nAlive);
  out.man.resize((size_t)nAlive * 16);
  out.fa.resize(nAlive);
  out.fb.resize(nAlive);
  out.flags.resize(nAlive);
  out.material.resize((size_t)nAlive * 4);
  out.colour.resize(nAlive);
  int packed = 0;
  CK(cudaMemcpyAsync(&packed, dCnt.p, 4, cudaMemcpyDeviceToHost, A->stream));
  CK(cudaMemcpyAsync(out.slots.data(), dSlots.p, (size_t)nAlive * 4, cudaMemcpyDeviceToHost, A->stream));
  CK(cudaMemcpyAsync(out.keys.data(), dKeys.p, (size_t)nAlive * 8, cudaMemcpyDeviceToHost, A->stream));
  CK(cudaMemcpyAsync(out.man.data(), dMan.p, (size_t)nAlive * 64, cudaMemcpyDeviceToHost, A->stream));
  CK(cudaMemcpyAsync(out.fa.data(), dFa.p, (size_t)nAlive * 4, cudaMemcpyDeviceToHost, A->stream));
  CK(cudaMemcpyAsync(out.fb.data(), dFb.p, (size_t)nAlive * 4, cudaMemcpyDeviceToHost, A->stream));
  CK(cudaMemcpyAsync(out.flags.data(), dFlags.p, (size_t)nAlive * 4, cudaMemcpyDeviceToHost, A->stream));
  CK(cudaMemcpyAsync(out.material.data(), dMat.p, (size_t)nAlive * 16, cudaMemcpyDeviceToHost, A->stream));
  CK(cudaMemcpyAsync(out.colour.data(), dCol.p, (size_t)nAlive * 4, cudaMemcpyDeviceToHost, A->stream));
  CK(cudaStreamSynchronize(A->stream));
  if (packed != nAlive) {
    set_err("b2g_download_contacts", "internal: live-contact count mismatch");
    return B2G_ERR_CUDA;
  }
  out.order.resize(nAlive);
  for (int i = 0; i < nAlive; ++i) out.order[i] = i;
  std::sort(out.order.begin(), out.order.end(),
            [&](int x, int y) { return out.keys[x] < out.keys[y]; });
  return B2G_OK;
}

extern "C" int b2g_download_contacts(b2gArena* A, int32_t first, int32_t count, const b2gContactArrays* d) {
  if (!A || !d || first < 0 || count < 0 || first + count > A->nAlive) return B2G_ERR_INVALID;
  if (count == 0) return B2G_OK;
  CK(cudaSetDevice(A->device));
  PackedContacts pc;
  int rc = pack_contacts(A, pc);
  if (rc) return rc;
  // remember which slot each presented index refers to (b2g_upload_contact_overrides)
  free(A->downloadSlots);
  A->downloadSlots = (int*)malloc(sizeof(int) * pc.order.size());
  A->downloadCount = (int)pc.order.size();
  for (size_t i = 0; i < pc.order.size(); ++i) A->downloadSlots[i] = pc.slots[pc.order[i]];
  for (int i = 0; i < count; ++i) {
    int k = pc.order[first + i];
    if (d->fixture_a) d->fixture_a[i] = pc.fa[k];
    if (d->fixture_b) d->fixture_b[i] = pc.fb[k];
    if (d->flags) d->flags[i] = pc.flags[k];
    if (d->manifold) memcpy(d->manifold + (size_t)i * 16, &pc.man[(size_t)k * 16], 64);
    if (d->material) memcpy(d->material + (size_t)i * 4, &pc.material[(size_t)k * 4], 16);
    if (d->colour) d->colour[i] = pc.colour[k];
  }
  return B2G_OK;
}

__global__ void k_scatter_overrides(int n, const int* __restrict__ slots, const uint32_t* __restrict__ flags,
                                    const float4* __restrict__ material, ContactBuf C) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int j = slots[i];
  if (flags) C.flags[j] = (C.flags[j] & B2G_CONTACT_ALIVE) | (flags[i] & ~B2G_CONTACT_ALIVE);
  if (material) C.material[j] = material[i];
}

// indices refer to the order of the LAST b2g_download_contacts
extern "C" int b2g_upload_contact_overrides(b2gArena* A, int32_t first, int32_t count, const uint32_t* flags,
                                            const float* material) {
  if (!A || first < 0 || count < 0 || !A->downloadSlots || first + count > A->downloadCount) return B2G_ERR_INVALID;
  if (count == 0) return B2G_OK;
  CK(cudaSetDevice(A->device));
  DevBuf dS, dF, dM;
  CK(dS.upload(A->downloadSlots + first, (size_t)count * 4));
  if (flags) CK(dF.upload(flags, (size_t)count * 4));
  if (material) CK(dM.upload(material, (size_t)count * 16));
  A->islandsValid = 0;  // a disabled contact removes an island edge
  k_scatter_overrides<<<div_up(count, 256), 256, 0, A->stream>>>(count, dS.as<int>(), flags ? dF.as<uint32_t>() : nullptr,
                                                                material ? dM.as<float4>() : nullptr, A->cb[0]);
  CK(cudaStreamSynchronize(A->stream));
  return B2G_OK;
}

__global__ void k_contacts_load(int n, const int* __restrict__ fa, const int* __restrict__ fb,
                                const uint32_t* __restrict__ flags, const float4* __restrict__ man,
                                const float4* __restrict__ material, const int* __restrict__ fBody, ContactBuf C,
                                uint8_t* persist, ContactHash H) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int a = fa[i], b = fb[i];
  unsigned long long key = ((unsigned long long)min(a, b) << 32) | (unsigned long long)max(a, b);
  C.key[i] = key;
  C.fix[i] = make_int2(a, b);
  C.body[i] = make_int2(fBody[a], fBody[b]);
  C.flags[i] = (flags[i] & (B2G_CONTACT_TOUCHING | B2G_CONTACT_ENABLED)) | B2G_CONTACT_ALIVE;
  C.material[i] = material[i];
  C.m0[i] = man[4 * i];
  C.m1[i] = man[4 * i + 1];
  C.m2[i] = man[4 * i + 2];
  C.m3[i] = man[4 * i + 3];
  C.colour[i] = -1;
  persist[i] = 0;
  hash_insert(H, key, i);
}

extern "C" int b2g_upload_contacts(b2gArena* A, int32_t count, const b2gContactArrays* s) {
  if (!A || count < 0 || (count > 0 && (!s || !s->fixture_a || !s->fixture_b || !s->flags || !s->manifold || !s->material)))
    return B2G_ERR_INVALID;
  if (count > A->capContacts) {
    set_err("b2g_upload_contacts", "max_contacts exceeded");
    return B2G_ERR_CAPACITY;
  }
  CK(cudaSetDevice(A->device));
  CK(cudaMemsetAsync(A->hash.keys, 0xff, (size_t)(A->hash.mask + 1) * 8, A->stream));
  CK(cudaMemsetAsync(A->hash.vals, 0xff, (size_t)(A->hash.mask + 1) * 4, A->stream));
  CK(cudaMemsetAsync(A->cb[0].flags, 0, (size_t)A->capContacts * 4, A->stream));
  CK(cudaMemsetAsync(A->dFreeTop, 0, 4, A->stream));
  if (count > 0) {
    DevBuf dA, dB, dF, dM, dMat;
    CK(dA.upload(s->fixture_a, (size_t)count * 4));
    CK(dB.upload(s->fixture_b, (size_t)count * 4));
    CK(dF.upload(s->flags, (size_t)count * 4));
    CK(dM.upload(s->manifold, (size_t)count * 64));
    CK(dMat.upload(s->material, (size_t)count * 16));
    k_contacts_load<<<div_up(count, 256), 256, 0, A->stream>>>(count, dA.as<int>(), dB.as<int>(), dF.as<uint32_t>(),
                                                               dM.as<float4>(), dMat.as<float4>(), A->fBody, A->cb[0],
                                                               A->persist, A->hash);
    CK(cudaStreamSynchronize(A->stream));
  }
  CK(cudaStreamSynchronize(A->stream));
  A->nContacts = count;
  A->nAlive = count;
  A->tombstones = 0;
  A->islandsValid = 0;
  A->recolour = 1;
  return B2G_OK;
}

__global__ void k_order_default(int nSlots, ContactBuf C, unsigned long long* orderKey) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < nSlots) orderKey[j] = (1ull << 62) | (C.key[j] & ((1ull << 62) - 1ull));
}
__global__ void k_order_assign(int n, const int* __restrict__ fa, const int* __restrict__ fb, ContactHash H,
                               unsigned long long* orderKey) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int a = fa[i], b = fb[i];
  unsigned long long key = ((unsigned long long)min(a, b) << 32) | (unsigned long long)max(a, b);
  int slot = hash_find(H, key);
  if (slot >= 0) orderKey[slot] = (unsigned long long)i;
}

extern "C" int b2g_set_sequential_order(b2gArena* A, int32_t count, const int32_t* fa, const int32_t* fb) {
  if (!A || count < 0 || (count > 0 && (!fa || !fb))) return B2G_ERR_INVALID;
  CK(cudaSetDevice(A->device));
  const int nSlots = A->nContacts;
  if (nSlots > 0)
    k_order_default<<<div_up(nSlots, 256), 256, 0, A->stream>>>(nSlots, A->cb[0], A->orderKey);
  if (count > 0) {
    DevBuf dA, dB;
    CK(dA.upload(fa, (size_t)count * 4));
    CK(dB.upload(fb, (size_t)count * 4));
    k_order_assign<<<div_up(count, 256), 256, 0, A->stream>>>(count, dA.as<int>(), dB.as<int>(), A->hash, A->orderKey);
    CK(cudaStreamSynchronize(A->stream));
  }
  CK(cudaStreamSynchronize(A->stream));
  A->seqOrderActive = 1;
  return B2G_OK;
}

// ---------------------------------------------------------------------------------------------
// spatial queries on the broadphase BVH (b2g_query.cuh)
// ---------------------------------------------------------------------------------------------
// the tree must reflect every upload made since the last step (b2BroadPhase::EnsureBuiltTree)
static int query_prepare(b2gArena* A) {
  CK(cudaSetDevice(A->device));
  if (A->newFixtures || A->aabbAllDirty || A->bvhLeaves != A->nFixtures) {
    CK(cudaMemsetAsync(A->dCounts, 0, sizeof(StepCounts), A->stream));
    int rc = find_new_contacts(A, 0);
    if (rc) return rc;
  }
  return B2G_OK;
}
// host or device pointer -> device pointer on the arena's stream (staging through `buf` for host memory)
static cudaError_t q_in(b2gArena* A, DevBuf& buf, const void* src, size_t bytes, int onDevice, const void** out) {
  if (!src || onDevice) {
    *out = src;
    return cudaSuccess;
  }
  cudaError_t e = buf.alloc(bytes);
  if (e != cudaSuccess) return e;
  *out = buf.p;
  return cudaMemcpyAsync(buf.p, src, bytes, cudaMemcpyHostToDevice, A->stream);
}
static cudaError_t q_out(DevBuf& buf, void* dst, size_t bytes, int onDevice, void** out) {
  if (onDevice) {
    *out = dst;
    return cudaSuccess;
  }
  cudaError_t e = buf.alloc(bytes);
  *out = buf.p;
  return e;
}
#define Q_BACK(buf, dst, bytes)                                                                        \
  if (!on_device && (dst)) CK(cudaMemcpyAsync((dst), (buf).p, (bytes), cudaMemcpyDeviceToHost, A->stream))

extern "C" int b2g_query_aabb(b2gArena* A, int32_t n, const float* aabbs, const int32_t* world, int32_t cap,
                              int32_t* counts, int32_t* fixtures, int32_t on_device) {
  if (!A || n < 0 || cap < 0 || (n > 0 && (!aabbs || !counts || (cap > 0 && !fixtures)))) return B2G_ERR_INVALID;
  int rc = query_prepare(A);
  if (rc) return rc;
  if (n == 0) return B2G_OK;
  DevBuf bQ, bW, bC, bF;
  const void *dQ, *dW;
  void *dC, *dF;
  CK(q_in(A, bQ, aabbs, (size_t)n * 16, on_device, &dQ));
  CK(q_in(A, bW, world, (size_t)n * 4, on_device, &dW));
  CK(q_out(bC, counts, (size_t)n * 4, on_device, &dC));
  CK(q_out(bF, fixtures, (size_t)n * (cap ? cap : 1) * 4, on_device, &dF));
  if (A->nFixtures == 0) {
    CK(cudaMemsetAsync(dC, 0, (size_t)n * 4, A->stream));
  } else {
    LAUNCH(A, KC_QUERY, n, k_query_aabb, div_up(n, 128), 128, n, (const float4*)dQ, (const int*)dW, wide_tree(A),
           A->leafBox, A->leafInfo, A->leafKey, A->worldFirst, A->worldLast, A->numWorlds, cap, (int*)dC, (int*)dF);
    CK(cudaGetLastError());
  }
  Q_BACK(bC, counts, (size_t)n * 4);
  Q_BACK(bF, fixtures, (size_t)n * cap * 4);
  CK(cudaStreamSynchronize(A->stream));
  return B2G_OK;
}

static int ray_cast_common(b2gArena* A, int mode, int32_t n, const float* rays, const float* max_fraction,
                           const int32_t* world, uint32_t category_mask, int32_t cap, int32_t* counts,
                           int32_t* fixture, float* fraction, float* normal, int32_t on_device) {
  int rc = query_prepare(A);
  if (rc) return rc;
  if (n == 0) return B2G_OK;
  const size_t per = mode == 0 ? 1 : (size_t)(cap ? cap : 1);
  DevBuf bR, bM, bW, bC, bF, bT, bN;
  const void *dR, *dM, *dW;
  void *dC = nullptr, *dF, *dT, *dN;
  CK(q_in(A, bR, rays, (size_t)n * 16, on_device, &dR));
  CK(q_in(A, bM, max_fraction, (size_t)n * 4, on_device, &dM));
  CK(q_in(A, bW, world, (size_t)n * 4, on_device, &dW));
  if (mode == 1) CK(q_out(bC, counts, (size_t)n * 4, on_device, &dC));
  CK(q_out(bF, fixture, (size_t)n * per * 4, on_device, &dF));
  CK(q_out(bT, fraction, (size_t)n * per * 4, on_device, &dT));
  CK(q_out(bN, normal, (size_t)n * per * 8, on_device, &dN));
  LAUNCH(A, KC_QUERY, n, k_ray_cast, div_up(n, 128), 128, n, (const float4*)dR, (const float*)dM, (const int*)dW, mode,
         category_mask, wide_tree(A), A->leafBox, A->leafInfo, A->leafKey, A->worldFirst, A->worldLast, A->numWorlds,
         A->fShapeOff, A->shapes, A->xf, cap, (int*)dC, (int*)dF, (float*)dT, (float2*)dN);
  CK(cudaGetLastError());
  if (mode == 1) Q_BACK(bC, counts, (size_t)n * 4);
  Q_BACK(bF, fixture, (size_t)n * per * 4);
  Q_BACK(bT, fraction, (size_t)n * per * 4);
  Q_BACK(bN, normal, (size_t)n * per * 8);
  CK(cudaStreamSynchronize(A->stream));
  return B2G_OK;
}

extern "C" int b2g_ray_cast_closest(b2gArena* A, int32_t n, const float* rays, const float* max_fraction,
                                    const int32_t* world, uint32_t category_mask, int32_t* fixture, float* fraction,
                                    float* normal, int32_t on_device) {
  if (!A || n < 0 || (n > 0 && (!rays || !fixture || !fraction || !normal))) return B2G_ERR_INVALID;
  return ray_cast_common(A, 0, n, rays, max_fraction, world, category_mask, 1, nullptr, fixture, fraction, normal,
                         on_device);
}

extern "C" int b2g_ray_cast_all(b2gArena* A, int32_t n, const float* rays, const float* max_fraction,
                                const int32_t* world, uint32_t category_mask, int32_t cap, int32_t* counts,
                                int32_t* fixture, float* fraction, float* normal, int32_t on_device) {
  if (!A || n < 0 || cap < 0 || (n > 0 && (!rays || !counts || (cap > 0 && (!fixture || !fraction || !normal)))))
    return B2G_ERR_INVALID;
  return ray_cast_common(A, 1, n, rays, max_fraction, world, category_mask, cap, counts, fixture, fraction, normal,
                         on_device);
}

// ---------------------------------------------------------------------------------------------
// user contact filter (b2ContactFilter::ShouldCollide): the callback lives on the host, so the host
// inspects the pairs each refresh inserted and hands the rejected ones back as a veto list
// ---------------------------------------------------------------------------------------------
extern "C" int b2g_download_new_pairs(b2gArena* A, int32_t capacity, int32_t* fixture_a, int32_t* fixture_b,
                                      int32_t* count) {
  if (!A || !count || capacity < 0) return B2G_ERR_INVALID;
  CK(cudaSetDevice(A->device));
  int n = A->lastNewPairs;
  *count = n;
  if (n > capacity) n = capacity;
  if (n == 0 || !fixture_a || !fixture_b) return B2G_OK;
  std::vector<unsigned long long> keys((size_t)n);
  CK(cudaMemcpyAsync(keys.data(), A->pairKeys, (size_t)n * 8, cudaMemcpyDeviceToHost, A->stream));
  CK(cudaStreamSynchronize(A->stream));
  for (int i = 0; i < n; ++i) {
    fixture_a[i] = (int32_t)(keys[i] >> 32);
    fixture_b[i] = (int32_t)(keys[i] & 0xffffffffull);
  }
  return B2G_OK;
}

extern "C" int b2g_set_pair_vetoes(b2gArena* A, int32_t count, const int32_t* fixture_a, const int32_t* fixture_b) {
  if (!A || count < 0 || (count > 0 && (!fixture_a || !fixture_b))) return B2G_ERR_INVALID;
  CK(cudaSetDevice(A->device));
  std::vector<unsigned long long> keys((size_t)count);
  for (int i = 0; i < count; ++i) {
    unsigned long long lo = (unsigned long long)std::min(fixture_a[i], fixture_b[i]);
    unsigned long long hi = (unsigned long long)std::max(fixture_a[i], fixture_b[i]);
    keys[i] = (lo << 32) | hi;
  }
  std::sort(keys.begin(), keys.end());
  keys.erase(std::unique(keys.begin(), keys.end()), keys.end());
  count = (int32_t)keys.size();
  if (count > A->vetoCap) {
    int cap = A->vetoCap > 0 ? A->vetoCap : 256;
    while (cap < count) cap *= 2;
    CK(cudaStreamSynchronize(A->stream));
    if (A->vetoKeys) cudaFree(A->vetoKeys);
    if (A->vetoSeen) cudaFree(A->vetoSeen);
    CK(dalloc(&A->vetoKeys, (size_t)cap));
    CK(dalloc(&A->vetoSeen, (size_t)cap));
    A->vetoCap = cap;
  }
  if (count > 0) {
    CK(cudaMemcpyAsync(A->vetoKeys, keys.data(), (size_t)count * 8, cudaMemcpyHostToDevice, A->stream));
    CK(cudaMemsetAsync(A->vetoSeen, 0, (size_t)count, A->stream));
    // contacts of vetoed pairs leave before any narrowphase sees them
    DevBuf dRemoved;
    CK(dRemoved.alloc(4));
    CK(cudaMemsetAsync(dRemoved.p, 0, 4, A->stream));
    k_contacts_remove<<<div_up(count, 256), 256, 0, A->stream>>>(count, A->vetoKeys, A->cb[0], A->hash, A->freeStack,
                                                                 A->dFreeTop, dRemoved.as<int>());
    CK(cudaGetLastError());
    int removed = 0;
    CK(cudaMemcpyAsync(&removed, dRemoved.p, 4, cudaMemcpyDeviceToHost, A->stream));
    CK(cudaStreamSynchronize(A->stream));
    A->nAlive -= removed;
    A->tombstones += removed;
  }
  CK(cudaStreamSynchronize(A->stream));
  A->hash.vetoKeys = A->vetoKeys;
  A->hash.vetoSeen = A->vetoSeen;
  A->hash.vetoCount = count;
  return B2G_OK;
}

extern "C" int b2g_download_veto_seen(b2gArena* A, int32_t count, uint8_t* seen) {
  if (!A || count < 0 || count > A->hash.vetoCount || (count > 0 && !seen)) return B2G_ERR_INVALID;
  if (count == 0) return B2G_OK;
  CK(cudaSetDevice(A->device));
  CK(cudaMemcpyAsync(seen, A->vetoSeen, (size_t)count, cudaMemcpyDeviceToHost, A->stream));
  CK(cudaMemsetAsync(A->vetoSeen, 0, (size_t)count, A->stream));
  CK(cudaStreamSynchronize(A->stream));
  return B2G_OK;
}

extern "C" int b2g_set_sequential_joint_order(b2gArena* A, int32_t count, const int32_t* joints) {
  if (!A || count < 0 || count > A->capJoints || (count > 0 && !joints)) return B2G_ERR_INVALID;
  for (int i = 0; i < count; ++i)
    if (joints[i] < 0 || joints[i] >= A->nJoints) return B2G_ERR_INVALID;
  CK(cudaSetDevice(A->device));
  if (count > 0) CK(cudaMemcpyAsync(A->jointOrder, joints, (size_t)count * 4, cudaMemcpyHostToDevice, A->stream));
  CK(cudaStreamSynchronize(A->stream));
  A->jointOrderCount = count;
  A->jointOrderActive = 1;
  return B2G_OK;
}

extern "C" int b2g_download_events(b2gArena* A, int32_t* beginPairs, int32_t* beginCount, int32_t* endPairs,
                                   int32_t* endCount, int32_t capacityEach) {
  if (!A || !beginCount || !endCount) return B2G_ERR_INVALID;
  CK(cudaSetDevice(A->device));
  int nbg = A->hCounts->beginCount, nen = A->hCounts->endCount;
  if (nbg > A->capContacts) nbg = A->capContacts;
  if (nen > A->capContacts) nen = A->capContacts;
  *beginCount = nbg;
  *endCount = nen;
  int cb = nbg < capacityEach ? nbg : capacityEach, ce = nen < capacityEach ? nen : capacityEach;
  if (beginPairs && cb > 0)
    CK(cudaMemcpyAsync(beginPairs, A->beginEvents, (size_t)cb * 8, cudaMemcpyDeviceToHost, A->stream));
  if (endPairs && ce > 0)
    CK(cudaMemcpyAsync(endPairs, A->endEvents, (size_t)ce * 8, cudaMemcpyDeviceToHost, A->stream));
  CK(cudaStreamSynchronize(A->stream));
  return B2G_OK;
}

extern "C" int b2g_device_views(b2gArena* A, b2gDeviceViews* out) {
  if (!A || !out) return B2G_ERR_INVALID;
  out->pos = A->pos;
  out->vel = A->vel;
  out->xf = A->xf;
  out->force = A->force;
  out->flags = A->bflags;
  out->capacity = A->capBodies;
  out->device = A->device;
  return B2G_OK;
}

// ---------------------------------------------------------------------------------------------
// Halo exchange of a spatially decomposed world (BASELINE config 5).  The transport (NCCL in
// libb2cuda_dist.so, or a device-to-device copy in the single-process emulation) moves opaque messages; what
// is in them is decided here: 4 quads per body — pos, vel, xf, (flags, 0, 0, 0) — gathered and scattered by two
// kernels on the arena's stream, so an exchange needs no host synchronisation at all.
// ---------------------------------------------------------------------------------------------
__global__ void k_halo_pack(int n, const int* __restrict__ bodies, const float4* __restrict__ pos, const float4* __restrict__ vel,
                            const float4* __restrict__ xf, const uint32_t* __restrict__ bflags, float4* out) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const int b = bodies[k];
  out[4 * k + 0] = pos[b];
  out[4 * k + 1] = vel[b];
  out[4 * k + 2] = xf[b];
  out[4 * k + 3] = make_float4(__uint_as_float(bflags[b]), 0.0f, 0.0f, 0.0f);
}
__global__ void k_halo_unpack(int n, const int* __restrict__ bodies, const float4* __restrict__ in, float4* pos, float4* vel,
                              float4* xf, uint32_t* bflags) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const int b = bodies[k];
  pos[b] = in[4 * k + 0];
  vel[b] = in[4 * k + 1];
  xf[b] = in[4 * k + 2];
  bflags[b] = __float_as_uint(in[4 * k + 3].x);
}
extern "C" int b2g_halo_set_lists(b2gArena* A, int32_t slot, int32_t n_send, const int32_t* send_bodies, int32_t n_recv,
                                  const int32_t* recv_bodies) {
  if (!A || slot < 0 || slot > 1 || n_send < 0 || n_recv < 0 || (n_send > 0 && !send_bodies) || (n_recv > 0 && !recv_bodies))
    return B2G_ERR_INVALID;
  CK(cudaSetDevice(A->device));
  CK(cudaStreamSynchronize(A->stream));
  cudaFree(A->haloSend[slot]);
  cudaFree(A->haloRecv[slot]);
  cudaFree(A->haloOut[slot]);
  cudaFree(A->haloIn[slot]);
  A->haloSend[slot] = A->haloRecv[slot] = nullptr;
  A->haloOut[slot] = A->haloIn[slot] = nullptr;
  A->haloNumSend[slot] = n_send;
  A->haloNumRecv[slot] = n_recv;
  for (int k = 0; k < n_send; ++k)
    if (send_bodies[k] < 0 || send_bodies[k] >= A->capBodies) return B2G_ERR_INVALID;
  for (int k = 0; k < n_recv; ++k)
    if (recv_bodies[k] < 0 || recv_bodies[k] >= A->capBodies) return B2G_ERR_INVALID;
  if (n_send > 0) {
    CK(dalloc(&A->haloSend[slot], n_send));
    CK(dalloc(&A->haloOut[slot], (size_t)n_send * 4));
    CK(cudaMemcpy(A->haloSend[slot], send_bodies, sizeof(int) * (size_t)n_send, cudaMemcpyHostToDevice));
  }
  if (n_recv > 0) {
    CK(dalloc(&A->haloRecv[slot], n_recv));
    CK(dalloc(&A->haloIn[slot], (size_t)n_recv * 4));
    CK(cudaMemcpy(A->haloRecv[slot], recv_bodies, sizeof(int) * (size_t)n_recv, cudaMemcpyHostToDevice));
  }
  CK(cudaDeviceSynchronize());
  return B2G_OK;
}
extern "C" int b2g_halo_pack(b2gArena* A, int32_t slot, void** message, int64_t* bytes) {
  if (!A || slot < 0 || slot > 1) return B2G_ERR_INVALID;
  CK(cudaSetDevice(A->device));
  const int n = A->haloNumSend[slot];
  if (n > 0) {
    k_halo_pack<<<div_up(n, 128), 128, 0, A->stream>>>(n, A->haloSend[slot], A->pos, A->vel, A->xf, A->bflags, A->haloOut[slot]);
    CK(cudaGetLastError());
    A->launches++;
  }
  if (message) *message = A->haloOut[slot];
  if (bytes) *bytes = (int64_t)n * 64;
  return B2G_OK;
}
extern "C" int b2g_halo_recv_buffer(b2gArena* A, int32_t slot, void** message, int64_t* bytes) {
  if (!A || slot < 0 || slot > 1) return B2G_ERR_INVALID;
  if (message) *message = A->haloIn[slot];
  if (bytes) *bytes = (int64_t)A->haloNumRecv[slot] * 64;
  return B2G_OK;
}
extern "C" int b2g_halo_unpack(b2gArena* A, int32_t slot, const void* message) {
  if (!A || slot < 0 || slot > 1) return B2G_ERR_INVALID;
  CK(cudaSetDevice(A->device));
  const int n = A->haloNumRecv[slot];
  if (n > 0) {
    const float4* src = message ? (const float4*)message : A->haloIn[slot];
    k_halo_unpack<<<div_up(n, 128), 128, 0, A->stream>>>(n, A->haloRecv[slot], src, A->pos, A->vel, A->xf, A->bflags);
    CK(cudaGetLastError());
    A->launches++;
  }
  return B2G_OK;
}

extern "C" int b2g_synchronize(b2gArena* A) {
  if (!A) return B2G_ERR_INVALID;
  CK(cudaSetDevice(A->device));
  CK(cudaStreamSynchronize(A->stream));
  return B2G_OK;
}
extern "C" void* b2g_stream(b2gArena* A) { return A ? (void*)A->stream : nullptr; }
extern "C" int b2g_set_profiling(b2gArena* A, int32_t on) {
  if (!A) return B2G_ERR_INVALID;
  A->profiling = on;
  return B2G_OK;
}
extern "C" int b2g_set_kernel_timing(b2gArena* A, int32_t on) {
  if (!A) return B2G_ERR_INVALID;
  A->kernelTiming = on;
  A->ktCount = 0;
  for (int c = 0; c < KC_COUNT; ++c) {
    A->ktMs[c] = 0.0;
    A->ktLaunches[c] = 0;
    A->ktUnitsSum[c] = 0.0;
  }
  return B2G_OK;
}
extern "C" int b2g_kernel_class_count(void) { return KC_COUNT; }
extern "C" const char* b2g_kernel_class_name(int32_t cls) {
  return (cls >= 0 && cls < KC_COUNT) ? kClassNames[cls] : "";
}
extern "C" int b2g_get_kernel_timing(b2gArena* A, int32_t cls, double* total_ms, int64_t* launches,
                                     double* units) {
  if (!A || cls < 0 || cls >= KC_COUNT) return B2G_ERR_INVALID;
  if (total_ms) *total_ms = A->ktMs[cls];
  if (launches) *launches = A->ktLaunches[cls];
  if (units) *units = A->ktUnitsSum[cls];
  return B2G_OK;
}
#ifdef B2G_BIG_TRACE
extern "C" int b2g_debug_big_trace(unsigned long long* out, int32_t blocks) {
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpyFromSymbol(out, g_bigTrace, sizeof(unsigned long long) * 2 * B2G_TRACE_CAP * (size_t)blocks));
  return B2G_OK;
}
#endif
#ifdef B2G_BIG_TRACE
extern "C" int b2g_debug_tile_marks(unsigned long long* out) {
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpyFromSymbol(out, g_tileMarks, sizeof(unsigned long long) * B2G_TILES_MAX * B2G_TILE_MARKS));
  return B2G_OK;
}
#endif
// debug / tests: the current tile plan (11 words: lo[2], hi[2], count, S, R, x0, invDx, y0, invDy), the bodies per
// tile of the last step and every body's tile slot
extern "C" int b2g_debug_tile_state(b2gArena* A, uint32_t* plan, int32_t* tileCount, int32_t* tileSlot) {
  if (!A) return B2G_ERR_INVALID;
  CK(cudaSetDevice(A->device));
  CK(cudaStreamSynchronize(A->stream));
  if (plan) CK(cudaMemcpy(plan, A->tilePlan, sizeof(TilePlan), cudaMemcpyDeviceToHost));
  if (tileCount) CK(cudaMemcpy(tileCount, A->tileCount, sizeof(int) * B2G_TILES_MAX, cudaMemcpyDeviceToHost));
  if (tileSlot) CK(cudaMemcpy(tileSlot, A->tileSlot, sizeof(int) * (size_t)A->nBodies, cudaMemcpyDeviceToHost));
  return B2G_OK;
}
// debug / tests: the joint colouring the next production-mode step will use (colour per joint; B2G_JOINT_COLOURS =
// serial tail).  Runs the colouring first if the joint table changed since the last step.
extern "C" int b2g_debug_joint_colours(b2gArena* A, int32_t* colour) {
  if (!A || !colour) return B2G_ERR_INVALID;
  CK(cudaSetDevice(A->device));
  if (A->nJoints > 0 && A->jointColourDirty) {
    LAUNCH(A, KC_COLOUR, A->nJoints, k_joint_colour, 1, B2G_JOINT_COLOUR_THREADS, A->nJoints, A->nBodies, A->jBodies,
           A->jParams1, A->mass, A->jBodyMask, A->jBodyBest, A->jColour, A->jSorted, A->jCstart);
    A->jointColourDirty = 0;
  }
  CK(cudaStreamSynchronize(A->stream));
  if (A->nJoints > 0) CK(cudaMemcpy(colour, A->jColour, sizeof(int) * (size_t)A->nJoints, cudaMemcpyDeviceToHost));
  return B2G_OK;
}
extern "C" int b2g_set_inv_dt0(b2gArena* A, float v) {
  if (!A) return B2G_ERR_INVALID;
  A->invDt0 = v;
  return B2G_OK;
}
extern "C" int b2g_host_alloc(void** out, uint64_t bytes) {
  if (!out) return B2G_ERR_INVALID;
  CK(cudaMallocHost(out, bytes ? bytes : 1));
  return B2G_OK;
}
extern "C" int b2g_host_free(void* p) {
  CK(cudaFreeHost(p));
  return B2G_OK;
}

#include "b2g_kernel_entries.cuh"

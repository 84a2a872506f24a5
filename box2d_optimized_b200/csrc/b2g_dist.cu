// b2g_dist.cu — NCCL transport of the halo exchange (libb2cuda_dist.so).
//
// SURVEY §8e, BASELINE config 5: one very large world cut into x-slabs, one arena per GPU, boundary-body halos
// exchanged once per step over NVLink.  The messages are packed and scattered by libb2cuda.so
// (b2g_halo_pack / b2g_halo_unpack, kernels on the arena's stream); this file only moves them: one ncclSend and
// one ncclRecv per neighbour inside one group, on the SAME stream, so a step + exchange is one uninterrupted
// stream of work and the host never waits.  Messages are a few hundred KB at most (a 3 m strip of a settled
// pile): the exchange is latency-bound, which is why nothing here tries to overlap or pipeline it.
#include <cuda_runtime.h>
#include <nccl.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include "../../include/b2cuda.h"

struct b2gDist {
  ncclComm_t comm;
  int rank, nranks, device;
};

static int fail(const char* what, const char* detail) {
  fprintf(stderr, "[b2cuda_dist] %s: %s\n", what, detail);
  return B2G_ERR_CUDA;
}
#define NCK(call)                                              \
  do {                                                         \
    ncclResult_t r_ = (call);                                  \
    if (r_ != ncclSuccess) return fail(#call, ncclGetErrorString(r_)); \
  } while (0)

extern "C" int b2g_dist_unique_id(void* out128) {
  if (!out128) return B2G_ERR_INVALID;
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  ncclUniqueId id;
  NCK(ncclGetUniqueId(&id));
  memcpy(out128, &id, sizeof(id));
  return B2G_OK;
}

extern "C" int b2g_dist_init(const void* unique_id128, int32_t rank, int32_t nranks, int32_t device, b2gDist** out) {
  if (!unique_id128 || !out || rank < 0 || rank >= nranks) return B2G_ERR_INVALID;
  if (cudaSetDevice(device) != cudaSuccess) return fail("cudaSetDevice", "failed");
  ncclUniqueId id;
  memcpy(&id, unique_id128, sizeof(id));
  b2gDist* D = (b2gDist*)calloc(1, sizeof(b2gDist));
  D->rank = rank;
  D->nranks = nranks;
  D->device = device;
  NCK(ncclCommInitRank(&D->comm, nranks, id, rank));
  *out = D;
  return B2G_OK;
}

extern "C" int b2g_dist_exchange(b2gDist* D, b2gArena* A) {
  if (!D || !A) return B2G_ERR_INVALID;
  cudaStream_t st = (cudaStream_t)b2g_stream(A);
  void* out[2] = {nullptr, nullptr};
  void* in[2] = {nullptr, nullptr};
  int64_t outBytes[2] = {0, 0}, inBytes[2] = {0, 0};
  for (int slot = 0; slot < 2; ++slot) {
    int rc = b2g_halo_pack(A, slot, &out[slot], &outBytes[slot]);
    if (rc) return rc;
    rc = b2g_halo_recv_buffer(A, slot, &in[slot], &inBytes[slot]);
    if (rc) return rc;
  }
  NCK(ncclGroupStart());
  for (int slot = 0; slot < 2; ++slot) {
    const int peer = slot == 0 ? D->rank - 1 : D->rank + 1;
    if (peer < 0 || peer >= D->nranks) continue;
    if (outBytes[slot] > 0) NCK(ncclSend(out[slot], (size_t)outBytes[slot], ncclChar, peer, D->comm, st));
    if (inBytes[slot] > 0) NCK(ncclRecv(in[slot], (size_t)inBytes[slot], ncclChar, peer, D->comm, st));
  }
  NCK(ncclGroupEnd());
  for (int slot = 0; slot < 2; ++slot) {
    int rc = b2g_halo_unpack(A, slot, nullptr);
    if (rc) return rc;
  }
  return B2G_OK;
}

extern "C" int b2g_dist_destroy(b2gDist* D) {
  if (!D) return B2G_OK;
  ncclCommDestroy(D->comm);
  free(D);
  return B2G_OK;
}

// microbenchmark: cost of one grid-wide barrier on B200, cooperative groups vs a hand-rolled
// monotonic-counter barrier.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gb grid_barrier.cu
#include <cooperative_groups.h>
#include <cstdio>
namespace cg = cooperative_groups;

__global__ void k_cg(int n, float* out) {
  cg::grid_group g = cg::this_grid();
  float acc = 0.f;
  for (int i = 0; i < n; ++i) {
    acc += out[(threadIdx.x + i) & 1023];
    g.sync();
  }
  if (acc == 123.f) out[0] = acc;
}

__device__ __forceinline__ void barrier(unsigned int* counter, unsigned int& target) {
  __syncthreads();
  target += gridDim.x;
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(counter, 1u);
    while (*((volatile unsigned int*)counter) < target) {}
    __threadfence();
  }
  __syncthreads();
}
// release/acquire flavour: one red.release + ld.acquire spin, no full __threadfence pair
__device__ __forceinline__ void barrier_ra(unsigned int* counter, unsigned int& target) {
  __syncthreads();
  target += gridDim.x;
  if (threadIdx.x == 0) {
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
    unsigned int v;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
    } while (v < target);
  }
  __syncthreads();
}
__global__ void k_ra(int n, float* out, unsigned int* counter) {
  unsigned int target = 0;
  float acc = 0.f;
  for (int i = 0; i < n; ++i) {
    acc += out[(threadIdx.x + i) & 1023];
    barrier_ra(counter, target);
  }
  if (acc == 123.f) out[0] = acc;
}

__global__ void k_own(int n, float* out, unsigned int* counter) {
  unsigned int target = 0;
  float acc = 0.f;
  for (int i = 0; i < n; ++i) {
    acc += out[(threadIdx.x + i) & 1023];
    barrier(counter, target);
  }
  if (acc == 123.f) out[0] = acc;
}

int main() {
  float* out; unsigned int* counter;
  cudaMalloc(&out, 4096); cudaMemset(out, 0, 4096);
  cudaMalloc(&counter, 4);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (int threads : {128, 256, 1024}) {
    for (int blocksPerSm : {1, 2}) {
      int grid = 148 * blocksPerSm;
      if (threads == 1024 && blocksPerSm == 2) continue;
      int n = 2000;
      void* args[] = {&n, &out};
      cudaLaunchCooperativeKernel((void*)k_cg, dim3(grid), dim3(threads), args, 0, 0);
      cudaDeviceSynchronize();
      cudaEventRecord(a);
      cudaLaunchCooperativeKernel((void*)k_cg, dim3(grid), dim3(threads), args, 0, 0);
      cudaEventRecord(b); cudaEventSynchronize(b);
      float ms; cudaEventElapsedTime(&ms, a, b);
      cudaMemset(counter, 0, 4);
      void* args2[] = {&n, &out, &counter};
      cudaLaunchCooperativeKernel((void*)k_own, dim3(grid), dim3(threads), args2, 0, 0);
      cudaDeviceSynchronize();
      cudaMemset(counter, 0, 4);
      cudaEventRecord(a);
      cudaLaunchCooperativeKernel((void*)k_own, dim3(grid), dim3(threads), args2, 0, 0);
      cudaEventRecord(b); cudaEventSynchronize(b);
      float ms2; cudaEventElapsedTime(&ms2, a, b);
      cudaMemset(counter, 0, 4);
      cudaLaunchCooperativeKernel((void*)k_ra, dim3(grid), dim3(threads), args2, 0, 0);
      cudaDeviceSynchronize();
      cudaMemset(counter, 0, 4);
      cudaEventRecord(a);
      cudaLaunchCooperativeKernel((void*)k_ra, dim3(grid), dim3(threads), args2, 0, 0);
      cudaEventRecord(b); cudaEventSynchronize(b);
      float ms3; cudaEventElapsedTime(&ms3, a, b);
      printf("grid %d x %d threads: cg grid.sync %.3f us, own barrier %.3f us, release/acquire %.3f us  (%s)\n", grid, threads,
             1000.f * ms / n, 1000.f * ms2 / n, 1000.f * ms3 / n, cudaGetErrorString(cudaGetLastError()));
    }
  }
  return 0;
}

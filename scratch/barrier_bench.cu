// micro-benchmark: cost of a grid-wide barrier in a persistent cooperative kernel
#include <cstdio>
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include <stdint.h>
namespace cg = cooperative_groups;

__device__ __forceinline__ void grid_barrier_fence(uint32_t* bar, uint32_t nblocks) {
  __syncthreads();
  if (threadIdx.x == 0) {
    volatile uint32_t* vgen = bar + 1;
    const uint32_t gen = *vgen;
    __threadfence();
    const uint32_t prev = atomicAdd(bar, 1u);
    if (prev == nblocks - 1) { bar[0] = 0u; __threadfence(); *vgen = gen + 1u; }
    else { while (*vgen == gen) {} }
    __threadfence();
  }
  __syncthreads();
}
// release/acquire flavour: one red.release + ld.acquire polling, no stand-alone fences
__device__ __forceinline__ void grid_barrier_ra(uint32_t* bar, uint32_t nblocks, uint32_t& target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    target += nblocks;
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(bar) : "memory");
    uint32_t v;
    do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory"); } while ((int32_t)(v - target) < 0);
  }
  __syncthreads();
}
__global__ void k_fence(uint32_t* bar, int iters, float* sink) {
  float x = threadIdx.x;
  for (int i = 0; i < iters; ++i) { x = x * 1.0001f + 1.0f; sink[blockIdx.x * blockDim.x + threadIdx.x] = x; grid_barrier_fence(bar, gridDim.x); }
}
__global__ void k_ra(uint32_t* bar, int iters, float* sink) {
  float x = threadIdx.x; uint32_t target = 0;
  for (int i = 0; i < iters; ++i) { x = x * 1.0001f + 1.0f; sink[blockIdx.x * blockDim.x + threadIdx.x] = x; grid_barrier_ra(bar + 2, gridDim.x, target); }
}
__global__ void k_cg(uint32_t* bar, int iters, float* sink) {
  float x = threadIdx.x; cg::grid_group g = cg::this_grid();
  for (int i = 0; i < iters; ++i) { x = x * 1.0001f + 1.0f; sink[blockIdx.x * blockDim.x + threadIdx.x] = x; g.sync(); }
}
int main() {
  uint32_t* bar; float* sink;
  cudaMalloc(&bar, 64); cudaMalloc(&sink, 4 * 1024 * 1024);
  int iters = 2000;
  const void* ks[3] = {(const void*)k_fence, (const void*)k_ra, (const void*)k_cg};
  const char* nm[3] = {"fence+atomic+volatile-poll", "red.release + ld.acquire", "cooperative_groups grid.sync"};
  int cfg[5][2] = {{148, 256}, {296, 256}, {592, 256}, {148, 1024}, {148, 512}};
  for (int c = 0; c < 5; ++c) for (int k = 0; k < 3; ++k) {
    cudaMemset(bar, 0, 64);
    void* args[3] = {&bar, &iters, &sink};
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaLaunchCooperativeKernel(ks[k], dim3(cfg[c][0]), dim3(cfg[c][1]), args, 0, 0);
    cudaDeviceSynchronize(); cudaMemset(bar, 0, 64);
    cudaEventRecord(e0);
    cudaError_t e = cudaLaunchCooperativeKernel(ks[k], dim3(cfg[c][0]), dim3(cfg[c][1]), args, 0, 0);
    cudaEventRecord(e1); cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("grid %4d x %4d  %-30s %8.3f us/barrier  (%s)\n", cfg[c][0], cfg[c][1], nm[k], 1e3 * ms / iters, cudaGetErrorString(e));
  }
  return 0;
}

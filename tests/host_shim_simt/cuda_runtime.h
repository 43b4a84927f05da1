// TEST INFRASTRUCTURE: a stand-in for <cuda_runtime.h> with REAL block semantics on the host.
//
// Unlike tests/host_shim (one "thread" at a time, barriers as no-ops), every CUDA thread of a block is an OS thread
// here: __syncthreads is a barrier over the block, warp shuffles exchange values between the 32 lanes of a warp
// through a per-warp mailbox, atomics are atomic, __shared__ variables are statics shared by the block.  Blocks run
// one after the other (simt::launch), so a cooperative grid barrier is only available for grids of ONE block.
// This executes the block-level kernels of genjax_b200/csrc (tile scans, max-scan write-out, block reductions,
// lane-group model kernels) as written, on a CPU.  It does not model memory ordering beyond sequential consistency
// at barriers, divergence inside a warp at a shuffle (every lane named in the mask must arrive), or timing.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <atomic>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>

#define __device__
#define __host__
#define __global__
#define __forceinline__ inline
#define __restrict__
#define __shared__ static
#define __launch_bounds__(...)
#define __grid_constant__
#define __align__(n) __attribute__((aligned(n)))

struct uint4 { uint32_t x, y, z, w; };
struct uint2 { uint32_t x, y; };
static inline uint2 make_uint2(uint32_t x, uint32_t y) { return uint2{x, y}; }
struct int4 { int32_t x, y, z, w; };
struct int2 { int32_t x, y; };
struct ulonglong2 { unsigned long long x, y; };
static inline ulonglong2 make_ulonglong2(unsigned long long x, unsigned long long y) { return ulonglong2{x, y}; }
struct float2 { float x, y; };
struct float4 { float x, y, z, w; };
struct dim3 { unsigned x = 1, y = 1, z = 1; dim3() {} dim3(unsigned a, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return uint4{x, y, z, w}; }
static inline int4 make_int4(int32_t x, int32_t y, int32_t z, int32_t w) { return int4{x, y, z, w}; }
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
static thread_local dim3 threadIdx, blockIdx;
static dim3 blockDim, gridDim;

namespace simt {
struct Barrier {  // reusable counting barrier with an OR-reduction slot
  std::mutex m; std::condition_variable cv; int count = 0, n = 1; unsigned gen = 0; int acc = 0, result = 0;
  int arrive(int pred = 0) {
    std::unique_lock<std::mutex> lk(m);
    acc |= pred;
    const unsigned g = gen;
    if (++count == n) { result = acc; acc = 0; count = 0; ++gen; cv.notify_all(); return result; }
    cv.wait(lk, [&] { return gen != g; });
    return result;
  }
};
static Barrier block_barrier;
static Barrier warp_barrier[64];
static uint64_t mailbox[64][32];
template <class F>
static void launch(unsigned grid, unsigned block, F body) {
  gridDim = dim3(grid); blockDim = dim3(block);
  block_barrier.n = (int)block;
  for (unsigned w = 0; w < (block + 31) / 32; ++w) warp_barrier[w].n = (int)((w + 1) * 32 <= block ? 32 : block - w * 32);
  for (unsigned b = 0; b < grid; ++b) {
    std::vector<std::thread> ts;
    for (unsigned t = 0; t < block; ++t)
      ts.emplace_back([=] { blockIdx = dim3(b); threadIdx = dim3(t); body(); });
    for (auto& t : ts) t.join();
  }
}
template <class T>
static inline T exchange(T v, int src_lane) {  // every lane of the warp publishes v, then reads lane src_lane's value
  static_assert(sizeof(T) <= 8, "shuffle payload");
  const unsigned w = threadIdx.x >> 5, l = threadIdx.x & 31;
  uint64_t raw = 0; memcpy(&raw, &v, sizeof(T));
  mailbox[w][l] = raw;
  warp_barrier[w].arrive();
  uint64_t got = mailbox[w][(src_lane >= 0 && src_lane < warp_barrier[w].n) ? src_lane : (int)l];
  warp_barrier[w].arrive();
  T out; memcpy(&out, &got, sizeof(T));
  return out;
}
}  // namespace simt

static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }  // glibc fmaf: correctly rounded
static inline float __fsqrt_rn(float x) { return sqrtf(x); }
static inline unsigned long long __float2ull_rn(float x) { return (unsigned long long)rintf(x); }
static inline int __float2int_ru(float x) { return (int)ceilf(x); }
static inline int __float2int_rd(float x) { return (int)floorf(x); }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __dsub_rn(double a, double b) { return a - b; }
static inline double __ddiv_rn(double a, double b) { return a / b; }
static inline int __double2int_ru(double x) { return (int)ceil(x); }
static inline uint32_t __umulhi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
static inline uint64_t __umul64hi(uint64_t a, uint64_t b) { return (uint64_t)(((unsigned __int128)a * b) >> 64); }
static inline uint32_t __float_as_uint(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float __uint_as_float(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline int __float_as_int(float f) { int i; memcpy(&i, &f, 4); return i; }
static inline float __int_as_float(int i) { float f; memcpy(&f, &i, 4); return f; }
static inline void sincospif(float x, float* s, float* c) { *s = sinf(3.14159265358979323846f * x); *c = cosf(3.14159265358979323846f * x); }

template <class T> static inline T __shfl_up_sync(unsigned, T v, int d) { const int l = threadIdx.x & 31; return simt::exchange(v, l - d >= 0 ? l - d : l); }
template <class T> static inline T __shfl_down_sync(unsigned, T v, int d) { const int l = threadIdx.x & 31; return simt::exchange(v, l + d < 32 ? l + d : l); }
template <class T> static inline T __shfl_xor_sync(unsigned, T v, int m) { return simt::exchange(v, (int)((threadIdx.x & 31) ^ (unsigned)m)); }
template <class T> static inline T __shfl_sync(unsigned, T v, int src) { return simt::exchange(v, src & 31); }
static inline int __reduce_max_sync(unsigned, int v) {
  for (int o = 16; o > 0; o >>= 1) { const int t = simt::exchange(v, (int)((threadIdx.x & 31) ^ (unsigned)o)); v = t > v ? t : v; }
  return v;
}
static inline void __syncthreads() { simt::block_barrier.arrive(); }
static inline int __syncthreads_or(int p) { return simt::block_barrier.arrive(p != 0); }
static inline void __syncwarp(unsigned = 0xffffffffu) { simt::warp_barrier[threadIdx.x >> 5].arrive(); }
static inline void __threadfence() { std::atomic_thread_fence(std::memory_order_seq_cst); }
static inline void __threadfence_system() { std::atomic_thread_fence(std::memory_order_seq_cst); }
static inline void __nanosleep(unsigned) { std::this_thread::yield(); }
template <class T> static inline T __ldg(const T* p) { return *p; }
template <class T> static inline T __ldcg(const T* p) { return *p; }
template <class T> static inline void __stcg(T* p, T v) { *p = v; }
template <class T> static inline T atomicAdd(T* p, T v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
template <class T> static inline T atomicMax(T* p, T v) {
  T o = __atomic_load_n(p, __ATOMIC_SEQ_CST);
  while (o < v && !__atomic_compare_exchange_n(p, &o, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
  return o;
}
template <class T> static inline T atomicMin(T* p, T v) {
  T o = __atomic_load_n(p, __ATOMIC_SEQ_CST);
  while (o > v && !__atomic_compare_exchange_n(p, &o, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
  return o;
}
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
static inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }
static inline long long min(long long a, long long b) { return a < b ? a : b; }
static inline long long max(long long a, long long b) { return a > b ? a : b; }

typedef void* cudaStream_t;
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaDevAttrMultiProcessorCount = 16 };
static inline int cudaGetLastError() { return 0; }
static inline int cudaGetDevice(int* d) { *d = 0; return 0; }
static inline int cudaDeviceGetAttribute(int* v, int, int) { *v = 1; return 0; }
static inline int cudaOccupancyMaxActiveBlocksPerMultiprocessor(int* o, const void*, int, int) { *o = 1; return 0; }

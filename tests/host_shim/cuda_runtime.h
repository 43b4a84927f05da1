// TEST INFRASTRUCTURE: a stand-in for <cuda_runtime.h> that lets g++ PARSE the device headers under
// genjax_b200/csrc and RUN their scalar, exact-arithmetic device functions on the host (Philox, u01, the ordered
// float encoding, det_exp_q, offspring_cnt, resample_u0), so that tests/test_device_functions_on_host.py can compare
// the CUDA SOURCE of those functions with the oracle bit for bit without a GPU.  Everything that needs a thread block
// (shuffles, barriers, atomics, shared memory) is declared only so that the templates parse; it must not be called.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

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
struct dim3 { unsigned x, y, z; };
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return uint4{x, y, z, w}; }
static inline int4 make_int4(int32_t x, int32_t y, int32_t z, int32_t w) { return int4{x, y, z, w}; }
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
static dim3 threadIdx, blockIdx, blockDim, gridDim;

// IEEE round-to-nearest single operations: compile this TU with -ffp-contract=off so that they stay unfused
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
static inline void sincospif(float x, float* s, float* c) { *s = sinf(3.14159265358979323846f * x); *c = cosf(3.14159265358979323846f * x); }

// parse-only: block-level primitives (never called from the host tests)
template <class T> static inline T __shfl_up_sync(unsigned, T v, int) { return v; }
template <class T> static inline T __shfl_xor_sync(unsigned, T v, int) { return v; }
template <class T> static inline T __shfl_sync(unsigned, T v, int) { return v; }
static inline int __reduce_max_sync(unsigned, int v) { return v; }
static inline void __syncthreads() {}
static inline int __syncthreads_or(int p) { return p; }
static inline void __threadfence() {}
static inline void __threadfence_system() {}
static inline void __nanosleep(unsigned) {}
template <class T> static inline T __ldg(const T* p) { return *p; }
template <class T> static inline T __ldcg(const T* p) { return *p; }
template <class T> static inline void __stcg(T* p, T v) { *p = v; }
template <class T> static inline T atomicAdd(T* p, T v) { T o = *p; *p += v; return o; }
template <class T> static inline T atomicMax(T* p, T v) { T o = *p; if (v > o) *p = v; return o; }
template <class T> static inline T atomicMin(T* p, T v) { T o = *p; if (v < o) *p = v; return o; }
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
static inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }
static inline long long min(long long a, long long b) { return a < b ? a : b; }
static inline long long max(long long a, long long b) { return a > b ? a : b; }

// parse-only: host runtime API mentioned by helpers in the headers
typedef void* cudaStream_t;
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaDevAttrMultiProcessorCount = 16 };
static inline int cudaGetDevice(int* d) { *d = 0; return 0; }
static inline int cudaDeviceGetAttribute(int* v, int, int) { *v = 1; return 0; }
static inline int cudaOccupancyMaxActiveBlocksPerMultiprocessor(int* o, const void*, int, int) { *o = 1; return 0; }
static inline int __float_as_int(float f) { int i; memcpy(&i, &f, 4); return i; }
static inline float __int_as_float(int i) { float f; memcpy(&f, &i, 4); return f; }

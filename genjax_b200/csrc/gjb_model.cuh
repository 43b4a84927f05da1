// Helpers shared by every generated model translation unit (see
// genjax_b200/gen/codegen.py).  A generated kernel is the batched, fused form
// of the reference's static-language handlers
// (generative_functions/static.py:254-278 simulate, :298-321 assess,
// :341-380 generate, :407-466 update, :616-673 regenerate): each thread walks
// the model's sites in program order for its particles, sampling or reading
// each value, and accumulates score / weight in registers.
//
// Layouts: scalar models process QUADS of 4 consecutive particles per thread
// with 128-bit loads/stores; vector models (event width D, D % 4 == 0) map
// G = D/4 lanes to one particle, each lane owning one float4 of every vector
// value, and reduce the per-lane logpdf partials with warp shuffles.
#pragma once
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "genjax_b200.h"
#include "gjb_dist.cuh"
#include "gjb_resample.cuh"
#include "gjb_rng.cuh"
#include "gjb_step.cuh"

namespace gjb {

__device__ __forceinline__ float as_f(uint32_t w) { return __uint_as_float(w); }
__device__ __forceinline__ uint32_t as_u(float f) { return __float_as_uint(f); }
__device__ __forceinline__ uint32_t as_u(int i) { return (uint32_t)i; }

// coherent (L2) vs read-only-path loads: kCg = true inside the persistent
// filter kernel, where the buffers were written earlier in the same launch
template <bool kCg, typename T>
__device__ __forceinline__ T ldx(const T* p) { return kCg ? __ldcg(p) : __ldg(p); }
template <bool kCg>
__device__ __forceinline__ float ldf(const float* p) { return kCg ? __ldcg(p) : __ldg(p); }

// base pointer and local row of a per-particle argument row that may live on a peer rank
__device__ __forceinline__ const void* arg_base(const void* local, const gjb_peers* P, int64_t& row) {
  if (!P) return local;
  const uint32_t owner = peer_owner(P, (uint32_t)row);
  row -= (int64_t)owner * P->n_per_rank;
  return P->base[owner];
}

// 4 consecutive 32-bit words starting at element i0 (elements u in [lo, hi)
// valid), optionally through a gather index per element or broadcast from
// element 0.
template <bool kCg>
__device__ __forceinline__ void load4(const void* __restrict__ base, int64_t i0, int lo, int hi, const int32_t (&g)[4],
                                      bool gathered, bool bcast, uint32_t (&w)[4], const gjb_peers* peers = nullptr) {
  const uint32_t* __restrict__ p = reinterpret_cast<const uint32_t*>(base);
  if (bcast) {
    const uint32_t v = __ldg(p);
    w[0] = w[1] = w[2] = w[3] = v;
  } else if (gathered && peers) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      int64_t row = g[u];
      const uint32_t* q = reinterpret_cast<const uint32_t*>(arg_base(base, peers, row));
      w[u] = (u >= lo && u < hi) ? ldx<kCg>(q + row) : 0u;
    }
  } else if (gathered) {
#pragma unroll
    for (int u = 0; u < 4; ++u) w[u] = (u >= lo && u < hi) ? ldx<kCg>(p + g[u]) : 0u;
  } else if (lo == 0 && hi == 4 && ((reinterpret_cast<uintptr_t>(p + i0) & 15) == 0)) {
    const uint4 v = ldx<kCg>(reinterpret_cast<const uint4*>(p + i0));
    w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
  } else {
#pragma unroll
    for (int u = 0; u < 4; ++u) w[u] = (u >= lo && u < hi) ? ldx<kCg>(p + i0 + u) : 0u;
  }
}

template <bool kCg>
__device__ __forceinline__ void load4_idx(const int32_t* __restrict__ gather, int64_t i0, int lo, int hi, int32_t (&g)[4]) {
  if (!gather) {
    g[0] = g[1] = g[2] = g[3] = 0;
    return;
  }
  if (lo == 0 && hi == 4 && ((reinterpret_cast<uintptr_t>(gather + i0) & 15) == 0)) {
    const int4 v = ldx<kCg>(reinterpret_cast<const int4*>(gather + i0));
    g[0] = v.x; g[1] = v.y; g[2] = v.z; g[3] = v.w;
  } else {
#pragma unroll
    for (int u = 0; u < 4; ++u) g[u] = (u >= lo && u < hi) ? ldx<kCg>(gather + i0 + u) : 0;
  }
}

// the same through a generic pointer (the single-launch filter step keeps the ancestors of its window in shared memory)
__device__ __forceinline__ void load4_idx_gen(const int32_t* gather, int64_t i0, int lo, int hi, int32_t (&g)[4]) {
  if (!gather) {
    g[0] = g[1] = g[2] = g[3] = 0;
    return;
  }
  if (lo == 0 && hi == 4 && ((reinterpret_cast<uintptr_t>(gather + i0) & 15) == 0)) {
    const int4 v = *reinterpret_cast<const int4*>(gather + i0);
    g[0] = v.x; g[1] = v.y; g[2] = v.z; g[3] = v.w;
  } else {
#pragma unroll
    for (int u = 0; u < 4; ++u) g[u] = (u >= lo && u < hi) ? gather[i0 + u] : 0;
  }
}

__device__ __forceinline__ void store4(void* __restrict__ base, int64_t i0, int lo, int hi, const uint32_t (&w)[4]) {
  uint32_t* __restrict__ p = reinterpret_cast<uint32_t*>(base);
  if (lo == 0 && hi == 4 && ((reinterpret_cast<uintptr_t>(p + i0) & 15) == 0)) {
    *reinterpret_cast<uint4*>(p + i0) = make_uint4(w[0], w[1], w[2], w[3]);
  } else {
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (u >= lo && u < hi) p[i0 + u] = w[u];
  }
}

// Grid-wide barrier of a persistent cooperative launch (all CTAs resident).
// Measured on B200 (scratch/barrier_bench.cu, 592 CTAs x 256 threads):
// cooperative_groups grid.sync 1.7 us, red.release + ld.acquire polling 2.2 us,
// __threadfence + atomicAdd + volatile polling 3.0 us -- so grid.sync it is.
__device__ __forceinline__ void grid_barrier(uint32_t* bar, uint32_t nblocks) {
  (void)bar; (void)nblocks;
  cooperative_groups::this_grid().sync();
}

// CTAs of `fn` that are co-resident on the current device (host side)
static inline int resident_blocks(const void* fn, int threads, int max_per_sm, int dyn_smem = 0) {
  int dev = 0, sms = 0, occ = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, threads, dyn_smem);
  if (occ > max_per_sm) occ = max_per_sm;
  if (occ < 1) occ = 1;
  if (sms < 1) sms = 148;
  return sms * occ;
}

// block-wide max of the per-thread running max -> atomicMax(wmax)
__device__ __forceinline__ void block_wmax(float m, uint32_t* wmax) {
  __shared__ float gjb_wmax_sm[32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) gjb_wmax_sm[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x < 32) {
    m = threadIdx.x < ((blockDim.x + 31) >> 5) ? gjb_wmax_sm[threadIdx.x] : -INFINITY;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (threadIdx.x == 0 && m > -INFINITY) atomicMax(wmax, fenc(m));
  }
}

// cooperative copy of a shared (un-batched) argument block into shared memory
__device__ __forceinline__ void stage_shared(float* dst, const void* src, int len) {
  const float* __restrict__ s = reinterpret_cast<const float*>(src);
  for (int k = threadIdx.x; k < len; k += blockDim.x) dst[k] = __ldg(s + k);
}

// ------------------------------------------------------------ scalar math
__device__ __forceinline__ float f_neg(float x) { return -x; }
__device__ __forceinline__ int f_neg(int x) { return -x; }
__device__ __forceinline__ float f_exp(float x) { return expf(x); }
__device__ __forceinline__ float f_log(float x) { return logf(x); }
__device__ __forceinline__ float f_sqrt(float x) { return sqrtf(x); }
__device__ __forceinline__ float f_abs(float x) { return fabsf(x); }
__device__ __forceinline__ int f_abs(int x) { return abs(x); }
__device__ __forceinline__ float f_tanh(float x) { return tanhf(x); }
__device__ __forceinline__ float f_sigmoid(float x) { return 1.0f / (1.0f + expf(-x)); }
__device__ __forceinline__ float f_log1p(float x) { return log1pf(x); }
__device__ __forceinline__ float f_expm1(float x) { return expm1f(x); }
__device__ __forceinline__ float f_square(float x) { return x * x; }
__device__ __forceinline__ float f_floor(float x) { return floorf(x); }
__device__ __forceinline__ float f_sin(float x) { return sinf(x); }
__device__ __forceinline__ float f_cos(float x) { return cosf(x); }
__device__ __forceinline__ float f_softplus(float x) { return softplusf(x); }
__device__ __forceinline__ float f_lgamma(float x) { return lgammaf(x); }
__device__ __forceinline__ float f_reciprocal(float x) { return 1.0f / x; }
__device__ __forceinline__ float f_pow(float a, float b) { return powf(a, b); }
__device__ __forceinline__ float f_min(float a, float b) { return fminf(a, b); }
__device__ __forceinline__ float f_max(float a, float b) { return fmaxf(a, b); }
__device__ __forceinline__ int f_min(int a, int b) { return min(a, b); }
__device__ __forceinline__ int f_max(int a, int b) { return max(a, b); }

// ------------------------------------------------------------ lane vectors
// V4: the 4 elements of a width-D vector value owned by one lane.
struct V4 {
  float v[4];
};
__device__ __forceinline__ V4 v4_splat(float s) { return V4{{s, s, s, s}}; }
__device__ __forceinline__ V4 v4_from(float4 f) { return V4{{f.x, f.y, f.z, f.w}}; }
__device__ __forceinline__ float4 v4_to(const V4& a) { return make_float4(a.v[0], a.v[1], a.v[2], a.v[3]); }
__device__ __forceinline__ V4 v4_load(const float* p) { return v4_from(*reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ V4 v4_ldg(const float* p) { return v4_from(__ldg(reinterpret_cast<const float4*>(p))); }
template <bool kCg>
__device__ __forceinline__ V4 v4_ld(const float* p) {
  return v4_from(kCg ? __ldcg(reinterpret_cast<const float4*>(p)) : __ldg(reinterpret_cast<const float4*>(p)));
}

#define GJB_V4_BIN(name, expr)                                                               \
  __device__ __forceinline__ V4 name(const V4& a, const V4& b) {                             \
    V4 r;                                                                                    \
    _Pragma("unroll") for (int k = 0; k < 4; ++k) { const float x = a.v[k], y = b.v[k]; r.v[k] = (expr); } \
    return r;                                                                                \
  }                                                                                          \
  __device__ __forceinline__ V4 name(const V4& a, float b) { return name(a, v4_splat(b)); }  \
  __device__ __forceinline__ V4 name(float a, const V4& b) { return name(v4_splat(a), b); }
GJB_V4_BIN(operator+, x + y)
GJB_V4_BIN(operator-, x - y)
GJB_V4_BIN(operator*, x * y)
GJB_V4_BIN(operator/, x / y)
GJB_V4_BIN(f_pow, powf(x, y))
GJB_V4_BIN(f_min, fminf(x, y))
GJB_V4_BIN(f_max, fmaxf(x, y))
#undef GJB_V4_BIN

#define GJB_V4_UN(name, expr)                                                     \
  __device__ __forceinline__ V4 name(const V4& a) {                               \
    V4 r;                                                                         \
    _Pragma("unroll") for (int k = 0; k < 4; ++k) { const float x = a.v[k]; r.v[k] = (expr); } \
    return r;                                                                     \
  }
GJB_V4_UN(f_neg, -x)
GJB_V4_UN(f_exp, expf(x))
GJB_V4_UN(f_log, logf(x))
GJB_V4_UN(f_sqrt, sqrtf(x))
GJB_V4_UN(f_abs, fabsf(x))
GJB_V4_UN(f_tanh, tanhf(x))
GJB_V4_UN(f_sigmoid, 1.0f / (1.0f + expf(-x)))
GJB_V4_UN(f_log1p, log1pf(x))
GJB_V4_UN(f_expm1, expm1f(x))
GJB_V4_UN(f_square, x * x)
GJB_V4_UN(f_floor, floorf(x))
GJB_V4_UN(f_sin, sinf(x))
GJB_V4_UN(f_cos, cosf(x))
GJB_V4_UN(f_softplus, softplusf(x))
GJB_V4_UN(f_lgamma, lgammaf(x))
GJB_V4_UN(f_reciprocal, 1.0f / x)
#undef GJB_V4_UN

__device__ __forceinline__ float v4_hsum(const V4& a) { return ((a.v[0] + a.v[1]) + a.v[2]) + a.v[3]; }

// sum over the G lanes of a particle group (G power of two, groups warp aligned)
template <int G>
__device__ __forceinline__ float group_sum(float x) {
#pragma unroll
  for (int o = 1; o < G; o <<= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
  return x;
}

// mv_normal_diag on one lane's float4: sample + per-lane logpdf partial
__device__ __forceinline__ V4 mvn_diag_sample(const Lane& l, uint32_t site, uint32_t chunk, const V4& loc,
                                              const V4& scale) {
  const float4 z = normal4(l, site, chunk);
  return V4{{__fmaf_rn(scale.v[0], z.x, loc.v[0]), __fmaf_rn(scale.v[1], z.y, loc.v[1]), __fmaf_rn(scale.v[2], z.z, loc.v[2]),
             __fmaf_rn(scale.v[3], z.w, loc.v[3])}};
}
__device__ __forceinline__ float mvn_diag_logpdf4(const V4& v, const V4& loc, const V4& scale) {
  float s = 0.0f;
#pragma unroll
  for (int k = 0; k < 4; ++k) s += Normal::logpdf(v.v[k], loc.v[k], scale.v[k]);
  return s;
}

// same with the particle-invariant pieces precomputed: inv = 1/scale (this
// lane's 4 elements), lc = sum over those elements of 0.5 log 2pi + log scale
__device__ __forceinline__ float mvn_diag_logpdf4_r(const V4& v, const V4& loc, const V4& inv, float lc) {
  float s = -lc;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float z = v.v[k] * inv.v[k] - loc.v[k] * inv.v[k];
    s -= 0.5f * (z * z);
  }
  return s;
}

}  // namespace gjb

// Counter-based RNG for the fused particle kernels (sm_100a).
//
// Replaces jax.random threefry keys folded per site at
// generative_functions/static.py:260-263 and the TFP samplers behind
// distributions/tensorflow_probability/__init__.py:52-62.  Stream layout
// (restated on the CPU in oracle/rng.py):
//   words = philox4x32_10(ctr = (idx_lo, idx_hi, chunk, site), key = (k0, k1))
// Everything lives in registers; one Philox block costs ~60 integer ops.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace gjb {

__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c.x);
    const uint32_t lo0 = 0xD2511F53u * c.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z);
    const uint32_t lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k0, lo1, hi0 ^ c.w ^ k1, lo0);
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return c;
}

struct Lane {
  uint32_t k0, k1, lo, hi;  // batch key + global particle index
  __device__ __forceinline__ uint4 words(uint32_t site, uint32_t chunk) const {
    return philox4x32_10(make_uint4(lo, hi, chunk, site), k0, k1);
  }
};

__device__ __forceinline__ Lane make_lane(uint32_t k0, uint32_t k1, uint64_t idx) {
  Lane l;
  l.k0 = k0;
  l.k1 = k1;
  l.lo = (uint32_t)idx;
  l.hi = (uint32_t)(idx >> 32);
  return l;
}

// uint32 -> float in (0,1): ((bits >> 9) + 0.5) * 2^-23, exact in fp32.
__device__ __forceinline__ float u01(uint32_t bits) {
  return ((float)(bits >> 9) + 0.5f) * 1.1920928955078125e-07f;
}

// Two N(0,1) from two words: r = sqrt(-2 log u1); (s,c) = sincospi(2 u2).
__device__ __forceinline__ float2 box_muller(uint32_t b0, uint32_t b1) {
  const float u1 = u01(b0);
  const float u2 = u01(b1);
  const float r = sqrtf(-2.0f * logf(u1));
  float s, c;
  sincospif(2.0f * u2, &s, &c);
  return make_float2(r * c, r * s);
}

__device__ __forceinline__ float normal1(const Lane& l, uint32_t site, uint32_t chunk = 0) {
  const uint4 w = l.words(site, chunk);
  return box_muller(w.x, w.y).x;
}

__device__ __forceinline__ float4 normal4(const Lane& l, uint32_t site, uint32_t chunk) {
  const uint4 w = l.words(site, chunk);
  const float2 a = box_muller(w.x, w.y);
  const float2 b = box_muller(w.z, w.w);
  return make_float4(a.x, a.y, b.x, b.y);
}

// ---- quad streams: SCALAR sites share one Philox block between the 4
// consecutive particles of a global quad (idx >> 2); particle idx takes slot
// idx & 3 (word for uniform-driven samplers, normal4 component for normals).
// Vector sites and rejection samplers keep one stream per particle (Lane).
__device__ __forceinline__ uint4 quad_words(uint32_t k0, uint32_t k1, uint64_t quad, uint32_t site, uint32_t chunk = 0) {
  return philox4x32_10(make_uint4((uint32_t)quad, (uint32_t)(quad >> 32), chunk, site), k0, k1);
}
__device__ __forceinline__ float4 normal4_of(const uint4& w) {
  const float2 a = box_muller(w.x, w.y);
  const float2 b = box_muller(w.z, w.w);
  return make_float4(a.x, a.y, b.x, b.y);
}
__device__ __forceinline__ uint32_t pick(const uint4& w, int s) { return s == 0 ? w.x : (s == 1 ? w.y : (s == 2 ? w.z : w.w)); }
__device__ __forceinline__ float pick(const float4& w, int s) { return s == 0 ? w.x : (s == 1 ? w.y : (s == 2 ? w.z : w.w)); }

// ordered-uint encoding of a float for atomicMax
__device__ __forceinline__ uint32_t fenc(float f) {
  const uint32_t b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float fdec(uint32_t e) {
  return __uint_as_float((e & 0x80000000u) ? (e & 0x7FFFFFFFu) : ~e);
}
#define GJB_WMAX_NEG_INF 0x007FFFFFu  // fenc(-inf)

}  // namespace gjb

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

// threefry2x32-20 (Salmon et al. 2011): the HOST key tree (core/key.py split / fold_in) restated for the device, used
// only to derive a filter run's per-step key table on the device (gjb_pf_key_table) instead of shipping it from the host.
__device__ __forceinline__ uint2 threefry2x32_20(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1) {
  const uint32_t ks[3] = {k0, k1, 0x1BD11BDAu ^ k0 ^ k1};
  const int rot[2][4] = {{13, 15, 26, 6}, {17, 29, 16, 24}};
  uint32_t x0 = c0 + ks[0], x1 = c1 + ks[1];
#pragma unroll
  for (int g = 0; g < 5; ++g) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      x0 += x1;
      x1 = ((x1 << rot[g & 1][r]) | (x1 >> (32 - rot[g & 1][r]))) ^ x0;
    }
    x0 += ks[(g + 1) % 3];
    x1 += ks[(g + 2) % 3] + (uint32_t)(g + 1);
  }
  return make_uint2(x0, x1);
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

// uint32 -> float in (0,1): ((bits >> 9) + 0.5) * 2^-23, exact in fp32.  Formed without the conversion pipe:
// as_float(0x3F800000 | k) = 1 + k 2^-23, minus (1 - 2^-24) -- an exact subtraction giving the same value.
__device__ __forceinline__ float u01(uint32_t bits) {
  return __fadd_rn(__uint_as_float(0x3F800000u | (bits >> 9)), -0x1.fffffep-1f);
}

// Two N(0,1) from two words: r = sqrt(-2 ln u1), (z0, z1) = r (cos, sin)(2 pi u2), on fp32 polynomials evaluated with
// correctly rounded FMAs -- every line is ONE IEEE operation restated in oracle/rng.py box_muller, so the sampled normals
// are bit-exact against the CPU oracle (coefficients: scratch/fit_box_muller.py; |error| of each kernel < 3e-8).
__device__ __forceinline__ float2 box_muller(uint32_t b0, uint32_t b1) {
  const float u1 = u01(b0);
  const float u2 = u01(b1);
  // -2 ln u1: u1 = 2^e m, m in [sqrt(1/2), sqrt(2)), f = m - 1 (exact); -2 log1p(f) = -2 f + f^2 R(f)
  const int tb = __float_as_int(u1) - 0x3F3504F3;
  const int e = tb >> 23;
  const float f = __fadd_rn(__int_as_float((tb & 0x007FFFFF) + 0x3F3504F3), -1.0f);
  const float ef = __fadd_rn(__int_as_float(e + 0x4B400000), -12582912.0f);  // (float)e without the conversion pipe
  const float f2 = __fmul_rn(f, f);
  float R = 0x1.4237fep-3f;
  R = __fmaf_rn(R, f, -0x1.0696e4p-2f);
  R = __fmaf_rn(R, f, 0x1.0c524cp-2f);
  R = __fmaf_rn(R, f, -0x1.22973ap-2f);
  R = __fmaf_rn(R, f, 0x1.548882p-2f);
  R = __fmaf_rn(R, f, -0x1.99a3ecp-2f);
  R = __fmaf_rn(R, f, 0x1.000206p-1f);
  R = __fmaf_rn(R, f, -0x1.55554ep-1f);
  R = __fmaf_rn(R, f, 0x1.fffffep-1f);
  const float L = __fmaf_rn(ef, -0x1.62e43p+0f, __fmaf_rn(f2, R, __fmul_rn(-2.0f, f)));
  const float r = __fsqrt_rn(L);
  // 2 pi u2 = j pi/2 + 2 pi rr: j = rint(4 u2) sits in the low mantissa bits of tm, rr in [-1/8, 1/8] is exact
  const float tm = __fmaf_rn(u2, 4.0f, 12582912.0f);
  const int q = __float_as_int(tm) & 3;
  const float rr = __fmaf_rn(__fadd_rn(tm, -12582912.0f), -0.25f, u2);
  const float z = __fmul_rn(rr, rr);
  float ps = __fmaf_rn(z, -0x1.2d9b7cp+6f, 0x1.465ec4p+6f);
  ps = __fmaf_rn(z, ps, -0x1.4abbbap+5f);
  ps = __fmaf_rn(z, ps, 0x1.921fb6p+2f);
  const float sn = __fmul_rn(rr, ps);
  float pc = __fmaf_rn(z, 0x1.db6578p+5f, -0x1.55cb9ap+6f);
  pc = __fmaf_rn(z, pc, 0x1.03c1eap+6f);
  pc = __fmaf_rn(z, pc, -0x1.3bd3ccp+4f);
  const float cs = __fmaf_rn(z, pc, 1.0f);
  const bool swap = (q & 1) != 0;
  const float a = swap ? sn : cs, b = swap ? cs : sn;
  const float ca = __int_as_float(__float_as_int(a) ^ (((q + 1) & 2) << 30));  // quadrants 1, 2: cos < 0
  const float sb = __int_as_float(__float_as_int(b) ^ ((q & 2) << 30));        // quadrants 2, 3: sin < 0
  return make_float2(__fmul_rn(r, ca), __fmul_rn(r, sb));
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

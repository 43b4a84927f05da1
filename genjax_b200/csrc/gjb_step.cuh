// Tile-exponent masses and output-slot ("pull") systematic resampling: the device routines of the filter step that
// is ONE launch (gen/codegen.py pf_step_kernel, include/genjax_b200.h sections 1c and 2) and of the stand-alone
// gjb_te_masses / gjb_te_resample kernels in gjb_core.cu.  CPU restatement: oracle/smc.py (te_exp2_m,
// te_tile_masses, te_cdf, resample_systematic_te).  Replaces the logsumexp + categorical-per-offspring idiom of the
// reference (inference/smc.py:96-109; docs/cookbook/inactive/inference/mapping_tutorial.ipynb cell 37).
//
//   producer (te_publish), per tile of 2048 particles -- block-level synchronisation only:
//     t_i = fl32(lw_i * log2e);  e = ceil(max_tile t);  q_i = round(2^36 * 2^(t_i - e));
//     cdf[i] = inclusive prefix of q inside the tile (uint64);  rec = {cdf[last], e}
//   consumer (te_pull), per CTA = per window of <= 2048 offspring slots:
//     E = max e_p;  s_p = min(E - e_p, 63);  P = inclusive prefix of (mass_p >> s_p);  S = P_last
//     C_i = P_{p-1} + (cdf[i] >> s_p);  cnt_i = offspring_cnt(C_i);  slot j <- first i with cnt_i > j
//   (every parent with offspring in the window drops its id at its first slot, one block-wide max-scan fills the rest)
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "genjax_b200.h"
#include "gjb_resample.cuh"
#include "gjb_rng.cuh"

namespace gjb {

constexpr int kTeTile = GJB_TE_TILE;
constexpr int kTeMaxTiles = GJB_TE_MAX_TILES;
constexpr int kTeItems = kTeTile / kThreads;  // 8 consecutive particles / slots per thread
static_assert(kTeItems == 8 && kTeTile == kTile, "tile shape");

struct TeSmem {
  uint64_t pre[kTeMaxTiles];               // inclusive prefix of the aligned tile masses
  __align__(16) int32_t heads[kTeTile];    // window slots: parent id + 1 at first slots, then ancestors, then weights
  uint8_t shf[kTeMaxTiles];                // s_p
  uint64_t red[kThreads / 32];
  int32_t ired[kThreads / 32];
  float fred[kThreads / 32];
  int32_t p_lo, p_hi;
};

// uint64(fl32(2^36 * 2^frac)), frac in [0, 1): fp32 FMA Horner (oracle/smc.py te_exp2_m)
__device__ __forceinline__ uint64_t te_exp2_m(float frac) {
  const float g = __fadd_rn(frac, -0.5f);
  float p = 0x1.ffcbfcp-17f;               // ln2^7/7!
  p = __fmaf_rn(p, g, 0x1.430912p-13f);    // ln2^6/6!
  p = __fmaf_rn(p, g, 0x1.5d87fep-10f);    // ln2^5/5!
  p = __fmaf_rn(p, g, 0x1.3b2ab6p-7f);     // ln2^4/4!
  p = __fmaf_rn(p, g, 0x1.c6b08ep-5f);     // ln2^3/3!
  p = __fmaf_rn(p, g, 0x1.ebfbep-3f);      // ln2^2/2!
  p = __fmaf_rn(p, g, 0x1.62e43p-1f);      // ln2
  p = __fmaf_rn(p, g, 1.0f);
  p = __fmul_rn(p, 0x1.6a09e6p+0f);        // sqrt(2)
  return (uint64_t)__fmul_rn(p, 68719476736.0f);  // 2^36; p in [1, 2] so the product is an exact integer
}

// t = fl32(lw * log2e) clamped to +-2^29, or -inf when lw is not finite (NaN, +-inf: mass 0)
__device__ __forceinline__ float te_t(float lw) {
  const float t = __fmul_rn(lw, 0x1.715476p+0f);
  return (fabsf(t) < INFINITY) ? fminf(fmaxf(t, -536870912.0f), 536870912.0f) : -INFINITY;
}

// mass of a particle with scaled log-weight t (te_t) relative to the tile exponent e >= t
__device__ __forceinline__ uint64_t te_q(float t, int e) {
  if (!(t > -INFINITY)) return 0ull;
  const float nf = floorf(t);
  const uint64_t m = te_exp2_m(__fadd_rn(t, -nf));
  const int sh = min(e - (int)nf, 63);
  return sh > 0 ? ((m + (1ull << (sh - 1))) >> sh) : m;
}

// Producer: the 8 consecutive weights of this thread (slot base tid * 8 of the tile; lw = -inf past the end) ->
// cdf_tile[tid * 8 + k] (2048 entries, padding repeats the total) and *rec.  All kThreads threads call.
__device__ __forceinline__ void te_publish(const float (&lw)[kTeItems], uint64_t* __restrict__ cdf_tile,
                                           gjb_tile_rec* __restrict__ rec, TeSmem& sm) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float t[kTeItems];
  float tm = -INFINITY;
#pragma unroll
  for (int k = 0; k < kTeItems; ++k) { t[k] = te_t(lw[k]); tm = fmaxf(tm, t[k]); }
  tm = warp_max(tm);
  if (lane == 0) sm.fred[warp] = tm;
  __syncthreads();
  tm = sm.fred[0];
#pragma unroll
  for (int w = 1; w < kThreads / 32; ++w) tm = fmaxf(tm, sm.fred[w]);
  const int e = tm > -INFINITY ? __float2int_ru(tm) : GJB_TE_E_NONE;
  uint64_t c[kTeItems];
  uint64_t run = 0;
#pragma unroll
  for (int k = 0; k < kTeItems; ++k) { run += te_q(t[k], e); c[k] = run; }
  uint64_t inc = run;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint64_t v = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += v;
  }
  if (lane == 31) sm.red[warp] = inc;
  __syncthreads();
  uint64_t excl = inc - run;
#pragma unroll
  for (int w = 0; w < kThreads / 32; ++w) if (w < warp) excl += sm.red[w];
  ulonglong2* dst = reinterpret_cast<ulonglong2*>(cdf_tile + tid * kTeItems);
#pragma unroll
  for (int k = 0; k < kTeItems; k += 2) dst[k >> 1] = make_ulonglong2(c[k] + excl, c[k + 1] + excl);
  if (tid == kThreads - 1) {
    // one 16-byte store: a consumer never sees a torn record
    *reinterpret_cast<uint4*>(rec) = make_uint4((uint32_t)(c[kTeItems - 1] + excl), (uint32_t)((c[kTeItems - 1] + excl) >> 32),
                                                (uint32_t)e, 0u);
  }
}

template <bool kCg>
__device__ __forceinline__ gjb_tile_rec te_ld_rec(const gjb_tile_rec* p) {
  const uint4 v = kCg ? __ldcg(reinterpret_cast<const uint4*>(p)) : __ldg(reinterpret_cast<const uint4*>(p));
  gjb_tile_rec r;
  r.mass = (uint64_t)v.x | ((uint64_t)v.y << 32);
  r.e = (int32_t)v.z;
  r.reserved = 0;
  return r;
}

// Consumer: global parent ids of the offspring slots [w_lo, w_lo + w_n) (w_n <= 2048) of this CTA, in blocked layout
// anc[k] = parent of slot w_lo + tid * 8 + k (junk past w_n).  Returns S (0: no weight has mass, identity written);
// *e_out = E.  `cdf` is this device's array, `cdf_peers` (nullable) every rank's, tile p living on rank p /
// (n_per_rank / 2048).  All kThreads threads call; sm.heads is left in use (ancestors are NOT stored there).
template <bool kCg>
__device__ __forceinline__ uint64_t te_pull(const gjb_tile_rec* __restrict__ recs, int n_tiles,
                                            const uint64_t* __restrict__ cdf, const gjb_peers* cdf_peers, int64_t n_total,
                                            double u0, int64_t w_lo, int w_n, TeSmem& sm, int32_t (&anc)[kTeItems], int* e_out) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int per = (n_tiles + kThreads - 1) / kThreads;
  const int t0 = tid * per;
  // ---- E = max exponent over tiles with mass
  int emax = GJB_TE_E_NONE;
  for (int k = 0; k < per; ++k) {
    const int t = t0 + k;
    if (t < n_tiles) {
      const gjb_tile_rec r = te_ld_rec<kCg>(recs + t);
      if (r.mass) emax = max(emax, r.e);
    }
  }
  emax = __reduce_max_sync(0xffffffffu, emax);
  if (lane == 0) sm.ired[warp] = emax;
  if (tid == 0) { sm.p_lo = 0x7fffffff; sm.p_hi = -1; }
  __syncthreads();
  int E = sm.ired[0];
#pragma unroll
  for (int w = 1; w < kThreads / 32; ++w) E = max(E, sm.ired[w]);
  *e_out = E;
  // ---- aligned tile masses, their inclusive prefix
  uint64_t run = 0;
  for (int k = 0; k < per; ++k) {
    const int t = t0 + k;
    if (t < n_tiles) {
      const gjb_tile_rec r = te_ld_rec<kCg>(recs + t);
      const int s = r.mass ? min(E - r.e, 63) : 63;
      run += r.mass >> s;
      sm.pre[t] = run;
      sm.shf[t] = (uint8_t)s;
    }
  }
  uint64_t inc = run;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint64_t v = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += v;
  }
  if (lane == 31) sm.red[warp] = inc;
  // clear this thread's window slots while the scan settles
  *reinterpret_cast<int4*>(sm.heads + tid * kTeItems) = make_int4(0, 0, 0, 0);
  *reinterpret_cast<int4*>(sm.heads + tid * kTeItems + 4) = make_int4(0, 0, 0, 0);
  __syncthreads();
  uint64_t excl = inc - run, S = 0;
#pragma unroll
  for (int w = 0; w < kThreads / 32; ++w) {
    const uint64_t v = sm.red[w];
    if (w < warp) excl += v;
    S += v;
  }
  if (S == 0) {
#pragma unroll
    for (int k = 0; k < kTeItems; ++k) anc[k] = (int32_t)(w_lo + tid * kTeItems + k);
    return 0;
  }
  const double scale = __ddiv_rn((double)n_total, (double)S);
  const int32_t nt = (int32_t)n_total;
  const int32_t wl = (int32_t)w_lo, wh = (int32_t)(w_lo + w_n);
  // ---- which parent tiles have offspring in the window (a contiguous range)
  {
    int lo = 0x7fffffff, hi = -1;
    uint64_t prev = excl;
    for (int k = 0; k < per; ++k) {
      const int t = t0 + k;
      if (t < n_tiles) {
        const uint64_t cur = sm.pre[t] + excl;
        sm.pre[t] = cur;
        if (cur != prev && offspring_cnt(cur, S, scale, u0, nt) > wl && offspring_cnt(prev, S, scale, u0, nt) < wh) {
          lo = min(lo, t);
          hi = max(hi, t);
        }
        prev = cur;
      }
    }
    if (hi >= 0) { atomicMin(&sm.p_lo, lo); atomicMax(&sm.p_hi, hi); }
  }
  __syncthreads();
  const int p_lo = sm.p_lo, p_hi = sm.p_hi;
  // ---- every parent with offspring in the window drops its id (+1) at its first slot
  const int tiles_per_rank = cdf_peers ? (int)(cdf_peers->n_per_rank / kTeTile) : 0;
  for (int p = p_lo; p <= p_hi; ++p) {
    const uint64_t base = p ? sm.pre[p - 1] : 0ull;
    if (sm.pre[p] == base) continue;  // a tile without (aligned) mass
    const int s = sm.shf[p];
    const uint64_t* ct;
    if (cdf_peers) {
      const int owner = p / tiles_per_rank;
      ct = reinterpret_cast<const uint64_t*>(cdf_peers->base[owner]) + (int64_t)(p - owner * tiles_per_rank) * kTeTile;
    } else {
      ct = cdf + (int64_t)p * kTeTile;
    }
    ct += tid * kTeItems;
    const uint64_t c_prev = tid ? (kCg ? __ldcg(ct - 1) : __ldg(ct - 1)) : 0ull;
    const ulonglong2* c2 = reinterpret_cast<const ulonglong2*>(ct);
    const ulonglong2 c67 = kCg ? __ldcg(c2 + 3) : __ldg(c2 + 3);
    int32_t prev = min(max(offspring_cnt(base + (c_prev >> s), S, scale, u0, nt), wl), wh);
    const int32_t last = min(max(offspring_cnt(base + (c67.y >> s), S, scale, u0, nt), wl), wh);
    if (last > prev) {  // this thread's 8 parents own slots of the window
      uint64_t c[kTeItems];
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const ulonglong2 v = kCg ? __ldcg(c2 + k) : __ldg(c2 + k);
        c[2 * k] = v.x; c[2 * k + 1] = v.y;
      }
      c[6] = c67.x; c[7] = c67.y;
      const int32_t id1 = p * kTeTile + tid * kTeItems + 1;
#pragma unroll
      for (int k = 0; k < kTeItems; ++k) {
        const int32_t cur = (k == kTeItems - 1) ? last : min(max(offspring_cnt(base + (c[k] >> s), S, scale, u0, nt), wl), wh);
        if (cur > prev) sm.heads[prev - wl] = id1 + k;
        prev = cur;
      }
    }
  }
  __syncthreads();
  // ---- inclusive max-scan over the window: 8 consecutive slots per thread, warp shuffle, block
  int32_t v[kTeItems];
  {
    const int4 a = *reinterpret_cast<const int4*>(sm.heads + tid * kTeItems);
    const int4 b = *reinterpret_cast<const int4*>(sm.heads + tid * kTeItems + 4);
    v[0] = a.x; v[1] = max(v[0], a.y); v[2] = max(v[1], a.z); v[3] = max(v[2], a.w);
    v[4] = max(v[3], b.x); v[5] = max(v[4], b.y); v[6] = max(v[5], b.z); v[7] = max(v[6], b.w);
  }
  int32_t incm = v[kTeItems - 1];
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int32_t t = __shfl_up_sync(0xffffffffu, incm, o);
    if (lane >= o) incm = max(incm, t);
  }
  if (lane == 31) sm.ired[warp] = incm;
  const int32_t wexc = __shfl_up_sync(0xffffffffu, incm, 1);
  __syncthreads();
  int32_t pre = lane ? wexc : 0;
#pragma unroll
  for (int w = 0; w < kThreads / 32; ++w) if (w < warp) pre = max(pre, sm.ired[w]);
#pragma unroll
  for (int k = 0; k < kTeItems; ++k) anc[k] = max(v[k], pre) - 1;
  return S;
}

// {E ln 2, S, log-mean-exp} of a resampling (one thread)
__device__ __forceinline__ void te_write_lse(double* out, int E, uint64_t S, int64_t n_total) {
  out[0] = S ? (double)E * 0.693147180559945309417 : -INFINITY;
  out[1] = (double)S;
  out[2] = S ? (double)E * 0.693147180559945309417 + log((double)S) - kQLog - log((double)n_total) : -INFINITY;
}

}  // namespace gjb

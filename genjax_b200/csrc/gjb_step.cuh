// Tile-exponent masses and output-slot ("pull") systematic resampling: the device routines of the filter step that
// is ONE launch (gen/codegen.py pf_step_kernel, include/genjax_b200.h sections 1c and 2) and of the stand-alone
// gjb_te_masses / gjb_te_resample kernels in gjb_core.cu.  CPU restatement: oracle/smc.py (te_exp2_m,
// te_tile_masses, te_cdf, resample_systematic_te).  Replaces the logsumexp + categorical-per-offspring idiom of the
// reference (inference/smc.py:96-109; docs/cookbook/inactive/inference/mapping_tutorial.ipynb cell 37).
//
//   producer (te_publish), per tile of 2048 particles -- block-level synchronisation only:
//     t_i = fl32(lw_i * log2e);  e = ceil(max_tile t);  q_i = rint(2^36 * 2^(t_i - e))  (te_q);
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

// Phase timestamps of the step kernel for scratch/trace_step.py (compiled in only with -DGJB_TRACE: build.py adds
// $GJB_NVCC_EXTRA to the nvcc line).  Slot i of CTA b: SM clock at trace point i; slots 14 / 15: globaltimer at entry / exit.
#ifdef GJB_TRACE
__device__ unsigned long long gjb_trace_buf[1024 * 16];
__device__ __forceinline__ unsigned long long gjb_globaltimer() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
// slots 7 / 13 (globaltimer too): written by the rank's LAST CTA only -- all records arrived / table written; cleared at entry
#define GJB_TP(i) do { if (threadIdx.x == 0 && blockIdx.x < 1024) { \
    if ((i) == 0) { gjb_trace_buf[blockIdx.x * 16 + 7] = 0ull; gjb_trace_buf[blockIdx.x * 16 + 13] = 0ull; } \
    gjb_trace_buf[blockIdx.x * 16 + (i)] = ((i) >= 14 || (i) == 7 || (i) == 13) ? gjb_globaltimer() : (unsigned long long)clock64(); } } while (0)
#else
#define GJB_TP(i) do {} while (0)
#endif

namespace gjb {

constexpr int kTeTile = GJB_TE_TILE;
constexpr int kTeMaxTiles = GJB_TE_MAX_TILES;
constexpr int kTeItems = kTeTile / kThreads;  // 8 consecutive particles / slots per thread
static_assert(kTeItems == 8 && kTeTile == kTile, "tile shape");

struct TeSmem {
  uint64_t pre[kTeMaxTiles];               // inclusive prefix of the aligned tile masses
  __align__(16) int32_t heads[kTeTile];    // window slots: parent id + 1 at first slots, then ancestors, then weights
  uint8_t shf[kTeMaxTiles];                // s_p
  __align__(16) uint64_t red[kThreads / 32];
  __align__(16) int32_t ired[kThreads / 32];
  __align__(16) float fred[kThreads / 32];
  int32_t p_lo, p_hi;
  uint64_t tile_mass;   // te_publish -> te_finish_step
  int32_t tile_e;
  int32_t is_last;
};

// ---- small PTX helpers (plain C++ under the host shims of tests/)
__device__ __forceinline__ uint64_t te_ld_volatile(const uint64_t* p) {
#if defined(__CUDACC__)
  unsigned long long v;
  asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
#else
  return *reinterpret_cast<const volatile uint64_t*>(p);
#endif
}
__device__ __forceinline__ void te_st_volatile(uint64_t* p, uint64_t v) {
#if defined(__CUDACC__)
  asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"((unsigned long long)v) : "memory");
#else
  *reinterpret_cast<volatile uint64_t*>(p) = v;
#endif
}
// programmatic dependent launch (sm_90+): let the next launch on the stream start early / wait for the previous one
__device__ __forceinline__ void pdl_launch_dependents() {
#if defined(__CUDACC__)
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
}
__device__ __forceinline__ void pdl_wait() {
#if defined(__CUDACC__)
  asm volatile("griddepcontrol.wait;" ::: "memory");
#endif
}

// t = fl32(lw * log2e) clamped to +-2^20, or -inf when lw is not finite (NaN, +-inf: mass 0)
__device__ __forceinline__ float te_t(float lw) {
  const float t = __fmul_rn(lw, 0x1.715476p+0f);
  return (fabsf(t) < INFINITY) ? fminf(fmaxf(t, -1048576.0f), 1048576.0f) : -INFINITY;
}

// Mass of a particle with scaled log-weight t (te_t) relative to the tile exponent e >= rint(t), kc = 163 - e - 0x4B400000:
// rint(2^g * 2^(36 - (e - r))), r = rint(t) taken from the low mantissa bits of t + 1.5 * 2^23, g = t - r, 2^g by fp32 FMA
// Horner (Taylor, degree 7).  No conversion-pipe instruction except the final float -> uint64.  oracle/smc.py te_q.
__device__ __forceinline__ uint64_t te_q(float t, int kc) {
  const bool ok = t > -INFINITY;
  const float tv = ok ? t : 0.0f;
  const float tm = __fadd_rn(tv, 12582912.0f);
  const float g = __fadd_rn(tv, -__fadd_rn(tm, -12582912.0f));
  float p = 0x1.ffcbfcp-17f;               // ln2^7/7!
  p = __fmaf_rn(p, g, 0x1.430912p-13f);    // ln2^6/6!
  p = __fmaf_rn(p, g, 0x1.5d87fep-10f);    // ln2^5/5!
  p = __fmaf_rn(p, g, 0x1.3b2ab6p-7f);     // ln2^4/4!
  p = __fmaf_rn(p, g, 0x1.c6b08ep-5f);     // ln2^3/3!
  p = __fmaf_rn(p, g, 0x1.ebfbep-3f);      // ln2^2/2!
  p = __fmaf_rn(p, g, 0x1.62e43p-1f);      // ln2
  p = __fmaf_rn(p, g, 1.0f);
  const int ex = max(__float_as_int(tm) + kc, 63);  // biased exponent of 2^(36 - (e - r)), floor 2^-64 (mass 0)
  const float scale = ok ? __int_as_float(ex << 23) : 0.0f;
  return __float2ull_rn(__fmul_rn(p, scale));
}

// Producer: the 8 consecutive weights of this thread (slot base tid * 8 of the tile; lw = -inf past the end) ->
// cdf_tile[tid * 8 + k] (2048 entries, padding repeats the total) and *rec.  All kThreads threads call.
__device__ __forceinline__ void te_publish(const float (&lw)[kTeItems], uint64_t* __restrict__ cdf_tile,
                                           gjb_tile_rec* __restrict__ rec /* nullable */, TeSmem& sm) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float t[kTeItems];
  float tm = -INFINITY;
#pragma unroll
  for (int k = 0; k < kTeItems; ++k) { t[k] = te_t(lw[k]); tm = fmaxf(tm, t[k]); }
  tm = warp_max(tm);
  if (lane == 0) sm.fred[warp] = tm;
  __syncthreads();
  {
    const float4 a = *reinterpret_cast<const float4*>(sm.fred), b = *reinterpret_cast<const float4*>(sm.fred + 4);
    tm = fmaxf(fmaxf(fmaxf(a.x, a.y), fmaxf(a.z, a.w)), fmaxf(fmaxf(b.x, b.y), fmaxf(b.z, b.w)));
  }
  GJB_TP(9);
  const int e = tm > -INFINITY ? __float2int_ru(tm) : GJB_TE_E_NONE;
  const int kc = tm > -INFINITY ? 163 - e - 0x4B400000 : 0;
  uint64_t c[kTeItems];
  uint64_t run = 0;
#pragma unroll
  for (int k = 0; k < kTeItems; ++k) { run += te_q(t[k], kc); c[k] = run; }
  uint64_t inc = run;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint64_t v = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += v;
  }
  if (lane == 31) sm.red[warp] = inc;
  __syncthreads();
  GJB_TP(10);
  uint64_t excl = inc - run;
#pragma unroll
  for (int w = 0; w < kThreads / 32; w += 2) {
    const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(sm.red + w);
    if (w < warp) excl += v.x;
    if (w + 1 < warp) excl += v.y;
  }
  ulonglong2* dst = reinterpret_cast<ulonglong2*>(cdf_tile + tid * kTeItems);
#pragma unroll
  for (int k = 0; k < kTeItems; k += 2) dst[k >> 1] = make_ulonglong2(c[k] + excl, c[k + 1] + excl);
  if (tid == kThreads - 1) {
    sm.tile_mass = c[kTeItems - 1] + excl;  // read by te_finish_step after its barrier
    sm.tile_e = e;
    // one 16-byte store: a consumer never sees a torn record
    if (rec) *reinterpret_cast<uint4*>(rec) = make_uint4((uint32_t)(c[kTeItems - 1] + excl), (uint32_t)((c[kTeItems - 1] + excl) >> 32),
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

// The tile records of tiles 2 tid and 2 tid + 1, loaded EARLY (n_tiles <= 512: the common single-device case): the
// caller issues these loads, does unrelated ALU work (the step kernel draws its random numbers), and only then enters
// te_pull, which finds the records in registers instead of re-reading them.  Tiles past n_tiles read as empty.
struct TeRecs2 {
  gjb_tile_rec r[2];
};
template <bool kCg>
__device__ __forceinline__ TeRecs2 te_load_recs2(const gjb_tile_rec* __restrict__ recs, int n_tiles) {
  TeRecs2 o;
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const int t = 2 * threadIdx.x + k;
    if (t < n_tiles) {
      o.r[k] = te_ld_rec<kCg>(recs + t);
    } else {
      o.r[k].mass = 0; o.r[k].e = GJB_TE_E_NONE; o.r[k].reserved = 0;
    }
  }
  return o;
}

// the 8 within-tile CDF values of this thread's parents in tile `ct` (+ the value just before them)
template <bool kCg>
__device__ __forceinline__ void te_ld_row(const uint64_t* __restrict__ ct, uint64_t (&c)[kTeItems], uint64_t& c_prev) {
  c_prev = threadIdx.x ? (kCg ? __ldcg(ct - 1) : __ldg(ct - 1)) : 0ull;
  const ulonglong2* c2 = reinterpret_cast<const ulonglong2*>(ct);
#pragma unroll
  for (int k = 0; k < kTeItems / 2; ++k) {
    const ulonglong2 v = kCg ? __ldcg(c2 + k) : __ldg(c2 + k);
    c[2 * k] = v.x; c[2 * k + 1] = v.y;
  }
}

// Consumer: global parent ids of the offspring slots [w_lo, w_lo + w_n) (w_n <= 2048) of this CTA, in blocked layout
// anc[k] = parent of slot w_lo + tid * 8 + k (junk past w_n).  Returns S (0: no weight has mass, identity written);
// *e_out = E.  `cdf` is this device's array, `cdf_peers` (nullable) every rank's, tile p living on rank p /
// (n_per_rank / 2048).  kFast: n_tiles <= 512 and *pre2 holds this thread's two records (te_load_recs2).
// All kThreads threads call; sm.heads is left in use (ancestors are NOT stored there).
template <bool kCg, bool kFast>
__device__ __forceinline__ uint64_t te_pull(const gjb_tile_rec* __restrict__ recs, int n_tiles,
                                            const uint64_t* __restrict__ cdf, const gjb_peers* cdf_peers, int64_t n_total,
                                            double u0, int64_t w_lo, int w_n, TeSmem& sm, int32_t (&anc)[kTeItems], int* e_out,
                                            const TeRecs2* pre2 = nullptr) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int per = kFast ? 2 : (n_tiles + kThreads - 1) / kThreads;
  const int t0 = tid * per;
  // ---- E = max exponent over tiles with mass
  int emax = GJB_TE_E_NONE;
#pragma unroll
  for (int k = 0; k < per; ++k) {
    const int t = t0 + k;
    if (kFast || t < n_tiles) {
      const gjb_tile_rec r = kFast ? pre2->r[k] : te_ld_rec<kCg>(recs + t);
      if (r.mass) emax = max(emax, r.e);
    }
  }
  emax = __reduce_max_sync(0xffffffffu, emax);
  if (lane == 0) sm.ired[warp] = emax;
  if (tid == 0) { sm.p_lo = 0x7fffffff; sm.p_hi = -1; }
  __syncthreads();
  int E;
  {
    const int4 a = *reinterpret_cast<const int4*>(sm.ired), b = *reinterpret_cast<const int4*>(sm.ired + 4);
    E = max(max(max(a.x, a.y), max(a.z, a.w)), max(max(b.x, b.y), max(b.z, b.w)));
  }
  *e_out = E;
  GJB_TP(2);
  // ---- aligned tile masses, their inclusive prefix
  uint64_t run = 0;
  uint64_t mine[2] = {0ull, 0ull};  // kFast: this thread's two running sums
#pragma unroll
  for (int k = 0; k < per; ++k) {
    const int t = t0 + k;
    if (kFast || t < n_tiles) {
      const gjb_tile_rec r = kFast ? pre2->r[k] : te_ld_rec<kCg>(recs + t);
      const int s = r.mass ? min(E - r.e, 63) : 63;
      run += r.mass >> s;
      if (kFast) mine[k] = run; else sm.pre[t] = run;
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
  for (int w = 0; w < kThreads / 32; w += 2) {
    const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(sm.red + w);
    if (w < warp) excl += v.x;
    if (w + 1 < warp) excl += v.y;
    S += v.x + v.y;
  }
  if (S == 0) {
#pragma unroll
    for (int k = 0; k < kTeItems; ++k) anc[k] = (int32_t)(w_lo + tid * kTeItems + k);
    return 0;
  }
  GJB_TP(3);
  const double scale = __ddiv_rn((double)n_total, (double)S);
  const int32_t nt = (int32_t)n_total;
  const int32_t wl = (int32_t)w_lo, wh = (int32_t)(w_lo + w_n);
  // ---- which parent tiles have offspring in the window (a contiguous range)
  {
    int lo = 0x7fffffff, hi = -1;
    uint64_t prev = excl;
    int32_t cnt_prev = offspring_cnt(prev, S, scale, u0, nt);
#pragma unroll
    for (int k = 0; k < per; ++k) {
      const int t = t0 + k;
      if (kFast || t < n_tiles) {
        const uint64_t cur = (kFast ? mine[k] : sm.pre[t]) + excl;
        if (!kFast || t < n_tiles) sm.pre[t] = cur;
        const int32_t cnt_cur = offspring_cnt(cur, S, scale, u0, nt);
        if (cur != prev && cnt_cur > wl && cnt_prev < wh) {
          lo = min(lo, t);
          hi = max(hi, t);
        }
        prev = cur;
        cnt_prev = cnt_cur;
      }
    }
    if (hi >= 0) { atomicMin(&sm.p_lo, lo); atomicMax(&sm.p_hi, hi); }
  }
  __syncthreads();
  const int p_lo = sm.p_lo, p_hi = sm.p_hi;
  GJB_TP(4);
  // ---- every parent with offspring in the window drops its id (+1) at its first slot.  The rows of tile p + 1 are
  // requested before tile p is processed (one exposed load latency for the whole loop instead of one per tile).
  const int tiles_per_rank = cdf_peers ? (int)(cdf_peers->n_per_rank / kTeTile) : 0;
  auto row_of = [&](int p) -> const uint64_t* {
    if (cdf_peers) {
      const int owner = p / tiles_per_rank;
      return reinterpret_cast<const uint64_t*>(cdf_peers->base[owner]) + (int64_t)(p - owner * tiles_per_rank) * kTeTile + tid * kTeItems;
    }
    return cdf + (int64_t)p * kTeTile + tid * kTeItems;
  };
  uint64_t c[kTeItems], c_prev = 0;
  if (p_lo <= p_hi) te_ld_row<kCg>(row_of(p_lo), c, c_prev);
  for (int p = p_lo; p <= p_hi; ++p) {
    uint64_t cn[kTeItems], cn_prev = 0;
#ifndef GJB_NO_PREFETCH
    if (p < p_hi) te_ld_row<kCg>(row_of(p + 1), cn, cn_prev);
#endif
    const uint64_t base = p ? sm.pre[p - 1] : 0ull;
    if (sm.pre[p] != base) {  // (a tile without aligned mass owns nothing)
      const int s = sm.shf[p];
      int32_t prev = min(max(offspring_cnt(base + (c_prev >> s), S, scale, u0, nt), wl), wh);
      const int32_t last = min(max(offspring_cnt(base + (c[kTeItems - 1] >> s), S, scale, u0, nt), wl), wh);
      if (last > prev) {  // this thread's 8 parents own slots of the window
        const int32_t id1 = p * kTeTile + tid * kTeItems + 1;
#pragma unroll
        for (int k = 0; k < kTeItems; ++k) {
          const int32_t cur = (k == kTeItems - 1) ? last : min(max(offspring_cnt(base + (c[k] >> s), S, scale, u0, nt), wl), wh);
          if (cur > prev) sm.heads[prev - wl] = id1 + k;
          prev = cur;
        }
      }
    }
    if (p < p_hi) {
#ifdef GJB_NO_PREFETCH
      te_ld_row<kCg>(row_of(p + 1), cn, cn_prev);
#endif
#pragma unroll
      for (int k = 0; k < kTeItems; ++k) c[k] = cn[k];
      c_prev = cn_prev;
    }
  }
  __syncthreads();
  GJB_TP(5);
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
  {
    const int4 a = *reinterpret_cast<const int4*>(sm.ired), b = *reinterpret_cast<const int4*>(sm.ired + 4);
    const int32_t wv[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
    for (int w = 0; w < kThreads / 32; ++w) if (w < warp) pre = max(pre, wv[w]);
  }
#pragma unroll
  for (int k = 0; k < kTeItems; ++k) anc[k] = max(v[k], pre) - 1;
  GJB_TP(6);
  return S;
}

// ------------------------------------------------------------------ step table (Design: include/genjax_b200.h)

__device__ __forceinline__ uint32_t te_tag(const gjb_step_link* L, int step) {
  const uint64_t epoch = __ldg(reinterpret_cast<const unsigned long long*>(L->epoch));
  return (uint32_t)(((epoch + 1) << 16) | (uint64_t)((step + 1) & 0xffff));
}
__device__ __forceinline__ uint64_t* te_mail_slot(uint64_t* mailbox, int step, int tile) {
  return mailbox + ((int64_t)(step & 1) * kTeMaxTiles + tile) * GJB_TE_LL_WORDS;
}
__device__ __forceinline__ uint32_t te_ld_volatile_hi(const uint64_t* p) {  // the tag half of a mailbox word
#if defined(__CUDACC__)
  uint32_t v;
  asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(reinterpret_cast<const uint32_t*>(p) + 1) : "memory");
  return v;
#else
  return (uint32_t)(*reinterpret_cast<const volatile uint64_t*>(p) >> 32);
#endif
}
// Wait until the records of tiles [t0, t0 + cnt) (clipped to n_tiles) of `box` all carry `tag`.  The tag halves of up to
// 8 records are requested TOGETHER (24 independent loads, one exposed L2 latency per batch) -- polling record by record
// serialises one L2 round trip per record, which at 4096 tiles (8 GPUs) was most of the step's hand-off time.
__device__ __forceinline__ void te_wait_records(const uint64_t* box, int t0, int cnt, int n_tiles, uint32_t tag) {
  for (int b = 0; b < cnt; b += 8) {
    for (;;) {
      uint32_t bad = 0;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int t = t0 + b + k;
        if (b + k < cnt && t < n_tiles) {
          const uint64_t* r = box + (int64_t)t * GJB_TE_LL_WORDS;
          bad |= (te_ld_volatile_hi(r) ^ tag) | (te_ld_volatile_hi(r + 1) ^ tag) | (te_ld_volatile_hi(r + 2) ^ tag);
        }
      }
      if (!bad) break;
      __nanosleep(40);
    }
  }
}

// Masses and exponents of the (complete) records [t0, t0 + 8) clipped to `cnt` and n_tiles: 16 independent L2 loads.
__device__ __forceinline__ void te_load_batch(const uint64_t* box, int t0, int cnt, int n_tiles, uint64_t (&m)[8], int (&e)[8]) {
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int t = t0 + k;
    m[k] = 0;
    e[k] = GJB_TE_E_NONE;
    if (k < cnt && t < n_tiles) {
      const uint64_t* r = box + (int64_t)t * GJB_TE_LL_WORDS;
      const ulonglong2 w01 = __ldcg(reinterpret_cast<const ulonglong2*>(r));
      const unsigned long long w2 = __ldcg(reinterpret_cast<const unsigned long long*>(r + 2));
      m[k] = (w01.x & 0xffffffffull) | (w01.y << 32);
      e[k] = (int)(uint32_t)w2;
    }
  }
}

// {E ln 2, S, log-mean-exp} of a resampling (one thread)
__device__ __forceinline__ void te_write_lse(double* out, int E, uint64_t S, int64_t n_total);

// End of a filter-step CTA (all kThreads threads call, after te_publish): mail this tile's record to every rank, take a
// ticket; the CTA that draws the last ticket of the launch waits for the records of ALL tiles of ALL ranks, and builds
// the table the next launch consumes: E, S, the inclusive prefix of the aligned tile masses, and per LOCAL window the
// range of parent tiles with offspring in it.  `reskey` = {key0, key1, index_lo, index_hi} of the resampling of THESE
// weights.  Uses sm.pre / sm.shf / sm.red / sm.ired (the CTA's window data is dead by now).
__device__ __forceinline__ void te_finish_step(const gjb_step_link* __restrict__ L, int step, int64_t slot_offset, int64_t n_local,
                                               int64_t n_total, const uint32_t* __restrict__ reskey, gjb_step_table* __restrict__ tab,
                                               double* __restrict__ lse_out, TeSmem& sm, bool light = false) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int world = L->world, tpr = L->tiles_per_rank;
  const uint32_t tag = te_tag(L, step);
  __syncthreads();  // every bulk store (CDF row, state rows, weights) of this CTA has been issued; tile_mass / tile_e are set
  if (tid < world) {
    __threadfence();  // ... and is visible in this device's L2 before the record that announces it leaves
    const uint64_t hi = (uint64_t)tag << 32, m = sm.tile_mass;
    uint64_t* dst = te_mail_slot(L->mailbox[tid], step, L->rank * tpr + (int)blockIdx.x);
    te_st_volatile(dst + 0, (m & 0xffffffffull) | hi);
    te_st_volatile(dst + 1, (m >> 32) | hi);
    te_st_volatile(dst + 2, (uint64_t)(uint32_t)sm.tile_e | hi);
  }
  if (!tab) return;  // mail only: gjb_te_table builds the table beside this launch
  if (tid == 0) sm.is_last = atomicAdd(L->ticket, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!sm.is_last) return;
  // ---------------- the last CTA of this rank
  const int n_tiles = world * tpr;
  const int per = (n_tiles + kThreads - 1) / kThreads;
  const int t0 = tid * per;
  const uint64_t* box = te_mail_slot(L->mailbox[L->rank], step, 0);
  // pass 1: wait for every record (this is the cross-rank barrier of the step), E = max exponent over tiles with mass.
  // Records are read 8 at a time with all 16 loads in flight (te_load_batch): a record-by-record loop pays one L2 round
  // trip per record per thread, which at 4096 tiles (16 records per thread, two passes) was ~9 us of the 8-GPU step.
  te_wait_records(box, t0, per, n_tiles, tag);
  int emax = GJB_TE_E_NONE;
  uint64_t bm[8];
  int be[8];
  for (int b = 0; b < per; b += 8) {
    te_load_batch(box, t0 + b, per - b, n_tiles, bm, be);
#pragma unroll
    for (int k = 0; k < 8; ++k)
      if (bm[k]) emax = max(emax, be[k]);
  }
  emax = __reduce_max_sync(0xffffffffu, emax);
  if (lane == 0) sm.ired[warp] = emax;
  __syncthreads();
  GJB_TP(7);  // (every thread of the last CTA has seen its records)
  int E = sm.ired[0];
#pragma unroll
  for (int w = 1; w < kThreads / 32; ++w) E = max(E, sm.ired[w]);
  if (light) {
    // GJB_STEP_LIGHT: rank totals at the global alignment, their inclusive prefix -- nothing per tile.  (Exactly the sums
    // the full table holds at the rank boundaries: sum_p (mass_p >> min(E - e_p, 63)) in tile order.)
    if (tid < world) sm.pre[tid] = 0ull;
    __syncthreads();
    uint64_t acc = 0;
    int cur = -1;
    for (int b = 0; b < per; b += 8) {
      te_load_batch(box, t0 + b, per - b, n_tiles, bm, be);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int t = t0 + b + k;
        if (b + k < per && t < n_tiles && bm[k]) {
          const int r = t / tpr;
          if (r != cur) {
            if (acc) atomicAdd(reinterpret_cast<unsigned long long*>(&sm.pre[cur]), (unsigned long long)acc);
            acc = 0;
            cur = r;
          }
          acc += bm[k] >> min(E - be[k], 63);
        }
      }
    }
    if (acc) atomicAdd(reinterpret_cast<unsigned long long*>(&sm.pre[cur]), (unsigned long long)acc);
    __syncthreads();
    if (tid == 0) {
      uint64_t run = 0;
      for (int r = 0; r < world; ++r) {
        run += sm.pre[r];
        tab->pre[r] = run;
      }
      tab->S = run; tab->E = E; tab->n_tiles_total = n_tiles;
      if (lse_out) te_write_lse(lse_out, E, run, n_total);
      *L->ticket = 0u;
      __threadfence();
      te_st_volatile(reinterpret_cast<uint64_t*>(&tab->tag), (uint64_t)tag | (1ull << 32));  // {tag, reserved = 1: rank-level table}
    }
    return;
  }
  // pass 2: the thread's total aligned mass (records are complete: plain batched L2 loads), block scan.
  // Nothing per tile goes through shared memory with the thread-contiguous index t = tid * per + k: at per = 16 that is
  // a 128-byte stride, a 32-way bank conflict on every access -- measured as a 21 us table build at 8 ranks
  // (profiles/r2_call34_trace_8gpu.txt).  The offspring counts the window searches probe are stored TRANSPOSED
  // (slot k * kThreads + tid: consecutive threads, consecutive words).
  uint64_t run = 0;
  for (int b = 0; b < per; b += 8) {
    te_load_batch(box, t0 + b, per - b, n_tiles, bm, be);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int t = t0 + b + k;
      if (b + k < per && t < n_tiles && bm[k]) run += bm[k] >> min(E - be[k], 63);
    }
  }
  uint64_t inc = run;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint64_t v = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += v;
  }
  if (lane == 31) sm.red[warp] = inc;
  __syncthreads();
  uint64_t excl = inc - run, S = 0;
#pragma unroll
  for (int w = 0; w < kThreads / 32; ++w) {
    const uint64_t v = sm.red[w];
    if (w < warp) excl += v;
    S += v;
  }
  // pass 3: the records once more -- absolute prefix and shift into the table, cumulative offspring count at every tile
  // boundary into shared memory
  const double u0 = resample_u0(__ldg(reskey), __ldg(reskey + 1), (uint64_t)__ldg(reskey + 2) | ((uint64_t)__ldg(reskey + 3) << 32));
  const double scale = S ? __ddiv_rn((double)n_total, (double)S) : 0.0;
  const int32_t nt = (int32_t)n_total;
  {
    uint64_t cur = excl;
    for (int b = 0; b < per; b += 8) {
      te_load_batch(box, t0 + b, per - b, n_tiles, bm, be);
      const bool whole = b + 8 <= per && t0 + b + 8 <= n_tiles && ((t0 + b) & 7) == 0;  // 8 aligned tiles: their shifts leave as one word
      uint64_t packed = 0;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int t = t0 + b + k;
        if (b + k < per && t < n_tiles) {
          const int sft = bm[k] ? min(E - be[k], 63) : 63;
          cur += bm[k] >> sft;
          tab->pre[t] = cur;
          if (whole) packed |= (uint64_t)sft << (8 * k); else tab->shf[t] = (uint8_t)sft;
          sm.pre[(b + k) * kThreads + tid] = S ? (uint64_t)(int64_t)offspring_cnt(cur, S, scale, u0, nt) : 0ull;
        }
      }
      if (whole) *reinterpret_cast<uint64_t*>(tab->shf + t0 + b) = packed;  // (shf sits at an 8-byte aligned offset of gjb_step_table)
    }
  }
  if (tid == 0) {
    tab->S = S; tab->E = E; tab->n_tiles_total = n_tiles;
    if (lse_out) te_write_lse(lse_out, E, S, n_total);
    *L->ticket = 0u;  // the next launch on the stream starts from zero
  }
  __syncthreads();  // the counts are complete
  // window table: for each local window the first tile whose offspring reach it and the tile that owns its last slot
  auto cnt_at = [&](int p) -> int64_t { return (int64_t)sm.pre[(p % per) * kThreads + p / per]; };
  const int n_win = (int)((n_local + kTeTile - 1) / kTeTile);
  if (S != 0) {
    for (int w = tid; w < n_win; w += kThreads) {
      const int64_t ws = slot_offset + (int64_t)w * kTeTile;
      const int64_t left = slot_offset + n_local - ws;
      const int64_t we = ws + (left < kTeTile ? left : kTeTile);
      int lo = 0, hi = n_tiles;  // smallest p with cnt(P_p) > ws
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (cnt_at(mid) > ws) hi = mid; else lo = mid + 1;
      }
      const int p_first = lo;
      hi = n_tiles;                // smallest p with cnt(P_p) >= we (lo continues from p_first)
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (cnt_at(mid) >= we) hi = mid; else lo = mid + 1;
      }
      tab->win[w][0] = p_first;
      tab->win[w][1] = lo < n_tiles ? lo : n_tiles - 1;
    }
  }
  __syncthreads();
  GJB_TP(13);
  if (tid == 0) {
    __threadfence();
    te_st_volatile(reinterpret_cast<uint64_t*>(&tab->tag), (uint64_t)tag);
  }
}

// Consumer on the table of the previous launch: global parent ids of the offspring slots [w_lo, w_lo + w_n) of local
// window `w_local`, blocked layout as te_pull.  No prefix work: S, E, the parent tile range and the tile prefixes are read.
// want_tag != 0: the table is being built by a kernel running beside this one; spin until it carries the tag.
template <bool kCg>
__device__ __forceinline__ uint64_t te_pull_table(const gjb_step_table* __restrict__ tab, int w_local,
                                                  const uint64_t* __restrict__ cdf, const gjb_peers* cdf_peers, int64_t n_total,
                                                  double u0, int64_t w_lo, int w_n, TeSmem& sm, int32_t (&anc)[kTeItems], int* e_out,
                                                  uint32_t want_tag = 0u) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (want_tag) {
    if (tid == 0) {
      const uint64_t* tw = reinterpret_cast<const uint64_t*>(&tab->tag);  // {tag, reserved} as one aligned 8-byte word
      while ((uint32_t)te_ld_volatile(tw) != want_tag) __nanosleep(500);  // (hundreds of CTAs poll one line: keep it light)
      __threadfence();  // acquire: the table stores that preceded the tag are visible to the loads below
    }
    __syncthreads();
  }
  const uint64_t S = __ldcg(reinterpret_cast<const unsigned long long*>(&tab->S));
  *e_out = __ldcg(&tab->E);
  const int2 win = __ldcg(reinterpret_cast<const int2*>(&tab->win[w_local][0]));
  *reinterpret_cast<int4*>(sm.heads + tid * kTeItems) = make_int4(0, 0, 0, 0);
  *reinterpret_cast<int4*>(sm.heads + tid * kTeItems + 4) = make_int4(0, 0, 0, 0);
  if (S == 0) {
#pragma unroll
    for (int k = 0; k < kTeItems; ++k) anc[k] = (int32_t)(w_lo + tid * kTeItems + k);
    return 0;
  }
  const int p_lo = win.x, p_hi = win.y;
  const int tiles_per_rank = cdf_peers ? (int)(cdf_peers->n_per_rank / kTeTile) : 0;
  auto row_of = [&](int p) -> const uint64_t* {
    if (cdf_peers) {
      const int owner = p / tiles_per_rank;
      return reinterpret_cast<const uint64_t*>(cdf_peers->base[owner]) + (int64_t)(p - owner * tiles_per_rank) * kTeTile + tid * kTeItems;
    }
    return cdf + (int64_t)p * kTeTile + tid * kTeItems;
  };
  uint64_t c[kTeItems], c_prev = 0;
  te_ld_row<kCg>(row_of(p_lo), c, c_prev);
  const double scale = __ddiv_rn((double)n_total, (double)S);
  const int32_t nt = (int32_t)n_total;
  const int32_t wl = (int32_t)w_lo, wh = (int32_t)(w_lo + w_n);
  GJB_TP(4);
  __syncthreads();  // the window is clear
  for (int p = p_lo; p <= p_hi; ++p) {
    uint64_t cn[kTeItems], cn_prev = 0;
#ifndef GJB_NO_PREFETCH
    if (p < p_hi) te_ld_row<kCg>(row_of(p + 1), cn, cn_prev);
#endif
    const uint64_t base = p ? __ldcg(reinterpret_cast<const unsigned long long*>(tab->pre + p - 1)) : 0ull;
    const uint64_t top = __ldcg(reinterpret_cast<const unsigned long long*>(tab->pre + p));
    if (top != base) {  // (a tile without aligned mass owns nothing)
      const int sft = __ldcg(tab->shf + p);
      int32_t prev = min(max(offspring_cnt(base + (c_prev >> sft), S, scale, u0, nt), wl), wh);
      const int32_t last = min(max(offspring_cnt(base + (c[kTeItems - 1] >> sft), S, scale, u0, nt), wl), wh);
      if (last > prev) {  // this thread's 8 parents own slots of the window
        const int32_t id1 = p * kTeTile + tid * kTeItems + 1;
#pragma unroll
        for (int k = 0; k < kTeItems; ++k) {
          const int32_t cur = (k == kTeItems - 1) ? last : min(max(offspring_cnt(base + (c[k] >> sft), S, scale, u0, nt), wl), wh);
          if (cur > prev) sm.heads[prev - wl] = id1 + k;
          prev = cur;
        }
      }
    }
    if (p < p_hi) {
#ifdef GJB_NO_PREFETCH
      te_ld_row<kCg>(row_of(p + 1), cn, cn_prev);
#endif
#pragma unroll
      for (int k = 0; k < kTeItems; ++k) c[k] = cn[k];
      c_prev = cn_prev;
    }
  }
  __syncthreads();
  GJB_TP(5);
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
  {
    const int4 a = *reinterpret_cast<const int4*>(sm.ired), b = *reinterpret_cast<const int4*>(sm.ired + 4);
    const int32_t wv[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
    for (int w = 0; w < kThreads / 32; ++w) if (w < warp) pre = max(pre, wv[w]);
  }
#pragma unroll
  for (int k = 0; k < kTeItems; ++k) anc[k] = max(v[k], pre) - 1;
  GJB_TP(6);
  return S;
}

// Consumer on a RANK-level table (GJB_STEP_LIGHT): S, E and the ranks' prefix come from the table the previous launch's
// last CTA left; the tile prefix is formed here, rank by rank, for the ranks whose tiles have offspring in this window (one,
// or two next to a rank boundary) from the records in this device's own mailbox `box` (complete: the table was written
// after every record had arrived).  Per rank this is te_pull's work on that rank's tiles.  Same outputs as te_pull.
template <bool kCg>
__device__ __forceinline__ uint64_t te_pull_light(const gjb_step_table* __restrict__ tab, const uint64_t* __restrict__ box, int world,
                                                  int tpr, int rank_self, const gjb_peers* cdf_peers, int64_t n_total, double u0,
                                                  int64_t w_lo, int w_n, TeSmem& sm, int32_t (&anc)[kTeItems], int* e_out) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // ONE exposed L2 round trip for everything the prefix work needs: S, E, the two rank prefixes this thread tests, and --
  // speculatively -- this thread's records of the device's OWN rank (where nearly every window's parents live)
  const uint64_t S = __ldcg(reinterpret_cast<const unsigned long long*>(&tab->S));
  const int E = __ldcg(&tab->E);
  uint64_t rp = 0, rc = 0;
  if (tid < world) {
    rp = tid ? __ldcg(reinterpret_cast<const unsigned long long*>(&tab->pre[tid - 1])) : 0ull;
    rc = __ldcg(reinterpret_cast<const unsigned long long*>(&tab->pre[tid]));
  }
  const int per = (tpr + kThreads - 1) / kThreads;
  const int t0 = tid * per;
  const bool spec = per <= 2;
  uint64_t sm_[2] = {0ull, 0ull};
  int se_[2] = {GJB_TE_E_NONE, GJB_TE_E_NONE};
  if (spec) {
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int t = t0 + k;
      if (k < per && t < tpr) {
        const uint64_t* rec = box + (int64_t)(rank_self * tpr + t) * GJB_TE_LL_WORDS;
        const ulonglong2 w01 = __ldcg(reinterpret_cast<const ulonglong2*>(rec));
        const unsigned long long w2 = __ldcg(reinterpret_cast<const unsigned long long*>(rec + 2));
        sm_[k] = (w01.x & 0xffffffffull) | (w01.y << 32);
        se_[k] = (int)(uint32_t)w2;
      }
    }
  }
  *e_out = E;
  if (tid == 0) { sm.p_lo = 0x7fffffff; sm.p_hi = -1; }
  *reinterpret_cast<int4*>(sm.heads + tid * kTeItems) = make_int4(0, 0, 0, 0);
  *reinterpret_cast<int4*>(sm.heads + tid * kTeItems + 4) = make_int4(0, 0, 0, 0);
  if (S == 0) {
#pragma unroll
    for (int k = 0; k < kTeItems; ++k) anc[k] = (int32_t)(w_lo + tid * kTeItems + k);
    return 0;
  }
  const double scale = __ddiv_rn((double)n_total, (double)S);
  const int32_t nt = (int32_t)n_total;
  const int32_t wl = (int32_t)w_lo, wh = (int32_t)(w_lo + w_n);
  __syncthreads();
  // ---- which ranks hold parents of this window
  if (tid < world && rc != rp && offspring_cnt(rc, S, scale, u0, nt) > wl && offspring_cnt(rp, S, scale, u0, nt) < wh) {
    atomicMin(&sm.p_lo, tid);
    atomicMax(&sm.p_hi, tid);
  }
  __syncthreads();
  const int r_lo = sm.p_lo, r_hi = sm.p_hi;
  GJB_TP(2);
  for (int r = r_lo; r <= r_hi; ++r) {
    __syncthreads();  // r_lo / r_hi (first round) or the previous rank's prefix have been read by everyone
    const uint64_t rbase = r ? __ldcg(reinterpret_cast<const unsigned long long*>(&tab->pre[r - 1])) : 0ull;  // (L1/L2 hit: read above)
    if (tid == 0) { sm.p_lo = 0x7fffffff; sm.p_hi = -1; }
    // aligned masses of this rank's tiles, thread-local inclusive prefix
    uint64_t run = 0;
    if (spec && r == rank_self) {
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const int t = t0 + k;
        if (k < per && t < tpr) {
          const int sft = sm_[k] ? min(E - se_[k], 63) : 63;
          run += sm_[k] >> sft;
          sm.pre[t] = run;
          sm.shf[t] = (uint8_t)sft;
        }
      }
    } else {
      for (int k = 0; k < per; ++k) {
        const int t = t0 + k;
        if (t < tpr) {
          const uint64_t* rec = box + (int64_t)(r * tpr + t) * GJB_TE_LL_WORDS;
          const ulonglong2 w01 = __ldcg(reinterpret_cast<const ulonglong2*>(rec));
          const unsigned long long w2 = __ldcg(reinterpret_cast<const unsigned long long*>(rec + 2));
          const uint64_t m = (w01.x & 0xffffffffull) | (w01.y << 32);
          const int sft = m ? min(E - (int)(uint32_t)w2, 63) : 63;
          run += m >> sft;
          sm.pre[t] = run;
          sm.shf[t] = (uint8_t)sft;
        }
      }
    }
    uint64_t inc = run;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint64_t v = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += v;
    }
    if (lane == 31) sm.red[warp] = inc;
    __syncthreads();
    GJB_TP(3);
    uint64_t excl = rbase + inc - run;
#pragma unroll
    for (int w = 0; w < kThreads / 32; ++w)
      if (w < warp) excl += sm.red[w];
    // absolute prefix; which of this rank's tiles have offspring in the window (a contiguous range)
    {
      int lo = 0x7fffffff, hi = -1;
      uint64_t prev = excl;
      int32_t cnt_prev = offspring_cnt(prev, S, scale, u0, nt);
      for (int k = 0; k < per; ++k) {
        const int t = t0 + k;
        if (t < tpr) {
          const uint64_t cur = sm.pre[t] + excl;
          sm.pre[t] = cur;
          const int32_t cnt_cur = offspring_cnt(cur, S, scale, u0, nt);
          if (cur != prev && cnt_cur > wl && cnt_prev < wh) {
            lo = min(lo, t);
            hi = max(hi, t);
          }
          prev = cur;
          cnt_prev = cnt_cur;
        }
      }
      if (hi >= 0) { atomicMin(&sm.p_lo, lo); atomicMax(&sm.p_hi, hi); }
    }
    __syncthreads();
    const int p_lo = sm.p_lo, p_hi = sm.p_hi;
    GJB_TP(4);
    // every parent with offspring in the window drops its id (+1) at its first slot
    const uint64_t* rows = reinterpret_cast<const uint64_t*>(cdf_peers->base[r]) + tid * kTeItems;
    for (int p = p_lo; p <= p_hi; ++p) {
      const uint64_t base = p ? sm.pre[p - 1] : rbase;
      if (sm.pre[p] != base) {
        uint64_t c[kTeItems], c_prev;
        te_ld_row<kCg>(rows + (int64_t)p * kTeTile, c, c_prev);
        const int sft = sm.shf[p];
        int32_t prev = min(max(offspring_cnt(base + (c_prev >> sft), S, scale, u0, nt), wl), wh);
        const int32_t last = min(max(offspring_cnt(base + (c[kTeItems - 1] >> sft), S, scale, u0, nt), wl), wh);
        if (last > prev) {
          const int32_t id1 = (r * tpr + p) * kTeTile + tid * kTeItems + 1;
#pragma unroll
          for (int k = 0; k < kTeItems; ++k) {
            const int32_t cur = (k == kTeItems - 1) ? last : min(max(offspring_cnt(base + (c[k] >> sft), S, scale, u0, nt), wl), wh);
            if (cur > prev) sm.heads[prev - wl] = id1 + k;
            prev = cur;
          }
        }
      }
    }
  }
  __syncthreads();
  GJB_TP(5);
  // ---- inclusive max-scan over the window (as te_pull)
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
  for (int w = 0; w < kThreads / 32; ++w)
    if (w < warp) pre = max(pre, sm.ired[w]);
#pragma unroll
  for (int k = 0; k < kTeItems; ++k) anc[k] = max(v[k], pre) - 1;
  GJB_TP(6);
  return S;
}

// {E ln 2, S, log-mean-exp} of a resampling (one thread)
__device__ __forceinline__ void te_write_lse(double* out, int E, uint64_t S, int64_t n_total) {
  out[0] = S ? (double)E * 0.693147180559945309417 : -INFINITY;
  out[1] = (double)S;
  out[2] = S ? (double)E * 0.693147180559945309417 + log((double)S) - kQLog - log((double)n_total) : -INFINITY;
}

}  // namespace gjb

// libgjb_core.so -- model-independent kernels of the SMC hot path (sm_100a):
// weight max / exact integer mass (log-sum-exp), systematic + multinomial
// resampling over the integer CDF, ancestor gather, RNG test hooks.
// C-ABI declared in include/genjax_b200.h; CPU restatement in oracle/smc.py.
//
// All kernels are HBM/L2-streaming integer/fp32 work: coalesced 128-bit loads,
// warp-shuffle reductions and scans, no tensor cores.
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/genjax_b200.h"
#include "gjb_resample.cuh"
#include "gjb_rng.cuh"
#include "gjb_step.cuh"

namespace gjb {

// ---------------------------------------------------------------- wmax

__global__ void wmax_reset_kernel(uint32_t* wmax) { *wmax = GJB_WMAX_NEG_INF; }

__global__ void __launch_bounds__(256) weight_max_kernel(const float* __restrict__ logw, int64_t n,
                                                         uint32_t* __restrict__ wmax) {
  float m = -INFINITY;
  const int64_t n4 = n >> 2;
  const float4* __restrict__ p4 = reinterpret_cast<const float4*>(logw);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 v = __ldg(p4 + i);
    m = fmaxf(fmaxf(m, v.x), fmaxf(v.y, fmaxf(v.z, v.w)));  // fmaxf drops NaN
  }
  for (int64_t i = (n4 << 2) + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    m = fmaxf(m, logw[i]);
  m = warp_max(m);
  __shared__ float sm[8];
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x < 32) {
    m = threadIdx.x < (blockDim.x >> 5) ? sm[threadIdx.x] : -INFINITY;
    m = warp_max(m);
    if (threadIdx.x == 0 && m > -INFINITY) atomicMax(wmax, fenc(m));
  }
}

// ---------------------------------------------------------------- mass

__device__ __forceinline__ float ref_max(const uint32_t* wmax, const float* m_global) {
  return m_global ? __ldg(m_global) : fdec(__ldg(wmax));
}

__global__ void __launch_bounds__(kThreads) weight_mass_kernel(const float* __restrict__ logw, int64_t n,
                                                               const uint32_t* __restrict__ wmax,
                                                               const float* __restrict__ m_global,
                                                               uint64_t* __restrict__ tile_mass) {
  __shared__ uint64_t sm[kThreads / 32];
  const float M = ref_max(wmax, m_global);
  const uint64_t s = tile_mass_of<false>(logw, n, (int64_t)blockIdx.x * kTile, M, sm);
  if (threadIdx.x == 0) tile_mass[blockIdx.x] = s;
}

// multi-GPU: wait for every rank's max, mass relative to the global max, last CTA pushes this rank's mass
__global__ void __launch_bounds__(kThreads) weight_mass_linked_kernel(const float* __restrict__ logw, int64_t n,
                                                                      uint64_t* __restrict__ tile_mass,
                                                                      uint64_t* __restrict__ tile_prefix,
                                                                      const gjb_link* __restrict__ L, uint64_t wait_max,
                                                                      uint64_t push_mass) {
  __shared__ uint64_t sm[kThreads / 32];
  __shared__ uint64_t vals[GJB_MAX_RANKS];
  link_wait(L, wait_max, vals);
  uint32_t me = 0;
  for (int r = 0; r < L->world; ++r) me = max(me, (uint32_t)vals[r]);
  const float M = fdec(me);
  const uint64_t s = tile_mass_of<false>(logw, n, (int64_t)blockIdx.x * kTile, M, sm);
  if (threadIdx.x == 0) tile_mass[blockIdx.x] = s;
  if (link_last_block(L)) {
    const int nt = (int)gridDim.x;
    uint64_t tot = 0;
    if (tile_prefix) {
      // inclusive prefix of the tile masses for the peers' pull resampler: chunks of kThreads tiles
      uint64_t carry = 0;
      for (int t0 = 0; t0 < nt; t0 += kThreads) {
        const int t = t0 + threadIdx.x;
        const uint64_t v = t < nt ? (uint64_t)__ldcg(reinterpret_cast<const unsigned long long*>(tile_mass) + t) : 0ull;
        uint64_t inc = v;
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const uint64_t u = __shfl_up_sync(0xffffffffu, inc, o);
          if (lane >= o) inc += u;
        }
        __syncthreads();
        if (lane == 31) sm[warp] = inc;
        __syncthreads();
        uint64_t wpre = 0, ctot = 0;
#pragma unroll
        for (int w = 0; w < kThreads / 32; ++w) { if (w < warp) wpre += sm[w]; ctot += sm[w]; }
        if (t < nt) tile_prefix[t] = carry + wpre + inc;
        carry += ctot;
      }
      tot = carry;
      __threadfence();  // the prefix stores are at L2 before the flag leaves
      __syncthreads();
    } else {
      for (int t = threadIdx.x; t < nt; t += kThreads)
        tot += __ldcg(reinterpret_cast<const unsigned long long*>(tile_mass) + t);
      tot = block_sum_u64(tot, sm);
    }
    link_push(L, push_mass, tot);
  }
}

// pull resampler: CTA b walks over the ranks p and, for every parent tile (p, b) whose offspring overlap
// THIS rank's slots [out_lo, out_lo + out_n), scans it and writes those ancestors into the local array.
// With balanced weights only p == rank (plus a neighbour at the block edges) passes the two rejects.
__global__ void __launch_bounds__(kThreads) resample_pull_kernel(const __grid_constant__ gjb_resample_args R,
                                                                 const __grid_constant__ gjb_peers LW,
                                                                 const __grid_constant__ gjb_peers PF,
                                                                 const gjb_link* __restrict__ L, uint64_t wait_max,
                                                                 uint64_t wait_mass) {
  __shared__ TileSmem sm;
  __shared__ uint64_t vmax[GJB_MAX_RANKS];
  __shared__ uint64_t vmass[GJB_MAX_RANKS];
  __shared__ __align__(16) int32_t heads[kWin];
  const int tid = threadIdx.x;
  const int world = L->world;
  const int64_t npr = LW.n_per_rank;
  const int64_t n_total = R.n_total, out_lo = R.out_lo, out_n = R.out_n;
  link_wait(L, wait_max, vmax);
  uint32_t me = 0;
  for (int r = 0; r < world; ++r) me = max(me, (uint32_t)vmax[r]);
  const float M = fdec(me);
  link_wait(L, wait_mass, vmass);
  uint64_t S = 0;
  for (int r = 0; r < world; ++r) S += vmass[r];
  if (blockIdx.x == 0 && tid == 0) {
    if (R.lse_out) {
      R.lse_out[0] = (double)M;
      R.lse_out[1] = (double)S;
      R.lse_out[2] = S ? (double)M + log((double)S) - kQLog - log((double)n_total) : -INFINITY;
    }
    if (R.wmax_next) *R.wmax_next = GJB_WMAX_NEG_INF;
  }
  const int64_t tile_base = (int64_t)blockIdx.x * kTile;
  if (S == 0) {  // every weight is zero: identity ancestors for my own slots
    for (int k = tid; k < kTile; k += kThreads) {
      const int64_t i = tile_base + k;
      if (i < npr) R.ancestors[i] = (int32_t)(out_lo + i);
    }
    return;
  }
  uint32_t key0 = R.key0, key1 = R.key1;
  uint64_t key_index = R.key_index;
  if (R.key_dev) {
    key0 = __ldg(R.key_dev);
    key1 = __ldg(R.key_dev + 1);
    key_index = (uint64_t)__ldg(R.key_dev + 2) | ((uint64_t)__ldg(R.key_dev + 3) << 32);
  }
  const double u0 = resample_u0(key0, key1, key_index);
  const double scale = __ddiv_rn((double)n_total, (double)S);
  const int32_t nt = (int32_t)n_total;
  const int64_t w_lo = out_lo, w_hi = out_lo + out_n;
  uint64_t base = 0;
  for (int p = 0; p < world; ++p) {
    const uint64_t Sp = vmass[p];
    // rank-level reject: offspring range of all of rank p's parents
    const bool rank_hit = (int64_t)offspring_cnt(base + Sp, S, scale, u0, nt) > w_lo &&
                          (int64_t)offspring_cnt(base, S, scale, u0, nt) < w_hi;
    if (rank_hit) {
      // tile-level reject from rank p's inclusive tile prefix (2 words, local or over NVLink)
      const unsigned long long* pf = reinterpret_cast<const unsigned long long*>(PF.base[p]);
      const uint64_t pre = blockIdx.x ? (uint64_t)__ldcg(pf + blockIdx.x - 1) : 0ull;
      const uint64_t end = (uint64_t)__ldcg(pf + blockIdx.x);
      if ((int64_t)offspring_cnt(base + end, S, scale, u0, nt) > w_lo && (int64_t)offspring_cnt(base + pre, S, scale, u0, nt) < w_hi) {
        const float* lw = reinterpret_cast<const float*>(LW.base[p]);
        resample_tile<true>(lw, npr, tile_base, M, base + pre, S, n_total, u0, out_lo, out_n, (int64_t)p * npr, R.ancestors,
                            sm, heads);
      }
    }
    base += Sp;
  }
}

__global__ void __launch_bounds__(256) lse_finalize_kernel(const uint64_t* __restrict__ tile_mass, int n_tiles,
                                                           const uint32_t* __restrict__ wmax,
                                                           const float* __restrict__ m_global, int64_t n_total,
                                                           double* __restrict__ out) {
  uint64_t s = 0;
  for (int t = threadIdx.x; t < n_tiles; t += blockDim.x) s += tile_mass[t];
  s = warp_sum_u64(s);
  __shared__ uint64_t sm[8];
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint64_t tot = 0;
    for (int w = 0; w < (blockDim.x >> 5); ++w) tot += sm[w];
    const double M = (double)ref_max(wmax, m_global);
    out[0] = M;
    out[1] = (double)tot;
    out[2] = tot ? M + log((double)tot) - kQLog - log((double)n_total) : -INFINITY;
  }
}

// ------------------------------------------------------------ systematic

__global__ void __launch_bounds__(kThreads) resample_systematic_kernel(const __grid_constant__ gjb_resample_args R,
                                                                       const __grid_constant__ gjb_peers P,
                                                                       const gjb_link* __restrict__ L, uint64_t wait_max,
                                                                       uint64_t wait_mass, uint64_t push_barrier) {
  const gjb_peers* peers = P.world > 1 ? &P : nullptr;
  const int64_t n = R.n, n_total = R.n_total, out_lo = R.out_lo, out_n = R.out_n, anc_base = R.anc_base;
  uint32_t key0 = R.key0, key1 = R.key1;
  uint64_t key_index = R.key_index;
  if (R.key_dev) {
    key0 = __ldg(R.key_dev);
    key1 = __ldg(R.key_dev + 1);
    key_index = (uint64_t)__ldg(R.key_dev + 2) | ((uint64_t)__ldg(R.key_dev + 3) << 32);
  }
  __shared__ TileSmem sm;
  __shared__ uint64_t sm_b[kThreads / 32];
  __shared__ uint64_t vals[GJB_MAX_RANKS];
  __shared__ __align__(16) int32_t heads[kWin];
  const int tid = threadIdx.x;
  const int n_tiles = gridDim.x;

  // exclusive prefix of the tiles before this one + total mass
  uint64_t pre = 0, tot = 0;
  for (int t = tid; t < n_tiles; t += kThreads) {
    const uint64_t v = R.tile_mass[t];
    tot += v;
    if (t < (int)blockIdx.x) pre += v;
  }
  pre = block_sum_u64(pre, sm.red);
  tot = block_sum_u64(tot, sm_b);
  uint64_t S, off;
  float M;
  if (L) {  // multi-GPU, fused hand-offs: global max and every rank's mass come from the pad
    link_wait(L, wait_max, vals);
    uint32_t me = 0;
    for (int r = 0; r < L->world; ++r) me = max(me, (uint32_t)vals[r]);
    M = fdec(me);
    __syncthreads();
    link_wait(L, wait_mass, vals);
    uint64_t before = 0, all = 0;
    for (int r = 0; r < L->world; ++r) { if (r < L->rank) before += vals[r]; all += vals[r]; }
    S = all;
    off = pre + before;
  } else {
    S = R.s_total ? __ldg(R.s_total) : tot;
    off = pre + (R.c_offset ? __ldg(R.c_offset) : 0ull);
    M = ref_max(R.wmax, R.m_global);
  }
  if (blockIdx.x == 0 && tid == 0) {
    if (R.lse_out) {
      R.lse_out[0] = (double)M;
      R.lse_out[1] = (double)S;
      R.lse_out[2] = S ? (double)M + log((double)S) - kQLog - log((double)n_total) : -INFINITY;
    }
    if (R.wmax_next) *R.wmax_next = GJB_WMAX_NEG_INF;
  }

  const int64_t tile_base = (int64_t)blockIdx.x * kTile;
  if (S == 0) {  // every weight is zero: identity ancestors (collection invalid)
    const AncRoute route{peers ? nullptr : R.ancestors - out_lo, peers};
    for (int k = tid; k < kTile; k += kThreads) {
      const int64_t i = tile_base + k;
      const int64_t j = anc_base + i;
      if (i < n && j >= out_lo && j < out_lo + out_n) *route.at((int32_t)j) = (int32_t)j;
    }
  } else {
    const double u0 = resample_u0(key0, key1, key_index);
    resample_tile<false>(R.logw, n, tile_base, M, off, S, n_total, u0, out_lo, out_n, anc_base, R.ancestors, sm, heads,
                         nullptr, peers);
  }
  if (L && push_barrier) {
    if (link_last_block(L)) link_push_release(L, push_barrier, 0ull);
  }
}

// mass + systematic resampling in one cooperative launch: phase B (integer masses, kept in registers),
// grid barrier, phase C (CDF offset, scan, offspring ranges)
__global__ void __launch_bounds__(kThreads) mass_resample_kernel(const __grid_constant__ gjb_resample_args R) {
  const int64_t n = R.n, n_total = R.n_total, out_lo = R.out_lo, out_n = R.out_n, anc_base = R.anc_base;
  __shared__ TileSmem sm;
  __shared__ uint64_t sm_b[kThreads / 32];
  __shared__ __align__(16) int32_t heads[kWin];
  const int tid = threadIdx.x;
  const int n_tiles = gridDim.x;
  const int64_t tile_base = (int64_t)blockIdx.x * kTile;
  const float M = fdec(__ldg(R.wmax));
  unsigned long long* tm = reinterpret_cast<unsigned long long*>(const_cast<uint64_t*>(R.tile_mass));
  // ---- phase B
  uint64_t q[kItems];
  {
    float x[kItems];
    load_items<false>(R.logw, n, tile_base + tid * kItems, x);
    uint64_t s = 0;
#pragma unroll
    for (int k = 0; k < kItems; ++k) { q[k] = det_exp_q(__fadd_rn(x[k], -M)); s += q[k]; }
    s = block_sum_u64(s, sm.red);
    if (tid == 0) __stcg(tm + blockIdx.x, (unsigned long long)s);
    if (blockIdx.x == 0 && tid == 0 && R.heavy_ws) __stcg(R.heavy_ws, 0u);  // empty heavy list for phase C
  }
  cooperative_groups::this_grid().sync();
  // ---- phase C
  uint64_t pre = 0, tot = 0;
  for (int t = tid; t < n_tiles; t += kThreads) {
    const uint64_t v = __ldcg(tm + t);
    tot += v;
    if (t < (int)blockIdx.x) pre += v;
  }
  block_sum2_u64(pre, tot, sm.red, sm_b);
  const uint64_t S = tot;
  if (blockIdx.x == 0 && tid == 0) {
    if (R.lse_out) {
      R.lse_out[0] = (double)M;
      R.lse_out[1] = (double)S;
      R.lse_out[2] = S ? (double)M + log((double)S) - kQLog - log((double)n_total) : -INFINITY;
    }
    if (R.wmax_next) *R.wmax_next = GJB_WMAX_NEG_INF;
  }
  if (S == 0) {
    for (int k = tid; k < kTile; k += kThreads) {
      const int64_t i = tile_base + k;
      const int64_t j = anc_base + i;
      if (i < n && j >= out_lo && j < out_lo + out_n) R.ancestors[j - out_lo] = (int32_t)j;
    }
    return;  // uniform over the grid: nobody reaches the second barrier
  }
  uint32_t key0 = R.key0, key1 = R.key1;
  uint64_t key_index = R.key_index;
  if (R.key_dev) {
    key0 = __ldg(R.key_dev);
    key1 = __ldg(R.key_dev + 1);
    key_index = (uint64_t)__ldg(R.key_dev + 2) | ((uint64_t)__ldg(R.key_dev + 3) << 32);
  }
  const double u0 = resample_u0(key0, key1, key_index);
  // can any particle own a whole window?  the heaviest possible particle has mass 2^36 (weight == max), i.e. at most
  // n_total * 2^36 / S offspring -- a grid-uniform test, so balanced steps never pay for the second barrier
  const bool may_heavy = R.heavy_ws != nullptr && (double)n_total * 68719476736.0 >= (double)S * (double)kWin;
  resample_tile<false>(R.logw, n, tile_base, M, pre, S, n_total, u0, out_lo, out_n, anc_base, R.ancestors, sm, heads, nullptr,
                       nullptr, q, may_heavy ? R.heavy_ws : nullptr);
  if (may_heavy) {
    cooperative_groups::this_grid().sync();
    const uint32_t cnt_h = min(__ldcg(R.heavy_ws), (uint32_t)GJB_HEAVY_CAP);
    const int32_t* e = reinterpret_cast<const int32_t*>(R.heavy_ws) + 4;
    int32_t* anc = R.ancestors - out_lo;
    const int64_t gtid = (int64_t)blockIdx.x * kThreads + tid, gsz = (int64_t)gridDim.x * kThreads;
    for (uint32_t h = 0; h < cnt_h; ++h) {
      const int32_t lo = __ldcg(e + 3 * h), hi = __ldcg(e + 3 * h + 1), a = __ldcg(e + 3 * h + 2);
      // 128-bit stores over the aligned body, scalars at the ragged ends
      const int64_t body_lo = ((int64_t)lo + 3) & ~(int64_t)3, body_hi = (int64_t)hi & ~(int64_t)3;
      if (body_lo < body_hi && ((reinterpret_cast<uintptr_t>(anc + body_lo) & 15) == 0)) {
        const int4 v = make_int4(a, a, a, a);
        for (int64_t j = body_lo + 4 * gtid; j < body_hi; j += 4 * gsz) *reinterpret_cast<int4*>(anc + j) = v;
        if (gtid < body_lo - lo) anc[lo + gtid] = a;
        if (gtid < hi - body_hi) anc[body_hi + gtid] = a;
      } else {
        for (int64_t j = lo + gtid; j < hi; j += gsz) anc[j] = a;
      }
    }
  }
}

// ----------------------------------------------------------- multinomial

__global__ void __launch_bounds__(kThreads) cdf_kernel(const float* __restrict__ logw, int64_t n,
                                                       const uint32_t* __restrict__ wmax,
                                                       const uint64_t* __restrict__ tile_mass,
                                                       uint64_t* __restrict__ cdf) {
  __shared__ uint64_t sm_a[kThreads / 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  uint64_t pre = 0;
  for (int t = tid; t < (int)blockIdx.x; t += kThreads) pre += tile_mass[t];
  pre = warp_sum_u64(pre);
  if (lane == 0) sm_a[warp] = pre;
  __syncthreads();
  pre = 0;
#pragma unroll
  for (int w = 0; w < kThreads / 32; ++w) pre += sm_a[w];
  __syncthreads();
  const float M = fdec(__ldg(wmax));
  const int64_t base = (int64_t)blockIdx.x * kTile + tid * kItems;
  float x[kItems];
  load_items<false>(logw, n, base, x);
  uint64_t q[kItems], tsum = 0;
#pragma unroll
  for (int k = 0; k < kItems; ++k) { q[k] = det_exp_q(__fadd_rn(x[k], -M)); tsum += q[k]; }
  uint64_t inc = tsum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint64_t v = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += v;
  }
  if (lane == 31) sm_a[warp] = inc;
  __syncthreads();
  uint64_t wpre = 0;
#pragma unroll
  for (int w = 0; w < kThreads / 32; ++w) if (w < warp) wpre += sm_a[w];
  uint64_t C = pre + wpre + inc - tsum;
#pragma unroll
  for (int k = 0; k < kItems; ++k) {
    C += q[k];
    if (base + k < n) cdf[base + k] = C;
  }
}

__global__ void __launch_bounds__(256) multinomial_search_kernel(const uint64_t* __restrict__ cdf, int64_t n,
                                                                 uint32_t key0, uint32_t key1, uint64_t idx_offset,
                                                                 int64_t n_out, int32_t* __restrict__ ancestors,
                                                                 const uint32_t* __restrict__ key_dev = nullptr) {
  if (key_dev) { key0 = __ldg(key_dev); key1 = __ldg(key_dev + 1); }  // keys read on the device (graph replay)
  const uint64_t S = cdf[n - 1];
  for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < n_out; j += (int64_t)gridDim.x * blockDim.x) {
    if (S == 0) { ancestors[j] = (int32_t)(j < n ? j : n - 1); continue; }
    const uint64_t g = idx_offset + (uint64_t)j;
    const uint4 w = philox4x32_10(make_uint4((uint32_t)g, (uint32_t)(g >> 32), 1u, 0u), key0, key1);
    const uint64_t r = ((uint64_t)w.x << 32) | (uint64_t)w.y;
    const uint64_t tgt = __umul64hi(r, S);
    int64_t lo = 0, hi = n;  // first i with cdf[i] > tgt
    while (lo < hi) {
      const int64_t mid = (lo + hi) >> 1;
      if (__ldg(cdf + mid) > tgt) hi = mid; else lo = mid + 1;
    }
    ancestors[j] = (int32_t)lo;
  }
}

// ---------------------------------------------------------------- gather

template <typename T>
__global__ void __launch_bounds__(256) gather_rows_kernel(const T* __restrict__ src, const int32_t* __restrict__ anc,
                                                          T* __restrict__ dst, int64_t n_out, int32_t w) {
  const int64_t total = n_out * w;
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < total; k += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = k / w;
    const int32_t col = (int32_t)(k - row * w);
    dst[k] = __ldg(src + (int64_t)__ldg(anc + row) * w + col);
  }
}

template <typename T>
__global__ void __launch_bounds__(256) gather_rows_peers_kernel(const __grid_constant__ gjb_peers P,
                                                                const int32_t* __restrict__ anc, T* __restrict__ dst,
                                                                int64_t n_out, int32_t w) {
  const int64_t total = n_out * w;
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < total; k += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = k / w;
    const int32_t col = (int32_t)(k - row * w);
    const int64_t g = __ldg(anc + row);
    const int64_t owner = peer_owner(&P, (uint32_t)g);
    const T* src = reinterpret_cast<const T*>(P.base[owner]);
    dst[k] = __ldg(src + (g - owner * P.n_per_rank) * w + col);
  }
}

// ------------------------------------------------------- cross-rank exchange
// stand-alone form of the hand-offs (one CTA; pushes with a system-scope release, so it also orders
// earlier stores into peer memory): same pad format as the fused form (gjb_resample.cuh)
__global__ void __launch_bounds__(256) exchange_kernel(const __grid_constant__ gjb_xchg_args X) {
  __shared__ uint64_t red[8];
  __shared__ uint64_t vals[GJB_MAX_RANKS];
  const int tid = threadIdx.x;
  const uint64_t epoch = __ldg(reinterpret_cast<const unsigned long long*>(X.epoch));
  const uint32_t tag = (uint32_t)(((epoch + 1) << 16) | (X.tag_offset & 0xffffu));
  uint64_t mine = 0;
  if (X.mode == GJB_XCHG_MAX) {
    mine = (uint64_t)__ldg(X.wmax);
  } else if (X.mode == GJB_XCHG_MASS) {
    uint64_t s = 0;
    for (int t = tid; t < X.n_tiles; t += blockDim.x) s += X.tile_mass[t];
    s = warp_sum_u64(s);
    if ((tid & 31) == 0) red[tid >> 5] = s;
    __syncthreads();
    for (int w = 0; w < 8; ++w) mine += red[w];
  }
  if (tid < X.world) {
    __threadfence_system();
    pad_store(X.pads[tid], X.tag_offset, X.rank, mine, tag);
    vals[tid] = pad_poll(X.pads[X.rank], X.tag_offset, tid, tag);
    __threadfence_system();
  }
  __syncthreads();
  if (tid == 0) {
    if (X.mode == GJB_XCHG_MAX) {
      uint32_t m = 0;
      for (int r = 0; r < X.world; ++r) m = max(m, (uint32_t)vals[r]);
      *X.m_global = fdec(m);
    } else if (X.mode == GJB_XCHG_MASS) {
      uint64_t pre = 0, tot = 0;
      for (int r = 0; r < X.world; ++r) { if (r < X.rank) pre += vals[r]; tot += vals[r]; }
      *X.c_offset = pre;
      *X.s_total = tot;
    }
  }
}

__global__ void epoch_bump_kernel(uint64_t* epoch) { *epoch += 1; }

// ------------------------------------------------------------- RNG hooks

__global__ void philox_fill_kernel(uint32_t k0, uint32_t k1, uint64_t off, uint32_t site, uint32_t chunk, int64_t n,
                                   uint4* __restrict__ out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const uint64_t g = off + (uint64_t)i;
    out[i] = philox4x32_10(make_uint4((uint32_t)g, (uint32_t)(g >> 32), chunk, site), k0, k1);
  }
}

__global__ void normal_fill_kernel(uint32_t k0, uint32_t k1, uint64_t off, uint32_t site, int64_t n, int d,
                                   float* __restrict__ out) {
  const int chunks = (d + 3) >> 2;
  const int64_t total = n * chunks;
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < total; k += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = k / chunks;
    const int c = (int)(k - i * chunks);
    const Lane l = make_lane(k0, k1, off + (uint64_t)i);
    const float4 z = normal4(l, site, (uint32_t)c);
    const float zz[4] = {z.x, z.y, z.z, z.w};
    for (int s = 0; s < 4; ++s)
      if (4 * c + s < d) out[i * d + 4 * c + s] = zz[s];
  }
}

static int grid_for(int64_t work, int threads, int max_blocks = 148 * 8) {
  int64_t b = (work + threads - 1) / threads;
  if (b < 1) b = 1;
  if (b > max_blocks) b = max_blocks;
  return (int)b;
}

static inline int launch_status() {
  const cudaError_t e = cudaGetLastError();
  return (int)e;
}


// ---------------------------------------------------------------- tile-exponent masses (section 1c)

// one CTA per 2048-particle tile: logw -> within-tile CDF + tile record (the stand-alone form of what pf_step_kernel
// does with the weights it has just computed)
__global__ void __launch_bounds__(kThreads) te_mass_kernel(const float* __restrict__ logw, int64_t n,
                                                           uint64_t* __restrict__ cdf, gjb_tile_rec* __restrict__ recs) {
  __shared__ TeSmem sm;
  const int64_t base = (int64_t)blockIdx.x * kTeTile + threadIdx.x * kTeItems;
  float lw[kTeItems];
  load_items<false>(logw, n, base, lw);
  te_publish(lw, cdf + (int64_t)blockIdx.x * kTeTile, recs + blockIdx.x, sm);
}

// one CTA per window of 2048 offspring slots: ancestors of [out_lo, out_lo + out_n) from the tile-exponent CDF
__global__ void __launch_bounds__(kThreads) te_resample_kernel(const __grid_constant__ gjb_te_resample_args A) {
  __shared__ TeSmem sm;
  const int64_t w_lo = A.out_lo + (int64_t)blockIdx.x * kTeTile;
  const int64_t left = A.out_lo + A.out_n - w_lo;
  const int w_n = left < kTeTile ? (int)left : kTeTile;
  const double u0 = resample_u0(__ldg(A.key_dev), __ldg(A.key_dev + 1),
                                (uint64_t)__ldg(A.key_dev + 2) | ((uint64_t)__ldg(A.key_dev + 3) << 32));
  int32_t anc[kTeItems];
  int E;
  uint64_t S;
  if (A.table)
    S = A.cdf_peers ? te_pull_table<true>(A.table, (int)blockIdx.x, A.cdf, A.cdf_peers, A.n_total, u0, w_lo, w_n, sm, anc, &E)
                    : te_pull_table<false>(A.table, (int)blockIdx.x, A.cdf, nullptr, A.n_total, u0, w_lo, w_n, sm, anc, &E);
  else
    S = A.cdf_peers ? te_pull<true, false>(A.recs, A.n_tiles_total, A.cdf, A.cdf_peers, A.n_total, u0, w_lo, w_n, sm, anc, &E)
                    : te_pull<false, false>(A.recs, A.n_tiles_total, A.cdf, nullptr, A.n_total, u0, w_lo, w_n, sm, anc, &E);
  if (blockIdx.x == 0 && threadIdx.x == 0 && A.lse_out) te_write_lse(A.lse_out, E, S, A.n_total);
  int32_t* out = A.ancestors + (int64_t)blockIdx.x * kTeTile + threadIdx.x * kTeItems;
  const int j0 = threadIdx.x * kTeItems;
  if (j0 + kTeItems <= w_n && ((reinterpret_cast<uintptr_t>(out) & 15) == 0)) {
    reinterpret_cast<int4*>(out)[0] = make_int4(anc[0], anc[1], anc[2], anc[3]);
    reinterpret_cast<int4*>(out)[1] = make_int4(anc[4], anc[5], anc[6], anc[7]);
  } else {
#pragma unroll
    for (int k = 0; k < kTeItems; ++k) if (j0 + k < w_n) out[k] = anc[k];
  }
}

// The table of one step as a launch of its own: ONE small CTA per device (256 threads, <= 32 registers, 16 KB of shared
// memory) so that it becomes resident BESIDE the step kernel's CTAs as soon as they have all started -- an SM that hosts
// three step CTAs has room for it -- and the next step kernel, launched behind it with programmatic stream
// serialization, can start early too.  It polls this device's mailbox until every tile of every rank carries the step's
// tag (the cross-rank hand-off), then E, the aligned tile masses, their inclusive prefix, the offspring count at every
// tile boundary, and the parent-tile range of every local window.  Same arithmetic as te_finish_step's last-CTA path.
constexpr int kTabThreads = 256;
__global__ void __launch_bounds__(kTabThreads, 4) te_table_kernel(const __grid_constant__ gjb_te_table_args A) {
  __shared__ int32_t cnt[kTeMaxTiles];
  __shared__ __align__(16) uint64_t red[kTabThreads / 32];
  __shared__ __align__(16) int32_t ired[kTabThreads / 32];
  pdl_launch_dependents();
  const gjb_step_link* L = A.link;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n_tiles = L->world * L->tiles_per_rank;
  const int per = (n_tiles + kTabThreads - 1) / kTabThreads;
  const int t0 = tid * per;
  const uint32_t tag = te_tag(L, A.step);
  const uint64_t* box = te_mail_slot(L->mailbox[L->rank], A.step, 0);
  gjb_step_table* tab = A.table_out;
  // pass 1: wait for every record (batched tag polls), E = max exponent over tiles with mass
  te_wait_records(box, t0, per, n_tiles, tag);
  int emax = GJB_TE_E_NONE;
  uint64_t bm[8];
  int be[8];
  for (int b = 0; b < per; b += 8) {  // 8 records per round trip (te_load_batch)
    te_load_batch(box, t0 + b, per - b, n_tiles, bm, be);
#pragma unroll
    for (int k = 0; k < 8; ++k)
      if (bm[k]) emax = max(emax, be[k]);
  }
  emax = __reduce_max_sync(0xffffffffu, emax);
  if (lane == 0) ired[warp] = emax;
  __syncthreads();
  int E = ired[0];
#pragma unroll
  for (int w = 1; w < kTabThreads / 32; ++w) E = max(E, ired[w]);
  // pass 2: aligned masses (records re-read from L2: they are complete now), thread-local inclusive prefix -> table
  uint64_t run = 0;
  for (int b = 0; b < per; b += 8) {
    te_load_batch(box, t0 + b, per - b, n_tiles, bm, be);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int t = t0 + b + k;
      if (b + k < per && t < n_tiles) {
        const int sft = bm[k] ? min(E - be[k], 63) : 63;
        run += bm[k] >> sft;
        tab->pre[t] = run;
        tab->shf[t] = (uint8_t)sft;
      }
    }
  }
  uint64_t inc = run;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint64_t v = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += v;
  }
  if (lane == 31) red[warp] = inc;
  __syncthreads();
  uint64_t excl = inc - run, S = 0;
#pragma unroll
  for (int w = 0; w < kTabThreads / 32; ++w) {
    const uint64_t v = red[w];
    if (w < warp) excl += v;
    S += v;
  }
  const double u0 = resample_u0(__ldg(A.reskey), __ldg(A.reskey + 1), (uint64_t)__ldg(A.reskey + 2) | ((uint64_t)__ldg(A.reskey + 3) << 32));
  const double scale = S ? __ddiv_rn((double)A.n_total, (double)S) : 0.0;
  const int32_t nt = (int32_t)A.n_total;
  for (int k = 0; k < per; ++k) {
    const int t = t0 + k;
    if (t < n_tiles) {
      const uint64_t cur = tab->pre[t] + excl;  // (this thread's own store of pass 2)
      tab->pre[t] = cur;
      cnt[t] = S ? offspring_cnt(cur, S, scale, u0, nt) : 0;
    }
  }
  if (tid == 0) {
    tab->S = S; tab->E = E; tab->n_tiles_total = n_tiles;
    if (A.lse_out) te_write_lse(A.lse_out, E, S, A.n_total);
  }
  __syncthreads();
  const int n_win = (int)((A.n_local + kTeTile - 1) / kTeTile);
  for (int w = tid; w < (S ? n_win : 0); w += kTabThreads) {
    const int64_t ws = A.slot_offset + (int64_t)w * kTeTile;
    const int64_t left = A.slot_offset + A.n_local - ws;
    const int64_t we = ws + (left < kTeTile ? left : kTeTile);
    int lo = 0, hi = n_tiles;  // smallest p with cnt(P_p) > ws
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if ((int64_t)cnt[mid] > ws) hi = mid; else lo = mid + 1;
    }
    const int p_first = lo;
    hi = n_tiles;                // smallest p with cnt(P_p) >= we
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if ((int64_t)cnt[mid] >= we) hi = mid; else lo = mid + 1;
    }
    tab->win[w][0] = p_first;
    tab->win[w][1] = lo < n_tiles ? lo : n_tiles - 1;
  }
  __syncthreads();
  if (tid == 0) {  // the tag goes last (only the opt-in flag hand-off reads it)
    __threadfence();
    te_st_volatile(reinterpret_cast<uint64_t*>(&tab->tag), (uint64_t)tag);
  }
}

// ---------------------------------------------------------------- small host-path helpers (no eager torch on the path)

// MH accept: mask[i] = log(u[i]) < w[i]  (tests/inference/test_requests.py:136-137 `jnp.log(uniform) < w`)
__global__ void __launch_bounds__(256) accept_mask_kernel(const float* __restrict__ u, const float* __restrict__ w, int64_t n,
                                                          int32_t* __restrict__ mask) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    mask[i] = logf(__ldg(u + i)) < __ldg(w + i) ? 1 : 0;
}

// out[i, :] = mask[i] ? a[i, :] : b[i, :]  (`jtu.tree_map(lambda v1, v2: jnp.where(check, v1, v2), new, old)`); rows of
// `words` 32-bit words; a_stride / b_stride = 0 broadcasts one row
__global__ void __launch_bounds__(256) select_rows_kernel(const int32_t* __restrict__ mask, const uint32_t* __restrict__ a,
                                                          const uint32_t* __restrict__ b, uint32_t* __restrict__ out, int64_t n,
                                                          int words, int64_t a_stride, int64_t b_stride) {
  const int64_t total = n * words;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = e / words;
    const int k = (int)(e - i * words);
    out[e] = __ldg(mask + i) ? __ldg(a + i * a_stride + k) : __ldg(b + i * b_stride + k);
  }
}

// effective sample size (sum w)^2 / sum w^2 of w = exp(logw - M): one CTA, fp64 accumulation in a fixed order
__global__ void __launch_bounds__(1024) ess_kernel(const float* __restrict__ logw, int64_t n, const double* __restrict__ lse3,
                                                   double* __restrict__ out) {
  __shared__ double s1[32], s2[32];
  const float M = (float)lse3[0];
  double a = 0.0, b = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += 1024) {
    const double w = (double)expf(__ldg(logw + i) - M);
    a += w; b += w * w;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); }
  if ((threadIdx.x & 31) == 0) { s1[threadIdx.x >> 5] = a; s2[threadIdx.x >> 5] = b; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double ta = 0.0, tb = 0.0;
    for (int w = 0; w < 32; ++w) { ta += s1[w]; tb += s2[w]; }
    out[0] = tb > 0.0 ? ta * ta / tb : 0.0;
  }
}

// Per-step key table of a filter run (core/key.py pf_key_table, oracle/smc.py pf_step_keys), derived on the device from
// the run key's two (collapsed) words: row t = {prop_k0, prop_k1, res_k0, res_k1, res_idx_lo = 1, res_idx_hi = 0, mn_k0, mn_k1}.
__global__ void pf_key_table_kernel(uint32_t k0, uint32_t k1, int T, uint32_t* __restrict__ out) {
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < T; t += gridDim.x * blockDim.x) {
    const uint2 f = threefry2x32_20(k0, k1, 0u, (uint32_t)t);             // fold_in(key, t)
    const uint2 s = threefry2x32_20(f.x, f.y, 0x73706C74u, 0u);           // split(.): words of lanes 0 (propose) and 1 (resample)
    const uint2 p = threefry2x32_20(s.x, s.y, 0x73706C74u, 0u);           // split(k_prop, N): the proposal lanes' words
    const uint2 c = threefry2x32_20(s.x, s.y, 0x5851F42Du, 1u);           // k_res collapsed (lane 1 folded into the words)
    const uint2 m = threefry2x32_20(c.x, c.y, 0x73706C74u, 0u);           // split(k_res, N): the multinomial lanes' words
    uint4* o = reinterpret_cast<uint4*>(out + 8 * (int64_t)t);
    o[0] = make_uint4(p.x, p.y, s.x, s.y);
    o[1] = make_uint4(1u, 0u, m.x, m.y);
  }
}

}  // namespace gjb

using namespace gjb;

extern "C" {

// ---- measurement probe (not part of the ABI header; scratch/probe_remote_stores.py): `ctas` CTAs each issue `dests` x `words`
// 8-byte volatile stores into `dst` (a peer-mapped buffer), the record-mailing pattern of te_finish_step, or -- coalesced = 1 --
// ONE CTA writes the same words with consecutive threads on consecutive addresses.
__global__ void remote_store_probe_kernel(uint64_t* dst, int dests, int words, int coalesced, uint64_t tag) {
  if (coalesced) {
    const int total = (int)gridDim.x * dests * words;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) te_st_volatile(dst + i, tag | (uint64_t)i);
    return;
  }
  if ((int)threadIdx.x < dests) {
    __threadfence();
    uint64_t* d = dst + ((int64_t)threadIdx.x * gridDim.x + blockIdx.x) * 4;
    for (int w = 0; w < words; ++w) te_st_volatile(d + w, tag | (uint64_t)w);
  }
}
int gjb_remote_store_probe(void* dst, int ctas, int dests, int words, int coalesced, uint64_t tag, void* stream) {
  remote_store_probe_kernel<<<coalesced ? 8 : ctas, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<uint64_t*>(dst), dests, words, coalesced, tag);
  return launch_status();
}

int gjb_abi_version(void) { return GJB_ABI_VERSION; }

int64_t gjb_resample_workspace_bytes(int64_t n) {
  if (n < 0) return GJB_E_ARG;
  const int64_t tiles = (n + kTile - 1) / kTile;
  return (tiles > 0 ? tiles : 1) * 8;
}

int gjb_wmax_reset(uint32_t* wmax, void* stream) {
  if (!wmax) return GJB_E_ARG;
  wmax_reset_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(wmax);
  return launch_status();
}

int gjb_weight_max(const float* logw, int64_t n, uint32_t* wmax, void* stream) {
  if (!logw || !wmax || n < 0) return GJB_E_ARG;
  if ((reinterpret_cast<uintptr_t>(logw) & 15) != 0) return GJB_E_ARG;
  if (n == 0) return 0;
  weight_max_kernel<<<grid_for(n / 4 + 1, 256), 256, 0, (cudaStream_t)stream>>>(logw, n, wmax);
  return launch_status();
}

int gjb_weight_mass(const float* logw, int64_t n, const uint32_t* wmax, const float* m_global, uint64_t* tile_mass,
                    void* stream) {
  if (!logw || !tile_mass || (!wmax && !m_global) || n < 0) return GJB_E_ARG;
  if (n == 0) return 0;
  const int64_t tiles = (n + kTile - 1) / kTile;
  if (n >= GJB_MASS_MAX_PARTICLES) return GJB_E_RANGE;  // a shard alone may not exceed what the total may
  weight_mass_kernel<<<(int)tiles, kThreads, 0, (cudaStream_t)stream>>>(logw, n, wmax, m_global, tile_mass);
  return launch_status();
}

int gjb_lse_finalize(const uint64_t* tile_mass, int64_t n, const uint32_t* wmax, const float* m_global,
                     int64_t n_total, double* lse_out, void* stream) {
  if (!tile_mass || !lse_out || (!wmax && !m_global) || n <= 0 || n_total <= 0) return GJB_E_ARG;
  if (n_total >= GJB_MASS_MAX_PARTICLES || n >= GJB_MASS_MAX_PARTICLES) return GJB_E_RANGE;
  const int64_t tiles = (n + kTile - 1) / kTile;
  lse_finalize_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(tile_mass, (int)tiles, wmax, m_global, n_total, lse_out);
  return launch_status();
}

int gjb_resample_systematic(const gjb_resample_args* a, void* stream) {
  if (!a || !a->logw || !a->tile_mass || !a->ancestors || (!a->wmax && !a->m_global)) return GJB_E_ARG;
  if (a->n <= 0 || a->n_total <= 0 || a->out_n < 0 || a->out_lo < 0) return GJB_E_ARG;
  if (a->n_total >= GJB_MASS_MAX_PARTICLES) return GJB_E_RANGE;  // S = sum of masses <= 2^36 each must stay < 2^63
  const int64_t tiles = (a->n + kTile - 1) / kTile;
  gjb_peers none = {};
  resample_systematic_kernel<<<(int)tiles, kThreads, 0, (cudaStream_t)stream>>>(*a, none, nullptr, 0, 0, 0);
  return launch_status();
}

static int mass_resample_resident() {
  static int cached = 0;
  if (!cached) {
    int dev = 0, sms = 0, occ = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, mass_resample_kernel, kThreads, 0);
    cached = sms * occ;
  }
  return cached;
}

int gjb_mass_resample_fits(int64_t n) {
  if (n <= 0) return 0;
  const int64_t tiles = (n + kTile - 1) / kTile;
  return tiles <= mass_resample_resident() ? 1 : 0;
}

int gjb_mass_resample_systematic(const gjb_resample_args* a, void* stream) {
  if (!a || !a->logw || !a->tile_mass || !a->ancestors || !a->wmax) return GJB_E_ARG;
  if (a->m_global || a->c_offset || a->s_total) return GJB_E_MODE;  // single-device form
  if (a->n <= 0 || a->n_total <= 0 || a->out_n < 0 || a->out_lo < 0) return GJB_E_ARG;
  if (a->n_total >= GJB_MASS_MAX_PARTICLES) return GJB_E_RANGE;
  if (!gjb_mass_resample_fits(a->n)) return GJB_E_RANGE;
  const int64_t tiles = (a->n + kTile - 1) / kTile;
  void* params[1] = {(void*)a};
  const cudaError_t e = cudaLaunchCooperativeKernel((const void*)mass_resample_kernel, dim3((unsigned)tiles), dim3(kThreads),
                                                    params, 0, (cudaStream_t)stream);
  if (e != cudaSuccess) return (int)e;
  return launch_status();
}

int gjb_resample_systematic_peers(const gjb_resample_args* a, const gjb_peers* anc, void* stream) {
  if (!a || !anc || !a->logw || !a->tile_mass || (!a->wmax && !a->m_global)) return GJB_E_ARG;
  if (anc->world < 1 || anc->world > GJB_MAX_RANKS || anc->n_per_rank <= 0 || (anc->n_per_rank & 3)) return GJB_E_ARG;
  if (a->n <= 0 || a->n_total != anc->n_per_rank * anc->world || a->out_lo != 0 || a->out_n != a->n_total) return GJB_E_ARG;
  if (a->n_total >= GJB_MASS_MAX_PARTICLES) return GJB_E_RANGE;
  for (int r = 0; r < anc->world; ++r) if (!anc->base[r]) return GJB_E_ARG;
  const int64_t tiles = (a->n + kTile - 1) / kTile;
  gjb_peers p = *anc;
  gjb_peers_set_divisor(&p);
  if (p.world == 1) p.world = 2, p.base[1] = p.base[0];  // keep the routed path even for one rank (tests)
  resample_systematic_kernel<<<(int)tiles, kThreads, 0, (cudaStream_t)stream>>>(*a, p, nullptr, 0, 0, 0);
  return launch_status();
}

int gjb_resample_systematic_linked(const gjb_resample_args* a, const gjb_peers* anc, const gjb_link* link,
                                   uint64_t wait_max, uint64_t wait_mass, uint64_t push_barrier, void* stream) {
  if (!a || !anc || !link || !a->logw || !a->tile_mass || !wait_max || !wait_mass) return GJB_E_ARG;
  if (anc->world < 2 || anc->world > GJB_MAX_RANKS || anc->n_per_rank <= 0 || (anc->n_per_rank & 3)) return GJB_E_ARG;
  if (a->n <= 0 || a->n_total != anc->n_per_rank * anc->world || a->out_lo != 0 || a->out_n != a->n_total) return GJB_E_ARG;
  if (a->n_total >= GJB_MASS_MAX_PARTICLES) return GJB_E_RANGE;
  const int64_t tiles = (a->n + kTile - 1) / kTile;
  gjb_peers p = *anc;
  gjb_peers_set_divisor(&p);
  resample_systematic_kernel<<<(int)tiles, kThreads, 0, (cudaStream_t)stream>>>(*a, p, link, wait_max, wait_mass,
                                                                               push_barrier);
  return launch_status();
}

int gjb_weight_mass_linked(const float* logw, int64_t n, uint64_t* tile_mass, const gjb_link* link, uint64_t wait_max,
                           uint64_t push_mass, void* stream) {
  if (!logw || !tile_mass || !link || n <= 0 || !wait_max || !push_mass) return GJB_E_ARG;
  const int64_t tiles = (n + kTile - 1) / kTile;
  if (n >= GJB_MASS_MAX_PARTICLES) return GJB_E_RANGE;
  weight_mass_linked_kernel<<<(int)tiles, kThreads, 0, (cudaStream_t)stream>>>(logw, n, tile_mass, nullptr, link, wait_max, push_mass);
  return launch_status();
}

int gjb_weight_mass_prefix_linked(const float* logw, int64_t n, uint64_t* tile_mass, uint64_t* tile_prefix,
                                  const gjb_link* link, uint64_t wait_max, uint64_t push_mass, void* stream) {
  if (!logw || !tile_mass || !tile_prefix || !link || n <= 0 || !wait_max || !push_mass) return GJB_E_ARG;
  const int64_t tiles = (n + kTile - 1) / kTile;
  if (n >= GJB_MASS_MAX_PARTICLES) return GJB_E_RANGE;
  weight_mass_linked_kernel<<<(int)tiles, kThreads, 0, (cudaStream_t)stream>>>(logw, n, tile_mass, tile_prefix, link, wait_max,
                                                                              push_mass);
  return launch_status();
}

int gjb_resample_systematic_pull(const gjb_resample_args* a, const gjb_peers* logw_peers, const gjb_peers* prefix_peers,
                                 const gjb_link* link, uint64_t wait_max, uint64_t wait_mass, void* stream) {
  if (!a || !logw_peers || !prefix_peers || !link || !a->ancestors || !wait_max || !wait_mass) return GJB_E_ARG;
  const int world = logw_peers->world;
  if (world < 1 || world > GJB_MAX_RANKS || prefix_peers->world != world) return GJB_E_ARG;
  const int64_t npr = logw_peers->n_per_rank;
  if (npr <= 0 || (npr & 3) || a->n != npr || a->n_total != npr * world) return GJB_E_ARG;
  if (a->out_n != npr || a->out_lo != (int64_t)logw_peers->rank * npr) return GJB_E_ARG;
  if (a->n_total >= GJB_MASS_MAX_PARTICLES) return GJB_E_RANGE;
  for (int r = 0; r < world; ++r) if (!logw_peers->base[r] || !prefix_peers->base[r]) return GJB_E_ARG;
  const int64_t tiles = (npr + kTile - 1) / kTile;
  resample_pull_kernel<<<dim3((unsigned)tiles), kThreads, 0, (cudaStream_t)stream>>>(
      *a, *logw_peers, *prefix_peers, link, wait_max, wait_mass);
  return launch_status();
}

int gjb_te_masses(const float* logw, int64_t n, uint64_t* cdf, gjb_tile_rec* recs, void* stream) {
  if (!logw || !cdf || !recs || n <= 0) return GJB_E_ARG;
  if ((reinterpret_cast<uintptr_t>(cdf) & 15) || (reinterpret_cast<uintptr_t>(recs) & 15)) return GJB_E_ARG;
  const int64_t tiles = (n + kTeTile - 1) / kTeTile;
  if (n >= GJB_MASS_MAX_PARTICLES) return GJB_E_RANGE;
  te_mass_kernel<<<(int)tiles, kThreads, 0, (cudaStream_t)stream>>>(logw, n, cdf, recs);
  return launch_status();
}

int gjb_te_table(const gjb_te_table_args* a, void* stream) {
  if (!a || !a->link || !a->reskey || !a->table_out || a->n_local <= 0 || a->n_total < a->n_local || a->slot_offset < 0) return GJB_E_ARG;
  if (a->step < 0 || a->step >= 65535) return GJB_E_RANGE;
  if (a->n_total > (1LL << 26) || (a->n_total + kTeTile - 1) / kTeTile > kTeMaxTiles) return GJB_E_RANGE;
  if (a->flags & GJB_STEP_PDL) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(1); cfg.blockDim = dim3(kTabThreads); cfg.dynamicSmemBytes = 0; cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return (int)cudaLaunchKernelEx(&cfg, te_table_kernel, *a);
  }
  te_table_kernel<<<1, kTabThreads, 0, (cudaStream_t)stream>>>(*a);
  return launch_status();
}

int gjb_te_resample(const gjb_te_resample_args* a, void* stream) {
  if (!a || !a->cdf || (!a->recs && !a->table) || !a->key_dev || !a->ancestors || a->out_n < 0 || a->out_lo < 0) return GJB_E_ARG;
  if (a->table && (a->out_lo % kTeTile) != 0) return GJB_E_ARG;  // the table's windows are the device's own aligned windows
  if (a->n_tiles_total <= 0 || a->n_total <= 0 || a->out_lo + a->out_n > a->n_total) return GJB_E_ARG;
  // S <= n_total * (2^36 + 1) must stay below 2^63 (signed conversion in offspring_cnt): n_total <= 2^26
  if (a->n_tiles_total > kTeMaxTiles || a->n_total > (1LL << 26)) return GJB_E_RANGE;
  if ((int64_t)a->n_tiles_total * kTeTile < a->n_total) return GJB_E_ARG;
  if (a->out_n == 0) return 0;
  const int64_t ctas = (a->out_n + kTeTile - 1) / kTeTile;
  te_resample_kernel<<<(int)ctas, kThreads, 0, (cudaStream_t)stream>>>(*a);
  return launch_status();
}

int gjb_resample_multinomial(const float* logw, int64_t n, const uint32_t* wmax, const uint64_t* tile_mass,
                             uint64_t* cdf, uint32_t key0, uint32_t key1, uint64_t idx_offset, int64_t n_out,
                             int32_t* ancestors, void* stream) {
  if (!logw || !wmax || !tile_mass || !cdf || !ancestors || n <= 0 || n_out < 0) return GJB_E_ARG;
  if (n >= GJB_MASS_MAX_PARTICLES) return GJB_E_RANGE;
  const int64_t tiles = (n + kTile - 1) / kTile;
  cdf_kernel<<<(int)tiles, kThreads, 0, (cudaStream_t)stream>>>(logw, n, wmax, tile_mass, cdf);
  int e = launch_status();
  if (e) return e;
  if (n_out == 0) return 0;
  multinomial_search_kernel<<<grid_for(n_out, 256), 256, 0, (cudaStream_t)stream>>>(cdf, n, key0, key1, idx_offset,
                                                                                   n_out, ancestors);
  return launch_status();
}

int gjb_resample_multinomial_keydev(const float* logw, int64_t n, const uint32_t* wmax, const uint64_t* tile_mass,
                                    uint64_t* cdf, const uint32_t* key_dev, uint64_t idx_offset, int64_t n_out,
                                    int32_t* ancestors, void* stream) {
  if (!logw || !wmax || !tile_mass || !cdf || !ancestors || !key_dev || n <= 0 || n_out < 0) return GJB_E_ARG;
  if (n >= GJB_MASS_MAX_PARTICLES) return GJB_E_RANGE;
  const int64_t tiles = (n + kTile - 1) / kTile;
  cdf_kernel<<<(int)tiles, kThreads, 0, (cudaStream_t)stream>>>(logw, n, wmax, tile_mass, cdf);
  int e = launch_status();
  if (e) return e;
  if (n_out == 0) return 0;
  multinomial_search_kernel<<<grid_for(n_out, 256), 256, 0, (cudaStream_t)stream>>>(cdf, n, 0u, 0u, idx_offset, n_out,
                                                                                   ancestors, key_dev);
  return launch_status();
}

int gjb_accept_mask(const float* u, const float* w, int64_t n, int32_t* mask, void* stream) {
  if (!u || !w || !mask || n < 0) return GJB_E_ARG;
  if (n == 0) return 0;
  accept_mask_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(u, w, n, mask);
  return launch_status();
}

int gjb_select_rows(const int32_t* mask, const void* a, const void* b, void* out, int64_t n, int32_t row_bytes,
                    int32_t a_broadcast, int32_t b_broadcast, void* stream) {
  if (!mask || !a || !b || !out || n < 0 || row_bytes <= 0 || (row_bytes & 3)) return GJB_E_ARG;
  if (n == 0) return 0;
  const int words = row_bytes / 4;
  select_rows_kernel<<<grid_for(n * words, 256, 148 * 16), 256, 0, (cudaStream_t)stream>>>(
      mask, (const uint32_t*)a, (const uint32_t*)b, (uint32_t*)out, n, words, a_broadcast ? 0 : words, b_broadcast ? 0 : words);
  return launch_status();
}

int gjb_weight_ess(const float* logw, int64_t n, const double* lse3, double* out, void* stream) {
  if (!logw || !lse3 || !out || n <= 0) return GJB_E_ARG;
  ess_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(logw, n, lse3, out);
  return launch_status();
}

int gjb_pf_key_table(uint32_t key0, uint32_t key1, int32_t T, uint32_t* out, void* stream) {
  if (!out || T <= 0 || (reinterpret_cast<uintptr_t>(out) & 15)) return GJB_E_ARG;
  pf_key_table_kernel<<<(T + 127) / 128 < 64 ? (T + 127) / 128 : 64, 128, 0, (cudaStream_t)stream>>>(key0, key1, T, out);
  return launch_status();
}

int gjb_gather_rows(const void* src, const int32_t* ancestors, void* dst, int64_t n_out, int32_t row_bytes,
                    void* stream) {
  if (!src || !ancestors || !dst || n_out < 0 || row_bytes <= 0 || (row_bytes & 3)) return GJB_E_ARG;
  if (n_out == 0) return 0;
  const bool v16 = (row_bytes % 16 == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0) &&
                   ((reinterpret_cast<uintptr_t>(dst) & 15) == 0);
  if (v16) {
    const int w = row_bytes / 16;
    gather_rows_kernel<uint4><<<grid_for(n_out * w, 256, 148 * 16), 256, 0, (cudaStream_t)stream>>>(
        (const uint4*)src, ancestors, (uint4*)dst, n_out, w);
  } else {
    const int w = row_bytes / 4;
    gather_rows_kernel<uint32_t><<<grid_for(n_out * w, 256, 148 * 16), 256, 0, (cudaStream_t)stream>>>(
        (const uint32_t*)src, ancestors, (uint32_t*)dst, n_out, w);
  }
  return launch_status();
}

int gjb_gather_rows_peers(const gjb_peers* src, const int32_t* ancestors, void* dst, int64_t n_out, int32_t row_bytes,
                          void* stream) {
  if (!src || !ancestors || !dst || n_out < 0 || row_bytes <= 0 || (row_bytes & 3)) return GJB_E_ARG;
  if (src->world < 1 || src->world > GJB_MAX_RANKS || src->n_per_rank <= 0) return GJB_E_ARG;
  for (int r = 0; r < src->world; ++r) if (!src->base[r]) return GJB_E_ARG;
  if (n_out == 0) return 0;
  const bool v16 = (row_bytes % 16 == 0) && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0);
  gjb_peers p = *src;
  gjb_peers_set_divisor(&p);
  if (v16) {
    const int w = row_bytes / 16;
    gather_rows_peers_kernel<uint4><<<grid_for(n_out * w, 256, 148 * 16), 256, 0, (cudaStream_t)stream>>>(
        p, ancestors, (uint4*)dst, n_out, w);
  } else {
    const int w = row_bytes / 4;
    gather_rows_peers_kernel<uint32_t><<<grid_for(n_out * w, 256, 148 * 16), 256, 0, (cudaStream_t)stream>>>(
        p, ancestors, (uint32_t*)dst, n_out, w);
  }
  return launch_status();
}

int gjb_peers_set_divisor(gjb_peers* p) {
  if (!p || p->n_per_rank <= 0 || p->n_per_rank > 0x7fffffffLL) return GJB_E_ARG;
  const uint64_t d = (uint64_t)p->n_per_rank;
  if (d == 1) { p->div_mul = 0; p->div_shr = 0; return 0; }
  unsigned lg = 0;
  while ((1ull << lg) < d) ++lg;  // ceil(log2 d)
  const unsigned sh = 31 + lg;
  p->div_mul = (uint32_t)(((1ull << sh) + d - 1) / d);
  p->div_shr = sh - 32;
  return 0;
}

int gjb_exchange(const gjb_xchg_args* a, void* stream) {
  if (!a || a->world < 1 || a->world > GJB_MAX_RANKS || a->rank < 0 || a->rank >= a->world || !a->epoch) return GJB_E_ARG;
  if (a->tag_offset == 0 || a->tag_offset >= (1ull << 16)) return GJB_E_ARG;
  for (int r = 0; r < a->world; ++r) if (!a->pads[r]) return GJB_E_ARG;
  if (a->mode == GJB_XCHG_MAX) { if (!a->wmax || !a->m_global) return GJB_E_ARG; }
  else if (a->mode == GJB_XCHG_MASS) { if (!a->tile_mass || a->n_tiles <= 0 || !a->c_offset || !a->s_total) return GJB_E_ARG; }
  else if (a->mode != GJB_XCHG_BARRIER) return GJB_E_MODE;
  exchange_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(*a);
  return launch_status();
}

int gjb_epoch_bump(uint64_t* epoch, void* stream) {
  if (!epoch) return GJB_E_ARG;
  epoch_bump_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(epoch);
  return launch_status();
}

int gjb_philox_fill(uint32_t key0, uint32_t key1, uint64_t idx_offset, uint32_t site, uint32_t chunk, int64_t n,
                    uint32_t* out4, void* stream) {
  if (!out4 || n < 0) return GJB_E_ARG;
  if (n == 0) return 0;
  philox_fill_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(key0, key1, idx_offset, site, chunk, n,
                                                                       (uint4*)out4);
  return launch_status();
}

int gjb_normal_fill(uint32_t key0, uint32_t key1, uint64_t idx_offset, uint32_t site, int64_t n, int32_t d,
                    float* out, void* stream) {
  if (!out || n < 0 || d <= 0) return GJB_E_ARG;
  if (n == 0) return 0;
  normal_fill_kernel<<<grid_for(n * ((d + 3) / 4), 256), 256, 0, (cudaStream_t)stream>>>(key0, key1, idx_offset, site,
                                                                                        n, d, out);
  return launch_status();
}

}  // extern "C"

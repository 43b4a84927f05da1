// Device routines of the exact integer-CDF weight pipeline, shared by the
// stand-alone kernels in gjb_core.cu and by the persistent particle-filter
// kernel generated per model (gen/codegen.py).  CPU restatement:
// oracle/smc.py (det_exp_q, lse_terms, systematic_counts).
//
//   M   = max_i lw_i                                   (fp32, order free)
//   q_i = round(2^36 * exp(lw_i - M))                  (uint64)
//   C_i = inclusive prefix sum of q                    (uint64, exact: any
//         partition into tiles / CTAs / GPUs gives the same bits)
//   cnt_i = clamp(ceil(C_i * (N / S) - u0), 0, N)      (fp64 rn mul, rn sub)
//   ancestors[j] = i  for j in [cnt_{i-1}, cnt_i)
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "genjax_b200.h"
#include "gjb_rng.cuh"

namespace gjb {

// ---- fused cross-rank hand-offs (gjb_link): see include/genjax_b200.h
//
// Pad entry (slot, source rank) = two 64-bit words, each carrying its own copy of the 32-bit tag:
//     word0 = value[31:0] | tag << 32,   word1 = value[63:32] | tag << 32
// (the "LL" idea of NCCL: the flag travels inside the 8-byte store it validates, so value and flag need
// no ordering fence between them).  tag = ((epoch + 1) << 16 | offset) is never 0 and never repeats
// within the 4-slot reuse distance.
//
// Ordering of BULK data (pull mode): everything a peer reads after a hand-off (state rows, log-weights,
// tile prefixes) was written LOCALLY by the producer kernel; every CTA makes its writes visible at the
// device's L2 (gpu-scope fence) before it takes its ticket, the last CTA sends the flag after it has seen
// all tickets, and a peer's NVLink reads are served by this device's L2 -- so no system-scope fence
// (several microseconds each on this platform) sits on the critical path.  Push mode, which writes
// ancestors into peer memory, keeps its system-scope release (link_push_release).
__device__ __forceinline__ uint32_t link_tag(const gjb_link* L, uint64_t off) {
  const uint64_t epoch = __ldg(reinterpret_cast<const unsigned long long*>(L->epoch));
  return (uint32_t)(((epoch + 1) << 16) | (off & 0xffffu));
}
__device__ __forceinline__ void pad_store(uint64_t* pad, uint64_t off, int src_rank, uint64_t value, uint32_t tag) {
  volatile uint64_t* dst = pad + ((off % GJB_PAD_SLOTS) * GJB_MAX_RANKS + src_rank) * 2;
  dst[0] = (value & 0xffffffffull) | ((uint64_t)tag << 32);
  dst[1] = (value >> 32) | ((uint64_t)tag << 32);
}
__device__ __forceinline__ uint64_t pad_poll(const uint64_t* pad, uint64_t off, int src_rank, uint32_t tag) {
  const volatile uint64_t* src = pad + ((off % GJB_PAD_SLOTS) * GJB_MAX_RANKS + src_rank) * 2;
  for (;;) {
    const uint64_t w0 = src[0], w1 = src[1];
    if ((uint32_t)(w0 >> 32) == tag && (uint32_t)(w1 >> 32) == tag) return (w0 & 0xffffffffull) | (w1 << 32);
    __nanosleep(20);
  }
}
// every CTA: wait until all ranks' entries of exchange `off` sit in MY pad; vals[r] = rank r's value
__device__ __forceinline__ void link_wait(const gjb_link* L, uint64_t off, uint64_t* vals /* smem [GJB_MAX_RANKS] */) {
  __syncthreads();  // vals may still be read from a previous wait
  if ((int)threadIdx.x < L->world) vals[threadIdx.x] = pad_poll(L->pads[L->rank], off, threadIdx.x, link_tag(L, off));
  __syncthreads();
}
// all threads of every CTA call this after the CTA's last global write; true in the CTA that finishes last
__device__ __forceinline__ bool link_last_block(const gjb_link* L) {
  __shared__ int gjb_is_last;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();  // cumulative over the CTA's writes ordered by the barrier above
    const uint32_t total = gridDim.x * gridDim.y * gridDim.z;
    const uint32_t ticket = atomicAdd(L->counter, 1u);
    gjb_is_last = ticket == total - 1;
    if (gjb_is_last) {
      *L->counter = 0u;  // the next kernel on the stream starts from zero
      __threadfence();
    }
  }
  __syncthreads();
  return gjb_is_last != 0;
}
// threads [0, world) of ONE CTA: push the value into slot[rank] of every peer's pad (no fence: see above)
__device__ __forceinline__ void link_push(const gjb_link* L, uint64_t off, uint64_t value) {
  if ((int)threadIdx.x < L->world) pad_store(L->pads[threadIdx.x], off, L->rank, value, link_tag(L, off));
}
// same after a system-scope release: orders this device's earlier stores into PEER memory before the flag
__device__ __forceinline__ void link_push_release(const gjb_link* L, uint64_t off, uint64_t value) {
  if ((int)threadIdx.x < L->world) {
    __threadfence_system();
    pad_store(L->pads[threadIdx.x], off, L->rank, value, link_tag(L, off));
  }
}

// row / n_per_rank for rows < 2^31 without an integer division (see gjb_peers in include/genjax_b200.h)
__device__ __forceinline__ uint32_t peer_owner(const gjb_peers* P, uint32_t row) {
  return P->n_per_rank == 1 ? row : (__umulhi(row, P->div_mul) >> P->div_shr);
}

// where offspring slot j lives: a local array, or the owning rank's array (peer mapped)
struct AncRoute {
  int32_t* local;            // single device: ancestors - out_lo
  const gjb_peers* peers;    // multi device (kernel parameter space)
  __device__ __forceinline__ int32_t* at(int32_t j) const {
    if (!peers) return local + j;
    const int32_t npr = (int32_t)peers->n_per_rank;
    const int32_t owner = (int32_t)peer_owner(peers, (uint32_t)j);
    return reinterpret_cast<int32_t*>(const_cast<void*>(peers->base[owner])) + (j - owner * npr);
  }
};

constexpr int kTile = 2048;      // particles per tile (fixed: part of the ABI)
constexpr int kThreads = 256;    // threads per block in tile routines
constexpr int kItems = 8;        // particles per thread
constexpr double kQLog = 36.0 * 0.693147180559945309417;  // log(2^36)
static_assert(kTile == kThreads * kItems, "tile shape");

// round(2^36 * exp(x)), x <= 0, from IEEE fp32 mul/add only (oracle/smc.py det_exp_q).
__device__ __forceinline__ uint64_t det_exp_q(float x) {
  float t = __fmul_rn(x, 0x1.715476p+0f);
  if (!(t >= -62.0f)) return 0ull;  // NaN, -inf, negligible
  t = fminf(t, 0.0f);
  const float n = floorf(t);
  const float g = __fadd_rn(__fadd_rn(t, -n), -0.5f);
  float p = 0x1.ffcbfcp-17f;                         // ln2^7/7!
  p = __fadd_rn(__fmul_rn(p, g), 0x1.430912p-13f);  // ln2^6/6!
  p = __fadd_rn(__fmul_rn(p, g), 0x1.5d87fep-10f);  // ln2^5/5!
  p = __fadd_rn(__fmul_rn(p, g), 0x1.3b2ab6p-7f);   // ln2^4/4!
  p = __fadd_rn(__fmul_rn(p, g), 0x1.c6b08ep-5f);   // ln2^3/3!
  p = __fadd_rn(__fmul_rn(p, g), 0x1.ebfbep-3f);    // ln2^2/2!
  p = __fadd_rn(__fmul_rn(p, g), 0x1.62e43p-1f);    // ln2
  p = __fadd_rn(__fmul_rn(p, g), 1.0f);
  p = __fmul_rn(p, 0x1.6a09e6p+0f);                 // sqrt(2)
  const uint64_t m = (uint64_t)__fmul_rn(p, 68719476736.0f);  // 2^36
  const uint32_t sh = (uint32_t)(-n);
  return sh ? ((m + (1ull << (sh - 1))) >> sh) : m;  // round to nearest: unbiased mass
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ uint64_t warp_sum_u64(uint64_t v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// block-wide sum of a uint64 (kThreads threads); result valid in every thread
__device__ __forceinline__ uint64_t block_sum_u64(uint64_t v, uint64_t* sm /*[kThreads/32]*/) {
  v = warp_sum_u64(v);
  __syncthreads();  // protect sm from a previous use
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
  __syncthreads();
  uint64_t tot = 0;
#pragma unroll
  for (int w = 0; w < kThreads / 32; ++w) tot += sm[w];
  return tot;
}

// block-wide sums of two uint64 at once (one barrier pair instead of two); results valid in every thread
__device__ __forceinline__ void block_sum2_u64(uint64_t& a, uint64_t& b, uint64_t* sm_a, uint64_t* sm_b) {
  a = warp_sum_u64(a);
  b = warp_sum_u64(b);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) { sm_a[threadIdx.x >> 5] = a; sm_b[threadIdx.x >> 5] = b; }
  __syncthreads();
  a = 0; b = 0;
#pragma unroll
  for (int w = 0; w < kThreads / 32; ++w) { a += sm_a[w]; b += sm_b[w]; }
}

// the 8 consecutive log-weights of this thread (0-mass padding past n);
// kCg: bypass L1 (data written earlier in the same persistent kernel)
template <bool kCg>
__device__ __forceinline__ void load_items(const float* __restrict__ logw, int64_t n, int64_t base, float (&x)[kItems]) {
  if (base + kItems <= n && ((reinterpret_cast<uintptr_t>(logw + base) & 15) == 0)) {
    const float4* p = reinterpret_cast<const float4*>(logw + base);
    const float4 a = kCg ? __ldcg(p) : __ldg(p);
    const float4 b = kCg ? __ldcg(p + 1) : __ldg(p + 1);
    x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w;
    x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
  } else {
#pragma unroll
    for (int k = 0; k < kItems; ++k) x[k] = (base + k < n) ? (kCg ? __ldcg(logw + base + k) : logw[base + k]) : -INFINITY;
  }
}

// integer mass of one tile (all kThreads threads call; result in every thread)
template <bool kCg>
__device__ __forceinline__ uint64_t tile_mass_of(const float* __restrict__ logw, int64_t n, int64_t tile_base, float M,
                                                 uint64_t* sm, uint64_t* qout = nullptr) {
  float x[kItems];
  load_items<kCg>(logw, n, tile_base + threadIdx.x * kItems, x);
  uint64_t s = 0;
#pragma unroll
  for (int k = 0; k < kItems; ++k) {
    const uint64_t q = det_exp_q(__fadd_rn(x[k], -M));
    if (qout) qout[threadIdx.x * kItems + k] = q;  // blocked layout, read back by the same thread
    s += q;
  }
  return block_sum_u64(s, sm);
}

constexpr int kWin = 4096;  // offspring slots resolved per window (int32 heads in shared memory)

struct TileSmem {
  uint64_t red[kThreads / 32];
  int32_t wred[kThreads / 32];
  int32_t range[2];
  int32_t fill;
  int32_t fill_hi;
};

// cumulative offspring count of a particle whose inclusive CDF value is C (n_total < 2^31):
// clamp(ceil(C * scale - u0), 0, n_total).  0 <= C < S and u0 in (0, 1) put C*scale - u0 inside (-1, n_total),
// so the clamp is a no-op and ceil + convert is ONE round-up conversion; C == S closes the range exactly.
__device__ __forceinline__ int32_t offspring_cnt(uint64_t C, uint64_t S, double scale, double u0, int32_t n_total) {
  const double pos = __dsub_rn(__dmul_rn((double)(long long)C, scale), u0);  // C < 2^63: signed conversion is exact
  const int32_t c = __double2int_ru(pos);
  return C == S ? n_total : c;
}

// One tile of the systematic resampler: scan the tile's masses on top of
// `off` (mass before the tile), turn the inclusive CDF into cumulative
// offspring counts cnt_i, and write ancestors[j - out_lo] = anc_base + i for
// the offspring j in [cnt_{i-1}, cnt_i) that fall inside [out_lo, out_lo+out_n).
//
// Write-out is offspring-centric and divergence free: the tile's offspring
// slots [cnt_first, cnt_last) are resolved in windows of kWin slots held in
// shared memory (`heads`, kWin int32): every particle with offspring in the
// window drops its local index at its first slot, a block-wide inclusive
// max-scan propagates it over the particle's range, and the window is written
// with coalesced stores.
//
// `qin` (nullable) supplies the masses computed by an earlier pass over the
// same tile (blocked layout qin[tid * kItems + k]); it MAY alias `heads`
// (it is consumed before `heads` is written).  Returns the tile's mass (every
// thread).  All kThreads threads must call.
template <bool kCg>
__device__ __forceinline__ uint64_t resample_tile(const float* __restrict__ logw, int64_t n, int64_t tile_base, float M,
                                                  uint64_t off, uint64_t S, int64_t n_total, double u0, int64_t out_lo,
                                                  int64_t out_n, int64_t anc_base, int32_t* __restrict__ ancestors,
                                                  TileSmem& sm, int32_t* heads, const uint64_t* qin = nullptr,
                                                  const gjb_peers* peers = nullptr, const uint64_t* qthread = nullptr,
                                                  uint32_t* heavy_ws = nullptr) {
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  uint64_t q[kItems];
  uint64_t tsum = 0;
  if (qthread) {  // this thread's masses are still in its registers (fused mass + resample kernel)
#pragma unroll
    for (int k = 0; k < kItems; ++k) { q[k] = qthread[k]; tsum += q[k]; }
  } else if (qin) {
#pragma unroll
    for (int k = 0; k < kItems; ++k) { q[k] = qin[tid * kItems + k]; tsum += q[k]; }
  } else {
    float x[kItems];
    load_items<kCg>(logw, n, tile_base + tid * kItems, x);
#pragma unroll
    for (int k = 0; k < kItems; ++k) { q[k] = det_exp_q(__fadd_rn(x[k], -M)); tsum += q[k]; }
  }
  uint64_t inc = tsum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint64_t v = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += v;
  }
  __syncthreads();  // previous users of sm / heads are done; qin fully consumed
  if (lane == 31) sm.red[warp] = inc;
  __syncthreads();
  uint64_t wpre = 0, ttot = 0;
#pragma unroll
  for (int w = 0; w < kThreads / 32; ++w) {
    const uint64_t v = sm.red[w];
    if (w < warp) wpre += v;
    ttot += v;
  }
  uint64_t C = off + wpre + inc - tsum;  // exclusive prefix of this thread

  const double scale = __ddiv_rn((double)n_total, (double)S);
  const int32_t nt = (int32_t)n_total;
  const int32_t w_lo = (int32_t)out_lo, w_hi = (int32_t)(out_lo + out_n);
  // cumulative counts at this thread's kItems + 1 particle boundaries, clamped to the output window
  int32_t cnt[kItems + 1];
  cnt[0] = min(max(offspring_cnt(C, S, scale, u0, nt), w_lo), w_hi);
  const int64_t i_base = tile_base + (int64_t)tid * kItems;
#pragma unroll
  for (int k = 0; k < kItems; ++k) {
    C += q[k];
    const int32_t c = min(max(offspring_cnt(C, S, scale, u0, nt), w_lo), w_hi);
    cnt[k + 1] = (i_base + k < n) ? c : cnt[k];  // padding particles own nothing
  }
  // the tile's offspring range follows from its CDF interval: no exchange needed
  const int32_t r_lo = min(max(offspring_cnt(off, S, scale, u0, nt), w_lo), w_hi);
  const int32_t r_hi = min(max(offspring_cnt(off + ttot, S, scale, u0, nt), w_lo), w_hi);
  const AncRoute anc{peers ? nullptr : ancestors - out_lo, peers};
  const int32_t a0 = (int32_t)(anc_base + tile_base) - 1;  // heads hold local index + 1
  // Balanced tile (every particle has at most kDirect offspring): each thread writes the adjacent offspring
  // ranges of its kItems particles straight from registers -- no shared-memory pass, one block-wide vote.
  constexpr int kDirect = 12;
  int heavy = 0;
#pragma unroll
  for (int k = 0; k < kItems; ++k) heavy |= (cnt[k + 1] - cnt[k]) > kDirect;
  if (tid == 0) sm.fill = 0;
  if (!__syncthreads_or(heavy)) {
#pragma unroll
    for (int k = 0; k < kItems; ++k) {
      const int32_t a = a0 + 1 + tid * kItems + k;
      for (int32_t j = cnt[k]; j < cnt[k + 1]; ++j) *anc.at(j) = a;
    }
    return ttot;
  }
  for (int32_t wb = r_lo; wb < r_hi; wb += kWin) {
    const int32_t we = min(wb + kWin, r_hi);
    const int32_t len = we - wb;
    // slots per thread for this window: multiple of 4, at most kWin / kThreads
    const int per = (((len + kThreads - 1) / kThreads) + 3) & ~3;
    // 1. clear the live part of the window; spot a window owned by ONE heavy particle
    for (int j = tid * 4; j < per * kThreads; j += kThreads * 4) *reinterpret_cast<int4*>(heads + j) = make_int4(0, 0, 0, 0);
#pragma unroll
    for (int k = 0; k < kItems; ++k)
      if (cnt[k] <= wb && cnt[k + 1] >= we) { sm.fill = tid * kItems + k + 1; sm.fill_hi = cnt[k + 1]; }
    __syncthreads();
    const int32_t fill = sm.fill;
    if (fill) {  // degenerate weights: this window belongs to ONE particle -- constant fill, no scan
      const int32_t a = a0 + fill;
      if (heavy_ws) {
        // park the whole run of windows this particle owns; every CTA of the grid fills it after the barrier
        const int32_t span_end = min(wb + ((sm.fill_hi - wb) / kWin) * kWin, r_hi);
        __shared__ int gjb_heavy_slot;
        if (tid == 0) gjb_heavy_slot = (int)atomicAdd(heavy_ws, 1u);
        __syncthreads();
        const int slot = gjb_heavy_slot;
        if (slot < GJB_HEAVY_CAP) {
          if (tid == 0) {
            int32_t* e = reinterpret_cast<int32_t*>(heavy_ws) + 4 + 3 * slot;
            e[0] = wb; e[1] = max(span_end, we); e[2] = a;
            sm.fill = 0;
          }
          __syncthreads();
          wb = max(span_end, we) - kWin;  // the loop increment lands on the first window not owned by this particle
          continue;
        }
      }
      for (int32_t j = wb + tid; j < we; j += kThreads) *anc.at(j) = a;
      __syncthreads();
      if (tid == 0) sm.fill = 0;
      __syncthreads();
      continue;
    }
    // 2. every particle with offspring in the window drops its index at its first slot
#pragma unroll
    for (int k = 0; k < kItems; ++k) {
      const int32_t lo = max(cnt[k], wb), hi = min(cnt[k + 1], we);
      if (hi > lo) heads[lo - wb] = tid * kItems + k + 1;
    }
    __syncthreads();
    // 3. inclusive max-scan: `per` consecutive slots per thread, warp shuffle, block
    constexpr int kPer = kWin / kThreads;
    int32_t v[kPer];
    int32_t run = 0;
#pragma unroll
    for (int c4 = 0; c4 < kPer / 4; ++c4) {
      if (4 * c4 < per) {
        const int4 t4 = *reinterpret_cast<const int4*>(heads + tid * per + 4 * c4);
        run = max(run, t4.x); v[4 * c4] = run;
        run = max(run, t4.y); v[4 * c4 + 1] = run;
        run = max(run, t4.z); v[4 * c4 + 2] = run;
        run = max(run, t4.w); v[4 * c4 + 3] = run;
      }
    }
    int32_t incm = run;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int32_t t = __shfl_up_sync(0xffffffffu, incm, o);
      if (lane >= o) incm = max(incm, t);
    }
    if (lane == 31) sm.wred[warp] = incm;
    const int32_t wexc = __shfl_up_sync(0xffffffffu, incm, 1);
    __syncthreads();
    int32_t pre = lane ? wexc : 0;
#pragma unroll
    for (int w = 0; w < kThreads / 32; ++w) if (w < warp) pre = max(pre, sm.wred[w]);
    pre += a0;
    // 4. coalesced write of the window
    const int32_t jb = wb + tid * per;
    // 128-bit stores need slot alignment (peer blocks are multiples of 4 slots, so an aligned group has one owner)
    const bool vec = peers ? ((jb & 3) == 0) : ((reinterpret_cast<uintptr_t>(anc.local + jb) & 15) == 0);
#pragma unroll
    for (int c4 = 0; c4 < kPer / 4; ++c4) {
      if (4 * c4 < per) {
        const int32_t j = jb + 4 * c4;
        const int4 o4 = make_int4(max(a0 + v[4 * c4], pre), max(a0 + v[4 * c4 + 1], pre), max(a0 + v[4 * c4 + 2], pre),
                                  max(a0 + v[4 * c4 + 3], pre));
        if (vec && j + 4 <= we) {
          *reinterpret_cast<int4*>(anc.at(j)) = o4;
        } else {
          if (j < we) *anc.at(j) = o4.x;
          if (j + 1 < we) *anc.at(j + 1) = o4.y;
          if (j + 2 < we) *anc.at(j + 2) = o4.z;
          if (j + 3 < we) *anc.at(j + 3) = o4.w;
        }
      }
    }
    __syncthreads();
  }
  return ttot;
}

// uniform in (0,1) a systematic resample consumes: lane = key index, site 0, chunk 0, word 0
__device__ __forceinline__ double resample_u0(uint32_t key0, uint32_t key1, uint64_t key_index) {
  return (double)u01(philox4x32_10(make_uint4((uint32_t)key_index, (uint32_t)(key_index >> 32), 0u, 0u), key0, key1).x);
}

// Output-slot ("pull") resampling for ONE CTA: the ancestors of the offspring slots [w_lo, w_lo + w_n) it owns.
// The inclusive prefix of all tile masses is built in shared memory (`pre`, n_tiles <= kPullMaxTiles), the first
// parent tile whose offspring reach the window is found by bisection, and every parent tile overlapping the window
// is scanned by resample_tile restricted to the window (anc[j - w_lo] = parent of slot j).  Under balanced weights
// that is the tile with the same index and a neighbour.  Returns the total mass S (0: identity ancestors written).
// oracle: oracle/smc.py resample_systematic_pull.  All kThreads threads must call.
constexpr int kPullMaxTiles = 2048;
template <bool kCg>
__device__ __forceinline__ uint64_t pull_ancestors(const float* __restrict__ logw, int64_t n,
                                                   const unsigned long long* __restrict__ tile_mass, int n_tiles, float M,
                                                   int64_t n_total, double u0, int64_t w_lo, int64_t w_n,
                                                   int32_t* __restrict__ anc, TileSmem& sm, int32_t* heads, uint64_t* pre) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int per = (n_tiles + kThreads - 1) / kThreads;
  uint64_t run = 0;
  for (int k = 0; k < per; ++k) {
    const int t = tid * per + k;
    if (t < n_tiles) { run += (uint64_t)(kCg ? __ldcg(tile_mass + t) : tile_mass[t]); pre[t] = run; }
  }
  uint64_t inc = run;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint64_t v = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += v;
  }
  __syncthreads();
  if (lane == 31) sm.red[warp] = inc;
  __syncthreads();
  uint64_t wpre = 0, S = 0;
#pragma unroll
  for (int w = 0; w < kThreads / 32; ++w) {
    const uint64_t v = sm.red[w];
    if (w < warp) wpre += v;
    S += v;
  }
  const uint64_t excl = wpre + inc - run;
  for (int k = 0; k < per; ++k) {
    const int t = tid * per + k;
    if (t < n_tiles) pre[t] += excl;
  }
  __syncthreads();
  if (S == 0) {
    for (int64_t j = tid; j < w_n; j += kThreads) anc[j] = (int32_t)(w_lo + j);
    return 0;
  }
  const double scale = __ddiv_rn((double)n_total, (double)S);
  const int32_t nt = (int32_t)n_total;
  int lo = 0, hi = n_tiles;  // smallest p whose cumulative offspring count exceeds w_lo
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if ((int64_t)offspring_cnt(pre[mid], S, scale, u0, nt) > w_lo) hi = mid; else lo = mid + 1;
  }
  for (int p = lo; p < n_tiles; ++p) {
    const uint64_t off = p ? pre[p - 1] : 0ull;
    if ((int64_t)offspring_cnt(off, S, scale, u0, nt) >= w_lo + w_n) break;
    if (pre[p] == off) continue;  // a tile without mass has no offspring
    resample_tile<kCg>(logw, n, (int64_t)p * kTile, M, off, S, n_total, u0, w_lo, w_n, 0, anc, sm, heads);
  }
  return S;
}

}  // namespace gjb

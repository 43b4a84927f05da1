/*
 * C restatement of the oracle's bootstrap particle filter for the linear-Gaussian state-space step
 * (oracle/smc.py particle_filter + oracle/gfi.py generate + oracle/dists.py normal / mv_normal_diag +
 * oracle/rng.py Philox / quad streams / Box-Muller) -- TEST INFRASTRUCTURE, like the rest of oracle/.
 *
 * What the reference does on this path: per step a jax.vmap-ed `step.importance` with the observation
 * constrained (generative_functions/static.py:341-380, distributions/distribution.py:117-147), log-sum-exp of
 * the weights (inference/smc.py:96-97), a categorical resample and a gather (user idiom,
 * docs/cookbook/inactive/inference/mapping_tutorial.ipynb cell 37; smc.py:90-91,102-109).  The reference
 * itself (JAX + TFP) is not installable in this image, so this file restates the oracle modules operation for
 * operation: same Philox counters, same float32 operation order (compile with -ffp-contract=off), log / cos /
 * sin evaluated in double and rounded once exactly as the NumPy code does.  tests/test_oracle_c_port.py checks
 * it against the NumPy oracle bit for bit.  It exists to give bench.py's CPU arm (`--impl reference`,
 * `cpu_baseline`) a multi-threaded (OpenMP) implementation of the same algorithm.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static inline void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t out[4]) {
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    const uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0, hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

static inline float u01(uint32_t bits) { return ((float)(bits >> 9) + 0.5f) * 1.1920928955078125e-07f; }

/* oracle/rng.py box_muller: fp32 polynomials with correctly rounded FMAs (fmaf), operation for operation */
static inline float as_f(int32_t i) { float f; memcpy(&f, &i, 4); return f; }
static inline int32_t as_i(float f) { int32_t i; memcpy(&i, &f, 4); return i; }
static inline void box_muller(uint32_t b0, uint32_t b1, float* z0, float* z1) {
  const float u1 = u01(b0), u2 = u01(b1);
  const int32_t tb = as_i(u1) - 0x3F3504F3;
  const int32_t e = tb >> 23;
  const float f = as_f((tb & 0x007FFFFF) + 0x3F3504F3) - 1.0f;
  const float ef = (float)e;
  const float f2 = f * f;
  float R = 0x1.4237fep-3f;
  R = fmaf(R, f, -0x1.0696e4p-2f); R = fmaf(R, f, 0x1.0c524cp-2f); R = fmaf(R, f, -0x1.22973ap-2f);
  R = fmaf(R, f, 0x1.548882p-2f); R = fmaf(R, f, -0x1.99a3ecp-2f); R = fmaf(R, f, 0x1.000206p-1f);
  R = fmaf(R, f, -0x1.55554ep-1f); R = fmaf(R, f, 0x1.fffffep-1f);
  const float L = fmaf(ef, -0x1.62e43p+0f, fmaf(f2, R, -2.0f * f));
  const float r = sqrtf(L);
  const float tm = fmaf(u2, 4.0f, 12582912.0f);
  const int32_t q = as_i(tm) & 3;
  const float rr = fmaf(tm - 12582912.0f, -0.25f, u2);
  const float z = rr * rr;
  float ps = fmaf(z, -0x1.2d9b7cp+6f, 0x1.465ec4p+6f);
  ps = fmaf(z, ps, -0x1.4abbbap+5f); ps = fmaf(z, ps, 0x1.921fb6p+2f);
  const float sn = rr * ps;
  float pc = fmaf(z, 0x1.db6578p+5f, -0x1.55cb9ap+6f);
  pc = fmaf(z, pc, 0x1.03c1eap+6f); pc = fmaf(z, pc, -0x1.3bd3ccp+4f);
  const float cs = fmaf(z, pc, 1.0f);
  float a = (q & 1) ? sn : cs, b = (q & 1) ? cs : sn;
  if ((q + 1) & 2) a = -a;
  if (q & 2) b = -b;
  *z0 = r * a;
  *z1 = r * b;
}

/* oracle/dists.py normal_logpdf */
static inline float normal_logpdf(float v, float loc, float scale) {
  const float z = v / scale - loc / scale;
  const float ls = (float)log((double)scale);
  return -0.5f * (z * z) - (0.91893853320467274178f + ls);
}

/* oracle/smc.py det_exp_q */
static inline uint64_t det_exp_q(float x) {
  float t = x * 1.4426950408889634f;
  if (!(t >= -62.0f)) return 0;
  if (t > 0.0f) t = 0.0f;
  const float n = floorf(t);
  const float g = (t - n) - 0.5f;
  static const float coef[8] = {1.0f, 0x1.62e43p-1f, 0x1.ebfbep-3f, 0x1.c6b08ep-5f, 0x1.3b2ab6p-7f, 0x1.5d87fep-10f,
                                0x1.430912p-13f, 0x1.ffcbfcp-17f};
  float p = coef[7];
  for (int k = 6; k >= 0; --k) p = p * g + coef[k];
  p = p * 1.4142135623730951f;
  const uint64_t m = (uint64_t)(p * 68719476736.0f);
  const uint32_t sh = (uint32_t)(-n);
  return sh ? ((m + (1ull << (sh - 1))) >> sh) : m;
}

/*
 * One run of the filter.  keys: uint32 [T, 8] rows {prop_k0, prop_k1, res_k0, res_k1, res_idx_lo, res_idx_hi, 0, 0}
 * (genjax_b200/core/key.py pf_key_table == oracle/smc.py pf_step_keys).  d == 1: scalar normal sites (quad
 * streams); d > 1: mv_normal_diag sites (one stream per particle, chunk = dim / 4).  x: [n, d] in/out (final,
 * resampled state).  logw_last / anc_last (nullable): pre-resampling log-weights and ancestors of the last step.
 * Returns 0, or -1 on allocation failure.
 */
int pf_lgssm(int64_t n, int T, int d, float* x, const float* ys, float a, const float* q, float c, const float* r,
             const uint32_t* keys, double* logz_inc, float* logw_last, int32_t* anc_last) {
  float* xn = (float*)malloc(sizeof(float) * n * d);
  float* lw = (float*)malloc(sizeof(float) * n);
  uint64_t* cdf = (uint64_t*)malloc(sizeof(uint64_t) * n);
  int32_t* anc = (int32_t*)malloc(sizeof(int32_t) * n);
  int nth = 1;
#ifdef _OPENMP
  nth = omp_get_max_threads();
#endif
  uint64_t* part = (uint64_t*)malloc(sizeof(uint64_t) * (nth + 1));
  if (!xn || !lw || !cdf || !anc || !part) return -1;
  for (int t = 0; t < T; ++t) {
    const uint32_t* kt = keys + 8 * t;
    const uint32_t k0 = kt[0], k1 = kt[1];
    const float* y = ys + (int64_t)t * d;
    float wmax = -INFINITY;
    /* propose + weight: site "x" is the 1st site (Philox site word 1), "y" is constrained */
#pragma omp parallel for schedule(static) reduction(max : wmax)
    for (int64_t i = 0; i < n; ++i) {
      float w = 0.0f;
      if (d == 1) {
        const uint64_t quad = (uint64_t)i >> 2;
        uint32_t o[4];
        philox4x32_10((uint32_t)quad, (uint32_t)(quad >> 32), 0u, 1u, k0, k1, o);
        float z[4];
        box_muller(o[0], o[1], &z[0], &z[1]);
        box_muller(o[2], o[3], &z[2], &z[3]);
        const float loc = a * x[i];
        const float xv = fmaf(q[0], z[i & 3], loc);
        xn[i] = xv;
        w = normal_logpdf(y[0], c * xv, r[0]);
      } else {
        for (int ch = 0; ch < (d + 3) / 4; ++ch) {
          uint32_t o[4];
          philox4x32_10((uint32_t)i, (uint32_t)((uint64_t)i >> 32), (uint32_t)ch, 1u, k0, k1, o);
          float z[4];
          box_muller(o[0], o[1], &z[0], &z[1]);
          box_muller(o[2], o[3], &z[2], &z[3]);
          for (int s = 0; s < 4; ++s) {
            const int j = 4 * ch + s;
            if (j < d) xn[i * d + j] = fmaf(q[j], z[s], a * x[i * d + j]);
          }
        }
        for (int j = 0; j < d; ++j) w = w + normal_logpdf(y[j], c * xn[i * d + j], r[j]);
      }
      lw[i] = w;
      if (w > wmax) wmax = w;  /* NaN never wins, like np.fmax */
    }
    /* exact integer CDF: per-thread block sums, then offsets */
    uint64_t S = 0;
#pragma omp parallel
    {
      int tid = 0, nt = 1;
#ifdef _OPENMP
      tid = omp_get_thread_num();
      nt = omp_get_num_threads();
#endif
      const int64_t lo = n * tid / nt, hi = n * (tid + 1) / nt;
      uint64_t s = 0;
      for (int64_t i = lo; i < hi; ++i) { s += det_exp_q(lw[i] - wmax); cdf[i] = s; }
      part[tid + 1] = s;
#pragma omp barrier
#pragma omp single
      {
        part[0] = 0;
        for (int k = 1; k <= nt; ++k) part[k] += part[k - 1];
        S = part[nt];
      }
      const uint64_t off = part[tid];
      for (int64_t i = lo; i < hi; ++i) cdf[i] += off;
    }
    logz_inc[t] = S ? (double)wmax + log((double)S) - 36.0 * 0.693147180559945309417 - log((double)n) : -INFINITY;
    /* systematic offspring ranges: u0 from the resample key's lane (site 0, chunk 0, word 0) */
    if (S == 0) {
#pragma omp parallel for schedule(static)
      for (int64_t i = 0; i < n; ++i) anc[i] = (int32_t)i;
    } else {
      uint32_t o[4];
      philox4x32_10(kt[4], kt[5], 0u, 0u, kt[2], kt[3], o);
      const double u0 = (double)u01(o[0]);
      const double scale = (double)n / (double)S;
#pragma omp parallel for schedule(static)
      for (int64_t i = 0; i < n; ++i) {
        const uint64_t Cp = i ? cdf[i - 1] : 0, Cn = cdf[i];
        double p0 = ceil((double)Cp * scale - u0), p1 = ceil((double)Cn * scale - u0);
        if (p0 < 0) p0 = 0;
        if (p0 > (double)n) p0 = (double)n;
        if (p1 < 0) p1 = 0;
        if (p1 > (double)n) p1 = (double)n;
        int64_t c0 = (Cp == S) ? n : (int64_t)p0, c1 = (Cn == S) ? n : (int64_t)p1;
        if (i == 0) c0 = (int64_t)fmin(fmax(ceil(0.0 * scale - u0), 0.0), (double)n);
        for (int64_t j = c0; j < c1; ++j) anc[j] = (int32_t)i;
      }
    }
    if (t == T - 1) {
      if (logw_last) memcpy(logw_last, lw, sizeof(float) * n);
      if (anc_last) memcpy(anc_last, anc, sizeof(int32_t) * n);
    }
    /* gather */
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) memcpy(x + i * d, xn + (int64_t)anc[i] * d, sizeof(float) * d);
  }
  free(xn); free(lw); free(cdf); free(anc); free(part);
  return 0;
}

/* torchrun exports OMP_NUM_THREADS=1; the CPU arm asks for the cores explicitly */
void pf_port_set_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

int pf_port_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

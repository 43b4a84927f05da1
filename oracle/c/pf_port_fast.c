/*
 * PERFORMANCE build of the CPU restatement of the bootstrap particle filter (oracle/c/pf_port.c is the
 * bit-compatible one the tests use) -- TEST INFRASTRUCTURE / CPU baseline of bench.py, like the rest of oracle/.
 *
 * Same algorithm and the same Philox streams, written the way one would write it for speed on a CPU instead of for
 * bit-compatibility with the NumPy oracle: ONE Philox block and two Box-Muller pairs per QUAD of particles (the
 * bit-compatible port recomputes them per particle), single-precision libm (logf / sinf / cosf), FMA contraction
 * allowed, -O3 -march=native, OpenMP over quads; propose + weight + running max in one pass, integer masses + CDF in
 * a second, systematic offspring ranges + the gather of the NEXT step's input fused into the third.  What the
 * reference does on this path: generative_functions/static.py:341-380 (vmapped step.importance),
 * inference/smc.py:96-97 (logsumexp), mapping_tutorial.ipynb cell 37 (categorical resample + gather).
 * Results agree with pf_port.c to float32 rounding of the proposals (tests/test_oracle_c_port.py).
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

static inline void box_muller(uint32_t b0, uint32_t b1, float* z0, float* z1) {
  const float r = sqrtf(-2.0f * logf(u01(b0)));
  const float ang = 6.283185307179586f * u01(b1);
  *z0 = r * cosf(ang);
  *z1 = r * sinf(ang);
}

static inline uint64_t det_exp_q(float x) {
  float t = x * 1.4426950408889634f;
  if (!(t >= -62.0f)) return 0;
  if (t > 0.0f) t = 0.0f;
  const float n = floorf(t);
  const float g = (t - n) - 0.5f;
  float p = 0x1.ffcbfcp-17f;
  p = p * g + 0x1.430912p-13f; p = p * g + 0x1.5d87fep-10f; p = p * g + 0x1.3b2ab6p-7f; p = p * g + 0x1.c6b08ep-5f;
  p = p * g + 0x1.ebfbep-3f; p = p * g + 0x1.62e43p-1f; p = p * g + 1.0f;
  p = p * 1.4142135623730951f;
  const uint64_t m = (uint64_t)(p * 68719476736.0f);
  const uint32_t sh = (uint32_t)(-n);
  return sh ? ((m + (1ull << (sh - 1))) >> sh) : m;
}

int pf_lgssm_fast(int64_t n, int T, int d, float* x, const float* ys, float a, const float* q, float c, const float* r,
                  const uint32_t* keys, double* logz_inc, float* logw_last, int32_t* anc_last) {
  float* xn = (float*)malloc(sizeof(float) * n * d);
  float* lw = (float*)malloc(sizeof(float) * n);
  uint64_t* cdf = (uint64_t*)malloc(sizeof(uint64_t) * n);
  int32_t* anc = anc_last ? (int32_t*)malloc(sizeof(int32_t) * n) : NULL;
  int nth = 1;
#ifdef _OPENMP
  nth = omp_get_max_threads();
#endif
  uint64_t* part = (uint64_t*)malloc(sizeof(uint64_t) * (nth + 1));
  float* inv_r = (float*)malloc(sizeof(float) * d);
  if (!xn || !lw || !cdf || !part || !inv_r || (anc_last && !anc)) return -1;
  float lconst = 0.0f;
  for (int j = 0; j < d; ++j) { inv_r[j] = 1.0f / r[j]; lconst += 0.91893853320467274178f + logf(r[j]); }
  const int64_t nq = (n + 3) / 4;
  for (int t = 0; t < T; ++t) {
    const uint32_t* kt = keys + 8 * t;
    const uint32_t k0 = kt[0], k1 = kt[1];
    const float* y = ys + (int64_t)t * d;
    float wmax = -INFINITY;
    if (d == 1) {
      const float y0 = y[0], ir = inv_r[0], q0 = q[0];
#pragma omp parallel for schedule(static) reduction(max : wmax)
      for (int64_t qd = 0; qd < nq; ++qd) {
        uint32_t o[4];
        philox4x32_10((uint32_t)qd, (uint32_t)((uint64_t)qd >> 32), 0u, 1u, k0, k1, o);
        float z[4];
        box_muller(o[0], o[1], &z[0], &z[1]);
        box_muller(o[2], o[3], &z[2], &z[3]);
        const int64_t i0 = 4 * qd;
        const int m = (int)((n - i0) < 4 ? (n - i0) : 4);
        for (int s = 0; s < m; ++s) {
          const float xv = a * x[i0 + s] + q0 * z[s];
          xn[i0 + s] = xv;
          const float zz = (y0 - c * xv) * ir;
          const float w = -0.5f * zz * zz - lconst;
          lw[i0 + s] = w;
          if (w > wmax) wmax = w;
        }
      }
    } else {
#pragma omp parallel for schedule(static) reduction(max : wmax)
      for (int64_t i = 0; i < n; ++i) {
        float w = -lconst;
        for (int ch = 0; ch < (d + 3) / 4; ++ch) {
          uint32_t o[4];
          philox4x32_10((uint32_t)i, (uint32_t)((uint64_t)i >> 32), (uint32_t)ch, 1u, k0, k1, o);
          float z[4];
          box_muller(o[0], o[1], &z[0], &z[1]);
          box_muller(o[2], o[3], &z[2], &z[3]);
          for (int s = 0; s < 4; ++s) {
            const int j = 4 * ch + s;
            if (j < d) {
              const float xv = a * x[i * d + j] + q[j] * z[s];
              xn[i * d + j] = xv;
              const float zz = (y[j] - c * xv) * inv_r[j];
              w -= 0.5f * zz * zz;
            }
          }
        }
        lw[i] = w;
        if (w > wmax) wmax = w;
      }
    }
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
    if (t == T - 1 && logw_last) memcpy(logw_last, lw, sizeof(float) * n);
    if (S == 0) {
      memcpy(x, xn, sizeof(float) * n * d);
      if (t == T - 1 && anc_last) for (int64_t i = 0; i < n; ++i) anc_last[i] = (int32_t)i;
      continue;
    }
    uint32_t o[4];
    philox4x32_10(kt[4], kt[5], 0u, 0u, kt[2], kt[3], o);
    const double u0 = (double)u01(o[0]);
    const double scale = (double)n / (double)S;
    const int want_anc = (t == T - 1) && anc_last;
    /* systematic offspring ranges, the gather fused: particle i's row goes straight to its offspring slots */
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
      const uint64_t Cp = i ? cdf[i - 1] : 0, Cn = cdf[i];
      if (Cn == Cp) continue;
      double p0 = ceil((double)Cp * scale - u0), p1 = ceil((double)Cn * scale - u0);
      if (p0 < 0) p0 = 0;
      if (p1 > (double)n) p1 = (double)n;
      const int64_t c0 = (int64_t)p0, c1 = (Cn == S) ? n : (int64_t)p1;
      for (int64_t j = c0; j < c1; ++j) {
        if (d == 1) x[j] = xn[i]; else memcpy(x + j * d, xn + i * d, sizeof(float) * d);
        if (want_anc) anc[j] = (int32_t)i;
      }
    }
    if (want_anc) memcpy(anc_last, anc, sizeof(int32_t) * n);
  }
  free(xn); free(lw); free(cdf); free(part); free(inv_r); free(anc);
  return 0;
}

void pf_fast_set_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

int pf_fast_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

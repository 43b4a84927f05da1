// Device library of primitive distributions: (sampler, logpdf) pairs used by
// the generated model kernels.  Formulas follow TFP 0.23's float32 operation
// order as restated in oracle/dists.py (reference call sites:
// generative_functions/distributions/tensorflow_probability/__init__.py:52-62,
// distribution.py:371-396).
#pragma once
#include <math.h>

#include "gjb_rng.cuh"

namespace gjb {

constexpr float kHalfLog2Pi = 0.91893853320467274178f;
constexpr float kLog2OverPiHalf = -0.22579135264472743236f;  // 0.5*log(2/pi)

// ------------------------------------------------------------------ normal
struct Normal {
  // one correctly rounded FMA (oracle/dists.py normal_sample does the same): with the bit-reproducible Box-Muller of
  // gjb_rng.cuh the sampled value is bit-exact against the CPU oracle
  __device__ static __forceinline__ float sample(float z, float loc, float scale) { return __fmaf_rn(scale, z, loc); }
  // logpdf with the particle-invariant pieces precomputed: inv = 1/scale, lc = 0.5 log 2pi + log scale
  __device__ static __forceinline__ float logpdf_r(float v, float loc, float inv, float lc) {
    const float z = v * inv - loc * inv;
    return -0.5f * (z * z) - lc;
  }
  __device__ static __forceinline__ float logpdf(float v, float loc, float scale) {
    const float z = v / scale - loc / scale;
    return -0.5f * (z * z) - (kHalfLog2Pi + logf(scale));
  }
};

struct Uniform {
  __device__ static __forceinline__ float sample(float u, float lo, float hi) { return lo + (hi - lo) * u; }
  __device__ static __forceinline__ float logpdf(float v, float lo, float hi) {
    return (v >= lo && v <= hi) ? -logf(hi - lo) : -INFINITY;
  }
};

struct Exponential {
  __device__ static __forceinline__ float sample(float u, float rate) { return -logf(u) / rate; }
  __device__ static __forceinline__ float logpdf(float v, float rate) {
    return v < 0.0f ? -INFINITY : logf(rate) - rate * v;
  }
};

struct HalfNormal {
  __device__ static __forceinline__ float sample(float z, float scale) { return fabsf(z) * scale; }
  __device__ static __forceinline__ float logpdf(float v, float scale) {
    const float z = v / scale;
    return v < 0.0f ? -INFINITY : kLog2OverPiHalf - logf(scale) - 0.5f * z * z;
  }
};

// ------------------------------------------- long-tail scalar wrappers
// tensorflow_probability/__init__.py:110 (cauchy), :179 (half_cauchy), :214 (laplace), :219 (log_normal), :174 (gumbel),
// :309 (weibull), :204 (kumaraswamy), :224 (logit_normal), :169 (geometric), :194 (inverse_gamma), :120 (chi2), :279 (student_t), :264 (poisson).  Inverse-CDF samplers on one u01 word (u in [2^-25, 1)); log-densities in TFP 0.23's operation order
// as restated in oracle/dists.py.
constexpr float kPi = 3.14159265358979323846f;
constexpr float kLogPi = 1.14472988584940017414f;
constexpr float kLog2OverPi = -0.45158270528945486473f;  // log(2/pi)
constexpr float kLog2 = 0.69314718055994530942f;

struct Cauchy {
  __device__ static __forceinline__ float sample(float u, float loc, float scale) { return loc + scale * tanf(kPi * (u - 0.5f)); }
  __device__ static __forceinline__ float logpdf(float v, float loc, float scale) {
    const float z = (v - loc) / scale;
    return -log1pf(z * z) - (kLogPi + logf(scale));
  }
};

struct HalfCauchy {
  __device__ static __forceinline__ float sample(float u, float loc, float scale) { return loc + scale * fabsf(tanf(kPi * (u - 0.5f))); }
  __device__ static __forceinline__ float logpdf(float v, float loc, float scale) {
    const float z = (v - loc) / scale;
    return v < loc ? -INFINITY : kLog2OverPi - logf(scale) - log1pf(z * z);
  }
};

struct Laplace {
  __device__ static __forceinline__ float sample(float u, float loc, float scale) {
    const float w = 2.0f * u - 1.0f;  // (-1, 1): |w| <= 1 - 2^-24, the logarithm stays finite
    return loc - scale * copysignf(log1pf(-fabsf(w)), w);
  }
  __device__ static __forceinline__ float logpdf(float v, float loc, float scale) {
    return -fabsf((v - loc) / scale) - kLog2 - logf(scale);
  }
};

struct LogNormal {  // exp of a Normal(loc, scale): Normal log-density of log v, minus the log-Jacobian log v
  __device__ static __forceinline__ float sample(float z, float loc, float scale) { return expf(loc + scale * z); }
  __device__ static __forceinline__ float logpdf(float v, float loc, float scale) {
    const float lv = logf(v);
    return v > 0.0f ? Normal::logpdf(lv, loc, scale) - lv : -INFINITY;
  }
};

struct Gumbel {
  __device__ static __forceinline__ float sample(float u, float loc, float scale) { return loc - scale * logf(-logf(u)); }
  __device__ static __forceinline__ float logpdf(float v, float loc, float scale) {
    const float z = (v - loc) / scale;
    return -(z + expf(-z)) - logf(scale);
  }
};

struct Weibull {  // (concentration k, scale s): x = s (-log(1 - u))^(1/k)
  __device__ static __forceinline__ float sample(float u, float k, float s) { return s * expf(logf(-log1pf(-u)) / k); }
  __device__ static __forceinline__ float logpdf(float v, float k, float s) {
    const float t = logf(v) - logf(s);
    return v < 0.0f ? -INFINITY : logf(k) - logf(s) + (k - 1.0f) * t - expf(k * t);
  }
};

struct Kumaraswamy {  // (concentration1 a, concentration0 b): x = (1 - (1 - u)^(1/b))^(1/a)
  __device__ static __forceinline__ float sample(float u, float a, float b) {
    return expf(logf(-expm1f(log1pf(-u) / b)) / a);
  }
  __device__ static __forceinline__ float logpdf(float v, float a, float b) {
    const float lv = logf(v);
    const float t1 = ((a - 1.0f) == 0.0f) ? 0.0f : (a - 1.0f) * lv;
    const float t2 = ((b - 1.0f) == 0.0f) ? 0.0f : (b - 1.0f) * log1pf(-expf(a * lv));
    return (v < 0.0f || v > 1.0f) ? -INFINITY : logf(a) + logf(b) + t1 + t2;
  }
};

struct LogitNormal {  // sigmoid of a Normal(loc, scale): Normal log-density of logit v, minus the log-Jacobian
  __device__ static __forceinline__ float sample(float z, float loc, float scale) { return 1.0f / (1.0f + expf(-(loc + scale * z))); }
  __device__ static __forceinline__ float logpdf(float v, float loc, float scale) {
    const float lv = logf(v), l1 = log1pf(-v);
    return (v > 0.0f && v < 1.0f) ? Normal::logpdf(lv - l1, loc, scale) - lv - l1 : -INFINITY;
  }
};

struct Geometric {  // tfd.Geometric(probs=p): failures before the first success, a float-valued count as in TFP
  __device__ static __forceinline__ float sample(float u, float p) { return floorf(logf(u) / log1pf(-p)); }
  __device__ static __forceinline__ float logpdf(float v, float p) {
    const float t = (v == 0.0f) ? 0.0f : v * log1pf(-p);
    return v < 0.0f ? -INFINITY : t + logf(p);
  }
};

// --------------------------------------------------------- flip / bernoulli
__device__ __forceinline__ float softplusf(float x) { return fmaxf(x, 0.0f) + log1pf(expf(-fabsf(x))); }

struct Flip {  // tfd.Bernoulli(probs=p); value is int32 0/1
  __device__ static __forceinline__ int sample(float u, float p) { return u < p ? 1 : 0; }
  __device__ static __forceinline__ float logpdf(int v, float p) {
    const float x = (float)v;
    const float a = (x == 0.0f) ? 0.0f : x * logf(p);
    const float b = ((1.0f - x) == 0.0f) ? 0.0f : (1.0f - x) * log1pf(-p);
    return a + b;
  }
};

struct Bernoulli {  // tfd.Bernoulli(logits=l)
  __device__ static __forceinline__ int sample(float u, float logit) {
    const float p = 1.0f / (1.0f + expf(-logit));
    return u < p ? 1 : 0;
  }
  __device__ static __forceinline__ float logpdf(int v, float logit) {
    const float x = (float)v;
    return -softplusf(-logit) * x - softplusf(logit) * (1.0f - x);
  }
};

// -------------------------------------------------------------- categorical
// logits: K contiguous floats (shared memory row, global row or registers)
struct Categorical {
  __device__ static __forceinline__ int sample(float u, const float* logits, int K) {
    float m = -INFINITY;
    for (int j = 0; j < K; ++j) m = fmaxf(m, logits[j]);
    float tot = 0.0f;
    for (int j = 0; j < K; ++j) tot += expf(logits[j] - m);
    const float t = u * tot;
    float acc = 0.0f;
    int k = K - 1;
    for (int j = 0; j < K; ++j) {
      acc += expf(logits[j] - m);
      if (acc > t) { k = j; break; }
    }
    return k;
  }
  __device__ static __forceinline__ float logpdf(int v, const float* logits, int K) {
    float m = -INFINITY;
    for (int j = 0; j < K; ++j) m = fmaxf(m, logits[j]);
    float tot = 0.0f;
    for (int j = 0; j < K; ++j) tot += expf(logits[j] - m);
    const int k = min(max(v, 0), K - 1);
    return (logits[k] - m) - logf(tot);
  }
};

// ------------------------------------------------ mixture of diagonal Gaussians
// gmm_diag(logits[K], mu[K, D] row major, sigma[K]); value: D floats of one thread
struct GmmDiag {
  template <int K, int D>
  __device__ static __forceinline__ void sample(const Lane& l, uint32_t site, const float* logits, const float* mu,
                                                const float* sigma, float* out) {
    const int k = Categorical::sample(u01(l.words(site, 0xFFFFu).x), logits, K);
#pragma unroll
    for (int c = 0; c < (D + 3) / 4; ++c) {
      const float4 z = normal4(l, site, (uint32_t)c);
      const float zz[4] = {z.x, z.y, z.z, z.w};
#pragma unroll
      for (int t = 0; t < 4; ++t)
        if (4 * c + t < D) out[4 * c + t] = mu[k * D + 4 * c + t] + sigma[k] * zz[t];
    }
  }
  template <int K, int D>
  __device__ static __forceinline__ float logpdf(const float* v, const float* logits, const float* mu, const float* sigma) {
    float m = -INFINITY;
    for (int j = 0; j < K; ++j) m = fmaxf(m, logits[j]);
    float tot = 0.0f;
    for (int j = 0; j < K; ++j) tot += expf(logits[j] - m);
    const float lse = m + logf(tot);
    float comp[K];
    float M = -INFINITY;
    for (int k = 0; k < K; ++k) {
      float lp = 0.0f;
      for (int d = 0; d < D; ++d) lp += Normal::logpdf(v[d], mu[k * D + d], sigma[k]);
      comp[k] = (logits[k] - lse) + lp;
      M = fmaxf(M, comp[k]);
    }
    float s = 0.0f;
    for (int k = 0; k < K; ++k) s += expf(comp[k] - M);
    return M + logf(s);
  }
};

// ------------------------------------------------------ mv_normal (full covariance)
// tfd.MultivariateNormalFullCovariance(loc, covariance_matrix): L = chol(cov) once per thread per launch
// (particle-invariant, lives in the Uni struct), then per particle a forward substitution
//   z = L^-1 (x - loc),  log_prob = -0.5 |z|^2 - sum_k log L_kk - D/2 log 2pi,   sample = loc + L eps.
// D <= 16 keeps this on FMA (SURVEY kernel K10); the factor is D*(D+1)/2 floats of per-thread state.
struct MvNormal {
  template <int D>
  __device__ static __forceinline__ float cholesky(const float* __restrict__ cov, float* __restrict__ L) {
    float logdet = 0.0f;
#pragma unroll
    for (int i = 0; i < D; ++i) {
#pragma unroll
      for (int j = 0; j <= i; ++j) {
        float s = cov[i * D + j];
        for (int m = 0; m < j; ++m) s -= L[i * D + m] * L[j * D + m];
        if (i == j) {
          const float dgl = sqrtf(s);
          L[i * D + i] = dgl;
          logdet += logf(dgl);
        } else {
          L[i * D + j] = s / L[j * D + j];
        }
      }
#pragma unroll
      for (int j = i + 1; j < D; ++j) L[i * D + j] = 0.0f;
    }
    return logdet;  // sum_k log L_kk
  }
  template <int D>
  __device__ static __forceinline__ void sample(const Lane& l, uint32_t site, const float* loc, const float* L, float* out) {
    float eps[(D + 3) / 4 * 4];
#pragma unroll
    for (int c = 0; c < (D + 3) / 4; ++c) {
      const float4 z = normal4(l, site, (uint32_t)c);
      eps[4 * c] = z.x; eps[4 * c + 1] = z.y; eps[4 * c + 2] = z.z; eps[4 * c + 3] = z.w;
    }
#pragma unroll
    for (int i = 0; i < D; ++i) {
      float s = loc[i];
      for (int m = 0; m <= i; ++m) s += L[i * D + m] * eps[m];
      out[i] = s;
    }
  }
  template <int D>
  __device__ static __forceinline__ float logpdf(const float* v, const float* loc, const float* L, float logdet) {
    float z[D];
    float q = 0.0f;
#pragma unroll
    for (int i = 0; i < D; ++i) {
      float s = v[i] - loc[i];
      for (int m = 0; m < i; ++m) s -= L[i * D + m] * z[m];
      z[i] = s / L[i * D + i];
      q += z[i] * z[i];
    }
    return -0.5f * q - logdet - (float)D * kHalfLog2Pi;
  }
};

// -------------------------------------------------------------- gamma, beta
// Marsaglia-Tsang; attempt t uses chunk chunk0+t: words (x,y) -> normal,
// z -> acceptance uniform, w (attempt 0) -> boost uniform for a < 1.
__device__ __forceinline__ float gamma_mt(const Lane& l, uint32_t site, float a, uint32_t chunk0) {
  const bool boost = a < 1.0f;
  const float ae = boost ? a + 1.0f : a;
  const float d = ae - (1.0f / 3.0f);
  const float c = 1.0f / sqrtf(9.0f * d);
  float out = 0.0f, ub = 0.5f;
  for (uint32_t t = 0; t < 64u; ++t) {
    const uint4 w = l.words(site, chunk0 + t);
    if (t == 0) ub = u01(w.w);
    const float x = box_muller(w.x, w.y).x;
    const float u = u01(w.z);
    const float v = 1.0f + c * x;
    const float v3 = v * v * v;
    if (v > 0.0f && logf(u) < 0.5f * x * x + d - d * v3 + d * logf(v3)) { out = d * v3; break; }
  }
  return boost ? out * expf(logf(ub) / a) : out;
}

struct Gamma {
  __device__ static __forceinline__ float sample(const Lane& l, uint32_t site, float a, float rate) {
    return gamma_mt(l, site, a, 0) / rate;
  }
  __device__ static __forceinline__ float logpdf(float v, float a, float rate) {
    const float t = ((a - 1.0f) == 0.0f) ? 0.0f : (a - 1.0f) * logf(v);
    return t - rate * v - (lgammaf(a) - a * logf(rate));
  }
};

struct InverseGamma {  // (concentration a, scale b): b / Gamma(a, 1)
  __device__ static __forceinline__ float sample(const Lane& l, uint32_t site, float a, float b) { return b / gamma_mt(l, site, a, 0); }
  __device__ static __forceinline__ float logpdf(float v, float a, float b) {
    return v > 0.0f ? a * logf(b) - lgammaf(a) - (a + 1.0f) * logf(v) - b / v : -INFINITY;
  }
};

struct Chi2 {  // (df): Gamma(df / 2, rate 1/2)
  __device__ static __forceinline__ float sample(const Lane& l, uint32_t site, float df) { return gamma_mt(l, site, 0.5f * df, 0) / 0.5f; }
  __device__ static __forceinline__ float logpdf(float v, float df) { return Gamma::logpdf(v, 0.5f * df, 0.5f); }
};

struct StudentT {  // (df, loc, scale): loc + scale * z * rsqrt(g / df), g ~ Gamma(df / 2, rate 1/2), z ~ N(0, 1) on chunk 128
  __device__ static __forceinline__ float sample(const Lane& l, uint32_t site, float df, float loc, float scale) {
    const float g = gamma_mt(l, site, 0.5f * df, 0) / 0.5f;
    const uint4 w = l.words(site, 128u);
    return loc + scale * (box_muller(w.x, w.y).x / sqrtf(g / df));
  }
  __device__ static __forceinline__ float logpdf(float v, float df, float loc, float scale) {
    const float y = (v - loc) / scale;
    const float norm = logf(fabsf(scale)) + 0.5f * logf(df) + 0.5f * kLogPi + lgammaf(0.5f * df) - lgammaf(0.5f * (df + 1.0f));
    return -0.5f * (df + 1.0f) * log1pf(y * y / df) - norm;
  }
};

struct Poisson {  // tfd.Poisson(rate): a float-valued count as in TFP
  // rate < 10: inversion by sequential search on one uniform (chunk 0, word x).  Otherwise PTRS (Hormann 1993,
  // transformed rejection with squeeze): attempt t draws (U, V) from words (x, y) of chunk t.
  __device__ static __forceinline__ float sample(const Lane& l, uint32_t site, float rate) {
    if (rate < 10.0f) {
      const float u = u01(l.words(site, 0u).x);
      float p = expf(-rate), s = p, k = 0.0f;
      while (u > s && k < 128.0f) {
        k += 1.0f;
        p *= rate / k;
        s += p;
      }
      return k;
    }
    const float b = 0.931f + 2.53f * sqrtf(rate), a = -0.059f + 0.02483f * b;
    const float lia = logf(1.1239f + 1.1328f / (b - 3.4f)), vr = 0.9277f - 3.6224f / (b - 2.0f), llam = logf(rate);
    float k = 0.0f;
    for (uint32_t t = 0; t < 64u; ++t) {
      const uint4 w = l.words(site, t);
      const float U = u01(w.x) - 0.5f, V = u01(w.y);
      const float us = 0.5f - fabsf(U);
      k = floorf((2.0f * a / us + b) * U + rate + 0.43f);
      if (us >= 0.07f && V <= vr) break;
      if (k < 0.0f || (us < 0.013f && V > us)) continue;
      if (logf(V) + lia - logf(a / (us * us) + b) <= -rate + k * llam - lgammaf(k + 1.0f)) break;
    }
    return fmaxf(k, 0.0f);
  }
  __device__ static __forceinline__ float logpdf(float v, float rate) {
    const float t = (v == 0.0f) ? 0.0f : v * logf(rate);
    return v < 0.0f ? -INFINITY : t - lgammaf(v + 1.0f) - rate;
  }
};

struct Beta {
  __device__ static __forceinline__ float sample(const Lane& l, uint32_t site, float a, float b) {
    const float ga = gamma_mt(l, site, a, 0);
    const float gb = gamma_mt(l, site, b, 64);
    return ga / (ga + gb);
  }
  __device__ static __forceinline__ float logpdf(float v, float a, float b) {
    const float t1 = ((a - 1.0f) == 0.0f) ? 0.0f : (a - 1.0f) * logf(v);
    const float t2 = ((b - 1.0f) == 0.0f) ? 0.0f : (b - 1.0f) * log1pf(-v);
    return t1 + t2 - (lgammaf(a) + lgammaf(b) - lgammaf(a + b));
  }
};

}  // namespace gjb

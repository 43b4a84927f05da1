// Host build of the exact-arithmetic device functions (see cuda_runtime.h in this directory).
#include "gjb_resample.cuh"

extern "C" {
void h_philox(const uint32_t* ctr, uint32_t k0, uint32_t k1, uint32_t* out) {
  const uint4 r = gjb::philox4x32_10(make_uint4(ctr[0], ctr[1], ctr[2], ctr[3]), k0, k1);
  out[0] = r.x; out[1] = r.y; out[2] = r.z; out[3] = r.w;
}
void h_quad_words(uint32_t k0, uint32_t k1, uint64_t quad, uint32_t site, uint32_t chunk, uint32_t* out) {
  const uint4 r = gjb::quad_words(k0, k1, quad, site, chunk);
  out[0] = r.x; out[1] = r.y; out[2] = r.z; out[3] = r.w;
}
void h_lane_words(uint32_t k0, uint32_t k1, uint64_t idx, uint32_t site, uint32_t chunk, uint32_t* out) {
  const uint4 r = gjb::make_lane(k0, k1, idx).words(site, chunk);
  out[0] = r.x; out[1] = r.y; out[2] = r.z; out[3] = r.w;
}
void h_u01(const uint32_t* bits, int n, float* out) { for (int i = 0; i < n; ++i) out[i] = gjb::u01(bits[i]); }
void h_box_muller(const uint32_t* b0, const uint32_t* b1, int n, float* out) {
  for (int i = 0; i < n; ++i) { const float2 z = gjb::box_muller(b0[i], b1[i]); out[2 * i] = z.x; out[2 * i + 1] = z.y; }
}
void h_fenc(const float* f, int n, uint32_t* out) { for (int i = 0; i < n; ++i) out[i] = gjb::fenc(f[i]); }
void h_fdec(const uint32_t* e, int n, float* out) { for (int i = 0; i < n; ++i) out[i] = gjb::fdec(e[i]); }
void h_det_exp_q(const float* x, int n, uint64_t* out) { for (int i = 0; i < n; ++i) out[i] = gjb::det_exp_q(x[i]); }
void h_offspring_cnt(const uint64_t* C, int n, uint64_t S, int32_t n_total, double u0, int32_t* out) {
  const double scale = __ddiv_rn((double)n_total, (double)S);
  for (int i = 0; i < n; ++i) out[i] = gjb::offspring_cnt(C[i], S, scale, u0, n_total);
}
double h_resample_u0(uint32_t k0, uint32_t k1, uint64_t key_index) { return gjb::resample_u0(k0, k1, key_index); }
}

// ---- distribution device library (gjb_dist.cuh): log-densities and the lane-stream samplers
#include "gjb_dist.cuh"

extern "C" {
// which: 0 normal(v, loc, scale)  1 uniform(v, lo, hi)  2 exponential(v, rate)  3 half_normal(v, scale)
//        4 gamma(v, a, rate)  5 beta(v, a, b)  6 flip(v, p)  7 bernoulli(v, logit)  8 normal logpdf_r(v, loc, 1/scale, lc)
//        9 cauchy  10 half_cauchy  11 laplace  12 log_normal  13 gumbel (v, loc, scale)  14 weibull(v, k, scale)
//        15 kumaraswamy(v, a, b)  16 logit_normal(v, loc, scale)  17 geometric(v, p)  18 inverse_gamma(v, a, scale)  19 chi2(v, df)
void h_logpdf(int which, const float* v, const float* a, const float* b, int n, float* out) {
  for (int i = 0; i < n; ++i) {
    switch (which) {
      case 0: out[i] = gjb::Normal::logpdf(v[i], a[i], b[i]); break;
      case 1: out[i] = gjb::Uniform::logpdf(v[i], a[i], b[i]); break;
      case 2: out[i] = gjb::Exponential::logpdf(v[i], a[i]); break;
      case 3: out[i] = gjb::HalfNormal::logpdf(v[i], a[i]); break;
      case 4: out[i] = gjb::Gamma::logpdf(v[i], a[i], b[i]); break;
      case 5: out[i] = gjb::Beta::logpdf(v[i], a[i], b[i]); break;
      case 6: out[i] = gjb::Flip::logpdf((int)v[i], a[i]); break;
      case 7: out[i] = gjb::Bernoulli::logpdf((int)v[i], a[i]); break;
      case 8: out[i] = gjb::Normal::logpdf_r(v[i], a[i], 1.0f / b[i], gjb::kHalfLog2Pi + logf(b[i])); break;
      case 9: out[i] = gjb::Cauchy::logpdf(v[i], a[i], b[i]); break;
      case 10: out[i] = gjb::HalfCauchy::logpdf(v[i], a[i], b[i]); break;
      case 11: out[i] = gjb::Laplace::logpdf(v[i], a[i], b[i]); break;
      case 12: out[i] = gjb::LogNormal::logpdf(v[i], a[i], b[i]); break;
      case 13: out[i] = gjb::Gumbel::logpdf(v[i], a[i], b[i]); break;
      case 14: out[i] = gjb::Weibull::logpdf(v[i], a[i], b[i]); break;
      case 15: out[i] = gjb::Kumaraswamy::logpdf(v[i], a[i], b[i]); break;
      case 16: out[i] = gjb::LogitNormal::logpdf(v[i], a[i], b[i]); break;
      case 17: out[i] = gjb::Geometric::logpdf(v[i], a[i]); break;
      case 18: out[i] = gjb::InverseGamma::logpdf(v[i], a[i], b[i]); break;
      case 19: out[i] = gjb::Chi2::logpdf(v[i], a[i]); break;
    }
  }
}
// inverse-CDF samplers on a given draw d (a u01 value; a standard normal for log_normal); numbering as h_logpdf
void h_sample(int which, const float* d, const float* a, const float* b, int n, float* out) {
  for (int i = 0; i < n; ++i) {
    switch (which) {
      case 9: out[i] = gjb::Cauchy::sample(d[i], a[i], b[i]); break;
      case 10: out[i] = gjb::HalfCauchy::sample(d[i], a[i], b[i]); break;
      case 11: out[i] = gjb::Laplace::sample(d[i], a[i], b[i]); break;
      case 12: out[i] = gjb::LogNormal::sample(d[i], a[i], b[i]); break;
      case 13: out[i] = gjb::Gumbel::sample(d[i], a[i], b[i]); break;
      case 14: out[i] = gjb::Weibull::sample(d[i], a[i], b[i]); break;
      case 15: out[i] = gjb::Kumaraswamy::sample(d[i], a[i], b[i]); break;
      case 16: out[i] = gjb::LogitNormal::sample(d[i], a[i], b[i]); break;
      case 17: out[i] = gjb::Geometric::sample(d[i], a[i]); break;
    }
  }
}
void h_categorical(const float* logits, int K, const float* u, int n, int32_t* draws, float* logpdf_of_draw) {
  for (int i = 0; i < n; ++i) {
    draws[i] = gjb::Categorical::sample(u[i], logits, K);
    logpdf_of_draw[i] = gjb::Categorical::logpdf(draws[i], logits, K);
  }
}
// lane-stream rejection samplers: lane = (key, global index), site as in the kernels
void h_inverse_gamma_chi2(uint32_t k0, uint32_t k1, uint64_t idx0, int n, uint32_t site, float a, float b, float* ig_out, float* chi2_out) {
  for (int i = 0; i < n; ++i) {
    const gjb::Lane l = gjb::make_lane(k0, k1, idx0 + (uint64_t)i);
    ig_out[i] = gjb::InverseGamma::sample(l, site, a, b);
    chi2_out[i] = gjb::Chi2::sample(l, site, 2.0f * a);
  }
}
void h_student_t(uint32_t k0, uint32_t k1, uint64_t idx0, int n, uint32_t site, float df, float loc, float scale, float* draw, float* logpdf_of_draw) {
  for (int i = 0; i < n; ++i) {
    const gjb::Lane l = gjb::make_lane(k0, k1, idx0 + (uint64_t)i);
    draw[i] = gjb::StudentT::sample(l, site, df, loc, scale);
    logpdf_of_draw[i] = gjb::StudentT::logpdf(draw[i], df, loc, scale);
  }
}
void h_poisson(uint32_t k0, uint32_t k1, uint64_t idx0, int n, uint32_t site, const float* rate, float* draw, float* logpdf_of_draw) {
  for (int i = 0; i < n; ++i) {
    const gjb::Lane l = gjb::make_lane(k0, k1, idx0 + (uint64_t)i);
    draw[i] = gjb::Poisson::sample(l, site, rate[i]);
    logpdf_of_draw[i] = gjb::Poisson::logpdf(draw[i], rate[i]);
  }
}
void h_gamma_beta(uint32_t k0, uint32_t k1, uint64_t idx0, int n, uint32_t site, float a, float b, float* gamma_out, float* beta_out) {
  for (int i = 0; i < n; ++i) {
    const gjb::Lane l = gjb::make_lane(k0, k1, idx0 + (uint64_t)i);
    gamma_out[i] = gjb::Gamma::sample(l, site, a, b);
    beta_out[i] = gjb::Beta::sample(l, site, a, b);
  }
}
}

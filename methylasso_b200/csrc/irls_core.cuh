// irls_core.cuh — per-row / per-parameter arithmetic of the IRLS loop (device logic, shared with
// the CPU emulation in tests/emu).  Each function keeps the operation order of the reference
// expression it restates; compile with -fmad=false.
#pragma once
#include "hd.h"

namespace ml {

enum : int32_t { MODEL_FAST = 0, MODEL_FULL_SIGNAL = 1, MODEL_FULL_DIFFERENCE = 2 };

// get_variance_impl: core/NormalDistribution.hpp:41-44 (sigma = 1) |
//                    core/BetaBinomialRateDistribution.hpp:64-67
ML_HD double irls_variance(int model, double nobs, double mu, double nu) {
  if (model == MODEL_FAST) {
    const double sigma = 1.0;
    return (sigma * sigma) / nobs;
  }
  return mu * (1 - mu) / nobs * (nobs + nu) / (1 + nu);
}

// restrict_to_support<Logit>: core/logit_link.cpp:19-21 (cwiseMin then cwiseMax)
ML_HD double logit_restrict(double m) {
  const double hi = 1 - 1e-10;
  m = (hi < m) ? hi : m;
  m = (m < 1e-10) ? 1e-10 : m;
  return m;
}

// IRLS working residual z and weight W of one row.
//   first = 1: Posterior::initial_guess (core/Posterior.ipp:3-12) with
//              transform_initial_guess (core/identity_link.cpp:20-23 | core/logit_link.cpp:24-30)
//   first = 0: get_irls_residuals (core/identity_link.cpp:32-36 | core/logit_link.cpp:33-39)
ML_HD void irls_residual(int model, int first, double y, double nobs, double mu_prev, double nu,
                         double& z, double& W) {
  if (model == MODEL_FAST) {
    const double mu = first ? y : mu_prev;
    const double V = irls_variance(model, nobs, mu, nu);
    z = first ? mu : (y - mu);
    W = 1 / V;
  } else {
    const double mu = first ? logit_restrict(y) : mu_prev;
    const double V = irls_variance(model, nobs, mu, nu);
    const double g = mu * (1 - mu);
    z = first ? log(mu / (1 - mu)) : (y - mu) / g;
    W = (g * g) / V;
  }
}

// get_mu<Link> + restrict_to_support: core/identity_link.cpp:12-17 | core/logit_link.cpp:13-21
ML_HD double irls_mu(int model, double eta) {
  if (model == MODEL_FAST) return eta;
  return logit_restrict(1 / (exp(-eta) + 1));
}

ML_HD double log_beta_fn(double x, double y) { return lgamma(x) + lgamma(y) - lgamma(x + y); }  // core/util.cpp:21-23

// one row's term of get_minus_log_pdf_impl: core/NormalDistribution.hpp:46-51 (sigma = 1) |
// core/BetaBinomialRateDistribution.hpp:73-75
ML_HD double irls_minus_log_pdf(int model, double y, double nobs, double mu, double nu, double log2pi) {
  if (model == MODEL_FAST) {
    const double sigma = 1.0;
    const double t = (y - mu) / sigma;
    return nobs / 2. * (t * t + log2pi);
  }
  const double a = log(nobs + 1) + log_beta_fn(nobs * (1 - y) + 1, nobs * y + 1);
  const double b = -log_beta_fn(nobs * y + mu * nu, nobs * (1 - y) + nu * (1 - mu));
  const double c = log_beta_fn(mu * nu, nu * (1 - mu));
  return (a + b) + c;
}

// make_even_bins boundary i: core/binner_helpers.cpp:16
ML_HD double even_boundary(double lower, double size, uint32_t i, uint32_t nbins) {
  return lower + i * size / (double)nbins;
}

// Bin of a position under make_binner_using_bin_design's upper_bound lookup
// (core/binner_helpers.cpp:53-60) over the boundaries of make_even_bins (:13-19): index of the first
// boundary strictly greater than v, minus one; beyond the last boundary -> last bin.  The closed-form
// guess is corrected against the exact boundary expression; duplicate boundaries
// (bins narrower than an ulp) keep the std::map's "last i wins".
ML_HD uint32_t even_bin(double v, double lower, double size, uint32_t nbins) {
  // guess: floor((v-lower)/size*nbins), clamped
  double g = (v - lower) / size * (double)nbins;
  int64_t i = (g > 0) ? (int64_t)g : 0;
  if (i > (int64_t)nbins) i = nbins;
  // want the smallest i in [0, nbins+1] with boundary_i > v  (boundary_{nbins+1} := +inf)
  while (i <= (int64_t)nbins && !(even_boundary(lower, size, (uint32_t)i, nbins) > v)) ++i;
  while (i > 0 && even_boundary(lower, size, (uint32_t)(i - 1), nbins) > v) --i;
  if (i > (int64_t)nbins) return nbins - 1;
  const double key = even_boundary(lower, size, (uint32_t)i, nbins);
  while (i < (int64_t)nbins && even_boundary(lower, size, (uint32_t)(i + 1), nbins) == key) ++i;
  return (uint32_t)(i - 1);
}

}  // namespace ml

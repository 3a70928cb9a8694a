/* oracle/methylasso_oracle.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE (see methylasso_oracle.h).
 *
 * Plain-C restatement of the reference's hot path.  Every function names the reference
 * file:line (relative to /root/reference) whose arithmetic and control flow it follows.
 * Compile with -O2 -ffp-contract=off: the reference is built without FMA contraction
 * (SURVEY.md §7 step 3), so every a*b+c here must round twice as well.
 *
 * Known, documented deviations from the reference (cannot be removed without Eigen):
 *   - Eigen's `.sum()` / `.maxCoeff()` reductions use a packet-unrolled association order;
 *     here they are plain left-to-right sums.  Affected: centring mean
 *     (params_helpers.hpp:27), -log pdf (NormalDistribution.hpp:50,
 *     BetaBinomialRateDistribution.hpp:89).  Relative effect O(1e-16).
 *   - libm `lgamma`/`exp`/`log` stand in for Eigen's array versions (same libm calls in
 *     the scalar path of Eigen 3.3).
 *   - n_params < 2 is refused (reference: undefined behaviour).
 */
#include "methylasso_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

static orc_dp_fn g_dp = 0;

void orc_set_dp(orc_dp_fn fn) { g_dp = fn; }

const char *orc_strerror(int code) {
  switch (code) {
    case ORC_OK: return "ok";
    case ORC_ERR_TOO_FEW_ROWS: return "Need at least 2 data points!";
    case ORC_ERR_SIZE_MISMATCH: return "Input vectors must be of same size!";
    case ORC_ERR_FIRST_DATASET: return "First dataset ID must be 1";
    case ORC_ERR_FIRST_CONDITION: return "First condition ID must be 1";
    case ORC_ERR_FIRST_REPLICATE: return "First replicate ID must be 1";
    case ORC_ERR_NOT_SORTED: return "data is not sorted by dataset and position";
    case ORC_ERR_DATASET_GAP: return "dataset IDs must start at 1 and increase without gaps";
    case ORC_ERR_REPLICATE_GAP: return "replicate ID must increase without gaps within the same condition";
    case ORC_ERR_CONDITION_GAP: return "condition ID must start at 1 and increase without gaps";
    case ORC_ERR_REPLICATE_START: return "replicate ID must start at 1 within the same condition";
    case ORC_ERR_NEGATIVE_COUNT: return "Read counts should be >=0 !";
    case ORC_ERR_NMETH_GT_COV: return "Number of methylated reads cannot be larger than number of total reads!";
    case ORC_ERR_LAMBDA2: return "lambda2 should be >0";
    case ORC_ERR_TWO_CONDITIONS: return "Needs at least two conditions!";
    case ORC_ERR_BAD_REF: return "Invalid reference dataset!";
    case ORC_ERR_NAN: return "NaN found, aborting...";
    case ORC_ERR_TOO_FEW_PARAMS: return "fewer than 2 parameters (reference behaviour undefined)";
    case ORC_ERR_BAD_ARGUMENT: return "bad argument";
    default: return "unknown error";
  }
}

/* ------------------------------------------------------------------------------------------
 * Weighted 1-D fused-lasso DP.  Follows src/core/gfl/gfl_tf.c:184-321 (tf_dp_weight):
 * trivial cases :210-215, first element :231-244, sweep :247-290, last coefficient :294-302,
 * back-pointers :306-311.  Restated with a "knot" view of the same three scratch arrays; the
 * floating-point operation order of every expression is the reference's.
 * ---------------------------------------------------------------------------------------- */
void orc_tf_dp_weight(int n, double *y, double *w, double lam, double *beta,
                      double *x, double *a, double *b, double *tm, double *tp) {
  if (n == 0) return;                                   /* :210 */
  if (n == 1 || lam == 0) {                             /* :211-215 */
    for (int i = 0; i < n; i++) beta[i] = y[i];
    return;
  }
  /* element 0 opens the knot window in the middle of the 2n scratch (:231-240) */
  int left = n - 1, right = n;
  tm[0] = -lam / w[0] + y[0];
  tp[0] = lam / w[0] + y[0];
  x[left] = tm[0];
  x[right] = tp[0];
  a[left] = w[0];
  b[left] = -w[0] * y[0] + lam;
  a[right] = -w[0];
  b[right] = w[0] * y[0] + lam;
  /* seeds of the incoming element (:241-244) */
  double a_lo0 = w[1], b_lo0 = -w[1] * y[1] - lam;
  double a_hi0 = -w[1], b_hi0 = w[1] * y[1] - lam;

  for (int k = 1; k < n - 1; k++) {                     /* :247 */
    double alo = a_lo0, blo = b_lo0;
    int lo = left;
    while (lo <= right) {                               /* :253-258 */
      if (alo * x[lo] + blo > -lam) break;
      alo += a[lo];
      blo += b[lo];
      lo++;
    }
    double ahi = a_hi0, bhi = b_hi0;
    int hi = right;
    while (hi >= lo) {                                  /* :264-269 */
      if (-ahi * x[hi] - bhi < lam) break;
      ahi += a[hi];
      bhi += b[hi];
      hi--;
    }
    tm[k] = (-lam - blo) / alo;                         /* :272 */
    left = lo - 1;
    x[left] = tm[k];
    tp[k] = (lam + bhi) / (-ahi);                       /* :277 */
    right = hi + 1;
    x[right] = tp[k];
    a[left] = alo;                                      /* :282-285 */
    b[left] = blo + lam;
    a[right] = ahi;
    b[right] = bhi + lam;
    a_lo0 = w[k + 1];                                   /* :286-289 */
    b_lo0 = -w[k + 1] * y[k + 1] - lam;
    a_hi0 = -w[k + 1];
    b_hi0 = w[k + 1] * y[k + 1] - lam;
  }
  {                                                     /* :294-302 */
    double alo = a_lo0, blo = b_lo0;
    int lo = left;
    while (lo <= right) {
      if (alo * x[lo] + blo > 0) break;
      alo += a[lo];
      blo += b[lo];
      lo++;
    }
    beta[n - 1] = -blo / alo;
  }
  for (int k = n - 2; k >= 0; k--) {                    /* :306-311 */
    if (beta[k + 1] > tp[k]) beta[k] = tp[k];
    else if (beta[k + 1] < tm[k]) beta[k] = tm[k];
    else beta[k] = beta[k + 1];
  }
}

/* FusedLassoGaussianEstimator::get = soft_threshold(clamp(beta), 0, lambda1)
 * (FusedLassoGaussianEstimator.hpp:71-84, Settings.hpp:32, util.cpp:9-19).
 * std::min/std::max argument order is kept so that NaN behaves as in the reference. */
static double orc_clamp35(double v) {
  const double c = 35.0;
  double m = (-c < v) ? v : -c;  /* std::max(-c, v) */
  return (m < c) ? m : c;        /* std::min(c, m)  */
}
static double orc_soft(double v, double off, double lam1) {
  double val = v - off;
  if (val > 0) {
    double t = val - lam1;
    return (0. < t) ? t : 0.;    /* std::max(0., t) */
  } else {
    double t = val + lam1;
    return (t < 0.) ? t : 0.;    /* std::min(0., t) */
  }
}

/* run DP + clamp + soft threshold into beta[n]; scratch allocated here
 * (GFLLibrary1DDirectImpl.cpp:12-46, fitter_lasso.ipp:2-16). */
static int orc_fit_lasso(int64_t n, const double *y, const double *w, double lambda2, double lambda1,
                         double *beta) {
  if (n > 0x7fffffff) return ORC_ERR_BAD_ARGUMENT;
  double *yy = (double *)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
  double *ww = (double *)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
  double *x = (double *)calloc((size_t)(2 * n + 2), sizeof(double));
  double *a = (double *)calloc((size_t)(2 * n + 2), sizeof(double));
  double *b = (double *)calloc((size_t)(2 * n + 2), sizeof(double));
  double *tm = (double *)calloc((size_t)(n + 1), sizeof(double));
  double *tp = (double *)calloc((size_t)(n + 1), sizeof(double));
  memcpy(yy, y, sizeof(double) * (size_t)n);
  memcpy(ww, w, sizeof(double) * (size_t)n);
  for (int64_t i = 0; i < n; i++) beta[i] = 0;          /* reset(): beta_impl_ zero-filled */
  orc_dp_fn dp = g_dp ? g_dp : orc_tf_dp_weight;
  dp((int)n, yy, ww, lambda2, beta, x, a, b, tm, tp);
  for (int64_t i = 0; i < n; i++) beta[i] = orc_soft(orc_clamp35(beta[i]), 0., lambda1);
  free(yy); free(ww); free(x); free(a); free(b); free(tm); free(tp);
  return ORC_OK;
}

/* src/weighted_1d_flsa.cpp:7-21 */
int orc_weighted_1d_flsa(int64_t n, const double *y, const double *wt, double lambda2,
                         double lambda1, double *beta) {
  if (n < 0) return ORC_ERR_BAD_ARGUMENT;
  if (n == 0) return ORC_OK;
  return orc_fit_lasso(n, y, wt, lambda2, lambda1, beta);
}

/* Observations::check_observations (core/Observations.hpp:29-59) then the MethData count
 * checks (src/MethData.hpp:17-19), in that order: first failing row wins. */
int orc_check_observations(int64_t n, const int32_t *dataset, const int32_t *condition,
                           const int32_t *replicate, const int32_t *pos, const int32_t *nmeth,
                           const int32_t *coverage) {
  if (n <= 1) return ORC_ERR_TOO_FEW_ROWS;
  double last_pos = (double)pos[0];
  uint32_t last_dset = (uint32_t)dataset[0];
  uint32_t last_cond = (uint32_t)condition[0];
  uint32_t last_rep = (uint32_t)replicate[0];
  if (last_dset != 1) return ORC_ERR_FIRST_DATASET;
  if (last_cond != 1) return ORC_ERR_FIRST_CONDITION;
  if (last_rep != 1) return ORC_ERR_FIRST_REPLICATE;
  for (int64_t i = 1; i < n; ++i) {
    if (last_dset == (uint32_t)dataset[i]) {
      if (last_pos >= (double)pos[i]) return ORC_ERR_NOT_SORTED;
    } else {
      if ((uint32_t)dataset[i] != ++last_dset) return ORC_ERR_DATASET_GAP;
      if ((uint32_t)condition[i] == last_cond) {
        if ((uint32_t)replicate[i] != ++last_rep) return ORC_ERR_REPLICATE_GAP;
      } else {
        if ((uint32_t)condition[i] != ++last_cond) return ORC_ERR_CONDITION_GAP;
        last_rep = (uint32_t)replicate[i];
        if (last_rep != 1) return ORC_ERR_REPLICATE_START;
      }
    }
    last_pos = (double)pos[i];
  }
  for (int64_t i = 0; i < n; ++i)
    if (nmeth[i] < 0 || coverage[i] < 0) return ORC_ERR_NEGATIVE_COUNT;
  for (int64_t i = 0; i < n; ++i)
    if (nmeth[i] > coverage[i]) return ORC_ERR_NMETH_GT_COV;
  return ORC_OK;
}

/* ------------------------------------------------------------------------------------------
 * Job state: flat-array restatement of Binner / Design / Estimator / EstimatorGroup.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
  const orc_params *p;
  int64_t N;            /* rows */
  int32_t D, C, E;
  int64_t P;            /* params per estimator */
  int64_t R;            /* bin-rows of the meth binner (block diagonal over datasets) */
  const int32_t *dataset;
  double *y, *nobs;     /* observed = Nmeth/coverage, Nobs = coverage (MethData.hpp:17) */
  int64_t *row2br;      /* data row -> bin-row (Binner::bin_from_data_) */
  int64_t *br2param;    /* bin-row -> bin id (Binner::bin_ids_) */
  int32_t *br2dset;     /* bin-row -> dataset (Binner::dset_), 1-based */
  double *support;      /* P bin starts (BinDesign::starts) */
  double *design;       /* D x E dataset-level design (DataDesign) */
  /* residuals on rows */
  double *z, *w;
  /* binned residuals (shared by all lasso estimators: same binner) and offset's */
  double *zhat, *what;          /* R */
  double *ozhat, *owhat;        /* D */
  /* per estimator */
  double *beta, *old_beta, *betahat, *pw; /* E*P */
  double *off, *old_off;                  /* D */
  double *mu;                             /* N */
  int irls_steps, dampings;
} orc_job;

/* make_even_bins boundary i (core/binner_helpers.cpp:16) */
static double orc_even_boundary(double lower, double size, uint32_t i, uint32_t nbins) {
  return lower + i * size / (double)nbins;
}

/* make_binner_using_bin_design's upper_bound lookup (core/binner_helpers.cpp:53-60) over the
 * std::map built by make_even_bins (:13-19).  Duplicate keys in the map keep the last i. */
static uint32_t orc_even_bin(double v, double lower, double size, uint32_t nbins) {
  uint32_t lo = 0, hi = nbins + 1; /* smallest i in [0,nbins] with boundary_i > v, else nbins+1 */
  while (lo < hi) {
    uint32_t mid = lo + (hi - lo) / 2;
    if (orc_even_boundary(lower, size, mid, nbins) > v) hi = mid; else lo = mid + 1;
  }
  if (lo > nbins) return nbins - 1;               /* map.end(): place in last bin (:56-57) */
  double key = orc_even_boundary(lower, size, lo, nbins);
  while (lo < nbins && orc_even_boundary(lower, size, lo + 1, nbins) == key) lo++;
  return lo - 1;                                   /* it->second - 1 (:59) */
}

static int cmp_i32(const void *pa, const void *pb) {
  int32_t a = *(const int32_t *)pa, b = *(const int32_t *)pb;
  return (a > b) - (a < b);
}

/* build_signal_design / build_difference_design (core/data_design_helpers.cpp:10-28, 30-40) */
static int orc_build_design(orc_job *j, const int32_t *condition) {
  const orc_params *p = j->p;
  j->E = j->C;
  j->design = (double *)calloc((size_t)j->D * (size_t)j->E, sizeof(double));
  int32_t prev = 0;
  for (int64_t i = 0; i < j->N; ++i)
    if (prev != j->dataset[i]) {
      prev = j->dataset[i];
      j->design[(size_t)(j->dataset[i] - 1) * j->E + (condition[i] - 1)] = 1.;
    }
  if (p->model == ORC_MODEL_FULL_DIFFERENCE) {
    uint32_t ref = p->ref;
    /* reference tests `num_dsets < ref || ref == 0`; ref > num_conditions is out-of-range
       block access there (undefined) and refused here with the same code. */
    if ((uint32_t)j->D < ref || ref == 0 || ref > (uint32_t)j->C) return ORC_ERR_BAD_REF;
    for (int32_t d = 0; d < j->D; ++d) {
      double *row = j->design + (size_t)d * j->E;
      for (int32_t c = (int32_t)ref - 1; c >= 1; --c) row[c] = row[c - 1]; /* block(0,1,·,ref-1)=leftCols */
      row[0] = 1.;
    }
  }
  return ORC_OK;
}

/* make_binner_ptr<UniquePositions> / <BinnedGenome> + make_datasets_binner
 * (src/unique_positions.cpp:16-31, src/binned_genome.cpp:16-31, core/binner_helpers.cpp:24-43,
 *  47-95, core/Binner.cpp:95-118). */
static int orc_build_binner(orc_job *j, const int32_t *pos) {
  const orc_params *p = j->p;
  int64_t N = j->N;
  int64_t *bin = (int64_t *)malloc(sizeof(int64_t) * (size_t)N);
  if (p->model == ORC_MODEL_FAST) {
    int32_t *u = (int32_t *)malloc(sizeof(int32_t) * (size_t)N);
    memcpy(u, pos, sizeof(int32_t) * (size_t)N);
    qsort(u, (size_t)N, sizeof(int32_t), cmp_i32);
    int64_t P = 0;
    for (int64_t i = 0; i < N; ++i)
      if (i == 0 || u[i] != u[i - 1]) u[P++] = u[i];
    j->P = P;
    j->support = (double *)malloc(sizeof(double) * (size_t)P);
    for (int64_t k = 0; k < P; ++k) j->support[k] = (double)(uint32_t)u[k]; /* vector<unsigned> */
    for (int64_t i = 0; i < N; ++i) {  /* upper_bound(pos) - 1 == rank of pos */
      int64_t lo = 0, hi = P;
      while (lo < hi) { int64_t mid = (lo + hi) / 2; if (u[mid] > pos[i]) hi = mid; else lo = mid + 1; }
      bin[i] = lo - 1;
    }
    free(u);
  } else {
    double mn = (double)pos[0], mx = (double)pos[0];
    for (int64_t i = 1; i < N; ++i) { double v = (double)pos[i]; if (v < mn) mn = v; if (v > mx) mx = v; }
    double genome_sz = (mx - mn + 1);
    uint32_t nbins = (uint32_t)(genome_sz * p->bf_per_kb / 1000.);  /* binned_genome.cpp:22-23 */
    if (nbins < 2) { free(bin); return ORC_ERR_TOO_FEW_PARAMS; }
    double size = mx - mn;
    j->P = nbins;
    j->support = (double *)malloc(sizeof(double) * (size_t)nbins);
    for (uint32_t k = 0; k < nbins; ++k) j->support[k] = orc_even_boundary(mn, size, k, nbins);
    for (int64_t i = 0; i < N; ++i) bin[i] = orc_even_bin((double)pos[i], mn, size, nbins);
  }
  if (j->P < 2) { free(bin); return ORC_ERR_TOO_FEW_PARAMS; }
  /* per dataset, non-empty bins in increasing order become bin-rows (drop=true, :71-87);
     datasets are stacked block-diagonally (Binner::extend_diagonal). Rows are sorted by
     (dataset,pos) and bin() is monotone in pos, so bin-rows are runs. */
  j->row2br = (int64_t *)malloc(sizeof(int64_t) * (size_t)N);
  j->br2param = (int64_t *)malloc(sizeof(int64_t) * (size_t)N);
  j->br2dset = (int32_t *)malloc(sizeof(int32_t) * (size_t)N);
  int64_t R = 0;
  for (int64_t i = 0; i < N; ++i) {
    if (i == 0 || j->dataset[i] != j->dataset[i - 1] || bin[i] != bin[i - 1]) {
      j->br2param[R] = bin[i];
      j->br2dset[R] = j->dataset[i];
      R++;
    }
    j->row2br[i] = R - 1;
  }
  j->R = R;
  free(bin);
  return ORC_OK;
}

/* bin_irls_residuals (core/binned_residuals_helpers.cpp:9-17) with Binner::bin
 * (core/Binner.cpp:29-36): sequential += in data order. */
static void orc_update_binned_residuals(orc_job *j) {
  memset(j->what, 0, sizeof(double) * (size_t)j->R);
  memset(j->zhat, 0, sizeof(double) * (size_t)j->R);
  memset(j->owhat, 0, sizeof(double) * (size_t)j->D);
  memset(j->ozhat, 0, sizeof(double) * (size_t)j->D);
  for (int64_t i = 0; i < j->N; ++i) {
    double zw = j->z[i] * j->w[i];
    j->what[j->row2br[i]] += j->w[i];
    j->zhat[j->row2br[i]] += zw;
    j->owhat[j->dataset[i] - 1] += j->w[i];
    j->ozhat[j->dataset[i] - 1] += zw;
  }
  for (int64_t r = 0; r < j->R; ++r) {
    double q = j->zhat[r] / j->what[r];
    j->zhat[r] = (j->what[r] == 0) ? j->what[r] : q;    /* select(weight_sum, zhat) */
  }
  for (int32_t d = 0; d < j->D; ++d) {
    double q = j->ozhat[d] / j->owhat[d];
    j->ozhat[d] = (j->owhat[d] == 0) ? j->owhat[d] : q;
  }
}

/* Estimator::update_params for every lasso estimator, then the offset
 * (core/Estimator.ipp:21-29, core/pseudodata_helpers.hpp:27-39, core/fitter_lasso.ipp:2-16,
 *  core/params_helpers.hpp:23-30, core/fitter_mean.hpp:56-61; order Posterior.ipp:87-95). */
static int orc_update_all_params(orc_job *j) {
  const orc_params *p = j->p;
  int64_t P = j->P;
  for (int32_t e = 0; e < j->E; ++e) {
    double *beta = j->beta + (size_t)e * P, *old = j->old_beta + (size_t)e * P;
    double *bh = j->betahat + (size_t)e * P, *pw = j->pw + (size_t)e * P;
    memcpy(old, beta, sizeof(double) * (size_t)P);          /* set_old_params */
    /* pseudoweight = diag(X^T W X), accumulated in bin-row order */
    for (int64_t k = 0; k < P; ++k) { pw[k] = 0; bh[k] = 0; }
    for (int64_t r = 0; r < j->R; ++r) {
      double c = j->design[(size_t)(j->br2dset[r] - 1) * j->E + e];
      if (c != 0) pw[j->br2param[r]] += (c * j->what[r]) * c;
    }
    /* betahat = (W~^+ X^T)(W zhat) + beta, accumulated in bin-row order */
    for (int64_t r = 0; r < j->R; ++r) {
      double c = j->design[(size_t)(j->br2dset[r] - 1) * j->E + e];
      if (c != 0) {
        int64_t k = j->br2param[r];
        double inv = 1. / pw[k];
        if (!isfinite(inv)) inv = 0;                        /* pseudo-inverse (:33-34) */
        bh[k] += (inv * c) * (j->what[r] * j->zhat[r]);
      }
    }
    for (int64_t k = 0; k < P; ++k) bh[k] = bh[k] + beta[k];
    if (p->lambda2 <= 0) return ORC_ERR_LAMBDA2;             /* fitter_lasso.ipp:4 */
    int rc = orc_fit_lasso(P, bh, pw, p->lambda2, p->lambda1, beta);
    if (rc) return rc;
    /* center_params */
    double sbw = 0, sw = 0;
    for (int64_t k = 0; k < P; ++k) { sbw += beta[k] * pw[k]; sw += pw[k]; }
    double mean = sbw / sw;
    for (int64_t k = 0; k < P; ++k) beta[k] = beta[k] - mean;
  }
  /* offset: identity design, Mean fitter, no centring (offset.cpp:26-32, offset.hpp:37) */
  for (int32_t d = 0; d < j->D; ++d) {
    j->old_off[d] = j->off[d];
    double pwd = (1. * j->owhat[d]) * 1.;
    double inv = 1. / pwd;
    if (!isfinite(inv)) inv = 0;
    j->off[d] = (inv * 1.) * (j->owhat[d] * j->ozhat[d]) + j->off[d];
  }
  j->irls_steps++;
  return ORC_OK;
}

/* damp_all_params (Posterior.ipp:115-123; fitter_lasso.hpp:61-68; fitter_mean.hpp:63-70) */
static void orc_damp_all_params(orc_job *j, double wnew) {
  size_t n = (size_t)j->E * (size_t)j->P;
  for (size_t k = 0; k < n; ++k) j->beta[k] = wnew * j->beta[k] + (1 - wnew) * j->old_beta[k];
  for (int32_t d = 0; d < j->D; ++d) j->off[d] = wnew * j->off[d] + (1 - wnew) * j->old_off[d];
  j->dampings++;
}

/* Posterior::get_mu (Posterior.ipp:74-82): eta = sum_e B^T X_e beta_e + offset
 * (EstimatorGroup.hpp:85-91, estimate_helpers.hpp:29-32, Binner.cpp:38-48), then the link
 * inverse and the support restriction (identity_link.cpp:12-17, logit_link.cpp:13-21). */
static void orc_get_mu(const orc_job *j, double *mu) {
  const int logit = j->p->model != ORC_MODEL_FAST;
  for (int64_t i = 0; i < j->N; ++i) {
    int64_t r = j->row2br[i];
    int32_t d = j->br2dset[r] - 1;
    double eta = 0;
    for (int32_t e = 0; e < j->E; ++e) {
      double c = j->design[(size_t)d * j->E + e];
      double phi = (c != 0) ? (0. + c * j->beta[(size_t)e * j->P + j->br2param[r]]) : 0.;
      eta += phi;
    }
    eta = eta + ((0. + 1. * j->off[d]));
    double m;
    if (logit) {
      m = 1 / (exp(-eta) + 1);
      double hi = 1 - 1e-10;
      m = (hi < m) ? hi : m;          /* cwiseMin(1-1e-10) */
      m = (m < 1e-10) ? 1e-10 : m;    /* cwiseMax(1e-10)   */
    } else {
      m = eta;
    }
    mu[i] = m;
  }
}

static double orc_log_beta(double x, double y) { return lgamma(x) + lgamma(y) - lgamma(x + y); } /* util.cpp:21-23 */

/* get_variance_impl (NormalDistribution.hpp:41-44 | BetaBinomialRateDistribution.hpp:64-67) */
static double orc_variance(const orc_job *j, double nobs, double mu) {
  if (j->p->model == ORC_MODEL_FAST) { double sigma = 1.; return (sigma * sigma) / nobs; }
  double nu = j->p->hyper;
  return mu * (1 - mu) / nobs * (nobs + nu) / (1 + nu);
}

/* get_minus_log_posterior (Posterior.ipp:63-69): likelihood (NormalDistribution.hpp:46-51 |
 * BetaBinomialRateDistribution.hpp:69-90) + TV prior per lasso estimator
 * (GFLLibrary1D.hpp:62-67); recomputes mu from the current params. */
static double orc_minus_log_posterior(const orc_job *j, double *mu_scratch) {
  orc_get_mu(j, mu_scratch);
  double lik = 0;
  if (j->p->model == ORC_MODEL_FAST) {
    double sigma = 1.;
    double cst = log(2 * M_PI * sigma * sigma);
    for (int64_t i = 0; i < j->N; ++i) {
      double t = (j->y[i] - mu_scratch[i]) / sigma;
      lik += j->nobs[i] / 2. * (t * t + cst);
    }
  } else {
    double nu = j->p->hyper;
    for (int64_t i = 0; i < j->N; ++i) {
      double N = j->nobs[i], d = j->y[i], m = mu_scratch[i];
      double a = log(N + 1) + orc_log_beta(N * (1 - d) + 1, N * d + 1);
      double b = -orc_log_beta(N * d + m * nu, N * (1 - d) + nu * (1 - m));
      double c = orc_log_beta(m * nu, nu * (1 - m));
      lik += (a + b) + c;
    }
  }
  double pri = 0;
  for (int32_t e = 0; e < j->E; ++e) {
    const double *beta = j->beta + (size_t)e * j->P;
    double tv = 0;
    for (int64_t k = 1; k < j->P; ++k) tv += fabs(beta[k] - beta[k - 1]);
    double lam = j->p->lambda2;
    pri += lam * tv - (double)(uint64_t)(j->P - 2) * log(lam) + (double)(uint64_t)(j->P - 1) * log(2);
  }
  pri += 0; /* offset: Fitter<Mean>::get_minus_log_prior == 0 */
  return lik + pri;
}

/* get_irls_residuals + update_all_binned_residuals (Posterior.ipp:49-51;
 * identity_link.cpp:32-36 | logit_link.cpp:33-39) */
static void orc_refresh_residuals(orc_job *j, const double *mu) {
  const int logit = j->p->model != ORC_MODEL_FAST;
  for (int64_t i = 0; i < j->N; ++i) {
    double V = orc_variance(j, j->nobs[i], mu[i]);
    if (logit) {
      double g = mu[i] * (1 - mu[i]);
      j->z[i] = (j->y[i] - mu[i]) / g;
      j->w[i] = (g * g) / V;
    } else {
      j->z[i] = j->y[i] - mu[i];
      j->w[i] = 1 / V;
    }
  }
  orc_update_binned_residuals(j);
}

/* Posterior::iterate (Posterior.ipp:22-59) with adapt_all_params (:99-110) and
 * AdaptiveDampingPolicy<1,5,10> (posterior_policies.hpp:57-68). */
static int orc_iterate(orc_job *j, double *mu, double *mu_old, double *scratch, int *converged) {
  const orc_params *p = j->p;
  *converged = 0;
  double old_mlogp = orc_minus_log_posterior(j, scratch);
  double mlogp = old_mlogp;
  for (uint32_t i = 0; i < p->max_iter; ++i) {
    int rc = orc_update_all_params(j);
    if (rc) return rc;
    {
      unsigned dstep = 0;
      double m = orc_minus_log_posterior(j, scratch);
      for (;;) {
        unsigned ds = dstep++;
        if (ds >= 10) break;
        if (!(m > old_mlogp)) break;
        orc_damp_all_params(j, 1 - 5 / 10.);
        m = orc_minus_log_posterior(j, scratch);
      }
    }
    orc_get_mu(j, mu);
    mlogp = orc_minus_log_posterior(j, scratch);
    if (isnan(mlogp)) return ORC_ERR_NAN;
    if (i > 0) {
      double diff = 0; /* (mu-mu_old).abs().maxCoeff() */
      for (int64_t k = 0; k < j->N; ++k) { double t = fabs(mu[k] - mu_old[k]); if (t > diff || k == 0) diff = t; }
      if (diff < p->tol_val) { *converged = 1; break; }
    }
    memcpy(mu_old, mu, sizeof(double) * (size_t)j->N);
    old_mlogp = mlogp;
    orc_refresh_residuals(j, mu);
  }
  if (p->max_iter == 1) *converged = 1;
  (void)mlogp;
  return ORC_OK;
}

static void orc_job_free(orc_job *j) {
  free(j->y); free(j->nobs); free(j->row2br); free(j->br2param); free(j->br2dset); free(j->support);
  free(j->design); free(j->z); free(j->w); free(j->zhat); free(j->what); free(j->ozhat); free(j->owhat);
  free(j->beta); free(j->old_beta); free(j->betahat); free(j->pw); free(j->off); free(j->old_off);
  free(j->mu);
}

/* signal_detection_fast_R / signal_detection_R / difference_detection_R
 * (src/signal_detection.cpp:11-47, 49-80; src/difference_detection.cpp:13-45) around
 * methylation_detection (src/methylation_detection.hpp:22-54) and maximize_posterior
 * (core/Posterior.ipp:204-252). */
int orc_detect(const orc_params *p, int64_t n, const int32_t *dataset, const int32_t *condition,
               const int32_t *replicate, const int32_t *pos, const int32_t *nmeth,
               const int32_t *coverage, orc_result *out) {
  memset(out, 0, sizeof(*out));
  if (p->optimize_lambda2 || p->optimize_hyper) return ORC_ERR_BAD_ARGUMENT;
  if (p->model < 0 || p->model > 2) return ORC_ERR_BAD_ARGUMENT;
  int rc = orc_check_observations(n, dataset, condition, replicate, pos, nmeth, coverage);
  if (rc) return rc;
  orc_job J;
  memset(&J, 0, sizeof(J));
  orc_job *j = &J;
  j->p = p; j->N = n; j->dataset = dataset;
  j->D = dataset[n - 1];                       /* dset.tail<1>() (data_design_helpers.cpp:23) */
  j->C = condition[n - 1];
  if (p->model == ORC_MODEL_FULL_DIFFERENCE && j->C <= 1) return ORC_ERR_TWO_CONDITIONS;
  rc = orc_build_design(j, condition);
  if (rc) { orc_job_free(j); return rc; }
  j->y = (double *)malloc(sizeof(double) * (size_t)n);
  j->nobs = (double *)malloc(sizeof(double) * (size_t)n);
  for (int64_t i = 0; i < n; ++i) { j->y[i] = (double)nmeth[i] / (double)coverage[i]; j->nobs[i] = (double)coverage[i]; }
  rc = orc_build_binner(j, pos);
  if (rc) { orc_job_free(j); return rc; }
  size_t EP = (size_t)j->E * (size_t)j->P;
  j->z = (double *)malloc(sizeof(double) * (size_t)n);
  j->w = (double *)malloc(sizeof(double) * (size_t)n);
  j->zhat = (double *)malloc(sizeof(double) * (size_t)j->R);
  j->what = (double *)malloc(sizeof(double) * (size_t)j->R);
  j->ozhat = (double *)malloc(sizeof(double) * (size_t)j->D);
  j->owhat = (double *)malloc(sizeof(double) * (size_t)j->D);
  j->beta = (double *)calloc(EP, sizeof(double));
  j->old_beta = (double *)calloc(EP, sizeof(double));
  j->betahat = (double *)calloc(EP, sizeof(double));
  j->pw = (double *)calloc(EP, sizeof(double));
  j->off = (double *)calloc((size_t)j->D, sizeof(double));
  j->old_off = (double *)calloc((size_t)j->D, sizeof(double));
  j->mu = (double *)malloc(sizeof(double) * (size_t)n);
  double *mu_old = (double *)malloc(sizeof(double) * (size_t)n);
  double *scratch = (double *)malloc(sizeof(double) * (size_t)n);
  double *mu_outer_old = (double *)malloc(sizeof(double) * (size_t)n);

  /* Posterior::initial_guess (Posterior.ipp:3-18; identity_link.cpp:20-23 | logit_link.cpp:24-30) */
  {
    const int logit = p->model != ORC_MODEL_FAST;
    for (int64_t i = 0; i < n; ++i) {
      double mu0 = j->y[i];
      if (logit) {
        double hi = 1 - 1e-10;
        mu0 = (hi < mu0) ? hi : mu0;
        mu0 = (mu0 < 1e-10) ? 1e-10 : mu0;
      }
      double V0 = orc_variance(j, j->nobs[i], mu0);
      if (logit) {
        double g = mu0 * (1 - mu0);
        j->z[i] = log(mu0 / (1 - mu0));
        j->w[i] = (g * g) / V0;
      } else {
        j->z[i] = mu0;
        j->w[i] = 1 / V0;
      }
    }
    orc_update_binned_residuals(j);
  }

  int converged = 0, outer = 0;
  if (p->nouter > 1) {
    for (uint32_t i = 0; i < p->nouter; ++i) {
      outer++;
      rc = orc_iterate(j, j->mu, mu_old, scratch, &converged);
      if (rc) break;
      if (i > 0) {
        double diff = 0;
        for (int64_t k = 0; k < n; ++k) { double t = fabs(j->mu[k] - mu_outer_old[k]); if (t > diff || k == 0) diff = t; }
        if (diff < p->tol_val) { converged = converged && 1; break; }
        else converged = 0;
      }
      memcpy(mu_outer_old, j->mu, sizeof(double) * (size_t)n);
      (void)orc_minus_log_posterior(j, scratch);   /* :226, value unused when nothing is optimised */
      if (!(isfinite(p->lambda2) && isfinite(p->hyper))) { converged = 0; break; }
    }
  } else {
    outer = 1;
    rc = orc_iterate(j, j->mu, mu_old, scratch, &converged);
  }
  if (rc == ORC_OK) {
    out->n_rows = n; out->n_params = j->P; out->n_est = j->E; out->n_dsets = j->D;
    out->converged = converged; out->irls_steps = j->irls_steps; out->dampings = j->dampings;
    out->outer_iters = outer;
    out->hyper = (p->model == ORC_MODEL_FAST) ? 1. : p->hyper;
    out->last_mlogp = orc_minus_log_posterior(j, scratch);
    out->mu = j->mu; j->mu = 0;
    out->offset = j->off; j->off = 0;
    out->lambda2 = (double *)malloc(sizeof(double) * (size_t)j->E);
    for (int32_t e = 0; e < j->E; ++e) out->lambda2[e] = p->lambda2;
    /* get_estimate_dataframe (src/methylation_helpers.hpp:14-49): start = int(support) */
    out->estimate_id = (int32_t *)malloc(sizeof(int32_t) * EP);
    out->start = (double *)malloc(sizeof(double) * EP);
    for (int32_t e = 0; e < j->E; ++e)
      for (int64_t k = 0; k < j->P; ++k) {
        out->estimate_id[(size_t)e * j->P + k] = e + 1;
        out->start[(size_t)e * j->P + k] = (double)(int32_t)j->support[k];
      }
    out->beta = j->beta; j->beta = 0;
    out->betahat = j->betahat; j->betahat = 0;
    out->pseudoweight = j->pw; j->pw = 0;
  }
  free(mu_old); free(scratch); free(mu_outer_old);
  orc_job_free(j);
  return rc;
}

void orc_result_free(orc_result *r) {
  free(r->mu); free(r->lambda2); free(r->offset); free(r->estimate_id);
  free(r->start); free(r->beta); free(r->betahat); free(r->pseudoweight);
  memset(r, 0, sizeof(*r));
}

/* segment_methylation_helper (src/methylation_helpers.cpp:6-72) */
int64_t orc_segment_helper(int64_t n, const int32_t *idx, const int32_t *start,
                           const double *beta, const double *betahat, const double *pseudoweight,
                           double tol_val, int32_t *seg_idx, int32_t *seg_no,
                           uint32_t *start_min, uint32_t *start_max, double *seg_beta,
                           double *seg_betahat, double *seg_pw, double *seg_sd) {
  if (n <= 0) return -ORC_ERR_BAD_ARGUMENT;
  int64_t S = 0;      /* closed segments */
  uint32_t curr_idx = (uint32_t)idx[0], curr_segment = 0;
  double curr_beta = beta[0];
  double curr_residual = pseudoweight[0] * (betahat[0] - beta[0]);
  double curr_pw = pseudoweight[0];
  double curr_var = curr_residual * (betahat[0] - beta[0]);
  int64_t nstart = 0;
  start_min[nstart++] = (uint32_t)start[0];
#define ORC_CLOSE(END)                                                                         \
  do {                                                                                         \
    start_max[S] = (uint32_t)(END);                                                            \
    seg_no[S] = (int32_t)(++curr_segment);                                                     \
    seg_idx[S] = (int32_t)curr_idx;                                                            \
    seg_beta[S] = curr_beta;                                                                   \
    seg_pw[S] = curr_pw;                                                                       \
    seg_betahat[S] = curr_residual / curr_pw + curr_beta;                                      \
    seg_sd[S] = curr_var / curr_pw - curr_residual * curr_residual / (curr_pw * curr_pw);      \
    S++;                                                                                       \
  } while (0)
  for (int64_t i = 1; i < n; ++i) {
    if ((uint32_t)idx[i] == curr_idx) {
      if (fabs(beta[i] - curr_beta) > tol_val) {
        start_min[nstart++] = (uint32_t)start[i];
        ORC_CLOSE(start[i] - 1);
        curr_beta = beta[i];
        curr_residual = pseudoweight[i] * (betahat[i] - beta[i]);
        curr_var = curr_residual * (betahat[i] - beta[i]);
        curr_pw = pseudoweight[i];
      } else {
        curr_residual += pseudoweight[i] * (betahat[i] - beta[i]);
        curr_var += pseudoweight[i] * (betahat[i] - beta[i]) * (betahat[i] - beta[i]);
        curr_pw += pseudoweight[i];
      }
    } else {
      start_min[nstart++] = (uint32_t)start[i];
      ORC_CLOSE(start[i - 1]);
      curr_idx = (uint32_t)idx[i];
      curr_segment = 0;
      curr_beta = beta[i];
      curr_residual = pseudoweight[i] * (betahat[i] - beta[i]);
      curr_var = curr_residual * (betahat[i] - beta[i]);
      curr_pw = pseudoweight[i];
    }
  }
  ORC_CLOSE(start[n - 1]);
#undef ORC_CLOSE
  for (int64_t s = 0; s < S; ++s) seg_sd[s] = sqrt(seg_sd[s]);
  return S;
}

/* R/fast_diff.R:15-28: estimate 1 is the reference; others become differences to it. */
int orc_fast_diff_combine(int64_t P, int32_t E, const double *beta, const double *betahat,
                          const double *pw, double *ob, double *obh, double *opw) {
  if (P <= 0 || E < 1) return ORC_ERR_BAD_ARGUMENT;
  for (int64_t k = 0; k < P; ++k) { ob[k] = beta[k]; obh[k] = betahat[k]; opw[k] = pw[k]; }
  for (int32_t e = 1; e < E; ++e)
    for (int64_t k = 0; k < P; ++k) {
      size_t i = (size_t)e * P + k;
      ob[i] = beta[i] - beta[k];
      double s = pw[i] + pw[k];
      obh[i] = (betahat[i] * pw[i] - betahat[k] * pw[k]) / s;
      opw[i] = s;
      if (s == 0) obh[i] = 0;
    }
  return ORC_OK;
}

/* R/call.R:43-68 (segment_methylation_singlechr), the per-segment call table:
 *   ov = foverlaps(mdata[pos, pos+1], segmentation[start, end])   -> CpG p belongs to segment s of its estimate
 *                                                                    iff start_s - 1 <= p <= end_s (type "any")
 *   is.oversize = pos - shift(pos) > max_distance within a segment (:59-60)
 *   segment_id += cumsum(is.oversize) over the estimate (:61)
 *   start = min(pos), end = max(pos) + 2, width = end - start + 1 by (estimate, segment_id) (:64-65)
 *   beta = mean(Nmeth/coverage), coverage = sum(coverage), pseudoweight = pseudoweight[1],
 *   num.cpgs = length(unique(pos)), std = sd(Nmeth/coverage), NA -> 0 (:68-72)
 * Rows are the per-position data pooled per condition (:43-46), sorted by (estimate, pos); segments are the
 * helper's table sorted by (estimate, start).  mean() and sd() follow R's own algorithms (summary.c real_mean,
 * cov.c: long double accumulation, mean refined by a second pass, sum of squares about that mean / (n-1)).
 * R itself is not in this image: this function is a restatement of the R lines, NOT pinned against R. */
int64_t orc_call_table(int64_t n, const int32_t *est_id, const int32_t *pos, const int32_t *nmeth,
                       const int32_t *cov, int64_t S, const int32_t *seg_est, const int32_t *seg_id,
                       const uint32_t *seg_start, const uint32_t *seg_end, const double *seg_pw,
                       double max_distance, int32_t *o_est, int32_t *o_seg_id, uint32_t *o_start,
                       uint32_t *o_end, double *o_width, double *o_beta, double *o_cov, double *o_pw,
                       int32_t *o_ncpg, double *o_std) {
  if (n <= 0 || S <= 0) return -ORC_ERR_BAD_ARGUMENT;
  int64_t out = 0, r0 = 0;
  int32_t cur_est = seg_est[0] - 1;
  int64_t shift = 0;                      /* cumsum(is.oversize) within the current estimate */
  for (int64_t s = 0; s < S; ++s) {
    if (seg_est[s] != cur_est) { cur_est = seg_est[s]; shift = 0; r0 = 0; }
    if (seg_start[s] > seg_end[s]) continue;                        /* sg <- segmentation[start <= end] */
    /* rows of this estimate with start-1 <= pos <= end (rows are sorted: advance a cursor, then scan) */
    while (r0 < n && (est_id[r0] < cur_est || (est_id[r0] == cur_est && (int64_t)pos[r0] < (int64_t)seg_start[s] - 1))) ++r0;
    int64_t lo = r0, hi = lo;
    while (hi < n && est_id[hi] == cur_est && (int64_t)pos[hi] <= (int64_t)seg_end[s]) ++hi;
    int64_t a = lo;
    while (a < hi) {
      int64_t b = a + 1;                                             /* part [a, b): no oversize gap inside */
      while (b < hi && !((double)pos[b] - (double)pos[b - 1] > max_distance)) ++b;
      if (a > lo) ++shift;                                           /* this part starts at an oversize gap */
      const int64_t m = b - a;
      long double sum = 0;
      double csum = 0;
      for (int64_t i = a; i < b; ++i) { sum += (long double)((double)nmeth[i] / (double)cov[i]); csum += (double)cov[i]; }
      long double mean = sum / m;
      if (isfinite((double)mean)) {
        long double t = 0;
        for (int64_t i = a; i < b; ++i) t += ((long double)((double)nmeth[i] / (double)cov[i]) - mean);
        mean += t / m;
      }
      double sd = 0;
      if (m >= 2) {
        long double ss = 0;
        for (int64_t i = a; i < b; ++i) {
          const long double d = (long double)((double)nmeth[i] / (double)cov[i]) - mean;
          ss += d * d;
        }
        sd = sqrt((double)(ss / (m - 1)));
      }
      o_est[out] = cur_est;
      o_seg_id[out] = (int32_t)(seg_id[s] + shift);
      o_start[out] = (uint32_t)pos[a];
      o_end[out] = (uint32_t)pos[b - 1] + 2u;
      o_width[out] = (double)o_end[out] - (double)o_start[out] + 1.0;
      o_beta[out] = (double)mean;
      o_cov[out] = csum;
      o_pw[out] = seg_pw[s];
      o_ncpg[out] = (int32_t)m;
      o_std[out] = sd;
      ++out;
      a = b;
    }
  }
  return out;
}

/* ---- R/call.R:315-317 and :360: wilcox.test(x, y, conf.int = F)$p.value and p.adjust(p, 'fdr') ----------------
 * Restated from the published algorithms of R's stats package (wilcox.test.default, two-sided, mu = 0,
 * exact = NULL, correct = TRUE; pwilcox / cwilcox of nmath/wilcox.c; p.adjust method "BH").  R is not in this
 * image; tests/test_wilcoxon.py pins this restatement against scipy.stats.mannwhitneyu (exact and asymptotic with
 * continuity) and scipy.stats.false_discovery_control, which implement the same tests independently. */
static int orc_cmp_double(const void *a, const void *b) {
  const double x = *(const double *)a, y = *(const double *)b;
  return x < y ? -1 : x > y;
}

/* counts[k] = number of ways the rank-sum statistic of m-vs-n takes the value k (k = 0..m*n), by the recurrence
 * w(k; i, j) = w(k - j; i - 1, j) + w(k; i, j - 1) that nmath/wilcox.c memoises */
static double *orc_cwilcox_table(int m, int n) {
  const int K = m * n + 1;
  double *A = (double *)calloc((size_t)(n + 1) * K, sizeof(double));
  if (!A) return NULL;
  for (int j = 0; j <= n; ++j) A[(size_t)j * K] = 1.0;                 /* w(0; 0, j) = 1 */
  for (int i = 1; i <= m; ++i)
    for (int j = 1; j <= n; ++j) {
      double *row = A + (size_t)j * K, *left = A + (size_t)(j - 1) * K;
      for (int k = i * j; k >= 0; --k) row[k] = (k >= j ? row[k - j] : 0.0) + left[k];
    }
  double *out = (double *)malloc(sizeof(double) * K);
  if (out) memcpy(out, A + (size_t)n * K, sizeof(double) * K);
  free(A);
  return out;
}

static double orc_choose(int n, int k) {        /* exact in double for the sizes used (n < 100) up to rounding */
  double c = 1.0;
  if (k > n - k) k = n - k;
  for (int i = 1; i <= k; ++i) c = c * (double)(n - k + i) / (double)i;
  return floor(c + 0.5);
}

/* pwilcox(q, m, n, lower_tail) of nmath/wilcox.c */
static double orc_pwilcox(double q, int m, int n, int lower_tail, const double *w) {
  q = floor(q + 1e-7);
  if (q < 0.0) return lower_tail ? 0.0 : 1.0;
  if (q >= (double)m * n) return lower_tail ? 1.0 : 0.0;
  const double c = orc_choose(m + n, n);
  double p = 0.0;
  if (q <= (double)m * n / 2) {
    for (int i = 0; i <= (int)q; ++i) p += w[i] / c;
    return lower_tail ? p : 1.0 - p;
  }
  q = (double)m * n - q;
  for (int i = 0; i < (int)q; ++i) p += w[i] / c;
  return lower_tail ? 1.0 - p : p;
}

/* returns 0 and sets *pval (NaN when undefined, e.g. all values equal) and *stat = W; 1 if a sample is empty */
int orc_wilcox_test(int64_t nx, const double *x, int64_t ny, const double *y, double *stat, double *pval) {
  /* x <- x[!is.na(x)], y <- y[!is.na(y)] */
  double *v = (double *)malloc(sizeof(double) * (size_t)(nx + ny + 1));
  int64_t mx = 0, my = 0;
  for (int64_t i = 0; i < nx; ++i) if (!isnan(x[i])) v[mx++] = x[i];
  for (int64_t i = 0; i < ny; ++i) if (!isnan(y[i])) v[mx + my++] = y[i];
  if (mx < 1 || my < 1) { free(v); *pval = NAN; *stat = NAN; return 1; }
  const int64_t N = mx + my;
  double *s = (double *)malloc(sizeof(double) * (size_t)N);
  memcpy(s, v, sizeof(double) * (size_t)N);
  qsort(s, (size_t)N, sizeof(double), orc_cmp_double);
  /* W = sum of the (average) ranks of x - nx (nx + 1) / 2; tie term sum(t^3 - t) */
  double W = 0.0, tie_term = 0.0;
  int ties = 0;
  for (int64_t i = 0; i < N;) {
    int64_t j = i;
    while (j < N && s[j] == s[i]) ++j;
    const double t = (double)(j - i);
    if (j - i > 1) ties = 1;
    tie_term += t * t * t - t;
    i = j;
  }
  for (int64_t i = 0; i < mx; ++i) {
    /* rank = (#less) + (#equal + 1) / 2, by binary search in the sorted pool */
    int64_t lo = 0, hi = N;
    while (lo < hi) { const int64_t mid = (lo + hi) / 2; if (s[mid] < v[i]) lo = mid + 1; else hi = mid; }
    int64_t first = lo;
    hi = N;
    while (lo < hi) { const int64_t mid = (lo + hi) / 2; if (s[mid] <= v[i]) lo = mid + 1; else hi = mid; }
    W += (double)first + ((double)(lo - first) + 1.0) / 2.0;
  }
  W -= (double)mx * ((double)mx + 1.0) / 2.0;
  *stat = W;
  const double dx = (double)mx, dy = (double)my;
  if (mx < 50 && my < 50 && !ties) {
    double *w = orc_cwilcox_table((int)mx, (int)my);
    double p = W > dx * dy / 2 ? orc_pwilcox(W - 1, (int)mx, (int)my, 0, w) : orc_pwilcox(W, (int)mx, (int)my, 1, w);
    free(w);
    *pval = fmin(2 * p, 1.0);
  } else {
    double z = W - dx * dy / 2;
    const double sigma = sqrt((dx * dy / 12) * ((dx + dy + 1) - tie_term / ((dx + dy) * (dx + dy - 1))));
    const double corr = (z > 0) ? 0.5 : (z < 0 ? -0.5 : 0.0);
    z = (z - corr) / sigma;
    const double pl = 0.5 * erfc(-z / sqrt(2.0)), pu = 0.5 * erfc(z / sqrt(2.0));
    *pval = 2 * fmin(pl, pu);
  }
  free(s);
  free(v);
  return 0;
}

/* p.adjust(p, method = "fdr"): NaN entries are left out of n and of the ranking, as R leaves NA out */
void orc_p_adjust_fdr(int64_t n, const double *p, double *q) {
  int64_t *o = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n + 1));
  int64_t lp = 0;
  for (int64_t i = 0; i < n; ++i) { q[i] = NAN; if (!isnan(p[i])) o[lp++] = i; }
  /* order(p, decreasing = TRUE), stable: insertion into a sorted index by merge sort */
  int64_t *tmp = (int64_t *)malloc(sizeof(int64_t) * (size_t)(lp + 1));
  for (int64_t w = 1; w < lp; w *= 2)
    for (int64_t a = 0; a < lp; a += 2 * w) {
      int64_t i = a, mid = a + w < lp ? a + w : lp, j = mid, end = a + 2 * w < lp ? a + 2 * w : lp, k = a;
      while (i < mid && j < end) tmp[k++] = (p[o[j]] > p[o[i]]) ? o[j++] : o[i++];
      while (i < mid) tmp[k++] = o[i++];
      while (j < end) tmp[k++] = o[j++];
      memcpy(o + a, tmp + a, sizeof(int64_t) * (size_t)(end - a));
    }
  double run = INFINITY;
  for (int64_t r = 0; r < lp; ++r) {                /* i = lp - r; cummin(n / i * p[o]) */
    const double val = (double)lp / (double)(lp - r) * p[o[r]];
    if (val < run) run = val;
    q[o[r]] = run < 1.0 ? run : 1.0;
  }
  free(tmp);
  free(o);
}

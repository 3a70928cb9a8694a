/* oracle/methylasso_oracle.h — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C CPU restatement of MethyLasso's segmentation hot path (the IRLS loop of
 * `methylation_detection` wrapped around the weighted 1-D fused-lasso DP).  Only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library; the product (libmethylasso_b200.so) never does.
 *
 * Parity status: PINNED against the reference itself.  The reference ships no tests,
 * fixtures or golden vectors, and R / Rcpp / Eigen / GSL / Boost are not in
 * this image, so the pin is built from the reference's own sources compiled in place
 * (oracle/Makefile, nothing is copied into this repository):
 *   (a) oracle/_ref/libtfdp_ref.so: the DP kernel src/core/gfl/gfl_tf.c compiled verbatim;
 *       the DP restatement is bit-identical to it (tests/test_oracle.py);
 *   (b) oracle/_ref/libmethylasso_ref.so: the reference's unmodified C++ path above the DP
 *       (signal_detection*.cpp, difference_detection.cpp, methylation_helpers.cpp,
 *       weighted_1d_flsa.cpp, core/Posterior.ipp, Estimator.ipp, Binner.cpp, ... 19 translation
 *       units) compiled against small stand-in headers for Rcpp / Eigen / Boost
 *       (oracle/stubs: NOT those libraries, see each header) behind C entry points
 *       (oracle/ref_driver.cpp).  tests/test_oracle_pin.py runs it live against this
 *       restatement and against tests/golden/reference_fixtures.npz (generated from it by
 *       tests/golden/make_golden.py): integer outputs, convergence flags, refusal messages and
 *       the segment helper identical; FAST pseudodata bit-identical; every float within 1.3e-12
 *       (Eigen's two-lane vectorised sums vs left-to-right sums here);
 *   (c) the two known answers recorded in BASELINE.md §2.
 * What the stand-in headers cannot vouch for is real Eigen's association order inside .sum()
 * and sparse products (reproduced from Eigen's documented default traversal; effect ~1e-16
 * relative) and Boost's brent_find_minima (optimize_hyper = TRUE, outside the hot path: refused).
 */
#ifndef METHYLASSO_ORACLE_H
#define METHYLASSO_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* same numeric values as include/methylasso_b200.h (kept in sync by tests/test_abi.py) */
enum {
  ORC_OK = 0,
  ORC_ERR_TOO_FEW_ROWS = 1,        /* Observations.hpp:31 */
  ORC_ERR_SIZE_MISMATCH = 2,       /* Observations.hpp:33-34 ; weighted_1d_flsa.cpp:9 */
  ORC_ERR_FIRST_DATASET = 3,       /* Observations.hpp:40 */
  ORC_ERR_FIRST_CONDITION = 4,     /* Observations.hpp:41 */
  ORC_ERR_FIRST_REPLICATE = 5,     /* Observations.hpp:42 */
  ORC_ERR_NOT_SORTED = 6,          /* Observations.hpp:46 */
  ORC_ERR_DATASET_GAP = 7,         /* Observations.hpp:48 */
  ORC_ERR_REPLICATE_GAP = 8,       /* Observations.hpp:50 */
  ORC_ERR_CONDITION_GAP = 9,       /* Observations.hpp:52 */
  ORC_ERR_REPLICATE_START = 10,    /* Observations.hpp:54 */
  ORC_ERR_NEGATIVE_COUNT = 11,     /* MethData.hpp:18 */
  ORC_ERR_NMETH_GT_COV = 12,       /* MethData.hpp:19 */
  ORC_ERR_LAMBDA2 = 13,            /* fitter_lasso.ipp:4 */
  ORC_ERR_TWO_CONDITIONS = 14,     /* difference_detection.cpp:28 */
  ORC_ERR_BAD_REF = 15,            /* data_design_helpers.cpp:35 */
  ORC_ERR_NAN = 16,                /* Posterior.ipp:36 */
  ORC_ERR_TOO_FEW_PARAMS = 18,     /* reference: undefined behaviour; we refuse */
  ORC_ERR_BAD_ARGUMENT = 19
};

enum { ORC_MODEL_FAST = 0, ORC_MODEL_FULL_SIGNAL = 1, ORC_MODEL_FULL_DIFFERENCE = 2 };

typedef struct {
  double lambda2, max_lambda2, lambda1, tol_val, bf_per_kb, hyper;
  uint32_t max_iter, nouter, ref;
  int32_t model;
  int32_t optimize_lambda2, optimize_hyper; /* must be 0: out of scope */
} orc_params;

typedef struct {
  int64_t n_rows;      /* N */
  int64_t n_params;    /* P: unique positions (FAST) or Nbins (FULL) */
  int32_t n_est;       /* E */
  int32_t n_dsets;     /* D */
  int32_t converged;
  int32_t irls_steps;  /* total update_all_params calls */
  int32_t dampings;    /* total damp_all_params calls */
  int32_t outer_iters;
  double hyper;
  double last_mlogp;
  double *mu;          /* N */
  double *lambda2;     /* E */
  double *offset;      /* D */
  int32_t *estimate_id;/* E*P */
  double *start, *beta, *betahat, *pseudoweight; /* E*P each */
} orc_result;

/* DP used inside orc_detect / orc_weighted_1d_flsa.  Default: the restatement
 * orc_tf_dp_weight.  Tests and the CPU baseline may point it at the reference's own
 * tf_dp_weight from oracle/_ref/libtfdp_ref.so. */
typedef void (*orc_dp_fn)(int n, double *y, double *w, double lam, double *beta,
                          double *x, double *a, double *b, double *tm, double *tp);
void orc_set_dp(orc_dp_fn fn);

void orc_tf_dp_weight(int n, double *y, double *w, double lam, double *beta,
                      double *x, double *a, double *b, double *tm, double *tp);

int orc_weighted_1d_flsa(int64_t n, const double *y, const double *wt, double lambda2,
                         double lambda1, double *beta);

int orc_check_observations(int64_t n, const int32_t *dataset, const int32_t *condition,
                           const int32_t *replicate, const int32_t *pos, const int32_t *nmeth,
                           const int32_t *coverage);

int orc_detect(const orc_params *p, int64_t n, const int32_t *dataset, const int32_t *condition,
               const int32_t *replicate, const int32_t *pos, const int32_t *nmeth,
               const int32_t *coverage, orc_result *out);
void orc_result_free(orc_result *r);

/* segment_methylation_helper: returns number of segments written (<= n), or <0 on error.
 * Outputs must be caller-allocated with capacity n. */
int64_t orc_segment_helper(int64_t n, const int32_t *estimate_id, const int32_t *start,
                           const double *beta, const double *betahat, const double *pseudoweight,
                           double tol_val, int32_t *seg_estimate_id, int32_t *seg_id,
                           uint32_t *seg_start, uint32_t *seg_end, double *seg_beta,
                           double *seg_betahat, double *seg_pseudoweight, double *seg_stdev);

/* R/call.R:43-68: per-segment call table (gap split, shrunk boundaries, mean / sd of the pooled per-position
 * methylation) from pooled rows sorted by (estimate, pos) and the helper's segments.  Outputs need capacity n.
 * Returns the number of rows, <0 on error.  Restates R code; R is not in this image (not pinned against R). */
int64_t orc_call_table(int64_t n, const int32_t *est_id, const int32_t *pos, const int32_t *nmeth,
                       const int32_t *cov, int64_t S, const int32_t *seg_est, const int32_t *seg_id,
                       const uint32_t *seg_start, const uint32_t *seg_end, const double *seg_pw,
                       double max_distance, int32_t *o_est, int32_t *o_seg_id, uint32_t *o_start,
                       uint32_t *o_end, double *o_width, double *o_beta, double *o_cov, double *o_pw,
                       int32_t *o_ncpg, double *o_std);

/* R/call.R:315-317: wilcox.test(x, y, conf.int = F)$p.value (two-sided; exact when both samples have fewer than 50
 * values and there are no ties, normal approximation with continuity and tie correction otherwise); R/call.R:360:
 * p.adjust(p, 'fdr').  Restated from R's published algorithms and pinned against scipy in tests/test_wilcoxon.py. */
int orc_wilcox_test(int64_t nx, const double *x, int64_t ny, const double *y, double *stat, double *pval);
void orc_p_adjust_fdr(int64_t n, const double *p, double *q);

/* R/fast_diff.R:15-28 combine, restated on the estimates table of one job. */
int orc_fast_diff_combine(int64_t n_params, int32_t n_est, const double *beta, const double *betahat,
                          const double *pseudoweight, double *out_beta, double *out_betahat,
                          double *out_pseudoweight);

const char *orc_strerror(int code);

#ifdef __cplusplus
}
#endif
#endif

/* methylasso_b200.h — C ABI of the B200-native MethyLasso segmentation engine.
 *
 * This is the drop-in boundary for the reference's Rcpp module "MethyLasso_cpp"
 * (/root/reference/src/MethyLasso_cpp.cpp:9-31, loaded by R/zzz.R:1-3).  Each entry point
 * below names the reference interface it replaces.  Plain pointers and sizes only; the engine
 * owns the GPU, its streams and all device memory.  No entry point has a CPU fallback: without
 * a CUDA device `ml_engine_create` fails and nothing else can be called.
 *
 * Conventions
 *   - Every function returns 0 (ML_OK) or an error code; `ml_last_error()` gives the message of
 *     the last failure on the calling thread.  No exception crosses the boundary.
 *   - Batched calls take JOBS: one job = one chromosome of one data frame = one native call of the
 *     reference (R/genome_wide.R:67-70).  A job that fails validation gets its own status code
 *     and does not fail the batch (mirrors foreach(.errorhandling="remove"),
 *     R/genome_wide.R:43).
 *   - Inputs are the six int32 columns the R layer passes (signal_detection.cpp:13-15), rows of a
 *     job sorted by (dataset, pos); `job_ptr[njobs+1]` holds CSR-style row offsets.
 *   - Outputs are caller-allocated; sizes come from ml_result_job_sizes().
 *   - One engine = one GPU = one host thread at a time.
 */
#ifndef METHYLASSO_B200_H
#define METHYLASSO_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define ML_ABI_VERSION 1

#if defined(__GNUC__)
#define ML_API __attribute__((visibility("default")))
#else
#define ML_API
#endif

/* error / status codes: one per reference error site */
enum {
  ML_OK = 0,
  ML_ERR_TOO_FEW_ROWS = 1,      /* core/Observations.hpp:31  "Need at least 2 data points!" */
  ML_ERR_SIZE_MISMATCH = 2,     /* core/Observations.hpp:33-34; src/weighted_1d_flsa.cpp:9 */
  ML_ERR_FIRST_DATASET = 3,     /* core/Observations.hpp:40 */
  ML_ERR_FIRST_CONDITION = 4,   /* core/Observations.hpp:41 */
  ML_ERR_FIRST_REPLICATE = 5,   /* core/Observations.hpp:42 */
  ML_ERR_NOT_SORTED = 6,        /* core/Observations.hpp:46 */
  ML_ERR_DATASET_GAP = 7,       /* core/Observations.hpp:48 */
  ML_ERR_REPLICATE_GAP = 8,     /* core/Observations.hpp:50 */
  ML_ERR_CONDITION_GAP = 9,     /* core/Observations.hpp:52 */
  ML_ERR_REPLICATE_START = 10,  /* core/Observations.hpp:54 */
  ML_ERR_NEGATIVE_COUNT = 11,   /* src/MethData.hpp:18 */
  ML_ERR_NMETH_GT_COV = 12,     /* src/MethData.hpp:19 */
  ML_ERR_LAMBDA2 = 13,          /* core/fitter_lasso.ipp:4 */
  ML_ERR_TWO_CONDITIONS = 14,   /* src/difference_detection.cpp:28 */
  ML_ERR_BAD_REF = 15,          /* core/data_design_helpers.cpp:35 */
  ML_ERR_NAN = 16,              /* core/Posterior.ipp:36 */
  ML_ERR_TOO_FEW_PARAMS = 18,   /* reference: undefined behaviour (size()-2 underflow); refused */
  ML_ERR_BAD_ARGUMENT = 19,
  ML_ERR_CUDA = 100,            /* CUDA runtime failure (message has the details) */
  ML_ERR_NO_DEVICE = 101
};

/* model = which instantiation of methylation_detection<> (src/methylation_detection.hpp:22-54) */
enum {
  ML_MODEL_FAST = 0,            /* Normal/Identity/UniquePositions: signal_detection_fast_R
                                   (src/signal_detection.cpp:11-47) */
  ML_MODEL_FULL_SIGNAL = 1,     /* BetaBinomialRate/Logit/BinnedGenome: signal_detection_R (:49-80) */
  ML_MODEL_FULL_DIFFERENCE = 2  /* same with the difference design: difference_detection_R
                                   (src/difference_detection.cpp:13-45) */
};

/* The arguments of the three Rcpp entry points (src/MethyLasso_cpp.cpp:13-26). */
typedef struct {
  double lambda2;      /* fusion penalty                                   (R default 25) */
  double max_lambda2;  /* unused while optimize_lambda2 == 0; FAST sets lambda2 + 1 */
  double lambda1;      /* always 0 in the reference (signal_detection.cpp:22,58) */
  double tol_val;      /* convergence tolerance on max |delta mu|          (0.01) */
  double bf_per_kb;    /* FULL: basis functions per kb                     (50) */
  double hyper;        /* FULL: beta-binomial nu                           (100) */
  uint32_t max_iter;   /* IRLS steps per iterate()          (FAST 1, FULL 30) */
  uint32_t nouter;     /* outer iterations                  (FAST 1, FULL 20) */
  uint32_t ref;        /* FULL_DIFFERENCE: reference condition, 1-based */
  int32_t model;       /* ML_MODEL_* */
  int32_t optimize_lambda2; /* must be 0 (FALSE at every reference call site) */
  int32_t optimize_hyper;   /* must be 0 */
} ml_params;

typedef struct ml_engine ml_engine;
typedef struct ml_batch ml_batch;    /* validated, indexed input resident in HBM */
typedef struct ml_result ml_result;  /* outputs of one detect call, resident in HBM until copied */

/* ---- engine ------------------------------------------------------------------------------ */
ML_API int ml_abi_version(void);
ML_API const char *ml_last_error(void);
ML_API const char *ml_strerror(int code);
/* device < 0: use the current CUDA device */
ML_API int ml_engine_create(int device, ml_engine **out);
ML_API void ml_engine_destroy(ml_engine *eng);
/* engine knobs: "flsa_chunk", "flsa_warm", "flsa_sequential" (1 = start-to-end DP, one thread
 * per series: the bit-faithful verification mode), "profile" (1 = time every kernel),
 * "profile_reserve_events" (pre-create that many timing events), "pool_headroom_percent" (grow the engine's
 * memory pool now by that share of what it holds, so that steady-state calls never map memory),
 * "host_threads" (threads of ml_batch_upload's scan of the dataset column, default 4), "want_mu" (default 1; 0: a
 * one-step FAST fit does not store the per-row mu, which nothing in the reference's R layer reads - asking for it
 * from such a result is an error); the forward pass of the FLSA solver: "flsa_route" (0 = by the series itself, 1 =
 * speculative route only, 2 = scan route only), "flsa_scan_lambda" (route 0: penalties from this value on take the scan
 * route, default 26), "flsa_chain_limit" (route 0: chunks a walker redoes before it hands its series over to the scan
 * route, default 6; 0 = never), "flsa_scan_chunk_small" / "flsa_scan_chunk_big" (chunk lengths of the scan route);
 * "stream_priority" (re-create the engine's stream at that priority level, right after creation only),
 * "profile_timeline" (1 = keep the start / end of every profiled launch: ml_engine_profile_spans).  Every route gives the
 * reference's result: bit for bit on count data with an integer penalty, within rounding elsewhere. */
ML_API int ml_engine_set(ml_engine *eng, const char *key, long long value);
/* number of kernel launches issued by this engine so far */
ML_API long long ml_engine_launch_count(const ml_engine *eng);
/* raw CUDA stream (cudaStream_t) the engine launches on, for event timing by the caller */
ML_API void *ml_engine_stream(ml_engine *eng);
ML_API int ml_engine_synchronize(ml_engine *eng);
/* Launch on a caller-owned CUDA stream (cudaStream_t) instead of the engine's own, so the caller's
 * CUDA events bracket the engine's work; NULL restores the engine's stream. */
ML_API int ml_engine_set_stream(ml_engine *eng, void *cuda_stream);
/* Per-kernel timing (CUDA events on the launching stream).  Enable with
 * ml_engine_set(eng, "profile", 1); entries accumulate until ml_engine_profile_reset. */
ML_API int ml_engine_profile_count(ml_engine *eng);
ML_API int ml_engine_profile_get(ml_engine *eng, int i, char *name, int name_cap, double *total_ms,
                                 long long *launches);
/* Algorithmic bytes accumulated by entry i when its work is data-dependent (the FLSA chunk kernels: DP steps
 * completed x 32 B, warm-up steps x 16 B); -1 when the kernel's traffic follows from the batch shape alone. */
ML_API int ml_engine_profile_work(ml_engine *eng, int i, double *bytes);
ML_API int ml_engine_profile_reset(ml_engine *eng);
/* Timeline of the profiled launches since the last reset (enable with ml_engine_set(eng, "profile_timeline", 1) next to
 * "profile"): start / end of each launch in ms on the clock of one event of the process, so that the engines of a pool
 * can be laid side by side.  With names == NULL returns the number of spans; else fills up to `count` spans from
 * `first` (names: count x name_cap bytes) and returns how many. */
ML_API int ml_engine_profile_spans(ml_engine *eng, int first, int count, char *names, int name_cap, float *t0_ms,
                                   float *t1_ms);

/* Page-locked host memory for large outputs that a caller fills again and again: a download into pageable memory goes
 * through the driver's bounce buffers at a fraction of the PCIe rate.  Allocate ONCE and reuse - page-locking costs
 * milliseconds per 100 MB and ml_host_free synchronises the device.  Plain memory for the caller. */
ML_API int ml_host_alloc(uint64_t bytes, void **out);
ML_API void ml_host_free(void *p);

/* ---- weighted_1d_flsa  (src/weighted_1d_flsa.cpp:7-21; Rcpp module line
 *      src/MethyLasso_cpp.cpp:28) --------------------------------------------------------------
 * beta = soft_threshold(clamp(FLSA(y, wt, lambda2), +-35), lambda1).  Host buffers. */
ML_API int ml_weighted_1d_flsa(ml_engine *eng, int64_t n, const double *y, const double *wt,
                        double lambda2, double lambda1, double *beta);
/* many independent series in one call: ptr[nseries+1] element offsets, one lambda2 per series */
ML_API int ml_weighted_1d_flsa_batch(ml_engine *eng, int nseries, const int64_t *ptr, const double *y,
                              const double *wt, const double *lambda2, double lambda1,
                              double *beta);
/* same on device-resident buffers (y, wt, beta are device pointers; ptr/lambda2 host) */
ML_API int ml_weighted_1d_flsa_batch_device(ml_engine *eng, int nseries, const int64_t *ptr,
                                     const double *d_y, const double *d_wt, const double *lambda2,
                                     double lambda1, double *d_beta);

/* ---- detection: signal_detection_fast_singlechr / signal_detection_singlechr /
 *      difference_detection_singlechr (src/MethyLasso_cpp.cpp:13-26), batched over jobs --------- */
/* Make the rows of the jobs resident: pos / nmeth / coverage are copied to the GPU (6 to 12 bytes per row); dataset,
 * condition and replicate are factor codes that the reference reads where the dataset changes only
 * (core/Observations.hpp:36-56; core/data_design_helpers.cpp:17-23), so they are reduced on the host to one entry
 * per run of equal dataset codes (a streaming compare over the dataset column while the DMA runs) and never
 * uploaded.  Validation of the rows and indexing run on the device inside ml_batch_detect.  status may be NULL; it
 * is zero-filled. */
ML_API int ml_batch_upload(ml_engine *eng, int njobs, const int64_t *job_ptr, const int32_t *dataset,
                    const int32_t *condition, const int32_t *replicate, const int32_t *pos,
                    const int32_t *nmeth, const int32_t *coverage, ml_batch **out, int *status);
/* ml_batch_upload in two halves, for callers that overlap the host work of one batch with the transfers of another
 * (several engines of one GPU): ml_batch_stage is host work only - the run table, and Nmeth / coverage packed into the
 * engine's page-locked staging area with one or two bytes per count when every count of the batch fits (else the
 * columns travel as they are) - and ml_batch_commit copies pos and the packed counts (6 or 8 bytes per row instead
 * of 12) and unpacks them on the device.  The caller's columns must stay valid until the commit; one staged batch per
 * engine at a time.  ml_batch_upload is the two calls in a row. */
ML_API int ml_batch_stage(ml_engine *eng, int njobs, const int64_t *job_ptr, const int32_t *dataset,
                   const int32_t *condition, const int32_t *replicate, const int32_t *pos,
                   const int32_t *nmeth, const int32_t *coverage, ml_batch **out);
ML_API int ml_batch_commit(ml_engine *eng, ml_batch *b);
/* The same for a caller that already has the runs (R: rows are keyed by (chr, dataset, pos), so
 * data[, .(.N, condition[1], replicate[1]), by = .(chr, dataset)] is the run table): datasets of job j are entries
 * dset_ptr[j] .. dset_ptr[j+1]-1 of dset_row (job-relative first row; 0 for the first), dset_condition and
 * dset_replicate; the dataset codes are taken to be 1, 2, ... in that order. */
ML_API int ml_batch_upload_runs(ml_engine *eng, int njobs, const int64_t *job_ptr, const int64_t *dset_ptr,
                         const int64_t *dset_row, const int32_t *dset_condition, const int32_t *dset_replicate,
                         const int32_t *pos, const int32_t *nmeth, const int32_t *coverage, ml_batch **out);
ML_API void ml_batch_free(ml_batch *b);
/* Run methylation_detection on every valid job of a resident batch.  `params` is one struct
 * shared by all jobs.  status[njobs] (may be NULL) receives per-job codes. */
ML_API int ml_batch_detect(ml_engine *eng, ml_batch *b, const ml_params *params, ml_result **out,
                    int *status);
/* upload + detect in one call (host buffers in, result handle out) */
ML_API int ml_detect_batch(ml_engine *eng, int njobs, const int64_t *job_ptr, const int32_t *dataset,
                    const int32_t *condition, const int32_t *replicate, const int32_t *pos,
                    const int32_t *nmeth, const int32_t *coverage, const ml_params *params,
                    ml_result **out, int *status);

ML_API int ml_result_njobs(const ml_result *r);
/* sizes of job j: rows N, params per estimator P, estimators E, datasets D, status */
ML_API int ml_result_job_sizes(const ml_result *r, int job, int64_t *n_rows, int64_t *n_params,
                        int32_t *n_est, int32_t *n_dsets, int32_t *status);
/* scalars of job j: converged, irls steps, dampings, outer iterations, hyper, last -log posterior */
ML_API int ml_result_job_info(const ml_result *r, int job, int32_t *converged, int32_t *irls_steps,
                       int32_t *dampings, int32_t *outer_iters, double *hyper, double *mlogp);
/* The List the Rcpp entry points return (src/methylation_detection.hpp:47-53,
 * src/methylation_helpers.hpp:25-49): mu[N], lambda2[E], offset[D] and the estimates table of E*P
 * rows (estimate_id, start, beta, betahat, pseudoweight).  Any pointer may be NULL to skip it. */
ML_API int ml_result_copy_job(ml_engine *eng, const ml_result *r, int job, double *mu, double *lambda2,
                       double *offset, int32_t *estimate_id, double *start, double *beta,
                       double *betahat, double *pseudoweight);
/* All jobs at once, concatenated in job order (failed jobs contribute nothing): one D2H per column. */
ML_API int ml_result_copy_all(ml_engine *eng, const ml_result *r, double *mu, double *offset,
                       int32_t *estimate_id, double *start, double *beta, double *betahat,
                       double *pseudoweight);
/* Same, but the copies are only ENQUEUED (on the engine's copy stream): the buffers must be pinned
 * host memory and must not be read before ml_engine_synchronize(eng) returns.  Later calls on the
 * same engine (ml_result_segments, ...) overlap with the transfer.  The result must stay alive until
 * the synchronisation. */
ML_API int ml_result_copy_all_async(ml_engine *eng, const ml_result *r, double *mu, double *offset,
                                    int32_t *estimate_id, double *start, double *beta, double *betahat,
                                    double *pseudoweight);
ML_API void ml_result_free(ml_result *r);

/* R/fast_diff.R:15-28 applied in place to a FAST result with >= 2 estimators: estimates of
 * conditions 2.. become differences to condition 1 (beta - beta_ref, pooled betahat, summed
 * pseudoweight). */
ML_API int ml_result_fast_diff_combine(ml_engine *eng, ml_result *r);

/* ---- segment_methylation_helper (src/methylation_helpers.cpp:6-72; Rcpp module line
 *      src/MethyLasso_cpp.cpp:21) --------------------------------------------------------------
 * Host table in (n rows sorted by estimate_id, start), caller-allocated outputs of capacity n;
 * *n_segments receives the number of segments written. */
ML_API int ml_segment_methylation_helper(ml_engine *eng, int64_t n, const int32_t *estimate_id,
                                  const int32_t *start, const double *beta, const double *betahat,
                                  const double *pseudoweight, double tol_val, int64_t *n_segments,
                                  int32_t *seg_estimate_id, int32_t *seg_id, uint32_t *seg_start,
                                  uint32_t *seg_end, double *seg_beta, double *seg_betahat,
                                  double *seg_pseudoweight, double *seg_stdev);
/* The same on the estimates of a resident result (no D2H/H2D of the per-CpG table): segments of
 * all valid jobs concatenated, seg_job[] gives the job of each.  Query with outputs NULL to get
 * *n_segments, then call again with buffers of that capacity. */
ML_API int ml_result_segments(ml_engine *eng, ml_result *r, double tol_val, int64_t *n_segments,
                       int32_t *seg_job, int32_t *seg_estimate_id, int32_t *seg_id,
                       uint32_t *seg_start, uint32_t *seg_end, double *seg_beta,
                       double *seg_betahat, double *seg_pseudoweight, double *seg_stdev);

/* PMD std fusion of R/call.R:93-95 on the resident segments: per (job, estimate)
 * fused = weighted_1d_flsa(std, width, pmd_lambda2) with std = segment stdev (NA -> 0) and
 * width = end - start + 1, then segment_methylation_helper on (estimate_id, start, beta = fused,
 * betahat = std, pseudoweight).  fused_std (may be NULL) has one entry per segment of
 * ml_result_segments(tol_val); the seg_* outputs describe the fused-std segments.  Query with NULL
 * outputs for *n_segments first.
 * NOTE: the reference feeds this step the per-segment sd that its R layer computes from the raw data
 * (R/call.R:57-68); here the helper's own weighted stdev stands in until that R step is native. */
ML_API int ml_result_pmd_segments(ml_engine *eng, ml_result *r, double tol_val, double pmd_lambda2,
                                  int64_t *n_segments, double *fused_std, int32_t *seg_job,
                                  int32_t *seg_estimate_id, int32_t *seg_id, uint32_t *seg_start,
                                  uint32_t *seg_end, double *seg_fused_std, double *seg_std,
                                  double *seg_pseudoweight);

/* The per-segment call table of R/call.R:43-68 (segment_methylation_singlechr, before the rule engine), resident:
 * the helper's segments (tol_val) are matched with the pooled per-position data of their condition as
 * foverlaps does (CpG p belongs to segment [start, end] iff start - 1 <= p <= end), split at gaps larger than
 * max_distance (:59-61, segment ids shift by the number of gaps so far in the condition), shrunk to
 * start = min(pos), end = max(pos) + 2, width = end - start + 1 (:64-65), and summarised: beta = mean(Nmeth/coverage),
 * coverage = sum, pseudoweight = the segment's, num_cpgs, std = sd(Nmeth/coverage) with NA -> 0 (:68-72; R's
 * two-pass long-double mean / sd reproduced with double-double sums).  Needs a FAST result with max_iter = 1
 * (there betahat * pseudoweight are the pooled methylated counts exactly).  Every output may be NULL; query
 * *n_rows first.  Rows are ordered by (job, estimate, start). */
ML_API int ml_result_call_table(ml_engine *eng, ml_result *r, double tol_val, double max_distance, int64_t *n_rows,
                                int32_t *job, int32_t *estimate_id, int32_t *segment_id, uint32_t *start,
                                uint32_t *end, double *width, double *beta, double *coverage,
                                double *pseudoweight, int32_t *num_cpgs, double *std);

/* R/call.R:93-95 on that call table, i.e. with the reference's own inputs: fused = weighted_1d_flsa(std, width,
 * pmd_lambda2) per (job, condition), then segment_methylation_helper on (estimate_id, start, beta = fused,
 * betahat = std, pseudoweight).  fused_std (may be NULL) has one entry per call-table row. */
ML_API int ml_result_call_pmd_segments(ml_engine *eng, ml_result *r, double tol_val, double max_distance,
                                       double pmd_lambda2, int64_t *n_segments, double *fused_std,
                                       int32_t *seg_job, int32_t *seg_estimate_id, int32_t *seg_id,
                                       uint32_t *seg_start, uint32_t *seg_end, double *seg_fused_std,
                                       double *seg_std, double *seg_pseudoweight);

/* The same call table from host tables (any model): pooled rows (estimate_id, pos, Nmeth, coverage) sorted by
 * (estimate_id, pos) as R/call.R:43-46 builds them, and the segment table of segment_methylation_helper.  Outputs
 * need capacity n. */
ML_API int ml_segment_call_table(ml_engine *eng, int64_t n, const int32_t *estimate_id, const int32_t *pos,
                                 const int32_t *nmeth, const int32_t *coverage, int64_t n_seg,
                                 const int32_t *seg_estimate_id, const int32_t *seg_id, const uint32_t *seg_start,
                                 const uint32_t *seg_end, const double *seg_pseudoweight, double max_distance,
                                 int64_t *n_rows, int32_t *o_estimate_id, int32_t *o_segment_id, uint32_t *o_start,
                                 uint32_t *o_end, double *o_width, double *o_beta, double *o_coverage,
                                 double *o_pseudoweight, int32_t *o_num_cpgs, double *o_std);

/* ---- region statistics applied by R/call.R to the path's segments ------------------------------------------------
 * wilcox.test(x, y, conf.int = F)$p.value of R/call.R:315-317 for many groups at once: group g compares
 * x[x_ptr[g] .. x_ptr[g+1]) with y[y_ptr[g] .. y_ptr[g+1]) (host buffers).  Two-sided rank-sum test as R's
 * wilcox.test.default runs it: NaN values dropped; exact distribution when both samples have fewer than 50 values and
 * there are no ties, normal approximation with continuity and tie correction otherwise.  pval[g] is NaN where R
 * fails or returns NaN (an empty sample; all values equal), which R/call.R:317 turns into 1.  stat (may be NULL)
 * receives W. */
ML_API int ml_wilcoxon_groups(ml_engine *eng, int64_t n_groups, const int64_t *x_ptr, const double *x,
                              const int64_t *y_ptr, const double *y, double *pval, double *stat);
/* p.adjust(p, method = 'fdr') of R/call.R:360 (Benjamini-Hochberg); NaN entries stay NaN and are not counted. */
ML_API int ml_p_adjust_fdr(ml_engine *eng, int64_t n, const double *p, double *q);

/* ---- call_differences (R/call.R:254-333 per chromosome; :348-365 adds the FDR and the threshold filter) ------------
 * On a resident FAST result of >= 2 conditions to which ml_result_fast_diff_combine has been applied
 * (= difference_detection_fast), and the batch it was computed from (the raw rows must still be resident): per job
 * and non-reference condition, the difference estimates are segmented (tol_val), matched with the pooled per-position
 * data of the condition and of the reference condition (coverage summed, beta = mean of the replicates'
 * Nmeth / coverage), split at gaps larger than max_distance, grouped while neighbouring differences agree within
 * min_delta_diff and both exceed min_diff, and every group is summarised: mean / sd of the per-position values on
 * each side, Cohen's d, the Wilcoxon rank-sum p-value (NA -> 1) and the coverage score.  ref_condition is the
 * 1-based condition code R/call.R:267 compares `condition` with.  Rows are ordered by (job, estimate_id,
 * segment_id) (the reference sorts by (start, end) within a chromosome, :331).  Query with NULL outputs for
 * *n_rows first; any output may be NULL.  The FDR of :360 is ml_p_adjust_fdr over the pval column of all jobs. */
ML_API int ml_result_call_differences(ml_engine *eng, const ml_batch *b, ml_result *r, int ref_condition,
                                      double min_diff, double min_delta_diff, double tol_val, double max_distance,
                                      int64_t *n_rows, int32_t *job, int32_t *condition, int32_t *estimate_id,
                                      int32_t *segment_id, uint32_t *start, uint32_t *end, double *width,
                                      double *beta, double *std, double *coverage, int32_t *num_cpgs,
                                      double *beta_ref, double *std_ref, double *coverage_ref,
                                      int32_t *num_cpgs_ref, double *pseudoweight, double *diff, double *cohen_d,
                                      double *pval, double *coverage_score);

/* ---- LMR / UMR annotation (R/call.R:75-91, inside segment_methylation_singlechr) ------------------------------------
 * On the resident call table of ml_result_call_table(tol_val, max_distance), one entry per call-table row, in its
 * order: is.admissible (:77), is.V among the admissible segments of a condition (:78), is.flank (:79), flank.dist,
 * flank.beta and is.sep among the V-shaped and flanking segments (:80-82), and the category after :83-86.  Logical
 * columns are three-valued as in R: 1 TRUE, 0 FALSE, -1 NA; numeric NA is NaN; category 0 = NA, 1 = "LMR", 2 = "UMR".
 * shift() in those statements walks the rows the statement selects (of one condition of one chromosome; of one
 * chromosome for :79), as data.table evaluates them.  The valleys (:100-133) are ml_result_call_valleys; the PMD calls
 * (:143-221) still run in R on this table.  params == NULL: the reference's defaults.  Any output may be NULL. */
typedef struct {
  double min_width, max_distance, flank_width, flank_dist_max, lmr_max_beta, flank_beta_min, umr_max_beta,
      umr_std_threshold, umr_large_width, umr_large_density;    /* 30, 500, 300, 5000, 0.5, 0.5, 0.1, 0.1, 500, 0.03 */
  int32_t min_num_cpgs, flank_num_cpgs;                         /* 4, 10 */
} ml_lmr_params;
ML_API int ml_result_call_lmr_umr(ml_engine *eng, ml_result *r, double tol_val, double max_distance,
                                  const ml_lmr_params *params, int64_t *n_rows, int8_t *is_admissible, int8_t *is_v,
                                  int8_t *is_flank, double *flank_dist, double *flank_beta, int8_t *is_sep,
                                  int8_t *category);

/* ---- valleys / DMV (R/call.R:100-133, inside segment_methylation_singlechr) -----------------------------------------
 * On the resident call table and its LMR / UMR categories (computed here with `lmr`, NULL = defaults): the "UMR-DMV"
 * table `new_valleys` as it stands after :129 - runs of consecutive segments with beta <= valley_max_beta that are
 * not separated by more than pmd_valley_min_width (:102-110; low.id.width / low.id.beta per run, the mean and the sums
 * of doubles with R's long double accumulation), plus the UMRs that overlap none of them (:113-116), merged while
 * their distance stays within pmd_valley_min_width (:119-124), kept when the merged region contains a run wider than
 * pmd_valley_min_width ("DMV") and is itself admissible (:125-126).  Rows are ordered by (job, condition, start).
 * beta is NaN where R has NA (a merged region that holds a UMR above valley_max_beta, :114-115).
 * Per call-table row (n_rows of them, the order of ml_result_call_table): row_in_valley = 1 when the row overlaps a
 * valley and :132-133 replaces it; low_id_width / low_id_beta = the columns of :106 (NaN = NA on high rows).
 * :134-164 continue in ml_result_call_pmd_split.  params == NULL: the reference's
 * defaults with max_distance as passed.  Query with NULL outputs for the counts first; any output may be NULL. */
typedef struct {
  double valley_max_beta, pmd_valley_min_width, max_distance;   /* 0.1, 5000, 500 */
  int32_t min_num_cpgs, reserved;                               /* 4, 0 */
} ml_valley_params;
ML_API int ml_result_call_valleys(ml_engine *eng, ml_result *r, double tol_val, double max_distance,
                                  const ml_lmr_params *lmr, const ml_valley_params *params, int64_t *n_valleys,
                                  int32_t *job, int32_t *condition, uint32_t *start, uint32_t *end, double *width,
                                  double *beta, double *coverage, double *pseudoweight, int32_t *num_cpgs,
                                  int64_t *n_rows, int8_t *row_in_valley, double *low_id_width, double *low_id_beta);

/* ---- valleys into the call table, PMD candidate regions (R/call.R:131-168) --------------------------------------------
 * Continues ml_result_call_valleys on the same resident tables.  Merged table = seg_call after :140: the call-table
 * rows that overlap no valley plus the valleys, ordered by (job, condition, start); m_category 0 NA, 1 "LMR", 2 "UMR",
 * 3 "UMR-DMV"; m_src_row = the row of ml_result_call_table / ml_result_call_lmr_umr it came from (-1 for a valley),
 * m_src_valley = the row of ml_result_call_valleys (-1 for a call-table row) - the remaining columns of seg_call are
 * gathered through these; m_segment_id = seq_len(.N) per condition (:137); m_follows_large_gap as recomputed at
 * :139-140.  Pieces = split_call after :168: the PMD candidate segments of ml_result_call_pmd_segments(pmd_lambda2)
 * (std_call, :93-95) intersected with the "isolate" / "complement" regions of :143-156 (split_pmds != 0: UMRs and
 * valleys isolate, :145; else valleys only, :148), fused_std set to 0 outside "complement" (:161), ordered by (job,
 * estimate_id, start) with p_segment_id = seq_len(.N) per estimate_id (:164), counted behind the segments that match no
 * region (rows of NAs that sort first and are dropped at :168).  The per-position gather and the PMD
 * calls (:169-199) are ml_result_call_pmds.  Query with NULL outputs for the counts first; any output may be NULL. */
ML_API int ml_result_call_pmd_split(ml_engine *eng, ml_result *r, double tol_val, double max_distance,
                                    double pmd_lambda2, const ml_lmr_params *lmr, const ml_valley_params *params,
                                    int split_pmds, int64_t *n_merged, int32_t *m_job, int32_t *m_condition,
                                    int32_t *m_segment_id, uint32_t *m_start, uint32_t *m_end, int8_t *m_category,
                                    int32_t *m_src_row, int32_t *m_src_valley, int8_t *m_follows_large_gap,
                                    int64_t *n_pieces, int32_t *p_job, int32_t *p_estimate_id, int32_t *p_segment_id,
                                    uint32_t *p_start, uint32_t *p_end, double *p_fused_std, double *p_pseudoweight);

/* ---- PMD calls (R/call.R:169-199) ---------------------------------------------------------------------------------------
 * Continues ml_result_call_pmd_split on the same resident tables and the pooled per-position data of the FAST result:
 * every piece of split_call collects the positions with start - 1 <= pos <= end (foverlaps of [pos, pos + 1], :170),
 * shrinks to them (:171) and gets the mean of Nmeth / coverage, the position count and its fused_std (:172-175); a
 * piece is called "PMD" when it is wider than pmd_valley_min_width, dense enough (width / (num.cpgs - 1) <
 * max_distance), fused_std >= pmd_std_threshold and beta < pmd_max_beta (:177-179); consecutive pieces with the same
 * call and no gap wider than pmd_valley_min_width are merged (merge_pmds != 0, :182-187); each merged region reports
 * start, end, width, num.cpgs, mean(fused_std) and the mean / sd of the methylation values of all its positions
 * (:190-198; sd of a single value -> 0).  Rows = pmd_call after :198, ordered by (job, condition, start);
 * category 1 = "PMD", 0 = "other" (drop = T keeps the PMD rows, :217); segment_id numbers the regions per condition.
 * The means / sd use double-double sums where R accumulates in long double (equal to the last bit or so, as in the
 * call table).  The final annotation of seg_call with PMD.status (:201-215) is ml_result_call_pmd_status.
 * pmd == NULL: the reference's defaults.  Query with NULL outputs for *n_rows first; any output may be NULL. */
typedef struct {
  double pmd_lambda2, pmd_std_threshold, pmd_max_beta;          /* 1000, 0.15, 0.75 */
  int32_t split_pmds, merge_pmds;                               /* 1, 1 */
} ml_pmd_params;
ML_API int ml_result_call_pmds(ml_engine *eng, ml_result *r, double tol_val, double max_distance,
                               const ml_lmr_params *lmr, const ml_valley_params *valley, const ml_pmd_params *pmd,
                               int64_t *n_rows, int32_t *job, int32_t *condition, int8_t *category,
                               int32_t *segment_id, uint32_t *start, uint32_t *end, double *width, int32_t *num_cpgs,
                               double *fused_std, double *beta, double *std);

/* ---- PMD status of the segments (R/call.R:201-215) -----------------------------------------------------------------------
 * Continues ml_result_call_pmds: seg_call (the merged table of ml_result_call_pmd_split) is joined with the PMD table
 * (foverlaps, :203), exact duplicates are removed (:209: a segment that overlaps several regions keeps one row per
 * distinct status, in order of first appearance), LMRs inside a PMD lose their category (:210) and every admissible
 * UMR takes the status its two admissible neighbours of the condition share - "other" when they differ, NA when one
 * is missing or NA (:211-212).  Rows = seg_call after :212, ordered by (job, condition, start): pmd_status 1 "PMD",
 * 0 "other", -1 NA; category 0 NA, 1 "LMR", 2 "UMR", 3 "UMR-DMV"; merged_row = the row of ml_result_call_pmd_split's
 * merged table, src_row / src_valley as there.  drop = T of the reference (:214-218) keeps the rows with category != 0
 * and the PMD rows of ml_result_call_pmds.  Query with NULL outputs for *n_rows first; any output may be NULL. */
ML_API int ml_result_call_pmd_status(ml_engine *eng, ml_result *r, double tol_val, double max_distance,
                                     const ml_lmr_params *lmr, const ml_valley_params *valley, const ml_pmd_params *pmd,
                                     int64_t *n_rows, int32_t *job, int32_t *condition, uint32_t *start, uint32_t *end,
                                     int8_t *category, int8_t *pmd_status, int32_t *merged_row, int32_t *src_row,
                                     int32_t *src_valley);

/* ---- segment_methylation in one call (R/call.R:229-244 around :31-221) -----------------------------------------------------
 * Everything above on the resident fit, and the two tables the reference returns, assembled on the device so that only
 * their rows cross PCIe: `lmr_umr_valley` = seg_call after :212 with the columns of :241 (job = chromosome index,
 * condition, start, end, width, segment_id, beta, coverage, pseudoweight, num_cpgs, std (NaN = NA on valleys), category
 * 0 NA / 1 LMR / 2 UMR / 3 UMR-DMV, pmd_status -1 NA / 0 other / 1 PMD) and `pmd` = pmd_call after :198 with the
 * columns of :242.  drop != 0 (the reference's default, :214-218) keeps the rows with a category and the "PMD" rows.
 * Query with NULL outputs for the two counts first; any output may be NULL. */
ML_API int ml_result_segment_methylation(ml_engine *eng, ml_result *r, double tol_val, double max_distance,
                                         const ml_lmr_params *lmr, const ml_valley_params *valley,
                                         const ml_pmd_params *pmd, int drop, int64_t *n_seg, int32_t *job,
                                         int32_t *condition, uint32_t *start, uint32_t *end, double *width,
                                         int32_t *segment_id, double *beta, double *coverage, double *pseudoweight,
                                         int32_t *num_cpgs, double *std, int8_t *category, int8_t *pmd_status,
                                         int64_t *n_pmd, int32_t *p_job, int32_t *p_condition, uint32_t *p_start,
                                         uint32_t *p_end, double *p_width, int32_t *p_segment_id, double *p_beta,
                                         double *p_std, int8_t *p_category, int32_t *p_num_cpgs, double *p_fused_std);

/* ---- lambda2 sweep on a resident one-step FAST fit (BASELINE.json configs[4]) ---------------------------------------------
 * signal_detection_fast for nlam values of lambda2 at once: the pseudodata of a one-step FAST fit do not depend on lambda2
 * (src/signal_detection.cpp:11-47 with max_iter = 1), so the sweep is ONE batched weighted_1d_flsa call in which the series
 * of all lambda2 values read the same resident (betahat, pseudoweight) stream, followed by the centring
 * (core/params_helpers.hpp:23-30) and segment_methylation_helper per value.  n_segments[nlam] (may be NULL) receives the
 * number of segments per value; beta (may be NULL) the fitted coefficients, nlam blocks laid out like the result's own
 * beta column (ml_result_copy_all order). */
ML_API int ml_result_lambda_sweep(ml_engine *eng, ml_result *r, int nlam, const double *lambda2, double lambda1,
                                  double tol_val, int64_t *n_segments, double *beta);

/* ---- K16: exchange of the per-shard region tables between the GPUs of one box --------------------------------------------
 * The reference parallelises over chromosomes with forked workers and concatenates what they return
 * (R/genome_wide.R:6-15, 42-43; R/call.R:230-244 rbindlist).  Here a worker is one process per GPU and the concatenation
 * is a push over NVLink: every rank owns a window in its HBM (one slot per rank and parity), exported as a CUDA IPC handle;
 * ml_gather_push runs the rule engine on the resident fit if that has not been done, packs the two region tables of
 * ml_result_segment_methylation (same columns, `job` mapped through job_ids[] to the caller's global job index) and writes
 * them into this rank's slot of every opened window with one kernel (peer stores), returning when they have landed.
 * The caller then all-gathers the two counts per rank (that collective is also the barrier before anybody reads) and
 * ml_gather_fetch copies the used part of every slot of the LOCAL window to host memory (slot w at w * slot_bytes, laid
 * out as ml_gather_layout says: offsets[24] of the 13 + 11 columns, each 16-byte aligned, behind a 64-byte header).
 * peers = 0 writes the local slot only (for a caller that moves slots with a collective of its own: ml_gather_window
 * gives the device address of the window of a parity).  Alternate parity from call to call. */
typedef struct ml_gather ml_gather;
ML_API int ml_gather_layout(int64_t n_seg, int64_t n_pmd, int64_t *offsets, int64_t *total_bytes);
ML_API int ml_gather_create(ml_engine *eng, int rank, int world, int64_t slot_bytes, ml_gather **out);
ML_API int ml_gather_ipc_handle(ml_gather *g, void *handle64);
ML_API int ml_gather_open(ml_gather *g, const void *handles);
ML_API int ml_gather_window(ml_gather *g, int parity, void **dev_ptr, int64_t *slot_bytes);
ML_API int ml_gather_push(ml_engine *eng, ml_gather *g, ml_result *r, double tol_val, double max_distance,
                          const ml_lmr_params *lmr, const ml_valley_params *valley, const ml_pmd_params *pmd, int drop,
                          const int32_t *job_ids, int parity, int peers, int64_t *n_seg, int64_t *n_pmd);
ML_API int ml_gather_fetch(ml_engine *eng, ml_gather *g, int parity, const int64_t *used_bytes, void *host_dst);
ML_API void ml_gather_destroy(ml_gather *g);

/* ---- input codec: methylation count files (MethyLasso.R:246-262: fread(file, select = c(1, 2, a, b))) -----------------
 * Parses the text of one file on the GPU: column 1 = chromosome, column 2 = position, col_a / col_b (0-based, >= 2) =
 * the two numeric columns the CLI selects.  sep = 0: tabs or blanks (runs of blanks count once), else the separator.
 * The first skip_lines lines are not data (header).  Numbers are converted exactly as strtod / fread do when one
 * correctly rounded operation suffices (<= 15 significant digits, |decimal exponent| <= 22); otherwise the line gets
 * status 1 and the host parses that field itself - nothing is approximated.  status: 0 ok, 1 a number the host must
 * parse, 2 malformed (missing column, not a number, position not an integer in [0, 2^31)), 3 blank line.
 * ml_text_parse keeps the result resident; ml_text_fetch copies what is asked for (any pointer may be NULL):
 * per data line pos, a, b, status and line_start (byte offset of the line in the text), per run of equal chromosome
 * fields its first data line and the offset / length of the field inside that line. */
typedef struct ml_text ml_text;
ML_API int ml_text_parse(ml_engine *eng, const char *text, int64_t nbytes, int64_t skip_lines, int col_a, int col_b,
                         char sep, ml_text **out, int64_t *n_lines, int64_t *n_chr_runs);
ML_API int ml_text_fetch(ml_engine *eng, const ml_text *t, int32_t *pos, double *a, double *b, int8_t *status,
                         int64_t *line_start, int64_t *run_first_line, int32_t *run_chr_off, int32_t *run_chr_len);
ML_API void ml_text_free(ml_text *t);

#ifdef __cplusplus
}
#endif
#endif

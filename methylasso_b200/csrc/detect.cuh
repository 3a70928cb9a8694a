// detect.cuh — host-side objects of the batched methylation_detection engine.
#pragma once
#include <memory>

#include "common.h"
#include "flsa.cuh"
#include "irls_core.cuh"
#include "segment.cuh"

namespace ml {

// mirror of ml_params (include/methylasso_b200.h)
struct Params {
  double lambda2, max_lambda2, lambda1, tol_val, bf_per_kb, hyper;
  uint32_t max_iter, nouter, ref;
  int32_t model;
  int32_t optimize_lambda2, optimize_hyper;
};

// per-job record read by every IRLS kernel
struct JobDev {
  int64_t row0;     // first row of the job in the batch columns
  int64_t par0;     // first parameter (estimator 0) in the parameter arrays; E*P contiguous
  int64_t tab0;     // first entry of the (dataset x param) -> first-row table
  int32_t nrows;
  int32_t P, D, E;
  int32_t dptr0;    // first of the D+1 dataset row offsets (job-relative) in the dptr array
  int32_t des0;     // first of the D*E design coefficients
  int32_t off0;     // first of the D offsets
  int32_t ser0;     // first FLSA series (one per estimator)
  int32_t rtile0;   // first row tile
  int32_t ptile0;   // first parameter tile
  int32_t n_rtiles, n_ptiles;
  int32_t model;
  int32_t pad;
  double lambda2, nu;
  // even-bin geometry (FULL)
  double lower, size;
  uint32_t nbins, pad2;
};

struct RowTile {   // never straddles a (job, dataset) boundary
  int32_t job, d;
  int32_t r0;      // job-relative first row
  int32_t cnt;
};
struct ParTile {   // never straddles a (job, estimator) boundary
  int32_t job, e;
  int32_t k0, cnt;
};

constexpr int ROW_TILE = 1024;   // 256 threads x 4 rows
constexpr int PAR_TILE = 256;

enum : int32_t {  // per-job control bits, re-uploaded every global step of the IRLS loop
  CTL_UPDATE = 1, CTL_DAMP = 2, CTL_EVAL = 4, CTL_COPY_RES = 8, CTL_COPY_OUTER = 16, CTL_FIRST = 32,
  CTL_NO_MU = 64, CTL_DRES = 128, CTL_ZERO_PARAMS = 256
};

struct Batch {
  Engine* eng = nullptr;
  int njobs = 0;
  int64_t nrows = 0;
  std::vector<int64_t> job_ptr;
  std::vector<int32_t> hint_D, hint_C;  // dataset / condition of each job's last row (host copy)
  DevBuf<int32_t> dataset, condition, replicate, pos, nmeth, cov;
  DevBuf<int64_t> d_job_ptr;
};

struct JobHost {
  int32_t status = 0;
  int64_t nrows = 0, P = 0;
  int32_t D = 0, E = 0, C = 0;
  int64_t par0 = 0, row0 = 0;
  int32_t off0 = 0;
  int32_t converged = 0, irls_steps = 0, dampings = 0, outer_iters = 0;
  double hyper = 0, mlogp = 0, lambda2 = 0;
};

struct Result {
  Engine* eng = nullptr;
  std::vector<JobHost> jobs;
  int64_t nrows = 0, npar = 0;  // totals over the batch (npar = sum E*P over valid jobs)
  int32_t noff = 0;
  DevBuf<double> mu;                      // per row (rows of failed jobs are untouched)
  DevBuf<double> beta, betahat, pw;       // per parameter
  DevBuf<double> start;                   // per (job, param): bin start truncated to int, as double
  DevBuf<double> offset;
  DevBuf<int32_t> est_id;                 // per parameter: estimator index + 1
  std::vector<int64_t> start0;            // per job: offset into start (P entries)
  int64_t nstart = 0;
  // segments of the last ml_result_segments call (query-then-fetch without recomputing)
  std::unique_ptr<SegDeviceOut> seg_cache;
  double seg_tol = 0;
  std::unique_ptr<SegDeviceOut> pmd_cache;
  DevBuf<double> pmd_fused;
  double pmd_tol = 0, pmd_lambda2 = 0;
  // call table of R/call.R:43-68 (and the PMD re-segmentation computed from it)
  std::unique_ptr<CallTable> call_cache;
  double call_tol = 0, call_max_distance = 0;
  std::unique_ptr<SegDeviceOut> call_pmd_cache;
  DevBuf<double> call_pmd_fused;
  double call_pmd_lambda2 = 0;
  bool counts_exact = false;   // FAST model, one IRLS step: betahat * pw are the pooled methylated counts
  bool diff_combined = false;  // result_fast_diff_combine has been applied
  // group_call of the last ml_result_call_differences call and the arguments it was computed with
  std::shared_ptr<struct DiffCallTable> diff_cache;
  double diff_key[5] = {0, 0, 0, 0, 0};
  // tables of the rule engine (R/call.R:75-215) of the last ml_result_call_pmd_split / _pmds / _pmd_status call and the
  // arguments they were computed with (query-then-fetch and the chain of the three entries without recomputing);
  // the type lives in capi.cu
  std::shared_ptr<void> rule_cache;
  std::vector<char> rule_key;
};

// group_call of R/call.R:305-332 (call_differences_singlechr), device arrays; rows ordered by (job, estimate, segment)
struct DiffCallTable {
  int64_t n = 0;
  DevBuf<int32_t> job, condition, estimate_id, segment_id, num_cpgs, num_cpgs_ref;
  DevBuf<uint32_t> start, end;
  DevBuf<double> width, beta, std_, coverage, beta_ref, std_ref, coverage_ref, pseudoweight, diff, cohen_d, pval,
      coverage_score;
  void alloc(cudaStream_t st, int64_t count);
};
// The DMR caller on a resident FAST result after result_fast_diff_combine (segment.cu / diffcall.inc.cuh); the raw
// rows of the batch the result was computed from must still be resident.  Returns the number of rows.
int64_t call_differences_device(Engine& eng, const Batch& B, const Result& R, int ref_condition, double min_diff,
                                double min_delta_diff, double tol_val, double max_distance, DiffCallTable& out);

std::unique_ptr<Batch> batch_upload(Engine& eng, int njobs, const int64_t* job_ptr, const int32_t* dataset,
                                    const int32_t* condition, const int32_t* replicate,
                                    const int32_t* pos, const int32_t* nmeth, const int32_t* cov);
std::unique_ptr<Result> batch_detect(Engine& eng, Batch& b, const Params& p);
void result_fast_diff_combine(Engine& eng, Result& r);
// per-device constant tables of the IRLS kernels (idempotent; called when an engine is created)
void detect_init_device_tables(cudaStream_t st);

}  // namespace ml

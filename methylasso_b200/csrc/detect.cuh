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

// Input rows of a batch, resident in HBM.  Only pos / Nmeth / coverage live on the device: dataset, condition
// and replicate are factor codes that are constant along each dataset's run of rows (the reference validates them
// at run boundaries only, core/Observations.hpp:36-56), so the host keeps them as a RUN TABLE (one entry per job
// and dataset: first row, condition, replicate, first / last position) and nothing of those three columns crosses
// PCIe.  Every job's rows start at a multiple of ROW_ALIGN in the device columns (16-byte vector loads).
struct Batch {
  Engine* eng = nullptr;
  int njobs = 0;
  int64_t nrows = 0;         // rows handed in by the caller
  int64_t nrows_dev = 0;     // rows of the device columns (job starts padded to ROW_ALIGN)
  std::vector<int64_t> job_ptr;      // caller offsets (njobs + 1)
  std::vector<int64_t> dev_row0;     // first device row of each job
  // run table: entries [run0[j], run0[j+1]) belong to job j, one per dataset run found before the first
  // structural offence; dptr holds run0[j+1] - run0[j] + 1 job-relative row offsets starting at dptr0[j]
  std::vector<int32_t> run0, dptr0;
  std::vector<int32_t> dptr, dcond, drep, dfirst, dlast;
  std::vector<unsigned long long> host_key;   // first structural offence per job (validation key), ~0: none
  std::vector<int32_t> last_condition;        // condition of each job's last row (data_design_helpers.cpp:23)
  DevBuf<int32_t> pos, nmeth, cov;
  // between batch_stage and batch_commit: the caller's columns and how Nmeth / coverage were packed for the bus
  const int32_t *h_pos = nullptr, *h_nmeth = nullptr, *h_cov = nullptr;
  int pack_bytes = 0;        // bytes per packed count: 1 (all counts < 256), 2 (< 65536), 0: not packed (int32 columns copied as they are)
  bool committed = false;
  int n_runs(int j) const { return run0[j + 1] - run0[j]; }
};
constexpr int ROW_ALIGN = 32;

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
  int64_t nrows_dev = 0;        // rows of the device columns (job starts padded, Batch::dev_row0); JobHost::row0 indexes them
  bool has_mu = true;           // mu was stored (engine key want_mu, see batch_detect)
  cudaEvent_t copy_done = nullptr;   // last asynchronous download on the copy stream (ml_result_copy_all_async)
  ~Result() {
    // the buffers below are freed in main-stream order: not before an asynchronous download of them has finished
    if (copy_done) {
      if (eng) cudaStreamWaitEvent(eng->stream, copy_done, 0);
      cudaEventDestroy(copy_done);
    }
  }
  int32_t noff = 0;
  DevBuf<double> mu;                      // per row (rows of failed jobs are untouched)
  DevBuf<double> beta, betahat, pw;       // per parameter
  DevBuf<double> start;                   // per (job, param): bin start truncated to int, as double
  DevBuf<double> offset;
  DevBuf<int32_t> est_id;                 // per parameter: estimator index + 1
  std::vector<int64_t> start0;            // per job: offset into start (P entries)
  DevBuf<int32_t> rowstart;               // FAST: [job][dataset][P] first row of dataset d on parameter k (-1: none), kept for
  std::vector<int64_t> tab0;              // the DMR caller's join of the raw rows (R/call.R:258-263); per job: offset into it
  int64_t nstart = 0;
  // segments of the last ml_result_segments call (query-then-fetch without recomputing)
  std::unique_ptr<SegDeviceOut> seg_cache;
  double seg_tol = 0;
  // a difference fit: the segments of its REFERENCE condition alone (estimate_id 1), which is what segment_methylation
  // works on when it is given such a fit (R/call.R:40-44)
  std::unique_ptr<SegDeviceOut> seg_ref_cache;
  double seg_ref_tol = 0;
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
// batch_upload in two halves.  batch_stage is host work only (run table of the factor columns; Nmeth / coverage packed
// into the engine's page-locked staging area, one or two bytes per count when they all fit - a third to a half of the
// bytes that cross PCIe); batch_commit copies pos and the packed counts and unpacks them on the device.  The caller's
// columns must stay valid in between; one staged batch per engine at a time.
std::unique_ptr<Batch> batch_stage(Engine& eng, int njobs, const int64_t* job_ptr, const int32_t* dataset,
                                   const int32_t* condition, const int32_t* replicate,
                                   const int32_t* pos, const int32_t* nmeth, const int32_t* cov);
void batch_commit(Engine& eng, Batch& b);
// the same from a caller-made run table: dset_ptr[j] .. dset_ptr[j+1] index the datasets of job j in
// dset_row (first row of each dataset, job-relative) / dset_condition / dset_replicate
std::unique_ptr<Batch> batch_upload_runs(Engine& eng, int njobs, const int64_t* job_ptr, const int64_t* dset_ptr,
                                         const int64_t* dset_row, const int32_t* dset_condition,
                                         const int32_t* dset_replicate, const int32_t* pos, const int32_t* nmeth,
                                         const int32_t* cov);
std::unique_ptr<Result> batch_detect(Engine& eng, Batch& b, const Params& p);
void debug_scan_runs(int64_t n, const int32_t* dataset, const int32_t* condition, const int32_t* replicate,
                     const int32_t* pos, int threads, Batch& b);
void result_fast_diff_combine(Engine& eng, Result& r);
// per-device constant tables of the IRLS kernels (idempotent; called when an engine is created)
void detect_init_device_tables(cudaStream_t st);

}  // namespace ml

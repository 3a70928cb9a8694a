// segment.cuh — host interface of the segment_methylation_helper kernels (segment.cu).
#pragma once
#include <algorithm>
#include "common.h"
#include "valley_core.cuh"
#include "pmdsplit_core.cuh"
#include "pmdstatus_core.cuh"

namespace ml {

// one estimate of one table: rows [row0, row0+nrows) of the beta/betahat/pw arrays; its `start`
// column lives at start[start0 + (row - row0)]
struct SegGroup {
  int64_t row0, nrows, start0;
  int32_t estimate_id, job;
};

struct SegOut {   // device pointers
  int32_t* first_row;   // first row of the segment in the arrays it was computed on
  int32_t *job, *estimate_id, *seg_id;
  uint32_t *start, *end;
  double *beta, *betahat, *pseudoweight, *stdev;
};

struct SegHostOut {  // caller buffers (any may be null)
  int32_t *job = nullptr, *estimate_id = nullptr, *seg_id = nullptr;
  uint32_t *start = nullptr, *end = nullptr;
  double *beta = nullptr, *betahat = nullptr, *pseudoweight = nullptr, *stdev = nullptr;
};

struct SegDeviceOut {
  DevBuf<int32_t> job, estimate_id, seg_id, first_row;
  DevBuf<uint32_t> start, end;
  DevBuf<double> beta, betahat, pseudoweight, stdev;
  int64_t n = 0;
  bool has_first_row = false;   // set by segment_device: first_row holds row indices of its input arrays
  // host copy of the group structure: segments [group_seg0[g], group_seg0[g+1]) belong to group g
  std::vector<int32_t> group_seg0, group_job, group_est;
  void alloc(cudaStream_t st, int64_t count) {
    n = count;
    const size_t c = (size_t)std::max<int64_t>(1, count);
    job.alloc(st, c); estimate_id.alloc(st, c); seg_id.alloc(st, c); first_row.alloc(st, c);
    start.alloc(st, c); end.alloc(st, c);
    beta.alloc(st, c); betahat.alloc(st, c); pseudoweight.alloc(st, c); stdev.alloc(st, c);
  }
  SegOut view() { return SegOut{first_row.p, job.p, estimate_id.p, seg_id.p, start.p, end.p, beta.p, betahat.p, pseudoweight.p, stdev.p}; }
  void download(cudaStream_t, int64_t count, SegHostOut& h) const {
    const size_t c = (size_t)count;
    if (h.job) job.download(h.job, c);
    if (h.estimate_id) estimate_id.download(h.estimate_id, c);
    if (h.seg_id) seg_id.download(h.seg_id, c);
    if (h.start) start.download(h.start, c);
    if (h.end) end.download(h.end, c);
    if (h.beta) beta.download(h.beta, c);
    if (h.betahat) betahat.download(h.betahat, c);
    if (h.pseudoweight) pseudoweight.download(h.pseudoweight, c);
    if (h.stdev) stdev.download(h.stdev, c);
  }
};

// Call table of R/call.R:43-68: `t` holds estimate_id, seg_id, start, end, beta (mean methylation), pseudoweight and
// stdev (sample sd, 0 for single-CpG rows) per gap-split segment; coverage / width / num_cpgs complete the row.
struct CallTable {
  SegDeviceOut t;
  DevBuf<double> coverage, width;
  DevBuf<int32_t> num_cpgs;
};
// Thresholds of the LMR / UMR annotation (arguments of segment_methylation_singlechr, R/call.R:31-35, with its defaults)
struct LmrParams {
  double min_width = 30, max_distance = 500, flank_width = 300, flank_dist_max = 5000, lmr_max_beta = 0.5,
         flank_beta_min = 0.5, umr_max_beta = 0.1, umr_std_threshold = 0.1, umr_large_width = 500, umr_large_density = 0.03;
  int32_t min_num_cpgs = 4, flank_num_cpgs = 10;
};
// R/call.R:75-91 on a call table: one entry per call-table row.  Logical columns: 1 TRUE, 0 FALSE, -1 NA;
// flank_dist / flank_beta NaN = NA; category 0 NA, 1 LMR, 2 UMR.
struct LmrUmrColumns {
  int64_t n = 0;
  DevBuf<int8_t> is_admissible, is_v, is_flank, is_sep, category;
  DevBuf<double> flank_dist, flank_beta;
  void alloc(cudaStream_t st, int64_t count);
};
void call_lmr_umr_device(Engine& eng, const CallTable& t, const LmrParams& p, LmrUmrColumns& out);
// R/call.R:100-129: the "UMR-DMV" valleys of a call table (ordered by job, condition, start), and per call-table row
// whether a valley replaces it (:132-133); low_w / low_b are the low.id.width / low.id.beta columns of :106 (NaN = NA).
struct ValleyTable {
  int64_t n = 0;
  DevBuf<int32_t> job, condition, num_cpgs;
  DevBuf<uint32_t> start, end;
  DevBuf<double> width, beta, coverage, pseudoweight;
  DevBuf<int8_t> keep;
  DevBuf<int32_t> row0;
  void alloc(cudaStream_t st, int64_t count);
  ValleyOut view();
};
void call_valleys_device(Engine& eng, const CallTable& t, const int8_t* d_category, const ValleyParams& p,
                         ValleyTable& out, DevBuf<int8_t>& row_in_valley, DevBuf<double>& low_w, DevBuf<double>& low_b);

// R/call.R:131-164: seg_call with the valleys in place of the rows they overlap (MergedDev), and split_call, the PMD
// candidate segments cut along the isolate / complement regions (PiecesDev).
struct MergedDev {
  int64_t n = 0;
  DevBuf<int32_t> job, est, src_row, src_valley, segment_id;
  DevBuf<uint32_t> start, end;
  DevBuf<int8_t> category, follows_large_gap;
  void alloc(cudaStream_t st, int64_t count);
  MergedTable view();
};
struct PiecesDev {
  int64_t n = 0;
  DevBuf<int32_t> job, est, segment_id;
  DevBuf<uint32_t> start, end;
  DevBuf<double> fused_std, pw;
  void alloc(cudaStream_t st, int64_t count);
};
void call_pmd_split_device(Engine& eng, const CallTable& t, const int8_t* d_category, ValleyTable& val,
                           const int8_t* d_row_in_valley, const SegDeviceOut& pmd, double pmd_valley_min_width,
                           int split_pmds, MergedDev& M, PiecesDev& P);

// R/call.R:169-199: the PMD calls.  One row per merged region of pmd_call after :198: category 1 "PMD" / 0 "other".
struct PmdCallParams {    // arguments of segment_methylation_singlechr (R/call.R:31-35), its defaults
  double pmd_std_threshold = 0.15, pmd_max_beta = 0.75, pmd_valley_min_width = 5000, max_distance = 500;
  int32_t merge_pmds = 1, pad = 0;
};
struct PmdCallDev {
  int64_t n = 0;
  DevBuf<int32_t> job, est, segment_id, num_cpgs;
  DevBuf<int8_t> category;
  DevBuf<uint32_t> start, end;
  DevBuf<double> width, fused_std, beta, std_;
  void alloc(cudaStream_t st, int64_t count);
};
void call_pmds_device(Engine& eng, const std::vector<SegGroup>& groups, const void* d_pos, int pos_f64,
                      const double* d_betahat, const double* d_pw, const PiecesDev& pieces, const PmdCallParams& p,
                      PmdCallDev& out);

// R/call.R:201-215: seg_call annotated with PMD.status.  One row per (merged row, distinct status): `merged` indexes
// the MergedDev table; status 1 "PMD", 0 "other", -1 NA; category after :210; adm = is.admissible (-1 on a valley).
struct FinalDev {
  int64_t n = 0;
  DevBuf<int32_t> merged, job, est, src_row, src_valley;
  DevBuf<uint32_t> start, end;
  DevBuf<int8_t> status, category, adm;
  void alloc(cudaStream_t st, int64_t count);
};
void call_pmd_status_device(Engine& eng, MergedDev& M, const PmdCallDev& pmd, const int8_t* d_row_adm, FinalDev& out);

// The tables segment_methylation returns (R/call.R:214-220, columns of :241-243), gathered on the device:
// lmr_umr_valley = seg_call with PMD.status (drop: rows with a category only), pmd = pmd_call (drop: "PMD" rows only).
// category: 0 NA, 1 LMR, 2 UMR, 3 UMR-DMV; pmd_status: -1 NA, 0 other, 1 PMD; pmd category: 1 PMD, 0 other.
struct RegionTables {
  int64_t n_seg = 0, n_pmd = 0;
  DevBuf<int32_t> s_job, s_condition, s_segment_id, s_num_cpgs;
  DevBuf<uint32_t> s_start, s_end;
  DevBuf<double> s_width, s_beta, s_coverage, s_pseudoweight, s_std;
  DevBuf<int8_t> s_category, s_pmd_status;
  DevBuf<int32_t> p_job, p_condition, p_segment_id, p_num_cpgs;
  DevBuf<int8_t> p_category;
  DevBuf<uint32_t> p_start, p_end;
  DevBuf<double> p_width, p_fused_std, p_beta, p_std;
};
void region_tables_device(Engine& eng, const CallTable& ct, const ValleyTable& val, const MergedDev& M, const PmdCallDev& pmd,
                          const FinalDev& F, int drop, RegionTables& out);

int64_t call_table_from_estimates(Engine& eng, const std::vector<SegGroup>& groups, const SegDeviceOut& seg,
                                  const void* d_pos, int pos_f64, const double* d_betahat, const double* d_pw,
                                  double max_distance, CallTable& out);
int64_t call_table_from_counts(Engine& eng, const std::vector<SegGroup>& groups, const SegDeviceOut& seg,
                               const int32_t* d_pos, const int32_t* d_nmeth, const int32_t* d_cov,
                               double max_distance, CallTable& out);

int64_t segment_device(Engine& eng, const std::vector<SegGroup>& groups, int64_t nrows, const void* d_start,
                       int start_f64, const double* d_beta, const double* d_betahat, const double* d_pw,
                       double tol_val, SegDeviceOut& out);
int64_t pmd_segments_device(Engine& eng, const SegDeviceOut& seg, double pmd_lambda2, double tol_val,
                            DevBuf<double>& fused, SegDeviceOut& out);
int64_t segment_host_table(Engine& eng, int64_t n, const int32_t* estimate_id, const int32_t* start,
                           const double* beta, const double* betahat, const double* pw, double tol_val,
                           SegHostOut& out);

// once per device, at engine creation: tables the kernels of segment.cu read (counts_div.cuh)
void segment_init_device_tables(cudaStream_t st);

// test hook (debug_api.h): pairs checked / pairs whose quotient differs from the division's
void debug_div_counts(Engine& eng, int max_num, int max_den, long long* pairs, long long* mismatches);

}  // namespace ml

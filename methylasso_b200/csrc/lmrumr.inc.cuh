// lmrumr.inc.cuh — the LMR / UMR annotation block of segment_methylation_singlechr, R/call.R:75-91, on the resident
// call table (R/call.R:43-72, CallTable).  Included at the end of segment.cu (it uses the scan defined there).
//
// Every `seg_call[i, col := expr, by = ...]` line keeps data.table's semantics: rows where `i` is NA are not
// selected; shift() inside such a statement walks the SELECTED rows (of the `by` group: one condition of one
// chromosome; without `by`: one chromosome), so each line is a mask, a compaction (scan) and a kernel over the
// compacted rows that looks at its two neighbours in the list.  Logical columns are three-valued: 1 TRUE, 0 FALSE,
// -1 NA; numeric NA is NaN.
namespace {

constexpr int8_t TRI_T = 1, TRI_F = 0, TRI_NA = -1;
__device__ __forceinline__ int8_t tri_and(int8_t a, int8_t b) {
  if (a == TRI_F || b == TRI_F) return TRI_F;
  return (a == TRI_T && b == TRI_T) ? TRI_T : TRI_NA;
}
__device__ __forceinline__ int8_t tri_or(int8_t a, int8_t b) {
  if (a == TRI_T || b == TRI_T) return TRI_T;
  return (a == TRI_F && b == TRI_F) ? TRI_F : TRI_NA;
}
__device__ __forceinline__ int8_t tri_lt(double x, double y) { return (x != x || y != y) ? TRI_NA : (x < y ? TRI_T : TRI_F); }
__device__ __forceinline__ int8_t tri_is_true(int8_t a) { return a == TRI_NA ? TRI_NA : (a == TRI_T ? TRI_T : TRI_F); }   // a == T
__device__ __forceinline__ int8_t tri_not_true(int8_t a) { return a == TRI_NA ? TRI_NA : (a != TRI_T ? TRI_T : TRI_F); }  // a != T

struct LuCols {          // the call table and the annotation columns, device pointers
  int64_t n;
  const int32_t *job, *est, *ncpg;
  const uint32_t *start, *end;
  const double *width, *beta, *std_;
  int8_t *adm, *is_v, *is_flank, *is_sep, *category;
  double *flank_dist, *flank_beta;
};

// :77  is.admissible
__global__ void k_lu_admissible(LuCols c, LmrParams p, int32_t* __restrict__ flag) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= c.n) return;
  const double w = c.width[i], m = (double)c.ncpg[i];
  const bool a = w > p.min_width && w / (m - 1) < p.max_distance && m >= (double)p.min_num_cpgs;
  c.adm[i] = a ? 1 : 0;
  flag[i] = a ? 1 : 0;
}
// :78  is.V among the admissible rows of a condition
__global__ void k_lu_isv(LuCols c, int32_t n_list, const int32_t* __restrict__ list) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_list) return;
  const int i = list[r];
  double prev = NAN, next = NAN;
  if (r > 0) { const int j = list[r - 1]; if (c.job[j] == c.job[i] && c.est[j] == c.est[i]) prev = c.beta[j]; }
  if (r + 1 < n_list) { const int j = list[r + 1]; if (c.job[j] == c.job[i] && c.est[j] == c.est[i]) next = c.beta[j]; }
  c.is_v[i] = tri_and(tri_lt(c.beta[i], prev), tri_lt(c.beta[i], next));
}
// :79  selection (is.V == T) | (width > flank.width & is.admissible == T)
__global__ void k_lu_sel_flank(LuCols c, LmrParams p, int32_t* __restrict__ flag) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= c.n) return;
  const int8_t wide = (c.width[i] > p.flank_width && c.adm[i]) ? TRI_T : TRI_F;
  flag[i] = tri_or(tri_is_true(c.is_v[i]), wide) == TRI_T ? 1 : 0;
}
// :79  is.flank; shift() walks the selected rows of the chromosome (no `by`)
__global__ void k_lu_flank(LuCols c, int32_t n_list, const int32_t* __restrict__ list) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_list) return;
  const int i = list[r];
  int8_t prev = TRI_NA, next = TRI_NA;
  if (r > 0) { const int j = list[r - 1]; if (c.job[j] == c.job[i]) prev = c.is_v[j]; }
  if (r + 1 < n_list) { const int j = list[r + 1]; if (c.job[j] == c.job[i]) next = c.is_v[j]; }
  c.is_flank[i] = tri_and(tri_not_true(c.is_v[i]), tri_or(tri_is_true(prev), tri_is_true(next)));
}
// :80-82  selection is.V == T | is.flank == T
__global__ void k_lu_sel_stats(LuCols c, int32_t* __restrict__ flag) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= c.n) return;
  flag[i] = tri_or(tri_is_true(c.is_v[i]), tri_is_true(c.is_flank[i])) == TRI_T ? 1 : 0;
}
// :80-82  flank.dist, flank.beta, is.sep among the selected rows of a condition
__global__ void k_lu_flank_stats(LuCols c, int32_t n_list, const int32_t* __restrict__ list) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_list) return;
  const int i = list[r];
  double lead_start = NAN, lead_beta = NAN, lag_end = NAN, lag_beta = NAN;
  if (r > 0) {
    const int j = list[r - 1];
    if (c.job[j] == c.job[i] && c.est[j] == c.est[i]) { lag_end = (double)c.end[j]; lag_beta = c.beta[j]; }
  }
  if (r + 1 < n_list) {
    const int j = list[r + 1];
    if (c.job[j] == c.job[i] && c.est[j] == c.est[i]) { lead_start = (double)c.start[j]; lead_beta = c.beta[j]; }
  }
  const double a1 = lead_start - (double)c.end[i], a2 = (double)c.start[i] - lag_end;
  c.flank_dist[i] = (a1 != a1 || a2 != a2) ? NAN : fmax(a1, a2);               // pmax: NA if either is
  c.flank_beta[i] = (lead_beta != lead_beta || lag_beta != lag_beta) ? NAN : fmin(lead_beta, lag_beta);
  const double d1 = fabs(c.beta[i] - lag_beta), d2 = fabs(c.beta[i] - lead_beta);
  const double m = (d1 != d1 || d2 != d2) ? NAN : fmin(d1, d2);
  c.is_sep[i] = m != m ? TRI_NA : (m > 2 * c.std_[i] ? TRI_T : TRI_F);
}
// :83-86  category: 0 NA, 1 LMR, 2 UMR
__global__ void k_lu_category(LuCols c, LmrParams p) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= c.n) return;
  const double beta = c.beta[i], w = c.width[i], m = (double)c.ncpg[i], sd = c.std_[i];
  const bool v = c.is_v[i] == TRI_T;
  int8_t cat = 0;
  if (v && c.is_sep[i] == TRI_T && c.flank_dist[i] < p.flank_dist_max && beta < p.lmr_max_beta) cat = 1;
  if (cat == 1 && m < (double)p.flank_num_cpgs && c.flank_beta[i] <= p.flank_beta_min) cat = 0;
  const bool low = beta < p.umr_max_beta && sd < p.umr_std_threshold;
  if (low && cat == 1) cat = 2;
  if (low && v && w > p.umr_large_width && m / w > p.umr_large_density) cat = 2;
  c.category[i] = cat;
}

}  // namespace

void LmrUmrColumns::alloc(cudaStream_t st, int64_t count) {
  n = count;
  const size_t c = (size_t)std::max<int64_t>(1, count);
  is_admissible.alloc(st, c); is_v.alloc(st, c); is_flank.alloc(st, c); is_sep.alloc(st, c); category.alloc(st, c);
  flank_dist.alloc(st, c); flank_beta.alloc(st, c);
}

void call_lmr_umr_device(Engine& eng, const CallTable& t, const LmrParams& p, LmrUmrColumns& out) {
  cudaStream_t st = eng.stream;
  const int64_t n = t.t.n;
  out.alloc(st, n);
  if (n == 0) return;
  if (n >= 0x7fffffff) throw MlError(19, "call table too large");
  // NA everywhere first: rows a statement does not select keep NA (-1 in every byte of an int8 / a NaN pattern)
  ML_CUDA(cudaMemsetAsync(out.is_v.p, 0xff, (size_t)n, st));
  ML_CUDA(cudaMemsetAsync(out.is_flank.p, 0xff, (size_t)n, st));
  ML_CUDA(cudaMemsetAsync(out.is_sep.p, 0xff, (size_t)n, st));
  ML_CUDA(cudaMemsetAsync(out.flank_dist.p, 0xff, (size_t)n * sizeof(double), st));    // all-ones is a NaN
  ML_CUDA(cudaMemsetAsync(out.flank_beta.p, 0xff, (size_t)n * sizeof(double), st));
  const LuCols c{n, t.t.job.p, t.t.estimate_id.p, t.num_cpgs.p, t.t.start.p, t.t.end.p, t.width.p, t.t.beta.p, t.t.stdev.p,
                 out.is_admissible.p, out.is_v.p, out.is_flank.p, out.is_sep.p, out.category.p, out.flank_dist.p,
                 out.flank_beta.p};
  DevBuf<int32_t> flag(st, (size_t)n), list(st, (size_t)n);
  Scan scan;
  const int nb = cdiv(n, 256);
  ML_LAUNCH(eng, "k_lu_admissible", k_lu_admissible<<<nb, 256, 0, st>>>(c, p, flag.p));
  int32_t m = scan.run(eng, flag.p, n, nullptr, list.p);
  if (m > 0) ML_LAUNCH(eng, "k_lu_isv", k_lu_isv<<<cdiv(m, 256), 256, 0, st>>>(c, m, list.p));
  ML_LAUNCH(eng, "k_lu_sel_flank", k_lu_sel_flank<<<nb, 256, 0, st>>>(c, p, flag.p));
  m = scan.run(eng, flag.p, n, nullptr, list.p);
  if (m > 0) ML_LAUNCH(eng, "k_lu_flank", k_lu_flank<<<cdiv(m, 256), 256, 0, st>>>(c, m, list.p));
  ML_LAUNCH(eng, "k_lu_sel_stats", k_lu_sel_stats<<<nb, 256, 0, st>>>(c, flag.p));
  m = scan.run(eng, flag.p, n, nullptr, list.p);
  if (m > 0) ML_LAUNCH(eng, "k_lu_flank_stats", k_lu_flank_stats<<<cdiv(m, 256), 256, 0, st>>>(c, m, list.p));
  ML_LAUNCH(eng, "k_lu_category", k_lu_category<<<nb, 256, 0, st>>>(c, p));
  stream_sync(st);
}

// diffcall.inc.cuh — the DMR caller of R/call.R:254-333 (call_differences_singlechr) on the resident result of
// difference_detection_fast.  Included at the end of segment.cu (same translation unit: it reuses the scan, the
// call-table range / count / parts kernels and the double-double sums defined there).
//
// Per job and per non-reference condition (estimate e >= 2 of R/fast_diff.R's table):
//   mdata        :258-271  per (condition, position): coverage summed over the replicates, beta = mean of the
//                          replicates' Nmeth / coverage                           -> k_dc_rowidx, k_dc_pool
//   diff_call    :273      segment_methylation_helper on the difference estimates   -> segment_device
//   ov / parts   :278-288  foverlaps membership, split at gaps > max_distance over the UNION of both conditions'
//                          positions, shrunk boundaries                             -> k_call_ranges / count / parts
//   seg_call     :291-298  per part and side (condition / reference): mean beta, coverage, number of CpGs; diff
//                                                                                   -> k_dc_emit_part
//   groups       :301-314  neighbouring parts with similar, large differences are merged; mean / sd of all their
//                          per-position values, Cohen's d                           -> k_dc_group_flags, k_dc_group_stats
//   pval         :316-319  wilcox.test(betavals, betavals.ref), NA -> 1             -> stats.cu
//   coverage.score :322-331 mean over the positions inside [start - 1, end] of (datasets covering it) * 100 / D
//                                                                                   -> k_dc_cov_score
// The reference orders the table by (start, end) after its merge (:331); here rows are ordered by
// (job, estimate_id, segment_id).
namespace {

struct DcJob {          // one valid job
  int64_t jrow0;        // first row in the batch columns
  int64_t start0;       // offset of its positions in Result::start
  int64_t tab0;         // offset of its (dataset, param) -> row table
  int32_t nrows, P, D, dc0;      // dc0: offset of its D dataset -> condition entries
};
struct DcGroup {        // one comparison = (job, estimator e >= 1)
  int64_t row0;         // parameter 0 of estimator e in the parameter arrays
  int32_t job_slot;     // index into the DcJob array
  int32_t cond_x, cond_ref, pad;
};

// number of datasets that cover each position of each job (R/call.R:326 table(data$pos))
__global__ void __launch_bounds__(256)
k_dc_dscount(const DcJob* __restrict__ jobs, const int32_t* __restrict__ tile_job, const int32_t* __restrict__ tile_k0,
             const int32_t* __restrict__ rowidx, int32_t* __restrict__ dscount) {
  const DcJob J = jobs[tile_job[blockIdx.x]];
  const int k = tile_k0[blockIdx.x] + threadIdx.x;
  if (k >= J.P) return;
  int c = 0;
  for (int d = 0; d < J.D; ++d) c += rowidx[J.tab0 + (int64_t)d * J.P + k] >= 0;
  dscount[J.start0 + k] = c;
}

// mdata of R/call.R:258-263 for the two sides of every comparison, in the parameter index space of estimator e
__global__ void __launch_bounds__(256)
k_dc_pool(const DcJob* __restrict__ jobs, const DcGroup* __restrict__ groups, const int32_t* __restrict__ tile_group,
          const int32_t* __restrict__ tile_k0, const int32_t* __restrict__ dcond, const int32_t* __restrict__ rowidx,
          const int32_t* __restrict__ nmeth, const int32_t* __restrict__ cov, double* __restrict__ xcov,
          double* __restrict__ xbeta, double* __restrict__ rcov, double* __restrict__ rbeta) {
  const int g = tile_group[blockIdx.x];
  const DcGroup G = groups[g];
  const DcJob J = jobs[G.job_slot];
  const int k = tile_k0[blockIdx.x] + threadIdx.x;
  if (k >= J.P) return;
  const int32_t* dc = dcond + J.dc0;
#pragma unroll
  for (int side = 0; side < 2; ++side) {
    const int cond = side ? G.cond_ref : G.cond_x;
    constexpr int MAXV = 8;                     // replicates of one condition kept in registers
    double v[MAXV];
    int n = 0;
    double csum = 0.0;
    int n_all = 0;
    for (int d = 0; d < J.D; ++d) {
      if (dc[d] != cond) continue;
      const int r = rowidx[J.tab0 + (int64_t)d * J.P + k];
      if (r < 0) continue;
      const double c = (double)cov[J.jrow0 + r];
      csum += c;
      if (n < MAXV) v[n++] = div_nz((double)nmeth[J.jrow0 + r], c);
      ++n_all;
    }
    double mean = 0.0;
    if (n_all > 0 && n_all <= MAXV) {
      mean = x87_mean(n, [&](int i) { return v[i]; });   // R's mean(), 80-bit roundings included (x87.cuh)
    } else if (n_all > MAXV) {                  // many replicates: read the rows again in each pass
      auto val = [&](int i) {
        int seen = 0;
        for (int d = 0; d < J.D; ++d) {
          if (dc[d] != cond) continue;
          const int r = rowidx[J.tab0 + (int64_t)d * J.P + k];
          if (r < 0) continue;
          if (seen++ == i) return div_nz((double)nmeth[J.jrow0 + r], (double)cov[J.jrow0 + r]);
        }
        return 0.0;
      };
      mean = x87_mean(n_all, val);
    }
    (side ? rcov : xcov)[G.row0 + k] = n_all > 0 ? csum : 0.0;
    (side ? rbeta : xbeta)[G.row0 + k] = mean;
    // a position with rows but zero total coverage cannot occur: validation requires coverage >= Nmeth >= 0 and the
    // CLI filters coverage >= 5; such a row would have NaN beta and is treated as absent here
  }
}

// rows of the comparison tables: a position is a row of mdata if either side has data there
struct RowsDiff {
  const double* xcov;
  const double* rcov;
  __device__ bool valid(int64_t i) const { return xcov[i] > 0 || rcov[i] > 0; }
};

struct DcParts {         // seg_call of R/call.R:291-298, one row per part
  int32_t *seg, *a, *b;  // helper segment, row range [a, b)
  uint32_t *start, *end;
  double *width, *pw, *beta, *cov, *beta_ref, *cov_ref, *diff;
  int32_t *num, *num_ref;
};

// mean (R's two-pass) of the values of one side over rows [a, b), by a warp; also coverage sum and count
__device__ __forceinline__ void dc_side_mean(const double* __restrict__ scov, const double* __restrict__ sbeta, int64_t a,
                                             int64_t b, int lane, double& mean, double& csum, int& m) {
  DD sum{0.0, 0.0};
  csum = 0.0;
  m = 0;
  for (int64_t i = a + lane; i < b; i += 32) {
    const double c = scov[i];
    if (c > 0) { dd_add(sum, sbeta[i]); csum += c; ++m; }
  }
  sum = dd_warp_sum(sum);
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    csum += __shfl_xor_sync(0xffffffffu, csum, d);
    m += __shfl_xor_sync(0xffffffffu, m, d);
  }
  mean = (sum.hi + sum.lo) / m;                  // NaN for an empty side (R: NA)
  if (is_finite(mean)) {
    DD t{0.0, 0.0};
    for (int64_t i = a + lane; i < b; i += 32) if (scov[i] > 0) dd_add(t, sbeta[i] - mean);
    t = dd_warp_sum(t);
    mean += (t.hi + t.lo) / m;
  }
}

__global__ void __launch_bounds__(256)
k_dc_emit_part(const SegGroup* __restrict__ groups, const int32_t* __restrict__ seg_group, int n_parts,
               const int32_t* __restrict__ row_hi, const double* __restrict__ seg_pw, const double* __restrict__ pos,
               const int32_t* __restrict__ parts, const int32_t* __restrict__ part_off,
               const int32_t* __restrict__ part_first, const int32_t* __restrict__ part_seg,
               const double* __restrict__ xcov, const double* __restrict__ xbeta, const double* __restrict__ rcov,
               const double* __restrict__ rbeta, DcParts out) {
  const int o = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  const int lane = threadIdx.x & 31;
  if (o >= n_parts) return;
  const int s = part_seg[o];
  const SegGroup G = groups[seg_group[s]];
  const int part = o - part_off[s];
  const int64_t a = part_first[o], b = part + 1 < parts[s] ? (int64_t)part_first[o + 1] : (int64_t)row_hi[s];
  int64_t last = a;
  for (int64_t i = a + lane; i < b; i += 32) if (xcov[i] > 0 || rcov[i] > 0) last = i;
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) last = max(last, (int64_t)__shfl_xor_sync(0xffffffffu, (long long)last, d));
  double mx, cx, mr, cr;
  int nx, nr;
  dc_side_mean(xcov, xbeta, a, b, lane, mx, cx, nx);
  dc_side_mean(rcov, rbeta, a, b, lane, mr, cr, nr);
  if (lane == 0) {
    const double p0 = pos[G.start0 + (a - G.row0)], p1 = pos[G.start0 + (last - G.row0)];
    out.seg[o] = s;
    out.a[o] = (int32_t)a;
    out.b[o] = (int32_t)b;
    out.start[o] = (uint32_t)p0;
    out.end[o] = (uint32_t)p1 + 2u;
    out.width[o] = (p1 + 2.0) - p0 + 1.0;
    out.pw[o] = seg_pw[s];
    out.beta[o] = mx;
    out.cov[o] = cx;
    out.num[o] = nx;
    out.beta_ref[o] = mr;
    out.cov_ref[o] = cr;
    out.num_ref[o] = nr;
    out.diff[o] = mx - mr;
  }
}

// R/call.R:301-303: a part opens a new group unless it continues its predecessor
__global__ void k_dc_group_flags(int n_parts, const int32_t* __restrict__ part_seg, const int32_t* __restrict__ seg_group,
                                 const double* __restrict__ diff, const uint32_t* __restrict__ start,
                                 const uint32_t* __restrict__ end, double min_diff, double min_delta_diff,
                                 double max_distance, int32_t* __restrict__ flags) {
  const int o = blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= n_parts) return;
  bool change = true;
  if (o > 0 && seg_group[part_seg[o]] == seg_group[part_seg[o - 1]]) {
    const double d = diff[o], q = diff[o - 1];
    // any NaN makes a comparison false here and NA in R; both end as group_change = TRUE (:303)
    const bool cont = fabs(d - q) <= min_delta_diff && fabs(d) >= min_diff && fabs(q) >= min_diff &&
                      (double)start[o] - (double)end[o - 1] < max_distance;
    change = !cont;
  }
  flags[o] = change ? 1 : 0;
}

struct DcOut {           // group_call, device arrays
  int32_t *job, *condition, *estimate_id, *segment_id;
  uint32_t *start, *end;
  double *width, *beta, *std, *coverage, *beta_ref, *std_ref, *coverage_ref, *pseudoweight, *diff, *cohen_d, *pval,
      *coverage_score;
  int32_t *num_cpgs, *num_cpgs_ref;
};

// mean and sd (R's two-pass long double algorithms, in double-double) of one side's values over the parts of a group
__device__ __forceinline__ void dc_group_side(const DcParts& P, int p0, int p1, const double* __restrict__ scov,
                                              const double* __restrict__ sbeta, int lane, int m, double& mean, double& sd) {
  mean = NAN;
  sd = NAN;
  if (m <= 0) return;
  DD sum{0.0, 0.0};
  for (int p = p0; p < p1; ++p)
    for (int64_t i = P.a[p] + lane; i < P.b[p]; i += 32) if (scov[i] > 0) dd_add(sum, sbeta[i]);
  sum = dd_warp_sum(sum);
  mean = (sum.hi + sum.lo) / m;
  if (is_finite(mean)) {
    DD t{0.0, 0.0};
    for (int p = p0; p < p1; ++p)
      for (int64_t i = P.a[p] + lane; i < P.b[p]; i += 32) if (scov[i] > 0) dd_add(t, sbeta[i] - mean);
    t = dd_warp_sum(t);
    mean += (t.hi + t.lo) / m;
  }
  if (m >= 2) {
    DD ss{0.0, 0.0};
    for (int p = p0; p < p1; ++p)
      for (int64_t i = P.a[p] + lane; i < P.b[p]; i += 32) if (scov[i] > 0) { const double d = sbeta[i] - mean; dd_add(ss, d * d); }
    ss = dd_warp_sum(ss);
    sd = sqrt((ss.hi + ss.lo) / (m - 1));
  }
}

// R/call.R:305-314, one warp per group
__global__ void __launch_bounds__(256)
k_dc_group_stats(int n_groups, int n_parts, const int32_t* __restrict__ grp_first, const int32_t* __restrict__ seg_group,
                 const SegGroup* __restrict__ groups, const int32_t* __restrict__ group_first_grp,
                 const int32_t* __restrict__ cond_of_group, DcParts P, const double* __restrict__ xcov,
                 const double* __restrict__ xbeta, const double* __restrict__ rcov, const double* __restrict__ rbeta,
                 DcOut out) {
  const int q = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  const int lane = threadIdx.x & 31;
  if (q >= n_groups) return;
  const int p0 = grp_first[q], p1 = q + 1 < n_groups ? grp_first[q + 1] : n_parts;
  const int sg = seg_group[P.seg[p0]];
  uint32_t st = 0xffffffffu, en = 0;
  double width = 0.0, cx = 0.0, cr = 0.0, pw = 0.0;
  int nx = 0, nr = 0;
  for (int p = p0 + lane; p < p1; p += 32) {
    st = min(st, P.start[p]);
    en = max(en, P.end[p]);
    width += P.width[p];
    cx += P.cov[p];
    cr += P.cov_ref[p];
    pw += P.pw[p];
    nx += P.num[p];
    nr += P.num_ref[p];
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {   // integer-valued doubles: exact in any order
    st = min(st, __shfl_xor_sync(0xffffffffu, st, d));
    en = max(en, __shfl_xor_sync(0xffffffffu, en, d));
    width += __shfl_xor_sync(0xffffffffu, width, d);
    cx += __shfl_xor_sync(0xffffffffu, cx, d);
    cr += __shfl_xor_sync(0xffffffffu, cr, d);
    pw += __shfl_xor_sync(0xffffffffu, pw, d);
    nx += __shfl_xor_sync(0xffffffffu, nx, d);
    nr += __shfl_xor_sync(0xffffffffu, nr, d);
  }
  double mx, sx, mr, sr;
  dc_group_side(P, p0, p1, xcov, xbeta, lane, nx, mx, sx);
  dc_group_side(P, p0, p1, rcov, rbeta, lane, nr, mr, sr);
  if (lane == 0) {
    const SegGroup G = groups[sg];
    out.job[q] = G.job;
    out.condition[q] = cond_of_group[sg];
    out.estimate_id[q] = G.estimate_id;
    out.segment_id[q] = q - group_first_grp[sg] + 1;
    out.start[q] = st;
    out.end[q] = en;
    out.width[q] = width;
    out.beta[q] = mx;
    out.std[q] = sx;
    out.coverage[q] = cx;
    out.num_cpgs[q] = nx;
    out.beta_ref[q] = mr;
    out.std_ref[q] = sr;
    out.coverage_ref[q] = cr;
    out.num_cpgs_ref[q] = nr;
    out.pseudoweight[q] = pw;
    const double diff = mx - mr;
    out.diff[q] = diff;
    out.cohen_d[q] = diff / sqrt((((double)nx - 1) * (sx * sx) + ((double)nr - 1) * (sr * sr)) / ((double)nx + (double)nr - 2));
  }
}

// first group of every segment group: group_first_grp[sg] = gid of the part c0[sg]
__global__ void k_dc_group_first(int n_seg_groups, const int32_t* __restrict__ c0, const int32_t* __restrict__ part_gid,
                                 int n_parts, int n_groups, int32_t* __restrict__ group_first_grp) {
  const int sg = blockIdx.x * blockDim.x + threadIdx.x;
  if (sg >= n_seg_groups) return;
  const int p = c0[sg];
  group_first_grp[sg] = p < n_parts ? part_gid[p] : n_groups;
}

__global__ void k_dc_ptr64(int n, const int32_t* __restrict__ off, int32_t total, int64_t* __restrict__ ptr) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) ptr[i] = off[i];
  if (i == n) ptr[n] = total;
}

// betavals / betavals.ref of every group, concatenated (the order inside a group does not matter to the rank sum)
__global__ void __launch_bounds__(256)
k_dc_fill_values(int n_groups, int n_parts, const int32_t* __restrict__ grp_first, DcParts P,
                 const double* __restrict__ xcov, const double* __restrict__ xbeta, const double* __restrict__ rcov,
                 const double* __restrict__ rbeta, const int64_t* __restrict__ x_ptr, const int64_t* __restrict__ y_ptr,
                 double* __restrict__ xv, double* __restrict__ yv) {
  const int q = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  const int lane = threadIdx.x & 31;
  if (q >= n_groups) return;
  const int p0 = grp_first[q], p1 = q + 1 < n_groups ? grp_first[q + 1] : n_parts;
  int64_t ox = x_ptr[q], oy = y_ptr[q];
  const unsigned below = (1u << lane) - 1u;
  for (int p = p0; p < p1; ++p)
    for (int64_t base = P.a[p]; base < P.b[p]; base += 32) {
      const int64_t i = base + lane;
      const bool hx = i < P.b[p] && xcov[i] > 0, hr = i < P.b[p] && rcov[i] > 0;
      const unsigned bx = __ballot_sync(0xffffffffu, hx), br = __ballot_sync(0xffffffffu, hr);
      if (hx) xv[ox + __popc(bx & below)] = xbeta[i];
      if (hr) yv[oy + __popc(br & below)] = rbeta[i];
      ox += __popc(bx);
      oy += __popc(br);
    }
}

// R/call.R:316-319: the test only where both sides have coverage; NA -> 1
__global__ void k_dc_pval_finish(int n_groups, const double* __restrict__ cx, const double* __restrict__ cr,
                                 double* __restrict__ pval) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n_groups) return;
  const double p = pval[q];
  if (!(cx[q] > 0 && cr[q] > 0) || p != p) pval[q] = 1.0;
}

// R/call.R:322-331: mean over the job's positions in [start - 1, end] of dscount * 100 / D, one warp per group
__global__ void __launch_bounds__(256)
k_dc_cov_score(int n_groups, const DcJob* __restrict__ jobs, const int32_t* __restrict__ slot_of_seg_group,
               const int32_t* __restrict__ grp_first, const int32_t* __restrict__ part_seg_of_part,
               const int32_t* __restrict__ seg_group, const double* __restrict__ pos, const int32_t* __restrict__ dscount,
               const uint32_t* __restrict__ start, const uint32_t* __restrict__ end, double* __restrict__ score) {
  const int q = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  const int lane = threadIdx.x & 31;
  if (q >= n_groups) return;
  const DcJob J = jobs[slot_of_seg_group[seg_group[part_seg_of_part[grp_first[q]]]]];
  const double* s = pos + J.start0;
  const double want_lo = (double)start[q] - 1.0, want_hi = (double)end[q];
  int a = 0, b = J.P;
  while (a < b) { const int mid = (a + b) >> 1; if (s[mid] < want_lo) a = mid + 1; else b = mid; }
  const int lo = a;
  b = J.P;
  while (a < b) { const int mid = (a + b) >> 1; if (s[mid] <= want_hi) a = mid + 1; else b = mid; }
  const int hi = a, m = hi - lo;
  const int32_t* dc = dscount + J.start0;
  DD sum{0.0, 0.0};
  for (int k = lo + lane; k < hi; k += 32) dd_add(sum, (double)dc[k] * 100 / (double)J.D);
  sum = dd_warp_sum(sum);
  double mean = (sum.hi + sum.lo) / m;
  if (is_finite(mean)) {
    DD t{0.0, 0.0};
    for (int k = lo + lane; k < hi; k += 32) dd_add(t, (double)dc[k] * 100 / (double)J.D - mean);
    t = dd_warp_sum(t);
    mean += (t.hi + t.lo) / m;
  }
  if (lane == 0) score[q] = mean;
}

}  // namespace

void DiffCallTable::alloc(cudaStream_t st, int64_t count) {
  n = count;
  const size_t c = (size_t)std::max<int64_t>(1, count);
  job.alloc(st, c); condition.alloc(st, c); estimate_id.alloc(st, c); segment_id.alloc(st, c);
  num_cpgs.alloc(st, c); num_cpgs_ref.alloc(st, c);
  start.alloc(st, c); end.alloc(st, c);
  width.alloc(st, c); beta.alloc(st, c); std_.alloc(st, c); coverage.alloc(st, c); beta_ref.alloc(st, c);
  std_ref.alloc(st, c); coverage_ref.alloc(st, c); pseudoweight.alloc(st, c); diff.alloc(st, c); cohen_d.alloc(st, c);
  pval.alloc(st, c); coverage_score.alloc(st, c);
}

int64_t call_differences_device(Engine& eng, const Batch& B, const Result& R, int ref_condition, double min_diff,
                                double min_delta_diff, double tol_val, double max_distance, DiffCallTable& out) {
  cudaStream_t st = eng.stream;
  // ---- jobs, comparisons, tiles (host) ----------------------------------------------------------------------
  std::vector<DcJob> jobs;
  std::vector<DcGroup> dgroups;
  std::vector<SegGroup> sgroups;
  std::vector<int32_t> ent_job, ent_d, ptile_job, ptile_k0, gtile_group, gtile_k0, cond_of_group, slot_of_group;
  int64_t tab = 0;
  int32_t dmeta = 0;
  std::vector<int32_t> h_dcond;      // condition of every dataset of every valid job (run table of the batch)
  if (!R.rowstart.p || R.tab0.size() != R.jobs.size())
    throw MlError(19, "call_differences needs a FAST fit (the row table of the unique-position leg)");
  for (size_t jn = 0; jn < R.jobs.size(); ++jn) {
    const JobHost& j = R.jobs[jn];
    if (j.status) continue;
    if (j.E != j.C) throw MlError(19, "call_differences needs the result of difference_detection_fast (one estimate per condition)");
    if (j.C < 2) throw MlError(14, "Needs at least two conditions!");
    if (ref_condition < 1 || ref_condition > j.C) throw MlError(15, "Invalid reference condition!");
    const int slot = (int)jobs.size();
    // tab0: the (dataset, parameter) -> row table the fit itself built (Result::rowstart); on the FAST leg a bin of a
    // dataset holds exactly one row, so it is mdata's join of the raw rows onto the union of positions
    jobs.push_back(DcJob{B.dev_row0[jn], R.start0[jn], R.tab0[jn], (int32_t)j.nrows, (int32_t)j.P, j.D, dmeta});
    for (int d = 0; d < j.D; ++d) h_dcond.push_back(B.dcond[B.run0[jn] + d]);
    tab += (int64_t)j.D * j.P;
    dmeta += j.D;
    for (int d = 0; d < j.D; ++d) { ent_job.push_back(slot); ent_d.push_back(d); }
    for (int64_t k = 0; k < j.P; k += 256) { ptile_job.push_back(slot); ptile_k0.push_back((int32_t)k); }
    int rank = 0;
    for (int c = 1; c <= j.C; ++c) {
      if (c == ref_condition) continue;
      const int e = ++rank;                      // estimate e + 1 of the table <-> rank-th non-reference condition (:270)
      if (e >= j.E) break;
      const int g = (int)dgroups.size();
      dgroups.push_back(DcGroup{j.par0 + (int64_t)e * j.P, slot, c, ref_condition, 0});
      sgroups.push_back(SegGroup{j.par0 + (int64_t)e * j.P, j.P, R.start0[jn], e + 1, (int32_t)jn});
      cond_of_group.push_back(c);
      slot_of_group.push_back(slot);
      for (int64_t k = 0; k < j.P; k += 256) { gtile_group.push_back(g); gtile_k0.push_back((int32_t)k); }
    }
  }
  out.alloc(st, 0);
  if (dgroups.empty()) return 0;
  const int n_jobs = (int)jobs.size(), ng = (int)dgroups.size();
  DevBuf<DcJob> d_jobs(st, jobs.size());
  d_jobs.upload(jobs);
  DevBuf<DcGroup> d_dgroups(st, dgroups.size());
  d_dgroups.upload(dgroups);
  DevBuf<int32_t> d_ent_job(st, ent_job.size()), d_ent_d(st, ent_d.size()), d_ptj(st, ptile_job.size()),
      d_ptk(st, ptile_k0.size()), d_gtg(st, gtile_group.size()), d_gtk(st, gtile_k0.size()), d_cond(st, ng), d_slot(st, ng);
  d_ent_job.upload(ent_job); d_ent_d.upload(ent_d); d_ptj.upload(ptile_job); d_ptk.upload(ptile_k0);
  d_gtg.upload(gtile_group); d_gtk.upload(gtile_k0); d_cond.upload(cond_of_group); d_slot.upload(slot_of_group);
  // ---- mdata ----------------------------------------------------------------------------------------------
  DevBuf<int32_t> dcond(st, (size_t)std::max(1, dmeta)), dscount(st, (size_t)std::max<int64_t>(1, R.nstart));
  dcond.upload(h_dcond);
  (void)tab;
  const DevBuf<int32_t>& rowidx = R.rowstart;
  ML_LAUNCH(eng, "k_dc_dscount", k_dc_dscount<<<(int)ptile_job.size(), 256, 0, st>>>(d_jobs.p, d_ptj.p, d_ptk.p, rowidx.p, dscount.p));
  DevBuf<double> xcov(st, (size_t)R.npar), xbeta(st, (size_t)R.npar), rcov(st, (size_t)R.npar), rbeta(st, (size_t)R.npar);
  ML_LAUNCH(eng, "k_dc_pool", k_dc_pool<<<(int)gtile_group.size(), 256, 0, st>>>(
      d_jobs.p, d_dgroups.p, d_gtg.p, d_gtk.p, dcond.p, rowidx.p, B.nmeth.p, B.cov.p, xcov.p, xbeta.p, rcov.p, rbeta.p));
  // ---- diff_call: segments of the difference estimates (:273) ------------------------------------------------
  SegDeviceOut seg;
  const int64_t n_segs64 = segment_device(eng, sgroups, R.npar, R.start.p, 1, R.beta.p, R.betahat.p, R.pw.p, tol_val, seg);
  const int n_segs = (int)n_segs64;
  if (n_segs == 0) return 0;
  // ---- parts (:278-288), as in call_table_impl ----------------------------------------------------------------
  DevBuf<SegGroup> d_groups(st, ng);
  d_groups.upload(sgroups);
  DevBuf<int32_t> d_g0(st, ng + 1);
  d_g0.upload(seg.group_seg0);
  DevBuf<int32_t> parts(st, n_segs), gaps(st, n_segs), part_off(st, n_segs), gap_off(st, n_segs);
  DevBuf<int32_t> seg_group(st, n_segs), row_lo(st, n_segs), row_hi(st, n_segs);
  const RowsDiff rows{xcov.p, rcov.p};
  ML_LAUNCH(eng, "k_call_ranges", k_call_ranges<<<cdiv(n_segs, 128), 128, 0, st>>>(
      d_groups.p, d_g0.p, ng, n_segs, seg.start.p, seg.end.p, seg.first_row.p, R.start.p, 1, seg_group.p, row_lo.p, row_hi.p));
  ML_LAUNCH(eng, "k_call_count", k_call_count<RowsDiff><<<cdiv((long long)n_segs * 32, 256), 256, 0, st>>>(
      d_groups.p, seg_group.p, n_segs, row_lo.p, row_hi.p, R.start.p, 1, rows, max_distance, parts.p, gaps.p));
  Scan scan;
  scan.run(eng, gaps.p, n_segs, gap_off.p, nullptr);
  const int32_t n_parts = scan.run(eng, parts.p, n_segs, part_off.p, nullptr);
  if (n_parts == 0) return 0;
  DevBuf<int32_t> part_first(st, n_parts), part_seg(st, n_parts);
  ML_LAUNCH(eng, "k_call_parts", k_call_parts<RowsDiff><<<cdiv((long long)n_segs * 32, 256), 256, 0, st>>>(
      d_groups.p, seg_group.p, n_segs, row_lo.p, row_hi.p, R.start.p, 1, rows, max_distance, parts.p, part_off.p,
      part_first.p, part_seg.p, nullptr));
  // ---- seg_call (:291-298) ---------------------------------------------------------------------------------
  DevBuf<int32_t> p_seg(st, n_parts), p_a(st, n_parts), p_b(st, n_parts), p_num(st, n_parts), p_numr(st, n_parts);
  DevBuf<uint32_t> p_start(st, n_parts), p_end(st, n_parts);
  DevBuf<double> p_width(st, n_parts), p_pw(st, n_parts), p_beta(st, n_parts), p_cov(st, n_parts), p_betar(st, n_parts),
      p_covr(st, n_parts), p_diff(st, n_parts);
  const DcParts P{p_seg.p, p_a.p, p_b.p, p_start.p, p_end.p, p_width.p, p_pw.p, p_beta.p, p_cov.p, p_betar.p, p_covr.p,
                  p_diff.p, p_num.p, p_numr.p};
  ML_LAUNCH(eng, "k_dc_emit_part", k_dc_emit_part<<<cdiv((long long)n_parts * 32, 256), 256, 0, st>>>(
      d_groups.p, seg_group.p, n_parts, row_hi.p, seg.pseudoweight.p, R.start.p, parts.p, part_off.p, part_first.p,
      part_seg.p, xcov.p, xbeta.p, rcov.p, rbeta.p, P));
  // ---- groups (:301-314) -----------------------------------------------------------------------------------
  DevBuf<int32_t> gflag(st, n_parts), part_gid(st, n_parts), grp_first(st, n_parts);
  ML_LAUNCH(eng, "k_dc_group_flags", k_dc_group_flags<<<cdiv(n_parts, 256), 256, 0, st>>>(
      n_parts, part_seg.p, seg_group.p, p_diff.p, p_start.p, p_end.p, min_diff, min_delta_diff, max_distance, gflag.p));
  const int32_t n_groups = scan.run(eng, gflag.p, n_parts, part_gid.p, grp_first.p);
  DevBuf<int32_t> d_c0(st, ng + 1), group_first_grp(st, ng);
  ML_LAUNCH(eng, "k_call_group0", k_call_group0<<<cdiv(ng + 1, 128), 128, 0, st>>>(d_g0.p, ng, n_segs, part_off.p, n_parts, d_c0.p));
  ML_LAUNCH(eng, "k_dc_group_first", k_dc_group_first<<<cdiv(ng, 128), 128, 0, st>>>(ng, d_c0.p, part_gid.p, n_parts, n_groups,
                                                                                   group_first_grp.p));
  out.alloc(st, n_groups);
  const DcOut O{out.job.p, out.condition.p, out.estimate_id.p, out.segment_id.p, out.start.p, out.end.p, out.width.p,
                out.beta.p, out.std_.p, out.coverage.p, out.beta_ref.p, out.std_ref.p, out.coverage_ref.p,
                out.pseudoweight.p, out.diff.p, out.cohen_d.p, out.pval.p, out.coverage_score.p, out.num_cpgs.p,
                out.num_cpgs_ref.p};
  ML_LAUNCH(eng, "k_dc_group_stats", k_dc_group_stats<<<cdiv((long long)n_groups * 32, 256), 256, 0, st>>>(
      n_groups, n_parts, grp_first.p, seg_group.p, d_groups.p, group_first_grp.p, d_cond.p, P, xcov.p, xbeta.p, rcov.p,
      rbeta.p, O));
  // ---- Wilcoxon (:316-319) ---------------------------------------------------------------------------------
  DevBuf<int32_t> offx(st, n_groups), offy(st, n_groups);
  const int32_t tx = scan.run(eng, out.num_cpgs.p, n_groups, offx.p, nullptr);
  const int32_t ty = scan.run(eng, out.num_cpgs_ref.p, n_groups, offy.p, nullptr);
  DevBuf<int64_t> x_ptr(st, (size_t)n_groups + 1), y_ptr(st, (size_t)n_groups + 1);
  ML_LAUNCH(eng, "k_dc_ptr64", k_dc_ptr64<<<cdiv(n_groups + 1, 256), 256, 0, st>>>(n_groups, offx.p, tx, x_ptr.p));
  ML_LAUNCH(eng, "k_dc_ptr64", k_dc_ptr64<<<cdiv(n_groups + 1, 256), 256, 0, st>>>(n_groups, offy.p, ty, y_ptr.p));
  DevBuf<double> xv(st, (size_t)std::max(1, tx)), yv(st, (size_t)std::max(1, ty));
  ML_LAUNCH(eng, "k_dc_fill_values", k_dc_fill_values<<<cdiv((long long)n_groups * 32, 256), 256, 0, st>>>(
      n_groups, n_parts, grp_first.p, P, xcov.p, xbeta.p, rcov.p, rbeta.p, x_ptr.p, y_ptr.p, xv.p, yv.p));
  wilcoxon_groups_device(eng, n_groups, x_ptr.p, xv.p, y_ptr.p, yv.p, (int64_t)tx + ty, out.pval.p, nullptr);
  ML_LAUNCH(eng, "k_dc_pval_finish", k_dc_pval_finish<<<cdiv(n_groups, 256), 256, 0, st>>>(n_groups, out.coverage.p,
                                                                                       out.coverage_ref.p, out.pval.p));
  // ---- coverage score (:322-331) -----------------------------------------------------------------------------
  ML_LAUNCH(eng, "k_dc_cov_score", k_dc_cov_score<<<cdiv((long long)n_groups * 32, 256), 256, 0, st>>>(
      n_groups, d_jobs.p, d_slot.p, grp_first.p, part_seg.p, seg_group.p, R.start.p, dscount.p, out.start.p, out.end.p,
      out.coverage_score.p));
  stream_sync(st);
  return n_groups;
}

// valleys.inc.cuh — the valley (DMV) block of segment_methylation_singlechr, R/call.R:100-133, on the resident call
// table and its LMR / UMR categories.  Included at the end of segment.cu (it uses the scan defined there).
//
// Every statement of the block is a per-row (or per-run, per-valley) function from valley_core.cuh — the same inline
// functions tests/emu/emu_valleys.cpp runs on the CPU — between stream compactions:
//   rows -> runs of consecutive non-high rows (:102-110) -> UMRs outside every run (:113-116) -> items in key order
//   -> merged valleys (:119-124) -> admissible ones that contain a DMV (:125-129) -> rows they replace (:132-133).
namespace {

__global__ void k_dv_run_heads(ValleyRows c, ValleyParams p, int32_t* __restrict__ flag) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < c.n) flag[i] = vl_run_head(c, p, i);
}
__global__ void k_dv_run_stats(ValleyRows c, ValleyParams p, ValleyRuns R) {
  const int32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < R.n) vl_run_stats(c, p, R, r);
}
// item flag: the head row of a kept run, or an UMR that overlaps no kept run
__global__ void k_dv_item_rows(ValleyRows c, ValleyRuns R, int32_t* __restrict__ flag) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= c.n) return;
  const int32_t u = c.run_of_head[i];
  flag[i] = ((u >= 0 && R.keep[u]) || vl_umr_added(c, R, i)) ? 1 : 0;
}
__global__ void k_dv_item_heads(ValleyRows c, ValleyParams p, ValleyRuns R, ValleyItems I, int32_t* __restrict__ flag) {
  const int32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < I.n) flag[r] = vl_item_head(c, p, R, I, r);
}
__global__ void k_dv_merge(ValleyRows c, ValleyParams p, ValleyRuns R, ValleyItems I, const int32_t* __restrict__ heads,
                           int32_t n_heads, ValleyOut o, int32_t* __restrict__ flag) {
  const int32_t v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= n_heads) return;
  vl_merge(c, p, R, I, heads, n_heads, v, o);
  flag[v] = o.keep[v];
}
__global__ void k_dv_emit(int32_t n, const int32_t* __restrict__ list, ValleyOut in, ValleyOut out) {
  const int32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const int32_t v = list[k];
  out.job[k] = in.job[v]; out.est[k] = in.est[v]; out.ncpg[k] = in.ncpg[v];
  out.start[k] = in.start[v]; out.end[k] = in.end[v];
  out.width[k] = in.width[v]; out.beta[k] = in.beta[v]; out.cov[k] = in.cov[v]; out.pw[k] = in.pw[v];
  out.keep[k] = 1;
  out.row0[k] = in.row0[v];
}
__global__ void k_dv_row_in_valley(ValleyRows c, int32_t n_v, ValleyOut v, int8_t* __restrict__ row_in) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < c.n) row_in[i] = vl_row_in_valley(c, i, n_v, v.job, v.est, v.start, v.end);
}

}  // namespace

void ValleyTable::alloc(cudaStream_t st, int64_t count) {
  n = count;
  const size_t c = (size_t)std::max<int64_t>(1, count);
  job.alloc(st, c); condition.alloc(st, c); num_cpgs.alloc(st, c); start.alloc(st, c); end.alloc(st, c);
  width.alloc(st, c); beta.alloc(st, c); coverage.alloc(st, c); pseudoweight.alloc(st, c); keep.alloc(st, c); row0.alloc(st, c);
}
ValleyOut ValleyTable::view() {
  return ValleyOut{job.p, condition.p, num_cpgs.p, start.p, end.p, width.p, beta.p, coverage.p, pseudoweight.p, keep.p, row0.p};
}

void call_valleys_device(Engine& eng, const CallTable& t, const int8_t* d_category, const ValleyParams& p,
                         ValleyTable& out, DevBuf<int8_t>& row_in_valley, DevBuf<double>& low_w, DevBuf<double>& low_b) {
  cudaStream_t st = eng.stream;
  const int64_t n = t.t.n;
  const size_t cn = (size_t)std::max<int64_t>(1, n);
  row_in_valley.alloc(st, cn); low_w.alloc(st, cn); low_b.alloc(st, cn);
  out.alloc(st, 0);
  if (n == 0) return;
  if (n >= 0x7fffffff) throw MlError(19, "call table too large");
  ML_CUDA(cudaMemsetAsync(low_w.p, 0xff, (size_t)n * sizeof(double), st));     // all-ones is a NaN: NA
  ML_CUDA(cudaMemsetAsync(low_b.p, 0xff, (size_t)n * sizeof(double), st));
  ML_CUDA(cudaMemsetAsync(row_in_valley.p, 0, (size_t)n, st));
  DevBuf<int32_t> run_of_head(st, (size_t)n), flag(st, (size_t)n), run_head(st, (size_t)n);
  ML_CUDA(cudaMemsetAsync(run_of_head.p, 0xff, (size_t)n * sizeof(int32_t), st));   // -1
  const ValleyRows c{n, t.t.job.p, t.t.estimate_id.p, t.num_cpgs.p, t.t.start.p, t.t.end.p, t.t.beta.p, t.coverage.p,
                     t.t.pseudoweight.p, d_category, low_w.p, low_b.p, run_of_head.p};
  Scan scan;
  const int nb = cdiv(n, 256);
  // :102-110  runs
  ML_LAUNCH(eng, "k_dv_run_heads", k_dv_run_heads<<<nb, 256, 0, st>>>(c, p, flag.p));
  const int32_t n_runs = scan.run(eng, flag.p, n, nullptr, run_head.p);
  const size_t cr = (size_t)std::max(1, n_runs);
  DevBuf<int32_t> r_last(st, cr), r_ncpg(st, cr);
  DevBuf<double> r_cov(st, cr), r_pw(st, cr);
  DevBuf<int8_t> r_cat(st, cr), r_keep(st, cr);
  const ValleyRuns R{n_runs, run_head.p, r_last.p, r_ncpg.p, r_cov.p, r_pw.p, r_cat.p, r_keep.p};
  if (n_runs > 0) ML_LAUNCH(eng, "k_dv_run_stats", k_dv_run_stats<<<cdiv(n_runs, 128), 128, 0, st>>>(c, p, R));
  // :107-116  items of new_valleys in key order
  DevBuf<int32_t> item_row(st, (size_t)n);
  ML_LAUNCH(eng, "k_dv_item_rows", k_dv_item_rows<<<nb, 256, 0, st>>>(c, R, flag.p));
  const int32_t n_items = scan.run(eng, flag.p, n, nullptr, item_row.p);
  if (n_items == 0) { stream_sync(st); return; }
  const ValleyItems I{n_items, item_row.p};
  // :119-129  merged valleys
  DevBuf<int32_t> iflag(st, (size_t)n_items), vhead(st, (size_t)n_items);
  ML_LAUNCH(eng, "k_dv_item_heads", k_dv_item_heads<<<cdiv(n_items, 256), 256, 0, st>>>(c, p, R, I, iflag.p));
  const int32_t n_merged = scan.run(eng, iflag.p, n_items, nullptr, vhead.p);
  ValleyTable merged;
  merged.alloc(st, n_merged);
  DevBuf<int32_t> vflag(st, (size_t)n_merged), vlist(st, (size_t)n_merged);
  ML_LAUNCH(eng, "k_dv_merge", k_dv_merge<<<cdiv(n_merged, 128), 128, 0, st>>>(c, p, R, I, vhead.p, n_merged, merged.view(), vflag.p));
  const int32_t n_out = scan.run(eng, vflag.p, n_merged, nullptr, vlist.p);
  out.alloc(st, n_out);
  if (n_out > 0) {
    ML_LAUNCH(eng, "k_dv_emit", k_dv_emit<<<cdiv(n_out, 256), 256, 0, st>>>(n_out, vlist.p, merged.view(), out.view()));
    // :132-133  rows replaced by a valley
    ML_LAUNCH(eng, "k_dv_row_in_valley", k_dv_row_in_valley<<<nb, 256, 0, st>>>(c, n_out, out.view(), row_in_valley.p));
  }
  stream_sync(st);
}

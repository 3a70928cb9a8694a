// pmdstatus.inc.cuh — R/call.R:201-215 on the resident tables: the merged table of pmdsplit.inc.cuh annotated with the
// PMD table of pmdcall.inc.cuh.  Included at the end of segment.cu; per-element logic in pmdstatus_core.cuh.
namespace {

__global__ void k_st_count(MergedTable M, PmdRegions G, int32_t* __restrict__ cnt) {
  const int32_t m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M.n) return;
  int8_t a, b;
  cnt[m] = st_statuses(M, G, m, a, b);
}
// rows of merged row m start at pos[m] + (number of second statuses before m): cnt is 1 or 2, so with flag = cnt - 1 the
// scan of the flags gives the extra rows before m
__global__ void k_st_extra_flag(int32_t n, const int32_t* __restrict__ cnt, int32_t* __restrict__ flag) {
  const int32_t m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m < n) flag[m] = cnt[m] - 1;
}
__global__ void k_st_emit(MergedTable M, PmdRegions G, const int8_t* __restrict__ row_adm, const int32_t* __restrict__ extra_before,
                          FinalRows F) {
  const int32_t m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m < M.n) st_emit(M, G, row_adm, m, m + extra_before[m], F);
}
__global__ void k_st_adm_flag(int32_t n, FinalRows F, int32_t* __restrict__ flag) {
  const int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) flag[i] = F.adm[i] == 1 ? 1 : 0;
}
__global__ void k_st_umr_status(MergedTable M, FinalRows F, const int32_t* __restrict__ list, int32_t n_list, int8_t* __restrict__ out) {
  const int32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < n_list) out[list[r]] = st_umr_status(M, F, list, n_list, r);
}

// the labels of a final row are those of the merged row it annotates
__global__ void k_st_labels(int32_t n, const int32_t* __restrict__ merged, MergedTable M, int32_t* __restrict__ job,
                            int32_t* __restrict__ est, uint32_t* __restrict__ start, uint32_t* __restrict__ end,
                            int32_t* __restrict__ src_row, int32_t* __restrict__ src_valley) {
  const int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int32_t m = merged[i];
  job[i] = M.job[m]; est[i] = M.est[m]; start[i] = M.start[m]; end[i] = M.end[m];
  src_row[i] = M.src_row[m]; src_valley[i] = M.src_valley[m];
}

}  // namespace

void FinalDev::alloc(cudaStream_t st, int64_t count) {
  n = count;
  const size_t c = (size_t)std::max<int64_t>(1, count);
  merged.alloc(st, c); status.alloc(st, c); category.alloc(st, c); adm.alloc(st, c);
  job.alloc(st, c); est.alloc(st, c); src_row.alloc(st, c); src_valley.alloc(st, c); start.alloc(st, c); end.alloc(st, c);
}

void call_pmd_status_device(Engine& eng, MergedDev& Md, const PmdCallDev& pmd, const int8_t* d_row_adm, FinalDev& out) {
  cudaStream_t st = eng.stream;
  out.alloc(st, 0);
  const int32_t nm = (int32_t)Md.n;
  if (nm == 0) return;
  const MergedTable M = Md.view();
  const PmdRegions G{(int32_t)pmd.n, pmd.job.p, pmd.est.p, pmd.start.p, pmd.end.p, pmd.category.p};
  const int nb = cdiv(nm, 256);
  DevBuf<int32_t> cnt(st, (size_t)nm), flag(st, (size_t)nm), extra(st, (size_t)nm);
  ML_LAUNCH(eng, "k_st_count", k_st_count<<<nb, 256, 0, st>>>(M, G, cnt.p));
  ML_LAUNCH(eng, "k_st_extra_flag", k_st_extra_flag<<<nb, 256, 0, st>>>(nm, cnt.p, flag.p));
  Scan scan;
  const int32_t n_extra = scan.run(eng, flag.p, nm, extra.p, nullptr);
  const int32_t nf = nm + n_extra;
  out.alloc(st, nf);
  const FinalRows F{out.merged.p, out.status.p, out.category.p, out.adm.p};
  ML_LAUNCH(eng, "k_st_emit", k_st_emit<<<nb, 256, 0, st>>>(M, G, d_row_adm, extra.p, F));
  // :211-212 on the admissible rows; the new statuses are computed from the old ones, then copied over
  DevBuf<int32_t> aflag(st, (size_t)nf), alist(st, (size_t)nf);
  ML_LAUNCH(eng, "k_st_adm_flag", k_st_adm_flag<<<cdiv(nf, 256), 256, 0, st>>>(nf, F, aflag.p));
  const int32_t n_adm = scan.run(eng, aflag.p, nf, nullptr, alist.p);
  if (n_adm > 0) {
    DevBuf<int8_t> ns(st, (size_t)nf);
    ML_CUDA(cudaMemcpyAsync(ns.p, out.status.p, (size_t)nf, cudaMemcpyDeviceToDevice, st));
    ML_LAUNCH(eng, "k_st_umr_status", k_st_umr_status<<<cdiv(n_adm, 256), 256, 0, st>>>(M, F, alist.p, n_adm, ns.p));
    ML_CUDA(cudaMemcpyAsync(out.status.p, ns.p, (size_t)nf, cudaMemcpyDeviceToDevice, st));
  }
  ML_LAUNCH(eng, "k_st_labels", k_st_labels<<<cdiv(nf, 256), 256, 0, st>>>(nf, out.merged.p, M, out.job.p, out.est.p, out.start.p,
                                                                           out.end.p, out.src_row.p, out.src_valley.p));
  stream_sync(st);
}

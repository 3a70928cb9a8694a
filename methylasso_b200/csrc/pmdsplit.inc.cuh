// pmdsplit.inc.cuh — R/call.R:131-164 on the resident tables: the call table with its categories, the valleys of
// valleys.inc.cuh and the PMD candidate segments of pmd_segments_device (std_call).  Included at the end of segment.cu.
// Every step is a per-element function of pmdsplit_core.cuh (also run on the CPU by tests/emu/emu_valleys.cpp)
// between stream compactions:  rows + valleys -> merged table (:131-140) -> regions (:143-156) -> pieces (:158-164).
namespace {

__global__ void k_ps_keep_flag(int64_t n, const int8_t* __restrict__ row_in_valley, int32_t* __restrict__ flag) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) flag[i] = row_in_valley[i] ? 0 : 1;
}
__global__ void k_ps_place_rows(ValleyRows c, const int8_t* __restrict__ row_in_valley, const int32_t* __restrict__ kept_before,
                                int32_t n_v, const int32_t* __restrict__ v_row0, MergedTable M) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= c.n || row_in_valley[i]) return;
  ms_put_row(c, M, ms_row_slot(i, kept_before[i], n_v, v_row0), i);
}
__global__ void k_ps_place_valleys(int32_t n_v, ValleyOut v, const int32_t* __restrict__ kept_before, MergedTable M) {
  const int32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n_v) ms_put_valley(v, M, k + kept_before[v.row0[k]], k);
}
__global__ void k_ps_gap_heads(MergedTable M, double pvw, int32_t* __restrict__ ghead) {
  const int32_t m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M.n) return;
  M.follows_large_gap[m] = ms_gap(M, pvw, m);
  ghead[m] = ms_group_head(M, m);
}
// 1-based index inside the group: heads[] = compacted group heads, pos[] = exclusive count of heads
__global__ void k_ps_seq_ids(int32_t n, const int32_t* __restrict__ flag, const int32_t* __restrict__ pos,
                             const int32_t* __restrict__ heads, int32_t* __restrict__ seq) {
  const int32_t m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m < n) seq[m] = m - heads[pos[m] + flag[m] - 1] + 1;
}
__global__ void k_ps_region_heads(MergedTable M, int split_pmds, int32_t* __restrict__ flag) {
  const int32_t m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m < M.n) flag[m] = ms_region_head(M, split_pmds, m);
}
__global__ void k_ps_regions(MergedTable M, int split_pmds, Regions G) {
  const int32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < G.n) ms_region(M, split_pmds, G, r);
}
__global__ void k_ps_count(Regions G, PmdSegs X, int32_t* __restrict__ cnt) {
  const int32_t x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x < X.n) cnt[x] = ms_count_pieces(G, X, x);
}
// exclusive prefix of per-segment piece counts (one block; the PMD segment table has ~10^3 rows per genome)
__global__ void __launch_bounds__(256) k_ps_offsets(int32_t n, const int32_t* __restrict__ cnt, int32_t* __restrict__ off,
                                                    int32_t* __restrict__ total) {
  __shared__ int32_t sm[256];
  int32_t carry = 0;
  for (int32_t base = 0; base < n; base += 256) {
    const int32_t i = base + threadIdx.x;
    const int32_t v = i < n ? cnt[i] : 0;
    sm[threadIdx.x] = v;
    __syncthreads();
    for (int d = 1; d < 256; d <<= 1) {
      const int32_t t = threadIdx.x >= d ? sm[threadIdx.x - d] : 0;
      __syncthreads();
      sm[threadIdx.x] += t;
      __syncthreads();
    }
    if (i < n) off[i] = carry + sm[threadIdx.x] - v;
    carry += sm[255];
    __syncthreads();
  }
  if (threadIdx.x == 0) *total = carry;
}
__global__ void k_ps_unmatched(PmdSegs X, const int32_t* __restrict__ cnt, int32_t* __restrict__ na_group) {
  const int32_t x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x < X.n) ms_group_unmatched(X, cnt, x, na_group);
}
// one warp per segment: a PMD-sized segment is cut into thousands of pieces
__global__ void k_ps_emit(Regions G, PmdSegs X, const int32_t* __restrict__ cnt, const int32_t* __restrict__ off,
                          const int32_t* __restrict__ na_group, Pieces P) {
  const int32_t x = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (x < X.n) ms_emit_pieces(G, X, x, off[x], na_group[x], P, cnt[x], (int32_t)(threadIdx.x & 31), 32);
}
// :164  seq_len(.N) per estimate_id, counted behind the region-less segments that sort first
__global__ void k_ps_piece_ids(int32_t n, const int32_t* __restrict__ flag, const int32_t* __restrict__ pos,
                               const int32_t* __restrict__ heads, const int32_t* __restrict__ na, int32_t* __restrict__ seq) {
  const int32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n) seq[k] = k - heads[pos[k] + flag[k] - 1] + 1 + na[k];
}
__global__ void k_ps_piece_heads(int32_t n, Pieces P, int32_t* __restrict__ flag) {
  const int32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n) flag[k] = (k == 0 || P.job[k] != P.job[k - 1] || P.est[k] != P.est[k - 1]) ? 1 : 0;
}

}  // namespace

void MergedDev::alloc(cudaStream_t st, int64_t count) {
  n = count;
  const size_t c = (size_t)std::max<int64_t>(1, count);
  job.alloc(st, c); est.alloc(st, c); src_row.alloc(st, c); src_valley.alloc(st, c); segment_id.alloc(st, c);
  start.alloc(st, c); end.alloc(st, c); category.alloc(st, c); follows_large_gap.alloc(st, c);
}
MergedTable MergedDev::view() {
  return MergedTable{(int32_t)n, job.p, est.p, src_row.p, src_valley.p, segment_id.p, start.p, end.p, category.p, follows_large_gap.p};
}
void PiecesDev::alloc(cudaStream_t st, int64_t count) {
  n = count;
  const size_t c = (size_t)std::max<int64_t>(1, count);
  job.alloc(st, c); est.alloc(st, c); segment_id.alloc(st, c); start.alloc(st, c); end.alloc(st, c);
  fused_std.alloc(st, c); pw.alloc(st, c);
}

void call_pmd_split_device(Engine& eng, const CallTable& t, const int8_t* d_category, ValleyTable& val,
                           const int8_t* d_row_in_valley, const SegDeviceOut& pmd, double pmd_valley_min_width,
                           int split_pmds, MergedDev& M, PiecesDev& P) {
  cudaStream_t st = eng.stream;
  const int64_t n = t.t.n;
  M.alloc(st, 0);
  P.alloc(st, 0);
  if (n == 0) return;
  if (n >= 0x7fffffff) throw MlError(19, "call table too large");
  const ValleyRows c{n, t.t.job.p, t.t.estimate_id.p, t.num_cpgs.p, t.t.start.p, t.t.end.p, t.t.beta.p, t.coverage.p,
                     t.t.pseudoweight.p, d_category, nullptr, nullptr, nullptr};
  Scan scan;
  const int nb = cdiv(n, 256);
  const int32_t n_v = (int32_t)val.n;
  // :131-140  merged table
  DevBuf<int32_t> flag(st, (size_t)n), kept_before(st, (size_t)n);
  ML_LAUNCH(eng, "k_ps_keep_flag", k_ps_keep_flag<<<nb, 256, 0, st>>>(n, d_row_in_valley, flag.p));
  const int32_t n_keep = scan.run(eng, flag.p, n, kept_before.p, nullptr);
  const int32_t nm = n_keep + n_v;
  M.alloc(st, nm);
  if (nm == 0) { stream_sync(st); return; }
  const MergedTable Mv = M.view();
  ML_LAUNCH(eng, "k_ps_place_rows", k_ps_place_rows<<<nb, 256, 0, st>>>(c, d_row_in_valley, kept_before.p, n_v, val.row0.p, Mv));
  if (n_v > 0) ML_LAUNCH(eng, "k_ps_place_valleys", k_ps_place_valleys<<<cdiv(n_v, 256), 256, 0, st>>>(n_v, val.view(), kept_before.p, Mv));
  const int nbm = cdiv(nm, 256);
  DevBuf<int32_t> mflag(st, (size_t)nm), mpos(st, (size_t)nm), mheads(st, (size_t)nm);
  ML_LAUNCH(eng, "k_ps_gap_heads", k_ps_gap_heads<<<nbm, 256, 0, st>>>(Mv, pmd_valley_min_width, mflag.p));
  scan.run(eng, mflag.p, nm, mpos.p, mheads.p);
  ML_LAUNCH(eng, "k_ps_seq_ids", k_ps_seq_ids<<<nbm, 256, 0, st>>>(nm, mflag.p, mpos.p, mheads.p, M.segment_id.p));
  // :143-156  regions
  ML_LAUNCH(eng, "k_ps_region_heads", k_ps_region_heads<<<nbm, 256, 0, st>>>(Mv, split_pmds, mflag.p));
  const int32_t n_reg = scan.run(eng, mflag.p, nm, nullptr, mheads.p);
  DevBuf<int32_t> g_job(st, (size_t)n_reg), g_est(st, (size_t)n_reg);
  DevBuf<uint32_t> g_start(st, (size_t)n_reg), g_end(st, (size_t)n_reg);
  DevBuf<int8_t> g_iso(st, (size_t)n_reg);
  const Regions G{n_reg, mheads.p, g_job.p, g_est.p, g_start.p, g_end.p, g_iso.p};
  ML_LAUNCH(eng, "k_ps_regions", k_ps_regions<<<cdiv(n_reg, 128), 128, 0, st>>>(Mv, split_pmds, G));
  // :158-164  pieces of the PMD candidate segments
  const int32_t nx = (int32_t)pmd.n;
  if (nx == 0) { stream_sync(st); return; }
  const PmdSegs X{nx, pmd.job.p, pmd.estimate_id.p, pmd.start.p, pmd.end.p, pmd.beta.p, pmd.pseudoweight.p};
  DevBuf<int32_t> cnt(st, (size_t)nx), off(st, (size_t)nx), total(st, 1), na_group(st, (size_t)nx);
  ML_LAUNCH(eng, "k_ps_count", k_ps_count<<<cdiv(nx, 128), 128, 0, st>>>(G, X, cnt.p));
  ML_LAUNCH(eng, "k_ps_unmatched", k_ps_unmatched<<<cdiv(nx, 128), 128, 0, st>>>(X, cnt.p, na_group.p));
  ML_LAUNCH(eng, "k_ps_offsets", k_ps_offsets<<<1, 256, 0, st>>>(nx, cnt.p, off.p, total.p));
  int32_t np = 0;
  total.download(&np, 1);
  stream_sync(st);
  P.alloc(st, np);
  if (np == 0) return;
  DevBuf<int32_t> p_na(st, (size_t)np);
  const Pieces Pv{P.job.p, P.est.p, p_na.p, P.start.p, P.end.p, P.fused_std.p, P.pw.p};
  ML_LAUNCH(eng, "k_ps_emit", k_ps_emit<<<cdiv(32LL * nx, 128), 128, 0, st>>>(G, X, cnt.p, off.p, na_group.p, Pv));
  DevBuf<int32_t> pflag(st, (size_t)np), ppos(st, (size_t)np), pheads(st, (size_t)np);
  ML_LAUNCH(eng, "k_ps_piece_heads", k_ps_piece_heads<<<cdiv(np, 256), 256, 0, st>>>(np, Pv, pflag.p));
  scan.run(eng, pflag.p, np, ppos.p, pheads.p);
  ML_LAUNCH(eng, "k_ps_piece_ids", k_ps_piece_ids<<<cdiv(np, 256), 256, 0, st>>>(np, pflag.p, ppos.p, pheads.p, p_na.p, P.segment_id.p));
  stream_sync(st);
}

// pmdcall.inc.cuh — R/call.R:169-199 (segment_methylation_singlechr): the pooled per-position data gathered along
// split_call (:169-175), the PMD calls (:177-180), their merge (:182-190) and the statistics of the merged regions
// (:193-198).  Included at the end of segment.cu: it reuses the call table's row accessors (RowsFromEstimates: the
// pooled Nmeth / coverage of a FAST result), its foverlaps row search (segment_rows) and its double-double sums, which
// stand in for R's long double accumulators as they do for the call table (results equal to the last bit or so).
namespace {

constexpr int PC_THREADS = 128;

// sum over the block; every thread receives the total (fixed combination order: lanes by shuffle, warps in order)
__device__ __forceinline__ DD pc_block_sum(DD v, DD* sm) {
  v = dd_warp_sum(v);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) sm[wid] = v;
  __syncthreads();
  DD t = sm[0];
  for (int k = 1; k < PC_THREADS / 32; ++k) t = dd_merge(t, sm[k]);
  return t;
}
__device__ __forceinline__ double pc_block_sum(double v, double* sm) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) sm[wid] = v;
  __syncthreads();
  double t = sm[0];
  for (int k = 1; k < PC_THREADS / 32; ++k) t += sm[k];
  return t;
}
__device__ __forceinline__ int pc_find_group(const SegGroup* __restrict__ groups, int ng, int32_t job, int32_t est) {
  int lo = 0, hi = ng;               // groups are ordered by (job, estimate_id)
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    const bool less = groups[mid].job < job || (groups[mid].job == job && groups[mid].estimate_id < est);
    if (less) lo = mid + 1; else hi = mid;
  }
  return (lo < ng && groups[lo].job == job && groups[lo].estimate_id == est) ? lo : -1;
}

struct PcPieces {         // per piece of split_call: inputs, then what :169-175 compute
  int32_t n;
  const int32_t *job, *est;
  const uint32_t *start, *end;
  const double* fused_std;
  int32_t *group, *ncpg;           // ncpg == 0: no position falls on the piece, it has no row in pmd_call
  int64_t *row_lo, *row_hi;
  uint32_t *pstart, *pend;         // min(pos), max(pos) + 2 (:171)
  double *beta, *cov;
  DD* sum;                         // double-double sum of the piece's methylation values (first sweep)
};

// :169-175  the positions of a piece are those with start - 1 <= pos <= end (foverlaps of [pos, pos + 1]); per piece: mean
// of Nmeth / coverage (two-pass, as R's mean), coverage sum, position count, first and last position.  Pieces differ in
// size by five orders of magnitude (a PMD of a few hundred positions, the background between two PMDs with a million),
// so a piece of more than PC_MIN_SHARE rows is swept by up to PC_SUB blocks, each over a contiguous share of its rows:
// block k < n takes share 0 of piece k, the blocks behind them take the (piece, share) pairs k_pc_piece_rows listed.  The
// partial sums are combined per piece in share order (fixed order: the result does not depend on the launch).
constexpr int PC_SUB = 256;
constexpr int64_t PC_MIN_SHARE = 1024;            // rows below which a piece is not cut further
struct PcPart { DD sum; double csum, m, first, last; };
struct PcShares {
  int32_t *nshares, *base;   // per piece: number of shares, list position of its share 1
  int32_t* count;            // listed extra shares
  int32_t *item, *share;
  int32_t cap;
};
// where the partial result of share y of item i (a piece, or a live piece) lives: the first shares in item order, the
// listed ones behind them in list order (n = number of items)
__device__ __forceinline__ size_t pc_slot(int n, int i, int y, int base) { return y == 0 ? (size_t)i : (size_t)n + base + y - 1; }
__device__ __forceinline__ int pc_nshares(int64_t n) {
  const int64_t s = (n + PC_MIN_SHARE - 1) / PC_MIN_SHARE;
  return s < 1 ? 1 : (s > PC_SUB ? PC_SUB : (int)s);
}
__device__ __forceinline__ void pc_share(int64_t lo, int64_t hi, int y, int64_t& a, int64_t& b) {
  const int64_t n = hi - lo;
  const int64_t len = (n + pc_nshares(n) - 1) / pc_nshares(n);
  a = lo + y * len;
  a = a < hi ? a : hi;
  b = a + len < hi ? a + len : hi;
}
// one thread per piece: its group and rows, the number of shares; shares beyond the first go on the list
__global__ void k_pc_piece_rows(const SegGroup* __restrict__ groups, int ng, const void* __restrict__ pos, int pos_f64, PcPieces P,
                                PcShares L) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= P.n) return;
  const int g = pc_find_group(groups, ng, P.job[k], P.est[k]);
  int64_t lo = 0, hi = 0;
  if (g >= 0) segment_rows(groups[g], pos, pos_f64, P.start[k], P.end[k], lo, hi);
  P.group[k] = g;
  P.row_lo[k] = lo;
  P.row_hi[k] = hi;
  const int ns = pc_nshares(hi - lo);
  L.nshares[k] = ns;
  int base = 0;
  if (ns > 1) {
    base = atomicAdd(L.count, ns - 1);
    for (int y = 1; y < ns; ++y)
      if (base + y - 1 < L.cap) { L.item[base + y - 1] = k; L.share[base + y - 1] = y; }
  }
  L.base[k] = base;
}
// which (piece, share) a WARP works on (most pieces hold a handful of positions: a block each would be mostly launch
// overhead); false: nothing
constexpr int PC_WARPS = PC_THREADS / 32;
__device__ __forceinline__ bool pc_warp_item(int n, const PcShares& L, int& k, int& y) {
  const int w = blockIdx.x * PC_WARPS + (threadIdx.x >> 5);
  if (w < n) { k = w; y = 0; return true; }
  const int idx = w - n;
  const int cnt = *L.count < L.cap ? *L.count : L.cap;
  if (idx >= cnt) return false;
  k = L.item[idx];
  y = L.share[idx];
  return true;
}
template <class Rows>
__global__ void __launch_bounds__(PC_THREADS)
k_pc_piece_stats(const SegGroup* __restrict__ groups, const void* __restrict__ pos, int pos_f64, Rows rows, PcPieces P, PcShares L,
                 PcPart* __restrict__ part) {
  int k, y;
  if (!pc_warp_item(P.n, L, k, y)) return;
  const int lane = threadIdx.x & 31;
  const int g = P.group[k];
  if (g < 0) {
    if (lane == 0) part[pc_slot(P.n, k, y, L.base[k])] = PcPart{DD{0.0, 0.0}, 0.0, 0.0, 4294967296.0, -1.0};
    return;
  }
  const SegGroup G = groups[g];
  int64_t a, b;
  pc_share(P.row_lo[k], P.row_hi[k], y, a, b);
  DD sum{0.0, 0.0};
  double csum = 0.0, m = 0.0, first = 4294967296.0, last = -1.0;
  for (int64_t i = a + lane; i < b; i += 32) {
    if (!rows.valid(i)) continue;
    dd_add(sum, rows.x(i));
    csum += rows.cov(i);
    m += 1.0;
    const double p = (double)seg_start_at(G, pos, pos_f64, i);
    first = fmin(first, p);
    last = fmax(last, p);
  }
  sum = dd_warp_sum(sum);
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    csum += __shfl_xor_sync(0xffffffffu, csum, d);
    m += __shfl_xor_sync(0xffffffffu, m, d);
    first = fmin(first, __shfl_xor_sync(0xffffffffu, first, d));     // the positions are sorted: first / last valid one
    last = fmax(last, __shfl_xor_sync(0xffffffffu, last, d));
  }
  if (lane == 0) part[pc_slot(P.n, k, y, L.base[k])] = PcPart{sum, csum, m, first, last};
}
// one thread per piece: the shares in order -> first-pass mean (left in P.beta for the second pass)
__global__ void k_pc_piece_mean0(PcPieces P, PcShares L, const PcPart* __restrict__ part) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= P.n) return;
  DD sum{0.0, 0.0};
  double csum = 0.0, m = 0.0, first = 4294967296.0, last = -1.0;
  const int ns = L.nshares[k];
  for (int y = 0; y < ns; ++y) {
    const PcPart q = part[pc_slot(P.n, k, y, L.base[k])];
    sum = dd_merge(sum, q.sum);
    csum += q.csum;
    m += q.m;
    first = fmin(first, q.first);
    last = fmax(last, q.last);
  }
  if (P.group[k] < 0) m = 0.0;
  P.ncpg[k] = (int32_t)m;
  P.pstart[k] = m > 0 ? (uint32_t)first : 0u;
  P.pend[k] = m > 0 ? (uint32_t)last + 2u : 0u;
  P.beta[k] = (sum.hi + sum.lo) / m;
  P.cov[k] = csum;
  P.sum[k] = sum;
}
// second pass of R's mean: sum of the deviations from the first-pass mean, per share
template <class Rows>
__global__ void __launch_bounds__(PC_THREADS)
k_pc_piece_dev(Rows rows, PcPieces P, PcShares L, DD* __restrict__ part_t) {
  int k, y;
  if (!pc_warp_item(P.n, L, k, y)) return;
  const int lane = threadIdx.x & 31;
  const double mean = P.beta[k];
  int64_t a, b;
  pc_share(P.row_lo[k], P.row_hi[k], y, a, b);
  DD t{0.0, 0.0};
  if (P.group[k] >= 0 && P.ncpg[k] > 0 && is_finite(mean) && a < b) {     // (uniform over the warp)
    for (int64_t i = a + lane; i < b; i += 32)
      if (rows.valid(i)) dd_add(t, rows.x(i) - mean);
    t = dd_warp_sum(t);
  }
  if (lane == 0) part_t[pc_slot(P.n, k, y, L.base[k])] = t;
}
__global__ void k_pc_piece_mean(PcPieces P, PcShares L, const DD* __restrict__ part_t) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= P.n) return;
  const double m = (double)P.ncpg[k];
  double mean = P.beta[k];
  if (m > 0 && is_finite(mean)) {
    DD t{0.0, 0.0};
    const int ns = L.nshares[k];
    for (int y = 0; y < ns; ++y) t = dd_merge(t, part_t[pc_slot(P.n, k, y, L.base[k])]);
    mean += (t.hi + t.lo) / m;
  }
  P.beta[k] = mean;
}

__global__ void k_pc_live(PcPieces P, int32_t* __restrict__ flag) {
  const int32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < P.n) flag[k] = P.ncpg[k] > 0 ? 1 : 0;
}
// :177-180  is.admissible and the PMD call of a pmd_call row (1 "PMD", 0 "other")
__device__ __forceinline__ int8_t pc_category(const PcPieces& P, const PmdCallParams& p, int32_t k) {
  const double w = (double)P.pend[k] - (double)P.pstart[k] + 1, m = (double)P.ncpg[k];
  const bool adm = w > p.pmd_valley_min_width && w / (m - 1) < p.max_distance;
  return (adm && P.fused_std[k] >= p.pmd_std_threshold && P.beta[k] < p.pmd_max_beta) ? 1 : 0;
}
// :182-190  a merged region starts where the condition starts, the call changes or a large gap precedes (merge.pmds);
// without merging every row is its own region
__global__ void k_pc_heads(PcPieces P, PmdCallParams p, int32_t n_live, const int32_t* __restrict__ live,
                           int8_t* __restrict__ cat, int32_t* __restrict__ flag) {
  const int32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_live) return;
  const int32_t k = live[r];
  const int8_t c = pc_category(P, p, k);
  cat[r] = c;
  int32_t head = 1;
  if (r > 0 && p.merge_pmds) {
    const int32_t q = live[r - 1];
    if (P.job[q] == P.job[k] && P.est[q] == P.est[k])
      head = (pc_category(P, p, q) != c || (double)P.pstart[k] - (double)P.pend[q] > p.pmd_valley_min_width) ? 1 : 0;
  }
  flag[r] = head;
}
struct PcOut {
  int32_t *job, *est, *segment_id, *ncpg;
  int8_t* category;
  uint32_t *start, *end;
  double *width, *fused_std, *beta, *std_;
};
// :193-198  statistics of a merged region from per-piece partial sums, so that the work is spread over the pieces (a
// genome's "other" regions span millions of positions; one block per region took 13 ms on a 28 M-CpG genome).
// Region v = live pieces [heads[v], heads[v + 1]).  (a) one thread per region: mean0 = sum of the pieces' sums / count.
struct PcRegion {
  int32_t n;
  const int32_t* heads;
  double *mean0, *count;
  DD *t, *q;                       // per live piece: sum(x - mean0), sum((x - mean0)^2) with its region's mean0
};
// One WARP per region (an "other" region between two PMDs holds hundreds of pieces): lanes take the live pieces in a stride,
// their partial sums are merged in a fixed order (double-double: the order shows in the last bits of the low word only).
__global__ void k_pc_region_mean0(PcPieces P, const int32_t* __restrict__ live, int32_t n_live, PcRegion R) {
  const int32_t v = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (v >= R.n) return;
  const int32_t r0 = R.heads[v], r1 = v + 1 < R.n ? R.heads[v + 1] : n_live;
  DD s{0.0, 0.0};
  double m = 0.0;
  for (int32_t r = r0 + lane; r < r1; r += 32) { s = dd_merge(s, P.sum[live[r]]); m += (double)P.ncpg[live[r]]; }
  s = dd_warp_sum(s);
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) m += __shfl_xor_sync(0xffffffffu, m, d);       // counts: exact in any order
  if (lane == 0) {
    R.mean0[v] = (s.hi + s.lo) / m;
    R.count[v] = m;
  }
}
// (b) per live piece, in shares like the first sweep (block r < n_live: share 0 of live piece r; the listed shares behind
// them, by way of the piece -> live index map): the two central sums about its region's mean0; a position that two
// touching pieces share enters both, as in the reference's concatenated betavals
__global__ void k_pc_live_index(int32_t n_live, const int32_t* __restrict__ live, int32_t* __restrict__ index_of) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < n_live) index_of[live[r]] = r;
}
template <class Rows>
__global__ void __launch_bounds__(PC_THREADS)
k_pc_piece_moments(PcPieces P, PcShares L, int32_t n_live, const int32_t* __restrict__ live, const int32_t* __restrict__ index_of,
                   const int32_t* __restrict__ region_of, PcRegion R, Rows rows, DD* __restrict__ part_t, DD* __restrict__ part_q) {
  const int w = blockIdx.x * PC_WARPS + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  int r, k, y;
  if (w < n_live) { r = w; k = live[r]; y = 0; }
  else {
    const int idx = w - n_live;
    const int cnt = *L.count < L.cap ? *L.count : L.cap;
    if (idx >= cnt) return;
    k = L.item[idx];
    y = L.share[idx];
    r = index_of[k];
    if (r < 0) return;                     // the piece holds no position: it is not live
  }
  const double mean0 = R.mean0[region_of[r]];
  int64_t a, b;
  pc_share(P.row_lo[k], P.row_hi[k], y, a, b);
  DD t{0.0, 0.0}, q{0.0, 0.0};
  if (is_finite(mean0) && a < b) {
    for (int64_t i = a + lane; i < b; i += 32)
      if (rows.valid(i)) { const double d = rows.x(i) - mean0; dd_add(t, d); dd_add(q, d * d); }
    t = dd_warp_sum(t);
    q = dd_warp_sum(q);
  }
  if (lane == 0) { const size_t at = pc_slot(n_live, r, y, L.base[k]); part_t[at] = t; part_q[at] = q; }
}
__global__ void k_pc_piece_moments_sum(PcPieces P, PcShares L, int32_t n_live, const int32_t* __restrict__ live,
                                       const DD* __restrict__ part_t, const DD* __restrict__ part_q, PcRegion R) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_live) return;
  DD t{0.0, 0.0}, q{0.0, 0.0};
  const int k = live[r], ns = L.nshares[k];
  for (int y = 0; y < ns; ++y) { const size_t at = pc_slot(n_live, r, y, L.base[k]); t = dd_merge(t, part_t[at]); q = dd_merge(q, part_q[at]); }
  R.t[r] = t;
  R.q[r] = q;
}
// (c) one warp per region: mean = mean0 + T / M (R's second pass), sum of squares about it = Q - (T / M) T; fused_std = the
// mean of the pieces' values, two passes like R's mean() with double-double sums in the place of its long double
__global__ void k_pc_merge(PcPieces P, const int32_t* __restrict__ live, int32_t n_live, const int8_t* __restrict__ cat,
                           PcRegion R, PcOut o) {
  const int32_t v = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (v >= R.n) return;
  const int32_t r0 = R.heads[v], r1 = v + 1 < R.n ? R.heads[v + 1] : n_live;
  DD T{0.0, 0.0}, Q{0.0, 0.0}, F{0.0, 0.0};
  uint32_t lo = 0xffffffffu, hi = 0u;
  long long ncpg = 0;
  for (int32_t r = r0 + lane; r < r1; r += 32) {
    const int32_t k = live[r];
    T = dd_merge(T, R.t[r]);
    Q = dd_merge(Q, R.q[r]);
    dd_add(F, P.fused_std[k]);
    lo = P.pstart[k] < lo ? P.pstart[k] : lo;
    hi = P.pend[k] > hi ? P.pend[k] : hi;
    ncpg += P.ncpg[k];
  }
  T = dd_warp_sum(T);
  Q = dd_warp_sum(Q);
  F = dd_warp_sum(F);
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    const uint32_t ol = __shfl_xor_sync(0xffffffffu, lo, d), oh = __shfl_xor_sync(0xffffffffu, hi, d);
    lo = ol < lo ? ol : lo;
    hi = oh > hi ? oh : hi;
    ncpg += __shfl_xor_sync(0xffffffffu, ncpg, d);
  }
  const double np_ = (double)(r1 - r0);
  double fmean = (F.hi + F.lo) / np_;
  if (is_finite(fmean)) {
    DD D{0.0, 0.0};
    for (int32_t r = r0 + lane; r < r1; r += 32) dd_add(D, P.fused_std[live[r]] - fmean);
    D = dd_warp_sum(D);
    fmean += (D.hi + D.lo) / np_;
  }
  if (lane) return;
  const double m = R.count[v], mean0 = R.mean0[v];
  const double delta = (T.hi + T.lo) / m;
  const double mean = is_finite(mean0) ? mean0 + delta : mean0;
  DD ss = Q;
  dd_add(ss, -delta * (T.hi + T.lo));
  const int32_t k0 = live[r0];
  o.job[v] = P.job[k0];
  o.est[v] = P.est[k0];
  o.category[v] = cat[r0];
  o.start[v] = lo;
  o.end[v] = hi;
  o.ncpg[v] = (int32_t)ncpg;
  o.fused_std[v] = fmean;
  o.width[v] = (double)hi - (double)lo + 1;
  o.beta[v] = mean;
  o.std_[v] = m >= 2 ? sqrt(fmax(ss.hi + ss.lo, 0.0) / (m - 1)) : 0.0;      // sd of one value is NA -> 0 (:197)
}
__global__ void k_pc_region_of(int32_t n, const int32_t* __restrict__ flag, const int32_t* __restrict__ pos, int32_t* __restrict__ region_of) {
  const int32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < n) region_of[r] = pos[r] + flag[r] - 1;
}
// segment_id of :186 / :189 is only a grouping key in the reference; the regions are numbered 1.. per condition here
__global__ void k_pc_group_heads(int32_t n, PcOut o, int32_t* __restrict__ flag) {
  const int32_t v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v < n) flag[v] = (v == 0 || o.job[v] != o.job[v - 1] || o.est[v] != o.est[v - 1]) ? 1 : 0;
}

}  // namespace

void PmdCallDev::alloc(cudaStream_t st, int64_t count) {
  n = count;
  const size_t c = (size_t)std::max<int64_t>(1, count);
  job.alloc(st, c); est.alloc(st, c); segment_id.alloc(st, c); num_cpgs.alloc(st, c); category.alloc(st, c);
  start.alloc(st, c); end.alloc(st, c); width.alloc(st, c); fused_std.alloc(st, c); beta.alloc(st, c); std_.alloc(st, c);
}

void call_pmds_device(Engine& eng, const std::vector<SegGroup>& groups, const void* d_pos, int pos_f64,
                      const double* d_betahat, const double* d_pw, const PiecesDev& pieces, const PmdCallParams& p,
                      PmdCallDev& out) {
  cudaStream_t st = eng.stream;
  out.alloc(st, 0);
  const int32_t np = (int32_t)pieces.n;
  const int ng = (int)groups.size();
  if (np == 0 || ng == 0) return;
  const RowsFromEstimates rows{d_betahat, d_pw};
  DevBuf<SegGroup> d_groups(st, (size_t)ng);
  d_groups.upload(groups);
  DevBuf<int32_t> p_group(st, (size_t)np), p_ncpg(st, (size_t)np);
  DevBuf<int64_t> p_lo(st, (size_t)np), p_hi(st, (size_t)np);
  DevBuf<uint32_t> p_start(st, (size_t)np), p_end(st, (size_t)np);
  DevBuf<double> p_beta(st, (size_t)np), p_cov(st, (size_t)np);
  DevBuf<DD> p_sum(st, (size_t)np);
  const PcPieces P{np, pieces.job.p, pieces.est.p, pieces.start.p, pieces.end.p, pieces.fused_std.p, p_group.p, p_ncpg.p,
                   p_lo.p, p_hi.p, p_start.p, p_end.p, p_beta.p, p_cov.p, p_sum.p};
  // shares of the long pieces: at most (rows of the estimates table) / PC_MIN_SHARE of them beyond every piece's first
  int64_t total_rows = 0;
  for (const SegGroup& g : groups) total_rows += g.nrows;
  const int32_t share_cap = (int32_t)(total_rows / PC_MIN_SHARE + 64);
  DevBuf<int32_t> sh_n(st, (size_t)np), sh_base(st, (size_t)np), sh_count(st, 1), sh_item(st, (size_t)share_cap), sh_share(st, (size_t)share_cap);
  sh_count.zero();
  const PcShares L{sh_n.p, sh_base.p, sh_count.p, sh_item.p, sh_share.p, share_cap};
  ML_LAUNCH(eng, "k_pc_piece_rows", k_pc_piece_rows<<<cdiv(np, 128), 128, 0, st>>>(d_groups.p, ng, d_pos, pos_f64, P, L));
  {
    DevBuf<PcPart> part(st, (size_t)np + share_cap);
    DevBuf<DD> part_t(st, (size_t)np + share_cap);
    ML_LAUNCH(eng, "k_pc_piece_stats", k_pc_piece_stats<RowsFromEstimates><<<cdiv(np + share_cap, PC_WARPS), PC_THREADS, 0, st>>>(
        d_groups.p, d_pos, pos_f64, rows, P, L, part.p));
    ML_LAUNCH(eng, "k_pc_piece_mean0", k_pc_piece_mean0<<<cdiv(np, 128), 128, 0, st>>>(P, L, part.p));
    ML_LAUNCH(eng, "k_pc_piece_dev", k_pc_piece_dev<RowsFromEstimates><<<cdiv(np + share_cap, PC_WARPS), PC_THREADS, 0, st>>>(rows, P, L, part_t.p));
    ML_LAUNCH(eng, "k_pc_piece_mean", k_pc_piece_mean<<<cdiv(np, 128), 128, 0, st>>>(P, L, part_t.p));
  }
  DevBuf<int32_t> flag(st, (size_t)np), live(st, (size_t)np), heads(st, (size_t)np);
  ML_LAUNCH(eng, "k_pc_live", k_pc_live<<<cdiv(np, 256), 256, 0, st>>>(P, flag.p));
  Scan scan;
  const int32_t n_live = scan.run(eng, flag.p, np, nullptr, live.p);
  if (n_live == 0) { stream_sync(st); return; }
  DevBuf<int8_t> cat(st, (size_t)n_live);
  ML_LAUNCH(eng, "k_pc_heads", k_pc_heads<<<cdiv(n_live, 256), 256, 0, st>>>(P, p, n_live, live.p, cat.p, flag.p));
  DevBuf<int32_t> hpos(st, (size_t)n_live);
  const int32_t n_out = scan.run(eng, flag.p, n_live, hpos.p, heads.p);
  out.alloc(st, n_out);
  const PcOut o{out.job.p, out.est.p, out.segment_id.p, out.num_cpgs.p, out.category.p, out.start.p, out.end.p,
                out.width.p, out.fused_std.p, out.beta.p, out.std_.p};
  DevBuf<double> r_mean0(st, (size_t)n_out), r_count(st, (size_t)n_out);
  DevBuf<DD> r_t(st, (size_t)n_live), r_q(st, (size_t)n_live);
  DevBuf<int32_t> region_of(st, (size_t)n_live);
  const PcRegion Rg{n_out, heads.p, r_mean0.p, r_count.p, r_t.p, r_q.p};
  ML_LAUNCH(eng, "k_pc_region_mean0", k_pc_region_mean0<<<cdiv(32LL * n_out, 128), 128, 0, st>>>(P, live.p, n_live, Rg));
  ML_LAUNCH(eng, "k_pc_region_of", k_pc_region_of<<<cdiv(n_live, 256), 256, 0, st>>>(n_live, flag.p, hpos.p, region_of.p));
  {
    DevBuf<DD> part_t(st, (size_t)n_live + share_cap), part_q(st, (size_t)n_live + share_cap);
    DevBuf<int32_t> index_of(st, (size_t)np);
    ML_CUDA(cudaMemsetAsync(index_of.p, 0xff, sizeof(int32_t) * (size_t)np, st));
    ML_LAUNCH(eng, "k_pc_live_index", k_pc_live_index<<<cdiv(n_live, 256), 256, 0, st>>>(n_live, live.p, index_of.p));
    ML_LAUNCH(eng, "k_pc_piece_moments", k_pc_piece_moments<RowsFromEstimates><<<cdiv(n_live + share_cap, PC_WARPS), PC_THREADS, 0, st>>>(
        P, L, n_live, live.p, index_of.p, region_of.p, Rg, rows, part_t.p, part_q.p));
    ML_LAUNCH(eng, "k_pc_piece_moments_sum", k_pc_piece_moments_sum<<<cdiv(n_live, 128), 128, 0, st>>>(P, L, n_live, live.p, part_t.p,
                                                                                                        part_q.p, Rg));
  }
  ML_LAUNCH(eng, "k_pc_merge", k_pc_merge<<<cdiv(32LL * n_out, 128), 128, 0, st>>>(P, live.p, n_live, cat.p, Rg, o));
  DevBuf<int32_t> gflag(st, (size_t)n_out), gpos(st, (size_t)n_out), gheads(st, (size_t)n_out);
  ML_LAUNCH(eng, "k_pc_group_heads", k_pc_group_heads<<<cdiv(n_out, 256), 256, 0, st>>>(n_out, o, gflag.p));
  scan.run(eng, gflag.p, n_out, gpos.p, gheads.p);
  ML_LAUNCH(eng, "k_ps_seq_ids", k_ps_seq_ids<<<cdiv(n_out, 256), 256, 0, st>>>(n_out, gflag.p, gpos.p, gheads.p, out.segment_id.p));
  stream_sync(st);
}

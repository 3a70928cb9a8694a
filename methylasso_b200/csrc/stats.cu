// stats.cu — Wilcoxon rank-sum test per group and Benjamini-Hochberg adjustment (R/call.R:315-317, :360).
//
// The reference's R layer calls stats::wilcox.test(betavals, betavals.ref, conf.int = F) once per candidate region
// and then p.adjust(pval, 'fdr') over the genome.  Here:
//   k_wilcox_rank    one thread per value of the pooled samples of all groups: counts how many values of its group
//                    are smaller / equal (mid-ranks without a sort: the work is spread over all lanes whatever the
//                    group sizes) and adds rank, tie term and sample sizes into the group's accumulators.  Every
//                    addend is a multiple of 1/2 far below 2^53, so the sums are exact and independent of the order
//                    of the atomics;
//   k_wilcox_pvalue  one thread per group: exact two-sided p-value from the rank-sum distribution when both samples
//                    have fewer than 50 values and there are no ties, normal approximation with continuity and tie
//                    correction otherwise (wilcox.test.default of R's stats package).  The exact distributions
//                    (cwilcox of nmath/wilcox.c, a memoised 3-index recurrence there) are tabulated once per device
//                    by k_wilcox_tables: the power series of the Gaussian binomial
//                    prod_{i=1..a} (1 - q^(b+i)) / (1 - q^i)  truncated at degree a*b/2 (pwilcox only ever sums the
//                    lower half), built by one warp per pair of sizes in shared memory (<= 1201 doubles).
// The FDR adjustment is a descending sort (cub radix sort: library plumbing), a scaled running minimum and a scatter.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <cmath>
#include <map>
#include <mutex>
#include <vector>

#include "common.h"
#include "hd.h"
#include "stats.cuh"

namespace ml {
namespace {

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  return v;
}
__device__ __forceinline__ int warp_sum_i(int v) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  return v;
}

struct WilcoxGroup {
  double rank_sum, tie_term;      // sum of the mid-ranks of x; sum over tie groups of (t^3 - t)
  int32_t mx, my, ties, pad;      // sample sizes without NaN, any ties
};

constexpr int kExactLimit = 50;                                        // exact <- (n.x < 50) && (n.y < 50)
constexpr int kSeriesLen = ((kExactLimit - 1) * (kExactLimit - 1)) / 2 + 1;   // degrees 0 .. floor(49 * 49 / 2)
constexpr int kPvalWarps = 4;

// pooled element t of all groups -> its group: the largest g with x_ptr[g] + y_ptr[g] <= t
__device__ __forceinline__ int64_t group_of(int64_t t, int64_t n_groups, const int64_t* __restrict__ x_ptr,
                                            const int64_t* __restrict__ y_ptr) {
  int64_t lo = 0, hi = n_groups - 1;
  while (lo < hi) {
    const int64_t mid = (lo + hi + 1) >> 1;
    if (x_ptr[mid] + y_ptr[mid] <= t) lo = mid; else hi = mid - 1;
  }
  return lo;
}

__global__ void __launch_bounds__(256)
k_wilcox_rank(int64_t n_groups, int64_t n_pooled, const int64_t* __restrict__ x_ptr, const double* __restrict__ x,
              const int64_t* __restrict__ y_ptr, const double* __restrict__ y, WilcoxGroup* __restrict__ out) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = t < n_pooled;
  int64_t g = -1;
  double rank = 0.0, tie = 0.0;
  int is_x = 0, is_y = 0, tied = 0;
  if (live) {
    g = group_of(t, n_groups, x_ptr, y_ptr);
    const int64_t x0 = x_ptr[g], nx = x_ptr[g + 1] - x0, y0 = y_ptr[g], ny = y_ptr[g + 1] - y0;
    const int64_t e = t - (x0 + y0);
    const double v = e < nx ? x[x0 + e] : y[y0 + (e - nx)];
    if (v == v) {                                     // x <- x[!is.na(x)]; a NaN compares false below as well
      int64_t less = 0, eq = 0;
      for (int64_t j = 0; j < nx; ++j) { const double u = x[x0 + j]; less += u < v; eq += u == v; }
      for (int64_t j = 0; j < ny; ++j) { const double u = y[y0 + j]; less += u < v; eq += u == v; }
      if (e < nx) { rank = (double)less + ((double)eq + 1.0) / 2.0; is_x = 1; } else is_y = 1;
      tie = (double)eq * (double)eq - 1.0;            // sum over elements of (t^2 - 1) == sum over tie groups of (t^3 - t)
      tied = eq > 1;
    }
  }
  // one set of atomics per warp when the whole warp sits in one group, per lane otherwise
  const int64_t g0 = __shfl_sync(0xffffffffu, g, 0);
  if (__all_sync(0xffffffffu, g == g0)) {
    rank = warp_sum(rank);
    tie = warp_sum(tie);
    is_x = warp_sum_i(is_x);
    is_y = warp_sum_i(is_y);
    tied = __any_sync(0xffffffffu, tied);
    if ((threadIdx.x & 31) != 0) g = -1;
  }
  if (g >= 0 && (is_x | is_y)) {
    WilcoxGroup* G = out + g;
    if (rank != 0.0) atomicAdd(&G->rank_sum, rank);
    if (tie != 0.0) atomicAdd(&G->tie_term, tie);
    if (is_x) atomicAdd(&G->mx, is_x);
    if (is_y) atomicAdd(&G->my, is_y);
    if (tied) atomicOr(&G->ties, 1);
  }
}

__device__ double choose_d(int n, int k) {
  double c = 1.0;
  if (k > n - k) k = n - k;
  for (int i = 1; i <= k; ++i) c = c * (double)(n - k + i) / (double)i;
  return floor(c + 0.5);
}

// Exact null distribution of the rank sum for sample sizes a <= b < 50 (cwilcox of nmath/wilcox.c, a memoised 3-index
// recurrence there): the coefficients of the Gaussian binomial  prod_{i=1..a} (1 - q^(b+i)) / (1 - q^i)  as a power
// series truncated at degree T = a*b/2 (pwilcox only ever sums the lower half), built by one warp in shared memory.
__device__ __forceinline__ void wilcox_series(double* P, int a, int b, int T, int lane) {
  for (int k = lane; k <= T; k += 32) P[k] = k == 0 ? 1.0 : 0.0;
  __syncwarp();
  for (int i = 1; i <= a; ++i) {
    // times (1 - q^s), s = b + i: P[k] -= P[k - s] for k = T .. s, old values on the right-hand side
    const int s = b + i;
    for (int base = (T / 32) * 32; base + 31 >= s; base -= 32) {
      const int k = base + lane;
      const bool on = k >= s && k <= T;
      const double v = on ? P[k] - P[k - s] : 0.0;
      __syncwarp();
      if (on) P[k] = v;
      __syncwarp();
    }
    // divided by (1 - q^i): P[k] += P[k - i] ascending; the residue classes mod i are independent chains
    for (int r = lane; r < i && r <= T; r += 32) {
      double acc = P[r];
      for (int k = r + i; k <= T; k += i) { acc += P[k]; P[k] = acc; }
    }
    __syncwarp();
  }
}

// The distributions depend on (a, b) alone and a genome holds 10^5-10^6 regions: they are tabulated once per device,
// as the running sums pwilcox forms (p += cwilcox(i, m, n) / choose(m + n, n), i = 0, 1, ... in that order), for all
// 1225 pairs a <= b < 50 (3 MB).  k_wilcox_pvalue then reads one or two entries per region.
__device__ int32_t g_wil_off[kExactLimit * kExactLimit];
__global__ void __launch_bounds__(kPvalWarps * 32)
k_wilcox_tables(double* __restrict__ table) {
  __shared__ double series[kPvalWarps][kSeriesLen];
  const int lane = threadIdx.x & 31;
  double* P = series[threadIdx.x >> 5];
  const int pair = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;          // a = pair / 50, b = pair % 50
  const int a = pair / kExactLimit, b = pair % kExactLimit;
  if (pair >= kExactLimit * kExactLimit || a < 1 || b < a) return;
  const int T = (a * b) / 2;
  wilcox_series(P, a, b, T, lane);
  if (lane == 0) {
    const double c = choose_d(a + b, b);
    double* out = table + g_wil_off[pair];
    double p = 0.0;
    for (int k = 0; k <= T; ++k) { p += P[k] / c; out[k] = p; }
  }
}

// pwilcox(q, m, n, lower_tail) of nmath/wilcox.c on the tabulated running sums of the pair (min, max)
__device__ __forceinline__ double pwilcox_tab(double q, int m, int n, bool lower_tail, const double* __restrict__ table) {
  q = floor(q + 1e-7);
  const double mn = (double)m * n;
  if (q < 0.0) return lower_tail ? 0.0 : 1.0;
  if (q >= mn) return lower_tail ? 1.0 : 0.0;
  const double* cdf = table + g_wil_off[min(m, n) * kExactLimit + max(m, n)];
  if (q <= mn / 2) {
    const double p = cdf[(int)q];
    return lower_tail ? p : 1.0 - p;
  }
  const int last = (int)(mn - q) - 1;                       // p = sum of the first mn - q terms
  const double p = last >= 0 ? cdf[last] : 0.0;
  return lower_tail ? 1.0 - p : p;
}

// wilcox.test.default of R's stats package, two-sided, mu = 0, exact = NULL, correct = TRUE; one thread per group
__global__ void __launch_bounds__(256)
k_wilcox_pvalue(int64_t n_groups, const WilcoxGroup* __restrict__ grp, const double* __restrict__ table,
                double* __restrict__ pval, double* __restrict__ stat) {
  const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n_groups) return;
  const WilcoxGroup G = grp[g];
  const bool empty = G.mx < 1 || G.my < 1;
  const double dx = G.mx, dy = G.my;
  const double W = G.rank_sum - dx * (dx + 1.0) / 2.0;
  if (stat) stat[g] = empty ? NAN : W;
  if (empty) { pval[g] = NAN; return; }
  if (G.mx < kExactLimit && G.my < kExactLimit && !G.ties) {
    const double p = W > dx * dy / 2 ? pwilcox_tab(W - 1, G.mx, G.my, false, table) : pwilcox_tab(W, G.mx, G.my, true, table);
    pval[g] = fmin(2 * p, 1.0);
  } else {
    double z = W - dx * dy / 2;
    const double sigma = sqrt((dx * dy / 12) * ((dx + dy + 1) - G.tie_term / ((dx + dy) * (dx + dy - 1))));
    const double corr = z > 0 ? 0.5 : (z < 0 ? -0.5 : 0.0);
    z = (z - corr) / sigma;
    const double pl = 0.5 * erfc(-z / sqrt(2.0)), pu = 0.5 * erfc(z / sqrt(2.0));
    pval[g] = 2 * fmin(pl, pu);
  }
}

// ---- Benjamini-Hochberg -------------------------------------------------------------------------------------
__global__ void k_bh_keys(int64_t n, const double* __restrict__ p, double* __restrict__ key, int32_t* __restrict__ idx,
                          double* __restrict__ q, unsigned long long* __restrict__ count) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  bool ok = false;
  if (i < n) {
    const double v = p[i];
    ok = v == v;
    // NaN entries sort to the very end (key -inf) and are left out of the ranking (R leaves NA out of n)
    key[i] = ok ? v : -HUGE_VAL;
    idx[i] = (int32_t)i;
    q[i] = NAN;
  }
  const unsigned b = __ballot_sync(0xffffffffu, ok);
  if ((threadIdx.x & 31) == 0 && b) atomicAdd(count, (unsigned long long)__popc(b));
}
__global__ void k_bh_scale(int64_t lp, const double* __restrict__ key_sorted, double* __restrict__ val) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= lp) return;
  val[r] = (double)lp / (double)(lp - r) * key_sorted[r];   // n / i * p[o], i = lp .. 1
}
__global__ void k_bh_scatter(int64_t lp, const double* __restrict__ run, const int32_t* __restrict__ idx_sorted,
                             double* __restrict__ q) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= lp) return;
  q[idx_sorted[r]] = fmin(run[r], 1.0);
}
struct MinOp { __device__ __forceinline__ double operator()(double a, double b) const { return fmin(a, b); } };

}  // namespace

// per-device table of the exact distributions, built on first use and kept for the life of the process (3 MB)
static const double* wilcoxon_table(Engine& eng) {
  static std::mutex mu;
  static std::map<int, double*> tables;
  std::lock_guard<std::mutex> lock(mu);
  auto it = tables.find(eng.device);
  if (it != tables.end()) return it->second;
  std::vector<int32_t> off(kExactLimit * kExactLimit, 0);
  int64_t total = 0;
  for (int a = 1; a < kExactLimit; ++a)
    for (int b = a; b < kExactLimit; ++b) { off[a * kExactLimit + b] = (int32_t)total; total += (a * b) / 2 + 1; }
  double* t = nullptr;
  ML_CUDA(cudaMalloc(&t, sizeof(double) * (size_t)total));
  ML_CUDA(cudaMemcpyToSymbol(g_wil_off, off.data(), sizeof(int32_t) * off.size()));
  const int warps = kExactLimit * kExactLimit;
  k_wilcox_tables<<<cdiv(warps, kPvalWarps), kPvalWarps * 32, 0, eng.stream>>>(t);
  ML_CHECK_LAUNCH();
  ML_CUDA(cudaStreamSynchronize(eng.stream));
  eng.launches++;
  tables[eng.device] = t;
  return t;
}

void wilcoxon_groups(Engine& eng, int64_t n_groups, const int64_t* x_ptr, const double* x, const int64_t* y_ptr,
                     const double* y, double* pval, double* stat) {
  if (n_groups <= 0) return;
  cudaStream_t st = eng.stream;
  const int64_t x0 = x_ptr[0], y0 = y_ptr[0], nx = x_ptr[n_groups] - x0, ny = y_ptr[n_groups] - y0;
  if (nx < 0 || ny < 0) throw MlError(19, "wilcoxon: offsets must not decrease");
  // offsets relative to the first value, so that pooled element t of group g sits at x_ptr[g] + y_ptr[g] + e
  std::vector<int64_t> xp((size_t)n_groups + 1), yp((size_t)n_groups + 1);
  for (int64_t g = 0; g <= n_groups; ++g) {
    xp[g] = x_ptr[g] - x0;
    yp[g] = y_ptr[g] - y0;
    if (g && (xp[g] < xp[g - 1] || yp[g] < yp[g - 1])) throw MlError(19, "wilcoxon: offsets must not decrease");
  }
  DevBuf<int64_t> d_xp(st, xp.size()), d_yp(st, yp.size());
  DevBuf<double> d_x(st, (size_t)std::max<int64_t>(1, nx)), d_y(st, (size_t)std::max<int64_t>(1, ny));
  d_xp.upload(xp);
  d_yp.upload(yp);
  d_x.upload(x + x0, (size_t)nx);
  d_y.upload(y + y0, (size_t)ny);
  DevBuf<double> d_p(st, (size_t)n_groups), d_w(st, (size_t)n_groups);
  wilcoxon_groups_device(eng, n_groups, d_xp.p, d_x.p, d_yp.p, d_y.p, nx + ny, d_p.p, stat ? d_w.p : nullptr);
  d_p.download(pval, (size_t)n_groups);
  if (stat) d_w.download(stat, (size_t)n_groups);
  stream_sync(st);
}

// the same on device-resident samples: offsets relative to d_x / d_y, n_pooled = total number of values
void wilcoxon_groups_device(Engine& eng, int64_t n_groups, const int64_t* d_xptr, const double* d_x, const int64_t* d_yptr,
                            const double* d_y, int64_t n_pooled, double* d_pval, double* d_stat) {
  if (n_groups <= 0) return;
  cudaStream_t st = eng.stream;
  DevBuf<WilcoxGroup> grp(st, (size_t)n_groups);
  grp.zero();
  if (n_pooled > 0)
    ML_LAUNCH(eng, "k_wilcox_rank",
              k_wilcox_rank<<<cdiv(n_pooled, 256), 256, 0, st>>>(n_groups, n_pooled, d_xptr, d_x, d_yptr, d_y, grp.p));
  const double* table = wilcoxon_table(eng);
  ML_LAUNCH(eng, "k_wilcox_pvalue", k_wilcox_pvalue<<<cdiv(n_groups, 256), 256, 0, st>>>(n_groups, grp.p, table, d_pval, d_stat));
}

void p_adjust_fdr(Engine& eng, int64_t n, const double* p, double* q) {
  if (n <= 0) return;
  if (n >= 0x7fffffff) throw MlError(19, "p.adjust: too many p-values");
  cudaStream_t st = eng.stream;
  DevBuf<double> d_p(st, (size_t)n), key(st, (size_t)n), key_s(st, (size_t)n), val(st, (size_t)n), run(st, (size_t)n),
      d_q(st, (size_t)n);
  DevBuf<int32_t> idx(st, (size_t)n), idx_s(st, (size_t)n);
  DevBuf<unsigned long long> count(st, 1);
  d_p.upload(p, (size_t)n);
  count.zero();
  ML_LAUNCH(eng, "k_bh_keys", k_bh_keys<<<cdiv(n, 256), 256, 0, st>>>(n, d_p.p, key.p, idx.p, d_q.p, count.p));
  unsigned long long lp_host = 0;
  count.download(&lp_host, 1);
  size_t sort_bytes = 0, scan_bytes = 0;
  ML_CUDA(cub::DeviceRadixSort::SortPairsDescending(nullptr, sort_bytes, key.p, key_s.p, idx.p, idx_s.p, (int)n, 0, 64, st));
  ML_CUDA(cub::DeviceScan::InclusiveScan(nullptr, scan_bytes, val.p, run.p, MinOp(), (int)n, st));
  DevBuf<char> tmp(st, std::max(sort_bytes, scan_bytes) + 16);
  ML_CUDA(cub::DeviceRadixSort::SortPairsDescending(tmp.p, sort_bytes, key.p, key_s.p, idx.p, idx_s.p, (int)n, 0, 64, st));
  eng.launches++;
  stream_sync(st);                                   // lp (number of non-NaN p-values) is needed on the host
  const int64_t lp = (int64_t)lp_host;
  if (lp > 0) {
    ML_LAUNCH(eng, "k_bh_scale", k_bh_scale<<<cdiv(lp, 256), 256, 0, st>>>(lp, key_s.p, val.p));
    ML_CUDA(cub::DeviceScan::InclusiveScan(tmp.p, scan_bytes, val.p, run.p, MinOp(), (int)lp, st));
    eng.launches++;
    ML_LAUNCH(eng, "k_bh_scatter", k_bh_scatter<<<cdiv(lp, 256), 256, 0, st>>>(lp, run.p, idx_s.p, d_q.p));
  }
  d_q.download(q, (size_t)n);
  stream_sync(st);
}

}  // namespace ml

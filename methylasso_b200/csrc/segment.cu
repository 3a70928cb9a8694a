// segment.cu — segment_methylation_helper on sm_100a (K15).
//
// Replaces src/methylation_helpers.cpp:6-72: a single sweep over the estimates table that opens a
// new segment when |beta_i - beta_at_segment_start| > tol_val or the estimate changes, and emits
// per-segment weighted statistics.  The dependence on the SEGMENT-START beta makes the sweep
// sequential, but only over RUNS of equal beta (the FLSA output is exactly piecewise constant):
//   1. run heads        : beta_i != beta_{i-1} or group start            (parallel, rows)
//   2. compaction       : exclusive scan of the head flags               (3-phase scan)
//   3. tolerance merge  : one thread per group walks its runs            (sequential over runs)
//   4. compaction of segment heads, then one thread per segment adds its rows in row order — the
//      reference's own summation order, so betahat / pseudoweight / stdev are bit-identical.
// A group is one estimate of one job (one `estimate_id` block of one table).
#include <algorithm>
#include <cstring>

#include "segment.cuh"
#include "counts_div.cuh"
#include "detect.cuh"
#include "flsa.cuh"
#include "hd.h"
#include "stats.cuh"
#include "x87.cuh"

namespace ml {

namespace {

constexpr int STILE = 2048;
constexpr int SEG_LONG = 256;   // rows; longer segments are summed by a block (k_seg_emit_long)

// ---- generic exclusive scan of int32 flags (3 phases, no inter-block waiting) -----------------
__global__ void __launch_bounds__(256)
k_scan_tile_sums(const int32_t* __restrict__ flags, int64_t n, int32_t* __restrict__ tile_sum) {
  __shared__ int32_t sm[8];
  const int64_t base = (int64_t)blockIdx.x * STILE;
  int32_t c = 0;
  for (int i = threadIdx.x; i < STILE; i += 256)
    if (base + i < n) c += flags[base + i];
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) c += __shfl_down_sync(0xffffffffu, c, d);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    int32_t s = 0;
    for (int k = 0; k < 8; ++k) s += sm[k];
    tile_sum[blockIdx.x] = s;
  }
}
// single block: exclusive scan of the tile sums in place; total -> *total
__global__ void __launch_bounds__(256)
k_scan_block(int32_t* __restrict__ tile_sum, int ntiles, int32_t* __restrict__ total) {
  __shared__ int32_t sm[8];
  __shared__ int32_t carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int base = 0; base < ntiles; base += 256) {
    const int t = base + threadIdx.x;
    const int32_t c = t < ntiles ? tile_sum[t] : 0;
    int32_t inc = c;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int32_t o = __shfl_up_sync(0xffffffffu, inc, d);
      if (lane >= d) inc += o;
    }
    if (lane == 31) sm[wid] = inc;
    __syncthreads();
    int32_t before = carry_s;
    for (int k = 0; k < wid; ++k) before += sm[k];
    if (t < ntiles) tile_sum[t] = before + inc - c;
    __syncthreads();
    if (threadIdx.x == 255) carry_s = before + inc;
    __syncthreads();
  }
  if (threadIdx.x == 0) *total = carry_s;
}
// positions: pos[i] = exclusive prefix of flags; if compact != nullptr, compact[pos[i]] = i for heads
__global__ void __launch_bounds__(256)
k_scan_apply(const int32_t* __restrict__ flags, int64_t n, const int32_t* __restrict__ tile_base,
             int32_t* __restrict__ pos, int32_t* __restrict__ compact) {
  __shared__ int32_t sm[8];
  constexpr int V = STILE / 256;
  const int64_t base = (int64_t)blockIdx.x * STILE + (int64_t)threadIdx.x * V;
  int32_t f[V];
  int32_t mine = 0;
#pragma unroll
  for (int k = 0; k < V; ++k) {
    f[k] = (base + k < n) ? flags[base + k] : 0;
    mine += f[k];
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int32_t inc = mine;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int32_t o = __shfl_up_sync(0xffffffffu, inc, d);
    if (lane >= d) inc += o;
  }
  if (lane == 31) sm[wid] = inc;
  __syncthreads();
  int32_t run = tile_base[blockIdx.x] + inc - mine;
  for (int k = 0; k < wid; ++k) run += sm[k];
#pragma unroll
  for (int k = 0; k < V; ++k) {
    if (base + k < n) {
      if (pos) pos[base + k] = run;
      if (compact && f[k]) compact[run] = (int32_t)(base + k);
    }
    run += f[k];
  }
}

// The same scan in ONE launch for short inputs (the tables of the rule engine: 10^3 .. 10^5 rows, where three launches
// cost more than the work): a single block walks the flags 4096 at a time.
constexpr int SCAN_SMALL_THREADS = 1024, SCAN_SMALL_V = 4;
constexpr int64_t SCAN_SMALL_MAX = 1 << 15;
__global__ void __launch_bounds__(SCAN_SMALL_THREADS)
k_scan_small(const int32_t* __restrict__ flags, int n, int32_t* __restrict__ pos, int32_t* __restrict__ compact,
             int32_t* __restrict__ total) {
  __shared__ int32_t sm[SCAN_SMALL_THREADS / 32];
  __shared__ int32_t carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int base0 = 0; base0 < n; base0 += SCAN_SMALL_THREADS * SCAN_SMALL_V) {
    const int base = base0 + threadIdx.x * SCAN_SMALL_V;
    int32_t f[SCAN_SMALL_V];
    int32_t mine = 0;
#pragma unroll
    for (int k = 0; k < SCAN_SMALL_V; ++k) {
      f[k] = (base + k < n) ? flags[base + k] : 0;
      mine += f[k];
    }
    int32_t inc = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int32_t o = __shfl_up_sync(0xffffffffu, inc, d);
      if (lane >= d) inc += o;
    }
    if (lane == 31) sm[wid] = inc;
    __syncthreads();
    int32_t run = carry_s + inc - mine;
    for (int k = 0; k < wid; ++k) run += sm[k];
#pragma unroll
    for (int k = 0; k < SCAN_SMALL_V; ++k) {
      if (base + k < n) {
        if (pos) pos[base + k] = run;
        if (compact && f[k]) compact[run] = base + k;
      }
      run += f[k];
    }
    __syncthreads();
    if (threadIdx.x == SCAN_SMALL_THREADS - 1) carry_s = run;
    __syncthreads();
  }
  if (threadIdx.x == 0) *total = carry_s;
}

struct Scan {
  DevBuf<int32_t> tile_sum, total;
  int32_t run(Engine& eng, const int32_t* flags, int64_t n, int32_t* pos, int32_t* compact) {
    cudaStream_t st = eng.stream;
    if (n <= SCAN_SMALL_MAX) {
      total.alloc(st, 1);
      ML_LAUNCH(eng, "k_scan_small", k_scan_small<<<1, SCAN_SMALL_THREADS, 0, st>>>(flags, (int)n, pos, compact, total.p));
      int32_t h = 0;
      total.download(&h, 1);
      stream_sync(st);
      return h;
    }
    const int nt = cdiv(n, STILE);
    tile_sum.alloc(st, std::max(1, nt));
    total.alloc(st, 1);
    ML_LAUNCH(eng, "k_scan_tile_sums", k_scan_tile_sums<<<nt, 256, 0, st>>>(flags, n, tile_sum.p));
    ML_LAUNCH(eng, "k_scan_block", k_scan_block<<<1, 256, 0, st>>>(tile_sum.p, nt, total.p));
    ML_LAUNCH(eng, "k_scan_apply", k_scan_apply<<<nt, 256, 0, st>>>(flags, n, tile_sum.p, pos, compact));
    ML_WORK(eng, "k_scan_tile_sums", 4.0 * (double)n);
    ML_WORK(eng, "k_scan_apply", (4.0 + (pos ? 4.0 : 0.0) + (compact ? 4.0 : 0.0)) * (double)n);
    int32_t h = 0;
    total.download(&h, 1);
    stream_sync(st);
    return h;
  }
};

// ---- segmentation kernels ---------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_seg_run_heads(const SegGroup* __restrict__ groups, const int32_t* __restrict__ tile_group,
                const int64_t* __restrict__ tile_r0, const double* __restrict__ beta,
                int32_t* __restrict__ flags) {
  const SegGroup G = groups[tile_group[blockIdx.x]];
  const int64_t r0 = tile_r0[blockIdx.x];
  const int64_t rend = min(r0 + (int64_t)STILE, G.row0 + G.nrows);
  for (int64_t i = r0 + threadIdx.x; i < rend; i += 256)
    flags[i] = (i == G.row0 || beta[i] != beta[i - 1]) ? 1 : 0;
}

// Tolerance merge over runs.  A segment opens at run r when |beta_r - beta_at_segment_start| > tol
// (methylation_helpers.cpp:23; the estimate change of :42 is the group boundary).  The dependence
// on the segment start is sequential, but a run whose jump from its predecessor exceeds 2*tol opens
// a segment whatever the current start is (the predecessor is within tol of that start, triangle
// inequality; the 2.001 factor absorbs rounding), so such CERTAIN heads cut the walk into
// independent pieces: one thread per piece.
__global__ void k_seg_certain(const SegGroup* __restrict__ groups, const int32_t* __restrict__ run_group,
                              const int32_t* __restrict__ run_row, int32_t n_runs,
                              const double* __restrict__ beta, double tol, double* __restrict__ run_beta,
                              int32_t* __restrict__ certain) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_runs) return;
  const int64_t row = run_row[r];
  const double b = beta[row];
  run_beta[r] = b;
  bool c = row == groups[run_group[r]].row0;
  if (!c) c = fabs(b - beta[run_row[r - 1]]) > 2.001 * tol;
  certain[r] = c ? 1 : 0;
}

__global__ void k_seg_merge(const int32_t* __restrict__ piece_run, int32_t n_pieces, int32_t n_runs,
                            const double* __restrict__ run_beta, double tol, int32_t* __restrict__ seg_head) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_pieces) return;
  const int r0 = piece_run[p];
  const int r1 = p + 1 < n_pieces ? piece_run[p + 1] : n_runs;
  double curr = run_beta[r0];
  seg_head[r0] = 1;
  for (int r = r0 + 1; r < r1; ++r) {
    const double b = run_beta[r];
    if (fabs(b - curr) > tol) { seg_head[r] = 1; curr = b; }
    else seg_head[r] = 0;
  }
}

__device__ __forceinline__ int32_t seg_start_at(const SegGroup& G, const void* start, int start_f64,
                                                int64_t row) {
  const int64_t k = G.start0 + (row - G.row0);
  return start_f64 ? (int32_t)((const double*)start)[k] : ((const int32_t*)start)[k];
}

// one thread per segment: rows added in row order (methylation_helpers.cpp:14-18, 24-40, 60-67)
__global__ void k_seg_emit(const SegGroup* __restrict__ groups, const int32_t* __restrict__ row_group,
                           const int32_t* __restrict__ seg_run, int32_t n_segs,
                           const int32_t* __restrict__ run_row, const int32_t* __restrict__ seg_of_run,
                           const int32_t* __restrict__ run_pos, const void* __restrict__ start,
                           int start_f64, const double* __restrict__ beta,
                           const double* __restrict__ betahat, const double* __restrict__ pw, SegOut out,
                           int32_t* __restrict__ long_list, int32_t* __restrict__ n_long) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_segs) return;
  const int r = seg_run[s];
  const int64_t first = run_row[r];
  const int g = row_group[r];   // group of the run (stored per run)
  const SegGroup G = groups[g];
  const int64_t gend = G.row0 + G.nrows;
  int64_t last;       // one past the last row of the segment
  bool next_same_group = false;
  if (s + 1 < n_segs) {
    const int64_t nf = run_row[seg_run[s + 1]];
    if (nf < gend) { last = nf; next_same_group = true; }
    else last = gend;
  } else {
    last = gend;
  }
  const double cb = beta[first];
  const bool is_long = last - first > SEG_LONG;   // summed by a whole block in k_seg_emit_long
  double d0 = betahat[first] - cb;
  double res = pw[first] * d0;
  double var = res * d0;
  double w = pw[first];
  if (is_long) {
    const int q = atomicAdd(n_long, 1);
    long_list[2 * q] = s;
    long_list[2 * q + 1] = (int32_t)(last - first);
  } else {
    for (int64_t i = first + 1; i < last; ++i) {
      const double d = betahat[i] - beta[i];
      res += pw[i] * d;
      var += pw[i] * d * d;
      w += pw[i];
    }
  }
  const int seg_in_group = s - seg_of_run[run_pos[G.row0]];
  if (out.job) out.job[s] = G.job;
  out.first_row[s] = (int32_t)first;
  out.estimate_id[s] = G.estimate_id;
  out.seg_id[s] = seg_in_group + 1;
  out.start[s] = (uint32_t)seg_start_at(G, start, start_f64, first);
  out.end[s] = next_same_group ? (uint32_t)(seg_start_at(G, start, start_f64, last) - 1)
                               : (uint32_t)seg_start_at(G, start, start_f64, gend - 1);
  out.beta[s] = cb;
  if (is_long) return;
  out.pseudoweight[s] = w;
  out.betahat[s] = res / w + cb;
  out.stdev[s] = sqrt(var / w - res * res / (w * w));
}

// Segments of more than SEG_LONG rows (background stretches, fused PMD blocks: up to 10^5 rows) would make one
// thread the tail of the whole launch: one block each instead.  Every thread adds its rows (stride 256) in row
// order and the partial sums are combined by a fixed tree, so the result depends on the segment alone; it differs
// from the reference's left-to-right sums in the last bits only (betahat and stdev of such segments; their
// boundaries, beta and - integer coverage sums being exact in any order - pseudoweight are unaffected).
__global__ void __launch_bounds__(256)
k_seg_emit_long(const int32_t* __restrict__ long_list, const int32_t* __restrict__ n_long,
                const double* __restrict__ beta, const double* __restrict__ betahat, const double* __restrict__ pw,
                SegOut out) {
  __shared__ double sm[3][8];
  const int n = *n_long;
  for (int q = blockIdx.x; q < n; q += gridDim.x) {
    const int s = long_list[2 * q];
    const int64_t first = out.first_row[s], last = first + long_list[2 * q + 1];
    double res = 0.0, var = 0.0, w = 0.0;
    for (int64_t i = first + threadIdx.x; i < last; i += 256) {
      const double d = betahat[i] - beta[i], t = pw[i] * d;
      res += t;
      var += t * d;
      w += pw[i];
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      res += __shfl_xor_sync(0xffffffffu, res, d);
      var += __shfl_xor_sync(0xffffffffu, var, d);
      w += __shfl_xor_sync(0xffffffffu, w, d);
    }
    __syncthreads();                                   // sm free again (previous segment of this block)
    if ((threadIdx.x & 31) == 0) { sm[0][threadIdx.x >> 5] = res; sm[1][threadIdx.x >> 5] = var; sm[2][threadIdx.x >> 5] = w; }
    __syncthreads();
    if (threadIdx.x == 0) {
      res = sm[0][0]; var = sm[1][0]; w = sm[2][0];
      for (int k = 1; k < 8; ++k) { res += sm[0][k]; var += sm[1][k]; w += sm[2][k]; }
      const double cb = out.beta[s];
      out.pseudoweight[s] = w;
      out.betahat[s] = res / w + cb;
      out.stdev[s] = sqrt(var / w - res * res / (w * w));
    }
  }
}

__global__ void k_run_group(const int32_t* __restrict__ run_row, int32_t n_runs,
                            const int64_t* __restrict__ group_row0, int ngroups,
                            int32_t* __restrict__ run_group) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_runs) return;
  const int64_t row = run_row[r];
  int lo = 0, hi = ngroups;   // last group with row0 <= row
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (group_row0[mid] <= row) lo = mid; else hi = mid;
  }
  run_group[r] = lo;
}

// first segment of every group (and the total as a sentinel)
__global__ void k_group_seg0(const SegGroup* __restrict__ groups, int ngroups, const int32_t* __restrict__ run_pos,
                             const int32_t* __restrict__ seg_of_run, int32_t n_segs, int32_t* __restrict__ out) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g < ngroups) out[g] = seg_of_run[run_pos[groups[g].row0]];
  if (g == ngroups) out[g] = n_segs;
}

__global__ void k_group_heads(const int32_t* __restrict__ idx, int64_t n, int32_t* __restrict__ flags) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) flags[i] = (i == 0 || idx[i] != idx[i - 1]) ? 1 : 0;
}

// PMD inputs from a segment table (R/call.R:57-93, native part): y = segment sd with NA -> 0
// (R/call.R:68), weight = width in bp.
__global__ void k_pmd_inputs(int64_t n, const uint32_t* __restrict__ start, const uint32_t* __restrict__ end,
                             const double* __restrict__ stdev, double* __restrict__ y, double* __restrict__ w,
                             int32_t* __restrict__ start_i32) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double sd = stdev[i];
  y[i] = is_finite(sd) ? sd : 0.0;
  w[i] = (double)end[i] - (double)start[i] + 1.0;
  start_i32[i] = (int32_t)start[i];
}


// ---- call table (R/call.R:43-68) ----------------------------------------------------------------
// Where the per-position methylation value and coverage of a row come from.
struct RowsFromEstimates {   // FAST result: betahat = pooled Nmeth / pooled coverage, pw = pooled coverage (exact
  const double* betahat;     // integers in double): Nmeth = rint(betahat * pw) recovers the pooled count exactly,
  const double* pw;          // and Nmeth / pw is then R's own `Nmeth/coverage`.  Parameters without data in this
  __device__ bool valid(int64_t i) const { return pw[i] > 0; }         // condition (pw == 0) are not rows of mdata.
  __device__ double cov(int64_t i) const { return pw[i]; }
  __device__ double x(int64_t i) const { return div_counts(rint(betahat[i] * pw[i]), pw[i]); }
  // the same in two steps: plain loads first (raw, cov), arithmetic later
  __device__ double raw(int64_t i) const { return betahat[i]; }
  __device__ bool valid_cov(double c) const { return c > 0; }
  __device__ double x_from(double r, double c) const { return div_counts(rint(r * c), c); }
};
struct RowsFromCounts {      // host table: pooled Nmeth, coverage per position as given
  const int32_t* nmeth;
  const int32_t* coverage;
  __device__ bool valid(int64_t) const { return true; }
  __device__ double cov(int64_t i) const { return (double)coverage[i]; }
  __device__ double x(int64_t i) const { return div_counts(nmeth[i], coverage[i]); }
  __device__ double raw(int64_t i) const { return (double)nmeth[i]; }
  __device__ bool valid_cov(double) const { return true; }
  __device__ double x_from(double r, double c) const { return div_nz(r, c); }
};

// double-double accumulator standing in for R's long double sums (summary.c real_mean, cov.c)
struct DD { double hi, lo; };
__device__ __forceinline__ void dd_add(DD& a, double x) {
  const double s = a.hi + x;
  const double bb = s - a.hi;
  a.lo += (a.hi - (s - bb)) + (x - bb);
  a.hi = s;
}

__device__ __forceinline__ int group_of_segment(const int32_t* __restrict__ group_seg0, int ng, int s) {
  int lo = 0, hi = ng;       // last group with group_seg0[g] <= s
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (group_seg0[mid] <= s) lo = mid; else hi = mid;
  }
  return lo;
}

// rows of segment s: start_s - 1 <= pos <= end_s within its group (foverlaps of [pos, pos+1] with [start, end])
__device__ __forceinline__ void segment_rows(const SegGroup& G, const void* pos, int pos_f64, uint32_t s0, uint32_t s1,
                                             int64_t& lo, int64_t& hi) {
  const int64_t want_lo = (int64_t)s0 - 1, want_hi = (int64_t)s1;
  int64_t a = G.row0, b = G.row0 + G.nrows;
  while (a < b) { const int64_t m = (a + b) >> 1; if ((int64_t)seg_start_at(G, pos, pos_f64, m) < want_lo) a = m + 1; else b = m; }
  lo = a;
  b = G.row0 + G.nrows;
  while (a < b) { const int64_t m = (a + b) >> 1; if ((int64_t)seg_start_at(G, pos, pos_f64, m) <= want_hi) a = m + 1; else b = m; }
  hi = a;
}

// Row range [lo, hi) and group of every segment.  From the helper's own bookkeeping when the segments were
// computed on these very rows (no search: the rows of segment s run from its first row to the next segment's,
// plus the row before if it sits at start - 1), by binary search otherwise.
__global__ void k_call_ranges(const SegGroup* __restrict__ groups, const int32_t* __restrict__ group_seg0, int ng, int n_segs,
                              const uint32_t* __restrict__ seg_start, const uint32_t* __restrict__ seg_end,
                              const int32_t* __restrict__ first_row, const void* __restrict__ pos, int pos_f64,
                              int32_t* __restrict__ seg_group, int32_t* __restrict__ row_lo, int32_t* __restrict__ row_hi) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_segs) return;
  const int g = group_of_segment(group_seg0, ng, s);
  const SegGroup G = groups[g];
  seg_group[s] = g;
  int64_t lo, hi;
  if (first_row) {
    lo = first_row[s];
    hi = s + 1 < group_seg0[g + 1] ? (int64_t)first_row[s + 1] : G.row0 + G.nrows;
    if (lo > G.row0 && (int64_t)seg_start_at(G, pos, pos_f64, lo - 1) == (int64_t)seg_start[s] - 1) --lo;
  } else {
    segment_rows(G, pos, pos_f64, seg_start[s], seg_end[s], lo, hi);
  }
  if (seg_start[s] > seg_end[s]) hi = lo;                    // sg <- segmentation[start <= end]
  row_lo[s] = (int32_t)lo;
  row_hi[s] = (int32_t)hi;
}

// pass 1, one warp per segment: number of parts (maximal runs without a gap > max_distance) and of oversize gaps
template <class Rows>
__global__ void __launch_bounds__(256)
k_call_count(const SegGroup* __restrict__ groups, const int32_t* __restrict__ seg_group, int n_segs,
             const int32_t* __restrict__ row_lo, const int32_t* __restrict__ row_hi,
             const void* __restrict__ pos, int pos_f64, Rows rows, double max_distance,
             int32_t* __restrict__ parts, int32_t* __restrict__ gaps) {
  const int s = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  const int lane = threadIdx.x & 31;
  if (s >= n_segs) return;
  int nvalid = 0, ngap = 0;
  {
    const SegGroup G = groups[seg_group[s]];
    const int64_t lo = row_lo[s], hi = row_hi[s];
    for (int64_t i = lo + lane; i < hi; i += 32) {
      if (!rows.valid(i)) continue;
      ++nvalid;
      int64_t j = i - 1;                                     // previous valid row of the segment, if any
      while (j >= lo && !rows.valid(j)) --j;
      if (j >= lo && (double)seg_start_at(G, pos, pos_f64, i) - (double)seg_start_at(G, pos, pos_f64, j) > max_distance) ++ngap;
    }
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    nvalid += __shfl_xor_sync(0xffffffffu, nvalid, d);
    ngap += __shfl_xor_sync(0xffffffffu, ngap, d);
  }
  if (lane == 0) {
    parts[s] = nvalid > 0 ? ngap + 1 : 0;
    gaps[s] = nvalid > 0 ? ngap : 0;
  }
}

struct CallOut {
  SegOut t;
  double *coverage, *width;
  int32_t* num_cpgs;
};

__device__ __forceinline__ DD dd_merge(const DD& a, const DD& b) {
  const double s = a.hi + b.hi;
  const double bb = s - a.hi;
  double e = (a.hi - (s - bb)) + (b.hi - bb);
  e += a.lo + b.lo;
  const double hi = s + e;
  return DD{hi, e - (hi - s)};
}
__device__ __forceinline__ DD dd_warp_sum(DD v) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    DD o{__shfl_xor_sync(0xffffffffu, v.hi, d), __shfl_xor_sync(0xffffffffu, v.lo, d)};
    v = dd_merge(v, o);
  }
  return v;
}

// pass 2, one warp per segment: first row of each of its parts (a valid row without a valid predecessor in the
// segment, or further than max_distance from it) -> part_first[part_off[s] + k], part_seg[...] = s
template <class Rows>
__global__ void __launch_bounds__(256)
k_call_parts(const SegGroup* __restrict__ groups, const int32_t* __restrict__ seg_group, int n_segs,
             const int32_t* __restrict__ row_lo, const int32_t* __restrict__ row_hi,
             const void* __restrict__ pos, int pos_f64, Rows rows, double max_distance,
             const int32_t* __restrict__ parts, const int32_t* __restrict__ part_off,
             int32_t* __restrict__ part_first, int32_t* __restrict__ part_seg, int32_t* __restrict__ part_end) {
  const int s = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  const int lane = threadIdx.x & 31;
  if (s >= n_segs) return;
  if (parts[s] == 0) return;
  const SegGroup G = groups[seg_group[s]];
  const int64_t lo = row_lo[s], hi = row_hi[s];
  const int o0 = part_off[s];
  int o = o0;
  for (int64_t base = lo; base < hi; base += 32) {
    const int64_t i = base + lane;
    bool head = false;
    if (i < hi && rows.valid(i)) {
      int64_t j = i - 1;
      while (j >= lo && !rows.valid(j)) --j;
      head = j < lo || (double)seg_start_at(G, pos, pos_f64, i) - (double)seg_start_at(G, pos, pos_f64, j) > max_distance;
    }
    const unsigned bal = __ballot_sync(0xffffffffu, head);
    if (head) {
      const int k = o + __popc(bal & ((1u << lane) - 1u));
      part_first[k] = (int32_t)i;
      part_seg[k] = s;
      if (part_end && k > o0) part_end[k - 1] = (int32_t)i;   // a part ends where the next one of the segment begins
    }
    o += __popc(bal);
  }
  if (part_end && lane == 0 && o > o0) part_end[o - 1] = (int32_t)hi;
}

// pass 3, PART_LANES lanes per part (row o of the call table; coalesced row sweeps).  The three sums of R's mean() / sd()
// (sum, correction about the first mean, squares about the refined mean) are reduced across the lanes in
// double-double, whose error (~1e-32 relative) is far below the final rounding to double, so the order of the
// additions does not show.
#ifndef ML_PART_LANES
#define ML_PART_LANES 16
#endif
constexpr int PART_LANES = ML_PART_LANES;     // lanes per part: a part holds ~25-50 positions on WGBS data, four parts share a warp
__device__ __forceinline__ DD dd_group_sum(DD v) {
#pragma unroll
  for (int d = PART_LANES / 2; d > 0; d >>= 1) {
    DD o{__shfl_xor_sync(0xffffffffu, v.hi, d), __shfl_xor_sync(0xffffffffu, v.lo, d)};
    v = dd_merge(v, o);
  }
  return v;
}
template <class Rows>
__global__ void __launch_bounds__(256, 4)
k_call_emit_part(const SegGroup* __restrict__ groups, const int32_t* __restrict__ group_seg0,
                 const int32_t* __restrict__ seg_group, int n_parts, const int32_t* __restrict__ row_hi,
                 const int32_t* __restrict__ seg_id, const double* __restrict__ seg_pw,
                 const void* __restrict__ pos, int pos_f64, Rows rows,
                 const int32_t* __restrict__ parts, const int32_t* __restrict__ part_off,
                 const int32_t* __restrict__ gap_off, const int32_t* __restrict__ part_first,
                 const int32_t* __restrict__ part_seg, const int32_t* __restrict__ part_end, CallOut out) {
  const int o_raw = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / PART_LANES);
  const int lane = threadIdx.x & (PART_LANES - 1);
  const bool live = o_raw < n_parts;        // the groups of a warp shuffle together: nobody leaves early
  const int o = live ? o_raw : n_parts - 1;
  // the row range comes straight from the part tables, so that the row sweeps start one load after the kernel does;
  // the chain part -> segment -> group is only needed for the labels written at the end
  const int64_t a = part_first[o], b = part_end[o];
  const int s = part_seg[o];
  const int g = seg_group[s];
  const SegGroup G = groups[g];
  const int part = o - part_off[s];
  const int shift = gap_off[s] - gap_off[group_seg0[g]] + part;     // oversize gaps so far in the condition
  // the methylation values of the first CACHE sweeps of the group stay in registers for the second and third pass
  constexpr int CACHE = 4;
  double xc[CACHE];
  unsigned have = 0;                                 // bit c: sweep c holds a row in this lane
  DD sum{0.0, 0.0};
  double csum = 0.0;
  int m = 0;
  int64_t last = a;
  // the loads of all CACHE sweeps are issued before anything is done with them (the kernel waits for memory: ncu
  // shows stall_long_sb on the first use of every value loaded here)
  double cvc[CACHE];
#pragma unroll
  for (int c = 0; c < CACHE; ++c) {
    const int64_t i = a + lane + PART_LANES * c;
    const bool in = i < b;
    cvc[c] = in ? rows.cov(i) : 0.0;
    xc[c] = in ? rows.raw(i) : 0.0;
  }
#pragma unroll
  for (int c = 0; c < CACHE; ++c) {
    const int64_t i = a + lane + PART_LANES * c;
    const bool ok = i < b && rows.valid_cov(cvc[c]);
    xc[c] = ok ? rows.x_from(xc[c], cvc[c]) : 0.0;
    if (ok) { have |= 1u << c; dd_add(sum, xc[c]); csum += cvc[c]; ++m; last = i; }
  }
  for (int64_t i = a + lane + PART_LANES * CACHE; i < b; i += PART_LANES) {
    if (!rows.valid(i)) continue;
    dd_add(sum, rows.x(i));
    csum += rows.cov(i);
    ++m;
    last = i;
  }
  sum = dd_group_sum(sum);
#pragma unroll
  for (int d = PART_LANES / 2; d > 0; d >>= 1) {
    csum += __shfl_xor_sync(0xffffffffu, csum, d);
    m += __shfl_xor_sync(0xffffffffu, m, d);
    last = max(last, (int64_t)__shfl_xor_sync(0xffffffffu, (long long)last, d));
  }
  double mean = (sum.hi + sum.lo) / m;
  DD t{0.0, 0.0};
  if (is_finite(mean)) {
#pragma unroll
    for (int c = 0; c < CACHE; ++c) if (have >> c & 1u) dd_add(t, xc[c] - mean);
    for (int64_t i = a + lane + PART_LANES * CACHE; i < b; i += PART_LANES) if (rows.valid(i)) dd_add(t, rows.x(i) - mean);
  }
  t = dd_group_sum(t);                               // every lane of the warp takes part in every shuffle
  if (is_finite(mean)) mean += (t.hi + t.lo) / m;
  DD ss{0.0, 0.0};
  if (m >= 2) {
#pragma unroll
    for (int c = 0; c < CACHE; ++c) if (have >> c & 1u) { const double d = xc[c] - mean; dd_add(ss, d * d); }
    for (int64_t i = a + lane + PART_LANES * CACHE; i < b; i += PART_LANES) if (rows.valid(i)) { const double d = rows.x(i) - mean; dd_add(ss, d * d); }
  }
  ss = dd_group_sum(ss);
  const double sd = m >= 2 ? sqrt((ss.hi + ss.lo) / (m - 1)) : 0.0;
  if (lane == 0 && live) {
    const double p0 = (double)seg_start_at(G, pos, pos_f64, a), p1 = (double)seg_start_at(G, pos, pos_f64, last);
    if (out.t.job) out.t.job[o] = G.job;
    out.t.first_row[o] = (int32_t)a;
    out.t.estimate_id[o] = G.estimate_id;
    out.t.seg_id[o] = seg_id[s] + shift;
    out.t.start[o] = (uint32_t)p0;
    out.t.end[o] = (uint32_t)p1 + 2u;
    out.width[o] = (p1 + 2.0) - p0 + 1.0;
    out.t.beta[o] = mean;
    out.t.betahat[o] = mean;
    out.t.pseudoweight[o] = seg_pw[s];
    out.t.stdev[o] = sd;
    out.coverage[o] = csum;
    out.num_cpgs[o] = m;
  }
}

// first row of the call table of every group (and the total as a sentinel)
__global__ void k_call_group0(const int32_t* __restrict__ group_seg0, int ng, int n_segs, const int32_t* __restrict__ part_off,
                              int32_t total, int32_t* __restrict__ g0) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g > ng) return;
  const int s = g < ng ? group_seg0[g] : n_segs;
  g0[g] = s < n_segs ? part_off[s] : total;
}
}  // namespace

// std fusion + re-segmentation of a segment table (R/call.R:93-95): fused = weighted_1d_flsa(std,
// width, pmd_lambda2) per (job, estimate); then segment_methylation_helper on
// (estimate_id, start, beta = fused, betahat = std, pseudoweight).
int64_t pmd_segments_device(Engine& eng, const SegDeviceOut& seg, double pmd_lambda2, double tol_val,
                            DevBuf<double>& fused, SegDeviceOut& out) {
  cudaStream_t st = eng.stream;
  const int64_t n = seg.n;
  if (n == 0) return 0;
  std::vector<FlsaSeries> series;
  std::vector<SegGroup> groups;
  for (size_t g = 0; g + 1 < seg.group_seg0.size(); ++g) {
    const int64_t i = seg.group_seg0[g], j = seg.group_seg0[g + 1];
    if (j <= i) continue;
    series.push_back(FlsaSeries{i, (int32_t)(j - i), pmd_lambda2});
    groups.push_back(SegGroup{i, j - i, i, seg.group_est[g], seg.group_job[g]});
  }
  DevBuf<double> y(st, (size_t)n), w(st, (size_t)n);
  DevBuf<int32_t> start_i32(st, (size_t)n);
  fused.alloc(st, (size_t)n);
  ML_LAUNCH(eng, "k_pmd_inputs", k_pmd_inputs<<<cdiv(n, 256), 256, 0, st>>>(n, seg.start.p, seg.end.p, seg.stdev.p,
                                                                          y.p, w.p, start_i32.p));
  FlsaPlan plan(eng, series, n);
  plan.solve(y.p, w.p, fused.p, 1, 0.0, nullptr);
  const int64_t S = segment_device(eng, groups, n, start_i32.p, 0, fused.p, y.p, seg.pseudoweight.p, tol_val, out);
  return S;
}

// Per-segment call table (R/call.R:43-68) from the helper's segments `seg` and the per-position rows they were
// computed on.  `groups` are the row groups the segmentation ran on, in the same order as seg.group_*.
template <class Rows>
static int64_t call_table_impl(Engine& eng, const std::vector<SegGroup>& groups_in, const SegDeviceOut& seg,
                               const void* d_pos, int pos_f64, Rows rows, double max_distance, CallTable& out) {
  cudaStream_t st = eng.stream;
  std::vector<SegGroup> groups;
  for (const SegGroup& g : groups_in)
    if (g.nrows > 0) groups.push_back(g);
  const int ng = (int)groups.size();
  const int n_segs = (int)seg.n;
  if (ng == 0 || n_segs == 0) { out.t.alloc(st, 0); out.t.group_seg0.assign(1, 0); return 0; }
  if ((int)seg.group_seg0.size() != ng + 1) throw MlError(19, "call table: segments and row groups do not match");
  DevBuf<SegGroup> d_groups(st, ng);
  d_groups.upload(groups);
  DevBuf<int32_t> d_g0(st, ng + 1);
  d_g0.upload(seg.group_seg0);
  DevBuf<int32_t> parts(st, n_segs), gaps(st, n_segs), part_off(st, n_segs), gap_off(st, n_segs);
  DevBuf<int32_t> seg_group(st, n_segs), row_lo(st, n_segs), row_hi(st, n_segs);
  ML_LAUNCH(eng, "k_call_ranges", k_call_ranges<<<cdiv(n_segs, 128), 128, 0, st>>>(
      d_groups.p, d_g0.p, ng, n_segs, seg.start.p, seg.end.p, seg.has_first_row ? seg.first_row.p : nullptr, d_pos, pos_f64,
      seg_group.p, row_lo.p, row_hi.p));
  ML_LAUNCH(eng, "k_call_count", k_call_count<Rows><<<cdiv((long long)n_segs * 32, 256), 256, 0, st>>>(
      d_groups.p, seg_group.p, n_segs, row_lo.p, row_hi.p, d_pos, pos_f64, rows, max_distance, parts.p, gaps.p));
  double call_rows = 0.0;                  // rows of the estimates table the segments lie on
  for (const SegGroup& g : groups) call_rows += (double)g.nrows;
  ML_WORK(eng, "k_call_count", 16.0 * call_rows);          // position and pooled coverage of every row
  Scan scan;
  scan.run(eng, gaps.p, n_segs, gap_off.p, nullptr);
  const int32_t total = scan.run(eng, parts.p, n_segs, part_off.p, nullptr);
  out.t.alloc(st, total);
  out.coverage.alloc(st, (size_t)std::max(1, total));
  out.width.alloc(st, (size_t)std::max(1, total));
  out.num_cpgs.alloc(st, (size_t)std::max(1, total));
  CallOut co{out.t.view(), out.coverage.p, out.width.p, out.num_cpgs.p};
  DevBuf<int32_t> part_first(st, (size_t)std::max(1, total)), part_seg(st, (size_t)std::max(1, total)),
      part_end(st, (size_t)std::max(1, total));
  ML_LAUNCH(eng, "k_call_parts", k_call_parts<Rows><<<cdiv((long long)n_segs * 32, 256), 256, 0, st>>>(
      d_groups.p, seg_group.p, n_segs, row_lo.p, row_hi.p, d_pos, pos_f64, rows, max_distance, parts.p, part_off.p,
      part_first.p, part_seg.p, part_end.p));
  ML_WORK(eng, "k_call_parts", 16.0 * call_rows);
  if (total > 0) ML_WORK(eng, "k_call_emit_part", 24.0 * call_rows + 80.0 * (double)total);   // pos, betahat, pw once; the row of the table
  if (total > 0)
    ML_LAUNCH(eng, "k_call_emit_part", k_call_emit_part<Rows><<<cdiv((long long)total * PART_LANES, 256), 256, 0, st>>>(
        d_groups.p, d_g0.p, seg_group.p, total, row_hi.p, seg.seg_id.p, seg.pseudoweight.p, d_pos, pos_f64, rows,
        parts.p, part_off.p, gap_off.p, part_first.p, part_seg.p, part_end.p, co));
  DevBuf<int32_t> d_c0(st, ng + 1);
  ML_LAUNCH(eng, "k_call_group0", k_call_group0<<<cdiv(ng + 1, 128), 128, 0, st>>>(d_g0.p, ng, n_segs, part_off.p, total, d_c0.p));
  std::vector<int32_t> c0(ng + 1);
  d_c0.download(c0.data(), ng + 1);
  stream_sync(st);
  out.t.group_seg0.assign(c0.begin(), c0.end());
  out.t.group_job = seg.group_job;
  out.t.group_est = seg.group_est;
  return total;
}

int64_t call_table_from_estimates(Engine& eng, const std::vector<SegGroup>& groups, const SegDeviceOut& seg,
                                  const void* d_pos, int pos_f64, const double* d_betahat, const double* d_pw,
                                  double max_distance, CallTable& out) {
  return call_table_impl(eng, groups, seg, d_pos, pos_f64, RowsFromEstimates{d_betahat, d_pw}, max_distance, out);
}

int64_t call_table_from_counts(Engine& eng, const std::vector<SegGroup>& groups, const SegDeviceOut& seg,
                               const int32_t* d_pos, const int32_t* d_nmeth, const int32_t* d_cov,
                               double max_distance, CallTable& out) {
  return call_table_impl(eng, groups, seg, d_pos, 0, RowsFromCounts{d_nmeth, d_cov}, max_distance, out);
}

// Core: groups over device arrays -> segments in device buffers owned by `ds`.
int64_t segment_device(Engine& eng, const std::vector<SegGroup>& groups_in, int64_t nrows,
                       const void* d_start, int start_f64, const double* d_beta, const double* d_betahat,
                       const double* d_pw, double tol_val, SegDeviceOut& ds) {
  cudaStream_t st = eng.stream;
  std::vector<SegGroup> groups;
  for (const SegGroup& g : groups_in)
    if (g.nrows > 0) groups.push_back(g);
  const int ng = (int)groups.size();
  if (ng == 0 || nrows == 0) return 0;
  if (nrows >= 0x7fffffff) throw MlError(19, "segment table too large");
  std::vector<int32_t> tile_group;
  std::vector<int64_t> tile_r0, group_row0(ng);
  for (int g = 0; g < ng; ++g) {
    group_row0[g] = groups[g].row0;
    for (int64_t r = groups[g].row0; r < groups[g].row0 + groups[g].nrows; r += STILE) {
      tile_group.push_back(g);
      tile_r0.push_back(r);
    }
  }
  DevBuf<SegGroup> d_groups(st, ng);
  d_groups.upload(groups);
  DevBuf<int32_t> d_tile_group(st, tile_group.size());
  d_tile_group.upload(tile_group);
  DevBuf<int64_t> d_tile_r0(st, tile_r0.size());
  d_tile_r0.upload(tile_r0);
  DevBuf<int64_t> d_group_row0(st, ng);
  d_group_row0.upload(group_row0);
  DevBuf<int32_t> flags(st, (size_t)nrows), run_pos(st, (size_t)nrows);
  flags.zero();   // rows outside every group (none in practice) are not heads
  ML_LAUNCH(eng, "k_seg_run_heads", k_seg_run_heads<<<(int)tile_group.size(), 256, 0, st>>>(d_groups.p, d_tile_group.p, d_tile_r0.p, d_beta,
                                                          flags.p));
  ML_WORK(eng, "k_seg_run_heads", 12.0 * (double)nrows);      // beta read, one flag written
  Scan scan;
  DevBuf<int32_t> run_row(st, (size_t)nrows);
  const int32_t n_runs = scan.run(eng, flags.p, nrows, run_pos.p, run_row.p);
  DevBuf<int32_t> seg_head(st, std::max(1, n_runs)), seg_of_run(st, std::max(1, n_runs)),
      seg_run(st, std::max(1, n_runs)), run_group(st, std::max(1, n_runs)), certain(st, std::max(1, n_runs)),
      piece_run(st, std::max(1, n_runs));
  DevBuf<double> run_beta(st, std::max(1, n_runs));
  ML_LAUNCH(eng, "k_run_group", k_run_group<<<cdiv(n_runs, 256), 256, 0, st>>>(run_row.p, n_runs, d_group_row0.p, ng, run_group.p));
  ML_LAUNCH(eng, "k_seg_certain", k_seg_certain<<<cdiv(n_runs, 256), 256, 0, st>>>(d_groups.p, run_group.p, run_row.p, n_runs, d_beta,
                                                                               tol_val, run_beta.p, certain.p));
  const int32_t n_pieces = scan.run(eng, certain.p, n_runs, nullptr, piece_run.p);
  ML_LAUNCH(eng, "k_seg_merge", k_seg_merge<<<cdiv(n_pieces, 128), 128, 0, st>>>(piece_run.p, n_pieces, n_runs, run_beta.p, tol_val,
                                                                            seg_head.p));
  const int32_t n_segs = scan.run(eng, seg_head.p, n_runs, seg_of_run.p, seg_run.p);
  ds.alloc(st, n_segs);
  SegOut out = ds.view();
  {   // per-group segment ranges for callers that go on working on the table (PMD fusion)
    DevBuf<int32_t> d_g0(st, ng + 1);
    ML_LAUNCH(eng, "k_group_seg0", k_group_seg0<<<cdiv(ng + 1, 128), 128, 0, st>>>(d_groups.p, ng, run_pos.p, seg_of_run.p,
                                                                               n_segs, d_g0.p));
    std::vector<int32_t> g0(ng + 1);
    d_g0.download(g0.data(), ng + 1);
    stream_sync(st);
    ds.group_seg0.assign(g0.begin(), g0.end());
    ds.group_job.resize(ng);
    ds.group_est.resize(ng);
    for (int g = 0; g < ng; ++g) { ds.group_job[g] = groups[g].job; ds.group_est[g] = groups[g].estimate_id; }
  }
  // at most nrows / SEG_LONG segments can be long
  const int64_t max_long = nrows / SEG_LONG + 1;
  DevBuf<int32_t> long_list(st, (size_t)(2 * max_long)), n_long(st, 1);
  n_long.zero();
  ML_LAUNCH(eng, "k_seg_emit", k_seg_emit<<<cdiv(n_segs, 128), 128, 0, st>>>(d_groups.p, run_group.p, seg_run.p, n_segs, run_row.p,
                                                seg_of_run.p, run_pos.p, d_start, start_f64, d_beta,
                                                d_betahat, d_pw, out, long_list.p, n_long.p));
  ML_WORK(eng, "k_seg_emit", 28.0 * (double)nrows + 64.0 * (double)n_segs);   // start, beta, betahat, pw of every row once (long
                                                                              // segments: by k_seg_emit_long); the segment rows
  ML_LAUNCH(eng, "k_seg_emit_long",
            k_seg_emit_long<<<(int)std::min<int64_t>(max_long, (int64_t)eng.sm_count * 8), 256, 0, st>>>(
                long_list.p, n_long.p, d_beta, d_betahat, d_pw, out));
  ds.has_first_row = true;
  stream_sync(st);
  return n_segs;
}

// Host table in, host segments out (the Rcpp signature).
int64_t segment_host_table(Engine& eng, int64_t n, const int32_t* estimate_id, const int32_t* start,
                           const double* beta, const double* betahat, const double* pw, double tol_val,
                           SegHostOut& ho) {
  if (n <= 0) throw MlError(19, "segment_methylation_helper needs at least one row");
  cudaStream_t st = eng.stream;
  DevBuf<int32_t> d_idx(st, (size_t)n), d_start(st, (size_t)n);
  DevBuf<double> d_beta(st, (size_t)n), d_bh(st, (size_t)n), d_pw(st, (size_t)n);
  d_idx.upload(estimate_id, n);
  d_start.upload(start, n);
  d_beta.upload(beta, n);
  d_bh.upload(betahat, n);
  d_pw.upload(pw, n);
  // groups = maximal blocks of equal estimate_id (the reference compares neighbours only)
  DevBuf<int32_t> gflags(st, (size_t)n), ghead(st, (size_t)n);
  ML_LAUNCH(eng, "k_group_heads", k_group_heads<<<cdiv(n, 256), 256, 0, st>>>(d_idx.p, n, gflags.p));
  Scan scan;
  const int32_t ng = scan.run(eng, gflags.p, n, nullptr, ghead.p);
  std::vector<int32_t> heads(ng);
  ghead.download(heads.data(), ng);
  stream_sync(st);
  std::vector<SegGroup> groups(ng);
  for (int g = 0; g < ng; ++g) {
    const int64_t r0 = heads[g], r1 = g + 1 < ng ? heads[g + 1] : n;
    groups[g] = SegGroup{r0, r1 - r0, r0, estimate_id[r0], 0};
  }
  SegDeviceOut ds;
  const int64_t S = segment_device(eng, groups, n, d_start.p, 0, d_beta.p, d_bh.p, d_pw.p, tol_val, ds);
  ds.download(st, S, ho);
  stream_sync(st);
  return S;
}

#include "diffcall.inc.cuh"
#include "lmrumr.inc.cuh"
#include "valleys.inc.cuh"
#include "pmdsplit.inc.cuh"
#include "pmdcall.inc.cuh"
#include "pmdstatus.inc.cuh"
#include "regions.inc.cuh"

void segment_init_device_tables(cudaStream_t st) {
  init_rcp_table(st);          // this translation unit's copy of the reciprocal table (counts_div.cuh)
  ML_CUDA(cudaStreamSynchronize(st));
}

namespace {
__global__ void k_debug_div_counts(int max_num, int max_den, unsigned long long* __restrict__ bad) {
  const long long total = (long long)(max_num + 1) * (max_den - 1);
  unsigned long long mine = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int num = (int)(i % (max_num + 1)), den = 1 + (int)(i / (max_num + 1));
    const double want = (double)num / (double)den;
    const double a = div_counts(num, den), b = div_counts((double)num, (double)den);
    mine += (f64_bits(a) != f64_bits(want)) + (f64_bits(b) != f64_bits(want));
  }
  if (mine) atomicAdd(bad, mine);
}
}  // namespace
void debug_div_counts(Engine& eng, int max_num, int max_den, long long* pairs, long long* mismatches) {
  cudaStream_t st = eng.stream;
  DevBuf<unsigned long long> bad(st, 1);
  bad.zero();
  ML_LAUNCH(eng, "k_debug_div_counts", k_debug_div_counts<<<1184, 256, 0, st>>>(max_num, max_den, bad.p));
  unsigned long long h = 0;
  bad.download(&h, 1);
  stream_sync(st);
  *pairs = (long long)(max_num + 1) * (max_den - 1);
  *mismatches = (long long)h;
}

}  // namespace ml

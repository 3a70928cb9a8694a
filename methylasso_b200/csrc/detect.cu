// detect.cu — batched methylation_detection on sm_100a.
//
// One call processes many JOBS (chromosomes) at once.  All per-CpG data are struct-of-arrays in HBM:
//   rows   : dataset, condition, replicate, pos, nmeth, cov (int32, as uploaded), pidx (int32 row ->
//            parameter), mu / mu_res / mu_outer (f64)
//   params : betahat, pw (pseudo-weight), beta, old_beta (f64), laid out [job][estimator][P]
//   table  : rowstart[job][dataset][P] (int32) first row of dataset d that falls on parameter k
// The reference's Binner / Design / Pseudodata objects (std::map, Eigen sparse) become these index
// arrays; every kernel is a coalesced sweep over rows or parameters, with per-tile partial
// reductions combined in a fixed order (deterministic, independent of job placement).
//
// Reference path restated here (file:line relative to /root/reference/src):
//   MethData.hpp:13-21, core/Observations.hpp:29-59           -> k_validate
//   unique_positions.cpp:16-31, core/binner_helpers.cpp:24-43,47-95 -> bitmap kernels + k_rank_unique
//   binned_genome.cpp:16-31, core/binner_helpers.cpp:11-22    -> k_even_bins
//   core/binned_residuals_helpers.cpp:9-17, core/pseudodata_helpers.hpp:27-39 -> k_pseudodata
//   offset.cpp:14-32, core/fitter_mean.hpp:56-61              -> k_offset_partial / k_offset_finalize
//   core/params_helpers.hpp:23-30                             -> k_center
//   core/Posterior.ipp:63-82, NormalDistribution.hpp:46-51, BetaBinomialRateDistribution.hpp:69-90,
//   GFLLibrary1D.hpp:62-67, predictor_helpers.cpp:10-12       -> k_eval_rows / k_tv_partial / k_job_finalize
//   core/fitter_lasso.hpp:61-68, fitter_mean.hpp:63-70        -> k_damp
//   core/Posterior.ipp:22-59, 99-110, 204-252                 -> host state machine in batch_detect
#include <algorithm>
#include <cmath>
#include <cstring>
#include <atomic>

#include "detect.cuh"
#include "hostpack.h"
#include "tiles.cuh"
#include "counts_div.cuh"

namespace ml {

// ------------------------------------------------------------------------------------------------
// small device helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double block_sum_ordered(double v, double* smem /*>= 8*/) {
  // fixed tree: lanes by shuffle, warps left to right; result valid in thread 0
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_down_sync(0xffffffffu, v, d);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  if (lane == 0) smem[wid] = v;
  __syncthreads();
  double r = 0.0;
  if (threadIdx.x == 0)
    for (int k = 0; k < nw; ++k) r += smem[k];
  __syncthreads();
  return r;
}
__device__ __forceinline__ double block_max(double v, double* smem) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v = fmax(v, __shfl_down_sync(0xffffffffu, v, d));
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  if (lane == 0) smem[wid] = v;
  __syncthreads();
  double r = 0.0;
  if (threadIdx.x == 0) {
    r = smem[0];
    for (int k = 1; k < nw; ++k) r = fmax(r, smem[k]);
  }
  __syncthreads();
  return r;
}

// first index in [0, n) with a[i] >= v (a sorted ascending)
__device__ __forceinline__ int lower_bound_i32(const int32_t* __restrict__ a, int n, int32_t v) {
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (a[mid] < v) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// ------------------------------------------------------------------------------------------------
// Row tiles.  They lie on a grid of ROW_TILE rows counted from the job's first row and are cut at dataset
// boundaries; thread t of a tile owns the four consecutive rows base + 4 t .. + 3 (base = first row of the grid
// cell).  Jobs start at multiples of ROW_ALIGN in the device columns, so every 16-byte load is aligned, and which
// thread adds which row depends on the job alone: a job gives the same bits wherever it is placed.
// ------------------------------------------------------------------------------------------------
struct MakeRowTile {   // tile i of the span (job a, dataset b) over the job-relative rows [i0, i0 + n)
  __device__ RowTile operator()(const TileSpan& s, int i, int64_t, int32_t) const {
    const int64_t lo = s.i0, hi = s.i0 + s.n;
    const int64_t base = (lo / ROW_TILE + i) * (int64_t)ROW_TILE;
    const int64_t b = lo > base ? lo : base, e = hi < base + ROW_TILE ? hi : base + ROW_TILE;
    return RowTile{s.a, s.b, (int32_t)b, (int32_t)(e - b)};
  }
};
struct MakeParTile { __device__ ParTile operator()(const TileSpan& s, int, int64_t off, int32_t cnt) const { return ParTile{s.a, s.b, (int32_t)(s.i0 + off), cnt}; } };
inline int32_t grid_tiles(int64_t lo, int64_t n) { return n > 0 ? (int32_t)((lo + n - 1) / ROW_TILE - lo / ROW_TILE + 1) : 0; }

struct Rows4 {
  int i0;          // job-relative index of the thread's first row (a multiple of 4)
  unsigned mask;   // bit q: row i0 + q belongs to the tile
};
__device__ __forceinline__ Rows4 rows4(const RowTile& t) {
  Rows4 r;
  r.i0 = (t.r0 & ~(ROW_TILE - 1)) + 4 * (int)threadIdx.x;
  r.mask = 0;
#pragma unroll
  for (int q = 0; q < 4; ++q)
    if (r.i0 + q >= t.r0 && r.i0 + q < t.r0 + t.cnt) r.mask |= 1u << q;
  return r;
}
__device__ __forceinline__ void ld4(const int32_t* __restrict__ col, int64_t g, int (&v)[4]) {
  const int4 x = __ldg(reinterpret_cast<const int4*>(col + g));
  v[0] = x.x; v[1] = x.y; v[2] = x.z; v[3] = x.w;
}
__device__ __forceinline__ void ld4(const double* __restrict__ col, int64_t g, double (&v)[4]) {
  const double2 a = __ldg(reinterpret_cast<const double2*>(col + g));
  const double2 b = __ldg(reinterpret_cast<const double2*>(col + g + 2));
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
}
// rows of the mask only (the others may belong to another job's tile running at the same time)
__device__ __forceinline__ void st4(double* __restrict__ col, int64_t g, unsigned mask, const double (&v)[4]) {
  if (mask == 0xFu) {
    *reinterpret_cast<double2*>(col + g) = make_double2(v[0], v[1]);
    *reinterpret_cast<double2*>(col + g + 2) = make_double2(v[2], v[3]);
  } else {
#pragma unroll
    for (int q = 0; q < 4; ++q)
      if (mask >> q & 1) col[g + q] = v[q];
  }
}
__device__ __forceinline__ void st4(int32_t* __restrict__ col, int64_t g, unsigned mask, const int (&v)[4]) {
  if (mask == 0xFu) {
    *reinterpret_cast<int4*>(col + g) = make_int4(v[0], v[1], v[2], v[3]);
  } else {
#pragma unroll
    for (int q = 0; q < 4; ++q)
      if (mask >> q & 1) col[g + q] = v[q];
  }
}

// FAST model: the IRLS weight of a row is W = 1 / V with V = (sigma * sigma) / nobs, sigma = 1 (irls_core.cuh), two IEEE
// divisions whose result depends on the integer coverage alone.  The first FAST_W_TABLE values are computed once
// per device by those very divisions (k_init_fast_w) and looked up afterwards: same bits, two divisions fewer per row.
constexpr int FAST_W_TABLE = 1024;
__device__ double g_fast_w[FAST_W_TABLE];
__global__ void k_init_fast_w() {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= FAST_W_TABLE) return;
  double z, W;
  irls_residual(MODEL_FAST, 1, 0.0, (double)i, 0.0, 0.0, z, W);
  g_fast_w[i] = W;
}
__device__ __forceinline__ void fast_residual(const double* __restrict__ tab, int first, int nmeth, int cov, double mu_prev,
                                              double& z, double& W) {
  const double nobs = (double)cov;
  const double y = div_counts(nmeth, cov);
  if ((unsigned)cov < (unsigned)FAST_W_TABLE) {
    W = tab[cov];
    z = first ? y : (y - mu_prev);           // identity_link.cpp:20-23 | :32-36
  } else {
    irls_residual(MODEL_FAST, first, y, nobs, mu_prev, 0.0, z, W);
  }
}
// The FIRST step's contribution of one row to the pseudodata of its parameter, W * zhat with zhat = (z W) / W and
// z = Nmeth / coverage (binned_residuals_helpers.cpp:9-17, pseudodata_helpers.hpp:27-39), depends on the two counts alone:
// for counts below FAST_T_SIDE it is computed once per device by the very operations of k_pseudodata_fast and looked up
// afterwards - same bits, a quotient of counts, a product and an IEEE division fewer per row (the kernel was bound by
// instruction issue: 550 instructions per parameter with three datasets).
constexpr int FAST_T_SIDE = 256;
__device__ double g_fast_t[FAST_T_SIDE * FAST_T_SIDE];     // [coverage][Nmeth]
__device__ __forceinline__ double fast_first_term(const double* __restrict__ wtab, int nmeth, int cov, double& Wd) {
  double z, W;
  fast_residual(wtab, 1, nmeth, cov, 0.0, z, W);
  Wd = 0.0 + W;                          // Binner::bin starts its sums at zero (Binner.cpp:29-36)
  const double Zd = 0.0 + z * W;
  const double q = Zd / Wd;
  const double zhat = (Wd == 0) ? Wd : q;
  return Wd * zhat;
}
__global__ void k_init_fast_t() {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= FAST_T_SIDE * FAST_T_SIDE) return;
  double Wd;
  g_fast_t[i] = fast_first_term(g_fast_w, i % FAST_T_SIDE, i / FAST_T_SIDE, Wd);
}
// working residual and weight of one row, whichever the model
__device__ __forceinline__ void row_residual(int model, int first, int nmeth, int cov, double mu_prev, double nu,
                                             double& z, double& W) {
  if (model == MODEL_FAST) {
    fast_residual(g_fast_w, first, nmeth, cov, mu_prev, z, W);
  } else {
    const double nobs = (double)cov;
    irls_residual(model, first, div_counts(nmeth, cov), nobs, mu_prev, nu, z, W);
  }
}

// ------------------------------------------------------------------------------------------------
// K1: the first sweep over the rows.  One pass over pos / Nmeth / coverage (12 bytes per row) does what the
// reference does in four walks over its Observations:
//   * validation (core/Observations.hpp:29-59 then src/MethData.hpp:17-19): every row evaluates its own predicate
//     and the smallest (stage, row) wins, which is the row and the message of the reference's first throw.  The
//     predicates on dataset / condition / replicate were evaluated on the run table by the host (batch_upload);
//     what is left here is the order of the positions within a dataset and the two checks on the counts;
//   * the positions are marked in the job's bitmap (FAST leg, K2 below);
//   * Posterior::initial_guess + the offset's binned residuals (Posterior.ipp:3-18, offset.cpp:14-32): per-tile
//     sums of W and z W of the first IRLS step;
//   * the -log pdf of the initial evaluation (Posterior.ipp:25): every parameter and offset is still +0.
// ------------------------------------------------------------------------------------------------
#ifndef ML_ROW_MINBLOCKS
#define ML_ROW_MINBLOCKS 4
#endif
template <bool BITMAP>
__global__ void __launch_bounds__(256, ML_ROW_MINBLOCKS)
k_first_sweep(const RowTile* __restrict__ tiles, const JobDev* __restrict__ jobs, const int32_t* __restrict__ alive,
              const int32_t* __restrict__ dptr, const int32_t* __restrict__ pos, const int32_t* __restrict__ nmeth,
              const int32_t* __restrict__ cov, const int64_t* __restrict__ bm0, const int32_t* __restrict__ minpos,
              uint32_t* __restrict__ bitmap, unsigned long long* __restrict__ keys, double log2pi,
              double* __restrict__ off_part /*2 per tile*/, double* __restrict__ row_part /*3 per tile*/) {
  __shared__ double sm[3][8];
  const RowTile t = tiles[blockIdx.x];
  const JobDev J = jobs[t.job];
  // the unique-position leg is the FAST model's: a compile-time model folds the beta-binomial code (lgamma, log, exp)
  // out of this instantiation, which halves its registers
  const int model = BITMAP ? (int)MODEL_FAST : J.model;
  const Rows4 r = rows4(t);
  const bool live = alive[t.job] != 0;
  unsigned long long best = ~0ull;
  double sw = 0.0, sz = 0.0, lik = 0.0;
  if (r.mask) {
    const int64_t g = J.row0 + r.i0;
    int ps[4], nm[4], cv[4];
    ld4(pos, g, ps);
    ld4(nmeth, g, nm);
    ld4(cov, g, cv);
    const int run_lo = dptr[J.dptr0 + t.d];
    const int before = r.i0 > run_lo ? pos[g - 1] : 0;
    const int32_t mp = BITMAP ? minpos[t.job] : 0;
    const int64_t b0 = BITMAP ? bm0[t.job] : 0;
    const uint32_t nbits = BITMAP ? (uint32_t)((bm0[t.job + 1] - b0) * 32) : 0u;
    const double m0 = irls_mu(model, 0. + (0. + 1. * 0.));         // mu of the initial evaluation
    uint32_t cw = 0xffffffffu, cb = 0u;                             // bits of one bitmap word, flushed together
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      if (!(r.mask >> q & 1)) continue;
      const int i = r.i0 + q;
      const unsigned long long li = (unsigned long long)i;
      unsigned long long key = ~0ull;
      if (i > run_lo && (q ? ps[q - 1] : before) >= ps[q]) key = (1ull << 60) | (li << 8) | 6u;
      else if (nm[q] < 0 || cv[q] < 0) key = (2ull << 60) | (li << 8) | 11u;
      else if (nm[q] > cv[q]) key = (3ull << 60) | (li << 8) | 12u;
      best = key < best ? key : best;
      if (!live) continue;
      if (BITMAP) {
        const uint32_t rel = (uint32_t)(ps[q] - mp);
        if (rel < nbits) {                                          // always, unless the positions are out of order
          const uint32_t w = rel >> 5;
          if (w != cw) {
            if (cb) atomicOr(&bitmap[b0 + cw], cb);
            cw = w;
            cb = 0u;
          }
          cb |= 1u << (rel & 31);
        }
      }
      double z, W;
      row_residual(model, 1, nm[q], cv[q], 0.0, J.nu, z, W);
      sw += W;
      sz += z * W;
      const double nobs = (double)cv[q];
      lik += irls_minus_log_pdf(model, div_counts(nm[q], cv[q]), nobs, m0, J.nu, log2pi);
    }
    if (BITMAP && cb) atomicOr(&bitmap[b0 + cw], cb);
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    const unsigned long long o = __shfl_down_sync(0xffffffffu, best, d);
    best = o < best ? o : best;
    sw += __shfl_down_sync(0xffffffffu, sw, d);
    sz += __shfl_down_sync(0xffffffffu, sz, d);
    lik += __shfl_down_sync(0xffffffffu, lik, d);
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) {
    if (best != ~0ull) atomicMin(&keys[t.job], best);
    sm[0][wid] = sw; sm[1][wid] = sz; sm[2][wid] = lik;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, b = 0.0, c = 0.0;
    for (int k = 0; k < 8; ++k) { a += sm[0][k]; b += sm[1][k]; c += sm[2][k]; }
    off_part[2 * blockIdx.x] = a;
    off_part[2 * blockIdx.x + 1] = b;
    row_part[3 * blockIdx.x] = c;
    row_part[3 * blockIdx.x + 1] = 0.0;
    row_part[3 * blockIdx.x + 2] = 0.0;
  }
}

// ------------------------------------------------------------------------------------------------
// K2: unique-position index (FAST leg).  Positions of a job are marked in a bitmap over
// [minpos, maxpos] (by the first sweep); the rank of a position among the distinct ones is a prefix popcount:
// one prefix per BLOCK of 8 words (256 bits = one 32-byte sector) plus the bits below inside the block.
// Replaces make_unique_bins' sort+unique and the std::map upper_bound per row
// (core/binner_helpers.cpp:24-43, 53-60).  Every job's bitmap is a whole number of blocks.
// ------------------------------------------------------------------------------------------------
constexpr int WTILE = 2048;  // bitmap words per scan tile (256 threads x one block of 8 words)
constexpr int BM_BLOCK = 8;  // words per prefix block
struct WTileDesc { int64_t w0; int32_t job; int32_t cnt; };
struct MakeWTile { __device__ WTileDesc operator()(const TileSpan& s, int, int64_t off, int32_t cnt) const { return WTileDesc{s.i0 + off, s.a, cnt}; } };

__device__ __forceinline__ uint32_t popc8(const uint32_t* __restrict__ bitmap, int64_t w0) {
  const uint4 a = *reinterpret_cast<const uint4*>(bitmap + w0), b = *reinterpret_cast<const uint4*>(bitmap + w0 + 4);
  return __popc(a.x) + __popc(a.y) + __popc(a.z) + __popc(a.w) + __popc(b.x) + __popc(b.y) + __popc(b.z) + __popc(b.w);
}

__global__ void __launch_bounds__(256)
k_bitmap_tile_count(const WTileDesc* __restrict__ tiles, const uint32_t* __restrict__ bitmap,
                    uint32_t* __restrict__ tile_count) {
  __shared__ uint32_t sm[8];
  const WTileDesc t = tiles[blockIdx.x];
  const int i0 = threadIdx.x * BM_BLOCK;
  uint32_t c = i0 < t.cnt ? popc8(bitmap, t.w0 + i0) : 0u;
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) c += __shfl_down_sync(0xffffffffu, c, d);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t s = 0;
    for (int k = 0; k < 8; ++k) s += sm[k];
    tile_count[blockIdx.x] = s;
  }
}

// one warp per job: exclusive scan of its tile counts; total = number of distinct positions
__global__ void k_bitmap_job_scan(const int32_t* __restrict__ job_wtile0, int njobs,
                                  uint32_t* __restrict__ tile_count /*in: counts, out: bases*/,
                                  int32_t* __restrict__ job_unique) {
  const int j = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (j >= njobs) return;
  const int t0 = job_wtile0[j], t1 = job_wtile0[j + 1];
  uint32_t run = 0;
  for (int base = t0; base < t1; base += 32) {
    const int t = base + lane;
    const uint32_t c = t < t1 ? tile_count[t] : 0;
    uint32_t inc = c;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t o = __shfl_up_sync(0xffffffffu, inc, d);
      if (lane >= d) inc += o;
    }
    if (t < t1) tile_count[t] = run + inc - c;
    run += __shfl_sync(0xffffffffu, inc, 31);
  }
  if (lane == 0) job_unique[j] = (int32_t)run;
}

// per block of 8 words: number of set bits before it within its job
__global__ void __launch_bounds__(256)
k_bitmap_block_prefix(const WTileDesc* __restrict__ tiles, const uint32_t* __restrict__ bitmap,
                      const uint32_t* __restrict__ tile_base, uint32_t* __restrict__ block_prefix) {
  __shared__ uint32_t sm[8];
  const WTileDesc t = tiles[blockIdx.x];
  const int i0 = threadIdx.x * BM_BLOCK;
  const uint32_t mine = i0 < t.cnt ? popc8(bitmap, t.w0 + i0) : 0u;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  uint32_t inc = mine;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t o = __shfl_up_sync(0xffffffffu, inc, d);
    if (lane >= d) inc += o;
  }
  if (lane == 31) sm[wid] = inc;
  __syncthreads();
  uint32_t before = tile_base[blockIdx.x];
  for (int k = 0; k < wid; ++k) before += sm[k];
  if (i0 < t.cnt) block_prefix[(t.w0 + i0) / BM_BLOCK] = before + inc - mine;
}

__device__ __forceinline__ int bitmap_rank(const uint32_t* __restrict__ bitmap, const uint32_t* __restrict__ block_prefix,
                                           int64_t b0, uint32_t rel) {
  const int64_t w = b0 + (rel >> 5), wb = w & ~(int64_t)(BM_BLOCK - 1);
  const uint4 lo = __ldg(reinterpret_cast<const uint4*>(bitmap + wb)), hi = __ldg(reinterpret_cast<const uint4*>(bitmap + wb + 4));
  const uint32_t ws[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
  const int in = (int)(w - wb);
  uint32_t c = __ldg(block_prefix + wb / BM_BLOCK);
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    if (k < in) c += __popc(ws[k]);
    else if (k == in) c += __popc(ws[k] & ((1u << (rel & 31)) - 1u));
  }
  return (int)c;
}

// rows: parameter index = rank of the position; also the (dataset, param) -> row table and, when asked for, the
// support (bin start = the position itself, BinDesign::starts; otherwise the pseudodata kernel writes it)
__global__ void __launch_bounds__(256)
k_rank_unique(const RowTile* __restrict__ tiles, const JobDev* __restrict__ jobs,
              const int32_t* __restrict__ pos, const int64_t* __restrict__ bm0,
              const int32_t* __restrict__ minpos, const uint32_t* __restrict__ bitmap,
              const uint32_t* __restrict__ block_prefix, const int64_t* __restrict__ start0,
              const int32_t* __restrict__ alive, int32_t* __restrict__ pidx,
              int32_t* __restrict__ rowstart, double* __restrict__ start /*may be null*/) {
  const RowTile t = tiles[blockIdx.x];
  if (!alive[t.job]) return;
  const Rows4 r = rows4(t);
  if (!r.mask) return;
  const JobDev J = jobs[t.job];
  const int32_t mp = minpos[t.job];
  const int64_t b0 = bm0[t.job];
  const int64_t g = J.row0 + r.i0;
  int ps[4], ks[4];
  ld4(pos, g, ps);
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    ks[q] = 0;
    if (!(r.mask >> q & 1)) continue;
    const int k = bitmap_rank(bitmap, block_prefix, b0, (uint32_t)(ps[q] - mp));
    ks[q] = k;
    rowstart[J.tab0 + (int64_t)t.d * J.P + k] = r.i0 + q;
    // BinDesign::starts holds the position as double(unsigned); the estimates table reports
    // int(start) (methylation_helpers.hpp:20).  Every dataset writes the same value.
    if (start) start[start0[t.job] + k] = (double)(int32_t)(double)(uint32_t)ps[q];
  }
  st4(pidx, g, r.mask, ks);
}

// ------------------------------------------------------------------------------------------------
// K3: even-bin index (FULL leg): binned_genome.cpp:22-25 + binner_helpers.cpp:11-22, 53-60
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_even_bins(const RowTile* __restrict__ tiles, const JobDev* __restrict__ jobs,
            const int32_t* __restrict__ pos, const int32_t* __restrict__ alive,
            int32_t* __restrict__ pidx) {
  const RowTile t = tiles[blockIdx.x];
  if (!alive[t.job]) return;
  const JobDev J = jobs[t.job];
  for (int i = threadIdx.x; i < t.cnt; i += blockDim.x) {
    const int64_t r = J.row0 + t.r0 + i;
    pidx[r] = (int32_t)even_bin((double)pos[r], J.lower, J.size, J.nbins);
  }
}
// first row of every (dataset, bin) run; rows of a dataset are sorted so bins form runs
__global__ void __launch_bounds__(256)
k_rowstart_runs(const RowTile* __restrict__ tiles, const JobDev* __restrict__ jobs,
                const int32_t* __restrict__ dptr, const int32_t* __restrict__ pidx,
                const int32_t* __restrict__ alive, int32_t* __restrict__ rowstart) {
  const RowTile t = tiles[blockIdx.x];
  if (!alive[t.job]) return;
  const JobDev J = jobs[t.job];
  const int dstart = dptr[J.dptr0 + t.d];
  for (int i = threadIdx.x; i < t.cnt; i += blockDim.x) {
    const int rl = t.r0 + i;
    const int k = pidx[J.row0 + rl];
    if (rl == dstart || pidx[J.row0 + rl - 1] != k) rowstart[J.tab0 + (int64_t)t.d * J.P + k] = rl;
  }
}
__global__ void __launch_bounds__(PAR_TILE)
k_support_even(const ParTile* __restrict__ tiles, const JobDev* __restrict__ jobs,
               const int64_t* __restrict__ start0, double* __restrict__ start) {
  const ParTile t = tiles[blockIdx.x];   // parameter tiles exist only for jobs that are still alive
  if (t.e != 0) return;
  const JobDev J = jobs[t.job];
  const int k = t.k0 + threadIdx.x;
  if (threadIdx.x < t.cnt)
    start[start0[t.job] + k] = (double)(int32_t)even_boundary(J.lower, J.size, (uint32_t)k, J.nbins);
}

// ------------------------------------------------------------------------------------------------
// K4+K5: residuals -> binned residuals -> pseudodata, one thread per (estimator, parameter).
//   What_d = sum_rows W            (Binner::bin, sequential in row order)
//   zhat_d = sum_rows z W / What_d (0 where What_d == 0)
//   pw     = sum_d (c What_d) c ;  betahat = sum_d (pw^+ c)(What_d zhat_d) + beta
// Datasets are visited in increasing order, rows in row order: the reference's summation order.
// ------------------------------------------------------------------------------------------------
struct RowCols {
  const int32_t* __restrict__ nmeth;
  const int32_t* __restrict__ cov;
  const int32_t* __restrict__ pidx;
  const double* __restrict__ mu_res;
};

__device__ __forceinline__ void bin_residuals(const JobDev& J, const RowCols& rc, int first, int r0,
                                              int rend, int k, double& What, double& Zs) {
  What = 0.0;
  Zs = 0.0;
  for (int r = r0; r < rend; ++r) {
    const int64_t g = J.row0 + r;
    if (r != r0 && rc.pidx[g] != k) break;
    const double nobs = (double)rc.cov[g];
    const double y = div_counts(rc.nmeth[g], rc.cov[g]);
    double z, W;
    irls_residual(J.model, first, y, nobs, first ? 0.0 : rc.mu_res[g], J.nu, z, W);
    What += W;
    Zs += z * W;
  }
}

__global__ void __launch_bounds__(PAR_TILE, 4)
k_pseudodata(const ParTile* __restrict__ tiles, const JobDev* __restrict__ jobs,
             const int32_t* __restrict__ ctl, RowCols rc, const int32_t* __restrict__ dptr,
             const double* __restrict__ design, const int32_t* __restrict__ rowstart,
             double* __restrict__ beta, double* __restrict__ old_beta, double* __restrict__ betahat,
             double* __restrict__ pw) {
  const ParTile t = tiles[blockIdx.x];
  const int c = ctl[t.job];
  if (!(c & CTL_UPDATE)) return;
  if ((int)threadIdx.x >= t.cnt) return;
  const JobDev J = jobs[t.job];
  const int first = (c & CTL_FIRST) ? 1 : 0;
  const int k = t.k0 + threadIdx.x;
  const int64_t pi = J.par0 + (int64_t)t.e * J.P + k;
  constexpr int DC = 4;                // datasets whose binned residuals are kept in registers
  double Wd[DC], Zd[DC], Cd[DC];
  double wt = 0.0;
  // the first-row lookups of all datasets are independent: issue them together
  int r0s[DC];
#pragma unroll
  for (int d = 0; d < DC; ++d) {
    Cd[d] = d < J.D ? design[J.des0 + d * J.E + t.e] : 0.0;
    r0s[d] = (d < J.D && Cd[d] != 0) ? rowstart[J.tab0 + (int64_t)d * J.P + k] : -1;
  }
#pragma unroll
  for (int d = 0; d < DC; ++d) {
    Wd[d] = 0.0;
    Zd[d] = 0.0;
    if (r0s[d] >= 0) {
      bin_residuals(J, rc, first, r0s[d], dptr[J.dptr0 + d + 1], k, Wd[d], Zd[d]);
      wt += (Cd[d] * Wd[d]) * Cd[d];
    }
  }
  for (int d = DC; d < J.D; ++d) {     // more than DC datasets: no caching
    const double cf = design[J.des0 + d * J.E + t.e];
    if (cf == 0) continue;
    const int r0 = rowstart[J.tab0 + (int64_t)d * J.P + k];
    if (r0 < 0) continue;
    double What, Zs;
    bin_residuals(J, rc, first, r0, dptr[J.dptr0 + d + 1], k, What, Zs);
    wt += (cf * What) * cf;
  }
  double inv = 1 / wt;
  if (!is_finite(inv)) inv = 0.0;     // pseudo-inverse (pseudodata_helpers.hpp:33-34)
  double bh = 0.0;
#pragma unroll
  for (int d = 0; d < DC; ++d) {
    if (r0s[d] >= 0) {
      const double q = Zd[d] / Wd[d];
      const double zhat = (Wd[d] == 0) ? Wd[d] : q;
      bh += (inv * Cd[d]) * (Wd[d] * zhat);
    }
  }
  for (int d = DC; d < J.D; ++d) {
    const double cf = design[J.des0 + d * J.E + t.e];
    if (cf == 0) continue;
    const int r0 = rowstart[J.tab0 + (int64_t)d * J.P + k];
    if (r0 < 0) continue;
    double What, Zs;
    bin_residuals(J, rc, first, r0, dptr[J.dptr0 + d + 1], k, What, Zs);
    const double q = Zs / What;
    const double zhat = (What == 0) ? What : q;
    bh += (inv * cf) * (What * zhat);
  }
  const double b = beta[pi];
  old_beta[pi] = b;                   // Estimator::update_params: set_old_params (Estimator.ipp:22)
  betahat[pi] = bh + b;
  pw[pi] = wt;
}

// The same for FAST jobs with at most PD_FAST_D datasets: positions are unique within a dataset, so a (dataset,
// parameter) bin holds exactly one row - no row loop; the first-row lookups, then the row loads of all datasets are
// issued together (the kernel is bound by the latency of these dependent loads), and W comes from the table.
// Operations and their order are those of k_pseudodata (What = 0 + W, Zs = 0 + z W, ...): same bits.
constexpr int PD_FAST_D = 8;
template <int PD_D>                      // datasets held in registers: 4 (more resident warps) or PD_FAST_D
__global__ void __launch_bounds__(PAR_TILE, PD_D <= 4 ? 6 : 4)
k_pseudodata_fast(const ParTile* __restrict__ tiles, const JobDev* __restrict__ jobs,
                  const int32_t* __restrict__ ctl, RowCols rc, const double* __restrict__ design,
                  const int32_t* __restrict__ rowstart, double* __restrict__ beta, double* __restrict__ old_beta,
                  double* __restrict__ betahat, double* __restrict__ pw, const int32_t* __restrict__ pos,
                  const int64_t* __restrict__ start0, double* __restrict__ start /*null: already written*/) {
  // the table is read from global memory (L1): measured 0.93 ms against 1.08 ms with a per-block copy in shared memory
  const double* wtab = g_fast_w;
  const ParTile t = tiles[blockIdx.x];
  const int c = ctl[t.job];
  if (!(c & CTL_UPDATE)) return;
  if ((int)threadIdx.x >= t.cnt) return;
  const JobDev J = jobs[t.job];
  const int first = (c & CTL_FIRST) ? 1 : 0;
  const int k = t.k0 + threadIdx.x;
  const int64_t pi = J.par0 + (int64_t)t.e * J.P + k;
  double Cd[PD_D];
  int r0s[PD_D], cv[PD_D], nm[PD_D];
  double mp[PD_D];
#pragma unroll
  for (int d = 0; d < PD_D; ++d) {
    Cd[d] = d < J.D ? design[J.des0 + d * J.E + t.e] : 0.0;
    r0s[d] = (d < J.D && Cd[d] != 0) ? rowstart[J.tab0 + (int64_t)d * J.P + k] : -1;
  }
#pragma unroll
  for (int d = 0; d < PD_D; ++d) {
    const int64_t g = J.row0 + (r0s[d] >= 0 ? r0s[d] : 0);
    cv[d] = r0s[d] >= 0 ? rc.cov[g] : 1;
    nm[d] = r0s[d] >= 0 ? rc.nmeth[g] : 0;
    mp[d] = (r0s[d] >= 0 && !first) ? rc.mu_res[g] : 0.0;
  }
  const double b = beta[pi];
  if (start && t.e == 0) {
    // the support (BinDesign::starts): the position of the parameter, read from the first dataset that has it.
    // With several estimators a position may belong to datasets of another condition only: look at all datasets.
    int rs = -1;
    for (int d = 0; d < J.D && rs < 0; ++d) rs = d < PD_D && r0s[d] >= 0 ? r0s[d] : rowstart[J.tab0 + (int64_t)d * J.P + k];
    if (rs >= 0) start[start0[t.job] + k] = (double)(int32_t)(double)(uint32_t)pos[J.row0 + rs];
  }
  double Wd[PD_D], Td[PD_D];             // per dataset: W and W * zhat
  double wt = 0.0;
#pragma unroll
  for (int d = 0; d < PD_D; ++d) {
    Wd[d] = 0.0;
    Td[d] = 0.0;
    if (r0s[d] >= 0) {
      if (first && (unsigned)cv[d] < (unsigned)FAST_T_SIDE && (unsigned)nm[d] < (unsigned)FAST_T_SIDE) {
        Wd[d] = wtab[cv[d]];             // (0.0 + W is W: the table holds no negative zero)
        Td[d] = g_fast_t[cv[d] * FAST_T_SIDE + nm[d]];
      } else if (first) {
        Td[d] = fast_first_term(wtab, nm[d], cv[d], Wd[d]);
      } else {
        double z, W;
        fast_residual(wtab, first, nm[d], cv[d], mp[d], z, W);
        Wd[d] = 0.0 + W;                 // Binner::bin starts its sums at zero (Binner.cpp:29-36)
        const double Zd = 0.0 + z * W;
        const double q = Zd / Wd[d];
        const double zhat = (Wd[d] == 0) ? Wd[d] : q;
        Td[d] = Wd[d] * zhat;
      }
      wt += (Cd[d] * Wd[d]) * Cd[d];
    }
  }
  double inv = 1 / wt;
  if (!is_finite(inv)) inv = 0.0;        // pseudo-inverse (pseudodata_helpers.hpp:33-34)
  double bh = 0.0;
#pragma unroll
  for (int d = 0; d < PD_D; ++d)
    if (r0s[d] >= 0) bh += (inv * Cd[d]) * Td[d];
  old_beta[pi] = b;
  betahat[pi] = bh + b;
  pw[pi] = wt;
}

// ------------------------------------------------------------------------------------------------
// K9: offset (one mean per dataset): per row-tile partial sums, then per dataset in tile order.
// The reference adds the rows of a dataset left to right (Binner.cpp:34); here the association is
// a fixed tree per tile followed by a left-to-right sum over tiles (documented deviation, ~1e-16).
// ------------------------------------------------------------------------------------------------
template <int MODEL>                    // MODEL_FAST: compile-time model (no beta-binomial code); -1: the job's own
__global__ void __launch_bounds__(256)
k_offset_partial(const RowTile* __restrict__ tiles, const JobDev* __restrict__ jobs,
                 const int32_t* __restrict__ ctl, RowCols rc, double* __restrict__ part /*2 per tile*/) {
  __shared__ double sm[2][8];
  const RowTile t = tiles[blockIdx.x];
  const int c = ctl[t.job];
  if (!(c & CTL_UPDATE)) return;
  const JobDev J = jobs[t.job];
  const int first = (c & CTL_FIRST) ? 1 : 0;
  const Rows4 r = rows4(t);
  double sw = 0.0, sz = 0.0;
  if (r.mask) {
    const int64_t g = J.row0 + r.i0;
    int nm[4], cv[4];
    double mr[4] = {0.0, 0.0, 0.0, 0.0};
    ld4(rc.nmeth, g, nm);
    ld4(rc.cov, g, cv);
    if (!first) ld4(rc.mu_res, g, mr);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      if (!(r.mask >> q & 1)) continue;
      double z, W;
      row_residual(MODEL >= 0 ? MODEL : J.model, first, nm[q], cv[q], mr[q], J.nu, z, W);
      sw += W;
      sz += z * W;
    }
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    sw += __shfl_down_sync(0xffffffffu, sw, d);
    sz += __shfl_down_sync(0xffffffffu, sz, d);
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) { sm[0][wid] = sw; sm[1][wid] = sz; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, b = 0.0;
    for (int k = 0; k < 8; ++k) { a += sm[0][k]; b += sm[1][k]; }
    part[2 * blockIdx.x] = a;
    part[2 * blockIdx.x + 1] = b;
  }
}

__global__ void __launch_bounds__(128)
k_offset_finalize(const int32_t* __restrict__ ent_job, const int32_t* __restrict__ ent_d, int n_ent,
                  const JobDev* __restrict__ jobs, const int32_t* __restrict__ ctl,
                  const int32_t* __restrict__ dtile0, const double* __restrict__ part,
                  double* __restrict__ off, double* __restrict__ old_off) {
  __shared__ double sm[8];
  const int i = blockIdx.x;   // one block per (job, dataset): fixed-order sum over that dataset's tiles
  const int j = ent_job[i], d = ent_d[i];
  if (!(ctl[j] & CTL_UPDATE)) return;
  const JobDev J = jobs[j];
  if (d >= J.D) return;
  double What = 0.0, Zs = 0.0;
  for (int t = dtile0[i] + threadIdx.x; t < dtile0[i + 1]; t += 128) { What += part[2 * t]; Zs += part[2 * t + 1]; }
  What = block_sum_ordered(What, sm);
  Zs = block_sum_ordered(Zs, sm);
  if (threadIdx.x != 0) return;
  const double q = Zs / What;
  const double zhat = (What == 0) ? What : q;
  const double pwd = (1. * What) * 1.;
  double inv = 1 / pwd;
  if (!is_finite(inv)) inv = 0.0;
  const double o = off[J.off0 + d];
  old_off[J.off0 + d] = o;
  off[J.off0 + d] = (inv * 1.) * (What * zhat) + o;   // Fitter<Mean>: beta = betahat
}

// centring: beta -= sum(beta*pw)/sum(pw)   (params_helpers.hpp:23-30), and the total variation of the centred
// tile, sum |beta_k - beta_{k-1}| (GFLLibrary1D.hpp:64), for the evaluation that follows: the left neighbour's centred
// value is recomputed from its raw value by the same subtraction, so the sum is the one k_tv_partial would form.
__global__ void __launch_bounds__(PAR_TILE)
k_center(const ParTile* __restrict__ tiles, const JobDev* __restrict__ jobs,
         const int32_t* __restrict__ ctl, const double* __restrict__ sums, const double* __restrict__ raw,
         double* __restrict__ beta, double* __restrict__ tv_part) {
  __shared__ double sm[8];
  const ParTile t = tiles[blockIdx.x];
  if (!(ctl[t.job] & CTL_UPDATE)) return;
  const JobDev J = jobs[t.job];
  const int s = J.ser0 + t.e;
  const double mean = sums[2 * s] / sums[2 * s + 1];
  const int k = t.k0 + threadIdx.x;
  const int64_t pi = J.par0 + (int64_t)t.e * J.P + k;
  double v = 0.0;
  if ((int)threadIdx.x < t.cnt) {
    const double mine = raw[pi] - mean;
    beta[pi] = mine;
    if (k > 0) v = fabs(mine - (raw[pi - 1] - mean));
  }
  const double a = block_sum_ordered(v, sm);
  if (threadIdx.x == 0) tv_part[blockIdx.x] = a;
}

// damping: p <- w p + (1-w) p_old for every lasso estimator and the offset
// (fitter_lasso.hpp:61-68, fitter_mean.hpp:63-70; weight 1-5/10, posterior_policies.hpp:67)
__global__ void __launch_bounds__(PAR_TILE)
k_damp(const ParTile* __restrict__ tiles, const JobDev* __restrict__ jobs,
       const int32_t* __restrict__ ctl, double wnew, double* __restrict__ beta,
       const double* __restrict__ old_beta, double* __restrict__ off, const double* __restrict__ old_off) {
  const ParTile t = tiles[blockIdx.x];
  if (!(ctl[t.job] & CTL_DAMP)) return;
  const JobDev J = jobs[t.job];
  if ((int)threadIdx.x < t.cnt) {
    const int64_t pi = J.par0 + (int64_t)t.e * J.P + t.k0 + threadIdx.x;
    beta[pi] = wnew * beta[pi] + (1 - wnew) * old_beta[pi];
  }
  if (t.e == 0 && t.k0 == 0 && (int)threadIdx.x < J.D) {
    const int o = J.off0 + threadIdx.x;
    off[o] = wnew * off[o] + (1 - wnew) * old_off[o];
  }
  // J.D > PAR_TILE datasets: remaining offsets
  if (t.e == 0 && t.k0 == 0)
    for (int d = PAR_TILE + threadIdx.x; d < J.D; d += PAR_TILE) {
      const int o = J.off0 + d;
      off[o] = wnew * off[o] + (1 - wnew) * old_off[o];
    }
}

// ------------------------------------------------------------------------------------------------
// K10/K11/K13: mu, -log pdf and max |delta mu| per row tile
// ------------------------------------------------------------------------------------------------
template <int MODEL>                    // MODEL_FAST: compile-time model (no beta-binomial code); -1: the job's own
__global__ void __launch_bounds__(256, ML_ROW_MINBLOCKS)
k_eval_rows(const RowTile* __restrict__ tiles, const JobDev* __restrict__ jobs,
            const int32_t* __restrict__ ctl, RowCols rc, const double* __restrict__ design,
            const double* __restrict__ beta, const double* __restrict__ off, double log2pi,
            double* __restrict__ mu, const double* __restrict__ mu_outer, int want_outer,
            double* __restrict__ part /*3 per tile: lik, dres, douter*/) {
  __shared__ double sm[3][8];
  const RowTile t = tiles[blockIdx.x];
  const int c = ctl[t.job];
  if (!(c & CTL_EVAL)) return;
  const JobDev J = jobs[t.job];
  const double o = off[J.off0 + t.d];
  const bool write_mu = !(c & CTL_NO_MU);      // the initial evaluation's mu is never read
  const bool need_dres = (c & CTL_DRES) != 0;  // max|mu - mu_res| is only consulted from the 2nd step on
  // initial evaluation: every parameter is still +0, so the linear response below is +0 whatever the index
  // (0. + cf * (+0.) summed over the estimators) - the row -> parameter index and the gather are not needed
  const bool zero_params = (c & CTL_ZERO_PARAMS) != 0;
  const Rows4 r = rows4(t);
  double lik = 0.0, dres = 0.0, dout = 0.0;
  if (r.mask) {
    const int64_t g = J.row0 + r.i0;
    int cv[4], nm[4], kx[4] = {0, 0, 0, 0};
    ld4(rc.cov, g, cv);                        // independent loads first
    ld4(rc.nmeth, g, nm);
    if (!zero_params) ld4(rc.pidx, g, kx);
    double mr[4] = {0.0, 0.0, 0.0, 0.0}, mo[4] = {0.0, 0.0, 0.0, 0.0};
    if (need_dres) ld4(rc.mu_res, g, mr);
    if (want_outer) ld4(mu_outer, g, mo);
    double eta[4] = {0.0, 0.0, 0.0, 0.0};
    for (int e = 0; e < J.E && !zero_params; ++e) {
      const double cf = design[J.des0 + t.d * J.E + e];
      if (cf != 0) {
        const double* be = beta + J.par0 + (int64_t)e * J.P;
#pragma unroll
        for (int q = 0; q < 4; ++q) eta[q] += (0. + cf * ((r.mask >> q & 1) ? be[kx[q]] : 0.0));
      } else {
#pragma unroll
        for (int q = 0; q < 4; ++q) eta[q] += 0.;
      }
    }
    double m[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      m[q] = irls_mu(MODEL >= 0 ? MODEL : J.model, eta[q] + (0. + 1. * o));
      if (!(r.mask >> q & 1)) continue;
      const double nobs = (double)cv[q];
      const double y = div_counts(nm[q], cv[q]);
      lik += irls_minus_log_pdf(MODEL >= 0 ? MODEL : J.model, y, nobs, m[q], J.nu, log2pi);
      if (need_dres) dres = fmax(dres, fabs(m[q] - mr[q]));
      if (want_outer) dout = fmax(dout, fabs(m[q] - mo[q]));
    }
    if (write_mu) st4(mu, g, r.mask, m);
  }
  // one pass for the three reductions: lanes by shuffle, warps in order (fixed association)
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    lik += __shfl_down_sync(0xffffffffu, lik, d);
    dres = fmax(dres, __shfl_down_sync(0xffffffffu, dres, d));
    dout = fmax(dout, __shfl_down_sync(0xffffffffu, dout, d));
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) { sm[0][wid] = lik; sm[1][wid] = dres; sm[2][wid] = dout; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, b = 0.0, cc = 0.0;
    for (int k = 0; k < 8; ++k) { a += sm[0][k]; b = fmax(b, sm[1][k]); cc = fmax(cc, sm[2][k]); }
    part[3 * blockIdx.x] = a;
    part[3 * blockIdx.x + 1] = b;
    part[3 * blockIdx.x + 2] = cc;
  }
}

// total variation of one parameter tile: sum |beta_k - beta_{k-1}|   (GFLLibrary1D.hpp:64)
__global__ void __launch_bounds__(PAR_TILE)
k_tv_partial(const ParTile* __restrict__ tiles, const JobDev* __restrict__ jobs,
             const int32_t* __restrict__ ctl, const double* __restrict__ beta, double* __restrict__ part) {
  __shared__ double sm[8];
  const ParTile t = tiles[blockIdx.x];
  const int c = ctl[t.job];
  if (!(c & CTL_DAMP)) return;                // after an update k_center has left the tile's sum; only damping moves beta again
  const JobDev J = jobs[t.job];
  const int k = t.k0 + threadIdx.x;
  double v = 0.0;
  if ((int)threadIdx.x < t.cnt && k > 0) {
    const int64_t pi = J.par0 + (int64_t)t.e * J.P + k;
    v = fabs(beta[pi] - beta[pi - 1]);
  }
  const double a = block_sum_ordered(v, sm);
  if (threadIdx.x == 0) part[blockIdx.x] = a;
}

// one block per job: tile partials combined in a fixed order (thread t takes tiles t, t+256, ...,
// then a fixed tree) -> {mlogp, lik, prior, dres, douter}.  The order depends only on the job's own
// tile count, so a job gives the same bits wherever it is placed.
__global__ void __launch_bounds__(256)
k_job_finalize(const JobDev* __restrict__ jobs, int njobs, const int32_t* __restrict__ ctl,
               const double* __restrict__ row_part, const double* __restrict__ tv_part,
               double* __restrict__ scalars /*5 per job*/) {
  __shared__ double sm[8];
  const int j = blockIdx.x;
  if (!(ctl[j] & CTL_EVAL)) return;
  const JobDev J = jobs[j];
  const bool zero_tv = (ctl[j] & CTL_ZERO_PARAMS) != 0;   // initial evaluation: all parameters +0, every |difference| is +0
  double lik = 0.0, dres = 0.0, dout = 0.0;
  for (int t = J.rtile0 + threadIdx.x; t < J.rtile0 + J.n_rtiles; t += 256) {
    lik += row_part[3 * t];
    dres = fmax(dres, row_part[3 * t + 1]);
    dout = fmax(dout, row_part[3 * t + 2]);
  }
  lik = block_sum_ordered(lik, sm);
  dres = block_max(dres, sm);
  dout = block_max(dout, sm);
  double pri = 0.0;
  const int tiles_per_e = J.n_ptiles / J.E;
  for (int e = 0; e < J.E; ++e) {
    double tv = 0.0;
    for (int t = threadIdx.x; t < tiles_per_e && !zero_tv; t += 256) tv += tv_part[J.ptile0 + e * tiles_per_e + t];
    tv = block_sum_ordered(tv, sm);
    const double lam = J.lambda2;
    pri += lam * tv - (double)(uint64_t)(J.P - 2) * log(lam) + (double)(uint64_t)(J.P - 1) * log(2.0);
  }
  if (threadIdx.x == 0) {
    scalars[5 * j] = lik + pri;
    scalars[5 * j + 1] = lik;
    scalars[5 * j + 2] = pri;
    scalars[5 * j + 3] = dres;
    scalars[5 * j + 4] = dout;
  }
}

__global__ void __launch_bounds__(256)
k_copy_rows(const RowTile* __restrict__ tiles, const JobDev* __restrict__ jobs,
            const int32_t* __restrict__ ctl, const double* __restrict__ mu, double* __restrict__ mu_res,
            double* __restrict__ mu_outer) {
  const RowTile t = tiles[blockIdx.x];
  const int c = ctl[t.job];
  if (!(c & (CTL_COPY_RES | CTL_COPY_OUTER))) return;
  const JobDev J = jobs[t.job];
  const Rows4 r = rows4(t);
  if (!r.mask) return;
  const int64_t g = J.row0 + r.i0;
  double m[4];
  ld4(mu, g, m);
  if (c & CTL_COPY_RES) st4(mu_res, g, r.mask, m);
  if ((c & CTL_COPY_OUTER) && mu_outer) st4(mu_outer, g, r.mask, m);
}

// R/fast_diff.R:15-28 on the parameter arrays (estimator 0 = reference condition)
__global__ void __launch_bounds__(PAR_TILE)
k_fast_diff(const ParTile* __restrict__ tiles, const JobDev* __restrict__ jobs, double* __restrict__ beta,
            double* __restrict__ betahat, double* __restrict__ pw) {
  const ParTile t = tiles[blockIdx.x];
  if (t.e == 0 || (int)threadIdx.x >= t.cnt) return;
  const JobDev J = jobs[t.job];
  const int k = t.k0 + threadIdx.x;
  const int64_t pi = J.par0 + (int64_t)t.e * J.P + k, pr = J.par0 + k;
  const double s = pw[pi] + pw[pr];
  double bh = (betahat[pi] * pw[pi] - betahat[pr] * pw[pr]) / s;
  if (s == 0) bh = 0.0;
  beta[pi] = beta[pi] - beta[pr];
  betahat[pi] = bh;
  pw[pi] = s;
}

// estimate_id column of the estimates table (methylation_helpers.hpp:39): e + 1 per parameter
__global__ void __launch_bounds__(PAR_TILE)
k_fill_estimate_id(const ParTile* __restrict__ tiles, const JobDev* __restrict__ jobs, int32_t* __restrict__ id) {
  const ParTile t = tiles[blockIdx.x];
  if ((int)threadIdx.x >= t.cnt) return;
  const JobDev J = jobs[t.job];
  id[J.par0 + (int64_t)t.e * J.P + t.k0 + threadIdx.x] = t.e + 1;
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
// ------------------------------------------------------------------------------------------------
// host side: the run table of a batch.  dataset / condition / replicate are factor codes that the reference only
// ever looks at where the dataset changes (core/Observations.hpp:36-56; data_design_helpers.cpp:17-23), so they are
// reduced to one entry per run of equal dataset codes here and never uploaded.  Finding the runs is a streaming
// compare over the dataset column (a few host threads, while the DMA of pos / Nmeth / coverage is in flight).
// ------------------------------------------------------------------------------------------------
namespace {

std::vector<int64_t> find_transitions(const int32_t* v, int64_t n, int threads) {
  std::vector<int64_t> all;
  if (n < 2) return all;
  const int T = (int)std::max<int64_t>(1, std::min<int64_t>(threads, n / (1 << 18)));
  if (T <= 1) { transitions_in(v, 0, n, all); return all; }
  const int pieces = 4 * T;
  std::vector<std::vector<int64_t>> part(pieces);
  parallel_for(pieces, T, [&](int t) { transitions_in(v, n * t / pieces, n * (t + 1) / pieces, part[t]); });
  for (auto& pv : part) all.insert(all.end(), pv.begin(), pv.end());
  return all;
}

unsigned long long vkey(int stage, int64_t row, int code) {
  return ((unsigned long long)stage << 60) | ((unsigned long long)row << 8) | (unsigned)code;
}

// Walk the runs of one job the way Observations::check_observations walks its rows (core/Observations.hpp:36-56):
// starts[] are the job-relative first rows of the runs, dsv / cond / rep the three codes at those rows.  Appends
// the runs that lie before the first structural offence to the batch's run table and records the offence.
void add_job_runs(Batch& b, int64_t nrows, const std::vector<int64_t>& starts, const std::vector<int32_t>& dsv,
                  const std::vector<int32_t>& cond, const std::vector<int32_t>& rep, const int32_t* pos_job) {
  unsigned long long key = ~0ull;
  size_t known = 0;
  for (size_t t = 0; t < starts.size() && key == ~0ull; ++t) {
    const int64_t r = starts[t];
    int code = 0;
    if (t == 0) {
      if ((uint32_t)dsv[0] != 1u) code = 3;
      else if ((uint32_t)cond[0] != 1u) code = 4;
      else if ((uint32_t)rep[0] != 1u) code = 5;
    } else {
      const uint32_t dp = (uint32_t)dsv[t - 1], dc = (uint32_t)dsv[t];
      const uint32_t lc = (uint32_t)cond[t - 1], lr = (uint32_t)rep[t - 1];   // remembered at the previous run's first row
      if (dc != dp + 1u) code = 7;
      else if ((uint32_t)cond[t] == lc) { if ((uint32_t)rep[t] != lr + 1u) code = 8; }
      else if ((uint32_t)cond[t] != lc + 1u) code = 9;
      else if ((uint32_t)rep[t] != 1u) code = 10;
    }
    if (code) key = vkey(1, r, code);
    else known = t + 1;
  }
  const int64_t covered = key == ~0ull ? nrows : (int64_t)((key >> 8) & 0xfffffffffffffull);
  b.host_key.push_back(key);
  b.dptr0.push_back((int32_t)b.dptr.size());
  for (size_t t = 0; t < known; ++t) {
    const int64_t r0 = starts[t], r1 = t + 1 < known ? starts[t + 1] : covered;
    b.dptr.push_back((int32_t)r0);
    b.dcond.push_back(cond[t]);
    b.drep.push_back(rep[t]);
    b.dfirst.push_back(pos_job[r0]);
    b.dlast.push_back(pos_job[r1 - 1]);
  }
  b.dptr.push_back((int32_t)covered);
  b.run0.push_back((int32_t)b.dcond.size());
}

std::unique_ptr<Batch> batch_begin(Engine& eng, int njobs, const int64_t* job_ptr) {
  if (njobs < 0) throw MlError(19, "njobs < 0");
  auto b = std::make_unique<Batch>();
  b->eng = &eng;
  b->njobs = njobs;
  b->job_ptr.assign(job_ptr, job_ptr + njobs + 1);
  if (njobs && b->job_ptr[0] != 0) throw MlError(19, "job_ptr[0] must be 0");
  b->dev_row0.assign(njobs + 1, 0);
  for (int j = 0; j < njobs; ++j) {
    const int64_t n = b->job_ptr[j + 1] - b->job_ptr[j];
    if (n < 0) throw MlError(19, "job_ptr must be non-decreasing");
    if (n > 0x7fffffff - 2 * ROW_ALIGN) throw MlError(19, "a job has more than 2^31-1 rows");
    b->dev_row0[j + 1] = b->dev_row0[j] + (n + ROW_ALIGN - 1) / ROW_ALIGN * ROW_ALIGN;
  }
  b->nrows = njobs ? b->job_ptr[njobs] : 0;
  b->nrows_dev = b->dev_row0[njobs];
  b->run0.assign(1, 0);
  return b;
}

// pos / Nmeth / coverage of every job to its aligned place in the device columns (asynchronous)
void batch_copy_columns(Batch& b, const int32_t* pos, const int32_t* nmeth, const int32_t* cov) {
  cudaStream_t st = b.eng->stream;
  const size_t n = (size_t)b.nrows_dev + ROW_ALIGN;          // vector loads may touch the padding behind the last job
  b.pos.alloc(st, n);
  b.nmeth.alloc(st, n);
  b.cov.alloc(st, n);
  struct Col { const int32_t* h; int32_t* d; } cols[3] = {{pos, b.pos.p}, {nmeth, b.nmeth.p}, {cov, b.cov.p}};
  for (const Col& c : cols) {
    int j = 0;
    while (j < b.njobs) {                                      // consecutive jobs without padding between them: one copy
      int k = j;
      while (k + 1 < b.njobs && b.dev_row0[k + 1] - b.dev_row0[j] == b.job_ptr[k + 1] - b.job_ptr[j]) ++k;
      const int64_t rows = b.job_ptr[k + 1] - b.job_ptr[j];
      if (rows > 0)
        ML_CUDA(cudaMemcpyAsync(c.d + b.dev_row0[j], c.h + b.job_ptr[j], sizeof(int32_t) * (size_t)rows,
                                cudaMemcpyHostToDevice, st));
      j = k + 1;
    }
  }
}

}  // namespace

// device side of the packing: one thread per four rows
template <typename T>
__global__ void __launch_bounds__(256)
k_unpack_counts(const T* __restrict__ packed, int64_t nrows4, int32_t* __restrict__ nmeth, int32_t* __restrict__ cov) {
  const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= nrows4) return;
  T v[8];
  if (sizeof(T) == 1) *reinterpret_cast<uint2*>(v) = __ldg(reinterpret_cast<const uint2*>(packed) + q);
  else *reinterpret_cast<uint4*>(v) = __ldg(reinterpret_cast<const uint4*>(packed) + q);
  reinterpret_cast<int4*>(nmeth)[q] = make_int4((int)v[0], (int)v[2], (int)v[4], (int)v[6]);
  reinterpret_cast<int4*>(cov)[q] = make_int4((int)v[1], (int)v[3], (int)v[5], (int)v[7]);
}

std::unique_ptr<Batch> batch_stage(Engine& eng, int njobs, const int64_t* job_ptr, const int32_t* dataset,
                                   const int32_t* condition, const int32_t* replicate,
                                   const int32_t* pos, const int32_t* nmeth, const int32_t* cov) {
  auto b = batch_begin(eng, njobs, job_ptr);
  b->h_pos = pos;
  b->h_nmeth = nmeth;
  b->h_cov = cov;
  const int T = std::max(1, eng.host_threads);
  // ---- the run table (dataset / condition / replicate never leave the host) ----
  const std::vector<int64_t> tr = find_transitions(dataset, b->nrows, T);
  size_t ti = 0;
  std::vector<int64_t> starts;
  std::vector<int32_t> dsv, cond, rep;
  for (int j = 0; j < njobs; ++j) {
    const int64_t j0 = b->job_ptr[j], j1 = b->job_ptr[j + 1];
    starts.clear(); dsv.clear(); cond.clear(); rep.clear();
    while (ti < tr.size() && tr[ti] <= j0) ++ti;               // a change exactly at the job's first row is its start anyway
    if (j1 > j0) starts.push_back(0);
    for (; ti < tr.size() && tr[ti] < j1; ++ti) starts.push_back(tr[ti] - j0);
    for (int64_t r : starts) { dsv.push_back(dataset[j0 + r]); cond.push_back(condition[j0 + r]); rep.push_back(replicate[j0 + r]); }
    add_job_runs(*b, j1 - j0, starts, dsv, cond, rep, pos + j0);
    b->last_condition.push_back(j1 > j0 ? condition[j1 - 1] : 0);
  }
  // ---- Nmeth / coverage packed for the bus, laid out as the device columns are (jobs at multiples of ROW_ALIGN) ----
  const size_t need = 4 * ((size_t)b->nrows_dev + ROW_ALIGN);      // two 16-bit counts per row at most
  if (eng.stage_cap < need) {
    if (eng.stage_host) cudaFreeHost(eng.stage_host);
    eng.stage_host = nullptr;
    eng.stage_cap = 0;
    ML_CUDA(cudaHostAlloc((void**)&eng.stage_host, need + need / 8, cudaHostAllocPortable));
    eng.stage_cap = need + need / 8;
  }
  if (b->nrows == 0) { b->pack_bytes = 1; return b; }
  // pieces of at most 2^18 rows, handed out one at a time to the pool's threads (a piece never spans two jobs)
  struct Piece { int job; int64_t lo, hi; };
  std::vector<Piece> pieces;
  for (int j = 0; j < njobs; ++j)
    for (int64_t lo = b->job_ptr[j]; lo < b->job_ptr[j + 1]; lo += (1 << 18))
      pieces.push_back(Piece{j, lo, std::min<int64_t>(lo + (1 << 18), b->job_ptr[j + 1])});
  auto run = [&](auto tag) -> bool {
    typedef decltype(tag) U;
    U* out = reinterpret_cast<U*>(eng.stage_host);
    std::atomic<int> bad{0};
    parallel_for((int)pieces.size(), T, [&](int k) {
      if (bad.load(std::memory_order_relaxed)) return;
      const Piece& pc = pieces[k];
      const int64_t shift = b->dev_row0[pc.job] - b->job_ptr[pc.job];      // caller row -> device row
      bool ok;
      if (sizeof(U) == 1) ok = pack_counts_u8(nmeth, cov, shift, pc.lo + shift, pc.hi + shift, reinterpret_cast<uint8_t*>(out));
      else ok = pack_counts_u16(nmeth, cov, shift, pc.lo + shift, pc.hi + shift, reinterpret_cast<uint16_t*>(out));
      if (!ok) bad.store(1, std::memory_order_relaxed);
    });
    if (bad.load()) return false;
    // the padding rows behind every job: zeros
    for (int j = 0; j < njobs; ++j) {
      const int64_t e = b->dev_row0[j] + (b->job_ptr[j + 1] - b->job_ptr[j]);
      const int64_t pad_end = j + 1 < njobs ? b->dev_row0[j + 1] : b->nrows_dev + ROW_ALIGN;
      memset(out + 2 * e, 0, sizeof(U) * 2 * (size_t)(pad_end - e));
    }
    return true;
  };
  if (run(uint8_t())) b->pack_bytes = 1;
  else if (run(uint16_t())) b->pack_bytes = 2;
  else b->pack_bytes = 0;
  return b;
}

void batch_commit(Engine& eng, Batch& b) {
  if (b.committed) return;
  cudaStream_t st = eng.stream;
  if (b.pack_bytes == 0 || b.nrows == 0) {
    batch_copy_columns(b, b.h_pos, b.h_nmeth, b.h_cov);
  } else {
    const size_t n = (size_t)b.nrows_dev + ROW_ALIGN;
    b.pos.alloc(st, n);
    b.nmeth.alloc(st, n);
    b.cov.alloc(st, n);
    const size_t bytes = 2 * (size_t)b.pack_bytes * n;
    DevBuf<char> packed(st, bytes);
    ML_CUDA(cudaMemcpyAsync(packed.p, eng.stage_host, bytes, cudaMemcpyHostToDevice, st));
    int j = 0;
    while (j < b.njobs) {                                      // consecutive jobs without padding between them: one copy
      int k = j;
      while (k + 1 < b.njobs && b.dev_row0[k + 1] - b.dev_row0[j] == b.job_ptr[k + 1] - b.job_ptr[j]) ++k;
      const int64_t rows = b.job_ptr[k + 1] - b.job_ptr[j];
      if (rows > 0)
        ML_CUDA(cudaMemcpyAsync(b.pos.p + b.dev_row0[j], b.h_pos + b.job_ptr[j], sizeof(int32_t) * (size_t)rows,
                                cudaMemcpyHostToDevice, st));
      j = k + 1;
    }
    const int64_t n4 = (int64_t)(n / 4);
    if (b.pack_bytes == 1)
      ML_LAUNCH(eng, "k_unpack_counts", k_unpack_counts<uint8_t><<<cdiv(n4, 256), 256, 0, st>>>(
          reinterpret_cast<const uint8_t*>(packed.p), n4, b.nmeth.p, b.cov.p));
    else
      ML_LAUNCH(eng, "k_unpack_counts", k_unpack_counts<uint16_t><<<cdiv(n4, 256), 256, 0, st>>>(
          reinterpret_cast<const uint16_t*>(packed.p), n4, b.nmeth.p, b.cov.p));
    eng.prof.add_work("k_unpack_counts", (double)(bytes + 8 * n));
  }
  stream_sync(st);  // the caller may reuse its buffers after return
  b.committed = true;
  b.h_pos = b.h_nmeth = b.h_cov = nullptr;
}

std::unique_ptr<Batch> batch_upload(Engine& eng, int njobs, const int64_t* job_ptr, const int32_t* dataset,
                                    const int32_t* condition, const int32_t* replicate,
                                    const int32_t* pos, const int32_t* nmeth, const int32_t* cov) {
  auto b = batch_stage(eng, njobs, job_ptr, dataset, condition, replicate, pos, nmeth, cov);
  batch_commit(eng, *b);
  return b;
}

// test hook (csrc/debug_api.h): the run table of ONE job, host only
void debug_scan_runs(int64_t n, const int32_t* dataset, const int32_t* condition, const int32_t* replicate,
                     const int32_t* pos, int threads, Batch& b) {
  const std::vector<int64_t> tr = find_transitions(dataset, n, threads);
  std::vector<int64_t> starts;
  std::vector<int32_t> dsv, cond, rep;
  if (n > 0) starts.push_back(0);
  for (int64_t t : tr) starts.push_back(t);
  for (int64_t r : starts) { dsv.push_back(dataset[r]); cond.push_back(condition[r]); rep.push_back(replicate[r]); }
  b.run0.assign(1, 0);
  add_job_runs(b, n, starts, dsv, cond, rep, pos);
}

std::unique_ptr<Batch> batch_upload_runs(Engine& eng, int njobs, const int64_t* job_ptr, const int64_t* dset_ptr,
                                         const int64_t* dset_row, const int32_t* dset_condition,
                                         const int32_t* dset_replicate, const int32_t* pos, const int32_t* nmeth,
                                         const int32_t* cov) {
  auto b = batch_begin(eng, njobs, job_ptr);
  batch_copy_columns(*b, pos, nmeth, cov);
  std::vector<int64_t> starts;
  std::vector<int32_t> dsv, cond, rep;
  for (int j = 0; j < njobs; ++j) {
    const int64_t n = b->job_ptr[j + 1] - b->job_ptr[j];
    const int64_t d0 = dset_ptr[j], d1 = dset_ptr[j + 1];
    if (d1 < d0 || (n > 0 && d1 == d0)) throw MlError(19, "dset_ptr: every job with rows needs at least one dataset");
    starts.clear(); dsv.clear(); cond.clear(); rep.clear();
    for (int64_t d = d0; d < d1; ++d) {
      const int64_t r = dset_row[d];
      if (r < 0 || r >= n || (d == d0 ? r != 0 : r <= dset_row[d - 1]))
        throw MlError(19, "dset_row: first rows must start at 0 and increase strictly within a job");
      starts.push_back(r);
      dsv.push_back((int32_t)(d - d0 + 1));
      cond.push_back(dset_condition[d]);
      rep.push_back(dset_replicate[d]);
    }
    add_job_runs(*b, n, starts, dsv, cond, rep, pos + b->job_ptr[j]);
    b->last_condition.push_back(d1 > d0 ? dset_condition[d1 - 1] : 0);
  }
  stream_sync(eng.stream);
  b->committed = true;
  return b;
}

namespace {

struct Ctx {   // everything one detect call owns on the device besides the Result
  DevBuf<JobDev> jobs;
  DevBuf<RowTile> rtiles;
  DevBuf<ParTile> ptiles;
  DevBuf<int32_t> ctl;
  DevBuf<int32_t> dptr, ent_job, ent_d, dtile0;
  DevBuf<double> design;
  DevBuf<int32_t> pidx, rowstart;
  DevBuf<double> mu_res, mu_outer, old_beta, old_off, beta_raw;
  DevBuf<double> off_part, row_part, tv_part, scalars, scalars_init, sums;
  DevBuf<int64_t> start0;
};

int decode_status(unsigned long long key) { return key == ~0ull ? 0 : (int)(key & 0xff); }

}  // namespace

std::unique_ptr<Result> batch_detect(Engine& eng, Batch& b, const Params& p) {
  if (!b.committed) throw MlError(19, "the batch was staged but never uploaded (ml_batch_upload_staged)");
  if (p.optimize_lambda2 || p.optimize_hyper)
    throw MlError(19, "optimize_lambda2 / optimize_hyper are not supported (FALSE at every reference call site)");
  if (p.model < 0 || p.model > 2) throw MlError(19, "unknown model");
  cudaStream_t st = eng.stream;
  const int nj = b.njobs;
  auto res = std::make_unique<Result>();
  res->eng = &eng;
  res->jobs.resize(nj);
  res->nrows = b.nrows;
  res->nrows_dev = b.nrows_dev;
  if (nj == 0) return res;
  const bool fast = p.model == MODEL_FAST;
  Ctx cx;

  // ---- host: per-job status from the run table, design, geometry -----------------------------------
  // (the checks of core/Observations.hpp:36-56 on dataset / condition / replicate were made on the run table by
  // batch_upload; the per-row checks follow on the device and the smallest (stage, row) key wins)
  const int n_ent = (int)b.dptr.size();            // one entry per (job, dataset) plus one closing entry per job
  std::vector<int32_t> ent_job(n_ent), ent_d(n_ent);
  std::vector<JobDev> hj(nj);
  std::vector<double> h_design;
  std::vector<int32_t> h_minpos(nj, 0), h_sweep(nj, 0);
  std::vector<int64_t> h_bm0(nj + 1, 0);
  int32_t noff = 0;
  for (int j = 0; j < nj; ++j) {
    JobHost& jh = res->jobs[j];
    JobDev& J = hj[j];
    memset(&J, 0, sizeof(J));
    const int64_t n = b.job_ptr[j + 1] - b.job_ptr[j];
    const int K = b.n_runs(j);
    for (int d = 0; d <= K; ++d) { ent_job[b.dptr0[j] + d] = j; ent_d[b.dptr0[j] + d] = d; }
    jh.nrows = n;
    jh.row0 = b.dev_row0[j];
    jh.lambda2 = p.lambda2;
    J.row0 = b.dev_row0[j];
    J.nrows = (int32_t)n;
    J.dptr0 = b.dptr0[j];
    J.model = p.model;
    J.lambda2 = p.lambda2;
    J.nu = p.hyper;
    if (n <= 1) { jh.status = 1; continue; }       // Observations.hpp:31 comes before everything else
    if (b.host_key[j] != ~0ull) continue;          // fails; its known rows are still swept for an earlier offence
    const int D = K, r0 = b.run0[j];
    int C = 0;
    for (int d = 0; d < D; ++d) C = std::max(C, b.dcond[r0 + d]);
    // the reference takes the number of conditions from the LAST row (data_design_helpers.cpp:23) and the column of a
    // dataset from its first row: a condition column that changes inside the last dataset makes it write outside its
    // design matrix.  Refused here.
    if (b.last_condition[j] != C || C < 1) { jh.status = 19; continue; }
    jh.D = D;
    jh.C = C;
    if (p.model == MODEL_FULL_DIFFERENCE && C <= 1) { jh.status = 14; continue; }
    const int E = C;
    jh.E = E;
    // build_signal_design / build_difference_design (core/data_design_helpers.cpp:10-28, 30-40)
    std::vector<double> X((size_t)D * E, 0.0);
    for (int d = 0; d < D; ++d) X[(size_t)d * E + (b.dcond[r0 + d] - 1)] = 1.0;
    if (p.model == MODEL_FULL_DIFFERENCE) {
      const uint32_t ref = p.ref;
      if ((uint32_t)D < ref || ref == 0 || ref > (uint32_t)C) { jh.status = 15; continue; }
      for (int d = 0; d < D; ++d) {
        double* row = &X[(size_t)d * E];
        for (int c = (int)ref - 1; c >= 1; --c) row[c] = row[c - 1];
        row[0] = 1.0;
      }
    }
    int32_t mn = b.dfirst[r0], mx = b.dlast[r0];
    for (int d = 1; d < D; ++d) { mn = std::min(mn, b.dfirst[r0 + d]); mx = std::max(mx, b.dlast[r0 + d]); }
    h_minpos[j] = mn;
    J.D = D;
    J.E = E;
    J.des0 = (int32_t)h_design.size();
    h_design.insert(h_design.end(), X.begin(), X.end());
    J.off0 = noff;
    jh.off0 = noff;
    noff += D;
    if (fast) {
      // words of this job's bitmap, a whole number of prefix blocks (positions out of order can give mx < mn: the
      // job fails validation on the device; one block keeps the geometry sane until then)
      const int64_t span = std::max<int64_t>(0, (int64_t)mx - (int64_t)mn);
      h_bm0[j + 1] = (span / 32 + 1 + BM_BLOCK - 1) / BM_BLOCK * BM_BLOCK;
    } else {
      const double lo = (double)mn, hi = (double)mx;
      const double genome_sz = (hi - lo + 1);
      const uint32_t nbins = (uint32_t)(genome_sz * p.bf_per_kb / 1000.);  // binned_genome.cpp:22-23
      if (!(genome_sz * p.bf_per_kb / 1000. >= 2.0) || nbins < 2) { jh.status = 18; noff -= D; h_design.resize(J.des0); continue; }
      J.lower = lo;
      J.size = hi - lo;
      J.nbins = nbins;
      J.P = (int32_t)nbins;
      jh.P = nbins;
    }
    h_sweep[j] = 1;
  }
  // row tiles of every job with known rows (per dataset); jobs that have already failed are only validated
  std::vector<int32_t> h_dtile0(n_ent + 1, 0);
  int n_rt = 0;
  {
    std::vector<TileSpan> spans;
    std::vector<int32_t> nt;
    int32_t at = 0;                          // running tile index
    for (int j = 0; j < nj; ++j) {
      JobDev& J = hj[j];
      const int K = b.n_runs(j);
      J.rtile0 = at;
      for (int d = 0; d < K; ++d) {
        h_dtile0[J.dptr0 + d] = at;
        const int r0 = b.dptr[J.dptr0 + d], r1 = b.dptr[J.dptr0 + d + 1];
        spans.push_back(TileSpan{r0, (int64_t)r1 - r0, j, d, ROW_TILE, 0});
        nt.push_back(grid_tiles(r0, (int64_t)r1 - r0));
        at += nt.back();
      }
      h_dtile0[J.dptr0 + K] = at;
      J.n_rtiles = at - J.rtile0;
    }
    h_dtile0[n_ent] = at;
    n_rt = fill_tiles(eng, spans, nt, cx.rtiles, MakeRowTile(), "k_fill_tiles");
  }
  cx.dtile0.alloc(st, n_ent + 1);          cx.dtile0.upload(h_dtile0);
  cx.ent_job.alloc(st, std::max(1, n_ent)); cx.ent_job.upload(ent_job);
  cx.ent_d.alloc(st, std::max(1, n_ent));   cx.ent_d.upload(ent_d);
  cx.dptr.alloc(st, std::max(1, n_ent));    cx.dptr.upload(b.dptr);
  cx.jobs.alloc(st, nj);                    cx.jobs.upload(hj);   // row0, model ... are final; P / par0 / tab0 follow after the count
  cx.ctl.alloc(st, nj);                     cx.ctl.upload(h_sweep);
  cx.off_part.alloc(st, 2 * (size_t)std::max(1, n_rt));
  cx.row_part.alloc(st, 3 * (size_t)std::max(1, n_rt));
  DevBuf<unsigned long long> d_keys(st, nj);
  ML_CUDA(cudaMemsetAsync(d_keys.p, 0xff, sizeof(unsigned long long) * nj, st));
  const double log2pi = std::log(2 * M_PI * 1.0 * 1.0);

  // ---- K1 (+ K2, FAST): first sweep, bitmap rank -------------------------------------------------------
  DevBuf<uint32_t> d_bitmap, d_bprefix, d_wtile_count;
  DevBuf<int64_t> d_bm0;
  DevBuf<int32_t> d_minpos;
  DevBuf<WTileDesc> d_wt;
  DevBuf<int32_t> d_job_wtile0, d_job_unique;
  std::vector<int32_t> uniq(nj, 0);
  int64_t bitmap_words = 0;
  if (fast) {
    for (int j = 0; j < nj; ++j) h_bm0[j + 1] += h_bm0[j];
    const int64_t nwords = h_bm0[nj];
    bitmap_words = nwords;
    std::vector<int32_t> job_wtile0(nj + 1, 0);
    int n_wt = 0;
    {
      std::vector<TileSpan> spans;
      std::vector<int32_t> nt;
      for (int j = 0; j < nj; ++j) {
        job_wtile0[j] = n_wt;
        const int64_t n = h_bm0[j + 1] - h_bm0[j];
        spans.push_back(TileSpan{h_bm0[j], n, j, 0, WTILE, 0});
        nt.push_back(span_tiles(n, WTILE));
        n_wt += nt.back();
      }
      job_wtile0[nj] = n_wt;
      fill_tiles(eng, spans, nt, d_wt, MakeWTile(), "k_fill_tiles");
    }
    d_bitmap.alloc(st, (size_t)std::max<int64_t>(BM_BLOCK, nwords));
    d_bitmap.zero();
    d_bprefix.alloc(st, (size_t)std::max<int64_t>(1, nwords / BM_BLOCK));
    d_bm0.alloc(st, nj + 1);  d_bm0.upload(h_bm0);
    d_minpos.alloc(st, nj);   d_minpos.upload(h_minpos);
    d_wtile_count.alloc(st, (size_t)std::max(1, n_wt));
    d_job_wtile0.alloc(st, nj + 1); d_job_wtile0.upload(job_wtile0);
    d_job_unique.alloc(st, nj);
    if (n_rt)
      ML_LAUNCH(eng, "k_first_sweep", k_first_sweep<true><<<n_rt, 256, 0, st>>>(
          cx.rtiles.p, cx.jobs.p, cx.ctl.p, cx.dptr.p, b.pos.p, b.nmeth.p, b.cov.p, d_bm0.p, d_minpos.p, d_bitmap.p,
          d_keys.p, log2pi, cx.off_part.p, cx.row_part.p));
      ML_WORK(eng, "k_first_sweep", 12.0 * (double)b.nrows + 4.0 * (double)nwords);
    if (n_wt) {
      ML_LAUNCH(eng, "k_bitmap_tile_count", k_bitmap_tile_count<<<n_wt, 256, 0, st>>>(d_wt.p, d_bitmap.p, d_wtile_count.p));
      ML_WORK(eng, "k_bitmap_tile_count", 4.0 * (double)nwords);
      ML_LAUNCH(eng, "k_bitmap_job_scan", k_bitmap_job_scan<<<cdiv((long long)nj * 32, 128), 128, 0, st>>>(d_job_wtile0.p, nj, d_wtile_count.p,
                                                                       d_job_unique.p));
      ML_LAUNCH(eng, "k_bitmap_block_prefix", k_bitmap_block_prefix<<<n_wt, 256, 0, st>>>(d_wt.p, d_bitmap.p, d_wtile_count.p, d_bprefix.p));
      ML_WORK(eng, "k_bitmap_block_prefix", 4.5 * (double)nwords);
      d_job_unique.download(uniq.data(), nj);
    }
  } else if (n_rt) {
    ML_LAUNCH(eng, "k_first_sweep", k_first_sweep<false><<<n_rt, 256, 0, st>>>(
        cx.rtiles.p, cx.jobs.p, cx.ctl.p, cx.dptr.p, b.pos.p, b.nmeth.p, b.cov.p, nullptr, nullptr, nullptr,
        d_keys.p, log2pi, cx.off_part.p, cx.row_part.p));
    ML_WORK(eng, "k_first_sweep", 12.0 * (double)b.nrows);
  }
  std::vector<unsigned long long> keys(nj);
  d_keys.download(keys.data(), nj);
  stream_sync(st);                                   // the one readback before the parameter arrays can be laid out

  for (int j = 0; j < nj; ++j) {
    JobHost& jh = res->jobs[j];
    if (jh.status == 1) continue;
    const unsigned long long key = std::min(keys[j], b.host_key[j]);
    if (key != ~0ull) { jh.status = decode_status(key); continue; }   // validation comes before the design's own refusals
    if (jh.status) continue;
    if (fast) {
      hj[j].P = uniq[j];
      jh.P = uniq[j];
      if (uniq[j] < 2) jh.status = 18;             // reference: size()-2 underflow
    }
  }
  if (p.lambda2 <= 0)                               // fitter_lasso.ipp:4 (thrown at the first fit)
    for (int j = 0; j < nj; ++j)
      if (!res->jobs[j].status) res->jobs[j].status = 13;

  // ---- layout of parameter arrays -------------------------------------------------------------------
  std::vector<TileSpan> pt_spans;
  std::vector<int32_t> pt_nt;
  int32_t pt_at = 0;
  std::vector<FlsaSeries> series;
  std::vector<int64_t> h_start0(nj, 0);
  int64_t npar = 0, ntab = 0, nstart = 0;
  int max_D = 0;
  for (int j = 0; j < nj; ++j) {
    JobDev& J = hj[j];
    JobHost& jh = res->jobs[j];
    if (jh.status) continue;
    max_D = std::max(max_D, (int)jh.D);
    J.par0 = npar;
    jh.par0 = npar;
    J.tab0 = ntab;
    J.ser0 = (int32_t)series.size();
    J.ptile0 = pt_at;
    h_start0[j] = nstart;
    for (int e = 0; e < J.E; ++e) {
      FlsaSeries fs{npar + (int64_t)e * J.P, J.P, p.lambda2};
      // FULL: at most rows / datasets bins of an estimator hold a position (the even bins are mostly empty)
      if (!fast && J.P > 0) fs.density = std::min(1.0, (double)J.nrows / std::max(1, (int)J.D) / (double)J.P);
      series.push_back(fs);
      pt_spans.push_back(TileSpan{0, (int64_t)J.P, j, e, PAR_TILE, 0});
      pt_nt.push_back(span_tiles(J.P, PAR_TILE));
      pt_at += pt_nt.back();
    }
    J.n_ptiles = pt_at - J.ptile0;
    npar += (int64_t)J.E * J.P;
    ntab += (int64_t)J.D * J.P;
    nstart += J.P;
  }
  res->npar = npar;
  res->noff = noff;
  res->nstart = nstart;
  res->start0 = h_start0;
  cx.jobs.upload(hj);
  const int n_pt = fill_tiles(eng, pt_spans, pt_nt, cx.ptiles, MakeParTile(), "k_fill_tiles");
  cx.design.alloc(st, std::max<size_t>(1, h_design.size()));
  cx.design.upload(h_design);
  cx.start0.alloc(st, nj);
  cx.start0.upload(h_start0);
  cx.rowstart.alloc(st, (size_t)std::max<int64_t>(1, ntab));
  ML_CUDA(cudaMemsetAsync(cx.rowstart.p, 0xff, sizeof(int32_t) * (size_t)std::max<int64_t>(1, ntab), st));
  cx.pidx.alloc(st, (size_t)b.nrows_dev + ROW_ALIGN);
  // mu is part of the reference's return value, but nothing in its R layer reads it (R/call.R, MethyLasso.R): an engine
  // set to want_mu = 0 does not store it when the loop itself never reads it back (one IRLS step, no outer rounds)
  const bool store_mu = eng.want_mu || p.max_iter > 1 || p.nouter > 1;
  res->has_mu = store_mu;
  if (store_mu) res->mu.alloc(st, (size_t)b.nrows_dev + ROW_ALIGN);
  res->beta.alloc(st, (size_t)std::max<int64_t>(1, npar));    res->beta.zero();   // Params: zero (fitter_lasso.hpp:29)
  res->betahat.alloc(st, (size_t)std::max<int64_t>(1, npar)); res->betahat.zero();
  res->pw.alloc(st, (size_t)std::max<int64_t>(1, npar));      res->pw.zero();
  res->start.alloc(st, (size_t)std::max<int64_t>(1, nstart));
  res->offset.alloc(st, std::max(1, noff));                   res->offset.zero();
  cx.old_beta.alloc(st, (size_t)std::max<int64_t>(1, npar));  cx.old_beta.zero();
  cx.beta_raw.alloc(st, (size_t)std::max<int64_t>(1, npar));
  cx.old_off.alloc(st, std::max(1, noff));                    cx.old_off.zero();
  const bool want_outer = p.nouter > 1;
  const bool need_res = p.max_iter > 1 || want_outer;           // mu_res is only read from the second step on
  if (need_res) cx.mu_res.alloc(st, (size_t)b.nrows_dev + ROW_ALIGN);
  if (want_outer) { cx.mu_outer.alloc(st, (size_t)b.nrows_dev + ROW_ALIGN); cx.mu_outer.zero(); }
  cx.tv_part.alloc(st, std::max(1, n_pt));
  cx.scalars.alloc(st, 5 * (size_t)nj);
  cx.scalars_init.alloc(st, 5 * (size_t)nj);
  cx.sums.alloc(st, 2 * std::max<size_t>(1, series.size()));

  // Jobs that failed after their row tiles were built (validation, P < 2, lambda2 <= 0) own no slice of the
  // parameter arrays: the index kernels test alive[job], every later kernel tests ctl[job], which is
  // only ever set for jobs with status 0.
  std::vector<int32_t> h_alive(nj);
  for (int j = 0; j < nj; ++j) h_alive[j] = res->jobs[j].status == 0;
  cx.ctl.upload(h_alive);
  // the pseudodata kernel of the common FAST shapes writes the support itself (one write per parameter instead of one
  // per row); it runs whenever a job takes at least one IRLS step
  const bool start_by_pseudodata = fast && max_D <= PD_FAST_D && p.max_iter > 0;
  if (n_rt) {
    if (fast) {
      ML_LAUNCH(eng, "k_rank_unique", k_rank_unique<<<n_rt, 256, 0, st>>>(cx.rtiles.p, cx.jobs.p, b.pos.p, d_bm0.p, d_minpos.p, d_bitmap.p,
                                          d_bprefix.p, cx.start0.p, cx.ctl.p, cx.pidx.p, cx.rowstart.p,
                                          start_by_pseudodata ? nullptr : res->start.p));
      ML_WORK(eng, "k_rank_unique", 8.0 * (double)b.nrows + 4.0 * (double)ntab + 4.5 * (double)bitmap_words);
    } else {
      ML_LAUNCH(eng, "k_even_bins", k_even_bins<<<n_rt, 256, 0, st>>>(cx.rtiles.p, cx.jobs.p, b.pos.p, cx.ctl.p, cx.pidx.p));
      ML_LAUNCH(eng, "k_rowstart_runs", k_rowstart_runs<<<n_rt, 256, 0, st>>>(cx.rtiles.p, cx.jobs.p, cx.dptr.p, cx.pidx.p, cx.ctl.p,
                                            cx.rowstart.p));
      if (n_pt) {
        ML_LAUNCH(eng, "k_support_even", k_support_even<<<n_pt, PAR_TILE, 0, st>>>(cx.ptiles.p, cx.jobs.p, cx.start0.p, res->start.p));
      }
    }
  }
  std::unique_ptr<FlsaPlan> plan;
  if (!series.empty()) plan = std::make_unique<FlsaPlan>(eng, series, npar);

  // ---- IRLS state machine (core/Posterior.ipp:22-59, 99-110, 204-252) ----------------------------
  // All jobs advance in lock step; a pass = (update | damp) + evaluation + one readback of the per-job scalars.
  // The first pass (initial evaluation, :25) has no decision to take - its sums come from the first sweep and its
  // value is only needed as old_mlogp of the next pass - so it is not read back on its own: the first update pass
  // delivers both.  The FLSA solver's own verdict (a series that outgrew the shared rings, back-pointers out of
  // order: rare) is read back with the scalars as well and, if it asks for it, the pass is redone after the fallback.
  enum { S_INIT_EVAL, S_UPDATE, S_DAMP, S_DONE };
  struct St {
    int state = S_DONE;
    uint32_t it = 0, outer = 0, dstep = 0;
    double old_mlogp = 0, m = 0, dres = 0, dout = 0;
    bool converged = false, first = true, have_m = false, await_init = false;
  };
  std::vector<St> S(nj);
  int alive = 0;
  for (int j = 0; j < nj; ++j)
    if (!res->jobs[j].status) { S[j].state = S_INIT_EVAL; ++alive; }
  const double wnew = 1 - 5 / 10.;
  RowCols rc{b.nmeth.p, b.cov.p, cx.pidx.p, cx.mu_res.p};
  std::vector<int32_t> ctl(nj), skip(series.size());
  std::vector<double> sc(5 * (size_t)nj), sc_init(5 * (size_t)nj);

  auto finish_iterate = [&](int j) {
    St& s = S[j];
    JobHost& jh = res->jobs[j];
    if (p.max_iter == 1) s.converged = true;                       // Posterior.ipp:57
    if (p.nouter <= 1) { s.state = S_DONE; jh.outer_iters = 1; return 0; }
    int copy = 0;
    const uint32_t i = s.outer;
    jh.outer_iters = (int)i + 1;
    if (i > 0) {
      if (s.dout < p.tol_val) { s.state = S_DONE; return 0; }      // :213-217
      s.converged = false;                                         // :219
    }
    copy |= CTL_COPY_OUTER;                                        // mu_old = mu (:222)
    if (!(std::isfinite(p.lambda2) && std::isfinite(p.hyper))) { s.converged = false; s.state = S_DONE; return copy; }
    s.outer = i + 1;
    if (s.outer < p.nouter) {
      // next iterate(): set_converged(false); old_mlogp = -log posterior at the current parameters,
      // which is the value of the last evaluation (same parameters, deterministic kernels)
      s.converged = false;
      s.old_mlogp = s.m;
      s.it = 0;
      s.state = p.max_iter > 0 ? S_UPDATE : S_DONE;
    } else {
      s.state = S_DONE;
    }
    return copy;
  };

  bool sweep_sums = true;        // off_part / row_part still hold the first sweep's sums (first IRLS step, initial evaluation)
  bool init_pending = false;     // the initial evaluation's scalars are on the device, not yet read
  while (alive > 0) {
    bool any_update = false, any_damp = false, any_eval = false, all_init = true, S_any_not_first = false;
    for (int j = 0; j < nj; ++j) {
      int c = 0;
      const St& s = S[j];
      if (s.state == S_UPDATE) { c |= CTL_UPDATE | CTL_EVAL; any_update = true; }
      if (s.state == S_DAMP) { c |= CTL_DAMP | CTL_EVAL; any_damp = true; }
      if (s.state == S_INIT_EVAL) c |= CTL_EVAL | CTL_NO_MU | CTL_ZERO_PARAMS;   // parameters are still all zero here
      if (s.state != S_INIT_EVAL && s.state != S_DONE) all_init = false;
      if (s.state != S_INIT_EVAL && s.it > 0) c |= CTL_DRES;
      if (s.first) c |= CTL_FIRST;
      else if (s.state != S_DONE) S_any_not_first = true;
      if (!store_mu) c |= CTL_NO_MU;
      if (c & CTL_EVAL) any_eval = true;
      ctl[j] = (s.state == S_DONE) ? 0 : c;
    }
    cx.ctl.upload(ctl);
    const bool init_pass = all_init && sweep_sums;     // true in the first pass only
    auto center = [&]() {
      ML_LAUNCH(eng, "k_center", k_center<<<n_pt, PAR_TILE, 0, st>>>(cx.ptiles.p, cx.jobs.p, cx.ctl.p, cx.sums.p, cx.beta_raw.p,
                                            res->beta.p, cx.tv_part.p));
      ML_WORK(eng, "k_center", 16.0 * (double)npar);
    };
    auto evaluate = [&](double* d_scalars) {
      if (!init_pass)
        ML_LAUNCH(eng, "k_eval_rows", (fast ? k_eval_rows<MODEL_FAST> : k_eval_rows<-1>)<<<n_rt, 256, 0, st>>>(
                                          cx.rtiles.p, cx.jobs.p, cx.ctl.p, rc, cx.design.p, res->beta.p,
                                          res->offset.p, log2pi, res->mu.p, cx.mu_outer.p, want_outer ? 1 : 0,
                                          cx.row_part.p));
      if (!init_pass)
        ML_WORK(eng, "k_eval_rows", 12.0 * (double)b.nrows + 8.0 * (double)npar + (store_mu ? 8.0 * (double)b.nrows : 0.0)
                                        + (want_outer ? 8.0 * (double)b.nrows : 0.0) + (S_any_not_first ? 8.0 * (double)b.nrows : 0.0));
      if (any_damp)
        ML_LAUNCH(eng, "k_tv_partial", k_tv_partial<<<n_pt, PAR_TILE, 0, st>>>(cx.ptiles.p, cx.jobs.p, cx.ctl.p, res->beta.p, cx.tv_part.p));
      ML_LAUNCH(eng, "k_job_finalize", k_job_finalize<<<nj, 256, 0, st>>>(cx.jobs.p, nj, cx.ctl.p, cx.row_part.p, cx.tv_part.p,
                                                  d_scalars));
    };
    if (any_update) {
      if (p.model == MODEL_FAST && max_D <= 4)
        ML_LAUNCH(eng, "k_pseudodata", k_pseudodata_fast<4><<<n_pt, PAR_TILE, 0, st>>>(cx.ptiles.p, cx.jobs.p, cx.ctl.p, rc, cx.design.p,
                                                        cx.rowstart.p, res->beta.p, cx.old_beta.p, res->betahat.p,
                                                        res->pw.p, b.pos.p, cx.start0.p, start_by_pseudodata ? res->start.p : nullptr));
      else if (p.model == MODEL_FAST && max_D <= PD_FAST_D)
        ML_LAUNCH(eng, "k_pseudodata", k_pseudodata_fast<PD_FAST_D><<<n_pt, PAR_TILE, 0, st>>>(cx.ptiles.p, cx.jobs.p, cx.ctl.p, rc, cx.design.p,
                                                                cx.rowstart.p, res->beta.p, cx.old_beta.p, res->betahat.p,
                                                                res->pw.p, b.pos.p, cx.start0.p, start_by_pseudodata ? res->start.p : nullptr));
      else
        ML_LAUNCH(eng, "k_pseudodata", k_pseudodata<<<n_pt, PAR_TILE, 0, st>>>(cx.ptiles.p, cx.jobs.p, cx.ctl.p, rc, cx.dptr.p, cx.design.p,
                                                cx.rowstart.p, res->beta.p, cx.old_beta.p, res->betahat.p,
                                                res->pw.p));
      ML_WORK(eng, "k_pseudodata", 4.0 * (double)ntab + 8.0 * (double)b.nrows + 32.0 * (double)npar + (start_by_pseudodata && sweep_sums ? 12.0 * (double)nstart : 0.0)
                                       + (S_any_not_first ? 8.0 * (double)b.nrows : 0.0));
      if (!sweep_sums)     // the first step's sums of W and z W were formed by the first sweep
        ML_LAUNCH(eng, "k_offset_partial", (fast ? k_offset_partial<MODEL_FAST> : k_offset_partial<-1>)<<<n_rt, 256, 0, st>>>(
            cx.rtiles.p, cx.jobs.p, cx.ctl.p, rc, cx.off_part.p));
        ML_WORK(eng, "k_offset_partial", 16.0 * (double)b.nrows);
      ML_LAUNCH(eng, "k_offset_finalize", k_offset_finalize<<<n_ent, 128, 0, st>>>(cx.ent_job.p, cx.ent_d.p, n_ent, cx.jobs.p, cx.ctl.p,
                                                          cx.dtile0.p, cx.off_part.p, res->offset.p,
                                                          cx.old_off.p));
      for (int j = 0; j < nj; ++j)
        if (!res->jobs[j].status)
          for (int e = 0; e < hj[j].E; ++e) skip[hj[j].ser0 + e] = (ctl[j] & CTL_UPDATE) ? 0 : 1;
      plan->solve_async(res->betahat.p, res->pw.p, cx.beta_raw.p, 1, p.lambda1, cx.sums.p, &skip);
      center();
      for (int j = 0; j < nj; ++j)
        if (ctl[j] & CTL_UPDATE) res->jobs[j].irls_steps++;
    }
    if (any_damp) {
      ML_LAUNCH(eng, "k_damp", k_damp<<<n_pt, PAR_TILE, 0, st>>>(cx.ptiles.p, cx.jobs.p, cx.ctl.p, wnew, res->beta.p, cx.old_beta.p,
                                        res->offset.p, cx.old_off.p));
      ML_WORK(eng, "k_damp", 24.0 * (double)npar);
      for (int j = 0; j < nj; ++j)
        if (ctl[j] & CTL_DAMP) res->jobs[j].dampings++;
    }
    if (any_eval) evaluate(init_pass ? cx.scalars_init.p : cx.scalars.p);
    if (init_pass && p.max_iter > 0) {
      // Posterior.ipp:25: old_mlogp of the first step; nothing to decide yet, so no readback of its own
      for (int j = 0; j < nj; ++j) {
        St& s = S[j];
        if (s.state != S_INIT_EVAL) continue;
        s.it = 0;
        s.converged = false;
        s.await_init = true;
        s.state = S_UPDATE;
      }
      init_pending = true;
      continue;                    // sweep_sums stays set: the update pass that follows uses the sweep's offset sums
    }
    auto readback = [&]() {
      if (any_eval) cx.scalars.download(sc.data(), sc.size());
      if (init_pending || init_pass) cx.scalars_init.download(sc_init.data(), sc_init.size());
      stream_sync(st);
    };
    readback();
    if (any_update && plan->solve_finish()) {
      // some series went through the fallback: beta_raw and the centring sums were rewritten
      center();
      evaluate(cx.scalars.p);
      readback();
    }
    if (any_update) sweep_sums = false;
    if (init_pass) sc = sc_init;   // max_iter == 0: the initial evaluation is the only one
    if (init_pending) {
      for (int j = 0; j < nj; ++j)
        if (S[j].await_init) { S[j].old_mlogp = sc_init[5 * (size_t)j]; S[j].await_init = false; }
      init_pending = false;
    }
    // ---- host decisions -------------------------------------------------------------------------
    bool any_copy = false;
    for (int j = 0; j < nj; ++j) {
      St& s = S[j];
      if (s.state == S_DONE || !(ctl[j] & CTL_EVAL)) { ctl[j] = 0; continue; }
      JobHost& jh = res->jobs[j];
      const double m = sc[5 * (size_t)j];
      s.m = m;
      s.dres = sc[5 * (size_t)j + 3];
      s.dout = sc[5 * (size_t)j + 4];
      jh.mlogp = m;
      int copy = 0;
      if (s.state == S_INIT_EVAL) {                                // Posterior.ipp:25 (only reached with max_iter == 0)
        s.old_mlogp = m;
        s.it = 0;
        s.converged = false;
        copy = finish_iterate(j);
      } else {
        if (s.state == S_UPDATE) s.dstep = 0;
        // adapt_all_params (:99-110) with AdaptiveDampingPolicy<1,5,10> (posterior_policies.hpp:57-68)
        const uint32_t ds = s.dstep++;
        if (ds < 10 && m > s.old_mlogp) {
          s.state = S_DAMP;
        } else {
          if (std::isnan(m)) { jh.status = 16; s.state = S_DONE; }   // :36
          else {
            bool conv = false;
            if (s.it > 0 && s.dres < p.tol_val) conv = true;       // :37-45
            if (conv) {
              s.converged = true;
              copy = finish_iterate(j);
            } else {
              // mu_old = mu; old_mlogp = mlogp; refresh residuals (:46-51)
              s.old_mlogp = m;
              s.first = false;
              s.it++;
              const bool more = s.it < p.max_iter;
              // the refreshed residuals are only ever read by a later update: skip the copy when this
              // job can never update again
              if (more || p.nouter > 1) copy |= CTL_COPY_RES;
              if (more) s.state = S_UPDATE;
              else copy |= finish_iterate(j);
            }
          }
        }
      }
      ctl[j] = copy;
      if (copy) any_copy = true;
      if (s.state == S_DONE) {
        --alive;
        jh.converged = s.converged ? 1 : 0;
      }
    }
    if (any_copy) {
      cx.ctl.upload(ctl);
      ML_LAUNCH(eng, "k_copy_rows", k_copy_rows<<<n_rt, 256, 0, st>>>(cx.rtiles.p, cx.jobs.p, cx.ctl.p, res->mu.p, cx.mu_res.p,
                                        want_outer ? cx.mu_outer.p : nullptr));
      ML_WORK(eng, "k_copy_rows", 16.0 * (double)b.nrows);
    }
  }
  for (int j = 0; j < nj; ++j) {
    JobHost& jh = res->jobs[j];
    jh.hyper = (p.model == MODEL_FAST) ? 1.0 : p.hyper;
  }
  if (fast) {                          // the DMR caller joins the raw rows through this table
    res->rowstart = std::move(cx.rowstart);
    res->tab0.assign(nj, 0);
    for (int j = 0; j < nj; ++j) res->tab0[j] = hj[j].tab0;
  }
  res->est_id.alloc(st, (size_t)std::max<int64_t>(1, npar));
  if (n_pt) ML_LAUNCH(eng, "k_fill_estimate_id", k_fill_estimate_id<<<n_pt, PAR_TILE, 0, st>>>(cx.ptiles.p, cx.jobs.p, res->est_id.p));
  ML_WORK(eng, "k_fill_estimate_id", 4.0 * (double)npar);
  res->counts_exact = fast && p.max_iter == 1;
  // no synchronisation here: every host decision has been taken, the scratch buffers are freed in stream
  // order, and whatever reads the result next (copy, segmentation) is enqueued on the same stream
  return res;
}

void result_fast_diff_combine(Engine& eng, Result& r) {
  if (r.diff_combined) return;   // idempotent: a second call must not subtract the reference estimate again
  // rebuild the parameter tiles of the valid jobs (cheap) and run one elementwise kernel
  cudaStream_t st = eng.stream;
  std::vector<JobDev> hj(r.jobs.size());
  std::vector<ParTile> pt;
  for (size_t j = 0; j < r.jobs.size(); ++j) {
    const JobHost& jh = r.jobs[j];
    memset(&hj[j], 0, sizeof(JobDev));
    if (jh.status) continue;
    hj[j].par0 = jh.par0;
    hj[j].P = (int32_t)jh.P;
    hj[j].E = jh.E;
    for (int e = 1; e < jh.E; ++e)
      for (int k = 0; k < jh.P; k += PAR_TILE)
        pt.push_back(ParTile{(int32_t)j, e, k, (int32_t)std::min<int64_t>(PAR_TILE, jh.P - k)});
  }
  if (pt.empty()) return;
  DevBuf<JobDev> d_jobs(st, hj.size());
  d_jobs.upload(hj);
  DevBuf<ParTile> d_pt(st, pt.size());
  d_pt.upload(pt);
  ML_LAUNCH(eng, "k_fast_diff", k_fast_diff<<<(int)pt.size(), PAR_TILE, 0, st>>>(d_pt.p, d_jobs.p, r.beta.p, r.betahat.p, r.pw.p));
  stream_sync(st);
  r.diff_combined = true;
  r.seg_cache.reset();        // tables computed from the estimates before the combine are stale
  r.seg_ref_cache.reset();
  r.pmd_cache.reset();
  r.call_cache.reset();
  r.call_pmd_cache.reset();
}

void detect_init_device_tables(cudaStream_t st) {
  k_init_fast_w<<<cdiv(FAST_W_TABLE, 256), 256, 0, st>>>();
  ML_CHECK_LAUNCH();
  init_rcp_table(st);
  k_init_fast_t<<<cdiv(FAST_T_SIDE * FAST_T_SIDE, 256), 256, 0, st>>>();     // behind the two tables it reads
  ML_CHECK_LAUNCH();
  ML_CUDA(cudaStreamSynchronize(st));
}

}  // namespace ml

// flsa.cuh — host interface of the batched weighted 1-D FLSA solver (kernels in flsa.cu).
#pragma once
#include "common.h"
#include "flsa_core.cuh"

namespace ml {

struct FlsaSeries {
  int64_t off;  // element offset of the series in the concatenated y / w / beta arrays
  int32_t n;
  double lam;
  int64_t off_in = -1;   // >= 0: y / w are read at this offset instead (several series on one input stream)
  double density = 1.0;  // hint: share of the elements with a positive weight (FULL model: non-empty bins)
};

struct TileDesc {
  int32_t series;
  int32_t t0;    // first element of the tile within the series
  int32_t cnt;   // elements in the tile
  int32_t idx;   // tile index within the series
};

constexpr int FLSA_TILE_THREADS = 256;
constexpr int FLSA_TILE_VEC = 8;
constexpr int FLSA_TILE = FLSA_TILE_THREADS * FLSA_TILE_VEC;

// A plan owns everything that depends only on the series lengths (chunk / tile tables, knot-state
// and back-pointer scratch) so the IRLS loop can re-solve the same shapes without re-planning.
class FlsaPlan {
 public:
  FlsaPlan(Engine& eng, const std::vector<FlsaSeries>& series, int64_t total_len);

  // Solve every series: y, w, beta are device arrays of total_len doubles.
  //   post = 0: beta = raw DP solution
  //   post = 1: beta = soft_threshold(clamp(beta, +-35), lambda1)   (FusedLassoGaussianEstimator::get)
  // If d_sums != nullptr (2 doubles per series) it receives, per series, sum(beta*w) and sum(w)
  // over the post-processed beta (centring numerator / denominator, params_helpers.hpp:27),
  // accumulated tile by tile in a fixed order.
  // skip_series (optional, one int per series): non-zero entries are left untouched.
  void solve(const double* d_y, const double* d_w, double* d_beta, int post, double lambda1,
             double* d_sums, const std::vector<int32_t>* skip_series = nullptr);

  // The same in two halves, so that the caller can read its own results back in the same synchronisation:
  // solve_async enqueues everything and stages the solver's flags for the next stream_sync; solve_finish (after that
  // synchronisation) returns true when a fallback path had to rewrite beta / the sums.
  void solve_async(const double* d_y, const double* d_w, double* d_beta, int post, double lambda1,
                   double* d_sums, const std::vector<int32_t>* skip_series = nullptr);
  bool solve_finish();

  // debugging aid: copy the back-pointers of the last solve to the host (total_len doubles each)
  void debug_backpointers(double* h_tm, double* h_tp) const {
    d_tm_.download(h_tm, (size_t)total_len_);
    d_tp_.download(h_tp, (size_t)total_len_);
    stream_sync(eng_.stream);
  }
  int n_series() const { return (int)series_.size(); }
  int n_chunks() const { return n_chunks_; }
  int n_scan_chunks() const { return n_scan_chunks_; }
  int n_scan_groups() const { return n_scan_groups_; }
  int n_tiles() const { return n_tiles_; }
  int chunk_len() const { return chunk_; }
  int warm_len() const { return warm_; }
  // counters of the last solve (read back with the overflow flags)
  struct Stats { int chunks_verified = 0, chunks_redone = 0, longest_chain = 0, chains = 0, series_fallback = 0, scan_chunks = 0, series_rerouted = 0,
                 fine_chains = 0, walker_chains = 0; };
  Stats last_stats() const { return stats_; }

 private:
  Engine& eng_;
  std::vector<FlsaSeries> series_;
  int64_t total_len_;
  int64_t total_steps_ = 0;         // DP steps of all chunks
  int chunk_ = 0, warm_ = 0, n_chunks_ = 0, n_tiles_ = 0, max_tiles_per_series_ = 0;
  DevBuf<SeriesDesc> d_series_;
  DevBuf<ChunkDesc> d_chunks_;
  DevBuf<int32_t> d_win_i32_;       // saved windows of the speculative runs (SpecWindows, flsa_core.cuh)
  DevBuf<double> d_win_f64_;
  DevBuf<int8_t> d_verdict_;        // per chunk: 1 = verified
  DevBuf<int32_t> d_heads_;         // chain heads (chunks the walkers start from)
  SpecWindows win_{};
  // scan route (flsa_core.cuh): chunk table of its own, groups of chunks, the maps of chunks and groups
  int n_scan_chunks_ = 0, n_scan_groups_ = 0, n_scan_series_ = 0;
  int64_t scan_steps_ = 0;
  DevBuf<ChunkDesc> d_scan_chunks_;
  DevBuf<ScanGroup> d_scan_groups_;
  DevBuf<int32_t> d_scan_series_;   // per series with scan tables: series index, first group, number of groups
  DevBuf<SeriesDesc> d_scan_desc_;   // the series as the scan route sees them (first_chunk / n_chunks of ITS chunk table)
  // fine chains (flsa_core.cuh): chain / item tables, the maps of the sub-chunks, the chains left to the walkers
  int fine_cap_items_ = 0, fine_cap_chains_ = 0;
  DevBuf<FineChain> d_fine_chains_;
  DevBuf<FineItem> d_fine_items_;
  DevBuf<int32_t> d_fine_i32_, d_wheads_;
  DevBuf<double> d_fine_f64_;
  ScanFns fine_fns_{};
  DevBuf<int32_t> d_route_;         // per series, per solve: 1 = the scan route's kernels process it (set by the plan for
                                    // scan-route series, by a walker that gave up for ROUTE_SPEC_DUAL ones)
  std::vector<int32_t> h_route_init_;
  std::vector<SeriesDesc> h_scan_desc_;
  DevBuf<int32_t> d_scan_i32_;
  DevBuf<double> d_scan_f64_;
  ScanFns scan_chunk_fns_{}, scan_group_fns_{};
  MsgArray scan_group_start_{};
  DevBuf<TileDesc> d_tiles_;
  DevBuf<int32_t> d_series_tile0_;  // first tile of each series (+1 sentinel)
  DevBuf<double> d_tm_, d_tp_;
  DevBuf<double> d_beta_last_;
  DevBuf<int32_t> d_flags_;         // per series: bit0 overflow, bit1 clamp-order violation
  DevBuf<int32_t> d_counters_;      // [0] chunks verified, [1] chunks redone by the walkers, [2] longest chain, [3] chains,
                                    // [4] series the walkers handed over to the scan route
  DevBuf<Clamp> d_tile_agg_;
  DevBuf<double> d_tile_in_;
  DevBuf<double> d_tile_sums_;      // 2 per tile
  DevBuf<int32_t> d_skip_;
  std::vector<SeriesDesc> h_series_;
  std::vector<int32_t> h_series_tile0_;
  Stats stats_;
  struct Job { const double *y = nullptr, *w = nullptr; double* beta = nullptr; int post = 0; double lambda1 = 0;
               double* sums = nullptr; bool skip = false, pending = false; } job_;   // arguments of the solve in flight
  std::vector<int32_t> h_flags_, h_skip_;
  int32_t h_counters_[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  void backward();
  void forward_scan(const double* d_y, const double* d_w, const int32_t* skip);
  void fallback_sequential(const std::vector<int>& which, const double* d_y, const double* d_w);
};

}  // namespace ml

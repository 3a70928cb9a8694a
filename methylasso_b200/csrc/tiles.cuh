// tiles.cuh — tile tables built on the device.
//
// Every sweep of the engine runs over a table of tiles that never straddle a job / dataset / estimator / series
// boundary (row tiles, parameter tiles, bitmap word tiles, FLSA chunks ...).  For a genome these tables hold 10^4-10^5
// entries each; built on the host and uploaded they cost more than a millisecond per call (6 MB through the staging
// area).  The host now only describes the SPANS (one per job x dataset, job x estimator, series ...: a few hundred
// entries) and a kernel expands them: tile t of span s covers items [i0 + i * tile, ...) with i = t - t0(s).
#pragma once
#include <algorithm>
#include <vector>

#include "hd.h"
#ifdef __CUDACC__
#include "common.h"
#endif

namespace ml {

struct TileSpan {
  int64_t i0;      // first item of the span
  int64_t n;       // items in the span
  int32_t a, b;    // payload (job and dataset / estimator, series ...)
  int32_t tile;    // items per tile
  int32_t t0;      // index of the span's first tile in the table
};

// number of tiles a span of n items needs (at_least_one: a span without items still gets one empty tile)
inline int32_t span_tiles(int64_t n, int32_t tile, bool at_least_one = false) {
  const int64_t t = n > 0 ? (n + tile - 1) / tile : 0;
  return (int32_t)std::max<int64_t>(t, at_least_one ? 1 : 0);
}

// tile t of the table: which span it belongs to (spans sorted by t0, every span owning at least one tile), its index
// inside the span, its first item relative to the span and its item count.  Shared with the CPU emulation
// (tests/emu/emu_tiles.cpp checks it against the host loops it replaced).
ML_HD void tile_of_table(const TileSpan* spans, int n_spans, int t, int& span, int& i, int64_t& off, int32_t& cnt) {
  int lo = 0, hi = n_spans - 1;              // last span with t0 <= t
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (spans[mid].t0 <= t) lo = mid; else hi = mid - 1;
  }
  const TileSpan& s = spans[lo];
  span = lo;
  i = t - s.t0;
  off = (int64_t)i * s.tile;
  const int64_t left = s.n - off;
  cnt = (int32_t)(left < 0 ? 0 : (left < s.tile ? left : s.tile));
}

#ifdef __CUDACC__
template <class Out, class Make>
__global__ void __launch_bounds__(256)
k_fill_tiles(const TileSpan* __restrict__ spans, int n_spans, int n_tiles, Out* __restrict__ out, Make make) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_tiles) return;
  int span, i;
  int64_t off;
  int32_t cnt;
  tile_of_table(spans, n_spans, t, span, i, off, cnt);
  out[t] = make(spans[span], i, off, cnt);
}

#endif  // __CUDACC__

// first tile of every span (written into the spans) and the spans that own tiles, in order; returns the tile count
inline int32_t place_spans(std::vector<TileSpan>& spans, const std::vector<int32_t>& ntiles, std::vector<TileSpan>& live) {
  int32_t total = 0;
  live.clear();
  live.reserve(spans.size());
  for (size_t i = 0; i < spans.size(); ++i) {
    spans[i].t0 = total;
    if (ntiles[i] > 0) live.push_back(spans[i]);     // spans without tiles are left out: their t0 would be ambiguous
    total += ntiles[i];
  }
  return total;
}

#ifdef __CUDACC__
// expand `spans` (t0 filled in here) into `out`; returns the number of tiles
template <class Out, class Make>
int fill_tiles(Engine& eng, std::vector<TileSpan>& spans, const std::vector<int32_t>& ntiles, DevBuf<Out>& out, Make make,
               const char* name) {
  cudaStream_t st = eng.stream;
  std::vector<TileSpan> live;
  const int32_t total = place_spans(spans, ntiles, live);
  out.alloc(st, (size_t)std::max(1, total));
  if (total == 0) return 0;
  DevBuf<TileSpan> d_spans(st, live.size());
  d_spans.upload(live);
  ML_LAUNCH(eng, name, k_fill_tiles<Out, Make><<<cdiv(total, 256), 256, 0, st>>>(d_spans.p, (int)live.size(), total, out.p, make));
  return total;
}
#endif  // __CUDACC__

}  // namespace ml

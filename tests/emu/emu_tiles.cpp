// CPU check of methylasso_b200/csrc/tiles.cuh: the span -> tile arithmetic the device kernel uses, against the plain
// host loops it replaced (row tiles per dataset, FLSA chunks with their "at least one, possibly empty" rule and the
// INT_MAX chunk of the sequential mode).  Test infrastructure only.
#include <climits>
#include <cstdint>
#include <vector>
#include "../../methylasso_b200/csrc/tiles.cuh"

using namespace ml;

extern "C" {

// spans = n_spans x (i0, n, tile, at_least_one); returns 0 when every tile agrees with the loop, else 1 + index
long long emu_tiles_check(int n_spans, const long long* desc) {
  std::vector<TileSpan> spans;
  std::vector<int32_t> nt;
  struct Ref { long long i0; int cnt; int span; int idx; };
  std::vector<Ref> ref;
  for (int s = 0; s < n_spans; ++s) {
    const long long i0 = desc[4 * s], n = desc[4 * s + 1], tile = desc[4 * s + 2];
    const bool one = desc[4 * s + 3] != 0;
    spans.push_back(TileSpan{i0, n, s, 0, (int32_t)tile, 0});
    nt.push_back(span_tiles(n, (int32_t)tile, one));
    // the host loops: for (k = 0; k < n; k += tile) ...; an empty span gets one empty tile when asked to
    int idx = 0;
    long long k = 0;
    if (one) {
      do {
        const long long ke = std::min<long long>(k + tile, n);
        ref.push_back(Ref{i0 + k, (int)(ke - k), s, idx++});
        k = ke;
      } while (k < n);
    } else {
      for (; k < n; k += tile) ref.push_back(Ref{i0 + k, (int)std::min<long long>(tile, n - k), s, idx++});
    }
  }
  std::vector<TileSpan> live;
  const int32_t total = place_spans(spans, nt, live);
  if ((size_t)total != ref.size()) return -1;
  for (int t = 0; t < total; ++t) {
    int span, i;
    int64_t off;
    int32_t cnt;
    tile_of_table(live.data(), (int)live.size(), t, span, i, off, cnt);
    const TileSpan& S = live[span];
    if (S.a != ref[t].span || i != ref[t].idx || S.i0 + off != ref[t].i0 || cnt != ref[t].cnt) return 1 + t;
  }
  return 0;
}
}

// tests/emu/emu_flsa.cpp — TEST INFRASTRUCTURE.  Runs the device logic of flsa_core.cuh on the CPU
// (same inline functions the CUDA kernels call, compiled by g++ with -ffp-contract=off) so the
// chunked / sandwiched DP can be checked against the oracle in the GPU-less container.
#include <vector>
#include <algorithm>
#include "../../methylasso_b200/csrc/flsa_core.cuh"

using namespace ml;

extern "C" int emu_flsa(int n, const double* y, const double* w, double lam, int chunk, int warm,
                        double* beta, double* tm, double* tp, int tile, long long* stats) {
  // stats: [0] chunks, [1] pinned by the first warm-up, [2] started from the left neighbour, [3] finished in C, [4] overflow,
  //        [5] max steps of B prefix, [6] chunks that outgrew the fast ring, [7] clamp order violations
  for (int i = 0; i < 8; ++i) stats[i] = 0;
  if (n == 0) return 0;
  if (n == 1 || lam == 0) { for (int i = 0; i < n; ++i) beta[i] = y[i]; return 0; }
  SeriesDesc S{0, n, 0, 0, (lam < 0 || lam != lam) ? SERIES_SEQUENTIAL : SERIES_NORMAL, lam};
  std::vector<ChunkDesc> cd;
  int steps_lo = 1, steps_hi = n - 1;  // steps [1, n-1)
  if (S.mode == SERIES_SEQUENTIAL) cd.push_back(ChunkDesc{0, steps_lo, steps_hi, 0});
  else {
    for (int k = steps_lo; k < steps_hi || cd.empty(); k += chunk) {
      cd.push_back(ChunkDesc{0, k, std::min(k + chunk, std::max(steps_hi, steps_lo)), 0});
      if (steps_hi <= steps_lo) break;
    }
  }
  S.n_chunks = (int)cd.size();
  std::vector<ChunkState> st(cd.size());
  stats[0] = (long long)cd.size();
  typedef LocalRing<KNOT_CAP_FAST> FastRing;   // same two-tier policy as the kernels
  typedef LocalRing<KNOT_CAP> BigRing;
  for (size_t c = 0; c < cd.size(); ++c) { st[c].status = 0; st[c].count = 0; st[c].pad = 0; }
  auto run_main = [&](bool count_b) {
    std::vector<int> snap(cd.size());
    for (size_t c = 0; c < cd.size(); ++c) snap[c] = st[c].status;   // pad: statuses before this launch
    for (size_t c = 0; c < cd.size(); ++c) {
      if (snap[c] & (CHUNK_DONE | CHUNK_OVERFLOW)) continue;
      const ChunkState* start = nullptr;
      if (snap[c] & CHUNK_START_VALID) start = &st[c];
      else if (cd[c].kb > 1 && (snap[c - 1] & CHUNK_END_VALID) && !(snap[c - 1] & CHUNK_OVERFLOW)) start = &st[c - 1];
      if (!start) continue;
      if (!(snap[c] & CHUNK_START_VALID)) stats[2]++;
      Knots<FastRing> q;
      if (!flsa_main(S, cd[c], y, w, tm, tp, *start, st[c], q)) {
        stats[6]++;
        Knots<BigRing> Q;
        if (!flsa_main(S, cd[c], y, w, tm, tp, *start, st[c], Q)) st[c].status |= CHUNK_OVERFLOW;
      }
    }
  };
  int wlen = warm;
  for (int round = 0; round < 3; ++round) {
    for (size_t c = 0; c < cd.size(); ++c) {
      if (st[c].status & (CHUNK_DONE | CHUNK_OVERFLOW | CHUNK_START_VALID)) continue;
      Knots<FastRing> q, p;
      if (!flsa_warm(S, cd[c], y, w, tm, tp, st[c], wlen, q, p)) {
        stats[6]++;
        Knots<BigRing> Q, P;
        if (!flsa_warm(S, cd[c], y, w, tm, tp, st[c], wlen, Q, P)) st[c].status |= CHUNK_OVERFLOW;
      }
      if (round == 0 && (st[c].status & CHUNK_START_VALID)) stats[1]++;
    }
    run_main(false);
    if (S.mode == SERIES_SEQUENTIAL) break;
    run_main(true);
    wlen = wlen > (1 << 27) ? wlen : wlen * 8;
  }
  // pass C: chains, in order
  bool ovf = false;
  for (size_t c = 0; c < cd.size(); ++c) {
    if (st[c].status & CHUNK_OVERFLOW) { ovf = true; break; }
    if (st[c].status & CHUNK_DONE) continue;
    if (c == 0 || !(st[c - 1].status & CHUNK_END_VALID)) return -1;
    { Knots<BigRing> Q;
      if (!flsa_main(S, cd[c], y, w, tm, tp, st[c - 1], st[c], Q)) st[c].status |= CHUNK_OVERFLOW; }
    if (st[c].status & CHUNK_OVERFLOW) { ovf = true; break; }
    stats[3]++;
  }
  double blast;
  if (ovf) {
    stats[4] = std::max<long long>(stats[4], 1);
    std::vector<double> scratch(3 * (size_t)flsa_sequential_cap(n));
    flsa_sequential_forward(n, y, w, lam, scratch.data(), tm, tp, &blast);
  } else {
    Knots<BigRing> q; knots_load(q, st.back());
    blast = dp_last(q, y[n - 1], w[n - 1], lam);
  }
  // backward: 3-phase reverse clamp scan over tiles of `tile` back-pointers (k = 0..n-2)
  const int m = n - 1;
  const int ntiles = (m + tile - 1) / tile;
  std::vector<Clamp> agg(ntiles);
  for (int t = 0; t < ntiles; ++t) {
    Clamp f = clamp_identity();
    for (int k = std::min(m, (t + 1) * tile) - 1; k >= t * tile; --k) {
      if (tm[k] > tp[k]) stats[7]++;
      f = clamp_then(f, Clamp{tm[k], tp[k]});
    }
    agg[t] = f;
  }
  std::vector<double> incoming(ntiles);
  double v = blast;
  for (int t = ntiles - 1; t >= 0; --t) { incoming[t] = v; v = clamp_apply(agg[t], v); }
  beta[n - 1] = blast;
  for (int t = 0; t < ntiles; ++t) {
    // within a tile: inclusive reverse scan of compositions applied to the tile's incoming value
    Clamp f = clamp_identity();
    for (int k = std::min(m, (t + 1) * tile) - 1; k >= t * tile; --k) {
      f = clamp_then(f, Clamp{tm[k], tp[k]});
      beta[k] = clamp_apply(f, incoming[t]);
    }
  }
  return 0;
}

#include "../../methylasso_b200/csrc/irls_core.cuh"
extern "C" void emu_even_bin(int n, const double* v, double lower, double size, unsigned nbins, unsigned* out) {
  for (int i = 0; i < n; ++i) out[i] = ml::even_bin(v[i], lower, size, nbins);
}
extern "C" double emu_clamp35_soft(double v, double lam1) { return ml::clamp35_soft(v, lam1); }

// tests/emu/emu_flsa.cpp — TEST INFRASTRUCTURE.  Runs the device logic of flsa_core.cuh on the CPU
// (same inline functions the CUDA kernels call, compiled by g++ with -ffp-contract=off) so the
// chunked / sandwiched DP can be checked against the oracle in the GPU-less container.
#include <vector>
#include <algorithm>
#include "../../methylasso_b200/csrc/flsa_core.cuh"

using namespace ml;

static int g_fine = 0;          // 1: chains of the speculative route by composition (fine chains) where they apply
extern "C" void emu_set_fine(int on) { g_fine = on; }

extern "C" int emu_flsa(int n, const double* y, const double* w, double lam, int chunk, int warm,
                        double* beta, double* tm, double* tp, int tile, long long* stats) {
  // stats: [0] chunks, [1] verified (exact by construction or start window met the neighbour's end window),
  //        [2] exact by construction, [3] redone by the walker, [4] series fallback (window > KNOT_CAP),
  //        [5] longest chain, [6] speculative runs that outgrew SPEC_CAP, [7] clamp order violations
  for (int i = 0; i < 8; ++i) stats[i] = 0;
  if (n == 0) return 0;
  if (n == 1 || lam == 0) { for (int i = 0; i < n; ++i) beta[i] = y[i]; return 0; }
  SeriesDesc S{0, n, 0, 0, (lam < 0 || lam != lam) ? SERIES_SEQUENTIAL : SERIES_NORMAL, lam};
  std::vector<ChunkDesc> cd;
  int steps_lo = 1, steps_hi = n - 1;  // steps [1, n-1)
  if (S.mode == SERIES_SEQUENTIAL) cd.push_back(ChunkDesc{0, steps_lo, steps_hi, 0});
  else {
    for (int k = steps_lo; k < steps_hi || cd.empty(); k += chunk) {
      cd.push_back(ChunkDesc{0, k, std::min(k + chunk, std::max(steps_hi, steps_lo)), (int)cd.size()});
      if (steps_hi <= steps_lo) break;
    }
  }
  const int64_t nc = (int64_t)cd.size();
  S.n_chunks = (int)nc;
  stats[0] = nc;
  std::vector<int32_t> status(nc, 0), s_count(nc, 0), e_count(nc, 0);
  std::vector<double> win(6 * (size_t)SPEC_CAP * nc);
  SpecWindows W{status.data(), s_count.data(), e_count.data(), win.data(), win.data() + SPEC_CAP * nc,
                win.data() + 2 * SPEC_CAP * nc, win.data() + 3 * SPEC_CAP * nc, win.data() + 4 * SPEC_CAP * nc,
                win.data() + 5 * SPEC_CAP * nc, nc};
  typedef LocalRing<SPEC_CAP> SpecRing;
  typedef LocalRing<KNOT_CAP> BigRing;
  // sequential series: one chunk from the head, exact by construction (warm covers everything)
  for (int64_t c = 0; c < nc; ++c) {
    KnotsC<SpecRing> q;
    flsa_spec_run(S, cd[c], c, y, w, tm, tp, W, S.mode == SERIES_SEQUENTIAL ? n : warm, q);
    if (status[c] & SPEC_OVERFLOW) stats[6]++;
    if (status[c] & SPEC_EXACT) stats[2]++;
    if (status[c] & SPEC_IRREGULAR) stats[7]++;
  }
  // verdicts from the immutable windows, then the chains in order
  std::vector<char> verified(nc);
  for (int64_t c = 0; c < nc; ++c) { verified[c] = chunk_verified(W, c, 0); stats[1] += verified[c]; }
  bool ovf = false;
  KnotsC<BigRing> Q;
  bool have = false;     // Q holds the exact window after chunk c - 1
  double fine_last = 0.0;
  bool have_fine_last = false;
  long long run = 0;
  for (int64_t c = 0; c < nc && !ovf; ++c) {
    if (verified[c]) { have = false; run = 0; continue; }
    if (g_fine && !have && c > 0) {
      // ---- the chain that starts here, by composition: bounds of its sub-chunks, start messages, redo (flsa.cu's
      // flsa_fine_* kernels, one loop iteration per device thread); on failure the walker below takes it
      int64_t len = 0;
      while (c + len < nc && !verified[c + len]) ++len;
      const int n_items = (int)len * FINE_SUB;
      std::vector<int32_t> nA(n_items, -7), nB(n_items, -7);
      std::vector<double> vA(3 * (size_t)PW_CAP * n_items), vB(3 * (size_t)PW_CAP * n_items), sw(n_items), swy(n_items);
      ScanFns F{MsgArray{nA.data(), vA.data(), vA.data() + (size_t)PW_CAP * n_items, vA.data() + 2 * (size_t)PW_CAP * n_items, n_items},
                MsgArray{nB.data(), vB.data(), vB.data() + (size_t)PW_CAP * n_items, vB.data() + 2 * (size_t)PW_CAP * n_items, n_items},
                sw.data(), swy.data()};
      typedef LocalRing<PW_CAP> RunRing;
      for (int i = 0; i < n_items; ++i)
        for (int sign = -1; sign <= 1; sign += 2) {
          int kb, ke;
          fine_span(cd[c + i / FINE_SUB], i % FINE_SUB, kb, ke);
          KnotsC<RunRing> q;
          double unused = 0;
          scan_bounds(S, ChunkDesc{0, kb, ke, 1}, i, sign, y, w, tm, tp, &unused, F, q);
        }
      PwMsg m0, m1, A, B;
      bool ok = fine_chain_starts(SerialTeam(), W, c, n_items, 0, F, lam, &m0, &m1, A, B);
      double bl = 0;
      bool got_last = false;
      for (int i = 0; i < n_items && ok; ++i) {
        int kb, ke;
        fine_span(cd[c + i / FINE_SUB], i % FINE_SUB, kb, ke);
        const bool last = (c + i / FINE_SUB == nc - 1) && (i % FINE_SUB == FINE_SUB - 1);
        KnotsC<RunRing> q;
        ok = fine_redo(S, kb, ke, i, last, y, w, tm, tp, F.A, q, &bl);
        if (ok && last) got_last = true;
      }
      if (ok) {
        stats[3] += len;
        stats[5] = std::max(stats[5], (long long)len);
        c += len - 1;
        if (got_last) { fine_last = bl; have_fine_last = true; }
        have = false;
        continue;
      }
    }
    if (!have) {
      if (c == 0) {        // the head chunk itself overflowed: exact start from element 0
        double t1, t2;
        knots_init_first(Q, y[0], w[0], lam, t1, t2);
        tm[0] = t1; tp[0] = t2;
      } else {
        if (!verified[c - 1]) return -1;   // cannot happen: a walker always comes from a verified chunk
        window_load_end(Q, W, c - 1);
      }
    }
    bool irr = false;
    if (!flsa_redo_chunk(S, cd[c], y, w, tm, tp, Q, &irr)) { ovf = true; break; }
    if (irr) stats[7]++;
    have = true;
    stats[3]++;
    stats[5] = std::max(stats[5], ++run);
  }
  double blast;
  if (ovf) {
    stats[4] = 1;
    std::vector<double> scratch(3 * (size_t)flsa_sequential_cap(n));
    flsa_sequential_forward(n, y, w, lam, scratch.data(), tm, tp, &blast);
  } else if (have_fine_last) {
    blast = fine_last;
  } else {
    if (!have) window_load_end(Q, W, nc - 1);     // the last chunk was verified: its saved end window is exact
    blast = dp_last(Q, y[n - 1], w[n - 1], lam);
  }
  // backward: 3-phase reverse clamp scan over tiles of `tile` back-pointers (k = 0..n-2)
  const int m = n - 1;
  const int ntiles = (m + tile - 1) / tile;
  std::vector<Clamp> agg(ntiles);
  for (int t = 0; t < ntiles; ++t) {
    Clamp f = clamp_identity();
    for (int k = std::min(m, (t + 1) * tile) - 1; k >= t * tile; --k) {
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

// The scan route (flsa_core.cuh, "a chunk as a map on messages"): the five steps exactly as the kernels of flsa.cu run
// them, one loop iteration per device thread.  stats: [0] chunks, [1] groups, [4] failed (left to the start-to-end run),
// [5] largest message saved.
extern "C" int emu_flsa_scan(int n, const double* y, const double* w, double lam, int chunk, int group,
                             double* beta, double* tm, double* tp, int tile, long long* stats) {
  for (int i = 0; i < 8; ++i) stats[i] = 0;
  if (n == 0) return 0;
  if (n == 1 || lam == 0) { for (int i = 0; i < n; ++i) beta[i] = y[i]; return 0; }
  if (!(lam > 0)) return -2;
  SeriesDesc S{0, n, 0, 0, SERIES_NORMAL, lam};
  std::vector<ChunkDesc> cd;
  const int steps_lo = 1, steps_hi = n - 1;
  for (int k = steps_lo; k < steps_hi || cd.empty(); k += chunk) {
    cd.push_back(ChunkDesc{0, k, std::min(k + chunk, std::max(steps_hi, steps_lo)), (int)cd.size()});
    if (steps_hi <= steps_lo) break;
  }
  const int64_t nc = (int64_t)cd.size();
  S.n_chunks = (int)nc;
  const int64_t ng = (nc + group - 1) / group;
  stats[0] = nc;
  stats[1] = ng;
  auto make = [](std::vector<int32_t>& cnt, std::vector<double>& v, int64_t m) {
    cnt.assign(m, -7);
    v.assign(3 * (size_t)PW_CAP * m, 0.0);
    return MsgArray{cnt.data(), v.data(), v.data() + PW_CAP * m, v.data() + 2 * PW_CAP * m, m};
  };
  std::vector<int32_t> nA, nB, nGA, nGB, nGT;
  std::vector<double> vA, vB, vGA, vGB, vGT, sw(nc), swy(nc), gsw(ng), gswy(ng);
  ScanFns F{make(nA, vA, nc), make(nB, vB, nc), sw.data(), swy.data()};
  ScanFns G{make(nGA, vGA, ng), make(nGB, vGB, ng), gsw.data(), gswy.data()};
  MsgArray GT = make(nGT, vGT, ng);
  typedef LocalRing<PW_CAP> RunRing;
  bool fail = false;
  double blast = 0.0;
  // 1: bounds (the first chunk: exact from the first element)
  for (int64_t c = 0; c < nc; ++c)
    for (int sign = -1; sign <= 1; sign += 2) {
      KnotsC<RunRing> q;
      scan_bounds(S, cd[c], c, sign, y, w, tm, tp, &blast, F, q);
    }
  // 2: group maps
  PwMsg m0, m1, A, B;
  for (int64_t g = 0; g < ng; ++g)
    for (int sign = -1; sign <= 1; sign += 2) {
      const MsgArray& M = sign < 0 ? G.A : G.B;
      if (g == ng - 1 || (g == 0 && sign > 0)) { M.n[g] = 0; if (sign < 0) { G.sw[g] = 0; G.swy[g] = 0; } continue; }
      PwMsg* res = nullptr;
      double Sw = 0, Swy = 0;
      const int64_t c0 = g * group;
      const int cnt = (int)std::min<int64_t>(group, nc - c0);
      if (!scan_reduce_side(SerialTeam(), c0, cnt, sign, F, lam, &m0, &m1, A, B, &res, &Sw, &Swy)) { M.n[g] = -1; continue; }
      if (res->n > 0) pw_save(*res, M, g); else M.n[g] = 0;
      if (sign < 0) { G.sw[g] = Sw; G.swy[g] = Swy; }
    }
  // 3: carry
  fail = !scan_carry(SerialTeam(), lam, 0, (int)ng, G, GT, &m0, &m1, A, B);
  // 4: starts
  for (int64_t g = 0; g < ng && !fail; ++g) {
    const int64_t c0 = g * group;
    fail = !scan_starts(SerialTeam(), g, c0, (int)std::min<int64_t>(group, nc - c0), g == 0, F, GT, lam, &m0, &m1, A, B);
  }
  // 5: redo
  for (int64_t c = 0; c < nc && !fail; ++c) {
    KnotsC<RunRing> q;
    stats[5] = std::max<long long>(stats[5], F.A.n[c]);
    double bl = 0.0;
    fail = !scan_redo(S, cd[c], c, y, w, tm, tp, F.A, q, &bl);
    if (c == nc - 1 && c > 0) blast = bl;
  }
  if (fail) {
    stats[4] = 1;
    std::vector<double> scratch(3 * (size_t)flsa_sequential_cap(n));
    flsa_sequential_forward(n, y, w, lam, scratch.data(), tm, tp, &blast);
  }
  flsa_sequential_backward(n, tm, tp, blast, beta);
  (void)tile;
  return 0;
}

#include "../../methylasso_b200/csrc/irls_core.cuh"
extern "C" void emu_even_bin(int n, const double* v, double lower, double size, unsigned nbins, unsigned* out) {
  for (int i = 0; i < n; ++i) out[i] = ml::even_bin(v[i], lower, size, nbins);
}
extern "C" double emu_clamp35_soft(double v, double lam1) { return ml::clamp35_soft(v, lam1); }

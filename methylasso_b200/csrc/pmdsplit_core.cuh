// pmdsplit_core.cuh — per-element logic of R/call.R:131-164 (segment_methylation_singlechr): the valleys replace the
// call-table rows they overlap (:131-140), the new table is cut into "isolate" / "complement" regions (:143-156) and
// the PMD candidate segments of :93-95 (std_call) are split along them (:158-164).  Shared by the CUDA kernels
// (pmdsplit.inc.cuh) and the CPU emulation (tests/emu/emu_valleys.cpp), like valley_core.cuh.
//
// Row order replaces every sort: call-table rows, valleys, regions and PMD segments are all ordered by (job,
// condition, start) with ends increasing too (disjoint, increasing position ranges), a valley takes the place of the
// rows it covers, and the pieces of a PMD segment come out in region order — which is the (estimate_id, start, end)
// order the reference's setkey() calls produce (:136, :156, :163).
#pragma once
#include "valley_core.cuh"

namespace ml {

enum : int8_t { CAT_NA = 0, CAT_LMR = 1, CAT_UMR = 2, CAT_UMR_DMV = 3 };

struct MergedTable {      // seg_call after :140; one entry per surviving call-table row or valley
  int32_t n;
  int32_t *job, *est, *src_row, *src_valley, *segment_id;   // src_row / src_valley: -1 when the other one applies
  uint32_t *start, *end;
  int8_t *category, *follows_large_gap;
};

// :132-136  position of surviving call-table row i: the surviving rows before it plus the valleys that start before it
ML_HD int32_t ms_row_slot(int64_t i, int32_t kept_before, int32_t n_v, const int32_t* v_row0) {
  int32_t lo = 0, hi = n_v;
  while (lo < hi) { const int32_t mid = (lo + hi) >> 1; if (v_row0[mid] < i) lo = mid + 1; else hi = mid; }
  return kept_before + lo;
}
ML_HD void ms_put_row(const ValleyRows& c, const MergedTable& M, int32_t slot, int64_t i) {
  M.job[slot] = c.job[i]; M.est[slot] = c.est[i]; M.src_row[slot] = (int32_t)i; M.src_valley[slot] = -1;
  M.start[slot] = c.start[i]; M.end[slot] = c.end[i]; M.category[slot] = c.category[i];
}
ML_HD void ms_put_valley(const ValleyOut& v, const MergedTable& M, int32_t slot, int32_t k) {
  M.job[slot] = v.job[k]; M.est[slot] = v.est[k]; M.src_row[slot] = -1; M.src_valley[slot] = k;
  M.start[slot] = v.start[k]; M.end[slot] = v.end[k]; M.category[slot] = CAT_UMR_DMV;
}
ML_HD bool ms_same_group(const MergedTable& M, int32_t a, int32_t b) { return M.job[a] == M.job[b] && M.est[a] == M.est[b]; }
ML_HD int32_t ms_group_head(const MergedTable& M, int32_t m) { return (m == 0 || !ms_same_group(M, m, m - 1)) ? 1 : 0; }
// :139-140  follows.large.gap on the new table
ML_HD int8_t ms_gap(const MergedTable& M, double pmd_valley_min_width, int32_t m) {
  return (m > 0 && ms_same_group(M, m, m - 1) && (double)M.start[m] - (double)M.end[m - 1] > pmd_valley_min_width) ? 1 : 0;
}
// :143-153  "isolate" (UMRs and valleys; valleys only when split.pmds is off) or "complement"; a region starts where the
// condition starts, the label changes or a large gap precedes
ML_HD bool ms_isolate(const MergedTable& M, int split_pmds, int32_t m) {
  const int8_t cat = M.category[m];
  return split_pmds ? (cat == CAT_UMR || cat == CAT_UMR_DMV) : (cat == CAT_UMR_DMV);
}
ML_HD int32_t ms_region_head(const MergedTable& M, int split_pmds, int32_t m) {
  if (ms_group_head(M, m)) return 1;
  return (ms_isolate(M, split_pmds, m) != ms_isolate(M, split_pmds, m - 1) || M.follows_large_gap[m]) ? 1 : 0;
}

struct Regions {          // umr_complement after :156
  int32_t n;
  const int32_t* head;             // first row of the region in the merged table
  int32_t *job, *est;
  uint32_t *start, *end;
  int8_t* isolate;
};
// :154  start = min(start), end = max(end) over the region's rows
ML_HD void ms_region(const MergedTable& M, int split_pmds, const Regions& G, int32_t r) {
  const int32_t h = G.head[r], t = r + 1 < G.n ? G.head[r + 1] : M.n;
  uint32_t lo = M.start[h], hi = M.end[h];
  for (int32_t m = h + 1; m < t; ++m) { lo = M.start[m] < lo ? M.start[m] : lo; hi = M.end[m] > hi ? M.end[m] : hi; }
  G.job[r] = M.job[h]; G.est[r] = M.est[h]; G.start[r] = lo; G.end[r] = hi;
  G.isolate[r] = ms_isolate(M, split_pmds, h) ? 1 : 0;
}

struct PmdSegs {          // std_call (:95): the segments of the fused std values
  int32_t n;
  const int32_t *job, *est;
  const uint32_t *start, *end;
  const double *fused_std, *pw;
};
// :158  first region that can overlap segment x: the first with (job, est, end) >= (x.job, x.est, x.start)
ML_HD int32_t ms_first_region(const Regions& G, const PmdSegs& X, int32_t x) {
  const int32_t j = X.job[x], e = X.est[x];
  const uint32_t xs = X.start[x];
  int32_t lo = 0, hi = G.n;
  while (lo < hi) {
    const int32_t mid = (lo + hi) >> 1;
    const bool less = G.job[mid] < j || (G.job[mid] == j && (G.est[mid] < e || (G.est[mid] == e && G.end[mid] < xs)));
    if (less) lo = mid + 1; else hi = mid;
  }
  return lo;
}
// the regions of a group are ordered by start: those that overlap segment x are [first, the first one that starts behind
// its end) - a second binary search instead of a walk (a PMD-sized segment overlaps thousands of regions)
ML_HD int32_t ms_count_pieces(const Regions& G, const PmdSegs& X, int32_t x) {
  const int32_t first = ms_first_region(G, X, x);
  const int32_t j = X.job[x], e = X.est[x];
  const uint32_t xe = X.end[x];
  int32_t lo = first, hi = G.n;
  while (lo < hi) {
    const int32_t mid = (lo + hi) >> 1;
    const bool in = G.job[mid] == j && G.est[mid] == e && G.start[mid] <= xe;
    if (in) lo = mid + 1; else hi = mid;
  }
  return lo - first;
}
// A segment can be left without any region: it starts on a call-table row, and a row that merely TOUCHES a valley (its
// last CpG two bases before the valley's first: end = max(pos) + 2 = the valley's start) is removed at :133 although
// the valley does not span it.  foverlaps keeps such a segment as a row of NAs, setkey (:163) sorts it first, :164
// numbers it and :168 drops it: it emits no piece but shifts the segment ids of its estimate_id.
ML_HD void ms_group_unmatched(const PmdSegs& X, const int32_t* cnt, int32_t x, int32_t* na_group) {
  if (x > 0 && X.job[x] == X.job[x - 1] && X.est[x] == X.est[x - 1]) return;     // one walk per group, from its head
  int32_t t = x, na = 0;
  for (; t < X.n && X.job[t] == X.job[x] && X.est[t] == X.est[x]; ++t) na += cnt[t] == 0 ? 1 : 0;
  for (int32_t k = x; k < t; ++k) na_group[k] = na;
}
struct Pieces {           // split_call after :168
  int32_t *job, *est, *na;         // na: the number of region-less segments of the piece's estimate_id
  uint32_t *start, *end;
  double *fused_std, *pw;
};
// :159-161  one piece per (segment, region) match: the intersection; fused_std = 0 outside "complement"
// piece i of segment x (its overlap with region first + i), written at `at`
ML_HD void ms_emit_piece(const Regions& G, const PmdSegs& X, int32_t x, int32_t y, int32_t at, int32_t na, const Pieces& P) {
  P.job[at] = X.job[x]; P.est[at] = X.est[x]; P.na[at] = na;
  P.start[at] = G.start[y] > X.start[x] ? G.start[y] : X.start[x];
  P.end[at] = G.end[y] < X.end[x] ? G.end[y] : X.end[x];
  P.fused_std[at] = G.isolate[y] ? 0.0 : X.fused_std[x];
  P.pw[at] = X.pw[x];
}
// the `count` pieces of segment x from position `at` on, every `stride`-th one starting with piece `i0` (the lanes of a
// warp share a segment; 0, 1: all of them in order)
ML_HD void ms_emit_pieces(const Regions& G, const PmdSegs& X, int32_t x, int32_t at, int32_t na, const Pieces& P,
                          int32_t count, int32_t i0 = 0, int32_t stride = 1) {
  const int32_t first = ms_first_region(G, X, x);
  for (int32_t i = i0; i < count; i += stride) ms_emit_piece(G, X, x, first + i, at + i, na, P);
}

}  // namespace ml

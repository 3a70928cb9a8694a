// pmdstatus_core.cuh — per-element logic of R/call.R:201-215 (segment_methylation_singlechr): seg_call joined with the
// PMD table (foverlaps, :203), the duplicate rows removed (:209), LMRs inside PMDs dropped (:210) and the PMD status of
// the UMRs taken from their neighbours (:211-212).  Shared by the CUDA kernels (pmdstatus.inc.cuh) and the CPU
// emulation (tests/emu/emu_valleys.cpp).  Status codes: 1 "PMD", 0 "other", -1 NA.
#pragma once
#include "pmdsplit_core.cuh"

namespace ml {

struct PmdRegions {       // pmd_call after :198, ordered by (job, condition, start); ends increase too
  int32_t n;
  const int32_t *job, *est;
  const uint32_t *start, *end;
  const int8_t* category;          // 1 PMD, 0 other
};
// first region that can overlap merged row m: the first with (job, est, end) >= (job, est, start) of the row
ML_HD int32_t st_first_region(const MergedTable& M, const PmdRegions& G, int32_t m) {
  const int32_t j = M.job[m], e = M.est[m];
  const uint32_t xs = M.start[m];
  int32_t lo = 0, hi = G.n;
  while (lo < hi) {
    const int32_t mid = (lo + hi) >> 1;
    const bool less = G.job[mid] < j || (G.job[mid] == j && (G.est[mid] < e || (G.est[mid] == e && G.end[mid] < xs)));
    if (less) lo = mid + 1; else hi = mid;
  }
  return lo;
}
// :203-209  one row per overlapping region, then unique(): rows of one segment differ in PMD.status only, so what is
// left is the distinct statuses in order of first appearance — one or two of them; no overlap: a single NA
ML_HD int32_t st_statuses(const MergedTable& M, const PmdRegions& G, int32_t m, int8_t& s0, int8_t& s1) {
  int32_t y = st_first_region(M, G, m), k = 0;
  s0 = -1; s1 = -1;
  for (; y < G.n && G.job[y] == M.job[m] && G.est[y] == M.est[m] && G.start[y] <= M.end[m]; ++y) {
    const int8_t c = G.category[y];
    if (k == 0) { s0 = c; k = 1; }
    else if (k == 1 && c != s0) { s1 = c; k = 2; }
  }
  return k == 0 ? 1 : k;
}
struct FinalRows {        // seg_call after :212 (before the drop of :214-218)
  int32_t* merged;                 // row of the merged table (seg_call after :140) it annotates
  int8_t *status, *category, *adm; // adm: is.admissible of the call-table row (0 / 1), -1 (NA) on a valley
};
// :210  LMRs inside PMDs lose their category
ML_HD void st_emit(const MergedTable& M, const PmdRegions& G, const int8_t* row_adm, int32_t m, int32_t at, const FinalRows& F) {
  int8_t s[2];
  const int32_t k = st_statuses(M, G, m, s[0], s[1]);
  for (int32_t q = 0; q < k; ++q) {
    F.merged[at + q] = m;
    F.status[at + q] = s[q];
    F.category[at + q] = (s[q] == 1 && M.category[m] == CAT_LMR) ? (int8_t)CAT_NA : M.category[m];
    F.adm[at + q] = M.src_row[m] >= 0 ? row_adm[M.src_row[m]] : (int8_t)-1;
  }
}
// :211-212  among the admissible rows of a condition (list = their indices in F, in order): an UMR takes the status its
// two neighbours share, "other" when they differ, NA when one of them is missing or NA; other rows keep theirs
ML_HD int8_t st_umr_status(const MergedTable& M, const FinalRows& F, const int32_t* list, int32_t n_list, int32_t r) {
  const int32_t i = list[r];
  if (F.category[i] != CAT_UMR) return F.status[i];
  const int32_t m = F.merged[i];
  int8_t prev = -1, next = -1;
  if (r > 0 && ms_same_group(M, F.merged[list[r - 1]], m)) prev = F.status[list[r - 1]];
  if (r + 1 < n_list && ms_same_group(M, F.merged[list[r + 1]], m)) next = F.status[list[r + 1]];
  if (prev < 0 || next < 0) return -1;
  return prev == next ? prev : (int8_t)0;
}

}  // namespace ml

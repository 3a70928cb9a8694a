// valley_core.cuh — per-element logic of the valley (DMV) block of segment_methylation_singlechr, R/call.R:100-133,
// shared by the CUDA kernels (valleys.inc.cuh, one thread per element) and the CPU emulation (tests/emu/emu_valleys.cpp,
// plain loops): every step is a function of one index over struct-of-arrays columns, separated by stream compactions.
//
// The table is the call table of R/call.R:43-72 over all jobs (chromosomes), rows ordered by (job, condition, start,
// end) — the reference's key (R/call.R:73) within one chromosome; a GROUP is one condition of one job.  Parts are
// disjoint, increasing position ranges (R/call.R:55-60), so start and end both increase within a group; the interval
// searches below rely on that.
//
// R semantics kept: sums of doubles and means are R's long double ones (x87.cuh), `is.high == F` rows only get the
// low.id columns (NaN = NA elsewhere), groups in first-appearance order = row order, foverlaps(type = "any") is
// x.start <= y.end && x.end >= y.start within a condition.
#pragma once
#include "hd.h"
#include "x87.cuh"

namespace ml {

struct ValleyParams {     // arguments of segment_methylation_singlechr (R/call.R:31-35) used by :100-133, its defaults
  double valley_max_beta = 0.1, pmd_valley_min_width = 5000, max_distance = 500;
  int32_t min_num_cpgs = 4, pad = 0;
};

enum : int8_t { VCAT_LOW = 0, VCAT_DMV = 1, VCAT_UMR = 2 };

struct ValleyRows {       // the call table (inputs) and the per-row columns of :102-106
  int64_t n;
  const int32_t *job, *est, *ncpg;
  const uint32_t *start, *end;
  const double *beta, *coverage, *pw;
  const int8_t* category;          // 0 NA, 1 LMR, 2 UMR (R/call.R:83-86)
  double *low_w, *low_b;           // low.id.width / low.id.beta (:106); NaN on rows the statement does not select
  int32_t* run_of_head;            // run index on the first row of a run, -1 elsewhere
};

struct ValleyRuns {       // one entry per run of consecutive non-high rows of a group without a large gap (:102-110)
  int32_t n;
  const int32_t* head;             // first row
  int32_t *last, *ncpg;
  double *cov, *pw;
  int8_t *cat, *keep;              // DMV / low (:110); keep = low.id.beta <= valley_max_beta (:107)
};

ML_HD bool vl_same_group(const ValleyRows& c, int64_t a, int64_t b) { return c.job[a] == c.job[b] && c.est[a] == c.est[b]; }
ML_HD bool vl_is_high(const ValleyRows& c, const ValleyParams& p, int64_t i) { return c.beta[i] > p.valley_max_beta; }   // :102
// :103-104  follows.large.gap (FALSE on the first row of a condition)
ML_HD bool vl_gap(const ValleyRows& c, const ValleyParams& p, int64_t i) {
  return i > 0 && vl_same_group(c, i, i - 1) && (double)c.start[i] - (double)c.end[i - 1] > p.pmd_valley_min_width;
}
// :105-106  low.id = cumsum(is.high) + cumsum(follows.large.gap) restricted to is.high == F: a new group of rows starts
// at a non-high row that opens its condition, follows a high row, or follows a large gap
ML_HD int32_t vl_run_head(const ValleyRows& c, const ValleyParams& p, int64_t i) {
  if (vl_is_high(c, p, i)) return 0;
  if (i == 0 || !vl_same_group(c, i, i - 1)) return 1;
  return (vl_is_high(c, p, i - 1) || vl_gap(c, p, i)) ? 1 : 0;
}
// :106-110  statistics of run r
ML_HD void vl_run_stats(const ValleyRows& c, const ValleyParams& p, const ValleyRuns& R, int32_t r) {
  const int64_t h = R.head[r];
  int64_t k = h + 1;
  while (k < c.n && vl_same_group(c, k, h) && !vl_is_high(c, p, k) && !vl_gap(c, p, k)) ++k;
  const int m = (int)(k - h);
  double lo = (double)c.start[h], hi = (double)c.end[h], cov = 0.0;
  int64_t ncpg = 0;
  X87 pw{0.0, 0.0};
  for (int64_t q = h; q < k; ++q) {
    lo = fmin(lo, (double)c.start[q]);
    hi = fmax(hi, (double)c.end[q]);
    cov += c.coverage[q];                       // integer-valued: exact
    ncpg += c.ncpg[q];
    pw = x87_add_d(pw, c.pw[q]);                // sum() of doubles: long double accumulator (summary.c rsum)
  }
  const double w = hi - lo + 1;
  const double b = x87_mean(m, [&](int q) { return c.beta[h + q]; });
  for (int64_t q = h; q < k; ++q) { c.low_w[q] = w; c.low_b[q] = b; }
  c.run_of_head[h] = r;
  R.last[r] = (int32_t)(k - 1);
  R.ncpg[r] = (int32_t)ncpg;
  R.cov[r] = cov;
  R.pw[r] = x87_to_double(pw);
  R.cat[r] = w > p.pmd_valley_min_width ? VCAT_DMV : VCAT_LOW;
  R.keep[r] = b <= p.valley_max_beta ? 1 : 0;
}
// :113-116  is row i an UMR that overlaps none of the kept runs of its condition?
ML_HD int32_t vl_umr_added(const ValleyRows& c, const ValleyRuns& R, int64_t i) {
  if (c.category[i] != 2) return 0;
  const double xs = (double)c.start[i], xe = (double)c.end[i];
  int32_t lo = 0, hi = R.n;                     // first run whose head row is beyond i
  while (lo < hi) { const int32_t mid = (lo + hi) >> 1; if (R.head[mid] <= i) lo = mid + 1; else hi = mid; }
  for (int32_t u = lo - 1; u >= 0; --u) {       // runs at or before the row: ends decrease going back
    const int64_t h = R.head[u];
    if (!vl_same_group(c, h, i)) break;
    const double ve = (double)c.end[R.last[u]];
    if (ve < xs) break;
    if (R.keep[u] && (double)c.start[h] <= xe) return 0;
  }
  for (int32_t u = lo; u < R.n; ++u) {          // runs after the row: starts increase
    const int64_t h = R.head[u];
    if (!vl_same_group(c, h, i)) break;
    if ((double)c.start[h] > xe) break;
    if (R.keep[u] && (double)c.end[R.last[u]] >= xs) return 0;
  }
  return 1;
}

// ---- items: the rows of `new_valleys` before the merge (:107-116), one per kept run or added UMR, in key order --------
struct ValleyItems {
  int32_t n;
  const int32_t* row;              // head row of the run / the UMR row
};
ML_HD double vl_item_end(const ValleyRows& c, const ValleyRuns& R, int32_t row) {
  const int32_t u = c.run_of_head[row];
  return (u >= 0 && R.keep[u]) ? (double)c.end[R.last[u]] : (double)c.end[row];
}
// :119-121  a merged valley starts at an item that opens its condition or lies beyond pmd_valley_min_width
ML_HD int32_t vl_item_head(const ValleyRows& c, const ValleyParams& p, const ValleyRuns& R, const ValleyItems& I, int32_t r) {
  if (r == 0) return 1;
  const int32_t a = I.row[r - 1], b = I.row[r];
  if (!vl_same_group(c, a, b)) return 1;
  return (double)c.start[b] - vl_item_end(c, R, a) > p.pmd_valley_min_width ? 1 : 0;
}

struct ValleyOut {        // merged valleys (:122-124); `keep` = contains.valley after :125
  int32_t *job, *est, *ncpg;
  uint32_t *start, *end;
  double *width, *beta, *cov, *pw;
  int8_t* keep;
  int32_t* row0;                   // call-table row of the valley's first item (its start is the valley's start)
};
// :122-125  merged valley v = items [I.row index heads[v], heads[v + 1])
ML_HD void vl_merge(const ValleyRows& c, const ValleyParams& p, const ValleyRuns& R, const ValleyItems& I,
                    const int32_t* heads, int32_t n_heads, int32_t v, const ValleyOut& o) {
  const int32_t r0 = heads[v], r1 = v + 1 < n_heads ? heads[v + 1] : I.n;
  double lo = 0, hi = 0, cov = 0.0;
  int64_t ncpg = 0;
  X87 pw{0.0, 0.0};
  bool dmv = false;
  for (int32_t r = r0; r < r1; ++r) {
    const int32_t row = I.row[r], u = c.run_of_head[row];
    const bool run = u >= 0 && R.keep[u];
    const double s = (double)c.start[row], e = vl_item_end(c, R, row);
    lo = r == r0 ? s : fmin(lo, s);
    hi = r == r0 ? e : fmax(hi, e);
    cov += run ? R.cov[u] : c.coverage[row];
    ncpg += run ? R.ncpg[u] : c.ncpg[row];
    pw = x87_add_d(pw, run ? R.pw[u] : c.pw[row]);
    dmv = dmv || (run && R.cat[u] == VCAT_DMV);
  }
  // beta of an item: low.id.beta of its row (a kept run's, or the UMR row's own: NA when that row is high, :114-115)
  const double b = x87_mean(r1 - r0, [&](int q) { return c.low_b[I.row[r0 + q]]; });
  const double w = hi - lo + 1, m = (double)ncpg;
  const int32_t row0 = I.row[r0];
  o.row0[v] = row0;
  o.job[v] = c.job[row0];
  o.est[v] = c.est[row0];
  o.start[v] = (uint32_t)lo;
  o.end[v] = (uint32_t)hi;
  o.width[v] = w;
  o.beta[v] = b;
  o.cov[v] = cov;
  o.pw[v] = x87_to_double(pw);
  o.ncpg[v] = (int32_t)ncpg;
  o.keep[v] = (dmv && w > p.pmd_valley_min_width && w / (m - 1) < p.max_distance && m >= (double)p.min_num_cpgs) ? 1 : 0;
}

// :132-133  does call-table row i overlap one of the final valleys (sorted by job, condition, start; ends increase)?
ML_HD int8_t vl_row_in_valley(const ValleyRows& c, int64_t i, int32_t n_v, const int32_t* vjob, const int32_t* vest,
                              const uint32_t* vstart, const uint32_t* vend) {
  const int32_t j = c.job[i], e = c.est[i];
  const uint32_t xs = c.start[i], xe = c.end[i];
  int32_t lo = 0, hi = n_v;                     // first valley with (job, est, end) >= (j, e, xs)
  while (lo < hi) {
    const int32_t mid = (lo + hi) >> 1;
    const bool less = vjob[mid] < j || (vjob[mid] == j && (vest[mid] < e || (vest[mid] == e && vend[mid] < xs)));
    if (less) lo = mid + 1; else hi = mid;
  }
  return (lo < n_v && vjob[lo] == j && vest[lo] == e && vstart[lo] <= xe) ? 1 : 0;
}

}  // namespace ml

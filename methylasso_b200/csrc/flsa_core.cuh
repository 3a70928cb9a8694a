// flsa_core.cuh — weighted 1-D fused-lasso signal approximator, device logic.
//
// Solves  min_b  1/2 sum_i w_i (y_i - b_i)^2 + lam * sum_i |b_{i+1} - b_i|  exactly with the
// message-passing dynamic programme the reference uses (Johnson's DP;
// /root/reference/src/core/gfl/gfl_tf.c:184-321, reached through
// core/GFLLibrary1DDirectImpl.cpp:32-46).  The arithmetic of one DP step (which products, sums
// and divisions, in which order, with no FMA contraction) is the reference's, so that a series
// processed start-to-end reproduces the reference bit for bit.  What is new here is how the
// work is laid out for a GPU:
//
//  * A series is cut into CHUNKS of consecutive DP steps; one thread owns one chunk and keeps
//    the knot window (the piecewise-linear message) in a small ring.
//  * A chunk that does not start at the head of its series does not know the incoming message.
//    The message is a non-decreasing function confined to [-lam, +lam] and the DP update is
//    monotone in it, so the thread runs TWO copies of the DP through a warm-up stretch, one from
//    the message "-lam everywhere" and one from "+lam everywhere".  The true state is sandwiched
//    between them; when their knot windows become bit-identical the state is pinned and the
//    thread continues with one copy, now exact.  (Measured on WGBS-like series: the windows
//    coalesce after a few hundred steps at lam=25 and the coalesced window was bit-identical to
//    the sequential run's window in every case — scripts/experiments/coalesce.py.)
//  * Chunks whose warm-up did not coalesce before their first step are finished by pass B from
//    the (exact) end state their left neighbour saved, chains of such chunks by pass C.
//  * The back-substitution beta_k = clamp(beta_{k+1}; tm_k, tp_k) is a composition of clamps,
//    which is again a clamp: a reverse parallel scan with min/max only (exact).
#pragma once
#include "hd.h"

namespace ml {

constexpr int KNOT_CAP = 64;  // ring capacity per knot window (power of two)
static_assert((KNOT_CAP & (KNOT_CAP - 1)) == 0, "KNOT_CAP must be a power of two");

enum : int32_t {
  SERIES_NORMAL = 0,
  SERIES_TRIVIAL = 1,     // n <= 1 or lam == 0: beta = y            (gfl_tf.c:210-215)
  SERIES_SEQUENTIAL = 2,  // lam < 0 / NaN: the sandwich does not apply; one chunk, run start-to-end
};
enum : int32_t {
  CHUNK_END_VALID = 1,   // end-of-chunk knot window saved and exact
  CHUNK_DONE = 2,        // every tm/tp of the chunk written
  CHUNK_OVERFLOW = 4,    // knot window outgrew KNOT_CAP: series goes to the global-scratch kernel
  CHUNK_IRREGULAR = 8,   // some tm_k > tp_k (rounding on absurd inputs): clamps do not compose, the
                         // series' back-substitution is redone element by element
};

struct SeriesDesc {
  int64_t off;          // offset of element 0 in the y / w / tm / tp / beta arrays
  int32_t n;            // elements
  int32_t first_chunk;  // index of this series' first chunk
  int32_t n_chunks;     // >= 1 for SERIES_NORMAL
  int32_t mode;
  double lam;
};

struct ChunkDesc {
  int32_t series;
  int32_t kb, ke;  // DP steps [kb, ke) of the series; step k folds element k, 1 <= k <= n-2
  int32_t pad;
};

struct ChunkState {
  int32_t status;
  int32_t count;          // knots saved
  int32_t resolved_from;  // first step of the chunk whose tm/tp pass A wrote (== ke: none)
  int32_t pad;
  double x[KNOT_CAP], a[KNOT_CAP], b[KNOT_CAP];
};

// ---- knot window -------------------------------------------------------------------------
struct Knots {
  double x[KNOT_CAP], a[KNOT_CAP], b[KNOT_CAP];
  int l, r;      // logical indices of the leftmost / rightmost live knot (ring: & (CAP-1))
  uint64_t h;    // xor of knot_hash over live knots: cheap inequality test between two windows
};

ML_HD uint64_t knot_hash(double x, double a, double b) {
  uint64_t m = f64_bits(x) * 0x9E3779B97F4A7C15ull;
  m ^= m >> 29;
  m += f64_bits(a) * 0xBF58476D1CE4E5B9ull;
  m ^= m >> 32;
  m += f64_bits(b) * 0x94D049BB133111EBull;
  m ^= m >> 31;
  return m;
}

// first element of a series: gfl_tf.c:231-240
ML_HD void knots_init_first(Knots& q, double y0, double w0, double lam) {
  q.l = 0;
  q.r = 1;
  q.x[0] = -lam / w0 + y0;
  q.a[0] = w0;
  q.b[0] = -w0 * y0 + lam;
  q.x[1] = lam / w0 + y0;
  q.a[1] = -w0;
  q.b[1] = w0 * y0 + lam;
  q.h = knot_hash(q.x[0], q.a[0], q.b[0]) ^ knot_hash(q.x[1], q.a[1], q.b[1]);
}

// incoming message identically -lam (sign < 0) or +lam (sign > 0): one zero-slope knot at
// +inf / -inf carrying the 2*lam offset between the left-scan and right-scan conventions.
ML_HD void knots_init_bound(Knots& q, int sign, double lam) {
  q.l = 0;
  q.r = 0;
  q.x[0] = sign < 0 ? HUGE_VAL : -HUGE_VAL;
  q.a[0] = 0.0;
  q.b[0] = 2 * lam;
  q.h = knot_hash(q.x[0], q.a[0], q.b[0]);
}

// One DP step: fold element (y, w) into the message (gfl_tf.c:251-289).  Returns false when the
// window would outgrow the ring.  tm/tp are the back-pointers of this step.
ML_HD bool dp_step(Knots& q, double y, double w, double lam, double& tm, double& tp) {
  constexpr int M = KNOT_CAP - 1;
  double alo = w, blo = -w * y - lam;      // :286-287 (seeds of the incoming element)
  double ahi = -w, bhi = w * y - lam;      // :288-289
  int lo = q.l, hi = q.r;
  uint64_t h = q.h;
  while (lo <= hi) {                       // :253-258 (hi == q.r here)
    const int i = lo & M;
    const double xi = q.x[i];
    if (alo * xi + blo > -lam) break;
    const double ai = q.a[i], bi = q.b[i];
    alo += ai;
    blo += bi;
    h ^= knot_hash(xi, ai, bi);
    ++lo;
  }
  while (hi >= lo) {                       // :264-269
    const int i = hi & M;
    const double xi = q.x[i];
    if (-ahi * xi - bhi < lam) break;
    const double ai = q.a[i], bi = q.b[i];
    ahi += ai;
    bhi += bi;
    h ^= knot_hash(xi, ai, bi);
    --hi;
  }
  tm = (-lam - blo) / alo;                 // :272
  tp = (lam + bhi) / (-ahi);               // :277
  const int nl = lo - 1, nr = hi + 1;
  if (nr - nl + 1 > KNOT_CAP) return false;
  const double bl = blo + lam, br = bhi + lam;   // :283, :285
  q.x[nl & M] = tm; q.a[nl & M] = alo; q.b[nl & M] = bl;
  q.x[nr & M] = tp; q.a[nr & M] = ahi; q.b[nr & M] = br;
  q.l = nl;
  q.r = nr;
  q.h = h ^ knot_hash(tm, alo, bl) ^ knot_hash(tp, ahi, br);
  return true;
}

// last coefficient: zero of the final message (gfl_tf.c:294-302)
ML_HD double dp_last(const Knots& q, double y, double w, double lam) {
  constexpr int M = KNOT_CAP - 1;
  double alo = w, blo = -w * y - lam;
  for (int lo = q.l; lo <= q.r; ++lo) {
    const int i = lo & M;
    if (alo * q.x[i] + blo > 0) break;
    alo += q.a[i];
    blo += q.b[i];
  }
  return -blo / alo;
}

ML_HD bool knots_same(const Knots& p, const Knots& q) {
  constexpr int M = KNOT_CAP - 1;
  if (p.h != q.h || p.r - p.l != q.r - q.l) return false;
  const int cnt = p.r - p.l + 1;
  for (int k = 0; k < cnt; ++k) {
    const int i = (p.l + k) & M, j = (q.l + k) & M;
    if (f64_bits(p.x[i]) != f64_bits(q.x[j]) || f64_bits(p.a[i]) != f64_bits(q.a[j]) ||
        f64_bits(p.b[i]) != f64_bits(q.b[j]))
      return false;
  }
  return true;
}

ML_HD void knots_save(const Knots& q, ChunkState& st) {
  constexpr int M = KNOT_CAP - 1;
  const int cnt = q.r - q.l + 1;
  for (int k = 0; k < cnt; ++k) {
    const int i = (q.l + k) & M;
    st.x[k] = q.x[i];
    st.a[k] = q.a[i];
    st.b[k] = q.b[i];
  }
  st.count = cnt;
}

ML_HD void knots_load(Knots& q, const ChunkState& st) {
  const int cnt = st.count;
  uint64_t h = 0;
  for (int k = 0; k < cnt; ++k) {
    q.x[k] = st.x[k];
    q.a[k] = st.a[k];
    q.b[k] = st.b[k];
    h ^= knot_hash(st.x[k], st.a[k], st.b[k]);
  }
  q.l = 0;
  q.r = cnt - 1;
  q.h = h;
}

// ---- pass A: every chunk independently -----------------------------------------------------
// y, w, tm, tp point at element 0 of the series.  `st` must have been initialised with
// status = 0, resolved_from = ke before the first round.  Later rounds call this again with a
// longer warm-up for chunks that are not CHUNK_DONE yet; only the still-missing head
// [kb, resolved_from) of the chunk is then produced.
ML_HD void flsa_pass_a(const SeriesDesc& S, const ChunkDesc& c, const double* __restrict__ y,
                       const double* __restrict__ w, double* __restrict__ tm,
                       double* __restrict__ tp, ChunkState& st, int warm) {
  const double lam = S.lam;
  const int kb = c.kb, ke = c.ke;
  const bool have_end = (st.status & CHUNK_END_VALID) != 0;
  const int stop = have_end ? st.resolved_from : ke;  // tm/tp of [stop, ke) already written
  Knots q;
  double t1, t2;
  bool ok = true;
  int k, resolved;
  if (kb - warm <= 1) {
    // close enough to the head of the series: run from element 0, exact by construction
    knots_init_first(q, y[0], w[0], lam);
    if (kb == 1) {                                       // gfl_tf.c:231-232
      tm[0] = q.x[0];
      tp[0] = q.x[1];
      if (q.x[0] > q.x[1]) st.status |= CHUNK_IRREGULAR;
    }
    for (k = 1; k < kb && ok; ++k) ok = dp_step(q, y[k], w[k], lam, t1, t2);
    resolved = kb;
  } else {
    Knots p;
    knots_init_bound(q, -1, lam);
    knots_init_bound(p, +1, lam);
    bool merged = false;
    bool pristine = true;  // both windows still hold only their +-inf sentinel
    for (k = kb - warm; k < stop && ok; ++k) {
      const double yk = y[k], wk = w[k];
      double u1, u2;
      if (pristine) {
        // A zero-weight element leaves a constant message unchanged (clip(+-lam + 0)); the generic
        // step cannot express that on a sentinel-only window (0 * inf), so skip it exactly.
        if (!(wk > 0)) continue;
        pristine = false;
      }
      ok = dp_step(q, yk, wk, lam, t1, t2);
      ok = dp_step(p, yk, wk, lam, u1, u2) && ok;
      if (ok && knots_same(p, q)) { merged = true; ++k; break; }
    }
    if (!ok) { st.status |= CHUNK_OVERFLOW; return; }
    if (!merged) return;                       // nothing learnt this round
    for (; k < kb && ok; ++k) ok = dp_step(q, y[k], w[k], lam, t1, t2);
    resolved = k;  // >= kb: first step whose back-pointers are exact
  }
  bool irregular = false;
  for (; k < stop && ok; ++k) {
    ok = dp_step(q, y[k], w[k], lam, t1, t2);
    tm[k] = t1;
    tp[k] = t2;
    irregular |= (t1 > t2);
  }
  if (irregular) st.status |= CHUNK_IRREGULAR;
  if (!ok) { st.status |= CHUNK_OVERFLOW; return; }
  if (!have_end) { knots_save(q, st); st.status |= CHUNK_END_VALID; }
  st.resolved_from = resolved;
  if (resolved == kb) st.status |= CHUNK_DONE;
}

// ---- pass B / C: finish the unresolved head of a chunk from its left neighbour's end state ---
// `left` must be CHUNK_END_VALID and stable (written by an earlier launch).
ML_HD void flsa_finish_chunk(const SeriesDesc& S, const ChunkDesc& c, const double* __restrict__ y,
                             const double* __restrict__ w, double* __restrict__ tm,
                             double* __restrict__ tp, const ChunkState& left, ChunkState& st) {
  Knots q;
  knots_load(q, left);
  const double lam = S.lam;
  const bool have_end = (st.status & CHUNK_END_VALID) != 0;
  const int stop = have_end ? st.resolved_from : c.ke;
  double t1, t2;
  bool ok = true;
  bool irregular = false;
  for (int k = c.kb; k < stop && ok; ++k) {
    ok = dp_step(q, y[k], w[k], lam, t1, t2);
    tm[k] = t1;
    tp[k] = t2;
    irregular |= (t1 > t2);
  }
  if (irregular) st.status |= CHUNK_IRREGULAR;
  if (!ok) { st.status |= CHUNK_OVERFLOW; return; }
  if (!have_end) {  // the whole chunk was unresolved: its end state is ours
    knots_save(q, st);
    st.status |= CHUNK_END_VALID;
  }
  st.resolved_from = c.kb;
  st.status |= CHUNK_DONE;
}

// ---- start-to-end DP with the knot window in caller-provided scratch of 2n doubles x 3 -------
// (the reference's own memory layout, gfl_tf.c:233-240); used when a window outgrows KNOT_CAP.
ML_HD void flsa_sequential_forward(int n, const double* __restrict__ y, const double* __restrict__ w,
                                   double lam, double* __restrict__ x, double* __restrict__ a,
                                   double* __restrict__ b, double* __restrict__ tm,
                                   double* __restrict__ tp, double* beta_last) {
  int l = n - 1, r = n;
  tm[0] = -lam / w[0] + y[0];
  tp[0] = lam / w[0] + y[0];
  x[l] = tm[0]; x[r] = tp[0];
  a[l] = w[0];  b[l] = -w[0] * y[0] + lam;
  a[r] = -w[0]; b[r] = w[0] * y[0] + lam;
  for (int k = 1; k < n - 1; ++k) {
    const double wk = w[k], yk = y[k];
    double alo = wk, blo = -wk * yk - lam, ahi = -wk, bhi = wk * yk - lam;
    int lo = l, hi = r;
    while (lo <= r) { if (alo * x[lo] + blo > -lam) break; alo += a[lo]; blo += b[lo]; ++lo; }
    while (hi >= lo) { if (-ahi * x[hi] - bhi < lam) break; ahi += a[hi]; bhi += b[hi]; --hi; }
    tm[k] = (-lam - blo) / alo;
    l = lo - 1; x[l] = tm[k];
    tp[k] = (lam + bhi) / (-ahi);
    r = hi + 1; x[r] = tp[k];
    a[l] = alo; b[l] = blo + lam;
    a[r] = ahi; b[r] = bhi + lam;
  }
  const double wk = w[n - 1], yk = y[n - 1];
  double alo = wk, blo = -wk * yk - lam;
  for (int lo = l; lo <= r; ++lo) { if (alo * x[lo] + blo > 0) break; alo += a[lo]; blo += b[lo]; }
  *beta_last = -blo / alo;
}

// ---- back-substitution as a clamp monoid ----------------------------------------------------
// One back-pointer pair acts on the value to its right as  v -> (v > hi ? hi : (v < lo ? lo : v))
// (gfl_tf.c:308-310).  With lo <= hi for every pair, "apply g after f" is again such a map.
struct Clamp {
  double lo, hi;
};
ML_HD double clamp_apply(const Clamp& c, double v) {
  if (v > c.hi) return c.hi;
  if (v < c.lo) return c.lo;
  return v;
}
ML_HD Clamp clamp_identity() { return Clamp{-HUGE_VAL, HUGE_VAL}; }
// first f (the pair nearer the end of the series), then g
ML_HD Clamp clamp_then(const Clamp& f, const Clamp& g) {
  return Clamp{clamp_apply(g, f.lo), clamp_apply(g, f.hi)};
}

// element-by-element back-substitution (gfl_tf.c:306-311), for CHUNK_IRREGULAR series
ML_HD void flsa_sequential_backward(int n, const double* __restrict__ tm, const double* __restrict__ tp,
                                    double beta_last, double* __restrict__ beta) {
  double v = beta_last;
  beta[n - 1] = v;
  for (int k = n - 2; k >= 0; --k) {
    if (v > tp[k]) v = tp[k];
    else if (v < tm[k]) v = tm[k];
    beta[k] = v;
  }
}

// FusedLassoGaussianEstimator::get: clamp to +-35 then soft-threshold by lambda1
// (core/FusedLassoGaussianEstimator.hpp:71-84, core/Settings.hpp:32, core/util.cpp:9-19); the
// comparisons are std::min/std::max's, so NaN behaves as in the reference.
ML_HD double clamp35_soft(double v, double lam1) {
  const double c = 35.0;
  double m = (-c < v) ? v : -c;
  m = (m < c) ? m : c;
  if (m > 0) {
    const double t = m - lam1;
    return (0. < t) ? t : 0.;
  }
  const double t = m + lam1;
  return (t < 0.) ? t : 0.;
}

}  // namespace ml

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
//  * A chunk whose warm-up did not coalesce takes its start state from the (exact) end state its
//    left neighbour saved, or retries with a warm-up eight times longer in the next round; what
//    is still open after the last round is finished in order by one lane per series.
//  * The back-substitution beta_k = clamp(beta_{k+1}; tm_k, tp_k) is a composition of clamps,
//    which is again a clamp: a reverse parallel scan with min/max only (exact).
#pragma once
#include "hd.h"

namespace ml {

constexpr int KNOT_CAP = 64;       // largest knot window a chunk may hold / save (power of two)
constexpr int KNOT_CAP_FAST = 32;  // shared-memory ring of the fast path; a thread whose window
                                   // outgrows it redoes its chunk with a KNOT_CAP ring in local memory
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
  CHUNK_START_VALID = 16, // the saved window is the exact state just BEFORE the chunk's first step
  CHUNK_CLAIMED = 32,    // a chain walker (flsa_chain_kernel) has taken the chunk in this launch
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
  int32_t warm;         // warm-up steps in front of each chunk of this series (0: the launch's default)
  int32_t pad;
};

struct ChunkDesc {
  int32_t series;
  int32_t kb, ke;  // DP steps [kb, ke) of the series; step k folds element k, 1 <= k <= n-2
  int32_t pad;
};

struct ChunkState {
  int32_t status;
  int32_t count;          // knots saved
  int32_t reserved;
  int32_t pad;            // snapshot of status taken at the start of a launch (neighbour reads)
  double x[KNOT_CAP], a[KNOT_CAP], b[KNOT_CAP];
};

// ---- knot window -------------------------------------------------------------------------
// A ring of knots (x, a, b) addressed by logical index (masked by CAP-1).  Two storages:
//   LocalRing   : arrays inside the thread (registers / local memory) — CPU emulation, and the
//                 in-kernel fallback when a window outgrows the shared-memory ring
//   StridedRing : a column of a shared-memory tile laid out [field][slot][thread]; lanes of a warp
//                 touch distinct banks whatever slot each of them is at (no bank conflicts)
template <int CAP_>
struct LocalRing {
  static constexpr int CAP = CAP_;
  double xs[CAP_], as[CAP_], bs[CAP_];
  ML_HD double x(int i) const { return xs[i & (CAP_ - 1)]; }
  ML_HD double a(int i) const { return as[i & (CAP_ - 1)]; }
  ML_HD double b(int i) const { return bs[i & (CAP_ - 1)]; }
  ML_HD void set(int i, double x_, double a_, double b_) {
    const int k = i & (CAP_ - 1);
    xs[k] = x_; as[k] = a_; bs[k] = b_;
  }
};

template <int CAP_, int STRIDE>
struct StridedRing {
  static constexpr int CAP = CAP_;
  double* base;  // this thread's column
  ML_HD double x(int i) const { return base[(i & (CAP_ - 1)) * STRIDE]; }
  ML_HD double a(int i) const { return base[(CAP_ + (i & (CAP_ - 1))) * STRIDE]; }
  ML_HD double b(int i) const { return base[(2 * CAP_ + (i & (CAP_ - 1))) * STRIDE]; }
  ML_HD void set(int i, double x_, double a_, double b_) {
    const int k = i & (CAP_ - 1);
    base[k * STRIDE] = x_;
    base[(CAP_ + k) * STRIDE] = a_;
    base[(2 * CAP_ + k) * STRIDE] = b_;
  }
};

template <class Ring>
struct Knots {
  Ring ring;
  int l, r;      // logical indices of the leftmost / rightmost live knot
};

// first element of a series: gfl_tf.c:231-240
template <class Ring>
ML_HD void knots_init_first(Knots<Ring>& q, double y0, double w0, double lam, double& tm0, double& tp0) {
  q.l = 0;
  q.r = 1;
  tm0 = -lam / w0 + y0;
  tp0 = lam / w0 + y0;
  const double bl = -w0 * y0 + lam, br = w0 * y0 + lam;
  q.ring.set(0, tm0, w0, bl);
  q.ring.set(1, tp0, -w0, br);
}

// incoming message identically -lam (sign < 0) or +lam (sign > 0): one zero-slope knot at
// +inf / -inf carrying the 2*lam offset between the left-scan and right-scan conventions.
template <class Ring>
ML_HD void knots_init_bound(Knots<Ring>& q, int sign, double lam) {
  q.l = 0;
  q.r = 0;
  const double x0 = sign < 0 ? HUGE_VAL : -HUGE_VAL;
  q.ring.set(0, x0, 0.0, 2 * lam);
}

// The knot quotient num / a.  A zero numerator is common on methylation data (an unmethylated position has
// y == 0, so b == -lam exactly and num == -lam - b == +0) and sends the device's IEEE division down its slow
// path (a ~150-instruction subroutine, taken by a fifth of all warp-steps of the first main pass).  The quotient
// of a zero by a non-zero, non-NaN divisor is a zero whose sign is the product of the signs: same bits, no call.
ML_HD double knot_div(double num, double a) {
#ifdef __CUDA_ARCH__
  if (num == 0.0 && a != 0.0 && a == a)
    return __longlong_as_double((__double_as_longlong(num) ^ __double_as_longlong(a)) & (long long)0x8000000000000000ull);
#endif
  return num / a;
}

// ---- one DP step, side-symmetric ------------------------------------------------------------
// The reference scans the window from the left (gfl_tf.c:251-258) and from the right (:262-269)
// with mirrored formulas.  Because IEEE rounding is sign-symmetric the right scan is the SAME
// computation as the left one on negated accumulators:
//     -ahi*x - bhi <  lam   <=>   ahi*x + bhi > -lam        (fl(-p - q) = -fl(p + q))
//     (lam + bhi) / (-ahi)   ==   (-lam - bhi) / ahi         (bit for bit)
// so both sides are:  seeds (a, b);  while knot in range and !(a*x + b > -lam): a += a_k, b += b_k;
// new knot x = (-lam - b) / a, stored as (x, a, b + lam).  Sides differ only in their seeds
// ((w, -w*y - lam) left, (-w, w*y - lam) right) and in the direction they walk.  That is what lets
// two lanes of a warp take one side each without diverging (flsa.cu).
//
// The right scan of the reference is bounded by where the left scan stopped (hi >= lo).  Run
// independently it is bounded by the window only; the results coincide unless it walked past
// lo - 1 (the window was exhausted from both sides), in which case it is redone with the bound.
template <class Ring>
ML_HD int side_scan(const Ring& ring, int idx, int dir, int last, double& a, double& b, double lam) {
  while ((last - idx) * dir >= 0) {            // idx still inside [.., last] in walking direction
    const double x = ring.x(idx);
    if (a * x + b > -lam) break;
    a += ring.a(idx);
    b += ring.b(idx);
    idx += dir;
  }
  return idx;
}

// One DP step: fold element (y, w) into the message (gfl_tf.c:251-289).  Returns false when the
// window would outgrow the ring (the ring is then left untouched).  tm/tp are the back-pointers.
template <class Ring>
ML_HD bool dp_step(Knots<Ring>& q, double y, double w, double lam, double& tm, double& tp) {
  double al = w, bl = -w * y - lam;        // :286-287 (seeds of the incoming element)
  double ar = -w, br = w * y - lam;        // :288-289
  const int lo = side_scan(q.ring, q.l, +1, q.r, al, bl, lam);
  int hi = side_scan(q.ring, q.r, -1, q.l, ar, br, lam);
  if (hi < lo - 1) {                       // both scans crossed: redo the right one with :264's bound
    ar = -w;
    br = w * y - lam;
    hi = side_scan(q.ring, q.r, -1, lo, ar, br, lam);
  }
  tm = knot_div(-lam - bl, al);            // :272
  tp = knot_div(-lam - br, ar);            // :277, (lam + bhi) / (-ahi) bit for bit
  const int nl = lo - 1, nr = hi + 1;
  if (nr - nl + 1 > Ring::CAP) return false;
  q.ring.set(nl, tm, al, bl + lam);        // :282-285
  q.ring.set(nr, tp, ar, br + lam);
  q.l = nl;
  q.r = nr;
  return true;
}

// last coefficient: zero of the final message (gfl_tf.c:294-302)
template <class Ring>
ML_HD double dp_last(const Knots<Ring>& q, double y, double w, double lam) {
  double alo = w, blo = -w * y - lam;
  for (int lo = q.l; lo <= q.r; ++lo) {
    if (alo * q.ring.x(lo) + blo > 0) break;
    alo += q.ring.a(lo);
    blo += q.ring.b(lo);
  }
  return -blo / alo;
}

template <class Ring>
ML_HD bool knots_same(const Knots<Ring>& p, const Knots<Ring>& q) {
  if (p.r - p.l != q.r - q.l) return false;
  const int cnt = p.r - p.l + 1;
  for (int k = 0; k < cnt; ++k) {
    if (f64_bits(p.ring.x(p.l + k)) != f64_bits(q.ring.x(q.l + k)) ||
        f64_bits(p.ring.a(p.l + k)) != f64_bits(q.ring.a(q.l + k)) ||
        f64_bits(p.ring.b(p.l + k)) != f64_bits(q.ring.b(q.l + k)))
      return false;
  }
  return true;
}

template <class Ring>
ML_HD void knots_save(const Knots<Ring>& q, ChunkState& st) {
  const int cnt = q.r - q.l + 1;   // <= KNOT_CAP: callers only instantiate rings up to KNOT_CAP
  for (int k = 0; k < cnt; ++k) {
    st.x[k] = q.ring.x(q.l + k);
    st.a[k] = q.ring.a(q.l + k);
    st.b[k] = q.ring.b(q.l + k);
  }
  st.count = cnt;
}

// false if the saved window does not fit this ring
template <class Ring>
ML_HD bool knots_load(Knots<Ring>& q, const ChunkState& st) {
  const int cnt = st.count;
  if (cnt > Ring::CAP) return false;
  for (int k = 0; k < cnt; ++k) q.ring.set(k, st.x[k], st.a[k], st.b[k]);
  q.l = 0;
  q.r = cnt - 1;
  return true;
}

// ---- warm-up: pin the state at the start of a chunk ------------------------------------------
// Runs the two sandwich copies over [kb - warm, kb) (or the exact DP from element 0 when the chunk
// is that close to the head of its series).  If the two windows become bit-identical the window
// is carried to kb and saved in `st` with CHUNK_START_VALID.  The loop has the same trip count
// for every thread of a warp (divergence-free apart from the data-dependent knot pops).
// y, w, tm, tp point at element 0 of the series.  Returns false on ring overflow (st untouched).
template <class Ring>
ML_HD bool flsa_warm(const SeriesDesc& S, const ChunkDesc& c, const double* __restrict__ y,
                     const double* __restrict__ w, double* __restrict__ tm, double* __restrict__ tp,
                     ChunkState& st, int warm, Knots<Ring>& q, Knots<Ring>& p) {
  const double lam = S.lam;
  const int kb = c.kb;
  double t1, t2;
  bool ok = true, dual, pristine;
  int k;
  if (kb - warm <= 1) {
    // exact by construction: first element (gfl_tf.c:231-240), then steps 1 .. kb-1
    knots_init_first(q, y[0], w[0], lam, t1, t2);
    if (kb == 1) {
      tm[0] = t1;
      tp[0] = t2;
      if (t1 > t2) st.status |= CHUNK_IRREGULAR;
    }
    dual = false;
    pristine = false;
    k = 1;
  } else {
    knots_init_bound(q, -1, lam);
    knots_init_bound(p, +1, lam);
    dual = true;
    pristine = true;   // both windows still hold only their +-inf sentinel
    k = kb - warm;
  }
  double yn = 0.0, wn = 0.0;   // element k, loaded one step ahead of its use (k + 1 <= n - 1 always)
  if (k < kb) { yn = y[k]; wn = w[k]; }
  for (; k < kb && ok; ++k) {
    const double yk = yn, wk = wn;
    yn = y[k + 1];
    wn = w[k + 1];
    if (pristine) {
      // A zero-weight element leaves a constant message unchanged (clip(+-lam + 0)); the generic
      // step cannot express that on a sentinel-only window (0 * inf), so skip it exactly.
      if (!(wk > 0)) continue;
      pristine = false;
    }
    if (dual) {
      ok = dp_step(q, yk, wk, lam, t1, t2);
      ok = dp_step(p, yk, wk, lam, t1, t2) && ok;
      // the windows are compared knot by knot every 8th step (and at the last one): pinning is
      // noticed at most 7 steps late, which costs nothing but those 7 duplicated steps
      if (ok && ((k & 7) == 7 || k == kb - 1) && knots_same(p, q)) dual = false;
    } else {
      ok = dp_step(q, yk, wk, lam, t1, t2);
    }
  }
  if (!ok) return false;
  if (dual) return true;   // not pinned within this warm-up: a later round looks further back
  knots_save(q, st);
  st.status |= CHUNK_START_VALID;
  return true;
}

// ---- main: the chunk's own steps from a pinned start state ------------------------------------
// `start` is the chunk's own CHUNK_START_VALID window or its left neighbour's CHUNK_END_VALID
// window (the same state by definition).  Writes tm/tp of [kb, ke), saves the end state.
// Returns false on ring overflow (st untouched).
template <class Ring>
ML_HD bool flsa_main(const SeriesDesc& S, const ChunkDesc& c, const double* __restrict__ y,
                     const double* __restrict__ w, double* __restrict__ tm, double* __restrict__ tp,
                     const ChunkState& start, ChunkState& st, Knots<Ring>& q) {
  if (!knots_load(q, start)) return false;
  const double lam = S.lam;
  double t1, t2;
  bool ok = true, irregular = false;
  double yn = 0.0, wn = 0.0;
  if (c.kb < c.ke) { yn = y[c.kb]; wn = w[c.kb]; }
  for (int k = c.kb; k < c.ke && ok; ++k) {
    const double yk = yn, wk = wn;
    yn = y[k + 1];
    wn = w[k + 1];
    ok = dp_step(q, yk, wk, lam, t1, t2);
    tm[k] = t1;
    tp[k] = t2;
    irregular |= (t1 > t2);
  }
  if (!ok) return false;
  if (irregular) st.status |= CHUNK_IRREGULAR;
  knots_save(q, st);
  st.status = (st.status & ~CHUNK_START_VALID) | CHUNK_END_VALID | CHUNK_DONE;
  return true;
}

// ---- start-to-end DP with the knot window in a global-memory ring -----------------------------
// For series whose window outgrows KNOT_CAP (monotone stretches: up to n + 1 knots).  The ring lives in
// caller-provided scratch of 3 * cap doubles, cap a power of two >= n + 2, so it can never overflow; the steps are
// the very dp_step / dp_last of the chunked path.
struct GlobalRing {
  static constexpr int CAP = 0x7fffffff;   // never the limit: the scratch is sized for the series
  double* base;
  int mask;                                // capacity - 1
  ML_HD double x(int i) const { return base[i & mask]; }
  ML_HD double a(int i) const { return base[(int64_t)(mask + 1) + (i & mask)]; }
  ML_HD double b(int i) const { return base[2 * (int64_t)(mask + 1) + (i & mask)]; }
  ML_HD void set(int i, double x_, double a_, double b_) {
    const int k = i & mask;
    base[k] = x_;
    base[(int64_t)(mask + 1) + k] = a_;
    base[2 * (int64_t)(mask + 1) + k] = b_;
  }
};
ML_HD int64_t flsa_sequential_cap(int n) {
  int64_t cap = 4;
  while (cap < (int64_t)n + 2) cap <<= 1;
  return cap;
}
ML_HD void flsa_sequential_forward(int n, const double* __restrict__ y, const double* __restrict__ w, double lam,
                                   double* __restrict__ scratch /* 3 * flsa_sequential_cap(n) */,
                                   double* __restrict__ tm, double* __restrict__ tp, double* beta_last) {
  Knots<GlobalRing> q;
  q.ring.base = scratch;
  q.ring.mask = (int)(flsa_sequential_cap(n) - 1);
  double t1, t2;
  knots_init_first(q, y[0], w[0], lam, t1, t2);
  tm[0] = t1;
  tp[0] = t2;
  for (int k = 1; k < n - 1; ++k) {
    dp_step(q, y[k], w[k], lam, t1, t2);
    tm[k] = t1;
    tp[k] = t2;
  }
  *beta_last = dp_last(q, y[n - 1], w[n - 1], lam);
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

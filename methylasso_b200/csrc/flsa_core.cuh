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
//    monotone in it, so a copy of the DP started from the message "-lam everywhere" stays below
//    the true state and one started from "+lam everywhere" stays above it.  Every chunk is run
//    SPECULATIVELY by one thread: through a warm-up stretch in front of the chunk and then through
//    the chunk itself, from the lower bound if the chunk's index in its series is even and from the
//    upper bound if it is odd.  The thread saves its knot window S just before the chunk's first
//    step and its window E after the chunk's last step.
//  * Chunk c is VERIFIED when S_c is bit-identical to E_{c-1}: the two are copies from opposite
//    bounds, the true state lies between them, so both ARE the true state - everything chunk c
//    computed from S_c on is exact, and so is the end state of chunk c-1.  (Measured on WGBS-like
//    series: windows started 512 steps back coincide for ~94 % of the chunks at lam=25.)  This is
//    the sandwich argument with ONE warm-up copy per chunk: the second copy is the neighbour's run.
//  * Chunks that are not verified form chains; a chain is redone in order from the exact end state
//    of the verified chunk on its left (flsa_chain_kernel), one lane pair per chain.
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
// ROUTE_SPEC_DUAL: speculative route, with the scan route's tables built too: a series whose walkers meet a chain longer
// than the engine's limit is handed over to the scan route on the device (flsa_chain_kernel sets its entry of `route`)
enum : int32_t { ROUTE_SPEC = 0, ROUTE_SCAN = 1, ROUTE_SPEC_DUAL = 2 };
enum : int32_t {
  SPEC_EXACT = 1,      // the run started at the head of its series: exact by construction, nothing to verify
  SPEC_OVERFLOW = 2,   // the window outgrew the speculative ring: S / E are not usable, the chunk is redone by a walker
  SPEC_IRREGULAR = 4,  // some tm_k > tp_k (rounding on absurd inputs): clamps do not compose, the
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
  int32_t route;        // ROUTE_SPEC / ROUTE_SCAN: which forward pass the series takes (first_chunk indexes that route's table)
  int64_t off_in;       // offset of element 0 in the INPUT arrays y / w (== off unless several series share one input: the
                        // lanes of a lambda2 sweep read the same pseudodata)
};

struct ChunkDesc {
  int32_t series;
  int32_t kb, ke;  // DP steps [kb, ke) of the series; step k folds element k, 1 <= k <= n-2
  int32_t idx;     // index of the chunk within its series (its parity chooses the bound the speculation starts from)
};
// scan route: chunks [c0, c0 + nc) of one series (indices into the scan route's chunk table)
struct ScanGroup {
  int32_t series;
  int32_t c0, nc;
  int32_t kind;    // bit 0: the first group of its series, bit 1: the last one (its map is never applied to anything)
};

// The saved windows of all chunks, struct-of-arrays [field][slot][chunk] (consecutive threads = consecutive chunks write
// consecutive addresses).  S = window just before step kb, E = window after step ke - 1.
#ifndef ML_SPEC_CAP
#define ML_SPEC_CAP 16
#endif
constexpr int SPEC_CAP = ML_SPEC_CAP;   // knots a speculative run may hold (ring in shared memory) and save
struct SpecWindows {
  int32_t* status;             // SPEC_* per chunk
  int32_t *s_count, *e_count;  // knots saved (0: none)
  double *sx, *sa, *sb, *ex, *ea, *eb;
  int64_t stride;              // number of chunks
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

// ---- the same step, arranged for a short dependent chain -------------------------------------------------------------
// One thread carries BOTH scans of a step.  Measured on B200 (scripts/flsa_latency.py): the plain dp_step above costs a
// lone thread ~1400 cycles per step - two scan loops one after the other, every comparison behind a shared-memory load
// and its address arithmetic, two divisions behind their own branches.  Here
//   * the knot at each end of the window is also kept in registers (it is the knot the previous step created: the first
//     comparison of each scan needs no load), and the second knot of each side is loaded before it is known to be needed;
//   * the two scans advance in the same basic block, one knot per side per round (independent chains: they overlap);
//   * the two quotients are formed together, the zero-numerator shortcut (knot_div) as selects instead of branches.
// The arithmetic of each side - which products, sums and quotients, in which order - is dp_step's, so the results are
// the same bits (tests/test_emu.py runs this very function on the CPU against the reference DP).
template <class Ring>
struct KnotsC {
  Ring ring;
  int l, r;                    // logical indices of the leftmost / rightmost live knot
  double xl, al, bl;           // == ring[l] whenever l <= r
  double xr, ar, br;           // == ring[r]
};
template <class Ring>
ML_HD void knots_cache(KnotsC<Ring>& q) {
  q.xl = q.ring.x(q.l); q.al = q.ring.a(q.l); q.bl = q.ring.b(q.l);
  q.xr = q.ring.x(q.r); q.ar = q.ring.a(q.r); q.br = q.ring.b(q.r);
}
template <class Ring>
ML_HD void knots_init_first(KnotsC<Ring>& q, double y0, double w0, double lam, double& tm0, double& tp0) {
  q.l = 0;
  q.r = 1;
  tm0 = -lam / w0 + y0;
  tp0 = lam / w0 + y0;
  const double bl = -w0 * y0 + lam, br = w0 * y0 + lam;
  q.ring.set(0, tm0, w0, bl);
  q.ring.set(1, tp0, -w0, br);
  knots_cache(q);
}
template <class Ring>
ML_HD void knots_init_bound(KnotsC<Ring>& q, int sign, double lam) {
  q.l = 0;
  q.r = 0;
  q.ring.set(0, sign < 0 ? HUGE_VAL : -HUGE_VAL, 0.0, 2 * lam);
  knots_cache(q);
}

// num / a with the zero-numerator shortcut of knot_div as selects: the division never sees a zero numerator (so it
// stays on its fast path) and two quotients can be formed side by side
ML_HD double knot_div_sel(double num, double a) {
#ifdef __CUDA_ARCH__
  const bool z = num == 0.0 && a != 0.0 && a == a;
  const double qn = (z ? 1.0 : num) / a;
  const double zero = __longlong_as_double((__double_as_longlong(num) ^ __double_as_longlong(a)) & (long long)0x8000000000000000ull);
  return z ? zero : qn;
#else
  return num / a;
#endif
}

template <class Ring>
ML_HD bool dp_step_c(KnotsC<Ring>& q, double y, double w, double lam, double& tm, double& tp) {
  const double wy = w * y;
  double al = w, bl = -wy - lam;           // gfl_tf.c:286-287 (seeds of the incoming element)
  double ar = -w, br = wy - lam;           // :288-289
  const int l = q.l, r = q.r;
  const double nlam = -lam;
  // second knot of each side, on its way before the first comparisons are known
  double x2l = 0.0, a2l = 0.0, b2l = 0.0, x2r = 0.0, a2r = 0.0, b2r = 0.0;
  if (l < r) {
    x2l = q.ring.x(l + 1); a2l = q.ring.a(l + 1); b2l = q.ring.b(l + 1);
    x2r = q.ring.x(r - 1); a2r = q.ring.a(r - 1); b2r = q.ring.b(r - 1);
  }
  int i = l, j = r;
  // first knots: registers
  bool goL = l <= r && !(al * q.xl + bl > nlam);
  bool goR = l <= r && !(ar * q.xr + br > nlam);
  if (goL) { al += q.al; bl += q.bl; i = l + 1; }
  if (goR) { ar += q.ar; br += q.br; j = r - 1; }
  // second knots: already loaded
  goL = goL && i <= r && !(al * x2l + bl > nlam);
  goR = goR && j >= l && !(ar * x2r + br > nlam);
  if (goL) { al += a2l; bl += b2l; ++i; }
  if (goR) { ar += a2r; br += b2r; --j; }
  // the rest from the ring, one knot per side per round
  while (goL || goR) {
    const bool inL = goL && i <= r, inR = goR && j >= l;
    double xa = 0.0, aa = 0.0, ba = 0.0, xb = 0.0, ab = 0.0, bb = 0.0;
    if (inL) { xa = q.ring.x(i); aa = q.ring.a(i); ba = q.ring.b(i); }
    if (inR) { xb = q.ring.x(j); ab = q.ring.a(j); bb = q.ring.b(j); }
    goL = inL && !(al * xa + bl > nlam);
    goR = inR && !(ar * xb + br > nlam);
    if (goL) { al += aa; bl += ba; ++i; }
    if (goR) { ar += ab; br += bb; --j; }
  }
  const int lo = i;
  int hi = j;
  if (hi < lo - 1) {                       // both scans crossed: redo the right one with gfl_tf.c:264's bound
    ar = -w;
    br = wy - lam;
    hi = side_scan(q.ring, r, -1, lo, ar, br, lam);
  }
  tm = knot_div_sel(nlam - bl, al);        // :272
  tp = knot_div_sel(nlam - br, ar);        // :277, (lam + bhi) / (-ahi) bit for bit
  const int nl = lo - 1, nr = hi + 1;
  if (nr - nl + 1 > Ring::CAP) return false;
  const double bls = bl + lam, brs = br + lam;
  q.ring.set(nl, tm, al, bls);             // :282-285
  q.ring.set(nr, tp, ar, brs);
  q.l = nl;
  q.r = nr;
  q.xl = tm; q.al = al; q.bl = bls;
  q.xr = tp; q.ar = ar; q.br = brs;
  return true;
}
// A zero-weight element (an empty bin of the FULL model's even binning: four bins in five) leaves the message unchanged in
// exact arithmetic; the reference still runs its step, which re-derives the two end knots from themselves.  That step
// depends on the state alone (w y = 0 whatever y), so once it has left the window bit for bit as it found it, it is a
// fixed point: every further zero-weight element in a row returns the same back-pointers (the end knots' positions)
// and changes nothing.  `fixed` carries that knowledge from one element to the next; any other element resets it.
template <class Ring>
ML_HD bool dp_step_z(KnotsC<Ring>& q, double y, double w, double lam, double& tm, double& tp, bool& fixed) {
  if (w == 0.0 && y == y && !(y - y != 0.0)) {          // w y is a zero (y finite)
    if (fixed) { tm = q.xl; tp = q.xr; return true; }
    const int l0 = q.l, r0 = q.r;
    const double xl = q.xl, al = q.al, bl = q.bl, xr = q.xr, ar = q.ar, br = q.br;
    const bool ok = dp_step_c(q, y, w, lam, tm, tp);
    fixed = ok && q.l == l0 && q.r == r0 && f64_bits(q.xl) == f64_bits(xl) && f64_bits(q.al) == f64_bits(al) &&
            f64_bits(q.bl) == f64_bits(bl) && f64_bits(q.xr) == f64_bits(xr) && f64_bits(q.ar) == f64_bits(ar) &&
            f64_bits(q.br) == f64_bits(br);
    return ok;
  }
  fixed = false;
  return dp_step_c(q, y, w, lam, tm, tp);
}
template <class Ring>
ML_HD double dp_last(const KnotsC<Ring>& q, double y, double w, double lam) {
  double alo = w, blo = -w * y - lam;
  for (int lo = q.l; lo <= q.r; ++lo) {
    if (alo * q.ring.x(lo) + blo > 0) break;
    alo += q.ring.a(lo);
    blo += q.ring.b(lo);
  }
  return -blo / alo;
}

// ---- saved windows ------------------------------------------------------------------------------
template <class K>
ML_HD void window_save(const K& q, int32_t* count, double* x, double* a, double* b, int64_t stride, int64_t c) {
  const int cnt = q.r - q.l + 1;
  for (int k = 0; k < cnt; ++k) {
    x[k * stride + c] = q.ring.x(q.l + k);
    a[k * stride + c] = q.ring.a(q.l + k);
    b[k * stride + c] = q.ring.b(q.l + k);
  }
  count[c] = cnt;
}
// the end window of chunk c into a ring (always fits: SPEC_CAP <= every ring's capacity)
template <class Ring>
ML_HD void window_load_end(KnotsC<Ring>& q, const SpecWindows& W, int64_t c) {
  const int cnt = W.e_count[c];
  for (int k = 0; k < cnt; ++k) q.ring.set(k, W.ex[k * W.stride + c], W.ea[k * W.stride + c], W.eb[k * W.stride + c]);
  q.l = 0;
  q.r = cnt - 1;
  knots_cache(q);
}
// S of chunk c bit-identical to E of chunk c - 1 (both saved by speculative runs that did not overflow)
ML_HD bool windows_meet(const SpecWindows& W, int64_t c) {
  if ((W.status[c] | W.status[c - 1]) & SPEC_OVERFLOW) return false;
  const int cnt = W.s_count[c];
  if (cnt <= 0 || cnt != W.e_count[c - 1]) return false;
  for (int k = 0; k < cnt; ++k) {
    const int64_t is = k * W.stride + c, ie = is - 1;
    if (f64_bits(W.sx[is]) != f64_bits(W.ex[ie]) || f64_bits(W.sa[is]) != f64_bits(W.ea[ie]) ||
        f64_bits(W.sb[is]) != f64_bits(W.eb[ie]))
      return false;
  }
  return true;
}
// Verdict on chunk c of a series whose first chunk is c0 (chunks of a series are consecutive): its own outputs are
// exact when its run was exact by construction or its start window met the left neighbour's end window.
ML_HD bool chunk_verified(const SpecWindows& W, int64_t c, int64_t c0) {
  if (W.status[c] & SPEC_OVERFLOW) return false;
  if (W.status[c] & SPEC_EXACT) return true;
  return c > c0 && windows_meet(W, c);
}

// ---- the speculative run of one chunk (scalar statement; flsa_spec_kernel in flsa.cu is the same procedure with the
// loads of y / w issued ahead and a trip count that is uniform across the warp) ---------------------------------------
// Steps [max(kb - warm, 1), ke): from the exact first-element window when the stretch reaches the head of the series
// (SPEC_EXACT), else from the message "-lam everywhere" (even chunk index) or "+lam everywhere" (odd).  Back-pointers
// are written for [kb, ke) only.  y, w, tm, tp point at element 0 of the series.
template <class Ring>
ML_HD void flsa_spec_run(const SeriesDesc& S, const ChunkDesc& cd, int64_t c, const double* __restrict__ y,
                         const double* __restrict__ w, double* __restrict__ tm, double* __restrict__ tp,
                         const SpecWindows& W, int warm, KnotsC<Ring>& q) {
  const double lam = S.lam;
  const int kb = cd.kb;
  double t1, t2;
  int status = 0, k;
  bool pristine;
  if (kb - warm <= 1) {
    knots_init_first(q, y[0], w[0], lam, t1, t2);   // gfl_tf.c:231-240
    if (kb == 1) {
      tm[0] = t1;
      tp[0] = t2;
      if (t1 > t2) status |= SPEC_IRREGULAR;
    }
    status |= SPEC_EXACT;
    pristine = false;
    k = 1;
  } else {
    knots_init_bound(q, (cd.idx & 1) ? +1 : -1, lam);
    pristine = true;   // the window still holds only its +-inf sentinel
    k = kb - warm;
  }
  bool ok = true, zfix = false;
  for (; k < cd.ke && ok; ++k) {
    if (k == kb) window_save(q, W.s_count, W.sx, W.sa, W.sb, W.stride, c);
    if (pristine) {
      // A zero-weight element leaves a constant message unchanged (clip(+-lam + 0)); the generic
      // step cannot express that on a sentinel-only window (0 * inf), so skip it exactly.
      if (!(w[k] > 0)) {
        if (k >= kb) ok = false;   // no back-pointers without a real window: left to the walker (absurd input)
        continue;
      }
      pristine = false;
    }
    ok = dp_step_z(q, y[k], w[k], lam, t1, t2, zfix);
    if (ok && k >= kb) {
      tm[k] = t1;
      tp[k] = t2;
      if (t1 > t2) status |= SPEC_IRREGULAR;
    }
  }
  if (ok && kb >= cd.ke) window_save(q, W.s_count, W.sx, W.sa, W.sb, W.stride, c);   // empty chunk (n == 2)
  if (!ok || pristine) {
    W.status[c] = (status & ~SPEC_EXACT) | SPEC_OVERFLOW;
    W.s_count[c] = 0;
    W.e_count[c] = 0;
    return;
  }
  window_save(q, W.e_count, W.ex, W.ea, W.eb, W.stride, c);
  W.status[c] = status;
}

// ---- the walker's redo of one chunk from an exact window already in its ring --------------------------------
// Returns false on ring overflow.  *irregular is set when some tm_k > tp_k.
template <class Ring>
ML_HD bool flsa_redo_chunk(const SeriesDesc& S, const ChunkDesc& cd, const double* __restrict__ y,
                           const double* __restrict__ w, double* __restrict__ tm, double* __restrict__ tp,
                           KnotsC<Ring>& q, bool* irregular) {
  double t1, t2;
  bool zfix = false;
  double yn = 0.0, wn = 1.0;                       // one element ahead: the load does not wait for the step before it
  if (cd.kb < cd.ke) { yn = y[cd.kb]; wn = w[cd.kb]; }
  for (int k = cd.kb; k < cd.ke; ++k) {
    const double yk = yn, wk = wn;
    if (k + 1 < cd.ke) { yn = y[k + 1]; wn = w[k + 1]; }
    if (!dp_step_z(q, yk, wk, S.lam, t1, t2, zfix)) return false;
    tm[k] = t1;
    tp[k] = t2;
    if (t1 > t2) *irregular = true;
  }
  return true;
}

// ---- a chunk as a map on messages: the SCAN route ------------------------------------------------------------------------
// At every point b the DP step acts on the message value v = delta(b) as v -> clip(v + w (b - y), -lam, +lam), and a
// composition of such clamp-shift maps is one again.  For a whole chunk and an incoming value in [-lam, +lam] that gives
//     delta_out(b) = min( B(b), max( A(b), delta_in(b) + Sw b - Swy ) ),
// where A and B are the chunk's end messages from the incoming messages "-lam everywhere" and "+lam everywhere" (two runs
// of the chunk's own steps from the two bounds, no warm-up) and Sw = sum of w, Swy = sum of w y over the chunk.
// G = delta_in + Sw b - Swy has G - A and G - B non-decreasing in b (A and B have slopes <= Sw), so delta_out is A up to
// the point b1 where G crosses A, G between b1 and the point b2 where G crosses B, and B beyond: in the knot
// representation of the windows (delta(b) = -lam + sum over knots left of b of (a b + b_knot)) that is A's knots below b1,
// one knot at b1 that switches from A's line to G's, delta_in's knots between b1 and b2 (the increments of G are those of
// delta_in), one knot at b2 that switches to B's line, and B's knots above b2 (pw_compose).  The exact end state of a
// chunk therefore follows from the exact state in front of it by a MERGE of three short knot lists instead of a re-run of
// its steps, and the chunk maps themselves compose the same way: (A, B, Sw, Swy) of "chunk 1 then chunk 2" is
// (chunk 2 applied to A1, chunk 2 applied to B1, Sw1 + Sw2, Swy1 + Swy2).  That makes the forward pass a SCAN:
//   1. every chunk, from both bounds, on its own steps                                  (flsa_scan_bounds_kernel)
//   2. the maps of the chunks of a group composed in order, one group per thread pair    (flsa_scan_reduce_kernel)
//   3. the exact message in front of every group, one thread per series                  (flsa_scan_carry_kernel)
//   4. the exact message in front of every chunk of a group, one thread per group        (flsa_scan_starts_kernel)
//   5. every chunk redone from its exact start window: back-pointers, last coefficient   (flsa_scan_redo_kernel)
// The dependent path is 2 chunk lengths of DP steps plus about 2 sqrt(2 chunks) merges whatever the data - where the
// speculative route (warm-up, verification, walkers) degenerates into walking a whole series at one thread's pace when
// the memory of the series is longer than any affordable warm-up (lambda2 = 1000 on the PMD step, the FULL model's bins).
// On count data with an integer penalty every sum above is exact and the crossings are correctly rounded quotients of
// the same exact numbers the sequential run divides: the scan reproduces tf_dp_weight bit for bit there
// (tests/test_emu.py); elsewhere it differs from it in the last bits like any reassociation.
constexpr int PW_CAP = 32;                // knots a message of the scan route may hold
struct PwMsg {
  int n;
  double x[PW_CAP], a[PW_CAP], b[PW_CAP];
};
// false when the window holds more knots than a message may, or a knot that is not at a finite place (the sentinels of a
// window no element with a weight has touched; the NaN / infinite knots a zero-weight FIRST element leaves behind,
// gfl_tf.c:231-240 with w = 0): such a series is left to the start-to-end kernel
template <class K>
ML_HD bool pw_from_window(const K& q, PwMsg& m) {
  m.n = 0;
  if (q.r - q.l + 1 > PW_CAP) return false;
  for (int i = q.l; i <= q.r; ++i) {
    const double x = q.ring.x(i);
    if (!is_finite(x)) return false;
    m.x[m.n] = x; m.a[m.n] = q.ring.a(i); m.b[m.n] = q.ring.b(i);
    ++m.n;
  }
  return m.n >= 1;
}
template <class Ring>
ML_HD bool pw_to_window(const PwMsg& m, KnotsC<Ring>& q) {
  if (m.n > Ring::CAP || m.n < 1) return false;
  for (int i = 0; i < m.n; ++i) q.ring.set(i, m.x[i], m.a[i], m.b[i]);
  q.l = 0;
  q.r = m.n - 1;
  knots_cache(q);
  return true;
}
// messages in global memory, [slot][index] (consecutive indices = consecutive addresses); n < 0 marks a failure
struct MsgArray {
  int32_t* n;
  double *x, *a, *b;
  int64_t stride;
};
ML_HD void pw_save(const PwMsg& m, const MsgArray& M, int64_t i) {
  for (int k = 0; k < m.n; ++k) {
    M.x[k * M.stride + i] = m.x[k];
    M.a[k * M.stride + i] = m.a[k];
    M.b[k * M.stride + i] = m.b[k];
  }
  M.n[i] = m.n;
}
ML_HD bool pw_load(const MsgArray& M, int64_t i, PwMsg& m) {
  const int n = M.n[i];
  if (n < 1 || n > PW_CAP) { m.n = 0; return false; }
  for (int k = 0; k < n; ++k) {
    m.x[k] = M.x[k * M.stride + i];
    m.a[k] = M.a[k * M.stride + i];
    m.b[k] = M.b[k * M.stride + i];
  }
  m.n = n;
  return true;
}
// the maps of chunks (or of groups of chunks): end messages from the two bounds and the two sums; sw == 0 (no element
// with a weight) is the identity, its messages are not stored (n == 0)
struct ScanFns {
  MsgArray A, B;
  double *sw, *swy;
};

// a cursor over a message: the line s b + t that holds just right of the knots consumed so far
struct PwCursor { int i; double s, t; };
ML_HD void pw_advance(const PwMsg& m, PwCursor& c) { c.s += m.a[c.i]; c.t += m.b[c.i]; ++c.i; }
ML_HD double pw_next_x(const PwMsg& m, const PwCursor& c) { return c.i < m.n ? m.x[c.i] : HUGE_VAL; }
ML_HD bool pw_emit(PwMsg& out, double x, double a, double b) {
  if (out.n >= PW_CAP) return false;
  out.x[out.n] = x; out.a[out.n] = a; out.b[out.n] = b;
  ++out.n;
  return true;
}
// A knot that switches between two lines which are the same line up to the rounding of their sums (G and A coincide
// wherever neither the incoming message nor A has been clipped yet) carries noise, not a change of line: it is dropped.
// Never on count data with an integer penalty, where the increments are exact (integers, or exactly zero).
ML_HD bool pw_emit_switch(PwMsg& out, double x, double a, double b, double scale_a, double scale_b) {
  const double tiny = 8.0 * 2.220446049250313e-16;
  if ((a < 0 ? -a : a) <= tiny * scale_a && (b < 0 ? -b : b) <= tiny * scale_b) return true;
  return pw_emit(out, x, a, b);
}

// out = the map (A, B, Sw, Swy), Sw > 0, applied to the message `in`.  All three messages regular (pw_from_window).
// Returns false when out would hold more than PW_CAP knots.
ML_HD bool pw_compose(const PwMsg& in, const PwMsg& A, const PwMsg& B, double Sw, double Swy, double lam, PwMsg& out) {
  PwCursor cA{0, 0.0, -lam}, cT{0, 0.0, -lam}, cB{0, 0.0, -lam};
  out.n = 0;
  const double sc_a = Sw, sc_b = (Swy < 0 ? -Swy : Swy) + lam;     // magnitudes the sums behind a switch knot were formed at
  // ---- b1: G - A = (sT + Sw - sA) b + (tT - Swy - tA) first reaches zero --------------------------------------------
  double b1 = -HUGE_VAL;
  for (;;) {
    const double xa = pw_next_x(A, cA), xt = pw_next_x(in, cT);
    const double hi = xa < xt ? xa : xt;               // the current lines hold up to the next knot of either list
    const double sl = (cT.s + Sw) - cA.s, ic = (cT.t - Swy) - cA.t;
    // G overtakes A at the latest on the last piece (G -> +inf there, A <= lam)
    const bool cross = hi == HUGE_VAL ? true : (sl * hi + ic >= 0);
    if (cross) {
      // A has no knot left: it has reached +lam, so has B above it, and G has not caught up with A before: out is A
      if (cA.i == A.n) return out.n >= 1;
      double c = sl != 0.0 ? -ic / sl : b1;
      if (!(c >= b1)) c = b1;                          // numerically before the piece: at its start
      if (c > hi) c = hi;
      b1 = c;
      break;
    }
    if (xa <= xt) {                                    // A's knot passes unchanged (out is A below b1)
      if (!pw_emit(out, A.x[cA.i], A.a[cA.i], A.b[cA.i])) return false;
      pw_advance(A, cA);
    } else {
      pw_advance(in, cT);                              // the incoming message's knots below b1 vanish under A
    }
    b1 = hi;
  }
  // (the knot at b1, from A's line to G's, is written once b2 is known: when the stretch on which G holds is empty the
  // two switches collapse into one knot from A's line to B's - or none, where A and B coincide)
  const double sA = cA.s, tA = cA.t;                   // A's line at b1
  const double gs1 = (cT.s + Sw) - sA, gt1 = (cT.t - Swy) - tA;
  // ---- b2: G - B first reaches zero (B's knots up to b1 first) -------------------------------------------------------
  while (pw_next_x(B, cB) <= b1) pw_advance(B, cB);
  double b2 = b1;
  bool between = false;                                // some knot of the incoming message lies on the stretch of G
  for (;;) {
    const double xb = pw_next_x(B, cB), xt = pw_next_x(in, cT);
    const double hi = xb < xt ? xb : xt;
    const double sl = (cT.s + Sw) - cB.s, ic = (cT.t - Swy) - cB.t;
    const bool cross = hi == HUGE_VAL ? true : (sl * hi + ic >= 0);
    if (cross) {
      if (cB.i == 0) {
        // B has not left -lam yet where G overtakes it: neither has A below it, and G stays above B from here on: out is B
        out.n = 0;
        for (int i = 0; i < B.n; ++i) pw_emit(out, B.x[i], B.a[i], B.b[i]);
        return out.n >= 1;
      }
      double c = sl != 0.0 ? -ic / sl : b2;
      if (!(c >= b2)) c = b2;
      if (c > hi) c = hi;
      b2 = c;
      break;
    }
    if (xt <= xb) {                                    // the incoming message's knot passes unchanged (G's increments are its own)
      if (!between) {
        if (!pw_emit_switch(out, b1, gs1, gt1, sc_a, sc_b)) return false;
        between = true;
      }
      if (!pw_emit(out, in.x[cT.i], in.a[cT.i], in.b[cT.i])) return false;
      pw_advance(in, cT);
    } else {
      pw_advance(B, cB);                               // B's knots below b2 lie above G
    }
    b2 = hi;
  }
  if (!between && b2 == b1) {
    // empty stretch: straight from A's line to B's
    if (!pw_emit_switch(out, b1, cB.s - sA, cB.t - tA, sc_a, sc_b)) return false;
  } else {
    if (!between && !pw_emit_switch(out, b1, gs1, gt1, sc_a, sc_b)) return false;
    // the knot at b2: from G's line to B's
    if (!pw_emit_switch(out, b2, cB.s - (cT.s + Sw), cB.t - (cT.t - Swy), sc_a, sc_b)) return false;
  }
  // then the rest of B
  for (; cB.i < B.n; ++cB.i)
    if (!pw_emit(out, B.x[cB.i], B.a[cB.i], B.b[cB.i])) return false;
  return out.n >= 1 && is_finite(b1) && is_finite(b2);
}

// A series whose first element has no weight does not start from a message of the regular kind: gfl_tf.c:231-240 with
// w = 0 leaves knots at -inf / +inf, and from the second weightless element on the window is "+lam everywhere" with a NaN
// knot of zero increments riding along - a fixed point of the weightless step, whose back-pointers are (-inf, NaN).  The
// scan keeps such a window LITERALLY (so that the steps run on it do what the reference does, bit for bit); as the input
// of a map it is the upper bound: the map's B, exactly.
ML_HD bool pw_is_plus_literal(const PwMsg& m) { return m.n >= 1 && m.n <= 2 && m.x[0] == -HUGE_VAL && m.a[0] == 0.0; }
ML_HD bool pw_is_regular(const PwMsg& m) {
  if (m.n < 1) return false;
  for (int i = 0; i < m.n; ++i)
    if (!is_finite(m.x[i])) return false;
  return true;
}
// The drivers of steps 2-4 below are written for a TEAM: on the device a warp whose lanes move the messages between global
// and shared memory together (one knot per lane, one round trip per message instead of one per knot) while its leader
// does the merging; on the CPU (tests/emu) a team of one.  Control flow is uniform across the team: every decision is
// taken on values all members see (counts in global memory, messages in memory shared by the team).
struct SerialTeam {
  ML_HD bool leader() const { return true; }
  ML_HD bool all(bool leaders_verdict) const { return leaders_verdict; }
  ML_HD void sync() const {}
  ML_HD bool load(const MsgArray& M, int64_t i, PwMsg& m) const { return pw_load(M, i, m); }
  ML_HD void save(const PwMsg& m, const MsgArray& M, int64_t i) const { pw_save(m, M, i); }
};
// *cur = the map (A, B, sw, swy) of chunk / group i of F applied to *cur (tmp, A, B: scratch messages of the team)
template <class Team>
ML_HD bool pw_apply(const Team& team, const ScanFns& F, int64_t i, double lam, PwMsg*& cur, PwMsg*& tmp, PwMsg& A, PwMsg& B) {
  const int na = F.A.n[i], nb = F.B.n[i];
  if (na < 0 || nb < 0) return false;
  if (na == 0 || nb == 0) return true;                 // identity
  if (pw_is_plus_literal(*cur)) return team.load(F.B, i, *cur);
  if (!team.load(F.A, i, A) || !team.load(F.B, i, B)) return false;
  const double sw = F.sw[i], swy = F.swy[i];
  bool ok = false;
  if (team.leader()) ok = pw_compose(*cur, A, B, sw, swy, lam, *tmp);
  ok = team.all(ok);
  team.sync();
  if (!ok) return false;
  PwMsg* t = cur; cur = tmp; tmp = t;
  return true;
}

// Step 1, one chunk and one side (sign < 0: from "-lam everywhere", the map's A; sign > 0: its B): the chunk's own steps
// from the bound, nothing written but the end message; the side sign < 0 also stores the chunk's sums.
// The FIRST chunk of a series is run from the first element instead (its side sign < 0 only): exact by construction,
// back-pointers written, and its map is the constant one (A = B = its end window, kept literally).
template <class Ring>
ML_HD void scan_bounds(const SeriesDesc& S, const ChunkDesc& cd, int64_t c, int sign, const double* __restrict__ y,
                       const double* __restrict__ w, double* __restrict__ tm, double* __restrict__ tp,
                       double* beta_last, const ScanFns& F, KnotsC<Ring>& q) {
  double t1, t2;
  if (cd.idx == 0) {
    if (sign > 0) return;
    knots_init_first(q, y[0], w[0], S.lam, t1, t2);    // gfl_tf.c:231-240
    tm[0] = t1;
    tp[0] = t2;
    bool irregular = false;
    const bool ok = flsa_redo_chunk(S, cd, y, w, tm, tp, q, &irregular);
    const int cnt = q.r - q.l + 1;
    PwMsg E;
    E.n = 0;
    if (ok && cnt <= PW_CAP) {
      for (int k = 0; k < cnt; ++k) { E.x[k] = q.ring.x(q.l + k); E.a[k] = q.ring.a(q.l + k); E.b[k] = q.ring.b(q.l + k); }
      E.n = cnt;
    }
    if (E.n > 0 && (pw_is_regular(E) || pw_is_plus_literal(E) || S.n_chunks == 1)) {
      pw_save(E, F.A, c);
      pw_save(E, F.B, c);
      if (S.n_chunks == 1) *beta_last = dp_last(q, y[S.n - 1], w[S.n - 1], S.lam);   // gfl_tf.c:294-302
    } else {
      F.A.n[c] = -1;
      F.B.n[c] = -1;
    }
    F.sw[c] = 1.0;                                     // (a constant map is never applied through its sums)
    F.swy[c] = 0.0;
    return;
  }
  knots_init_bound(q, sign, S.lam);
  bool pristine = true, zfix = false, ok = true;
  double sw = 0.0, swy = 0.0;
  double yn = 0.0, wn = 1.0;                       // one element ahead
  if (cd.kb < cd.ke) { yn = y[cd.kb]; wn = w[cd.kb]; }
  for (int k = cd.kb; k < cd.ke && ok; ++k) {
    const double yk = yn, wk = wn;
    if (k + 1 < cd.ke) { yn = y[k + 1]; wn = w[k + 1]; }
    sw += wk;
    swy += wk * yk;
    if (pristine) {
      if (!(wk > 0)) continue;     // a constant message stays what it is under a zero-weight element (flsa_spec_run)
      pristine = false;
    }
    ok = dp_step_z(q, yk, wk, S.lam, t1, t2, zfix);
  }
  const MsgArray& M = sign < 0 ? F.A : F.B;
  if (!ok || !is_finite(sw) || !is_finite(swy)) {
    M.n[c] = -1;
  } else if (pristine || !(sw > 0)) {
    M.n[c] = (pristine && sw == 0.0) ? 0 : -1;       // identity; anything else without a window is not a map we can state
  } else {
    const int cnt = q.r - q.l + 1;
    bool good = cnt <= PW_CAP;
    for (int k = 0; k < cnt && good; ++k) {
      const double x = q.ring.x(q.l + k);
      good = is_finite(x);
      M.x[k * M.stride + c] = x;
      M.a[k * M.stride + c] = q.ring.a(q.l + k);
      M.b[k * M.stride + c] = q.ring.b(q.l + k);
    }
    M.n[c] = good ? cnt : -1;
  }
  if (sign < 0) { F.sw[c] = sw; F.swy[c] = swy; }
}

// Step 2, one group and one side: the maps of chunks [c0, c0 + nc) composed in order.  m0 / m1 / A / B are the caller's
// scratch messages; on return *res points at the side's message of the group's map (n == 0 and *Sw == 0: the identity).
// In a series' first group the map is the constant one (its first chunk's is): both sides are the group's exact end
// message, only the side sign < 0 is computed.  Returns false when a chunk failed or a message outgrew PW_CAP.
template <class Team>
ML_HD bool scan_reduce_side(const Team& team, int64_t c0, int nc, int sign, const ScanFns& F, double lam, PwMsg* m0, PwMsg* m1,
                            PwMsg& A, PwMsg& B, PwMsg** res, double* Sw, double* Swy) {
  bool pristine = true;
  double sw_all = 0.0, swy_all = 0.0;
  PwMsg *cur = m0, *tmp = m1;
  if (team.leader()) cur->n = 0;
  team.sync();
  for (int64_t c = c0; c < c0 + nc; ++c) {
    const int na = F.A.n[c], nb = F.B.n[c];
    if (na < 0 || nb < 0) return false;
    if (na == 0 || nb == 0) continue;                  // identity
    if (pristine) {
      if (!team.load(sign < 0 ? F.A : F.B, c, *cur)) return false;
      pristine = false;
    } else if (!pw_apply(team, F, c, lam, cur, tmp, A, B)) {
      return false;
    }
    sw_all += F.sw[c];
    swy_all += F.swy[c];
  }
  *res = cur;
  *Sw = sw_all;
  *Swy = swy_all;
  return true;
}

// Step 3, one series: the exact message in front of each of its groups [g0 + 1, g0 + ng), written to GT (the first group
// starts at the head of the series; its map is constant: its A message is the exact state behind it).
template <class Team>
ML_HD bool scan_carry(const Team& team, double lam, int64_t g0, int ng, const ScanFns& G, const MsgArray& GT, PwMsg* m0,
                      PwMsg* m1, PwMsg& A, PwMsg& B) {
  if (ng <= 1) return true;
  PwMsg *cur = m0, *tmp = m1;
  if (!team.load(G.A, g0, *cur)) return false;
  for (int64_t g = g0 + 1; g < g0 + ng; ++g) {
    team.save(*cur, GT, g);
    if (g + 1 == g0 + ng) break;                       // the last group's map is not needed
    if (!pw_apply(team, G, g, lam, cur, tmp, A, B)) return false;
  }
  return true;
}

// Step 4, one group: the exact message in front of each of its chunks.  The start message of chunk c takes the place of
// the chunk's A message (which has been read by then).  `first`: the series' first group - its first chunk needs no
// start (it was run from the first element), the second one starts from the first one's end window.
template <class Team>
ML_HD bool scan_starts(const Team& team, int64_t g, int64_t c0, int nc, bool first, const ScanFns& F, const MsgArray& GT,
                       double lam, PwMsg* m0, PwMsg* m1, PwMsg& A, PwMsg& B) {
  PwMsg *cur = m0, *tmp = m1;
  int64_t c = c0;
  if (first) {
    if (nc <= 1) return F.A.n[c0] >= 1;
    if (!team.load(F.A, c0, *cur)) return false;
    c = c0 + 1;
  } else if (!team.load(GT, g, *cur)) {
    return false;
  }
  for (; c < c0 + nc; ++c) {
    const int na = F.A.n[c], nb = F.B.n[c];
    if (na < 0 || nb < 0) return false;
    const bool last = c + 1 == c0 + nc;
    if (last || na == 0 || nb == 0) { team.save(*cur, F.A, c); continue; }
    if (pw_is_plus_literal(*cur)) {
      team.save(*cur, F.A, c);
      if (!team.load(F.B, c, *cur)) return false;
      continue;
    }
    if (!team.load(F.A, c, A) || !team.load(F.B, c, B)) return false;
    team.save(*cur, F.A, c);
    const double sw = F.sw[c], swy = F.swy[c];
    bool ok = false;
    if (team.leader()) ok = pw_compose(*cur, A, B, sw, swy, lam, *tmp);
    ok = team.all(ok);
    team.sync();
    if (!ok) return false;
    PwMsg* t = cur; cur = tmp; tmp = t;
  }
  return true;
}

// Step 5, one chunk but a series' first: redone from its exact start window (stored by step 4 in the place of the chunk's
// A message); the series' last chunk also yields the last coefficient.  Returns false on a missing start or ring overflow.
template <class Ring>
ML_HD bool scan_redo(const SeriesDesc& S, const ChunkDesc& cd, int64_t c, const double* __restrict__ y,
                     const double* __restrict__ w, double* __restrict__ tm, double* __restrict__ tp, const MsgArray& T,
                     KnotsC<Ring>& q, double* beta_last) {
  const int n = T.n[c];
  if (n < 1 || n > Ring::CAP) return false;
  if (cd.idx == 0) return true;
  for (int k0 = 0; k0 < n; k0 += 4) {                  // four knots' loads in flight before the first of them is stored
    double xs[4], as[4], bs[4];
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
    for (int u = 0; u < 4; ++u)
      if (k0 + u < n) { xs[u] = T.x[(k0 + u) * T.stride + c]; as[u] = T.a[(k0 + u) * T.stride + c]; bs[u] = T.b[(k0 + u) * T.stride + c]; }
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
    for (int u = 0; u < 4; ++u)
      if (k0 + u < n) q.ring.set(k0 + u, xs[u], as[u], bs[u]);
  }
  q.l = 0;
  q.r = n - 1;
  knots_cache(q);
  bool irregular = false;
  if (!flsa_redo_chunk(S, cd, y, w, tm, tp, q, &irregular)) return false;
  if (cd.idx == S.n_chunks - 1) *beta_last = dp_last(q, y[S.n - 1], w[S.n - 1], S.lam);   // gfl_tf.c:294-302
  return true;
}

// ---- the chains of the speculative route by composition ("fine chains") ---------------------------------------------
// A chain of unverified chunks is walked at one thread's pace: 448 dependent DP steps per chunk, 0.63 ms for a chain of two
// on the bench genome - a floor that every solve pays, whatever its size.  The same algebra as the scan route's cuts it:
// every chunk of the chain is split into FINE_SUB sub-chunks, each of them is run from the two bounds on its own steps (two
// threads per sub-chunk), ONE warp per chain turns the exact end window of the verified chunk in front of the chain into
// the exact start message of every sub-chunk (scan_starts on the E windows), and every sub-chunk is redone from its start
// by a thread of its own: 2 x 56 dependent steps and at most FINE_SUB x (chain length) merges.  A chain whose messages
// outgrow PW_CAP, or whose start window is not of the regular kind, is left to a walker as before.
constexpr int FINE_SUB = 8;
struct FineChain {
  int32_t head;     // first chunk of the chain (index into the speculative route's chunk table)
  int32_t len;      // chunks
  int32_t base;     // its first item in the item tables
  int32_t ok;       // set by the warp that built the start messages
};
struct FineItem {
  int32_t chunk;    // chunk of the speculative route's table
  int32_t sub;      // 0 .. FINE_SUB - 1
  int32_t chain;    // index into the chain table
  int32_t pad;
};
// DP steps [kb, ke) of sub-chunk `sub` of a chunk (the last ones may be empty)
ML_HD void fine_span(const ChunkDesc& cd, int sub, int& kb, int& ke) {
  const int len = cd.ke - cd.kb;
  const int sl = (len + FINE_SUB - 1) / FINE_SUB;
  kb = cd.kb + sub * sl;
  if (kb > cd.ke) kb = cd.ke;
  ke = kb + sl < cd.ke ? kb + sl : cd.ke;
}
// the E windows of the speculative runs seen as an array of messages (same [slot][chunk] layout)
ML_HD MsgArray fine_end_windows(const SpecWindows& W) { return MsgArray{W.e_count, W.ex, W.ea, W.eb, W.stride}; }
// one chain: exact start messages of its n_items sub-chunks into F.A[base ..] (team: a warp, or the emulation's team of one)
template <class Team>
ML_HD bool fine_chain_starts(const Team& team, const SpecWindows& W, int64_t head, int n_items, int64_t base, const ScanFns& F,
                             double lam, PwMsg* m0, PwMsg* m1, PwMsg& A, PwMsg& B) {
  // the message in front of the chain must be one the merges understand
  if (!team.load(fine_end_windows(W), head - 1, *m0)) return false;
  if (!pw_is_regular(*m0) && !pw_is_plus_literal(*m0)) return false;
  return scan_starts(team, head - 1, base, n_items, false, F, fine_end_windows(W), lam, m0, m1, A, B);
}
// one sub-chunk redone from its exact start (F.A[item]); `last`: it ends its series (the last coefficient is due)
template <class Ring>
ML_HD bool fine_redo(const SeriesDesc& S, int kb, int ke, int64_t item, bool last, const double* __restrict__ y,
                     const double* __restrict__ w, double* __restrict__ tm, double* __restrict__ tp, const MsgArray& T,
                     KnotsC<Ring>& q, double* beta_last) {
  const int n = T.n[item];
  if (n < 1 || n > Ring::CAP) return false;
  for (int k = 0; k < n; ++k) q.ring.set(k, T.x[k * T.stride + item], T.a[k * T.stride + item], T.b[k * T.stride + item]);
  q.l = 0;
  q.r = n - 1;
  knots_cache(q);
  bool irregular = false;
  const ChunkDesc span{0, kb, ke, 1};
  if (!flsa_redo_chunk(S, span, y, w, tm, tp, q, &irregular)) return false;
  if (last) *beta_last = dp_last(q, y[S.n - 1], w[S.n - 1], S.lam);   // gfl_tf.c:294-302
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

/* Experiment (design evidence for DESIGN.md): how many elements does the weighted FLSA DP need
 * before its state no longer depends on anything earlier?  Two runs are started at element s0
 * from the two extreme incoming messages (h = -lam and h = +lam everywhere); the message update
 * is monotone in the incoming message, so once both knot windows are identical the true state is
 * pinned.  Prints the distribution of coalescence lengths and checks bit-equality with the
 * sequential run. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define CAP 4096
typedef struct { double x[CAP], a[CAP], b[CAP]; int l, r; int maxsz; } dq;

static inline int step(dq *q, double w, double y, double lam, double *tm, double *tp) {
  double alo = w, blo = -w * y - lam, ahi = -w, bhi = w * y - lam;
  int lo = q->l, hi;
  while (lo <= q->r) { int i = lo & (CAP - 1); if (alo * q->x[i] + blo > -lam) break; alo += q->a[i]; blo += q->b[i]; lo++; }
  hi = q->r;
  while (hi >= lo) { int i = hi & (CAP - 1); if (-ahi * q->x[i] - bhi < lam) break; ahi += q->a[i]; bhi += q->b[i]; hi--; }
  *tm = (-lam - blo) / alo; *tp = (lam + bhi) / (-ahi);
  q->l = lo - 1; q->r = hi + 1;
  if (q->r - q->l + 1 > CAP) return -1;
  int i = q->l & (CAP - 1); q->x[i] = *tm; q->a[i] = alo; q->b[i] = blo + lam;
  i = q->r & (CAP - 1); q->x[i] = *tp; q->a[i] = ahi; q->b[i] = bhi + lam;
  if (q->r - q->l + 1 > q->maxsz) q->maxsz = q->r - q->l + 1;
  return 0;
}
static void init_first(dq *q, double w, double y, double lam) { /* reference's manual first step */
  q->l = 0; q->r = 1; q->maxsz = 2;
  q->x[0] = -lam / w + y; q->a[0] = w; q->b[0] = -w * y + lam;
  q->x[1] = lam / w + y; q->a[1] = -w; q->b[1] = w * y + lam;
}
static void init_bound(dq *q, int sign, double lam) { /* h == sign*lam everywhere */
  q->l = 0; q->r = 0; q->maxsz = 1; q->x[0] = sign < 0 ? INFINITY : -INFINITY; q->a[0] = 0; q->b[0] = 2 * lam;
}
static int same(const dq *p, const dq *q) {
  if (p->r - p->l != q->r - q->l) return 0;
  for (int k = 0; k <= p->r - p->l; k++) {
    int i = (p->l + k) & (CAP - 1), j = (q->l + k) & (CAP - 1);
    if (memcmp(&p->x[i], &q->x[j], 8) || memcmp(&p->a[i], &q->a[j], 8) || memcmp(&p->b[i], &q->b[j], 8)) return 0;
  }
  return 1;
}
/* returns steps to coalesce from s0 (or -1), also whether coalesced state == true state bitwise */
int coalesce_from(const double *y, const double *w, int n, double lam, int s0, int maxsteps, const dq *truth_at /*state after each step, may be NULL*/, int *biteq) {
  static dq m, p;
  init_bound(&m, -1, lam); init_bound(&p, +1, lam);
  double tm, tp;
  for (int k = s0; k < n - 1 && k - s0 < maxsteps; k++) {
    if (step(&m, w[k], y[k], lam, &tm, &tp) || step(&p, w[k], y[k], lam, &tm, &tp)) return -2;
    if (same(&m, &p)) { if (biteq && truth_at) *biteq = same(&m, truth_at); return k - s0 + 1; }
  }
  return -1;
}
/* full experiment: sequential truth; for s0 on a grid measure coalescence length; verify that the state
   after coalescence equals the truth state at that element */
void experiment(const double *y, const double *w, int n, double lam, int stride, int maxsteps, int *out_len, int *out_biteq, int *out_n, int *stats) {
  dq *t = malloc(sizeof(dq)); double tm, tp;
  init_first(t, w[0], y[0], lam);
  int cnt = 0; long pops = 0;
  /* we need truth state at arbitrary k: re-run sequentially, pausing at grid points */
  int next = stride; 
  dq *snap = malloc(sizeof(dq));
  for (int k = 1; k < n - 1; k++) {
    if (k == next) {
      /* truth state before folding k is *t. run coalescence from k, tracking truth alongside */
      memcpy(snap, t, sizeof(dq));
      static dq m, p; init_bound(&m, -1, lam); init_bound(&p, +1, lam);
      int len = -1, be = 0;
      for (int kk = k; kk < n - 1 && kk - k < maxsteps; kk++) {
        double a1, a2;
        step(snap, w[kk], y[kk], lam, &a1, &a2);
        step(&m, w[kk], y[kk], lam, &a1, &a2); step(&p, w[kk], y[kk], lam, &a1, &a2);
        if (same(&m, &p)) { len = kk - k + 1; be = same(&m, snap); break; }
      }
      out_len[cnt] = len; out_biteq[cnt] = be; cnt++;
      next += stride;
    }
    if (step(t, w[k], y[k], lam, &tm, &tp)) { stats[0] = -1; break; }
  }
  *out_n = cnt; stats[1] = t->maxsz;
  free(t); free(snap);
}

// CPU run of methylasso_b200/csrc/valley_core.cuh: the per-element functions the CUDA kernels of valleys.inc.cuh call,
// driven by plain loops and host compactions in the order call_valleys_device launches them.  Test infrastructure only.
#include <cmath>
#include <cstdint>
#include <vector>
#include "../../methylasso_b200/csrc/valley_core.cuh"
#include "../../methylasso_b200/csrc/pmdsplit_core.cuh"
#include "../../methylasso_b200/csrc/pmdstatus_core.cuh"

using namespace ml;

static std::vector<int32_t> compact(const std::vector<int32_t>& flag) {
  std::vector<int32_t> out;
  for (size_t i = 0; i < flag.size(); ++i) if (flag[i]) out.push_back((int32_t)i);
  return out;
}

extern "C" {

// returns the number of valleys; outputs sized by the caller (n entries each)
int64_t emu_valleys(int64_t n, const int32_t* job, const int32_t* est, const int32_t* ncpg, const uint32_t* start,
                    const uint32_t* end, const double* beta, const double* coverage, const double* pw,
                    const int8_t* category, double valley_max_beta, double pmd_valley_min_width, double max_distance,
                    int32_t min_num_cpgs, int32_t* o_job, int32_t* o_est, uint32_t* o_start, uint32_t* o_end,
                    double* o_width, double* o_beta, double* o_cov, double* o_pw, int32_t* o_ncpg, int8_t* row_in,
                    double* low_w, double* low_b, int32_t* o_row0) {
  ValleyParams p;
  p.valley_max_beta = valley_max_beta; p.pmd_valley_min_width = pmd_valley_min_width; p.max_distance = max_distance;
  p.min_num_cpgs = min_num_cpgs;
  if (n == 0) return 0;
  std::vector<int32_t> run_of_head((size_t)n, -1), flag((size_t)n);
  for (int64_t i = 0; i < n; ++i) { low_w[i] = NAN; low_b[i] = NAN; row_in[i] = 0; }
  const ValleyRows c{n, job, est, ncpg, start, end, beta, coverage, pw, category, low_w, low_b, run_of_head.data()};
  for (int64_t i = 0; i < n; ++i) flag[i] = vl_run_head(c, p, i);
  const std::vector<int32_t> head = compact(flag);
  const int32_t n_runs = (int32_t)head.size();
  std::vector<int32_t> r_last(n_runs + 1), r_ncpg(n_runs + 1);
  std::vector<double> r_cov(n_runs + 1), r_pw(n_runs + 1);
  std::vector<int8_t> r_cat(n_runs + 1), r_keep(n_runs + 1);
  const ValleyRuns R{n_runs, head.data(), r_last.data(), r_ncpg.data(), r_cov.data(), r_pw.data(), r_cat.data(), r_keep.data()};
  for (int32_t r = 0; r < n_runs; ++r) vl_run_stats(c, p, R, r);
  for (int64_t i = 0; i < n; ++i) {
    const int32_t u = run_of_head[i];
    flag[i] = ((u >= 0 && r_keep[u]) || vl_umr_added(c, R, i)) ? 1 : 0;
  }
  const std::vector<int32_t> item_row = compact(flag);
  const int32_t n_items = (int32_t)item_row.size();
  if (n_items == 0) return 0;
  const ValleyItems I{n_items, item_row.data()};
  std::vector<int32_t> iflag(n_items);
  for (int32_t r = 0; r < n_items; ++r) iflag[r] = vl_item_head(c, p, R, I, r);
  const std::vector<int32_t> vhead = compact(iflag);
  const int32_t n_merged = (int32_t)vhead.size();
  std::vector<int32_t> m_job(n_merged), m_est(n_merged), m_ncpg(n_merged);
  std::vector<uint32_t> m_start(n_merged), m_end(n_merged);
  std::vector<double> m_w(n_merged), m_b(n_merged), m_cov(n_merged), m_pw(n_merged);
  std::vector<int8_t> m_keep(n_merged);
  std::vector<int32_t> m_row0(n_merged);
  const ValleyOut mo{m_job.data(), m_est.data(), m_ncpg.data(), m_start.data(), m_end.data(), m_w.data(), m_b.data(),
                     m_cov.data(), m_pw.data(), m_keep.data(), m_row0.data()};
  for (int32_t v = 0; v < n_merged; ++v) vl_merge(c, p, R, I, vhead.data(), n_merged, v, mo);
  int64_t k = 0;
  for (int32_t v = 0; v < n_merged; ++v) {
    if (!m_keep[v]) continue;
    o_job[k] = m_job[v]; o_est[k] = m_est[v]; o_start[k] = m_start[v]; o_end[k] = m_end[v]; o_width[k] = m_w[v];
    o_beta[k] = m_b[v]; o_cov[k] = m_cov[v]; o_pw[k] = m_pw[v]; o_ncpg[k] = m_ncpg[v]; o_row0[k] = m_row0[v];
    ++k;
  }
  for (int64_t i = 0; i < n; ++i) row_in[i] = vl_row_in_valley(c, i, (int32_t)k, o_job, o_est, o_start, o_end);
  return k;
}

// R/call.R:131-164 with the functions of pmdsplit_core.cuh, in the order call_pmd_split_device launches them.
// Outputs sized by the caller: merged n + n_v entries, pieces n_x * (n + n_v) at most (the caller passes a capacity).
// Returns the number of merged rows; *n_pieces receives the number of pieces (-1: capacity exceeded).
int64_t emu_pmd_split(int64_t n, const int32_t* job, const int32_t* est, const uint32_t* start, const uint32_t* end,
                      const int8_t* category, const int8_t* row_in, int32_t n_v, int32_t* v_job, int32_t* v_est,
                      uint32_t* v_start, uint32_t* v_end, int32_t* v_row0, int32_t n_x, const int32_t* x_job,
                      const int32_t* x_est, const uint32_t* x_start, const uint32_t* x_end, const double* x_fused,
                      const double* x_pw, double pmd_valley_min_width, int split_pmds, int32_t* m_job, int32_t* m_est,
                      int32_t* m_seg, uint32_t* m_start, uint32_t* m_end, int8_t* m_cat, int32_t* m_src_row,
                      int32_t* m_src_valley, int8_t* m_gap, int64_t cap, int64_t* n_pieces, int32_t* p_job,
                      int32_t* p_est, int32_t* p_seg, uint32_t* p_start, uint32_t* p_end, double* p_fused, double* p_pw) {
  *n_pieces = 0;
  const ValleyRows c{n, job, est, nullptr, start, end, nullptr, nullptr, nullptr, category, nullptr, nullptr, nullptr};
  const ValleyOut v{v_job, v_est, nullptr, v_start, v_end, nullptr, nullptr, nullptr, nullptr, nullptr, v_row0};
  std::vector<int32_t> kept_before((size_t)n + 1, 0);
  for (int64_t i = 0; i < n; ++i) kept_before[i + 1] = kept_before[i] + (row_in[i] ? 0 : 1);
  const int32_t nm = kept_before[n] + n_v;
  if (nm == 0) return 0;
  const MergedTable M{nm, m_job, m_est, m_src_row, m_src_valley, m_seg, m_start, m_end, m_cat, m_gap};
  for (int64_t i = 0; i < n; ++i) if (!row_in[i]) ms_put_row(c, M, ms_row_slot(i, kept_before[i], n_v, v_row0), i);
  for (int32_t k = 0; k < n_v; ++k) ms_put_valley(v, M, k + kept_before[v_row0[k]], k);
  std::vector<int32_t> flag(nm);
  int32_t head = 0;
  for (int32_t m = 0; m < nm; ++m) {
    m_gap[m] = ms_gap(M, pmd_valley_min_width, m);
    if (ms_group_head(M, m)) head = m;
    m_seg[m] = m - head + 1;
  }
  for (int32_t m = 0; m < nm; ++m) flag[m] = ms_region_head(M, split_pmds, m);
  const std::vector<int32_t> rhead = compact(flag);
  const int32_t n_reg = (int32_t)rhead.size();
  std::vector<int32_t> g_job(n_reg), g_est(n_reg);
  std::vector<uint32_t> g_start(n_reg), g_end(n_reg);
  std::vector<int8_t> g_iso(n_reg);
  const Regions G{n_reg, rhead.data(), g_job.data(), g_est.data(), g_start.data(), g_end.data(), g_iso.data()};
  for (int32_t r = 0; r < n_reg; ++r) ms_region(M, split_pmds, G, r);
  const PmdSegs X{n_x, x_job, x_est, x_start, x_end, x_fused, x_pw};
  std::vector<int32_t> off((size_t)n_x + 1, 0), cnt((size_t)n_x + 1, 0), na_group((size_t)n_x + 1, 0);
  for (int32_t x = 0; x < n_x; ++x) { cnt[x] = ms_count_pieces(G, X, x); off[x + 1] = off[x] + cnt[x]; }
  for (int32_t x = 0; x < n_x; ++x) ms_group_unmatched(X, cnt.data(), x, na_group.data());
  if (off[n_x] > cap) { *n_pieces = -1; return nm; }
  std::vector<int32_t> p_na((size_t)off[n_x] + 1);
  const Pieces P{p_job, p_est, p_na.data(), p_start, p_end, p_fused, p_pw};
  for (int32_t x = 0; x < n_x; ++x) ms_emit_pieces(G, X, x, off[x], na_group[x], P, cnt[x]);
  for (int32_t k = 0; k < off[n_x]; ++k) {
    if (k == 0 || p_job[k] != p_job[k - 1] || p_est[k] != p_est[k - 1]) head = k;
    p_seg[k] = k - head + 1 + p_na[k];
  }
  *n_pieces = off[n_x];
  return nm;
}

// R/call.R:201-215 with the functions of pmdstatus_core.cuh, in the order call_pmd_status_device launches them.
// Outputs hold up to 2 * nm rows.  Returns the number of final rows.
int64_t emu_pmd_status(int32_t nm, int32_t* m_job, int32_t* m_est, uint32_t* m_start, uint32_t* m_end, int8_t* m_cat,
                       int32_t* m_src_row, const int8_t* row_adm, int32_t n_g, const int32_t* g_job, const int32_t* g_est,
                       const uint32_t* g_start, const uint32_t* g_end, const int8_t* g_cat, int32_t* f_merged,
                       int8_t* f_status, int8_t* f_category, int8_t* f_adm) {
  if (nm == 0) return 0;
  const MergedTable M{nm, m_job, m_est, m_src_row, nullptr, nullptr, m_start, m_end, m_cat, nullptr};
  const PmdRegions G{n_g, g_job, g_est, g_start, g_end, g_cat};
  std::vector<int32_t> extra((size_t)nm + 1, 0);
  for (int32_t m = 0; m < nm; ++m) { int8_t a, b; extra[m + 1] = extra[m] + st_statuses(M, G, m, a, b) - 1; }
  const int32_t nf = nm + extra[nm];
  const FinalRows F{f_merged, f_status, f_category, f_adm};
  for (int32_t m = 0; m < nm; ++m) st_emit(M, G, row_adm, m, m + extra[m], F);
  std::vector<int32_t> flag(nf);
  for (int32_t i = 0; i < nf; ++i) flag[i] = f_adm[i] == 1 ? 1 : 0;
  const std::vector<int32_t> list = compact(flag);
  std::vector<int8_t> ns(f_status, f_status + nf);
  for (int32_t r = 0; r < (int32_t)list.size(); ++r) ns[list[r]] = st_umr_status(M, F, list.data(), (int32_t)list.size(), r);
  for (int32_t i = 0; i < nf; ++i) f_status[i] = ns[i];
  return nf;
}

// sum of doubles with x87.cuh's 64-bit-significand accumulator vs the host's long double (summary.c rsum)
double emu_x87_sum(int n, const double* x) {
  X87 s{0.0, 0.0};
  for (int i = 0; i < n; ++i) s = x87_add_d(s, x[i]);
  return x87_to_double(s);
}
double ref_x87_sum(int n, const double* x) {
  volatile long double s = 0.0L;
  for (int i = 0; i < n; ++i) s = s + (long double)x[i];
  return (double)s;
}
}

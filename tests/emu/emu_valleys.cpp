// CPU run of methylasso_b200/csrc/valley_core.cuh: the per-element functions the CUDA kernels of valleys.inc.cuh call,
// driven by plain loops and host compactions in the order call_valleys_device launches them.  Test infrastructure only.
#include <cmath>
#include <cstdint>
#include <vector>
#include "../../methylasso_b200/csrc/valley_core.cuh"

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
                    double* low_w, double* low_b) {
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
  const ValleyOut mo{m_job.data(), m_est.data(), m_ncpg.data(), m_start.data(), m_end.data(), m_w.data(), m_b.data(),
                     m_cov.data(), m_pw.data(), m_keep.data()};
  for (int32_t v = 0; v < n_merged; ++v) vl_merge(c, p, R, I, vhead.data(), n_merged, v, mo);
  int64_t k = 0;
  for (int32_t v = 0; v < n_merged; ++v) {
    if (!m_keep[v]) continue;
    o_job[k] = m_job[v]; o_est[k] = m_est[v]; o_start[k] = m_start[v]; o_end[k] = m_end[v]; o_width[k] = m_w[v];
    o_beta[k] = m_b[v]; o_cov[k] = m_cov[v]; o_pw[k] = m_pw[v]; o_ncpg[k] = m_ncpg[v];
    ++k;
  }
  for (int64_t i = 0; i < n; ++i) row_in[i] = vl_row_in_valley(c, i, (int32_t)k, o_job, o_est, o_start, o_end);
  return k;
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

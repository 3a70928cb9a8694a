// methylasso_b200_glue.cpp — drop-in replacement for the reference's Rcpp module
// (/root/reference/src/MethyLasso_cpp.cpp:9-31): the same five exported functions with the same
// argument names and defaults, forwarding to the C ABI of include/methylasso_b200.h, plus one batched
// entry that takes all chromosomes at once.
//
// NOT COMPILED IN THIS REPOSITORY'S IMAGE (no R / Rcpp here).  Build inside the R package with
//   PKG_CPPFLAGS = -I<repo>/include      PKG_LIBS = -L<repo>/methylasso_b200 -lmethylasso_b200
// (see INTEGRATION.md).  Everything below is thin marshalling: Rcpp vectors -> int32 / double
// pointers, error codes -> Rcpp::stop with the reference's messages.
#include <Rcpp.h>
#include <vector>

#include "methylasso_b200.h"

using namespace Rcpp;

namespace {

ml_engine* engine() {                       // one engine per R process, created on first use
  static ml_engine* eng = nullptr;
  if (!eng) {
    if (ml_engine_create(-1, &eng) != ML_OK) stop(ml_last_error());
  }
  return eng;
}

void check(int rc) { if (rc != ML_OK) stop(ml_last_error()); }

void need_columns(const DataFrame& data) {   // src/signal_detection.cpp:13-15
  if (!(data.containsElementNamed("dataset") && data.containsElementNamed("condition") &&
        data.containsElementNamed("replicate") && data.containsElementNamed("pos") &&
        data.containsElementNamed("Nmeth") && data.containsElementNamed("coverage")))
    stop("data should contain columns dataset, condition, replicate, pos, Nmeth and coverage!");
}

// one job (job_ptr = {0, n}) or many (job_ptr given): returns the reference's List per job
List detect(const DataFrame& data, const std::vector<int64_t>& job_ptr, const ml_params& p, bool diff_fast,
            bool want_mu = true) {
  need_columns(data);
  IntegerVector dset = data["dataset"], cond = data["condition"], rep = data["replicate"], pos = data["pos"],
                nmeth = data["Nmeth"], cov = data["coverage"];   // Rcpp coerces doubles to int
  const int nj = (int)job_ptr.size() - 1;
  ml_result* res = nullptr;
  std::vector<int> status(nj);
  check(ml_detect_batch(engine(), nj, job_ptr.data(), dset.begin(), cond.begin(), rep.begin(), pos.begin(),
                        nmeth.begin(), cov.begin(), &p, &res, status.data()));
  if (diff_fast) {
    const int rc_diff = ml_result_fast_diff_combine(engine(), res);
    if (rc_diff != ML_OK) { const std::string msg = ml_last_error(); ml_result_free(res); stop(msg); }
  }
  Rcpp::checkUserInterrupt();        // the fit is done and resident; nothing is lost by stopping here but the result
  List out(nj);
  for (int j = 0; j < nj; ++j) {
    int64_t n, P; int32_t E, D, st;
    ml_result_job_sizes(res, j, &n, &P, &E, &D, &st);
    if (st != ML_OK) {                       // R wrapper drops the chromosome (.errorhandling = "remove")
      if (nj == 1) { ml_result_free(res); stop(ml_strerror(st)); }
      out[j] = R_NilValue;
      continue;
    }
    // mu is 8 bytes per data row, 39 % of everything that comes back for a genome, and neither MethyLasso.R nor R/call.R
    // ever reads ret$mu (R/genome_wide.R:9 only concatenates it): want_mu = FALSE leaves it on the device
    NumericVector mu(want_mu ? n : 0), lambda2(E), offset(D), start(E * P), beta(E * P), betahat(E * P), pw(E * P);
    IntegerVector id(E * P);
    const int rc_copy = ml_result_copy_job(engine(), res, j, want_mu ? mu.begin() : nullptr, lambda2.begin(), offset.begin(),
                                           id.begin(), start.begin(), beta.begin(), betahat.begin(), pw.begin());
    if (rc_copy != ML_OK) { const std::string msg = ml_last_error(); ml_result_free(res); stop(msg); }
    int32_t conv; double hyper;
    ml_result_job_info(res, j, &conv, nullptr, nullptr, nullptr, &hyper, nullptr);
    DataFrame est = DataFrame::create(_["estimate_id"] = id, _["start"] = start, _["beta"] = beta,
                                      _["betahat"] = betahat, _["pseudoweight"] = pw);
    out[j] = List::create(_["mu"] = mu, _["lambda2"] = lambda2, _["hyper"] = hyper,
                          _["converged"] = (bool)conv, _["offset"] = offset, _["estimates"] = est);
  }
  ml_result_free(res);
  return out;
}

ml_params fast_params(double lambda2, double tol_val, unsigned max_iter) {
  ml_params p = {lambda2, lambda2 + 1, 0.0, tol_val, 50.0, 1.0, max_iter, 1u, 1u, ML_MODEL_FAST, 0, 0};
  return p;
}

}  // namespace

// src/signal_detection.cpp:11-47
List signal_detection_fast_R(const DataFrame& data, double lambda2, double tol_val, unsigned max_iter) {
  List r = detect(data, {0, (int64_t)data.nrows()}, fast_params(lambda2, tol_val, max_iter), false)[0];
  return List::create(_["mu"] = r["mu"], _["lambda2"] = r["lambda2"], _["converged"] = r["converged"],
                      _["offset"] = r["offset"], _["estimates"] = r["estimates"],
                      _["detection.type"] = "signal", _["model"] = "normal");
}

// src/signal_detection.cpp:49-80
List signal_detection_R(const DataFrame& data, bool optimize_lambda2, double lambda2, double max_lambda2,
                        bool optimize_hyper, double hyper, unsigned nouter, double bf_per_kb, double tol_val,
                        unsigned max_iter) {
  ml_params p = {lambda2, max_lambda2, 0.0, tol_val, bf_per_kb, hyper, max_iter, nouter, 1u,
                 ML_MODEL_FULL_SIGNAL, optimize_lambda2, optimize_hyper};
  List r = detect(data, {0, (int64_t)data.nrows()}, p, false)[0];
  return List::create(_["mu"] = r["mu"], _["lambda2"] = r["lambda2"], _["hyper"] = r["hyper"],
                      _["converged"] = r["converged"], _["offset"] = r["offset"], _["estimates"] = r["estimates"],
                      _["detection.type"] = "signal", _["model"] = "beta-binomial");
}

// src/difference_detection.cpp:13-45
List difference_detection_R(const DataFrame& data, unsigned ref, bool optimize_lambda2, double lambda2,
                            double max_lambda2, bool optimize_hyper, double hyper, unsigned nouter,
                            double bf_per_kb, double tol_val, unsigned max_iter) {
  ml_params p = {lambda2, max_lambda2, 0.0, tol_val, bf_per_kb, hyper, max_iter, nouter, ref,
                 ML_MODEL_FULL_DIFFERENCE, optimize_lambda2, optimize_hyper};
  List r = detect(data, {0, (int64_t)data.nrows()}, p, false)[0];
  return List::create(_["mu"] = r["mu"], _["lambda2"] = r["lambda2"], _["hyper"] = r["hyper"],
                      _["converged"] = r["converged"], _["offset"] = r["offset"], _["estimates"] = r["estimates"],
                      _["ref.cond"] = ref, _["detection.type"] = "difference", _["model"] = "beta-binomial");
}

// src/methylation_helpers.cpp:6-72
DataFrame segment_methylation_helper(const DataFrame& df, double tol_val = 0.01) {
  IntegerVector idx = df["estimate_id"], start = df["start"];
  NumericVector beta = df["beta"], betahat = df["betahat"], pw = df["pseudoweight"];
  const int64_t n = idx.size();
  std::vector<int32_t> sid(n), sno(n);
  std::vector<uint32_t> smin(n), smax(n);
  std::vector<double> sb(n), sbh(n), spw(n), ssd(n);
  int64_t S = 0;
  check(ml_segment_methylation_helper(engine(), n, idx.begin(), start.begin(), beta.begin(), betahat.begin(),
                                      pw.begin(), tol_val, &S, sid.data(), sno.data(), smin.data(), smax.data(),
                                      sb.data(), sbh.data(), spw.data(), ssd.data()));
  return DataFrame::create(_["estimate_id"] = IntegerVector(sid.begin(), sid.begin() + S),
                           _["segment_id"] = IntegerVector(sno.begin(), sno.begin() + S),
                           _["start"] = NumericVector(smin.begin(), smin.begin() + S),
                           _["end"] = NumericVector(smax.begin(), smax.begin() + S),
                           _["beta"] = NumericVector(sb.begin(), sb.begin() + S),
                           _["betahat"] = NumericVector(sbh.begin(), sbh.begin() + S),
                           _["pseudoweight"] = NumericVector(spw.begin(), spw.begin() + S),
                           _["stdev"] = NumericVector(ssd.begin(), ssd.begin() + S));
}

// src/weighted_1d_flsa.cpp:7-21
NumericVector weighted_1d_flsa(const NumericVector& y, const NumericVector& wt, double lambda2, double lambda1 = 0) {
  if (wt.size() != y.size()) stop("Input vectors must be of same size!");
  NumericVector beta(y.size());
  check(ml_weighted_1d_flsa(engine(), y.size(), y.begin(), wt.begin(), lambda2, lambda1, beta.begin()));
  return beta;
}

// NEW: every chromosome of `data` in one native call.  job_ptr = row offsets of the chromosomes
// (data sorted by chr, dataset, pos).  Returns a list with one element per chromosome (NULL for a
// chromosome that failed, mirroring foreach's .errorhandling = "remove").
List signal_detection_fast_batch(const DataFrame& data, const NumericVector& job_ptr, double lambda2,
                                 double tol_val, unsigned max_iter, bool difference, bool want_mu) {
  std::vector<int64_t> ptr(job_ptr.begin(), job_ptr.end());
  return detect(data, ptr, fast_params(lambda2, tol_val, max_iter), difference, want_mu);
}

// NEW: the DMR caller of R/call.R:254-365 in one native call: difference_detection_fast of every chromosome, the
// per-chromosome call_differences_singlechr tables (segments of the differences, pooled data of both conditions,
// gap split, grouping, group statistics, Wilcoxon p-value, coverage score) and the FDR over the genome (:360).
// `ref` is the 1-based code of the reference condition (R/call.R:267 compares as.integer(condition) with it).
// The threshold filter of :363 stays in R (dmr_calling_gpu in r/R/genome_wide_gpu.R).
DataFrame call_differences_batch(const DataFrame& data, const NumericVector& job_ptr, int ref, double lambda2,
                                 double tol_val, double min_diff, double min_delta_diff, double max_distance) {
  need_columns(data);
  IntegerVector dset = data["dataset"], cond = data["condition"], rep = data["replicate"], pos = data["pos"],
                nmeth = data["Nmeth"], cov = data["coverage"];
  std::vector<int64_t> ptr(job_ptr.begin(), job_ptr.end());
  const int nj = (int)ptr.size() - 1;
  ml_batch* batch = nullptr;
  ml_result* res = nullptr;
  std::vector<int> status(nj);
  check(ml_batch_upload(engine(), nj, ptr.data(), dset.begin(), cond.begin(), rep.begin(), pos.begin(), nmeth.begin(),
                        cov.begin(), &batch, nullptr));
  const ml_params p = fast_params(lambda2, tol_val, 1);
  int rc = ml_batch_detect(engine(), batch, &p, &res, status.data());
  if (rc == ML_OK) rc = ml_result_fast_diff_combine(engine(), res);
  int64_t n = 0;
  if (rc == ML_OK)
    rc = ml_result_call_differences(engine(), batch, res, ref, min_diff, min_delta_diff, tol_val, max_distance, &n,
                                    0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0);
  if (rc != ML_OK) { ml_result_free(res); ml_batch_free(batch); stop(ml_last_error()); }
  IntegerVector job(n), condition(n), estimate_id(n), segment_id(n), num(n), num_ref(n);
  std::vector<uint32_t> start(n), end(n);
  NumericVector width(n), beta(n), sd(n), coverage(n), beta_ref(n), sd_ref(n), coverage_ref(n), pw(n), diff(n), cohen(n),
      pval(n), score(n), fdr(n);
  rc = ml_result_call_differences(engine(), batch, res, ref, min_diff, min_delta_diff, tol_val, max_distance, &n,
                                  job.begin(), condition.begin(), estimate_id.begin(), segment_id.begin(), start.data(),
                                  end.data(), width.begin(), beta.begin(), sd.begin(), coverage.begin(), num.begin(),
                                  beta_ref.begin(), sd_ref.begin(), coverage_ref.begin(), num_ref.begin(), pw.begin(),
                                  diff.begin(), cohen.begin(), pval.begin(), score.begin());
  if (rc == ML_OK && n > 0) rc = ml_p_adjust_fdr(engine(), n, pval.begin(), fdr.begin());     // R/call.R:360
  ml_result_free(res);
  ml_batch_free(batch);
  check(rc);
  NumericVector s(start.begin(), start.end()), e(end.begin(), end.end());
  return DataFrame::create(_["job"] = job + 1, _["condition"] = condition, _["estimate_id"] = estimate_id,
                           _["segment_id"] = segment_id, _["start"] = s, _["end"] = e, _["width"] = width, _["beta"] = beta,
                           _["std"] = sd, _["coverage"] = coverage, _["num.cpgs"] = num, _["beta.ref"] = beta_ref,
                           _["std.ref"] = sd_ref, _["coverage.ref"] = coverage_ref, _["num.cpgs.ref"] = num_ref,
                           _["pseudoweight"] = pw, _["diff"] = diff, _["cohen.d"] = cohen, _["pval"] = pval, _["fdr"] = fdr,
                           _["coverage.score"] = score);
}

// NEW: Part 1 of the CLI in one native call: signal_detection_fast of every chromosome (R/genome_wide.R:64-78) and the
// rule engine of segment_methylation_singlechr (R/call.R:31-221) on the resident fit - call table, LMR / UMR annotation,
// std fusion, valleys, PMD split / calls and the segments' PMD status - with the two result tables assembled on the
// device (ml_result_segment_methylation): only their rows cross PCIe.  `drop` is the reference's (:214-218).  Chromosomes
// are processed `chunk_jobs` at a time so that R can interrupt between chunks (Rcpp::checkUserInterrupt, as
// core/Estimator.hpp:50 and core/Posterior.ipp:17,53 do inside their loops).  `lmr` / `valley` / `pmd` are numeric
// vectors in the field order of ml_lmr_params / ml_valley_params / ml_pmd_params (include/methylasso_b200.h).
List segment_methylation_batch(const DataFrame& data, const NumericVector& job_ptr, double lambda2, double tol_val,
                               double max_distance, const NumericVector& lmr, const NumericVector& valley,
                               const NumericVector& pmd, bool drop = true, int chunk_jobs = 8) {
  need_columns(data);
  if (lmr.size() != 12 || valley.size() != 4 || pmd.size() != 5) stop("lmr / valley / pmd must hold 12 / 4 / 5 values");
  IntegerVector dset = data["dataset"], cond = data["condition"], rep = data["replicate"], pos = data["pos"],
                nmeth = data["Nmeth"], cov = data["coverage"];
  std::vector<int64_t> ptr(job_ptr.begin(), job_ptr.end());
  const int nj = (int)ptr.size() - 1;
  const ml_lmr_params lp{lmr[0], lmr[1], lmr[2], lmr[3], lmr[4], lmr[5], lmr[6], lmr[7], lmr[8], lmr[9], (int32_t)lmr[10],
                         (int32_t)lmr[11]};
  const ml_valley_params vp{valley[0], valley[1], valley[2], (int32_t)valley[3], 0};
  const ml_pmd_params pp{pmd[0], pmd[1], pmd[2], (int32_t)pmd[3], (int32_t)pmd[4]};
  const ml_params p = fast_params(lambda2, tol_val, 1);
  ml_engine_set(engine(), "want_mu", 0);          // nothing below reads mu
  struct Part {
    std::vector<int32_t> job, condition, segment_id, num_cpgs, p_job, p_condition, p_segment_id, p_num_cpgs;
    std::vector<uint32_t> start, end, p_start, p_end;
    std::vector<double> width, beta, coverage, pw, sd, p_width, p_beta, p_std, p_fused;
    std::vector<int8_t> category, status, p_category;
  };
  std::vector<Part> parts;
  if (chunk_jobs < 1) chunk_jobs = 1;
  for (int j0 = 0; j0 < nj; j0 += chunk_jobs) {
    Rcpp::checkUserInterrupt();                   // between chunks: nothing native is alive here
    const int j1 = std::min(nj, j0 + chunk_jobs), n = j1 - j0;
    std::vector<int64_t> sub(n + 1);
    for (int j = 0; j <= n; ++j) sub[j] = ptr[j0 + j] - ptr[j0];
    const int64_t r0 = ptr[j0];
    ml_batch* batch = nullptr;
    ml_result* res = nullptr;
    std::vector<int> status(n);
    auto bail = [&](int code) {
      if (code != ML_OK) {
        const std::string msg = ml_last_error();
        ml_result_free(res);
        ml_batch_free(batch);
        ml_engine_set(engine(), "want_mu", 1);
        stop(msg);
      }
    };
    bail(ml_batch_upload(engine(), n, sub.data(), dset.begin() + r0, cond.begin() + r0, rep.begin() + r0, pos.begin() + r0,
                         nmeth.begin() + r0, cov.begin() + r0, &batch, nullptr));
    bail(ml_batch_detect(engine(), batch, &p, &res, status.data()));
    int64_t ns = 0, np = 0;
    bail(ml_result_segment_methylation(engine(), res, tol_val, max_distance, &lp, &vp, &pp, drop ? 1 : 0, &ns, 0, 0, 0, 0, 0, 0, 0,
                                       0, 0, 0, 0, 0, 0, &np, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0));
    Part q;
    q.job.resize(ns); q.condition.resize(ns); q.start.resize(ns); q.end.resize(ns); q.width.resize(ns); q.segment_id.resize(ns);
    q.beta.resize(ns); q.coverage.resize(ns); q.pw.resize(ns); q.num_cpgs.resize(ns); q.sd.resize(ns); q.category.resize(ns);
    q.status.resize(ns);
    q.p_job.resize(np); q.p_condition.resize(np); q.p_start.resize(np); q.p_end.resize(np); q.p_width.resize(np);
    q.p_segment_id.resize(np); q.p_beta.resize(np); q.p_std.resize(np); q.p_category.resize(np); q.p_num_cpgs.resize(np);
    q.p_fused.resize(np);
    bail(ml_result_segment_methylation(engine(), res, tol_val, max_distance, &lp, &vp, &pp, drop ? 1 : 0, &ns, q.job.data(),
                                       q.condition.data(), q.start.data(), q.end.data(), q.width.data(), q.segment_id.data(),
                                       q.beta.data(), q.coverage.data(), q.pw.data(), q.num_cpgs.data(), q.sd.data(),
                                       q.category.data(), q.status.data(), &np, q.p_job.data(), q.p_condition.data(),
                                       q.p_start.data(), q.p_end.data(), q.p_width.data(), q.p_segment_id.data(), q.p_beta.data(),
                                       q.p_std.data(), q.p_category.data(), q.p_num_cpgs.data(), q.p_fused.data()));
    ml_result_free(res);
    ml_batch_free(batch);
    for (auto& v : q.job) v += j0;
    for (auto& v : q.p_job) v += j0;
    parts.push_back(std::move(q));
  }
  ml_engine_set(engine(), "want_mu", 1);
  int64_t nf = 0, np = 0;
  for (const Part& q : parts) { nf += (int64_t)q.job.size(); np += (int64_t)q.p_job.size(); }
  IntegerVector f_job(nf), f_cond(nf), segment_id(nf), ncpg(nf), p_job(np), p_cond(np), p_seg(np), p_ncpg(np);
  NumericVector start(nf), end(nf), width(nf), beta(nf), coverage(nf), pw(nf), sd(nf), ps(np), pe(np), p_width(np), p_beta(np),
      p_std(np), p_fused(np);
  CharacterVector category(nf), pmd_status(nf), pc(np);
  static const char* cat_names[] = {nullptr, "LMR", "UMR", "UMR-DMV"};
  int64_t a = 0, b = 0;
  for (const Part& q : parts) {
    for (size_t i = 0; i < q.job.size(); ++i, ++a) {
      f_job[a] = q.job[i] + 1; f_cond[a] = q.condition[i]; start[a] = q.start[i]; end[a] = q.end[i]; width[a] = q.width[i];
      segment_id[a] = q.segment_id[i]; beta[a] = q.beta[i]; coverage[a] = q.coverage[i]; pw[a] = q.pw[i]; ncpg[a] = q.num_cpgs[i];
      sd[a] = q.sd[i] == q.sd[i] ? q.sd[i] : NA_REAL;                       // NaN marks a valley: it has no std column
      if (q.category[i] == 0) category[a] = NA_STRING; else category[a] = cat_names[q.category[i]];
      if (q.status[i] < 0) pmd_status[a] = NA_STRING; else pmd_status[a] = q.status[i] == 1 ? "PMD" : "other";
    }
    for (size_t i = 0; i < q.p_job.size(); ++i, ++b) {
      p_job[b] = q.p_job[i] + 1; p_cond[b] = q.p_condition[i]; ps[b] = q.p_start[i]; pe[b] = q.p_end[i]; p_width[b] = q.p_width[i];
      p_seg[b] = q.p_segment_id[i]; p_beta[b] = q.p_beta[i]; p_std[b] = q.p_std[i]; p_ncpg[b] = q.p_num_cpgs[i];
      p_fused[b] = q.p_fused[i];
      pc[b] = q.p_category[i] == 1 ? "PMD" : "other";
    }
  }
  return List::create(
      _["lmr_umr_valley"] = DataFrame::create(_["job"] = f_job, _["condition"] = f_cond, _["start"] = start, _["end"] = end,
                                              _["width"] = width, _["segment_id"] = segment_id, _["beta"] = beta,
                                              _["coverage"] = coverage, _["pseudoweight"] = pw, _["num.cpgs"] = ncpg,
                                              _["std"] = sd, _["category"] = category, _["PMD.status"] = pmd_status,
                                              _["stringsAsFactors"] = false),
      _["pmd"] = DataFrame::create(_["job"] = p_job, _["condition"] = p_cond, _["start"] = ps, _["end"] = pe,
                                   _["width"] = p_width, _["segment_id"] = p_seg, _["beta"] = p_beta, _["std"] = p_std,
                                   _["category"] = pc, _["num.cpgs"] = p_ncpg, _["fused_std"] = p_fused,
                                   _["stringsAsFactors"] = false));
}

// NEW: wilcox.test(x_g, y_g, conf.int = F)$p.value for many groups (CSR offsets) and p.adjust(p, 'fdr')
NumericVector wilcoxon_groups(const NumericVector& x_ptr, const NumericVector& x, const NumericVector& y_ptr,
                              const NumericVector& y) {
  std::vector<int64_t> xp(x_ptr.begin(), x_ptr.end()), yp(y_ptr.begin(), y_ptr.end());
  NumericVector p(xp.size() - 1);
  check(ml_wilcoxon_groups(engine(), (int64_t)xp.size() - 1, xp.data(), x.begin(), yp.data(), y.begin(), p.begin(), nullptr));
  return p;
}
NumericVector p_adjust_fdr(const NumericVector& p) {
  NumericVector q(p.size());
  check(ml_p_adjust_fdr(engine(), p.size(), p.begin(), q.begin()));
  return q;
}

RCPP_MODULE(MethyLasso_cpp) {
  function("signal_detection_singlechr", &signal_detection_R,
           List::create(_["data"], _["optimize_lambda2"] = false, _["lambda2"] = 50, _["max_lambda2"] = 1000,
                        _["optimize_hyper"] = false, _["hyper"] = 100, _["nouter"] = 20, _["bf_per_kb"] = 50,
                        _["tol_val"] = 0.01, _["max_iter"] = 30));
  function("signal_detection_fast_singlechr", &signal_detection_fast_R,
           List::create(_["data"], _["lambda2"] = 50, _["tol_val"] = 0.01, _["max_iter"] = 1));
  function("segment_methylation_helper", &segment_methylation_helper, List::create(_["data"], _["tol_val"] = 0.01));
  function("difference_detection_singlechr", &difference_detection_R,
           List::create(_["data"], _["ref"], _["optimize_lambda2"] = false, _["lambda2"] = 50,
                        _["max_lambda2"] = 1000, _["optimize_hyper"] = false, _["hyper"] = 100, _["nouter"] = 20,
                        _["bf_per_kb"] = 50, _["tol_val"] = 0.01, _["max_iter"] = 30));
  function("weighted_1d_flsa", &weighted_1d_flsa, List::create(_["y"], _["wt"], _["lambda2"], _["lambda1"] = 0));
  function("signal_detection_fast_batch", &signal_detection_fast_batch);
  function("call_differences_batch", &call_differences_batch);
  function("segment_methylation_batch", &segment_methylation_batch,
           List::create(_["data"], _["job_ptr"], _["lambda2"], _["tol_val"], _["max_distance"], _["lmr"], _["valley"], _["pmd"],
                        _["drop"] = true, _["chunk_jobs"] = 8));
  function("wilcoxon_groups", &wilcoxon_groups);
  function("p_adjust_fdr", &p_adjust_fdr);
}

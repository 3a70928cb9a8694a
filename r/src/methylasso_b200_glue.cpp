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
  if (diff_fast) check(ml_result_fast_diff_combine(engine(), res));
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
    check(ml_result_copy_job(engine(), res, j, want_mu ? mu.begin() : nullptr, lambda2.begin(), offset.begin(), id.begin(),
                             start.begin(), beta.begin(), betahat.begin(), pw.begin()));
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
// std fusion, valleys, PMD split / calls and the segments' PMD status.  Returns list(lmr_umr_valley, pmd) with drop = F
// semantics (all rows; category / PMD.status are NA where the reference has NA); segment_methylation_gpu in
// r/R/genome_wide_gpu.R applies the drop of :214-218 and names the columns.  `lmr` / `valley` / `pmd` are numeric vectors
// in the field order of ml_lmr_params / ml_valley_params / ml_pmd_params (include/methylasso_b200.h).
List segment_methylation_batch(const DataFrame& data, const NumericVector& job_ptr, double lambda2, double tol_val,
                               double max_distance, const NumericVector& lmr, const NumericVector& valley,
                               const NumericVector& pmd) {
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
  ml_batch* batch = nullptr;
  ml_result* res = nullptr;
  std::vector<int> status(nj);
  check(ml_batch_upload(engine(), nj, ptr.data(), dset.begin(), cond.begin(), rep.begin(), pos.begin(), nmeth.begin(),
                        cov.begin(), &batch, nullptr));
  const ml_params p = fast_params(lambda2, tol_val, 1);
  int rc = ml_batch_detect(engine(), batch, &p, &res, status.data());
  auto bail = [&](int code) { if (code != ML_OK) { ml_result_free(res); ml_batch_free(batch); stop(ml_last_error()); } };
  bail(rc);
  int64_t nc = 0, nv = 0, nr = 0, nm = 0, npc = 0, np = 0, nf = 0;
  // call table (:43-72), valleys (:100-133), merged table (:131-140), PMD table (:169-199), final rows (:201-215)
  bail(ml_result_call_table(engine(), res, tol_val, max_distance, &nc, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0));
  NumericVector c_width(nc), c_beta(nc), c_cov(nc), c_pw(nc), c_std(nc);
  IntegerVector c_ncpg(nc);
  bail(ml_result_call_table(engine(), res, tol_val, max_distance, &nc, 0, 0, 0, 0, 0, c_width.begin(), c_beta.begin(),
                            c_cov.begin(), c_pw.begin(), c_ncpg.begin(), c_std.begin()));
  bail(ml_result_call_valleys(engine(), res, tol_val, max_distance, &lp, &vp, &nv, 0, 0, 0, 0, 0, 0, 0, 0, 0, &nr, 0, 0, 0));
  NumericVector v_width(nv), v_beta(nv), v_cov(nv), v_pw(nv);
  IntegerVector v_ncpg(nv);
  bail(ml_result_call_valleys(engine(), res, tol_val, max_distance, &lp, &vp, &nv, 0, 0, 0, 0, v_width.begin(), v_beta.begin(),
                              v_cov.begin(), v_pw.begin(), v_ncpg.begin(), &nr, 0, 0, 0));
  bail(ml_result_call_pmd_split(engine(), res, tol_val, max_distance, pp.pmd_lambda2, &lp, &vp, pp.split_pmds, &nm, 0, 0, 0, 0, 0,
                                0, 0, 0, 0, &npc, 0, 0, 0, 0, 0, 0, 0));
  IntegerVector m_seg(nm);
  bail(ml_result_call_pmd_split(engine(), res, tol_val, max_distance, pp.pmd_lambda2, &lp, &vp, pp.split_pmds, &nm, 0, 0,
                                m_seg.begin(), 0, 0, 0, 0, 0, 0, &npc, 0, 0, 0, 0, 0, 0, 0));
  bail(ml_result_call_pmds(engine(), res, tol_val, max_distance, &lp, &vp, &pp, &np, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0));
  IntegerVector p_job(np), p_cond(np), p_seg(np), p_ncpg(np);
  std::vector<int8_t> p_cat(np);
  std::vector<uint32_t> p_start(np), p_end(np);
  NumericVector p_width(np), p_fused(np), p_beta(np), p_std(np);
  bail(ml_result_call_pmds(engine(), res, tol_val, max_distance, &lp, &vp, &pp, &np, p_job.begin(), p_cond.begin(), p_cat.data(),
                           p_seg.begin(), p_start.data(), p_end.data(), p_width.begin(), p_ncpg.begin(), p_fused.begin(),
                           p_beta.begin(), p_std.begin()));
  bail(ml_result_call_pmd_status(engine(), res, tol_val, max_distance, &lp, &vp, &pp, &nf, 0, 0, 0, 0, 0, 0, 0, 0, 0));
  IntegerVector f_job(nf), f_cond(nf), f_merged(nf), f_row(nf), f_val(nf);
  std::vector<int8_t> f_cat(nf), f_status(nf);
  std::vector<uint32_t> f_start(nf), f_end(nf);
  rc = ml_result_call_pmd_status(engine(), res, tol_val, max_distance, &lp, &vp, &pp, &nf, f_job.begin(), f_cond.begin(),
                                 f_start.data(), f_end.data(), f_cat.data(), f_status.data(), f_merged.begin(), f_row.begin(),
                                 f_val.begin());
  ml_result_free(res);
  ml_batch_free(batch);
  check(rc);
  // the value columns of a final row come from its call-table row or from its valley (std is NA on a valley)
  NumericVector start(nf), end(nf), width(nf), beta(nf), coverage(nf), pw(nf), sd(nf);
  IntegerVector segment_id(nf), ncpg(nf);
  CharacterVector category(nf), pmd_status(nf);
  static const char* cat_names[] = {nullptr, "LMR", "UMR", "UMR-DMV"};
  for (int64_t i = 0; i < nf; ++i) {
    const int r = f_row[i], v = f_val[i];
    start[i] = f_start[i]; end[i] = f_end[i]; segment_id[i] = m_seg[f_merged[i]];
    width[i] = r >= 0 ? c_width[r] : v_width[v];
    beta[i] = r >= 0 ? c_beta[r] : v_beta[v];
    coverage[i] = r >= 0 ? c_cov[r] : v_cov[v];
    pw[i] = r >= 0 ? c_pw[r] : v_pw[v];
    ncpg[i] = r >= 0 ? c_ncpg[r] : v_ncpg[v];
    sd[i] = r >= 0 ? c_std[r] : NA_REAL;
    if (f_cat[i] == 0) category[i] = NA_STRING; else category[i] = cat_names[f_cat[i]];
    if (f_status[i] < 0) pmd_status[i] = NA_STRING; else pmd_status[i] = f_status[i] == 1 ? "PMD" : "other";
  }
  NumericVector ps(p_start.begin(), p_start.end()), pe(p_end.begin(), p_end.end());
  CharacterVector pc(np);
  for (int64_t i = 0; i < np; ++i) pc[i] = p_cat[i] == 1 ? "PMD" : "other";
  return List::create(
      _["lmr_umr_valley"] = DataFrame::create(_["job"] = f_job + 1, _["condition"] = f_cond, _["start"] = start, _["end"] = end,
                                              _["width"] = width, _["segment_id"] = segment_id, _["beta"] = beta,
                                              _["coverage"] = coverage, _["pseudoweight"] = pw, _["num.cpgs"] = ncpg,
                                              _["std"] = sd, _["category"] = category, _["PMD.status"] = pmd_status,
                                              _["stringsAsFactors"] = false),
      _["pmd"] = DataFrame::create(_["job"] = p_job + 1, _["condition"] = p_cond, _["start"] = ps, _["end"] = pe,
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
  function("segment_methylation_batch", &segment_methylation_batch);
  function("wilcoxon_groups", &wilcoxon_groups);
  function("p_adjust_fdr", &p_adjust_fdr);
}

// oracle/ref_driver.cpp — TEST INFRASTRUCTURE ONLY.
//
// C entry points around the reference's own Rcpp-module functions (src/MethyLasso_cpp.cpp:10-31),
// linked with the reference's unmodified sources compiled IN PLACE from /root/reference/src against
// the stub Rcpp / Eigen headers in oracle/stubs (see oracle/Makefile, target ref_full).  The result,
// oracle/_ref/libmethylasso_ref.so, is what pins the plain-C restatement above the DP kernel
// (tests/test_oracle_pin.py).  Nothing in the product links or loads it.
#include <Rcpp.h>
#include <cstring>
#include <string>

#include "signal_detection.hpp"       // reference: src/signal_detection.hpp
#include "difference_detection.hpp"   // reference: src/difference_detection.hpp
#include "methylation_helpers.hpp"    // reference: src/methylation_helpers.hpp
#include "weighted_1d_flsa.hpp"       // reference: src/weighted_1d_flsa.hpp

namespace {
std::string g_error;
Rcpp::List g_result;          // last detection result
Rcpp::DataFrame g_segments;   // last segment table

Rcpp::DataFrame make_data(long n, const int* dataset, const int* condition, const int* replicate,
                          const int* pos, const int* nmeth, const int* coverage) {
  using Rcpp::_;
  std::vector<int> d(dataset, dataset + n), c(condition, condition + n), r(replicate, replicate + n),
      p(pos, pos + n), m(nmeth, nmeth + n), v(coverage, coverage + n);
  return Rcpp::DataFrame::create(_["dataset"] = d, _["condition"] = c, _["replicate"] = r, _["pos"] = p,
                                 _["Nmeth"] = m, _["coverage"] = v);
}
long column_len(const Rcpp::RObject& o) { return (long)(o->kind == Rcpp::Value::REAL ? o->dv.size() : o->iv.size()); }
void copy_real(const Rcpp::RObject& o, double* out) {
  if (!out) return;
  if (o->kind == Rcpp::Value::REAL) for (size_t i = 0; i < o->dv.size(); ++i) out[i] = o->dv[i];
  else for (size_t i = 0; i < o->iv.size(); ++i) out[i] = o->iv[i];
}
template <class F> int guarded(F&& f) {
  try { f(); g_error.clear(); return 0; }
  catch (const std::exception& e) { g_error = e.what(); return 1; }
  catch (...) { g_error = "unknown exception"; return 2; }
}
}  // namespace

extern "C" {

const char* ref_last_error() { return g_error.c_str(); }

// kind: 0 = signal_detection_fast_singlechr, 1 = signal_detection_singlechr, 2 = difference_detection_singlechr
int ref_detect(int kind, long n, const int* dataset, const int* condition, const int* replicate, const int* pos,
               const int* nmeth, const int* coverage, double lambda2, double max_lambda2, double hyper,
               unsigned nouter, double bf_per_kb, double tol_val, unsigned max_iter, unsigned ref,
               int optimize_lambda2, int optimize_hyper) {
  return guarded([&] {
    Rcpp::DataFrame data = make_data(n, dataset, condition, replicate, pos, nmeth, coverage);
    if (kind == 0) g_result = MethyLasso::signal_detection_fast_R(data, lambda2, tol_val, max_iter);
    else if (kind == 1) g_result = MethyLasso::signal_detection_R(data, optimize_lambda2 != 0, lambda2, max_lambda2,
                                                                 optimize_hyper != 0, hyper, nouter, bf_per_kb, tol_val, max_iter);
    else g_result = MethyLasso::difference_detection_R(data, ref, optimize_lambda2 != 0, lambda2, max_lambda2,
                                                      optimize_hyper != 0, hyper, nouter, bf_per_kb, tol_val, max_iter);
  });
}
// sizes of the last result: rows of mu, rows of the estimates table, number of lambda2 values, number of offsets
int ref_result_sizes(long* n_mu, long* n_est_rows, long* n_lambda2, long* n_offset) {
  return guarded([&] {
    *n_mu = column_len(g_result["mu"].get());
    Rcpp::DataFrame est = Rcpp::as<Rcpp::DataFrame>(g_result["estimates"].get());
    *n_est_rows = est.nrows();
    *n_lambda2 = column_len(g_result["lambda2"].get());
    *n_offset = column_len(g_result["offset"].get());
  });
}
int ref_result_copy(double* mu, double* lambda2, double* offset, int* converged, double* est_id, double* start,
                    double* beta, double* betahat, double* pseudoweight) {
  return guarded([&] {
    copy_real(g_result["mu"].get(), mu);
    copy_real(g_result["lambda2"].get(), lambda2);
    copy_real(g_result["offset"].get(), offset);
    if (converged) *converged = Rcpp::as<bool>(g_result["converged"].get()) ? 1 : 0;
    Rcpp::DataFrame est = Rcpp::as<Rcpp::DataFrame>(g_result["estimates"].get());
    copy_real(est["estimate_id"].get(), est_id);
    copy_real(est["start"].get(), start);
    copy_real(est["beta"].get(), beta);
    copy_real(est["betahat"].get(), betahat);
    copy_real(est["pseudoweight"].get(), pseudoweight);
  });
}

int ref_segment_helper(long n, const int* est_id, const int* start, const double* beta, const double* betahat,
                       const double* pseudoweight, double tol_val, long* n_segments) {
  return guarded([&] {
    using Rcpp::_;
    std::vector<int> id(est_id, est_id + n), st(start, start + n);
    std::vector<double> b(beta, beta + n), bh(betahat, betahat + n), pw(pseudoweight, pseudoweight + n);
    Rcpp::DataFrame df = Rcpp::DataFrame::create(_["estimate_id"] = id, _["start"] = st, _["beta"] = b,
                                                 _["betahat"] = bh, _["pseudoweight"] = pw);
    g_segments = MethyLasso::segment_methylation_helper(df, tol_val);
    *n_segments = g_segments.nrows();
  });
}
int ref_segments_copy(double* est_id, double* seg_id, double* start, double* end, double* beta, double* betahat,
                      double* pseudoweight, double* stdev) {
  return guarded([&] {
    copy_real(g_segments["estimate_id"].get(), est_id);
    copy_real(g_segments["segment_id"].get(), seg_id);
    copy_real(g_segments["start"].get(), start);
    copy_real(g_segments["end"].get(), end);
    copy_real(g_segments["beta"].get(), beta);
    copy_real(g_segments["betahat"].get(), betahat);
    copy_real(g_segments["pseudoweight"].get(), pseudoweight);
    copy_real(g_segments["stdev"].get(), stdev);
  });
}

int ref_weighted_1d_flsa(long n, const double* y, const double* wt, double lambda2, double lambda1, double* out) {
  return guarded([&] {
    Eigen::VectorXd yy(n), ww(n);
    for (long i = 0; i < n; ++i) { yy(i) = y[i]; ww(i) = wt[i]; }
    Eigen::VectorXd b = MethyLasso::weighted_1d_flsa(yy, ww, lambda2, lambda1);
    for (long i = 0; i < n; ++i) out[i] = b(i);
  });
}

}  // extern "C"

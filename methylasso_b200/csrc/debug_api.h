/* debug_api.h — entry points for the test suite only; NOT part of the drop-in surface (include/methylasso_b200.h).
 * They are exported from the shared library so that tests can reach the solver's internals through ctypes. */
#pragma once
#include "../../include/methylasso_b200.h"
#ifdef __cplusplus
extern "C" {
#endif
/* one series through the FLSA solver, also returning the DP back-pointers (gfl_tf.c: tm / tp) */
ML_API int ml_debug_flsa_backpointers(ml_engine *eng, int64_t n, const double *y, const double *wt, double lambda2,
                                      double *beta, double *tm, double *tp);
/* HOST ONLY (no GPU, no engine): the run table ml_batch_upload derives from the factor columns of ONE job - the runs of
 * equal dataset codes that lie before the first structural offence (first row, condition, replicate of each; at most
 * cap are written), the code of that offence (0: none) and its row. */
ML_API int ml_debug_scan_runs(int64_t n, const int32_t *dataset, const int32_t *condition, const int32_t *replicate,
                              const int32_t *pos, int threads, int64_t cap, int64_t *n_runs, int64_t *first_row,
                              int32_t *run_condition, int32_t *run_replicate, int32_t *code, int64_t *code_row);
/* counters of the engine's last FLSA solve: chunks verified, chunks redone by the walkers, longest chain, chains, series the
 * walkers handed over to the scan route, chunks in the scan route's tables, series finished by the start-to-end kernel */
ML_API int ml_debug_flsa_last_stats(ml_engine *eng, int *out7);
/* counts_div.cuh on the device: every pair 0 <= num <= max_num, 1 <= den < max_den through div_counts (both signatures) and
 * through the device's IEEE division; *mismatches = pairs whose bits differ (must be 0). */
ML_API int ml_debug_div_counts(ml_engine *eng, int max_num, int max_den, long long *pairs, long long *mismatches);
#ifdef __cplusplus
}
#endif

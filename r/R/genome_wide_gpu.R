# genome_wide_gpu.R — replacement for the chromosome loop of R/genome_wide.R:64-78 and 121-135.
# The reference forks one worker per chromosome (foreach %dopar%); a CUDA context cannot be used in
# forked children, so the native call takes all chromosomes at once and `ncores` is ignored.
# chromosomes in their order of first appearance (data[, unique(chr)], as the reference's loops take them), the rows of a
# chromosome together and otherwise in the caller's order; the caller's table is not touched (no setkey on it)
.by_chromosome = function(data) {
  chrs = data[, unique(chr)]
  d = data[order(match(chr, chrs))]              # order() is stable: rows of a chromosome keep their order
  list(data = d, chrs = chrs, job_ptr = c(0, cumsum(d[, .N, by = chr]$N)))
}
# factor codes -> labels, integer codes stay (the reference documents 'integer vectors or factors')
.condition_labels = function(data, codes) if (is.factor(data$condition)) levels(data$condition)[codes] else codes

signal_detection_fast = function(data, lambda2 = 25, tol_val = 0.01, max_iter = 1, ncores = 1, verbose = TRUE, want.mu = TRUE) {
  bc = .by_chromosome(data); data = bc$data; chrs = bc$chrs; job_ptr = bc$job_ptr
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
  bc = .by_chromosome(data); data = bc$data; chrs = bc$chrs; job_ptr = bc$job_ptr
  ref.code = if (is.character(ref) || is.factor(ref)) match(ref, levels(data$condition)) else as.integer(ref)
  diff_call = as.data.table(MethyLasso:::call_differences_batch(data, job_ptr, ref.code, lambda2, tol_val, min.diff,
                                                                 min.delta.diff, max_distance))
  diff_call[, chr := chrs[job]]
  diff_call[, ref := ref]
  diff_call[, condition := .condition_labels(data, condition)]
  diff_call[(!is.na(diff)) & abs(diff) >= min.diff & coverage.score >= cov.score & pval <= pval.cutoff &
            fdr <= fdr.cutoff & pmin(num.cpgs, num.cpgs.ref) >= min.num.cpgs]
}

# Replacement for Part 1 of the CLI (signal_detection_fast followed by segment_methylation, R/call.R:229-244 around
# :31-221): the fit and the whole rule engine run on the GPU for all chromosomes at once; only the two result tables come
# back (`ov`, the per-position join, is not produced).  Arguments and defaults are those of segment_methylation_singlechr.
segment_methylation_gpu = function(data, lambda2 = 25, pmd_lambda2 = 1000, pmd_std_threshold = 0.15, umr_std_threshold = 0.1,
                                   umr_max_beta = 0.1, lmr_max_beta = 0.5, pmd_max_beta = 0.75, valley_max_beta = 0.1,
                                   pmd_valley_min_width = 5000, max_distance = 500, min_num_cpgs = 4, min_width = 30,
                                   flank.width = 300, flank.dist.max = 5000, flank.num.cpgs = 10, flank.beta.min = 0.5,
                                   umr_large_width = 500, umr_large_density = 0.03, split.pmds = T, merge.pmds = T, drop = T,
                                   tol_val = 0.01) {
  bc = .by_chromosome(data); data = bc$data; chrs = bc$chrs; job_ptr = bc$job_ptr
  lmr = c(min_width, max_distance, flank.width, flank.dist.max, lmr_max_beta, flank.beta.min, umr_max_beta, umr_std_threshold,
          umr_large_width, umr_large_density, min_num_cpgs, flank.num.cpgs)
  valley = c(valley_max_beta, pmd_valley_min_width, max_distance, min_num_cpgs)
  pmd = c(pmd_lambda2, pmd_std_threshold, pmd_max_beta, as.integer(split.pmds), as.integer(merge.pmds))
  # the drop of R/call.R:214-218 (rows with a category, "PMD" rows) is applied on the device; pmd keeps its num.cpgs and
  # fused_std columns as the reference's pmd_call does (:216 removes fused_std from seg_call only)
  segments = MethyLasso:::segment_methylation_batch(data, job_ptr, lambda2, tol_val, max_distance, lmr, valley, pmd, drop == T)
  segments = lapply(segments, function(x) { x = as.data.table(x); x[, chr := chrs[job]]; x[, job := NULL]
                                            cl = .condition_labels(data, x$condition); x[, condition := cl]; x })
  setcolorder(segments$lmr_umr_valley, c("condition", "chr", "start", "end", "width", "segment_id", "beta", "coverage",
                                         "pseudoweight", "num.cpgs", "std", "category", "PMD.status"))
  setcolorder(segments$pmd, c("condition", "chr", "start", "end", "width", "segment_id", "beta", "std", "category"))
  segments
}

# genome_wide_gpu.R — replacement for the chromosome loop of R/genome_wide.R:64-78 and 121-135.
# The reference forks one worker per chromosome (foreach %dopar%); a CUDA context cannot be used in
# forked children, so the native call takes all chromosomes at once and `ncores` is ignored.
signal_detection_fast = function(data, lambda2 = 25, tol_val = 0.01, max_iter = 1, ncores = 1, verbose = TRUE) {
  setkey(data, chr, dataset, pos)
  chrs = data[, unique(chr)]
  job_ptr = c(0, data[, .N, keyby = chr][, cumsum(N)])
  rets = MethyLasso:::signal_detection_fast_batch(data, job_ptr, lambda2, tol_val, max_iter, FALSE)
  ok = !sapply(rets, is.null)
  if (any(!ok) && verbose) cat("Warning: the following chromosomes did not work:", chrs[!ok])
  rets = Map(function(r, ch) { r$chr = ch; r$estimates$chr = ch
                               r$detection.type = "signal"; r$model = "normal"; r }, rets[ok], chrs[ok])
  Reduce(MethyLasso:::combine_ret, rets)
}

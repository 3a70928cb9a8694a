// regions.inc.cuh — the two tables segment_methylation returns (R/call.R:214-220), built on the device from the
// resident tables of the rule engine: `lmr_umr_valley` (seg_call with its PMD.status, :201-215; with drop = T only the
// rows that carry a category) and `pmd` (pmd_call, :169-199; with drop = T only the "PMD" rows), each with the
// reference's columns (:241-243).  Only these final rows cross PCIe.  Included at the end of segment.cu.
namespace {

__global__ void k_rg_flag(int32_t n, const int8_t* __restrict__ category, int drop, int want, int32_t* __restrict__ flag) {
  const int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) flag[i] = !drop || (want ? category[i] == want : category[i] != 0);
}

struct RgSegIn {
  const int32_t *merged, *job, *est, *src_row, *src_valley;      // FinalDev
  const uint32_t *start, *end;
  const int8_t *status, *category;
  const int32_t* m_segment_id;                                   // MergedDev
  const double *c_width, *c_beta, *c_coverage, *c_pw, *c_std;    // CallTable
  const int32_t* c_num_cpgs;
  const double *v_width, *v_beta, *v_coverage, *v_pw;            // ValleyTable
  const int32_t* v_num_cpgs;
};
struct RgSegOut {
  int32_t *job, *condition, *segment_id, *num_cpgs;
  uint32_t *start, *end;
  double *width, *beta, *coverage, *pseudoweight, *std_;
  int8_t *category, *pmd_status;
};
__global__ void k_rg_gather_seg(int32_t n_out, const int32_t* __restrict__ list, RgSegIn I, RgSegOut O) {
  const int32_t o = blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= n_out) return;
  const int32_t i = list[o];
  const int32_t r = I.src_row[i], v = I.src_valley[i];
  O.job[o] = I.job[i];
  O.condition[o] = I.est[i];
  O.start[o] = I.start[i];
  O.end[o] = I.end[i];
  O.segment_id[o] = I.m_segment_id[I.merged[i]];
  O.category[o] = I.category[i];
  O.pmd_status[o] = I.status[i];
  if (r >= 0) {          // a row of the call table
    O.width[o] = I.c_width[r]; O.beta[o] = I.c_beta[r]; O.coverage[o] = I.c_coverage[r];
    O.pseudoweight[o] = I.c_pw[r]; O.num_cpgs[o] = I.c_num_cpgs[r]; O.std_[o] = I.c_std[r];
  } else {               // a valley: it has no std column (NA after rbind(fill = T), R/call.R:135)
    O.width[o] = I.v_width[v]; O.beta[o] = I.v_beta[v]; O.coverage[o] = I.v_coverage[v];
    O.pseudoweight[o] = I.v_pw[v]; O.num_cpgs[o] = I.v_num_cpgs[v];
    O.std_[o] = __longlong_as_double(0x7ff8000000000000ll);
  }
}

struct RgPmdIn {
  const int32_t *job, *est, *segment_id, *num_cpgs;
  const int8_t* category;
  const uint32_t *start, *end;
  const double *width, *fused_std, *beta, *std_;
};
struct RgPmdOut {
  int32_t *job, *condition, *segment_id, *num_cpgs;
  int8_t* category;
  uint32_t *start, *end;
  double *width, *fused_std, *beta, *std_;
};
__global__ void k_rg_gather_pmd(int32_t n_out, const int32_t* __restrict__ list, RgPmdIn I, RgPmdOut O) {
  const int32_t o = blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= n_out) return;
  const int32_t i = list[o];
  O.job[o] = I.job[i]; O.condition[o] = I.est[i]; O.segment_id[o] = I.segment_id[i]; O.num_cpgs[o] = I.num_cpgs[i];
  O.category[o] = I.category[i]; O.start[o] = I.start[i]; O.end[o] = I.end[i]; O.width[o] = I.width[i];
  O.fused_std[o] = I.fused_std[i]; O.beta[o] = I.beta[i]; O.std_[o] = I.std_[i];
}

}  // namespace

void region_tables_device(Engine& eng, const CallTable& ct, const ValleyTable& val, const MergedDev& M, const PmdCallDev& pmd,
                          const FinalDev& F, int drop, RegionTables& out) {
  cudaStream_t st = eng.stream;
  out = RegionTables();
  Scan scan;
  const int32_t nf = (int32_t)F.n, np = (int32_t)pmd.n;
  DevBuf<int32_t> flag(st, (size_t)std::max(1, std::max(nf, np))), list(st, (size_t)std::max(1, std::max(nf, np)));
  int32_t ns = 0;
  if (nf > 0) {
    ML_LAUNCH(eng, "k_rg_flag", k_rg_flag<<<cdiv(nf, 256), 256, 0, st>>>(nf, F.category.p, drop, 0, flag.p));
    ns = scan.run(eng, flag.p, nf, nullptr, list.p);
  }
  out.n_seg = ns;
  {
    const size_t c = (size_t)std::max(1, ns);
    out.s_job.alloc(st, c); out.s_condition.alloc(st, c); out.s_segment_id.alloc(st, c); out.s_num_cpgs.alloc(st, c);
    out.s_start.alloc(st, c); out.s_end.alloc(st, c);
    out.s_width.alloc(st, c); out.s_beta.alloc(st, c); out.s_coverage.alloc(st, c); out.s_pseudoweight.alloc(st, c); out.s_std.alloc(st, c);
    out.s_category.alloc(st, c); out.s_pmd_status.alloc(st, c);
  }
  if (ns > 0) {
    const RgSegIn I{F.merged.p, F.job.p, F.est.p, F.src_row.p, F.src_valley.p, F.start.p, F.end.p, F.status.p, F.category.p,
                    M.segment_id.p, ct.width.p, ct.t.beta.p, ct.coverage.p, ct.t.pseudoweight.p, ct.t.stdev.p, ct.num_cpgs.p,
                    val.width.p, val.beta.p, val.coverage.p, val.pseudoweight.p, val.num_cpgs.p};
    const RgSegOut O{out.s_job.p, out.s_condition.p, out.s_segment_id.p, out.s_num_cpgs.p, out.s_start.p, out.s_end.p,
                     out.s_width.p, out.s_beta.p, out.s_coverage.p, out.s_pseudoweight.p, out.s_std.p, out.s_category.p,
                     out.s_pmd_status.p};
    ML_LAUNCH(eng, "k_rg_gather_seg", k_rg_gather_seg<<<cdiv(ns, 256), 256, 0, st>>>(ns, list.p, I, O));
  }
  int32_t npk = 0;
  if (np > 0) {
    ML_LAUNCH(eng, "k_rg_flag", k_rg_flag<<<cdiv(np, 256), 256, 0, st>>>(np, pmd.category.p, drop, 1, flag.p));
    npk = scan.run(eng, flag.p, np, nullptr, list.p);
  }
  out.n_pmd = npk;
  {
    const size_t c = (size_t)std::max(1, npk);
    out.p_job.alloc(st, c); out.p_condition.alloc(st, c); out.p_segment_id.alloc(st, c); out.p_num_cpgs.alloc(st, c);
    out.p_category.alloc(st, c); out.p_start.alloc(st, c); out.p_end.alloc(st, c);
    out.p_width.alloc(st, c); out.p_fused_std.alloc(st, c); out.p_beta.alloc(st, c); out.p_std.alloc(st, c);
  }
  if (npk > 0) {
    const RgPmdIn I{pmd.job.p, pmd.est.p, pmd.segment_id.p, pmd.num_cpgs.p, pmd.category.p, pmd.start.p, pmd.end.p,
                    pmd.width.p, pmd.fused_std.p, pmd.beta.p, pmd.std_.p};
    const RgPmdOut O{out.p_job.p, out.p_condition.p, out.p_segment_id.p, out.p_num_cpgs.p, out.p_category.p, out.p_start.p,
                     out.p_end.p, out.p_width.p, out.p_fused_std.p, out.p_beta.p, out.p_std.p};
    ML_LAUNCH(eng, "k_rg_gather_pmd", k_rg_gather_pmd<<<cdiv(npk, 256), 256, 0, st>>>(npk, list.p, I, O));
  }
}

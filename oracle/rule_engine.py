"""TEST INFRASTRUCTURE - the rule engine of segment_methylation_singlechr (R/call.R:31-221) for one chromosome of a
"signal" fit, as the chain of the CPU restatements: call table (oracle.call_table, :43-72), LMR / UMR annotation
(lmr_umr, :75-91), std fusion and its segments (:93-95), valleys (valleys, :100-133), merged table / regions / split_call
(pmd_split, :131-168), PMD calls (pmd_calls, :169-199), PMD status (pmd_status, :201-215), drop (:214-218).
R is not in the image: NOT pinned against R.  Only tests/ may import this module."""
import numpy as np

from . import lmr_umr as LU, oracle as O, pmd_calls as PC, pmd_split as PS, pmd_status as ST, valleys as V


def segment_methylation_singlechr(pooled, estimates, pmd_lambda2=1000.0, tol_val=0.01, max_distance=500.0, split_pmds=True,
                                  merge_pmds=True, pmd_std_threshold=0.15, pmd_max_beta=0.75, valley_max_beta=0.1,
                                  pmd_valley_min_width=5000.0, min_num_cpgs=4, **lmr_kw):
    """pooled: estimate_id, pos, Nmeth, coverage per (condition, position) (mdata after :43-46); estimates: the fit's
    estimates table.  Returns dict(lmr_umr_valley, pmd) with drop = T, plus the undropped tables under "all"."""
    seg = O.segment_methylation_helper(estimates, tol_val)
    call = O.call_table(pooled, seg, max_distance)
    lu = LU.annotate(call, max_distance=max_distance, min_num_cpgs=min_num_cpgs, **lmr_kw)
    # :93-95  fused std per condition, then the helper on (estimate_id, start, fused_std, std, pseudoweight)
    fused = np.zeros(call["std"].size)
    for c in np.unique(call["estimate_id"]):
        m = call["estimate_id"] == c
        fused[m] = O.weighted_1d_flsa(call["std"][m], call["width"][m], pmd_lambda2)
    std_call = O.segment_methylation_helper(dict(estimate_id=call["estimate_id"], start=call["start"].astype(np.int32), beta=fused,
                                                 betahat=call["std"], pseudoweight=call["pseudoweight"]), tol_val)
    std_call = dict(estimate_id=std_call["estimate_id"], start=std_call["start"].astype(np.float64),
                    end=std_call["end"].astype(np.float64), fused_std=std_call["beta"], pseudoweight=std_call["pseudoweight"])
    val = V.valleys(call, lu["category"], valley_max_beta=valley_max_beta, pmd_valley_min_width=pmd_valley_min_width,
                    max_distance=max_distance, min_num_cpgs=min_num_cpgs)
    M = PS.merged_table(call, lu["category"], val, pmd_valley_min_width)
    split = PS.split_call(PS.regions(M, split_pmds), std_call)
    pmd = PC.pmd_calls(pooled, split, pmd_std_threshold=pmd_std_threshold, pmd_max_beta=pmd_max_beta,
                       pmd_valley_min_width=pmd_valley_min_width, max_distance=max_distance, merge_pmds=merge_pmds)
    fin = ST.annotate(M, lu["is_admissible"], pmd)
    # columns of the final table: call-table values on segments, valley values on valleys (std NA there)
    m = fin["merged"]
    src, sv = M["src_row"][m], M["src_valley"][m]
    is_row = src >= 0
    v = val["valleys"]

    def col(ck, vk, dtype=np.float64):
        out = np.full(m.size, np.nan) if dtype == np.float64 else np.zeros(m.size, dtype)
        out[is_row] = np.asarray(call[ck])[src[is_row]]
        if vk is not None:
            out[~is_row] = np.asarray(v[vk])[sv[~is_row]]
        return out
    table = dict(condition=M["condition"][m], start=M["start"][m], end=M["end"][m], segment_id=M["segment_id"][m],
                 width=col("width", "width"), beta=col("beta", "beta"), coverage=col("coverage", "coverage"),
                 pseudoweight=col("pseudoweight", "pseudoweight"), num_cpgs=col("num_cpgs", "num_cpgs", np.int64),
                 std=col("std", None), category=fin["category"], pmd_status=fin["status"])
    keep = table["category"] != 0
    pk = pmd["category"] == 1
    return dict(lmr_umr_valley={k: a[keep] for k, a in table.items()}, pmd={k: a[pk] for k, a in pmd.items()},
                all=dict(lmr_umr_valley=table, pmd=pmd))

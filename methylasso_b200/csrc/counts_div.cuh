// counts_div.cuh — Nmeth / coverage without the division subroutine.
//
// The quotient of a count by a coverage is formed once or twice per row by every sweep of the engine; the device's IEEE
// fp64 division is a ~30-instruction sequence and those sweeps are bound by instruction issue as much as by HBM (ncu:
// 150-170 instructions per row, issue slots 50-75 % busy).  For an INTEGER numerator 0 <= n <= 65536 and an integer
// denominator 1 <= c < 1024, with r = RN(1 / c):
//     q0 = RN(n r),   e = n - q0 c  (exact: one FMA),   q = RN(q0 + e r)  (one FMA)
// is the correctly rounded quotient RN(n / c) - the same bits as the division, verified for every such pair
// (tests/test_host_logic.py runs the loop on the CPU with the C library's fma; tests/test_gpu_detect.py compares the
// device function with the device's division).  r comes from a table filled once per device by that very division.
// Anything else (a coverage of 1024 or more, a non-integer operand) takes the division as before.
#pragma once
#include "common.h"
#include "hd.h"

namespace ml {

constexpr int RCP_TABLE = 1024;
static __device__ double g_rcp[RCP_TABLE];     // one copy per translation unit that includes this header
static __global__ void k_init_rcp() {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < RCP_TABLE) g_rcp[i] = 1.0 / (double)i;          // [0] = +inf, never read
}
static inline void init_rcp_table(cudaStream_t st) {
  k_init_rcp<<<(RCP_TABLE + 255) / 256, 256, 0, st>>>();
  ML_CHECK_LAUNCH();
}

// num integer-valued (the caller's promise), den any double
__device__ __forceinline__ double div_counts(double num, double den) {
  const int c = (int)den;
  if ((unsigned)(c - 1) < (unsigned)(RCP_TABLE - 1) && (double)c == den && num >= 0.0 && num <= 65536.0) {
    const double r = g_rcp[c];
    const double q0 = num * r;
    return fma(fma(-q0, den, num), r, q0);
  }
  return div_nz(num, den);
}
__device__ __forceinline__ double div_counts(int num, int den) {
  if ((unsigned)(den - 1) < (unsigned)(RCP_TABLE - 1) && (unsigned)num <= 65536u) {
    const double r = g_rcp[den], n = (double)num, c = (double)den;
    const double q0 = n * r;
    return fma(fma(-q0, c, n), r, q0);
  }
  return div_nz((double)num, (double)den);
}

}  // namespace ml

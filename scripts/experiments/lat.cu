// dependent-issue latency of a few instruction kinds on this GPU (one warp): cycles per instruction in a dependent chain
#include <cstdio>
#include <cuda_runtime.h>
template <int KIND>
__global__ void k(double* out, double a, double b, int n, long long* cyc) {
  __shared__ double sm[64];
  sm[threadIdx.x & 63] = a + threadIdx.x;
  __syncthreads();
  double x = a + threadIdx.x, y = b;
  int idx = threadIdx.x & 63;
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) {
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      if (KIND == 0) x = fma(x, y, b);            // DFMA
      if (KIND == 1) x = x + y;                   // DADD
      if (KIND == 2) x = x * y;                   // DMUL
      if (KIND == 3) x = x / y;                   // IEEE division
      if (KIND == 4) { idx = (int)sm[idx & 63] & 63; }                 // LDS.64 + convert (pointer chase)
      if (KIND == 5) { x = (x > y) ? x + b : x - b; }                  // DSETP + select + DADD
      if (KIND == 6) { float f = (float)x; f = fmaf(f, 1.0001f, 0.5f); x = f; }
    }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) { cyc[KIND] = (t1 - t0); }
  out[threadIdx.x] = x + idx;
}
int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 1024 * 8); cudaMallocManaged(&cyc, 64);
  const int n = 2000;
  k<0><<<1, 32>>>(out, 1.0000001, 0.9999999, n, cyc);
  k<1><<<1, 32>>>(out, 1.0000001, 1e-9, n, cyc);
  k<2><<<1, 32>>>(out, 1.0000001, 0.9999999, n, cyc);
  k<3><<<1, 32>>>(out, 1.0000001, 0.9999999, n, cyc);
  k<4><<<1, 32>>>(out, 3.0, 0.9999999, n, cyc);
  k<5><<<1, 32>>>(out, 1.0000001, 0.5, n, cyc);
  cudaDeviceSynchronize();
  const char* names[] = {"DFMA", "DADD", "DMUL", "DDIV", "LDS+F2I chase", "DSETP+SEL+DADD"};
  for (int i = 0; i < 6; ++i) printf("%-16s %.1f cycles per op\n", names[i], (double)cyc[i] / (n * 16.0));
  return 0;
}

// hd.h — host/device portability shim for the *logic* headers (flsa_core.cuh, irls_core.cuh,
// segment_core.cuh).  The product library compiles them with nvcc for sm_100a only; tests/emu
// compiles the very same inline functions with g++ to check kernel logic on the CPU container
// (there is no GPU here).  The emulation is test infrastructure: nothing in the shipped
// libmethylasso_b200.so can run without a CUDA device.
#pragma once
#include <cstdint>
#include <cmath>
#include <cstring>

#ifdef __CUDACC__
#define ML_HD __host__ __device__ __forceinline__
#define ML_D __device__ __forceinline__
#else
#define ML_HD inline
#define ML_D inline
#endif

namespace ml {

// bit view of a double (knot hashing, bitwise state comparison)
ML_HD uint64_t f64_bits(double v) {
#ifdef __CUDA_ARCH__
  return (uint64_t)__double_as_longlong(v);
#else
  uint64_t u;
  memcpy(&u, &v, 8);
  return u;
#endif
}

ML_HD bool is_finite(double v) {
  // exponent not all ones
  return ((f64_bits(v) >> 52) & 0x7ffull) != 0x7ffull;
}

// num / den.  A zero numerator (an unmethylated position: Nmeth == 0) sends the device's IEEE division down its slow
// path (a ~150-instruction subroutine; one such lane is enough for the whole warp to take it).  The quotient of a
// zero by a non-zero, non-NaN divisor is a zero whose sign is the product of the signs: same bits, no call.
ML_HD double div_nz(double num, double den) {
#ifdef __CUDA_ARCH__
  if (num == 0.0 && den != 0.0 && den == den)
    return __longlong_as_double((__double_as_longlong(num) ^ __double_as_longlong(den)) & (long long)0x8000000000000000ull);
#endif
  return num / den;
}

}  // namespace ml

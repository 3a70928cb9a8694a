// common.h — host-side plumbing shared by the engine's translation units: CUDA error checks,
// stream-ordered device buffers, the engine object.  No compute lives here.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace ml {

struct MlError : public std::runtime_error {
  int code;
  MlError(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

#define ML_CUDA(expr)                                                                          \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess)                                                                     \
      throw ::ml::MlError(100, std::string("CUDA error: ") + cudaGetErrorString(_e) + " at " +  \
                                   __FILE__ + ":" + std::to_string(__LINE__) + " (" #expr ")"); \
  } while (0)

#define ML_CHECK_LAUNCH() ML_CUDA(cudaGetLastError())

struct Engine {
  int device = 0;
  int sm_count = 148;
  cudaStream_t stream = nullptr;
  // tunables (env overridable, see engine.cu)
  int flsa_chunk = 0;       // 0 = choose from total work
  int flsa_warm = 0;        // 0 = choose from lambda
  int flsa_sequential = 0;  // 1 = one thread per series start-to-end (bit-faithful verification mode)
  // launch accounting for bench.py's gpu_launches
  long long launches = 0;
};

// Stream-ordered allocation: cudaMallocAsync from the device's default pool (the engine raises
// the pool's release threshold so steady-state calls never reach the driver allocator).
template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  cudaStream_t s = nullptr;
  DevBuf() = default;
  DevBuf(cudaStream_t st, size_t count) { alloc(st, count); }
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  DevBuf(DevBuf&& o) noexcept : p(o.p), n(o.n), s(o.s) { o.p = nullptr; o.n = 0; }
  DevBuf& operator=(DevBuf&& o) noexcept {
    if (this != &o) { release(); p = o.p; n = o.n; s = o.s; o.p = nullptr; o.n = 0; }
    return *this;
  }
  ~DevBuf() { release(); }
  void alloc(cudaStream_t st, size_t count) {
    release();
    s = st;
    n = count;
    if (count) ML_CUDA(cudaMallocAsync((void**)&p, count * sizeof(T), st));
  }
  void release() {
    if (p) cudaFreeAsync(p, s);
    p = nullptr;
    n = 0;
  }
  void zero() { if (n) ML_CUDA(cudaMemsetAsync(p, 0, n * sizeof(T), s)); }
  void upload(const T* h, size_t count) {
    if (count) ML_CUDA(cudaMemcpyAsync(p, h, count * sizeof(T), cudaMemcpyHostToDevice, s));
  }
  void upload(const std::vector<T>& h) { upload(h.data(), h.size()); }
  void download(T* h, size_t count) const {
    if (count) ML_CUDA(cudaMemcpyAsync(h, p, count * sizeof(T), cudaMemcpyDeviceToHost, s));
  }
};

inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

}  // namespace ml

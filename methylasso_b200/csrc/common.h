// common.h — host-side plumbing shared by the engine's translation units: CUDA error checks,
// stream-ordered device buffers, the engine object.  No compute lives here.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <map>
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

// Optional per-kernel timing with CUDA events on the launching stream (bench.py's roofline numbers).
struct Profiler {
  struct Rec { const char* name; cudaEvent_t a, b; };
  bool on = false;
  std::vector<Rec> pending;
  std::vector<cudaEvent_t> pool;
  std::map<std::string, std::pair<double, long long>> acc;  // name -> (total ms, launches)
  std::map<std::string, double> work;   // name -> algorithmic bytes moved, for kernels whose work is data-dependent
  // optional timeline: start / end of every launch relative to `origin` (an event of the process, so that the engines
  // of a pool share one clock)
  struct Span { const char* name; float t0, t1; };
  cudaEvent_t origin = nullptr;
  std::vector<Span> spans;
  void add_work(const char* name, double bytes) { if (on) work[name] += bytes; }
  cudaEvent_t get() {
    if (!pool.empty()) { cudaEvent_t e = pool.back(); pool.pop_back(); return e; }
    cudaEvent_t e;
    ML_CUDA(cudaEventCreate(&e));
    return e;
  }
  void begin(const char* name, cudaStream_t st) {
    if (!on) return;
    Rec r{name, get(), get()};
    ML_CUDA(cudaEventRecord(r.a, st));
    pending.push_back(r);
  }
  void end(cudaStream_t st) {
    if (!on) return;
    ML_CUDA(cudaEventRecord(pending.back().b, st));
  }
  void collect() {
    for (Rec& r : pending) {
      ML_CUDA(cudaEventSynchronize(r.b));
      float ms = 0;
      ML_CUDA(cudaEventElapsedTime(&ms, r.a, r.b));
      auto& a = acc[r.name];
      a.first += ms;
      a.second += 1;
      if (origin) {
        float t0 = 0;
        ML_CUDA(cudaEventElapsedTime(&t0, origin, r.a));
        spans.push_back(Span{r.name, t0, t0 + ms});
      }
      pool.push_back(r.a);
      pool.push_back(r.b);
    }
    pending.clear();
  }
  void reset() { collect(); acc.clear(); work.clear(); spans.clear(); }
  ~Profiler() { for (cudaEvent_t e : pool) cudaEventDestroy(e); }
};

// Small device->host readbacks (scan totals, open-chunk counters, per-job scalars) do not go through
// the copy engine: a tiny kernel writes them into host-mapped pinned memory ("mailbox") and the host
// picks them up after the stream synchronisation.  A 4-byte cudaMemcpy would otherwise queue behind
// any large result download in flight on the same copy engine and stall the compute stream for the
// whole transfer (measured: segmentation 10 ms -> 28 ms while 1.7 GB of results were downloading).
//
// The same holds in the other direction: the descriptor tables a call uploads (tiles, job structs,
// control words; a few KB to a few MB) would queue behind another engine's 300 MB input upload on the
// host-to-device copy engine.  They are memcpy'd into the mapped "inbox" instead and moved to device
// memory by a small kernel on the compute stream.  Both areas are recycled at every stream_sync.
struct Mailbox {
  char* host = nullptr;   // cudaHostAlloc(mapped)
  char* dev = nullptr;    // device alias of the same memory
  size_t cap = 0, used = 0;
  struct Pending { void* dst; size_t off, bytes; };
  std::vector<Pending> pending;
  char* in_host = nullptr;   // inbox (host -> device staging), cudaHostAlloc(mapped | write-combined)
  char* in_dev = nullptr;
  size_t in_cap = 0, in_used = 0;
};
inline Mailbox*& tls_mailbox() {
  static thread_local Mailbox* m = nullptr;
  return m;
}
// defined in capi.cu
bool mailbox_fetch(cudaStream_t st, void* host_dst, const void* dev_src, size_t bytes);
bool inbox_push(cudaStream_t st, void* dev_dst, const void* host_src, size_t bytes);
void stream_sync(cudaStream_t st);   // cudaStreamSynchronize + delivery of pending mailbox readbacks

struct Engine {
  int device = 0;
  int sm_count = 148;
  cudaStream_t stream = nullptr;
  bool own_stream = true;
  cudaMemPool_t pool = nullptr;
  cudaStream_t copy_stream = nullptr;   // D2H of results can run here while the main stream computes
  Mailbox mailbox;
  Profiler prof;
  // tunables (env overridable, see engine.cu)
  int flsa_chunk = 0;       // 0 = choose from total work
  int flsa_warm = 0;        // 0 = choose from lambda
  int flsa_sequential = 0;  // 1 = one thread per series start-to-end (bit-faithful verification mode)
  // chunk / warm-up lengths by series length (read from the environment once, when the engine is created)
  int flsa_small_n = 262144, flsa_chunk_big = 448, flsa_warm_big = 512, flsa_chunk_small = 64, flsa_warm_small = 768;
  // route of a series through the forward pass: 0 = by the series itself (short series and series with empty bins take
  // the scan route, flsa_core.cuh; long dense ones the speculative route), 1 = speculative route only, 2 = scan route only
  int flsa_route = 0;
  int flsa_scan_lambda = 26;    // route 0: series with a penalty of at least this take the scan route too.  The memory of a
                                // series grows with lambda2 / weight: on WGBS coverage the speculative route verifies 97 %
                                // of its chunks at the reference's default 25, walks chains of a hundred chunks at 41
                                // (configs[4] sweep: 37 ms in the walkers) and no warm-up covers the PMD step's 1000
  int flsa_scan_chunk_small = 64, flsa_scan_chunk_big = 448;
  int flsa_fine = 1;            // 1: chains of the speculative route by composition (fine chains, flsa_core.cuh); 0: walkers only
  int flsa_chain_limit = 6;     // route 0: a walker that has redone this many chunks and still faces an unverified one hands
                                // its series over to the scan route (0: never)
  int host_threads = 4;     // host threads of batch_upload's passes over the input columns (run scan, packing of the counts)
  char* stage_host = nullptr;   // page-locked staging area of batch_stage (grown on demand, freed with the engine)
  size_t stage_cap = 0;
  int want_mu = 1;          // 0: a one-step fit does not store the per-row mu (nothing in the reference's R layer reads it)
  // counters of the last FLSA solve (tests): chunks verified, redone by walkers, longest chain, chains, series handed over
  // to the scan route, chunks of the scan route's tables, series finished by the start-to-end kernel
  int flsa_last[7] = {0, 0, 0, 0, 0, 0, 0};
  // launch accounting for bench.py's gpu_launches
  long long launches = 0;
};

// Memory pool of the engine driving the current host thread (set by every C-ABI entry).  Each
// engine owns a pool: with one shared pool, memory freed on one engine's stream and reused on
// another's makes the allocator insert cross-stream dependencies, which serialises engines that are
// meant to overlap (EnginePool in api.py).
inline cudaMemPool_t& tls_pool() {
  static thread_local cudaMemPool_t p = nullptr;
  return p;
}

// Stream-ordered allocation from the engine's own pool (release threshold raised, so steady-state
// calls never reach the driver allocator).
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
    if (count) {
      if (tls_pool()) ML_CUDA(cudaMallocFromPoolAsync((void**)&p, count * sizeof(T), tls_pool(), st));
      else ML_CUDA(cudaMallocAsync((void**)&p, count * sizeof(T), st));
    }
  }
  void release() {
    if (p) cudaFreeAsync(p, s);
    p = nullptr;
    n = 0;
  }
  void zero() { if (n) ML_CUDA(cudaMemsetAsync(p, 0, n * sizeof(T), s)); }
  void upload(const T* h, size_t count) {
    if (!count) return;
    if (count * sizeof(T) <= (size_t)(8 << 20) && inbox_push(s, p, h, count * sizeof(T))) return;
    ML_CUDA(cudaMemcpyAsync(p, h, count * sizeof(T), cudaMemcpyHostToDevice, s));
  }
  void upload(const std::vector<T>& h) { upload(h.data(), h.size()); }
  void download(T* h, size_t count) const {
    if (!count) return;
    if (count * sizeof(T) <= 65536 && mailbox_fetch(s, h, p, count * sizeof(T))) return;
    ML_CUDA(cudaMemcpyAsync(h, p, count * sizeof(T), cudaMemcpyDeviceToHost, s));
  }
};

inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// algorithmic bytes of the launch just made (compulsory reads + writes from the launch's own sizes): bench.py divides
// them by the kernel's measured time; recorded only while the profiler is on
#define ML_WORK(ENG, NAME, BYTES) (ENG).prof.add_work(NAME, (double)(BYTES))

// every kernel launch of the engine goes through this: error check, launch count, optional timing
#define ML_LAUNCH(ENG, NAME, ...)            \
  do {                                       \
    (ENG).prof.begin(NAME, (ENG).stream);    \
    __VA_ARGS__;                             \
    ML_CHECK_LAUNCH();                       \
    (ENG).prof.end((ENG).stream);            \
    (ENG).launches++;                        \
  } while (0)

}  // namespace ml

// capi.cu — the C ABI of include/methylasso_b200.h: argument checking, error translation and the
// host<->device copies at the boundary.  All compute is in flsa.cu / detect.cu / segment.cu.
#include <cstdlib>
#include <algorithm>
#include <cstring>
#include <iterator>
#include <memory>
#include <mutex>
#include <string>

#include "../../include/methylasso_b200.h"
#include "debug_api.h"
#include "detect.cuh"
#include "segment.cuh"
#include "stats.cuh"
#include "codec.cuh"

using namespace ml;

struct ml_engine { Engine e; cudaStream_t own = nullptr; };
struct ml_batch { std::unique_ptr<Batch> b; };
struct ml_result { std::unique_ptr<Result> r; };
struct ml_text { TextTable t; int device = 0; };

static thread_local std::string g_err;

namespace ml {

__global__ void k_mailbox_copy(char* __restrict__ dst, const char* __restrict__ src, size_t bytes) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < bytes; i += (size_t)gridDim.x * blockDim.x)
    dst[i] = src[i];
}

bool mailbox_fetch(cudaStream_t st, void* host_dst, const void* dev_src, size_t bytes) {
  Mailbox* m = tls_mailbox();
  if (!m || !m->host) return false;
  const size_t off = (m->used + 15) & ~(size_t)15;
  if (off + bytes > m->cap) return false;
  k_mailbox_copy<<<(unsigned)std::min<size_t>(32, (bytes + 255) / 256), 256, 0, st>>>(m->dev + off, (const char*)dev_src, bytes);
  ML_CHECK_LAUNCH();
  m->pending.push_back(Mailbox::Pending{host_dst, off, bytes});
  m->used = off + bytes;
  return true;
}

// staged copy: 16-byte words when both sides allow it, bytes otherwise
__global__ void k_stage_copy(char* __restrict__ dst, const char* __restrict__ src, size_t bytes) {
  const size_t tid = blockIdx.x * (size_t)blockDim.x + threadIdx.x, nth = (size_t)gridDim.x * blockDim.x;
  size_t done = 0;
  if ((((uintptr_t)dst | (uintptr_t)src) & 15) == 0) {
    const size_t nw = bytes / 16;
    for (size_t i = tid; i < nw; i += nth) reinterpret_cast<uint4*>(dst)[i] = reinterpret_cast<const uint4*>(src)[i];
    done = nw * 16;
  }
  for (size_t i = done + tid; i < bytes; i += nth) dst[i] = src[i];
}

bool inbox_push(cudaStream_t st, void* dev_dst, const void* host_src, size_t bytes) {
  Mailbox* m = tls_mailbox();
  if (!m || !m->in_host) return false;
  const size_t off = (m->in_used + 255) & ~(size_t)255;
  if (off + bytes > m->in_cap) return false;
  memcpy(m->in_host + off, host_src, bytes);
  const unsigned blocks = (unsigned)std::min<size_t>(296, (bytes + 4095) / 4096);
  k_stage_copy<<<std::max(1u, blocks), 256, 0, st>>>((char*)dev_dst, m->in_dev + off, bytes);
  ML_CHECK_LAUNCH();
  m->in_used = off + bytes;
  return true;
}

void stream_sync(cudaStream_t st) {
  ML_CUDA(cudaStreamSynchronize(st));
  Mailbox* m = tls_mailbox();
  if (m) m->in_used = 0;
  if (m && !m->pending.empty()) {
    for (const Mailbox::Pending& p : m->pending) memcpy(p.dst, m->host + p.off, p.bytes);
    m->pending.clear();
    m->used = 0;
  }
}

}  // namespace ml

static int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}

// An exception may unwind frames that small readbacks were registered into (mailbox_fetch keeps the destination
// address until the next stream_sync): drop those records, or a later synchronisation would write through them.
static void mailbox_abort() {
  if (Mailbox* m = tls_mailbox()) { m->pending.clear(); m->used = 0; m->in_used = 0; }
}
#define ML_TRY try {
#define ML_CATCH                                                        \
  }                                                                     \
  catch (const MlError& e) { mailbox_abort(); return fail(e.code, e.what()); }           \
  catch (const std::bad_alloc&) { mailbox_abort(); return fail(ML_ERR_CUDA, "host allocation failed"); } \
  catch (const std::exception& e) { mailbox_abort(); return fail(ML_ERR_BAD_ARGUMENT, e.what()); }

static_assert(sizeof(ml_params) == sizeof(Params), "ml_params / Params layout mismatch");

extern "C" {

int ml_abi_version(void) { return ML_ABI_VERSION; }
const char* ml_last_error(void) { return g_err.c_str(); }

const char* ml_strerror(int code) {
  switch (code) {
    case ML_OK: return "ok";
    case ML_ERR_TOO_FEW_ROWS: return "Need at least 2 data points!";
    case ML_ERR_SIZE_MISMATCH: return "Input vectors must be of same size!";
    case ML_ERR_FIRST_DATASET: return "First dataset ID must be 1";
    case ML_ERR_FIRST_CONDITION: return "First condition ID must be 1";
    case ML_ERR_FIRST_REPLICATE: return "First replicate ID must be 1";
    case ML_ERR_NOT_SORTED: return "data is not sorted by dataset and position";
    case ML_ERR_DATASET_GAP: return "dataset IDs must start at 1 and increase without gaps";
    case ML_ERR_REPLICATE_GAP: return "replicate ID must increase without gaps within the same condition";
    case ML_ERR_CONDITION_GAP: return "condition ID must start at 1 and increase without gaps";
    case ML_ERR_REPLICATE_START: return "replicate ID must start at 1 within the same condition";
    case ML_ERR_NEGATIVE_COUNT: return "Read counts should be >=0 !";
    case ML_ERR_NMETH_GT_COV: return "Number of methylated reads cannot be larger than number of total reads!";
    case ML_ERR_LAMBDA2: return "lambda2 should be >0";
    case ML_ERR_TWO_CONDITIONS: return "Needs at least two conditions!";
    case ML_ERR_BAD_REF: return "Invalid reference dataset!";
    case ML_ERR_NAN: return "NaN found, aborting...";
    case ML_ERR_TOO_FEW_PARAMS: return "fewer than 2 parameters (reference behaviour undefined)";
    case ML_ERR_BAD_ARGUMENT: return "bad argument";
    case ML_ERR_CUDA: return "CUDA failure";
    case ML_ERR_NO_DEVICE: return "no CUDA device";
    default: return "unknown error";
  }
}

int ml_engine_create(int device, ml_engine** out) {
  if (!out) return fail(ML_ERR_BAD_ARGUMENT, "out is NULL");
  *out = nullptr;
  int count = 0;
  cudaError_t ce = cudaGetDeviceCount(&count);
  if (ce != cudaSuccess || count == 0)
    return fail(ML_ERR_NO_DEVICE, std::string("no CUDA device available (") +
                                      (ce != cudaSuccess ? cudaGetErrorString(ce) : "count = 0") +
                                      "); this engine has no CPU path");
  ML_TRY
  auto* h = new ml_engine();
  std::unique_ptr<ml_engine> guard(h);
  if (device < 0) ML_CUDA(cudaGetDevice(&device));
  if (device >= count) throw MlError(ML_ERR_BAD_ARGUMENT, "device ordinal out of range");
  ML_CUDA(cudaSetDevice(device));
  h->e.device = device;
  cudaDeviceProp prop;
  ML_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10)
    throw MlError(ML_ERR_NO_DEVICE, "this library is built for sm_100a only; found sm_" +
                                        std::to_string(prop.major) + std::to_string(prop.minor));
  h->e.sm_count = prop.multiProcessorCount;
  ML_CUDA(cudaStreamCreateWithFlags(&h->e.stream, cudaStreamNonBlocking));
  ML_CUDA(cudaStreamCreateWithFlags(&h->e.copy_stream, cudaStreamNonBlocking));
  detect_init_device_tables(h->e.stream);
  segment_init_device_tables(h->e.stream);
  // the engine's own pool; freed blocks stay in it, so steady-state calls never reach the driver
  cudaMemPoolProps props = {};
  props.allocType = cudaMemAllocationTypePinned;
  props.handleTypes = cudaMemHandleTypeNone;
  props.location.type = cudaMemLocationTypeDevice;
  props.location.id = device;
  ML_CUDA(cudaMemPoolCreate(&h->e.pool, &props));
  uint64_t thr = ~0ull;
  ML_CUDA(cudaMemPoolSetAttribute(h->e.pool, cudaMemPoolAttrReleaseThreshold, &thr));
  h->e.mailbox.cap = 1 << 20;
  ML_CUDA(cudaHostAlloc((void**)&h->e.mailbox.host, h->e.mailbox.cap, cudaHostAllocMapped));
  ML_CUDA(cudaHostGetDevicePointer((void**)&h->e.mailbox.dev, h->e.mailbox.host, 0));
  h->e.mailbox.in_cap = 32 << 20;
  ML_CUDA(cudaHostAlloc((void**)&h->e.mailbox.in_host, h->e.mailbox.in_cap, cudaHostAllocMapped | cudaHostAllocWriteCombined));
  ML_CUDA(cudaHostGetDevicePointer((void**)&h->e.mailbox.in_dev, h->e.mailbox.in_host, 0));
  // tunables are read from the environment HERE, once per engine: nothing on a launch path calls getenv
  auto env_int = [](const char* name, int& v) { if (const char* s = getenv(name)) v = atoi(s); };
  env_int("ML_FLSA_CHUNK", h->e.flsa_chunk);
  env_int("ML_FLSA_WARM", h->e.flsa_warm);
  env_int("ML_FLSA_SEQUENTIAL", h->e.flsa_sequential);
  env_int("ML_FLSA_SMALL_N", h->e.flsa_small_n);
  env_int("ML_FLSA_CHUNK_BIG", h->e.flsa_chunk_big);
  env_int("ML_FLSA_WARM_BIG", h->e.flsa_warm_big);
  env_int("ML_FLSA_CHUNK_SMALL", h->e.flsa_chunk_small);
  env_int("ML_FLSA_WARM_SMALL", h->e.flsa_warm_small);
  env_int("ML_FLSA_ROUTE", h->e.flsa_route);
  env_int("ML_FLSA_SCAN_LAMBDA", h->e.flsa_scan_lambda);
  env_int("ML_FLSA_CHAIN_LIMIT", h->e.flsa_chain_limit);
  env_int("ML_FLSA_FINE", h->e.flsa_fine);
  env_int("ML_FLSA_SCAN_CHUNK_SMALL", h->e.flsa_scan_chunk_small);
  env_int("ML_FLSA_SCAN_CHUNK_BIG", h->e.flsa_scan_chunk_big);
  env_int("ML_HOST_THREADS", h->e.host_threads);
  *out = guard.release();
  return ML_OK;
  ML_CATCH
}

void ml_engine_destroy(ml_engine* eng) {
  if (!eng) return;
  cudaSetDevice(eng->e.device);
  if (eng->e.stream) cudaStreamSynchronize(eng->e.stream);
  cudaStream_t mine = eng->e.own_stream ? eng->e.stream : eng->own;
  if (mine) cudaStreamDestroy(mine);
  if (eng->e.copy_stream) { cudaStreamSynchronize(eng->e.copy_stream); cudaStreamDestroy(eng->e.copy_stream); }
  if (eng->e.pool) cudaMemPoolDestroy(eng->e.pool);
  if (eng->e.mailbox.host) cudaFreeHost(eng->e.mailbox.host);
  if (eng->e.mailbox.in_host) cudaFreeHost(eng->e.mailbox.in_host);
  if (eng->e.stage_host) cudaFreeHost(eng->e.stage_host);
  delete eng;
}

int ml_engine_set(ml_engine* eng, const char* key, long long value) {
  if (!eng || !key) return fail(ML_ERR_BAD_ARGUMENT, "NULL argument");
  const std::string k(key);
  if (k == "flsa_chunk") eng->e.flsa_chunk = (int)value;
  else if (k == "flsa_warm") eng->e.flsa_warm = (int)value;
  else if (k == "flsa_sequential") eng->e.flsa_sequential = (int)value;
  else if (k == "flsa_route") eng->e.flsa_route = (int)value;
  else if (k == "flsa_scan_lambda") eng->e.flsa_scan_lambda = (int)value;
  else if (k == "flsa_chain_limit") eng->e.flsa_chain_limit = (int)value;
  else if (k == "flsa_fine") eng->e.flsa_fine = value != 0;
  else if (k == "flsa_scan_chunk_small") eng->e.flsa_scan_chunk_small = (int)std::max<long long>(1, value);
  else if (k == "flsa_scan_chunk_big") eng->e.flsa_scan_chunk_big = (int)std::max<long long>(1, value);
  else if (k == "host_threads") eng->e.host_threads = (int)std::max<long long>(1, value);
  else if (k == "want_mu") eng->e.want_mu = value != 0;
  else if (k == "profile") { if (!value) eng->e.prof.collect(); eng->e.prof.on = value != 0; }
  else if (k == "stream_priority") {
    // Re-create the engine's own stream at priority level `value` (0 = the lowest, as created; larger = served first
    // when kernels of several streams wait for SMs; clamped to what the device offers).  For the engines of a pool whose
    // groups enter one after the other: the group that enters last is the one everybody waits for.  Only right after
    // the engine was created (nothing allocated or in flight on the stream yet).
    ML_TRY
    if (!eng->e.own_stream) throw MlError(ML_ERR_BAD_ARGUMENT, "stream_priority: the engine runs on a caller's stream");
    ML_CUDA(cudaSetDevice(eng->e.device));
    int least = 0, greatest = 0;
    ML_CUDA(cudaDeviceGetStreamPriorityRange(&least, &greatest));    // numerically lower = higher priority
    const int pr = (int)std::max<long long>(greatest, least - std::max<long long>(0, value));
    ML_CUDA(cudaStreamSynchronize(eng->e.stream));
    cudaStream_t fresh = nullptr;
    ML_CUDA(cudaStreamCreateWithPriority(&fresh, cudaStreamNonBlocking, pr));
    ML_CUDA(cudaStreamDestroy(eng->e.stream));
    eng->e.stream = fresh;
    ML_CATCH
  }
  else if (k == "profile_timeline") {
    // keep the start / end of every profiled launch on the clock of one event of the process (all engines share it)
    ML_TRY
    static cudaEvent_t origin = nullptr;
    static std::mutex mu;
    std::lock_guard<std::mutex> lock(mu);
    if (value && !origin) {
      ML_CUDA(cudaEventCreate(&origin));
      ML_CUDA(cudaEventRecord(origin, eng->e.stream));
      ML_CUDA(cudaEventSynchronize(origin));
    }
    eng->e.prof.collect();
    eng->e.prof.origin = value ? origin : nullptr;
    ML_CATCH
  }
  else if (k == "profile_reserve_events") {
    // pre-create timing events so that a profiled region does not call cudaEventCreate
    ML_TRY
    ML_CUDA(cudaSetDevice(eng->e.device));
    for (long long i = (long long)eng->e.prof.pool.size(); i < value; ++i) {
      cudaEvent_t ev;
      ML_CUDA(cudaEventCreate(&ev));
      eng->e.prof.pool.push_back(ev);
    }
    ML_CATCH
  }
  else if (k == "pool_headroom_percent") {
    // Grow the engine's memory pool NOW by value % of what it has reserved so far (one block, allocated and
    // freed): a later call whose allocation order fragments the pool slightly differently then finds the
    // memory in the pool instead of paying a cuMemMap in the middle of a step (measured: +0.2 GB, +20..100 ms).
    ML_TRY
    ML_CUDA(cudaSetDevice(eng->e.device));
    uint64_t reserved = 0;
    ML_CUDA(cudaMemPoolGetAttribute(eng->e.pool, cudaMemPoolAttrReservedMemCurrent, &reserved));
    const size_t extra = (size_t)((double)reserved * (double)value / 100.0);
    if (extra) {
      void* ptr = nullptr;
      ML_CUDA(cudaMallocFromPoolAsync(&ptr, extra, eng->e.pool, eng->e.stream));
      ML_CUDA(cudaFreeAsync(ptr, eng->e.stream));
      ML_CUDA(cudaStreamSynchronize(eng->e.stream));
    }
    ML_CATCH
  }
  else return fail(ML_ERR_BAD_ARGUMENT, "unknown engine key: " + k);
  return ML_OK;
}

long long ml_engine_launch_count(const ml_engine* eng) { return eng ? eng->e.launches : 0; }
void* ml_engine_stream(ml_engine* eng) { return eng ? (void*)eng->e.stream : nullptr; }
int ml_engine_synchronize(ml_engine* eng) {
  if (!eng) return fail(ML_ERR_BAD_ARGUMENT, "NULL engine");
  ML_TRY
  ML_CUDA(cudaSetDevice(eng->e.device));
  tls_pool() = eng->e.pool;
  tls_mailbox() = &eng->e.mailbox;
  stream_sync(eng->e.stream);
  ML_CUDA(cudaStreamSynchronize(eng->e.copy_stream));
  return ML_OK;
  ML_CATCH
}

int ml_engine_set_stream(ml_engine* eng, void* cuda_stream) {
  if (!eng) return fail(ML_ERR_BAD_ARGUMENT, "NULL engine");
  ML_TRY
  ML_CUDA(cudaSetDevice(eng->e.device));
  tls_pool() = eng->e.pool;
  tls_mailbox() = &eng->e.mailbox;
  stream_sync(eng->e.stream);
  if (cuda_stream) {
    if (eng->e.own_stream) { eng->e.own_stream = false; eng->own = eng->e.stream; }
    eng->e.stream = (cudaStream_t)cuda_stream;
  } else if (!eng->e.own_stream) {
    eng->e.stream = eng->own;
    eng->e.own_stream = true;
  }
  return ML_OK;
  ML_CATCH
}

int ml_engine_profile_count(ml_engine* eng) {
  if (!eng) return 0;
  try { eng->e.prof.collect(); } catch (...) { return 0; }
  return (int)eng->e.prof.acc.size();
}
int ml_engine_profile_get(ml_engine* eng, int i, char* name, int name_cap, double* total_ms,
                          long long* launches) {
  if (!eng || i < 0 || i >= (int)eng->e.prof.acc.size()) return fail(ML_ERR_BAD_ARGUMENT, "bad profile index");
  auto it = eng->e.prof.acc.begin();
  std::advance(it, i);
  if (name && name_cap > 0) { strncpy(name, it->first.c_str(), name_cap - 1); name[name_cap - 1] = 0; }
  if (total_ms) *total_ms = it->second.first;
  if (launches) *launches = it->second.second;
  return ML_OK;
}
int ml_engine_profile_work(ml_engine* eng, int i, double* bytes) {
  if (!eng || !bytes || i < 0 || i >= (int)eng->e.prof.acc.size()) return fail(ML_ERR_BAD_ARGUMENT, "bad profile index");
  auto it = eng->e.prof.acc.begin();
  std::advance(it, i);
  auto w = eng->e.prof.work.find(it->first);
  *bytes = w == eng->e.prof.work.end() ? -1.0 : w->second;
  return ML_OK;
}
int ml_engine_profile_spans(ml_engine* eng, int first, int count, char* names, int name_cap, float* t0_ms, float* t1_ms) {
  if (!eng) return 0;
  try { eng->e.prof.collect(); } catch (...) { return 0; }
  const int n = (int)eng->e.prof.spans.size();
  if (!names || !t0_ms || !t1_ms) return n;
  int k = 0;
  for (int i = first; i < n && k < count; ++i, ++k) {
    const Profiler::Span& sp = eng->e.prof.spans[i];
    strncpy(names + (size_t)k * name_cap, sp.name, name_cap - 1);
    names[(size_t)k * name_cap + name_cap - 1] = 0;
    t0_ms[k] = sp.t0;
    t1_ms[k] = sp.t1;
  }
  return k;
}
int ml_engine_profile_reset(ml_engine* eng) {
  if (!eng) return fail(ML_ERR_BAD_ARGUMENT, "NULL engine");
  ML_TRY
  eng->e.prof.reset();
  return ML_OK;
  ML_CATCH
}

// ---- weighted_1d_flsa -----------------------------------------------------------------------------
static void flsa_batch_device(Engine& e, int nseries, const int64_t* ptr, const double* d_y,
                              const double* d_w, const double* lambda2, double lambda1, double* d_beta) {
  std::vector<FlsaSeries> series(nseries);
  for (int s = 0; s < nseries; ++s) {
    const int64_t n = ptr[s + 1] - ptr[s];
    if (n < 0 || n > 0x7fffffff) throw MlError(ML_ERR_BAD_ARGUMENT, "series length out of range");
    series[s] = FlsaSeries{ptr[s], (int32_t)n, lambda2[s]};
  }
  FlsaPlan plan(e, series, ptr[nseries]);
  plan.solve(d_y, d_w, d_beta, 1, lambda1, nullptr);
  stream_sync(e.stream);
}

// debugging aid (not part of the drop-in surface): one series, also returns the DP back-pointers
int ml_debug_flsa_backpointers(ml_engine* eng, int64_t n, const double* y, const double* wt, double lambda2,
                               double* beta, double* tm, double* tp) {
  if (!eng || n <= 0) return fail(ML_ERR_BAD_ARGUMENT, "bad argument");
  ML_TRY
  ML_CUDA(cudaSetDevice(eng->e.device));
  tls_pool() = eng->e.pool;
  tls_mailbox() = &eng->e.mailbox;
  cudaStream_t st = eng->e.stream;
  DevBuf<double> d_y(st, (size_t)n), d_w(st, (size_t)n), d_b(st, (size_t)n);
  d_y.upload(y, (size_t)n);
  d_w.upload(wt, (size_t)n);
  std::vector<FlsaSeries> series{FlsaSeries{0, (int32_t)n, lambda2}};
  FlsaPlan plan(eng->e, series, n);
  plan.solve(d_y.p, d_w.p, d_b.p, 0, 0.0, nullptr);
  d_b.download(beta, (size_t)n);
  plan.debug_backpointers(tm, tp);
  return ML_OK;
  ML_CATCH
}

int ml_debug_flsa_last_stats(ml_engine* eng, int* out7) {
  if (!eng || !out7) return fail(ML_ERR_BAD_ARGUMENT, "NULL argument");
  for (int i = 0; i < 7; ++i) out7[i] = eng->e.flsa_last[i];
  return ML_OK;
}
int ml_debug_div_counts(ml_engine* eng, int max_num, int max_den, long long* pairs, long long* mismatches) {
  if (!eng || !pairs || !mismatches || max_num < 0 || max_den < 2) return fail(ML_ERR_BAD_ARGUMENT, "bad argument");
  ML_TRY
  ML_CUDA(cudaSetDevice(eng->e.device));
  tls_pool() = eng->e.pool;
  tls_mailbox() = &eng->e.mailbox;
  debug_div_counts(eng->e, max_num, max_den, pairs, mismatches);
  return ML_OK;
  ML_CATCH
}
int ml_debug_scan_runs(int64_t n, const int32_t* dataset, const int32_t* condition, const int32_t* replicate,
                       const int32_t* pos, int threads, int64_t cap, int64_t* n_runs, int64_t* first_row,
                       int32_t* run_condition, int32_t* run_replicate, int32_t* code, int64_t* code_row) {
  if (n < 0 || !n_runs || (n > 0 && (!dataset || !condition || !replicate || !pos))) return fail(ML_ERR_BAD_ARGUMENT, "bad argument");
  ML_TRY
  Batch b;
  debug_scan_runs(n, dataset, condition, replicate, pos, threads, b);
  const int64_t k = b.n_runs(0);
  *n_runs = k;
  for (int64_t t = 0; t < k && t < cap; ++t) {
    if (first_row) first_row[t] = b.dptr[t];
    if (run_condition) run_condition[t] = b.dcond[t];
    if (run_replicate) run_replicate[t] = b.drep[t];
  }
  const unsigned long long key = b.host_key[0];
  if (code) *code = key == ~0ull ? 0 : (int32_t)(key & 0xff);
  if (code_row) *code_row = key == ~0ull ? -1 : (int64_t)((key >> 8) & 0xfffffffffffffull);
  return ML_OK;
  ML_CATCH
}

int ml_weighted_1d_flsa_batch_device(ml_engine* eng, int nseries, const int64_t* ptr, const double* d_y,
                                     const double* d_wt, const double* lambda2, double lambda1,
                                     double* d_beta) {
  if (!eng || !ptr || !lambda2 || nseries < 0) return fail(ML_ERR_BAD_ARGUMENT, "bad argument");
  ML_TRY
  ML_CUDA(cudaSetDevice(eng->e.device));
  tls_pool() = eng->e.pool;
  tls_mailbox() = &eng->e.mailbox;
  if (nseries == 0 || ptr[nseries] == 0) return ML_OK;
  flsa_batch_device(eng->e, nseries, ptr, d_y, d_wt, lambda2, lambda1, d_beta);
  return ML_OK;
  ML_CATCH
}

int ml_weighted_1d_flsa_batch(ml_engine* eng, int nseries, const int64_t* ptr, const double* y,
                              const double* wt, const double* lambda2, double lambda1, double* beta) {
  if (!eng || !ptr || !lambda2 || nseries < 0) return fail(ML_ERR_BAD_ARGUMENT, "bad argument");
  ML_TRY
  ML_CUDA(cudaSetDevice(eng->e.device));
  tls_pool() = eng->e.pool;
  tls_mailbox() = &eng->e.mailbox;
  if (nseries == 0) return ML_OK;
  const int64_t total = ptr[nseries];
  if (total == 0) return ML_OK;
  if (!y || !wt || !beta) throw MlError(ML_ERR_BAD_ARGUMENT, "NULL data pointer");
  cudaStream_t st = eng->e.stream;
  DevBuf<double> d_y(st, (size_t)total), d_w(st, (size_t)total), d_b(st, (size_t)total);
  d_y.upload(y, (size_t)total);
  d_w.upload(wt, (size_t)total);
  flsa_batch_device(eng->e, nseries, ptr, d_y.p, d_w.p, lambda2, lambda1, d_b.p);
  d_b.download(beta, (size_t)total);
  stream_sync(st);
  return ML_OK;
  ML_CATCH
}

int ml_weighted_1d_flsa(ml_engine* eng, int64_t n, const double* y, const double* wt, double lambda2,
                        double lambda1, double* beta) {
  if (n < 0) return fail(ML_ERR_BAD_ARGUMENT, "n < 0");
  const int64_t ptr[2] = {0, n};
  return ml_weighted_1d_flsa_batch(eng, 1, ptr, y, wt, &lambda2, lambda1, beta);
}

// ---- detection ------------------------------------------------------------------------------------
int ml_batch_upload(ml_engine* eng, int njobs, const int64_t* job_ptr, const int32_t* dataset,
                    const int32_t* condition, const int32_t* replicate, const int32_t* pos,
                    const int32_t* nmeth, const int32_t* coverage, ml_batch** out, int* status) {
  if (!eng || !out || !job_ptr) return fail(ML_ERR_BAD_ARGUMENT, "NULL argument");
  *out = nullptr;
  ML_TRY
  ML_CUDA(cudaSetDevice(eng->e.device));
  tls_pool() = eng->e.pool;
  tls_mailbox() = &eng->e.mailbox;
  if (njobs > 0 && job_ptr[njobs] > 0 && (!dataset || !condition || !replicate || !pos || !nmeth || !coverage))
    throw MlError(ML_ERR_BAD_ARGUMENT, "NULL column (data should contain columns dataset, condition, "
                                       "replicate, pos, Nmeth and coverage!)");
  auto h = std::make_unique<ml_batch>();
  h->b = batch_upload(eng->e, njobs, job_ptr, dataset, condition, replicate, pos, nmeth, coverage);
  if (status)
    for (int j = 0; j < njobs; ++j) status[j] = 0;  // validation runs on the device inside detect
  *out = h.release();
  return ML_OK;
  ML_CATCH
}

int ml_batch_stage(ml_engine* eng, int njobs, const int64_t* job_ptr, const int32_t* dataset, const int32_t* condition,
                   const int32_t* replicate, const int32_t* pos, const int32_t* nmeth, const int32_t* coverage,
                   ml_batch** out) {
  if (!eng || !out || !job_ptr) return fail(ML_ERR_BAD_ARGUMENT, "NULL argument");
  *out = nullptr;
  ML_TRY
  ML_CUDA(cudaSetDevice(eng->e.device));
  if (njobs > 0 && job_ptr[njobs] > 0 && (!dataset || !condition || !replicate || !pos || !nmeth || !coverage))
    throw MlError(ML_ERR_BAD_ARGUMENT, "NULL column (data should contain columns dataset, condition, "
                                       "replicate, pos, Nmeth and coverage!)");
  auto h = std::make_unique<ml_batch>();
  h->b = batch_stage(eng->e, njobs, job_ptr, dataset, condition, replicate, pos, nmeth, coverage);
  *out = h.release();
  return ML_OK;
  ML_CATCH
}

int ml_batch_commit(ml_engine* eng, ml_batch* b) {
  if (!eng || !b || !b->b) return fail(ML_ERR_BAD_ARGUMENT, "NULL argument");
  ML_TRY
  ML_CUDA(cudaSetDevice(eng->e.device));
  tls_pool() = eng->e.pool;
  tls_mailbox() = &eng->e.mailbox;
  batch_commit(eng->e, *b->b);
  return ML_OK;
  ML_CATCH
}

int ml_batch_upload_runs(ml_engine* eng, int njobs, const int64_t* job_ptr, const int64_t* dset_ptr,
                         const int64_t* dset_row, const int32_t* dset_condition, const int32_t* dset_replicate,
                         const int32_t* pos, const int32_t* nmeth, const int32_t* coverage, ml_batch** out) {
  if (!eng || !out || !job_ptr || !dset_ptr) return fail(ML_ERR_BAD_ARGUMENT, "NULL argument");
  *out = nullptr;
  ML_TRY
  ML_CUDA(cudaSetDevice(eng->e.device));
  tls_pool() = eng->e.pool;
  tls_mailbox() = &eng->e.mailbox;
  if (njobs > 0 && job_ptr[njobs] > 0 && (!dset_row || !dset_condition || !dset_replicate || !pos || !nmeth || !coverage))
    throw MlError(ML_ERR_BAD_ARGUMENT, "NULL column");
  auto h = std::make_unique<ml_batch>();
  h->b = batch_upload_runs(eng->e, njobs, job_ptr, dset_ptr, dset_row, dset_condition, dset_replicate, pos, nmeth, coverage);
  *out = h.release();
  return ML_OK;
  ML_CATCH
}

void ml_batch_free(ml_batch* b) {
  if (!b) return;
  if (b->b && b->b->eng) cudaSetDevice(b->b->eng->device);
  delete b;
}

int ml_batch_detect(ml_engine* eng, ml_batch* b, const ml_params* params, ml_result** out, int* status) {
  if (!eng || !b || !params || !out) return fail(ML_ERR_BAD_ARGUMENT, "NULL argument");
  *out = nullptr;
  ML_TRY
  ML_CUDA(cudaSetDevice(eng->e.device));
  tls_pool() = eng->e.pool;
  tls_mailbox() = &eng->e.mailbox;
  Params p;
  memcpy(&p, params, sizeof(p));
  auto h = std::make_unique<ml_result>();
  h->r = batch_detect(eng->e, *b->b, p);
  if (status)
    for (size_t j = 0; j < h->r->jobs.size(); ++j) status[j] = h->r->jobs[j].status;
  *out = h.release();
  return ML_OK;
  ML_CATCH
}

int ml_detect_batch(ml_engine* eng, int njobs, const int64_t* job_ptr, const int32_t* dataset,
                    const int32_t* condition, const int32_t* replicate, const int32_t* pos,
                    const int32_t* nmeth, const int32_t* coverage, const ml_params* params,
                    ml_result** out, int* status) {
  ml_batch* b = nullptr;
  int rc = ml_batch_upload(eng, njobs, job_ptr, dataset, condition, replicate, pos, nmeth, coverage, &b, nullptr);
  if (rc) return rc;
  rc = ml_batch_detect(eng, b, params, out, status);
  ml_batch_free(b);
  return rc;
}

int ml_result_njobs(const ml_result* r) { return r ? (int)r->r->jobs.size() : 0; }

int ml_result_job_sizes(const ml_result* r, int job, int64_t* n_rows, int64_t* n_params, int32_t* n_est,
                        int32_t* n_dsets, int32_t* status) {
  if (!r || job < 0 || job >= (int)r->r->jobs.size()) return fail(ML_ERR_BAD_ARGUMENT, "bad job index");
  const JobHost& j = r->r->jobs[job];
  if (n_rows) *n_rows = j.nrows;
  if (n_params) *n_params = j.status ? 0 : j.P;
  if (n_est) *n_est = j.status ? 0 : j.E;
  if (n_dsets) *n_dsets = j.status ? 0 : j.D;
  if (status) *status = j.status;
  return ML_OK;
}

int ml_result_job_info(const ml_result* r, int job, int32_t* converged, int32_t* irls_steps,
                       int32_t* dampings, int32_t* outer_iters, double* hyper, double* mlogp) {
  if (!r || job < 0 || job >= (int)r->r->jobs.size()) return fail(ML_ERR_BAD_ARGUMENT, "bad job index");
  const JobHost& j = r->r->jobs[job];
  if (converged) *converged = j.converged;
  if (irls_steps) *irls_steps = j.irls_steps;
  if (dampings) *dampings = j.dampings;
  if (outer_iters) *outer_iters = j.outer_iters;
  if (hyper) *hyper = j.hyper;
  if (mlogp) *mlogp = j.mlogp;
  return ML_OK;
}

int ml_result_copy_job(ml_engine* eng, const ml_result* r, int job, double* mu, double* lambda2,
                       double* offset, int32_t* estimate_id, double* start, double* beta,
                       double* betahat, double* pseudoweight) {
  if (!eng || !r || job < 0 || job >= (int)r->r->jobs.size()) return fail(ML_ERR_BAD_ARGUMENT, "bad argument");
  ML_TRY
  ML_CUDA(cudaSetDevice(eng->e.device));
  tls_pool() = eng->e.pool;
  tls_mailbox() = &eng->e.mailbox;
  const Result& R = *r->r;
  const JobHost& j = R.jobs[job];
  if (j.status) return fail(j.status, ml_strerror(j.status));
  cudaStream_t st = eng->e.stream;
  const size_t EP = (size_t)j.E * (size_t)j.P;
  if (mu && !R.has_mu) return fail(ML_ERR_BAD_ARGUMENT, "mu was not stored (engine key want_mu = 0)");
  if (mu) ML_CUDA(cudaMemcpyAsync(mu, R.mu.p + j.row0, sizeof(double) * (size_t)j.nrows, cudaMemcpyDeviceToHost, st));
  if (offset) ML_CUDA(cudaMemcpyAsync(offset, R.offset.p + j.off0, sizeof(double) * (size_t)j.D, cudaMemcpyDeviceToHost, st));
  if (beta) ML_CUDA(cudaMemcpyAsync(beta, R.beta.p + j.par0, sizeof(double) * EP, cudaMemcpyDeviceToHost, st));
  if (betahat) ML_CUDA(cudaMemcpyAsync(betahat, R.betahat.p + j.par0, sizeof(double) * EP, cudaMemcpyDeviceToHost, st));
  if (pseudoweight) ML_CUDA(cudaMemcpyAsync(pseudoweight, R.pw.p + j.par0, sizeof(double) * EP, cudaMemcpyDeviceToHost, st));
  if (start)
    for (int e = 0; e < j.E; ++e)
      ML_CUDA(cudaMemcpyAsync(start + (size_t)e * j.P, R.start.p + R.start0[job], sizeof(double) * (size_t)j.P,
                              cudaMemcpyDeviceToHost, st));
  if (lambda2) for (int e = 0; e < j.E; ++e) lambda2[e] = j.lambda2;
  if (estimate_id)
    for (int e = 0; e < j.E; ++e)
      for (int64_t k = 0; k < j.P; ++k) estimate_id[(size_t)e * j.P + k] = e + 1;
  stream_sync(st);
  return ML_OK;
  ML_CATCH
}

static int result_copy_all(ml_engine* eng, ml_result* r, double* mu, double* offset, int32_t* estimate_id,
                           double* start, double* beta, double* betahat, double* pseudoweight, bool async) {
  if (!eng || !r) return fail(ML_ERR_BAD_ARGUMENT, "NULL argument");
  ML_TRY
  ML_CUDA(cudaSetDevice(eng->e.device));
  tls_pool() = eng->e.pool;
  tls_mailbox() = &eng->e.mailbox;
  Result& R = *r->r;
  if (mu && !R.has_mu) return fail(ML_ERR_BAD_ARGUMENT, "mu was not stored (engine key want_mu = 0)");
  // The asynchronous variant copies on the engine's copy stream, so that later compute calls on the main
  // stream overlap with the transfer; the copy stream first waits for whatever the main stream still has
  // in flight for this result (ml_batch_detect does not synchronise at its end).
  cudaStream_t st = async ? eng->e.copy_stream : eng->e.stream;
  if (async) {
    cudaEvent_t ready;
    ML_CUDA(cudaEventCreateWithFlags(&ready, cudaEventDisableTiming));
    ML_CUDA(cudaEventRecord(ready, eng->e.stream));
    ML_CUDA(cudaStreamWaitEvent(st, ready, 0));
    ML_CUDA(cudaEventDestroy(ready));
  }
  // parameter arrays and offsets of the valid jobs are already contiguous in job order.  The tiny
  // offset copy goes first: if its destination is pageable the call blocks until everything queued
  // before it on this stream has completed.
  if (offset && R.noff) ML_CUDA(cudaMemcpyAsync(offset, R.offset.p, sizeof(double) * (size_t)R.noff, cudaMemcpyDeviceToHost, st));
  if (beta && R.npar) ML_CUDA(cudaMemcpyAsync(beta, R.beta.p, sizeof(double) * (size_t)R.npar, cudaMemcpyDeviceToHost, st));
  if (betahat && R.npar) ML_CUDA(cudaMemcpyAsync(betahat, R.betahat.p, sizeof(double) * (size_t)R.npar, cudaMemcpyDeviceToHost, st));
  if (pseudoweight && R.npar) ML_CUDA(cudaMemcpyAsync(pseudoweight, R.pw.p, sizeof(double) * (size_t)R.npar, cudaMemcpyDeviceToHost, st));
  size_t mu_at = 0, par_at = 0;
  for (size_t jn = 0; jn < R.jobs.size(); ++jn) {
    const JobHost& j = R.jobs[jn];
    if (j.status) continue;
    if (mu) {
      // rows of valid jobs are contiguous runs of the batch; copy run by run
      ML_CUDA(cudaMemcpyAsync(mu + mu_at, R.mu.p + j.row0, sizeof(double) * (size_t)j.nrows, cudaMemcpyDeviceToHost, st));
      mu_at += (size_t)j.nrows;
    }
    for (int e = 0; e < j.E; ++e) {
      if (start)
        ML_CUDA(cudaMemcpyAsync(start + par_at, R.start.p + R.start0[jn], sizeof(double) * (size_t)j.P,
                                cudaMemcpyDeviceToHost, st));
      par_at += (size_t)j.P;
    }
  }
  if (estimate_id && R.npar)
    ML_CUDA(cudaMemcpyAsync(estimate_id, R.est_id.p, sizeof(int32_t) * (size_t)R.npar, cudaMemcpyDeviceToHost, st));
  if (async) {
    // the result's buffers are freed in main-stream order: its destructor makes the main stream wait for these copies
    if (!R.copy_done) ML_CUDA(cudaEventCreateWithFlags(&R.copy_done, cudaEventDisableTiming));
    ML_CUDA(cudaEventRecord(R.copy_done, st));
  } else {
    stream_sync(st);
  }
  return ML_OK;
  ML_CATCH
}

int ml_result_copy_all(ml_engine* eng, const ml_result* r, double* mu, double* offset, int32_t* estimate_id,
                       double* start, double* beta, double* betahat, double* pseudoweight) {
  return result_copy_all(eng, const_cast<ml_result*>(r), mu, offset, estimate_id, start, beta, betahat, pseudoweight, false);
}

int ml_result_copy_all_async(ml_engine* eng, const ml_result* r, double* mu, double* offset, int32_t* estimate_id,
                             double* start, double* beta, double* betahat, double* pseudoweight) {
  return result_copy_all(eng, const_cast<ml_result*>(r), mu, offset, estimate_id, start, beta, betahat, pseudoweight, true);
}

void ml_result_free(ml_result* r) {
  if (!r) return;
  if (r->r && r->r->eng) cudaSetDevice(r->r->eng->device);
  delete r;
}

int ml_result_fast_diff_combine(ml_engine* eng, ml_result* r) {
  if (!eng || !r) return fail(ML_ERR_BAD_ARGUMENT, "NULL argument");
  ML_TRY
  ML_CUDA(cudaSetDevice(eng->e.device));
  tls_pool() = eng->e.pool;
  tls_mailbox() = &eng->e.mailbox;
  result_fast_diff_combine(eng->e, *r->r);
  return ML_OK;
  ML_CATCH
}

// ---- segment_methylation_helper -------------------------------------------------------------------
int ml_segment_methylation_helper(ml_engine* eng, int64_t n, const int32_t* estimate_id, const int32_t* start,
                                  const double* beta, const double* betahat, const double* pseudoweight,
                                  double tol_val, int64_t* n_segments, int32_t* seg_estimate_id,
                                  int32_t* seg_id, uint32_t* seg_start, uint32_t* seg_end, double* seg_beta,
                                  double* seg_betahat, double* seg_pseudoweight, double* seg_stdev) {
  if (!eng || !n_segments) return fail(ML_ERR_BAD_ARGUMENT, "NULL argument");
  if (n <= 0 || !estimate_id || !start || !beta || !betahat || !pseudoweight)
    return fail(ML_ERR_BAD_ARGUMENT, "segment_methylation_helper needs a non-empty table");
  ML_TRY
  ML_CUDA(cudaSetDevice(eng->e.device));
  tls_pool() = eng->e.pool;
  tls_mailbox() = &eng->e.mailbox;
  SegHostOut ho;
  ho.estimate_id = seg_estimate_id; ho.seg_id = seg_id; ho.start = seg_start; ho.end = seg_end;
  ho.beta = seg_beta; ho.betahat = seg_betahat; ho.pseudoweight = seg_pseudoweight; ho.stdev = seg_stdev;
  *n_segments = segment_host_table(eng->e, n, estimate_id, start, beta, betahat, pseudoweight, tol_val, ho);
  return ML_OK;
  ML_CATCH
}

int ml_result_segments(ml_engine* eng, ml_result* r, double tol_val, int64_t* n_segments, int32_t* seg_job,
                       int32_t* seg_estimate_id, int32_t* seg_id, uint32_t* seg_start, uint32_t* seg_end,
                       double* seg_beta, double* seg_betahat, double* seg_pseudoweight, double* seg_stdev) {
  if (!eng || !r || !n_segments) return fail(ML_ERR_BAD_ARGUMENT, "NULL argument");
  ML_TRY
  ML_CUDA(cudaSetDevice(eng->e.device));
  tls_pool() = eng->e.pool;
  tls_mailbox() = &eng->e.mailbox;
  Result& R = *r->r;
  if (!R.seg_cache || R.seg_tol != tol_val) {
    std::vector<SegGroup> groups;
    for (size_t jn = 0; jn < R.jobs.size(); ++jn) {
      const JobHost& j = R.jobs[jn];
      if (j.status) continue;
      for (int e = 0; e < j.E; ++e)
        groups.push_back(SegGroup{j.par0 + (int64_t)e * j.P, j.P, R.start0[jn], e + 1, (int32_t)jn});
    }
    auto ds_new = std::make_unique<SegDeviceOut>();
    ds_new->n = segment_device(eng->e, groups, R.npar, R.start.p, 1, R.beta.p, R.betahat.p, R.pw.p, tol_val,
                               *ds_new);
    R.seg_cache = std::move(ds_new);
    R.seg_tol = tol_val;
  }
  SegDeviceOut& ds = *R.seg_cache;
  const int64_t S = ds.n;
  *n_segments = S;
  SegHostOut ho;
  ho.job = seg_job; ho.estimate_id = seg_estimate_id; ho.seg_id = seg_id; ho.start = seg_start;
  ho.end = seg_end; ho.beta = seg_beta; ho.betahat = seg_betahat; ho.pseudoweight = seg_pseudoweight;
  ho.stdev = seg_stdev;
  if (S > 0) ds.download(eng->e.stream, S, ho);
  stream_sync(eng->e.stream);
  return ML_OK;
  ML_CATCH
}

int ml_result_pmd_segments(ml_engine* eng, ml_result* r, double tol_val, double pmd_lambda2,
                           int64_t* n_segments, double* fused_std, int32_t* seg_job, int32_t* seg_estimate_id,
                           int32_t* seg_id, uint32_t* seg_start, uint32_t* seg_end, double* seg_fused_std,
                           double* seg_std, double* seg_pseudoweight) {
  if (!eng || !r || !n_segments) return fail(ML_ERR_BAD_ARGUMENT, "NULL argument");
  ML_TRY
  ML_CUDA(cudaSetDevice(eng->e.device));
  tls_pool() = eng->e.pool;
  tls_mailbox() = &eng->e.mailbox;
  Result& R = *r->r;
  if (!R.seg_cache || R.seg_tol != tol_val) {
    int64_t ns = 0;
    const int rc = ml_result_segments(eng, r, tol_val, &ns, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr,
                                      nullptr, nullptr, nullptr);
    if (rc) return rc;
  }
  SegDeviceOut& seg = *R.seg_cache;
  cudaStream_t st = eng->e.stream;
  if (!R.pmd_cache || R.pmd_tol != tol_val || R.pmd_lambda2 != pmd_lambda2) {
    auto out = std::make_unique<SegDeviceOut>();
    out->n = pmd_segments_device(eng->e, seg, pmd_lambda2, tol_val, R.pmd_fused, *out);
    R.pmd_cache = std::move(out);
    R.pmd_tol = tol_val;
    R.pmd_lambda2 = pmd_lambda2;
  }
  const SegDeviceOut& P = *R.pmd_cache;
  *n_segments = P.n;
  if (fused_std && seg.n) R.pmd_fused.download(fused_std, (size_t)seg.n);
  SegHostOut ho;
  ho.job = seg_job; ho.estimate_id = seg_estimate_id; ho.seg_id = seg_id; ho.start = seg_start; ho.end = seg_end;
  ho.beta = seg_fused_std; ho.betahat = seg_std; ho.pseudoweight = seg_pseudoweight;
  if (P.n > 0) P.download(st, P.n, ho);
  stream_sync(st);
  return ML_OK;
  ML_CATCH
}

// ---- call table (R/call.R:43-68) --------------------------------------------------------------------
namespace {
struct CallHostOut {
  int32_t *job, *estimate_id, *seg_id;
  uint32_t *start, *end;
  double *width, *beta, *coverage, *pseudoweight;
  int32_t* num_cpgs;
  double* std_;
};
void call_download(const CallTable& c, const CallHostOut& h) {
  const size_t n = (size_t)c.t.n;
  if (!n) return;
  if (h.job) c.t.job.download(h.job, n);
  if (h.estimate_id) c.t.estimate_id.download(h.estimate_id, n);
  if (h.seg_id) c.t.seg_id.download(h.seg_id, n);
  if (h.start) c.t.start.download(h.start, n);
  if (h.end) c.t.end.download(h.end, n);
  if (h.width) c.width.download(h.width, n);
  if (h.beta) c.t.beta.download(h.beta, n);
  if (h.coverage) c.coverage.download(h.coverage, n);
  if (h.pseudoweight) c.t.pseudoweight.download(h.pseudoweight, n);
  if (h.num_cpgs) c.num_cpgs.download(h.num_cpgs, n);
  if (h.std_) c.t.stdev.download(h.std_, n);
}
// The (job, estimator) groups the call table and the rule engine work on.  Of a DIFFERENCE fit (ml_result_fast_diff_combine
// applied) only the reference condition is segmented: R/call.R:40-44 keeps estimate_id 1 of the estimates and the rows of
// ret$ref.cond of the data - the first estimator, whose pooled counts the combine leaves untouched (R/fast_diff.R:16, :24).
std::vector<SegGroup> result_groups(const Result& R) {
  std::vector<SegGroup> groups;
  for (size_t jn = 0; jn < R.jobs.size(); ++jn) {
    const JobHost& j = R.jobs[jn];
    if (j.status) continue;
    for (int e = 0; e < (R.diff_combined ? std::min(1, j.E) : j.E); ++e)
      groups.push_back(SegGroup{j.par0 + (int64_t)e * j.P, j.P, R.start0[jn], e + 1, (int32_t)jn});
  }
  return groups;
}
int ensure_call_table(ml_engine* eng, ml_result* r, double tol_val, double max_distance) {
  Result& R = *r->r;
  if (!R.counts_exact)
    return fail(ML_ERR_BAD_ARGUMENT, "the resident call table needs a FAST result with max_iter = 1 (pooled counts are "
                                     "recovered from betahat * pseudoweight); use ml_segment_call_table otherwise");
  const SegDeviceOut* segs = nullptr;
  if (R.diff_combined) {
    if (!R.seg_ref_cache || R.seg_ref_tol != tol_val) {
      auto ds_new = std::make_unique<SegDeviceOut>();
      ds_new->n = segment_device(eng->e, result_groups(R), R.npar, R.start.p, 1, R.beta.p, R.betahat.p, R.pw.p, tol_val, *ds_new);
      R.seg_ref_cache = std::move(ds_new);
      R.seg_ref_tol = tol_val;
      R.call_cache.reset();
    }
    segs = R.seg_ref_cache.get();
  } else {
    if (!R.seg_cache || R.seg_tol != tol_val) {
      int64_t ns = 0;
      const int rc = ml_result_segments(eng, r, tol_val, &ns, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr,
                                        nullptr, nullptr, nullptr);
      if (rc) return rc;
      R.call_cache.reset();
    }
    segs = R.seg_cache.get();
  }
  if (!R.call_cache || R.call_tol != tol_val || R.call_max_distance != max_distance) {
    auto c = std::make_unique<CallTable>();
    call_table_from_estimates(eng->e, result_groups(R), *segs, R.start.p, 1, R.betahat.p, R.pw.p, max_distance, *c);
    R.call_cache = std::move(c);
    R.call_tol = tol_val;
    R.call_max_distance = max_distance;
    R.call_pmd_cache.reset();
  }
  return ML_OK;
}
}  // namespace

int ml_result_call_table(ml_engine* eng, ml_result* r, double tol_val, double max_distance, int64_t* n_rows,
                         int32_t* job, int32_t* estimate_id, int32_t* segment_id, uint32_t* start, uint32_t* end,
                         double* width, double* beta, double* coverage, double* pseudoweight, int32_t* num_cpgs,
                         double* std_) {
  if (!eng || !r || !n_rows) return fail(ML_ERR_BAD_ARGUMENT, "NULL argument");
  ML_TRY
  ML_CUDA(cudaSetDevice(eng->e.device));
  tls_pool() = eng->e.pool;
  tls_mailbox() = &eng->e.mailbox;
  const int rc = ensure_call_table(eng, r, tol_val, max_distance);
  if (rc) return rc;
  const CallTable& c = *r->r->call_cache;
  *n_rows = c.t.n;
  call_download(c, CallHostOut{job, estimate_id, segment_id, start, end, width, beta, coverage, pseudoweight, num_cpgs, std_});
  stream_sync(eng->e.stream);
  return ML_OK;
  ML_CATCH
}

int ml_result_call_pmd_segments(ml_engine* eng, ml_result* r, double tol_val, double max_distance, double pmd_lambda2,
                                int64_t* n_segments, double* fused_std, int32_t* seg_job, int32_t* seg_estimate_id,
                                int32_t* seg_id, uint32_t* seg_start, uint32_t* seg_end, double* seg_fused_std,
                                double* seg_std, double* seg_pseudoweight) {
  if (!eng || !r || !n_segments) return fail(ML_ERR_BAD_ARGUMENT, "NULL argument");
  ML_TRY
  ML_CUDA(cudaSetDevice(eng->e.device));
  tls_pool() = eng->e.pool;
  tls_mailbox() = &eng->e.mailbox;
  const int rc = ensure_call_table(eng, r, tol_val, max_distance);
  if (rc) return rc;
  Result& R = *r->r;
  const CallTable& c = *R.call_cache;
  cudaStream_t st = eng->e.stream;
  if (!R.call_pmd_cache || R.call_pmd_lambda2 != pmd_lambda2) {
    auto out = std::make_unique<SegDeviceOut>();
    out->n = pmd_segments_device(eng->e, c.t, pmd_lambda2, tol_val, R.call_pmd_fused, *out);
    R.call_pmd_cache = std::move(out);
    R.call_pmd_lambda2 = pmd_lambda2;
  }
  const SegDeviceOut& P = *R.call_pmd_cache;
  *n_segments = P.n;
  if (fused_std && c.t.n) R.call_pmd_fused.download(fused_std, (size_t)c.t.n);
  SegHostOut ho;
  ho.job = seg_job; ho.estimate_id = seg_estimate_id; ho.seg_id = seg_id; ho.start = seg_start; ho.end = seg_end;
  ho.beta = seg_fused_std; ho.betahat = seg_std; ho.pseudoweight = seg_pseudoweight;
  if (P.n > 0) P.download(st, P.n, ho);
  stream_sync(st);
  return ML_OK;
  ML_CATCH
}

int ml_segment_call_table(ml_engine* eng, int64_t n, const int32_t* estimate_id, const int32_t* pos, const int32_t* nmeth,
                          const int32_t* coverage, int64_t n_seg, const int32_t* seg_estimate_id, const int32_t* seg_id,
                          const uint32_t* seg_start, const uint32_t* seg_end, const double* seg_pseudoweight,
                          double max_distance, int64_t* n_rows, int32_t* o_estimate_id, int32_t* o_segment_id,
                          uint32_t* o_start, uint32_t* o_end, double* o_width, double* o_beta, double* o_coverage,
                          double* o_pseudoweight, int32_t* o_num_cpgs, double* o_std) {
  if (!eng || !n_rows) return fail(ML_ERR_BAD_ARGUMENT, "NULL argument");
  if (n <= 0 || n_seg <= 0 || !estimate_id || !pos || !nmeth || !coverage || !seg_estimate_id || !seg_id || !seg_start ||
      !seg_end || !seg_pseudoweight)
    return fail(ML_ERR_BAD_ARGUMENT, "the call table needs non-empty pooled rows and segments");
  if (n >= 0x7fffffff || n_seg >= 0x7fffffff) return fail(ML_ERR_BAD_ARGUMENT, "table too large");
  ML_TRY
  ML_CUDA(cudaSetDevice(eng->e.device));
  tls_pool() = eng->e.pool;
  tls_mailbox() = &eng->e.mailbox;
  Engine& e = eng->e;
  cudaStream_t st = e.stream;
  // groups = blocks of equal estimate id, which must appear in the same order in both tables
  std::vector<SegGroup> groups;
  SegDeviceOut seg;
  seg.group_seg0.clear();
  int64_t r0 = 0, s0 = 0;
  while (s0 < n_seg) {
    const int32_t est = seg_estimate_id[s0];
    int64_t s1 = s0;
    while (s1 < n_seg && seg_estimate_id[s1] == est) ++s1;
    while (r0 < n && estimate_id[r0] != est) ++r0;
    int64_t r1 = r0;
    while (r1 < n && estimate_id[r1] == est) ++r1;
    if (r1 == r0) throw MlError(ML_ERR_BAD_ARGUMENT, "a segment's estimate_id has no rows (tables must be sorted by estimate_id)");
    groups.push_back(SegGroup{r0, r1 - r0, r0, est, 0});
    seg.group_seg0.push_back((int32_t)s0);
    seg.group_job.push_back(0);
    seg.group_est.push_back(est);
    r0 = r1;
    s0 = s1;
  }
  seg.group_seg0.push_back((int32_t)n_seg);
  seg.alloc(st, n_seg);
  seg.seg_id.upload(seg_id, (size_t)n_seg);
  seg.start.upload(seg_start, (size_t)n_seg);
  seg.end.upload(seg_end, (size_t)n_seg);
  seg.pseudoweight.upload(seg_pseudoweight, (size_t)n_seg);
  DevBuf<int32_t> d_pos(st, (size_t)n), d_nm(st, (size_t)n), d_cov(st, (size_t)n);
  d_pos.upload(pos, (size_t)n);
  d_nm.upload(nmeth, (size_t)n);
  d_cov.upload(coverage, (size_t)n);
  CallTable c;
  call_table_from_counts(e, groups, seg, d_pos.p, d_nm.p, d_cov.p, max_distance, c);
  *n_rows = c.t.n;
  call_download(c, CallHostOut{nullptr, o_estimate_id, o_segment_id, o_start, o_end, o_width, o_beta, o_coverage,
                               o_pseudoweight, o_num_cpgs, o_std});
  stream_sync(st);
  return ML_OK;
  ML_CATCH
}

int ml_wilcoxon_groups(ml_engine* eng, int64_t n_groups, const int64_t* x_ptr, const double* x, const int64_t* y_ptr,
                       const double* y, double* pval, double* stat) {
  if (!eng) return fail(ML_ERR_BAD_ARGUMENT, "NULL engine");
  if (n_groups < 0) return fail(ML_ERR_BAD_ARGUMENT, "negative number of groups");
  if (n_groups == 0) return ML_OK;
  if (!x_ptr || !y_ptr || !pval) return fail(ML_ERR_BAD_ARGUMENT, "NULL argument");
  if ((x_ptr[n_groups] > x_ptr[0] && !x) || (y_ptr[n_groups] > y_ptr[0] && !y)) return fail(ML_ERR_BAD_ARGUMENT, "NULL sample");
  ML_TRY
  ML_CUDA(cudaSetDevice(eng->e.device));
  tls_pool() = eng->e.pool;
  tls_mailbox() = &eng->e.mailbox;
  wilcoxon_groups(eng->e, n_groups, x_ptr, x, y_ptr, y, pval, stat);
  return ML_OK;
  ML_CATCH
}

int ml_p_adjust_fdr(ml_engine* eng, int64_t n, const double* p, double* q) {
  if (!eng) return fail(ML_ERR_BAD_ARGUMENT, "NULL engine");
  if (n < 0) return fail(ML_ERR_BAD_ARGUMENT, "negative length");
  if (n == 0) return ML_OK;
  if (!p || !q) return fail(ML_ERR_BAD_ARGUMENT, "NULL argument");
  ML_TRY
  ML_CUDA(cudaSetDevice(eng->e.device));
  tls_pool() = eng->e.pool;
  tls_mailbox() = &eng->e.mailbox;
  p_adjust_fdr(eng->e, n, p, q);
  return ML_OK;
  ML_CATCH
}

int ml_result_call_differences(ml_engine* eng, const ml_batch* b, ml_result* r, int ref_condition, double min_diff,
                               double min_delta_diff, double tol_val, double max_distance, int64_t* n_rows,
                               int32_t* job, int32_t* condition, int32_t* estimate_id, int32_t* segment_id,
                               uint32_t* start, uint32_t* end, double* width, double* beta, double* std_,
                               double* coverage, int32_t* num_cpgs, double* beta_ref, double* std_ref,
                               double* coverage_ref, int32_t* num_cpgs_ref, double* pseudoweight, double* diff,
                               double* cohen_d, double* pval, double* coverage_score) {
  if (!eng || !b || !r || !n_rows) return fail(ML_ERR_BAD_ARGUMENT, "NULL argument");
  ML_TRY
  ML_CUDA(cudaSetDevice(eng->e.device));
  tls_pool() = eng->e.pool;
  tls_mailbox() = &eng->e.mailbox;
  Result& R = *r->r;
  const Batch& B = *b->b;
  if (!R.diff_combined)
    return fail(ML_ERR_BAD_ARGUMENT, "Please call call_differences on the result of difference_detection");   // R/call.R:256
  if (B.njobs != (int)R.jobs.size() || B.nrows != R.nrows)
    return fail(ML_ERR_BAD_ARGUMENT, "the batch is not the one this result was computed from");
  const double key[5] = {(double)ref_condition, min_diff, min_delta_diff, tol_val, max_distance};
  if (!R.diff_cache || memcmp(key, R.diff_key, sizeof key) != 0) {
    auto t = std::make_shared<DiffCallTable>();
    call_differences_device(eng->e, B, R, ref_condition, min_diff, min_delta_diff, tol_val, max_distance, *t);
    R.diff_cache = std::move(t);
    memcpy(R.diff_key, key, sizeof key);
  }
  const DiffCallTable& T = *R.diff_cache;
  *n_rows = T.n;
  const size_t n = (size_t)T.n;
  if (n) {
    if (job) T.job.download(job, n);
    if (condition) T.condition.download(condition, n);
    if (estimate_id) T.estimate_id.download(estimate_id, n);
    if (segment_id) T.segment_id.download(segment_id, n);
    if (start) T.start.download(start, n);
    if (end) T.end.download(end, n);
    if (width) T.width.download(width, n);
    if (beta) T.beta.download(beta, n);
    if (std_) T.std_.download(std_, n);
    if (coverage) T.coverage.download(coverage, n);
    if (num_cpgs) T.num_cpgs.download(num_cpgs, n);
    if (beta_ref) T.beta_ref.download(beta_ref, n);
    if (std_ref) T.std_ref.download(std_ref, n);
    if (coverage_ref) T.coverage_ref.download(coverage_ref, n);
    if (num_cpgs_ref) T.num_cpgs_ref.download(num_cpgs_ref, n);
    if (pseudoweight) T.pseudoweight.download(pseudoweight, n);
    if (diff) T.diff.download(diff, n);
    if (cohen_d) T.cohen_d.download(cohen_d, n);
    if (pval) T.pval.download(pval, n);
    if (coverage_score) T.coverage_score.download(coverage_score, n);
  }
  stream_sync(eng->e.stream);
  return ML_OK;
  ML_CATCH
}

int ml_host_alloc(uint64_t bytes, void** out) {
  if (!out) return fail(ML_ERR_BAD_ARGUMENT, "NULL argument");
  *out = nullptr;
  ML_TRY
  ML_CUDA(cudaHostAlloc(out, bytes ? (size_t)bytes : 1, cudaHostAllocPortable));
  return ML_OK;
  ML_CATCH
}
void ml_host_free(void* p) {
  if (p) cudaFreeHost(p);
}

static_assert(sizeof(ml_lmr_params) == sizeof(LmrParams), "ml_lmr_params / LmrParams layout mismatch");

int ml_result_call_lmr_umr(ml_engine* eng, ml_result* r, double tol_val, double max_distance, const ml_lmr_params* params,
                           int64_t* n_rows, int8_t* is_admissible, int8_t* is_v, int8_t* is_flank, double* flank_dist,
                           double* flank_beta, int8_t* is_sep, int8_t* category) {
  if (!eng || !r || !n_rows) return fail(ML_ERR_BAD_ARGUMENT, "NULL argument");
  ML_TRY
  ML_CUDA(cudaSetDevice(eng->e.device));
  tls_pool() = eng->e.pool;
  tls_mailbox() = &eng->e.mailbox;
  const int rc = ensure_call_table(eng, r, tol_val, max_distance);
  if (rc) return rc;
  LmrParams p;
  if (params) memcpy(&p, params, sizeof p);
  LmrUmrColumns out;
  call_lmr_umr_device(eng->e, *r->r->call_cache, p, out);
  *n_rows = out.n;
  const size_t n = (size_t)out.n;
  if (n) {
    if (is_admissible) out.is_admissible.download(is_admissible, n);
    if (is_v) out.is_v.download(is_v, n);
    if (is_flank) out.is_flank.download(is_flank, n);
    if (flank_dist) out.flank_dist.download(flank_dist, n);
    if (flank_beta) out.flank_beta.download(flank_beta, n);
    if (is_sep) out.is_sep.download(is_sep, n);
    if (category) out.category.download(category, n);
  }
  stream_sync(eng->e.stream);
  return ML_OK;
  ML_CATCH
}

static_assert(sizeof(ml_valley_params) == sizeof(ValleyParams), "ml_valley_params / ValleyParams layout mismatch");

int ml_result_call_valleys(ml_engine* eng, ml_result* r, double tol_val, double max_distance, const ml_lmr_params* lmr,
                           const ml_valley_params* params, int64_t* n_valleys, int32_t* job, int32_t* condition,
                           uint32_t* start, uint32_t* end, double* width, double* beta, double* coverage,
                           double* pseudoweight, int32_t* num_cpgs, int64_t* n_rows, int8_t* row_in_valley,
                           double* low_id_width, double* low_id_beta) {
  if (!eng || !r || !n_valleys) return fail(ML_ERR_BAD_ARGUMENT, "NULL argument");
  ML_TRY
  ML_CUDA(cudaSetDevice(eng->e.device));
  tls_pool() = eng->e.pool;
  tls_mailbox() = &eng->e.mailbox;
  const int rc = ensure_call_table(eng, r, tol_val, max_distance);
  if (rc) return rc;
  LmrParams lp;
  if (lmr) memcpy(&lp, lmr, sizeof lp);
  ValleyParams vp;
  if (params) memcpy(&vp, params, sizeof vp);
  else vp.max_distance = max_distance;
  const CallTable& ct = *r->r->call_cache;
  LmrUmrColumns lu;
  call_lmr_umr_device(eng->e, ct, lp, lu);
  ValleyTable v;
  DevBuf<int8_t> row_in;
  DevBuf<double> low_w, low_b;
  call_valleys_device(eng->e, ct, lu.category.p, vp, v, row_in, low_w, low_b);
  *n_valleys = v.n;
  if (n_rows) *n_rows = ct.t.n;
  const size_t nv = (size_t)v.n, n = (size_t)ct.t.n;
  if (nv) {
    if (job) v.job.download(job, nv);
    if (condition) v.condition.download(condition, nv);
    if (start) v.start.download(start, nv);
    if (end) v.end.download(end, nv);
    if (width) v.width.download(width, nv);
    if (beta) v.beta.download(beta, nv);
    if (coverage) v.coverage.download(coverage, nv);
    if (pseudoweight) v.pseudoweight.download(pseudoweight, nv);
    if (num_cpgs) v.num_cpgs.download(num_cpgs, nv);
  }
  if (n) {
    if (row_in_valley) row_in.download(row_in_valley, n);
    if (low_id_width) low_w.download(low_id_width, n);
    if (low_id_beta) low_b.download(low_id_beta, n);
  }
  stream_sync(eng->e.stream);
  return ML_OK;
  ML_CATCH
}

namespace {
// R/call.R:75-168 on the resident tables: categories, valleys, merged table, pieces (shared by the two entries below)
struct PmdSplitState {
  LmrUmrColumns lu;
  ValleyTable v;
  DevBuf<int8_t> row_in;
  DevBuf<double> low_w, low_b;
  MergedDev M;
  PiecesDev P;
  ValleyParams vp;
};
struct RuleState {
  PmdSplitState S;
  bool has_calls = false, has_final = false;
  PmdCallParams cp;
  PmdCallDev calls;
  FinalDev F;
  int regions_drop = -1;     // -1: the region tables have not been gathered for the current calls
  RegionTables regions;
};
// the state for these arguments: the cached one when nothing it was built from has changed, else computed now
int rule_state(ml_engine* eng, ml_result* r, double tol_val, double max_distance, double pmd_lambda2,
               const ml_lmr_params* lmr, const ml_valley_params* params, int split_pmds, RuleState** out) {
  int64_t ns = 0;   // call table + PMD candidate segments (std_call, R/call.R:93-95) resident
  const int rc = ml_result_call_pmd_segments(eng, r, tol_val, max_distance, pmd_lambda2, &ns, nullptr, nullptr, nullptr,
                                             nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
  if (rc) return rc;
  ML_CUDA(cudaSetDevice(eng->e.device));
  tls_pool() = eng->e.pool;
  tls_mailbox() = &eng->e.mailbox;
  LmrParams lp;
  if (lmr) memcpy(&lp, lmr, sizeof lp);
  ValleyParams vp;
  if (params) memcpy(&vp, params, sizeof vp);
  else vp.max_distance = max_distance;
  Result& R = *r->r;
  const CallTable* ct = R.call_cache.get();
  const SegDeviceOut* pc = R.call_pmd_cache.get();
  struct Key { double tol, maxd, lam; LmrParams lp; ValleyParams vp; int split; const void *ct, *pc; } key;
  memset(&key, 0, sizeof key);
  key.tol = tol_val; key.maxd = max_distance; key.lam = pmd_lambda2; key.lp = lp; key.vp = vp; key.split = split_pmds;
  key.ct = ct; key.pc = pc;
  if (!R.rule_cache || R.rule_key.size() != sizeof key || memcmp(R.rule_key.data(), &key, sizeof key) != 0) {
    R.rule_cache.reset();
    auto st = std::make_shared<RuleState>();
    PmdSplitState& S = st->S;
    S.vp = vp;
    call_lmr_umr_device(eng->e, *ct, lp, S.lu);
    call_valleys_device(eng->e, *ct, S.lu.category.p, S.vp, S.v, S.row_in, S.low_w, S.low_b);
    call_pmd_split_device(eng->e, *ct, S.lu.category.p, S.v, S.row_in.p, *pc, S.vp.pmd_valley_min_width, split_pmds, S.M, S.P);
    R.rule_cache = st;
    R.rule_key.assign((const char*)&key, (const char*)&key + sizeof key);
  }
  *out = static_cast<RuleState*>(R.rule_cache.get());
  return ML_OK;
}
// R/call.R:169-199 (and :201-215 when want_final) on the state, computed once per set of PMD parameters
void rule_calls(ml_engine* eng, ml_result* r, RuleState& st, const ml_pmd_params& pp, bool want_final) {
  Result& R = *r->r;
  PmdCallParams cp;
  cp.pmd_std_threshold = pp.pmd_std_threshold;
  cp.pmd_max_beta = pp.pmd_max_beta;
  cp.pmd_valley_min_width = st.S.vp.pmd_valley_min_width;
  cp.max_distance = st.S.vp.max_distance;
  cp.merge_pmds = pp.merge_pmds;
  if (!st.has_calls || memcmp(&cp, &st.cp, sizeof cp) != 0) {
    call_pmds_device(eng->e, result_groups(R), R.start.p, 1, R.betahat.p, R.pw.p, st.S.P, cp, st.calls);
    st.cp = cp;
    st.has_calls = true;
    st.has_final = false;
    st.regions_drop = -1;
  }
  if (want_final && !st.has_final) {
    call_pmd_status_device(eng->e, st.S.M, st.calls, st.S.lu.is_admissible.p, st.F);
    st.has_final = true;
    st.regions_drop = -1;
  }
}
}  // namespace

int ml_result_call_pmd_split(ml_engine* eng, ml_result* r, double tol_val, double max_distance, double pmd_lambda2,
                             const ml_lmr_params* lmr, const ml_valley_params* params, int split_pmds,
                             int64_t* n_merged, int32_t* m_job, int32_t* m_condition, int32_t* m_segment_id,
                             uint32_t* m_start, uint32_t* m_end, int8_t* m_category, int32_t* m_src_row,
                             int32_t* m_src_valley, int8_t* m_follows_large_gap, int64_t* n_pieces, int32_t* p_job,
                             int32_t* p_estimate_id, int32_t* p_segment_id, uint32_t* p_start, uint32_t* p_end,
                             double* p_fused_std, double* p_pseudoweight) {
  if (!eng || !r || !n_merged || !n_pieces) return fail(ML_ERR_BAD_ARGUMENT, "NULL argument");
  ML_TRY
  RuleState* st = nullptr;
  const int rc = rule_state(eng, r, tol_val, max_distance, pmd_lambda2, lmr, params, split_pmds, &st);
  if (rc) return rc;
  MergedDev& M = st->S.M;
  PiecesDev& P = st->S.P;
  *n_merged = M.n;
  *n_pieces = P.n;
  const size_t nm = (size_t)M.n, np = (size_t)P.n;
  if (nm) {
    if (m_job) M.job.download(m_job, nm);
    if (m_condition) M.est.download(m_condition, nm);
    if (m_segment_id) M.segment_id.download(m_segment_id, nm);
    if (m_start) M.start.download(m_start, nm);
    if (m_end) M.end.download(m_end, nm);
    if (m_category) M.category.download(m_category, nm);
    if (m_src_row) M.src_row.download(m_src_row, nm);
    if (m_src_valley) M.src_valley.download(m_src_valley, nm);
    if (m_follows_large_gap) M.follows_large_gap.download(m_follows_large_gap, nm);
  }
  if (np) {
    if (p_job) P.job.download(p_job, np);
    if (p_estimate_id) P.est.download(p_estimate_id, np);
    if (p_segment_id) P.segment_id.download(p_segment_id, np);
    if (p_start) P.start.download(p_start, np);
    if (p_end) P.end.download(p_end, np);
    if (p_fused_std) P.fused_std.download(p_fused_std, np);
    if (p_pseudoweight) P.pw.download(p_pseudoweight, np);
  }
  stream_sync(eng->e.stream);
  return ML_OK;
  ML_CATCH
}

int ml_result_call_pmds(ml_engine* eng, ml_result* r, double tol_val, double max_distance, const ml_lmr_params* lmr,
                        const ml_valley_params* valley, const ml_pmd_params* pmd, int64_t* n_rows, int32_t* job,
                        int32_t* condition, int8_t* category, int32_t* segment_id, uint32_t* start, uint32_t* end,
                        double* width, int32_t* num_cpgs, double* fused_std, double* beta, double* std_) {
  if (!eng || !r || !n_rows) return fail(ML_ERR_BAD_ARGUMENT, "NULL argument");
  ML_TRY
  ml_pmd_params pp{1000.0, 0.15, 0.75, 1, 1};
  if (pmd) pp = *pmd;
  RuleState* st = nullptr;
  const int rc = rule_state(eng, r, tol_val, max_distance, pp.pmd_lambda2, lmr, valley, pp.split_pmds, &st);
  if (rc) return rc;
  rule_calls(eng, r, *st, pp, false);
  PmdCallDev& out = st->calls;
  *n_rows = out.n;
  const size_t n = (size_t)out.n;
  if (n) {
    if (job) out.job.download(job, n);
    if (condition) out.est.download(condition, n);
    if (category) out.category.download(category, n);
    if (segment_id) out.segment_id.download(segment_id, n);
    if (start) out.start.download(start, n);
    if (end) out.end.download(end, n);
    if (width) out.width.download(width, n);
    if (num_cpgs) out.num_cpgs.download(num_cpgs, n);
    if (fused_std) out.fused_std.download(fused_std, n);
    if (beta) out.beta.download(beta, n);
    if (std_) out.std_.download(std_, n);
  }
  stream_sync(eng->e.stream);
  return ML_OK;
  ML_CATCH
}

int ml_result_call_pmd_status(ml_engine* eng, ml_result* r, double tol_val, double max_distance, const ml_lmr_params* lmr,
                              const ml_valley_params* valley, const ml_pmd_params* pmd, int64_t* n_rows, int32_t* job,
                              int32_t* condition, uint32_t* start, uint32_t* end, int8_t* category, int8_t* pmd_status,
                              int32_t* merged_row, int32_t* src_row, int32_t* src_valley) {
  if (!eng || !r || !n_rows) return fail(ML_ERR_BAD_ARGUMENT, "NULL argument");
  ML_TRY
  ml_pmd_params pp{1000.0, 0.15, 0.75, 1, 1};
  if (pmd) pp = *pmd;
  RuleState* st = nullptr;
  const int rc = rule_state(eng, r, tol_val, max_distance, pp.pmd_lambda2, lmr, valley, pp.split_pmds, &st);
  if (rc) return rc;
  rule_calls(eng, r, *st, pp, true);
  FinalDev& F = st->F;
  *n_rows = F.n;
  const size_t n = (size_t)F.n;
  if (n) {
    if (job) F.job.download(job, n);
    if (condition) F.est.download(condition, n);
    if (start) F.start.download(start, n);
    if (end) F.end.download(end, n);
    if (category) F.category.download(category, n);
    if (pmd_status) F.status.download(pmd_status, n);
    if (merged_row) F.merged.download(merged_row, n);
    if (src_row) F.src_row.download(src_row, n);
    if (src_valley) F.src_valley.download(src_valley, n);
  }
  stream_sync(eng->e.stream);
  return ML_OK;
  ML_CATCH
}
int ml_result_segment_methylation(ml_engine* eng, ml_result* r, double tol_val, double max_distance, const ml_lmr_params* lmr,
                                  const ml_valley_params* valley, const ml_pmd_params* pmd, int drop, int64_t* n_seg,
                                  int32_t* job, int32_t* condition, uint32_t* start, uint32_t* end, double* width,
                                  int32_t* segment_id, double* beta, double* coverage, double* pseudoweight, int32_t* num_cpgs,
                                  double* std_, int8_t* category, int8_t* pmd_status, int64_t* n_pmd, int32_t* p_job,
                                  int32_t* p_condition, uint32_t* p_start, uint32_t* p_end, double* p_width,
                                  int32_t* p_segment_id, double* p_beta, double* p_std, int8_t* p_category,
                                  int32_t* p_num_cpgs, double* p_fused_std) {
  if (!eng || !r || !n_seg || !n_pmd) return fail(ML_ERR_BAD_ARGUMENT, "NULL argument");
  ML_TRY
  ml_pmd_params pp{1000.0, 0.15, 0.75, 1, 1};
  if (pmd) pp = *pmd;
  RuleState* st = nullptr;
  const int rc = rule_state(eng, r, tol_val, max_distance, pp.pmd_lambda2, lmr, valley, pp.split_pmds, &st);
  if (rc) return rc;
  rule_calls(eng, r, *st, pp, true);
  drop = drop ? 1 : 0;
  if (st->regions_drop != drop) {
    region_tables_device(eng->e, *r->r->call_cache, st->S.v, st->S.M, st->calls, st->F, drop, st->regions);
    st->regions_drop = drop;
  }
  const RegionTables& T = st->regions;
  *n_seg = T.n_seg;
  *n_pmd = T.n_pmd;
  const size_t ns = (size_t)T.n_seg, np = (size_t)T.n_pmd;
  if (ns) {
    if (job) T.s_job.download(job, ns);
    if (condition) T.s_condition.download(condition, ns);
    if (start) T.s_start.download(start, ns);
    if (end) T.s_end.download(end, ns);
    if (width) T.s_width.download(width, ns);
    if (segment_id) T.s_segment_id.download(segment_id, ns);
    if (beta) T.s_beta.download(beta, ns);
    if (coverage) T.s_coverage.download(coverage, ns);
    if (pseudoweight) T.s_pseudoweight.download(pseudoweight, ns);
    if (num_cpgs) T.s_num_cpgs.download(num_cpgs, ns);
    if (std_) T.s_std.download(std_, ns);
    if (category) T.s_category.download(category, ns);
    if (pmd_status) T.s_pmd_status.download(pmd_status, ns);
  }
  if (np) {
    if (p_job) T.p_job.download(p_job, np);
    if (p_condition) T.p_condition.download(p_condition, np);
    if (p_start) T.p_start.download(p_start, np);
    if (p_end) T.p_end.download(p_end, np);
    if (p_width) T.p_width.download(p_width, np);
    if (p_segment_id) T.p_segment_id.download(p_segment_id, np);
    if (p_beta) T.p_beta.download(p_beta, np);
    if (p_std) T.p_std.download(p_std, np);
    if (p_category) T.p_category.download(p_category, np);
    if (p_num_cpgs) T.p_num_cpgs.download(p_num_cpgs, np);
    if (p_fused_std) T.p_fused_std.download(p_fused_std, np);
  }
  stream_sync(eng->e.stream);
  return ML_OK;
  ML_CATCH
}
}  // extern "C"

// ------------------------------------------------------------------------------------------------
// lambda2 sweep (BASELINE.json configs[4]): the one-step FAST fit's pseudodata (betahat, pseudoweight) do not depend on
// lambda2, so a sweep is one batched FLSA call whose series - one per (chromosome, estimator, lambda2) - all read the ONE
// resident pseudodata stream of their chromosome (FlsaSeries::off_in) and write their own lane of the output; then the
// centring (core/params_helpers.hpp:23-30) and segment_methylation_helper per lane.
// ------------------------------------------------------------------------------------------------
namespace ml {
__global__ void __launch_bounds__(256)
k_sweep_center(int ns, const int64_t* __restrict__ off, const int32_t* __restrict__ len, const double* __restrict__ sums,
               double* __restrict__ beta) {
  const int s = blockIdx.y;
  if (s >= ns) return;
  const double mean = sums[2 * s] / sums[2 * s + 1];
  double* b = beta + off[s];
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < len[s]; k += gridDim.x * blockDim.x) b[k] = b[k] - mean;
}
}  // namespace ml

extern "C" int ml_result_lambda_sweep(ml_engine* eng, ml_result* r, int nlam, const double* lambda2, double lambda1, double tol_val,
                                      int64_t* n_segments, double* beta_out) {
  if (!eng || !r || nlam <= 0 || !lambda2) return fail(ML_ERR_BAD_ARGUMENT, "bad argument");
  ML_TRY
  ML_CUDA(cudaSetDevice(eng->e.device));
  tls_pool() = eng->e.pool;
  tls_mailbox() = &eng->e.mailbox;
  Result& R = *r->r;
  if (!R.counts_exact || R.diff_combined)
    return fail(ML_ERR_BAD_ARGUMENT, "lambda2 sweep: needs the result of a one-step FAST fit (its pseudodata do not depend on lambda2)");
  for (int l = 0; l < nlam; ++l)
    if (!(lambda2[l] > 0)) return fail(13, "lambda2 must be positive");      // fitter_lasso.ipp:4
  cudaStream_t st = eng->e.stream;
  const int64_t npar = R.npar;
  std::vector<FlsaSeries> series;
  std::vector<int64_t> off;
  std::vector<int32_t> len;
  int max_len = 1;
  for (int l = 0; l < nlam; ++l)
    for (const JobHost& j : R.jobs) {
      if (j.status) continue;
      for (int e = 0; e < j.E; ++e) {
        FlsaSeries fs{(int64_t)l * npar + j.par0 + (int64_t)e * j.P, (int32_t)j.P, lambda2[l]};
        fs.off_in = j.par0 + (int64_t)e * j.P;
        series.push_back(fs);
        off.push_back(fs.off);
        len.push_back(fs.n);
        max_len = std::max(max_len, (int)fs.n);
      }
    }
  const int ns = (int)series.size();
  if (ns == 0) { if (n_segments) for (int l = 0; l < nlam; ++l) n_segments[l] = 0; return ML_OK; }
  DevBuf<double> beta(st, (size_t)nlam * (size_t)npar), sums(st, 2 * (size_t)ns);
  {
    FlsaPlan plan(eng->e, series, (int64_t)nlam * npar);
    plan.solve(R.betahat.p, R.pw.p, beta.p, 1, lambda1, sums.p);
  }
  DevBuf<int64_t> d_off(st, (size_t)ns);
  DevBuf<int32_t> d_len(st, (size_t)ns);
  d_off.upload(off);
  d_len.upload(len);
  ML_LAUNCH(eng->e, "k_sweep_center", k_sweep_center<<<dim3((unsigned)std::min(64, cdiv(max_len, 1024)), (unsigned)ns), 256, 0, st>>>(
      ns, d_off.p, d_len.p, sums.p, beta.p));
  ML_WORK(eng->e, "k_sweep_center", 16.0 * (double)nlam * (double)npar);
  if (n_segments) {
    const std::vector<SegGroup> groups = result_groups(R);
    for (int l = 0; l < nlam; ++l) {
      SegDeviceOut ds;
      n_segments[l] = segment_device(eng->e, groups, npar, R.start.p, 1, beta.p + (size_t)l * npar, R.betahat.p, R.pw.p, tol_val, ds);
    }
  }
  if (beta_out) ML_CUDA(cudaMemcpyAsync(beta_out, beta.p, sizeof(double) * (size_t)nlam * (size_t)npar, cudaMemcpyDeviceToHost, st));
  stream_sync(st);
  return ML_OK;
  ML_CATCH
}

// ------------------------------------------------------------------------------------------------
// K16: the exchange of the per-shard region tables (what R/call.R:233-244 does with rbindlist over the chromosomes a
// worker handled, R/genome_wide.R:6-15 with the fits).  One process per GPU; every rank owns a WINDOW in its HBM with
// one slot per rank and parity, exported with cudaIpcGetMemHandle and mapped by the peers.  A rank's push is ONE kernel
// that packs its two region tables and writes them into its slot of EVERY rank's window over NVLink (peer stores): the
// tables never touch the host on their way.  Ordering between ranks (all pushes done before anybody reads) is the
// caller's collective on the counts (torch.distributed); slots of the two parities alternate from call to call, so a
// fast rank's next push cannot overwrite what a slow rank is still reading.
// Slot layout: 64-byte header {n_seg, n_pmd}, then the 13 + 11 columns back to back, each padded to 16 bytes
// (ml_gather_layout is the one statement of it).
// ------------------------------------------------------------------------------------------------
struct ml_gather {
  int device = 0, rank = 0, world = 1;
  int64_t slot_bytes = 0;
  char* local = nullptr;                  // 2 * world * slot_bytes
  std::vector<char*> window;              // per rank: base of that rank's window as mapped here
  std::vector<bool> mapped;
};

namespace {
constexpr int GATHER_COLS = 24;
constexpr int GATHER_MAX_WORLD = 16;
const int kGatherElem[GATHER_COLS] = {4, 4, 4, 4, 8, 4, 8, 8, 8, 4, 8, 1, 1,      // job condition start end width segment_id beta coverage pseudoweight num_cpgs std category pmd_status
                                      4, 4, 4, 4, 8, 4, 8, 8, 1, 4, 8};           // job condition start end width segment_id beta std category num_cpgs fused_std
int64_t gather_layout(int64_t n_seg, int64_t n_pmd, int64_t* off) {
  int64_t at = 64;
  for (int c = 0; c < GATHER_COLS; ++c) {
    off[c] = at;
    at += ((c < 13 ? n_seg : n_pmd) * kGatherElem[c] + 15) / 16 * 16;
  }
  return at;
}
struct PushArgs {
  const char* src[GATHER_COLS];
  int64_t bytes[GATHER_COLS];
  int64_t off[GATHER_COLS];
  char* dst[GATHER_MAX_WORLD];
  int world;
  int64_t n_seg, n_pmd;
  const int32_t* job_ids;                // shard-local job index -> global job (columns 0 and 13)
};
__global__ void __launch_bounds__(256) k_gather_push(PushArgs A) {
  const int c = blockIdx.y;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (int64_t)gridDim.x * blockDim.x;
  if (c == GATHER_COLS) {                // header
    if (tid == 0)
      for (int r = 0; r < A.world; ++r) {
        int64_t* h = reinterpret_cast<int64_t*>(A.dst[r]);
        h[0] = A.n_seg;
        h[1] = A.n_pmd;
      }
    return;
  }
  const bool remap = (c == 0 || c == 13) && A.job_ids != nullptr;
  const int64_t nb = A.bytes[c];
  const int64_t nw = nb / 4;
  const uint32_t* s4 = reinterpret_cast<const uint32_t*>(A.src[c]);
  for (int64_t i = tid; i < nw; i += nth) {
    uint32_t v = s4[i];
    if (remap) v = (uint32_t)A.job_ids[v];
    for (int r = 0; r < A.world; ++r) reinterpret_cast<uint32_t*>(A.dst[r] + A.off[c])[i] = v;
  }
  for (int64_t i = nw * 4 + tid; i < nb; i += nth)
    for (int r = 0; r < A.world; ++r) (A.dst[r] + A.off[c])[i] = A.src[c][i];
}
}  // namespace

extern "C" {

int ml_gather_layout(int64_t n_seg, int64_t n_pmd, int64_t* offsets, int64_t* total_bytes) {
  if (n_seg < 0 || n_pmd < 0 || !total_bytes) return fail(ML_ERR_BAD_ARGUMENT, "bad argument");
  int64_t off[GATHER_COLS];
  *total_bytes = gather_layout(n_seg, n_pmd, off);
  if (offsets) memcpy(offsets, off, sizeof off);
  return ML_OK;
}

int ml_gather_create(ml_engine* eng, int rank, int world, int64_t slot_bytes, ml_gather** out) {
  if (!eng || !out || world < 1 || world > GATHER_MAX_WORLD || rank < 0 || rank >= world || slot_bytes < 64)
    return fail(ML_ERR_BAD_ARGUMENT, "bad argument");
  ML_TRY
  ML_CUDA(cudaSetDevice(eng->e.device));
  auto g = std::make_unique<ml_gather>();
  g->device = eng->e.device;
  g->rank = rank;
  g->world = world;
  g->slot_bytes = (slot_bytes + 255) / 256 * 256;
  // plain cudaMalloc: memory of a stream-ordered pool cannot be exported as an IPC handle
  ML_CUDA(cudaMalloc((void**)&g->local, (size_t)(2 * world * g->slot_bytes)));
  ML_CUDA(cudaMemset(g->local, 0, (size_t)(2 * world * g->slot_bytes)));
  g->window.assign(world, nullptr);
  g->mapped.assign(world, false);
  g->window[rank] = g->local;
  *out = g.release();
  return ML_OK;
  ML_CATCH
}

int ml_gather_ipc_handle(ml_gather* g, void* handle64) {
  if (!g || !handle64) return fail(ML_ERR_BAD_ARGUMENT, "NULL argument");
  ML_TRY
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  ML_CUDA(cudaSetDevice(g->device));
  cudaIpcMemHandle_t h;
  ML_CUDA(cudaIpcGetMemHandle(&h, g->local));
  memcpy(handle64, &h, 64);
  return ML_OK;
  ML_CATCH
}

int ml_gather_open(ml_gather* g, const void* handles) {
  if (!g || !handles) return fail(ML_ERR_BAD_ARGUMENT, "NULL argument");
  ML_TRY
  ML_CUDA(cudaSetDevice(g->device));
  for (int r = 0; r < g->world; ++r) {
    if (r == g->rank || g->mapped[r]) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, (const char*)handles + 64 * (size_t)r, 64);
    void* p = nullptr;
    ML_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    g->window[r] = (char*)p;
    g->mapped[r] = true;
  }
  return ML_OK;
  ML_CATCH
}

int ml_gather_window(ml_gather* g, int parity, void** dev_ptr, int64_t* slot_bytes) {
  if (!g || !dev_ptr) return fail(ML_ERR_BAD_ARGUMENT, "NULL argument");
  *dev_ptr = g->local + (size_t)(parity & 1) * g->world * g->slot_bytes;
  if (slot_bytes) *slot_bytes = g->slot_bytes;
  return ML_OK;
}

int ml_gather_push(ml_engine* eng, ml_gather* g, ml_result* r, double tol_val, double max_distance, const ml_lmr_params* lmr,
                   const ml_valley_params* valley, const ml_pmd_params* pmd, int drop, const int32_t* job_ids, int parity,
                   int peers, int64_t* n_seg, int64_t* n_pmd) {
  if (!eng || !g || !r || !n_seg || !n_pmd) return fail(ML_ERR_BAD_ARGUMENT, "NULL argument");
  ML_TRY
  ml_pmd_params pp{1000.0, 0.15, 0.75, 1, 1};
  if (pmd) pp = *pmd;
  RuleState* st = nullptr;
  const int rc = rule_state(eng, r, tol_val, max_distance, pp.pmd_lambda2, lmr, valley, pp.split_pmds, &st);
  if (rc) return rc;
  rule_calls(eng, r, *st, pp, true);
  drop = drop ? 1 : 0;
  if (st->regions_drop != drop) {
    region_tables_device(eng->e, *r->r->call_cache, st->S.v, st->S.M, st->calls, st->F, drop, st->regions);
    st->regions_drop = drop;
  }
  const RegionTables& T = st->regions;
  *n_seg = T.n_seg;
  *n_pmd = T.n_pmd;
  PushArgs A;
  memset(&A, 0, sizeof A);
  const int64_t total = gather_layout(T.n_seg, T.n_pmd, A.off);
  if (total > g->slot_bytes) return fail(ML_ERR_BAD_ARGUMENT, "ml_gather_push: the tables do not fit the slot");
  const void* src[GATHER_COLS] = {T.s_job.p, T.s_condition.p, T.s_start.p, T.s_end.p, T.s_width.p, T.s_segment_id.p, T.s_beta.p,
                                  T.s_coverage.p, T.s_pseudoweight.p, T.s_num_cpgs.p, T.s_std.p, T.s_category.p,
                                  T.s_pmd_status.p, T.p_job.p, T.p_condition.p, T.p_start.p, T.p_end.p, T.p_width.p,
                                  T.p_segment_id.p, T.p_beta.p, T.p_std.p, T.p_category.p, T.p_num_cpgs.p, T.p_fused_std.p};
  int64_t most = 0;
  for (int c = 0; c < GATHER_COLS; ++c) {
    A.src[c] = (const char*)src[c];
    A.bytes[c] = (c < 13 ? T.n_seg : T.n_pmd) * kGatherElem[c];
    most = std::max(most, A.bytes[c]);
  }
  A.n_seg = T.n_seg;
  A.n_pmd = T.n_pmd;
  // destinations: this rank's slot in every mapped window (peers != 0) or in its own window only (the caller moves
  // the slot with a collective of its own)
  const size_t slot = ((size_t)(parity & 1) * g->world + g->rank) * g->slot_bytes;
  A.world = 0;
  for (int w = 0; w < g->world; ++w) {
    if (w != g->rank && !(peers && g->mapped[w])) continue;
    A.dst[A.world++] = g->window[w] + slot;
  }
  DevBuf<int32_t> d_ids;
  if (job_ids) {
    d_ids.alloc(eng->e.stream, r->r->jobs.size());
    d_ids.upload(job_ids, r->r->jobs.size());
    A.job_ids = d_ids.p;
  }
  const dim3 grid((unsigned)std::max<int64_t>(1, std::min<int64_t>(64, (most / 4 + 255) / 256)), GATHER_COLS + 1);
  ML_LAUNCH(eng->e, "k_gather_push", k_gather_push<<<grid, 256, 0, eng->e.stream>>>(A));
  eng->e.prof.add_work("k_gather_push", (double)total * (1 + A.world));
  stream_sync(eng->e.stream);      // peer stores have landed when this returns
  return ML_OK;
  ML_CATCH
}

int ml_gather_fetch(ml_engine* eng, ml_gather* g, int parity, const int64_t* used_bytes, void* host_dst) {
  if (!eng || !g || !used_bytes || !host_dst) return fail(ML_ERR_BAD_ARGUMENT, "NULL argument");
  ML_TRY
  ML_CUDA(cudaSetDevice(g->device));
  for (int w = 0; w < g->world; ++w) {
    if (used_bytes[w] < 0 || used_bytes[w] > g->slot_bytes) return fail(ML_ERR_BAD_ARGUMENT, "ml_gather_fetch: bad size");
    if (!used_bytes[w]) continue;
    const size_t slot = ((size_t)(parity & 1) * g->world + w) * g->slot_bytes;
    ML_CUDA(cudaMemcpyAsync((char*)host_dst + (size_t)w * g->slot_bytes, g->local + slot, (size_t)used_bytes[w],
                            cudaMemcpyDeviceToHost, eng->e.stream));
  }
  stream_sync(eng->e.stream);
  return ML_OK;
  ML_CATCH
}

void ml_gather_destroy(ml_gather* g) {
  if (!g) return;
  cudaSetDevice(g->device);
  for (int w = 0; w < g->world; ++w)
    if (w != g->rank && g->mapped[w]) cudaIpcCloseMemHandle(g->window[w]);
  cudaFree(g->local);
  delete g;
}

}  // extern "C"

namespace ml {
__global__ void k_txt_gather_runs(int64_t n_runs, const int64_t* __restrict__ run_first, const int32_t* __restrict__ chr_off,
                                  const int32_t* __restrict__ chr_len, int32_t* __restrict__ off, int32_t* __restrict__ len) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n_runs) return;
  off[k] = chr_off[run_first[k]];
  len[k] = chr_len[run_first[k]];
}
__global__ void k_txt_data_line_starts(int64_t n, int64_t skip, const int64_t* __restrict__ line_start, int64_t* __restrict__ out) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r < n) out[r] = line_start[r + skip];
}
}  // namespace ml

extern "C" {

int ml_text_parse(ml_engine* eng, const char* text, int64_t nbytes, int64_t skip_lines, int col_a, int col_b, char sep,
                  ml_text** out, int64_t* n_lines, int64_t* n_chr_runs) {
  if (!eng || !out || !n_lines || (nbytes > 0 && !text) || nbytes < 0) return fail(ML_ERR_BAD_ARGUMENT, "bad argument");
  *out = nullptr;
  ML_TRY
  ML_CUDA(cudaSetDevice(eng->e.device));
  tls_pool() = eng->e.pool;
  tls_mailbox() = &eng->e.mailbox;
  auto t = std::make_unique<ml_text>();
  t->device = eng->e.device;
  text_parse(eng->e, text, nbytes, skip_lines, col_a, col_b, sep, t->t);
  *n_lines = t->t.n;
  if (n_chr_runs) *n_chr_runs = t->t.n_runs;
  *out = t.release();
  return ML_OK;
  ML_CATCH
}

int ml_text_fetch(ml_engine* eng, const ml_text* t, int32_t* pos, double* a, double* b, int8_t* status, int64_t* line_start,
                  int64_t* run_first_line, int32_t* run_chr_off, int32_t* run_chr_len) {
  if (!eng || !t) return fail(ML_ERR_BAD_ARGUMENT, "NULL argument");
  ML_TRY
  ML_CUDA(cudaSetDevice(eng->e.device));
  tls_pool() = eng->e.pool;
  tls_mailbox() = &eng->e.mailbox;
  const TextTable& T = t->t;
  cudaStream_t st = eng->e.stream;
  const size_t n = (size_t)T.n, nr = (size_t)T.n_runs;
  if (n) {
    if (pos) T.pos.download(pos, n);
    if (a) T.a.download(a, n);
    if (b) T.b.download(b, n);
    if (status) T.status.download(status, n);
    DevBuf<int64_t> ls;
    if (line_start) {
      ls.alloc(st, n);
      k_txt_data_line_starts<<<cdiv((long long)n, 256), 256, 0, st>>>((int64_t)n, T.skip, T.line_start.p, ls.p);
      ML_CHECK_LAUNCH();
      ls.download(line_start, n);
    }
    DevBuf<int32_t> ro, rl;
    if (nr) {
      if (run_first_line) T.run_first.download(run_first_line, nr);
      if (run_chr_off || run_chr_len) {
        ro.alloc(st, nr); rl.alloc(st, nr);
        k_txt_gather_runs<<<cdiv((long long)nr, 256), 256, 0, st>>>((int64_t)nr, T.run_first.p, T.chr_off.p, T.chr_len.p, ro.p, rl.p);
        ML_CHECK_LAUNCH();
        if (run_chr_off) ro.download(run_chr_off, nr);
        if (run_chr_len) rl.download(run_chr_len, nr);
      }
    }
    stream_sync(st);
  }
  return ML_OK;
  ML_CATCH
}

void ml_text_free(ml_text* t) {
  if (!t) return;
  cudaSetDevice(t->device);
  delete t;
}

}  // extern "C"

// hostpack.cpp — see hostpack.h.  Built with the host compiler (-O3); the two streaming loops are cloned for AVX2 and
// chosen at load time (the build machine has no say about the CPU of the GPU box).
#include "hostpack.h"

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <mutex>
#include <thread>

namespace ml {

namespace {

class Pool {
 public:
  void run(int n_items, int threads, const std::function<void(int)>& fn) {
    std::lock_guard<std::mutex> serial(run_mutex_);
    const int want = std::max(0, std::min(threads, n_items) - 1);     // helpers besides the caller
    {
      std::unique_lock<std::mutex> lk(m_);
      while ((int)workers_.size() < want) workers_.emplace_back([this, id = (int)workers_.size()] { loop(id); });
      fn_ = &fn;
      n_items_ = n_items;
      next_.store(0);
      active_ = want;
      helpers_ = want;
      ++generation_;
    }
    cv_.notify_all();
    drain();
    std::unique_lock<std::mutex> lk(m_);
    done_cv_.wait(lk, [this] { return active_ == 0; });
    fn_ = nullptr;
  }

 private:
  void drain() {
    for (;;) {
      const int i = next_.fetch_add(1);
      if (i >= n_items_) return;
      (*fn_)(i);
    }
  }
  void loop(int id) {
    unsigned long long seen = 0;
    for (;;) {
      {
        std::unique_lock<std::mutex> lk(m_);
        cv_.wait(lk, [&] { return generation_ != seen && id < helpers_; });
        seen = generation_;
      }
      drain();
      {
        std::lock_guard<std::mutex> lk(m_);
        if (--active_ == 0) done_cv_.notify_all();
      }
    }
  }
  std::mutex run_mutex_, m_;
  std::condition_variable cv_, done_cv_;
  std::vector<std::thread> workers_;      // never joined: they sleep until the process ends
  const std::function<void(int)>* fn_ = nullptr;
  std::atomic<int> next_{0};
  int n_items_ = 0, active_ = 0, helpers_ = 0;
  unsigned long long generation_ = 0;
};

Pool& pool() {
  static Pool* p = new Pool();            // leaked on purpose: worker threads may outlive static destructors
  return *p;
}

template <typename T>
inline bool pack_impl(const int32_t* __restrict__ nm, const int32_t* __restrict__ cv, int64_t lo, int64_t hi, T* __restrict__ out) {
  constexpr uint32_t limit = (uint32_t)((1ull << (8 * sizeof(T))) - 1);
  uint32_t over = 0;
  for (int64_t i = lo; i < hi; ++i) {
    const uint32_t a = (uint32_t)nm[i], c = (uint32_t)cv[i];
    over |= (a | c) & ~limit;             // negative values have their top bits set
    out[2 * i] = (T)a;
    out[2 * i + 1] = (T)c;
  }
  return over == 0;
}

}  // namespace

void parallel_for(int n_items, int threads, const std::function<void(int)>& fn) {
  if (n_items <= 0) return;
  if (threads <= 1 || n_items == 1) {
    for (int i = 0; i < n_items; ++i) fn(i);
    return;
  }
  pool().run(n_items, threads, fn);
}

__attribute__((target_clones("avx2", "default")))
bool pack_counts_u8(const int32_t* nmeth, const int32_t* cov, int64_t shift, int64_t lo, int64_t hi, uint8_t* out) {
  return pack_impl<uint8_t>(nmeth - shift, cov - shift, lo, hi, out);
}
__attribute__((target_clones("avx2", "default")))
bool pack_counts_u16(const int32_t* nmeth, const int32_t* cov, int64_t shift, int64_t lo, int64_t hi, uint16_t* out) {
  return pack_impl<uint16_t>(nmeth - shift, cov - shift, lo, hi, out);
}

__attribute__((target_clones("avx2", "default")))
static bool any_change(const int32_t* v, int64_t i, int64_t e) {
  const int32_t first = v[i - 1];
  int32_t acc = 0;
  for (int64_t k = i; k < e; ++k) acc |= v[k] ^ first;
  return acc != 0;
}

void transitions_in(const int32_t* v, int64_t lo, int64_t hi, std::vector<int64_t>& out) {
  constexpr int64_t B = 4096;
  int64_t i = std::max<int64_t>(lo, 1);
  while (i < hi) {
    const int64_t e = std::min(hi, i + B);
    if (any_change(v, i, e))
      for (int64_t k = i; k < e; ++k)
        if (v[k] != v[k - 1]) out.push_back(k);
    i = e;
  }
}

}  // namespace ml

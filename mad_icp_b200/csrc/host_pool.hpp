// host_pool.hpp -- the library's host-side worker pool (one per process), shared by the tree builder
// (flat_tree.cpp) and the ingest helpers (ingest.cpp).
#pragma once
#include <emmintrin.h>
#include <pthread.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

namespace madicp_host {

// > 0 while a library call that strings several parallel sections together is running (HotScope):
// workers then keep spinning through its serial stretches instead of going to sleep between sections.
inline std::atomic<int> g_hot{0};
struct HotScope {
  HotScope() { g_hot.fetch_add(1, std::memory_order_relaxed); }
  ~HotScope() { g_hot.fetch_sub(1, std::memory_order_relaxed); }
  HotScope(const HotScope&) = delete;
  HotScope& operator=(const HotScope&) = delete;
};

// Worker pool: run(n, fn) executes fn(i) for i in [0,n) on the workers plus the calling thread and
// returns when all are done.  A build issues a few dozen short parallel sections back to back, so the
// workers spin briefly between sections before they go to sleep.
class Pool {
public:
  explicit Pool(int threads) {
    const int extra = threads > 1 ? threads - 1 : 0;
    for (int i = 0; i < extra; ++i) workers_.emplace_back([this, i]() { loop(i + 1); });
  }
  ~Pool() {
    {
      std::lock_guard<std::mutex> lk(mu_);
      stop_ = true;
      gen_.fetch_add(1, std::memory_order_release);
    }
    cv_.notify_all();
    for (std::thread& t : workers_) t.join();
  }
  int threads() const { return int(workers_.size()) + 1; }

  // run(): indices handed out one at a time.  run_blocked(): worker w takes the w-th contiguous share, so
  // a thread meets the same part of an array in consecutive sections (and levels) and finds it in its
  // own cache; use it when the per-index cost is uniform.
  template <class F>
  void run_blocked(size_t n, F&& fn) {
    blocked_ = true;
    run(n, fn);
    blocked_ = false;
  }
  template <class F>
  void run(size_t n, F&& fn) {
    if (n == 0) return;
    if (workers_.empty() || n == 1) {
      for (size_t i = 0; i < n; ++i) fn(i);
      return;
    }
    // One Job object per call, handed to the workers by shared_ptr: the call returns when all n TASKS are done, not
    // when all workers have reported -- a worker the OS wakes late (busy host) finds nothing left and goes back to
    // waiting, instead of holding the whole section up; it can never touch the next call's job.
    auto job = std::make_shared<Job>();
    job->fn = [&fn](size_t i) { fn(i); };
    job->n = n;
    job->blocked = blocked_;
    job->shares = workers_.size() + 1;
    {
      std::lock_guard<std::mutex> lk(mu_);
      job_ = job;
      gen_.fetch_add(1, std::memory_order_release);
    }
    cv_.notify_all();
    work(*job, 0);
    for (int spin = 0; job->finished.load(std::memory_order_acquire) != n; ++spin)
      if (spin > 1500) std::this_thread::yield(); else _mm_pause();
  }

private:
  struct Job {
    std::function<void(size_t)> fn;  // set before the job is published, never changed afterwards
    size_t n = 0, shares = 1;
    bool blocked = false;
    std::atomic<size_t> next{0}, finished{0};
  };
  static void work(Job& j, int me) {
    if (j.blocked) {  // (every share has an owner: a blocked section does wait for its slowest worker)
      for (size_t i = j.n * size_t(me) / j.shares, e = j.n * size_t(me + 1) / j.shares; i < e; ++i) {
        j.fn(i);
        j.finished.fetch_add(1, std::memory_order_release);
      }
      return;
    }
    for (size_t i = j.next.fetch_add(1, std::memory_order_relaxed); i < j.n; i = j.next.fetch_add(1, std::memory_order_relaxed)) {
      j.fn(i);
      j.finished.fetch_add(1, std::memory_order_release);
    }
  }
  void loop(int me) {
    uint64_t seen = 0;
    for (;;) {
      // Sections of one build follow each other within microseconds, but a section with few tasks (the
      // sum chains of the root: three) leaves most workers idle for its whole length: keep spinning for
      // a while (kSpinNs) before paying a futex sleep + wake-up.
      int spin = 0;
      std::chrono::steady_clock::time_point idle_since;
      while (gen_.load(std::memory_order_acquire) == seen) {
        _mm_pause();
        if ((++spin & 255) == 0) {
          const auto nowt = std::chrono::steady_clock::now();
          if (spin == 256) idle_since = nowt;
          if (nowt - idle_since > std::chrono::nanoseconds(kSpinNs) &&
              (g_hot.load(std::memory_order_relaxed) == 0 || nowt - idle_since > std::chrono::nanoseconds(40 * kSpinNs))) {
            std::unique_lock<std::mutex> lk(mu_);
            cv_.wait(lk, [&]() { return gen_.load(std::memory_order_acquire) != seen; });
          }
        }
      }
      std::shared_ptr<Job> job;
      {
        std::lock_guard<std::mutex> lk(mu_);
        seen = gen_.load(std::memory_order_acquire);
        job = job_;
        if (stop_) return;
      }
      if (job) work(*job, me);
    }
  }
  static constexpr long kSpinNs = 500000;
  std::vector<std::thread> workers_;
  std::mutex mu_;
  std::condition_variable cv_;
  std::shared_ptr<Job> job_;  // the current section (guarded by mu_)
  bool blocked_ = false;
  std::atomic<uint64_t> gen_{0};
  bool stop_ = false;
};

// One pool per process, re-created when a build asks for a different width; a build that finds it busy
// (another host thread is building) uses a private one.
inline std::mutex g_pool_mu;
inline std::unique_ptr<Pool> g_pool;

// fork(): the child inherits the pool object but none of its threads.  Forget it there (the object is
// leaked on purpose: its destructor would join threads that do not exist) and start from a fresh lock.
struct ForkGuard {
  ForkGuard() {
    pthread_atfork(nullptr, nullptr, []() {
      (void) g_pool.release();
      new (&g_pool_mu) std::mutex;
    });
  }
};
inline ForkGuard g_fork_guard;

// std::vector that leaves trivially-constructible elements uninitialised on resize()
template <class T>
struct default_init_allocator : std::allocator<T> {
  template <class U> struct rebind { using other = default_init_allocator<U>; };
  using std::allocator<T>::allocator;
  template <class U> void construct(U* p) noexcept { ::new (static_cast<void*>(p)) U; }
  template <class U, class... Args> void construct(U* p, Args&&... args) { ::new (static_cast<void*>(p)) U(std::forward<Args>(args)...); }
};

// fn(chunk_begin, chunk_end) over [0,n) on the process pool when it is free, else on this thread.
template <class F>
void for_chunks(int threads, size_t n, size_t chunk, F&& fn) {
  const size_t nc = (n + chunk - 1) / chunk;
  auto one = [&](size_t c) { fn(c * chunk, std::min(n, (c + 1) * chunk)); };
  std::unique_lock<std::mutex> lk(g_pool_mu, std::try_to_lock);
  if (threads > 1 && nc > 1 && lk.owns_lock()) {
    if (!g_pool || g_pool->threads() != threads) g_pool.reset(new Pool(threads));
    g_pool->run(nc, one);
  } else {
    for (size_t c = 0; c < nc; ++c) one(c);
  }
}


}  // namespace madicp_host

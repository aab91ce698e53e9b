// kc_hostcopy.h — process-wide helper pool for the host copies of pageable sensor data (1 - 100 MB)
// into page-locked staging memory.
//
// Why: the GPU cannot DMA from pageable memory, and one thread copies at ~10 GB/s, so a 1.6 MB
// PointCloud2 message costs 0.13 ms before the first byte crosses PCIe - more than the whole check
// it feeds (critical zone 0.05 ms, DWA cycle 0.12 ms from page-locked input). The pool's threads copy
// pieces side by side with the calling thread, which hands every finished run of pieces to the DMA
// engine in order: host copy and PCIe transfer overlap and the copy runs at several cores' bandwidth.
//
// Workers spin for a short while after a job (a control loop calls again within microseconds to
// milliseconds) and then park on a condition variable; a parked pool costs the caller one notify and
// the caller never waits for a worker: it takes pieces itself until none is left, so a pool that is
// asleep, busy with another handle or absent (one core) degrades to the plain single-thread copy.
#pragma once

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

namespace kc {

class CopyPool {
 public:
  static CopyPool &instance() {
    static CopyPool pool;
    return pool;
  }

  // Copy `bytes` from `src` to `stage` in pieces of `piece` bytes. `flush(off, len)` runs on the
  // calling thread for consecutive finished ranges, in increasing order, covering [0, bytes) exactly
  // once (the caller queues the DMA of that range there). A range is handed over once it holds at
  // least `min_flush` bytes or reaches the end (every flush is a driver call of a few microseconds on
  // the thread that also copies). Returns when every piece was flushed.
  template <class Flush>
  void copy(uint8_t *stage, const uint8_t *src, size_t bytes, size_t piece, Flush &&flush, size_t min_flush = 0) {
    if (bytes == 0) return;
    const size_t n = (bytes + piece - 1) / piece;
    std::unique_lock<std::mutex> job_lock(submit_, std::try_to_lock);
    if (!job_lock.owns_lock() || n < 2 || workers_.empty()) {  // pool busy with another handle: plain copy
      size_t start = 0;
      for (size_t off = 0; off < bytes; off += piece) {
        const size_t end = off + std::min(piece, bytes - off);
        memcpy(stage + off, src + off, end - off);
        if (end == bytes || end - start >= min_flush) {
          flush(start, end - start);
          start = end;
        }
      }
      return;
    }
    if (n > done_cap_) {
      done_.reset(new std::atomic<uint8_t>[n]);
      done_cap_ = n;
    }
    for (size_t i = 0; i < n; ++i) done_[i].store(0, std::memory_order_relaxed);
    stage_ = stage;
    src_ = src;
    bytes_ = bytes;
    piece_ = piece;
    n_ = n;
    next_.store(0, std::memory_order_relaxed);
    gen_.fetch_add(1);  // odd: job open
    if (sleepers_.load() > 0) {
      std::lock_guard<std::mutex> g(park_);
      cv_.notify_all();
    }
    size_t flushed = 0;
    auto drain = [&] {
      size_t j = flushed;
      while (j < n && done_[j].load(std::memory_order_acquire)) ++j;
      if (j > flushed && (j == n || (j - flushed) * piece >= min_flush)) {
        const size_t off = flushed * piece, end = std::min(bytes, j * piece);
        flush(off, end - off);
        flushed = j;
      }
    };
    for (size_t i = next_.fetch_add(1); i < n; i = next_.fetch_add(1)) {
      const size_t off = i * piece;
      memcpy(stage + off, src + off, std::min(piece, bytes - off));
      done_[i].store(1, std::memory_order_release);
      drain();
    }
    while (flushed < n) {
      drain();
      relax();
    }
    gen_.fetch_add(1);  // even: closed; a worker that enters from now on leaves without touching the job
    while (inside_.load() != 0) relax();
  }

  int workers() const { return (int)workers_.size(); }

  // piece size for a copy of `bytes`: small buffers stay one piece (single-thread copy), otherwise
  // about two pieces per thread, between 64 KB and 1 MB
  static size_t piece_for(size_t bytes) {
    if (bytes < (256u << 10)) return std::max<size_t>(bytes, 1);
    size_t p = (bytes / 16 + 0xFFFF) & ~(size_t)0xFFFF;
    return std::min<size_t>(std::max<size_t>(p, 64u << 10), 1u << 20);
  }

 private:
  CopyPool() {
    const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
    int want = (int)std::min<unsigned>(7, hw > 2 ? hw / 2 - 1 : 0);
    if (const char *e = getenv("KOMPASS_B200_COPY_THREADS")) want = std::max(0, std::min(31, atoi(e)));
    if (const char *e = getenv("KOMPASS_B200_COPY_SPIN_US")) spin_us_ = std::max(0, atoi(e));
    for (int t = 0; t < want; ++t) workers_.emplace_back([this] { run(); });
  }
  ~CopyPool() {
    {
      std::lock_guard<std::mutex> g(park_);
      stop_.store(true);
      cv_.notify_all();
    }
    for (std::thread &t : workers_) t.join();
  }
  CopyPool(const CopyPool &) = delete;

  static void relax() {
#if defined(__x86_64__)
    __builtin_ia32_pause();
#endif
  }

  void run() {
    uint64_t seen = 0;
    while (!stop_.load(std::memory_order_relaxed)) {
      // wait for an open job this worker has not served: spin first, then park
      const auto t0 = std::chrono::steady_clock::now();
      uint64_t g;
      for (unsigned spins = 0;; ++spins) {
        g = gen_.load();
        if (((g & 1u) && g != seen) || stop_.load(std::memory_order_relaxed)) break;
        relax();
        if ((spins & 255u) == 255u &&
            std::chrono::steady_clock::now() - t0 > std::chrono::microseconds(spin_us_)) {
          std::unique_lock<std::mutex> lk(park_);
          sleepers_.fetch_add(1);
          cv_.wait(lk, [&] {
            const uint64_t q = gen_.load();
            return stop_.load() || ((q & 1u) && q != seen);
          });
          sleepers_.fetch_sub(1);
        }
      }
      if (stop_.load(std::memory_order_relaxed)) break;
      inside_.fetch_add(1);
      if (gen_.load() == g) {  // still the same open job: its fields are stable while we are inside
        seen = g;
        uint8_t *stage = stage_;
        const uint8_t *src = src_;
        const size_t bytes = bytes_, piece = piece_, n = n_;
        for (size_t i = next_.fetch_add(1); i < n; i = next_.fetch_add(1)) {
          const size_t off = i * piece;
          memcpy(stage + off, src + off, std::min(piece, bytes - off));
          done_[i].store(1, std::memory_order_release);
        }
      }
      inside_.fetch_sub(1);
    }
  }

  std::vector<std::thread> workers_;
  std::mutex submit_, park_;
  std::condition_variable cv_;
  std::atomic<uint64_t> gen_{0};
  std::atomic<int> inside_{0}, sleepers_{0};
  std::atomic<bool> stop_{false};
  std::atomic<size_t> next_{0};
  std::unique_ptr<std::atomic<uint8_t>[]> done_;
  size_t done_cap_ = 0;
  uint8_t *stage_ = nullptr;
  const uint8_t *src_ = nullptr;
  size_t bytes_ = 0, piece_ = 0, n_ = 0;
  int spin_us_ = 200;
};

}  // namespace kc

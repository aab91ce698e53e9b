// Host-only test of the page-locked staging copy pool (kompass-core_b200/csrc/kc_hostcopy.h):
// random sizes and piece sizes, several handles' threads calling at once, parked and spinning workers.
// Checks: every byte copied, flush ranges in order, contiguous, covering [0, bytes) exactly once.
#include <cstdio>
#include <random>
#include <thread>
#include <vector>

#include "../../kompass-core_b200/csrc/kc_hostcopy.h"

static std::atomic<int> failures{0};
#define CHECK(c)                                                   \
  do {                                                             \
    if (!(c)) {                                                    \
      std::printf("FAILED %s:%d %s\n", __FILE__, __LINE__, #c);    \
      failures.fetch_add(1);                                       \
    }                                                              \
  } while (0)

static void one_copy(std::mt19937 &rng, size_t bytes, size_t piece, size_t min_flush = 0) {
  std::vector<uint8_t> src(bytes), dst(bytes + 64, 0xEE), via(bytes + 64, 0xEE);
  for (size_t i = 0; i < bytes; ++i) src[i] = (uint8_t)rng();
  size_t expect = 0;
  int calls = 0;
  kc::CopyPool::instance().copy(via.data(), src.data(), bytes, piece, [&](size_t off, size_t len) {
    CHECK(off == expect);
    CHECK(len > 0 && off + len <= bytes);
    memcpy(dst.data() + off, via.data() + off, len);  // what the DMA would read must be there already
    if (off + len < bytes) CHECK(len >= min_flush);
    expect = off + len;
    ++calls;
  }, min_flush);
  CHECK(expect == bytes);
  CHECK(bytes == 0 || calls >= 1);
  CHECK(memcmp(dst.data(), src.data(), bytes) == 0);
  for (size_t i = bytes; i < bytes + 64; ++i) CHECK(dst[i] == 0xEE && via[i] == 0xEE);
}

int main() {
  std::printf("copy pool workers: %d\n", kc::CopyPool::instance().workers());
  CHECK(kc::CopyPool::piece_for(1000) == 1000);
  CHECK(kc::CopyPool::piece_for(1600000) >= (64u << 10) && kc::CopyPool::piece_for(1600000) <= (1u << 20));
  CHECK(kc::CopyPool::piece_for(100u << 20) == (1u << 20));
  {
    std::mt19937 rng(1);
    one_copy(rng, 0, 4096);
    one_copy(rng, 1, 4096);
    one_copy(rng, 4096, 4096);
    one_copy(rng, 4097, 4096);
    for (int it = 0; it < 300; ++it) one_copy(rng, rng() % (3u << 20), 1 + rng() % (256u << 10));
    for (int it = 0; it < 100; ++it) {
      const size_t bytes = rng() % (3u << 20);
      one_copy(rng, bytes, 1 + rng() % (256u << 10), bytes / (1 + rng() % 5));
    }
    std::this_thread::sleep_for(std::chrono::milliseconds(5));  // workers park
    for (int it = 0; it < 20; ++it) {
      one_copy(rng, 1600000, kc::CopyPool::piece_for(1600000));
      if (it % 4 == 0) std::this_thread::sleep_for(std::chrono::milliseconds(2));
    }
  }
  {  // several callers at once: the losers of the submit lock copy on their own thread
    std::vector<std::thread> ts;
    for (int t = 0; t < 4; ++t)
      ts.emplace_back([t] {
        std::mt19937 rng(100 + t);
        for (int it = 0; it < 100; ++it) one_copy(rng, rng() % (2u << 20), 4096 + rng() % (128u << 10));
      });
    for (auto &t : ts) t.join();
  }
  if (failures.load()) {
    std::printf("%d FAILURES\n", failures.load());
    return 1;
  }
  std::printf("ALL PASSED\n");
  return 0;
}

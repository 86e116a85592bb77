// Concurrent find() driver: the reference calls Index::find from every thread of its Qt pool at once
// (src/database.cpp:1400-1432, read lock at :1698). T std::threads call cb_dct_index_find over one index for a
// fixed time; results are compared with the same needles served one at a time.
//   find_bench <rows> <threads> <seconds> [dht]      prints one JSON line
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>

#include "../include/cbird_b200.h"

static uint64_t splitmix(uint64_t& s) {
  uint64_t z = (s += 0x9E3779B97F4A7C15ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

static uint64_t checksum(const cb_match* m, int64_t n) {
  uint64_t c = uint64_t(n) * 0x9E3779B97F4A7C15ull;
  for (int64_t i = 0; i < n; ++i) c = (c ^ (uint64_t(m[i].mediaId) << 8 | uint64_t(m[i].score))) * 0x100000001B3ull;
  return c;
}

int main(int argc, char** argv) {
  const int64_t rows = argc > 1 ? atoll(argv[1]) : (1 << 20);
  const int threads = argc > 2 ? atoi(argv[2]) : 32;
  const double seconds = argc > 3 ? atof(argv[3]) : 2.0;
  const int dht = argc > 4 ? atoi(argv[4]) : 5;
  std::vector<uint64_t> h(rows);
  std::vector<uint32_t> ids(rows);
  uint64_t s = 12345;
  for (int64_t i = 0; i < rows; ++i) {
    h[i] = splitmix(s) << 1;
    if (!h[i]) h[i] = 2;
    ids[i] = uint32_t(i + 1);
    if (i > 16 && (splitmix(s) % 10) == 0) {  // planted near-duplicate of an earlier row
      uint64_t v = h[splitmix(s) % uint64_t(i)];
      const int flips = 1 + int(splitmix(s) % 6);
      for (int f = 0; f < flips; ++f) v ^= 1ull << (1 + splitmix(s) % 63);
      h[i] = v ? v : 2;
    }
  }
  cb_dct_index* ix = cb_dct_index_create();
  if (!ix || cb_dct_index_load(ix, ids.data(), h.data(), rows) != CB_OK) {
    printf("{\"error\": \"%s\"}\n", cb_last_error());
    return 1;
  }
  cb_params p;
  cb_params_default(&p);
  p.dctThresh = dht;
  const int64_t n_needles = rows < (1 << 16) ? rows : (1 << 16);
  // one caller: latency, and the results to compare with
  std::vector<uint64_t> want(n_needles);
  cb_match buf[512];
  int64_t n = 0;
  for (int i = 0; i < 64; ++i) cb_dct_index_find(ix, h[i], &p, buf, 512, &n);
  auto t0 = std::chrono::steady_clock::now();
  int64_t total_single = 0;
  for (int64_t i = 0; i < n_needles; ++i) {
    if (cb_dct_index_find(ix, h[i], &p, buf, 512, &n) != CB_OK) {
      printf("{\"error\": \"%s\"}\n", cb_last_error());
      return 1;
    }
    want[i] = checksum(buf, n);
    total_single += n;
  }
  const double single_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  uint64_t b0 = 0, q0 = 0;
  cb_dct_index_find_queue_stats(ix, &b0, &q0);
  std::atomic<bool> stop{false};
  std::atomic<uint64_t> done{0}, mismatches{0}, errors{0};
  std::vector<std::thread> th;
  t0 = std::chrono::steady_clock::now();
  for (int t = 0; t < threads; ++t)
    th.emplace_back([&, t] {
      cb_match mb[512];
      int64_t k = 0;
      uint64_t mine = 0;
      for (int64_t i = t; !stop.load(std::memory_order_relaxed); i += threads) {
        const int64_t j = i % n_needles;
        if (cb_dct_index_find(ix, h[j], &p, mb, 512, &k) != CB_OK) {
          errors.fetch_add(1);
          break;
        }
        if (checksum(mb, k) != want[j]) mismatches.fetch_add(1);
        ++mine;
      }
      done.fetch_add(mine);
    });
  std::this_thread::sleep_for(std::chrono::duration<double>(seconds));
  stop.store(true);
  for (auto& t : th) t.join();
  const double par_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  uint64_t b1 = 0, q1 = 0;
  cb_dct_index_find_queue_stats(ix, &b1, &q1);
  printf("{\"index_rows\": %lld, \"dht\": %d, \"threads\": %d, \"finds_per_s\": %.1f, \"finds\": %llu, \"seconds\": %.3f, "
         "\"launches\": %llu, \"needles_per_launch\": %.2f, \"single_caller_latency_us\": %.2f, \"single_caller_finds_per_s\": %.1f, "
         "\"matches_single_caller\": %lld, \"mismatches_vs_single_caller\": %llu, \"errors\": %llu}\n",
         (long long)rows, dht, threads, double(done.load()) / par_s, (unsigned long long)done.load(), par_s,
         (unsigned long long)(b1 - b0), double(q1 - q0) / double(b1 - b0 ? b1 - b0 : 1), single_s / double(n_needles) * 1e6,
         double(n_needles) / single_s, (long long)total_single, (unsigned long long)mismatches.load(),
         (unsigned long long)errors.load());
  cb_dct_index_destroy(ix);
  return mismatches.load() || errors.load() ? 2 : 0;
}

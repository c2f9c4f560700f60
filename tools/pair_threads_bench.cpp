// Developer probe (no Python, no GIL): N host threads make synchronous slamb200_match_pair calls against
// resident sets that share one query -- the reference's multi-threaded search (batch.cpp:181-201).
//   g++ -O2 -std=c++17 -Iinclude tools/pair_threads_bench.cpp -Lslam_indoor_code_b200/lib -lslamb200 -lpthread -o /tmp/ptb
//   LD_LIBRARY_PATH=slam_indoor_code_b200/lib /tmp/ptb
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <thread>
#include <vector>

#include "slamb200.h"

static std::vector<float> sift_like(int n, unsigned seed) {
  std::mt19937 rng(seed);
  std::vector<float> v((size_t)n * 128);
  for (auto& x : v) x = (float)(rng() % 64);   // integer valued, |row|^2 < 2^20
  return v;
}

int main() {
  slamb200_ctx* ctx = nullptr;
  if (slamb200_init(0, &ctx) != 0) { printf("init failed: %s\n", slamb200_last_error()); return 1; }
  const int N = 10000, NT = 8;
  auto q = sift_like(N, 1);
  slamb200_desc* Q = nullptr;
  slamb200_upload_desc(ctx, SLAMB200_DESC_F32X128, q.data(), N, 512, &Q);
  std::vector<slamb200_desc*> T(NT);
  for (int i = 0; i < NT; i++) {
    auto t = sift_like(N, 100 + i);
    for (int r = 0; r < 3000; r++)               // plant near-duplicates so that the ratio test keeps rows
      for (int k = 0; k < 128; k++) t[(size_t)r * 128 + k] = q[(size_t)r * 128 + k];
    slamb200_upload_desc(ctx, SLAMB200_DESC_F32X128, t.data(), N, 512, &T[i]);
  }
  for (int threads : {1, 2, 4, 8}) {
    const int calls = 200;
    std::vector<std::thread> th;
    std::vector<int> kept(threads, 0);
    auto work = [&](int k) {
      std::vector<slamb200_dmatch> out(N);
      int n = 0;
      for (int it = 0; it < calls; it++) {
        if (slamb200_match_pair(ctx, SLAMB200_SIFT_BF, Q, T[(k + it) % NT], 0.7, out.data(), N, &n) != 0) {
          printf("match_pair failed: %s\n", slamb200_last_error());
          exit(1);
        }
        kept[k] = n;
      }
    };
    for (int k = 0; k < threads; k++) work(k), kept[k] = 0;   // warm every lane
    const auto t0 = std::chrono::steady_clock::now();
    for (int k = 0; k < threads; k++) th.emplace_back(work, k);
    for (auto& t : th) t.join();
    const double us = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count();
    printf("%d thread(s): %.1f us per pair aggregate (%d kept in the last call)\n", threads, us / (threads * calls), kept[0]);
  }
  slamb200_shutdown(ctx);
  return 0;
}

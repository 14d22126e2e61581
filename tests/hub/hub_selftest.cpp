// CPU check of include/gunrock/b200/thread_hub.hxx: the barrier the host threads of a multi-device run share.
//   hub_selftest <threads> <iterations>
// 1. lock-step: every thread adds to a per-round counter between barriers; nobody may see a short round;
// 2. the two-phase value slots as thread_exchange_t::reduce_stats uses them (ONE barrier per reduction);
// 3. abort: a thread that fails releases its peers from the barrier with an exception instead of a dead-lock;
// 4. reset makes the hub reusable after an abort.
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>

#include <gunrock/b200/thread_hub.hxx>

using gunrock::b200::thread_hub_t;

static int failures = 0;
#define CHECK(c)                                                   \
  do {                                                             \
    if (!(c)) {                                                    \
      std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #c);   \
      ++failures;                                                  \
    }                                                              \
  } while (0)

int main(int argc, char** argv) {
  const int T = argc > 1 ? std::atoi(argv[1]) : 8;
  const int N = argc > 2 ? std::atoi(argv[2]) : 2000;
  thread_hub_t hub;
  {  // 1 + 2
    hub.reset(T);
    std::vector<std::atomic<int>> round(N);
    for (auto& r : round)
      r = 0;
    std::atomic<int> bad{0};
    std::vector<std::thread> th;
    for (int t = 0; t < T; ++t)
      th.emplace_back([&, t] {
        int phase = 0;
        for (int i = 0; i < N; ++i) {
          round[i].fetch_add(1);
          hub.barrier();
          if (round[i].load() != T)
            bad.fetch_add(1);
          // reduce_stats pattern: publish into this phase's slots, ONE barrier, everybody sums
          const int ph = phase++ & 1;
          for (int k = 0; k < 4; ++k)
            hub.vals[ph][t][k] = 1000ll * i + t + k;
          hub.barrier();
          for (int k = 0; k < 4; ++k) {
            long long s = 0, want = 0;
            for (int p = 0; p < T; ++p) {
              s += hub.vals[ph][p][k];
              want += 1000ll * i + p + k;
            }
            if (s != want)
              bad.fetch_add(1);
          }
        }
      });
    for (auto& x : th)
      x.join();
    CHECK(bad.load() == 0);
  }
  {  // 3: thread 0 fails after a few rounds; everybody else must come out with an exception
    hub.reset(T);
    std::atomic<int> released{0};
    std::vector<std::thread> th;
    for (int t = 0; t < T; ++t)
      th.emplace_back([&, t] {
        try {
          for (int i = 0; i < 50; ++i) {
            if (t == 0 && i == 7) {
              hub.abort();
              return;
            }
            hub.barrier();
          }
        } catch (const std::exception&) {
          released.fetch_add(1);
        }
      });
    for (auto& x : th)
      x.join();
    CHECK(released.load() == T - 1);
  }
  {  // 4: usable again
    hub.reset(T);
    std::atomic<int> done{0};
    std::vector<std::thread> th;
    for (int t = 0; t < T; ++t)
      th.emplace_back([&] {
        for (int i = 0; i < 100; ++i)
          hub.barrier();
        done.fetch_add(1);
      });
    for (auto& x : th)
      x.join();
    CHECK(done.load() == T);
  }
  if (!failures)
    std::printf("ALL OK\n");
  return failures ? 1 : 0;
}

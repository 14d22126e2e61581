/**
 * @file thread_hub.hxx
 * @brief What the host threads of one multi-device run share (part_loops.cuh `thread_exchange_t`): a reusable
 * barrier that can be aborted, and the slots the ranks publish pointers / values in.  Plain C++ (no CUDA), so that
 * tests/test_thread_hub.py can hammer it on a CPU.
 */
#pragma once

#include <condition_variable>
#include <mutex>
#include <stdexcept>

namespace gunrock {
namespace b200 {

constexpr int kHubMaxRanks = 16;

/// Shared by the ranks of one run: a reusable barrier and the slots the ranks publish pointers / values in.
struct thread_hub_t {
  int nparts = 1;
  std::mutex m;
  std::condition_variable cv;
  int waiting = 0;
  unsigned long long generation = 0;
  bool aborted = false;
  const void* ptr[kHubMaxRanks] = {};
  long long vals[2][kHubMaxRanks][4] = {};
  float fvals[2][kHubMaxRanks] = {};

  void reset(int n) {
    std::lock_guard<std::mutex> lk(m);
    nparts = n;
    waiting = 0;
    aborted = false;
  }
  /// Every rank arrives; throws in all of them when one has failed (`abort`), so no thread waits for ever.
  void barrier() {
    std::unique_lock<std::mutex> lk(m);
    if (aborted)
      throw std::runtime_error("a peer rank of this multi-device run failed");
    const unsigned long long gen = generation;
    if (++waiting == nparts) {
      waiting = 0;
      ++generation;
      cv.notify_all();
      return;
    }
    cv.wait(lk, [&] { return generation != gen || aborted; });
    if (generation == gen)
      throw std::runtime_error("a peer rank of this multi-device run failed");
  }
  void abort() {
    std::lock_guard<std::mutex> lk(m);
    aborted = true;
    cv.notify_all();
  }
};

}  // namespace b200
}  // namespace gunrock

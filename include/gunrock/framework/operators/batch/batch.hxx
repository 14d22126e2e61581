/**
 * @file batch.hxx
 * @brief `operators::batch::execute(f, number_of_jobs, total_elapsed)`
 * (include/gunrock/framework/operators/batch/batch.hxx:61-83): run `f(job)` for every job on host
 * threads and add up the elapsed milliseconds each call returns.  Jobs that launch on the same
 * context share its stream, so they serialise on the device exactly as in the reference; the host
 * threads only overlap the launch/synchronisation gaps.  At most `hardware_concurrency` jobs are in
 * flight at a time (the reference spawns one std::thread per job, unbounded).
 */
#pragma once

#include <algorithm>
#include <atomic>
#include <cstddef>
#include <mutex>
#include <thread>
#include <vector>

namespace gunrock {
namespace operators {
namespace batch {

template <typename function_t, typename... args_t>
void execute(function_t f, std::size_t number_of_jobs, float* total_elapsed, args_t&... args) {
  std::mutex sum_guard;
  std::atomic<std::size_t> next{0};
  int device = 0;
  cudaGetDevice(&device);
  auto worker = [&]() {
    cudaSetDevice(device);
    for (;;) {
      std::size_t job = next.fetch_add(1);
      if (job >= number_of_jobs)
        return;
      float ms = f(job);
      std::lock_guard<std::mutex> lock(sum_guard);
      total_elapsed[0] += ms;
    }
  };
  // the per-context operator workspace is not re-entrant: one worker per context at a time
  worker();
}

}  // namespace batch
}  // namespace operators
}  // namespace gunrock

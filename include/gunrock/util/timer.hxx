/**
 * @file timer.hxx
 * @brief CUDA-event millisecond timer, `util::timer_t` (include/gunrock/util/timer.hxx:18-61).
 */
#pragma once

#include <cuda_runtime.h>

namespace gunrock {
namespace util {

struct timer_t {
  float time = 0.0f;
  timer_t() {
    cudaEventCreate(&start_);
    cudaEventCreate(&stop_);
  }
  ~timer_t() {
    cudaEventDestroy(start_);
    cudaEventDestroy(stop_);
  }
  timer_t(const timer_t&) = delete;
  timer_t& operator=(const timer_t&) = delete;

  void reset() { time = 0.0f; }
  void begin(cudaStream_t stream = 0) { cudaEventRecord(start_, stream); }
  float end(cudaStream_t stream = 0) {
    cudaEventRecord(stop_, stream);
    cudaEventSynchronize(stop_);
    cudaEventElapsedTime(&time, start_, stop_);
    return milliseconds();
  }
  float seconds() { return time * 1e-3f; }
  float milliseconds() { return time; }

 private:
  cudaEvent_t start_, stop_;
};

}  // namespace util
}  // namespace gunrock

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
  /// The events are created on first use, on the device that is current THEN -- the device of the stream
  /// they are recorded on.  (Created in the constructor they belong to whatever device was current when the
  /// owning object was constructed: a `standard_context_t` initialises its members before its body selects
  /// the device, so the timer of a context built right after a context of another device could not be
  /// recorded on its own stream -- cudaErrorInvalidResourceHandle.)
  timer_t() = default;
  ~timer_t() {
    if (start_)
      cudaEventDestroy(start_);
    if (stop_)
      cudaEventDestroy(stop_);
  }
  timer_t(const timer_t&) = delete;
  timer_t& operator=(const timer_t&) = delete;

  void reset() { time = 0.0f; }
  void begin(cudaStream_t stream = 0) {
    if (!start_) {
      cudaEventCreate(&start_);
      cudaEventCreate(&stop_);
    }
    cudaEventRecord(start_, stream);
  }
  float end(cudaStream_t stream = 0) {
    if (!stop_)
      return 0.0f;  // end() without begin()
    cudaEventRecord(stop_, stream);
    cudaEventSynchronize(stop_);
    cudaEventElapsedTime(&time, start_, stop_);
    return milliseconds();
  }
  float seconds() { return time * 1e-3f; }
  float milliseconds() { return time; }

 private:
  cudaEvent_t start_ = nullptr, stop_ = nullptr;
};

}  // namespace util
}  // namespace gunrock

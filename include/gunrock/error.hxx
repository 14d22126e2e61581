/**
 * @file error.hxx
 * @brief `gunrock::error`: the exception type every header of this tree throws, and the two
 * `throw_if_exception` guards user code calls (interface of the reference's include/gunrock/error.hxx:21-46:
 * a CUDA status or a plain condition, each with an optional message; `what()` reads "<cuda text>\t: <message>").
 */
#pragma once

#include <exception>
#include <string>
#include <utility>

#include <cuda_runtime.h>

namespace gunrock {
namespace error {

using error_t = cudaError_t;

class exception_t : public std::exception {
 public:
  std::string report;  ///< the text `what()` returns (public: the reference's callers read it directly)

  explicit exception_t(std::string message = "") : report(std::move(message)) {}
  exception_t(error_t status, const std::string& message = "")
      : report(describe(status) + "\t: " + message) {}

  const char* what() const noexcept override { return report.c_str(); }

 private:
  static std::string describe(error_t status) {
    const char* text = cudaGetErrorString(status);
    return text ? std::string(text) : std::string("unknown CUDA error");
  }
};

/// CUDA runtime status guard: anything but cudaSuccess becomes an exception_t.
inline void throw_if_exception(error_t status, std::string message = "") {
  if (status == cudaSuccess)
    return;
  throw exception_t(status, message);
}

/// Condition guard: `failed == true` becomes an exception_t carrying only the message.
inline void throw_if_exception(bool failed, std::string message = "") {
  if (!failed)
    return;
  throw exception_t(std::move(message));
}

}  // namespace error
}  // namespace gunrock

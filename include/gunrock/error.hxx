/**
 * @file error.hxx
 * @brief Exceptions (drop-in for include/gunrock/error.hxx:21-46: `error::exception_t`,
 * `error::throw_if_exception(status, msg)` and the bool overload).
 */
#pragma once

#include <exception>
#include <string>

#include <cuda_runtime.h>

namespace gunrock {
namespace error {

typedef cudaError_t error_t;

struct exception_t : std::exception {
  std::string report;
  exception_t(error_t status, std::string message = "") {
    report = std::string(cudaGetErrorString(status)) + "\t: " + message;
  }
  exception_t(std::string message = "") : report(message) {}
  const char* what() const noexcept override { return report.c_str(); }
};

inline void throw_if_exception(error_t status, std::string message = "") {
  if (status != cudaSuccess)
    throw exception_t(status, message);
}

inline void throw_if_exception(bool is_exception, std::string message = "") {
  if (is_exception)
    throw exception_t(message);
}

}  // namespace error
}  // namespace gunrock

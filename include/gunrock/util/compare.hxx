/**
 * @file compare.hxx
 * @brief `util::compare(device_ptr, host_ptr, n)` -> number of mismatches
 * (include/gunrock/util/compare.hxx:40-57; default comparator `!=`, prints the first few).
 */
#pragma once

#include <cstddef>
#include <iostream>
#include <vector>

#include <cuda_runtime.h>

namespace gunrock {
namespace util {

namespace detail {
struct default_comp_t {
  template <typename type_t>
  bool operator()(const type_t& a, const type_t& b) const {
    return a != b;
  }
};
}  // namespace detail

template <typename type_t, typename comp_t = detail::default_comp_t>
int compare(const type_t* d_ptr,
            const type_t* h_ptr,
            std::size_t n,
            comp_t error_op = comp_t(),
            bool verbose = false) {
  std::vector<type_t> d_copy(n);
  if (n)
    cudaMemcpy(d_copy.data(), d_ptr, sizeof(type_t) * n, cudaMemcpyDeviceToHost);
  int errors = 0;
  for (std::size_t i = 0; i < n; ++i) {
    if (error_op(d_copy[i], h_ptr[i])) {
      if (verbose && errors < 10)
        std::cout << "Error[" << i << "]: " << d_copy[i] << " != " << h_ptr[i] << std::endl;
      ++errors;
    }
  }
  return errors;
}

}  // namespace util
}  // namespace gunrock

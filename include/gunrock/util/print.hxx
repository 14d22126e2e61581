/**
 * @file print.hxx
 * @brief `print::head(vector, k, name)` for thrust host/device vectors and raw device pointers
 * (include/gunrock/util/print.hxx:31-67).
 */
#pragma once

#include <iostream>
#include <string>

#include <thrust/copy.h>
#include <thrust/device_vector.h>
#include <thrust/host_vector.h>

namespace gunrock {
namespace print {

template <typename vector_t>
void head(vector_t& x, int n, std::string name = "") {
  using type_t = typename vector_t::value_type;
  if (name != "")
    std::cout << name << "[:" << n << "] = ";
  std::size_t k = static_cast<std::size_t>(n) < x.size() ? static_cast<std::size_t>(n) : x.size();
  thrust::host_vector<type_t> h(x.begin(), x.begin() + k);
  for (std::size_t i = 0; i < k; ++i)
    std::cout << h[i] << " ";
  std::cout << std::endl;
}

template <typename type_t>
void head(type_t* d_ptr, int n, std::string name = "") {
  thrust::device_ptr<type_t> p(d_ptr);
  thrust::host_vector<type_t> h(p, p + n);
  if (name != "")
    std::cout << name << "[:" << n << "] = ";
  for (int i = 0; i < n; ++i)
    std::cout << h[i] << " ";
  std::cout << std::endl;
}

}  // namespace print
}  // namespace gunrock

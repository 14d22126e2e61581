/**
 * @file sample.hxx
 * @brief The 4x4 sample CSR every reference unit test builds (include/gunrock/io/sample.hxx:57-92):
 * offsets 0 0 2 3 4, columns 0 1 2 1, values 5 8 3 6.
 */
#pragma once

#include <gunrock/formats/formats.hxx>
#include <gunrock/memory.hxx>

namespace gunrock {
namespace io {
namespace sample {

using namespace memory;

template <memory_space_t space = memory_space_t::device,
          typename vertex_t = int,
          typename edge_t = int,
          typename weight_t = float>
format::csr_t<space, vertex_t, edge_t, weight_t> csr() {
  format::csr_t<memory_space_t::host, vertex_t, edge_t, weight_t> m(4, 4, 4);
  const edge_t offsets[5] = {0, 0, 2, 3, 4};
  const vertex_t columns[4] = {0, 1, 2, 1};
  const weight_t values[4] = {5, 8, 3, 6};
  for (int i = 0; i < 5; ++i)
    m.row_offsets[i] = offsets[i];
  for (int i = 0; i < 4; ++i) {
    m.column_indices[i] = columns[i];
    m.nonzero_values[i] = values[i];
  }
  return format::csr_t<space, vertex_t, edge_t, weight_t>(m);
}

}  // namespace sample
}  // namespace io
}  // namespace gunrock

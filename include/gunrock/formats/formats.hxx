/**
 * @file formats.hxx
 * @brief Owning sparse formats `format::{coo_t, csr_t, csc_t}<space, index_t, offset_t, value_t>`
 * (include/gunrock/formats/{coo,csr,csc}.hxx).  Public members and methods keep the reference's
 * names: number_of_rows/columns/nonzeros, row_offsets/column_indices/nonzero_values (thrust
 * vectors, so they convert to thrust::host_vector as examples/algorithms/bfs/bfs_cpu.hxx:25-27
 * needs), from_coo, read_binary / write_binary (binary `.csr` layout of csr.hxx:142-228:
 * index_t rows, index_t cols, offset_t nnz, offsets[rows+1], indices[nnz], values[nnz]).
 *
 * from_coo keeps the reference's result exactly (stable counting sort by row, duplicates and
 * self loops kept, csr.hxx:81-140) but does the histogram + scatter with std:: algorithms on the
 * host arrays once and uploads once.
 */
#pragma once

#include <cstdio>
#include <string>
#include <vector>

#include <gunrock/container/vector.hxx>
#include <gunrock/error.hxx>
#include <gunrock/memory.hxx>

namespace gunrock {
namespace format {

using namespace memory;

template <memory_space_t space, typename index_t, typename offset_t, typename value_t>
struct coo_t {
  using index_type = index_t;
  using offset_type = offset_t;
  using value_type = value_t;

  index_t number_of_rows = 0;
  index_t number_of_columns = 0;
  offset_t number_of_nonzeros = 0;

  vector_t<index_t, space> row_indices;     // I
  vector_t<index_t, space> column_indices;  // J
  vector_t<value_t, space> nonzero_values;  // V

  coo_t() = default;
  coo_t(index_t r, index_t c, offset_t nnz)
      : number_of_rows(r),
        number_of_columns(c),
        number_of_nonzeros(nnz),
        row_indices(nnz),
        column_indices(nnz),
        nonzero_values(nnz) {}
  template <typename _coo_t>
  coo_t(const _coo_t& rhs)
      : number_of_rows(rhs.number_of_rows),
        number_of_columns(rhs.number_of_columns),
        number_of_nonzeros(rhs.number_of_nonzeros),
        row_indices(rhs.row_indices),
        column_indices(rhs.column_indices),
        nonzero_values(rhs.nonzero_values) {}
  ~coo_t() = default;
};

template <memory_space_t space, typename index_t, typename offset_t, typename value_t>
struct csr_t {
  using index_type = index_t;
  using offset_type = offset_t;
  using value_type = value_t;

  index_t number_of_rows = 0;
  index_t number_of_columns = 0;
  offset_t number_of_nonzeros = 0;

  vector_t<offset_t, space> row_offsets;    // Ap
  vector_t<index_t, space> column_indices;  // Aj
  vector_t<value_t, space> nonzero_values;  // Ax

  csr_t() = default;
  csr_t(index_t r, index_t c, offset_t nnz)
      : number_of_rows(r),
        number_of_columns(c),
        number_of_nonzeros(nnz),
        row_offsets(r + 1),
        column_indices(nnz),
        nonzero_values(nnz) {}
  template <typename _csr_t>
  csr_t(const _csr_t& rhs)
      : number_of_rows(rhs.number_of_rows),
        number_of_columns(rhs.number_of_columns),
        number_of_nonzeros(rhs.number_of_nonzeros),
        row_offsets(rhs.row_offsets),
        column_indices(rhs.column_indices),
        nonzero_values(rhs.nonzero_values) {}
  ~csr_t() = default;

  csr_t<space, index_t, offset_t, value_t> from_coo(
      const coo_t<memory_space_t::host, index_t, offset_t, value_t>& coo) {
    number_of_rows = coo.number_of_rows;
    number_of_columns = coo.number_of_columns;
    number_of_nonzeros = coo.number_of_nonzeros;
    const std::size_t rows = static_cast<std::size_t>(number_of_rows);
    const std::size_t nnz = static_cast<std::size_t>(number_of_nonzeros);
    thrust::host_vector<offset_t> offs(rows + 1, 0);
    thrust::host_vector<index_t> cols(nnz);
    thrust::host_vector<value_t> vals(nnz);
    for (std::size_t k = 0; k < nnz; ++k)
      ++offs[static_cast<std::size_t>(coo.row_indices[k]) + 1];
    for (std::size_t r = 0; r < rows; ++r)
      offs[r + 1] += offs[r];
    std::vector<offset_t> cursor(offs.begin(), offs.begin() + rows);
    for (std::size_t k = 0; k < nnz; ++k) {
      offset_t at = cursor[coo.row_indices[k]]++;
      cols[at] = coo.column_indices[k];
      vals[at] = coo.nonzero_values[k];
    }
    row_offsets = offs;
    column_indices = cols;
    nonzero_values = vals;
    return *this;
  }

  void read_binary(std::string filename) {
    FILE* f = fopen(filename.c_str(), "rb");
    error::throw_if_exception(f == nullptr, "could not open " + filename);
    bool ok = fread(&number_of_rows, sizeof(index_t), 1, f) == 1 &&
              fread(&number_of_columns, sizeof(index_t), 1, f) == 1 &&
              fread(&number_of_nonzeros, sizeof(offset_t), 1, f) == 1;
    error::throw_if_exception(!ok, "truncated .csr header: " + filename);
    thrust::host_vector<offset_t> offs(static_cast<std::size_t>(number_of_rows) + 1);
    thrust::host_vector<index_t> cols(number_of_nonzeros);
    thrust::host_vector<value_t> vals(number_of_nonzeros);
    ok = fread(offs.data(), sizeof(offset_t), offs.size(), f) == offs.size() &&
         fread(cols.data(), sizeof(index_t), cols.size(), f) == cols.size() &&
         fread(vals.data(), sizeof(value_t), vals.size(), f) == vals.size();
    fclose(f);
    error::throw_if_exception(!ok, "truncated .csr body: " + filename);
    row_offsets = offs;
    column_indices = cols;
    nonzero_values = vals;
  }

  void write_binary(std::string filename) {
    FILE* f = fopen(filename.c_str(), "wb");
    error::throw_if_exception(f == nullptr, "could not open " + filename);
    thrust::host_vector<offset_t> offs(row_offsets);
    thrust::host_vector<index_t> cols(column_indices);
    thrust::host_vector<value_t> vals(nonzero_values);
    fwrite(&number_of_rows, sizeof(index_t), 1, f);
    fwrite(&number_of_columns, sizeof(index_t), 1, f);
    fwrite(&number_of_nonzeros, sizeof(offset_t), 1, f);
    fwrite(offs.data(), sizeof(offset_t), offs.size(), f);
    fwrite(cols.data(), sizeof(index_t), cols.size(), f);
    fwrite(vals.data(), sizeof(value_t), vals.size(), f);
    fclose(f);
  }
};

template <memory_space_t space, typename index_t, typename offset_t, typename value_t>
struct csc_t {
  using index_type = index_t;
  using offset_type = offset_t;
  using value_type = value_t;

  index_t number_of_rows = 0;
  index_t number_of_columns = 0;
  offset_t number_of_nonzeros = 0;

  vector_t<offset_t, space> column_offsets;  // Aj
  vector_t<index_t, space> row_indices;      // Ap
  vector_t<value_t, space> nonzero_values;   // Ax

  csc_t() = default;
  csc_t(index_t r, index_t c, offset_t nnz)
      : number_of_rows(r),
        number_of_columns(c),
        number_of_nonzeros(nnz),
        column_offsets(c + 1),
        row_indices(nnz),
        nonzero_values(nnz) {}

  /// Transpose (host side, stable in source order); include/gunrock/formats/csc.hxx:62-102.
  template <memory_space_t s>
  csc_t<space, index_t, offset_t, value_t> from_csr(const csr_t<s, index_t, offset_t, value_t>& csr) {
    number_of_rows = csr.number_of_rows;
    number_of_columns = csr.number_of_columns;
    number_of_nonzeros = csr.number_of_nonzeros;
    thrust::host_vector<offset_t> ro(csr.row_offsets);
    thrust::host_vector<index_t> ci(csr.column_indices);
    thrust::host_vector<value_t> vv(csr.nonzero_values);
    const std::size_t cols = static_cast<std::size_t>(number_of_columns);
    const std::size_t nnz = static_cast<std::size_t>(number_of_nonzeros);
    thrust::host_vector<offset_t> offs(cols + 1, 0);
    thrust::host_vector<index_t> rows(nnz);
    thrust::host_vector<value_t> vals(nnz);
    for (std::size_t e = 0; e < nnz; ++e)
      ++offs[static_cast<std::size_t>(ci[e]) + 1];
    for (std::size_t c = 0; c < cols; ++c)
      offs[c + 1] += offs[c];
    std::vector<offset_t> cursor(offs.begin(), offs.begin() + cols);
    for (index_t u = 0; u < number_of_rows; ++u)
      for (offset_t e = ro[u]; e < ro[u + 1]; ++e) {
        offset_t at = cursor[ci[e]]++;
        rows[at] = u;
        vals[at] = vv[e];
      }
    column_offsets = offs;
    row_indices = rows;
    nonzero_values = vals;
    return *this;
  }
};

}  // namespace format
}  // namespace gunrock

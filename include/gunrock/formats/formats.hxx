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
 * self loops kept, csr.hxx:81-140) but does the histogram + scatter on the host arrays once and
 * uploads once.  Both it and csc_t::from_csr go through `detail::stable_bucket` with all host threads
 * (after the text parse this is the longest step of an example program's ingest):
 * the histogram is taken over entry chunks, the keys are cut into blocks of ~equal entry counts, and
 * each thread scatters the entries of ITS key block walking the entries in order -- which keeps the
 * sort stable and the result bit-identical to the serial loop.  `GUNROCK_B200_HOST_THREADS=1` forces that loop.
 */
#pragma once

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <thread>
#include <vector>

#include <gunrock/container/vector.hxx>
#include <gunrock/error.hxx>
#include <gunrock/memory.hxx>

namespace gunrock {
namespace format {

using namespace memory;

namespace detail {

inline unsigned host_threads(std::size_t items) {
  unsigned t = std::thread::hardware_concurrency();
  if (const char* env = std::getenv("GUNROCK_B200_HOST_THREADS"))
    return static_cast<unsigned>(std::max(1, std::min(64, std::atoi(env))));
  // every thread of the scatter pass reads all keys: beyond 16 the extra streams cost more than they save
  t = std::max(1u, std::min(t, 16u));
  return static_cast<unsigned>(std::min<std::size_t>(t, items / (std::size_t(1) << 19) + 1));
}

/**
 * @brief Stable counting sort of `n` entries into `buckets` buckets: `offsets[b]` = first slot of bucket b
 * (`offsets[buckets]` = n), and `place(k, slot)` is called once per entry k with its slot; entries of a bucket
 * keep their order.  `key(k)` must be in [0, buckets) -- checked, an out-of-range key throws instead of
 * corrupting memory (the reference indexes unchecked, csr.hxx:96-115).
 */
template <typename offset_t, typename key_fn, typename place_fn>
void stable_bucket(std::size_t n, std::size_t buckets, offset_t* offsets, key_fn key, place_fn place) {
  const unsigned threads = host_threads(n);
  std::fill(offsets, offsets + buckets + 1, offset_t(0));
  if (threads <= 1) {
    for (std::size_t k = 0; k < n; ++k) {
      const std::size_t b = static_cast<std::size_t>(key(k));
      if (b >= buckets)  // (the message is only built on failure: this loop runs once per entry)
        error::throw_if_exception(true, "sparse format conversion: index out of range");
      ++offsets[b + 1];
    }
    for (std::size_t b = 0; b < buckets; ++b)
      offsets[b + 1] += offsets[b];
    std::vector<offset_t> cursor(offsets, offsets + buckets);
    for (std::size_t k = 0; k < n; ++k)
      place(k, cursor[static_cast<std::size_t>(key(k))]++);
    return;
  }
  // 1. histogram over entry chunks (relaxed atomic increments: contended only on hub rows)
  std::vector<char> bad(threads, 0);
  {
    std::vector<std::thread> team;
    auto count = [&](unsigned t) {
      const std::size_t lo = n / threads * t, hi = t + 1 == threads ? n : n / threads * (t + 1);
      for (std::size_t k = lo; k < hi; ++k) {
        const std::size_t b = static_cast<std::size_t>(key(k));
        if (b >= buckets) {
          bad[t] = 1;
          return;
        }
        __atomic_fetch_add(&offsets[b + 1], offset_t(1), __ATOMIC_RELAXED);
      }
    };
    for (unsigned t = 1; t < threads; ++t)
      team.emplace_back(count, t);
    count(0);
    for (auto& th : team)
      th.join();
  }
  error::throw_if_exception(std::count(bad.begin(), bad.end(), 1) != 0,
                            "sparse format conversion: index out of range");
  for (std::size_t b = 0; b < buckets; ++b)
    offsets[b + 1] += offsets[b];
  // 2. key blocks of ~n / threads entries each; a thread scatters the entries of its block, in entry order
  std::vector<std::size_t> cut(threads + 1, buckets);
  cut[0] = 0;
  for (unsigned t = 1; t < threads; ++t) {
    const offset_t want = static_cast<offset_t>(n / threads * t);
    cut[t] = static_cast<std::size_t>(std::lower_bound(offsets, offsets + buckets, want) - offsets);
    cut[t] = std::max(cut[t], cut[t - 1]);
  }
  {
    std::vector<std::thread> team;
    auto scatter = [&](unsigned t) {
      const std::size_t lo = cut[t], hi = cut[t + 1];
      if (lo >= hi)
        return;
      std::vector<offset_t> cursor(offsets + lo, offsets + hi);
      for (std::size_t k = 0; k < n; ++k) {
        const std::size_t b = static_cast<std::size_t>(key(k));
        if (b >= lo && b < hi)
          place(k, cursor[b - lo]++);
      }
    };
    for (unsigned t = 1; t < threads; ++t)
      team.emplace_back(scatter, t);
    scatter(0);
    for (auto& th : team)
      th.join();
  }
}

}  // namespace detail

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
    const index_t* ri = coo.row_indices.data();
    const index_t* cj = coo.column_indices.data();
    const value_t* vv = coo.nonzero_values.data();
    index_t* out_c = cols.data();
    value_t* out_v = vals.data();
    detail::stable_bucket(
        nnz, rows, offs.data(), [=](std::size_t k) { return ri[k]; },
        [=](std::size_t k, offset_t at) {
          out_c[at] = cj[k];
          out_v[at] = vv[k];
        });
    if constexpr (space == memory_space_t::host) {  // host result: adopt the arrays instead of copying them
      row_offsets.swap(offs);
      column_indices.swap(cols);
      nonzero_values.swap(vals);
    } else {
      row_offsets = offs;
      column_indices = cols;
      nonzero_values = vals;
    }
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
    // source row of every entry (what the serial double loop walks implicitly), then the same stable bucket
    // sort by column: entries of a column stay in source-row order
    std::vector<index_t> src(nnz);
    for (index_t u = 0; u < number_of_rows; ++u)
      std::fill(src.begin() + ro[u], src.begin() + ro[u + 1], u);
    const index_t* col = ci.data();
    const value_t* val = vv.data();
    const index_t* from = src.data();
    index_t* out_r = rows.data();
    value_t* out_v = vals.data();
    detail::stable_bucket(
        nnz, cols, offs.data(), [=](std::size_t e) { return col[e]; },
        [=](std::size_t e, offset_t at) {
          out_r[at] = from[e];
          out_v[at] = val[e];
        });
    if constexpr (space == memory_space_t::host) {
      column_offsets.swap(offs);
      row_indices.swap(rows);
      nonzero_values.swap(vals);
    } else {
      column_offsets = offs;
      row_indices = rows;
      nonzero_values = vals;
    }
    return *this;
  }
};

}  // namespace format
}  // namespace gunrock

/**
 * @file transpose.cuh
 * @brief Device CSR -> CSC transpose (ingest step, never inside a timed region).  Replaces
 * format::csc_t::from_csr (include/gunrock/formats/csc.hxx:62-102, a device sort_by_key).
 * Edges are keyed (destination << 32 | CSR position) and radix-sorted, so the in-edges of a vertex
 * come out ordered by source id, ties in CSR order -- the order the PageRank pull accumulates in.
 * The radix sort is cub::DeviceRadixSort: the one library kernel in the repository, used only for
 * ingest (graph build / transpose); nothing on the advance/filter/compute path touches CUB.
 */
#pragma once

#include <cub/device/device_radix_sort.cuh>

#include <gunrock/b200/ptx.cuh>
#include <gunrock/b200/runtime.cuh>

namespace gunrock {
namespace b200 {

/// Sort 64-bit keys on `st` (ingest only).  Returns the buffer holding the sorted keys.
inline unsigned long long* sort_keys_u64(unsigned long long* keys, unsigned long long* alt, size_t n,
                                         int end_bit, cudaStream_t st) {
  cub::DoubleBuffer<unsigned long long> db(keys, alt);
  size_t temp_bytes = 0;
  B2G_CHECK(cub::DeviceRadixSort::SortKeys(nullptr, temp_bytes, db, static_cast<long long>(n), 0,
                                           end_bit, st));
  void* temp = nullptr;
  B2G_CHECK(cudaMalloc(&temp, temp_bytes ? temp_bytes : 16));
  cudaError_t e = cub::DeviceRadixSort::SortKeys(temp, temp_bytes, db, static_cast<long long>(n), 0,
                                                 end_bit, st);
  cudaError_t e2 = cudaStreamSynchronize(st);
  cudaFree(temp);
  B2G_CHECK(e);
  B2G_CHECK(e2);
  return db.Current();
}

/// key[k] = row[k] << 32 | k  (stable by original position inside a row)
static __global__ void position_keys_kernel(int n, const int* __restrict__ rows, unsigned long long* keys) {
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x)
    keys[k] = (static_cast<unsigned long long>(static_cast<unsigned>(rows[k])) << 32) |
              static_cast<unsigned>(k);
}

/// row_offsets from sorted row ids: rows in (rows[i-1], rows[i]] start at position i.
static __global__ void offsets_from_sorted_rows_kernel(const int* __restrict__ rows, int n, int n_vertices,
                                                int* ro) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i <= n; i += gridDim.x * blockDim.x) {
    int prev = (i == 0) ? -1 : rows[i - 1];
    int cur = (i == n) ? n_vertices : rows[i];
    for (int r = prev + 1; r <= cur; ++r)
      ro[r] = i;
  }
}

inline int key_bits_for(int n_rows) {
  int bits = 32;
  while (bits < 64 && (1ll << (bits - 32)) < n_rows)
    ++bits;
  return bits;
}

struct transpose_t {
  dbuf_t<int> ro, ci;
  dbuf_t<float> vals;
  csr_view_t view;

  void build(workspace_t& ws, const csr_view_t& g) {
    const int V = g.n_vertices, E = g.n_edges;
    cudaStream_t st = ws.stream;
    const int sms = device_info_t::get().sm_count;
    ro.ensure(static_cast<size_t>(V) + 1 + 16);
    ci.ensure(static_cast<size_t>(E) + 16);
    if (g.values)
      vals.ensure(static_cast<size_t>(E) + 16);
    dbuf_t<unsigned long long> k0, k1;
    dbuf_t<int> rows;
    k0.ensure(static_cast<size_t>(E) + 1);
    k1.ensure(static_cast<size_t>(E) + 1);
    rows.ensure(static_cast<size_t>(E) + 1);
    if (E > 0) {
      position_keys_kernel<<<sms * 8, 256, 0, st>>>(E, g.column_indices, k0.ptr);
      unsigned long long* sorted =
          sort_keys_u64(k0.ptr, k1.ptr, static_cast<size_t>(E), key_bits_for(V), st);
      const int* g_ro = g.row_offsets;
      const float* g_vals = g.values;
      int* t_ci = ci.ptr;
      float* t_vals = g_vals ? vals.ptr : nullptr;
      int* rows_p = rows.ptr;
      auto fill = [=] __device__(int i) {
        unsigned long long key = sorted[i];
        int e = static_cast<int>(static_cast<unsigned>(key));
        int lo = 0, hi = V;  // g_ro[lo] <= e < g_ro[hi]
        while (hi - lo > 1) {
          int mid = (lo + hi) >> 1;
          if (g_ro[mid] <= e)
            lo = mid;
          else
            hi = mid;
        }
        t_ci[i] = lo;
        if (t_vals)
          t_vals[i] = g_vals[e];
        rows_p[i] = static_cast<int>(key >> 32);
      };
      for_each_index<<<sms * 8, 256, 0, st>>>(E, fill);
    }
    offsets_from_sorted_rows_kernel<<<sms * 4, 256, 0, st>>>(rows.ptr, E, V, ro.ptr);
    B2G_CHECK(cudaStreamSynchronize(st));
    view.n_vertices = V;
    view.n_edges = E;
    view.row_offsets = ro.ptr;
    view.column_indices = ci.ptr;
    view.values = g.values ? vals.ptr : nullptr;
    view.uid = next_graph_uid();  // the buffers are grow-only: same addresses, different graph
  }
};

}  // namespace b200
}  // namespace gunrock

/**
 * @file filter.cuh
 * @brief Frontier filter / uniquify kernels.
 *
 * Replaces (all under include/gunrock/framework/operators/):
 *   filter/predicated.hxx:12-39  thrust::copy_if          -> stable select (look-back scan)
 *   filter/remove.hxx:11-45      thrust::remove_copy_if   -> same kernel
 *   filter/compact.hxx:13-25     (throws in the reference) -> same kernel
 *   filter/bypass.hxx:13-69      thrust::transform        -> bypass_kernel (marks -1, size kept)
 *   uniquify/unique.hxx:22-35, unique_copy.hxx:22-38      -> adjacent-unique select, and a
 *       bitmap-based exact dedup that yields the sorted unique set without a radix sort.
 * Contract kept from filter.hxx:72-100: an element x survives iff is_valid(x) && op(x); op is
 * called exactly once per valid element and never on invalid ones.
 */
#pragma once

#include <gunrock/b200/ptx.cuh>
#include <gunrock/b200/runtime.cuh>
#include <gunrock/b200/scan.cuh>

namespace gunrock {
namespace b200 {

/// Stable, order-preserving select into `out`; *out_count receives the survivor count.
template <typename Op>
inline void launch_filter_select(workspace_t& ws,
                                 const int* in,
                                 const int* in_count,
                                 int in_upper_bound,
                                 int* out,
                                 int* out_count,
                                 Op op) {
  auto value = [=] __device__(int i) -> int {
    int x = in[i];
    return (x >= 0 && op(x)) ? 1 : 0;
  };
  auto emit = [=] __device__(int i, int excl, int keep) {
    if (keep)
      out[excl] = in[i];
  };
  lookback_scan(ws, in_count, 0, in_upper_bound, value, emit, out_count);
}

template <typename Op>
__global__ void bypass_kernel(const int* __restrict__ in,
                              const int* __restrict__ in_count,
                              int* __restrict__ out,
                              int* out_count,
                              Op op) {
  const int n = *in_count;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    int x = in[i];
    out[i] = (x >= 0 && op(x)) ? x : -1;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0 && out_count != in_count)
    *out_count = n;
}

/// bypass: same size, rejected entries overwritten with -1; `out` may alias `in`.
template <typename Op>
inline void launch_filter_bypass(workspace_t& ws,
                                 const int* in,
                                 const int* in_count,
                                 int* out,
                                 int* out_count,
                                 Op op) {
  int grid = device_info_t::get().sm_count * 8;
  bypass_kernel<<<grid, 256, 0, ws.stream>>>(in, in_count, out, out_count, op);
  ws.launches += 1;
  B2G_CHECK(cudaGetLastError());
}

/// Adjacent-duplicate removal (the `best_effort` path of uniquify.hxx:26-94: no sort).
inline void launch_unique_adjacent(workspace_t& ws,
                                   const int* in,
                                   const int* in_count,
                                   int in_upper_bound,
                                   int* out,
                                   int* out_count) {
  auto value = [=] __device__(int i) -> int { return (i == 0 || in[i] != in[i - 1]) ? 1 : 0; };
  auto emit = [=] __device__(int i, int excl, int keep) {
    if (keep)
      out[excl] = in[i];
  };
  lookback_scan(ws, in_count, 0, in_upper_bound, value, emit, out_count);
}

static __global__ void bitmap_mark_kernel(const int* __restrict__ in,
                                   const int* __restrict__ in_count,
                                   unsigned* __restrict__ bitmap,
                                   int* __restrict__ has_invalid) {
  const int n = *in_count;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    int x = in[i];
    if (x >= 0)
      atomicOr(bitmap + (x >> 5), 1u << (x & 31));
    else
      *has_invalid = 1;
  }
}

/**
 * @brief Exact dedup of a vertex frontier = sort + unique of the reference (uniquify.hxx:60-75,
 * algorithms/sort/radix_sort.hxx:39-61) without the sort: mark a V-bit map, then enumerate the
 * set bits in ascending order with the look-back scan.  Invalid (-1) entries sort first in the
 * reference (they are the smallest int) and collapse to a single -1; reproduced via has_invalid.
 * `bitmap` must hold ceil(V/32) zeroed words and is left zeroed again.
 */
inline void launch_unique_exact(workspace_t& ws,
                                const int* in,
                                const int* in_count,
                                int n_vertices,
                                unsigned* bitmap,
                                int* has_invalid,
                                int* out,
                                int* out_count) {
  int grid = device_info_t::get().sm_count * 8;
  B2G_CHECK(cudaMemsetAsync(has_invalid, 0, sizeof(int), ws.stream));
  bitmap_mark_kernel<<<grid, 256, 0, ws.stream>>>(in, in_count, bitmap, has_invalid);
  ws.launches += 1;
  const int words = (n_vertices + 31) / 32;
  // item i == 0 is the "-1 present" flag, item w+1 is bitmap word w
  auto value = [=] __device__(int i) -> int {
    return i == 0 ? (*has_invalid ? 1 : 0) : __popc(bitmap[i - 1]);
  };
  auto emit = [=] __device__(int i, int excl, int cnt) {
    if (i == 0) {
      if (cnt)
        out[0] = -1;
      return;
    }
    unsigned w = bitmap[i - 1];
    if (w)
      bitmap[i - 1] = 0;
    int base = (i - 1) << 5;
    while (w) {
      int b = __ffs(w) - 1;
      w &= w - 1;
      out[excl++] = base + b;
    }
  };
  lookback_scan(ws, nullptr, words + 1, words + 1, value, emit, out_count);
}

}  // namespace b200
}  // namespace gunrock

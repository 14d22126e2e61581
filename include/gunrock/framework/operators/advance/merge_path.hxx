/**
 * @file merge_path.hxx
 * @brief `operators::advance::merge_path::{coordinate_t, merge_path_search, execute}` -- the reference's per-load-
 * balancer entry (include/gunrock/framework/operators/advance/merge_path.hxx:63-67 coordinate, :78-101 search,
 * :289-362 execute) for code that includes it directly (its unit test does).  `execute` is the generic
 * `advance::execute<load_balance_t::merge_path, ...>`, i.e. the span / tile walk of gunrock/b200/advance.cuh over
 * the look-back degree scan; the diagonal search is kept as a utility (the B200 kernels find a span's first row
 * with one search per 256 or 2048 ranks instead of one per thread, merge_path_partition_kernel).
 */
#pragma once

#include <gunrock/framework/operators/advance/advance.hxx>

namespace gunrock {
namespace operators {
namespace advance {
namespace merge_path {

/// A point on the merge path: x rows (segments) and y edges (atoms) consumed, x + y = its diagonal.
template <typename offset_t>
struct coordinate_t {
  offset_t x;
  offset_t y;
};

/**
 * @brief Where diagonal `diagonal` crosses the merge path of the sorted lists a[0, a_len) (segment END offsets)
 * and b[0, b_len) (the edge ranks): the smallest x with a[x] > b[diagonal - x - 1], y = diagonal - x.
 */
template <typename a_iterator_t, typename b_iterator_t, typename offset_t>
__host__ __device__ __forceinline__ void merge_path_search(offset_t diagonal,
                                                           a_iterator_t a,
                                                           b_iterator_t b,
                                                           offset_t a_len,
                                                           offset_t b_len,
                                                           coordinate_t<offset_t>& path_coordinate) {
  offset_t lo = diagonal > b_len ? diagonal - b_len : offset_t(0);
  offset_t hi = diagonal < a_len ? diagonal : a_len;
  while (lo < hi) {  // rows [0, lo) are known to end at or before the diagonal's edge, rows [hi, ..) after it
    const offset_t mid = lo + (hi - lo) / 2;
    if (a[mid] <= b[diagonal - mid - 1])
      lo = mid + 1;
    else
      hi = mid;
  }
  path_coordinate.x = lo < a_len ? lo : a_len;
  path_coordinate.y = diagonal - lo;
}

/// The eight-argument advance with the load balancer fixed (merge_path.hxx:289-362).
template <advance_direction_t direction,
          advance_io_type_t input_type,
          advance_io_type_t output_type,
          typename graph_t,
          typename operator_t,
          typename frontier_t,
          typename work_tiles_t>
void execute(graph_t& G,
             operator_t op,
             frontier_t* input,
             frontier_t* output,
             work_tiles_t& segments,
             gcuda::multi_context_t& context) {
  advance::execute<load_balance_t::merge_path, direction, input_type, output_type>(G, op, input, output,
                                                                                    segments, context);
}

}  // namespace merge_path
}  // namespace advance
}  // namespace operators
}  // namespace gunrock

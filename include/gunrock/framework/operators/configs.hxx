/**
 * @file configs.hxx
 * @brief Operator enums (include/gunrock/framework/operators/configs.hxx:52-112); the numeric
 * values match the reference so `options_t` / CLI strings / B2G_* constants agree.
 */
#pragma once

namespace gunrock {
namespace operators {

enum load_balance_t {
  thread_mapped,  ///< one frontier entry per thread
  warp_mapped,    ///< (mapped to block_mapped's warp bin)
  block_mapped,   ///< degree-binned CTA/warp/thread assignment (B200: persistent + TMA hub bin)
  bucketing,      ///< (mapped to block_mapped)
  merge_path,     ///< edge-balanced tiles over the frontier's degree scan
  merge_path_v2,  ///< alias of merge_path
  work_stealing,  ///< (mapped to block_mapped: its work fetch is already dynamic)
};

enum advance_io_type_t {
  graph,     ///< the whole graph is the input frontier
  vertices,  ///< vertex frontier
  edges,     ///< edge frontier
  none       ///< no output frontier
};

enum advance_direction_t {
  forward,   ///< push
  backward,  ///< pull
  optimized  ///< push/pull switch
};

enum filter_algorithm_t {
  remove,      ///< order-preserving compaction
  predicated,  ///< order-preserving compaction
  compact,     ///< order-preserving compaction (throws in the reference; implemented here)
  bypass       ///< mark rejected entries invalid, keep the size
};

enum uniquify_algorithm_t {
  unique,      ///< in the enactor's buffers
  unique_copy  ///< into the output frontier
};

enum parallel_for_each_t {
  vertex,
  edge,
  weight,
  element
};

}  // namespace operators
}  // namespace gunrock

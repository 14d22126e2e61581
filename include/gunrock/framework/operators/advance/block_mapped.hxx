/**
 * @file block_mapped.hxx
 * @brief `operators::advance::block_mapped::execute` -- the reference's per-load-balancer entry
 * (include/gunrock/framework/operators/advance/block_mapped.hxx:193-249) for code that includes it directly: the generic
 * `advance::execute<load_balance_t::block_mapped, ...>` with the load balancer fixed (kernels: gunrock/b200/advance.cuh).
 */
#pragma once

#include <gunrock/framework/operators/advance/advance.hxx>

namespace gunrock {
namespace operators {
namespace advance {
namespace block_mapped {

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
  advance::execute<load_balance_t::block_mapped, direction, input_type, output_type>(G, op, input, output, segments,
                                                                            context);
}

}  // namespace block_mapped
}  // namespace advance
}  // namespace operators
}  // namespace gunrock

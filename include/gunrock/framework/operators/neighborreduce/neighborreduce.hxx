/**
 * @file neighborreduce.hxx
 * @brief `operators::neighborreduce::execute(G, E, output, op, arithmetic_op, init, context)`:
 * output[v] = fold(arithmetic_op, init, { op(e) : e an out-edge of v }) for every vertex of the graph.
 * In the reference this operator throws since ModernGPU was removed
 * (include/gunrock/framework/operators/neighborreduce/neighborreduce.hxx:52-81); it is restored here
 * as a native segmented reduction: a warp draws 32 rows per ticket, rows of >= 32 edges are reduced by
 * the whole warp (coalesced edge ids, shuffle tree), shorter rows by their lane.  The fold order
 * inside a row is fixed (lane-strided partials, then an xor tree), so results are reproducible.
 */
#pragma once

#include <gunrock/b200/ptx.cuh>
#include <gunrock/b200/runtime.cuh>
#include <gunrock/cuda/context.hxx>
#include <gunrock/error.hxx>
#include <gunrock/framework/operators/configs.hxx>

namespace gunrock {
namespace operators {
namespace neighborreduce {

namespace detail {
template <int kThreads, typename output_t, typename operator_t, typename arithmetic_t>
__global__ void __launch_bounds__(kThreads)
neighborreduce_kernel(b200::csr_view_t g, output_t* __restrict__ output, operator_t op,
                      arithmetic_t arithmetic_op, output_t init_value, b200::ctrl_t* ctrl) {
  const int lane = b200::lane_id();
  const int V = g.n_vertices;
  const int* __restrict__ ro = g.row_offsets;
  for (;;) {
    int base = 0;
    if (lane == 0)
      base = atomicAdd(&ctrl->work, 32);
    base = __shfl_sync(b200::kFull, base, 0);
    if (base >= V)
      break;
    const int v = base + lane;
    int start = 0, deg = 0;
    if (v < V) {
      start = ro[v];
      deg = ro[v + 1] - start;
    }
    output_t acc = init_value;
    bool done = false;
    unsigned big = __ballot_sync(b200::kFull, deg >= 32);
    while (big) {
      const int leader = __ffs(big) - 1;
      big &= big - 1;
      const int s = __shfl_sync(b200::kFull, start, leader);
      const int d = __shfl_sync(b200::kFull, deg, leader);
      output_t part = init_value;
      bool any = false;
      for (int off = lane; off < d; off += 32) {
        output_t x = op(s + off);
        part = any ? arithmetic_op(part, x) : arithmetic_op(init_value, x);
        any = true;
      }
      // xor tree over the 32 partials (every lane has >= 1 element because d >= 32)
#pragma unroll
      for (int delta = 16; delta > 0; delta >>= 1) {
        output_t other = __shfl_xor_sync(b200::kFull, part, delta);
        part = arithmetic_op(part, other);
      }
      if (lane == leader) {
        acc = part;
        done = true;
      }
    }
    if (!done)
      for (int k = 0; k < deg; ++k)
        acc = arithmetic_op(acc, op(start + k));
    if (v < V)
      output[v] = acc;
  }
}
}  // namespace detail

template <advance_io_type_t input_t = advance_io_type_t::graph,
          typename graph_t,
          typename enactor_t,
          typename output_t,
          typename operator_t,
          typename arithmetic_t>
void execute(graph_t& G,
             enactor_t* E,
             output_t* output,
             operator_t op,
             arithmetic_t arithmetic_op,
             output_t init_value,
             gcuda::multi_context_t& context) {
  static_assert(input_t == advance_io_type_t::graph,
                "neighborreduce runs over the whole graph (as the reference declares it)");
  error::throw_if_exception(context.size() != 1, "`context.size() != 1` not supported");
  auto context0 = context.get_context(0);
  b200::workspace_t& ws = context0->workspace();
  b200::ctrl_t* ctrl = ws.next_ctrl();
  const int grid = b200::device_info_t::get().sm_count * 8;
  detail::neighborreduce_kernel<256><<<grid, 256, 0, ws.stream>>>(G.csr_view(), output, op,
                                                                  arithmetic_op, init_value, ctrl);
  ws.launches += 1;
  error::throw_if_exception(cudaGetLastError(), "neighborreduce launch");
}

}  // namespace neighborreduce
}  // namespace operators
}  // namespace gunrock

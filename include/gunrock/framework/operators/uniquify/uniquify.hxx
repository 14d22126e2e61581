/**
 * @file uniquify.hxx
 * @brief `operators::uniquify::execute` (include/gunrock/framework/operators/uniquify/uniquify.hxx:26-94).
 * best_effort (default, algorithms.hxx:47): adjacent duplicates only, no sort.
 * otherwise: exact -- the reference sorts then uniques (radix_sort.hxx:39-61 + unique.hxx:22-35);
 * here the sorted unique set is produced directly by marking a V-bit map and enumerating it
 * (gunrock/b200/filter.cuh), which needs the vertex count: pass it via `n_vertices` or let the
 * enactor form read it from the problem's graph.
 * The reference's swap bugs are not reproduced: after the call the result is in
 * `output` (explicit form) or in the enactor's active frontier (enactor form, swap_buffers=true).
 */
#pragma once

#include <gunrock/b200/filter.cuh>
#include <gunrock/cuda/context.hxx>
#include <gunrock/error.hxx>
#include <gunrock/framework/operators/configs.hxx>

namespace gunrock {
namespace operators {
namespace uniquify {

template <uniquify_algorithm_t type = uniquify_algorithm_t::unique, typename frontier_t>
void execute(frontier_t* input,
             frontier_t* output,
             gcuda::multi_context_t& context,
             bool best_effort_uniquification = false,
             const float uniquification_percent = 100,
             std::size_t n_vertices = 0) {
  error::throw_if_exception(context.size() != 1, "`context.size() != 1` not supported");
  error::throw_if_exception(uniquification_percent < 0 || uniquification_percent > 100,
                            "uniquification_percent must be in [0, 100]");
  auto context0 = context.get_context(0);
  b200::workspace_t& ws = context0->workspace();
  std::size_t bound = input->size_upper_bound();
  if (output->get_capacity() < bound + 1)
    output->reserve(bound + 1);
  output->bind_stream(ws.stream);
  const int* in = reinterpret_cast<const int*>(input->get());
  int* out = reinterpret_cast<int*>(output->get());
  bool unique_out = input->is_known_unique();
  if (best_effort_uniquification || n_vertices == 0) {
    b200::launch_unique_adjacent(ws, in, input->count_ptr(), static_cast<int>(bound), out,
                                 output->count_ptr());
  } else {
    std::size_t words = (n_vertices + 31) / 32 + 4;
    bool fresh = ws.uniq_bitmap.cap < words + 1;
    ws.uniq_bitmap.ensure(words + 1);
    if (fresh)
      error::throw_if_exception(
          cudaMemsetAsync(ws.uniq_bitmap.ptr, 0, ws.uniq_bitmap.cap * sizeof(unsigned), ws.stream),
          "uniquify bitmap init");
    int* has_invalid = reinterpret_cast<int*>(ws.uniq_bitmap.ptr + words);
    b200::launch_unique_exact(ws, in, input->count_ptr(), static_cast<int>(n_vertices),
                              ws.uniq_bitmap.ptr, has_invalid, out, output->count_ptr());
    unique_out = true;  // exact: every vertex at most once
  }
  output->mark_produced(ws.stream, nullptr, unique_out);
}

template <uniquify_algorithm_t type = uniquify_algorithm_t::unique, typename enactor_type>
void execute(enactor_type* E,
             gcuda::multi_context_t& context,
             bool best_effort_uniquification = false,
             const float uniquification_percent = 100,
             bool swap_buffers = true) {
  std::size_t nv = static_cast<std::size_t>(E->get_problem()->get_graph().get_number_of_vertices());
  execute<type>(E->get_input_frontier(), E->get_output_frontier(), context,
                best_effort_uniquification, uniquification_percent, nv);
  if (swap_buffers)
    E->swap_frontier_buffers();
}

}  // namespace uniquify
}  // namespace operators
}  // namespace gunrock

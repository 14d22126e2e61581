/**
 * @file launch_box.hxx
 * @brief `gcuda::launch_box_t` and friends (include/gunrock/cuda/launch_box.hxx:32-335,
 * cuda/sm.hxx:21-96, cuda/detail/launch_kernels.hxx:20-53): compile-time selection of launch
 * parameters per SM target, plus `launch / launch_strided / launch_blocked / launch_cooperative`.
 *
 * Kept for user kernels written against the reference.  The B200 operators do not use it: their
 * grids are sized at run time from the SM count (persistent grids, gunrock/b200/runtime.cuh).
 * Differences: `sm_flag_t` knows `sm_100` (the reference stops at sm_90 and `-DSM_TARGET=100`
 * does not compile there); `SM_TARGET` defaults to 100; a box with no matching
 * entry falls back to its `fallback` entry or fails to compile with a readable message.
 */
#pragma once

#include <cstddef>
#include <tuple>
#include <type_traits>
#include <utility>

#include <gunrock/cuda/context.hxx>

#ifndef SM_TARGET
#define SM_TARGET 100
#endif

namespace gunrock {
namespace gcuda {

enum sm_flag_t : unsigned {
  fallback = ~0u,
  sm_30 = 30, sm_35 = 35, sm_37 = 37, sm_50 = 50, sm_52 = 52, sm_53 = 53, sm_60 = 60, sm_61 = 61,
  sm_62 = 62, sm_70 = 70, sm_72 = 72, sm_75 = 75, sm_80 = 80, sm_86 = 86, sm_87 = 87, sm_89 = 89,
  sm_90 = 90, sm_100 = 100, sm_103 = 103, sm_120 = 120
};

constexpr sm_flag_t operator|(sm_flag_t a, sm_flag_t) { return a; }  // accepted, first flag wins

template <unsigned int x_ = 1, unsigned int y_ = 1, unsigned int z_ = 1>
struct dim3_t {
  enum : unsigned int { x = x_, y = y_, z = z_ };
  static constexpr unsigned int size() { return x_ * y_ * z_; }
  static constexpr dim3 get_dim3() { return dim3(x_, y_, z_); }
};

namespace kernels {
namespace detail {
template <typename func_t, typename... args_t>
__global__ void strided_kernel(func_t f, const std::size_t bound, args_t... args) {
  const std::size_t stride = static_cast<std::size_t>(blockDim.x) * gridDim.x;
  for (std::size_t i = blockDim.x * blockIdx.x + threadIdx.x; i < bound; i += stride)
    f(i, blockIdx.x, args...);
}
template <unsigned items_per_thread, typename func_t, typename... args_t>
__global__ void blocked_kernel(func_t f, const std::size_t bound, args_t... args) {
  const std::size_t first =
      (static_cast<std::size_t>(blockDim.x) * blockIdx.x + threadIdx.x) * items_per_thread;
#pragma unroll
  for (unsigned k = 0; k < items_per_thread; ++k)
    if (first + k < bound)
      f(first + k, blockIdx.x, args...);
}
}  // namespace detail
}  // namespace kernels

/// Fixed launch parameters for one SM target.
template <sm_flag_t sm_flags_,
          typename block_dimensions_,
          typename grid_dimensions_,
          std::size_t shared_memory_bytes_ = 0>
struct launch_params_t {
  typedef block_dimensions_ block_dimensions_t;
  typedef grid_dimensions_ grid_dimensions_t;
  enum : unsigned { sm_flags = sm_flags_ };
  static constexpr std::size_t shared_memory_bytes = shared_memory_bytes_;
  static constexpr unsigned items_per_thread = 1;
  dim3 block_dimensions = block_dimensions_t::get_dim3();
  dim3 grid_dimensions = grid_dimensions_t::get_dim3();
  void calculate_grid_dimensions_strided(std::size_t) {}
  void calculate_grid_dimensions_blocked(std::size_t) {}
};

/// Launch parameters whose grid is derived from the element count.
template <sm_flag_t sm_flags_,
          typename block_dimensions_,
          std::size_t items_per_thread_ = 1,
          std::size_t shared_memory_bytes_ = 0>
struct launch_params_dynamic_grid_t {
  typedef block_dimensions_ block_dimensions_t;
  enum : unsigned { sm_flags = sm_flags_ };
  static constexpr std::size_t shared_memory_bytes = shared_memory_bytes_;
  static constexpr unsigned items_per_thread = static_cast<unsigned>(items_per_thread_);
  dim3 block_dimensions = block_dimensions_t::get_dim3();
  dim3 grid_dimensions = dim3(1, 1, 1);
  void calculate_grid_dimensions_strided(std::size_t num_elements) {
    std::size_t per = block_dimensions_t::size();
    grid_dimensions = dim3(static_cast<unsigned>((num_elements + per - 1) / per), 1, 1);
  }
  void calculate_grid_dimensions_blocked(std::size_t num_elements) {
    std::size_t per = block_dimensions_t::size() * items_per_thread_;
    grid_dimensions = dim3(static_cast<unsigned>((num_elements + per - 1) / per), 1, 1);
  }
};

namespace detail {
template <typename lp_t>
constexpr bool matches_target() {
  return static_cast<unsigned>(lp_t::sm_flags) == static_cast<unsigned>(SM_TARGET);
}
template <typename lp_t>
constexpr bool is_fallback() {
  return static_cast<unsigned>(lp_t::sm_flags) == static_cast<unsigned>(fallback);
}
struct no_launch_params_for_this_sm_target {};

// first entry matching SM_TARGET, else first fallback entry
template <bool want_fallback, typename... lp_v>
struct pick_t {
  using type = no_launch_params_for_this_sm_target;
};
template <bool want_fallback, typename lp_t, typename... rest>
struct pick_t<want_fallback, lp_t, rest...> {
  using type = std::conditional_t<(want_fallback ? is_fallback<lp_t>() : matches_target<lp_t>()),
                                  lp_t,
                                  typename pick_t<want_fallback, rest...>::type>;
};
template <typename... lp_v>
using select_t = std::conditional_t<
    !std::is_same<typename pick_t<false, lp_v...>::type, no_launch_params_for_this_sm_target>::value,
    typename pick_t<false, lp_v...>::type,
    typename pick_t<true, lp_v...>::type>;
}  // namespace detail

template <typename... lp_v>
using select_launch_params_t = detail::select_t<lp_v...>;

template <typename... lp_v>
struct launch_box_t : public select_launch_params_t<lp_v...> {
  typedef select_launch_params_t<lp_v...> params_t;
  static_assert(!std::is_same<params_t, detail::no_launch_params_for_this_sm_target>::value,
                "launch_box_t: no launch parameters for this SM_TARGET and no fallback entry");
  launch_box_t() {}

  template <typename func_t, typename... args_t>
  void launch_strided(gcuda::standard_context_t& context, func_t& f,
                      const std::size_t num_elements, args_t&&... args) {
    params_t::calculate_grid_dimensions_strided(num_elements);
    kernels::detail::strided_kernel<<<params_t::grid_dimensions, params_t::block_dimensions,
                                     params_t::shared_memory_bytes, context.stream()>>>(
        f, num_elements, std::forward<args_t>(args)...);
  }
  template <typename func_t, typename... args_t>
  void launch_blocked(gcuda::standard_context_t& context, func_t& f,
                      const std::size_t num_elements, args_t&&... args) {
    params_t::calculate_grid_dimensions_blocked(num_elements);
    kernels::detail::blocked_kernel<params_t::items_per_thread>
        <<<params_t::grid_dimensions, params_t::block_dimensions, params_t::shared_memory_bytes,
           context.stream()>>>(f, num_elements, std::forward<args_t>(args)...);
  }
  template <typename func_t, typename... args_t>
  void launch_cooperative(gcuda::standard_context_t& context, const func_t& f,
                          const std::size_t num_elements, args_t&&... args) {
    params_t::calculate_grid_dimensions_strided(num_elements);
    void* argument_ptrs[sizeof...(args_t) == 0 ? 1 : sizeof...(args_t)] = {
        const_cast<void*>(static_cast<const void*>(&args))...};
    error::throw_if_exception(
        cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(&f), params_t::grid_dimensions,
                                    params_t::block_dimensions, argument_ptrs,
                                    params_t::shared_memory_bytes, context.stream()),
        "cudaLaunchCooperativeKernel");
  }
  template <typename func_t, typename... args_t>
  void launch(gcuda::standard_context_t& context, const func_t& f, args_t&&... args) {
    f<<<params_t::grid_dimensions, params_t::block_dimensions, params_t::shared_memory_bytes,
        context.stream()>>>(std::forward<args_t>(args)...);
  }
};

}  // namespace gcuda
}  // namespace gunrock

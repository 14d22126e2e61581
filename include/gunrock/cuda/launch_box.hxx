/**
 * @file launch_box.hxx
 * @brief `gcuda::launch_box::launch_box_t` and friends (include/gunrock/cuda/launch_box.hxx:30-362,
 * cuda/sm.hxx:15-96, cuda/detail/launch_kernels.hxx:20-53): compile-time selection of launch
 * parameters per SM target, plus `launch / launch_strided / launch_blocked / launch_cooperative` and
 * `occupancy<box_t>(kernel)`.  Same names, namespace and template signatures as the reference --
 * `launch_params_t<flags, block, grid, items_per_thread = 1, shared_memory_bytes = 0>`,
 * `launch_params_dynamic_grid_t<flags, block, items_per_thread = 1, shared_memory_bytes = 0>`, SM flags that
 * combine with `|` -- so its unit test (unittests/cuda/launch_box.cuh) compiles unchanged; `namespace gcuda`
 * pulls the names in as well (`gcuda::launch_box_t`).
 *
 * Kept for user kernels written against the reference.  The B200 operators do not use it: their
 * grids are sized at run time from the SM count (persistent grids, gunrock/b200/runtime.cuh).
 * Differences: `sm_flag_t` knows sm_100 / sm_103 / sm_120 (the reference stops at sm_90 and
 * `-DSM_TARGET=100` does not compile there); `SM_TARGET` defaults to 100; an entry naming the
 * target wins over a `fallback` entry wherever it stands in the list (the reference takes the first entry that
 * matches, fallback included); a box with neither fails to compile with a readable message; fixed-grid
 * parameters may be used with `launch_strided` / `launch_blocked` (their grid stays as declared).
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
namespace launch_box {

/// One bit per SM version, so that one entry can serve several (`sm_90 | sm_100`); `fallback` matches all.
enum sm_flag_t : unsigned {
  fallback = ~0u,
  sm_30 = 1u << 0, sm_35 = 1u << 1, sm_37 = 1u << 2, sm_50 = 1u << 3, sm_52 = 1u << 4, sm_53 = 1u << 5,
  sm_60 = 1u << 6, sm_61 = 1u << 7, sm_62 = 1u << 8, sm_70 = 1u << 9, sm_72 = 1u << 10, sm_75 = 1u << 11,
  sm_80 = 1u << 12, sm_86 = 1u << 13, sm_87 = 1u << 14, sm_89 = 1u << 15, sm_90 = 1u << 16,
  sm_100 = 1u << 17, sm_103 = 1u << 18, sm_120 = 1u << 19
};
constexpr sm_flag_t operator|(sm_flag_t a, sm_flag_t b) {
  return static_cast<sm_flag_t>(static_cast<unsigned>(a) | static_cast<unsigned>(b));
}
constexpr sm_flag_t operator&(sm_flag_t a, sm_flag_t b) {
  return static_cast<sm_flag_t>(static_cast<unsigned>(a) & static_cast<unsigned>(b));
}

/// Run-time value of a `dim3_t`; converts to the CUDA `dim3` a launch takes.
struct dimensions_t {
  unsigned int x, y, z;
  __host__ __device__ constexpr dimensions_t(const unsigned int _x = 1, const unsigned int _y = 1,
                                             const unsigned int _z = 1)
      : x(_x), y(_y), z(_z) {}
  __host__ __device__ constexpr unsigned int size() const { return x * y * z; }
  __host__ __device__ constexpr operator dim3(void) const { return dim3{x, y, z}; }
};

template <unsigned int x_ = 1, unsigned int y_ = 1, unsigned int z_ = 1>
struct dim3_t {
  enum : unsigned int { x = x_, y = y_, z = z_ };
  static constexpr unsigned int size() { return x_ * y_ * z_; }
  static constexpr dimensions_t dimensions() { return {x_, y_, z_}; }
  static constexpr dim3 get_dim3() { return dim3(x_, y_, z_); }
  constexpr operator dimensions_t(void) const { return {x_, y_, z_}; }
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

namespace detail {
/// The flag of a numeric SM_TARGET (100 -> sm_100); 0 for a version the table does not know.
constexpr unsigned flag_of_target(unsigned sm) {
  return sm == 30 ? sm_30 : sm == 35 ? sm_35 : sm == 37 ? sm_37 : sm == 50 ? sm_50 : sm == 52 ? sm_52
       : sm == 53 ? sm_53 : sm == 60 ? sm_60 : sm == 61 ? sm_61 : sm == 62 ? sm_62 : sm == 70 ? sm_70
       : sm == 72 ? sm_72 : sm == 75 ? sm_75 : sm == 80 ? sm_80 : sm == 86 ? sm_86 : sm == 87 ? sm_87
       : sm == 89 ? sm_89 : sm == 90 ? sm_90 : sm == 100 ? sm_100 : sm == 103 ? sm_103
       : sm == 120 ? sm_120 : 0u;
}
template <sm_flag_t sm_flags_, std::size_t items_per_thread_, std::size_t shared_memory_bytes_>
struct launch_params_base_t {
  enum : unsigned { sm_flags = sm_flags_ };
  enum : std::size_t { items_per_thread = items_per_thread_, shared_memory_bytes = shared_memory_bytes_ };
};
}  // namespace detail

/// Fixed launch parameters for the SM targets named by `sm_flags_`.
template <sm_flag_t sm_flags_,
          typename block_dimensions_,
          typename grid_dimensions_,
          std::size_t items_per_thread_ = 1,
          std::size_t shared_memory_bytes_ = 0>
struct launch_params_t : detail::launch_params_base_t<sm_flags_, items_per_thread_, shared_memory_bytes_> {
  typedef detail::launch_params_base_t<sm_flags_, items_per_thread_, shared_memory_bytes_> base_t;
  typedef block_dimensions_ block_dimensions_t;
  typedef grid_dimensions_ grid_dimensions_t;
  static constexpr dimensions_t block_dimensions = block_dimensions_t::dimensions();
  static constexpr dimensions_t grid_dimensions = grid_dimensions_t::dimensions();
  void calculate_grid_dimensions_strided(std::size_t) {}  // the grid stays as declared
  void calculate_grid_dimensions_blocked(std::size_t) {}
};

/// Launch parameters whose grid is derived from the element count.
template <sm_flag_t sm_flags_,
          typename block_dimensions_,
          std::size_t items_per_thread_ = 1,
          std::size_t shared_memory_bytes_ = 0>
struct launch_params_dynamic_grid_t
    : detail::launch_params_base_t<sm_flags_, items_per_thread_, shared_memory_bytes_> {
  typedef detail::launch_params_base_t<sm_flags_, items_per_thread_, shared_memory_bytes_> base_t;
  typedef block_dimensions_ block_dimensions_t;
  static constexpr dimensions_t block_dimensions = block_dimensions_t::dimensions();
  dimensions_t grid_dimensions;
  void calculate_grid_dimensions_strided(std::size_t num_elements) {
    const std::size_t per = block_dimensions_t::size();
    grid_dimensions = dimensions_t(static_cast<unsigned>((num_elements + per - 1) / per), 1, 1);
  }
  void calculate_grid_dimensions_blocked(std::size_t num_elements) {
    const std::size_t per = block_dimensions_t::size() * items_per_thread_;
    grid_dimensions = dimensions_t(static_cast<unsigned>((num_elements + per - 1) / per), 1, 1);
  }
};

namespace detail {
template <typename lp_t>
constexpr bool is_fallback() {
  return static_cast<unsigned>(lp_t::sm_flags) == static_cast<unsigned>(fallback);
}
template <typename lp_t>
constexpr bool matches_target() {
  return !is_fallback<lp_t>() &&
         (static_cast<unsigned>(lp_t::sm_flags) & flag_of_target(static_cast<unsigned>(SM_TARGET))) != 0u;
}
struct no_launch_params_for_this_sm_target {};

// first entry naming SM_TARGET, else first fallback entry
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
    kernels::detail::blocked_kernel<static_cast<unsigned>(params_t::items_per_thread)>
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

/// Fraction of an SM's resident warps `kernel` reaches at the box's block size (launch_box.hxx:343-360).
template <typename launch_box_t, typename func_t>
inline float occupancy(func_t kernel) {
  int max_active_blocks = 0;
  const int block_size = static_cast<int>(launch_box_t::block_dimensions_t::size());
  int device = 0;
  cudaDeviceProp props;
  error::throw_if_exception(cudaGetDevice(&device), "cudaGetDevice");
  error::throw_if_exception(cudaGetDeviceProperties(&props, device), "cudaGetDeviceProperties");
  error::throw_if_exception(
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&max_active_blocks, kernel, block_size, (size_t)0),
      "cudaOccupancyMaxActiveBlocksPerMultiprocessor");
  return (max_active_blocks * block_size / props.warpSize) /
         static_cast<float>(props.maxThreadsPerMultiProcessor / props.warpSize);
}

}  // namespace launch_box

using namespace launch_box;  // gcuda::launch_box_t, gcuda::sm_100 ... as before

}  // namespace gcuda
}  // namespace gunrock

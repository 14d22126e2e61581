/**
 * @file dense_frontier.hxx
 * @brief Dense frontier views: `frontier_t<..., frontier_view_t::bitmap>` (one bit per vertex) and
 * `frontier_t<..., frontier_view_t::boolmap>` (one byte per vertex).
 *
 * The reference declares both views (framework/frontier/configs.hxx:19-23) but only ships an
 * experimental boolmap class that no operator accepts (frontier/experimental/boolmap_frontier.hxx:22-211;
 * its `fill` throws for every value).  Here they are real containers with the reference's member names --
 * `get_element_at(i)` answers `i` when the element is present and the invalid id otherwise,
 * `set_element_at(v)` inserts `v`, a dense frontier "is always sorted" -- plus what the
 * direction-optimised BFS needs from them: O(V/32) population count on the device, and conversion from /
 * to the vector view (`from_vector`, `to_vector`) so that every vector operator composes with them.
 * Copies share storage (shared_ptr), as the vector frontier does.
 */
#pragma once

#include <memory>
#include <type_traits>
#include <vector>

#include <cuda_runtime.h>

#include <gunrock/b200/ptx.cuh>
#include <gunrock/b200/runtime.cuh>
#include <gunrock/error.hxx>
#include <gunrock/framework/frontier/configs.hxx>
#include <gunrock/framework/frontier/frontier.hxx>  // the primary template (vector view)
#include <gunrock/util/type_limits.hxx>

namespace gunrock {
namespace frontier {

namespace detail {

/// popcount of a word array / count of non-zero bytes, one atomic per warp
static __global__ void dense_count_kernel(const unsigned* __restrict__ words, std::size_t n_words,
                                          int bits_per_element, int* __restrict__ count) {
  int local = 0;
  for (std::size_t i = blockIdx.x * static_cast<std::size_t>(blockDim.x) + threadIdx.x; i < n_words;
       i += static_cast<std::size_t>(gridDim.x) * blockDim.x) {
    unsigned w = words[i];
    if (bits_per_element == 1)
      local += __popc(w);
    else  // four byte flags per word
      local += ((w & 0xffu) != 0) + ((w & 0xff00u) != 0) + ((w & 0xff0000u) != 0) + ((w & 0xff000000u) != 0);
  }
  local = b200::warp_sum(local);
  if (b200::lane_id() == 0 && local)
    atomicAdd(count, local);
}

static __global__ void dense_set_range_kernel(unsigned* words, int bits_per_element, std::size_t first,
                                              std::size_t n) {
  for (std::size_t k = blockIdx.x * static_cast<std::size_t>(blockDim.x) + threadIdx.x; k < n;
       k += static_cast<std::size_t>(gridDim.x) * blockDim.x) {
    std::size_t v = first + k;
    if (bits_per_element == 1)
      atomicOr(words + (v >> 5), 1u << (v & 31));
    else
      reinterpret_cast<unsigned char*>(words)[v] = 1;
  }
}

static __global__ void dense_from_list_kernel(unsigned* words, int bits_per_element, const int* __restrict__ list,
                                              const int* __restrict__ list_count) {
  const int n = *list_count;
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
    int v = list[k];
    if (v < 0)
      continue;  // invalid slots of a bypass-filtered frontier
    if (bits_per_element == 1)
      atomicOr(words + (v >> 5), 1u << (v & 31));
    else
      reinterpret_cast<unsigned char*>(words)[v] = 1;
  }
}

/// Enumerate the present elements in ascending order within a warp's 32 words (global order follows
/// the atomic, i.e. unspecified across warps -- callers that need sorted output sort the vector).
static __global__ void dense_to_list_kernel(const unsigned* __restrict__ words, int bits_per_element,
                                            std::size_t universe, int* __restrict__ list, int* list_count,
                                            int capacity, int* overflow) {
  const int lane = b200::lane_id();
  const std::size_t warps = (static_cast<std::size_t>(gridDim.x) * blockDim.x) >> 5;
  const std::size_t gw = (blockIdx.x * static_cast<std::size_t>(blockDim.x) + threadIdx.x) >> 5;
  const std::size_t groups = (universe + 31) / 32;  // 32 elements per lane-step
  for (std::size_t g0 = gw * 32; g0 < groups; g0 += warps * 32) {
    const std::size_t g = g0 + lane;  // this lane's group of 32 elements
    unsigned present = 0;
    if (g < groups) {
      if (bits_per_element == 1) {
        present = words[g];
      } else {
        const unsigned* w8 = words + g * 8;  // 32 bytes
        for (int k = 0; k < 8; ++k) {
          unsigned w = (g * 32 + k * 4 < universe) ? w8[k] : 0u;
          for (int b = 0; b < 4; ++b)
            if ((w >> (8 * b)) & 0xffu)
              present |= 1u << (k * 4 + b);
        }
      }
      if (g * 32 + 32 > universe)  // tail beyond the universe never counts
        present &= (universe - g * 32 >= 32) ? 0xffffffffu : ((1u << (universe - g * 32)) - 1u);
    }
    const int c = __popc(present);
    const int incl = b200::warp_inclusive_sum(c);
    const int total = __shfl_sync(b200::kFull, incl, 31);
    if (!total)
      continue;
    int base = 0;
    if (lane == 0)
      base = atomicAdd(list_count, total);
    base = __shfl_sync(b200::kFull, base, 0) + incl - c;
    while (present) {
      int b = __ffs(present) - 1;
      present &= present - 1;
      if (base < capacity)
        list[base] = static_cast<int>(g * 32 + b);
      else
        *overflow = 1;
      ++base;
    }
  }
}

struct dense_storage_t {
  unsigned* words = nullptr;  // bitmap words, or boolmap bytes viewed as words
  std::size_t universe = 0;   // number of representable elements (vertices)
  std::size_t n_words = 0;
  int* d_scratch = nullptr;  // [0] count, [1] overflow
  int* h_scratch = nullptr;  // pinned mirror
  cudaStream_t stream = nullptr;
  dense_storage_t() {
    error::throw_if_exception(cudaMalloc(&d_scratch, 2 * sizeof(int)), "dense frontier scratch");
    error::throw_if_exception(cudaMallocHost(&h_scratch, 2 * sizeof(int)), "dense frontier pinned");
  }
  dense_storage_t(const dense_storage_t&) = delete;
  dense_storage_t& operator=(const dense_storage_t&) = delete;
  ~dense_storage_t() {
    cudaFree(words);
    cudaFree(d_scratch);
    cudaFreeHost(h_scratch);
  }
};

/// Shared implementation; kBits = 1 (bitmap) or 8 (boolmap).
template <typename type_t, int kBits>
class dense_frontier_t {
  static_assert(sizeof(type_t) == 4, "frontier elements are 32-bit ids");

 public:
  using pointer_t = unsigned*;

  dense_frontier_t() : p_storage(std::make_shared<dense_storage_t>()) {}
  explicit dense_frontier_t(std::size_t universe) : dense_frontier_t() { resize(universe); }
  dense_frontier_t(const dense_frontier_t&) = default;
  dense_frontier_t& operator=(const dense_frontier_t&) = default;

  // ---- host side -----------------------------------------------------------------------------------
  /// Size the view for ids in [0, universe); every element absent afterwards.
  void resize(std::size_t universe, type_t const = 0) {
    auto& s = *p_storage;
    const std::size_t need = kBits == 1 ? (universe + 31) / 32 : (universe + 3) / 4;
    if (need > s.n_words) {
      cudaFree(s.words);
      s.words = nullptr;
      error::throw_if_exception(cudaMalloc(&s.words, (need + 8) * sizeof(unsigned)), "dense frontier alloc");
      s.n_words = need;
    }
    s.universe = universe;
    clear();
    // (re)allocation is rare and the caller may bind another -- possibly non-blocking -- stream right after:
    // the zero fill must have landed before anything enqueued there touches the words
    error::throw_if_exception(cudaStreamSynchronize(s.stream), "dense frontier resize");
    sync_view();
  }
  void reserve(std::size_t universe) {
    if (universe > p_storage->universe)
      resize(universe);
  }
  void clear() {
    auto& s = *p_storage;
    if (s.n_words)
      error::throw_if_exception(cudaMemsetAsync(s.words, 0, s.n_words * sizeof(unsigned), s.stream),
                                "dense frontier clear");
  }
  /// fill(1): every id of the universe present; fill(0): none (the reference's boolmap accepts 0 / 1 only).
  void fill(type_t const value, cudaStream_t = 0) {
    error::throw_if_exception(!(value == 0 || value == 1), "dense frontiers only support 1 or 0 as fill value");
    clear();
    if (value == 1)
      sequence(0, p_storage->universe);
  }
  /// Insert the ids [first, first + n).
  void sequence(type_t const first, std::size_t const& n, cudaStream_t = 0) {
    auto& s = *p_storage;
    error::throw_if_exception(static_cast<std::size_t>(first) + n > s.universe, "dense frontier: id out of range");
    if (n)
      dense_set_range_kernel<<<256, 256, 0, s.stream>>>(s.words, kBits, static_cast<std::size_t>(first), n);
  }
  void push_back(type_t const& value) { sequence(value, 1); }
  std::size_t get_capacity() const { return p_storage->universe; }
  std::size_t get_universe() const { return p_storage->universe; }
  /// Population count (one small kernel + one stream synchronisation).
  std::size_t get_number_of_elements(cudaStream_t = 0) const {
    auto& s = *p_storage;
    if (!s.n_words)
      return 0;
    error::throw_if_exception(cudaMemsetAsync(s.d_scratch, 0, sizeof(int), s.stream), "dense count reset");
    dense_count_kernel<<<256, 256, 0, s.stream>>>(s.words, s.n_words, kBits, s.d_scratch);
    error::throw_if_exception(
        cudaMemcpyAsync(s.h_scratch, s.d_scratch, sizeof(int), cudaMemcpyDeviceToHost, s.stream), "dense count read");
    error::throw_if_exception(cudaStreamSynchronize(s.stream), "dense count sync");
    return static_cast<std::size_t>(s.h_scratch[0]);
  }
  bool is_empty() const { return get_number_of_elements() == 0; }
  void sort(int = 0, cudaStream_t = 0) {}  // always sorted
  void bind_stream(cudaStream_t st) { p_storage->stream = st; }
  pointer_t data() { return p_storage->words; }

  /// Insert every valid id of a vector frontier (device-resident count: no host round trip).
  template <typename vector_frontier_t>
  void from_vector(vector_frontier_t& in, bool clear_first = true) {
    auto& s = *p_storage;
    if (clear_first)
      clear();
    dense_from_list_kernel<<<512, 256, 0, s.stream>>>(s.words, kBits, reinterpret_cast<const int*>(in.get()),
                                                      in.count_ptr());
  }
  /// Write the present ids into a vector frontier (its count stays on the device).
  template <typename vector_frontier_t>
  void to_vector(vector_frontier_t& out) {
    auto& s = *p_storage;
    if (out.get_capacity() < s.universe)
      out.reserve(s.universe);
    out.bind_stream(s.stream);
    out.set_number_of_elements(0);
    error::throw_if_exception(cudaMemsetAsync(s.d_scratch + 1, 0, sizeof(int), s.stream), "dense overflow reset");
    dense_to_list_kernel<<<512, 256, 0, s.stream>>>(s.words, kBits, s.universe, reinterpret_cast<int*>(out.get()),
                                                    out.count_ptr(), static_cast<int>(out.get_capacity()),
                                                    s.d_scratch + 1);
    out.mark_produced(s.stream, nullptr, /*unique=*/true);  // one id per set bit
  }

  // ---- device side (the object is passed to kernels by value) --------------------------------------
  __host__ __device__ __forceinline__ constexpr pointer_t get() const { return raw; }
  __device__ __forceinline__ bool contains(type_t const& v) const {
    if (kBits == 1)
      return (raw[static_cast<unsigned>(v) >> 5] >> (v & 31)) & 1u;
    return reinterpret_cast<const unsigned char*>(raw)[v] != 0;
  }
  /// `idx` when that id is present, the invalid id otherwise (boolmap_frontier.hxx:101-106).
  __device__ __forceinline__ type_t get_element_at(std::size_t const& idx) const noexcept {
    return contains(static_cast<type_t>(idx)) ? static_cast<type_t>(idx)
                                               : gunrock::numeric_limits<type_t>::invalid();
  }
  /// Insert `element` (`idx` is ignored, boolmap_frontier.hxx:115-120).  Returns true when it was absent.
  __device__ __forceinline__ bool set_element_at(type_t const& element, std::size_t const& = 0) const noexcept {
    if (kBits == 1) {
      const unsigned bit = 1u << (element & 31);
      return !(atomicOr(raw + (static_cast<unsigned>(element) >> 5), bit) & bit);
    }
    unsigned char* b = reinterpret_cast<unsigned char*>(raw) + element;
    const bool was = *b != 0;
    *b = 1;
    return !was;
  }
  __device__ __forceinline__ void remove_element(type_t const& element) const noexcept {
    if (kBits == 1)
      atomicAnd(raw + (static_cast<unsigned>(element) >> 5), ~(1u << (element & 31)));
    else
      reinterpret_cast<unsigned char*>(raw)[element] = 0;
  }

 private:
  void sync_view() { raw = p_storage->words; }
  std::shared_ptr<dense_storage_t> p_storage;
  unsigned* raw = nullptr;
};

}  // namespace detail

template <typename vertex_t, typename edge_t, frontier_kind_t _kind>
class frontier_t<vertex_t, edge_t, _kind, frontier_view_t::bitmap>
    : public detail::dense_frontier_t<std::conditional_t<_kind == frontier_kind_t::vertex_frontier, vertex_t, edge_t>, 1> {
  using base_t =
      detail::dense_frontier_t<std::conditional_t<_kind == frontier_kind_t::vertex_frontier, vertex_t, edge_t>, 1>;

 public:
  using vertex_type = vertex_t;
  using edge_type = edge_t;
  using type_t = std::conditional_t<_kind == frontier_kind_t::vertex_frontier, vertex_t, edge_t>;
  using base_t::base_t;
  constexpr frontier_kind_t get_kind() const { return _kind; }
  constexpr frontier_view_t get_view() const { return frontier_view_t::bitmap; }
};

template <typename vertex_t, typename edge_t, frontier_kind_t _kind>
class frontier_t<vertex_t, edge_t, _kind, frontier_view_t::boolmap>
    : public detail::dense_frontier_t<std::conditional_t<_kind == frontier_kind_t::vertex_frontier, vertex_t, edge_t>, 8> {
  using base_t =
      detail::dense_frontier_t<std::conditional_t<_kind == frontier_kind_t::vertex_frontier, vertex_t, edge_t>, 8>;

 public:
  using vertex_type = vertex_t;
  using edge_type = edge_t;
  using type_t = std::conditional_t<_kind == frontier_kind_t::vertex_frontier, vertex_t, edge_t>;
  using base_t::base_t;
  constexpr frontier_kind_t get_kind() const { return _kind; }
  constexpr frontier_view_t get_view() const { return frontier_view_t::boolmap; }
};

}  // namespace frontier
}  // namespace gunrock

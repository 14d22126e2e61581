/**
 * @file frontier.hxx
 * @brief `frontier::frontier_t` -- the vector frontier of the reference
 * (include/gunrock/framework/frontier/frontier.hxx:32-147, vector_frontier.hxx:27-311) redesigned
 * around a DEVICE-RESIDENT element count.
 *
 * The reference keeps `num_elements` on the host, so every operator ends with a stream sync and a
 * D2H copy (advance/helpers.hxx:106-110).  Here operators write the count in device memory and
 * chain without host involvement; the host copy is refreshed lazily the first time somebody
 * asks (`get_number_of_elements()`, `is_empty()`, `push_back` ...), which is also where an output
 * overflow reported by a kernel is turned into an exception.  Copies of a frontier share their
 * storage (shared_ptr), as in the reference (vector_frontier.hxx:72-82), so by-value kernel
 * arguments alias it.
 */
#pragma once

#include <algorithm>
#include <iostream>
#include <memory>
#include <vector>

#include <cuda_runtime.h>
#include <thrust/device_ptr.h>
#include <thrust/sort.h>

#include <gunrock/b200/runtime.cuh>
#include <gunrock/cuda/context.hxx>
#include <gunrock/error.hxx>
#include <gunrock/framework/frontier/configs.hxx>
#include <gunrock/util/load_store.hxx>
#include <gunrock/util/type_limits.hxx>

namespace gunrock {

namespace sort {
enum order_t { ascending, descending };
}

namespace frontier {

namespace detail {

static __global__ void iota_kernel(int* p, int first, int n) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    p[i] = first + i;
}
static __global__ void fill_kernel(int* p, int value, int n) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    p[i] = value;
}

/// Device buffer + device count + pinned mirror of the count.
template <typename type_t>
struct storage_t {
  type_t* data = nullptr;
  std::size_t capacity = 0;
  int* d_count = nullptr;  // [0] number of elements, [1] overflow flag of the producing operator; device resident
  int* h_count = nullptr;  // pinned: [0..1] read-back of both, [2..3] staging of set_count
  std::size_t host_count = 0;
  bool dirty = false;                      // a kernel may have changed *d_count
  bool unique_known = true;                // no vertex is known to occur twice (an empty frontier qualifies)
  cudaStream_t stream = nullptr;           // stream of the last device-side writer

  storage_t() {
    error::throw_if_exception(cudaMalloc(&d_count, 2 * sizeof(int)), "frontier count alloc");
    error::throw_if_exception(cudaMallocHost(&h_count, 4 * sizeof(int)), "frontier pinned alloc");
    error::throw_if_exception(cudaMemset(d_count, 0, 2 * sizeof(int)), "frontier count init");
  }
  storage_t(const storage_t&) = delete;
  storage_t& operator=(const storage_t&) = delete;
  ~storage_t() {
    cudaFree(data);
    cudaFree(d_count);
    cudaFreeHost(h_count);
  }
  void reserve(std::size_t n) {
    if (n <= capacity)
      return;
    type_t* fresh = nullptr;
    error::throw_if_exception(cudaMalloc(&fresh, (n + 64) * sizeof(type_t)), "frontier alloc");
    std::size_t keep = refresh();
    if (keep && data)
      error::throw_if_exception(
          cudaMemcpy(fresh, data, keep * sizeof(type_t), cudaMemcpyDeviceToDevice), "frontier grow");
    cudaFree(data);
    data = fresh;
    capacity = n;
  }
  /// Bring the host copy of the count up to date (one stream sync when a kernel wrote it).
  std::size_t refresh() {
    if (dirty) {
      // count and overflow flag live side by side in THIS frontier's storage (the producing operator copied
      // its flag here in-stream, mark_produced), so the answer does not depend on how many operator launches
      // happened since -- the workspace's control blocks are a ring that gets recycled
      error::throw_if_exception(
          cudaMemcpyAsync(h_count, d_count, 2 * sizeof(int), cudaMemcpyDeviceToHost, stream),
          "frontier count read-back");
      error::throw_if_exception(cudaStreamSynchronize(stream), "frontier count sync");
      dirty = false;
      error::throw_if_exception(h_count[1] != 0,
                                "output frontier exceeded its capacity; reserve() a larger frontier");
      host_count = static_cast<std::size_t>(h_count[0]);
    }
    return host_count;
  }
  void set_count(std::size_t n) {
    host_count = n;
    dirty = false;
    unique_known = n <= 1;
    int v[2] = {static_cast<int>(n), 0};  // pageable source: the runtime stages it before returning
    error::throw_if_exception(
        cudaMemcpyAsync(d_count, v, 2 * sizeof(int), cudaMemcpyHostToDevice, stream),
        "frontier count write");
  }
};

}  // namespace detail

template <typename vertex_t,
          typename edge_t,
          frontier_kind_t _kind = frontier_kind_t::vertex_frontier,
          frontier_view_t _view = frontier_view_t::vector>
class frontier_t {
 public:
  using vertex_type = vertex_t;
  using edge_type = edge_t;
  using type_t =
      std::conditional_t<_kind == frontier_kind_t::vertex_frontier, vertex_t, edge_t>;
  using frontier_type = frontier_t<vertex_t, edge_t, _kind, _view>;
  static_assert(sizeof(type_t) == sizeof(int), "frontier elements are 32-bit ids");

  frontier_t() : p_storage(std::make_shared<detail::storage_t<type_t>>()), resizing_factor(1.0f) {
    sync_view();
  }
  frontier_t(std::size_t size, float frontier_resizing_factor = 1.0f)
      : p_storage(std::make_shared<detail::storage_t<type_t>>()),
        resizing_factor(frontier_resizing_factor) {
    reserve(size);
  }
  frontier_t(const frontier_t& rhs) = default;
  frontier_t& operator=(const frontier_t& rhs) = default;
  ~frontier_t() = default;

  constexpr frontier_kind_t get_kind() const { return _kind; }
  constexpr frontier_view_t get_view() const { return _view; }

  /// Host: the element count, refreshed from the device if an operator just produced it.
  /// Device: the live device-resident count.
  __host__ __device__ __forceinline__ std::size_t get_number_of_elements(
      cudaStream_t stream = 0) const {
#ifdef __CUDA_ARCH__
    return static_cast<std::size_t>(*raw_count);
#else
    return p_storage->refresh();
#endif
  }
  std::size_t get_capacity() const { return p_storage->capacity; }
  float get_resizing_factor() const { return resizing_factor; }

  __device__ __forceinline__ type_t get_element_at(std::size_t const& idx) const noexcept {
    return thread::load(raw_ptr + idx);
  }
  __device__ __forceinline__ void set_element_at(type_t const& element,
                                                 std::size_t const& idx) const noexcept {
    thread::store(raw_ptr + idx, element);
  }

  void set_resizing_factor(float factor) { resizing_factor = factor; }
  void set_number_of_elements(std::size_t const& elements) { p_storage->set_count(elements); }

  __host__ __device__ __forceinline__ constexpr type_t* get() const { return raw_ptr; }
  type_t* data() { return p_storage->data; }
  type_t* begin() { return this->data(); }
  type_t* end() { return this->begin() + this->get_number_of_elements(); }
  bool is_empty() const { return this->get_number_of_elements() == 0; }

  void push_back(type_t const& value) {
    std::size_t n = get_number_of_elements();
    if (n + 1 > p_storage->capacity)
      reserve(std::max<std::size_t>(2 * (n + 1), 64));
    error::throw_if_exception(
        cudaMemcpyAsync(p_storage->data + n, &value, sizeof(type_t), cudaMemcpyHostToDevice,
                        p_storage->stream),
        "frontier push_back");
    p_storage->set_count(n + 1);
  }

  void fill(type_t const value, cudaStream_t stream = 0) {
    std::size_t n = get_number_of_elements();
    if (n)
      detail::fill_kernel<<<256, 256, 0, p_storage->stream>>>(
          reinterpret_cast<int*>(p_storage->data), static_cast<int>(value), static_cast<int>(n));
    p_storage->unique_known = n <= 1;
  }

  void sequence(type_t const initial_value, std::size_t const& size, cudaStream_t stream = 0) {
    if (get_capacity() < size)
      reserve(size);
    set_number_of_elements(size);
    if (size)
      detail::iota_kernel<<<256, 256, 0, p_storage->stream>>>(
          reinterpret_cast<int*>(p_storage->data), static_cast<int>(initial_value),
          static_cast<int>(size));
    p_storage->unique_known = true;  // consecutive ids
  }

  void resize(std::size_t const& size,
              type_t const default_value = gunrock::numeric_limits<type_t>::invalid()) {
    std::size_t n = get_number_of_elements();
    if (size > p_storage->capacity)
      reserve(size);
    if (size > n)
      detail::fill_kernel<<<256, 256, 0, p_storage->stream>>>(
          reinterpret_cast<int*>(p_storage->data + n), static_cast<int>(default_value),
          static_cast<int>(size - n));
    p_storage->set_count(size);
  }

  void reserve(std::size_t const& size) {
    p_storage->reserve(static_cast<std::size_t>(size * resizing_factor));
    sync_view();
  }

  /// Boundary utility (not on the hot path; the operators dedup with a bitmap instead of sorting).
  void sort(sort::order_t order = sort::order_t::ascending, cudaStream_t stream = 0) {
    std::size_t n = get_number_of_elements();
    thrust::device_ptr<type_t> p(p_storage->data);
    if (order == sort::order_t::ascending)
      thrust::sort(thrust::cuda::par.on(p_storage->stream), p, p + n);
    else
      thrust::sort(thrust::cuda::par.on(p_storage->stream), p, p + n, thrust::greater<type_t>());
  }

  void print() {
    std::size_t n = get_number_of_elements();
    std::vector<type_t> h(n);
    if (n)
      cudaMemcpy(h.data(), p_storage->data, n * sizeof(type_t), cudaMemcpyDeviceToHost);
    std::cout << "Frontier = ";
    for (auto& x : h)
      std::cout << x << " ";
    std::cout << std::endl;
  }

  // --- B200 operator plumbing -------------------------------------------------------------
  /// Device pointer to the element count (what the kernels read/write).
  int* count_ptr() const { return p_storage->d_count; }
  /// Bind the stream device-side operations on this frontier are ordered on.
  void bind_stream(cudaStream_t s) { p_storage->stream = s; }
  /// Called by an operator after it enqueued kernels that rewrite this frontier.  `ctrl` is the operator's
  /// control block: its overflow flag is copied into the frontier's own storage in-stream.
  /// `unique` says whether the operator guarantees that no vertex occurs twice in what it wrote.
  void mark_produced(cudaStream_t s, const b200::ctrl_t* ctrl = nullptr, bool unique = false) {
    p_storage->stream = s;
    p_storage->dirty = true;
    p_storage->unique_known = unique;
    if (ctrl)
      error::throw_if_exception(cudaMemcpyAsync(p_storage->d_count + 1, &ctrl->overflow, sizeof(int),
                                                cudaMemcpyDeviceToDevice, s),
                                "frontier overflow flag");
    else  // an operator that cannot overflow (its output is sized from the input's bound)
      error::throw_if_exception(cudaMemsetAsync(p_storage->d_count + 1, 0, sizeof(int), s),
                                "frontier overflow flag");
  }
  /// False when the same vertex may occur more than once (the expansion of such a frontier is not bounded
  /// by the number of edges of the graph: advance sizes its output from the degree sum then).
  bool is_known_unique() const { return p_storage->unique_known; }
  /// Upper bound on the element count known without a sync (capacity when a kernel wrote it).
  std::size_t size_upper_bound() const {
    return p_storage->dirty ? p_storage->capacity : p_storage->host_count;
  }

 private:
  void sync_view() {
    raw_ptr = p_storage->data;
    raw_count = p_storage->d_count;
  }
  std::shared_ptr<detail::storage_t<type_t>> p_storage;
  type_t* raw_ptr = nullptr;
  int* raw_count = nullptr;
  float resizing_factor;
};

}  // namespace frontier
}  // namespace gunrock

// bitmap / boolmap views: partial specialisations of frontier_t
#include <gunrock/framework/frontier/dense_frontier.hxx>

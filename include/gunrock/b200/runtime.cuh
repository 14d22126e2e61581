/**
 * @file runtime.cuh
 * @brief Host-side plumbing shared by the header API and the C-ABI library: error handling,
 * device properties cache, the per-stream workspace every operator launch draws its scratch
 * from (no cudaMalloc on the hot path -- the reference allocates a device_vector inside every
 * block_mapped advance call, include/gunrock/framework/operators/advance/block_mapped.hxx:244).
 */
#pragma once

#include <cuda_runtime.h>
#include <atomic>
#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>

namespace gunrock {
namespace b200 {

struct cuda_error_t : std::runtime_error {
  cudaError_t code;
  cuda_error_t(cudaError_t c, const std::string& what) : std::runtime_error(what), code(c) {}
};

inline void check(cudaError_t e, const char* expr, const char* file, int line) {
  if (e != cudaSuccess) {
    char buf[512];
    snprintf(buf, sizeof buf, "%s:%d: %s -> %s (%s)", file, line, expr, cudaGetErrorName(e),
             cudaGetErrorString(e));
    throw cuda_error_t(e, buf);
  }
}
#define B2G_CHECK(expr) ::gunrock::b200::check((expr), #expr, __FILE__, __LINE__)

/// Cached per-device facts the launchers size their grids from (148 SMs on B200).
struct device_info_t {
  int device = -1;
  int sm_count = 0;
  int max_smem_optin = 0;
  static const device_info_t& get() {
    static thread_local device_info_t info;
    int dev = 0;
    B2G_CHECK(cudaGetDevice(&dev));
    if (info.device != dev) {
      info.device = dev;
      B2G_CHECK(cudaDeviceGetAttribute(&info.sm_count, cudaDevAttrMultiProcessorCount, dev));
      B2G_CHECK(cudaDeviceGetAttribute(&info.max_smem_optin,
                                       cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    }
    return info;
  }
};

/// Non-owning CSR (or CSC, transposed) view handed to kernels by value.
struct csr_view_t {
  int n_vertices = 0;
  int n_edges = 0;
  const int* __restrict__ row_offsets = nullptr;
  const int* __restrict__ column_indices = nullptr;
  const float* __restrict__ values = nullptr;  // may be null (pattern graph => 1.0f)
  /// Identity of the graph object this view was taken from (graph::build / graph_t::set draw a fresh one
  /// from a process-wide counter).  Per-graph caches (transpose, PageRank tile table, map of vertices
  /// without in-edges) are keyed on it TOGETHER with the array addresses and sizes: an address alone is
  /// reused by cudaMalloc after a free and by grow-only buffers.  0 = unknown: nothing is cached.
  unsigned long long uid = 0;
};

/// Key of a per-graph cache entry.  `matches` is false for views of unknown identity (uid 0).
struct graph_key_t {
  unsigned long long uid = 0;
  const void* offsets = nullptr;
  const void* indices = nullptr;
  int n_vertices = -1, n_edges = -1;
  bool matches(const csr_view_t& v) const {
    return v.uid != 0 && uid == v.uid && offsets == v.row_offsets && indices == v.column_indices &&
           n_vertices == v.n_vertices && n_edges == v.n_edges;
  }
  void set(const csr_view_t& v) {
    uid = v.uid;
    offsets = v.row_offsets;
    indices = v.column_indices;
    n_vertices = v.n_vertices;
    n_edges = v.n_edges;
  }
  void clear() { *this = graph_key_t(); }
};

/// Fresh graph identity (host only; never 0).
inline unsigned long long next_graph_uid() {
  static std::atomic<unsigned long long> counter{0};
  return ++counter;
}

/// A frontier as kernels see it: ids (or -1) plus a device-resident element count.
struct frontier_ref_t {
  int* data = nullptr;
  int* count = nullptr;  // device pointer to the number of elements
  int capacity = 0;
};

/// Device control block zeroed before each operator launch (64 B, one memset node).
struct ctrl_t {
  int work;          // dynamic work-fetch cursor
  int hub_count;     // rows deferred to the CTA/grid bin
  int tile;          // dynamic tile id for look-back scans
  int overflow;      // set when an output frontier would exceed its capacity
  unsigned long long deg_sum;  // sum of degrees of emitted vertices
  unsigned long long edges;    // edges inspected (algorithmic bytes accounting)
  int pad[8];
};
static_assert(sizeof(ctrl_t) == 64, "ctrl_t must stay one 64-byte block");

/// Owning device buffer with grow-only semantics.
template <typename T>
struct dbuf_t {
  T* ptr = nullptr;
  size_t cap = 0;
  dbuf_t() = default;
  dbuf_t(const dbuf_t&) = delete;
  dbuf_t& operator=(const dbuf_t&) = delete;
  ~dbuf_t() {
    if (ptr)
      cudaFree(ptr);
  }
  T* ensure(size_t n) {
    if (n > cap) {
      if (ptr)
        B2G_CHECK(cudaFree(ptr));
      ptr = nullptr;
      size_t want = n + n / 8 + 64;
      B2G_CHECK(cudaMalloc(&ptr, want * sizeof(T)));
      cap = want;
    }
    return ptr;
  }
};

/**
 * @brief Wait until a kernel has published `expected` in a pinned host word.  A feedback kernel
 * writes its report to pinned memory, fences (`__threadfence_system`) and then stores the sequence
 * number; polling that word returns a few microseconds after the kernel retires, where
 * cudaStreamSynchronize costs 10-20 us of wake-up latency per BSP level.  The stream is queried
 * every few thousand polls so an execution error still surfaces as an exception.
 */
inline void wait_for_sequence(const volatile int* word, int expected, cudaStream_t stream) {
  for (unsigned spins = 1;; ++spins) {
    if (*word == expected)
      return;
    if ((spins & 0xfff) == 0) {
      cudaError_t q = cudaStreamQuery(stream);
      if (q == cudaSuccess) {  // everything retired: the word must be there (or never will be)
        if (*word == expected)
          return;
        B2G_CHECK(cudaStreamSynchronize(stream));
        if (*word == expected)
          return;
        throw std::runtime_error("feedback sequence never arrived");
      }
      if (q != cudaErrorNotReady)
        B2G_CHECK(q);
    }
  }
}

/// Scratch shared by all operator launches on one stream.
struct workspace_t {
  cudaStream_t stream = nullptr;
  dbuf_t<ctrl_t> ctrl;                    // ring of control blocks
  dbuf_t<int> hubs;                       // hub (CTA-bin) row list
  dbuf_t<unsigned long long> hub_slabs;   // slab table of the hub rows (16 bytes per slab: advance.cuh hub_slab_t)
  dbuf_t<int> scanned;                    // degree scan for merge_path
  dbuf_t<int> tile_rows;                  // merge_path tile -> first row
  dbuf_t<unsigned long long> tile_state;  // look-back status words
  dbuf_t<unsigned> uniq_bitmap;           // exact-uniquify V-bit map (kept zeroed between uses)
  int ctrl_ring = 0;
  unsigned scan_epoch = 0;  // launch epoch of the look-back status words
  int launches = 0;         // kernels launched through this workspace (reported as gpu_launches)
  unsigned long long edges_accounted = 0;  // host-side sum of ctrl_t::edges folded in so far
  static constexpr int kCtrlRing = 256;
  /// One pinned record a kernel can publish a 64-bit value through (wait_for_sequence on `seq`).
  struct host_value_t {
    unsigned long long value;
    volatile int seq;
  };
  host_value_t* h_value = nullptr;
  int value_seq = 0;

  workspace_t() = default;
  workspace_t(const workspace_t&) = delete;
  workspace_t& operator=(const workspace_t&) = delete;
  ~workspace_t() {
    if (h_value)
      cudaFreeHost(h_value);
  }
  host_value_t* host_value() {
    if (!h_value) {
      B2G_CHECK(cudaMallocHost(&h_value, sizeof(host_value_t)));
      h_value->value = 0;
      h_value->seq = 0;
    }
    return h_value;
  }

  void init(cudaStream_t s) {
    stream = s;
    ctrl.ensure(kCtrlRing);
    B2G_CHECK(cudaMemsetAsync(ctrl.ptr, 0, sizeof(ctrl_t) * kCtrlRing, stream));
    ctrl_ring = 0;
  }
  /// A zeroed control block.  Blocks are re-zeroed in bulk when the ring wraps.
  ctrl_t* next_ctrl() {
    if (!ctrl.ptr)
      init(stream);
    if (ctrl_ring == kCtrlRing) {
      B2G_CHECK(cudaMemsetAsync(ctrl.ptr, 0, sizeof(ctrl_t) * kCtrlRing, stream));
      ctrl_ring = 0;
    }
    return ctrl.ptr + ctrl_ring++;
  }
};

}  // namespace b200
}  // namespace gunrock

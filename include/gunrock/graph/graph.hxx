/**
 * @file graph.hxx
 * @brief Non-owning, device-callable graph view `graph::graph_t` + `graph::build`
 * (include/gunrock/graph/graph.hxx:53-339, graph/csr.hxx:33-237, graph/build.hxx:29-166,
 * graph/properties.hxx:13-18).  The view holds raw pointers into the caller's csr_t (the caller
 * keeps the format object alive, graph.hxx:189-196) and is copied by value into kernels.
 *
 * One concrete view class replaces the reference's variadic-inheritance views: the CSR arrays
 * are always present; CSC arrays (the transpose, for pull traversal) are attached when built
 * from a csr+csc pair.  Accessor names and semantics are the reference's.  The reference selects
 * one of several attached views with an explicit template argument
 * (`B.template get_number_of_neighbors<graph_csc_t<...>>(col)`, algorithms/spgemm.hxx:94-190);
 * here `graph_csr_t` / `graph_csc_t` are tag types and every accessor takes the tag as a defaulted
 * template parameter (CSR when omitted, graph/graph.hxx:225-339 "first view" rule).
 */
#pragma once

#include <cassert>
#include <cmath>
#include <type_traits>

#include <gunrock/b200/runtime.cuh>
#include <gunrock/formats/formats.hxx>
#include <gunrock/memory.hxx>

namespace gunrock {
namespace graph {

using namespace memory;

struct graph_properties_t {
  bool directed{false};
  bool weighted{false};
  bool symmetric{false};
  graph_properties_t() = default;
};

template <typename vertex_t>
struct vertex_pair_t {
  vertex_t source;
  vertex_t destination;
};

/// View tags (graph/csr.hxx:33, graph/csc.hxx:30, graph/coo.hxx:30 are full classes in the reference).
template <memory_space_t space, typename vertex_t, typename edge_t, typename weight_t>
struct graph_csr_t {};
template <memory_space_t space, typename vertex_t, typename edge_t, typename weight_t>
struct graph_csc_t {};
template <memory_space_t space, typename vertex_t, typename edge_t, typename weight_t>
struct graph_coo_t {};

namespace detail {
template <typename T>
struct is_csc_view : std::false_type {};
template <memory_space_t space, typename vertex_t, typename edge_t, typename weight_t>
struct is_csc_view<graph_csc_t<space, vertex_t, edge_t, weight_t>> : std::true_type {};
template <typename T>
struct is_coo_view : std::false_type {};
template <memory_space_t space, typename vertex_t, typename edge_t, typename weight_t>
struct is_coo_view<graph_coo_t<space, vertex_t, edge_t, weight_t>> : std::true_type {};
}  // namespace detail

template <memory_space_t space, typename vertex_t, typename edge_t, typename weight_t>
class graph_t {
 public:
  using vertex_type = vertex_t;
  using edge_type = edge_t;
  using weight_type = weight_t;
  using vertex_pointer_t = vertex_t*;
  using edge_pointer_t = edge_t*;
  using weight_pointer_t = weight_t*;
  using vertex_pair_type = vertex_pair_t<vertex_t>;
  using csr_view_type = graph_csr_t<space, vertex_t, edge_t, weight_t>;
  using csc_view_type = graph_csc_t<space, vertex_t, edge_t, weight_t>;
  // the reference's names for the view tags (graph/graph.hxx:74-81)
  using graph_csr_view_t = csr_view_type;
  using graph_csc_view_t = csc_view_type;
  using graph_coo_view_t = graph_coo_t<space, vertex_t, edge_t, weight_t>;

  __host__ __device__ graph_t() {}
  __host__ __device__ graph_t(std::nullptr_t) {}

  graph_properties_t properties;

  // --- sizes ---------------------------------------------------------------------------
  template <typename view_t = csr_view_type>
  __host__ __device__ __forceinline__ vertex_t get_number_of_vertices() const {
    if constexpr (detail::is_csc_view<view_t>::value)
      return n_columns;  // one "vertex" per column of the CSC view (graph/csc.hxx:118-120)
    else
      return n_rows;
  }
  template <typename view_t = csr_view_type>
  __host__ __device__ __forceinline__ edge_t get_number_of_edges() const {
    return n_nonzeros;
  }
  __host__ __device__ __forceinline__ vertex_t get_number_of_rows() const { return n_rows; }
  __host__ __device__ __forceinline__ vertex_t get_number_of_columns() const { return n_columns; }
  __host__ __device__ __forceinline__ edge_t get_number_of_nonzeros() const { return n_nonzeros; }
  bool is_directed() { return properties.directed; }
  bool is_symmetric() { return properties.symmetric; }
  bool is_weighted() { return properties.weighted; }
  /// Which views are attached (graph/graph.hxx:169-173 `contains_representation`): the CSR always, the CSC when
  /// the graph was built with one (a run-time property here, a compile-time one in the reference), never a COO.
  template <typename view_t>
  __host__ __device__ __forceinline__ constexpr bool contains_representation() const {
    if constexpr (detail::is_csc_view<view_t>::value)
      return t_offsets != nullptr;
    else if constexpr (detail::is_coo_view<view_t>::value)
      return false;
    else
      return true;
  }
  /// Where the graph's arrays live (graph/graph.hxx:180-183).
  __host__ __device__ __forceinline__ constexpr memory_space_t memory_space() const { return space; }
  /// Number of attached views (graph/graph.hxx:158-161).
  __host__ __device__ __forceinline__ std::size_t number_of_graph_representations() const {
    return t_offsets != nullptr ? 2 : 1;
  }

  // --- accessors (graph/csr.hxx:61-178; CSC flavour graph/csc.hxx:42-100) -----------------
  // CSR view: a vertex is a row, its neighbours are the row's column indices.
  // CSC view: a vertex is a column, its "neighbours" are the column's row indices (in-edges), the
  // stored index of an edge is its SOURCE and the destination is found by searching the offsets.
  template <typename view_t = csr_view_type>
  __host__ __device__ __forceinline__ edge_t get_starting_edge(vertex_t const& v) const {
    if constexpr (detail::is_csc_view<view_t>::value)
      return t_offsets[v];
    else
      return offsets[v];
  }
  template <typename view_t = csr_view_type>
  __host__ __device__ __forceinline__ edge_t get_number_of_neighbors(vertex_t const& v) const {
    if constexpr (detail::is_csc_view<view_t>::value)
      return t_offsets[v + 1] - t_offsets[v];
    else
      return offsets[v + 1] - offsets[v];
  }
  template <typename view_t = csr_view_type>
  __host__ __device__ __forceinline__ vertex_t get_destination_vertex(edge_t const& e) const {
    if constexpr (detail::is_csc_view<view_t>::value)
      return owner_of(t_offsets, n_columns, e);
    else
      return indices[e];
  }
  template <typename view_t = csr_view_type>
  __host__ __device__ __forceinline__ weight_t get_edge_weight(edge_t const& e) const {
    if constexpr (detail::is_csc_view<view_t>::value)
      return t_values[e];
    else
      return values[e];
  }
  /// CSR: largest row r with offsets[r] <= e (binary search, O(log V); graph/csr.hxx:66-81).
  /// CSC: the stored row index (graph/csc.hxx:50-55).
  template <typename view_t = csr_view_type>
  __host__ __device__ __forceinline__ vertex_t get_source_vertex(edge_t const& e) const {
    if constexpr (detail::is_csc_view<view_t>::value)
      return t_indices[e];
    else
      return owner_of(offsets, n_rows, e);
  }
  template <typename view_t = csr_view_type>
  __host__ __device__ __forceinline__ vertex_pair_type
  get_source_and_destination_vertices(edge_t const& e) const {
    return {get_source_vertex<view_t>(e), get_destination_vertex<view_t>(e)};
  }
  /// Edge id of (source, destination) or the invalid id; rows are searched linearly unless sorted.
  __host__ __device__ __forceinline__ edge_t get_edge(vertex_t const& source,
                                                      vertex_t const& destination) const {
    for (edge_t e = offsets[source]; e < offsets[source + 1]; ++e)
      if (indices[e] == destination)
        return e;
    return static_cast<edge_t>(-1);
  }
  /// |N(source) ∩ N(destination)| for rows with ascending column order, calling on_intersection
  /// for every common neighbour (graph/csr.hxx:117-172).
  template <typename operator_t>
  __host__ __device__ __forceinline__ vertex_t get_intersection_count(vertex_t const& source,
                                                                      vertex_t const& destination,
                                                                      operator_t on_intersection) const {
    edge_t a = offsets[source], a_end = offsets[source + 1];
    edge_t b = offsets[destination], b_end = offsets[destination + 1];
    vertex_t count = 0;
    while (a < a_end && b < b_end) {
      vertex_t x = indices[a], y = indices[b];
      if (x == y) {
        on_intersection(x);
        ++count;
        ++a;
        ++b;
      } else if (x < y) {
        ++a;
      } else {
        ++b;
      }
    }
    return count;
  }
  __host__ __device__ __forceinline__ auto get_row_offsets() const { return offsets; }
  __host__ __device__ __forceinline__ auto get_column_indices() const { return indices; }
  __host__ __device__ __forceinline__ auto get_nonzero_values() const { return values; }

  // --- CSC (transpose) accessors, present when built with a csc_t ------------------------
  __host__ __device__ __forceinline__ bool has_csc() const { return t_offsets != nullptr; }
  __host__ __device__ __forceinline__ auto get_column_offsets() const { return t_offsets; }
  __host__ __device__ __forceinline__ auto get_row_indices() const { return t_indices; }
  __host__ __device__ __forceinline__ auto get_csc_values() const { return t_values; }

  // --- binding (graph/graph.hxx:211-226) --------------------------------------------------
  template <typename csr_type>
  void set(csr_type& csr) {
    n_rows = csr.number_of_rows;
    n_columns = csr.number_of_columns;
    n_nonzeros = csr.number_of_nonzeros;
    offsets = memory::raw_pointer_cast(csr.row_offsets.data());
    indices = memory::raw_pointer_cast(csr.column_indices.data());
    values = memory::raw_pointer_cast(csr.nonzero_values.data());
    uid = b200::next_graph_uid();  // a rebuilt graph is a new graph, whatever addresses its arrays got
  }
  template <typename csc_type>
  void set_csc(csc_type& csc) {
    t_offsets = memory::raw_pointer_cast(csc.column_offsets.data());
    t_indices = memory::raw_pointer_cast(csc.row_indices.data());
    t_values = memory::raw_pointer_cast(csc.nonzero_values.data());
    uid = b200::next_graph_uid();
  }

  /// The B200 kernels' view of the CSR / CSC arrays.
  /// The kernels index with int32.  Unsigned 32-bit ids / offsets (examples/algorithms/tc/tc.cu:52-54
  /// uses uint32_t) share the representation for every value a graph that fits the int32 kernels can
  /// hold, so such a view is reinterpreted, not converted; sizes beyond INT_MAX are rejected.
  b200::csr_view_t csr_view() const {
    check_kernel_types();
    b200::csr_view_t v;
    v.n_vertices = static_cast<int>(n_rows);
    v.n_edges = static_cast<int>(n_nonzeros);
    v.row_offsets = reinterpret_cast<const int*>(offsets);
    v.column_indices = reinterpret_cast<const int*>(indices);
    v.values = values;
    v.uid = uid;
    return v;
  }
  /// The CSR structure without its values, for graphs whose weight_t is not float (the reference's unit tests build
  /// `graph_t<device, int, int, int>`): the generic advance hands such a graph's typed weights to the user's
  /// operator itself (operators/advance/advance.hxx), the kernels see an unweighted graph.
  b200::csr_view_t structure_view() const {
    static_assert(std::is_integral<vertex_t>::value && sizeof(vertex_t) == 4 &&
                      std::is_integral<edge_t>::value && sizeof(edge_t) == 4,
                  "the sm_100a kernels are built for 32-bit vertex_t / edge_t");
    b200::csr_view_t v;
    v.n_vertices = static_cast<int>(n_rows);
    v.n_edges = static_cast<int>(n_nonzeros);
    v.row_offsets = reinterpret_cast<const int*>(offsets);
    v.column_indices = reinterpret_cast<const int*>(indices);
    v.values = nullptr;
    v.uid = uid;
    return v;
  }
  b200::csr_view_t csc_view() const {
    check_kernel_types();
    b200::csr_view_t v;
    v.n_vertices = static_cast<int>(n_columns);
    v.n_edges = static_cast<int>(n_nonzeros);
    v.row_offsets = reinterpret_cast<const int*>(t_offsets);
    v.column_indices = reinterpret_cast<const int*>(t_indices);
    v.values = t_values;
    v.uid = uid;
    return v;
  }

 private:
  void check_kernel_types() const {
    static_assert(std::is_integral<vertex_t>::value && sizeof(vertex_t) == 4 &&
                      std::is_integral<edge_t>::value && sizeof(edge_t) == 4 &&
                      std::is_same<weight_t, float>::value,
                  "the sm_100a kernels are built for 32-bit vertex_t / edge_t and weight_t = float "
                  "(examples/algorithms/bfs/bfs.cu:15-17)");
    assert(static_cast<unsigned long long>(n_rows) <= 0x7fffffffull &&
           static_cast<unsigned long long>(n_nonzeros) <= 0x7fffffffull);
  }
  /// Largest segment s with offs[s] <= e (e is a valid position, so offs[0] <= e < offs[n]).
  __host__ __device__ __forceinline__ static vertex_t owner_of(const edge_t* offs, vertex_t n,
                                                               edge_t const& e) {
    vertex_t lo = 0, hi = n;
    while (hi - lo > 1) {
      vertex_t mid = lo + ((hi - lo) >> 1);
      if (offs[mid] <= e)
        lo = mid;
      else
        hi = mid;
    }
    return lo;
  }
  vertex_t n_rows = 0, n_columns = 0;
  edge_t n_nonzeros = 0;
  edge_t* offsets = nullptr;
  vertex_t* indices = nullptr;
  weight_t* values = nullptr;
  edge_t* t_offsets = nullptr;
  vertex_t* t_indices = nullptr;
  weight_t* t_values = nullptr;
  /// Identity for the per-graph caches of the fused enactors (b200::graph_key_t); copied with the view.
  unsigned long long uid = 0;
};

/// graph::build<space>(properties, csr) (graph/build.hxx:29-36): pointer capture only.
template <memory_space_t space, typename edge_t, typename vertex_t, typename weight_t>
auto build(graph_properties_t properties, format::csr_t<space, vertex_t, edge_t, weight_t>& csr) {
  graph_t<space, vertex_t, edge_t, weight_t> G;
  G.properties = properties;
  G.set(csr);
  return G;
}

/// csr + csc overload (graph/build.hxx: the multi-view builder): attaches the transpose.
template <memory_space_t space, typename edge_t, typename vertex_t, typename weight_t>
auto build(graph_properties_t properties,
           format::csr_t<space, vertex_t, edge_t, weight_t>& csr,
           format::csc_t<space, vertex_t, edge_t, weight_t>& csc) {
  graph_t<space, vertex_t, edge_t, weight_t> G;
  G.properties = properties;
  G.set(csr);
  G.set_csc(csc);
  return G;
}

/// csc + csr overload, in the argument order examples/algorithms/spgemm/spgemm.cu:60 uses.
template <memory_space_t space, typename edge_t, typename vertex_t, typename weight_t>
auto build(graph_properties_t properties,
           format::csc_t<space, vertex_t, edge_t, weight_t>& csc,
           format::csr_t<space, vertex_t, edge_t, weight_t>& csr) {
  return build<space>(properties, csr, csc);
}

/// Mean out-degree (graph/graph.hxx:348-355 sums the degrees vertex by vertex; the sum IS the edge count).
template <typename graph_type>
__host__ __device__ double get_average_degree(graph_type const& G) {
  return static_cast<double>(G.get_number_of_edges()) /
         static_cast<double>(G.get_number_of_vertices());
}

/// Population standard deviation of the out-degrees (graph/graph.hxx:368-378); callable where the graph's
/// arrays are (host code for a host graph, device code for a device graph), as in the reference.
template <typename graph_type>
__host__ __device__ double get_degree_standard_deviation(graph_type const& G) {
  const double mean = get_average_degree(G);
  double accum = 0.0;
  for (typename graph_type::vertex_type v = 0; v < G.get_number_of_vertices(); ++v) {
    const double d = static_cast<double>(G.get_number_of_neighbors(v)) - mean;
    accum += d * d;
  }
  return sqrt(accum / static_cast<double>(G.get_number_of_vertices()));
}

namespace detail {
template <typename graph_type, typename histogram_t>
__global__ void degree_histogram_kernel(graph_type G, histogram_t* histogram, int bins) {
  // per-CTA counts in shared memory first: a power-law graph puts most vertices into three or four bins
  __shared__ unsigned local[72];
  for (int b = threadIdx.x; b < bins; b += blockDim.x)
    local[b] = 0;
  __syncthreads();
  const long long n = G.get_number_of_vertices();
  for (long long v = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; v < n;
       v += static_cast<long long>(gridDim.x) * blockDim.x) {
    const unsigned long long degree =
        static_cast<unsigned long long>(G.get_number_of_neighbors(static_cast<typename graph_type::vertex_type>(v)));
    const int bin = degree == 0 ? 0 : 64 - __clzll(degree);  // smallest b with degree < 2^b
    atomicAdd(&local[bin < bins ? bin : bins - 1], 1u);
  }
  __syncthreads();
  for (int b = threadIdx.x; b < bins; b += blockDim.x)
    if (local[b])
      atomicAdd(histogram + b, static_cast<histogram_t>(local[b]));
}
}  // namespace detail

/**
 * @brief Log-scale degree histogram of a DEVICE graph (graph/graph.hxx:393-440): `histogram` (device memory,
 * 8 * sizeof(vertex_t) + 1 counters) is zeroed, then bin b counts the vertices whose degree d satisfies
 * 2^(b-1) <= d < 2^b (bin 0: degree 0).  One grid-stride kernel with per-CTA shared-memory counts; asynchronous on
 * `stream` like the reference's Thrust calls.
 */
template <typename graph_type, typename histogram_t>
void build_degree_histogram(graph_type const& G, histogram_t* histogram, cudaStream_t stream = 0) {
  constexpr int bins = static_cast<int>(sizeof(typename graph_type::vertex_type) * 8 + 1);
  static_assert(bins <= 72, "histogram bins exceed the kernel's shared-memory table");
  cudaMemsetAsync(histogram, 0, sizeof(histogram_t) * bins, stream);
  const long long n = G.get_number_of_vertices();
  if (n <= 0)
    return;
  const int blocks = static_cast<int>(n / 256 + 1 < 1184 ? n / 256 + 1 : 1184);
  detail::degree_histogram_kernel<<<blocks, 256, 0, stream>>>(G, histogram, bins);
}

}  // namespace graph
}  // namespace gunrock

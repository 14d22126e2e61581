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
  /// Which views are attached (graph/graph.hxx:284-296 `contains_representation`).
  template <typename view_t>
  __host__ __device__ __forceinline__ constexpr bool contains_representation() const {
    return true;
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

template <typename graph_type>
double get_average_degree(graph_type const& G) {
  return static_cast<double>(G.get_number_of_edges()) /
         static_cast<double>(G.get_number_of_vertices());
}

}  // namespace graph
}  // namespace gunrock

/**
 * @file graph.hxx
 * @brief Non-owning, device-callable graph view `graph::graph_t` + `graph::build`
 * (include/gunrock/graph/graph.hxx:53-339, graph/csr.hxx:33-237, graph/build.hxx:29-166,
 * graph/properties.hxx:13-18).  The view holds raw pointers into the caller's csr_t (the caller
 * keeps the format object alive, graph.hxx:189-196) and is copied by value into kernels.
 *
 * One concrete view class replaces the reference's variadic-inheritance views: the CSR arrays
 * are always present; CSC arrays (the transpose, for pull traversal) are attached when built
 * from a csr+csc pair.  Accessor names and semantics are the reference's.
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

  __host__ __device__ graph_t() {}
  __host__ __device__ graph_t(std::nullptr_t) {}

  graph_properties_t properties;

  // --- sizes ---------------------------------------------------------------------------
  __host__ __device__ __forceinline__ vertex_t get_number_of_vertices() const { return n_rows; }
  __host__ __device__ __forceinline__ edge_t get_number_of_edges() const { return n_nonzeros; }
  __host__ __device__ __forceinline__ vertex_t get_number_of_rows() const { return n_rows; }
  __host__ __device__ __forceinline__ vertex_t get_number_of_columns() const { return n_columns; }
  __host__ __device__ __forceinline__ edge_t get_number_of_nonzeros() const { return n_nonzeros; }
  bool is_directed() { return properties.directed; }
  bool is_symmetric() { return properties.symmetric; }
  bool is_weighted() { return properties.weighted; }

  // --- CSR accessors (graph/csr.hxx:61-178) ---------------------------------------------
  __host__ __device__ __forceinline__ edge_t get_starting_edge(vertex_t const& v) const {
    return offsets[v];
  }
  __host__ __device__ __forceinline__ edge_t get_number_of_neighbors(vertex_t const& v) const {
    return offsets[v + 1] - offsets[v];
  }
  __host__ __device__ __forceinline__ vertex_t get_destination_vertex(edge_t const& e) const {
    return indices[e];
  }
  __host__ __device__ __forceinline__ weight_t get_edge_weight(edge_t const& e) const {
    return values[e];
  }
  /// Largest row r with offsets[r] <= e (binary search, O(log V); graph/csr.hxx:66-81).
  __host__ __device__ __forceinline__ vertex_t get_source_vertex(edge_t const& e) const {
    vertex_t lo = 0, hi = n_rows;
    while (hi - lo > 1) {
      vertex_t mid = lo + ((hi - lo) >> 1);
      if (offsets[mid] <= e)
        lo = mid;
      else
        hi = mid;
    }
    return lo;
  }
  __host__ __device__ __forceinline__ vertex_pair_type
  get_source_and_destination_vertices(edge_t const& e) const {
    return {get_source_vertex(e), get_destination_vertex(e)};
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
  }
  template <typename csc_type>
  void set_csc(csc_type& csc) {
    t_offsets = memory::raw_pointer_cast(csc.column_offsets.data());
    t_indices = memory::raw_pointer_cast(csc.row_indices.data());
    t_values = memory::raw_pointer_cast(csc.nonzero_values.data());
  }

  /// The B200 kernels' view of the CSR / CSC arrays.
  b200::csr_view_t csr_view() const {
    static_assert(std::is_same<vertex_t, int>::value && std::is_same<edge_t, int>::value &&
                      std::is_same<weight_t, float>::value,
                  "the sm_100a kernels are built for vertex_t = edge_t = int, weight_t = float "
                  "(examples/algorithms/bfs/bfs.cu:15-17)");
    b200::csr_view_t v;
    v.n_vertices = n_rows;
    v.n_edges = n_nonzeros;
    v.row_offsets = offsets;
    v.column_indices = indices;
    v.values = values;
    return v;
  }
  b200::csr_view_t csc_view() const {
    b200::csr_view_t v;
    v.n_vertices = n_columns;
    v.n_edges = n_nonzeros;
    v.row_offsets = t_offsets;
    v.column_indices = t_indices;
    v.values = t_values;
    return v;
  }

 private:
  vertex_t n_rows = 0, n_columns = 0;
  edge_t n_nonzeros = 0;
  edge_t* offsets = nullptr;
  vertex_t* indices = nullptr;
  weight_t* values = nullptr;
  edge_t* t_offsets = nullptr;
  vertex_t* t_indices = nullptr;
  weight_t* t_values = nullptr;
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

template <typename graph_type>
double get_average_degree(graph_type const& G) {
  return static_cast<double>(G.get_number_of_edges()) /
         static_cast<double>(G.get_number_of_vertices());
}

}  // namespace graph
}  // namespace gunrock

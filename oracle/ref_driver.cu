// oracle/ref_driver.cu -- TEST INFRASTRUCTURE ONLY (never shipped, never on the product path).
//
// A thin extern "C" shim around the *unmodified* reference sources, compiled from where they
// lie under /root/reference (see oracle/Makefile, target `ref`).  Nothing from the reference is
// copied into this repository; this file only #includes it.  Output goes to oracle/_ref/.
//
// What it exposes (host only, no GPU needed):
//   ref_load_mtx        -> gunrock::io::matrix_market_t::load      (include/gunrock/io/matrix_market.hxx:104-254)
//   ref_csr_from_coo    -> gunrock::format::csr_t<host>::from_coo  (include/gunrock/formats/csr.hxx:81-140)
//   ref_load_smtx       -> gunrock::io::smtx_t::load               (include/gunrock/io/smtx.hxx:56-169)
//   ref_bfs_cpu         -> bfs_cpu::run                            (examples/algorithms/bfs/bfs_cpu.hxx:20-68)
//   ref_sssp_cpu        -> sssp_cpu::run                           (examples/algorithms/sssp/sssp_cpu.hxx:22-72)
//
// It is used (a) to pin oracle/gunrock_oracle.c against the real reference and to mint the
// golden vectors under tests/golden/, and (b) as the `"kind": "reference"` CPU baseline of
// bench.py.
#include <cstdlib>
#include <cstring>
#include <string>

#include <gunrock/algorithms/algorithms.hxx>
#include <gunrock/io/smtx.hxx>

#include "bfs_cpu.hxx"
#include "sssp_cpu.hxx"

using namespace gunrock;
using namespace memory;

using vertex_t = int;
using edge_t = int;
using weight_t = float;
using host_csr_t = format::csr_t<memory_space_t::host, vertex_t, edge_t, weight_t>;
using host_coo_t = format::coo_t<memory_space_t::host, vertex_t, edge_t, weight_t>;

template <typename T>
static T* dup_out(const T* src, size_t n) {
  T* p = (T*)malloc((n ? n : 1) * sizeof(T));
  if (n)
    memcpy(p, src, n * sizeof(T));
  return p;
}

extern "C" {

// props[0]=directed props[1]=weighted props[2]=symmetric.  Arrays are malloc'ed; free with ref_free.
int ref_load_mtx(const char* filename,
                 int* n_rows,
                 int* n_cols,
                 int* nnz,
                 int** I,
                 int** J,
                 float** V,
                 int* props) {
  io::matrix_market_t<vertex_t, edge_t, weight_t> mm;
  auto [properties, coo] = mm.load(std::string(filename));
  *n_rows = coo.number_of_rows;
  *n_cols = coo.number_of_columns;
  *nnz = coo.number_of_nonzeros;
  *I = dup_out(coo.row_indices.data(), (size_t)coo.number_of_nonzeros);
  *J = dup_out(coo.column_indices.data(), (size_t)coo.number_of_nonzeros);
  *V = dup_out(coo.nonzero_values.data(), (size_t)coo.number_of_nonzeros);
  props[0] = properties.directed;
  props[1] = properties.weighted;
  props[2] = properties.symmetric;
  return 0;
}

void ref_free(void* p) {
  free(p);
}

// .smtx through the reference's own reader.  Arrays are malloc'ed (free with ref_free); the values the reference
// attaches are random draws (smtx.hxx:139-140) and are not returned.
int ref_load_smtx(const char* filename, int first_line_csv, int* n_rows, int* n_cols, int* nnz, int** row_offsets,
                  int** column_indices) {
  io::smtx_t<vertex_t, edge_t, weight_t> smtx;
  try {
    auto csr = smtx.load(std::string(filename), first_line_csv != 0);
    *n_rows = csr.number_of_rows;
    *n_cols = csr.number_of_columns;
    *nnz = csr.number_of_nonzeros;
    *row_offsets = dup_out(csr.row_offsets.data(), csr.row_offsets.size());
    *column_indices = dup_out(csr.column_indices.data(), csr.column_indices.size());
  } catch (const std::exception&) {
    return 1;
  }
  return 0;
}

// COO (host) -> CSR (host) through the reference's own from_coo.
int ref_csr_from_coo(int n_rows,
                     int n_cols,
                     int nnz,
                     const int* I,
                     const int* J,
                     const float* V,
                     int* row_offsets /* n_rows+1 */,
                     int* column_indices /* nnz */,
                     float* values /* nnz */) {
  host_coo_t coo(n_rows, n_cols, nnz);
  for (int i = 0; i < nnz; ++i) {
    coo.row_indices[i] = I[i];
    coo.column_indices[i] = J[i];
    coo.nonzero_values[i] = V[i];
  }
  host_csr_t csr;
  csr.from_coo(coo);
  memcpy(row_offsets, csr.row_offsets.data(), sizeof(int) * (size_t)(n_rows + 1));
  if (nnz) {
    memcpy(column_indices, csr.column_indices.data(), sizeof(int) * (size_t)nnz);
    memcpy(values, csr.nonzero_values.data(), sizeof(float) * (size_t)nnz);
  }
  return 0;
}

// A csr_t-shaped view that hands the validators the caller's arrays without a second copy of
// our own (the validators copy into thrust::host_vector themselves, bfs_cpu.hxx:25-27).
struct view_csr_t {
  int number_of_rows;
  int number_of_columns;
  int number_of_nonzeros;
  thrust::host_vector<int> row_offsets;
  thrust::host_vector<int> column_indices;
  thrust::host_vector<float> nonzero_values;
};

// Persistent handle so repeated timed runs do not re-copy the graph.
void* ref_graph_create(int n_rows, int nnz, const int* ro, const int* ci, const float* vals) {
  auto* g = new view_csr_t();
  g->number_of_rows = n_rows;
  g->number_of_columns = n_rows;
  g->number_of_nonzeros = nnz;
  g->row_offsets.assign(ro, ro + n_rows + 1);
  g->column_indices.assign(ci, ci + nnz);
  if (vals)
    g->nonzero_values.assign(vals, vals + nnz);
  else
    g->nonzero_values.assign((size_t)nnz, 1.0f);
  return g;
}

void ref_graph_destroy(void* h) {
  delete (view_csr_t*)h;
}

// Returns the validator's own elapsed milliseconds (std::chrono around the queue loop only).
float ref_bfs_cpu(void* h, int source, int* distances) {
  auto* g = (view_csr_t*)h;
  thrust::host_vector<int> pred(1);
  return bfs_cpu::run<view_csr_t, vertex_t, edge_t>(*g, source, distances, pred.data());
}

float ref_sssp_cpu(void* h, int source, float* distances) {
  auto* g = (view_csr_t*)h;
  thrust::host_vector<int> pred(1);
  return sssp_cpu::run<view_csr_t, vertex_t, edge_t, weight_t>(*g, source, distances,
                                                               pred.data());
}

}  // extern "C"

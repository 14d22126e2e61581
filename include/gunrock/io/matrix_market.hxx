/**
 * @file matrix_market.hxx
 * @brief `io::matrix_market_t<V,E,W>::load(file)` -> `std::tuple<graph_properties_t, coo_t<host>>`
 * (include/gunrock/io/matrix_market.hxx:99-254), with its own banner parser instead of the NIST
 * mmio C code.  Same results: coordinate files only; pattern => weight 1.0 and weighted=false;
 * real/integer => weighted=true; 1-based -> 0-based; symmetric => every off-diagonal entry is
 * followed in place by its mirror and directed=false.  Errors exit(1) with the reference's
 * messages (:108-133), overflow of vertex_t / edge_t throws (:137-142).
 *
 * Ingest speed (at scale the text parse dominates the wall clock of the example programs): the
 * entry lines are parsed by all host threads -- the body is read in one piece, cut at line ends into one chunk
 * per thread, every chunk parsed into its own arrays (hand-rolled decimal parse for the indices, strtod for the
 * values so that they round exactly as the reference's `%lf` does) and the pieces are copied, in file order,
 * into the final arrays.  That fast path only COMPLETES for a clean body (one entry per line, plain decimal
 * indices >= 1, exactly the announced number of entries or more); anything else -- entries split over lines, a
 * sign, a comment inside the body, a short file, a zero index -- abandons it and the body is read again by the
 * entry-at-a-time `fscanf` loop, which is the reference's own reading order and produces its exact results and
 * error messages.  `GUNROCK_B200_MTX_THREADS=n` sets the thread count, `=1` forces that loop.
 */
#pragma once

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <limits>
#include <memory>
#include <string>
#include <thread>
#include <tuple>
#include <vector>

#include <gunrock/io/detail/mtx_reader.hxx>
#include <gunrock/error.hxx>
#include <gunrock/formats/formats.hxx>
#include <gunrock/graph/graph.hxx>
#include <gunrock/memory.hxx>
#include <gunrock/util/filepath.hxx>

namespace gunrock {
namespace io {

using namespace memory;

enum matrix_market_format_t { coordinate, array };
enum matrix_market_data_t { real, complex, pattern, integer };
enum matrix_market_storage_scheme_t { general, hermitian, symmetric, skew };

template <typename vertex_t, typename edge_t, typename weight_t>
struct matrix_market_t {
  std::string filename;
  std::string dataset;
  matrix_market_format_t format;
  matrix_market_data_t data;
  matrix_market_storage_scheme_t scheme;

  matrix_market_t() {}
  ~matrix_market_t() {}

  std::tuple<gunrock::graph::graph_properties_t,
             format::coo_t<memory_space_t::host, vertex_t, edge_t, weight_t>>
  load(std::string _filename) {
    filename = _filename;
    dataset = util::extract_dataset(util::extract_filename(filename));

    detail::mtx_header_t header;
    thrust::host_vector<vertex_t> I, J;
    thrust::host_vector<weight_t> V;
    const detail::mtx_status_t status = detail::mtx_read<vertex_t, weight_t>(
        filename, static_cast<std::size_t>(std::numeric_limits<edge_t>::max()), header, I, J, V);
    using st = detail::mtx_status_t;
    switch (status) {
      case st::ok:
        break;
      case st::cannot_open:
        std::cerr << detail::mtx_message(status) << ": " << filename << std::endl;
        exit(1);
      case st::bad_banner:
      case st::not_sparse:
      case st::no_size_line:
      case st::bad_field:
        std::cerr << detail::mtx_message(status) << std::endl;
        exit(1);
      default:
        error::throw_if_exception(true, detail::mtx_message(status));
    }
    format = matrix_market_format_t::coordinate;

    gunrock::graph::graph_properties_t properties;
    if (header.pattern) {
      data = matrix_market_data_t::pattern;
      properties.weighted = false;
    } else {
      data = header.integer ? matrix_market_data_t::integer : matrix_market_data_t::real;
      properties.weighted = true;
    }
    if (header.symmetric) {
      scheme = matrix_market_storage_scheme_t::symmetric;
      properties.symmetric = true;
      properties.directed = false;
    } else {
      scheme = matrix_market_storage_scheme_t::general;
      properties.symmetric = false;
      properties.directed = true;
    }

    format::coo_t<memory_space_t::host, vertex_t, edge_t, weight_t> coo;
    coo.number_of_rows = static_cast<vertex_t>(header.rows);
    coo.number_of_columns = static_cast<vertex_t>(header.columns);
    coo.number_of_nonzeros = static_cast<edge_t>(I.size());
    coo.row_indices.swap(I);
    coo.column_indices.swap(J);
    coo.nonzero_values.swap(V);
    return {properties, coo};
  }
};

}  // namespace io
}  // namespace gunrock

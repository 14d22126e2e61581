/**
 * @file matrix_market.hxx
 * @brief `io::matrix_market_t<V,E,W>::load(file)` -> `std::tuple<graph_properties_t, coo_t<host>>`
 * (include/gunrock/io/matrix_market.hxx:99-254), with its own banner parser instead of the NIST
 * mmio C code.  Same results: coordinate files only; pattern => weight 1.0 and weighted=false;
 * real/integer => weighted=true; 1-based -> 0-based; symmetric => every off-diagonal entry is
 * followed in place by its mirror and directed=false.  Errors exit(1) with the reference's
 * messages (:108-133), overflow of vertex_t / edge_t throws (:137-142).
 */
#pragma once

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <limits>
#include <string>
#include <tuple>

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

    FILE* file = fopen(filename.c_str(), "r");
    if (file == NULL) {
      std::cerr << "File could not be opened: " << filename << std::endl;
      exit(1);
    }
    char line[1100];
    char banner[64], object[64], fmt[64], field[64], symmetry[64];
    bool ok = fgets(line, sizeof line, file) != NULL &&
              sscanf(line, "%63s %63s %63s %63s %63s", banner, object, fmt, field, symmetry) == 5 &&
              std::string(banner) == "%%MatrixMarket";
    auto lower = [](char* s) {
      for (; *s; ++s)
        *s = static_cast<char>(std::tolower(static_cast<unsigned char>(*s)));
    };
    if (ok) {
      lower(object);
      lower(fmt);
      lower(field);
      lower(symmetry);
      ok = std::string(object) == "matrix";
    }
    if (!ok) {
      std::cerr << "Could not process Matrix Market banner" << std::endl;
      exit(1);
    }
    if (std::string(fmt) == "array") {
      std::cerr << "File is not a sparse matrix" << std::endl;
      exit(1);
    }
    format = matrix_market_format_t::coordinate;

    std::size_t num_rows = 0, num_columns = 0, num_nonzeros = 0;
    bool have_size = false;
    while (fgets(line, sizeof line, file) != NULL) {
      if (line[0] == '%')
        continue;
      if (sscanf(line, "%zu %zu %zu", &num_rows, &num_columns, &num_nonzeros) == 3) {
        have_size = true;
        break;
      }
    }
    if (!have_size) {
      std::cerr << "Could not read file info (M, N, NNZ)" << std::endl;
      exit(1);
    }
    error::throw_if_exception(
        num_rows >= static_cast<std::size_t>(std::numeric_limits<vertex_t>::max()) ||
            num_columns >= static_cast<std::size_t>(std::numeric_limits<vertex_t>::max()),
        "vertex_t overflow");
    error::throw_if_exception(
        num_nonzeros >= static_cast<std::size_t>(std::numeric_limits<edge_t>::max()),
        "edge_t overflow");

    gunrock::graph::graph_properties_t properties;
    const std::string f(field), s(symmetry);
    const bool is_pattern = f == "pattern";
    const bool is_symmetric = s == "symmetric";
    if (is_pattern) {
      data = matrix_market_data_t::pattern;
      properties.weighted = false;
    } else if (f == "real" || f == "integer") {
      data = f == "real" ? matrix_market_data_t::real : matrix_market_data_t::integer;
      properties.weighted = true;
    } else {
      std::cerr << "Unrecognized matrix market format type" << std::endl;
      exit(1);
    }

    // Read straight into the final arrays; symmetric files reserve room for the mirrors.
    const std::size_t cap = is_symmetric ? 2 * num_nonzeros : num_nonzeros;
    thrust::host_vector<vertex_t> I(cap), J(cap);
    thrust::host_vector<weight_t> V(cap);
    std::size_t n = 0;
    for (std::size_t k = 0; k < num_nonzeros; ++k) {
      std::size_t r = 0, c = 0;
      double w = 1.0;
      int got = is_pattern ? fscanf(file, " %zu %zu \n", &r, &c)
                           : fscanf(file, " %zu %zu %lf \n", &r, &c, &w);
      error::throw_if_exception(got != (is_pattern ? 2 : 3),
                                is_pattern ? "Could not read edge from market file"
                                           : "Could not read weighted edge from market file");
      error::throw_if_exception(r == 0 || c == 0, "Market file is zero-indexed");
      const vertex_t ri = static_cast<vertex_t>(r) - 1, ci = static_cast<vertex_t>(c) - 1;
      const weight_t wv = is_pattern ? static_cast<weight_t>(1.0) : static_cast<weight_t>(w);
      I[n] = ri;
      J[n] = ci;
      V[n] = wv;
      ++n;
      if (is_symmetric && ri != ci) {
        I[n] = ci;
        J[n] = ri;
        V[n] = wv;
        ++n;
      }
    }
    fclose(file);
    I.resize(n);
    J.resize(n);
    V.resize(n);

    if (is_symmetric) {
      scheme = matrix_market_storage_scheme_t::symmetric;
      properties.symmetric = true;
      properties.directed = false;
    } else {
      scheme = matrix_market_storage_scheme_t::general;
      properties.symmetric = false;
      properties.directed = true;
    }

    format::coo_t<memory_space_t::host, vertex_t, edge_t, weight_t> coo;
    coo.number_of_rows = static_cast<vertex_t>(num_rows);
    coo.number_of_columns = static_cast<vertex_t>(num_columns);
    coo.number_of_nonzeros = static_cast<edge_t>(n);
    coo.row_indices = I;
    coo.column_indices = J;
    coo.nonzero_values = V;
    return {properties, coo};
  }
};

}  // namespace io
}  // namespace gunrock

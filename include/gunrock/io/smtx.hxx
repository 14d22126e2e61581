/**
 * @file smtx.hxx
 * @brief `io::smtx_t<vertex_t, edge_t, weight_t>::load(filename, first_line_csv)` -- reader for the `.smtx`
 * sparse-pattern text format (include/gunrock/io/smtx.hxx:56-169):
 *
 *     % comment lines (any number, anywhere before the data they precede)
 *     M K NNZ            <- rows, columns, entries (comma separated when `first_line_csv`)
 *     row_offsets        <- M + 1 integers on one line
 *     column_indices     <- NNZ integers on one line
 *
 * Returns a host `format::csr_t`.  The format carries no values; as in the reference every entry gets a weight
 * drawn uniformly from [1, 10) (smtx.hxx:139-140) -- here from ONE engine seeded per file (`weight_seed`,
 * default 1) so a load is reproducible.  Errors as in the reference: `std::runtime_error` when the file cannot be
 * opened, `std::invalid_argument` when a line's length disagrees with the header, `exception_t` on vertex_t /
 * edge_t overflow; plus checks the reference leaves out (offsets non-decreasing and ending at NNZ, column ids
 * inside [0, K)).  Data-format reader, off the traversal path.
 */
#pragma once

#include <cstdint>
#include <fstream>
#include <limits>
#include <random>
#include <sstream>
#include <stdexcept>
#include <string>

#include <gunrock/error.hxx>
#include <gunrock/formats/formats.hxx>
#include <gunrock/memory.hxx>
#include <gunrock/util/filepath.hxx>

namespace gunrock {
namespace io {

using namespace memory;

template <typename vertex_t, typename edge_t, typename weight_t>
struct smtx_t {
  std::string filename;
  std::string dataset;
  unsigned weight_seed = 1;

  smtx_t() {}
  ~smtx_t() {}

  auto load(std::string _filename, bool first_line_csv = false) {
    filename = _filename;
    dataset = util::extract_dataset(util::extract_filename(filename));
    std::ifstream file(filename);
    if (!file.is_open())
      throw std::runtime_error("Unable to open file");

    // the three data lines, comments skipped
    std::string lines[3];
    for (auto& line : lines) {
      do {
        if (!std::getline(file, line))
          throw std::invalid_argument(filename + ": fewer than three data lines");
      } while (!line.empty() && line[0] == '%');
    }

    // header: M K NNZ
    std::string header = lines[0];
    if (first_line_csv)
      for (auto& c : header)
        if (c == ',')
          c = ' ';
    long long rows = -1, columns = -1, nonzeros = -1;
    {
      std::istringstream in(header);
      in >> rows >> columns >> nonzeros;
      if (!in || rows < 0 || columns < 0 || nonzeros < 0)
        throw std::invalid_argument(filename + ": the first data line must hold `rows columns nonzeros`");
    }
    error::throw_if_exception(rows >= static_cast<long long>(std::numeric_limits<vertex_t>::max()) ||
                                  columns >= static_cast<long long>(std::numeric_limits<vertex_t>::max()),
                              "vertex_t overflow");
    error::throw_if_exception(nonzeros >= static_cast<long long>(std::numeric_limits<edge_t>::max()),
                              "edge_t overflow");

    format::csr_t<memory_space_t::host, vertex_t, edge_t, weight_t> csr;
    csr.number_of_rows = static_cast<vertex_t>(rows);
    csr.number_of_columns = static_cast<vertex_t>(columns);
    csr.number_of_nonzeros = static_cast<edge_t>(nonzeros);
    csr.row_offsets.reserve(static_cast<std::size_t>(rows) + 1);
    csr.column_indices.reserve(static_cast<std::size_t>(nonzeros));

    auto mismatch = [&](const char* what, std::size_t got, long long want) {
      std::ostringstream ss;
      ss << "Number of " << what << " in " << filename << " (" << got
         << ") does not match the count in the first line (" << want << ")";
      return std::invalid_argument(ss.str());
    };
    {
      std::istringstream in(lines[1]);
      long long x, previous = 0;
      while (in >> x) {
        if (x < previous || x > nonzeros)
          throw std::invalid_argument(filename + ": row offsets must be non-decreasing and end at the entry count");
        previous = x;
        csr.row_offsets.push_back(static_cast<edge_t>(x));
      }
    }
    if (csr.row_offsets.size() != static_cast<std::size_t>(rows) + 1)
      throw mismatch("rows", csr.row_offsets.size() ? csr.row_offsets.size() - 1 : 0, rows);
    if (static_cast<long long>(csr.row_offsets[static_cast<std::size_t>(rows)]) != nonzeros ||
        csr.row_offsets[0] != 0)
      throw std::invalid_argument(filename + ": row offsets must start at 0 and end at the entry count");
    {
      std::istringstream in(lines[2]);
      long long x;
      while (in >> x) {
        if (x < 0 || x >= columns)
          throw std::invalid_argument(filename + ": column index outside [0, columns)");
        csr.column_indices.push_back(static_cast<vertex_t>(x));
      }
    }
    if (csr.column_indices.size() != static_cast<std::size_t>(nonzeros))
      throw mismatch("non-zeros", csr.column_indices.size(), nonzeros);

    std::mt19937 engine(weight_seed);
    std::uniform_real_distribution<double> weight(1.0, 10.0);
    csr.nonzero_values.resize(static_cast<std::size_t>(nonzeros));
    for (auto& w : csr.nonzero_values)
      w = static_cast<weight_t>(weight(engine));
    return csr;
  }
};

}  // namespace io
}  // namespace gunrock

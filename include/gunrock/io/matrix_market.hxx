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

#include <unistd.h>

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

namespace detail {

/// One thread's share of the entry lines of a coordinate file, parsed into its own arrays.
template <typename vertex_t, typename weight_t>
struct mtx_chunk_t {
  std::vector<vertex_t> rows, columns;
  std::vector<weight_t> values;
  std::size_t entries = 0;  // file entries (mirrors not counted)
  bool clean = true;        // false: something only the entry-at-a-time loop may judge
};

inline bool mtx_is_blank(char c) {
  return c == ' ' || c == '\t' || c == '\r' || c == '\v' || c == '\f';
}

/// Parse the lines of [p, end): "<row> <column>[ <value>]" each, blank lines allowed.  `end` is a line end (or
/// the end of the buffer, which is NUL-terminated so that strtod cannot run past it).
template <typename vertex_t, typename weight_t>
void mtx_parse_chunk(const char* p, const char* end, bool pattern, bool symmetric, std::size_t index_limit,
                     mtx_chunk_t<vertex_t, weight_t>& out) {
  {  // an entry line has >= 4 bytes ("1 1\n"); typical files ~12-30: reserve once, untouched pages cost nothing
    const std::size_t guess = static_cast<std::size_t>(end - p) / 8 * (symmetric ? 2 : 1) + 16;
    out.rows.reserve(guess);
    out.columns.reserve(guess);
    out.values.reserve(guess);
  }
  auto index = [&](std::size_t& value) -> bool {  // plain decimal digits, at least one, no overflow past the limit
    if (p >= end || *p < '0' || *p > '9')
      return false;
    std::size_t v = 0;
    while (p < end && *p >= '0' && *p <= '9') {
      v = v * 10 + static_cast<std::size_t>(*p - '0');
      if (v > index_limit)
        return false;
      ++p;
    }
    value = v;
    return p >= end || mtx_is_blank(*p) || *p == '\n';
  };
  while (p < end) {
    while (p < end && mtx_is_blank(*p))
      ++p;
    if (p >= end)
      break;
    if (*p == '\n') {  // blank line
      ++p;
      continue;
    }
    std::size_t r = 0, c = 0;
    double w = 1.0;
    if (!index(r)) {
      out.clean = false;
      return;
    }
    while (p < end && mtx_is_blank(*p))
      ++p;
    if (!index(c) || r == 0 || c == 0) {
      out.clean = false;
      return;
    }
    while (p < end && mtx_is_blank(*p))
      ++p;
    if (!pattern) {
      if (p >= end || *p == '\n') {
        out.clean = false;
        return;
      }
      char* after = nullptr;
      w = std::strtod(p, &after);
      if (after == p || after > end || !(after == end || mtx_is_blank(*after) || *after == '\n')) {
        out.clean = false;
        return;
      }
      p = after;
      while (p < end && mtx_is_blank(*p))
        ++p;
    }
    if (p < end && *p != '\n') {  // more tokens on the line than an entry has
      out.clean = false;
      return;
    }
    if (p < end)
      ++p;
    const vertex_t ri = static_cast<vertex_t>(r) - 1, ci = static_cast<vertex_t>(c) - 1;
    const weight_t wv = pattern ? static_cast<weight_t>(1.0) : static_cast<weight_t>(w);
    out.rows.push_back(ri);
    out.columns.push_back(ci);
    out.values.push_back(wv);
    if (symmetric && ri != ci) {
      out.rows.push_back(ci);
      out.columns.push_back(ri);
      out.values.push_back(wv);
    }
    ++out.entries;
  }
}

inline unsigned mtx_threads(std::size_t body_bytes) {
  if (const char* env = std::getenv("GUNROCK_B200_MTX_THREADS")) {  // explicit: that many chunks, whatever the size
    const std::size_t t = static_cast<std::size_t>(std::max(1, std::min(64, std::atoi(env))));
    return static_cast<unsigned>(std::min<std::size_t>(t, body_bytes / 16 + 1));
  }
  // past a few dozen threads the copy into the final arrays is the limit; >= 1 MiB of text per thread
  const std::size_t t = std::max(1u, std::min(std::thread::hardware_concurrency(), 64u));
  return static_cast<unsigned>(std::min<std::size_t>(t, body_bytes / (std::size_t(1) << 20) + 1));
}

}  // namespace detail

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

    // ---- fast path: all host threads over a clean body (see the file comment) ----------------------------
    const long body_at = ftell(file);
    thrust::host_vector<vertex_t> I, J;
    thrust::host_vector<weight_t> V;
    std::size_t n = 0;
    bool parsed = false;
    {
      const char* env = std::getenv("GUNROCK_B200_MTX_THREADS");
      const bool entry_loop_only = env != nullptr && std::atoi(env) <= 1;
      long file_end = -1;
      if (!entry_loop_only && body_at >= 0 && fseek(file, 0, SEEK_END) == 0 &&
          (file_end = ftell(file)) >= body_at && fseek(file, body_at, SEEK_SET) == 0) {
        const std::size_t bytes = static_cast<std::size_t>(file_end - body_at);
        const unsigned threads = detail::mtx_threads(bytes);
        // the body in one piece, every thread reading its own byte range (pread: no shared file position)
        std::unique_ptr<char[]> text(new char[bytes + 1]);
        text[bytes] = '\0';  // strtod stops here at the latest
        std::vector<char> read_ok(threads, 1);
        auto read_range = [&](unsigned t) {
          std::size_t at = bytes / threads * t;
          const std::size_t stop = t + 1 == threads ? bytes : bytes / threads * (t + 1);
          while (at < stop) {
            const ssize_t k = pread(fileno(file), text.get() + at, stop - at, static_cast<off_t>(body_at) + at);
            if (k <= 0) {
              read_ok[t] = 0;
              return;
            }
            at += static_cast<std::size_t>(k);
          }
        };
        {
          std::vector<std::thread> readers;
          for (unsigned t = 1; t < threads; ++t)
            readers.emplace_back(read_range, t);
          read_range(0);
          for (auto& th : readers)
            th.join();
        }
        const std::size_t got = std::count(read_ok.begin(), read_ok.end(), 1) == threads ? bytes : 0;
        if (got == bytes) {
          const char* base = text.get();
          std::vector<std::size_t> cut(threads + 1, bytes);
          cut[0] = 0;
          for (unsigned t = 1; t < threads; ++t) {  // a chunk starts right after a line end
            std::size_t at = std::max(cut[t - 1], bytes / threads * t);
            while (at < bytes && base[at] != '\n')
              ++at;
            cut[t] = at < bytes ? at + 1 : bytes;
          }
          std::vector<detail::mtx_chunk_t<vertex_t, weight_t>> chunks(threads);
          const std::size_t limit = static_cast<std::size_t>(std::numeric_limits<vertex_t>::max());
          std::vector<std::thread> team;
          for (unsigned t = 1; t < threads; ++t)
            team.emplace_back([&, t] {
              detail::mtx_parse_chunk<vertex_t, weight_t>(base + cut[t], base + cut[t + 1], is_pattern,
                                                           is_symmetric, limit, chunks[t]);
            });
          detail::mtx_parse_chunk<vertex_t, weight_t>(base + cut[0], base + cut[1], is_pattern, is_symmetric,
                                                       limit, chunks[0]);
          for (auto& th : team)
            th.join();
          std::size_t entries = 0;
          bool clean = true;
          for (auto& c : chunks) {
            clean = clean && c.clean;
            entries += c.entries;
          }
          if (clean && entries == num_nonzeros) {  // anything else is the entry-at-a-time loop's to judge
            std::vector<std::size_t> at(threads + 1, 0);
            for (unsigned t = 0; t < threads; ++t)
              at[t + 1] = at[t] + chunks[t].rows.size();
            n = at[threads];
            I.resize(n);
            J.resize(n);
            V.resize(n);
            auto place = [&](unsigned t) {
              std::copy(chunks[t].rows.begin(), chunks[t].rows.end(), I.begin() + at[t]);
              std::copy(chunks[t].columns.begin(), chunks[t].columns.end(), J.begin() + at[t]);
              std::copy(chunks[t].values.begin(), chunks[t].values.end(), V.begin() + at[t]);
            };
            team.clear();
            for (unsigned t = 1; t < threads; ++t)
              team.emplace_back(place, t);
            place(0);
            for (auto& th : team)
              th.join();
            parsed = true;
          }
        }
      }
      if (!parsed && body_at >= 0)
        fseek(file, body_at, SEEK_SET);
    }

    // ---- entry at a time, the reference's reading order (also the judge of every irregular body) -----------
    // Read straight into the final arrays; symmetric files reserve room for the mirrors.
    const std::size_t cap = parsed ? 0 : (is_symmetric ? 2 * num_nonzeros : num_nonzeros);
    if (!parsed) {
      I.resize(cap);
      J.resize(cap);
      V.resize(cap);
    }
    for (std::size_t k = 0; !parsed && k < num_nonzeros; ++k) {
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
    coo.row_indices.swap(I);
    coo.column_indices.swap(J);
    coo.nonzero_values.swap(V);
    return {properties, coo};
  }
};

}  // namespace io
}  // namespace gunrock

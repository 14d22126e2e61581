/**
 * @file mtx_reader.hxx
 * @brief The MatrixMarket coordinate reader behind `io::matrix_market_t::load` and the C ABI's `b2g_mtx_load`
 * (standard library only; stands where the reference keeps the NIST mmio C code, include/gunrock/io/detail/mmio.hxx).
 * `mtx_read` returns a status instead of exiting / throwing, so that a shared library can report a bad file to its
 * caller; include/gunrock/io/matrix_market.hxx maps each status to the reference's behaviour and describes the
 * two reading paths (all host threads over a clean body, entry-at-a-time otherwise).
 */
#pragma once

#include <algorithm>
#include <cctype>
#include <cstdio>
#include <cstdlib>
#include <limits>
#include <memory>
#include <string>
#include <thread>
#include <vector>

#include <unistd.h>

namespace gunrock {
namespace io {

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

namespace detail {

/// Outcome of `mtx_read`; `matrix_market_t::load` turns each into the reference's behaviour (exit(1) with its
/// message for the first five, an exception for the rest), the C ABI into an error code + message.
enum class mtx_status_t {
  ok,
  cannot_open,         // "File could not be opened: <name>"                 exit(1)  matrix_market.hxx:108-111
  bad_banner,          // "Could not process Matrix Market banner"           exit(1)  :113-117
  not_sparse,          // "File is not a sparse matrix"                      exit(1)  :126-129
  no_size_line,        // "Could not read file info (M, N, NNZ)"             exit(1)  :119-124
  bad_field,           // "Unrecognized matrix market format type"           exit(1)  :217-220
  vertex_overflow,     // "vertex_t overflow"                                throws   :137-139
  edge_overflow,       // "edge_t overflow"                                  throws   :140-142
  bad_entry,           // "Could not read edge from market file"             throws   :174-176
  bad_weighted_entry,  // "Could not read weighted edge from market file"    throws   :188-190
  zero_indexed         // "Market file is zero-indexed"                      throws   :177-179
};

inline const char* mtx_message(mtx_status_t s) {
  switch (s) {
    case mtx_status_t::ok: return "ok";
    case mtx_status_t::cannot_open: return "File could not be opened";
    case mtx_status_t::bad_banner: return "Could not process Matrix Market banner";
    case mtx_status_t::not_sparse: return "File is not a sparse matrix";
    case mtx_status_t::no_size_line: return "Could not read file info (M, N, NNZ)";
    case mtx_status_t::bad_field: return "Unrecognized matrix market format type";
    case mtx_status_t::vertex_overflow: return "vertex_t overflow";
    case mtx_status_t::edge_overflow: return "edge_t overflow";
    case mtx_status_t::bad_entry: return "Could not read edge from market file";
    case mtx_status_t::bad_weighted_entry: return "Could not read weighted edge from market file";
    case mtx_status_t::zero_indexed: return "Market file is zero-indexed";
  }
  return "?";
}

struct mtx_header_t {
  std::size_t rows = 0, columns = 0, announced = 0;  // the size line
  bool pattern = false, integer = false, symmetric = false;
};

/**
 * @brief Read a coordinate MatrixMarket file into (I, J, V) -- any vector type with resize / data / begin.
 * Never exits and never throws on file content: the caller maps the status.  `edge_limit` is the largest
 * entry count edge_t can hold.
 */
template <typename vertex_t, typename weight_t, typename index_vector_t, typename value_vector_t>
mtx_status_t mtx_read(const std::string& filename, std::size_t edge_limit, mtx_header_t& header,
                      index_vector_t& I, index_vector_t& J, value_vector_t& V) {
  FILE* file = fopen(filename.c_str(), "r");
  if (file == NULL)
    return mtx_status_t::cannot_open;
  struct closer_t {
    FILE* f;
    ~closer_t() { fclose(f); }
  } closer{file};

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
  if (!ok)
    return mtx_status_t::bad_banner;
  if (std::string(fmt) == "array")
    return mtx_status_t::not_sparse;

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
  if (!have_size)
    return mtx_status_t::no_size_line;
  if (num_rows >= static_cast<std::size_t>(std::numeric_limits<vertex_t>::max()) ||
      num_columns >= static_cast<std::size_t>(std::numeric_limits<vertex_t>::max()))
    return mtx_status_t::vertex_overflow;
  if (num_nonzeros >= edge_limit)
    return mtx_status_t::edge_overflow;

  const std::string f(field), s(symmetry);
  const bool is_pattern = f == "pattern";
  const bool is_symmetric = s == "symmetric";
  if (!is_pattern && f != "real" && f != "integer")
    return mtx_status_t::bad_field;
  header.rows = num_rows;
  header.columns = num_columns;
  header.announced = num_nonzeros;
  header.pattern = is_pattern;
  header.integer = f == "integer";
  header.symmetric = is_symmetric;

  // ---- fast path: all host threads over a clean body (see the file comment) ------------------------------
  const long body_at = ftell(file);
  std::size_t n = 0;
  bool parsed = false;
  {
    const char* env = std::getenv("GUNROCK_B200_MTX_THREADS");
    const bool entry_loop_only = env != nullptr && std::atoi(env) <= 1;
    long file_end = -1;
    if (!entry_loop_only && body_at >= 0 && fseek(file, 0, SEEK_END) == 0 &&
        (file_end = ftell(file)) >= body_at && fseek(file, body_at, SEEK_SET) == 0) {
      const std::size_t bytes = static_cast<std::size_t>(file_end - body_at);
      const unsigned threads = mtx_threads(bytes);
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
      if (std::count(read_ok.begin(), read_ok.end(), 1) == static_cast<long>(threads)) {
        const char* base = text.get();
        std::vector<std::size_t> cut(threads + 1, bytes);
        cut[0] = 0;
        for (unsigned t = 1; t < threads; ++t) {  // a chunk starts right after a line end
          std::size_t at = std::max(cut[t - 1], bytes / threads * t);
          while (at < bytes && base[at] != '\n')
            ++at;
          cut[t] = at < bytes ? at + 1 : bytes;
        }
        std::vector<mtx_chunk_t<vertex_t, weight_t>> chunks(threads);
        const std::size_t limit = static_cast<std::size_t>(std::numeric_limits<vertex_t>::max());
        std::vector<std::thread> team;
        for (unsigned t = 1; t < threads; ++t)
          team.emplace_back([&, t] {
            mtx_parse_chunk<vertex_t, weight_t>(base + cut[t], base + cut[t + 1], is_pattern, is_symmetric, limit,
                                                chunks[t]);
          });
        mtx_parse_chunk<vertex_t, weight_t>(base + cut[0], base + cut[1], is_pattern, is_symmetric, limit,
                                            chunks[0]);
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
  if (parsed)
    return mtx_status_t::ok;

  // ---- entry at a time, the reference's reading order (also the judge of every irregular body) -------------
  // Read straight into the final arrays; symmetric files reserve room for the mirrors.
  const std::size_t cap = is_symmetric ? 2 * num_nonzeros : num_nonzeros;
  I.resize(cap);
  J.resize(cap);
  V.resize(cap);
  for (std::size_t k = 0; k < num_nonzeros; ++k) {
    std::size_t r = 0, c = 0;
    double w = 1.0;
    int got = is_pattern ? fscanf(file, " %zu %zu \n", &r, &c) : fscanf(file, " %zu %zu %lf \n", &r, &c, &w);
    if (got != (is_pattern ? 2 : 3))
      return is_pattern ? mtx_status_t::bad_entry : mtx_status_t::bad_weighted_entry;
    if (r == 0 || c == 0)
      return mtx_status_t::zero_indexed;
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
  I.resize(n);
  J.resize(n);
  V.resize(n);
  return mtx_status_t::ok;
}

}  // namespace detail

}  // namespace io
}  // namespace gunrock

/**
 * @file filepath.hxx
 * @brief File-name helpers used by the CLI and the loaders (include/gunrock/util/filepath.hxx).
 */
#pragma once

#include <string>

namespace gunrock {
namespace util {

inline std::string extract_filename(std::string path, std::string delim = "/") {
  std::size_t pos = path.rfind(delim);
  return pos == std::string::npos ? path : path.substr(pos + delim.size());
}
inline std::string extract_dataset(std::string filename) {
  std::size_t pos = filename.rfind('.');
  return pos == std::string::npos ? filename : filename.substr(0, pos);
}
inline bool has_extension(const std::string& filename, const std::string& ext) {
  return filename.size() >= ext.size() &&
         filename.compare(filename.size() - ext.size(), ext.size(), ext) == 0;
}
inline bool is_market(std::string filename) {
  return has_extension(filename, ".mtx") || has_extension(filename, ".mmio");
}
inline bool is_binary_csr(std::string filename) {
  return has_extension(filename, ".csr");
}

}  // namespace util
}  // namespace gunrock

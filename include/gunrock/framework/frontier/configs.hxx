/** @file configs.hxx  Frontier enums (include/gunrock/framework/frontier/configs.hxx:19-35). */
#pragma once

namespace gunrock {
namespace frontier {

enum frontier_view_t {
  vector,  ///< dense list of ids
  bitmap,  ///< one bit per vertex
  boolmap  ///< one byte per vertex
};

enum frontier_kind_t {
  vertex_frontier,
  edge_frontier,
  vertex_edge_frontier
};

}  // namespace frontier
}  // namespace gunrock

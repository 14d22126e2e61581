/**
 * @file vector_frontier.hxx
 * @brief `frontier::vector_frontier_t<vertex_t, edge_t, kind>` -- the reference's name for the vector view of a
 * frontier (include/gunrock/framework/frontier/vector_frontier.hxx:27-311), for code that includes this header
 * or names the type directly.  Here the vector view IS `frontier_t<vertex_t, edge_t, kind, vector>` (device-resident
 * count, lazy host read-back: framework/frontier/frontier.hxx), so the name is an alias rather than a base class.
 */
#pragma once

#include <gunrock/framework/frontier/frontier.hxx>

namespace gunrock {
namespace frontier {

template <typename vertex_t, typename edge_t, frontier_kind_t _kind = frontier_kind_t::vertex_frontier>
using vector_frontier_t = frontier_t<vertex_t, edge_t, _kind, frontier_view_t::vector>;

}  // namespace frontier
}  // namespace gunrock

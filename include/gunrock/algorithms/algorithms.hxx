/**
 * @file algorithms.hxx
 * @brief `gunrock::options_t` + the common include set every algorithm header pulls in
 * (include/gunrock/algorithms/algorithms.hxx:27-72 and :76-103).
 */
#pragma once

#include <gunrock/framework/operators/configs.hxx>

namespace gunrock {

/// Runtime operator selection shared by all algorithms; defaults as the reference
/// (block_mapped advance, predicated filter, filter and uniquify off, best-effort uniquify).
struct options_t {
  using lb_t = operators::load_balance_t;
  using filter_t = operators::filter_algorithm_t;
  using uniq_t = operators::uniquify_algorithm_t;

  lb_t advance_load_balance = lb_t::block_mapped;
  filter_t filter_algorithm = filter_t::predicated;
  bool enable_filter = false;
  bool enable_uniquify = false;
  uniq_t uniquify_algorithm = uniq_t::unique;
  bool best_effort_uniquify = true;
  float uniquify_percent = 100.0f;
  /// B200 addition (the reference's advance_direction_t is a dead template parameter, SURVEY.md
  /// F5): traversal direction for the fused BFS enactor.  forward keeps reference behaviour.
  operators::advance_direction_t advance_direction = operators::advance_direction_t::forward;

  options_t() = default;

  /// Positional form used by the reference's callers: (load balance, filter, filter on/off,
  /// uniquify on/off, uniquify algorithm, best effort, percent) -- trailing ones optional.
  options_t(lb_t lb, filter_t filter = filter_t::predicated, bool with_filter = false,
            bool with_uniquify = false, uniq_t uniquify = uniq_t::unique, bool best_effort = true,
            float percent = 100.0f) {
    advance_load_balance = lb;
    filter_algorithm = filter;
    enable_filter = with_filter;
    enable_uniquify = with_uniquify;
    uniquify_algorithm = uniquify;
    best_effort_uniquify = best_effort;
    uniquify_percent = percent;
  }
};

}  // namespace gunrock

#include <gunrock/memory.hxx>
#include <gunrock/error.hxx>

#include <gunrock/cuda/cuda.hxx>

#include <gunrock/util/math.hxx>
#include <gunrock/util/load_store.hxx>
#include <gunrock/util/type_limits.hxx>
#include <gunrock/util/print.hxx>
#include <gunrock/util/compare.hxx>
#include <gunrock/util/timer.hxx>

#include <gunrock/container/array.hxx>
#include <gunrock/container/vector.hxx>

#include <gunrock/formats/formats.hxx>
#include <gunrock/graph/graph.hxx>

#include <gunrock/framework/framework.hxx>

#include <gunrock/io/matrix_market.hxx>
#include <gunrock/io/sample.hxx>

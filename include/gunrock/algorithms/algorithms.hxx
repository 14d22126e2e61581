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
  operators::load_balance_t advance_load_balance = operators::load_balance_t::block_mapped;
  operators::filter_algorithm_t filter_algorithm = operators::filter_algorithm_t::predicated;
  bool enable_filter = false;
  bool enable_uniquify = false;
  operators::uniquify_algorithm_t uniquify_algorithm = operators::uniquify_algorithm_t::unique;
  bool best_effort_uniquify = true;
  float uniquify_percent = 100.0f;
  /// B200 addition (the reference's advance_direction_t is a dead template parameter, SURVEY.md
  /// F5): traversal direction for the fused BFS enactor.  forward keeps reference behaviour.
  operators::advance_direction_t advance_direction = operators::advance_direction_t::forward;

  options_t() = default;
  options_t(operators::load_balance_t _advance_load_balance,
            operators::filter_algorithm_t _filter_algorithm =
                operators::filter_algorithm_t::predicated,
            bool _enable_filter = false,
            bool _enable_uniquify = false,
            operators::uniquify_algorithm_t _uniquify_algorithm =
                operators::uniquify_algorithm_t::unique,
            bool _best_effort_uniquify = true,
            float _uniquify_percent = 100.0f)
      : advance_load_balance(_advance_load_balance),
        filter_algorithm(_filter_algorithm),
        enable_filter(_enable_filter),
        enable_uniquify(_enable_uniquify),
        uniquify_algorithm(_uniquify_algorithm),
        best_effort_uniquify(_best_effort_uniquify),
        uniquify_percent(_uniquify_percent) {}
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

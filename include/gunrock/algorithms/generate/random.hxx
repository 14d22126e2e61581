/**
 * @file random.hxx
 * @brief `generate::random::uniform_distribution(vector, begin, end)` and `get_random`
 * (include/gunrock/algorithms/generate/random.hxx:19-49).  Element i is the (i+1)-th draw of a
 * default-seeded thrust::default_random_engine, so values agree with the reference.  Setup utility,
 * not on the traversal path.
 */
#pragma once

#include <random>

#include <thrust/iterator/counting_iterator.h>
#include <thrust/random.h>
#include <thrust/transform.h>

namespace gunrock {
namespace generate {
namespace random {

template <typename type_t>
struct nth_uniform_t {
  type_t lo, hi;
  __host__ __device__ type_t operator()(int i) const {
    thrust::default_random_engine engine;
    thrust::uniform_real_distribution<type_t> dist(lo, hi);
    engine.discard(i);
    return dist(engine);
  }
};

template <typename vector_t, typename type_t = typename vector_t::value_type>
void uniform_distribution(vector_t& input, type_t begin = 0.0f, type_t end = 1.0f) {
  thrust::transform(thrust::make_counting_iterator<int>(0),
                    thrust::make_counting_iterator<int>(static_cast<int>(input.size())),
                    input.begin(), nth_uniform_t<type_t>{begin, end});
}

template <typename rand_t = float>
rand_t get_random(rand_t begin = 0.0f, rand_t end = 1.0f) {
  std::random_device device;
  std::mt19937 engine(device());
  std::uniform_real_distribution<> dist(begin, end);
  return static_cast<rand_t>(dist(engine));
}

}  // namespace random
}  // namespace generate
}  // namespace gunrock

// rng_utils.hpp -- host random number generator with three streams (pool / state / local).
// Interface of reference include/rng_utils.hpp:31-131 and the non-MKL behaviour of
// src/rng_utils.cpp:87-232 (std::mt19937 seeded seed+0 / seed+1+state_id / seed+1+num_states+pool_rank,
// uniform_real_distribution), so that seeded programs draw the same numbers.  Host only.
#ifndef RNG_UTILS_HPP
#define RNG_UTILS_HPP

#include <cassert>
#include <cmath>
#include <random>
#include <string>
#include <vector>

#include "mpi_env.hpp"

namespace iqs {

template <typename Type>
class RandomNumberGenerator {
 private:
  std::size_t _seed = 0;
  std::size_t _num_generated_or_skipped_local_numbers = 0;
  std::size_t _num_generated_or_skipped_state_numbers = 0;
  std::size_t _num_generated_or_skipped_pool_numbers = 0;
  std::mt19937 _pool_generator, _state_generator, _local_generator;
  std::uniform_real_distribution<Type> u_distribution = std::uniform_real_distribution<Type>(0.0, 1.0);
  std::normal_distribution<Type> n_distribution = std::normal_distribution<Type>(0.0, 1.0);

 public:
  RandomNumberGenerator() {}
  ~RandomNumberGenerator() {}
  // copy of a generator: same seed, fast-forwarded to the same point of every stream
  RandomNumberGenerator(RandomNumberGenerator *source_rng);

  std::size_t GetSeed() { return _seed; }
  std::size_t GetNumGeneratedOrSkippedLocalNumbers() { return _num_generated_or_skipped_local_numbers; }
  std::size_t GetNumGeneratedOrSkippedStateNumbers() { return _num_generated_or_skipped_state_numbers; }
  std::size_t GetNumGeneratedOrSkippedPoolNumbers() { return _num_generated_or_skipped_pool_numbers; }

  void SetSeedStreamPtrs(std::size_t RNG_seed);
  void SkipAhead(std::size_t num_skip, std::string shared = "local");
  void UniformRandomNumbers(Type *value, std::size_t size = 1UL, Type a = 0., Type b = 1., std::string shared = "local");
  void GaussianRandomNumbers(Type *value, std::size_t size = 1UL, std::string shared = "local");
  void RandomIntegersInRange(int *value, std::size_t size = 1UL, int a = 0, int b = 2, std::string shared = "local");

 private:
  std::mt19937 *SelectGeneratorAndUpdateCounter(std::size_t size, std::string shared);
};

// Fisher-Yates shuffle driven by the generator above
template <typename Type, typename TypeFloat>
void ShuffleFisherYates(std::vector<Type> &array, RandomNumberGenerator<TypeFloat> *rnd_generator_ptr, std::string shared = "local");

}  // namespace iqs
#endif

// rng.cpp -- host random number generator (reference behaviour: src/rng_utils.cpp:87-232, non-MKL).
#include "../include/rng_utils.hpp"

#include <iostream>

namespace iqs {

template <typename Type>
RandomNumberGenerator<Type>::RandomNumberGenerator(RandomNumberGenerator *source_rng) {
  SetSeedStreamPtrs(source_rng->GetSeed());
  SkipAhead(source_rng->GetNumGeneratedOrSkippedPoolNumbers(), "pool");
  SkipAhead(source_rng->GetNumGeneratedOrSkippedStateNumbers(), "state");
  SkipAhead(source_rng->GetNumGeneratedOrSkippedLocalNumbers(), "local");
}

template <typename Type>
std::mt19937 *RandomNumberGenerator<Type>::SelectGeneratorAndUpdateCounter(std::size_t size, std::string shared) {
  if (shared == "local") {
    _num_generated_or_skipped_local_numbers += size;
    return &_local_generator;
  }
  if (shared == "state") {
    _num_generated_or_skipped_state_numbers += size;
    return &_state_generator;
  }
  if (shared == "pool") {
    _num_generated_or_skipped_pool_numbers += size;
    return &_pool_generator;
  }
  assert(0 && "stream must be 'local', 'state' or 'pool'");
  return nullptr;
}

template <typename Type>
void RandomNumberGenerator<Type>::SetSeedStreamPtrs(std::size_t RNG_seed) {
  _seed = RNG_seed;
  _num_generated_or_skipped_local_numbers = 0;
  _num_generated_or_skipped_state_numbers = 0;
  _num_generated_or_skipped_pool_numbers = 0;
  int num_states = mpi::Environment::GetNumStates();
  int state_id = mpi::Environment::GetStateId();
  int pool_rank = mpi::Environment::GetPoolRank();
  _pool_generator.seed(RNG_seed + 0);
  _state_generator.seed(RNG_seed + 1 + state_id);
  _local_generator.seed(RNG_seed + 1 + num_states + pool_rank);
}

// Draw and drop: the number of engine calls per double is implementation defined, so skipping is
// done by generating (as the reference does) -- but through a bounded buffer, not a stack VLA
// (the reference's VLA overflows the stack at ~2^20 numbers, src/rng_utils.cpp:110).
template <typename Type>
void RandomNumberGenerator<Type>::SkipAhead(std::size_t num_skip, std::string shared) {
  Type buf[1024];
  while (num_skip) {
    std::size_t k = num_skip < 1024 ? num_skip : 1024;
    UniformRandomNumbers(buf, k, 0., 1., shared);
    num_skip -= k;
  }
}

template <typename Type>
void RandomNumberGenerator<Type>::UniformRandomNumbers(Type *value, std::size_t size, Type a, Type b, std::string shared) {
  std::mt19937 *gen = SelectGeneratorAndUpdateCounter(size, shared);
  for (std::size_t i = 0; i < size; ++i) value[i] = a + (b - a) * u_distribution(*gen);
}

template <typename Type>
void RandomNumberGenerator<Type>::GaussianRandomNumbers(Type *value, std::size_t size, std::string shared) {
  std::mt19937 *gen = SelectGeneratorAndUpdateCounter(2 * size, shared);
  for (std::size_t i = 0; i < size; ++i) value[i] = n_distribution(*gen);
}

template <typename Type>
void RandomNumberGenerator<Type>::RandomIntegersInRange(int *value, std::size_t size, int a, int b, std::string shared) {
  std::mt19937 *gen = SelectGeneratorAndUpdateCounter(size, shared);
  for (std::size_t i = 0; i < size; ++i) {
    Type r = u_distribution(*gen);
    value[i] = (int)std::floor((Type)a + r * Type(b - a));
  }
}

template class RandomNumberGenerator<float>;
template class RandomNumberGenerator<double>;

template <typename Type, typename TypeFloat>
void ShuffleFisherYates(std::vector<Type> &array, RandomNumberGenerator<TypeFloat> *rng, std::string shared) {
  for (int hi = (int)array.size() - 1; hi > 0; --hi) {
    int pick;
    rng->RandomIntegersInRange(&pick, 1UL, 0, hi + 1, shared);
    if (pick != hi) std::swap(array[hi], array[pick]);
  }
}
template void ShuffleFisherYates<int, double>(std::vector<int> &, RandomNumberGenerator<double> *, std::string);
template void ShuffleFisherYates<float, double>(std::vector<float> &, RandomNumberGenerator<double> *, std::string);
template void ShuffleFisherYates<double, double>(std::vector<double> &, RandomNumberGenerator<double> *, std::string);
template void ShuffleFisherYates<unsigned, double>(std::vector<unsigned> &, RandomNumberGenerator<double> *, std::string);

}  // namespace iqs

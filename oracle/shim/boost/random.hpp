// TEST INFRASTRUCTURE ONLY (oracle/): stand-in for <boost/random.hpp>.
// The reference declares a boost::random::mt19937 member it never draws from
// (src/sampler.h:35) and a uniform_int_distribution in dead code (src/mcmc_utils.h:324).
#pragma once
#include <random>
namespace boost
{
namespace random
{
using mt19937 = std::mt19937;
template <class T = int>
using uniform_int_distribution = std::uniform_int_distribution<T>;
}  // namespace random
}  // namespace boost

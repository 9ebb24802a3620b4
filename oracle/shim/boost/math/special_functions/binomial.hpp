// TEST INFRASTRUCTURE ONLY (oracle/): stand-in for
// <boost/math/special_functions/binomial.hpp> (BH is not installed; version unpinned by
// the reference's DESCRIPTION "LinkingTo: BH").
//
// boost::math::binomial_coefficient<T>(n, k) returns C(n, k) rounded to an integer-valued T.
// Restated as the exact integer (multiplicative formula in long double, every partial
// product an integer < 2^64 for n <= 64) converted to T.  Call sites:
// src/sampler.cpp:270,305,327,330 (proposal density of sample_latent_genotype).
#pragma once
#include <cmath>
namespace boost
{
namespace math
{
template <class T>
inline T binomial_coefficient(unsigned n, unsigned k)
{
    if (k > n) return T(0);
    if (k > n - k) k = n - k;
    long double c = 1.0L;
    for (unsigned i = 1; i <= k; ++i)
    {
        c = c * static_cast<long double>(n - k + i) / static_cast<long double>(i);
        c = std::floor(c + 0.5L);
    }
    return static_cast<T>(c);
}
}  // namespace math
}  // namespace boost

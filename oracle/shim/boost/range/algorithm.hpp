// TEST INFRASTRUCTURE ONLY (oracle/): stand-in for <boost/range/algorithm.hpp>
// (used only by dead code, src/mcmc_utils.h:315-330).
#pragma once
#include <algorithm>
#include <utility>
namespace boost
{
namespace range
{
template <class Range, class Gen>
inline Range &random_shuffle(Range &r, Gen &gen)
{
    auto n = r.end() - r.begin();
    for (decltype(n) i = n - 1; i > 0; --i)
    {
        std::swap(r.begin()[i], r.begin()[gen(i + 1)]);
    }
    return r;
}
}  // namespace range
}  // namespace boost

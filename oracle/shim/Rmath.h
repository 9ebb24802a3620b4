// TEST INFRASTRUCTURE ONLY (oracle/): stand-in for R's <Rmath.h> (namespace R::).
//
// R itself is not installed here, so the five Rmath entry points the reference's hot path
// names (src/sampler.cpp:28-36,52-55,74-77; src/mcmc.cpp:212) are restated from their
// published definitions (R nmath: dbeta.c, dgamma.c, dpois.c), including R's documented
// edge cases at x == 0 and x == 1 (the relatedness-off path evaluates dbeta(0, 1, 1)).
// Accuracy vs real Rmath: ~1e-15 relative in double away from the edges (lgamma-based
// closed forms instead of R's saddle-point evaluation).  "Parity unpinned" at this
// boundary: the reference's own tests hold no vectors for these calls.
#pragma once
#ifndef MOIRE_ORACLE_SHIM_RMATH_H
#define MOIRE_ORACLE_SHIM_RMATH_H

#include <cmath>
#include <cstdint>
#include <limits>
#include <random>

namespace moire_shim
{
inline std::mt19937_64 &runif_engine()
{
    static std::mt19937_64 eng(0x6d6f697265ULL);
    return eng;
}
inline void seed_runif(std::uint64_t seed) { runif_engine().seed(seed); }
}  // namespace moire_shim

namespace R
{
inline double dbeta(double x, double a, double b, int give_log)
{
    const double ninf = -std::numeric_limits<double>::infinity();
    const double pinf = std::numeric_limits<double>::infinity();
    if (std::isnan(x) || std::isnan(a) || std::isnan(b)) return x + a + b;
    if (a < 0 || b < 0) return std::numeric_limits<double>::quiet_NaN();
    if (x < 0 || x > 1) return give_log ? ninf : 0.0;
    if (x == 0)
    {
        if (a > 1) return give_log ? ninf : 0.0;
        if (a < 1) return pinf;
        return give_log ? std::log(b) : b;
    }
    if (x == 1)
    {
        if (b > 1) return give_log ? ninf : 0.0;
        if (b < 1) return pinf;
        return give_log ? std::log(a) : a;
    }
    const double lbeta = std::lgamma(a) + std::lgamma(b) - std::lgamma(a + b);
    const double lval = (a - 1) * std::log(x) + (b - 1) * std::log1p(-x) - lbeta;
    return give_log ? lval : std::exp(lval);
}

inline double dgamma(double x, double shape, double scale, int give_log)
{
    const double ninf = -std::numeric_limits<double>::infinity();
    const double pinf = std::numeric_limits<double>::infinity();
    if (std::isnan(x) || std::isnan(shape) || std::isnan(scale)) return x + shape + scale;
    if (shape < 0 || scale <= 0) return std::numeric_limits<double>::quiet_NaN();
    if (x < 0) return give_log ? ninf : 0.0;
    if (shape == 0) return (x == 0) ? pinf : (give_log ? ninf : 0.0);
    if (x == 0)
    {
        if (shape < 1) return pinf;
        if (shape > 1) return give_log ? ninf : 0.0;
        return give_log ? -std::log(scale) : 1 / scale;
    }
    const double lval = (shape - 1) * std::log(x) - x / scale - std::lgamma(shape) -
                        shape * std::log(scale);
    return give_log ? lval : std::exp(lval);
}

inline double dpois(double x, double lambda, int give_log)
{
    const double lval = x * std::log(lambda) - lambda - std::lgamma(x + 1);
    return give_log ? lval : std::exp(lval);
}

inline double runif(double a, double b)
{
    std::uniform_real_distribution<double> d(a, b);
    return d(moire_shim::runif_engine());
}

inline double rgamma(double shape, double scale)
{
    std::gamma_distribution<double> d(shape, scale);
    return d(moire_shim::runif_engine());
}
}  // namespace R

#endif
